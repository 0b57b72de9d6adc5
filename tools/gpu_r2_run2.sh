#!/bin/bash
mkdir -p gpurun_out
(cd tools/ubench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../openvis_b200/csrc -o umma_probe umma_probe.cu && timeout 60 ./umma_probe > ../../gpurun_out/r2_umma_probe.txt 2>&1; echo "rc=$?" >> ../../gpurun_out/r2_umma_probe.txt)
cat gpurun_out/r2_umma_probe.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run2_pytest.txt
tail -8 gpurun_out/r2_run2_pytest.txt
