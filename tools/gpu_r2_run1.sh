#!/bin/bash
# round-2 GPU run 1: micro-benchmark of the candidate softmax streams, GPU test tier, default bench (sweep + other configs)
mkdir -p gpurun_out
(cd tools/ubench && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o softmax_loop softmax_loop.cu && ./softmax_loop > ../../gpurun_out/r2_softmax_loop.txt 2>&1)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run1_pytest.txt
timeout 900 python bench.py > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err; echo "bench rc=$?" >> gpurun_out/r2_run1_bench.err
tail -5 gpurun_out/r2_run1_pytest.txt; tail -3 gpurun_out/r2_run1_bench.err; head -c 1500 gpurun_out/r2_run1_bench.json
