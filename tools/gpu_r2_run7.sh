#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decoder_gpu.py -x -q -k "chain or cfg1_shape" > gpurun_out/r2_run7_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_run7_pytest.txt
tail -25 gpurun_out/r2_run7_pytest.txt
