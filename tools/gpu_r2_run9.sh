#!/bin/bash
# new rows: pixel decoder (f-2), zero-shot decoder; then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixel_decoder_gpu.py tests/test_ov_tails_gpu.py tests/test_msda_gpu.py -m gpu -q -s > gpurun_out/r2_run9_new.txt 2>&1
echo "new tests rc=$?"; grep -E "max err|passed|failed|Error|error" gpurun_out/r2_run9_new.txt | tail -40
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run9_all.txt 2>&1
echo "all rc=$?"; tail -5 gpurun_out/r2_run9_all.txt
