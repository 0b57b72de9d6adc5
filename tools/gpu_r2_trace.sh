#!/bin/bash
mkdir -p gpurun_out
NVCC_EXTRA=-DOVIS_XATTN_TRACE_BUILD python -m openvis_b200.build --force > /dev/null
python tools/trace_xattn_t.py > gpurun_out/r2_tc3_trace.txt 2>&1
cat gpurun_out/r2_tc3_trace.txt
