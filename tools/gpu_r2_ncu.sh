#!/bin/bash
mkdir -p gpurun_out
NVCC_EXTRA="-DX3_POLY_MOD=0" python -m openvis_b200.build --force > /dev/null
ncu --set full --clock-control none --import-source on -k regex:xattn_tc3 -s 2 -c 1 -o gpurun_out/r2_tc3_v3 python tools/prof_xattn_t.py 4 100 529920 > gpurun_out/r2_tc3_v3.log 2>&1
tail -2 gpurun_out/r2_tc3_v3.log
