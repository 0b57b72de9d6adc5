#!/bin/bash
mkdir -p gpurun_out
for fl in "-DX3_POLY_MOD=0" "-DX3_POLY_MOD=4" "-DX3_POLY_MOD=3" "-DX3_POLY_MOD=2"; do
  NVCC_EXTRA="$fl" python -m openvis_b200.build --force > /dev/null
  echo "== flags: $fl"
  python tools/prof_xattn_t.py 1 100 529920
  python tools/prof_xattn_t.py 4 100 529920
done > gpurun_out/r2_tc3_exp.txt 2>&1
cat gpurun_out/r2_tc3_exp.txt
NVCC_EXTRA="-DX3_POLY_MOD=4" python -m openvis_b200.build --force > /dev/null
timeout 600 python -m pytest tests/test_xattn_t_gpu.py -q -x 2>&1 | tail -3
