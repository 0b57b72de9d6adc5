#!/bin/bash
mkdir -p gpurun_out
for fl in "" "-DX3_EXPERIMENT_NO_L"; do
  NVCC_EXTRA="$fl" python -m openvis_b200.build --force > /dev/null
  echo "== flags: $fl"
  python tools/prof_xattn_t.py 1 100 529920
  python tools/prof_xattn_t.py 4 100 529920
done > gpurun_out/r2_tc3_exp.txt 2>&1
cat gpurun_out/r2_tc3_exp.txt
