#!/bin/bash
mkdir -p gpurun_out
for fl in "-DX3_EXPERIMENT_SKELETON -DX3_EXP_PV_STEPS=4" "-DX3_EXPERIMENT_SKELETON -DX3_EXP_PV_N32" "-DX3_EXPERIMENT_SKELETON -DX3_EXP_PV_STEPS=1"; do
  NVCC_EXTRA="$fl" python -m openvis_b200.build --force > /dev/null
  echo "== flags: $fl"
  python tools/prof_xattn_t.py 4 100 529920 | sed 's/max |tc2.*//'
done > gpurun_out/r2_tc3_exp.txt 2>&1
cat gpurun_out/r2_tc3_exp.txt
