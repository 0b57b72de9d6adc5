#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_cfg2.csv \
  python bench.py --workload openvis_video_36x720x1280_q100_k40 --clips 1 --streams 1 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_r2_cfg2.log 2>&1
OVIS_PROF_ONCE=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_brivis.csv \
  python tools/prof_brivis.py > gpurun_out/launches_r2_brivis.log 2>&1
wc -l gpurun_out/launches_r2_cfg2.csv gpurun_out/launches_r2_brivis.csv
