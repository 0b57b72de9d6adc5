#!/bin/bash
mkdir -p gpurun_out
for w in brivis_frame_36x360x640_q100_k1196 san_online_36x720x1280_q200_k1196; do
  for s in 1 2; do
    echo "== $w streams=$s"
    timeout 600 python bench.py --workload $w --streams $s --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'launches',d['gpu_launches'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})"
  done
done > gpurun_out/r2_lanes.txt 2>&1
cat gpurun_out/r2_lanes.txt
