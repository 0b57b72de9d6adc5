#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_scale2.json 2> gpurun_out/r2_scale2.err; echo "torchrun rc=$?"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2_scale2.err | tail -30
head -c 600 gpurun_out/r2_scale2.json
