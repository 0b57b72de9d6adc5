"""Host-to-device copy ceiling of the box with N ranks copying at once (no compute): what bounds bench.py's e2e leg at N > 1.
torchrun --nproc-per-node N tools/prof_h2d_ceiling.py   -> per-rank and aggregate GB/s, alone (rank 0 only) and all together."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import bench
    aff = bench.bind_to_gpu_numa_node(local)
except Exception as e:          # the measurement still stands without the binding
    aff = {"error": str(e)[:80]}
n = 1 << 30                     # 1 GiB pinned
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.fill_(1)
devb = torch.empty(n, dtype=torch.uint8, device="cuda")


def rate(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        devb.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


rate(2)
alone = None
for r in range(world):          # every rank alone, in turn
    if world > 1:
        dist.barrier()
    if r == rank:
        alone = rate(8)
if world > 1:
    dist.barrier()
together = rate(16)
res = torch.tensor([alone, together], device="cuda")
if world > 1:
    allr = [torch.zeros_like(res) for _ in range(world)]
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    a = [round(float(t[0]), 1) for t in allr]
    b = [round(float(t[1]), 1) for t in allr]
    print(f"ranks {world}: H2D alone per rank {a} GB/s; all together per rank {b} GB/s, aggregate {sum(b):.0f} GB/s; rank-0 affinity {aff}")
if world > 1:
    dist.destroy_process_group()
