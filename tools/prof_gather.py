"""The sweep's result gather alone (scores + top-10 + packed masks of 64 clips) under torchrun: ms per gather and GB/s received."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from openvis_b200.sharding import gather_clip_dict
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 64 // world
res = {"scores": torch.randn(n, 100, 1196, device="cuda"), "top": torch.randn(n, 3, 10, device="cuda"),
       "masks": torch.randint(0, 2 ** 31 - 1, (n, 10, 36, 720, 40), dtype=torch.int32, device="cuda")}
for _ in range(3):
    out = gather_clip_dict(res, 64)
torch.cuda.synchronize(); dist.barrier()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for a, b in ev:
    a.record(); out = gather_clip_dict(res, 64); b.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
byts = sum(v.numel() * v.element_size() for v in res.values()) * (world - 1)
if rank == 0:
    print(f"world {world} [{os.environ.get('TAG', 'default')}]: gather median {ts[5]:.2f} ms (min {ts[0]:.2f}), {byts / ts[5] / 1e6:.0f} GB/s received per rank")
dist.destroy_process_group()
