"""Pure-write / pure-read / copy HBM rates of this B200 (torch fill_, sum, copy_ on 2 GB): the ceilings the write-heavy
projections (kv_proj: 26 % read / 74 % write) should be read against."""
import torch
n = 1 << 29          # 2 GiB of fp32
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.fill_(1.0)); print(f"write only (fill_): {n*4/ms/1e6:.0f} GB/s")
ms = t(lambda: torch.cuda.memset if False else a.zero_()); print(f"write only (zero_ / memset): {n*4/ms/1e6:.0f} GB/s")
ms = t(lambda: a.sum()); print(f"read only (sum): {n*4/ms/1e6:.0f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy (read + write): {2*n*4/ms/1e6:.0f} GB/s")
h = a.view(torch.float16)[:n]   # 1 GiB fp16 source -> fp32? no: fp32 -> fp16 cast = 4 B read, 2 B write
c = torch.empty(n, dtype=torch.float16, device="cuda")
ms = t(lambda: c.copy_(a)); print(f"cast fp32->fp16 (67 % read): {n*6/ms/1e6:.0f} GB/s")
d = torch.empty(n, device="cuda")
ms = t(lambda: d.copy_(c)); print(f"cast fp16->fp32 (33 % read / 67 % write): {n*6/ms/1e6:.0f} GB/s")
