import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
T, N2 = 36, 14720
g = torch.Generator().manual_seed(0)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x2 = torch.randn(T, 256, 92, 160, generator=g).cuda()
pos = torch.randn(N2, 256, generator=g).cuda(); pz = torch.randn(T, 256, generator=g).cuda()
pcn = pos.t().contiguous()
o1 = torch.empty(T, N2, 256, dtype=torch.float16, device="cuda"); o2 = torch.empty_like(o1)
for name, fn, byts in [
    ("old xt+xp", lambda: L.nchw_to_tokens_f16(x2, out=o1, out_pos=o2, pos=pos, pos_t=pz), 8),
    ("old xt only", lambda: L.nchw_to_tokens_f16(x2, out=o1), 6),
    ("tma xt+xp+pz", lambda: L.nchw_to_tokens_hw_f16(x2, out=o1, out_pos=o2, pos_cn=pcn, pos_t=pz), 8),
    ("tma xt+xp", lambda: L.nchw_to_tokens_hw_f16(x2, out=o1, out_pos=o2, pos_cn=pcn), 8),
    ("tma xt only", lambda: L.nchw_to_tokens_hw_f16(x2, out=o1), 6)]:
    ms = timeit(fn)
    print(f"{name}: {ms*1e3:.1f} us  {x2.numel()*byts/ms/1e6:.1f} GB/s")
