"""A/B of the two cross-attention kernels at one launch size: xattn_tc2 (thread = query, [word][Q] mask bits) and xattn_tc3
(transposed scores, key-major mask bits).  python tools/prof_xattn_t.py [G Q keys]   (SPARSE=1: block-sparse masks)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
G, Q, keys = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (1, 100, 529920)
g = torch.Generator().manual_seed(0)
q = (torch.randn(G * Q, 256, generator=g) * 0.6).half().cuda()
k = torch.randn(G * keys, 256, generator=g).half().cuda()
v = torch.randn(G * keys, 256, generator=g).half().cuda()
W = (keys + 31) // 32
qw = 4 * ((Q + 127) // 128)
gd = torch.Generator(device="cuda").manual_seed(1)
if os.environ.get("SPARSE"):
    blocked = torch.ones(G, Q, keys, dtype=torch.bool, device="cuda")
    for qb in range((Q + 31) // 32):
        k0 = (qb * keys) // 8 % keys
        blocked[:, qb * 32:(qb + 1) * 32, k0:k0 + keys // 8] = torch.rand(G, min(32, Q - qb * 32), keys // 8, device="cuda", generator=gd) < 0.5
else:
    blocked = torch.rand(G, Q, keys, device="cuda", generator=gd) < 0.5
flags = (~blocked).any(-1).to(torch.uint8).contiguous()
r = torch.arange(keys, device="cuda")
bits = torch.zeros(G, W, Q, dtype=torch.int64, device="cuda")
bits.scatter_add_(1, (r // 32)[None, :, None].expand(G, keys, Q), (blocked.permute(0, 2, 1).long() << (r % 32)[None, :, None]))
bits = bits.to(torch.int32).contiguous()
full = torch.ones(G, qw * 32, keys, dtype=torch.bool, device="cuda")
full[:, :Q] = blocked
bt = torch.zeros(G, keys, qw, dtype=torch.int64, device="cuda")
for b in range(32):
    bt += full[:, b::32].permute(0, 2, 1).long() << b
bits_t = torch.where(bt >= 2 ** 31, bt - 2 ** 32, bt).to(torch.int32).contiguous()
pad = torch.full((G, W * 32, qw), -1, dtype=torch.int32, device="cuda")
pad[:, :keys] = bits_t
blockand = pad.view(G, W, 32, qw)[:, :, 0].clone()
for i in range(1, 32):
    blockand &= pad.view(G, W, 32, qw)[:, :, i]
del full, bt, pad, blocked
out = torch.empty(G * Q, 256, dtype=torch.float16, device="cuda")
out_t = torch.empty(G * Q, 256, dtype=torch.float16, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


splits, q_pad, o_n, ml_n = L.xattn_plan(G, Q, keys)
o_part = torch.empty(o_n, device="cuda"); ml_part = torch.empty(ml_n, device="cuda")
ms2 = timeit(lambda: L.xattn(q, k, v, bits, flags, G, Q, Q, keys, splits, o_part, ml_part, out))
use_t, splits_t, q_pad_t, o_nt, ml_nt = L.xattn_plan_t(G, Q, keys)
o_pt = torch.empty(o_nt, device="cuda"); ml_pt = torch.empty(ml_nt, device="cuda")
stats = torch.zeros(2, dtype=torch.int32, device="cuda")
ms3 = timeit(lambda: L.xattn_t(q, k, v, bits_t, blockand, flags, G, Q, Q, keys, splits_t, o_pt, ml_pt, out_t, stats=stats))
st = stats.tolist()
err = (out.float() - out_t.float()).abs().max().item()
print(f"G={G} Q={Q} keys={keys}: tc2 {ms2*1e3:.1f} us ({splits} partials), tc3 {ms3*1e3:.1f} us ({splits_t} partials, plan use_t={use_t}, "
      f"retried CTAs {st[1]}/{st[0]}), tc2/tc3 = {ms2/ms3:.2f}, max |tc2 - tc3| = {err:.2e}, "
      f"tc3: {4*Q*keys*256*G/ms3/1e9:.1f} TFLOP/s")
