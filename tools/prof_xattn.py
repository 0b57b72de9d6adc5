"""One cross-attention launch at the cfg-2 level-2 size (for ncu / timing).  python tools/prof_xattn.py [G Q keys]"""
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
bits = torch.randint(-2**31, 2**31 - 1, (G, W, Q), generator=g, dtype=torch.int64).to(torch.int32).cuda()
if os.environ.get("SPARSE"):
    # object-like masks: each block of 32 queries sees one contiguous 1/8 of the keys (random bits inside), nothing else
    bits = torch.full((G, W, Q), -1, dtype=torch.int32)
    for qb in range((Q + 31) // 32):
        w0 = (qb * W) // 8 % W
        bits[:, w0:w0 + W // 8, qb * 32:(qb + 1) * 32] = torch.randint(-2**31, 2**31 - 1, (G, W // 8, min(32, Q - qb * 32)), generator=g, dtype=torch.int64).to(torch.int32)
    bits = bits.cuda()
flags = torch.ones(G, Q, dtype=torch.uint8).cuda()
splits, q_pad, o_n, ml_n = L.xattn_plan(G, Q, keys)
o_part = torch.empty(o_n, device="cuda"); ml_part = torch.empty(ml_n, device="cuda")
out = torch.empty(G * Q, 256, dtype=torch.float16, device="cuda")
for _ in range(3):
    L.xattn(q, k, v, bits, flags, G, Q, Q, keys, splits, o_part, ml_part, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.xattn(q, k, v, bits, flags, G, Q, Q, keys, splits, o_part, ml_part, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"G={G} Q={Q} keys={keys} splits={splits}: {ms*1e3:.1f} us per xattn (split+combine), {4*Q*keys*256*G/ms/1e9:.1f} TFLOP/s")
