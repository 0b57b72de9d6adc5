"""clock64 timeline of block 0 of xattn_tc2_kernel (roles: 4 softmax warpgroups, S issuer, PV issuer).
   Needs a library built with the stamps compiled in:
       NVCC_EXTRA=-DOVIS_XATTN_TRACE_BUILD python -m openvis_b200.build --force
   python tools/trace_xattn.py            dense random masks;   SPARSE=1 -> 7/8 of the (warp, half tile) pairs skip"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ts = torch.zeros(6 * 64 * 8, dtype=torch.int64, device="cuda")
os.environ["OVIS_XATTN_TRACE"] = str(ts.data_ptr())
from openvis_b200 import _lib as L
G, Q, keys = 1, 100, 529920
g = torch.Generator().manual_seed(0)
q = (torch.randn(G * Q, 256, generator=g) * 0.6).half().cuda()
k = torch.randn(G * keys, 256, generator=g).half().cuda(); v = torch.randn(G * keys, 256, generator=g).half().cuda()
W = (keys + 31) // 32
bits = torch.randint(-2**31, 2**31 - 1, (G, W, Q), generator=g, dtype=torch.int64).to(torch.int32)
if os.environ.get("SPARSE"):
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
t = ts.cpu().view(6, 64, 8)
base = t[0, 20, 0].item()
print("cycles relative to warpgroup 0's step 20.  wg rows: [step begin, S available, softmax done, P buffer free, P handed over]")
for step in range(20, 28):
    for role in range(4):
        ev = [t[role, step, e].item() - base for e in range(5)]
        d = [ev[i + 1] - ev[i] for i in range(4)]
        print(f"wg{role} n={step}: {ev}  deltas wait_S={d[0]} softmax={d[1]} wait_P={d[2]} store+fence={d[3]}")
    print(f"   S issued  n={step}: " + str([t[4, step, w].item() - base for w in range(4)]))
    print(f"   PV issued n={step}: " + str([t[5, step, w].item() - base for w in range(4)]))
for role in range(4):
    per = (t[role, 50, 0].item() - t[role, 20, 0].item()) / 30
    print(f"wg{role}: {per:.0f} cycles per step (steps 20..50)")
