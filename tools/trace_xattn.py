import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ts = torch.zeros(6 * 64 * 8, dtype=torch.int64, device="cuda")
os.environ["OVIS_XATTN_TS"] = str(ts.data_ptr())
from openvis_b200 import _lib as L
G, Q, keys = 1, 100, 529920
g = torch.Generator().manual_seed(0)
q = (torch.randn(G * Q, 256, generator=g) * 0.6).half().cuda()
k = torch.randn(G * keys, 256, generator=g).half().cuda(); v = torch.randn(G * keys, 256, generator=g).half().cuda()
W = (keys + 31) // 32
bits = torch.randint(-2**31, 2**31 - 1, (G, W, Q), generator=g, dtype=torch.int64).to(torch.int32).cuda()
flags = torch.ones(G, Q, dtype=torch.uint8).cuda()
splits, q_pad, o_n, ml_n = L.xattn_plan(G, Q, keys)
o_part = torch.empty(o_n, device="cuda"); ml_part = torch.empty(ml_n, device="cuda")
out = torch.empty(G * Q, 256, dtype=torch.float16, device="cuda")
for _ in range(3):
    L.xattn(q, k, v, bits, flags, G, Q, Q, keys, splits, o_part, ml_part, out)
torch.cuda.synchronize()
t = ts.cpu().view(6, 64, 8)
base = t[0, 8, 0].item()
names = {0: "S-is", 1: "WG00", 2: "WG10", 3: "WG01", 4: "WG11", 5: "PV-i"}
for step in range(12, 20):
    for role in (0, 1, 2, 3, 4, 5):
        ev = [(t[role, step, e].item() - base) if t[role, step, e].item() else None for e in range(8)]
        print(names[role], step, ev)
