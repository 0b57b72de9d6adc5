"""clock64 timeline of block 0 of xattn_tc3_kernel (roles 0-2: softmax warpgroups, 3: S issuer, 4: PV issuer, 5 / 6: K / V
   producers).  Needs a library built with the stamps compiled in:
       NVCC_EXTRA=-DOVIS_XATTN_TRACE_BUILD python -m openvis_b200.build --force"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ts = torch.zeros(7 * 128 * 8, dtype=torch.int64, device="cuda")
os.environ["OVIS_XATTN_TRACE"] = str(ts.data_ptr())
from openvis_b200 import _lib as L
G, Q, keys = 1, 100, 529920
g = torch.Generator().manual_seed(0)
q = (torch.randn(G * Q, 256, generator=g) * 0.6).half().cuda()
k = torch.randn(G * keys, 256, generator=g).half().cuda(); v = torch.randn(G * keys, 256, generator=g).half().cuda()
qw = 4
bits_t = torch.randint(-2**31, 2**31 - 1, (G, keys, qw), generator=g, dtype=torch.int64).to(torch.int32).cuda()
flags = torch.ones(G, Q, dtype=torch.uint8).cuda()
use_t, splits, q_pad, o_n, ml_n = L.xattn_plan_t(G, Q, keys)
o_part = torch.empty(o_n, device="cuda"); ml_part = torch.empty(ml_n, device="cuda")
out = torch.empty(G * Q, 256, dtype=torch.float16, device="cuda")
for _ in range(3):
    L.xattn_t(q, k, v, bits_t, None, flags, G, Q, Q, keys, splits, o_part, ml_part, out)
torch.cuda.synchronize()
t = ts.cpu().view(7, 128, 8)
n0 = 10
base = t[0, n0, 0].item()
print("cycles relative to warpgroup 0's step", n0, " wg rows: [begin, S available, 2 chunks done, P buffer free, S drained, P handed over]")
for n in range(n0, n0 + 8):
    for wg in range(2):
        u = wg + 2 * n
        ev = [t[wg, n, e].item() - base for e in range(6)]
        print(f"wg{wg} n={n} u={u} (tile {u >> 1} head {u & 1}): {ev}  wait_S={ev[1]-ev[0]} chunks01={ev[2]-ev[1]} wait_P={ev[3]-ev[2]} "
              f"rest={ev[5]-ev[3]}  | S committed {t[3, u, 0].item() - base}  PV committed {t[4, u, 0].item() - base}")
    for tile in range(n, n + 1):
        print(f"      tile {tile}: K wait/issue {t[5, tile, 0].item() - base} {t[5, tile, 1].item() - base}   V wait/issue {t[6, tile, 0].item() - base} {t[6, tile, 1].item() - base}")
for wg in range(2):
    print(f"wg{wg}: {(t[wg, 35, 0].item() - t[wg, 10, 0].item()) / 25:.0f} cycles per unit (steps 10..35)")
