"""mask_logits in the Frame decoders' layout (one group per frame, output [1, Q, BT, H, W]) against the Video layout."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
g = torch.Generator().manual_seed(0)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (BT, M, Q) in ((36, 58880, 200), (36, 58880, 100), (144, 15360, 100), (36, 15360, 100)):
    ft = torch.randn(BT * M, 256, generator=g).half().cuda()
    me = (torch.randn(BT * Q, 256, generator=g) * 0.1).half().cuda()
    byts = BT * M * (512 + 4 * Q)
    out = torch.empty(1, Q, BT, M, device="cuda")
    pf = torch.zeros(BT, Q, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: L.mask_logits(ft, BT, M, me, Q, Q, out, M, BT * M, posflags=pf, rows_per_frame=M))
    print(f"frame layout BT={BT} M={M} Q={Q}: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: L.mask_logits(ft, BT, M, me, Q, Q, out, M, BT * M))
    print(f"   no posflags: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s")
    out2 = torch.empty(1, Q, BT, M, device="cuda")
    ms = timeit(lambda: L.mask_logits(ft, 1, BT * M, me, 0, Q, out2, Q * BT * M, BT * M))
    print(f"   video layout (one group): {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s")
    del ft, out, out2
