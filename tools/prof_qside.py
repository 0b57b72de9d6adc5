"""Query-side chain timing (linear+LN variants, self-attention) at the Video decoder's row counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
g = torch.Generator().manual_seed(0)
for rows in (100, 400):
    for K, two in ((256, False), (2048, True)):
        x = torch.randn(rows, K, generator=g).half().cuda(); w = (torch.randn(256, K, generator=g) * K ** -0.5).half().cuda()
        b = torch.randn(256, generator=g).cuda(); resid = torch.randn(rows, 256, generator=g).cuda()
        ln = (torch.ones(256).cuda(), torch.zeros(256).cuda()); pe = torch.randn(100, 256, generator=g).cuda()
        y32 = torch.empty(rows, 256).cuda(); y16 = torch.empty(rows, 256).half().cuda(); ye = torch.empty_like(y16)
        d32 = torch.empty_like(y32); d16 = torch.empty_like(y16)
        ws = torch.empty((K // 256) * ((rows + 127) // 128) * 128 * 256).cuda()
        f = lambda s: L.linear_ln_f16(x, w, b, resid, ln, ln if two else None, pe, y32, y16, ye, d32 if two else None, d16 if two else None, split_ws=s)
        print(f"rows={rows} K={K} two={two}: fused {timeit(lambda: f(None)):.1f} us, split {timeit(lambda: f(ws)):.1f} us")
    qk = torch.randn(rows, 512, generator=g).half().cuda(); v = torch.randn(rows, 256, generator=g).half().cuda(); o = torch.empty_like(v)
    print(f"rows={rows} self_attn: {timeit(lambda: L.self_attn(qk, v, o, rows // 100, 100)):.1f} us")
    x = torch.randn(rows, 256, generator=g).half().cuda(); w = (torch.randn(256, 256, generator=g) / 16).half().cuda(); o = torch.empty(rows, 256).half().cuda()
    print(f"rows={rows} linear 256->256: {timeit(lambda: L.linear_f16(x, w, None, out=o)):.1f} us")
    w = (torch.randn(2048, 256, generator=g) / 16).half().cuda(); o = torch.empty(rows, 2048).half().cuda()
    print(f"rows={rows} linear 256->2048: {timeit(lambda: L.linear_f16(x, w, None, relu=True, out=o)):.1f} us")
