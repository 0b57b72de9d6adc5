"""Compute-bound GEMMs of the CLIP ViT-B/16 tower (500 crops x 197 tokens = 98 500 rows): qkv (768 -> 2304), out (768 -> 768),
fc1 (768 -> 3072, QuickGELU), fc2 (3072 -> 768, + residual).  python tools/prof_gemm_tower.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
g = torch.Generator().manual_seed(0)
rows = 98500
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, K, N, act in (("qkv", 768, 2304, 0), ("out", 768, 768, 0), ("fc1", 768, 3072, 2), ("fc2", 3072, 768, 0)):
    x = torch.randn(rows, K, generator=g).half().cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().cuda()
    b = torch.randn(N, generator=g).cuda()
    out = torch.empty(rows, N, dtype=torch.float16, device="cuda")
    ms = timeit(lambda: L.linear_act_f16(x, w, b, act=act, out=out))
    print(f"{name}: rows={rows} K={K} N={N}: {ms*1e3:.1f} us  {2*rows*K*N/ms/1e9:.0f} TFLOP/s  ({rows*(K+N)*2/ms/1e6:.0f} GB/s of operands)")
