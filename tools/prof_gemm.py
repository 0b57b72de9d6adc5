"""GEMM-family kernels at cfg-2 sizes (for timing / ncu).  python tools/prof_gemm.py [which]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
which = sys.argv[1] if len(sys.argv) > 1 else "all"
T, Q = 36, 100
N2, M = 14720, 58880
g = torch.Generator().manual_seed(0)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
if which in ("all", "kv"):
    rows = T * N2
    xk = torch.randn(rows, 256, generator=g).half().cuda(); xv = torch.randn(rows, 256, generator=g).half().cuda()
    w = (torch.randn(1536, 256, generator=g) / 16).half().cuda()
    outs = [torch.empty(rows, 256, dtype=torch.float16, device="cuda") for _ in range(6)]
    biases = [torch.randn(256, generator=g).cuda() for _ in range(6)]
    ms = timeit(lambda: L.kv_proj_f16(xk, xv, w, outs, biases))
    byts = rows * 256 * 2 * 2 + rows * 1536 * 2
    print(f"kv_proj L2 rows={rows}: {ms*1e3:.1f} us  {2*rows*1536*256/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: L.kv_proj_f16(xk, xv, w, outs, None))
    print(f"kv_proj L2 (no bias): {ms*1e3:.1f} us  {2*rows*1536*256/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.1f} GB/s")
    del xk, xv, outs
if which in ("all", "ml"):
    rows = T * M
    ft = torch.randn(rows, 256, generator=g).half().cuda()
    me = (torch.randn(Q, 256, generator=g) * 0.1).half().cuda()
    out = torch.empty(Q, rows, device="cuda")
    pf = torch.zeros(T, Q, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: L.mask_logits(ft, 1, rows, me, Q, Q, out, rows, rows, posflags=pf, rows_per_frame=M))
    byts = rows * 256 * 2 + Q * rows * 4
    print(f"mask_logits rows={rows}: {ms*1e3:.1f} us  {byts/ms/1e6:.1f} GB/s")
    ms = timeit(lambda: L.mask_logits(ft, 1, rows, me, Q, Q, out, rows, rows))
    print(f"mask_logits (no posflags): {ms*1e3:.1f} us  {byts/ms/1e6:.1f} GB/s")
    del out
    bits = torch.zeros(1, (rows + 31) // 32, Q, dtype=torch.int32, device="cuda"); flags = torch.zeros(1, Q, dtype=torch.uint8, device="cuda")
    rows2 = T * N2
    ms = timeit(lambda: L.mask_bits(ft[:rows2], 1, rows2, me, Q, bits, flags, Q))
    print(f"mask_bits rows={rows2}: {ms*1e3:.1f} us  {(rows2*512 + Q*rows2/8)/ms/1e6:.1f} GB/s")
if which in ("all", "prep"):
    F = torch.randn(T, 256, 184, 320, generator=g).cuda()
    ms = timeit(lambda: L.maskfeat_prep(F))
    byts = F.numel() * 6 + (T * 19320) * 512
    print(f"maskfeat_prep: {ms*1e3:.1f} us  {byts/ms/1e6:.1f} GB/s")
    x2 = torch.randn(T, 256, 92, 160, generator=g).cuda()
    pos = torch.randn(N2, 256, generator=g).cuda(); pz = torch.randn(T, 256, generator=g).cuda()
    o1 = torch.empty(T, N2, 256, dtype=torch.float16, device="cuda"); o2 = torch.empty_like(o1)
    ms = timeit(lambda: L.nchw_to_tokens_f16(x2, out=o1, out_pos=o2, pos=pos, pos_t=pz))
    print(f"nchw_to_tokens L2: {ms*1e3:.1f} us  {x2.numel()*8/ms/1e6:.1f} GB/s")
    pcn = pos.t().contiguous()
    ms = timeit(lambda: L.nchw_to_tokens_hw_f16(x2, out=o1, out_pos=o2, pos_cn=pcn, pos_t=pz))
    print(f"nchw_to_tokens_hw (TMA) L2: {ms*1e3:.1f} us  {x2.numel()*8/ms/1e6:.1f} GB/s")
