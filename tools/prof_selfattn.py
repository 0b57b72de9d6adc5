"""Self-attention kernel timing: decoder shapes (groups = frames or clips, rows = queries) and the resampler's
(groups = instances, rows = frames)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
for G, Q in ((4, 100), (36, 100), (36, 200), (100, 36), (144, 100), (400, 36)):
    qk = torch.randn(G * Q, 512, device="cuda").half()
    v = torch.randn(G * Q, 256, device="cuda").half()
    out = torch.empty(G * Q, 256, dtype=torch.float16, device="cuda")
    for _ in range(3): L.self_attn(qk, v, out, G, Q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): L.self_attn(qk, v, out, G, Q)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = G * 8 * Q * Q * 32 * 4
    print(f"self_attn G={G:4d} rows={Q:4d}: {us:7.1f} us  ({fl / us / 1e6:.1f} TFLOP/s)")
