"""Device post-processing timing at the cfg-2 shape: top-10 + composed up-sample / crop / resize / threshold / bit-pack."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
Q, K, T = 100, 40, 36
g = torch.Generator().manual_seed(0)
scores = torch.rand(Q, K, generator=g).softmax(-1).cuda()
masks = (torch.randn(Q, T, 184, 320, generator=g) * 4).cuda()
for out in ((720, 1280), (1080, 1920)):
    for _ in range(2):
        vs, qi, lb, en = L.topk_scores(scores, 10); bits = L.mask_postprocess(masks, qi, (736, 1280), (720, 1280), out)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(10): vs, qi, lb, en = L.topk_scores(scores, 10)
    e1.record()
    for _ in range(10): bits = L.mask_postprocess(masks, qi, (736, 1280), (720, 1280), out)
    e2.record(); torch.cuda.synchronize()
    px = 10 * T * out[0] * out[1]
    ms = e1.elapsed_time(e2) / 10
    print(f"out={out}: topk {e0.elapsed_time(e1) / 10 * 1e3:.0f} us, mask_postprocess {ms * 1e3:.0f} us = {px / ms / 1e6:.1f} Gpixel/s, packed {bits.numel() * 4 / 1e6:.1f} MB (bool would be {px / 1e6:.0f} MB)")
