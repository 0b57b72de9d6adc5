"""SAN / BriVIS side path timing: post-split CLIP blocks + tail for one 36-frame clip (cfg 3: Q = 100, cfg 4: Q = 200)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
from openvis_b200.ov_head import SideAdapterBlocks
from openvis_b200.synthetic import seeded_clip_block_params
n = 36
for Q in (100, 200):
    g = torch.Generator().manual_seed(0)
    sd = {f"transformer.resblocks.{k}": v for k, v in seeded_clip_block_params(1).items()}
    sd.update({"ln_post.weight": torch.ones(768), "ln_post.bias": torch.zeros(768), "proj": torch.randn(768, 512, generator=g) * 768 ** -0.5})
    m = SideAdapterBlocks(num_queries=Q).load_clip_visual_state_dict(sd)
    cls = torch.randn(1, n, 768, generator=g).cuda(); pix = torch.randn(n, 768, 14, 14, generator=g).cuda()
    bias = (3 * torch.randn(n, 12, Q, 46, 80, generator=g)).cuda()
    for _ in range(2): m.post_encode_image((cls, pix), bias)
    torch.cuda.synchronize()
    L.reset_timers() if hasattr(L, "reset_timers") else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m.post_encode_image((cls, pix), bias)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    Lt = Q + 1 + 196
    flops = n * Lt * 2 * (768 * 2304 + 768 * 768 + 2 * 768 * 3072) * 3 + n * 12 * 3 * 4 * Lt * 197 * 64
    print(f"Q={Q}: {ms:.3f} ms per 36-frame clip = {n / ms * 1e3:.0f} frames/s, {flops / ms / 1e9:.0f} TFLOP/s on {n * Lt} token rows")
    # attention kernel alone
    qkv = torch.randn(n * Lt, 2304, generator=g).half().cuda(); out = torch.empty(n * Lt, 768, dtype=torch.float16, device="cuda")
    pooled = L.san_pool_bias(bias, (14, 14))
    for _ in range(2): L.san_attn(qkv, pooled, out, n, Q, 196, 12)
    e0.record()
    for _ in range(10): L.san_attn(qkv, pooled, out, n, Q, 196, 12)
    e1.record(); torch.cuda.synchronize()
    print(f"   san_attn_kernel: {e0.elapsed_time(e1) / 10 * 1e3:.0f} us per block")
