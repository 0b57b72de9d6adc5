"""Per-stage device time of the pixel decoder (row f-2) at a BASELINE frame size: python tools/prof_pixdec.py [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from openvis_b200 import _lib as L
from openvis_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
from openvis_b200.synthetic import seeded_pixel_decoder_params

Nf = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ch, Hp, Wp = (256, 512, 1024, 2048), 736, 1280
dev = torch.device("cuda:0")
pd = MSDeformAttnPixelDecoder({f"res{i + 2}": ShapeSpec(channels=c, stride=4 << i) for i, c in enumerate(ch)})
pd.load_state_dict(seeded_pixel_decoder_params(2, in_channels=ch))
pd = pd.to(dev)
g = torch.Generator(device=dev).manual_seed(13)
feats = {f"res{i + 2}": torch.randn(Nf, c, Hp // (4 << i), Wp // (4 << i), generator=g, device=dev) for i, c in enumerate(ch)}
pd.forward_features(feats)
torch.cuda.synchronize()

# wrap every C-ABI wrapper the schedule uses with CUDA events
names = ["nchw_to_tokens_f16", "linear_f16", "linear_ln_f16", "group_norm_tokens", "tokens_to_nchw", "conv3x3_unfold_f16", "mask_logits",
         "cast_f16", "msda_prepare", "ms_deform_attn_forward", "msda_fused_f16"]
acc = {}
orig = {n: getattr(L, n) for n in names}


def wrap(n):
    f = orig[n]

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f(*a, **k)
        e1.record()
        acc.setdefault(n, []).append((e0, e1))
        return r
    return w


for n in names:
    setattr(L, n, wrap(n))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pd.forward_features(feats)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
print(f"pixel decoder, {Nf} frames of {Hp}x{Wp}: {tot:.2f} ms = {Nf / tot * 1e3:.0f} frames/s")
for n, ev in sorted(acc.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
    t = sum(a.elapsed_time(b) for a, b in ev)
    print(f"  {n:26s} x{len(ev):3d}  {t:8.3f} ms  {100 * t / tot:5.1f} %")
