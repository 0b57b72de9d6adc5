"""pixel decoder + Video decoder at one 36-frame 736x1280 clip: fp32 NCHW interface vs fp16 token-major hand-off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L
from openvis_b200.decoder import VideoMultiScaleMaskedTransformerDecoder
from openvis_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
from openvis_b200.synthetic import decoder_param_shapes, seeded_params, seeded_pixel_decoder_params
Nf = int(sys.argv[1]) if len(sys.argv) > 1 else 36
ch, Hp, Wp = (256, 512, 1024, 2048), 736, 1280
dev = torch.device("cuda:0")
pd = MSDeformAttnPixelDecoder({f"res{i + 2}": ShapeSpec(channels=c, stride=4 << i) for i, c in enumerate(ch)})
pd.load_state_dict(seeded_pixel_decoder_params(2, in_channels=ch)); pd = pd.to(dev)
dec = VideoMultiScaleMaskedTransformerDecoder(in_channels=256, mask_classification=True, num_classes=40, hidden_dim=256, num_queries=100,
                                              nheads=8, dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256,
                                              enforce_input_project=False, num_frames=Nf).eval().to(dev)
dec.load_state_dict(seeded_params(decoder_param_shapes("video", num_classes=40), seed=0))
g = torch.Generator(device=dev).manual_seed(13)
feats = {f"res{i + 2}": torch.randn(Nf, c, Hp // (4 << i), Wp // (4 << i), generator=g, device=dev) for i, c in enumerate(ch)}
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]
mf, _, ms = pd.forward_features(feats)
tok = pd.forward_tokens(feats)
t_pd, t_pdt = timed(lambda: pd.forward_features(feats)), timed(lambda: pd.forward_tokens(feats))
t_dec, t_dect = timed(lambda: dec(ms, mf)), timed(lambda: dec.forward_tokens(tok))
print(f"{Nf} frames of {Hp}x{Wp}: pixel decoder {t_pd:.2f} ms (NCHW fp32 out) / {t_pdt:.2f} ms (token hand-off); "
      f"Video decoder {t_dec:.2f} ms (NCHW fp32 in) / {t_dect:.2f} ms (tokens in); "
      f"pipeline {Nf / (t_pd + t_dec) * 1e3:.0f} -> {Nf / (t_pdt + t_dect) * 1e3:.0f} frames/s; decoder alone {Nf / t_dec * 1e3:.0f} -> {Nf / t_dect * 1e3:.0f} frames/s")
