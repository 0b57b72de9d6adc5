"""Per-configuration decoder timing (BASELINE configs 1, 3, 4 and the Frame variant of 2): frames/s and per-family time."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L, decoder as D
from openvis_b200.synthetic import decoder_param_shapes, seeded_params
CASES = [("video", 5, 384, 640, 100), ("frame", 36, 736, 1280, 100), ("san_frame", 36, 384, 640, 100),
         ("san_frame", 36, 736, 1280, 200), ("san_video", 36, 736, 1280, 100)]
CLS = {"video": D.VideoMultiScaleMaskedTransformerDecoder, "frame": D.FrameMultiScaleMaskedTransformerDecoder,
       "san_frame": D.SideAdapterFrameMultiScaleMaskedTransformerDecoder, "san_video": D.SideAdapterVideoMultiScaleMaskedTransformerDecoder}
for kind, T, Hp, Wp, Q in CASES:
    kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=Q, nheads=8,
              dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)
    if kind.startswith("san"):
        kw["clip_heads"] = 12
    m = CLS[kind](**kw)
    m.load_state_dict(seeded_params(decoder_param_shapes(kind, Q=Q), 0))
    m = m.cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = [torch.randn(T, 256, Hp // 32 * 2 ** l, Wp // 32 * 2 ** l, generator=g, device="cuda") for l in range(3)]
    mf = torch.randn(T, 256, Hp // 4, Wp // 4, generator=g, device="cuda")
    for _ in range(3):
        out = m(x, mf)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(5):
        out = m(x, mf)
        _ = out["pred_masks"]
    g1.record(); torch.cuda.synchronize()
    ms_graph = g0.elapsed_time(g1) / 5
    ev = []
    L.PROFILE = ev
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = m(x, mf)
        _ = out["pred_masks"]
    e1.record(); torch.cuda.synchronize()
    L.PROFILE = None
    ms = e0.elapsed_time(e1) / 3
    fam = collections.defaultdict(float)
    for f, a, b in ev:
        fam[f] += a.elapsed_time(b) / 3
    print(f"{kind:10s} T={T} {Hp}x{Wp} Q={Q}: {ms_graph:.2f} ms per call = {T / ms_graph * 1e3:.0f} frames/s with the layer loop as a CUDA graph; "
          f"eager {ms:.2f} ms = {T / ms * 1e3:.0f} frames/s | " +
          ", ".join(f"{k} {v:.2f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))
    del m, x, mf, out
    torch.cuda.empty_cache()
