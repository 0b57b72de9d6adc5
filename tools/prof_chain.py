"""Per-phase device time of the query-side chain (csrc/chain.cuh) of one decoder layer: each phase launched on its own
(ovis_chain_run(first + j, 1)), and the layer's chain as one launch, for the Video decoder (G = clips) and the Frame decoder
(G = frames).  python tools/prof_chain.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L, decoder as D
from openvis_b200.synthetic import decoder_param_shapes, seeded_params

NAMES = ["xo+LN", "sqk", "sv", "self_attn", "so+LN", "ffn1", "ffn2+LN+dn", "xq(next)", "me0", "me1", "me2"]


def ev(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for kind, T, clips, Hp, Wp in (("video", 8, 4, 128, 192), ("frame", 144, 1, 128, 192), ("frame", 36, 1, 128, 192), ("video", 2, 1, 128, 192)):
    kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=100, nheads=8,
              dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)
    cls = D.VideoMultiScaleMaskedTransformerDecoder if kind == "video" else D.FrameMultiScaleMaskedTransformerDecoder
    m = cls(**kw)
    m.load_state_dict(seeded_params(decoder_param_shapes(kind, Q=100), 0))
    m = m.cuda().eval()
    m.clips_per_call = clips
    m.use_cuda_graph = False
    m.use_chain = True
    g = torch.Generator(device="cuda").manual_seed(1)
    x = [torch.randn(T, 256, Hp // 32 * 2 ** l, Wp // 32 * 2 ** l, generator=g, device="cuda") for l in range(3)]
    mf = torch.randn(T, 256, Hp // 4, Wp // 4, generator=g, device="cuda")
    m(x, mf)
    ws = next(iter(m._ws.values()))
    c = ws["chain"]
    ch = c["chain"]
    first, count = c["first"][3], c["count"][3]
    per = [ev(lambda j=j: ch.run(first + j, 1)) for j in range(count)]
    whole = ev(lambda: ch.run(first, count))
    tr = ch.run_traced(first, count).cpu()
    cyc = (tr[1:count + 1] - tr[:count]).tolist()
    print("    in-kernel cycles per phase (CTA of group 0): " + ", ".join(f"{NAMES[j]} {cyc[j]}" for j in range(count)) + f"  total {sum(cyc)}")
    print(f"{kind} G={ws['G']} (rows {ws['R']}): layer chain {whole:.1f} us as one launch; phases alone: " +
          ", ".join(f"{NAMES[j]} {per[j]:.1f}" for j in range(count)) + f"  (sum {sum(per):.1f})")
    # the launch-per-op schedule of the same layer for comparison
    m.use_chain = False
    W = m._weights()
    lw = W["layers"][3]
    def ops():
        L.linear_ln_f16(ws["att16"], lw["xo_w"], lw["xo_b"], ws["z32"], lw["ln_x"], None, W["qe"], y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"], split_ws=ws["split"])
        L.linear_f16(ws["ze16"], lw["sqk_w"], lw["sqk_b"], out=ws["qk16"])
        L.linear_f16(ws["z16"], lw["sv_w"], lw["sv_b"], out=ws["v16"])
        L.self_attn(ws["qk16"], ws["v16"], ws["sa16"], ws["G"], 100)
        L.linear_ln_f16(ws["sa16"], lw["so_w"], lw["so_b"], ws["z32"], lw["ln_s"], None, W["qe"], y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"], split_ws=ws["split"])
        L.linear_f16(ws["z16"], lw["f1_w"], lw["f1_b"], relu=True, out=ws["h16"])
        L.linear_ln_f16(ws["h16"], lw["f2_w"], lw["f2_b"], ws["z32"], lw["ln_f"], W["dn"], W["qe"], y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"], d32=ws["d32"], d16=ws["d16"][4], split_ws=ws["split"])
        m._mlp3(W["mask_embed"], ws["d16"][4], ws["m1"], ws["m2"], ws["me16"])
        L.linear_f16(ws["ze16"], lw["xq_w"], lw["xq_b"], scale=0.25, out=ws["q16"])
    print(f"    launch-per-op schedule of the same layer: {ev(ops):.1f} us")
    # the wide chain (tiles of every phase over all CTAs of one cooperative launch, grid barriers between phases)
    m.use_chain = "wide"
    m(x, mf)
    cw = ws["chain"]
    assert ws["chain_mode"] == "wide"
    chw = cw["chain"]
    fw, nw = cw["first"][3], cw["count"][3]
    perw = [ev(lambda j=j: chw.run(fw + j, 1)) for j in range(nw)]
    print(f"    wide chain: {ev(lambda: chw.run(fw, nw)):.1f} us as one launch; phases alone: " +
          ", ".join(f"{NAMES[j]} {perw[j]:.1f}" for j in range(nw)) + f"  (sum {sum(perw):.1f})")
