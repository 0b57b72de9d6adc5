"""End-to-end parity (GPU): the B200 decoders against the CPU oracle (oracle/decoder_ref.py, pinned to the
reference by tests/test_oracle_golden.py) on identical seeded weights and inputs, and against the committed
golden fixtures generated from the reference's own modules.

Tolerances (north_star): fp16 operands / fp32 accumulation on the GPU vs fp32 on the CPU:
  * binarised attention masks of every layer:  >= 99.9 % agreement
  * top-1 class per query:                     >= 99.9 % agreement
  * mask logits:   |err| <= 0.25 absolute on values of magnitude ~30-50 for 99.9 % of the elements
  * class logits:  |err| <= 3e-2 ; pred_embeds |err| <= 3e-2
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from openvis_b200 import _lib as L  # noqa: E402
from openvis_b200 import decoder as D  # noqa: E402
from oracle import decoder_ref as O  # noqa: E402

KINDS = {
    "frame": "FrameMultiScaleMaskedTransformerDecoder",
    "video": "VideoMultiScaleMaskedTransformerDecoder",
    "san_frame": "SideAdapterFrameMultiScaleMaskedTransformerDecoder",
    "san_video": "SideAdapterVideoMultiScaleMaskedTransformerDecoder",
}


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def build(kind, Q=100, pseed=0):
    kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=Q, nheads=8,
              dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)
    if kind.startswith("san"):
        kw["clip_heads"] = 12
    m = D.TRANSFORMER_DECODER_REGISTRY[KINDS[kind]](**kw)
    P = O.seeded_params(O.decoder_param_shapes(kind, Q=Q), pseed)
    m.load_state_dict(P)
    return m.cuda().eval(), P


def unpack_bits(bits, keys):
    """bits [G, W, Q] int32 -> bool [G, Q, keys]"""
    r = torch.arange(keys, device=bits.device)
    return (((bits.long()[:, r // 32, :] >> (r % 32)[None, :, None]) & 1).bool()).permute(0, 2, 1)


def frac_within(a, b, tol):
    return ((a.float() - b.float()).abs() <= tol).float().mean().item()


def run_case(kind, T, Hp, Wp, Q=100, pseed=0, iseed=1234):
    m, P = build(kind, Q, pseed)
    x, mf = O.seeded_inputs(T, Hp, Wp, seed=iseed)
    ref = O.decoder_forward(P, x, mf, kind=kind)
    m.debug_capture = []
    out = m([t.cuda() for t in x], mf.cuda())
    torch.cuda.synchronize()
    return m, ref, out


# Small spatial sizes have 4-60 keys per attention; a single mask bit that flips at fp16 noise level (|logit| < 0.01)
# then changes that query's attention output visibly and the perturbation propagates through self-attention.  The
# strict north_star bars are asserted at the real shapes (config-1 and larger); the tiny golden / Q=200 cases use the
# loose set below, which still catches any indexing / layout / weight-mapping error (those give O(1) mismatches).
STRICT = dict(mask=0.999, pm_tol=0.25, pm_frac=0.999, sign=0.999, logit=3e-2, emb_frac=0.999, bias_frac=0.999)
LOOSE = dict(mask=0.995, pm_tol=0.25, pm_frac=0.95, sign=0.995, logit_frac=0.97, emb_frac=0.95, bias_frac=0.97)


def check_case(kind, m, ref, out, T, Hp, Wp, Q, tol=STRICT):
    # ---- per-layer binarised attention masks
    sizes = [(Hp // 32 * 2 ** l) * (Wp // 32 * 2 ** l) for l in range(3)]
    agree = []
    for hidx, level, bits, flags in m.debug_capture:
        keys = sizes[level] * (T if kind.endswith("video") else 1)
        got = unpack_bits(bits, keys).cpu()
        want = ref["attn_masks"][hidx]
        assert got.shape == want.shape, (got.shape, want.shape)
        agree.append((got == want).float().mean().item())
        assert torch.equal(flags.bool().cpu(), (~got).any(-1))
    assert len(agree) == 9
    assert min(agree) >= tol['mask'], agree
    # ---- final mask logits
    pm, rm = out["pred_masks"].cpu(), ref["pred_masks"]
    assert pm.shape == rm.shape
    assert frac_within(pm, rm, tol['pm_tol']) >= tol['pm_frac'], ((pm - rm).abs().max().item(), frac_within(pm, rm, 0.25))
    assert ((pm > 0) == (rm > 0)).float().mean().item() >= tol['sign']
    # ---- class logits / attention biases / embeds
    if "pred_logits" in ref:
        pl, rl = out["pred_logits"].cpu(), ref["pred_logits"]
        assert pl.shape == rl.shape
        if 'logit_frac' in tol:      # chaotic small cases: fraction of agreeing logits instead of the worst one
            assert frac_within(pl, rl, 0.05) >= tol['logit_frac'], (pl - rl).abs().max().item()
        else:
            assert (pl - rl).abs().max().item() <= tol['logit'], (pl - rl).abs().max().item()
        assert (pl.argmax(-1) == rl.argmax(-1)).float().mean().item() >= 0.999
    if "class_attn_biases" in ref:
        pb, rb = out["class_attn_biases"].cpu(), ref["class_attn_biases"]
        assert pb.shape == rb.shape
        assert frac_within(pb, rb, 3e-2) >= tol['bias_frac'], (pb - rb).abs().max().item()
    if "pred_embeds" in ref:
        pe, re_ = out["pred_embeds"].cpu(), ref["pred_embeds"]
        assert pe.shape == re_.shape
        assert frac_within(pe, re_, 3e-2) >= tol['emb_frac'], (pe - re_).abs().max().item()
    if "attn_feats" in ref:
        assert frac_within(out["attn_feats"].cpu(), ref["attn_feats"], 5e-3) >= 0.999
    return agree


@pytest.mark.parametrize("kind", ["frame", "video", "san_frame", "san_video"])
def test_decoder_parity_cfg1_shape(kind):
    """BASELINE config-1 shape: 5 frames, 360x640 padded to 384x640, Q = 100."""
    T, Hp, Wp = 5, 384, 640
    m, ref, out = run_case(kind, T, Hp, Wp)
    agree = check_case(kind, m, ref, out, T, Hp, Wp, 100)
    print(kind, "mask agreement per layer:", ["%.5f" % a for a in agree])


def test_decoder_parity_q200():
    """SAN-online uses 200 queries (BASELINE config 4); small spatial size."""
    T, Hp, Wp = 2, 96, 160
    m, ref, out = run_case("san_frame", T, Hp, Wp, Q=200, pseed=4, iseed=99)
    check_case("san_frame", m, ref, out, T, Hp, Wp, 200, tol=LOOSE)


def test_san_frame_q200_cfg1_shape_strict():
    """SAN-online with 200 queries (BASELINE config 4) at the config-1 spatial size: strict north_star bars."""
    T, Hp, Wp = 2, 384, 640
    m, ref, out = run_case("san_frame", T, Hp, Wp, Q=200, pseed=4, iseed=99)
    check_case("san_frame", m, ref, out, T, Hp, Wp, 200)


def test_video_decoder_cfg2_full_shape_against_oracle():
    """BASELINE configs[1] at its own shape -- 36 frames of 720x1280 padded to 736x1280, Q = 100, joint attention over
    33 120 / 132 480 / 529 920 keys -- against the CPU oracle (one fp32 pass, ~10-20 s on the box's cores): strict bars."""
    T, Hp, Wp = 36, 736, 1280
    m, ref, out = run_case("video", T, Hp, Wp)
    agree = check_case("video", m, ref, out, T, Hp, Wp, 100)
    print("cfg2 full shape, mask agreement per layer:", ["%.5f" % a for a in agree])


def test_san_frame_cfg4_full_resolution_q200_against_oracle():
    """BASELINE configs[3]'s decoder at its own resolution (736x1280, Q = 200); two frames (frames are independent)."""
    T, Hp, Wp = 2, 736, 1280
    m, ref, out = run_case("san_frame", T, Hp, Wp, Q=200, pseed=4, iseed=99)
    check_case("san_frame", m, ref, out, T, Hp, Wp, 200)


def test_frame_decoder_cfg5b_full_resolution_against_oracle():
    """The Frame decoder at 736x1280, Q = 100 (configs[4], Frame reading), three frames."""
    T, Hp, Wp = 3, 736, 1280
    m, ref, out = run_case("frame", T, Hp, Wp)
    check_case("frame", m, ref, out, T, Hp, Wp, 100)


def test_full_size_clip_properties():
    """BASELINE config 2 at full size (36 frames, 736x1280, Q = 100): too large for the CPU oracle to finish in
    seconds, so the check is through size-independent properties of the path."""
    T, Hp, Wp, Q = 36, 736, 1280, 100
    m, P = build("video", Q, 0)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = [torch.randn(T, 256, Hp // 32 * 2 ** l, Wp // 32 * 2 ** l, generator=g, device="cuda") for l in range(3)]
    mf = torch.randn(T, 256, Hp // 4, Wp // 4, generator=g, device="cuda")
    m.debug_capture = []
    out = m(x, mf)
    pm = out["pred_masks"]
    assert pm.shape == (1, Q, T, Hp // 4, Wp // 4) and bool(torch.isfinite(pm).all())
    # (1) the final head is linear in the mask features: pred_masks == mask_embed . F  (fp16 operands)
    frames = (0, 17, 35)
    # recover mask_embed from the decoder state: re-run the MLP on the saved decoder_norm output of the last head
    ws = next(iter(m._ws.values()))
    d = ws["d16"][m.num_layers].float()
    me = d
    for i, lyr in enumerate(m.mask_embed.layers):
        me = me.half().float() @ lyr.weight.half().float().T + lyr.bias
        if i < 2:
            me = me.relu()
    me = me.half().float()
    for t in frames:
        ref_t = torch.einsum("qc,chw->qhw", me, mf[t].half().float())
        assert (pm[0, :, t] - ref_t).abs().max().item() < 0.05
    # (2) the epilogue's non-empty flags agree with the logits it wrote
    assert torch.equal(out["mask_valid"].bool(), (pm[0] > 0).flatten(2).any(-1).T)
    # (3) every layer's attention mask leaves at least one key for rows flagged as non-empty, bits are ~50 % dense
    for hidx, level, bits, flags in m.debug_capture:
        dens = torch.stack([((bits >> s) & 1).float().mean() for s in range(0, 32, 5)]).mean().item()
        assert 0.3 < dens < 0.7
        assert bool(flags.bool().all())
    # (4) determinism, and (5) two clips in one call == the clips one by one
    out2 = m(x, mf)
    assert torch.equal(out2["pred_masks"], pm) and torch.equal(out2["pred_logits"], out["pred_logits"])
    del out2
    half = T // 2
    m.clips_per_call = 1
    a = m([t[:half] for t in x], mf[:half])["pred_masks"].clone()
    m.clips_per_call = 2
    both = m([torch.cat([t[:half], t[:half]]) for t in x], torch.cat([mf[:half], mf[:half]]))["pred_masks"]
    # the key-split count of the attention differs between one and two clips per call (148 / clips), so the partial
    # sums are merged in a different order: equal up to fp32 summation order, not bit-exact
    assert torch.equal(both[0], both[1])
    assert frac_within(both[0], a[0], 0.05) >= 0.999


def test_frame_outputs_api():
    T, Hp, Wp = 2, 64, 96
    m, ref, out = run_case("frame", T, Hp, Wp)
    assert set(out.keys()) == {"pred_logits", "pred_masks", "mask_feats", "ms_feats", "ms_pos", "size_list", "aux_outputs",
                               "pred_embeds", "mask_valid"}
    pm = out["pred_masks"][0]                                     # [Q, T, H, W]
    assert torch.equal(out["mask_valid"].bool(), (pm > 0).flatten(2).any(-1).T)
    for a, b in zip(out["ms_feats"], ref["ms_feats"]):
        assert torch.allclose(a.cpu(), b, atol=1e-6)
    for a, b in zip(out["ms_pos"], ref["ms_pos"]):
        assert torch.allclose(a.cpu(), b, atol=1e-5)
    assert [tuple(s) for s in out["size_list"]] == [tuple(s) for s in ref["size_list"]]
    # lazily computed aux_outputs match the oracle's intermediate heads
    assert len(out["aux_outputs"]) == 9
    for i in (0, 4, 8):
        a, b = out["aux_outputs"][i], ref["aux_outputs"][i]
        assert frac_within(a["pred_masks"].cpu(), b["pred_masks"], 0.25) >= LOOSE["pm_frac"]
        # 64x96 input = 6..96 keys per attention: chaotic (see LOOSE); a query whose mask bit flipped moves by ~0.5, so
        # the check is on the fraction of class logits that agree, not on the worst one
        assert frac_within(a["pred_logits"].cpu(), b["pred_logits"], 0.05) >= 0.97
    # a second forward invalidates un-read aux entries of the first
    out2 = m([t.cuda() for t in O.seeded_inputs(T, Hp, Wp, seed=5)[0]], O.seeded_inputs(T, Hp, Wp, seed=5)[1].cuda())
    with pytest.raises(RuntimeError):
        out["aux_outputs"][1]
    assert out2["aux_outputs"][1]["pred_masks"].shape == ref["aux_outputs"][1]["pred_masks"].shape


@pytest.mark.parametrize("kind", ["video", "san_video"])
def test_video_multi_clip_call_equals_single_clip_calls(kind):
    """clips_per_call > 1 (throughput extension) must reproduce the per-clip results."""
    T, Hp, Wp, Q = 3, 128, 192, 100
    m, P = build(kind, Q, 0)
    m.use_chain = False              # one schedule for both calls: bit-equality holds within a schedule (single-clip calls
                                     # would otherwise take the chained query side, a different fp32 summation order)
    clips = [O.seeded_inputs(T, Hp, Wp, seed=50 + i) for i in range(2)]
    singles = []
    for x, mf in clips:
        o = m([t.cuda() for t in x], mf.cuda())
        singles.append({k: o[k].clone() for k in ("pred_masks", "mask_valid") + (("pred_logits",) if kind == "video" else ("class_attn_biases",))})
    m.clips_per_call = 2
    xs = [torch.cat([clips[0][0][l], clips[1][0][l]]).cuda() for l in range(3)]
    mfs = torch.cat([clips[0][1], clips[1][1]]).cuda()
    o = m(xs, mfs)
    assert o["pred_masks"].shape == (2, Q, T, Hp // 4, Wp // 4)
    for g in range(2):
        assert torch.equal(o["pred_masks"][g], singles[g]["pred_masks"][0])
        assert torch.equal(o["mask_valid"][g * T:(g + 1) * T], singles[g]["mask_valid"])
        if kind == "video":
            assert torch.equal(o["pred_logits"][g], singles[g]["pred_logits"][0])
        else:
            assert torch.equal(o["class_attn_biases"][g], singles[g]["class_attn_biases"][0])


GOLDEN_CASES = [("dec_frame_q100", "frame"), ("dec_video_q100", "video"), ("dec_san_frame_q100", "san_frame"),
                ("dec_san_video_q100", "san_video"), ("dec_frame_q200", "frame")]


@pytest.mark.parametrize("name,kind", GOLDEN_CASES)
def test_decoder_matches_reference_golden(name, kind, golden_dir):
    """Directly against outputs of the reference's own modules (tests/golden, oracle/make_golden.py)."""
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    T, Hp, Wp, Q, pseed, iseed = [int(v) for v in gold["meta"]]
    m, _ = build(kind, Q, pseed)
    x, mf = O.seeded_inputs(T, Hp, Wp, seed=iseed)
    out = m([t.cuda() for t in x], mf.cuda())
    g = lambda k: torch.as_tensor(gold[k]).float()
    t = LOOSE                                         # fixtures are small (128x192 inputs): see LOOSE above
    pm, gm = out["pred_masks"].cpu(), g("pred_masks")
    assert frac_within(pm, gm, t["pm_tol"]) >= t["pm_frac"]
    assert ((pm > 0) == (gm > 0)).float().mean().item() >= t["sign"]
    if "pred_logits" in gold.files:
        pl, gl = out["pred_logits"].cpu(), g("pred_logits")
        assert frac_within(pl, gl, 0.05) >= t["logit_frac"], (pl - gl).abs().max().item()
        assert (pl.argmax(-1) == gl.argmax(-1)).float().mean().item() >= 0.99
    if "pred_embeds" in gold.files:
        assert frac_within(out["pred_embeds"].cpu(), g("pred_embeds"), 3e-2) >= t["emb_frac"]
    if "class_attn_biases" in gold.files:
        assert frac_within(out["class_attn_biases"].cpu(), g("class_attn_biases"), 3e-2) >= t["bias_frac"]
    # the first head depends on no attention at all: it must match tightly
    assert frac_within(out["aux_outputs"][0]["pred_masks"].cpu(), g("aux0_pred_masks"), 0.08) >= 0.9999
    for i in (4, 8):
        assert frac_within(out["aux_outputs"][i]["pred_masks"].cpu()[..., ::4, ::4], g(f"aux{i}_pred_masks"), 0.3) >= t["pm_frac"]


def test_no_cpu_fallback_and_training_refused():
    m, _ = build("frame")
    x, mf = O.seeded_inputs(1, 64, 64)
    with pytest.raises(Exception):
        m(x, mf)                                   # CPU tensors
    m.train()
    with pytest.raises(RuntimeError):
        m([t.cuda() for t in x], mf.cuda())


@pytest.mark.parametrize("kind", ["video", "san_frame"])
def test_cuda_graph_replay_equals_eager(kind):
    """The layer loop replayed as a CUDA graph (second and later calls of a workspace) gives bit-identical outputs to
    the eager launches, also after the inputs change between calls."""
    T, Hp, Wp, Q = 2, 128, 192, 100
    m, P = build(kind, Q, 0)
    ins = [O.seeded_inputs(T, Hp, Wp, seed=70 + i) for i in range(3)]
    m.use_cuda_graph = False
    eager = []
    for x, mf in ins:
        out = m([t.cuda() for t in x], mf.cuda())
        eager.append((out["pred_masks"].clone(), out["pred_embeds"].clone() if "pred_embeds" in out else None))
    m.use_cuda_graph = True
    n0 = L.launch_count()
    for (x, mf), (pm, pe) in zip(ins, eager):          # call 1 is still eager (already warm), 2 captures, 3 replays
        out = m([t.cuda() for t in x], mf.cuda())
        assert torch.equal(out["pred_masks"], pm)
        if pe is not None:
            assert torch.equal(out["pred_embeds"], pe)
    per_call = (L.launch_count() - n0) / 3
    assert per_call > 50, per_call                      # replayed launches are still counted (one group: chained layers)


@pytest.mark.parametrize("mode", [True, "wide"])
@pytest.mark.parametrize("kind,T", [("video", 2), ("frame", 3), ("san_frame", 2)])
def test_query_side_chain_equals_launch_per_op(kind, T, mode):
    """The query-side chain (one launch per layer: csrc/chain.cuh) against the launch-per-operation schedule on the same
    weights and inputs: same kernels' arithmetic, only the LayerNorm GEMMs take the fused-epilogue path instead of the
    split-K one, so the results agree to fp32 summation order (at the config-1 resolution: tiny inputs amplify a single
    flipped mask bit, see LOOSE)."""
    Hp, Wp, Q = 384, 640, 100
    m, P = build(kind, Q, 0)
    x, mf = O.seeded_inputs(T, Hp, Wp, seed=31)
    xs, mfs = [t.cuda() for t in x], mf.cuda()
    m.use_cuda_graph = False
    m.use_chain = False
    n0 = L.launch_count()
    a = m(xs, mfs)
    n_ops = L.launch_count() - n0
    ref = {k: a[k].clone() for k in ("pred_masks",) + (("pred_logits",) if "pred_logits" in a else ("class_attn_biases",))}
    m.use_chain = mode               # True: one CTA per group; "wide": tiles over all CTAs + grid barriers (cooperative launch)
    n0 = L.launch_count()
    b = m(xs, mfs)
    n_chain = L.launch_count() - n0
    assert n_chain <= n_ops - 100, (n_chain, n_ops)             # ~14 launches per layer become one
    for k, v in ref.items():
        assert v.shape == b[k].shape
        assert frac_within(b[k], v, 0.05) >= 0.995, (k, (b[k] - v).abs().max().item())
    ref_o = O.decoder_forward(P, x, mf, kind=kind, return_attn_masks=False)
    assert frac_within(b["pred_masks"].cpu(), ref_o["pred_masks"], 0.25) >= STRICT["pm_frac"]
    # graph replay of the chained schedule
    m.use_cuda_graph = True
    outs = [m(xs, mfs)["pred_masks"].clone() for _ in range(3)]
    assert torch.equal(outs[1], outs[0]) and torch.equal(outs[2], outs[0])
