"""CPU tests of the host-side mirror of the reference interface (no kernels run here)."""
import types

import pytest
import torch

from openvis_b200 import decoder as D
from oracle import decoder_ref as O

KW = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=100, nheads=8,
          dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)


def _cfg(name="FrameMultiScaleMaskedTransformerDecoder", queries=100):
    ns = types.SimpleNamespace
    return ns(MODEL=ns(SEM_SEG_HEAD=ns(NUM_CLASSES=1, MASK_DIM=256),
                       MASK_FORMER=ns(HIDDEN_DIM=256, NUM_OBJECT_QUERIES=queries, NHEADS=8, DIM_FEEDFORWARD=2048,
                                      DEC_LAYERS=10, PRE_NORM=False, ENFORCE_INPUT_PROJ=False,
                                      TRANSFORMER_DECODER_NAME=name),
                       CLIP_ADAPTER=ns(CLIP_NUM_HEADS=12, CLIP_EMBED_DIMS=512)),
              INPUT=ns(SAMPLING_FRAME_NUM=2))


@pytest.mark.parametrize("name,kind", [("FrameMultiScaleMaskedTransformerDecoder", "frame"),
                                       ("VideoMultiScaleMaskedTransformerDecoder", "video"),
                                       ("SideAdapterFrameMultiScaleMaskedTransformerDecoder", "san_frame"),
                                       ("SideAdapterVideoMultiScaleMaskedTransformerDecoder", "san_video")])
def test_state_dict_contract(name, kind):
    """Parameter names and shapes equal the reference's; built through the registry from a cfg,
    exactly as MaskFormerHead.from_config does (mask_former_head.py:112-116)."""
    m = D.build_transformer_decoder(_cfg(name), 256, True)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == O.decoder_param_shapes(kind)
    assert m.num_layers == 9 and m.num_queries == 100 and m.num_heads == 8
    m.load_state_dict(O.seeded_params(O.decoder_param_shapes(kind), 0))      # strict


def test_registry_has_reference_names():
    want = {"VideoMultiScaleMaskedTransformerDecoder", "EmbeddingVideoMultiScaleMaskedTransformerDecoder",
            "ProposalVideoMultiScaleMaskedTransformerDecoder", "FrameMultiScaleMaskedTransformerDecoder",
            "EmbeddingFrameMultiScaleMaskedTransformerDecoder", "ProposalFrameMultiScaleMaskedTransformerDecoder",
            "SideAdapterVideoMultiScaleMaskedTransformerDecoder", "SideAdapterFrameMultiScaleMaskedTransformerDecoder"}
    assert want <= set(D.TRANSFORMER_DECODER_REGISTRY)
    target = {}
    D.register_into(target)
    assert want <= set(target)


def test_embedding_and_proposal_variants():
    e = D.EmbeddingFrameMultiScaleMaskedTransformerDecoder(clip_dims=512, **KW)
    assert e.class_embed.layers[0].weight.shape == (1024, 256) and e.class_embed.layers[1].weight.shape == (512, 1024)
    p = D.ProposalVideoMultiScaleMaskedTransformerDecoder(**KW)
    assert p.class_embed.weight.shape == (2, 256)
    e2 = D.build_transformer_decoder(_cfg("EmbeddingVideoMultiScaleMaskedTransformerDecoder"), 256, True)
    assert e2.class_embed.layers[1].weight.shape == (512, 1024)


def test_legacy_static_query_key_is_upgraded():
    m = D.FrameMultiScaleMaskedTransformerDecoder(**KW)
    sd = m.state_dict()
    sd["static_query.weight"] = sd.pop("query_feat.weight") + 1.0
    if hasattr(sd, "_metadata"):
        sd._metadata[""] = {"version": 1}
    m2 = D.FrameMultiScaleMaskedTransformerDecoder(**KW)
    m2.load_state_dict(sd)
    assert torch.equal(m2.query_feat.weight, sd["query_feat.weight"] if "query_feat.weight" in sd else m.query_feat.weight + 1.0)


def test_unsupported_configs_fail_loudly():
    for bad in (dict(hidden_dim=128, mask_dim=128, in_channels=128), dict(nheads=4), dict(pre_norm=True),
                dict(enforce_input_project=True), dict(num_queries=300)):
        with pytest.raises(NotImplementedError):
            D.FrameMultiScaleMaskedTransformerDecoder(**{**KW, **bad})


def test_position_tables_match_oracle():
    for (h, w) in ((12, 20), (23, 40), (4, 6)):
        a = D.sine_pos_2d(h, w, "cpu")                          # [h*w, 256]
        b = O.sine_pos_2d(h, w).flatten(1).T
        assert torch.allclose(a, b, atol=1e-6)
    T, h, w = 5, 6, 8
    p3 = O.sine_pos_3d(T, h, w).flatten(2).permute(0, 2, 1)      # [T, hw, C]
    mine = D.sine_pos_2d(h, w, "cpu")[None] + D.sine_pos_z(T, "cpu")[:, None]
    assert torch.allclose(mine, p3, atol=1e-6)


def test_lazy_containers():
    calls = []
    aux = D.LazyAuxOutputs(3, lambda i: calls.append(i) or {"i": i})
    assert len(aux) == 3 and calls == []
    assert aux[1] == {"i": 1} and aux[1] == {"i": 1} and calls == [1]
    assert aux[-1] == {"i": 2}
    assert [a["i"] for a in aux] == [0, 1, 2] and sorted(calls) == [0, 1, 2]
    d = D._LazyDict()
    d["a"] = D._lazy(lambda: calls.append("a") or 7)
    d["b"] = 3
    assert d["a"] == 7 and d["a"] == 7 and calls.count("a") == 1
    assert dict(d.items()) == {"a": 7, "b": 3}


def test_inference_only_and_no_cpu_path():
    m = D.FrameMultiScaleMaskedTransformerDecoder(**KW)
    x, mf = O.seeded_inputs(1, 64, 64)
    with pytest.raises(RuntimeError):
        m(x, mf)                                  # training mode
    m.eval()
    with pytest.raises(Exception) as ei:
        m(x, mf)                                  # CPU tensors: refused, never computed on the host
    assert "CPU" in str(ei.value) or "CUDA" in str(ei.value)
    with pytest.raises(NotImplementedError):
        m([x[0], x[1], x[2][..., :-1]], mf)       # not the stride-32/16/8 pyramid of a /32-padded input


def test_resampler_state_dict_contract():
    """TemporalInstanceResampler: the reference's parameter names / shapes
    (openvis/modeling/resampler.py:191-236), loadable strictly; CPU calls are refused, never computed on the host."""
    from openvis_b200.synthetic import resampler_param_shapes, seeded_resampler_params
    from openvis_b200.temporal import TemporalInstanceResampler, batch_video_match_via_embeds
    from openvis_b200 import _lib as L
    m = TemporalInstanceResampler(hidden_dim=256, feed_dim=2048, nheads=8, nlayers=6)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == resampler_param_shapes()
    m.load_state_dict(seeded_resampler_params(0))
    assert sum(p.numel() for p in m.parameters()) == 11437568          # the reference module's count
    with pytest.raises(RuntimeError):
        m.train()(torch.zeros(1, 2, 3, 256), torch.zeros(2, 256, 8, 8), torch.zeros(2, 12, 256, 2, 2))
    with pytest.raises(L.OvisError):
        m.eval()(torch.zeros(1, 2, 3, 256), torch.zeros(2, 256, 8, 8), torch.zeros(2, 12, 256, 2, 2))
    with pytest.raises(L.OvisError):
        batch_video_match_via_embeds(torch.randn(1, 2, 5, 256))
    with pytest.raises(NotImplementedError):
        TemporalInstanceResampler(hidden_dim=128)


def test_shared_operands_bookkeeping():
    """decoder.shared_operands hands its fp16 operand copies only to a consumer holding exactly the tensors of the most
    recent forward (identity + version), and does not keep those tensors alive."""
    import gc
    import weakref
    m = D.SideAdapterFrameMultiScaleMaskedTransformerDecoder(
        in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=10, nheads=8,
        dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2,
        clip_heads=12)
    mf, af = torch.zeros(2, 256, 8, 8), torch.zeros(2, 12, 256, 2, 2)
    ft, af16 = torch.zeros(1), torch.zeros(1)
    assert m.shared_operands(mf, af) is None                                  # no forward yet
    m._generation = 3
    m._last = dict(gen=3, mf=weakref.ref(mf), mf_ver=mf._version, ft=ft, af32=weakref.ref(af), af_ver=af._version,
                   af16=af16)
    got = m.shared_operands(mf, af)
    assert got is not None and got[0] is ft and got[1] is af16
    assert m.shared_operands(mf.clone(), af) is None and m.shared_operands(mf, af.clone()) is None
    mf.add_(1)                                                                # modified in place since
    assert m.shared_operands(mf, af) is None
    m._last["mf_ver"] = mf._version
    assert m.shared_operands(mf, af) is not None
    af.mul_(2)                                                                # attn_feats edited in place: stale af16
    assert m.shared_operands(mf, af) is None
    m._last["af_ver"] = af._version
    m._generation = 4                                                         # a later forward reused the workspace
    assert m.shared_operands(mf, af) is None
    m._generation = 3
    r = weakref.ref(af)
    del af
    gc.collect()
    assert r() is None                                                        # the decoder did not keep it alive
    assert m.shared_operands(mf, torch.zeros(1)) is None


def test_pixel_decoder_state_dict_contract_and_config():
    """MSDeformAttnPixelDecoder: parameter names / shapes of the reference (msdeformattn.py:182-306), (cfg, input_shape)
    construction like @configurable (:308-327), scope errors, registry."""
    from types import SimpleNamespace as NS
    from openvis_b200 import pixel_decoder as PD
    from openvis_b200.synthetic import pixel_decoder_param_shapes
    shape = {f"res{i + 2}": PD.ShapeSpec(channels=c, stride=4 << i) for i, c in enumerate((256, 512, 1024, 2048))}
    cfg = NS(MODEL=NS(SEM_SEG_HEAD=NS(IN_FEATURES=["res2", "res3", "res4", "res5"], CONVS_DIM=256, MASK_DIM=256, NORM="GN",
                                      TRANSFORMER_ENC_LAYERS=6, DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES=["res3", "res4", "res5"],
                                      COMMON_STRIDE=4, PIXEL_DECODER_NAME="MSDeformAttnPixelDecoder"),
                      MASK_FORMER=NS(DROPOUT=0.0, NHEADS=8)))
    m = PD.build_pixel_decoder(cfg, shape)
    want = pixel_decoder_param_shapes()
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    assert m.in_features == ["res2", "res3", "res4", "res5"] and m.transformer_in_features == ["res3", "res4", "res5"]
    assert m.num_fpn_levels == 1 and m.maskformer_num_feature_levels == 3 and not m.training
    reg = {}
    PD.register_into(reg)
    assert reg["MSDeformAttnPixelDecoder"] is PD.MSDeformAttnPixelDecoder
    import pytest
    with pytest.raises(NotImplementedError):
        PD.MSDeformAttnPixelDecoder(shape, conv_dim=128)
    with pytest.raises(NotImplementedError):
        PD.MSDeformAttnPixelDecoder({**shape, "res2": PD.ShapeSpec(channels=96, stride=4)})
    import torch
    with pytest.raises(Exception):                      # CPU tensors: no CPU path
        m.forward_features({k: torch.zeros(1, v.channels, 64 // v.stride, 64 // v.stride) for k, v in shape.items()})


def test_chain_mode_selection():
    """decoder._chain_mode: wide chain up to WIDE_CHAIN_AUTO_ROWS query rows (and <= 256 queries, workspace present), the
    one-CTA-per-group chain only when forced (or as the round-2a rule with the wide default off), launch per op otherwise."""
    from openvis_b200 import decoder as D
    kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, nheads=8, dim_feedforward=2048,
              dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)
    m = D.FrameMultiScaleMaskedTransformerDecoder(num_queries=100, **kw)
    ws = lambda G, R, split=True: {"G": G, "R": R, "split": object() if split else None}
    assert m.use_chain is None and m.wide_chain_default
    assert m._chain_mode(ws(1, 100)) == "wide" and m._chain_mode(ws(4, 400)) == "wide" and m._chain_mode(ws(20, 2000)) == "wide"
    assert m._chain_mode(ws(36, 3600)) is None and m._chain_mode(ws(144, 14400)) is None
    assert m._chain_mode(ws(4, 400, split=False)) is None
    m.use_chain = "wide"
    assert m._chain_mode(ws(144, 14400)) == "wide" and m._chain_mode(ws(200, 20000)) is None
    m.use_chain = True
    assert m._chain_mode(ws(144, 14400)) == "group"
    m.use_chain = False
    assert m._chain_mode(ws(1, 100)) is None and not m._chain_on(ws(1, 100))
    m.use_chain, m.wide_chain_default = None, False
    assert m._chain_mode(ws(1, 100)) == "group" and m._chain_mode(ws(4, 400)) is None
    m200 = D.FrameMultiScaleMaskedTransformerDecoder(num_queries=200, **kw)
    assert m200._chain_mode(ws(2, 400)) == "wide"              # 200-query groups fit the wide chain's CTA-wide self-attention
    m200.use_chain = True
    assert m200._chain_mode(ws(2, 400)) is None                # ... but not the one-warp-per-head group chain


def test_forward_tokens_argument_checks():
    """decoder.forward_tokens: Video decoders only, strides 32 / 16 / 8 / 4 of a /32-padded input -- rejected before any
    device work (no GPU needed for the error paths)."""
    import pytest
    import torch
    from openvis_b200 import decoder as D
    from openvis_b200.pixel_decoder import DecoderTokens
    kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=100, nheads=8,
              dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)
    H4, W4 = 32, 48
    ok_sizes = [(H4 // s, W4 // s) for s in (8, 4, 2)]
    mk = lambda sizes, h4=H4, w4=W4: DecoderTokens([torch.zeros(2, h * w, 256, dtype=torch.float16) for h, w in sizes],
                                                   torch.zeros(2, h4 * w4, 256, dtype=torch.float16), sizes, h4, w4)
    v = D.VideoMultiScaleMaskedTransformerDecoder(**kw).eval()
    with pytest.raises(NotImplementedError):
        v.forward_tokens(mk([(4, 6), (8, 12), (15, 24)]))                      # not the /2 pyramid of the mask features
    with pytest.raises(NotImplementedError):
        v.forward_tokens(mk([(3, 5), (7, 11), (15, 23)], 30, 46))              # H/4, W/4 not multiples of 8
    with pytest.raises(NotImplementedError):
        D.FrameMultiScaleMaskedTransformerDecoder(**kw).eval().forward_tokens(mk(ok_sizes))
    with pytest.raises(NotImplementedError):
        D.SideAdapterVideoMultiScaleMaskedTransformerDecoder(clip_heads=12, **kw).eval().forward_tokens(mk(ok_sizes))
    with pytest.raises(RuntimeError):
        v.train().forward_tokens(mk(ok_sizes))


def test_bench_workloads_cover_every_baseline_config():
    """bench.py: one workload per BASELINE.json config (metric / unit as BASELINE names them, the default = the 64-clip scaling
    sweep on the LV-VIS vocabulary, shapes of the named configs)."""
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    assert len(base["configs"]) == 5 and base["metric"].startswith(bench.METRIC) and bench.UNIT == "frames/s"
    named = " ".join(bench.BASELINE_CONFIG.values())
    assert all(f"configs[{i}]" in named for i in range(5))
    W = bench.WORKLOADS
    assert set(W) == set(bench.BASELINE_CONFIG) and bench.DEFAULT_WORKLOAD in W and set(bench.OTHER_CONFIGS) <= set(W)
    pipe, kind, T, Hp, Wp, out_hw, Q, K, total = W[bench.DEFAULT_WORKLOAD]
    assert (T, out_hw, Q, K, total) == (36, (720, 1280), 100, 1196, 64) and Hp % 32 == 0 and Wp % 32 == 0      # configs[4] on configs[1]'s clip
    assert W["openvis_video_5x360x640_q100_k40"][2:8] == (5, 384, 640, (360, 640), 100, 40)                      # configs[0]
    assert W["openvis_video_36x720x1280_q100_k40"][2:8] == (36, 736, 1280, (720, 1280), 100, 40)                 # configs[1]
    assert W["brivis_frame_36x360x640_q100_k1196"][0] == "brivis" and W["brivis_frame_36x360x640_q100_k1196"][7] == 1196    # configs[2]
    assert W["san_online_36x720x1280_q200_k1196"][6] == 200                                                      # configs[3]
