"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference modules.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The fixtures hold only reference OUTPUTS; weights and inputs are regenerated from seeds by
oracle.decoder_ref.seeded_params / seeded_inputs (no reference needed), so the fixtures stay small.
"""
import os

import numpy as np
import torch

from . import decoder_ref as O
from . import ref_shim as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, kind, T, Hp, Wp, Q, param seed, input seed)
# Spatial size 128x192 (24 / 96 / 384 keys per frame): large enough that an isolated attention-mask bit flip at fp16
# noise level does not dominate the comparison (at 64x64 the coarsest level has 4 keys), small enough to commit.
DECODER_CASES = [
    ("dec_frame_q100", "frame", 2, 128, 192, 100, 0, 1234),
    ("dec_video_q100", "video", 3, 128, 192, 100, 1, 1235),
    ("dec_san_frame_q100", "san_frame", 2, 128, 192, 100, 2, 1236),
    ("dec_san_video_q100", "san_video", 2, 128, 192, 100, 3, 1237),
    ("dec_frame_q200", "frame", 1, 128, 192, 200, 4, 1238),
    # SimpleBaseline / proposal-network variants (row A17): only class_embed differs
    ("dec_embedding_frame_q100", "embedding_frame", 2, 128, 192, 100, 5, 1241),
    ("dec_proposal_video_q100", "proposal_video", 2, 128, 192, 100, 6, 1240),
]


def _ref_decoder(kind, Q):
    d = R.decoders()
    cls = {"frame": d.FrameMultiScaleMaskedTransformerDecoder,
           "video": d.VideoMultiScaleMaskedTransformerDecoder,
           "san_frame": d.SideAdapterFrameMultiScaleMaskedTransformerDecoder,
           "san_video": d.SideAdapterVideoMultiScaleMaskedTransformerDecoder,
           "embedding_frame": d.frame.EmbeddingFrameMultiScaleMaskedTransformerDecoder,
           "embedding_video": d.video.EmbeddingVideoMultiScaleMaskedTransformerDecoder,
           "proposal_frame": d.frame.ProposalFrameMultiScaleMaskedTransformerDecoder,
           "proposal_video": d.video.ProposalVideoMultiScaleMaskedTransformerDecoder}[kind]
    kw = R.decoder_kwargs(num_queries=Q)
    if kind.startswith("san"):
        kw["clip_heads"] = 12
    if kind.startswith("embedding"):
        kw["clip_dims"] = 512
    return cls(**kw).eval()


def run_reference_decoder(kind, T, Hp, Wp, Q, pseed, iseed):
    m = _ref_decoder(kind, Q)
    P = O.seeded_params(O.decoder_param_shapes(kind, Q=Q), pseed)
    m.load_state_dict(P)
    x, mf = O.seeded_inputs(T, Hp, Wp, seed=iseed)
    with torch.no_grad():
        return m(x, mf)


def make_decoder_fixture(name, kind, T, Hp, Wp, Q, pseed, iseed):
    out = run_reference_decoder(kind, T, Hp, Wp, Q, pseed, iseed)
    rec = {"meta": np.array([T, Hp, Wp, Q, pseed, iseed])}
    for k in ("pred_logits", "pred_masks", "pred_embeds", "class_attn_biases"):
        if k in out:
            rec[k] = out[k].numpy() if k == "pred_logits" else out[k].numpy().astype(np.float16)
    rec["aux0_pred_masks"] = out["aux_outputs"][0]["pred_masks"].numpy().astype(np.float16)
    for i in (4, 8):   # every 4th pixel in both directions
        rec[f"aux{i}_pred_masks"] = out["aux_outputs"][i]["pred_masks"][..., ::4, ::4].numpy().astype(np.float16)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **rec)
    print("wrote", name, {k: v.shape for k, v in rec.items()})


def make_zero_shot_fixture(T=2, Hp=128, Wp=192, Q=100, pseed=8, iseed=1243):
    """tests/golden/dec_zero_shot_q100.npz: the reference's ZeroShotMultiScaleMaskedTransformerDecoder
    (zero_shot_mask2former_transformer_decoder.py:172-277) on seeded weights / inputs."""
    kw = R.decoder_kwargs(num_queries=Q)
    kw.pop("num_frames")
    m = R.zero_shot_decoder()(**kw).eval()
    m.load_state_dict(O.seeded_params(O.decoder_param_shapes("zero_shot", Q=Q), pseed))
    x, mf = O.seeded_inputs(T, Hp, Wp, seed=iseed)
    with torch.no_grad():
        out = m(x, mf)
    rec = dict(meta=np.array([T, Hp, Wp, Q, pseed, iseed]), pred_object_logits=out["pred_object_logits"].numpy(),
               pred_logits=out["pred_logits"].numpy(), pred_embeds=out["pred_embeds"].numpy(),
               pred_masks=out["pred_masks"].numpy().astype(np.float16),
               aux0_pred_masks=out["aux_outputs"][0]["pred_masks"].numpy().astype(np.float16),
               aux4_pred_object_logits=out["aux_outputs"][4]["pred_object_logits"].numpy(),
               aux4_pred_logits=out["aux_outputs"][4]["pred_logits"].numpy())
    np.savez_compressed(os.path.join(GOLDEN, "dec_zero_shot_q100.npz"), **rec)
    print("wrote dec_zero_shot_q100", {k: v.shape for k, v in rec.items()})


def make_san_tail_fixture():
    """SideAdapter._build_attn_biases + post_encode_image tail + cal_sim_logits on seeded inputs
    (side_adapter.py:201-207, 234-270).  The three CLIP blocks in between are out of scope, so the tail
    is exercised on a synthetic SOS-token tensor."""
    s = R.side_adapter_module()
    torch.manual_seed(3)
    sa = s.SideAdapter(num_queries=7).eval()
    g = torch.Generator().manual_seed(77)
    bias = torch.randn(2, 12, 7, 24, 40, generator=g)
    full = sa._build_attn_biases([bias], sa.num_heads, 3, target_shape=(14, 14))[0]
    sos = torch.randn(2, 7, 768, generator=g)
    text = torch.nn.functional.normalize(torch.randn(41, 512, generator=g), dim=-1)
    cm = sa.clip_model
    with torch.no_grad():
        # the same three statements as side_adapter.py:203-205, on the reference's own parameters
        f = torch.nn.functional.normalize(cm.visual.ln_post(sos) @ cm.visual.proj, dim=-1)
        logits = sa.cal_sim_logits(text, f)
    rec = dict(pooled=full[:, :7, -196:].numpy(), corner=full[0, :9, :9].numpy(),
               row_last=full[0, -1].numpy(), clip_feats=f.numpy(), logits=logits.numpy(),
               ln_w=cm.visual.ln_post.weight.detach().numpy(), ln_b=cm.visual.ln_post.bias.detach().numpy(),
               proj=cm.visual.proj.detach().numpy().astype(np.float32),
               logit_scale_exp=np.array(cm.logit_scale.exp().item()))
    np.savez_compressed(os.path.join(GOLDEN, "san_tail.npz"), **rec)
    print("wrote san_tail", {k: v.shape for k, v in rec.items()})


def ov_tail_inputs(T=7, Q=9, K=11, seed=21):
    """Seeded inputs of the OpenVIS / SimpleBaseline / OV2Seg tails: un-normalised region features [T, Q, 512], unit-norm
    text [K, 512], stride-4 mask logits [Q, T, 8, 8] with one never-valid query and one partly valid query."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(T, Q, 512, generator=g)
    text = torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1)
    masks = torch.randn(Q, T, 8, 8, generator=g)
    masks[3] = -5.0
    masks[5, :3] = -5.0
    return feats, text, masks


def run_reference_ov_tails(T=7, Q=9, K=11, seed=21):
    """The reference's own OpenVIS.open_vocabulary_inference (openvis.py:110-147) with `self.clip_adapter` replaced by
    ClipAdapter.forward's arithmetic (adapter.py:56-71) minus the CLIP image tower -- the region features are given --
    i.e. valid flags (adapter.py:85-88), ClipAdapter.normalize and cal_sim_logits, all the reference's own functions;
    plus SimpleBaseline's tail (simplebsl.py:68-69: the same two functions on the decoder's embeddings) and OV2Seg's
    ZeroShotClassifier.forward (ov2seg.py:515-529) with `linear` = identity and the text matrix given."""
    import types
    ov = R.ov_tails()
    feats, text, masks = ov_tail_inputs(T, Q, K, seed)

    def clip_adapter(part_frames, class_names, part_masks):
        bin_masks = part_masks > 0.5                               # adapter.py:85
        valid = bin_masks.sum(dim=(-1, -2)) > 0                     # adapter.py:86
        if torch.sum(valid) == 0:
            return None, valid
        return ov.cal_sim_logits(None, text, ov.normalize(None, part_frames[valid])), valid

    self_ = types.SimpleNamespace(clip_adapter=clip_adapter, device="cpu")
    probs, kept = ov.open_vocabulary_inference(self_, torch.ones(Q), masks, feats, list(range(K)))
    simple = ov.cal_sim_logits(None, text, ov.normalize(None, feats))          # simplebsl.py:69
    zs = types.SimpleNamespace(linear=lambda x: x, norm_weight=True, norm_temperature=50.0, use_bias=False,
                               frame_clip_adapter=types.SimpleNamespace(get_text_features=lambda texts: text))
    ov2 = ov.zero_shot_forward(zs, feats, None)
    return dict(probs=probs, kept_masks=kept, simple=simple, ov2seg=ov2)


def make_ov_tails_fixture():
    out = run_reference_ov_tails()
    rec = {k: v.numpy() for k, v in out.items()}
    np.savez_compressed(os.path.join(GOLDEN, "ov_tails.npz"), **rec)
    print("wrote ov_tails", {k: v.shape for k, v in rec.items()})


def san_blocks_inputs(n=2, Q=12, seed=91):
    """Seeded inputs of the post-split CLIP blocks: CLS token, 14x14 patch features, per-head attention biases."""
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(1, n, 768, generator=g)
    pix = torch.randn(n, 768, 14, 14, generator=g)
    bias = 3.0 * torch.randn(n, 12, Q, 24, 40, generator=g)
    return cls, pix, bias


def make_san_blocks_fixture(n=2, Q=12, pseed=6):
    """SideAdapter.post_encode_image (side_adapter.py:176-209) of the reference, with the three post-split blocks
    loaded from oracle.decoder_ref.seeded_clip_block_params (ln_post / proj stay the seeded CLIP's own = san_tail.npz)."""
    s = R.side_adapter_module()
    torch.manual_seed(3)
    sa = s.SideAdapter(num_queries=Q).eval()
    P = O.seeded_clip_block_params(pseed)
    missing, unexpected = sa.clip_model.visual.transformer.resblocks.load_state_dict(P, strict=False)
    assert not unexpected and all(not k.startswith(("9.", "10.", "11.")) for k in missing)
    cls, pix, bias = san_blocks_inputs(n, Q)
    with torch.no_grad():
        f = sa.post_encode_image((cls, pix), bias)
    rec = dict(meta=np.array([n, Q, pseed]), clip_feats=f.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "san_blocks.npz"), **rec)
    print("wrote san_blocks", {k: v.shape for k, v in rec.items()})


MSDA_CASES = {
    # the reference's own test configuration (ops/test.py:24-31, seed 3)
    "ref_test": dict(N=1, M=2, D=2, Lq=2, L=2, P=2, shapes=[(6, 4), (3, 2)], seed=3, scale=0.01),
    # the pixel decoder's configuration (8 heads x 32, 3 levels, 4 points) at a small size, queries = all positions
    "pixdec": dict(N=2, M=8, D=32, Lq=None, L=3, P=4, shapes=[(4, 6), (8, 12), (16, 24)], seed=11, scale=1.0),
    # odd channel count and sampling locations that leave the map
    "odd": dict(N=1, M=3, D=7, Lq=5, L=2, P=3, shapes=[(5, 3), (2, 7)], seed=12, scale=1.0, spread=1.6),
}


def msda_inputs(N, M, D, Lq, L, P, shapes, seed, scale, spread=1.0):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    Lq = S if Lq is None else Lq
    value = torch.rand(N, S, M, D, generator=g) * scale
    loc = (torch.rand(N, Lq, M, L, P, 2, generator=g) - 0.5) * spread + 0.5
    w = torch.rand(N, Lq, M, L, P, generator=g) + 1e-5
    w = w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)
    return value, torch.as_tensor(shapes, dtype=torch.long), loc, w


def make_msda_fixture():
    """Outputs of the reference's ms_deform_attn_core_pytorch (fp64 evaluation, stored as fp32)."""
    core = R.msda_core_pytorch()
    rec = {}
    for name, cfg in MSDA_CASES.items():
        value, shapes, loc, w = msda_inputs(**cfg)
        with torch.no_grad():
            rec[name] = core(value.double(), shapes, loc.double(), w.double()).float().numpy()
    np.savez_compressed(os.path.join(GOLDEN, "msda.npz"), **rec)
    print("wrote msda", {k: v.shape for k, v in rec.items()})


def temporal_match_inputs(b=2, t=6, Q=100, seed=41, noise=0.6):
    """Frame embeddings of `Q` instances that drift over time and change slot every frame (what the online decoders
    produce): e[b, i] = (base + drift_i)[perm_i] + noise."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(b, Q, 256, generator=g)
    out = []
    for i in range(t):
        perm = torch.stack([torch.randperm(Q, generator=g) for _ in range(b)])
        e = base + noise * torch.randn(b, Q, 256, generator=g)
        out.append(torch.gather(e, 1, perm[:, :, None].expand(-1, -1, 256)))
    return torch.stack(out, dim=1)                                    # [b, t, Q, 256]


def make_temporal_match_fixture():
    """Outputs of the reference's own batch_video_match_via_embeds / match_via_embeds (minvis.py:28-72) and of
    batch_index as used by BriVIS.reset_image_output_order (brivis.py:231-240)."""
    T_ = R.temporal()
    rec = {}
    for name, kw in (("q100", dict(b=2, t=6, Q=100, seed=41)), ("q200", dict(b=1, t=4, Q=200, seed=42)),
                     ("q7", dict(b=3, t=5, Q=7, seed=43, noise=1.5))):
        e = temporal_match_inputs(**kw)
        idx, emb = T_.batch_video_match_via_embeds(e)
        rec[name + "_indices"] = idx.numpy().astype(np.int16)
        rec[name + "_embeds_sum"] = emb.sum(-1).numpy()
        rec[name + "_pair"] = np.array(T_.match_via_embeds(e[0, 0], e[0, 1]), dtype=np.int16)
    e = temporal_match_inputs(b=2, t=3, Q=7, seed=44)
    idx, _ = T_.batch_video_match_via_embeds(e)
    g = torch.Generator().manual_seed(45)
    logits, masks = torch.randn(2, 3, 7, 5, generator=g), torch.randn(2, 7, 3, 4, 6, generator=g)
    fl = T_.batch_index(logits.flatten(0, 1), idx.flatten(0, 1))
    fm = T_.batch_index(masks.transpose(2, 1).flatten(0, 1), idx.flatten(0, 1))
    rec["reorder_logits"] = fl.view(2, 3, 7, 5).numpy()
    rec["reorder_masks"] = fm.view(2, 3, 7, 4, 6).transpose(1, 2).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "temporal_match.npz"), **rec)
    print("wrote temporal_match", {k: v.shape for k, v in rec.items()})


def resampler_inputs(t=4, Q=12, K=9, seed=51, hw=(32, 48)):
    """Seeded inputs of TemporalInstanceResampler.forward for one clip (b = 1)."""
    g = torch.Generator().manual_seed(seed)
    frame_embeds = torch.randn(1, t, Q, 256, generator=g)
    mask_feats = torch.randn(t, 256, hw[0], hw[1], generator=g)
    attn_feats = 0.2 * torch.randn(t, 12, 256, hw[0] // 4, hw[1] // 4, generator=g)
    cls = torch.randn(1, t, 768, generator=g)
    pix = torch.randn(t, 768, 14, 14, generator=g)
    text = torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1)
    return frame_embeds, mask_feats, attn_feats, (cls, pix), text


def make_resampler_fixture(t=4, Q=12, pseed=21, bseed=6):
    """TemporalInstanceResampler.forward (resampler.py:244-302) of the reference with the reference's own SideAdapter as
    `adapter` (post-split blocks = seeded_clip_block_params(bseed), ln_post / proj = the seeded CLIP's, san_tail.npz)."""
    from openvis_b200.synthetic import seeded_resampler_params
    T_ = R.temporal()
    s = R.side_adapter_module()
    torch.manual_seed(3)
    sa = s.SideAdapter(num_queries=Q).eval()
    sa.clip_model.visual.transformer.resblocks.load_state_dict(O.seeded_clip_block_params(bseed), strict=False)
    m = T_.TemporalInstanceResampler().eval()
    m.load_state_dict(seeded_resampler_params(pseed))
    fe, mf, af, bk, text = resampler_inputs(t, Q)
    with torch.no_grad():
        out = m(fe, mf, af, sa, bk, text)
    rec = dict(meta=np.array([t, Q, pseed, bseed]), pred_logits=out["pred_logits"].numpy(),
               pred_masks=out["pred_masks"].numpy().astype(np.float16), pred_embeds=out["pred_embeds"].numpy())
    for i in (0, 3):
        rec[f"aux{i}_pred_logits"] = out["aux_outputs"][i]["pred_logits"].numpy()
        rec[f"aux{i}_pred_masks"] = out["aux_outputs"][i]["pred_masks"][..., ::4, ::4].numpy().astype(np.float16)
    np.savez_compressed(os.path.join(GOLDEN, "temporal_resampler.npz"), **rec)
    print("wrote temporal_resampler", {k: v.shape for k, v in rec.items()})


def msda_module_case(seed=13, ref_dim=2):
    """Seeded MSDeformAttn module parameters + an encoder-style call (queries = all positions of three levels, one
    reference point per level, a padding mask on the last columns of the second sample)."""
    g = torch.Generator().manual_seed(seed)
    shapes = torch.tensor([(8, 12), (4, 6), (2, 3)])
    start = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    N, C, M, L_, P_ = 2, 256, 8, 3, 4
    P = {"sampling_offsets.weight": 0.05 * torch.randn(M * L_ * P_ * 2, C, generator=g),
         "sampling_offsets.bias": 1.5 * torch.randn(M * L_ * P_ * 2, generator=g),
         "attention_weights.weight": 0.1 * torch.randn(M * L_ * P_, C, generator=g),
         "attention_weights.bias": 0.1 * torch.randn(M * L_ * P_, generator=g),
         "value_proj.weight": C ** -0.5 * torch.randn(C, C, generator=g), "value_proj.bias": 0.02 * torch.randn(C, generator=g),
         "output_proj.weight": C ** -0.5 * torch.randn(C, C, generator=g), "output_proj.bias": 0.02 * torch.randn(C, generator=g)}
    query = torch.randn(N, S, C, generator=g)
    src = torch.randn(N, S, C, generator=g)
    ref = torch.rand(N, S, L_, ref_dim, generator=g)
    if ref_dim == 4:
        ref[..., 2:] = 0.1 + 0.3 * ref[..., 2:]
    pad = torch.zeros(N, S, dtype=torch.bool)
    pad[1, -20:] = True
    return P, query, ref, src, shapes, start, pad


def make_msda_module_fixture():
    """tests/golden/msda_module.npz: the reference MSDeformAttn.forward (ops/modules/ms_deform_attn.py:83-125) on
    msda_module_case(), for 2-d reference points and for reference boxes."""
    cls = R.msda_module()
    rec = {}
    for ref_dim in (2, 4):
        P, query, ref, src, shapes, start, pad = msda_module_case(ref_dim=ref_dim)
        m = cls(d_model=256, n_levels=3, n_heads=8, n_points=4).eval()
        m.load_state_dict(P)
        with torch.no_grad():
            rec[f"out{ref_dim}"] = m(query, ref, src, shapes, start, pad)[:, :, ::2].numpy()     # every second channel
    np.savez_compressed(os.path.join(GOLDEN, "msda_module.npz"), **rec)
    print("wrote msda_module", {k: v.shape for k, v in rec.items()})


def clip_adapter_case(seed=3, K=7):
    """Seeded inputs of the crop classifier fixture: (visual params, frames, mask logits, text matrix)."""
    from openvis_b200.synthetic import seeded_clip_visual_params, seeded_crop_inputs
    P = seeded_clip_visual_params(seed)
    frames, logits = seeded_crop_inputs(seed=seed)
    text = torch.nn.functional.normalize(torch.randn(K, 512, generator=torch.Generator().manual_seed(seed + 7)), dim=-1)
    return P, frames, logits, text


def run_reference_clip_adapter(fp32):
    """The reference ClipAdapter's own _preprocess_image / encode_image / cal_sim_logits (adapter.py:73-147) and
    OpenVIS.open_vocabulary_inference (openvis.py:110-147) on clip_adapter_case(); fp32 = its `.half()` sections in fp32."""
    import types
    P, frames, logits, text = clip_adapter_case()
    a = R.clip_adapter(P)
    ov = R.ov_tails()
    with torch.no_grad(), R.on_cpu(fp32=fp32):
        masks = logits.sigmoid().transpose(0, 1).contiguous()
        regions, valid = a._preprocess_image(frames, masks)
        feats = a.encode_image(regions.float())
        sim = a.cal_sim_logits(text, feats)

        def clip_adapter(part_frames, class_names, part_masks):
            r, v = a._preprocess_image(part_frames, part_masks)
            if r is None:
                return None, v
            return a.cal_sim_logits(text, a.encode_image(r.float())), v

        self_ = types.SimpleNamespace(clip_adapter=clip_adapter, device="cpu")
        probs, kept = ov.open_vocabulary_inference(self_, torch.ones(logits.shape[0]), logits, frames, list(range(text.shape[0])))
    return dict(valid=valid, regions=regions.float(), feats=feats, sim=sim, probs=probs, kept_shape=torch.tensor(kept.shape))


def make_clip_adapter_fixture():
    """tests/golden/clip_adapter.npz: the reference's outputs evaluated in fp32 (pins oracle/clip_ref.py) and as written
    (fp16 roi_align on the CPU; the band the fp16 sections move the results by).  Regions are stored sub-sampled."""
    f32 = run_reference_clip_adapter(True)
    f16 = run_reference_clip_adapter(False)
    rec = dict(valid=f32["valid"].numpy(), regions_sub=f32["regions"][:, :, 3::7, 2::7].numpy(),
               regions_mean=f32["regions"].mean(dim=(-1, -2)).numpy(), feats=f32["feats"].numpy(), sim=f32["sim"].numpy(),
               probs=f32["probs"].numpy(), kept_shape=f32["kept_shape"].numpy(),
               regions_sub_h=f16["regions"][:, :, 3::7, 2::7].numpy(), sim_h=f16["sim"].numpy(), probs_h=f16["probs"].numpy())
    np.savez_compressed(os.path.join(GOLDEN, "clip_adapter.npz"), **rec)
    print("wrote clip_adapter", {k: v.shape for k, v in rec.items()},
          "fp16-vs-fp32 band: regions", float(np.abs(rec["regions_sub"] - rec["regions_sub_h"]).max()),
          "sim", float(np.abs(rec["sim"] - rec["sim_h"]).max()), "probs", float(np.abs(rec["probs"] - rec["probs_h"]).max()))


PIXDEC_CH = (64, 128, 192, 256)


def pixel_decoder_case(seed=5, layers=2, T=2, Hp=64, Wp=96):
    """Seeded pixel-decoder weights (narrow stand-in backbone channels, 2 encoder layers) + res2..res5 maps."""
    from openvis_b200.synthetic import seeded_backbone_features, seeded_pixel_decoder_params
    P = seeded_pixel_decoder_params(seed, in_channels=PIXDEC_CH, L=layers)
    feats = seeded_backbone_features(T, Hp, Wp, in_channels=PIXDEC_CH, seed=seed + 70)
    return P, feats


def pixel_decoder_extra(T=2, seed=21):
    """extra_features of the SAN-fused call (mask_former_head.py:120): one map per encoder level (res5, res4, res3 order), the
    middle one at another resolution so that the reference interpolates it."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(T, 256, h, w, generator=g) for (h, w) in ((2, 3), (7, 5), (8, 12))]


def reference_pixel_decoder(P, layers=2):
    ns = R.pixel_decoder()
    shape = {f"res{i + 2}": ns.ShapeSpec(channels=c, stride=4 << i) for i, c in enumerate(PIXDEC_CH)}
    m = ns.MSDeformAttnPixelDecoder(shape, transformer_dropout=0.0, transformer_nheads=8, transformer_dim_feedforward=1024,
                                    transformer_enc_layers=layers, conv_dim=256, mask_dim=256, norm="GN",
                                    transformer_in_features=["res3", "res4", "res5"], common_stride=4).eval()
    m.load_state_dict(P)
    return m


def make_pixel_decoder_fixture():
    """tests/golden/pixel_decoder.npz: the reference MSDeformAttnPixelDecoder.forward_features (msdeformattn.py:329-380) and,
    separately, its MSDeformAttnTransformerEncoderOnly (:76-104) on pixel_decoder_case()."""
    P, feats = pixel_decoder_case()
    m = reference_pixel_decoder(P)
    with torch.no_grad():
        mf, o0, ms = m.forward_features(feats)
        mf_ex, _, ms_ex = m.forward_features(feats, pixel_decoder_extra())
        g = torch.Generator().manual_seed(11)
        srcs = [torch.randn(2, 256, h, w, generator=g) for (h, w) in ((2, 3), (4, 6), (8, 12))]
        pos = [m.pe_layer(s_) for s_ in srcs]
        mem, shapes, start = m.transformer(srcs, pos)
    rec = dict(mask_features=mf.numpy(), ms0=ms[0].numpy(), ms1=ms[1].numpy(), ms2=ms[2].numpy(), enc_memory=mem.numpy(),
               mask_features_ex=mf_ex[:, ::4].numpy(), ms1_ex=ms_ex[1].numpy(),
               enc_shapes=shapes.numpy(), enc_start=start.numpy())
    assert torch.equal(o0, ms[0])
    np.savez_compressed(os.path.join(GOLDEN, "pixel_decoder.npz"), **rec)
    print("wrote pixel_decoder", {k: v.shape for k, v in rec.items()})


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for case in DECODER_CASES:
        make_decoder_fixture(*case)
    make_san_tail_fixture()
    make_san_blocks_fixture()
    make_msda_fixture()
    make_temporal_match_fixture()
    make_resampler_fixture()
    make_ov_tails_fixture()
    make_clip_adapter_fixture()
    make_msda_module_fixture()
    make_pixel_decoder_fixture()
    make_zero_shot_fixture()


if __name__ == "__main__":
    main()
