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
]


def _ref_decoder(kind, Q):
    d = R.decoders()
    cls = {"frame": d.FrameMultiScaleMaskedTransformerDecoder,
           "video": d.VideoMultiScaleMaskedTransformerDecoder,
           "san_frame": d.SideAdapterFrameMultiScaleMaskedTransformerDecoder,
           "san_video": d.SideAdapterVideoMultiScaleMaskedTransformerDecoder}[kind]
    kw = R.decoder_kwargs(num_queries=Q)
    if kind.startswith("san"):
        kw["clip_heads"] = 12
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


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    for case in DECODER_CASES:
        make_decoder_fixture(*case)
    make_san_tail_fixture()
    make_san_blocks_fixture()
    make_msda_fixture()


if __name__ == "__main__":
    main()
