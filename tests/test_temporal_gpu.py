"""Temporal association on the B200 kernels against the oracle restatement and the committed
outputs of the reference's own functions: query matching (index work: bit-exact), re-ordering (bit-exact), the temporal
Conv1d unfold (bit-exact) and TemporalInstanceResampler (fp16 operands / fp32 accumulation: tolerances below)."""
import os

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

pytestmark = pytest.mark.gpu

from openvis_b200 import _lib as L  # noqa: E402
from openvis_b200 import temporal as T  # noqa: E402
from openvis_b200.ov_head import SideAdapterBlocks  # noqa: E402
from openvis_b200.synthetic import seeded_clip_block_params, seeded_resampler_params  # noqa: E402

MATCH_CASES = {"q100": dict(b=2, t=6, Q=100, seed=41), "q200": dict(b=1, t=4, Q=200, seed=42),
               "q7": dict(b=3, t=5, Q=7, seed=43, noise=1.5)}


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L.device_check()


@pytest.mark.parametrize("G,Tn,taps", [(7, 9, 5), (3, 1, 5), (5, 2, 3), (100, 36, 3)])
def test_unfold_equals_replicate_padding(G, Tn, taps):
    x = torch.randn(G, Tn, 256, generator=torch.Generator().manual_seed(G)).half()
    ref = torch.nn.functional.pad(x.float().transpose(1, 2), (taps // 2, taps // 2), mode="replicate")    # [G, C, T + taps - 1]
    ref = torch.stack([ref[:, :, k:k + Tn] for k in range(taps)], dim=1)                                   # [G, taps, C, T]
    ref = ref.permute(0, 3, 1, 2).reshape(G, Tn, taps * 256).half()
    got = L.temporal_unfold_f16(x.cuda(), taps).cpu()
    assert torch.equal(got, ref)


@pytest.mark.parametrize("name", list(MATCH_CASES))
def test_matching_equals_reference_golden(name, golden_dir):
    from oracle import temporal_ref as TR
    from oracle.make_golden import temporal_match_inputs
    g = np.load(os.path.join(golden_dir, "temporal_match.npz"))
    e = temporal_match_inputs(**MATCH_CASES[name])
    idx, emb, pi, cost = T.batch_video_match_via_embeds(e.cuda(), return_cost=True)
    assert idx.dtype == torch.int64 and idx.is_cuda
    # cost matrices: fp32 dot products in a different summation order
    b, t, q, _ = e.shape
    for bi in range(b):
        for i in range(t):
            ref_c = TR.match_cost(e[bi, max(i - 1, 0)], e[bi, i]).T            # rows = target
            assert (cost[bi, i].cpu() - ref_c).abs().max().item() < 2e-6
            # the solver is exact on its own cost matrix: same assignment as scipy's
            assert np.array_equal(pi[bi, i].cpu().numpy(), linear_sum_assignment(cost[bi, i].cpu().numpy())[1])
    # and the chain equals the reference's (index work: bit-exact)
    assert np.array_equal(idx.cpu().numpy(), g[name + "_indices"].astype(np.int64))
    oi, oe = TR.batch_video_match_via_embeds(e)
    assert torch.equal(idx.cpu(), oi) and torch.equal(emb.cpu(), oe)
    assert T.match_via_embeds(e[0, 0].cuda(), e[0, 1].cuda()) == g[name + "_pair"].tolist()


def test_matching_large_n_global_cost_path():
    """n = 230: the cost matrix does not fit shared memory and lives in the scratch buffer."""
    from oracle import temporal_ref as TR
    from oracle.make_golden import temporal_match_inputs
    e = temporal_match_inputs(b=1, t=3, Q=230, seed=77)
    idx, emb = T.batch_video_match_via_embeds(e.cuda())
    oi, oe = TR.batch_video_match_via_embeds(e)
    assert torch.equal(idx.cpu(), oi) and torch.equal(emb.cpu(), oe)


def test_matching_unstructured_embeddings():
    """i.i.d. embeddings (no instance structure: every assignment is a long augmenting-path search)."""
    from oracle import temporal_ref as TR
    e = torch.randn(2, 4, 100, 256, generator=torch.Generator().manual_seed(5))
    idx, _ = T.batch_video_match_via_embeds(e.cuda())
    assert torch.equal(idx.cpu(), TR.batch_video_match_via_embeds(e)[0])


def test_matching_full_size_properties():
    """BASELINE config 5 scale: 64 clips x 36 frames x 100 queries in one call.  Size-independent properties: every
    row is a permutation; frame 0 maps to itself; the chain is the composition of the raw assignments; a sample of the
    2304 problems equals scipy on the kernel's cost matrix; matched cost <= cost of the identity and of a random
    permutation."""
    from oracle.make_golden import temporal_match_inputs
    e = temporal_match_inputs(b=64, t=36, Q=100, seed=123).cuda()
    idx, emb, pi, cost = T.batch_video_match_via_embeds(e, return_cost=True)
    torch.cuda.synchronize()
    ar = torch.arange(100, device="cuda")
    assert torch.equal(idx.sort(-1).values, ar.expand_as(idx))
    assert torch.equal(pi.long().sort(-1).values, ar.expand_as(idx))
    assert torch.equal(idx[:, 0], ar.expand(64, 100))
    comp = ar.expand(64, 100)
    for i in range(36):
        comp = torch.gather(pi[:, i].long(), 1, comp)
        assert torch.equal(comp, idx[:, i])
    assert torch.equal(emb, torch.gather(e, 2, idx[..., None].expand_as(e)))
    matched = torch.gather(cost, 3, pi.long()[..., None]).sum((-1, -2))
    ident = torch.diagonal(cost, dim1=-2, dim2=-1).sum(-1)
    rnd = torch.gather(cost, 3, torch.randperm(100, device="cuda").expand(64, 36, 100)[..., None]).sum((-1, -2))
    assert (matched <= ident + 1e-4).all() and (matched <= rnd + 1e-4).all()
    g = torch.Generator().manual_seed(0)
    for _ in range(12):
        bi, i = int(torch.randint(64, (1,), generator=g)), int(torch.randint(36, (1,), generator=g))
        assert np.array_equal(pi[bi, i].cpu().numpy(), linear_sum_assignment(cost[bi, i].cpu().numpy())[1])


def test_reorder_equals_reference_golden(golden_dir):
    from oracle.make_golden import temporal_match_inputs
    g = np.load(os.path.join(golden_dir, "temporal_match.npz"))
    e = temporal_match_inputs(b=2, t=3, Q=7, seed=44)
    idx, _ = T.batch_video_match_via_embeds(e.cuda())
    gen = torch.Generator().manual_seed(45)
    logits, masks = torch.randn(2, 3, 7, 5, generator=gen), torch.randn(2, 7, 3, 4, 6, generator=gen)
    out = T.reset_image_output_order({"pred_logits": logits.cuda(), "pred_masks": masks.cuda()}, idx)
    assert np.array_equal(out["pred_logits"].cpu().numpy(), g["reorder_logits"])
    assert np.array_equal(out["pred_masks"].cpu().numpy(), g["reorder_masks"])


def _adapter(Q, bseed, st):
    sd = {f"transformer.resblocks.{k}": v for k, v in seeded_clip_block_params(bseed).items()}
    sd.update({"ln_post.weight": torch.tensor(st["ln_w"]), "ln_post.bias": torch.tensor(st["ln_b"]), "proj": torch.tensor(st["proj"])})
    a = SideAdapterBlocks(num_queries=Q).load_clip_visual_state_dict(sd)
    a.tail.logit_scale_exp = float(st["logit_scale_exp"])
    return a


def _check_resampler(out, ref_logits, ref_masks, ref_embeds):
    # tolerances: LayerNorm-ed embeddings (|x| ~ 1) through six fp16-operand layers; mask logits relative to their range;
    # class logits = 14.3 x cosine of unit features that went through three fp16-operand CLIP blocks
    assert (out["pred_embeds"].cpu() - ref_embeds).abs().max().item() < 3e-2
    pm = out["pred_masks"].cpu()
    tol = 1e-2 * ref_masks.abs().max().item()
    assert ((pm - ref_masks).abs() <= tol).float().mean().item() >= 0.999, (pm - ref_masks).abs().max().item()
    assert ((pm > 0) == (ref_masks > 0)).float().mean().item() >= 0.995
    assert (out["pred_logits"].cpu() - ref_logits).abs().max().item() < 0.1


def test_resampler_matches_reference_golden(golden_dir):
    from oracle.make_golden import resampler_inputs
    g = np.load(os.path.join(golden_dir, "temporal_resampler.npz"))
    st = np.load(os.path.join(golden_dir, "san_tail.npz"))
    t, Q, pseed, bseed = [int(v) for v in g["meta"]]
    fe, mf, af, bk, text = resampler_inputs(t, Q)
    m = T.TemporalInstanceResampler().eval()
    m.load_state_dict(seeded_resampler_params(pseed))
    m = m.cuda()
    n0 = L.launch_count()
    out = m(fe.cuda(), mf.cuda(), af.cuda(), _adapter(Q, bseed, st), (bk[0].cuda(), bk[1].cuda()), text.cuda())
    assert L.launch_count() - n0 > 60
    assert out["pred_logits"].shape == (1, t, Q, text.shape[0]) and out["pred_masks"].shape == (1, Q, t, 32, 48)
    _check_resampler(out, torch.tensor(g["pred_logits"]), torch.tensor(g["pred_masks"].astype(np.float32)),
                     torch.tensor(g["pred_embeds"]))
    assert len(out["aux_outputs"]) == 6
    for i in (0, 3):                                                     # lazily computed intermediate heads
        a = out["aux_outputs"][i]
        assert (a["pred_logits"].cpu() - torch.tensor(g[f"aux{i}_pred_logits"])).abs().max().item() < 0.1
        am = torch.tensor(g[f"aux{i}_pred_masks"].astype(np.float32))
        assert ((a["pred_masks"].cpu()[..., ::4, ::4] - am).abs() <= 1e-2 * am.abs().max()).float().mean().item() >= 0.999


def test_resampler_matches_oracle_clip_scale(golden_dir):
    """36 frames, 100 queries (BASELINE config 3's clip shape at a reduced mask resolution) against the oracle."""
    from oracle import decoder_ref as O
    from oracle import temporal_ref as TR
    from oracle.make_golden import resampler_inputs
    st = np.load(os.path.join(golden_dir, "san_tail.npz"))
    t, Q = 36, 100
    fe, mf, af, bk, text = resampler_inputs(t, Q, K=41, seed=52, hw=(48, 80))
    P = seeded_resampler_params(22)
    CP = seeded_clip_block_params(7)
    ln_w, ln_b, proj = (torch.tensor(st[k]) for k in ("ln_w", "ln_b", "proj"))
    scale = float(st["logit_scale_exp"])
    post = lambda biases: O.san_sos_tail(O.san_post_blocks(CP, bk[0], bk[1], biases, Q), ln_w, ln_b, proj, text, scale)[0]
    with torch.no_grad():
        ref = TR.resampler_forward(P, fe, mf, af, post, lambda f: scale * f @ text.T, heads_at=())
    m = T.TemporalInstanceResampler().eval()
    m.load_state_dict(P)
    out = m.cuda()(fe.cuda(), mf.cuda(), af.cuda(), _adapter(Q, 7, st), (bk[0].cuda(), bk[1].cuda()), text.cuda())
    _check_resampler(out, ref["pred_logits"], ref["pred_masks"], ref["pred_embeds"])
    assert (out["pred_logits"].argmax(-1).cpu() == ref["pred_logits"].argmax(-1)).float().mean().item() >= 0.99


def test_resampler_long_clip_layers():
    """300 frames (self-attention over the frame axis in two row passes): embeddings against the oracle's layers."""
    from oracle import temporal_ref as TR
    t, Q = 300, 20
    g = torch.Generator().manual_seed(3)
    fe = torch.randn(1, t, Q, 256, generator=g)
    mf, af = torch.randn(t, 256, 8, 8, generator=g), torch.randn(t, 12, 256, 2, 4, generator=g)
    P = seeded_resampler_params(24)

    class _Stub:                       # heads without a CLIP side path: this test is about the temporal layers
        def post_encode_image(self, bk, biases): return biases.mean(1).flatten(2)

        def cal_sim_logits(self, text, f): return f

    with torch.no_grad():
        ref = TR.resampler_forward(P, fe, mf, af, lambda b: b.mean(1).flatten(2), lambda f: f, heads_at=())
    m = T.TemporalInstanceResampler().eval()
    m.load_state_dict(P)
    out = m.cuda()(fe.cuda(), mf.cuda(), af.cuda(), _Stub(), None, None)
    err = (out["pred_embeds"].cpu() - ref["pred_embeds"]).abs()
    assert (err <= 3e-2).float().mean().item() >= 0.999 and err.max().item() < 0.15, err.max().item()
    rm = ref["pred_masks"]
    assert ((out["pred_masks"].cpu() - rm).abs() <= 1e-2 * rm.abs().max()).float().mean().item() >= 0.999


def test_online_post_processing_equals_oracle():
    """MinVIS.post_processing (minvis.py:320-338): matching + re-ordering of logits and masks in one call."""
    from oracle import temporal_ref as TR
    from oracle.make_golden import temporal_match_inputs
    e = temporal_match_inputs(b=2, t=5, Q=100, seed=61)
    g = torch.Generator().manual_seed(62)
    logits, masks = torch.randn(2, 5, 100, 41, generator=g), torch.randn(2, 100, 5, 24, 40, generator=g)
    out = T.post_processing({"pred_logits": logits.cuda(), "pred_masks": masks.cuda(), "pred_embeds": e.cuda()})
    idx, _ = TR.batch_video_match_via_embeds(e)
    rl, rm = TR.reset_image_output_order(logits, masks, idx)
    assert torch.equal(out["pred_logits"].cpu(), rl) and torch.equal(out["pred_masks"].cpu(), rm)
