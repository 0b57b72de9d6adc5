"""Pins oracle/temporal_ref.py (query matching + TemporalInstanceResampler) against the
fixtures generated from the reference's own functions / module (oracle/make_golden.py) and, when /root/reference is
mounted, against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_ref as O
from oracle import ref_shim as R
from oracle import temporal_ref as TR
from oracle.make_golden import resampler_inputs, temporal_match_inputs
from openvis_b200.synthetic import seeded_resampler_params

torch.set_grad_enabled(False)

MATCH_CASES = {"q100": dict(b=2, t=6, Q=100, seed=41), "q200": dict(b=1, t=4, Q=200, seed=42),
               "q7": dict(b=3, t=5, Q=7, seed=43, noise=1.5)}


@pytest.mark.parametrize("name", list(MATCH_CASES))
def test_matching_equals_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "temporal_match.npz"))
    e = temporal_match_inputs(**MATCH_CASES[name])
    idx, emb = TR.batch_video_match_via_embeds(e)
    assert np.array_equal(idx.numpy(), g[name + "_indices"].astype(np.int64))          # index work: bit-exact
    assert np.allclose(emb.sum(-1).numpy(), g[name + "_embeds_sum"], atol=1e-4)
    assert TR.match_via_embeds(e[0, 0], e[0, 1]) == g[name + "_pair"].tolist()
    # the chain is not the identity (the inputs shuffle slots every frame), and frame 0 matches itself
    assert np.array_equal(idx[:, 0].numpy(), np.tile(np.arange(idx.shape[-1]), (idx.shape[0], 1)))
    assert (idx[:, 1:] != torch.arange(idx.shape[-1])).float().mean() > 0.5


def test_reorder_equals_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "temporal_match.npz"))
    e = temporal_match_inputs(b=2, t=3, Q=7, seed=44)
    idx, _ = TR.batch_video_match_via_embeds(e)
    gen = torch.Generator().manual_seed(45)
    logits, masks = torch.randn(2, 3, 7, 5, generator=gen), torch.randn(2, 7, 3, 4, 6, generator=gen)
    fl, fm = TR.reset_image_output_order(logits, masks, idx)
    assert np.array_equal(fl.numpy(), g["reorder_logits"]) and np.array_equal(fm.numpy(), g["reorder_masks"])


def _oracle_adapter(bseed, st, Q, bk, text):
    P = O.seeded_clip_block_params(bseed)
    ln_w, ln_b, proj = (torch.tensor(st[k]) for k in ("ln_w", "ln_b", "proj"))
    scale = float(st["logit_scale_exp"])

    def post(biases):
        sos = O.san_post_blocks(P, bk[0], bk[1], biases, Q)
        return O.san_sos_tail(sos, ln_w, ln_b, proj, text, scale)[0]

    return post, lambda f: scale * f @ text.T


def test_resampler_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "temporal_resampler.npz"))
    st = np.load(os.path.join(golden_dir, "san_tail.npz"))
    t, Q, pseed, bseed = [int(v) for v in g["meta"]]
    fe, mf, af, bk, text = resampler_inputs(t, Q)
    post, sim = _oracle_adapter(bseed, st, Q, bk, text)
    out = TR.resampler_forward(seeded_resampler_params(pseed), fe, mf, af, post, sim)
    assert np.abs(out["pred_embeds"].numpy() - g["pred_embeds"]).max() < 2e-5
    assert np.abs(out["pred_logits"].numpy() - g["pred_logits"]).max() < 2e-4
    pm = g["pred_masks"].astype(np.float32)
    assert np.abs(out["pred_masks"].numpy() - pm).max() < 2e-3 + 1e-3 * np.abs(pm).max()     # fixture stored in fp16
    for i in (0, 3):                                                  # aux_outputs[i] = head i (resampler.py:318-323)
        lg, m = out["heads"][i]
        assert np.abs(lg.numpy() - g[f"aux{i}_pred_logits"]).max() < 2e-4
        am = g[f"aux{i}_pred_masks"].astype(np.float32)
        assert np.abs(m[..., ::4, ::4].numpy() - am).max() < 2e-3 + 1e-3 * np.abs(am).max()


@pytest.mark.skipif(not R.available(), reason="reference not mounted")
@pytest.mark.parametrize("t,b", [(9, 1), (1, 1), (2, 1), (4, 2)])
def test_layers_match_live_reference(t, b):
    """Full-width temporal layers, Q = 20, against the reference module with a stand-in adapter (the heads are pinned by
    the golden test above).  t = 1 and t = 2: clips shorter than the Conv1d kernels (replicate padding only); b = 2: two
    clips per call."""
    T_ = R.temporal()
    m = T_.TemporalInstanceResampler().eval()
    P = seeded_resampler_params(5)
    m.load_state_dict(P)
    gen = torch.Generator().manual_seed(9)
    fe = torch.randn(b, t, 20, 256, generator=gen)
    mf, af = torch.randn(b * t, 256, 8, 8, generator=gen), torch.randn(b * t, 12, 256, 2, 2, generator=gen)

    class _Adapter:
        def post_encode_image(self, bk, biases):
            return biases.mean(1).flatten(2)                     # [n, Q, 4]

        def cal_sim_logits(self, text, f):
            return f @ text.T

    text = torch.randn(3, 4, generator=gen)
    ref = m(fe, mf, af, _Adapter(), None, text)
    ad = _Adapter()
    out = TR.resampler_forward(P, fe, mf, af, lambda b: ad.post_encode_image(None, b), lambda f: ad.cal_sim_logits(text, f))
    for k in ("pred_logits", "pred_masks", "pred_embeds"):
        assert (out[k] - ref[k]).abs().max().item() < 1e-4 * max(1.0, ref[k].abs().max().item()), k


@pytest.mark.skipif(not R.available(), reason="reference not mounted")
def test_matching_live_reference_random_embeds():
    T_ = R.temporal()
    e = torch.randn(2, 5, 33, 256, generator=torch.Generator().manual_seed(3))
    ri, re_ = T_.batch_video_match_via_embeds(e)
    oi, oe = TR.batch_video_match_via_embeds(e)
    assert torch.equal(ri, oi) and torch.equal(re_, oe)
