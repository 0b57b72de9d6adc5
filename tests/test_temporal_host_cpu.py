"""Host-side schedule of openvis_b200.temporal on the CPU test box (no GPU here).

The product has NO CPU path: every tensor operation of `temporal.py` is a C-ABI kernel call (`openvis_b200._lib`) and CPU
tensors are refused (tests/test_host_logic.py).  To exercise the *host logic* -- row layouts ([(b q) t] vs [(b t) q]),
weight re-packing of the Conv1d taps, per-clip offsets of the mask GEMM, permutation composition, lazy aux heads -- this
test replaces those C-ABI wrappers, for the duration of one test, with plain-torch stand-ins that mimic their contracts
(fp16 rounding of the operands included) and checks the result against the committed reference outputs.  The stand-ins
exist only in this file.
"""
import contextlib
import os

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from openvis_b200 import _lib as L
from openvis_b200 import temporal as T
from openvis_b200.synthetic import seeded_clip_block_params, seeded_resampler_params
from oracle import decoder_ref as O
from oracle.make_golden import resampler_inputs, temporal_match_inputs

torch.set_grad_enabled(False)


def _ln(x, g, b):
    return O.layer_norm(x, g, b)


def _rownorm(x, g=None, b=None, layer_norm=False, l2=False, want32=True, want16=True, out16=None):
    y = _ln(x, g, b) if layer_norm else x
    if l2:
        y = y / y.norm(dim=-1, keepdim=True)
    if out16 is not None:
        out16.copy_(y.half())
    return (y if want32 else None), (y.half() if want16 else None)


def _cast_f16(x, out=None):
    return x.half() if out is None else out.copy_(x.half())


def _linear_f16(x, w, bias=None, scale=1.0, relu=False, out=None, out_f32=False):
    y = (x.float() @ w.float().T + (bias if bias is not None else 0)) * scale
    y = y.relu() if relu else y
    y = y if out_f32 else y.half()
    return y if out is None else out.copy_(y)


def _linear_ln_f16(x, w, bias, resid, ln1, ln2=None, pe=None, y32=None, y16=None, ype16=None, d32=None, d16=None, split_ws=None):
    y = _ln(x.float() @ w.float().T + bias + resid, *ln1)
    if ln2 is not None:
        d = _ln(y, *ln2)
        if d32 is not None:
            d32.copy_(d)
        if d16 is not None:
            d16.copy_(d.half())
    if y32 is not None:
        y32.copy_(y)
    if y16 is not None:
        y16.copy_(y.half())


def _self_attn(qk, v, out, G, Q):
    q = qk[:, :256].float().view(G, Q, 8, 32).transpose(1, 2) * 32 ** -0.5
    k = qk[:, 256:].float().view(G, Q, 8, 32).transpose(1, 2)
    vv = v.float().view(G, Q, 8, 32).transpose(1, 2)
    out.copy_(((q @ k.transpose(-1, -2)).softmax(-1) @ vv).transpose(1, 2).reshape(G * Q, 256).half())


def _unfold(x, taps, out=None):
    G, Tn, C = x.shape
    idx = (torch.arange(Tn)[:, None] + torch.arange(taps)[None] - taps // 2).clamp(0, Tn - 1)
    y = x[:, idx].reshape(G, Tn, taps * C)
    return y if out is None else out.copy_(y)


def _tokens(x, **_):
    return x.flatten(2).transpose(1, 2).contiguous().half()


def _mask_logits(ft, groups, rpg, me, mgs, Q, out, tgs, ldt, bias=None, posflags=None, rows_per_frame=0):
    flat = out.view(-1)
    for g in range(groups):
        lg = ft.reshape(groups, rpg, 256)[g].float() @ me.reshape(-1, 256)[g * mgs: g * mgs + Q].float().T
        for qi in range(Q):
            flat[g * tgs + qi * ldt: g * tgs + qi * ldt + rpg] = lg[:, qi]


def _san_bias_logits(af, B, P, heads, ae, Q, out):
    out.copy_(torch.einsum("bpnc,bqc->bnqp", af.float().view(B, P, heads, 256), ae.float().view(B, Q, 256)).view(out.shape))


def _match_embeds(en, want_cost=False):
    B, Tn, n, _ = en.shape
    pi = torch.empty(B, Tn, n, dtype=torch.int32)
    for b in range(B):
        for i in range(Tn):
            c = 1 - en[b, max(i - 1, 0)] @ en[b, i].T
            pi[b, i] = torch.as_tensor(linear_sum_assignment(c.numpy())[1], dtype=torch.int32)
    return pi, None


def _match_compose(pi):
    B, Tn, n = pi.shape
    idx = torch.empty(B, Tn, n, dtype=torch.int64)
    cur = torch.arange(n).expand(B, n)
    for i in range(Tn):
        cur = torch.gather(pi[:, i].long(), 1, cur)
        idx[:, i] = cur
    return idx


def _reorder(x, idx, layout="btq"):
    ix = idx.view(*idx.shape, *([1] * (x.dim() - 3)))
    if layout == "btq":
        return torch.gather(x, 2, ix.expand_as(x))
    xt = x.transpose(1, 2)
    return torch.gather(xt, 2, ix.expand_as(xt)).transpose(1, 2).contiguous()


@pytest.fixture
def host_only(monkeypatch):
    for name, fn in dict(rownorm=_rownorm, cast_f16=_cast_f16, linear_f16=_linear_f16, linear_ln_f16=_linear_ln_f16,
                         self_attn=_self_attn, temporal_unfold_f16=_unfold, nchw_to_tokens_hw_f16=_tokens,
                         nchw_to_tokens_f16=_tokens, mask_logits=_mask_logits, san_bias_logits=_san_bias_logits,
                         match_embeds=_match_embeds, match_compose=_match_compose, reorder_queries=_reorder).items():
        monkeypatch.setattr(L, name, fn)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    yield


def test_matching_and_reordering_host_logic(host_only, golden_dir):
    g = np.load(os.path.join(golden_dir, "temporal_match.npz"))
    e = temporal_match_inputs(b=2, t=6, Q=100, seed=41)
    idx, emb = T.batch_video_match_via_embeds(e)
    assert np.array_equal(idx.numpy(), g["q100_indices"].astype(np.int64))
    assert np.allclose(emb.sum(-1).numpy(), g["q100_embeds_sum"], atol=1e-4)
    assert T.match_via_embeds(e[0, 0], e[0, 1]) == g["q100_pair"].tolist()
    e = temporal_match_inputs(b=2, t=3, Q=7, seed=44)
    gen = torch.Generator().manual_seed(45)
    logits, masks = torch.randn(2, 3, 7, 5, generator=gen), torch.randn(2, 7, 3, 4, 6, generator=gen)
    out = T.post_processing({"pred_logits": logits, "pred_masks": masks, "pred_embeds": e})
    assert np.array_equal(out["pred_logits"].numpy(), g["reorder_logits"])
    assert np.array_equal(out["pred_masks"].numpy(), g["reorder_masks"])


def test_resampler_host_schedule_against_reference_outputs(host_only, golden_dir):
    g = np.load(os.path.join(golden_dir, "temporal_resampler.npz"))
    st = np.load(os.path.join(golden_dir, "san_tail.npz"))
    t, Q, pseed, bseed = [int(v) for v in g["meta"]]
    fe, mf, af, bk, text = resampler_inputs(t, Q)
    CP = seeded_clip_block_params(bseed)
    ln_w, ln_b, proj = (torch.tensor(st[k]) for k in ("ln_w", "ln_b", "proj"))
    scale = float(st["logit_scale_exp"])

    class _Adapter:                                   # the CLIP side path is not under test here: oracle arithmetic
        def post_encode_image(self, bk_, biases):
            return O.san_sos_tail(O.san_post_blocks(CP, bk_[0], bk_[1], biases, Q), ln_w, ln_b, proj, text, scale)[0]

        def cal_sim_logits(self, text_, f):
            return scale * f @ text_.T

    m = T.TemporalInstanceResampler().eval()
    m.load_state_dict(seeded_resampler_params(pseed))
    m.use_cuda_graph = False
    out = m(fe, mf, af, _Adapter(), bk, text)
    pm = torch.tensor(g["pred_masks"].astype(np.float32))
    assert (out["pred_embeds"] - torch.tensor(g["pred_embeds"])).abs().max().item() < 2e-2        # fp16 operand rounding
    assert (out["pred_logits"] - torch.tensor(g["pred_logits"])).abs().max().item() < 5e-3
    assert (out["pred_masks"] - pm).abs().max().item() < 5e-3 * pm.abs().max().item()
    a3 = out["aux_outputs"][3]                                                                   # lazily computed head
    assert (a3["pred_logits"] - torch.tensor(g["aux3_pred_logits"])).abs().max().item() < 5e-3
    # two clips in one call: each clip equals its own single-clip call
    feB, mfB, afB, bkB, _ = resampler_inputs(t, Q, seed=99)
    outB = m(feB, mfB, afB, _Adapter(), bkB, text)
    ref_masks, ref_logits = out["pred_masks"].clone(), out["pred_logits"].clone()
    out2 = m(torch.cat([fe, feB]), torch.cat([mf, mfB]), torch.cat([af, afB]), _Adapter(),
             (torch.cat([bk[0], bkB[0]], 1), torch.cat([bk[1], bkB[1]])), text)
    assert out2["pred_masks"].shape[0] == 2 and out2["pred_logits"].shape[:2] == (2, t)
    for c, single in enumerate(((ref_logits, ref_masks), (outB["pred_logits"], outB["pred_masks"]))):
        assert (out2["pred_logits"][c] - single[0][0]).abs().max().item() < 5e-3
        assert (out2["pred_masks"][c] - single[1][0]).abs().max().item() < 5e-3 * pm.abs().max().item()
