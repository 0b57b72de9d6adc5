"""Multi-scale deformable attention kernel against the committed outputs of the reference's
ms_deform_attn_core_pytorch and against the oracle restatement; mirrors the forward checks of the reference's
ops/test.py (check_forward_equal_with_pytorch_float and the channel sweep of its gradient check)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from openvis_b200 import _lib as L  # noqa: E402
from openvis_b200.msda import MSDeformAttnFunction  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L.device_check()


def _run(value, shapes, loc, w):
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    return MSDeformAttnFunction.apply(value.cuda(), shapes.cuda(), start.cuda(), loc.cuda(), w.cuda(), 2).cpu()


@pytest.mark.parametrize("name", ["ref_test", "pixdec", "odd"])
def test_matches_reference_golden(name, golden_dir):
    from oracle.make_golden import MSDA_CASES, msda_inputs
    g = np.load(os.path.join(golden_dir, "msda.npz"))
    value, shapes, loc, w = msda_inputs(**MSDA_CASES[name])
    out = _run(value, shapes, loc, w)
    ref = torch.tensor(g[name])
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-6), (out - ref).abs().max().item()


@pytest.mark.parametrize("D", [2, 4, 8, 16, 30, 32, 64, 71, 128, 1025])
def test_channel_sweep_matches_oracle(D):
    """The reference's test sweeps the channel count (30, 32, 64, 71, 1025, ...): vector path for 4 * 2^k <= 128, scalar
    path otherwise."""
    from oracle import decoder_ref as O
    from oracle.make_golden import msda_inputs
    value, shapes, loc, w = msda_inputs(N=2, M=2, D=D, Lq=37, L=2, P=3, shapes=[(6, 4), (3, 5)], seed=40 + D, scale=1.0, spread=1.3)
    out = _run(value, shapes, loc, w)
    ref = O.ms_deform_attn(value, shapes, loc, w)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-6), (out - ref).abs().max().item()


def test_pixel_decoder_shape():
    """One frame of the pixel decoder's encoder at 384 x 640: 5040 queries, 8 heads x 32, 3 levels, 4 points."""
    from oracle import decoder_ref as O
    from oracle.make_golden import msda_inputs
    value, shapes, loc, w = msda_inputs(N=1, M=8, D=32, Lq=None, L=3, P=4, shapes=[(12, 20), (24, 40), (48, 80)], seed=5, scale=1.0, spread=1.2)
    out = _run(value, shapes, loc, w)
    ref = O.ms_deform_attn(value, shapes, loc, w)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-6), (out - ref).abs().max().item()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_forward_against_reference_golden_and_oracle(golden_dir, ref_dim, fused):
    """MSDeformAttn.forward (ops/modules/ms_deform_attn.py:83-125) -- value projection, one GEMM for the sampling offsets and
    attention logits, then either the fused kernel (softmax + sampling locations + gather of the fp16 value map) or
    ovis_msda_prepare + the fp32 sampling kernel, output projection -- with the reference module's own state dict, against its
    committed output and the oracle restatement."""
    import os
    from oracle import decoder_ref as O
    from oracle.make_golden import msda_module_case
    from openvis_b200.msda import MSDeformAttn
    g = np.load(os.path.join(golden_dir, "msda_module.npz"))
    P, query, ref, src, shapes, start, pad = msda_module_case(ref_dim=ref_dim)
    m = MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4).eval()
    m.load_state_dict(P)                                   # the reference's parameter names
    m = m.cuda()
    m.fused = fused
    n0 = L.launch_count()
    out = m(query.cuda(), ref.cuda(), src.cuda(), shapes.cuda(), start.cuda(), pad.cuda()).cpu()
    assert L.launch_count() - n0 >= (6 if fused else 8)
    want = O.ms_deform_attn_module(P, query, ref, src, shapes, pad)
    # fp16 GEMM operands (three projections in a row), outputs of magnitude ~1: a sampling location that moves by an
    # fp16 ulp of the offsets shifts a bilinear sample, so the check is on the bulk and on the worst case separately
    d = (out - want).abs()
    assert (d <= 1e-2 + 1e-2 * want.abs()).float().mean().item() > 0.999, d.max().item()
    assert d.max().item() < 6e-2, d.max().item()
    assert np.abs(out[:, :, ::2].numpy() - g[f"out{ref_dim}"]).max() < 6e-2
    with pytest.raises(L.OvisError):
        m(query, ref, src, shapes, start, pad)                # CPU tensors: no CPU path
