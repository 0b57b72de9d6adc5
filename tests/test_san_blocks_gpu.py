"""SAN / BriVIS side path: the post-split CLIP blocks on the B200 kernels against the oracle
restatement of SideAdapter.post_encode_image and against the committed reference outputs."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from openvis_b200 import _lib as L  # noqa: E402
from openvis_b200.ov_head import SideAdapterBlocks  # noqa: E402
from openvis_b200.synthetic import seeded_clip_block_params  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L.device_check()


def _module(Q, pseed, ln_w, ln_b, proj):
    sd = {f"transformer.resblocks.{k}": v for k, v in seeded_clip_block_params(pseed).items()}
    sd.update({"ln_post.weight": ln_w, "ln_post.bias": ln_b, "proj": proj})
    return SideAdapterBlocks(num_queries=Q).load_clip_visual_state_dict(sd)


def test_matches_reference_golden(golden_dir):
    from oracle.make_golden import san_blocks_inputs
    g = np.load(os.path.join(golden_dir, "san_blocks.npz"))
    st = np.load(os.path.join(golden_dir, "san_tail.npz"))
    n, Q, pseed = [int(v) for v in g["meta"]]
    m = _module(Q, pseed, torch.tensor(st["ln_w"]), torch.tensor(st["ln_b"]), torch.tensor(st["proj"]))
    cls, pix, bias = san_blocks_inputs(n, Q)
    f = m.post_encode_image((cls.cuda(), pix.cuda()), bias.cuda()).cpu()
    ref = torch.tensor(g["clip_feats"])
    # unit-norm 512-d features (|component| ~ 0.04): fp16 operands through three blocks
    assert (f - ref).abs().max().item() < 3e-3, (f - ref).abs().max().item()
    assert torch.nn.functional.cosine_similarity(f, ref, dim=-1).min().item() > 0.9995


@pytest.mark.parametrize("n,Q", [(3, 100), (1, 200)])
def test_matches_oracle_full_queries(n, Q):
    from oracle import decoder_ref as O
    gen = torch.Generator().manual_seed(17)
    P = seeded_clip_block_params(8)
    ln_w, ln_b = 1 + 0.1 * torch.randn(768, generator=gen), 0.1 * torch.randn(768, generator=gen)
    proj = torch.randn(768, 512, generator=gen) * 768 ** -0.5
    cls = torch.randn(1, n, 768, generator=gen)
    pix = torch.randn(n, 768, 14, 14, generator=gen)
    bias = 3.0 * torch.randn(n, 12, Q, 24, 40, generator=gen)
    sos_ref = O.san_post_blocks(P, cls, pix, bias, Q)
    f_ref, _ = O.san_sos_tail(sos_ref, ln_w, ln_b, proj, torch.zeros(1, 512), 1.0)
    m = _module(Q, 8, ln_w, ln_b, proj)
    sos = m.post_blocks((cls.cuda(), pix.cuda()), [bias.cuda()]).cpu()
    assert ((sos - sos_ref).abs() <= 2e-2 + 1e-2 * sos_ref.abs()).float().mean().item() >= 0.999
    f = m.post_encode_image((cls.cuda(), pix.cuda()), bias.cuda()).cpu()
    assert (f - f_ref).abs().max().item() < 3e-3
    assert torch.nn.functional.cosine_similarity(f, f_ref, dim=-1).min().item() > 0.9995


@pytest.mark.parametrize("n,Q,K", [(3, 100, 1197), (2, 37, 41)])
def test_fused_tail_logits(n, Q, K):
    """Kernel 3 in its fused form (SideAdapterBlocks.post_encode_logits: ln_post on the strided SOS rows -> proj GEMM with
    row sums of squares -> logits GEMM with the normalise + exp(logit_scale) epilogue, three launches) against the composed
    reference calls post_encode_image + cal_sim_logits and against the oracle's fp32 tail."""
    from oracle import decoder_ref as O
    from openvis_b200.ov_head import ClipLogitHead
    gen = torch.Generator().manual_seed(23)
    P = seeded_clip_block_params(8)
    ln_w, ln_b = 1 + 0.1 * torch.randn(768, generator=gen), 0.1 * torch.randn(768, generator=gen)
    proj = torch.randn(768, 512, generator=gen) * 768 ** -0.5
    cls = torch.randn(1, n, 768, generator=gen)
    pix = torch.randn(n, 768, 14, 14, generator=gen)
    bias = 3.0 * torch.randn(n, 12, Q, 24, 40, generator=gen)
    text = torch.nn.functional.normalize(torch.randn(K, 512, generator=gen), dim=-1)
    m = _module(Q, 8, ln_w, ln_b, proj)
    feats = (cls.cuda(), pix.cuda())
    n0 = L.launch_count()
    fused = m.post_encode_logits(feats, bias.cuda(), text.cuda())
    n_fused = L.launch_count() - n0
    n0 = L.launch_count()
    composed = m.cal_sim_logits(text.cuda(), m.post_encode_image(feats, bias.cuda()))
    n_composed = L.launch_count() - n0
    assert fused.shape == composed.shape == (n, Q, K)
    assert n_fused < n_composed                            # tail: 3 launches instead of 5 + the torch copy of the SOS rows
    assert (fused - composed).abs().max().item() < 2e-2    # |logit| <= 14.3
    sos_ref = O.san_post_blocks(P, cls, pix, bias, Q)
    _, ref = O.san_sos_tail(sos_ref, ln_w, ln_b, proj, text, m.tail.logit_scale_exp)
    assert (fused.cpu() - ref).abs().max().item() < 4e-2
    assert (fused.cpu().argmax(-1) == ref.argmax(-1)).float().mean().item() >= 0.99
    # crop path (ClipAdapter.normalize + cal_sim_logits): the normalise-in-epilogue form against the oracle
    f = 3.0 * torch.randn(n * Q, 512, generator=gen)
    lg = ClipLogitHead().cal_sim_logits(text.cuda(), f.cuda(), 100, normalized=False).cpu()
    want = O.ov_cosine_logits(f, text, 100.0)
    assert (lg - want).abs().max().item() < 0.06
