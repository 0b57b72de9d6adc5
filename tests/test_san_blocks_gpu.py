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
