"""OpenVIS crop classifier on the B200 kernels against the oracle restatement
(oracle/clip_ref.py) and the committed outputs of the reference's own ClipAdapter (tests/golden/clip_adapter.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from openvis_b200 import _lib as L  # noqa: E402
from openvis_b200.clip_adapter import ClipAdapter, ClipVisualEncoder  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L.device_check()


@pytest.fixture(scope="module")
def case():
    from oracle.make_golden import clip_adapter_case
    P, frames, logits, text = clip_adapter_case()
    ad = ClipAdapter(ClipVisualEncoder().load_state_dict(P))
    return P, frames, logits, text, ad


def test_mask_boxes_bit_exact(case):
    from oracle import clip_ref as C
    P, frames, logits, text, ad = case
    masks = logits.sigmoid().transpose(0, 1).contiguous()                   # [T, N, H, W] soft masks
    T, N = masks.shape[:2]
    ref_valid = (masks > 0.5).sum(dim=(-1, -2)) > 0
    ref_boxes = C.mask_boxes((masks > 0.5).flatten(0, 1)).view(T, N, 4).int()
    valid, boxes = L.mask_boxes(masks.cuda(), 0.5)
    assert torch.equal(valid.cpu(), ref_valid) and torch.equal(boxes.cpu(), ref_boxes)
    # the decoder's [N, T, H, W] logits read in place, also through a strided slice of the frames
    valid2, boxes2 = L.mask_boxes(logits.cuda(), 0.5, layout="nt", logits=True)
    assert torch.equal(valid2.cpu(), ref_valid) and torch.equal(boxes2.cpu(), ref_boxes)
    valid3, boxes3 = L.mask_boxes(logits.cuda()[:, 1:], 0.5, layout="nt", logits=True)
    assert torch.equal(valid3.cpu(), ref_valid[1:]) and torch.equal(boxes3.cpu(), ref_boxes[1:])
    # odd width: scalar path
    m = torch.zeros(1, 2, 9, 13)
    m[0, 1, 2:5, 3:11] = 0.9
    v, b = L.mask_boxes(m.cuda(), 0.5)
    assert v.cpu().tolist() == [[False, True]] and b.cpu()[0, 1].tolist() == [3, 2, 11, 5] and b.cpu()[0, 0].tolist() == [0, 0, 0, 0]


def test_regions_against_oracle_and_reference(case, golden_dir):
    from oracle import clip_ref as C
    P, frames, logits, text, ad = case
    g = np.load(os.path.join(golden_dir, "clip_adapter.npz"))
    masks = logits.sigmoid().transpose(0, 1).contiguous()
    ref, ref_valid, _ = C.preprocess_image(frames, masks, half_io=True)
    regions, valid = ad._preprocess_image(frames.cuda(), masks.cuda())
    assert torch.equal(valid.cpu(), ref_valid) and regions.dtype == torch.float16
    assert tuple(regions.shape) == tuple(ref.shape)
    d = (regions.float().cpu() - ref).abs()
    # values up to 255 in fp16 (ulp 0.125 above 128), two roundings on either side; the sigmoid differs by an fp32 ulp
    assert d.max().item() <= 0.5 and (d <= 0.13).float().mean().item() > 0.999, (d.max().item(), (d <= 0.13).float().mean().item())
    # the reference's own output (its fp16 sections evaluated in fp32), sub-sampled in the fixture
    dg = np.abs(regions.float().cpu()[:, :, 3::7, 2::7].numpy() - g["regions_sub"])
    assert dg.max() <= 0.5, dg.max()
    # same crops from the logits in place (no sigmoid / transpose copies)
    regions2, valid2 = ad._preprocess_image(frames.cuda(), logits.cuda(), layout="nt", logits=True)
    d2 = (regions2.float() - regions.float()).abs()
    assert torch.equal(valid2, valid) and d2.max().item() <= 0.26        # __expf sigmoid vs torch's: fp16 rounding flips only


def test_visual_tower_against_oracle(case, golden_dir):
    from oracle import clip_ref as C
    P, frames, logits, text, ad = case
    g = np.load(os.path.join(golden_dir, "clip_adapter.npz"))
    masks = logits.sigmoid().transpose(0, 1).contiguous()
    ref_regions, _, _ = C.preprocess_image(frames, masks, half_io=True)
    with torch.no_grad():
        f_ref = C.encode_image(P, ref_regions)
    f = ad.encode_image(ref_regions.cuda()).cpu()
    # unit-norm 512-d features: fp16 operands through the patch GEMM and twelve blocks
    assert torch.nn.functional.cosine_similarity(f, f_ref, dim=-1).min().item() > 0.9995
    assert (f - f_ref).abs().max().item() < 5e-3, (f - f_ref).abs().max().item()
    assert np.abs(f.numpy() - g["feats"]).max() < 5e-3


def test_forward_and_open_vocabulary_inference(case, golden_dir):
    P, frames, logits, text, ad = case
    g = np.load(os.path.join(golden_dir, "clip_adapter.npz"))
    masks = logits.sigmoid().transpose(0, 1).contiguous()
    n0 = L.launch_count()
    sim, valid = ad(frames.cuda(), text.cuda(), masks.cuda())
    assert L.launch_count() - n0 > 80                                        # the CUDA path ran (12 blocks x 7 launches + front end)
    assert np.array_equal(valid.cpu().numpy(), g["valid"])
    assert np.abs(sim.cpu().numpy() - g["sim"]).max() < 0.25, np.abs(sim.cpu().numpy() - g["sim"]).max()      # logits of scale 100
    probs, kept = ad.open_vocabulary_inference(torch.ones(logits.shape[0]), logits.cuda(), frames.cuda(), text.cuda())
    assert tuple(kept.shape) == tuple(int(v) for v in g["kept_shape"])
    assert np.abs(probs.cpu().numpy() - g["probs"]).max() < 1e-2, np.abs(probs.cpu().numpy() - g["probs"]).max()
    # a part length that splits the clip differently gives the same result (per-region work is independent)
    probs1, _ = ad.open_vocabulary_inference(torch.ones(logits.shape[0]), logits.cuda(), frames.cuda(), text.cuda(), part_len=1)
    assert (probs1 - probs).abs().max().item() < 1e-5


def test_empty_masks_and_errors(case):
    P, frames, logits, text, ad = case
    empty = torch.full_like(logits, -9.0).cuda()
    sim, valid = ad(frames.cuda(), text.cuda(), empty.sigmoid().transpose(0, 1).contiguous())
    assert sim is None and not bool(valid.any())
    assert ad.open_vocabulary_inference(torch.ones(logits.shape[0]), empty, frames.cuda(), text.cuda()) == ([], [])
    assert ad.open_vocabulary_inference([], empty, frames.cuda(), text.cuda()) == ([], [])
    with pytest.raises(L.OvisError):
        L.mask_boxes(logits, 0.5)                                            # CPU tensor: no CPU path
    with pytest.raises(NotImplementedError):
        ad.visual(torch.zeros(1, 3, 192, 192, device="cuda"))
