"""Device-side post-processing against the restatement of the reference's
VideoMaskFormer.postprocess + inference_video (oracle.decoder_ref.video_postprocess, plain F.interpolate / topk)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from openvis_b200 import _lib as L  # noqa: E402
from openvis_b200 import postprocess as PP  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L.device_check()


@pytest.mark.parametrize("Q,K,T,pad,img,out", [
    (100, 40, 3, (96, 160), (90, 160), (90, 160)),        # output = image: second interpolation is the identity
    (100, 40, 2, (96, 160), (90, 160), (180, 320)),       # MIN_SIZE_TEST-style: frames resized back up
    (20, 1196, 2, (64, 96), (61, 93), (47, 75)),          # odd sizes, down-scaling, ragged last word
    (7, 3, 1, (32, 32), (32, 32), (32, 33)),              # fewer candidates than 10 * ... (k = 10 <= Q*K = 21)
])
def test_inference_video_matches_oracle(Q, K, T, pad, img, out):
    from oracle import decoder_ref as O
    g = torch.Generator().manual_seed(5)
    scores = torch.rand(Q, K, generator=g).softmax(-1) * torch.rand(Q, 1, generator=g)
    masks = torch.randn(Q, T, pad[0] // 4, pad[1] // 4, generator=g) * 4
    sc, lab, qi, ent, ref_mask, ref_logit = O.video_postprocess(scores, masks, pad, img, out)
    res = PP.inference_video(Q, K, scores.cuda(), masks.cuda(), pad, img, out[0], out[1])
    assert res["image_size"] == tuple(out)
    assert res["pred_queries"] == qi.tolist() and res["pred_labels"] == lab.tolist()
    assert torch.allclose(torch.tensor(res["pred_scores"]), sc, atol=0, rtol=0)
    assert torch.allclose(torch.tensor(res["pred_entropys"]), ent, atol=1e-5, rtol=1e-5)
    got = res["pred_masks"].unpack()
    assert got.shape == ref_mask.shape
    # identical except where the resized logit is within fp32 evaluation-order noise of the threshold
    diff = got != ref_mask
    assert bool((~diff | (ref_logit.abs() < 1e-4)).all()), int(diff.sum())
    assert diff.float().mean().item() < 1e-4


def test_packed_masks_roundtrip():
    g = torch.Generator().manual_seed(0)
    m = torch.rand(2, 3, 5, 70, generator=g) > 0.5
    words = (70 + 31) // 32
    pad = torch.zeros(2, 3, 5, words * 32, dtype=torch.bool)
    pad[..., :70] = m
    bits = (pad.view(2, 3, 5, words, 32).long() << torch.arange(32)).sum(-1)
    bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
    assert torch.equal(PP.PackedMasks(bits, 70).unpack(), m)


@pytest.mark.parametrize("T,pad,img", [(3, (96, 160), (90, 160)), (2, (736, 1280), (720, 1280)), (1, (64, 96), (64, 96)), (2, (32, 64), (29, 37))])
def test_x4_fast_path_against_torch_and_general_kernel(T, pad, img, monkeypatch):
    """output = image size with an exact x4 first interpolation takes mask_postprocess_x4_kernel: against F.interpolate + crop
    (the reference's arithmetic) including the first / last two rows and columns (clamped samples), ragged widths, and bit for
    bit against the general kernel except at |logit| ~ 0."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    Q = 12
    masks = (torch.randn(Q, T, pad[0] // 4, pad[1] // 4, generator=g) * 4).cuda()
    qi = torch.tensor([3, 0, 11, 7, 5, 1, 2, 9, 10, 4], dtype=torch.int32).cuda()
    n0 = L.launch_count()
    bits = L.mask_postprocess(masks, qi, pad, img, img)
    assert L.launch_count() - n0 == 1
    got = PP.PackedMasks(bits, img[1]).unpack().cpu()
    up = F.interpolate(masks[qi.long()], size=pad, mode="bilinear", align_corners=False)[:, :, :img[0], :img[1]].cpu()
    want = up > 0
    diff = got != want
    assert got.shape == want.shape and bool((~diff | (up.abs() < 1e-4)).all()), int(diff.sum())
    # rows / columns 0, 1 and the last ones exercise the clamped samples
    assert torch.equal(got[..., :2, :][up[..., :2, :].abs() > 1e-4], want[..., :2, :][up[..., :2, :].abs() > 1e-4])
