"""Transposed-score masked cross-attention (xattn_tc3_kernel, csrc/xattn_tc3.cuh) and the key-major mask epilogue
(EPI_SIGNBITS_T) against fp64 torch: dense / sparse / block masks, fully blocked rows (all-masked-row rule,
frame_mask2former_transformer_decoder.py:87), partial last tiles, Q = 200 (two query tiles), tile skipping, and the
retry path (scores that rise far above the first key tile's maximum)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from openvis_b200 import _lib as L  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L.device_check()


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _randn(*shape, seed=0, scale=1.0):
    return (torch.randn(*shape, generator=_g(seed)) * scale).cuda()


def pack_bits_t(blocked, Q):
    """blocked [G, Q, keys] bool -> bits_t [G, keys, qw] int32 (1 past Q), blockand [G, ceil(keys/32), qw]."""
    G, _, keys = blocked.shape
    qw = 4 * ((Q + 127) // 128)
    full = torch.ones(G, qw * 32, keys, dtype=torch.bool, device=blocked.device)
    full[:, :Q] = blocked
    b = full.permute(0, 2, 1).reshape(G, keys, qw, 32).long()
    words = (b << torch.arange(32, device=blocked.device)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    W = (keys + 31) // 32
    padded = torch.full((G, W * 32, qw), -1, dtype=torch.int32, device=blocked.device)
    padded[:, :keys] = words
    blockand = padded.view(G, W, 32, qw)[:, :, 0].clone()
    for i in range(1, 32):
        blockand &= padded.view(G, W, 32, qw)[:, :, i]
    return words.contiguous(), blockand.contiguous()


def ref_xattn(q, k, v, blocked, G, Q, keys):
    qh = q.view(G, Q, 8, 32).permute(0, 2, 1, 3).double()
    kh = k.view(G, keys, 8, 32).permute(0, 2, 1, 3).double()
    vh = v.view(G, keys, 8, 32).permute(0, 2, 1, 3).double()
    s = qh @ kh.transpose(-1, -2) * math.log(2.0)
    full = blocked.all(-1, keepdim=True)
    s = s.masked_fill((blocked & ~full)[:, None], float("-inf"))
    return (s.softmax(-1) @ vh).permute(0, 2, 1, 3).reshape(G * Q, 256)


def run_t(q, k, v, blocked, G, Q, keys, want_stats=False):
    bits_t, blockand = pack_bits_t(blocked, Q)
    flags = (~blocked).any(-1).to(torch.uint8).contiguous()
    use_t, splits, q_pad, o_n, ml_n = L.xattn_plan_t(G, Q, keys)
    o_part = torch.full((o_n,), float("nan"), device="cuda")
    ml_part = torch.full((ml_n,), float("nan"), device="cuda")
    out = torch.empty(G * Q, 256, dtype=torch.float16, device="cuda")
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    L.xattn_t(q, k, v, bits_t, blockand, flags, G, Q, Q, keys, splits, o_part, ml_part, out, stats=stats)
    torch.cuda.synchronize()
    return (out, stats.tolist(), (splits, q_pad, ml_part)) if want_stats else out


def _maxerr(a, b):
    return (a.double() - b.double()).abs().max().item()


@pytest.mark.parametrize("G,Q,keys,density", [(3, 100, 240, 0.5), (2, 100, 3840, 0.5), (1, 100, 19200, 0.9), (2, 200, 920, 0.5),
                                              (1, 100, 77, 0.3), (4, 100, 960, 0.99), (1, 100, 132480, 0.5), (1, 128, 5000, 0.7),
                                              (2, 16, 700, 0.5), (1, 100, 128, 0.5)])
def test_xattn_t(G, Q, keys, density):
    q = _randn(G * Q, 256, seed=1, scale=0.6).half()
    k = _randn(G * keys, 256, seed=2).half()
    v = _randn(G * keys, 256, seed=3).half()
    blocked = torch.rand(G, Q, keys, generator=_g(4)).cuda() < density
    blocked[0, 3] = True          # a fully blocked row -> must attend everywhere
    blocked[-1, Q - 1] = True
    blocked[0, 5] = False
    out, stats, _ = run_t(q, k, v, blocked, G, Q, keys, want_stats=True)
    ref = ref_xattn(q, k, v, blocked, G, Q, keys)
    assert _maxerr(out, ref) < 6e-3, (_maxerr(out, ref), stats)
    assert stats[0] > 0 and (density > 0.9 or stats[1] == 0), stats      # benign scores and dense masks: no CTA retries


def test_xattn_t_retry_path_when_scores_rise():
    """Keys late in a chunk score far above anything in the chunk's first tile (2^40 in probability): the first pass
    saturates, the row sums give it away and the CTA re-runs with the exact reference.  Also a query whose only unblocked
    keys sit in the last tile and score far BELOW the first tile (underflow direction)."""
    G, Q, keys = 1, 100, 6000
    q = _randn(G * Q, 256, seed=1, scale=0.6).half()
    k = _randn(G * keys, 256, seed=2).half()
    v = _randn(G * keys, 256, seed=3).half()
    # boost: keys 4000.. get a large component along each head's query direction of query 10
    qh = q.view(Q, 8, 32).float()
    kk = k.view(keys, 8, 32).float()
    kk[4000:4100] += 12.0 * qh[10][None] / qh[10].norm(dim=-1, keepdim=True)[None]
    kk[5900:] -= 30.0 * qh[20][None] / qh[20].norm(dim=-1, keepdim=True)[None]
    k = kk.reshape(keys, 256).half().contiguous()
    blocked = torch.rand(G, Q, keys, generator=_g(4)).cuda() < 0.5
    blocked[0, 20, :5900] = True                               # query 20 sees only the (very low-scoring) tail
    blocked[0, 20, 5900:] = False
    out, stats, _ = run_t(q, k, v, blocked, G, Q, keys, want_stats=True)
    ref = ref_xattn(q, k, v, blocked, G, Q, keys)
    assert _maxerr(out, ref) < 6e-3, (_maxerr(out, ref), stats)
    assert stats[1] >= 1, stats                               # at least one CTA took the retry


@pytest.mark.parametrize("G,Q,keys", [(2, 100, 3840), (1, 200, 3850), (1, 100, 64000), (4, 100, 14720)])
def test_xattn_t_skips_fully_masked_tiles(G, Q, keys):
    q = _randn(G * Q, 256, seed=1, scale=0.6).half()
    k = _randn(G * keys, 256, seed=2).half()
    v = _randn(G * keys, 256, seed=3).half()
    gen = _g(11)
    blocked = torch.ones(G, Q, keys, dtype=torch.bool)
    for g_ in range(G):
        for qi in range(Q):
            lo, hi = (0.10, 0.35) if qi < 128 else (0.60, 0.80)
            for _ in range(int(torch.randint(1, 4, (1,), generator=gen))):
                a0 = int(keys * lo) + int(torch.randint(0, max(1, int(keys * (hi - lo)) - 50), (1,), generator=gen))
                blocked[g_, qi, a0:a0 + int(torch.randint(1, 50, (1,), generator=gen))] = False
            if qi >= 128 and qi % 7 == 0:
                blocked[g_, qi, keys - 3:] = False
    for second_pass in (False, True):
        if second_pass:
            blocked[0, 7] = True                          # attends everywhere: nothing skippable for (group 0, tile 0)
        bl = blocked.cuda()
        out, stats, (splits, q_pad, ml_part) = run_t(q, k, v, bl, G, Q, keys, want_stats=True)
        ref = ref_xattn(q, k, v, bl, G, Q, keys)
        assert _maxerr(out, ref) < 6e-3, (_maxerr(out, ref), splits, second_pass)
        tiles = (keys + 127) // 128
        qtiles = (Q + 127) // 128
        m = ml_part[G * splits * 8 * q_pad * 2:].view(torch.int32).view(G, qtiles, 256)
        marked = sum(bin(int(wd) & 0xffffffff).count("1") for wd in m[G - 1, 0, :(tiles + 31) // 32].tolist())
        if not (second_pass and G == 1):
            assert marked >= tiles // 2, (marked, tiles)
        if second_pass:
            assert int((m[0, 0, :(tiles + 31) // 32] != 0).sum()) == 0


@pytest.mark.parametrize("G,Q,keys", [(2, 100, 700), (1, 200, 1000), (3, 40, 96)])
def test_mask_bits_t_epilogue(G, Q, keys):
    """EPI_SIGNBITS_T: bits_t / blockand / flags from the mask GEMM equal the packing of (g . mask_embed < 0)."""
    gt = _randn(G * keys, 256, seed=5).half()
    me = _randn(G * Q, 256, seed=6).half()
    qw = 4 * ((Q + 127) // 128)
    W = (keys + 31) // 32
    bits_t = torch.zeros(G, keys, qw, dtype=torch.int32, device="cuda")
    blockand = torch.zeros(G, W, qw, dtype=torch.int32, device="cuda")
    flags = torch.zeros(G, Q, dtype=torch.uint8, device="cuda")
    L.mask_bits_t(gt, G, keys, me, Q, bits_t, blockand, flags, Q)
    logits = torch.einsum("gkc,gqc->gqk", gt.view(G, keys, 256).double(), me.view(G, Q, 256).double())
    blocked = logits < 0
    sure = logits.abs() > 1e-2                                   # fp32-accumulation noise around zero
    got = torch.zeros(G, Q, keys, dtype=torch.bool, device="cuda")
    for w in range(qw):
        for b in range(32):
            qi = w * 32 + b
            if qi < Q:
                got[:, qi] = ((bits_t[:, :, w] >> b) & 1).bool()
            else:
                assert bool((((bits_t[:, :, w] >> b) & 1) == 1).all())          # queries past Q: blocked
    assert bool((got == blocked)[sure].all())
    want_bits, want_and = pack_bits_t(got, Q)
    assert torch.equal(want_bits, bits_t) and torch.equal(want_and, blockand)
    assert torch.equal(flags.bool(), (~got).any(-1))
