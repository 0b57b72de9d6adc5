"""Temporal association timing: query matching for 1 and 64 clips of 36 x 100 queries, against
the reference's host loop (scipy per frame) on the same embeddings; TemporalInstanceResampler for one BriVIS clip
(cfg 3: 36 frames of 360x640 -> 96x160 mask features, Q = 100, K = 1197) split into layers / last head."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy.optimize import linear_sum_assignment
from openvis_b200 import _lib as L
from openvis_b200 import temporal as T
from openvis_b200.ov_head import SideAdapterBlocks
from openvis_b200.synthetic import seeded_clip_block_params, seeded_resampler_params


def ev(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def embeds(b, t, Q, noise, seed=1):
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(b, 1, Q, 256, generator=g)
    e = base + noise * torch.randn(b, t, Q, 256, generator=g)
    perm = torch.stack([torch.stack([torch.randperm(Q, generator=g) for _ in range(t)]) for _ in range(b)])
    return torch.gather(e, 2, perm[..., None].expand_as(e))


for b, noise in ((1, 0.6), (64, 0.6), (1, 1e3), (64, 1e3)):
    e = embeds(b, 36, 100, noise).cuda()
    ms = ev(lambda: T.batch_video_match_via_embeds(e))
    en = torch.nn.functional.normalize(e, dim=-1)
    ms_assign = ev(lambda: L.match_embeds(en))
    # the reference's loop on the host (minvis.py:44-72): cost on the device, .cpu(), scipy, per frame
    t0 = time.perf_counter()
    for bi in range(min(b, 4)):
        last = e[bi, 0]
        for i in range(36):
            c = 1 - torch.nn.functional.normalize(e[bi, i], dim=1) @ torch.nn.functional.normalize(last, dim=1).T
            idx = linear_sum_assignment(c.cpu().T.numpy())[1]
            last = e[bi, i][torch.as_tensor(idx, device="cuda")]
    torch.cuda.synchronize()
    host = (time.perf_counter() - t0) * 1e3 / min(b, 4)
    kind = "instances + noise" if noise < 10 else "unstructured"
    print(f"matching {b:3d} clips x 36 x 100 ({kind}): {ms:.3f} ms ({ms / b * 1e3:.1f} us per clip; assign kernel {ms_assign:.3f} ms); "
          f"reference host loop {host:.2f} ms per clip")

g = torch.Generator().manual_seed(0)
t, Q, K = 36, 100, 1197
sd = {f"transformer.resblocks.{k}": v for k, v in seeded_clip_block_params(1).items()}
sd.update({"ln_post.weight": torch.ones(768), "ln_post.bias": torch.zeros(768), "proj": torch.randn(768, 512, generator=g) * 768 ** -0.5})
ad = SideAdapterBlocks(num_queries=Q).load_clip_visual_state_dict(sd)
m = T.TemporalInstanceResampler().eval()
m.load_state_dict(seeded_resampler_params(0))
m = m.cuda()
fe = torch.randn(1, t, Q, 256, generator=g).cuda()
mf = torch.randn(t, 256, 96, 160, generator=g).cuda()
af = (0.2 * torch.randn(t, 12, 256, 24, 40, generator=g)).cuda()
bk = (torch.randn(1, t, 768, generator=g).cuda(), torch.randn(t, 768, 14, 14, generator=g).cuda())
text = torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1).cuda()
n0 = L.launch_count()
m(fe, mf, af, ad, bk, text)
nl = L.launch_count() - n0
ms = ev(lambda: m(fe, mf, af, ad, bk, text))


class _NoClip:      # heads without the CLIP side path: isolates the temporal layers + two einsums
    def post_encode_image(self, bk, biases): return torch.zeros(t, Q, 512, device="cuda")
    def cal_sim_logits(self, text, f): return torch.zeros(t, Q, K, device="cuda")


ms2 = ev(lambda: m(fe, mf, af, _NoClip(), bk, text))
print(f"resampler, 36 frames x 100 queries, 96x160 masks: {ms:.3f} ms per clip ({nl} launches) = {t / ms * 1e3:.0f} frames/s; "
      f"without the CLIP side path {ms2:.3f} ms")
m.materialize_aux = True
ms3 = ev(lambda: m(fe, mf, af, ad, bk, text), n=3, warm=1)
print(f"   API-exact (all seven heads, as the reference computes): {ms3:.3f} ms per clip")
