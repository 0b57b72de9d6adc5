#!/usr/bin/env python
"""Headline benchmark: decoder + open-vocabulary head frames/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repository's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...    # reference arm: the CPU restatement of the reference
                                                                  # path (oracle/) on the host cores, bounded sample

Workload (config.workload): BASELINE configs[1] -- OpenVIS R50 Video decoder on synthetic 36-frame 720x1280 clips
(padded to 736x1280), 100 queries, followed by the OpenVIS OV tail (L2-normalise region features, 100 * f @ text^T
against a cached 40-class text matrix, per-query mean over non-empty frames, softmax).  One step = one clip.
Synthetic data: N(0,1) pixel-decoder outputs, seeded weights (oracle.decoder_ref.seeded_params), unit-norm text.

One process per GPU (torchrun for N > 1); clips shard across ranks with no data-path collective; a single NCCL
all_gather of the per-clip class scores closes the timed region.  value = frames of all ranks / max-over-ranks time.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (decoder kind, T, Hp, Wp, Q, K_vocab)
    "openvis_video_36x720x1280_q100_k40": ("video", 36, 736, 1280, 100, 40),
    "openvis_video_5x360x640_q100_k40": ("video", 5, 384, 640, 100, 40),
}
DEFAULT_WORKLOAD = "openvis_video_36x720x1280_q100_k40"
METRIC = "decoder+OV-head frames/sec"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--api-exact", action="store_true",
                    help="secondary, labelled mode (SURVEY 8d): materialise the nine aux_outputs (full-resolution mask logits "
                         "+ class logits of every intermediate head) as the reference's forward does; default is inference-minimal")
    ap.add_argument("--streams", type=int, default=2, help="independent decoder calls in flight per GPU (one workspace each)")
    ap.add_argument("--clips", type=int, default=4, help="clips stacked into one decoder call (step = this many clips)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {}
        if nv is not None:
            for n in dir(nv):
                if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                    v = getattr(nv, n)
                    if isinstance(v, int) and v:
                        names[v] = n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, n in names.items():
                        if r & bit and n not in ("GpuIdle", "None", "All"):
                            self.reasons.add(n)
                else:
                    import subprocess
                    o = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                        "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    a, b = [int(v) for v in o.strip().split(",")]
                    self.samples.append(a)
                    self.max_mhz = b
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------ helpers
def algorithmic_flops_per_frame(T, Hp, Wp, Q, K, C=256, F=2048, L=9, cls=2):
    """SURVEY.md section 8(d), Video decoder: query-side terms divided by T."""
    N = [Hp * Wp // 1024, Hp * Wp // 256, Hp * Wp // 64]
    M = Hp * Wp // 16
    f_x = sum(4 * N[i % 3] * C * C + 4 * Q * N[i % 3] * C for i in range(L)) + L * 4 * Q * C * C / T
    f_self = L * (8 * Q * C * C + 4 * Q * Q * C) / T
    f_ffn = L * 4 * Q * C * F / T
    f_head = (L + 1) * (6 * Q * C * C + 2 * Q * C * cls) / T + (L + 1) * 2 * Q * C * M
    f_ov = 2 * Q * 512 * K
    return f_x + f_self + f_ffn + f_head + f_ov


def make_text(K, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1)


def make_clip(T, Hp, Wp, Q, seed):
    from openvis_b200.synthetic import seeded_inputs
    x, mf = seeded_inputs(T, Hp, Wp, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    feats = torch.randn(T, Q, 512, generator=g)
    return x, mf, feats


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_step(P, x, mf, feats, text, kind):
    from oracle import decoder_ref as O
    out = O.decoder_forward(P, x, mf, kind=kind, return_attn_masks=False)
    masks = out["pred_masks"][0]                                  # [Q, T, H, W]
    valid = (masks > 0).flatten(2).any(-1).T                      # [T, Q]
    logits = O.ov_cosine_logits(feats.reshape(-1, 512), text, 100.0)
    if valid.any():
        O.openvis_clip_aggregate(logits[valid.flatten()], valid)
    return out


def run_reference(args):
    """CPU implementation of the path (the oracle port; the reference itself is Python and cannot travel to the GPU
    box) on all host threads, each step a bounded sample (a 4-frame sub-clip at the workload's resolution)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import decoder_ref as O
    kind, T, Hp, Wp, Q, K = WORKLOADS[args.workload]
    Ts = min(T, 4)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_grad_enabled(False)
    P = O.seeded_params(O.decoder_param_shapes(kind, Q=Q), 0)
    text = make_text(K)
    x, mf, feats = make_clip(Ts, Hp, Wp, Q, 1234)
    for _ in range(max(1, min(args.warmup, 2))):
        oracle_step(P, x, mf, feats, text, kind)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(P, x, mf, feats, text, kind)
    dt = time.perf_counter() - t0
    v = args.steps * Ts / dt
    sample = f"{Ts}-frame sub-clip of the {T}-frame {Hp}x{Wp} clip per step, fp32, torch CPU ({torch.get_num_threads()} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "frames_per_step": Ts},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(kind, T, Hp, Wp, Q, K):
    from oracle import decoder_ref as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Ts = min(T, 4)
    P = O.seeded_params(O.decoder_param_shapes(kind, Q=Q), 0)
    text = make_text(K)
    x, mf, feats = make_clip(Ts, Hp, Wp, Q, 1234)
    with torch.no_grad():
        oracle_step(P, x, mf, feats, text, kind)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            oracle_step(P, x, mf, feats, text, kind)
            best = min(best, time.perf_counter() - t0)
    return {"value": Ts / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port, {Ts}-frame sub-clip at {Hp}x{Wp}, fp32, best of 3, {torch.get_num_threads()} threads"}


# ------------------------------------------------------------------------------------------------ B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    # persistent kernels leave a few SMs to the other in-flight call's small latency-bound kernels
    if args.streams > 1:
        os.environ.setdefault("OVIS_SM_BUDGET", "140")
    import torch.distributed as dist
    from openvis_b200 import _lib as L
    from openvis_b200 import decoder as D
    from openvis_b200.ov_head import ClipLogitHead
    from openvis_b200 import synthetic as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.device_check()
    torch.set_grad_enabled(False)

    kind, T, Hp, Wp, Q, K = WORKLOADS[args.workload]
    C_ = max(1, args.clips)            # clips per decoder call
    kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=Q, nheads=8,
              dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)
    sd = O.seeded_params(O.decoder_param_shapes(kind, Q=Q), 0)
    decs = []
    for _ in range(max(1, args.streams)):
        d_ = D.VideoMultiScaleMaskedTransformerDecoder(**kw)
        d_.load_state_dict(sd)
        d_.clips_per_call = C_
        # several decoder calls in flight: their eagerly launched small kernels interleave with the other call's heavy
        # ones; replaying each call's layer loop as one CUDA graph removes that interleaving (measured: 2 streams 11.2 k
        # eager vs 10.9 k graph frames/s, 3 streams 11.3 k vs 4.8 k), so graph replay is kept for single-stream use
        d_.use_cuda_graph = d_.use_cuda_graph and args.streams <= 1
        d_.materialize_aux = bool(args.api_exact)
        decs.append(d_.to(dev).eval())
    dec = decs[0]
    streams = [torch.cuda.Stream() for _ in decs]
    head = ClipLogitHead()
    text = make_text(K).to(dev)

    # two distinct clips per rank, alternated, resident in HBM (each clip's inputs are ~2.9 GB >> 126 MB L2)
    # Synthetic inputs are drawn on the device (seeded), which keeps start-up short and host memory small at 8 ranks;
    # the end-to-end leg below copies them once into pinned host memory and uploads from there every step.
    def device_clip(seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        TT_ = T * C_
        x = [torch.randn(TT_, 256, Hp // 32 * 2 ** l, Wp // 32 * 2 ** l, generator=g, device=dev) for l in range(3)]
        mf = torch.randn(TT_, 256, Hp // 4, Wp // 4, generator=g, device=dev)
        feats = torch.randn(TT_, Q, 512, generator=g, device=dev)
        return x, mf, feats

    dev_clips = [device_clip(1234 + 2 * rank + j) for j in range(2)]

    xattn_events = []

    def step(clip, record=False, dec=dec):
        x, mf, feats = clip
        if record:
            L.PROFILE = xattn_events
        out = dec(x, mf)
        L.PROFILE = None
        # OpenVIS OV tail: one logits GEMM for all frames of the call, then the per-clip aggregation
        logits = head.cal_sim_logits(text, feats, 100, normalized=False)            # [C*T, Q, K]
        valid = out["mask_valid"]
        pq = [L.clip_aggregate(logits[c * T:(c + 1) * T], valid[c * T:(c + 1) * T]) for c in range(C_)]
        probs = torch.stack([p for p, _ in pq])
        qvalid = torch.stack([v for _, v in pq])
        return out, probs, qvalid

    def run_steps(n, record=False):
        """n clips; with --streams S > 1 they are issued round-robin on S streams (independent clips in flight)."""
        res = []
        if len(decs) == 1:
            for i in range(n):
                res.append(step(dev_clips[i % 2], record=record)[1])
            return res
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for i in range(n):
            j = i % len(decs)
            with torch.cuda.stream(streams[j]):
                res.append(step(dev_clips[i % 2], record=False, dec=decs[j])[1])
        for st in streams:
            cur.wait_stream(st)
        return res

    # ---- device-resident throughput
    run_steps(max(args.warmup, len(decs)))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = L.launch_count()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    scores = run_steps(args.steps, record=True)
    if world > 1:   # the only collective of the path: gather of per-clip results
        allp = [torch.empty_like(scores[-1]) for _ in range(world)]
        dist.all_gather(allp, scores[-1])
    e1.record()
    torch.cuda.synchronize()
    launches = L.launch_count() - l0
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    value = world * args.steps * C_ * T / (ms * 1e-3)

    # ---- per-kernel-family device time (CUDA events recorded around every C-ABI call inside the timed region;
    #      with --streams > 1 kernels of different clips overlap, so the families are timed in a single-stream pass)
    ms_prof = ms
    if len(decs) > 1:
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(args.steps):
            step(dev_clips[i % 2], record=True)
        p1.record()
        torch.cuda.synchronize()
        ms_prof = p0.elapsed_time(p1)
    N3 = [Hp * Wp // 1024, Hp * Wp // 256, Hp * Wp // 64]
    M = Hp * Wp // 16
    rows3 = [C_ * T * n for n in N3]
    TT = C_ * T
    work = {   # algorithmic work of ONE step (one clip), see DESIGN.md
        "xattn": ("tensor", sum(4.0 * Q * rows3[i % 3] * 256 for i in range(9))),
        "kv_proj": ("tensor", sum(2.0 * rows3[l] * 1536 * 256 for l in range(3))),
        "prep": ("hbm", sum(r * 256 * (4 + 2 + 2) for r in rows3) + TT * M * 256 * (4 + 2) + sum(rows3) * 256 * 2),
        "mask_logits": ("hbm", (TT * M * 256 * 2 + Q * TT * M * 4) * (10 if args.api_exact else 1)),   # API-exact: ten heads
        "mask_bits": ("hbm", sum(rows3[(i) % 3] * 256 * 2 + Q * rows3[i % 3] / 8 for i in range(9))),
    }
    fam_ms = {}
    for (fam, a_, b_) in xattn_events:
        fam_ms[fam] = fam_ms.get(fam, 0.0) + a_.elapsed_time(b_)
    pk, src = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    peak_bw = pk.get("hbm_gbs")
    kernels = {}
    for fam, t_ms in fam_ms.items():
        ent = {"ms_per_step": t_ms / args.steps, "share_of_step": t_ms / ms_prof if ms_prof > 0 else None}
        if fam in work and t_ms > 0:
            bound, amount = work[fam]
            if bound == "tensor":
                ach = amount * args.steps / (t_ms * 1e-3) / 1e12
                ent.update(bound="tensor", achieved=ach, peak=peak_tf, unit="TFLOP/s", frac=ach / peak_tf)
            else:
                ach = amount * args.steps / (t_ms * 1e-3) / 1e9
                ent.update(bound="hbm", achieved=ach, peak=peak_bw, unit="GB/s", frac=ach / peak_bw)
        kernels[fam] = ent
    dom = max((f for f in kernels if f in work), key=lambda f: kernels[f]["ms_per_step"], default=None)
    kname = {"xattn": "xattn_tc2_kernel+xattn_combine_kernel", "kv_proj": "gemm_tn_bs_kernel<256> (key/value projection)",
             "prep": "maskfeat_prep_tma_kernel+tokens_prep_tma_kernel", "mask_logits": "gemm_tn_bs_kernel<128> (final mask logits)",
             "mask_bits": "gemm_tn_bs_kernel<128> (mask sign bits)"}
    if "kv_proj" in kernels and kernels["kv_proj"]["ms_per_step"] > 0:
        # the K/V projection sits at the ridge (190 flop per byte): report the HBM view next to the tensor view
        byts = sum(rows3[l] * (2 * 512 + 1536 * 2) for l in range(3)) * args.steps
        ach = byts / (kernels["kv_proj"]["ms_per_step"] * args.steps * 1e-3) / 1e9
        kernels["kv_proj"]["hbm_view"] = {"achieved": ach, "peak": peak_bw, "unit": "GB/s", "frac": ach / peak_bw}
    if "xattn" in kernels and kernels["xattn"]["ms_per_step"] > 0:
        # d = 32 heads make the masked attention exp-bound, not tensor-bound: 128 flop per exponential.  The binding
        # pipe is the XU (MUFU.EX2: 16 lanes/clk/SM, measured 16.5 by tools/ubench/pipes.cu); reported next to the
        # tensor fraction.  Exponentials counted on the 128-row UMMA tile the kernel has to process (Q = 100 padded).
        sms, clk = torch.cuda.get_device_properties(dev).multi_processor_count, (clocks.get("sm_mhz") or 1965) * 1e6
        qpad = ((Q + 127) // 128) * 128
        nexp = sum(qpad * 8.0 * rows3[i % 3] for i in range(9)) * args.steps
        ach = nexp / (kernels["xattn"]["ms_per_step"] * args.steps * 1e-3) / 1e9
        peak_x = 16.0 * sms * clk / 1e9
        kernels["xattn"]["xu_bound"] = {"pipe": "XU (MUFU.EX2)", "achieved": ach, "peak": peak_x, "unit": "Gexp/s",
                                        "frac": ach / peak_x, "note": "includes the split-combine launches"}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic_r1.json")
    if os.path.isfile(tp):
        try:
            traffic = json.load(open(tp))
        except Exception:
            traffic = None
    roofline = None
    if dom is not None:
        d = kernels[dom]
        roofline = {"kernel": kname[dom], "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                    "frac": d["frac"],
                    "traffic": ((traffic or {}).get(dom, {}).get("dram_bytes_per_clip") or 0) * C_ or None,
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this family's launches (ncu, one clip: "
                                    "profiles/ncu_traffic_r1.json) x clips_per_step, i.e. per step like `achieved`",
                    "peak_source": f"{src} ({'bf16_tflops_sustained' if d['bound'] == 'tensor' else 'hbm_gbs'}; kernel timed inside a long step)",
                    "share_of_step": d["share_of_step"]}
        if dom == "xattn" and "xu_bound" in d:
            # the schema knows "hbm" and "tensor"; this kernel's binding pipe is neither (d = 32: 128 flop per exp)
            roofline["binding_pipe"] = d["xu_bound"]
            roofline["note"] = ("masked attention with 32-wide heads is bound by the exp pipe (XU / MUFU.EX2, 16 lanes/clk/SM), "
                                "which saturates at ~12 % of the tensor peak; `binding_pipe` is the fraction of that pipe's peak, "
                                "ncu: profiles/ncu_r1_kernels.txt (XU 61 %, issue 65 % for the level-2 launch)")

    # ---- end to end through the public API with host buffers (pinned), H2D + forward + D2H every step
    e2e = None
    if not args.no_e2e:
        def to_pinned(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        # ONE clip (T frames, 2.9 GB) in pinned host memory, uploaded once per clip of the step into that clip's slice of
        # the device buffers: the same H2D bytes as a full pinned step, a quarter of the pinned footprint (8 ranks per box)
        x0, mf0, f0 = dev_clips[0]
        one = ([to_pinned(t[:T]) for t in x0], to_pinned(mf0[:T]), to_pinned(f0[:T]))
        torch.cuda.synchronize()
        h2d = C_ * (sum(t.numel() * 4 for t in one[0]) + one[1].numel() * 4 + one[2].numel() * 4)
        copy_s = torch.cuda.Stream()
        bufs = dev_clips                       # reuse the two resident buffers as the double-buffered staging area
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        # post-processing of every clip: top-10 + up-sample / crop / threshold / bit-pack on the device,
        # the packed masks of the 720x1280 frames are part of the per-step device-to-host traffic
        OUT_HW = (720, 1280) if (Hp, Wp) == (736, 1280) else (Hp, Wp)
        res_host = [(torch.empty(C_, Q, K).pin_memory(), torch.empty(C_, Q, 2).pin_memory(),
                     torch.empty(C_, Q, dtype=torch.bool).pin_memory(),
                     torch.empty(C_, 10, T, OUT_HW[0], (OUT_HW[1] + 31) // 32, dtype=torch.int32).pin_memory(),
                     torch.empty(C_, 3, 10).pin_memory()) for _ in range(2)]
        post_dev = [torch.empty(C_, 10, T, OUT_HW[0], (OUT_HW[1] + 31) // 32, dtype=torch.int32, device=dev) for _ in range(2)]
        d2h = sum(t.numel() * t.element_size() for t in res_host[0])

        def upload(j):
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(done[j])
                for c in range(C_):
                    for dst, srcx in zip(bufs[j][0], one[0]):
                        dst[c * T:(c + 1) * T].copy_(srcx, non_blocking=True)
                    bufs[j][1][c * T:(c + 1) * T].copy_(one[1], non_blocking=True)
                    bufs[j][2][c * T:(c + 1) * T].copy_(one[2], non_blocking=True)
                ready[j].record(copy_s)

        def e2e_loop(n):
            cur = torch.cuda.current_stream()
            for j in range(2):
                done[j].record(cur)
            upload(0)
            for i in range(n):
                j = i % 2
                if i + 1 < n:
                    upload((i + 1) % 2)        # overlaps the next clip's H2D with this clip's kernels
                cur.wait_event(ready[j])
                out, probs, qvalid = step(bufs[j])
                res_host[j][0].copy_(probs, non_blocking=True)
                res_host[j][1].copy_(out["pred_logits"], non_blocking=True)
                res_host[j][2].copy_(qvalid, non_blocking=True)
                pm = out["pred_masks"]                                       # [clips, Q, T, H/4, W/4]
                for c in range(C_):
                    vs, qi, lb, en = L.topk_scores(probs[c], 10)
                    L.mask_postprocess(pm[c], qi, (Hp, Wp), OUT_HW, OUT_HW, out=post_dev[j][c])
                    res_host[j][4][c, 0].copy_(vs, non_blocking=True)
                    res_host[j][4][c, 1].copy_(lb, non_blocking=True)
                    res_host[j][4][c, 2].copy_(en, non_blocking=True)
                res_host[j][3].copy_(post_dev[j], non_blocking=True)
                done[j].record(cur)
            torch.cuda.synchronize()

        e2e_loop(2)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * args.steps * C_ * T / dt.item(), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "note": "pinned host inputs -> device (double-buffered on a copy stream) -> decoder + OV head -> device "
                       "post-processing (top-10, x4 up-sample, crop, threshold, bit-pack) -> scores / labels / packed "
                       "720x1280 masks to pinned host; H2D of the fp32 inputs (80 MB per frame) is the PCIe bound"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(kind, T, Hp, Wp, Q, K)
    flops_frame = algorithmic_flops_per_frame(T, Hp, Wp, Q, K)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": args.workload, "frames_per_step_per_gpu": C_ * T, "clips_per_step": C_, "queries": Q, "vocab": K,
                       "l2": "inputs larger than L2 (2.9 GB per clip, two input sets alternated)",
                       "aux_outputs": "materialised: nine intermediate heads written per call (API-exact)" if args.api_exact else "lazy (inference-minimal)", "parallelism": f"clip-sharded dp{world}",
                       "decoder_calls_in_flight_per_gpu": len(decs), "cuda_graph_layer_loop": bool(dec.use_cuda_graph),
                       "sm_budget": os.environ.get("OVIS_SM_BUDGET"),
                       "masked_tile_skipping": "off (OVIS_XATTN_SKIP=0)" if os.environ.get("OVIS_XATTN_SKIP") == "0" else
                       "on; the synthetic N(0,1) features give ~50 % dense masks, so no 128-query x 64-key tile is fully "
                       "masked and nothing is skipped here (sparse-mask A/B: profiles/xattn_skip_ab_r1.txt)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "whole_path": {"gflop_per_frame": flops_frame / 1e9,
                           "tensor_frac_of_sustained": value / world * flops_frame / 1e12 / peak_tf if peak_tf else None}}
    line["ms_per_frame"] = ms / (args.steps * C_ * T)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
