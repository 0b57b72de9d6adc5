#!/usr/bin/env python
"""Headline benchmark: decoder + open-vocabulary head frames/sec (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]     # this repository's CUDA path
  python bench.py --impl reference [--gpus N] [--steps K] ...                # reference arm: the CPU restatement of the
                                                                              # reference path (oracle/) on the host cores

Workloads (config.workload), one per BASELINE.json config:

  scale64_openvis_video_36x720x1280_q100_k1196   DEFAULT.  BASELINE configs[4] (the scaling sweep the metric "frames/sec at
        1/2/4/8 B200" is quoted on): 64 synthetic 36-frame 720x1280 clips of configs[1]'s shape (OpenVIS R50 Video decoder,
        100 queries, padded to 736x1280) with the LV-VIS vocabulary (K = 1196), STRONG-sharded over the ranks in contiguous
        blocks (reference: InferenceSampler, openvis/data/build.py:238-247).  One step = all 64 clips: decoder + OV head +
        device post-processing per clip, closed by the per-clip result gather over NCCL (scores, top-10 ids, bit-packed
        720x1280 masks; reference: comm.gather of the per-video results, openvis/data/evals/ytvis_eval.py:117-128).
  openvis_video_36x720x1280_q100_k40             configs[1]: the same decoder shapes, 40-class YTVIS vocabulary, 4 clips per step
  openvis_video_5x360x640_q100_k40               configs[0]: the reference's CPU-runnable case
  brivis_frame_36x360x640_q100_k1196             configs[2]: BriVIS online (SAN frame decoder -> query matching ->
                                                  TemporalInstanceResampler with the CLIP side path -> post-processing)
  san_online_36x720x1280_q200_k1196              configs[3]: SAN-online, 200 queries, CLIP side path, query matching
  scale64_san_online_36x720x1280_q100_k1196      configs[4], second reading ("run for both"): the sweep on the Frame path

The default N = 1 run also measures configs[0..3] briefly and reports them under "other_configs" (so every BASELINE
config has a number measured by whoever runs this file); --no-other-configs skips that.

Synthetic data: N(0,1) pixel-decoder outputs, seeded weights (openvis_b200.synthetic), unit-norm text.
One process per GPU (torchrun for N > 1).  value = frames of all ranks / max-over-ranks device time.
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

# name: (pipeline, decoder kind, T, Hp, Wp, out_hw, Q, K classes (without background), total clips per step or None)
WORKLOADS = {
    "scale64_openvis_video_36x720x1280_q100_k1196": ("openvis", "video", 36, 736, 1280, (720, 1280), 100, 1196, 64),
    "openvis_video_36x720x1280_q100_k40": ("openvis", "video", 36, 736, 1280, (720, 1280), 100, 40, None),
    "openvis_video_5x360x640_q100_k40": ("openvis", "video", 5, 384, 640, (360, 640), 100, 40, None),
    "brivis_frame_36x360x640_q100_k1196": ("brivis", "san_frame", 36, 384, 640, (360, 640), 100, 1196, None),
    "san_online_36x720x1280_q200_k1196": ("san_online", "san_frame", 36, 736, 1280, (720, 1280), 200, 1196, None),
    "scale64_san_online_36x720x1280_q100_k1196": ("san_online", "san_frame", 36, 736, 1280, (720, 1280), 100, 1196, 64),
}
BASELINE_CONFIG = {
    "scale64_openvis_video_36x720x1280_q100_k1196": "configs[4] (64-clip scaling sweep, LV-VIS vocab) on configs[1]'s clip shape",
    "openvis_video_36x720x1280_q100_k40": "configs[1]",
    "openvis_video_5x360x640_q100_k40": "configs[0]",
    "brivis_frame_36x360x640_q100_k1196": "configs[2]",
    "san_online_36x720x1280_q200_k1196": "configs[3]",
    "scale64_san_online_36x720x1280_q100_k1196": "configs[4] on the SAN-online Frame path",
}
DEFAULT_WORKLOAD = "scale64_openvis_video_36x720x1280_q100_k1196"
OTHER_CONFIGS = ["openvis_video_5x360x640_q100_k40", "openvis_video_36x720x1280_q100_k40", "brivis_frame_36x360x640_q100_k1196",
                 "san_online_36x720x1280_q200_k1196"]
METRIC = "decoder+OV-head frames/sec"
UNIT = "frames/s"
TOPK = 10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--api-exact", action="store_true",
                    help="secondary, labelled mode (SURVEY 8d): materialise the nine aux_outputs (full-resolution mask logits "
                         "+ class logits of every intermediate head) as the reference's forward does; default is inference-minimal")
    ap.add_argument("--streams", type=int, default=None, help="independent pipeline calls in flight per GPU (one workspace each)")
    ap.add_argument("--clips", type=int, default=None, help="clips stacked into one pipeline call")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks / affinity
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
            self._stop.wait(0.05)

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


def bind_to_gpu_numa_node(index):
    """Pins this process (and therefore its pinned host allocations, first-touch) to the CPUs of the GPU's NUMA node:
    with all ranks on node 0 the host->device copies of 8 ranks shared one node's memory bandwidth (round-1 SCALE)."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [i for i in range(n_cpu) if (int(words[i // 64]) >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        bind_to_gpu_numa_node.previous = allowed
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"bound": True, "cpus": len(cpus), "first_cpu": cpus[0], "last_cpu": cpus[-1]}
            try:
                node = None
                for nd in sorted(os.listdir("/sys/devices/system/node")):
                    if nd.startswith("node") and os.path.isdir(f"/sys/devices/system/node/{nd}/cpu{cpus[0]}"):
                        node = int(nd[4:])
                info["numa_node"] = node
            except Exception:
                pass
    except Exception as e:                                       # pragma: no cover - best effort
        info["error"] = str(e)[:120]
    return info


# ------------------------------------------------------------------------------------------------ helpers
def algorithmic_flops_per_frame(kind, T, Hp, Wp, Q, K, C=256, F=2048, L=9, cls=2, nh=12):
    """SURVEY.md section 8(d).  Video decoders: query-side terms divided by T; SAN: attention-bias branch + side-path tail."""
    N = [Hp * Wp // 1024, Hp * Wp // 256, Hp * Wp // 64]
    M = Hp * Wp // 16
    qdiv = T if kind.endswith("video") else 1
    f_x = sum(4 * N[i % 3] * C * C + 4 * Q * N[i % 3] * C for i in range(L)) + L * 4 * Q * C * C / qdiv
    f_self = L * (8 * Q * C * C + 4 * Q * Q * C) / qdiv
    f_ffn = L * 4 * Q * C * F / qdiv
    f_head = (L + 1) * (6 * Q * C * C + 2 * Q * C * cls) / qdiv + (L + 1) * 2 * Q * C * M
    f_ov = 2 * Q * 512 * K
    f_san = 0.0
    if kind.startswith("san"):
        f_san = 2 * (M / 16) * (2 * C * C + nh * C * C) + (L + 1) * (6 * Q * C * C / qdiv + 2 * Q * C * nh * M / 16) + 2 * Q * 768 * 512
    return f_x + f_self + f_ffn + f_head + f_ov + f_san


def make_text(K, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(K, 512, generator=g), dim=-1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


DEC_KW = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, nheads=8, dim_feedforward=2048,
              dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2)


# ------------------------------------------------------------------------------------------------ reference arm
def oracle_clip(name, Ts, seed=1234):
    """One CPU pass of the oracle over a Ts-frame clip of workload `name` (decoder + the OV tail of its pipeline)."""
    from oracle import decoder_ref as O
    pipe, kind, T, Hp, Wp, out_hw, Q, K, _ = WORKLOADS[name]
    P = oracle_clip.cache.get((kind, Q))
    if P is None:
        P = oracle_clip.cache[(kind, Q)] = O.seeded_params(O.decoder_param_shapes(kind, Q=Q), 0)
    x, mf = O.seeded_inputs(Ts, Hp, Wp, seed=seed)
    out = O.decoder_forward(P, x, mf, kind=kind, return_attn_masks=False)
    g = torch.Generator().manual_seed(seed + 1)
    if pipe == "openvis":
        text = make_text(K)
        feats = torch.randn(Ts, Q, 512, generator=g)
        masks = out["pred_masks"][0]
        valid = (masks > 0).flatten(2).any(-1).T
        logits = O.ov_cosine_logits(feats.reshape(-1, 512), text, 100.0)
        if valid.any():
            O.openvis_clip_aggregate(logits[valid.flatten()], valid)
    else:
        # SAN tail arithmetic on the decoder's biases: pooled attention biases + cosine logits against K + 1 rows (the three
        # frozen CLIP blocks between them are the side path; their CPU cost is not part of this baseline)
        text = make_text(K + 1)
        O.san_pool_bias(out["class_attn_biases"][0], (14, 14))
        f = torch.nn.functional.normalize(torch.randn(Ts * Q, 512, generator=g), dim=-1)
        O.ov_cosine_logits(f, text, 1.0 / 0.07, normalize=False)
    return out


oracle_clip.cache = {}


def run_reference(args):
    """CPU implementation of the path (the oracle port; the reference itself is Python + Detectron2 and cannot travel to
    the GPU box) on all host threads.  Each step is a bounded sample -- a sub-clip of the workload's clip, as many frames
    as keep the whole run within a few minutes -- and the full clip is run once next to it (same_config)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    pipe, kind, T, Hp, Wp, out_hw, Q, K, total = WORKLOADS[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_grad_enabled(False)
    t0 = time.perf_counter()
    oracle_clip(name, min(T, 2))                                         # warm-up + calibration
    t0 = time.perf_counter()
    oracle_clip(name, min(T, 2))
    per_frame = (time.perf_counter() - t0) / min(T, 2)
    budget = 150.0
    Ts = T
    for cand in (T, 18, 9, 4, 2, 1):
        if cand <= T and (args.steps + min(args.warmup, 1)) * cand * per_frame <= budget:
            Ts = cand
            break
    else:
        Ts = 1
    for _ in range(min(args.warmup, 1)):
        oracle_clip(name, Ts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_clip(name, Ts)
    dt = time.perf_counter() - t0
    v = args.steps * Ts / dt
    full = None
    if Ts < T and T * per_frame < 120.0:
        t1 = time.perf_counter()
        oracle_clip(name, T)
        d1 = time.perf_counter() - t1
        full = {"frames": T, "seconds": d1, "frames_per_s": T / d1, "same_config": True}
    sample = (f"{Ts}-frame {'clip' if Ts == T else 'sub-clip'} of the {T}-frame {Hp}x{Wp} clip per step, fp32, torch CPU "
              f"({torch.get_num_threads()} threads)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if total else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "baseline_config": BASELINE_CONFIG[name], "frames_per_step": Ts,
                       "same_config": Ts == T, "full_clip_once": full},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(name):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    pipe, kind, T, Hp, Wp, out_hw, Q, K, _ = WORKLOADS[name]
    Ts = min(T, 4)
    with torch.no_grad():
        oracle_clip(name, Ts)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            oracle_clip(name, Ts)
            best = min(best, time.perf_counter() - t0)
    return {"value": Ts / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port, {Ts}-frame sub-clip at {Hp}x{Wp}, fp32, best of 3, {torch.get_num_threads()} threads"}


# ------------------------------------------------------------------------------------------------ B200 pipelines
class Pipeline:
    """One workload on one GPU: `lanes` independent model instances (one CUDA stream each), each call processing
    `clips` clips.  call(lane, inputs, out_slot) runs decoder + OV head + device post-processing for the call's clips and
    leaves the per-clip results (scores [Q, K], top-10 (score, label, entropy), bit-packed masks) in res[out_slot]."""

    def __init__(self, name, dev, clips, lanes, api_exact=False):
        from openvis_b200 import _lib as L
        from openvis_b200 import decoder as D
        from openvis_b200 import synthetic as S
        self.L = L
        self.name, self.dev, self.C, self.lanes = name, dev, clips, lanes
        (self.pipe, self.kind, self.T, self.Hp, self.Wp, self.out_hw, self.Q, self.K, self.total) = WORKLOADS[name]
        T, Q = self.T, self.Q
        kw = dict(DEC_KW, num_queries=Q)
        sd = S.seeded_params(S.decoder_param_shapes(self.kind, Q=Q), 0)
        self.decs, self.extra = [], []
        for _ in range(lanes):
            if self.pipe == "openvis":
                d_ = D.VideoMultiScaleMaskedTransformerDecoder(**kw)
                d_.clips_per_call = clips
            else:
                d_ = D.SideAdapterFrameMultiScaleMaskedTransformerDecoder(clip_heads=12, **kw)
            d_.load_state_dict(sd)
            # several calls in flight: their eagerly launched small kernels interleave with the other call's heavy ones;
            # replaying each call's layer loop as one CUDA graph removes that interleaving (measured in round 1), so graph
            # replay is kept for single-lane use
            d_.use_cuda_graph = d_.use_cuda_graph and lanes <= 1
            d_.materialize_aux = bool(api_exact)
            self.decs.append(d_.to(dev).eval())
        if self.pipe == "openvis":
            from openvis_b200.ov_head import ClipLogitHead
            self.head = ClipLogitHead()
            self.text = make_text(self.K).to(dev)
        else:
            from openvis_b200.ov_head import SideAdapterBlocks
            from openvis_b200 import temporal as TP
            self.TP = TP
            cg = torch.Generator().manual_seed(0)
            csd = {f"transformer.resblocks.{k}": v for k, v in S.seeded_clip_block_params(1).items()}
            csd.update({"ln_post.weight": torch.ones(768), "ln_post.bias": torch.zeros(768),
                        "proj": torch.randn(768, 512, generator=cg) * 768 ** -0.5})
            self.adapter = SideAdapterBlocks(num_queries=Q).load_clip_visual_state_dict(csd)
            self.text = make_text(self.K + 1).to(dev)                     # LV-VIS classes + the learned background row
            if self.pipe == "brivis":
                self.res_mods = []
                for _ in range(lanes):
                    r = TP.TemporalInstanceResampler().eval()
                    r.load_state_dict(S.seeded_resampler_params(0))
                    r.use_cuda_graph = r.use_cuda_graph and lanes <= 1
                    self.res_mods.append(r.to(dev))
        self.streams = [torch.cuda.Stream() for _ in range(lanes)]
        oh, ow = self.out_hw
        self.words = (ow + 31) // 32

    # -- synthetic inputs ------------------------------------------------------------------------------------------
    def device_inputs(self, seed):
        """Inputs of one call (C clips), drawn on the device (seeded): keeps start-up short and host memory small at 8
        ranks; the end-to-end leg copies one clip into pinned host memory and uploads from there every step."""
        g = torch.Generator(device=self.dev).manual_seed(seed)
        n, Hp, Wp, dev = self.T * self.C, self.Hp, self.Wp, self.dev
        x = [torch.randn(n, 256, Hp // 32 * 2 ** l, Wp // 32 * 2 ** l, generator=g, device=dev) for l in range(3)]
        mf = torch.randn(n, 256, Hp // 4, Wp // 4, generator=g, device=dev)
        if self.pipe == "openvis":
            extra = [torch.randn(n, self.Q, 512, generator=g, device=dev)]         # CLIP features of the masked crops
        else:
            extra = [torch.randn(1, n, 768, generator=g, device=dev), torch.randn(n, 768, 14, 14, generator=g, device=dev)]
        return x + [mf] + extra

    def input_bytes_per_clip(self, inputs):
        return sum(t.numel() * t.element_size() for t in inputs) // self.C

    def result_buffers(self, n_clips):
        oh, ow = self.out_hw
        dev = self.dev
        return {"scores": torch.zeros(n_clips, self.Q, self.K, device=dev),
                "top": torch.zeros(n_clips, 3, TOPK, device=dev),
                "masks": torch.zeros(n_clips, TOPK, self.T, oh, self.words, dtype=torch.int32, device=dev)}

    # -- one call --------------------------------------------------------------------------------------------------
    def call(self, lane, inputs, res, slot, post=True):
        """Runs the pipeline on `inputs` (C clips); results of clip c go to res[...][slot + c]."""
        L, T, C, Q = self.L, self.T, self.C, self.Q
        dec = self.decs[lane]
        pad, oh_ow = (self.Hp, self.Wp), self.out_hw
        if self.pipe == "openvis":
            x, mf, feats = inputs[:3], inputs[3], inputs[4]
            out = dec(x, mf)
            # OpenVIS OV tail: one logits GEMM for all frames of the call, then the per-clip aggregation
            logits = self.head.cal_sim_logits(self.text, feats, 100, normalized=False)            # [C*T, Q, K]
            valid = out["mask_valid"]
            pm = out["pred_masks"]                                                                # [C, Q, T, H/4, W/4]
            for c in range(C):
                probs, _ = L.clip_aggregate(logits[c * T:(c + 1) * T], valid[c * T:(c + 1) * T])
                res["scores"][slot + c].copy_(probs)
                if post:
                    vs, qi, lb, en = L.topk_scores(probs, TOPK)
                    L.mask_postprocess(pm[c], qi, pad, oh_ow, oh_ow, out=res["masks"][slot + c])
                    res["top"][slot + c, 0].copy_(vs)
                    res["top"][slot + c, 1].copy_(lb)
                    res["top"][slot + c, 2].copy_(en)
            return out
        x, mf, bk = inputs[:3], inputs[3], (inputs[4], inputs[5])
        if self.pipe == "brivis":
            vids, outputs, _ = self.TP.brivis_video_inference(dec, self.adapter, self.res_mods[lane], x, mf, bk, self.text, pad,
                                                              oh_ow, oh_ow[0], oh_ow[1], num_clips=C, to_host=False)
        else:
            vids, outputs, _ = self.TP.san_online_video_inference(dec, self.adapter, x, mf, bk, self.text, pad, oh_ow,
                                                                  oh_ow[0], oh_ow[1], num_clips=C, to_host=False)
        vids = vids if isinstance(vids, list) else [vids]
        sc = outputs["mask_cls_result"]
        sc = sc if sc.dim() == 3 else sc[None]
        for c in range(C):
            res["scores"][slot + c].copy_(sc[c])
            res["masks"][slot + c].copy_(vids[c]["pred_masks"].bits)
            res["top"][slot + c, 0].copy_(vids[c]["pred_scores"])
            res["top"][slot + c, 1].copy_(vids[c]["pred_labels"])
            res["top"][slot + c, 2].copy_(vids[c]["pred_entropys"])
        return outputs

    def work(self, calls, api_exact=False):
        """Algorithmic work of `calls` pipeline calls per kernel family."""
        Q, T, C = self.Q, self.T, self.C
        N3 = [self.Hp * self.Wp // 1024, self.Hp * self.Wp // 256, self.Hp * self.Wp // 64]
        M = self.Hp * self.Wp // 16
        rows3 = [C * T * n for n in N3]
        TT = C * T
        w = {
            "xattn": ("tensor", sum(4.0 * Q * rows3[i % 3] * 256 for i in range(9))),
            "kv_proj": ("tensor", sum(2.0 * rows3[l] * 1536 * 256 for l in range(3))),
            "prep": ("hbm", sum(r * 256 * (4 + 2 + 2) for r in rows3) + TT * M * 256 * (4 + 2) + sum(rows3) * 256 * 2),
            "mask_logits": ("hbm", (TT * M * 256 * 2 + Q * TT * M * 4) * (10 if api_exact else 1)),
            "mask_bits": ("hbm", sum(rows3[i % 3] * 256 * 2 + Q * rows3[i % 3] / 8 for i in range(9))),
        }
        return {k: (b, a * calls) for k, (b, a) in w.items()}, rows3


def measure(args, name, dev, world, rank, local, steps, warmup, clips, lanes, want_e2e=True, e2e_steps=None, detail=True):
    """Device-resident throughput, per-kernel-family device times and the end-to-end leg for one workload."""
    import torch.distributed as dist
    from openvis_b200 import _lib as L
    from openvis_b200.sharding import gather_clip_dict, shard_range
    P = Pipeline(name, dev, clips, lanes, api_exact=args.api_exact)
    T, C = P.T, P.C
    total = P.total
    # clips of this rank per step: the strong-sharded block of the 64-clip sweep, or `lanes` calls of C clips (weak)
    if total:
        mine = shard_range(total, rank, world)
        n_local = len(mine)
        if n_local % C:
            raise SystemExit(f"--clips {C} must divide this rank's block of {n_local} clips")
    else:
        n_local = C * max(1, lanes)
    calls = n_local // C
    sets = [P.device_inputs(1234 + 2 * rank + j) for j in range(2)]        # two input sets, alternated (inputs >> L2)
    res = P.result_buffers(n_local)
    events = []

    def run_step(record=False, post=True):
        """All calls of one step, round-robin over the lanes (independent clips in flight)."""
        if record:
            L.PROFILE = events
        if len(P.streams) == 1 or record:
            for i in range(calls):
                P.call(0, sets[i % 2], res, i * C, post=post)
        else:
            cur = torch.cuda.current_stream()
            for st in P.streams:
                st.wait_stream(cur)
            for i in range(calls):
                j = i % len(P.streams)
                with torch.cuda.stream(P.streams[j]):
                    P.call(j, sets[i % 2], res, i * C, post=post)
            for st in P.streams:
                cur.wait_stream(st)
        L.PROFILE = None

    def gather():
        """The only collective of the path: per-clip results of every rank's block into global clip order."""
        return gather_clip_dict(res, total) if (total and world > 1) else res

    # the sweep's step includes the device post-processing and the result gather; the per-config workloads time
    # decoder + OV head (+ the temporal stages of their pipeline), post-processing belongs to their e2e leg
    post_in_value = bool(total) or P.pipe != "openvis"
    for _ in range(max(warmup, 1)):
        run_step(post=post_in_value)
    if total and world > 1:
        gather()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = L.launch_count()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g_ev = []
    e0.record()
    for _ in range(steps):
        run_step(post=post_in_value)
        if total and world > 1:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gathered = gather()
            b.record()
            g_ev.append((a, b))
    e1.record()
    torch.cuda.synchronize()
    launches = L.launch_count() - l0
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    gather_ms = sum(a.elapsed_time(b) for a, b in g_ev) / max(1, len(g_ev)) if g_ev else 0.0
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    frames_per_step = (total if total else world * n_local) * T
    value = frames_per_step * steps / (ms * 1e-3)
    out = {"value": value, "ms_per_step": ms / steps, "frames_per_step": frames_per_step, "clips_per_rank_per_step": n_local,
           "launches": int(launches), "clocks": clocks, "timed_region_s": ms * 1e-3}
    if total:
        gbytes = sum(v.numel() * v.element_size() for v in res.values()) // max(1, n_local)
        out["gather"] = {"ms_per_step": gather_ms, "bytes_per_clip": gbytes, "clips": total,
                         "collective": "all_gather (NCCL)" if world > 1 else "none (one rank)"}

    # ---- per-kernel-family device time: CUDA events around every C-ABI call of a single-lane pass
    if detail:
        psteps = max(1, min(steps, 3) if total else steps)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(psteps):
            run_step(record=True, post=post_in_value)
        p1.record()
        torch.cuda.synchronize()
        ms_prof = p0.elapsed_time(p1)
        work, rows3 = P.work(calls * psteps, args.api_exact)
        fam_ms = {}
        for (fam, a_, b_) in events:
            fam_ms[fam] = fam_ms.get(fam, 0.0) + a_.elapsed_time(b_)
        pk, src = peaks()
        # the fractions are quoted against the BURST tensor peak: the clocks of these runs sit at 1.9+ GHz, while the
        # "sustained" figure of MEASURED_PEAKS.json was taken at a 1.34 GHz median
        peak_tf = pk.get("bf16_tflops", pk.get("bf16_tflops_sustained"))
        peak_bw = pk.get("hbm_gbs")
        kernels = {}
        for fam, t_ms in fam_ms.items():
            ent = {"ms_per_step": t_ms / psteps, "share_of_step": t_ms / ms_prof if ms_prof > 0 else None}
            if fam in work and t_ms > 0:
                bound, amount = work[fam]
                if bound == "tensor":
                    ach = amount / (t_ms * 1e-3) / 1e12
                    ent.update(bound="tensor", achieved=ach, peak=peak_tf, unit="TFLOP/s", frac=ach / peak_tf)
                else:
                    ach = amount / (t_ms * 1e-3) / 1e9
                    ent.update(bound="hbm", achieved=ach, peak=peak_bw, unit="GB/s", frac=ach / peak_bw)
            kernels[fam] = ent
        kname = {"xattn": L.xattn_kernel_name() if hasattr(L, "xattn_kernel_name") else "xattn_tc2_kernel+xattn_combine_kernel",
                 "kv_proj": "gemm_tn_bs_kernel<256> (key/value projection)",
                 "prep": "maskfeat_prep_tma_kernel+tokens_prep_tma_kernel", "mask_logits": "gemm_tn_bs_kernel<128> (final mask logits)",
                 "mask_bits": "gemm_tn_bs_kernel<128> (mask sign bits)"}
        if "kv_proj" in kernels and kernels["kv_proj"]["ms_per_step"] > 0:
            byts = sum(rows3[l] * (2 * 512 + 1536 * 2) for l in range(3)) * calls
            ach = byts / (kernels["kv_proj"]["ms_per_step"] * 1e-3) / 1e9
            kernels["kv_proj"]["hbm_view"] = {"achieved": ach, "peak": peak_bw, "unit": "GB/s", "frac": ach / peak_bw}
        if "xattn" in kernels and kernels["xattn"]["ms_per_step"] > 0:
            # d = 32 heads: 128 useful flop per exponential, so the exp pipe (MUFU.EX2, 16 lanes/clk/SM) caps this kernel at
            # ~1/3 of the tensor peak (16 x 148 x 1.9e9 exp/s x 100 flop); reported next to the tensor fraction
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            clk = (clocks.get("sm_mhz") or 1965) * 1e6
            nexp = sum(P.Q * 8.0 * rows3[i % 3] for i in range(9)) * calls
            ach = nexp / (kernels["xattn"]["ms_per_step"] * 1e-3) / 1e9
            peak_x = 16.0 * sms * clk / 1e9
            kernels["xattn"]["xu_bound"] = {"pipe": "XU (MUFU.EX2)", "achieved": ach, "peak": peak_x, "unit": "Gexp/s",
                                            "frac": ach / peak_x,
                                            "note": "useful exponentials (Q rows, not the padded tile); includes the split-combine launches"}
        dom = max((f for f in kernels if f in work), key=lambda f: kernels[f]["ms_per_step"], default=None)
        traffic = None
        for tp in ("ncu_traffic_r2.json", "ncu_traffic_r1.json"):
            tp = os.path.join(ROOT, "profiles", tp)
            if os.path.isfile(tp):
                try:
                    traffic = json.load(open(tp))
                    traffic["_file"] = os.path.basename(tp)
                    break
                except Exception:
                    traffic = None
        roofline = None
        if dom is not None:
            d = kernels[dom]
            per_clip = ((traffic or {}).get(dom, {}) or {}).get("dram_bytes_per_clip")
            roofline = {"kernel": kname[dom], "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                        "frac": d["frac"],
                        "traffic": per_clip * n_local if (per_clip and P.pipe == "openvis" and (P.Hp, P.Wp) == (736, 1280)) else None,
                        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this family's launches (ncu, one clip: "
                                        f"profiles/{(traffic or {}).get('_file')}) x clips per step, i.e. per step like `achieved`",
                        "peak_source": f"{src} ({'bf16_tflops (burst; SM clocks of this run ~1.9 GHz)' if d['bound'] == 'tensor' else 'hbm_gbs'})",
                        "share_of_step": d["share_of_step"]}
            if dom == "xattn" and "xu_bound" in d:
                roofline["binding_pipe"] = d["xu_bound"]
                roofline["note"] = ("masked attention with 32-wide heads: 100 useful flop per exponential, the exp pipe (MUFU.EX2, "
                                    "16 lanes/clk/SM) saturates at ~1/3 of the tensor peak; `binding_pipe` is the fraction of that pipe's peak")
        out["kernels"], out["roofline"] = kernels, roofline

    # ---- end to end through the public API with host buffers (pinned): H2D + pipeline + D2H of the results every step
    if want_e2e:
        n_e2e = e2e_steps or steps

        def to_pinned(t):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            return h
        # ONE clip in pinned host memory, uploaded once per clip of every call into that clip's slice of the device
        # buffers: the same H2D bytes as fully distinct host clips, a fraction of the pinned footprint (8 ranks per box)
        def clip_slice(t, c):
            if t.dim() == 3 and t.shape[0] == 1:                 # CLS tokens [1, frames, 768]
                return t[:, c * T:(c + 1) * T]
            return t[c * T:(c + 1) * T]
        one = [to_pinned(clip_slice(t, 0)) for t in sets[0]]
        torch.cuda.synchronize()
        h2d = n_local * sum(t.numel() * t.element_size() for t in one)
        copy_s = torch.cuda.Stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        res_host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in res.items()}
        d2h = sum(t.numel() * t.element_size() for t in res_host.values())

        def upload(j):
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(done[j])
                for c in range(C):
                    for dst, srcx in zip(sets[j], one):
                        clip_slice(dst, c).copy_(srcx, non_blocking=True)
                ready[j].record(copy_s)

        def e2e_loop(n):
            cur = torch.cuda.current_stream()
            for j in range(2):
                done[j].record(cur)
            upload(0)
            k = 0
            for _ in range(n):
                for i in range(calls):
                    j = k % 2
                    if not (_ == n - 1 and i == calls - 1):
                        upload((k + 1) % 2)    # overlaps the next call's H2D with this call's kernels
                    cur.wait_event(ready[j])
                    P.call(0, sets[j], res, i * C, post=True)
                    done[j].record(cur)
                    k += 1
                for key in res:
                    res_host[key].copy_(res[key], non_blocking=True)
            torch.cuda.synchronize()

        e2e_loop(1)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_loop(n_e2e)
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": frames_per_step * n_e2e / dt.item(), "unit": UNIT, "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "steps": n_e2e,
                      "h2d_gbs_per_rank": h2d * n_e2e / dt.item() / 1e9,
                      "note": "pinned host inputs -> device (double-buffered on a copy stream) -> pipeline incl. device "
                              "post-processing (top-10, x4 up-sample, crop, threshold, bit-pack) -> scores / top-10 / packed "
                              "masks to pinned host; H2D of the fp32 pixel-decoder outputs (80 MB per 720x1280 frame) is the PCIe bound"}
    del P, sets, res
    torch.cuda.empty_cache()
    return out


def measure_next_rows(dev):
    """The SURVEY section 8(f) rows that exist as device code, timed briefly (CUDA events, after a warm-up call) so that the
    default line carries a number for each: f-4 OpenVIS crop classifier (ClipAdapter.forward on one part of 5 frames x 100
    queries at 720 x 1280: every (frame, query) has a non-empty mask = 500 crops through CLIP ViT-B/16) and f-2
    MSDeformAttn.forward at the pixel decoder's encoder shapes (4 frames, three levels of a 736 x 1280 input)."""
    import torch
    from openvis_b200 import _lib as L
    from openvis_b200.clip_adapter import ClipAdapter, ClipVisualEncoder
    from openvis_b200.msda import MSDeformAttn
    from openvis_b200.synthetic import seeded_clip_visual_params
    pk, _ = peaks()
    out = {}
    torch.cuda.empty_cache()

    def timed(fn, n):
        """median device time of n calls after a warm-up call (per-call CUDA events: one call that has to go back to
        cudaMalloc -- the caching allocator after the big sweeps -- must not triple the figure)"""
        fn()
        torch.cuda.synchronize(dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in ev:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize(dev)
        ts = sorted(a.elapsed_time(b) for a, b in ev)
        return ts[len(ts) // 2]

    try:
        T, N, H, W, K = 5, 100, 720, 1280, 1196
        g = torch.Generator(device=dev).manual_seed(11)
        ad = ClipAdapter(ClipVisualEncoder().load_state_dict(seeded_clip_visual_params(3)))
        frames = torch.rand(T, 3, H, W, generator=g, device=dev) * 255
        logits = torch.full((N, T, H, W), -6.0, device=dev)
        hw = torch.randint(24, 400, (N, T, 2), generator=g, device=dev).cpu()
        for n in range(N):
            for t in range(T):
                h, w = int(hw[n, t, 0]), int(hw[n, t, 1])
                y0, x0 = (n * 37 + t * 11) % (H - h), (n * 53 + t * 7) % (W - w)
                logits[n, t, y0:y0 + h, x0:x0 + w] = 6.0
        text = make_text(K).to(dev)
        n0 = L.launch_count()
        ms_all = timed(lambda: ad(frames, text, logits, layout="nt", logits=True), 3)
        launches = (L.launch_count() - n0) // 4
        ms_pre = timed(lambda: ad._preprocess_image(frames, logits, layout="nt", logits=True), 3)
        crops = T * N
        Lt, Wd = 197, 768
        flops = crops * (2 * 196 * Wd * Wd + 12 * (2 * Lt * Wd * (3 * Wd + Wd + 4 * Wd + 4 * Wd) + 4 * Lt * Lt * Wd) + 2 * Wd * 512 + 2 * 512 * K)
        ms_vit = ms_all - ms_pre
        out["f4_crop_classifier"] = {
            "workload": "openvis_crop_classifier_5x720x1280_q100_k1196 (one part of open_vocabulary_inference, openvis.py:112-122)",
            "crops": crops, "ms": ms_all, "crops_per_s": crops / ms_all * 1e3, "frames_per_s": T / ms_all * 1e3,
            "ms_preprocess": ms_pre, "preprocess_hbm_gbs": (logits.numel() * 4 * 1.0 + frames.numel() * 4) / ms_pre / 1e6,
            "ms_clip_tower": ms_vit, "tflops_clip_tower": flops / ms_vit / 1e9,
            "tensor_frac_of_burst_peak": flops / ms_vit / 1e9 / pk.get("bf16_tflops", 1667.8), "gpu_launches": launches}
        del ad, frames, logits
    except Exception as e:
        out["f4_crop_classifier"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    try:
        g = torch.Generator(device=dev).manual_seed(12)
        shapes = torch.tensor([(92, 160), (46, 80), (23, 40)], device=dev)
        start = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
        S, Nf = int((shapes[:, 0] * shapes[:, 1]).sum()), 4
        m = MSDeformAttn(256, 3, 8, 4).to(dev).eval()
        with torch.no_grad():
            m.sampling_offsets.weight.normal_(0, 0.02, generator=None)
        src = torch.randn(Nf, S, 256, generator=g, device=dev)
        ref = torch.rand(Nf, S, 3, 2, generator=g, device=dev)
        n0 = L.launch_count()
        ms = timed(lambda: m(src, ref, src, shapes, start), 5)
        out["f2_msdeformattn_module"] = {
            "workload": "MSDeformAttn.forward, 4 frames x (92x160 + 46x80 + 23x40) positions, 8 heads x 3 levels x 4 points",
            "ms": ms, "ms_per_frame": ms / Nf, "gpu_launches": (L.launch_count() - n0) // 6,
            "positions_per_s": Nf * S / ms * 1e3}
    except Exception as e:
        out["f2_msdeformattn_module"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    try:
        # the whole pixel decoder (input projections + GroupNorm, six deformable encoder layers, FPN level, mask features) on
        # ResNet-50-shaped backbone maps of 12 frames of 736 x 1280, handed straight to the Video decoder
        from openvis_b200.pixel_decoder import MSDeformAttnPixelDecoder, ShapeSpec
        from openvis_b200.synthetic import seeded_pixel_decoder_params
        ch, Nf, Hp, Wp = (256, 512, 1024, 2048), 12, 736, 1280
        pd = MSDeformAttnPixelDecoder({f"res{i + 2}": ShapeSpec(channels=c, stride=4 << i) for i, c in enumerate(ch)})
        pd.load_state_dict(seeded_pixel_decoder_params(2, in_channels=ch))
        pd = pd.to(dev)
        g = torch.Generator(device=dev).manual_seed(13)
        feats = {f"res{i + 2}": torch.randn(Nf, c, Hp // (4 << i), Wp // (4 << i), generator=g, device=dev) for i, c in enumerate(ch)}
        torch.cuda.empty_cache()
        n0 = L.launch_count()
        ms = timed(lambda: pd.forward_features(feats), 5)
        S = sum((Hp // s) * (Wp // s) for s in (8, 16, 32))
        M4 = (Hp // 4) * (Wp // 4)
        flops = Nf * (2 * 256 * sum(c * (Hp // (4 << i)) * (Wp // (4 << i)) for i, c in enumerate(ch))     # 1x1 projections
                      + 6 * S * 2 * 256 * (256 + 288 + 256 + 2 * 1024)                                     # encoder GEMMs
                      + M4 * 2 * 256 * (9 * 256 + 256))                                                   # 3x3 + mask features
        out["f2_pixel_decoder"] = {
            "workload": "MSDeformAttnPixelDecoder.forward_features, 12 frames of 736x1280, ResNet-50 channel counts, 6 encoder layers",
            "ms": ms, "frames_per_s": Nf / ms * 1e3, "gpu_launches": (L.launch_count() - n0) // 6,
            "gemm_tflops": flops / ms / 1e9, "tensor_frac_of_burst_peak": flops / ms / 1e9 / pk.get("bf16_tflops", 1667.8)}
        # pixel decoder -> Video decoder: through the reference's fp32 NCHW maps, and through the fp16 token-major hand-off
        from openvis_b200.decoder import VideoMultiScaleMaskedTransformerDecoder
        from openvis_b200.synthetic import decoder_param_shapes, seeded_params
        dec = VideoMultiScaleMaskedTransformerDecoder(in_channels=256, mask_classification=True, num_classes=40, hidden_dim=256,
                                                      num_queries=100, nheads=8, dim_feedforward=2048, dec_layers=9, pre_norm=False,
                                                      mask_dim=256, enforce_input_project=False, num_frames=Nf).eval().to(dev)
        dec.load_state_dict(seeded_params(decoder_param_shapes("video", num_classes=40), seed=0))

        def nchw():
            mf, _, ms = pd.forward_features(feats)
            return dec(ms, mf)

        ms_nchw = timed(nchw, 5)
        ms_tok = timed(lambda: dec.forward_tokens(pd.forward_tokens(feats)), 5)
        out["f2_pixel_decoder_to_decoder"] = {
            "workload": "pixel decoder + Video decoder, one 12-frame 736x1280 clip (device-resident backbone maps)",
            "ms_nchw_fp32_interface": ms_nchw, "frames_per_s_nchw": Nf / ms_nchw * 1e3,
            "ms_token_handoff_fp16": ms_tok, "frames_per_s_token_handoff": Nf / ms_tok * 1e3}
        del pd, feats, dec
    except Exception as e:
        out["f2_pixel_decoder"] = out.get("f2_pixel_decoder") or {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        out.setdefault("f2_pixel_decoder_to_decoder", {"error": f"{type(e).__name__}: {str(e)[:200]}"})
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    numa = bind_to_gpu_numa_node(local)
    pipe = WORKLOADS[args.workload][0]
    lanes = args.streams if args.streams is not None else (1 if pipe == "san_online" else 2)
    clips = args.clips if args.clips is not None else 4
    # persistent kernels leave a few SMs to the other in-flight call's small latency-bound kernels
    if lanes > 1:
        os.environ.setdefault("OVIS_SM_BUDGET", "140")
    from openvis_b200 import _lib as L
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the result gather runs alone at the end of a step: give NCCL all the channels it can use (measured with
        # tools/prof_gather.py: 2 GPUs 2.90 -> 2.20 ms, 8 GPUs 3.75 -> 3.46 ms for the 64 clips' 2.7 GB)
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "64")
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "64")
        dist.init_process_group("nccl", device_id=dev)
    L.device_check()
    torch.set_grad_enabled(False)

    name = args.workload
    _, kind, T, Hp, Wp, out_hw, Q, K, total = WORKLOADS[name]
    e2e_steps = max(2, min(args.steps, 4)) if total else None            # the sweep uploads 185 GB per step at N = 1
    r = measure(args, name, dev, world, rank, local, args.steps, args.warmup, clips, lanes, want_e2e=not args.no_e2e,
                e2e_steps=e2e_steps)

    others = None
    next_rows = None
    if world == 1 and name == DEFAULT_WORKLOAD and not args.no_other_configs:
        others = {}
        for on in OTHER_CONFIGS:
            op = WORKLOADS[on][0]
            ol = 1 if op == "san_online" else 2      # (two calls in flight: +9 % for BriVIS, nothing for SAN-online, r2_lanes)
            oc = 1 if on == "openvis_video_5x360x640_q100_k40" else (2 if WORKLOADS[on][6] == 200 else 4)
            try:
                m = measure(args, on, dev, 1, 0, local, max(5, args.steps // 2), 3, oc, ol, want_e2e=not args.no_e2e,
                            e2e_steps=3, detail=True)
                kern = {k: {"ms_per_step": round(v["ms_per_step"], 4), **({"frac": round(v["frac"], 4), "bound": v["bound"]} if "frac" in v else {})}
                        for k, v in (m.get("kernels") or {}).items()}
                others[on] = {"baseline_config": BASELINE_CONFIG[on], "value": m["value"], "unit": UNIT,
                              "ms_per_step": m["ms_per_step"], "frames_per_step": m["frames_per_step"], "clips_per_call": oc,
                              "calls_in_flight": ol, "gpu_launches": m["launches"],
                              "e2e": (m.get("e2e") or {}).get("value"), "kernels": kern,
                              "gflop_per_frame": algorithmic_flops_per_frame(WORKLOADS[on][1], *WORKLOADS[on][2:5], WORKLOADS[on][6], WORKLOADS[on][7]) / 1e9}
            except Exception as e:                                  # a secondary measurement must not lose the headline
                others[on] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        try:
            next_rows = measure_next_rows(dev)
        except Exception as e:
            next_rows = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        prev = getattr(bind_to_gpu_numa_node, "previous", None)
        if prev:                                                    # the CPU baseline may use every host core again
            os.sched_setaffinity(0, prev)
        cpu = cpu_baseline(name)
    pk, _ = peaks()
    flops_frame = algorithmic_flops_per_frame(kind, T, Hp, Wp, Q, K)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if total else "weak",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": name, "baseline_config": BASELINE_CONFIG[name], "frames_per_step": r["frames_per_step"],
                       "clips_per_step": r["frames_per_step"] // T, "clips_per_rank_per_step": r["clips_per_rank_per_step"],
                       "clips_per_call": clips, "queries": Q, "vocab": K,
                       "l2": "inputs larger than L2 (2.9 GB per 720x1280 clip, two input sets alternated)",
                       "aux_outputs": "materialised: nine intermediate heads written per call (API-exact)" if args.api_exact else "lazy (inference-minimal)",
                       "parallelism": f"clip-sharded dp{world}" + (" (contiguous blocks of the 64 clips, result gather closes the step)" if total else ""),
                       "calls_in_flight_per_gpu": lanes, "sm_budget": os.environ.get("OVIS_SM_BUDGET"),
                       "timed_region_s": r["timed_region_s"], "host_affinity": numa,
                       "masked_tile_skipping": "off (OVIS_XATTN_SKIP=0)" if os.environ.get("OVIS_XATTN_SKIP") == "0" else
                       "on; the synthetic N(0,1) features give ~50 % dense masks, so no fully-masked tile exists to skip here"},
            "clocks": r["clocks"], "e2e": r.get("e2e"), "gpu_launches": r["launches"],
            "roofline": r.get("roofline"), "kernels": r.get("kernels"), "cpu_baseline": cpu,
            "gather": r.get("gather"), "other_configs": others, "next_rows": next_rows,
            "whole_path": {"gflop_per_frame": flops_frame / 1e9,
                           "tensor_frac_of_burst_peak": r["value"] / world * flops_frame / 1e12 / pk.get("bf16_tflops", 1667.8)}}
    line["ms_per_frame"] = r["ms_per_step"] / r["frames_per_step"] * world
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
