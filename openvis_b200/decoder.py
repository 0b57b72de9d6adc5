"""B200-native drop-in replacements for the reference's masked transformer decoders.

Same class names, constructor keywords, ``from_config``, parameter names / shapes, ``forward``
signature and output dictionaries as

  VideoMultiScaleMaskedTransformerDecoder              openvis/modeling/transformer_decoder/video_mask2former_transformer_decoder.py:219-484
  FrameMultiScaleMaskedTransformerDecoder              .../frame_mask2former_transformer_decoder.py:12-154
  SideAdapterFrameMultiScaleMaskedTransformerDecoder   .../side_adapter_frame_mask2former_transformer_decoder.py:29-176
  SideAdapterVideoMultiScaleMaskedTransformerDecoder   .../side_adapter_video_mask2former_transformer_decoder.py:28-149
  (+ the Embedding*/Proposal* variants that only swap ``class_embed``)

so that ``MaskFormerHead`` (mask_former_head.py:112-124) can build and call them unchanged.  The modules are
parameter containers plus a launch schedule: every tensor operation on the hot path is one of the hand-written
sm_100a kernels behind the C ABI (openvis_b200/_lib.py); PyTorch only allocates memory and provides the stream.
Inference only (eval mode, no autograd), CUDA sm_100 only, no fallback.

Design notes (DESIGN.md has the full account):
  * a "group" is the set of frames that share one set of queries: the whole clip for the Video decoders (joint
    attention over T*HW keys), a single frame for the Frame decoders;
  * the per-layer attention mask is never materialised as bools: the mask head's GEMM epilogue writes one sign bit
    per (query, key) and a per-row "has an unblocked key" flag; the cross-attention kernel expands them in-kernel;
  * because bilinear down-sampling by an integer factor is the mean of the centre 2x2 pixels and the mask logit is
    linear in the mask features, the intermediate heads multiply mask_embed with centre-pooled features (1/64, 1/16,
    1/4 of the pixels) instead of the full-resolution map; only the last head produces full-resolution logits.
    ``aux_outputs`` (never read in eval by the reference's meta-architectures) are computed on first access.
"""
import math
import weakref
from typing import List

import os

import torch
from torch import nn

from . import _lib as L

HIDDEN = 256          # the kernels are specialised for hidden_dim 256 = 8 heads x 32
NHEADS = 8
LOG2E = 1.4426950408889634


# ------------------------------------------------------------------------------------------------ registry
class Registry(dict):
    """Minimal stand-in for detectron2.utils.registry.Registry (video_..._decoder.py:16)."""

    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        def deco(o):
            self[o.__name__] = o
            return o
        return deco if obj is None else deco(obj)


TRANSFORMER_DECODER_REGISTRY = Registry("TRANSFORMER_MODULE")


def build_transformer_decoder(cfg, in_channels, mask_classification=True):
    """Same contract as video_mask2former_transformer_decoder.py:21-26."""
    name = cfg.MODEL.MASK_FORMER.TRANSFORMER_DECODER_NAME
    return TRANSFORMER_DECODER_REGISTRY.get(name)(cfg, in_channels, mask_classification)


def register_into(registry):
    """Registers (overrides) the B200 decoders in the reference's own TRANSFORMER_DECODER_REGISTRY
    (a detectron2 Registry or any mapping), see INTEGRATION.md."""
    for name, cls in TRANSFORMER_DECODER_REGISTRY.items():
        obj_map = getattr(registry, "_obj_map", registry)
        obj_map[name] = cls


# ------------------------------------------------------------------------------------------------ pos. embeddings
def _sincos(coord, npf, temperature=10000.0):
    i = torch.arange(npf, dtype=torch.float32, device=coord.device)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / npf)
    ang = coord[..., None] / dim_t
    return torch.stack((ang[..., 0::2].sin(), ang[..., 1::2].cos()), dim=-1).flatten(-2)


def sine_pos_2d(h, w, device, npf=HIDDEN // 2):
    """PositionEmbeddingSine2D(normalize=True) with no padding mask (position_encoding.py:78-103) -> [h*w, 2*npf]."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32, device=device)
    x = torch.arange(1, w + 1, dtype=torch.float32, device=device)
    py = _sincos(y / (y[-1] + eps) * scale, npf)
    px = _sincos(x / (x[-1] + eps) * scale, npf)
    return torch.cat([py[:, None, :].expand(h, w, npf), px[None, :, :].expand(h, w, npf)], dim=-1).reshape(h * w, 2 * npf)


def sine_pos_z(t, device, npf=HIDDEN // 2):
    """frame term of PositionEmbeddingSine3D (position_encoding.py:141-163) -> [t, 2*npf]; pos3d = pos2d + pos_z."""
    eps, scale = 1e-6, 2 * math.pi
    z = torch.arange(1, t + 1, dtype=torch.float32, device=device)
    return _sincos(z / (z[-1] + eps) * scale, 2 * npf)


# ------------------------------------------------------------------------------------------------ containers
class _AttnLayer(nn.Module):
    """Parameter container with the reference layer's names (SelfAttentionLayer / CrossAttentionLayer)."""

    def __init__(self, attn_name, d_model, nhead):
        super().__init__()
        setattr(self, attn_name, nn.MultiheadAttention(d_model, nhead, dropout=0.0))
        self.norm = nn.LayerNorm(d_model)
        for p in self.parameters():          # _reset_parameters (video_..._decoder.py:44-47)
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class _FFNLayer(nn.Module):
    def __init__(self, d_model, dim_feedforward):
        super().__init__()
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)


class MLP(nn.Module):
    """Container matching MLP (video_..._decoder.py:204-216)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, affine_func=nn.Linear):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(affine_func(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


class _LazyDict(dict):
    """dict whose values may be zero-argument callables resolved (once) on first access."""

    def __getitem__(self, k):
        v = super().__getitem__(k)
        if callable(v) and getattr(v, "_ovis_lazy", False):
            v = v()
            super().__setitem__(k, v)
        return v

    def get(self, k, default=None):
        return self[k] if k in self else default

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def values(self):
        return [self[k] for k in self.keys()]

    def copy(self):
        """A plain dict with every lazy entry resolved (the reference's training path calls image_outputs.copy());
        dict(out) / {**out} bypass __getitem__ and would expose the raw callables, so resolve through copy() or items()."""
        return dict(self.items())


def _lazy(fn):
    fn._ovis_lazy = True
    return fn


class LazyAuxOutputs(list):
    """``aux_outputs`` of the reference (``_set_aux_loss``, video_..._decoder.py:473-484): one dict per intermediate
    prediction head.  The reference's eval paths never read them (openvis.py:84-85, minvis.py:352-353), and writing
    nine full-resolution mask tensors would make the whole path HBM-bound, so each entry is computed from the saved
    per-layer query state the first time it is indexed."""

    def __init__(self, n, compute):
        super().__init__([None] * n)
        self._compute = compute

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        v = super().__getitem__(i)
        if v is None:
            v = self._compute(i if i >= 0 else len(self) + i)
            super().__setitem__(i, v)
        return v

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _words_from_bits_t(bits_t, Q):
    """tests' debug capture only: key-major mask bits [G, keys, qw] -> the query-major packing [G, ceil(keys/32), Q]."""
    G, keys, qw = bits_t.shape
    sh = torch.arange(32, device=bits_t.device, dtype=torch.int32)
    blocked = ((bits_t[..., None] >> sh) & 1).flatten(-2)[..., :Q]                      # [G, keys, Q]
    W = (keys + 31) // 32
    pad = torch.zeros(G, W * 32, Q, dtype=torch.int64, device=bits_t.device)
    pad[:, :keys] = blocked
    words = (pad.view(G, W, 32, Q) << torch.arange(32, device=bits_t.device)[None, None, :, None]).sum(2)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


# ------------------------------------------------------------------------------------------------ the decoder
class _B200MaskedDecoderBase(nn.Module):
    _version = 2
    VIDEO = False     # joint attention over the clip (Video decoders) vs per-frame (Frame decoders)
    SAN = False       # side-adapter variant (attention-bias branch instead of class_embed)

    def __init__(self, in_channels=None, mask_classification=True, *, num_classes: int = None, hidden_dim: int = None,
                 num_queries: int = None, nheads: int = None, dim_feedforward: int = None, dec_layers: int = None,
                 pre_norm: bool = None, mask_dim: int = None, enforce_input_project: bool = None, num_frames=None,
                 **extra):
        super().__init__()
        if hidden_dim != HIDDEN or nheads != NHEADS or mask_dim != HIDDEN:
            raise NotImplementedError("openvis_b200 kernels are specialised for hidden_dim = mask_dim = 256, nheads = 8 "
                                      f"(got hidden_dim={hidden_dim}, nheads={nheads}, mask_dim={mask_dim})")
        if pre_norm:
            raise NotImplementedError("PRE_NORM True is not supported (all shipped configs use post-norm)")
        if in_channels != hidden_dim or enforce_input_project:
            raise NotImplementedError("input_proj 1x1 convs are not supported (in_channels must equal hidden_dim, "
                                      "ENFORCE_INPUT_PROJ False, as in every shipped config)")
        if num_queries > 256:
            raise NotImplementedError("at most 256 object queries")
        self.mask_classification = mask_classification
        self.num_frames = num_frames
        self.num_heads = nheads
        self.num_layers = dec_layers
        self.num_queries = num_queries
        self.num_feature_levels = 3
        self.dim_feedforward = dim_feedforward
        self.transformer_self_attention_layers = nn.ModuleList(
            _AttnLayer("self_attn", hidden_dim, nheads) for _ in range(dec_layers))
        self.transformer_cross_attention_layers = nn.ModuleList(
            _AttnLayer("multihead_attn", hidden_dim, nheads) for _ in range(dec_layers))
        self.transformer_ffn_layers = nn.ModuleList(_FFNLayer(hidden_dim, dim_feedforward) for _ in range(dec_layers))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.level_embed = nn.Embedding(3, hidden_dim)
        self.input_proj = nn.ModuleList(nn.Sequential() for _ in range(3))
        if self.mask_classification:
            self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)
        # non-reference knobs
        self.materialize_aux = False      # True: compute all nine aux_outputs eagerly (API-exact mode)
        # Video decoders only: number of independent clips stacked along the frame axis of one call.  The reference
        # forces bs = 1 in eval (video_..._decoder.py:382-383); >1 is a throughput extension (same arithmetic as its
        # training-mode bs > 1): the query-side GEMMs then run on clips_per_call * Q rows per launch.
        self.clips_per_call = 1
        self.debug_capture = None         # tests: set to a list to receive (head, level, bits, flags) clones
        self.use_cuda_graph = os.environ.get("OVIS_NO_CUDA_GRAPH") is None   # replay the layer loop as a CUDA graph (_run_layers)
        # One launch per layer for everything between two cross-attentions (csrc/chain.cuh).  Two forms:
        #   "wide"  -- the layer's phases as ONE cooperative launch, every phase's tiles spread over all CTAs, grid barriers
        #              between dependent phases, LayerNorms as split-K partials + a row-parallel reduction: 118-124 us per
        #              layer at 100-400 query rows against 135-186 us for the launch-per-op schedule; loses beyond ~2000
        #              rows, where the separate kernels already fill the GPU (profiles/experiments/chain_r2.md);
        #   "group" -- one CTA per group (<= 128 queries) runs the layer alone (round-2 first version, 254-265 us per layer;
        #              kept for A/B).
        # use_chain: None = wide up to WIDE_CHAIN_AUTO_ROWS query rows, else launch per op; "wide" / True ("group") / False
        # force a schedule (OVIS_CHAIN=wide / 1 / 0; OVIS_CHAIN_WIDE_DEFAULT=0 restores the round-2a rule: group chain for
        # single-group calls).
        env = os.environ.get("OVIS_CHAIN")
        self.use_chain = None if env is None else ("wide" if env == "wide" else env == "1")
        self.wide_chain_default = os.environ.get("OVIS_CHAIN_WIDE_DEFAULT", "1") == "1"
        self._wcache = None
        self._pcache = {}
        self._ws = {}
        self._generation = 0

    # -- config ------------------------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        """video_mask2former_transformer_decoder.py:351-378."""
        ret = {"in_channels": in_channels, "mask_classification": mask_classification}
        ret["num_classes"] = cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES
        ret["hidden_dim"] = cfg.MODEL.MASK_FORMER.HIDDEN_DIM
        ret["num_queries"] = cfg.MODEL.MASK_FORMER.NUM_OBJECT_QUERIES
        ret["nheads"] = cfg.MODEL.MASK_FORMER.NHEADS
        ret["dim_feedforward"] = cfg.MODEL.MASK_FORMER.DIM_FEEDFORWARD
        assert cfg.MODEL.MASK_FORMER.DEC_LAYERS >= 1
        ret["dec_layers"] = cfg.MODEL.MASK_FORMER.DEC_LAYERS - 1
        ret["pre_norm"] = cfg.MODEL.MASK_FORMER.PRE_NORM
        ret["enforce_input_project"] = cfg.MODEL.MASK_FORMER.ENFORCE_INPUT_PROJ
        ret["mask_dim"] = cfg.MODEL.SEM_SEG_HEAD.MASK_DIM
        ret["num_frames"] = cfg.INPUT.SAMPLING_FRAME_NUM
        return ret

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """legacy-key upgrade of the reference (video_..._decoder.py:224-245): static_query -> query_feat."""
        version = local_metadata.get("version", None)
        if version is None or version < 2:
            for k in list(state_dict.keys()):
                if k.startswith(prefix) and "static_query" in k:
                    state_dict[k.replace("static_query", "query_feat")] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    # -- weight / table caches ---------------------------------------------------------------------------------
    def _wkey(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _weights(self):
        """fp16 operand copies of the weights + fused biases; rebuilt whenever a parameter changes."""
        key = self._wkey()
        if self._wcache is not None and self._wcache["key"] == key:
            return self._wcache
        f16 = lambda t: L.cast_f16(t.detach().float().contiguous())
        f32 = lambda t: t.detach().float().contiguous()
        W = {"key": key, "layers": []}
        C = HIDDEN
        for i in range(self.num_layers):
            ca = self.transformer_cross_attention_layers[i]
            sa = self.transformer_self_attention_layers[i]
            ff = self.transformer_ffn_layers[i]
            cw, cb = ca.multihead_attn.in_proj_weight, ca.multihead_attn.in_proj_bias
            sw, sb = sa.self_attn.in_proj_weight, sa.self_attn.in_proj_bias
            W["layers"].append(dict(
                xq_w=f16(cw[:C]), xq_b=f32(cb[:C]),
                xk_w32=f32(cw[C:2 * C]), xk_b=f32(cb[C:2 * C]), xv_w32=f32(cw[2 * C:]), xv_b=f32(cb[2 * C:]),
                xo_w=f16(ca.multihead_attn.out_proj.weight), xo_b=f32(ca.multihead_attn.out_proj.bias),
                ln_x=(f32(ca.norm.weight), f32(ca.norm.bias)),
                sqk_w=f16(sw[:2 * C]), sqk_b=f32(sb[:2 * C]), sv_w=f16(sw[2 * C:]), sv_b=f32(sb[2 * C:]),
                so_w=f16(sa.self_attn.out_proj.weight), so_b=f32(sa.self_attn.out_proj.bias),
                ln_s=(f32(sa.norm.weight), f32(sa.norm.bias)),
                f1_w=f16(ff.linear1.weight), f1_b=f32(ff.linear1.bias),
                f2_w=f16(ff.linear2.weight), f2_b=f32(ff.linear2.bias),
                ln_f=(f32(ff.norm.weight), f32(ff.norm.bias)),
            ))
        # K/V projection weights of the layers that read level l, stacked [K_i, V_i, K_i+3, V_i+3, ...]
        W["kv_w"], W["kv_layers"] = [], []
        for l in range(3):
            ids = [i for i in range(self.num_layers) if i % 3 == l]
            W["kv_layers"].append(ids)
            if ids:
                stack = torch.cat([torch.cat([W["layers"][i]["xk_w32"], W["layers"][i]["xv_w32"]]) for i in ids])
                W["kv_w"].append(L.cast_f16(stack.contiguous()))
            else:
                W["kv_w"].append(None)
        le = f32(self.level_embed.weight)
        for i, lw in enumerate(W["layers"]):
            e = le[i % 3]
            lw["v_bias"] = (lw["xv_b"] + lw["xv_w32"] @ e).contiguous()      # value = (x + level_embed) W_v^T + b_v
        W["dn"] = (f32(self.decoder_norm.weight), f32(self.decoder_norm.bias))
        W["qf"], W["qe"] = f32(self.query_feat.weight), f32(self.query_embed.weight)
        W["mask_embed"] = [(f16(m.weight), f32(m.bias)) for m in self.mask_embed.layers]
        if hasattr(self, "class_embed"):
            ce = self.class_embed
            W["class_embed"] = [(f16(m.weight), f32(m.bias)) for m in (ce.layers if isinstance(ce, MLP) else [ce])]
        if hasattr(self, "object_embed"):
            W["object_embed"] = [(f16(m.weight), f32(m.bias)) for m in self.object_embed.layers]
        if self.SAN:
            W["attn_embed"] = [(f16(m.weight), f32(m.bias)) for m in self.attn_embed.layers]
            W["attn_mlp"] = [(f16(m.weight.flatten(1)), f32(m.bias)) for m in self.attn_mlp.layers]
        self._wcache = W
        self._pcache = {}
        return W

    def _pos_tables(self, T, sizes, device):
        """Input-independent additive tables of the key operand: padd[l] = (pos2d_l + level_embed_l)^T  [256, N_l] fp32 and,
        for the Video decoders, pz = frame term [T, 256] (pos3d = pos2d + pos_z, position_encoding.py:163).
        Cached per (level_embed version, T, sizes)."""
        le = self.level_embed.weight
        key = (T if self.VIDEO else 0, tuple(sizes), le.data_ptr(), le._version)
        hit = self._pcache.get(key)
        if hit is not None:
            return hit
        p2 = [sine_pos_2d(h, w, device) for (h, w) in sizes]
        pz = sine_pos_z(T, device).contiguous() if self.VIDEO else None
        lef = le.detach().float()
        # channel-major [256, N_l] (the layout of the NCHW inputs): what the TMA-fed layout kernel reads coalesced
        padd = [(p2[l] + lef[l][None, :]).t().contiguous() for l in range(3)]
        if len(self._pcache) > 4:
            self._pcache.clear()
        self._pcache[key] = (padd, p2, pz)
        return self._pcache[key]

    def _groups(self, BT):
        if not self.VIDEO:
            return BT
        G = int(self.clips_per_call)
        if G < 1 or BT % G:
            raise ValueError(f"clips_per_call={G} does not divide the {BT} frames of this call")
        return G

    def _workspace(self, BT, H4, W4, device):
        key = (BT, H4, W4, str(device), self._groups(BT))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        if len(self._ws) >= 2:           # keep at most two shapes resident (a clip-sized workspace is GBs)
            self._ws.pop(next(iter(self._ws)))
        Q, C = self.num_queries, HIDDEN
        G = self._groups(BT)
        Tg = BT // G
        R = G * Q
        N = [(H4 // s) * (W4 // s) for s in (8, 4, 2)]
        M = H4 * W4
        h = lambda *s: torch.empty(*s, dtype=torch.float16, device=device)
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        ws = dict(G=G, Tg=Tg, R=R, N=N, M=M)
        ws["xt"] = [h(BT * n, C) for n in N]              # value operand: x
        ws["xp"] = [h(BT * n, C) for n in N]              # key operand:   x + level_embed + pos
        ws["ft"] = h(BT * M, C)
        ws["gt"] = [h(BT * n, C) for n in N]
        ws["k"] = [h(BT * N[i % 3], C) for i in range(self.num_layers)]
        ws["v"] = [h(BT * N[i % 3], C) for i in range(self.num_layers)]
        ws["z32"], ws["z16"], ws["ze16"] = f(R, C), h(R, C), h(R, C)
        ws["d32"] = f(R, C)
        ws["d16"] = h(self.num_layers + 1, R, C)            # decoder_norm output of every head (lazy aux needs them)
        ws["q16"], ws["att16"], ws["qk16"], ws["v16"], ws["sa16"] = h(R, C), h(R, C), h(R, 2 * C), h(R, C), h(R, C)
        ws["h16"] = h(R, self.dim_feedforward)
        ws["m1"], ws["m2"], ws["me16"] = h(R, C), h(R, C), h(R, C)
        # scratch of the split linear+LayerNorm path (split-K partials); beyond 16384 rows the fused epilogue is used
        ws["split"] = f((self.dim_feedforward // 256) * ((R + 127) // 128) * 128 * 256) if R <= 16384 else None
        ws["flags"] = torch.zeros(self.num_layers + 1, G, Q, dtype=torch.uint8, device=device)
        # Per level: the transposed-score attention kernel (xattn_tc3, key-major mask bits) where a CTA walks a long run of
        # keys, else the query-major kernel (xattn_tc2, [word][Q] mask bits); ovis_xattn_plan_t decides.
        plans_t = [L.xattn_plan_t(G, Q, Tg * n) for n in N]
        plans = [L.xattn_plan(G, Q, Tg * n) for n in N]
        ws["use_t"] = [pt[0] for pt in plans_t]
        ws["splits"] = [pt[1] if pt[0] else p[0] for pt, p in zip(plans_t, plans)]
        qw = 4 * ((Q + 127) // 128)
        ws["bits"] = [None if ut else torch.zeros(G, (Tg * n + 31) // 32, Q, dtype=torch.int32, device=device)
                      for n, ut in zip(N, ws["use_t"])]
        ws["bits_t"] = [torch.zeros(G, Tg * n, qw, dtype=torch.int32, device=device) if ut else None for n, ut in zip(N, ws["use_t"])]
        ws["blockand"] = [torch.zeros(G, (Tg * n + 31) // 32, qw, dtype=torch.int32, device=device) if ut else None
                          for n, ut in zip(N, ws["use_t"])]
        ws["o_part"] = f(max(max(p[2], pt[3]) for p, pt in zip(plans, plans_t)))
        ws["ml_part"] = f(max(max(p[3], pt[4]) for p, pt in zip(plans, plans_t)))
        self._ws[key] = ws
        return ws

    # -- pieces of the schedule ----------------------------------------------------------------------------------
    @staticmethod
    def _mlp3(params, x16, t1, t2, out16):
        L.linear_f16(x16, params[0][0], params[0][1], relu=True, out=t1)
        L.linear_f16(t1, params[1][0], params[1][1], relu=True, out=t2)
        L.linear_f16(t2, params[2][0], params[2][1], relu=False, out=out16)
        return out16

    def _check_inputs(self, x, mask_features):
        if self.training:
            raise RuntimeError("openvis_b200 decoders are inference-only: call .eval() (training / autograd is out of scope)")
        if torch.is_grad_enabled() and any(t.requires_grad for t in list(x) + [mask_features]):
            raise RuntimeError("openvis_b200 decoders do not support autograd; wrap the call in torch.no_grad()")
        assert len(x) == self.num_feature_levels
        BT, C, H4, W4 = mask_features.shape
        if C != HIDDEN or H4 % 8 or W4 % 8:
            raise NotImplementedError(f"mask_features must be [BT, 256, H/4, W/4] with H, W multiples of 32, got {tuple(mask_features.shape)}")
        sizes = []
        for l, s in enumerate((8, 4, 2)):
            exp = (BT, C, H4 // s, W4 // s)
            if tuple(x[l].shape) != exp:
                raise NotImplementedError(
                    f"multi-scale feature {l} must be {exp} (strides 32/16/8 of a /32-padded input, coarsest first); "
                    f"got {tuple(x[l].shape)}: the centre-2x2 mask down-sampling identity needs integer factors")
            sizes.append((H4 // s, W4 // s))
        if not mask_features.is_cuda or any(not t.is_cuda for t in x):
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        return BT, H4, W4, sizes

    def forward(self, x: List[torch.Tensor], mask_features: torch.Tensor, mask=None):
        del mask                                  # "disable mask, it does not affect performance" (frame_...:59-60)
        BT, H4, W4, sizes = self._check_inputs(x, mask_features)     # (autograd guard: evaluated before no_grad below)
        with torch.no_grad():
            return self._forward_nograd(x, mask_features, BT, H4, W4, sizes)

    def _forward_nograd(self, x, mask_features, BT, H4, W4, sizes):
        dev = mask_features.device
        x = [t.float().contiguous() for t in x]
        mask_features_in = mask_features
        mf = mask_features.float().contiguous()
        with torch.cuda.device(dev):
            return self._forward_impl(x, mf, mask_features_in, BT, H4, W4, sizes, dev)

    def _build_chain(self, W, ws):
        """Phase list of the query-side chain: [mask_embed MLP of head 0, q-projection of layer 0], then per layer
        [out-proj + LN, self-attention in-projections, self-attention, its out-proj + LN, FFN1, FFN2 + LN + decoder_norm,
        mask_embed MLP of the next head, q-projection of the next layer]."""
        nl, R = self.num_layers, ws["R"]
        qscale = (HIDDEN // NHEADS) ** -0.5 * LOG2E
        me = W["mask_embed"]
        first, count = [], []
        n = 4 + sum(10 + (1 if i + 1 < nl else 0) for i in range(nl))
        wide = self._chain_mode(ws) == "wide"
        ch = L.Chain(n, ws["G"], self.num_queries, wide=wide)
        if wide:
            ch.set_scratch(ws["split"])
        k = 0

        def mlp3_and_q(k, hidx, nxt):
            # (the next layer's query projection does not depend on the mask-embed MLP: first, and beside its first layer)
            if nxt is not None:
                lw = W["layers"][nxt]
                ch.set_linear(k, ws["ze16"], lw["xq_w"], lw["xq_b"], ws["q16"], scale=qscale)
                if wide:
                    ch.set_parallel(k)
                k += 1
            ch.set_linear(k, ws["d16"][hidx], me[0][0], me[0][1], ws["m1"], relu=True)
            ch.set_linear(k + 1, ws["m1"], me[1][0], me[1][1], ws["m2"], relu=True)
            ch.set_linear(k + 2, ws["m2"], me[2][0], me[2][1], ws["me16"])
            return k + 3

        pre_first = k
        k = mlp3_and_q(k, 0, 0)
        pre_count = k - pre_first
        for i in range(nl):
            lw = W["layers"][i]
            first.append(k)
            ch.set_linear_ln(k, ws["att16"], lw["xo_w"], lw["xo_b"], ws["z32"], lw["ln_x"], None, W["qe"],
                             y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"])
            ch.set_linear(k + 1, ws["ze16"], lw["sqk_w"], lw["sqk_b"], ws["qk16"])
            if wide:
                ch.set_parallel(k + 1)                       # the key / query and the value projections side by side
            ch.set_linear(k + 2, ws["z16"], lw["sv_w"], lw["sv_b"], ws["v16"])
            ch.set_self_attn(k + 3, ws["qk16"], ws["v16"], ws["sa16"])
            ch.set_linear_ln(k + 4, ws["sa16"], lw["so_w"], lw["so_b"], ws["z32"], lw["ln_s"], None, W["qe"],
                             y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"])
            ch.set_linear(k + 5, ws["z16"], lw["f1_w"], lw["f1_b"], ws["h16"], relu=True)
            ch.set_linear_ln(k + 6, ws["h16"], lw["f2_w"], lw["f2_b"], ws["z32"], lw["ln_f"], W["dn"], W["qe"],
                             y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"], d32=ws["d32"], d16=ws["d16"][i + 1])
            k = mlp3_and_q(k + 7, i + 1, i + 1 if i + 1 < nl else None)
            count.append(k - first[-1])
        assert k == n
        ch.upload()
        return dict(chain=ch, pre=(pre_first, pre_count), first=first, count=count)

    def _ensure_chain(self, W, ws):
        """(re)builds the phase list when the workspace is new or the weights changed; never inside a graph capture"""
        mode = self._chain_mode(ws)
        if mode != ws.get("chain_mode", None) or (mode and ws.get("chain_W") is not W):
            # a captured layer loop belongs to the schedule (and, for the wide chain, to the barrier words) it was captured
            # with: drop it before the chain it refers to goes away
            ws["graph"] = None
            ws["chain"], ws["chain_W"], ws["chain_mode"] = (self._build_chain(W, ws) if mode else None), W, mode

    WIDE_CHAIN_MAX_ROWS = 16384          # what the wide chain can take (its split workspace)
    WIDE_CHAIN_AUTO_ROWS = 2048          # where it beats the launch-per-op schedule (profiles/experiments/chain_r2.md)

    def _chain_mode(self, ws):
        """None (launch per op), "group" (one CTA per group runs a layer's query side: single-group calls) or "wide" (the
        layer's tiles spread over all CTAs of one cooperative launch, grid barriers between the phases: many groups)."""
        wide_ok = self.num_queries <= 256 and ws["R"] <= self.WIDE_CHAIN_MAX_ROWS and ws.get("split") is not None
        uc = self.use_chain
        if uc is None:
            if self.wide_chain_default and wide_ok and ws["R"] <= self.WIDE_CHAIN_AUTO_ROWS:
                return "wide"
            return "group" if ws["G"] == 1 and self.num_queries <= 128 else None
        if uc == "wide":
            return "wide" if wide_ok else None
        return "group" if uc and self.num_queries <= 128 else None

    def _chain_on(self, ws):
        return self._chain_mode(ws) is not None

    def _layer_loop_chain(self, W, ws):
        """_layer_loop with the query side of every layer as one launch (csrc/chain.cuh)."""
        G, Tg, N = ws["G"], ws["Tg"], ws["N"]
        Q, nl = self.num_queries, self.num_layers
        self._ensure_chain(W, ws)
        c = ws["chain"]
        ch = c["chain"]
        ws["flags"].zero_()
        L.init_queries(W["qf"], W["qe"], W["dn"][0], W["dn"][1], G, (ws["z32"], ws["z16"], ws["ze16"], ws["d32"], ws["d16"][0]))

        def bits(hidx, level):
            if ws["use_t"][level]:
                L.mask_bits_t(ws["gt"][level], G, Tg * N[level], ws["me16"], Q, ws["bits_t"][level], ws["blockand"][level],
                              ws["flags"][hidx], Q)
            else:
                L.mask_bits(ws["gt"][level], G, Tg * N[level], ws["me16"], Q, ws["bits"][level], ws["flags"][hidx], Q)
            if self.debug_capture is not None:
                b = _words_from_bits_t(ws["bits_t"][level], Q) if ws["use_t"][level] else ws["bits"][level].clone()
                self.debug_capture.append((hidx, level, b, ws["flags"][hidx].clone()))

        ch.run(*c["pre"])
        bits(0, 0)
        for i in range(nl):
            l = i % 3
            if ws["use_t"][l]:
                L.xattn_t(ws["q16"], ws["k"][i], ws["v"][i], ws["bits_t"][l], ws["blockand"][l], ws["flags"][i], G, Q, Q,
                          Tg * N[l], ws["splits"][l], ws["o_part"], ws["ml_part"], ws["att16"])
            else:
                L.xattn(ws["q16"], ws["k"][i], ws["v"][i], ws["bits"][l], ws["flags"][i], G, Q, Q, Tg * N[l], ws["splits"][l],
                        ws["o_part"], ws["ml_part"], ws["att16"])
            ch.run(c["first"][i], c["count"][i])
            if i + 1 < nl:
                bits(i + 1, (i + 1) % 3)

    def _layer_loop(self, W, ws):
        """Query initialisation, the first head's mask bits and the nine decoder layers.  Everything here reads and writes
        the per-shape workspace and the weight cache only (no caller tensors, no allocation), so it can be replayed as a
        CUDA graph."""
        if self._chain_on(ws):
            return self._layer_loop_chain(W, ws)
        G, Tg, N = ws["G"], ws["Tg"], ws["N"]
        Q, C, nl = self.num_queries, HIDDEN, self.num_layers
        ws["flags"].zero_()
        L.init_queries(W["qf"], W["qe"], W["dn"][0], W["dn"][1], G, (ws["z32"], ws["z16"], ws["ze16"], ws["d32"], ws["d16"][0]))

        def head_bits(hidx, level):
            me = self._mlp3(W["mask_embed"], ws["d16"][hidx], ws["m1"], ws["m2"], ws["me16"])
            if ws["use_t"][level]:
                L.mask_bits_t(ws["gt"][level], G, Tg * N[level], me, Q, ws["bits_t"][level], ws["blockand"][level],
                              ws["flags"][hidx], Q)
            else:
                L.mask_bits(ws["gt"][level], G, Tg * N[level], me, Q, ws["bits"][level], ws["flags"][hidx], Q)
            if self.debug_capture is not None:
                bits = _words_from_bits_t(ws["bits_t"][level], Q) if ws["use_t"][level] else ws["bits"][level].clone()
                self.debug_capture.append((hidx, level, bits, ws["flags"][hidx].clone()))

        head_bits(0, 0)
        qscale = (C // NHEADS) ** -0.5 * LOG2E
        for i in range(nl):
            l = i % 3
            lw = W["layers"][i]
            # masked cross-attention (video_..._decoder.py:110-122)
            L.linear_f16(ws["ze16"], lw["xq_w"], lw["xq_b"], scale=qscale, out=ws["q16"])
            if ws["use_t"][l]:
                L.xattn_t(ws["q16"], ws["k"][i], ws["v"][i], ws["bits_t"][l], ws["blockand"][l], ws["flags"][i], G, Q, Q,
                          Tg * N[l], ws["splits"][l], ws["o_part"], ws["ml_part"], ws["att16"])
            else:
                L.xattn(ws["q16"], ws["k"][i], ws["v"][i], ws["bits"][l], ws["flags"][i], G, Q, Q, Tg * N[l], ws["splits"][l],
                        ws["o_part"], ws["ml_part"], ws["att16"])
            L.linear_ln_f16(ws["att16"], lw["xo_w"], lw["xo_b"], ws["z32"], lw["ln_x"], None, W["qe"],
                            y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"], split_ws=ws["split"])
            # self-attention (video_..._decoder.py:52-62)
            L.linear_f16(ws["ze16"], lw["sqk_w"], lw["sqk_b"], out=ws["qk16"])
            L.linear_f16(ws["z16"], lw["sv_w"], lw["sv_b"], out=ws["v16"])
            L.self_attn(ws["qk16"], ws["v16"], ws["sa16"], G, Q)
            L.linear_ln_f16(ws["sa16"], lw["so_w"], lw["so_b"], ws["z32"], lw["ln_s"], None, W["qe"],
                            y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"], split_ws=ws["split"])
            # FFN (video_..._decoder.py:175-179) + decoder_norm of the following prediction head
            L.linear_f16(ws["z16"], lw["f1_w"], lw["f1_b"], relu=True, out=ws["h16"])
            L.linear_ln_f16(ws["h16"], lw["f2_w"], lw["f2_b"], ws["z32"], lw["ln_f"], W["dn"], W["qe"],
                            y32=ws["z32"], y16=ws["z16"], ype16=ws["ze16"], d32=ws["d32"], d16=ws["d16"][i + 1],
                            split_ws=ws["split"])
            if i + 1 < nl:
                head_bits(i + 1, (i + 1) % 3)

    def _run_layers(self, W, ws):
        """The layer loop is ~130 small dependent launches: issued from Python they are launch-rate bound (2.5 ms of host
        time per call, more than the GPU needs).  After one eager call per workspace the loop is captured into a CUDA
        graph and replayed; tests' bit capture and the bench's per-launch profiling run it eagerly."""
        self._ensure_chain(W, ws)
        eager = (not self.use_cuda_graph) or self.debug_capture is not None or L.PROFILE is not None
        if eager or not ws.get("warm"):
            ws["warm"] = True
            return self._layer_loop(W, ws)
        if ws.get("graph") is None or ws.get("graph_W") is not W:
            n0 = L.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._layer_loop(W, ws)
            ws["graph"], ws["graph_W"], ws["graph_launches"] = g, W, L.launch_count() - n0
            L.add_launch_count(-ws["graph_launches"])      # capture only recorded them; replays are counted below
        ws["graph"].replay()
        L.add_launch_count(ws["graph_launches"])

    def forward_tokens(self, tokens):
        """Video decoders: the forward pass on the token-major fp16 hand-off of `pixel_decoder.MSDeformAttnPixelDecoder.
        forward_tokens` (`DecoderTokens`) instead of fp32 NCHW maps -- same layers, same outputs; the layout kernels are replaced
        by a centre-2x2 pooling of the fp16 mask features and a position add (0.7 GB of traffic per 36-frame 720 x 1280 clip
        instead of 4.9 GB).  The g_l operands are then means of fp16-rounded mask features (the NCHW path averages in fp32
        before rounding): mask bits agree to the same 99.9 % bar (tests/test_pixel_decoder_gpu.py)."""
        if self.training:
            raise RuntimeError("openvis_b200 decoders are inference-only: call .eval()")
        if not self.VIDEO or self.SAN:
            raise NotImplementedError("forward_tokens: Video decoders (the Frame / SAN variants return NCHW-derived extras)")
        BT, H4, W4 = tokens.ft.shape[0], tokens.H4, tokens.W4
        sizes = [tuple(s_) for s_ in tokens.sizes]
        if H4 % 8 or W4 % 8 or sizes != [(H4 // s_, W4 // s_) for s_ in (8, 4, 2)] or tokens.ft.shape[1:] != (H4 * W4, HIDDEN):
            raise NotImplementedError("forward_tokens: strides 32 / 16 / 8 / 4 of a /32-padded input, 256 channels")
        dev = tokens.ft.device
        with torch.no_grad(), torch.cuda.device(dev):
            return self._forward_impl(None, None, None, BT, H4, W4, sizes, dev, tokens=tokens)

    def _forward_impl(self, x, mf, mask_features_in, BT, H4, W4, sizes, dev, tokens=None):
        W = self._weights()
        ws = self._workspace(BT, H4, W4, dev)
        padd, p2, pz = self._pos_tables(BT // self._groups(BT), sizes, dev)
        if pz is not None and self._groups(BT) > 1:
            pz = pz.repeat(self._groups(BT), 1)            # frame b of the call is frame b % Tg of its clip
        self._generation += 1
        gen = self._generation
        G, Tg, R, N, M = ws["G"], ws["Tg"], ws["R"], ws["N"], ws["M"]
        Q, C, nl = self.num_queries, HIDDEN, self.num_layers

        # ---- layout preparation (HBM-bound, once per call)
        xt, ft = ws["xt"], ws["ft"]
        if tokens is not None:
            xt, ft = [t.reshape(BT * n, C) for t, n in zip(tokens.xt, N)], tokens.ft.reshape(BT * M, C)
            pk = ("tok", id(padd))
            if pk not in self._pcache:                       # token-major copies of the position tables (cached with them)
                self._pcache[pk] = [p.t().contiguous() for p in padd]
            for l, s_ in enumerate((8, 4, 2)):
                L.tokens_pool_f16(ft, BT, H4, W4, s_, ws["gt"][l])
                L.tokens_add_pos_f16(tokens.xt[l], self._pcache[pk][l], pz, ws["xp"][l])
        else:
            for l in range(3):
                if x[l].shape[-1] % 4 == 0:
                    L.nchw_to_tokens_hw_f16(x[l], out=ws["xt"][l], out_pos=ws["xp"][l], pos_cn=padd[l], pos_t=pz)
                else:                                      # odd widths: no 16-byte row pitch for the tensor maps
                    L.nchw_to_tokens_f16(x[l], out=ws["xt"][l], out_pos=ws["xp"][l], pos=padd[l].t().contiguous(), pos_t=pz)
            L.maskfeat_prep(mf, (ws["ft"], ws["gt"][0], ws["gt"][1], ws["gt"][2]))
        ws["ft_cur"] = ft                                  # the final / lazy mask heads read the mask features from here
        # ---- key / value projections of all layers, one launch per level
        for l in range(3):
            ids = W["kv_layers"][l]
            if not ids:
                continue
            outs, biases = [], []
            for i in ids:
                outs += [ws["k"][i], ws["v"][i]]
                biases += [W["layers"][i]["xk_b"], W["layers"][i]["v_bias"]]
            L.kv_proj_f16(ws["xp"][l], xt[l], W["kv_w"][l], outs, biases)

        san = self._san_prepare(W, ws, BT, H4, W4) if self.SAN else None

        self._run_layers(W, ws)

        # ---- final prediction head: full-resolution mask logits, class logits / attention biases
        out = _LazyDict()
        valid = torch.zeros(BT, Q, dtype=torch.uint8, device=dev)
        pred_masks = self._full_masks(W, ws, nl, BT, H4, W4, posflags=valid)
        # extension (not a reference key): [T, Q] "mask is non-empty" = ClipAdapter._preprocess_image's `valid`
        # (clip_adapter/adapter.py:86-88) evaluated on the stride-4 logits, produced by the mask GEMM's epilogue
        out["mask_valid"] = valid
        pred_embeds = ws["d32"].clone()
        cls = self._class_outputs(W, ws, nl, BT, san)
        self._pack_outputs(out, cls, pred_masks, pred_embeds, x, mask_features_in, sizes, p2, pz, BT, san)

        def compute_aux(j):
            if gen != self._generation:
                raise RuntimeError("aux_outputs must be read before the next forward() of the same decoder "
                                   "(set decoder.materialize_aux = True to compute them eagerly)")
            with torch.cuda.device(dev):
                m = self._full_masks(W, ws, j, BT, H4, W4)
                c = self._class_outputs(W, ws, j, BT, san)
            d = {}
            self._pack_head(d, c, m, BT)
            return d

        aux = LazyAuxOutputs(nl, compute_aux)
        if self.materialize_aux:
            list(aux)
        out["aux_outputs"] = aux
        # (weak references: remembering which tensors the operand copies belong to must not keep GBs of them alive)
        self._last = dict(gen=gen, mf=weakref.ref(mask_features_in) if mask_features_in is not None else (lambda: None),
                          mf_ver=mask_features_in._version if mask_features_in is not None else -1, ft=ws["ft_cur"],
                          af32=weakref.ref(san["attn_feats"]) if san else None,
                          af_ver=san["attn_feats"]._version if san else None, af16=ws.get("af16"))
        return out

    def shared_operands(self, mask_feats, attn_feats):
        """The token-major fp16 copies of `mask_feats` / `attn_feats` made by the most recent forward(), for a consumer
        that is handed exactly those tensors (temporal.TemporalInstanceResampler.operand_source); None when they are
        other tensors, were modified since, or the workspace has been reused by a later call."""
        last = getattr(self, "_last", None)
        if last is None or last["gen"] != self._generation or last["af16"] is None or last["af32"] is None:
            return None
        if mask_feats is not last["mf"]() or mask_feats._version != last["mf_ver"]:
            return None
        if attn_feats is not last["af32"]() or attn_feats._version != last["af_ver"]:
            return None
        return last["ft"], last["af16"]

    def _full_masks(self, W, ws, hidx, BT, H4, W4, posflags=None):
        """einsum("bqc,bchw->bqhw") of head `hidx`, written directly in the reference's eval layout [1, Q, T, H, W]
        ('(b t) q h w -> b q t h w', frame_...:117-118; video_...:459)."""
        Q = self.num_queries
        me = self._mlp3(W["mask_embed"], ws["d16"][hidx], ws["m1"], ws["m2"], ws["me16"])
        G, Tg, M = ws["G"], ws["Tg"], ws["M"]
        if self.VIDEO:
            # [clips, Q, Tg, H, W]: out[g][q][t*M + p]
            out = torch.empty(G, Q, Tg, H4, W4, dtype=torch.float32, device=me.device)
            L.mask_logits(ws.get("ft_cur", ws["ft"]), G, Tg * M, me, Q, Q, out, Q * Tg * M, Tg * M,
                          posflags=posflags, rows_per_frame=M if posflags is not None else 0)
        else:
            # frames as groups, written as [1, Q, T, H, W]: out[q][g*M + p]
            out = torch.empty(1, Q, BT, H4, W4, dtype=torch.float32, device=me.device)
            L.mask_logits(ws.get("ft_cur", ws["ft"]), G, M, me, Q, Q, out, M, BT * M,
                          posflags=posflags, rows_per_frame=M if posflags is not None else 0)
        return out

    def _class_outputs(self, W, ws, hidx, BT, san):
        if "class_embed" not in W:
            return None
        x = ws["d16"][hidx]
        params = W["class_embed"]
        for j, (w, b) in enumerate(params):
            last = j == len(params) - 1
            x = L.linear_f16(x, w, b, relu=not last, out_f32=last)
        return x                                                   # [G*Q, cls] fp32

    def _pack_head(self, d, cls, masks, BT):
        Q = self.num_queries
        if cls is not None:
            d["pred_logits"] = cls.view(-1, Q, cls.shape[-1]) if self.VIDEO else cls.view(1, BT, Q, -1)
        d["pred_masks"] = masks

    def _pack_outputs(self, out, cls, pred_masks, pred_embeds, x, mask_features, sizes, p2, pz, BT, san):
        self._pack_head(out, cls, pred_masks, BT)
        if not self.VIDEO:
            Q = self.num_queries
            le = self.level_embed.weight.detach().float()
            out["mask_feats"] = mask_features
            # (frame_...:64-69) returned for API compatibility only; built on first access with torch glue
            out["ms_feats"] = _lazy(lambda: [(x[l].flatten(2) + le[l][None, :, None]).permute(2, 0, 1) for l in range(3)])
            out["ms_pos"] = _lazy(lambda: [p2[l][:, None, :].expand(-1, BT, -1) for l in range(3)])
            out["size_list"] = [torch.Size(s) for s in sizes]
            out["pred_embeds"] = pred_embeds.view(1, BT, Q, HIDDEN)

    def _san_prepare(self, W, ws, BT, H4, W4):   # pragma: no cover - overridden
        return None


# ------------------------------------------------------------------------------------------------ public classes
def _configurable_new(cls):
    """Emulates detectron2's @configurable: Cls(cfg, in_channels, mask_classification) -> Cls(**from_config(...))."""
    orig_init = cls.__init__

    def __init__(self, *args, **kwargs):
        if args and hasattr(args[0], "MODEL"):
            kw = type(self).from_config(*args, **kwargs)
            orig_init(self, **kw)
        else:
            orig_init(self, *args, **kwargs)

    cls.__init__ = __init__
    return cls


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class VideoMultiScaleMaskedTransformerDecoder(_B200MaskedDecoderBase):
    VIDEO = True


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class FrameMultiScaleMaskedTransformerDecoder(_B200MaskedDecoderBase):
    VIDEO = False


class _EmbeddingMixin:
    """Embedding* variants: class_embed = MLP(hidden, 2*clip_dims, clip_dims, 2) (video_..._decoder.py:487-523)."""

    def __init__(self, clip_dims=None, mask_classification=True, **kwargs):
        super().__init__(mask_classification=False, **kwargs)
        self.mask_classification = mask_classification
        if self.mask_classification:
            self.class_embed = MLP(kwargs["hidden_dim"], clip_dims * 2, clip_dims, 2)

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        ret = super().from_config(cfg, in_channels, mask_classification)
        ret["clip_dims"] = cfg.MODEL.CLIP_ADAPTER.CLIP_EMBED_DIMS
        return ret


class _ProposalMixin:
    """Proposal* variants: class_embed = Linear(hidden, 2) (video_..._decoder.py:526-537)."""

    def __init__(self, mask_classification=True, **kwargs):
        super().__init__(mask_classification=False, **kwargs)
        self.mask_classification = mask_classification
        if self.mask_classification:
            self.class_embed = nn.Linear(kwargs["hidden_dim"], 1 + 1)


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class EmbeddingVideoMultiScaleMaskedTransformerDecoder(_EmbeddingMixin, _B200MaskedDecoderBase):
    VIDEO = True


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class ProposalVideoMultiScaleMaskedTransformerDecoder(_ProposalMixin, _B200MaskedDecoderBase):
    VIDEO = True


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class EmbeddingFrameMultiScaleMaskedTransformerDecoder(_EmbeddingMixin, _B200MaskedDecoderBase):
    VIDEO = False


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class ProposalFrameMultiScaleMaskedTransformerDecoder(_ProposalMixin, _B200MaskedDecoderBase):
    VIDEO = False


class _ZeroShotMixin:
    """ZeroShotMultiScaleMaskedTransformerDecoder (zero_shot_mask2former_transformer_decoder.py:44-143, 172-277): a still-image
    decoder -- every image of the batch is its own attention group and the outputs carry no frame axis.  There is no
    class_embed: `pred_logits` IS the decoder_norm embedding (the zero-shot classifier consumes it, :256-265) and a two-layer
    `object_embed` MLP gives the 2-way objectness `pred_object_logits` (:142, 249)."""

    def __init__(self, mask_classification=True, **kwargs):
        kwargs.pop("num_frames", None)
        super().__init__(mask_classification=False, **kwargs)
        self.mask_classification = mask_classification
        self.object_embed = MLP(kwargs["hidden_dim"], kwargs["hidden_dim"], 2, 2)

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        """zero_shot_...:145-170 (no num_frames)."""
        m = cfg.MODEL.MASK_FORMER
        assert m.DEC_LAYERS >= 1
        return dict(in_channels=in_channels, mask_classification=mask_classification, num_classes=cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES,
                    hidden_dim=m.HIDDEN_DIM, num_queries=m.NUM_OBJECT_QUERIES, nheads=m.NHEADS, dim_feedforward=m.DIM_FEEDFORWARD,
                    dec_layers=m.DEC_LAYERS - 1, pre_norm=m.PRE_NORM, enforce_input_project=m.ENFORCE_INPUT_PROJ,
                    mask_dim=cfg.MODEL.SEM_SEG_HEAD.MASK_DIM)

    def _full_masks(self, W, ws, hidx, BT, H4, W4, posflags=None):
        """einsum("bqc,bchw->bqhw") (:251), one group per image: out[b][q][p]."""
        Q, M = self.num_queries, ws["M"]
        me = self._mlp3(W["mask_embed"], ws["d16"][hidx], ws["m1"], ws["m2"], ws["me16"])
        out = torch.empty(BT, Q, H4, W4, dtype=torch.float32, device=me.device)
        L.mask_logits(ws.get("ft_cur", ws["ft"]), BT, M, me, Q, Q, out, Q * M, M, posflags=posflags, rows_per_frame=M if posflags is not None else 0)
        return out

    def _class_outputs(self, W, ws, hidx, BT, san):
        x = ws["d16"][hidx]
        (w0, b0), (w1, b1) = W["object_embed"]
        obj = L.linear_f16(L.linear_f16(x, w0, b0, relu=True), w1, b1, out_f32=True)
        # the embedding of the last head is available in fp32; earlier heads (aux_outputs) keep their fp16 operand copy
        emb = ws["d32"].clone() if hidx == self.num_layers else x.float()
        return obj, emb

    def _pack_head(self, d, cls, masks, BT):
        Q = self.num_queries
        if self.mask_classification:
            d["pred_object_logits"] = cls[0].view(BT, Q, 2)
            d["pred_logits"] = cls[1].view(BT, Q, HIDDEN)
        d["pred_masks"] = masks

    def _pack_outputs(self, out, cls, pred_masks, pred_embeds, x, mask_features, sizes, p2, pz, BT, san):
        out["pred_object_logits"] = cls[0].view(BT, self.num_queries, 2)
        out["pred_logits"] = cls[1].view(BT, self.num_queries, HIDDEN)
        out["pred_masks"] = pred_masks
        out["pred_embeds"] = pred_embeds.view(BT, self.num_queries, HIDDEN)


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class ZeroShotMultiScaleMaskedTransformerDecoder(_ZeroShotMixin, _B200MaskedDecoderBase):
    VIDEO = False


class _SideAdapterMixin:
    """SAN decoders (side_adapter_frame_..._decoder.py:29-55): no class_embed; `attn_embed` MLP and the `attn_mlp`
    1x1-conv branch that turns quarter-resolution mask features into per-CLIP-head attention-bias features."""
    SAN = True

    def __init__(self, clip_heads=None, mask_classification=True, **kwargs):
        super().__init__(mask_classification=False, **kwargs)
        hidden_dim = kwargs["hidden_dim"]
        self.clip_heads = clip_heads
        self.attn_embed = MLP(hidden_dim, hidden_dim, hidden_dim, 3)
        self.attn_mlp = MLP(hidden_dim, hidden_dim, hidden_dim * clip_heads, 3,
                            affine_func=lambda n, k: nn.Conv2d(n, k, kernel_size=1))

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        ret = super().from_config(cfg, in_channels, mask_classification)
        ret["clip_heads"] = cfg.MODEL.CLIP_ADAPTER.CLIP_NUM_HEADS
        return ret

    def _san_prepare(self, W, ws, BT, H4, W4):
        """attn_features = attn_mlp(bilinear 1/4 (mask_features)) (side_adapter_frame_...:67-71).  The bilinear 1/4
        map is the centre-2x2 mean at stride 4, i.e. exactly the level-1 pooled features g1 already computed."""
        nh, C = self.clip_heads, HIDDEN
        P = ws["N"][1]
        dev = ws["ft"].device
        if "af_h1" not in ws:
            ws["af_h1"] = torch.empty(BT * P, C, dtype=torch.float16, device=dev)
            ws["af_h2"] = torch.empty(BT * P, C, dtype=torch.float16, device=dev)
            ws["af16"] = torch.empty(BT * P, nh * C, dtype=torch.float16, device=dev)
            ws["ae16"] = torch.empty(ws["R"], C, dtype=torch.float16, device=dev)
        (w0, b0), (w1, b1), (w2, b2) = W["attn_mlp"]
        L.linear_f16(ws["gt"][1], w0, b0, relu=True, out=ws["af_h1"])
        L.linear_f16(ws["af_h1"], w1, b1, relu=True, out=ws["af_h2"])
        # token-major fp16 copy (operand of the bias einsum) ...
        L.linear_f16(ws["af_h2"], w2, b2, out=ws["af16"])
        # ... and the fp32 NCHW tensor the reference returns as `attn_feats` [BT, heads, C, h, w]
        attn_feats = torch.empty(BT, nh, C, H4 // 4, W4 // 4, dtype=torch.float32, device=dev)
        L.mask_logits(ws["af_h2"], BT, P, w2, 0, nh * C, attn_feats, nh * C * P, P, bias=b2)
        return dict(attn_feats=attn_feats, P=P, hw=(H4 // 4, W4 // 4))

    def _class_outputs(self, W, ws, hidx, BT, san):
        """class_attn_biases = einsum("bqc,bnchw->bnqhw") (side_adapter_frame_...:154-157) /
        "bqc,btnchw->btnqhw" (side_adapter_video_...:128)."""
        Q, nh = self.num_queries, self.clip_heads
        ae = self._mlp3(W["attn_embed"], ws["d16"][hidx], ws["m1"], ws["m2"], ws["ae16"])
        h, w = san["hw"]
        out = torch.empty(BT, nh, Q, h, w, dtype=torch.float32, device=ae.device)
        if self.VIDEO:
            # one set of queries per clip: replicate its [Q, 256] embedding block for each of its frames (tiny)
            G = ws["G"]
            ae = ae.view(G, 1, Q, HIDDEN).expand(G, BT // G, Q, HIDDEN).reshape(BT * Q, HIDDEN).contiguous()
        L.san_bias_logits(ws["af16"], BT, san["P"], nh, ae, Q, out)
        return out

    def _pack_head(self, d, cls, masks, BT):
        G = self._groups(BT) if self.VIDEO else 1
        d["class_attn_biases"] = cls.view(G, BT // G, *cls.shape[1:])    # [1, T, n, Q, h, w] (reference: b = 1)
        d["pred_masks"] = masks

    def _pack_outputs(self, out, cls, pred_masks, pred_embeds, x, mask_features, sizes, p2, pz, BT, san):
        super()._pack_outputs(out, cls, pred_masks, pred_embeds, x, mask_features, sizes, p2, pz, BT, san)
        if not self.VIDEO:
            out["attn_feats"] = san["attn_feats"]


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class SideAdapterFrameMultiScaleMaskedTransformerDecoder(_SideAdapterMixin, _B200MaskedDecoderBase):
    VIDEO = False


@TRANSFORMER_DECODER_REGISTRY.register()
@_configurable_new
class SideAdapterVideoMultiScaleMaskedTransformerDecoder(_SideAdapterMixin, _B200MaskedDecoderBase):
    VIDEO = True
