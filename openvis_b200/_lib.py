"""ctypes binding of the C-ABI library (include/openvis_b200.h) + thin torch-tensor wrappers.

There is NO fallback: if the shared library is missing or the device is not sm_100, calls raise.
PyTorch is used only for device memory (tensors) and the current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# OVIS_LIB_PATH: A/B timing of an older build of the same C ABI (tools/); never set in tests or the bench
LIB_PATH = os.environ.get("OVIS_LIB_PATH") or os.path.join(_HERE, "libopenvis_b200.so")

_c_int, _c_ll, _c_float, _vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); must list every symbol declared in include/openvis_b200.h
SIGNATURES = {
    "ovis_version": (_c_int, []),
    "ovis_last_error": (ctypes.c_char_p, []),
    "ovis_device_check": (_c_int, []),
    "ovis_launch_count": (_c_ll, []),
    "ovis_add_launch_count": (None, [_c_ll]),
    "ovis_nchw_to_tokens_f16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _vp]),
    "ovis_nchw_to_tokens_hw_f16": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_maskfeat_prep": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_cast_f16": (_c_int, [_vp, _vp, _c_ll, _vp]),
    "ovis_init_queries": (_c_int, [_vp] * 9 + [_c_int, _c_int, _vp]),
    "ovis_rownorm": (_c_int, [_vp] * 5 + [_c_int, _c_int, _c_int, _vp]),
    "ovis_rowstats": (_c_int, [_vp] * 6 + [_c_int, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_linear_rowscale_f16": (_c_int, [_vp, _c_ll, _c_int, _c_int, _vp, _c_int, _vp, _c_float, _vp, _vp, _vp, _c_int, _c_int, _vp]),
    "ovis_linear_f16": (_c_int, [_vp, _c_ll, _c_int, _c_int, _vp, _c_int, _vp, _c_float, _c_int, _vp, _c_int, _c_int, _vp]),
    "ovis_linear_act_f16": (_c_int, [_vp, _c_ll, _c_int, _c_int, _vp, _c_int, _vp, ctypes.c_float, _c_int, _vp, _vp, _c_int,
                                     _c_int, _vp]),
    "ovis_linear_ln_f16": (_c_int, [_vp, _c_ll, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int,
                                    _vp, _vp, _vp, _vp, _vp, _vp, _c_ll, _vp]),
    "ovis_kv_proj_f16": (_c_int, [_vp, _vp, _c_ll, _vp, _c_int, _vp, _vp, _vp]),
    "ovis_mask_bits": (_c_int, [_vp, _c_int, _c_int, _vp, _c_int, _vp, _vp, _c_int, _vp]),
    "ovis_mask_logits": (_c_int, [_vp, _c_int, _c_int, _vp, _c_int, _c_int, _vp, _vp, _c_ll, _c_ll, _vp, _c_int, _vp]),
    "ovis_san_bias_logits": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _c_int, _vp, _vp]),
    "ovis_xattn_plan": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int),
                                 ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll)]),
    "ovis_xattn": (_c_int, [_vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp]),
    "ovis_xattn_plan_t": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int), ctypes.POINTER(_c_int),
                                   ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll)]),
    "ovis_mask_bits_t": (_c_int, [_vp, _c_int, _c_int, _vp, _c_int, _vp, _vp, _vp, _c_int, _vp]),
    "ovis_xattn_t": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp]),
    "ovis_chain_create": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(_vp)]),
    "ovis_chain_set_wide": (_c_int, [_vp, _c_int]),
    "ovis_chain_set_scratch": (_c_int, [_vp, _vp, _c_ll]),
    "ovis_chain_set_parallel": (_c_int, [_vp, _c_int, _c_int]),
    "ovis_chain_set_linear": (_c_int, [_vp, _c_int, _vp, _c_int, _c_int, _vp, _c_int, _vp, _c_float, _c_int, _vp, _c_int, _c_int]),
    "ovis_chain_set_linear_ln": (_c_int, [_vp, _c_int, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int, _vp, _vp, _vp, _vp, _vp]),
    "ovis_chain_set_self_attn": (_c_int, [_vp, _c_int, _vp, _vp, _vp]),
    "ovis_chain_upload": (_c_int, [_vp]),
    "ovis_chain_run": (_c_int, [_vp, _c_int, _c_int, _vp]),
    "ovis_chain_run_traced": (_c_int, [_vp, _c_int, _c_int, _vp, _vp]),
    "ovis_chain_destroy": (_c_int, [_vp]),
    "ovis_self_attn": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _vp]),
    "ovis_clip_aggregate": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _vp]),
    "ovis_mask_boxes": (_c_int, [_vp, _c_int, _c_int, _c_ll, _c_ll, _c_int, _c_int, _c_int, _c_float, _vp, _vp, _vp]),
    "ovis_crop_blend": (_c_int, [_vp, _vp, _c_ll, _c_ll, _c_int, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp]),
    "ovis_clip_patchify": (_c_int, [_vp, _c_ll, _c_int, _c_int, _vp, _vp, _vp, _vp]),
    "ovis_clip_embed": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_ll, _c_int, _c_int, _vp]),
    "ovis_ms_deform_attn_forward": (_c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                             _c_int, _vp]),
    "ovis_msda_prepare": (_c_int, [_vp, _vp, _vp, _c_ll, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "ovis_msda_fused_f16": (_c_int, [_vp, _vp, _vp, _c_ll, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_topk_scores": (_c_int, [_vp, _c_int, _c_int, _c_int, _vp, _vp, _vp, _vp, _vp]),
    "ovis_mask_postprocess": (_c_int, [_vp, _c_ll, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                       _c_int, _c_int, _vp, _vp]),
    "ovis_san_pool_bias": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_san_attn": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_san_attn_bias": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_temporal_unfold_f16": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_match_embeds": (_c_int, [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp]),
    "ovis_match_compose": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _vp]),
    "ovis_reorder_queries_f32": (_c_int, [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_ll, _c_ll, _c_ll, _c_ll, _vp]),
    "ovis_gn_stats": (_c_int, [_vp, _vp, _c_int, _c_int, _vp]),
    "ovis_gn_apply": (_c_int, [_vp, _vp, _vp, _vp, _c_float, _c_int, _c_int, _c_int, _c_int, _vp, _c_ll, _c_ll, _c_ll, _c_int, _c_int,
                               _vp, _vp, _c_ll, _c_ll, _vp]),
    "ovis_tokens_to_nchw_f32": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_ll, _c_ll, _vp]),
    "ovis_conv3x3_unfold_f16": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_tokens_pool_f16": (_c_int, [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp]),
    "ovis_tokens_add_pos_f16": (_c_int, [_vp, _vp, _vp, _vp, _c_int, _c_int, _vp]),
}

_lib = None


def load():
    """Loads libopenvis_b200.so (built in-tree by `python -m openvis_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension must be built (python -m openvis_b200.build); "
            "openvis_b200 has no CPU / PyTorch fallback path")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("OVIS_LIB_PATH") and not hasattr(lib, name):
            continue                     # older A/B build: entry points added since are simply absent
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class OvisError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise OvisError(f"openvis_b200 error {rc}: {_lib.ovis_last_error().decode()}")


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise OvisError(f"{name}: expected a CUDA tensor (no CPU path)")
    if t.dtype != dtype:
        raise OvisError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise OvisError(f"{name}: expected a contiguous tensor")


PROFILE = None   # bench.py: set to a list to receive (family, start_event, end_event) for every wrapped call


def _timed(family):
    def deco(fn):
        def wrapper(*a, **k):
            prof = PROFILE
            if prof is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            prof.append((family, e0, e1))
            return r
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


def launch_count():
    return load().ovis_launch_count()


def add_launch_count(n):
    load().ovis_add_launch_count(int(n))


def device_check():
    _check(load().ovis_device_check())


# ------------------------------------------------------------------------------------------ wrappers
@_timed("prep")
def nchw_to_tokens_f16(x, out=None, out_pos=None, pos=None, pos_t=None):
    """[B, C, h, w] fp32 -> [B, h*w, C] fp16 (and optionally out_pos = fp16(x + pos[n] + pos_t[b]))."""
    lib = load()
    _req(x, torch.float32, "x")
    B, C = x.shape[:2]
    N = x[0, 0].numel()
    if out is None:
        out = torch.empty(B, N, C, dtype=torch.float16, device=x.device)
    _check(lib.ovis_nchw_to_tokens_f16(_p(x), _p(out), _p(out_pos), _p(pos), _p(pos_t), B, C, N, _stream()))
    return out


@_timed("prep")
def nchw_to_tokens_hw_f16(x, out=None, out_pos=None, pos_cn=None, pos_t=None):
    """TMA-fed variant of nchw_to_tokens_f16; pos_cn is the position table in channel-major layout [C, h*w]."""
    lib = load()
    _req(x, torch.float32, "x")
    B, C, h, w = x.shape
    if out is None:
        out = torch.empty(B, h * w, C, dtype=torch.float16, device=x.device)
    _check(lib.ovis_nchw_to_tokens_hw_f16(_p(x), _p(out), _p(out_pos), _p(pos_cn), _p(pos_t), B, C, h, w, _stream()))
    return out


@_timed("prep")
def maskfeat_prep(F, outs=None):
    """mask_features [B, C, H, W] fp32 -> (ft [B,HW,C], g0 [B,HW/64,C], g1 [B,HW/16,C], g2 [B,HW/4,C]) fp16."""
    lib = load()
    _req(F, torch.float32, "mask_features")
    B, C, H, W = F.shape
    if outs is None:
        mk = lambda n: torch.empty(B, n, C, dtype=torch.float16, device=F.device)
        outs = (mk(H * W), mk((H // 8) * (W // 8)), mk((H // 4) * (W // 4)), mk((H // 2) * (W // 2)))
    ft, g0, g1, g2 = outs
    _check(lib.ovis_maskfeat_prep(_p(F), _p(ft), _p(g0), _p(g1), _p(g2), B, C, H, W, _stream()))
    return outs


def cast_f16(x, out=None):
    lib = load()
    _req(x, torch.float32, "x")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float16, device=x.device)
    _check(lib.ovis_cast_f16(_p(x), _p(out), x.numel(), _stream()))
    return out


def init_queries(query_feat, query_embed, dn_g, dn_b, G, outs):
    lib = load()
    Q = query_feat.shape[0]
    z32, z16, ze16, d32, d16 = outs
    _check(lib.ovis_init_queries(_p(query_feat), _p(query_embed), _p(dn_g), _p(dn_b), _p(z32), _p(z16), _p(ze16),
                                 _p(d32), _p(d16), Q, G * Q, _stream()))


def rownorm(x, g=None, b=None, layer_norm=False, l2=False, want32=True, want16=True, out16=None):
    lib = load()
    _req(x, torch.float32, "x")
    rows, D = x.shape
    o32 = torch.empty_like(x) if want32 else None
    o16 = out16 if out16 is not None else (torch.empty(rows, D, dtype=torch.float16, device=x.device) if want16 else None)
    mode = (1 if layer_norm else 0) | (2 if l2 else 0)
    _check(lib.ovis_rownorm(_p(x), _p(g), _p(b), _p(o32), _p(o16), rows, D, mode, _stream()))
    return o32, o16


@_timed("ov_tail")
def rowstats(x, g=None, b=None, layer_norm=False, want_ss=True, zero=None, groups=None):
    """x [rows, D] fp32 -> (fp16 copy (after LayerNorm when asked), ss [rows] = its rows' sums of squares or None);
    `zero` [rows] fp32 is cleared (accumulator for linear_rowscale_f16(row_ss_out=...)).
    groups = (n_groups, group_rows, group_stride): take the first group_rows rows of each of n_groups blocks of group_stride rows."""
    lib = load()
    _req(x, torch.float32, "x")
    D = x.shape[1]
    rows, gr, gs = (x.shape[0], 0, 0) if groups is None else (groups[0] * groups[1], groups[1], groups[2])
    o16 = torch.empty(rows, D, dtype=torch.float16, device=x.device)
    ss = torch.empty(rows, dtype=torch.float32, device=x.device) if want_ss else None
    _check(lib.ovis_rowstats(_p(x), _p(g), _p(b), _p(o16), _p(ss), _p(zero), rows, D, int(layer_norm), gr, gs, _stream()))
    return o16, ss


@_timed("ov_tail")
def linear_rowscale_f16(x, w, bias=None, scale=1.0, row_ss_in=None, row_ss_out=None, out=None, out_f32=False):
    """(x @ w^T + bias) * scale, rows divided by sqrt(row_ss_in) when given; row_ss_out += output rows' sums of squares."""
    lib = load()
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.stride(-1) == 1 and w.is_contiguous()
    rows, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(rows, N, dtype=torch.float32 if out_f32 else torch.float16, device=x.device)
    _check(lib.ovis_linear_rowscale_f16(_p(x), rows, K, x.stride(0), _p(w), N, _p(bias), float(scale), _p(row_ss_in),
                                        _p(row_ss_out), _p(out), out.stride(0), int(out_f32), _stream()))
    return out


@_timed("query_side")
def linear_f16(x, w, bias=None, scale=1.0, relu=False, out=None, out_f32=False):
    """x [rows, K] fp16 (row stride may exceed K), w [N, K] fp16 -> [rows, N] fp16 / fp32."""
    lib = load()
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.stride(-1) == 1 and w.is_contiguous()
    rows, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(rows, N, dtype=torch.float32 if out_f32 else torch.float16, device=x.device)
    _check(lib.ovis_linear_f16(_p(x), rows, K, x.stride(0), _p(w), N, _p(bias), float(scale), int(relu), _p(out),
                               out.stride(0), int(out_f32), _stream()))
    return out


def linear_act_f16(x, w, bias=None, scale=1.0, act=0, resid=None, out=None, out_f32=False):
    """out = act((x @ w^T + bias) * scale) + resid; act 0 none / 1 ReLU / 2 QuickGELU; resid fp32 [rows, N] (may be `out`)."""
    lib = load()
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.stride(-1) == 1 and w.is_contiguous()
    rows, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(rows, N, dtype=torch.float32 if out_f32 else torch.float16, device=x.device)
    if resid is not None:
        _req(resid, torch.float32, "resid")
        assert resid.shape == (rows, N) and resid.stride(0) == out.stride(0)
    _check(lib.ovis_linear_act_f16(_p(x), rows, K, x.stride(0), _p(w), N, _p(bias), float(scale), int(act), _p(resid),
                                   _p(out), out.stride(0), int(out_f32), _stream()))
    return out


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """value [N, S, M, D] fp32, spatial_shapes [L, 2] int64, level_start_index [L] int64, sampling_locations
    [N, Lq, M, L, P, 2], attention_weights [N, Lq, M, L, P] -> [N, Lq, M*D] fp32."""
    lib = load()
    for t, n in ((value, "value"), (sampling_locations, "sampling_locations"), (attention_weights, "attention_weights")):
        _req(t, torch.float32, n)
    _req(spatial_shapes, torch.int64, "spatial_shapes")
    _req(level_start_index, torch.int64, "level_start_index")
    N, S, M, D = value.shape
    _, Lq, M2, L_, P, two = sampling_locations.shape
    assert M2 == M and two == 2 and attention_weights.shape == (N, Lq, M, L_, P) and spatial_shapes.shape == (L_, 2)
    out = torch.empty(N, Lq, M * D, dtype=torch.float32, device=value.device)
    _check(lib.ovis_ms_deform_attn_forward(_p(value), _p(spatial_shapes), _p(level_start_index), _p(sampling_locations),
                                           _p(attention_weights), _p(out), N, S, M, D, Lq, L_, P, _stream()))
    return out


def msda_prepare(proj, reference_points, spatial_shapes, M, L_, P):
    """proj [N, Lq, M*L*P*3] fp32 (offsets | attention logits), reference_points [N, Lq, L, 2 or 4] fp32 ->
    (sampling_locations [N, Lq, M, L, P, 2], attention_weights [N, Lq, M, L, P])."""
    lib = load()
    _req(proj, torch.float32, "proj")
    _req(reference_points, torch.float32, "reference_points")
    _req(spatial_shapes, torch.int64, "spatial_shapes")
    N, Lq, cols = proj.shape
    assert cols == M * L_ * P * 3 and reference_points.shape[:3] == (N, Lq, L_)
    loc = torch.empty(N, Lq, M, L_, P, 2, dtype=torch.float32, device=proj.device)
    w = torch.empty(N, Lq, M, L_, P, dtype=torch.float32, device=proj.device)
    _check(lib.ovis_msda_prepare(_p(proj), _p(reference_points), _p(spatial_shapes), N * Lq, M, L_, P, reference_points.shape[-1],
                                 _p(loc), _p(w), _stream()))
    return loc, w


def msda_fused_f16(value16, proj, reference_points, spatial_shapes, level_start_index, N, S, Lq, M, L_, P):
    """value16 [N*S, M*32] fp16, proj [N*Lq, 3*M*L*P] fp32, reference_points [N or 1, Lq, L, 2 or 4] fp32 -> [N*Lq, M*32] fp16."""
    lib = load()
    assert value16.dtype == torch.float16 and value16.is_contiguous() and value16.shape == (N * S, M * 32)
    _req(proj, torch.float32, "proj")
    _req(reference_points, torch.float32, "reference_points")
    _req(spatial_shapes, torch.int64, "spatial_shapes")
    _req(level_start_index, torch.int64, "level_start_index")
    assert proj.shape == (N * Lq, 3 * M * L_ * P) and reference_points.shape[1:3] == (Lq, L_) and reference_points.shape[0] in (1, N)
    rd = reference_points.shape[-1]
    out = torch.empty(N * Lq, M * 32, dtype=torch.float16, device=proj.device)
    _check(lib.ovis_msda_fused_f16(_p(value16), _p(proj), _p(reference_points), 0 if reference_points.shape[0] == 1 else Lq * L_ * rd,
                                   _p(spatial_shapes), _p(level_start_index), _p(out), N, S, M, Lq, L_, P, rd, _stream()))
    return out


@_timed("postproc")
def topk_scores(scores, k=10):
    """scores [Q, K] fp32 -> (top scores [k], query index [k], label [k], entropy [k]) sorted by score."""
    lib = load()
    _req(scores, torch.float32, "scores")
    Q, K = scores.shape
    dev = scores.device
    vs, qi = torch.empty(k, device=dev), torch.empty(k, dtype=torch.int32, device=dev)
    lb, en = torch.empty(k, dtype=torch.int32, device=dev), torch.empty(k, device=dev)
    _check(lib.ovis_topk_scores(_p(scores), Q, K, k, _p(vs), _p(qi), _p(lb), _p(en), _stream()))
    return vs, qi, lb, en


@_timed("postproc")
def mask_postprocess(masks, query, pad_hw, img_hw, out_hw, out=None):
    """masks [Q, T, h4, w4] fp32 stride-4 logits (a frame slice of a larger [Q, T', h4, w4] tensor is fine),
    query [n] int32 -> bits [n, T, out_h, ceil(out_w/32)] int32."""
    lib = load()
    assert masks.dtype == torch.float32 and masks.is_cuda and masks.stride(-1) == 1, "masks: fp32 CUDA tensor"
    Q, T, h4, w4 = masks.shape
    assert masks.numel() > 0 and masks.stride(2) == w4 and masks.stride(1) == h4 * w4, "frames of a query must be contiguous"
    assert query.dtype == torch.int32 and query.is_cuda
    n = query.numel()
    if out is None:
        out = torch.empty(n, T, out_hw[0], (out_hw[1] + 31) // 32, dtype=torch.int32, device=masks.device)
    _check(lib.ovis_mask_postprocess(_p(masks), masks.stride(0), _p(query), n, T, h4, w4, pad_hw[0], pad_hw[1],
                                     img_hw[0], img_hw[1], out_hw[0], out_hw[1], _p(out), _stream()))
    return out


def san_pool_bias(bias, grid_hw):
    """bias [B, n, Q, h, w] fp32 -> pooled [B*n, Q, gh*gw] (adaptive max-pool to the CLIP grid)."""
    lib = load()
    _req(bias, torch.float32, "bias")
    B, n, Q, h, w = bias.shape
    gh, gw = grid_hw
    out = torch.empty(B * n, Q, gh * gw, dtype=torch.float32, device=bias.device)
    _check(lib.ovis_san_pool_bias(_p(bias), _p(out), B * n, Q, h, w, gh, gw, _stream()))
    return out


def san_attn(qkv, pooled, out, B, Q, L, heads=12):
    lib = load()
    assert qkv.dtype == torch.float16 and qkv.is_contiguous() and out.dtype == torch.float16 and out.is_contiguous()
    _check(lib.ovis_san_attn(_p(qkv), _p(pooled), _p(out), B, Q, L, heads, _stream()))
    return out


@_timed("query_side")
def linear_ln_f16(x, w, bias, resid, ln1, ln2=None, pe=None, y32=None, y16=None, ype16=None, d32=None, d16=None,
                  split_ws=None):
    """split_ws: optional fp32 scratch (>= K/256 * ceil128(rows) * 256 elements) enabling the few-rows split path."""
    lib = load()
    rows, K = x.shape
    _check(lib.ovis_linear_ln_f16(_p(x), rows, K, _p(w), _p(bias), _p(resid), _p(ln1[0]), _p(ln1[1]),
                                  _p(ln2[0]) if ln2 else None, _p(ln2[1]) if ln2 else None,
                                  _p(pe), pe.shape[0] if pe is not None else 0,
                                  _p(y32), _p(y16), _p(ype16), _p(d32), _p(d16),
                                  _p(split_ws), split_ws.numel() if split_ws is not None else 0, _stream()))


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = _p(t)
    return arr


@_timed("kv_proj")
def kv_proj_f16(xk, xv, w, outs, biases=None):
    """xk, xv [rows, 256] fp16 (key / value operands); w [n_tiles*256, 256] fp16; outs: n_tiles tensors [rows, 256] fp16
    (even tiles read xk, odd tiles read xv)."""
    lib = load()
    n_tiles = len(outs)
    _check(lib.ovis_kv_proj_f16(_p(xk), _p(xv), xk.shape[0], _p(w), n_tiles, _ptr_array(outs),
                                _ptr_array(biases or [None] * n_tiles), _stream()))


@_timed("mask_bits")
def mask_bits(gt, groups, rows_per_group, me, Q, bits, flags, q_stride):
    lib = load()
    _check(lib.ovis_mask_bits(_p(gt), groups, rows_per_group, _p(me), Q, _p(bits), _p(flags), q_stride, _stream()))


@_timed("mask_logits")
def mask_logits(ft, groups, rows_per_group, me, me_group_stride, Q, out, t_group_stride, ldt, bias=None,
                posflags=None, rows_per_frame=0):
    lib = load()
    _check(lib.ovis_mask_logits(_p(ft), groups, rows_per_group, _p(me), me_group_stride, Q, _p(bias), _p(out),
                                t_group_stride, ldt, _p(posflags), rows_per_frame, _stream()))


@_timed("san_bias")
def san_bias_logits(af, B, P, heads, ae, Q, out):
    lib = load()
    _check(lib.ovis_san_bias_logits(_p(af), B, P, heads, _p(ae), Q, _p(out), _stream()))


def xattn_plan(G, Q, keys):
    lib = load()
    s, qp, o, ml = _c_int(), _c_int(), _c_ll(), _c_ll()
    _check(lib.ovis_xattn_plan(G, Q, keys, ctypes.byref(s), ctypes.byref(qp), ctypes.byref(o), ctypes.byref(ml)))
    return s.value, qp.value, o.value, ml.value


@_timed("xattn")
def xattn(q, k, v, bits, flags, G, Q, q_stride, keys, splits, o_part, ml_part, out):
    lib = load()
    _check(lib.ovis_xattn(_p(q), _p(k), _p(v), _p(bits), _p(flags), G, Q, q_stride, keys, splits, _p(o_part),
                          _p(ml_part), _p(out), _stream()))


class Chain:
    """Query-side chain (csrc/chain.cuh): phases set once on fixed operands, then runs of consecutive phases are single
    launches (one persistent CTA per group of Q <= 128 query rows)."""

    def __init__(self, nphases, G, Q, wide=False):
        """wide: phases spread over all CTAs of one cooperative launch with grid barriers between them (many groups)."""
        self._lib = load()
        h = _vp()
        _check(self._lib.ovis_chain_create(int(nphases), int(G), int(Q), ctypes.byref(h)))
        self._h = h
        self._keep = []              # operands must stay where they are for the life of the chain
        if wide:
            _check(self._lib.ovis_chain_set_wide(h, 1))

    def set_scratch(self, ws):
        _req(ws, torch.float32, "scratch")
        self._keep.append(ws)
        _check(self._lib.ovis_chain_set_scratch(self._h, _p(ws), ws.numel()))

    def set_linear(self, idx, x, w, bias, out, scale=1.0, relu=False, out_f32=False):
        assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.stride(-1) == 1 and w.is_contiguous()
        self._keep += [x, w, bias, out]
        _check(self._lib.ovis_chain_set_linear(self._h, idx, _p(x), x.shape[1], x.stride(0), _p(w), w.shape[0], _p(bias), float(scale),
                                               int(relu), _p(out), out.stride(0), int(out_f32)))

    def set_linear_ln(self, idx, x, w, bias, resid, ln1, ln2=None, pe=None, y32=None, y16=None, ype16=None, d32=None, d16=None):
        self._keep += [x, w, bias, resid, ln1, ln2, pe, y32, y16, ype16, d32, d16]
        _check(self._lib.ovis_chain_set_linear_ln(self._h, idx, _p(x), x.shape[1], _p(w), _p(bias), _p(resid), _p(ln1[0]), _p(ln1[1]),
                                                  _p(ln2[0]) if ln2 else None, _p(ln2[1]) if ln2 else None, _p(pe),
                                                  pe.shape[0] if pe is not None else 0, _p(y32), _p(y16), _p(ype16), _p(d32), _p(d16)))

    def set_parallel(self, idx, parallel=True):
        """wide chains: phase idx (a plain GEMM, already set) runs beside phase idx + 1 (no barrier between them)"""
        _check(self._lib.ovis_chain_set_parallel(self._h, int(idx), int(parallel)))

    def set_self_attn(self, idx, qk, v, out):
        self._keep += [qk, v, out]
        _check(self._lib.ovis_chain_set_self_attn(self._h, idx, _p(qk), _p(v), _p(out)))

    def upload(self):
        _check(self._lib.ovis_chain_upload(self._h))

    def run(self, first, count):
        prof = PROFILE
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _check(self._lib.ovis_chain_run(self._h, int(first), int(count), _stream()))
        if prof is not None:
            e1.record()
            prof.append(("query_side", e0, e1))

    def run_traced(self, first, count):
        """profiling: SM cycles at the start of each phase and at the end, for group 0's CTA -> int64 tensor (first count + 1 entries)"""
        tr = torch.zeros(66 + 8000, dtype=torch.int64, device="cuda")     # [count + 1] phase starts; [64] event counter, events
        _check(self._lib.ovis_chain_run_traced(self._h, int(first), int(count), _p(tr), _stream()))
        return tr

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ovis_chain_destroy(self._h)
                self._h = None
        except Exception:
            pass


def xattn_kernel_name():
    return "xattn_tc3_kernel (transposed scores) / xattn_tc2_kernel + xattn_combine_kernel"


def xattn_plan_t(G, Q, keys):
    """-> (use_t, splits, q_pad, o_part floats, ml_part floats) for the transposed-score kernel (xattn_tc3)."""
    lib = load()
    u, s, qp, o, ml = _c_int(), _c_int(), _c_int(), _c_ll(), _c_ll()
    _check(lib.ovis_xattn_plan_t(G, Q, keys, ctypes.byref(u), ctypes.byref(s), ctypes.byref(qp), ctypes.byref(o), ctypes.byref(ml)))
    return bool(u.value), s.value, qp.value, o.value, ml.value


@_timed("mask_bits")
def mask_bits_t(gt, groups, rows_per_group, me, Q, bits_t, blockand, flags, q_stride):
    """Key-major mask bits for xattn_t: bits_t [G, keys, qw] int32, blockand [G, ceil(keys/32), qw] int32."""
    lib = load()
    _check(lib.ovis_mask_bits_t(_p(gt), groups, rows_per_group, _p(me), Q, _p(bits_t), _p(blockand), _p(flags), q_stride, _stream()))


@_timed("xattn")
def xattn_t(q, k, v, bits_t, blockand, flags, G, Q, q_stride, keys, splits, o_part, ml_part, out, stats=None):
    lib = load()
    _check(lib.ovis_xattn_t(_p(q), _p(k), _p(v), _p(bits_t), _p(blockand), _p(flags), G, Q, q_stride, keys, splits,
                            _p(o_part), _p(ml_part), _p(out), _p(stats), _stream()))


@_timed("query_side")
def self_attn(qk, v, out, G, Q):
    lib = load()
    _check(lib.ovis_self_attn(_p(qk), _p(v), _p(out), G, Q, _stream()))


@_timed("ov_tail")
def clip_aggregate(logits, valid):
    """logits [T, Q, K] fp32, valid [T, Q] bool/uint8 -> (probs [Q, K], qvalid [Q] bool)."""
    lib = load()
    _req(logits, torch.float32, "logits")
    T, Q, K = logits.shape
    valid = valid.to(torch.uint8).contiguous()
    probs = torch.empty(Q, K, dtype=torch.float32, device=logits.device)
    qv = torch.empty(Q, dtype=torch.uint8, device=logits.device)
    _check(lib.ovis_clip_aggregate(_p(logits), _p(valid), _p(probs), _p(qv), T, Q, K, _stream()))
    return probs, qv.bool()


def _mask_view(masks, layout):
    """(T, N, H, W, stride_t, stride_n) of a mask tensor in layout "tn" ([T, N, H, W]) or "nt" ([N, T, H, W]); the two
    leading dimensions may be strided views (a slice of frames), the H x W planes must be dense."""
    if not masks.is_cuda:
        raise OvisError("masks: expected a CUDA tensor (no CPU path)")
    if masks.dtype != torch.float32 or masks.dim() != 4:
        raise OvisError(f"masks: expected a 4-D float32 tensor, got {masks.dtype} {tuple(masks.shape)}")
    a, b, H, W = masks.shape
    if masks.stride(3) != 1 or masks.stride(2) != W:
        raise OvisError("masks: the H x W planes must be contiguous")
    if layout == "tn":
        return a, b, H, W, masks.stride(0), masks.stride(1)
    if layout == "nt":
        return b, a, H, W, masks.stride(1), masks.stride(0)
    raise OvisError(f"unknown mask layout {layout!r}")


def mask_boxes(masks, thresh=0.5, layout="tn", logits=False):
    """masks fp32: [T, N, H, W] soft masks (layout "tn") or [N, T, H, W] (layout "nt"); logits=True applies the sigmoid on load.
    Returns (valid [T, N] bool, boxes [T, N, 4] int32 = x_min, y_min, x_max + 1, y_max + 1)."""
    lib = load()
    T, N, H, W, st, sn = _mask_view(masks, layout)
    boxes = torch.empty(T * N, 4, dtype=torch.int32, device=masks.device)
    valid = torch.empty(T * N, dtype=torch.uint8, device=masks.device)
    _check(lib.ovis_mask_boxes(_p(masks), T, N, st, sn, H, W, int(logits), float(thresh), _p(boxes), _p(valid), _stream()))
    return valid.bool().view(T, N), boxes.view(T, N, 4)


def crop_blend(frames, masks, ids, boxes, R, layout="tn", logits=False):
    """frames [T, 3, H, W] fp32, masks as in mask_boxes, ids [M, 2] int32 (frame, query), boxes [T, N, 4] int32
    -> regions [M, 3, R, R] fp16."""
    lib = load()
    _req(frames, torch.float32, "frames")
    _req(ids, torch.int32, "ids")
    _req(boxes, torch.int32, "boxes")
    T, N, H, W, st, sn = _mask_view(masks, layout)
    if tuple(frames.shape) != (T, 3, H, W):
        raise OvisError(f"frames {tuple(frames.shape)} do not match masks {tuple(masks.shape)} ({layout})")
    M = ids.shape[0]
    regions = torch.empty(M, 3, R, R, dtype=torch.float16, device=frames.device)
    for m0 in range(0, M, 65535):
        m1 = min(M, m0 + 65535)
        _check(lib.ovis_crop_blend(_p(frames), _p(masks), st, sn, int(logits), _p(ids[m0:m1]), _p(boxes), m1 - m0, T, N, H, W, R,
                                   _p(regions[m0:m1]), _stream()))
    return regions


def clip_patchify(regions, patch, mean, std):
    """regions [M, 3, R, R] fp16 (0..255) -> [M * (R/patch)^2, 3 * patch^2] fp16 normalised patch rows."""
    lib = load()
    _req(regions, torch.float16, "regions")
    M, _, R, _ = regions.shape
    out = torch.empty(M * (R // patch) ** 2, 3 * patch * patch, dtype=torch.float16, device=regions.device)
    mean_c = (ctypes.c_float * 3)(*[float(v) for v in mean])
    std_c = (ctypes.c_float * 3)(*[float(v) for v in std])
    _check(lib.ovis_clip_patchify(_p(regions), M, R, patch, ctypes.cast(mean_c, _vp), ctypes.cast(std_c, _vp), _p(out), _stream()))
    return out


def clip_embed(patch_tokens, cls, pos, ln_g, ln_b, M, Lp):
    """patch_tokens [M * Lp, W] fp32 -> x [M * (1 + Lp), W] fp32 = ln_pre([cls | tokens] + pos)."""
    lib = load()
    _req(patch_tokens, torch.float32, "patch_tokens")
    Wd = patch_tokens.shape[1]
    x = torch.empty(M * (1 + Lp), Wd, dtype=torch.float32, device=patch_tokens.device)
    _check(lib.ovis_clip_embed(_p(patch_tokens), _p(cls), _p(pos), _p(ln_g), _p(ln_b), _p(x), M, Lp, Wd, _stream()))
    return x


def san_attn_bias(bias, grid_hw):
    """bias [B, n, Q, h, w] fp32 -> [B*n, Q+1+L, Q+1+L] fp32."""
    lib = load()
    _req(bias, torch.float32, "bias")
    B, n, Q, h, w = bias.shape
    gh, gw = grid_hw
    L = gh * gw
    out = torch.empty(B * n, Q + 1 + L, Q + 1 + L, dtype=torch.float32, device=bias.device)
    _check(lib.ovis_san_attn_bias(_p(bias), _p(out), B * n, Q, h, w, gh, gw, _stream()))
    return out


# ---- temporal association ------------------------------------------------------------
def temporal_unfold_f16(x, taps, out=None):
    """x [G, T, C] fp16 -> [G, T, taps*C] fp16, replicate padding along T (Conv1d padding='same' as a GEMM operand)."""
    lib = load()
    _req(x, torch.float16, "x")
    G, T, C = x.shape
    if out is None:
        out = torch.empty(G, T, taps * C, dtype=torch.float16, device=x.device)
    _check(lib.ovis_temporal_unfold_f16(_p(x), _p(out), G, T, C, int(taps), _stream()))
    return out


def match_embeds(en, want_cost=False):
    """en [B, T, n, C] fp32 L2-normalised -> pi [B, T, n] int32 (raw frame-to-previous-frame assignments) and the cost
    matrices [B, T, n, n] (None unless requested or needed as scratch)."""
    lib = load()
    _req(en, torch.float32, "en")
    B, T, n, C = en.shape
    pi = torch.empty(B, T, n, dtype=torch.int32, device=en.device)
    cost = torch.empty(B, T, n, n, dtype=torch.float32, device=en.device) if (want_cost or n > 200) else None
    _check(lib.ovis_match_embeds(_p(en), B, T, n, C, _p(cost), _p(pi), _stream()))
    return pi, cost


def match_compose(pi):
    lib = load()
    _req(pi, torch.int32, "pi")
    B, T, n = pi.shape
    idx = torch.empty(B, T, n, dtype=torch.int64, device=pi.device)
    _check(lib.ovis_match_compose(_p(pi), _p(idx), B, T, n, _stream()))
    return idx


def reorder_queries(x, idx, layout="btq"):
    """out[b, t, q] = x[b, t, idx[b, t, q]] for x [b, t, q, ...] (layout "btq"), x [b, q, t, ...] ("bqt") or
    x [q, b, t, ...] ("qbt": the Frame decoders' pred_masks [Q, (b t), h, w] of a multi-clip call); same layout out."""
    lib = load()
    _req(x, torch.float32, "x")
    _req(idx, torch.int64, "idx")
    B, T, n = idx.shape
    inner = 1
    for d in x.shape[3:]:
        inner *= d
    if layout == "btq":
        assert tuple(x.shape[:3]) == (B, T, n)
        sb, st, sq = T * n * inner, n * inner, inner
    elif layout == "bqt":
        assert tuple(x.shape[:3]) == (B, n, T)
        sb, st, sq = n * T * inner, inner, T * inner
    else:
        assert layout == "qbt" and tuple(x.shape[:3]) == (n, B, T)
        sb, st, sq = T * inner, inner, B * T * inner
    out = torch.empty_like(x)
    _check(lib.ovis_reorder_queries_f32(_p(x), _p(idx), _p(out), B, T, n, inner, sb, st, sq, _stream()))
    return out


# ------------------------------------------------------------------------------------------ pixel-decoder glue (row f-2)
def group_norm_tokens(x, B, H, W, gamma, beta, eps=1e-5, relu=False, add=None, add_layout=None, out32=None, out16=None,
                      out_bs=None, out_off=0):
    """GroupNorm(32, 256) of x [B*H*W, 256] fp32 (token-major) (+ bilinear(add) + ReLU).  add_layout = ("tokens", rows per
    sample, first row, hs, ws) for a token-major fp32 map, ("nchw", hs, ws) for a [B, 256, hs, ws] one.  Row (b, r) goes to
    row b*out_bs + out_off + r of out32 / out16."""
    lib = load()
    _req(x, torch.float32, "x")
    S = H * W
    assert x.shape == (B * S, 256)
    stats = torch.zeros(B, 32, 2, dtype=torch.float64, device=x.device)
    _check(lib.ovis_gn_stats(_p(x), _p(stats), B, S, _stream()))
    a_ptr, a_bs, a_cs, a_ps, hs, ws = None, 0, 0, 0, 0, 0
    if add is not None:
        _req(add, torch.float32, "add")
        if add_layout[0] == "tokens":
            _, rows, first, hs, ws = add_layout
            a_ptr, a_bs, a_cs, a_ps = add.data_ptr() + first * 256 * 4, rows * 256, 1, 256
        else:
            _, hs, ws = add_layout
            assert add.shape == (B, 256, hs, ws)
            a_ptr, a_bs, a_cs, a_ps = add.data_ptr(), 256 * hs * ws, hs * ws, 1
    if out32 is None and out16 is None:
        out16 = torch.empty(B * S, 256, dtype=torch.float16, device=x.device)
    _check(lib.ovis_gn_apply(_p(x), _p(stats), _p(gamma), _p(beta), float(eps), B, H, W, int(relu), a_ptr, a_bs, a_cs, a_ps, hs, ws,
                             _p(out32), _p(out16), S if out_bs is None else out_bs, out_off, _stream()))
    return out32, out16


def tokens_to_nchw(x, B, C, N, in_bs, in_off, out=None):
    lib = load()
    _req(x, torch.float32, "x")
    if out is None:
        out = torch.empty(B, C, N, dtype=torch.float32, device=x.device)
    _check(lib.ovis_tokens_to_nchw_f32(_p(x), _p(out), B, C, N, in_bs, in_off, _stream()))
    return out


def conv3x3_unfold_f16(x, B, H, W, out=None):
    """x [B*H*W, C] fp16 (token-major maps) -> [B*H*W, 9*C] fp16."""
    lib = load()
    assert x.dtype == torch.float16 and x.is_contiguous()
    C = x.shape[-1]
    if out is None:
        out = torch.empty(B * H * W, 9 * C, dtype=torch.float16, device=x.device)
    _check(lib.ovis_conv3x3_unfold_f16(_p(x), _p(out), B, H, W, C, _stream()))
    return out


@_timed("prep")
def tokens_pool_f16(ft, B, H, W, s, out):
    """ft [B*H*W, 256] fp16 token-major -> out [B*(H/s)*(W/s), 256] fp16: mean of the centre 2x2 pixels of every s x s block."""
    lib = load()
    assert ft.dtype == torch.float16 and out.dtype == torch.float16 and ft.is_contiguous() and out.is_contiguous()
    assert ft.numel() == B * H * W * 256 and out.numel() == B * (H // s) * (W // s) * 256
    _check(lib.ovis_tokens_pool_f16(_p(ft), _p(out), B, H, W, s, _stream()))
    return out


@_timed("prep")
def tokens_add_pos_f16(xt, pos, pos_t, out):
    """xt [B, N, 256] fp16 + pos [N, 256] fp32 (+ pos_t [B, 256] fp32) -> out fp16 [B, N, 256]."""
    lib = load()
    assert xt.dtype == torch.float16 and out.dtype == torch.float16 and xt.is_contiguous() and out.is_contiguous()
    _req(pos, torch.float32, "pos")
    _req(pos_t, torch.float32, "pos_t")
    B, N = xt.shape[0], xt.shape[1]
    assert pos.shape == (N, 256) and (pos_t is None or pos_t.shape == (B, 256)) and out.numel() == xt.numel()
    _check(lib.ovis_tokens_add_pos_f16(_p(xt), _p(pos), _p(pos_t), _p(out), B, N, _stream()))
    return out
