"""Multi-scale deformable attention on the B200 kernel: drop-in for the reference's
``MSDeformAttnFunction`` (openvis/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py:32-52), forward only.

Reference-side change (ops/functions/ms_deform_attn_func.py:22-30): replace ``import MultiScaleDeformableAttention as
MSDA`` by ``from openvis_b200.msda import MSDA`` -- ``MSDA.ms_deform_attn_forward`` keeps the extension's signature
(``im2col_step`` is accepted and ignored: the whole batch is one launch).  Training needs the reference's backward.
"""
import torch

from . import _lib as L


class _MSDA:
    """Stand-in for the compiled ``MultiScaleDeformableAttention`` module (vision.cpp:18-21)."""

    @staticmethod
    def ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                               im2col_step=None):
        if not value.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        with torch.cuda.device(value.device):
            return L.ms_deform_attn_forward(value.float().contiguous(), value_spatial_shapes.long().contiguous(),
                                            value_level_start_index.long().contiguous(),
                                            sampling_locations.float().contiguous(), attention_weights.float().contiguous())

    @staticmethod
    def ms_deform_attn_backward(*_a, **_k):
        raise NotImplementedError("openvis_b200.msda is inference-only (training keeps the reference's extension)")


MSDA = _MSDA()


class MSDeformAttnFunction(torch.autograd.Function):
    """``MSDeformAttnFunction.apply(value, shapes, level_start_index, sampling_locations, attention_weights, im2col_step)``."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step):
        return MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                           attention_weights, im2col_step)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("openvis_b200.msda is inference-only")


class MSDeformAttn(torch.nn.Module):
    """Drop-in for the reference ``MSDeformAttn`` at inference (ops/modules/ms_deform_attn.py:35-125): same constructor,
    parameter names (``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj``) and ``forward`` arguments,
    so a reference state dict loads with ``load_state_dict``.  Four launches + the casts: value projection (tcgen05 GEMM,
    fp16 operands, fp32 accumulate and output) -- ONE GEMM for sampling_offsets | attention_weights -- softmax + sampling
    locations (``ovis_msda_prepare``) -- the sampling kernel -- output projection."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        self.im2col_step = 128
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = torch.nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = torch.nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = torch.nn.Linear(d_model, d_model)
        self.output_proj = torch.nn.Linear(d_model, d_model)
        self._wc = None

    def _weights(self):
        ps = (self.sampling_offsets.weight, self.sampling_offsets.bias, self.attention_weights.weight, self.attention_weights.bias,
              self.value_proj.weight, self.value_proj.bias, self.output_proj.weight, self.output_proj.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._wc is None or self._wc[0] != key:
            f = lambda t: t.detach().float().contiguous()
            self._wc = (key, dict(
                qw=L.cast_f16(torch.cat([f(ps[0]), f(ps[2])]).contiguous()), qb=torch.cat([f(ps[1]), f(ps[3])]).contiguous(),
                vw=L.cast_f16(f(ps[4])), vb=f(ps[5]), ow=L.cast_f16(f(ps[6])), ob=f(ps[7])))
        return self._wc[1]

    @torch.no_grad()
    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        if not query.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        N, Len_q, C = query.shape
        _, Len_in, _ = input_flatten.shape
        if int((input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum()) != Len_in:
            raise ValueError("input_spatial_shapes do not add up to the flattened input length")
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")
        with torch.cuda.device(query.device):
            W = self._weights()
            value = L.linear_f16(L.cast_f16(input_flatten.float().reshape(N * Len_in, C).contiguous()), W["vw"], W["vb"], out_f32=True)
            if input_padding_mask is not None:
                value = value.masked_fill(input_padding_mask.reshape(-1, 1), 0.0)
            proj = L.linear_f16(L.cast_f16(query.float().reshape(N * Len_q, C).contiguous()), W["qw"], W["qb"], out_f32=True)
            shapes = input_spatial_shapes.long().contiguous()
            loc, w = L.msda_prepare(proj.view(N, Len_q, -1), reference_points.float().contiguous(), shapes, self.n_heads,
                                    self.n_levels, self.n_points)
            out = L.ms_deform_attn_forward(value.view(N, Len_in, self.n_heads, C // self.n_heads), shapes,
                                           input_level_start_index.long().contiguous(), loc, w)
            out = L.linear_f16(L.cast_f16(out.view(N * Len_q, C)), W["ow"], W["ob"], out_f32=True)
            return out.view(N, Len_q, C)
