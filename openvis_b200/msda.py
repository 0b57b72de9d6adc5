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
