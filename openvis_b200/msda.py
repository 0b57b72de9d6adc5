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
    so a reference state dict loads with ``load_state_dict``.  Four launches: value projection (tcgen05 GEMM, fp16 operands
    and output, fp32 accumulate) -- ONE GEMM for sampling_offsets | attention_weights (fp32 out) -- ``ovis_msda_fused_f16``
    (softmax + sampling locations in registers, bilinear gather of the fp16 value map with fp32 weights / accumulators, fp16
    rows out) -- output projection.  Head widths other than 32 take the fp32 value map through ``ovis_msda_prepare`` +
    ``ovis_ms_deform_attn_forward``."""

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
        self.fused = True       # False: value map in fp32 + separate softmax / location kernel (A/B, other head widths)
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
                input_padding_mask=None, _skip_output_proj=False, _query_f16=None, _input_f16=None, _trusted=False):
        if not query.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        N, Len_q, C = query.shape
        _, Len_in, _ = input_flatten.shape
        # (the encoder validates its shapes once per call: this check reads the device tensor back, i.e. synchronises)
        if not _trusted and int((input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum()) != Len_in:
            raise ValueError("input_spatial_shapes do not add up to the flattened input length")
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")
        with torch.cuda.device(query.device):
            W = self._weights()
            x16 = _input_f16 if _input_f16 is not None else L.cast_f16(input_flatten.float().reshape(N * Len_in, C).contiguous())
            q16 = _query_f16 if _query_f16 is not None else L.cast_f16(query.float().reshape(N * Len_q, C).contiguous())
            proj = L.linear_f16(q16, W["qw"], W["qb"], out_f32=True)
            shapes = input_spatial_shapes.long().contiguous()
            start = input_level_start_index.long().contiguous()
            if C // self.n_heads == 32 and self.n_levels * self.n_points <= 16 and self.fused:
                # fp16 value map, softmax + sampling locations + gather in one kernel, fp16 rows for output_proj
                value = L.linear_f16(x16, W["vw"], W["vb"])
                if input_padding_mask is not None:
                    value = value.masked_fill(input_padding_mask.reshape(-1, 1), 0.0)
                out16 = L.msda_fused_f16(value, proj, reference_points.float().contiguous(), shapes, start, N, Len_in, Len_q,
                                         self.n_heads, self.n_levels, self.n_points)
            else:
                value = L.linear_f16(x16, W["vw"], W["vb"], out_f32=True)
                if input_padding_mask is not None:
                    value = value.masked_fill(input_padding_mask.reshape(-1, 1), 0.0)
                loc, w = L.msda_prepare(proj.view(N, Len_q, -1), reference_points.float().expand(N, -1, -1, -1).contiguous(), shapes, self.n_heads,
                                        self.n_levels, self.n_points)
                out = L.ms_deform_attn_forward(value.view(N, Len_in, self.n_heads, C // self.n_heads), shapes, start, loc, w)
                out16 = L.cast_f16(out.view(N * Len_q, C))
            if _skip_output_proj:
                return out16                # [N * Len_q, C] fp16: the caller fuses output_proj into its residual + LayerNorm
            out = L.linear_f16(out16, W["ow"], W["ob"], out_f32=True)
            return out.view(N, Len_q, C)


class MSDeformAttnTransformerEncoderLayer(torch.nn.Module):
    """Drop-in for the reference encoder layer at inference (pixel_decoder/msdeformattn.py:107-146): ``self_attn`` /
    ``norm1`` / ``linear1`` / ``linear2`` / ``norm2`` keep their names.  ``output_proj`` + residual + ``norm1`` and
    ``linear2`` + residual + ``norm2`` are one GEMM each with the LayerNorm in the epilogue (``ovis_linear_ln_f16``), which
    also emits the fp16 operand of the next GEMM and ``src + pos`` for the next layer's query projections."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model != 256 or activation != "relu":
            raise NotImplementedError("openvis_b200: d_model = 256 and ReLU (every shipped pixel-decoder config)")
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.norm1 = torch.nn.LayerNorm(d_model)
        self.linear1 = torch.nn.Linear(d_model, d_ffn)
        self.linear2 = torch.nn.Linear(d_ffn, d_model)
        self.norm2 = torch.nn.LayerNorm(d_model)
        self._wc = None

    def _weights(self):
        ps = (self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias, self.norm1.weight, self.norm1.bias,
              self.norm2.weight, self.norm2.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._wc is None or self._wc[0] != key:
            f = lambda t: t.detach().float().contiguous()
            self._wc = (key, dict(w1=L.cast_f16(f(ps[0])), b1=f(ps[1]), w2=L.cast_f16(f(ps[2])), b2=f(ps[3]),
                                  n1=(f(ps[4]), f(ps[5])), n2=(f(ps[6]), f(ps[7]))))
        return self._wc[1]

    @torch.no_grad()
    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None, _state=None, _split_ws=None):
        """`_state` (encoder-internal): (src fp32 [rows, C], src fp16, (src + pos) fp16, pos table [S, C] or None) of the
        previous layer's epilogue, so that no operand is converted twice."""
        N, S, C = src.shape
        with torch.cuda.device(src.device):
            W, A = self._weights(), self.self_attn._weights()
            if _state is None:
                x32 = src.float().reshape(N * S, C).contiguous()
                x16 = L.cast_f16(x32)
                q16 = x16 if pos is None else L.cast_f16((src + pos).float().reshape(N * S, C).contiguous())
                pe = None
            else:
                x32, x16, q16, pe = _state
            att16 = self.self_attn(src, reference_points, x32.view(N, S, C), spatial_shapes, level_start_index, padding_mask,
                                   _skip_output_proj=True, _query_f16=q16, _input_f16=x16, _trusted=_state is not None)
            y32 = torch.empty_like(x32)
            y16 = torch.empty_like(x16)
            L.linear_ln_f16(att16, A["ow"], A["ob"], x32, W["n1"], y32=y32, y16=y16, split_ws=_split_ws)
            h16 = L.linear_f16(y16, W["w1"], W["b1"], relu=True)
            o32, o16 = torch.empty_like(x32), torch.empty_like(x16)
            oq16 = torch.empty_like(x16) if pe is not None else None
            L.linear_ln_f16(h16, W["w2"], W["b2"], y32, W["n2"], pe=pe, y32=o32, y16=o16, ype16=oq16, split_ws=_split_ws)
            self._last_state = (o32, o16, oq16 if oq16 is not None else o16, pe)
            return o32.view(N, S, C)


class MSDeformAttnTransformerEncoder(torch.nn.Module):
    """pixel_decoder/msdeformattn.py:149-177: the layer stack + reference points of every position at every level."""

    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        import copy
        self.layers = torch.nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        ref_list = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            H_, W_ = int(H_), int(W_)
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device), indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            ref_list.append(torch.stack((ref_x, ref_y), -1))
        reference_points = torch.cat(ref_list, 1)
        return reference_points[:, :, None] * valid_ratios[:, None]

    @torch.no_grad()
    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None, _reference_points=None):
        """`_reference_points` (pixel decoder: cached per input size) [1 or N, S, L, 2] replaces get_reference_points."""
        if not src.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        N, S, C = src.shape
        if _reference_points is None:
            if int((spatial_shapes[:, 0] * spatial_shapes[:, 1]).sum()) != S:
                raise ValueError("spatial_shapes do not add up to the flattened input length")
            reference_points = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device).contiguous()
        else:
            reference_points = _reference_points
        state = None
        # the position term is usually the same table for every sample (sine embedding + level embedding): then the
        # LayerNorm epilogue adds it for the next layer's query operand
        shared_pos = pos is not None and (pos.shape[0] == 1 or bool((pos == pos[:1]).all()))
        out = src
        with torch.cuda.device(src.device):
            # fp32 scratch of the linear + LayerNorm products (split-K partials for few rows, the plain product for many)
            rows_pad = (N * S + 127) // 128 * 128
            d_ffn = self.layers[0].linear1.out_features
            split_ws = torch.empty((d_ffn // 256 if N * S <= 16384 else 1) * rows_pad * 256, dtype=torch.float32, device=src.device)
            for i, layer in enumerate(self.layers):
                if i == 0 and shared_pos:
                    x32 = src.float().reshape(N * S, C).contiguous()
                    state = (x32, L.cast_f16(x32), L.cast_f16((src + pos).float().reshape(N * S, C).contiguous()),
                             pos[0].float().contiguous())
                out = layer(out, pos, reference_points, spatial_shapes, level_start_index, padding_mask,
                            _state=state if shared_pos else None, _split_ws=split_ws)
                state = layer._last_state if shared_pos else None
        return out


class MSDeformAttnTransformerEncoderOnly(torch.nn.Module):
    """pixel_decoder/msdeformattn.py:38-104: flattens the multi-scale maps, adds the level embedding to the position
    embeddings and runs the encoder; returns (memory, spatial_shapes, level_start_index)."""

    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1, activation="relu",
                 num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model, self.nhead = d_model, nhead
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_feature_levels, nhead,
                                                    enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = torch.nn.Parameter(torch.zeros(num_feature_levels, d_model))

    @torch.no_grad()
    def forward(self, srcs, pos_embeds):
        src_flatten, pos_flatten, shapes = [], [], []
        for lvl, (src, pos_embed) in enumerate(zip(srcs, pos_embeds)):
            shapes.append(tuple(src.shape[-2:]))
            src_flatten.append(src.flatten(2).transpose(1, 2))
            pos_flatten.append(pos_embed.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        src_flatten, pos_flatten = torch.cat(src_flatten, 1), torch.cat(pos_flatten, 1)
        spatial_shapes = torch.as_tensor(shapes, dtype=torch.long, device=src_flatten.device)
        level_start_index = torch.cat((spatial_shapes.new_zeros((1,)), spatial_shapes.prod(1).cumsum(0)[:-1]))
        # masks are all-False in the reference (msdeformattn.py:77): every valid ratio is 1
        valid_ratios = torch.ones(src_flatten.shape[0], len(shapes), 2, device=src_flatten.device)
        memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, pos_flatten, None)
        return memory, spatial_shapes, level_start_index
