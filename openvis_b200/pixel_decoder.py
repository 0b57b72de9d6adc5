"""MSDeformAttnPixelDecoder on the device: a drop-in for the inference path of the reference's
pixel decoder (openvis/modeling/pixel_decoder/msdeformattn.py:182-380) with the reference's parameter names, so its state
dict loads by name.

Schedule of ``forward_features`` (every map lives token-major [frames, positions, 256]; the convolutions are tcgen05 GEMMs):

  res5 / res4 / res3  --nchw_to_tokens-->  fp16 operand  --1x1 conv GEMM-->  fp32  --GroupNorm (+ extra feature)-->  rows of the
                       flattened encoder input [B, S, 256]                                          (msdeformattn.py:337-343)
  encoder             msda.MSDeformAttnTransformerEncoder (6 layers: deformable attention + FFN)    (:345)
  multi_scale_features  token rows of every level  --tokens_to_nchw-->  [B, 256, h, w] fp32        (:358-359, 376-378)
  res2                --nchw_to_tokens--> 1x1 lateral GEMM --> GroupNorm + bilinear(stride-8 level) --> fp16
                      --3x3 unfold + GEMM (K = 2304)--> GroupNorm + ReLU --> fp16                   (:364-373)
  mask_features       1x1 conv GEMM with the transposed fp32 store: written once, directly NCHW    (:380)

There is no CPU path and no autograd (inference only), like the decoders.
"""
from typing import Dict, List

import torch
from torch import nn

from . import _lib as L
from .decoder import Registry, _configurable_new, sine_pos_2d
from .msda import MSDeformAttnTransformerEncoderOnly

SEM_SEG_HEADS_REGISTRY = Registry("SEM_SEG_HEADS")


class ShapeSpec:
    """Minimal stand-in for detectron2.layers.ShapeSpec (only .channels / .stride are read)."""

    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class DecoderTokens:
    """Token-major hand-off from `MSDeformAttnPixelDecoder.forward_tokens` to a B200 decoder's `forward_tokens`: what the
    decoder's own layout kernels would otherwise rebuild from the fp32 NCHW maps.  xt[l] [BT, h_l*w_l, 256] fp16 (multi-scale
    features, coarsest first), ft [BT, H/4*W/4, 256] fp16 (mask features), sizes [(h_l, w_l)], (H4, W4)."""

    def __init__(self, xt, ft, sizes, H4, W4):
        self.xt, self.ft, self.sizes, self.H4, self.W4 = xt, ft, sizes, H4, W4


class _ConvGN(nn.Conv2d):
    """Parameter container with the names of detectron2's Conv2d wrapper: .weight (.bias) and .norm.{weight, bias}."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__(cin, cout, k, padding=padding, bias=False)
        self.norm = nn.GroupNorm(32, cout)


@SEM_SEG_HEADS_REGISTRY.register()
@_configurable_new
class MSDeformAttnPixelDecoder(nn.Module):
    def __init__(self, input_shape: Dict[str, ShapeSpec] = None, *, transformer_dropout: float = 0.0, transformer_nheads: int = 8,
                 transformer_dim_feedforward: int = 1024, transformer_enc_layers: int = 6, conv_dim: int = 256,
                 mask_dim: int = 256, norm="GN", transformer_in_features: List[str] = ("res3", "res4", "res5"),
                 common_stride: int = 4):
        super().__init__()
        items = sorted(input_shape.items(), key=lambda kv: kv[1].stride)
        self.in_features = [k for k, _ in items]
        self.feature_strides = [v.stride for _, v in items]
        self.feature_channels = [v.channels for _, v in items]
        t_items = [(k, v) for k, v in items if k in transformer_in_features]
        self.transformer_in_features = [k for k, _ in t_items]
        t_channels = [v.channels for _, v in t_items]
        self.transformer_feature_strides = [v.stride for _, v in t_items]
        self.transformer_num_feature_levels = len(t_items)
        self.common_stride = common_stride
        stride = min(self.transformer_feature_strides)
        self.num_fpn_levels = 0
        while (common_stride << self.num_fpn_levels) < stride:
            self.num_fpn_levels += 1
        if (conv_dim != 256 or mask_dim != 256 or norm != "GN" or transformer_nheads != 8 or self.transformer_num_feature_levels != 3
                or self.num_fpn_levels != 1 or (common_stride << 1) != stride):
            raise NotImplementedError("openvis_b200 pixel decoder: conv_dim = mask_dim = 256, GroupNorm, 8 heads, the three coarsest "
                                      "maps in the encoder and one FPN level at stride 4 (every shipped config)")
        if any(c % 64 for c in self.feature_channels):
            raise NotImplementedError("backbone channel counts must be multiples of 64 (ResNet, Swin-B / Swin-L)")
        self.input_proj = nn.ModuleList([nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim))
                                         for c in t_channels[::-1]])
        self.transformer = MSDeformAttnTransformerEncoderOnly(d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
                                                              dim_feedforward=transformer_dim_feedforward,
                                                              num_encoder_layers=transformer_enc_layers,
                                                              num_feature_levels=self.transformer_num_feature_levels)
        self.mask_dim = mask_dim
        self.mask_features = nn.Conv2d(conv_dim, mask_dim, kernel_size=1)
        self.maskformer_num_feature_levels = 3
        self.adapter_1 = _ConvGN(self.feature_channels[0], conv_dim, 1)
        self.layer_1 = _ConvGN(conv_dim, conv_dim, 3, padding=1)
        self.unfold_frames = 4          # frames per 3x3-unfold GEMM (bounds the [frames*H*W, 2304] scratch)
        self._wc = None
        self._tc = None
        self.eval()

    @classmethod
    def from_config(cls, cfg, input_shape):
        """msdeformattn.py:308-327."""
        h = cfg.MODEL.SEM_SEG_HEAD
        return dict(input_shape={k: v for k, v in input_shape.items() if k in h.IN_FEATURES}, conv_dim=h.CONVS_DIM,
                    mask_dim=h.MASK_DIM, norm=h.NORM, transformer_dropout=cfg.MODEL.MASK_FORMER.DROPOUT,
                    transformer_nheads=cfg.MODEL.MASK_FORMER.NHEADS, transformer_dim_feedforward=1024,
                    transformer_enc_layers=h.TRANSFORMER_ENC_LAYERS,
                    transformer_in_features=h.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES, common_stride=h.COMMON_STRIDE)

    def _weights(self):
        ps = [self.input_proj[i][j].weight for i in range(3) for j in (0, 1)] + [self.adapter_1.weight, self.layer_1.weight,
                                                                                  self.mask_features.weight]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._wc is None or self._wc[0] != key:
            f = lambda t: t.detach().float().contiguous()
            mat = lambda w: L.cast_f16(f(w).reshape(w.shape[0], -1).contiguous())
            W = dict(proj=[(mat(self.input_proj[i][0].weight), f(self.input_proj[i][0].bias), f(self.input_proj[i][1].weight),
                            f(self.input_proj[i][1].bias), float(self.input_proj[i][1].eps)) for i in range(3)],
                     lat=(mat(self.adapter_1.weight), f(self.adapter_1.norm.weight), f(self.adapter_1.norm.bias),
                          float(self.adapter_1.norm.eps)),
                     # [C_out][(ky, kx, C_in)]: the operand order of ovis_conv3x3_unfold_f16
                     out=(L.cast_f16(f(self.layer_1.weight).permute(0, 2, 3, 1).reshape(256, -1).contiguous()),
                          f(self.layer_1.norm.weight), f(self.layer_1.norm.bias), float(self.layer_1.norm.eps)),
                     mf=(mat(self.mask_features.weight), f(self.mask_features.bias)))
            self._wc = (key, W)
        return self._wc[1]

    def _tables(self, shapes, starts, dev):
        """Input-size-dependent constants, cached: position table (sine embedding + level embedding) [1, S, 256], the shape /
        start-index tensors, and the encoder's reference points for all-valid maps (msdeformattn.py:155-169 with every valid
        ratio = 1: the pixel centres, identical for every level and frame) [1, S, 3, 2]."""
        le = self.transformer.level_embed
        key = (tuple(shapes), str(dev), le.data_ptr(), le._version)
        if self._tc is None or self._tc[0] != key:
            lef = le.detach().float()
            pos = torch.cat([sine_pos_2d(h, w, dev) + lef[l][None, :] for l, (h, w) in enumerate(shapes)], 0)[None].contiguous()
            pts = []
            for (h, w) in shapes:
                y = (torch.arange(h, dtype=torch.float32, device=dev) + 0.5) / h
                x = (torch.arange(w, dtype=torch.float32, device=dev) + 0.5) / w
                pts.append(torch.stack([x[None, :].expand(h, w), y[:, None].expand(h, w)], -1).reshape(h * w, 2))
            ref = torch.cat(pts, 0)[None, :, None, :].expand(1, -1, len(shapes), 2).contiguous()
            self._tc = (key, (pos, torch.as_tensor(shapes, dtype=torch.long, device=dev),
                              torch.as_tensor(starts, dtype=torch.long, device=dev), ref))
        return self._tc[1]

    @torch.no_grad()
    def forward_tokens(self, features, extra_features=None):
        """The same computation as forward_features, handed over token-major in fp16 (`DecoderTokens`) for
        `decoder.forward_tokens`: the multi-scale maps are the encoder's own fp16 rows, the mask-feature convolution stores fp16
        rows instead of an fp32 NCHW map -- no NCHW store here and no layout pass (2.9 GB read + 2.0 GB written per 36-frame
        720 x 1280 clip) in the decoder."""
        return self.forward_features(features, extra_features, _tokens=True)

    @torch.no_grad()
    def forward_features(self, features, extra_features=None, _tokens=False):
        if self.training:
            raise RuntimeError("openvis_b200 pixel decoder is inference-only: call .eval()")
        names = self.transformer_in_features[::-1]                    # res5, res4, res3: low to high resolution
        xs = [features[f].float().contiguous() for f in names]
        x2 = features[self.in_features[0]].float().contiguous()
        if not x2.is_cuda or any(not x.is_cuda for x in xs):
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        dev = x2.device
        B = x2.shape[0]
        shapes = [tuple(x.shape[-2:]) for x in xs]
        S = sum(h * w for h, w in shapes)
        starts = [sum(h * w for h, w in shapes[:i]) for i in range(3)]
        with torch.cuda.device(dev):
            W = self._weights()
            # ---- input projections: 1x1 conv + GroupNorm (+ extra feature), written as rows of the encoder input
            src = torch.empty(B, S, 256, dtype=torch.float32, device=dev)
            for i, x in enumerate(xs):
                h, w = shapes[i]
                wt, bias, g, b, eps = W["proj"][i]
                y = L.linear_f16(L.nchw_to_tokens_f16(x).view(B * h * w, -1), wt, bias, out_f32=True)
                add, lay = None, None
                if extra_features is not None:
                    add = extra_features[i].float().contiguous()
                    lay = ("nchw", add.shape[-2], add.shape[-1])
                L.group_norm_tokens(y, B, h, w, g, b, eps, add=add, add_layout=lay, out32=src.view(B * S, 256), out_bs=S,
                                    out_off=starts[i])
            # ---- deformable encoder; position term = sine embedding + level embedding, one table for every frame
            pos, spatial_shapes, level_start, ref = self._tables(shapes, starts, dev)
            mem = self.transformer.encoder(src, spatial_shapes, level_start, None, pos, None, _reference_points=ref).reshape(B * S, 256)
            # ---- maps returned in the reference's layout (token hand-off: the encoder's fp16 rows, level by level)
            if _tokens:
                m16 = self.transformer.encoder.layers[-1]._last_state[1].view(B, S, 256)
                xt = [m16[:, starts[i]:starts[i] + h * w].contiguous() for i, (h, w) in enumerate(shapes)]
            else:
                out = [L.tokens_to_nchw(mem, B, 256, h * w, S, starts[i]).view(B, 256, h, w) for i, (h, w) in enumerate(shapes)]
            # ---- FPN level (res2): lateral conv + GN + top-down bilinear addition, 3x3 output conv + GN + ReLU
            H, Wd = x2.shape[-2:]
            wt, g, b, eps = W["lat"]
            lat = L.linear_f16(L.nchw_to_tokens_f16(x2).view(B * H * Wd, -1), wt, None, out_f32=True)
            h3, w3 = shapes[2]
            _, y16 = L.group_norm_tokens(lat, B, H, Wd, g, b, eps, add=mem, add_layout=("tokens", S, starts[2], h3, w3))
            del lat
            wt, g, b, eps = W["out"]
            conv = torch.empty(B * H * Wd, 256, dtype=torch.float32, device=dev)
            step = max(1, int(self.unfold_frames))
            scratch = torch.empty(min(step, B) * H * Wd, 9 * 256, dtype=torch.float16, device=dev)
            for b0 in range(0, B, step):
                nb = min(step, B - b0)
                rows = slice(b0 * H * Wd, (b0 + nb) * H * Wd)
                u = L.conv3x3_unfold_f16(y16[rows], nb, H, Wd, out=scratch[:nb * H * Wd])
                L.linear_f16(u, wt, None, out=conv[rows], out_f32=True)
            del scratch
            _, y16 = L.group_norm_tokens(conv, B, H, Wd, g, b, eps, relu=True, out16=y16)
            del conv
            # ---- mask features: 1x1 conv, transposed fp32 store = NCHW directly
            wt, bias = W["mf"]
            if _tokens:
                ft = L.linear_f16(y16, wt, bias).view(B, H * Wd, 256)
                return DecoderTokens(xt, ft, [tuple(s_) for s_ in shapes], H, Wd)
            mf = torch.empty(B, 256, H, Wd, dtype=torch.float32, device=dev)
            L.mask_logits(y16, B, H * Wd, wt, 0, 256, mf, 256 * H * Wd, H * Wd, bias=bias)
        return mf, out[0], out[:self.maskformer_num_feature_levels]


def build_pixel_decoder(cfg, input_shape):
    """msdeformattn.py:22-35."""
    return SEM_SEG_HEADS_REGISTRY[cfg.MODEL.SEM_SEG_HEAD.PIXEL_DECODER_NAME](cfg, input_shape)


def register_into(registry):
    """Registers the class under the reference's name in a Detectron2-style registry (replacing the reference entry)."""
    store = getattr(registry, "_obj_map", registry)
    store["MSDeformAttnPixelDecoder"] = MSDeformAttnPixelDecoder
    return registry
