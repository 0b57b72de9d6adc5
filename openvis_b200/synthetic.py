"""Seeded synthetic weights / inputs for tests and bench.py (no reference, no oracle needed to regenerate them)."""
import math

import torch

def decoder_param_shapes(kind="frame", Q=100, C=256, F_=2048, L=9, num_classes=1, clip_heads=12, clip_dims=512):
    """state_dict contract of the reference decoders.  kind: frame / video / san_frame / san_video,
    or embedding_{frame,video} (class_embed = MLP(C, 2*clip_dims, clip_dims, 2), video_..._decoder.py:513-515) /
    proposal_{frame,video} (class_embed = Linear(C, 2), :536-537) / zero_shot (object_embed = MLP(C, C, 2, 2), no class_embed,
    zero_shot_mask2former_transformer_decoder.py:142)."""
    s = {}
    for i in range(L):
        for pre, att in ((f"transformer_self_attention_layers.{i}", "self_attn"),
                         (f"transformer_cross_attention_layers.{i}", "multihead_attn")):
            s[f"{pre}.{att}.in_proj_weight"] = (3 * C, C)
            s[f"{pre}.{att}.in_proj_bias"] = (3 * C,)
            s[f"{pre}.{att}.out_proj.weight"] = (C, C)
            s[f"{pre}.{att}.out_proj.bias"] = (C,)
            s[f"{pre}.norm.weight"] = (C,)
            s[f"{pre}.norm.bias"] = (C,)
        pre = f"transformer_ffn_layers.{i}"
        s[f"{pre}.linear1.weight"] = (F_, C)
        s[f"{pre}.linear1.bias"] = (F_,)
        s[f"{pre}.linear2.weight"] = (C, F_)
        s[f"{pre}.linear2.bias"] = (C,)
        s[f"{pre}.norm.weight"] = (C,)
        s[f"{pre}.norm.bias"] = (C,)
    s["decoder_norm.weight"] = (C,)
    s["decoder_norm.bias"] = (C,)
    s["query_feat.weight"] = (Q, C)
    s["query_embed.weight"] = (Q, C)
    s["level_embed.weight"] = (3, C)
    for i in range(3):
        s[f"mask_embed.layers.{i}.weight"] = (C, C)
        s[f"mask_embed.layers.{i}.bias"] = (C,)
    if kind in ("san_frame", "san_video"):
        for i in range(3):
            s[f"attn_embed.layers.{i}.weight"] = (C, C)
            s[f"attn_embed.layers.{i}.bias"] = (C,)
        for i in range(3):
            o = C * clip_heads if i == 2 else C
            s[f"attn_mlp.layers.{i}.weight"] = (o, C, 1, 1)
            s[f"attn_mlp.layers.{i}.bias"] = (o,)
    elif kind == "zero_shot":
        s["object_embed.layers.0.weight"] = (C, C)
        s["object_embed.layers.0.bias"] = (C,)
        s["object_embed.layers.1.weight"] = (2, C)
        s["object_embed.layers.1.bias"] = (2,)
    elif kind.startswith("embedding"):
        s["class_embed.layers.0.weight"] = (2 * clip_dims, C)
        s["class_embed.layers.0.bias"] = (2 * clip_dims,)
        s["class_embed.layers.1.weight"] = (clip_dims, 2 * clip_dims)
        s["class_embed.layers.1.bias"] = (clip_dims,)
    elif kind.startswith("proposal"):
        s["class_embed.weight"] = (2, C)
        s["class_embed.bias"] = (2,)
    else:
        s["class_embed.weight"] = (num_classes + 1, C)
        s["class_embed.bias"] = (num_classes + 1,)
    return s


def clip_block_param_shapes(width=768, blocks=(9, 10, 11)):
    """state_dict names of the post-split CLIP blocks the SAN side path runs (ViT-B/16: width 768, 12 heads;
    side_adapter.py:188 `resblocks[self.broken_idx:]`), relative to `clip_model.visual.transformer.resblocks`."""
    s = {}
    for i in blocks:
        s[f"{i}.ln_1.weight"] = (width,)
        s[f"{i}.ln_1.bias"] = (width,)
        s[f"{i}.attn.in_proj_weight"] = (3 * width, width)
        s[f"{i}.attn.in_proj_bias"] = (3 * width,)
        s[f"{i}.attn.out_proj.weight"] = (width, width)
        s[f"{i}.attn.out_proj.bias"] = (width,)
        s[f"{i}.ln_2.weight"] = (width,)
        s[f"{i}.ln_2.bias"] = (width,)
        s[f"{i}.mlp.c_fc.weight"] = (4 * width, width)
        s[f"{i}.mlp.c_fc.bias"] = (4 * width,)
        s[f"{i}.mlp.c_proj.weight"] = (width, 4 * width)
        s[f"{i}.mlp.c_proj.bias"] = (width,)
    return s


def seeded_clip_block_params(seed=0, width=768, blocks=(9, 10, 11)):
    """Deterministic CLIP-block weights with the scales of CLIP.initialize_parameters (attn std width^-0.5, output
    projections damped by (2 * layers)^-0.5, fc std (2 * width)^-0.5), LayerNorm near identity."""
    g = torch.Generator().manual_seed(seed)
    shapes = clip_block_param_shapes(width, blocks)
    attn_std, proj_std, fc_std = width ** -0.5, (width ** -0.5) * (2 * 12) ** -0.5, (2 * width) ** -0.5
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("ln_1.weight") or name.endswith("ln_2.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shp, generator=g)
        elif name.endswith("in_proj_weight"):
            t = attn_std * torch.randn(shp, generator=g)
        elif name.endswith("c_fc.weight"):
            t = fc_std * torch.randn(shp, generator=g)
        else:
            t = proj_std * torch.randn(shp, generator=g)
        out[name] = t
    return out


def seeded_clip_visual_params(seed=0, width=768, layers=12, patch=16, resolution=224, output_dim=512):
    """Deterministic state dict of a CLIP VisionTransformer (mask_adapted_clip model.py:288-325 key names: conv1.weight,
    class_embedding, positional_embedding, ln_pre.*, transformer.resblocks.*, ln_post.*, proj) with the initial scales of
    the reference model (width^-0.5 embeddings and projection) and near-identity LayerNorms."""
    g = torch.Generator().manual_seed(seed)
    scale = width ** -0.5
    out = {"conv1.weight": (3 * patch * patch) ** -0.5 * torch.randn(width, 3, patch, patch, generator=g),
           "class_embedding": scale * torch.randn(width, generator=g),
           "positional_embedding": scale * torch.randn((resolution // patch) ** 2 + 1, width, generator=g),
           "ln_pre.weight": 1.0 + 0.1 * torch.randn(width, generator=g), "ln_pre.bias": 0.02 * torch.randn(width, generator=g),
           "ln_post.weight": 1.0 + 0.1 * torch.randn(width, generator=g), "ln_post.bias": 0.02 * torch.randn(width, generator=g),
           "proj": scale * torch.randn(width, output_dim, generator=g)}
    for k, v in seeded_clip_block_params(seed + 1, width, tuple(range(layers))).items():
        out["transformer.resblocks." + k] = v
    return out


def seeded_crop_inputs(T=2, N=5, H=320, W=416, seed=0):
    """Frames [T, 3, H, W] in 0..255 (smooth colour fields + noise) and mask logits [N, T, H, W]: query 0 covers most of the
    image (crop side > 224: more than one roi_align sample per bin), query 1 a small box near the right border (the square
    crop box leaves the image), query 2 an ellipse, query 3 is empty everywhere, query 4 is non-empty in frame 0 only."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    frames = torch.stack([torch.stack([127.5 + 100 * torch.sin(xx / (17 + 3 * c + t) + c) * torch.cos(yy / (23 + 2 * c) + t)
                                       for c in range(3)]) for t in range(T)])
    frames = (frames + 12 * torch.randn(T, 3, H, W, generator=g)).clamp(0, 255).round()
    logits = torch.full((N, T, H, W), -6.0)
    for t in range(T):
        logits[0, t, 10 + t:H - 14, 8:W - 20 - t] = 5.0
        logits[1, t, 40:90 + 5 * t, W - 36:W - 2] = 4.0
        e = ((yy - H / 2 - 9 * t) / 60) ** 2 + ((xx - W / 3) / 35) ** 2
        logits[2, t] = 4.0 * (1.0 - e)
    if N > 4:
        logits[4, 0, H - 50:H - 20, 30:75] = 3.0
    logits = logits + 0.3 * torch.randn(N, T, H, W, generator=g)
    logits[3] = -6.0
    if N > 4:
        logits[4, 1:] = -6.0
    return frames, logits


def seeded_params(shapes, seed=0):
    """Deterministic weights that do not need the reference to regenerate: names in sorted order, one
    torch.Generator.  Scales mimic the reference's inits (xavier-like for matrices, N(0,1) embeddings,
    LayerNorm weight near 1) so that activations / mask densities look like a random-init reference."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias") or name.endswith("in_proj_bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif name in ("query_feat.weight", "query_embed.weight", "level_embed.weight"):
            t = torch.randn(shp, generator=g)
        else:
            fan_out, fan_in = shp[0], shp[1]
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        out[name] = t
    return out


def seeded_inputs(T, Hp, Wp, C=256, seed=1234):
    """N(0,1) multi-scale features (coarsest first) and mask features."""
    g = torch.Generator().manual_seed(seed)
    x = [torch.randn(T, C, Hp // 32 * 2 ** l, Wp // 32 * 2 ** l, generator=g) for l in range(3)]
    mf = torch.randn(T, C, Hp // 4, Wp // 4, generator=g)
    return x, mf


def resampler_param_shapes(C=256, F_=2048, L=6):
    """state_dict contract of TemporalInstanceResampler (openvis/modeling/resampler.py:191-236)."""
    s = {}
    for i in range(L):
        pre = f"long_aggregate_layers.{i}"
        s[f"{pre}.self_attn.in_proj_weight"] = (3 * C, C)
        s[f"{pre}.self_attn.in_proj_bias"] = (3 * C,)
        s[f"{pre}.self_attn.out_proj.weight"] = (C, C)
        s[f"{pre}.self_attn.out_proj.bias"] = (C,)
        s[f"{pre}.norm.weight"] = (C,)
        s[f"{pre}.norm.bias"] = (C,)
        s[f"short_aggregate_layers.{i}.0.weight"] = (C, C, 5)
        s[f"short_aggregate_layers.{i}.0.bias"] = (C,)
        s[f"short_aggregate_layers.{i}.2.weight"] = (C, C, 3)
        s[f"short_aggregate_layers.{i}.2.bias"] = (C,)
        s[f"aggregate_norms.{i}.weight"] = (C,)
        s[f"aggregate_norms.{i}.bias"] = (C,)
        pre = f"transformer_ffn_layers.{i}"
        s[f"{pre}.linear1.weight"] = (F_, C)
        s[f"{pre}.linear1.bias"] = (F_,)
        s[f"{pre}.linear2.weight"] = (C, F_)
        s[f"{pre}.linear2.bias"] = (C,)
        s[f"{pre}.norm.weight"] = (C,)
        s[f"{pre}.norm.bias"] = (C,)
    s["decode_norm.weight"] = (C,)
    s["decode_norm.bias"] = (C,)
    for name in ("attn_embed", "mask_embed"):
        for i in range(3):
            s[f"{name}.layers.{i}.weight"] = (C, C)
            s[f"{name}.layers.{i}.bias"] = (C,)
    return s


def seeded_resampler_params(seed=0, **kw):
    """Deterministic resampler weights: xavier-like matrices / conv taps, small biases, LayerNorm weight near 1."""
    g = torch.Generator().manual_seed(seed)
    shapes = resampler_param_shapes(**kw)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) == 1 and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:
            t = 0.05 * torch.randn(shp, generator=g)
        else:
            taps = shp[2] if len(shp) == 3 else 1
            bound = math.sqrt(6.0 / ((shp[0] + shp[1]) * taps))
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        out[name] = t
    return out


def pixel_decoder_param_shapes(in_channels=(256, 512, 1024, 2048), C=256, F_=1024, L=6, M=8, P=4):
    """state_dict names / shapes of the reference MSDeformAttnPixelDecoder (pixel_decoder/msdeformattn.py:182-306) with
    res2..res5 inputs, the three coarsest fed to the deformable encoder and one FPN level (res2), norm "GN"."""
    s = {}
    for i, cin in enumerate(in_channels[:0:-1]):              # res5, res4, res3
        s[f"input_proj.{i}.0.weight"] = (C, cin, 1, 1)
        s[f"input_proj.{i}.0.bias"] = (C,)
        s[f"input_proj.{i}.1.weight"] = (C,)
        s[f"input_proj.{i}.1.bias"] = (C,)
    s["transformer.level_embed"] = (3, C)
    for i in range(L):
        pre = f"transformer.encoder.layers.{i}"
        s[f"{pre}.self_attn.sampling_offsets.weight"] = (M * 3 * P * 2, C)
        s[f"{pre}.self_attn.sampling_offsets.bias"] = (M * 3 * P * 2,)
        s[f"{pre}.self_attn.attention_weights.weight"] = (M * 3 * P, C)
        s[f"{pre}.self_attn.attention_weights.bias"] = (M * 3 * P,)
        for n in ("value_proj", "output_proj"):
            s[f"{pre}.self_attn.{n}.weight"] = (C, C)
            s[f"{pre}.self_attn.{n}.bias"] = (C,)
        s[f"{pre}.linear1.weight"] = (F_, C)
        s[f"{pre}.linear1.bias"] = (F_,)
        s[f"{pre}.linear2.weight"] = (C, F_)
        s[f"{pre}.linear2.bias"] = (C,)
        for n in ("norm1", "norm2"):
            s[f"{pre}.{n}.weight"] = (C,)
            s[f"{pre}.{n}.bias"] = (C,)
    s["mask_features.weight"] = (C, C, 1, 1)
    s["mask_features.bias"] = (C,)
    s["adapter_1.weight"] = (C, in_channels[0], 1, 1)
    s["adapter_1.norm.weight"] = (C,)
    s["adapter_1.norm.bias"] = (C,)
    s["layer_1.weight"] = (C, C, 3, 3)
    s["layer_1.norm.weight"] = (C,)
    s["layer_1.norm.bias"] = (C,)
    return s


def seeded_pixel_decoder_params(seed=0, **kw):
    """Deterministic pixel-decoder weights: xavier-like matrices / convolutions (fan = channels x kernel area), norm weights
    near 1, sampling-offset biases of a pixel or two (the reference initialises them to a ring of unit steps)."""
    g = torch.Generator().manual_seed(seed)
    shapes = pixel_decoder_param_shapes(**kw)
    out = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) == 1 and (".norm" in name or name.endswith(".1.weight") or name.endswith(".1.bias")):
            t = (1.0 if name.endswith("weight") else 0.0) + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("sampling_offsets.bias"):
            t = 1.5 * torch.randn(shp, generator=g)
        elif name.endswith("sampling_offsets.weight"):
            t = 0.03 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shp, generator=g)
        elif name == "transformer.level_embed":
            t = torch.randn(shp, generator=g)
        else:
            area = shp[2] * shp[3] if len(shp) == 4 else 1
            bound = math.sqrt(6.0 / ((shp[0] + shp[1]) * area))
            t = (torch.rand(shp, generator=g) * 2 - 1) * bound
        out[name] = t
    return out


def seeded_backbone_features(T, Hp, Wp, in_channels=(256, 512, 1024, 2048), seed=77):
    """N(0,1) stand-ins for the backbone's res2..res5 maps (strides 4 / 8 / 16 / 32) of a [T, 3, Hp, Wp] batch."""
    g = torch.Generator().manual_seed(seed)
    return {f"res{i + 2}": torch.randn(T, c, Hp // (4 << i), Wp // (4 << i), generator=g) for i, c in enumerate(in_channels)}
