"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32) of the reference's pixel decoder:
MSDeformAttnTransformerEncoderOnly and MSDeformAttnPixelDecoder.forward_features
(/root/reference/openvis/modeling/pixel_decoder/msdeformattn.py).  Nothing under openvis_b200/ imports this file; it is the
checker of tests/ only.  Pinned against tests/golden/pixel_decoder.npz = outputs of the reference's own classes
(oracle/make_golden.make_pixel_decoder_fixture) and live against them when /root/reference is mounted.

P = the reference module's state dict (names of MSDeformAttnPixelDecoder, see synthetic.pixel_decoder_param_shapes).
"""
import torch
import torch.nn.functional as F

from .decoder_ref import layer_norm, ms_deform_attn_module, sine_pos_2d


def encoder_reference_points(shapes):
    """MSDeformAttnTransformerEncoder.get_reference_points (msdeformattn.py:155-169) with all valid ratios = 1 (the encoder
    is fed all-False padding masks, :77): pixel centres in [0, 1], the same point for every level -> [1, S, L, 2] (x, y)."""
    pts = []
    for (h, w) in shapes:
        y = (torch.arange(h, dtype=torch.float32) + 0.5) / h
        x = (torch.arange(w, dtype=torch.float32) + 0.5) / w
        pts.append(torch.stack([x[None, :].expand(h, w), y[:, None].expand(h, w)], -1).reshape(h * w, 2))
    return torch.cat(pts, 0)[None, :, None, :].expand(1, -1, len(shapes), 2)


def encoder_layer(P, pre, src, pos, ref, shapes):
    """MSDeformAttnTransformerEncoderLayer.forward (msdeformattn.py:136-146), eval mode (dropout = identity)."""
    A = {k[len(pre) + len(".self_attn."):]: v for k, v in P.items() if k.startswith(pre + ".self_attn.")}
    src2 = ms_deform_attn_module(A, src + pos, ref.expand(src.shape[0], -1, -1, -1), src, shapes)
    src = layer_norm(src + src2, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    h = torch.relu(src @ P[pre + ".linear1.weight"].T + P[pre + ".linear1.bias"])
    src2 = h @ P[pre + ".linear2.weight"].T + P[pre + ".linear2.bias"]
    return layer_norm(src + src2, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])


def encoder_only(P, srcs, pos_embeds, pre="transformer"):
    """MSDeformAttnTransformerEncoderOnly.forward (msdeformattn.py:76-104): srcs / pos_embeds lists of [B, C, h, w]
    -> memory [B, S, C], shapes [L, 2]."""
    shapes = torch.tensor([tuple(s.shape[-2:]) for s in srcs])
    src = torch.cat([s.flatten(2).transpose(1, 2) for s in srcs], 1)
    pos = torch.cat([p.flatten(2).transpose(1, 2) + P[pre + ".level_embed"][l].view(1, 1, -1)
                     for l, p in enumerate(pos_embeds)], 1)
    ref = encoder_reference_points([tuple(int(v) for v in s) for s in shapes])
    n_layers = 1 + max(int(k.split(".layers.")[1].split(".")[0]) for k in P if k.startswith(pre + ".encoder.layers."))
    for i in range(n_layers):
        src = encoder_layer(P, f"{pre}.encoder.layers.{i}", src, pos, ref, shapes)
    return src, shapes


def group_norm(x, w, b, groups=32, eps=1e-5):
    """nn.GroupNorm(32, C) on [B, C, h, w]: statistics over (C / 32 channels x h x w) per sample, biased variance."""
    B, C = x.shape[:2]
    g = x.reshape(B, groups, -1)
    mu = g.mean(-1, keepdim=True)
    var = ((g - mu) ** 2).mean(-1, keepdim=True)
    y = ((g - mu) / torch.sqrt(var + eps)).reshape(x.shape)
    return y * w.view(1, C, 1, 1) + b.view(1, C, 1, 1)


def pixel_decoder_forward(P, features, extra_features=None):
    """MSDeformAttnPixelDecoder.forward_features (msdeformattn.py:329-380) for res2..res5 inputs (encoder on res3..res5, one
    FPN level): returns (mask_features [B, C, H/4, W/4], out[0] (stride 32), multi_scale_features (strides 32, 16, 8))."""
    srcs, pos = [], []
    for i, name in enumerate(("res5", "res4", "res3")):
        x = features[name].float()
        y = F.conv2d(x, P[f"input_proj.{i}.0.weight"], P[f"input_proj.{i}.0.bias"])
        srcs.append(group_norm(y, P[f"input_proj.{i}.1.weight"], P[f"input_proj.{i}.1.bias"]))
        h, w = x.shape[-2:]
        if extra_features is not None:                         # msdeformattn.py:338-343
            ex = extra_features[i]
            if tuple(ex.shape[-2:]) != (h, w):
                ex = F.interpolate(ex, size=(h, w), mode="bilinear", align_corners=False)
            srcs[-1] = srcs[-1] + ex
        pos.append(sine_pos_2d(h, w)[None].expand(x.shape[0], -1, -1, -1))
    mem, shapes = encoder_only(P, srcs, pos)
    B = mem.shape[0]
    out, start = [], 0
    for (h, w) in shapes.tolist():
        out.append(mem[:, start:start + h * w].transpose(1, 2).reshape(B, -1, h, w))
        start += h * w
    x = features["res2"].float()
    cur = group_norm(F.conv2d(x, P["adapter_1.weight"]), P["adapter_1.norm.weight"], P["adapter_1.norm.bias"])
    y = cur + F.interpolate(out[-1], size=cur.shape[-2:], mode="bilinear", align_corners=False)
    y = torch.relu(group_norm(F.conv2d(y, P["layer_1.weight"], padding=1), P["layer_1.norm.weight"], P["layer_1.norm.bias"]))
    mask_features = F.conv2d(y, P["mask_features.weight"], P["mask_features.bias"])
    return mask_features, out[0], out[:3]
