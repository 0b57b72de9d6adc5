"""TEST INFRASTRUCTURE -- CPU restatement (torch fp32) of the OpenVIS crop classifier.

Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.

Pinned (tests/test_oracle_golden.py, oracle/make_golden.py:make_clip_adapter_fixture) against the reference's own
ClipAdapter._preprocess_image / encode_image / cal_sim_logits (openvis/modeling/clip_adapter/adapter.py:73-147, loaded
through oracle/ref_shim.py with torchvision's roi_align and the vendored mask_adapted_clip VisionTransformer) and
OpenVIS.open_vocabulary_inference (openvis/openvis.py:110-147); the committed fixture is tests/golden/clip_adapter.npz.
detectron2 is not installed here: BitMasks.get_bounding_boxes is restated from its published definition
(detectron2/structures/masks.py: boxes = [x_min, y_min, x_max + 1, y_max + 1] of the non-zero pixels, zeros when empty)."""
import math

import torch

from . import decoder_ref as O

PIXEL_MEAN = (0.48145466, 0.4578275, 0.40821073)     # adapter.py:20-21
PIXEL_STD = (0.26862954, 0.26130258, 0.27577711)


def mask_boxes(bin_masks):
    """bin_masks [M, H, W] bool -> float boxes [M, 4] (detectron2 BitMasks.get_bounding_boxes)."""
    boxes = torch.zeros(bin_masks.shape[0], 4)
    x_any, y_any = bin_masks.any(dim=1), bin_masks.any(dim=2)
    for i in range(bin_masks.shape[0]):
        x, y = torch.where(x_any[i])[0], torch.where(y_any[i])[0]
        if len(x) > 0 and len(y) > 0:
            boxes[i] = torch.tensor([x[0], y[0], x[-1] + 1, y[-1] + 1], dtype=torch.float32)
    return boxes


def roi_align(inp, rois, R):
    """torchvision.ops.roi_align(inp, rois, (R, R)) with its defaults spatial_scale = 1, sampling_ratio = -1, aligned = False
    (torchvision/csrc/ops/cpu/roi_align_kernel.cpp): inp [B, C, H, W], rois [K, 5] = (batch index, x1, y1, x2, y2)."""
    B, C, H, W = inp.shape
    out = torch.zeros(rois.shape[0], C, R, R)
    for k in range(rois.shape[0]):
        bi, x1, y1, x2, y2 = [float(v) for v in rois[k]]
        roi_w, roi_h = max(x2 - x1, 1.0), max(y2 - y1, 1.0)
        bw, bh = roi_w / R, roi_h / R
        gh, gw = math.ceil(roi_h / R), math.ceil(roi_w / R)
        # sample coordinates of every (bin, sample): [R, g]
        ys = y1 + torch.arange(R, dtype=torch.float32)[:, None] * bh + (torch.arange(gh, dtype=torch.float32)[None] + 0.5) * bh / gh
        xs = x1 + torch.arange(R, dtype=torch.float32)[:, None] * bw + (torch.arange(gw, dtype=torch.float32)[None] + 0.5) * bw / gw
        ys, xs = ys.reshape(-1), xs.reshape(-1)

        def axis(v, n):
            ok = (v >= -1.0) & (v <= n)
            v = v.clamp(min=0.0)
            lo = v.floor().long()
            top = lo >= n - 1
            lo = torch.where(top, torch.full_like(lo, n - 1), lo)
            hi = torch.where(top, lo, lo + 1)
            v = torch.where(top, lo.float(), v)
            return ok, lo, hi, v - lo.float()

        oky, yl, yh, ly = axis(ys, H)
        okx, xl, xh, lx = axis(xs, W)
        img = inp[int(bi)]
        v = ((1 - ly)[:, None] * (1 - lx)[None] * img[:, yl][:, :, xl] + (1 - ly)[:, None] * lx[None] * img[:, yl][:, :, xh] +
             ly[:, None] * (1 - lx)[None] * img[:, yh][:, :, xl] + ly[:, None] * lx[None] * img[:, yh][:, :, xh])
        v = v * (oky[:, None] & okx[None]).float()
        out[k] = v.reshape(C, R, gh, R, gw).sum(dim=(2, 4)) / max(gh * gw, 1)
    return out


def preprocess_image(frames, masks, R=224, half_io=True):
    """ClipAdapter._preprocess_image (adapter.py:73-116): frames [T, 3, H, W], masks [T, N, H, W] soft masks ->
    (regions [M, 3, R, R] or None, valid [T, N], sboxes [M, 4]).  half_io: inputs and both roi_align outputs rounded to
    fp16 as the reference's `.half()` tensors are (the arithmetic in between stays fp32)."""
    h = (lambda t: t.half().float()) if half_io else (lambda t: t)
    bin_masks = masks > 0.5
    valid = bin_masks.sum(dim=(-1, -2)) > 0
    if valid.sum() == 0:
        return None, valid, None
    boxes = mask_boxes(bin_masks[valid])
    s = boxes.clone()
    s[:, 2] = s[:, 2] - s[:, 0]
    s[:, 3] = s[:, 3] - s[:, 1]
    s[:, 3] = s[:, 2] = torch.max(s[:, 2], s[:, 3])
    s[:, 2] = s[:, 0] + s[:, 2]
    s[:, 3] = s[:, 1] + s[:, 3]
    ids = torch.nonzero(valid)
    regions = h(roi_align(h(frames), torch.cat([ids[:, 0:1].float(), s], dim=-1), R))
    mask_regions = h(roi_align(h(masks[valid])[:, None], torch.cat([torch.arange(len(s))[:, None].float(), s], dim=-1), R))
    out = mask_regions * regions
    return (h(out) if half_io else out), valid, s


def clip_visual(P, image, nheads=12):
    """VisionTransformer.forward with m = None (mask_adapted_clip model.py:327-362): image [M, 3, R, R] normalised."""
    x = torch.nn.functional.conv2d(image, P["conv1.weight"], stride=P["conv1.weight"].shape[-1])
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([P["class_embedding"] + torch.zeros(x.shape[0], 1, x.shape[-1]), x], dim=1) + P["positional_embedding"]
    x = O.layer_norm(x, P["ln_pre.weight"], P["ln_pre.bias"]).permute(1, 0, 2)
    blocks = {k[len("transformer.resblocks."):]: v for k, v in P.items() if k.startswith("transformer.resblocks.")}
    for i in range(1 + max(int(k.split(".")[0]) for k in blocks)):
        x = O.clip_block(blocks, i, x, None, nheads)
    x = O.layer_norm(x.permute(1, 0, 2)[:, 0], P["ln_post.weight"], P["ln_post.bias"])
    return x @ P["proj"]


def encode_image(P, regions):
    """ClipAdapter.encode_image (adapter.py:140-144); the bicubic resize is the identity for R x R regions."""
    R = regions.shape[-1]
    image = torch.nn.functional.interpolate(regions / 255.0, (R, R), mode="bicubic")
    mean, std = torch.tensor(PIXEL_MEAN).view(1, 3, 1, 1), torch.tensor(PIXEL_STD).view(1, 3, 1, 1)
    f = clip_visual(P, (image - mean) / std)
    return f / f.norm(dim=-1, keepdim=True)


def clip_adapter_forward(P, frames, text, masks, R=224):
    """ClipAdapter.forward (adapter.py:56-71) with the text matrix given: (sim_logits [M, K] or None, valid [T, N])."""
    regions, valid, _ = preprocess_image(frames, masks, R)
    if regions is None:
        return None, valid
    return 100.0 * encode_image(P, regions) @ text.T, valid


def open_vocabulary_inference(P, mask_logits, frames, text, part_len=5, R=224):
    """OpenVIS.open_vocabulary_inference (openvis.py:110-147): mask_logits [N, T, H, W] -> (probs [N_valid, K], valid_query [N])."""
    clip_cls, valid_flag = [], []
    for idx in range(0, len(frames), part_len):
        part_masks = mask_logits[:, idx:idx + part_len].sigmoid().transpose(0, 1).contiguous()
        sim, valid = clip_adapter_forward(P, frames[idx:idx + part_len], text, part_masks, R)
        clip_cls.append(sim if sim is not None else torch.empty(0, text.shape[0]))
        valid_flag.append(valid)
    clip_cls, valid_flag = torch.cat(clip_cls), torch.cat(valid_flag)
    if valid_flag.sum() == 0:
        return None, valid_flag.sum(dim=0) > 0
    probs_all, qvalid = O.openvis_clip_aggregate(clip_cls, valid_flag)
    return probs_all, qvalid
