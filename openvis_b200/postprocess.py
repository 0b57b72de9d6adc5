"""Device-side post-processing: the tail of the reference's video inference on the B200
kernels -- VideoMaskFormer.postprocess + inference_video (openvis/modeling/video_maskformer.py:215-229, 262-298; the
same up-sampling sits in OpenVIS.forward, openvis/openvis.py:87-96).

The reference up-samples ALL queries' stride-4 mask logits to the padded input size (fp32 [Q, T, Hp, Wp]), picks the
top-10 (query, class) pairs, crops, resizes to the output size, thresholds and copies one byte per pixel to the host.
Here the top-10 selection runs first and one kernel evaluates the composed interpolation for the selected queries only,
writing one BIT per output pixel.  No CPU path.
"""
import torch

from . import _lib as L


class PackedMasks:
    """Bit-packed boolean masks [n, T, H, W]: word x // 32 of a row holds pixel x in bit x % 32."""

    def __init__(self, bits: torch.Tensor, width: int):
        self.bits, self.width = bits, width

    @property
    def shape(self):
        n, t, h, _ = self.bits.shape
        return (n, t, h, self.width)

    def cpu(self):
        """Device -> host through pinned memory (torch's caching host allocator: the block is recycled once the caller
        drops the result); a pageable copy of the 41 MB of a 36-frame 720x1280 clip costs 20 ms, this one 2 ms."""
        if not self.bits.is_cuda:
            return self
        host = torch.empty(self.bits.shape, dtype=self.bits.dtype, pin_memory=True)
        host.copy_(self.bits, non_blocking=True)
        torch.cuda.current_stream(self.bits.device).synchronize()
        return PackedMasks(host, self.width)

    def unpack(self) -> torch.Tensor:
        """-> bool [n, T, H, W] (on the tensor's device; plain torch, for consumers that want the reference's format)."""
        b = self.bits
        sh = torch.arange(32, device=b.device, dtype=torch.int32)
        m = ((b[..., None] >> sh) & 1).bool().flatten(-2)
        return m[..., : self.width]


@torch.no_grad()
def inference_video(num_queries, num_classes, pred_cls, pred_masks, padded_size, img_size, output_height, output_width,
                    topk=10, to_host=True, out_bits=None):
    """Mirror of ``VideoMaskFormer.inference_video`` (video_maskformer.py:262-298) fed with the decoder's own outputs.

    pred_cls   [Q, num_classes] fp32 scores (softmax(...)[:, :-1] or the open-vocabulary scores), on the GPU
    pred_masks [Q, T, H/4, W/4] fp32 stride-4 mask logits (``outputs["pred_masks"][0]``) -- NOT up-sampled
    padded_size (Hp, Wp) of the network input (what ``postprocess`` up-samples to), img_size the un-padded size,
    (output_height, output_width) the size of the original frames.
    Returns the reference's dictionary; ``pred_masks`` is a ``PackedMasks`` on the host (``.unpack()`` gives the bool
    tensor), entries ordered by descending score.
    to_host=False (throughput callers that batch their device-to-host traffic): nothing is copied or synchronised, the
    scores / labels / entropies / query ids stay device tensors and ``pred_masks`` is a ``PackedMasks`` on the device
    (written into `out_bits` [topk, T, out_h, ceil(out_w/32)] int32 when given)."""
    assert pred_cls.shape == (num_queries, num_classes)
    if pred_cls.numel() == 0:
        return {"image_size": (output_height, output_width), "pred_entropys": [], "pred_scores": [], "pred_labels": [],
                "pred_masks": []}
    if not (pred_cls.is_cuda and pred_masks.is_cuda):
        raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
    with torch.cuda.device(pred_masks.device):
        k = min(topk, pred_cls.numel())
        scores, qidx, labels, ent = L.topk_scores(pred_cls.float().contiguous(), k)
        bits = L.mask_postprocess(pred_masks.float(), qidx, padded_size, img_size,
                                  (output_height, output_width), out=out_bits)
        if not to_host:
            return {"image_size": (output_height, output_width), "pred_entropys": ent, "pred_scores": scores,
                    "pred_labels": labels, "pred_masks": PackedMasks(bits, output_width), "pred_queries": qidx}
        packed = PackedMasks(bits, output_width).cpu()
    return {"image_size": (output_height, output_width), "pred_entropys": ent.tolist(), "pred_scores": scores.tolist(),
            "pred_labels": labels.tolist(), "pred_masks": packed, "pred_queries": qidx.tolist()}
