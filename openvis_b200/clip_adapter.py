"""OpenVIS crop classifier: the reference's ``ClipAdapter`` (openvis/modeling/clip_adapter/
adapter.py:34-147) and ``OpenVIS.open_vocabulary_inference`` (openvis/openvis.py:110-147) over the C ABI.

  ClipVisualEncoder   CLIP VisionTransformer.forward (third_parties/mask_adapted_clip/.../model.py:327-362, m = None):
                      conv1 as a patch GEMM, [class | patches] + positional embedding, ln_pre, the residual attention blocks,
                      ln_post on the class token, @ proj
  ClipAdapter         forward / _preprocess_image / encode_image / normalize / cal_sim_logits with the reference's names
  open_vocabulary_inference   crops of every (frame, query) with a non-empty mask -> CLIP logits -> per-query mean over the
                      valid frames -> softmax

fp16 GEMM operands, fp32 accumulation; LayerNorm / softmax statistics / residual stream in fp32.  No CPU path."""
from typing import List

import torch

from . import _lib as L
from .ov_head import ClipLogitHead, SideAdapterBlocks

PIXEL_MEAN = (0.48145466, 0.4578275, 0.40821073)     # adapter.py:20-21
PIXEL_STD = (0.26862954, 0.26130258, 0.27577711)


class ClipVisualEncoder(SideAdapterBlocks):
    """``clip_model.visual`` of the reference (ViT-B/16 by default).  ``load_state_dict(clip_model.visual.state_dict())``
    reads conv1.weight, class_embedding, positional_embedding, ln_pre.*, transformer.resblocks.*, ln_post.*, proj."""

    def __init__(self, input_resolution=224, patch_size=16, width=768, layers=12, heads=12, output_dim=512):
        super().__init__(num_queries=0, broken_idx=0, num_layers=layers, width=width, heads=heads)
        self.input_resolution, self.patch_size, self.output_dim = input_resolution, patch_size, output_dim
        self.grid = input_resolution // patch_size
        self.max_images_per_pass = 256              # activations: ~3.3 MB per 224 x 224 image

    def load_state_dict(self, sd):
        self.load_clip_visual_state_dict(sd)
        f32 = lambda t: t.detach().float().contiguous().cuda()
        W = self._w
        W["conv1"] = L.cast_f16(f32(sd["conv1.weight"]).reshape(self.width, -1).contiguous())      # [width, 3 * P * P]
        W["cls"] = f32(sd["class_embedding"])
        W["pos"] = f32(sd["positional_embedding"])
        W["ln_pre"] = (f32(sd["ln_pre.weight"]), f32(sd["ln_pre.bias"]))
        if tuple(W["pos"].shape) != (1 + self.grid * self.grid, self.width):
            raise ValueError(f"positional_embedding {tuple(W['pos'].shape)} does not match a {self.grid} x {self.grid} grid")
        return self

    @torch.no_grad()
    def tokens(self, image):
        """image [M, 3, R, R] in 0..255 (fp16, or fp32 rounded to fp16 like the reference's half regions) -> token matrix
        [M * (1 + grid^2), width] fp32 after the last block."""
        if self._w is None or "conv1" not in self._w:
            raise RuntimeError("ClipVisualEncoder: load_state_dict() first")
        if not image.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        R = self.input_resolution
        if tuple(image.shape[1:]) != (3, R, R):
            raise NotImplementedError(f"regions must already be 3 x {R} x {R} (the bicubic resize of adapter.py:141 is the identity "
                                      f"there); got {tuple(image.shape)}")
        W = self._w
        M, Lp = image.shape[0], self.grid * self.grid
        with torch.cuda.device(image.device):
            a16 = L.clip_patchify(image.half().contiguous(), self.patch_size, PIXEL_MEAN, PIXEL_STD)
            pt = L.linear_f16(a16, W["conv1"], None, out_f32=True)                 # conv1 has no bias (model.py:302-308)
            X = L.clip_embed(pt, W["cls"], W["pos"], W["ln_pre"][0], W["ln_pre"][1], M, Lp)
            return self._run_blocks(X, M, 0, Lp, None)

    @torch.no_grad()
    def forward(self, image):
        """``clip_model.visual(image)`` after the Normalize of adapter.py:142: features [M, output_dim] fp32 (not normalised)."""
        outs = []
        for m0 in range(0, image.shape[0], self.max_images_per_pass):
            part = image[m0:m0 + self.max_images_per_pass]
            X = self.tokens(part)
            M = part.shape[0]
            with torch.cuda.device(image.device):
                cls_rows = X.view(M, -1, self.width)[:, 0].contiguous()
                _, x16 = L.rownorm(cls_rows, self._w["ln_post"][0], self._w["ln_post"][1], layer_norm=True, want32=False)
                outs.append(L.linear_f16(x16, self.tail._proj_f16(self._w["proj"]), None, out_f32=True))
        return torch.cat(outs) if len(outs) != 1 else outs[0]

    __call__ = forward


class ClipAdapter(ClipLogitHead):
    """Drop-in for the reference ``ClipAdapter`` at inference (adapter.py:34-147).  The text tower is not part of the path:
    ``set_text_embeddings(names, embeds)`` fills the cache ``encode_text`` reads (adapter.py:47, 121-138), or pass the
    [K, D] text matrix itself as ``text``."""

    def __init__(self, visual: ClipVisualEncoder):
        super().__init__()
        self.visual = visual
        self.input_resolution = visual.input_resolution

    def _text(self, text):
        return text if torch.is_tensor(text) else self.encode_text(list(text))

    @torch.no_grad()
    def _preprocess_image(self, frames: torch.Tensor, masks: torch.Tensor, layout="tn", logits=False):
        """adapter.py:73-116: (regions [M, 3, R, R] fp16, valid [T, N]); (None, valid) when no mask is non-empty."""
        frames = frames.float().contiguous()
        masks = masks.to(frames.device).float()
        if masks.stride(3) != 1 or masks.stride(2) != masks.shape[3]:
            masks = masks.contiguous()
        with torch.cuda.device(frames.device):
            valid, boxes = L.mask_boxes(masks, 0.5, layout=layout, logits=logits)
            ids = torch.nonzero(valid).to(torch.int32).contiguous()       # [M, 2] (frame, query), row-major like adapter.py:103
            if ids.shape[0] == 0:
                return None, valid
            return L.crop_blend(frames, masks, ids, boxes.contiguous(), self.input_resolution, layout=layout, logits=logits), valid

    @torch.no_grad()
    def encode_image(self, image: torch.Tensor):
        """adapter.py:140-144: / 255, resize (identity at R x R), Normalize, visual tower, normalize."""
        return self.normalize(self.visual(image))

    @torch.no_grad()
    def forward(self, frames: torch.Tensor, text, masks: torch.Tensor, layout="tn", logits=False):
        """frames [T, 3, H, W] (0..255), masks [T, N, H, W] soft masks -> (sim_logits [M, K] or None, valid_flag [T, N])."""
        regions, valid = self._preprocess_image(frames, masks, layout, logits)
        if regions is None:
            return None, valid
        text_feature = self._text(text)
        feats = self.visual(regions)
        # normalize + cal_sim_logits fused (the GEMM's epilogue divides each row by its norm)
        return self.cal_sim_logits(text_feature, feats, 100, normalized=False), valid

    __call__ = forward

    @torch.no_grad()
    def open_vocabulary_inference(self, scores, masks: torch.Tensor, frames: torch.Tensor, class_names, part_len: int = 5):
        """OpenVIS.open_vocabulary_inference (openvis.py:110-147).  masks [N, T, H, W] mask LOGITS at the frames' resolution
        (the reference applies the sigmoid and transposes per part of 5 frames; here the kernels read the logits in place),
        frames [T, 3, H, W].  Returns (probs [N_valid, K], masks[valid_query]) or ([], [])."""
        if len(scores) == 0:
            return [], []
        text = self._text(class_names)
        N, T = masks.shape[:2]
        K = text.shape[0]
        frames = frames.float().contiguous()
        masks_c = masks.float()
        logits_all = torch.zeros(T, N, K, dtype=torch.float32, device=frames.device)
        valid_all = torch.zeros(T, N, dtype=torch.bool, device=frames.device)
        for idx in range(0, T, part_len):            # parts bound the number of crops in flight, like the reference
            sim, valid = self.forward(frames[idx:idx + part_len], text, masks_c[:, idx:idx + part_len], layout="nt", logits=True)
            valid_all[idx:idx + part_len] = valid
            if sim is not None:
                logits_all[idx:idx + part_len][valid] = sim
        if not bool(valid_all.any()):
            return [], []
        probs, qvalid = L.clip_aggregate(logits_all, valid_all)
        return probs[qvalid], masks[qvalid]
