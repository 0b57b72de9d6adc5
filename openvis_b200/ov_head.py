"""Open-vocabulary classifier tails on the B200 kernels (host-side mirror of the reference adapters' arithmetic).

ClipLogitHead     -- ClipAdapter.normalize / cal_sim_logits / text cache (clip_adapter/adapter.py:118-138, 146-147) and
                     OpenVIS.open_vocabulary_inference's per-query aggregation (openvis/openvis.py:123-141).
SideAdapterTail   -- SideAdapter._build_attn_biases, the ln_post/proj/normalize tail of post_encode_image and
                     cal_sim_logits (clip_adapter/side_adapter.py:201-207, 234-270).

The frozen CLIP towers themselves (text encoder, ViT blocks, crop/roi_align preprocessing) are out of scope
(SURVEY.md section 8): text embeddings are supplied already encoded ("cached text embeddings", north_star) and
region / SOS-token features come from the caller.
"""
from typing import List

import torch

from . import _lib as L


class _TextCache:
    """Python-dict text cache keyed by class string, like the reference (adapter.py:47,122,134)."""

    def __init__(self):
        self.text_cache = {}
        self._mat = {}

    def set_text_embeddings(self, names: List[str], embeds: torch.Tensor):
        """Registers already-encoded, L2-normalised text embeddings [K, D] for `names`."""
        assert embeds.shape[0] == len(names)
        self.text_cache.update(dict(zip(names, embeds)))
        self._mat.clear()

    def encode_text(self, noun_list: List[str]) -> torch.Tensor:
        missing = [w for w in noun_list if w not in self.text_cache]
        if missing:
            raise KeyError(f"no cached text embedding for {missing[:3]}...: the CLIP text encoder is out of scope; "
                           "register embeddings with set_text_embeddings()")
        return torch.stack([self.text_cache[w] for w in noun_list])

    def _text_f16(self, text):
        key = (text.data_ptr(), text._version, tuple(text.shape))
        hit = self._mat.get(key)
        if hit is None:
            if len(self._mat) > 8:
                self._mat.clear()
            hit = L.cast_f16(text.detach().float().contiguous())
            self._mat[key] = hit
        return hit


class ClipLogitHead(_TextCache):
    def normalize(self, feat: torch.Tensor):
        """feat / feat.norm(dim=-1, keepdim=True) (adapter.py:118-119)."""
        shp = feat.shape
        o32, _ = L.rownorm(feat.reshape(-1, shp[-1]).float().contiguous(), l2=True, want16=False)
        return o32.view(shp)

    def cal_sim_logits(self, text_features: torch.Tensor, image_features: torch.Tensor, temperature: float = 100,
                       normalized: bool = True):
        """temperature * image_features @ text_features.T (adapter.py:146-147); fp16 operands, fp32 accumulate.
        With normalized=False the L2 normalisation of the image features is fused in front."""
        shp = image_features.shape
        f = image_features.reshape(-1, shp[-1]).float().contiguous()
        if normalized:
            f16 = L.cast_f16(f)
        else:
            _, f16 = L.rownorm(f, l2=True, want32=False)
        out = L.linear_f16(f16, self._text_f16(text_features), None, scale=float(temperature), out_f32=True)
        return out.view(*shp[:-1], text_features.shape[0])

    def open_vocabulary_scores(self, region_feats: torch.Tensor, valid: torch.Tensor, text_features: torch.Tensor):
        """region_feats [T, Q, D] (un-normalised CLIP features of the masked crops; rows of invalid regions are ignored),
        valid [T, Q].  Returns (probs [Q, K] with zero rows for queries without a valid frame, valid_query [Q])."""
        logits = self.cal_sim_logits(text_features, region_feats, 100, normalized=False)
        return L.clip_aggregate(logits.contiguous(), valid)


class SideAdapterTail(_TextCache):
    def __init__(self, grid_size=14, logit_scale_exp=1.0 / 0.07):
        super().__init__()
        self.grid_size = grid_size
        self.logit_scale_exp = logit_scale_exp

    def build_attn_biases(self, attn_bias: torch.Tensor, num_layers: int = 3, target_shape=None):
        """attn_bias [B, n, Q, h, w] -> list of num_layers references to one [B*n, Q+1+L, Q+1+L] matrix
        (side_adapter.py:237-270; the same tensor is reused for every block, :268-269)."""
        gs = target_shape or (self.grid_size, self.grid_size)
        m = L.san_attn_bias(attn_bias.float().contiguous(), gs)
        return [m for _ in range(num_layers)]

    def sos_tail(self, sos_token: torch.Tensor, ln_w, ln_b, proj):
        """ln_post -> @ visual.proj -> F.normalize (side_adapter.py:203-205).  sos_token [B, Q, W]; proj [W, D]."""
        B, Q, Wd = sos_token.shape
        _, x16 = L.rownorm(sos_token.reshape(-1, Wd).float().contiguous(), ln_w, ln_b, layer_norm=True, want32=False)
        key = ("proj", proj.data_ptr(), proj._version)
        pt = self._mat.get(key)
        if pt is None:
            pt = L.cast_f16(proj.detach().float().T.contiguous())
            self._mat[key] = pt
        e = L.linear_f16(x16, pt, None, out_f32=True)
        e32, _ = L.rownorm(e, l2=True, want16=False)
        return e32.view(B, Q, -1)

    def cal_sim_logits(self, text_feats: torch.Tensor, image_feats: torch.Tensor):
        """logit_scale.exp() * image_feats @ text_feats.T (side_adapter.py:234-235)."""
        shp = image_feats.shape
        f16 = L.cast_f16(image_feats.reshape(-1, shp[-1]).float().contiguous())
        out = L.linear_f16(f16, self._text_f16(text_feats), None, scale=float(self.logit_scale_exp), out_f32=True)
        return out.view(*shp[:-1], text_feats.shape[0])
