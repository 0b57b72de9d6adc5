"""Open-vocabulary classifier tails on the B200 kernels (host-side mirror of the reference adapters' arithmetic).

ClipLogitHead     -- ClipAdapter.normalize / cal_sim_logits / text cache (clip_adapter/adapter.py:118-138, 146-147) and
                     OpenVIS.open_vocabulary_inference's per-query aggregation (openvis/openvis.py:123-141).
SideAdapterTail   -- SideAdapter._build_attn_biases, the ln_post/proj/normalize tail of post_encode_image and
                     cal_sim_logits (clip_adapter/side_adapter.py:201-207, 234-270).

SideAdapterBlocks -- SURVEY.md section 8 (f) rank 1: the post-split CLIP blocks of SideAdapter.post_encode_image
                     (side_adapter.py:176-209) on [Q SOS | CLS | patches] tokens, the additive bias applied inside the
                     attention kernel from the pooled per-head biases (never materialised), then the tail above.

The rest of the frozen CLIP towers (text encoder, the ViT blocks before the split, crop/roi_align preprocessing) is out
of scope: text embeddings are supplied already encoded ("cached text embeddings", north_star) and
region features / the split-point CLS + patch features come from the caller.
"""
import weakref
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
        """fp16 operand copy of the caller's [K, D] text matrix.  A hit needs the SAME tensor object (weak reference),
        unmodified (_version): the caching allocator hands a freed matrix's address to the next vocabulary of the same
        shape, so (data_ptr, shape) alone would return another vocabulary's embeddings."""
        key = (text.data_ptr(), tuple(text.shape))
        hit = self._mat.get(key)
        if hit is not None and hit[0]() is text and hit[1] == text._version:
            return hit[2]
        if len(self._mat) > 8:
            self._mat.clear()
        f16 = L.cast_f16(text.detach().float().contiguous())
        self._mat[key] = (weakref.ref(text), text._version, f16)
        return f16


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
            out = L.linear_f16(L.cast_f16(f), self._text_f16(text_features), None, scale=float(temperature), out_f32=True)
        else:
            # kernel 3, fused: the GEMM's epilogue divides every row by its norm (one pass makes the fp16 operand and the
            # rows' sums of squares, nothing normalised is written)
            f16, ss = L.rowstats(f)
            out = L.linear_rowscale_f16(f16, self._text_f16(text_features), None, scale=float(temperature), row_ss_in=ss,
                                        out_f32=True)
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
        pt = self._proj_f16(proj)
        e = L.linear_f16(x16, pt, None, out_f32=True)
        e32, _ = L.rownorm(e, l2=True, want16=False)
        return e32.view(B, Q, -1)

    def cal_sim_logits(self, text_feats: torch.Tensor, image_feats: torch.Tensor):
        """logit_scale.exp() * image_feats @ text_feats.T (side_adapter.py:234-235)."""
        shp = image_feats.shape
        f16 = L.cast_f16(image_feats.reshape(-1, shp[-1]).float().contiguous())
        out = L.linear_f16(f16, self._text_f16(text_feats), None, scale=float(self.logit_scale_exp), out_f32=True)
        return out.view(*shp[:-1], text_feats.shape[0])

    def _proj_f16(self, proj):
        key = ("proj", proj.data_ptr())
        hit = self._mat.get(key)
        if hit is not None and hit[0]() is proj and hit[1] == proj._version:
            return hit[2]
        pt = L.cast_f16(proj.detach().float().T.contiguous())
        self._mat[key] = (weakref.ref(proj), proj._version, pt)
        return pt

    def sos_logits(self, tokens: torch.Tensor, n: int, Q: int, Lt: int, ln_w, ln_b, proj, text_feats):
        """The whole SAN tail in three launches (side_adapter.py:203-207, 234-235): ln_post on the Q SOS rows at the head of
        every frame's [Lt, W] token block (fp16 out) -> @ visual.proj with the rows' sums of squares from the GEMM epilogue
        -> logits GEMM whose epilogue divides by the norm and multiplies by exp(logit_scale).  tokens [n * Lt, W] fp32.
        Returns logits [n, Q, K + 1]; the normalised clip_feats are never written."""
        ss = torch.empty(n * Q, dtype=torch.float32, device=tokens.device)
        x16, _ = L.rowstats(tokens, ln_w, ln_b, layer_norm=True, want_ss=False, zero=ss, groups=(n, Q, Lt))
        e16 = L.linear_rowscale_f16(x16, self._proj_f16(proj), None, row_ss_out=ss)
        out = L.linear_rowscale_f16(e16, self._text_f16(text_feats), None, scale=float(self.logit_scale_exp), row_ss_in=ss,
                                    out_f32=True)
        return out.view(n, Q, text_feats.shape[0])


class SideAdapterBlocks:
    """Drop-in for the post-split half of ``SideAdapter.post_encode_image`` (side_adapter.py:176-209).

    Weights are the frozen CLIP visual tower's (``freeze_params``, side_adapter.py:105): load them with
    ``load_clip_visual_state_dict(clip_model.visual.state_dict())`` -- the keys read are
    ``transformer.resblocks.{i}.*`` for i >= broken_idx, ``ln_post.*`` and ``proj``.  fp16 GEMM operands, fp32
    accumulation, LayerNorm / softmax statistics / residual stream in fp32.  No CPU path."""

    def __init__(self, num_queries=100, broken_idx=9, num_layers=12, width=768, heads=12):
        self.sos_token_num = num_queries
        self.blocks = tuple(range(broken_idx, num_layers))
        self.width, self.heads = width, heads
        self.tail = SideAdapterTail()
        self._w = None

    def load_clip_visual_state_dict(self, sd, prefix="transformer.resblocks."):
        f32 = lambda t: t.detach().float().contiguous().cuda()
        w16 = lambda t: L.cast_f16(f32(t))
        W = {}
        for i in self.blocks:
            g = lambda k: sd[f"{prefix}{i}.{k}"]
            W[i] = dict(ln1=(f32(g("ln_1.weight")), f32(g("ln_1.bias"))), ln2=(f32(g("ln_2.weight")), f32(g("ln_2.bias"))),
                        in_w=w16(g("attn.in_proj_weight")), in_b=f32(g("attn.in_proj_bias")),
                        out_w=w16(g("attn.out_proj.weight")), out_b=f32(g("attn.out_proj.bias")),
                        fc_w=w16(g("mlp.c_fc.weight")), fc_b=f32(g("mlp.c_fc.bias")),
                        pj_w=w16(g("mlp.c_proj.weight")), pj_b=f32(g("mlp.c_proj.bias")))
        if "ln_post.weight" in sd:
            W["ln_post"] = (f32(sd["ln_post.weight"]), f32(sd["ln_post.bias"]))
            W["proj"] = f32(sd["proj"])
        self._w = W
        return self

    def _run_blocks(self, X, n, Q, Lp, pooled):
        """The residual attention blocks `self.blocks` in place on the fp32 token matrix X [n * (Q + 1 + Lp), W]
        (ResidualAttentionBlock.forward, mask_adapted_clip model.py:237-268; BiasedResidualAttentionBlock,
        side_adapter.py:70-78 when `pooled` carries the SOS rows' attention bias)."""
        att16 = torch.empty(X.shape[0], X.shape[1], dtype=torch.float16, device=X.device)
        for i in self.blocks:
            p = self._w[i]
            _, y16 = L.rownorm(X, p["ln1"][0], p["ln1"][1], layer_norm=True, want32=False)
            qkv = L.linear_f16(y16, p["in_w"], p["in_b"])
            L.san_attn(qkv, pooled, att16, n, Q, Lp, self.heads)
            L.linear_act_f16(att16, p["out_w"], p["out_b"], resid=X, out=X, out_f32=True)
            _, y16 = L.rownorm(X, p["ln2"][0], p["ln2"][1], layer_norm=True, want32=False)
            h16 = L.linear_act_f16(y16, p["fc_w"], p["fc_b"], act=2)
            L.linear_act_f16(h16, p["pj_w"], p["pj_b"], resid=X, out=X, out_f32=True)
        return X

    @torch.no_grad()
    def post_blocks(self, feats, attn_bias, return_tokens=False):
        """feats = (cls_token [1, n, W], pix_feat [n, W, h, w]); attn_bias [n, heads, Q, H', W'] fp32 (or a one-element
        list, or None).  Returns the SOS tokens [n, Q, W] fp32 after the post-split blocks (before ln_post);
        return_tokens: the whole token matrix [n * (Q + 1 + L), W] instead (no copy of the SOS rows)."""
        if self._w is None:
            raise RuntimeError("SideAdapterBlocks: load_clip_visual_state_dict() first")
        cls_token, pix = feats
        if not pix.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        if isinstance(attn_bias, (list, tuple)):
            if len(attn_bias) != 1:
                raise NotImplementedError("one attention-bias tensor shared by all blocks (side_adapter.py:268-269)")
            attn_bias = attn_bias[0]
        n, Wd, h, w = pix.shape
        Q, Lp = self.sos_token_num, h * w
        Lt = Q + 1 + Lp
        with torch.cuda.device(pix.device):
            # token assembly [Q SOS copies of CLS | CLS | patches] (side_adapter.py:182-187, 196), frame-major rows
            x = torch.empty(n, Lt, Wd, dtype=torch.float32, device=pix.device)
            x[:, :Q + 1] = cls_token[0].float()[:, None, :]
            x[:, Q + 1:] = pix.float().flatten(2).transpose(1, 2)
            X = x.view(n * Lt, Wd)
            pooled = None
            if attn_bias is not None:
                if attn_bias.shape[1] == 1:
                    attn_bias = attn_bias.expand(-1, self.heads, -1, -1, -1)
                pooled = L.san_pool_bias(attn_bias.float().contiguous(), (h, w))
            self._run_blocks(X, n, Q, Lp, pooled)
            return X if return_tokens else x[:, :Q].contiguous()

    @torch.no_grad()
    def post_encode_logits(self, feats, attn_bias, text_feats):
        """post_encode_image + cal_sim_logits (san.py:230-231, resampler.py:313-314) as one call: the tail runs fused
        (SideAdapterTail.sos_logits) and neither the SOS-row copy nor the normalised features are materialised."""
        tokens = self.post_blocks(feats, attn_bias, return_tokens=True)
        n = feats[1].shape[0]
        Q = self.sos_token_num
        Lt = tokens.shape[0] // n
        with torch.cuda.device(tokens.device):
            return self.tail.sos_logits(tokens, n, Q, Lt, self._w["ln_post"][0], self._w["ln_post"][1], self._w["proj"], text_feats)

    @torch.no_grad()
    def post_encode_image(self, feats, attn_bias):
        """SideAdapter.post_encode_image: post-split blocks -> ln_post -> @ proj -> F.normalize; returns [n, Q, D]."""
        sos = self.post_blocks(feats, attn_bias)
        lw, lb = self._w["ln_post"]
        return self.tail.sos_tail(sos, lw, lb, self._w["proj"])

    def cal_sim_logits(self, text_feats, image_feats):
        """SideAdapter.cal_sim_logits (side_adapter.py:234-235), so that this object can be passed as the `adapter`
        argument of TemporalInstanceResampler.forward (resampler.py:244, 313-314)."""
        return self.tail.cal_sim_logits(text_feats, image_feats)


class ZeroShotClassifier(torch.nn.Module):
    """OV2Seg's classifier head (openvis/ov2seg.py:489-529), row A17: ``linear`` (Linear -> ReLU -> Linear, the reference's
    parameter names), a zero row appended to the text matrix, L2-normalise, ``norm_temperature`` (50) * x @ zs_weight^T.
    `texts` is the already-encoded [K, D] matrix (the reference asks its adapter for ``get_text_features``, which no
    adapter defines; the text tower is out of scope here either way).  use_bias < 0 adds the learned scalar ``cls_bias``."""

    def __init__(self, input_size=256, zs_weight_dim=512, use_bias=0.0, norm_weight=True, norm_temperature=50.0):
        super().__init__()
        nn = torch.nn
        self.norm_weight, self.norm_temperature = norm_weight, norm_temperature
        self.use_bias = use_bias < 0
        if self.use_bias:
            self.cls_bias = nn.Parameter(torch.ones(1) * use_bias)
        self.linear = nn.Sequential(nn.Linear(input_size, zs_weight_dim // 2), nn.ReLU(),
                                    nn.Linear(zs_weight_dim // 2, zs_weight_dim))
        self._head = ClipLogitHead()

    @torch.no_grad()
    def forward(self, x, texts):
        if not x.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        shp = x.shape
        with torch.cuda.device(x.device):
            w16 = lambda m: L.cast_f16(m.weight.detach().float().contiguous())
            h = L.linear_f16(L.cast_f16(x.reshape(-1, shp[-1]).float().contiguous()), w16(self.linear[0]),
                             self.linear[0].bias.detach().float(), relu=True)
            e = L.linear_f16(h, w16(self.linear[2]), self.linear[2].bias.detach().float(), out_f32=True)
            zs = torch.cat([texts, torch.zeros_like(texts)[0:1]])
            if self.norm_weight:
                logits = self._head.cal_sim_logits(zs, e, self.norm_temperature, normalized=False)
            else:
                logits = self._head.cal_sim_logits(zs, e, 1.0, normalized=True)
            if self.use_bias:
                logits = logits + self.cls_bias
        return logits.view(*shp[:-1], zs.shape[0])
