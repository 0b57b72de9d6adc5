"""Temporal association on the B200 kernels: host-side mirror of

  match_via_embeds / batch_video_match_via_embeds     openvis/modeling/minvis.py:28-72
  BriVIS.reset_image_output_order                     openvis/brivis.py:231-240
  TemporalInstanceResampler                           openvis/modeling/resampler.py:189-323

with the reference's names, arguments, parameter names / shapes and return values, so `BriVIS.forward`
(brivis.py:173-176) can call them unchanged.  It stays on the rank that owns the clip (no collective).

Matching.  The reference solves one assignment per frame on the host (scipy, one device->host copy each), each against
the *re-ordered* previous frame.  Re-ordering the rows of a cost matrix only re-labels them, so the chain equals the
composition of the assignments between the raw frames i-1 and i: all B*T problems are solved at once on the device
(one CTA each, exact shortest-augmenting-path solver with double potentials), then composed.  No host round trip.

Resampler.  Rows are kept instance-major, [(b q) t, 256], through the six temporal layers: self-attention over the t
frames of an instance is `ovis_self_attn` with groups = instances; the Conv1d(5) -> ReLU -> Conv1d(3) aggregation is two
tcgen05 GEMMs over a replicate-padded unfold (`ovis_temporal_unfold_f16`), the second one fused with the residual and
`aggregate_norms`; the FFN's second GEMM is fused with its LayerNorm and `decode_norm`.  Of the seven prediction heads
(resampler.py:304-316: mask einsum, attention-bias einsum, CLIP side path, logits) the reference's eval path reads only
the last (brivis.py:247-249 use outputs['pred_logits'] / ['pred_masks']); the other six are `aux_outputs`, computed on
first access from the saved per-layer `decode_norm` outputs.
Inference only, CUDA sm_100 only, no fallback.
"""
import os

import torch
from torch import nn

from . import _lib as L
from .decoder import HIDDEN, NHEADS, MLP, LazyAuxOutputs, _AttnLayer, _FFNLayer


# ------------------------------------------------------------------------------------------------ query matching
def _raw_assignments(embeds, want_cost=False):
    e = embeds.detach().float().contiguous()
    if not e.is_cuda:
        raise L.OvisError("openvis_b200 has no CPU path: embeddings must be CUDA tensors on an sm_100 device")
    b, t, q, c = e.shape
    with torch.cuda.device(e.device):
        en, _ = L.rownorm(e.view(-1, c), l2=True, want16=False)        # x / ||x||  (minvis.py:29-30)
        pi, cost = L.match_embeds(en.view(b, t, q, c), want_cost=want_cost)
    return e, pi, cost


def match_via_embeds(tgt_embeds, cur_embeds):
    """minvis.py:28-41: for every target slot, the current query aligned to it (a Python list, like the reference --
    this single-pair form is the one place that copies indices to the host)."""
    if tuple(tgt_embeds.shape) != tuple(cur_embeds.shape):
        raise NotImplementedError(
            f"match_via_embeds: square problems only (tgt {tuple(tgt_embeds.shape)} vs cur {tuple(cur_embeds.shape)}); every "
            "caller in the reference matches the same number of queries per frame (minvis.py:60-66)")
    _, pi, _ = _raw_assignments(torch.stack([tgt_embeds, cur_embeds])[None])
    return pi[0, 1].tolist()


def batch_video_match_via_embeds(orig_embeds, return_cost=False):
    """minvis.py:44-72.  orig_embeds [b, t, q, c] -> (batch_indices [b, t, q] int64, re-ordered embeds [b, t, q, c]),
    both on the device; nothing synchronises."""
    e, pi, cost = _raw_assignments(orig_embeds, want_cost=return_cost)
    with torch.cuda.device(e.device):
        idx = L.match_compose(pi)
        out = L.reorder_queries(e, idx, "btq")
    return (idx, out, pi, cost) if return_cost else (idx, out)


def reset_image_output_order(outputs, indices):
    """brivis.py:231-240: pred_logits [b, t, q, k] and pred_masks [b, q, t, h, w] gathered along q by indices [b, t, q]."""
    with torch.cuda.device(indices.device):
        idx = indices.contiguous()
        outputs["pred_logits"] = L.reorder_queries(outputs["pred_logits"].float().contiguous(), idx, "btq")
        outputs["pred_masks"] = L.reorder_queries(outputs["pred_masks"].float().contiguous(), idx, "bqt")
    return outputs


def post_processing(outputs):
    """MinVIS.post_processing (openvis/modeling/minvis.py:320-338), inherited by the online meta-architectures
    (openvis.py:214, san.py:255, simplebsl.py:270): match the frames' queries through their embeddings and bring
    pred_logits [b, t, q, k] and pred_masks [b, q, t, h, w] into the matched order.  pred_embeds [b, t, q, c]."""
    indices, _ = batch_video_match_via_embeds(outputs["pred_embeds"])
    return reset_image_output_order(outputs, indices)


# ------------------------------------------------------------------------------------------------ resampler
class TemporalInstanceResampler(nn.Module):
    """Drop-in for openvis/modeling/resampler.py:189-323 (constructed as in brivis.py:47)."""

    def __init__(self, hidden_dim=256, feed_dim=2048, nheads=8, nlayers=6):
        super().__init__()
        if hidden_dim != HIDDEN or nheads != NHEADS:
            raise NotImplementedError("openvis_b200 kernels are specialised for hidden_dim 256, 8 heads")
        self.num_heads, self.num_layers, self.feed_dim = nheads, nlayers, feed_dim
        self.long_aggregate_layers = nn.ModuleList(_AttnLayer("self_attn", hidden_dim, nheads) for _ in range(nlayers))
        self.short_aggregate_layers = nn.ModuleList(
            nn.Sequential(nn.Conv1d(hidden_dim, hidden_dim, kernel_size=5, stride=1, padding="same", padding_mode="replicate"),
                          nn.ReLU(inplace=True),
                          nn.Conv1d(hidden_dim, hidden_dim, kernel_size=3, stride=1, padding="same", padding_mode="replicate"))
            for _ in range(nlayers))
        self.aggregate_norms = nn.ModuleList(nn.LayerNorm(hidden_dim) for _ in range(nlayers))
        self.transformer_ffn_layers = nn.ModuleList(_FFNLayer(hidden_dim, feed_dim) for _ in range(nlayers))
        self.decode_norm = nn.LayerNorm(hidden_dim)
        self.attn_embed = MLP(hidden_dim, hidden_dim, hidden_dim, 3)
        self.mask_embed = MLP(hidden_dim, hidden_dim, hidden_dim, 3)
        self.adapter = None
        self.text_feats = None
        # non-reference knobs
        self.materialize_aux = False     # True: compute the six aux heads eagerly (API-exact mode)
        self.use_cuda_graph = os.environ.get("OVIS_NO_CUDA_GRAPH") is None      # replay the layer loop as a CUDA graph
        self.operand_source = None       # the SAN decoder whose outputs are passed in: its fp16 operand copies are reused
        self._wcache = None
        self._ws = {}
        self._generation = 0

    def _weights(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._wcache is not None and self._wcache["key"] == key:
            return self._wcache
        f16 = lambda t: L.cast_f16(t.detach().float().contiguous())
        f32 = lambda t: t.detach().float().contiguous()
        ln = lambda m: (f32(m.weight), f32(m.bias))
        conv = lambda m: f16(m.weight.detach().permute(0, 2, 1).flatten(1))        # [out, in, k] -> [out, k*in]
        C = HIDDEN
        W = {"key": key, "layers": []}
        for i in range(self.num_layers):
            sa, ff, cv = self.long_aggregate_layers[i], self.transformer_ffn_layers[i], self.short_aggregate_layers[i]
            w, b = sa.self_attn.in_proj_weight, sa.self_attn.in_proj_bias
            W["layers"].append(dict(
                qk_w=f16(w[:2 * C]), qk_b=f32(b[:2 * C]), v_w=f16(w[2 * C:]), v_b=f32(b[2 * C:]),
                o_w=f16(sa.self_attn.out_proj.weight), o_b=f32(sa.self_attn.out_proj.bias), ln_a=ln(sa.norm),
                c5_w=conv(cv[0]), c5_b=f32(cv[0].bias), c3_w=conv(cv[2]), c3_b=f32(cv[2].bias),
                ln_c=ln(self.aggregate_norms[i]),
                f1_w=f16(ff.linear1.weight), f1_b=f32(ff.linear1.bias), f2_w=f16(ff.linear2.weight), f2_b=f32(ff.linear2.bias),
                ln_f=ln(ff.norm)))
        W["dn"] = ln(self.decode_norm)
        W["mask_embed"] = [(f16(m.weight), f32(m.bias)) for m in self.mask_embed.layers]
        W["attn_embed"] = [(f16(m.weight), f32(m.bias)) for m in self.attn_embed.layers]
        self._wcache = W
        return W

    @staticmethod
    def _mlp3(params, x16):
        x = L.linear_f16(x16, params[0][0], params[0][1], relu=True)
        x = L.linear_f16(x, params[1][0], params[1][1], relu=True)
        return L.linear_f16(x, params[2][0], params[2][1])

    @torch.no_grad()
    def forward(self, frame_embeds, mask_feats, attn_feats, adapter=None, clip_bk_feats=None, text_feats=None):
        """frame_embeds [b, t, q, 256]; mask_feats [(b t), 256, h, w]; attn_feats [(b t), n, 256, h', w'];
        adapter: object with post_encode_image(clip_bk_feats, attn_biases) and cal_sim_logits(text_feats, clip_feats)
        (ov_head.SideAdapterBlocks, or the reference's SideAdapter).  Returns the reference's dictionary:
        pred_logits [b, t, q, K], pred_masks [b, q, t, h, w], pred_embeds [b, t, q, 256], aux_outputs (6 heads)."""
        if self.training:
            raise RuntimeError("openvis_b200 TemporalInstanceResampler is inference-only: call .eval()")
        if not frame_embeds.is_cuda:
            raise L.OvisError("openvis_b200 has no CPU path: inputs must be CUDA tensors on an sm_100 device")
        bs, t, q, c = frame_embeds.shape
        if c != HIDDEN or t > 1536 or q > 256:
            raise NotImplementedError(f"frame_embeds must be [b, t <= 1536, q <= 256, 256], got {tuple(frame_embeds.shape)}")
        BT, _, H, Wd = mask_feats.shape
        nh = attn_feats.shape[1]
        assert BT == bs * t and attn_feats.shape[0] == BT and attn_feats.shape[2] == HIDDEN
        dev = frame_embeds.device
        with torch.cuda.device(dev):
            return self._forward_impl(frame_embeds, mask_feats, attn_feats, adapter, clip_bk_feats, text_feats,
                                      bs, t, q, BT, H, Wd, nh, dev)

    def _workspace(self, bs, t, q, dev):
        key = (bs, t, q, str(dev))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        if len(self._ws) >= 2:
            self._ws.pop(next(iter(self._ws)))
        C, nl = HIDDEN, self.num_layers
        G, R = bs * q, bs * q * t
        h16 = lambda *s: torch.empty(*s, dtype=torch.float16, device=dev)
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        ws = dict(G=G, R=R, x32=f32(R, C), x16=h16(R, C), l32=f32(R, C), l16=h16(R, C), d32=f32(R, C),
                  d16=h16(nl + 1, R, C),                # decode_norm output feeding each of the seven heads
                  qk16=h16(R, 2 * C), v16=h16(R, C), sa16=h16(R, C), c16=h16(R, C), f16=h16(R, self.feed_dim),
                  u5=h16(G, t, 5 * C), u3=h16(G, t, 3 * C),
                  split=f32((self.feed_dim // 256) * ((R + 127) // 128) * 128 * 256) if R <= 16384 else None)
        self._ws[key] = ws
        return ws

    def _layer_loop(self, W, ws, t):
        """The six temporal layers on the instance-major residual stream ws["x32"] [(b q) t, 256].  Reads / writes the
        workspace and the weight cache only, so it can be replayed as a CUDA graph."""
        C, G, R = HIDDEN, ws["G"], ws["R"]
        x32, x16, l32, l16, split = ws["x32"], ws["x16"], ws["l32"], ws["l16"], ws["split"]
        L.cast_f16(x32, out=x16)
        L.rownorm(x32, W["dn"][0], W["dn"][1], layer_norm=True, want32=False, out16=ws["d16"][0])
        for i in range(self.num_layers):
            lw = W["layers"][i]
            # long-term aggregation: self-attention over the frames of each instance (resampler.py:258-262)
            L.linear_f16(x16, lw["qk_w"], lw["qk_b"], out=ws["qk16"])
            L.linear_f16(x16, lw["v_w"], lw["v_b"], out=ws["v16"])
            L.self_attn(ws["qk16"], ws["v16"], ws["sa16"], G, t)
            L.linear_ln_f16(ws["sa16"], lw["o_w"], lw["o_b"], x32, lw["ln_a"], y32=l32, y16=l16, split_ws=split)
            # short-term aggregation: Conv1d(5) -> ReLU -> Conv1d(3) over t + residual + aggregate_norms (:264-267)
            L.temporal_unfold_f16(l16.view(G, t, C), 5, out=ws["u5"])
            L.linear_f16(ws["u5"].view(R, 5 * C), lw["c5_w"], lw["c5_b"], relu=True, out=ws["c16"])
            L.temporal_unfold_f16(ws["c16"].view(G, t, C), 3, out=ws["u3"])
            L.linear_ln_f16(ws["u3"].view(R, 3 * C), lw["c3_w"], lw["c3_b"], l32, lw["ln_c"], y32=x32, y16=x16, split_ws=split)
            # FFN (:270) + decode_norm of the following head (:305)
            L.linear_f16(x16, lw["f1_w"], lw["f1_b"], relu=True, out=ws["f16"])
            L.linear_ln_f16(ws["f16"], lw["f2_w"], lw["f2_b"], x32, lw["ln_f"], W["dn"], y32=x32, y16=x16, d32=ws["d32"],
                            d16=ws["d16"][i + 1], split_ws=split)

    def _run_layers(self, W, ws, t):
        """~60 small dependent launches: replayed as a CUDA graph after one eager call per workspace (as the decoders do)."""
        eager = (not self.use_cuda_graph) or L.PROFILE is not None
        if eager or not ws.get("warm"):
            ws["warm"] = True
            return self._layer_loop(W, ws, t)
        if ws.get("graph") is None or ws.get("graph_W") is not W:
            n0 = L.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._layer_loop(W, ws, t)
            ws["graph"], ws["graph_W"], ws["graph_launches"] = g, W, L.launch_count() - n0
            L.add_launch_count(-ws["graph_launches"])
        ws["graph"].replay()
        L.add_launch_count(ws["graph_launches"])

    def _forward_impl(self, frame_embeds, mask_feats, attn_feats, adapter, clip_bk_feats, text_feats,
                      bs, t, q, BT, H, Wd, nh, dev):
        W = self._weights()
        ws = self._workspace(bs, t, q, dev)
        C, nl = HIDDEN, self.num_layers
        self._generation += 1
        gen = self._generation
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)

        # operands of the two head einsums, token-major fp16: the decoder that produced mask_feats / attn_feats already
        # holds them (operand_source); otherwise one layout pass each
        ah, aw = attn_feats.shape[-2:]
        M, P = H * Wd, ah * aw
        shared = self.operand_source.shared_operands(mask_feats, attn_feats) if self.operand_source is not None else None
        if shared is not None:
            ft, af = shared
        else:
            tok = lambda x: (L.nchw_to_tokens_hw_f16 if x.shape[-1] % 4 == 0 and x.shape[1] % 32 == 0 else L.nchw_to_tokens_f16)(x)
            ft = tok(mask_feats.float().contiguous())                                        # [BT, M, 256]
            af = tok(attn_feats.float().contiguous().view(BT, nh * C, ah, aw))               # [BT, P, n*256]
        ft, af = ft.view(BT, M, C), af.view(BT, P, nh * C)

        # instance-major residual stream [(b q) t, 256]
        ws["x32"].view(bs, q, t, C).copy_(frame_embeds.detach().permute(0, 2, 1, 3))
        self._run_layers(W, ws, t)
        d16, d32 = ws["d16"], ws["d32"]

        def head(hidx):
            """forward_prediction_heads (resampler.py:304-316) from the saved decode_norm output of head `hidx`."""
            d = d16[hidx].view(bs, q, t, C).transpose(1, 2).contiguous().view(BT * q, C)          # rows (b t) q
            me = self._mlp3(W["mask_embed"], d)
            masks = f32(bs, q, t, H, Wd)
            for b in range(bs):                                                                    # out[q][frame][pixel]
                L.mask_logits(ft[b * t:(b + 1) * t], t, M, me[b * t * q:(b + 1) * t * q], q, q, masks[b], M, t * M)
            ae = self._mlp3(W["attn_embed"], d)
            biases = f32(BT, nh, q, ah, aw)
            L.san_bias_logits(af, BT, P, nh, ae, q, biases)
            if hasattr(adapter, "post_encode_logits"):        # fused tail (ov_head.SideAdapterBlocks)
                logits = adapter.post_encode_logits(clip_bk_feats, biases, text_feats)
            else:                                             # the reference's two calls (resampler.py:313-314)
                logits = adapter.cal_sim_logits(text_feats, adapter.post_encode_image(clip_bk_feats, biases))
            return logits.reshape(bs, t, q, -1), masks

        logits, masks = head(nl)
        out = {"pred_logits": logits, "pred_masks": masks,
               "pred_embeds": d32.view(bs, q, t, C).transpose(1, 2).contiguous()}

        def compute_aux(j):
            if gen != self._generation:
                raise RuntimeError("aux_outputs must be read before the next forward() of the same resampler "
                                   "(set resampler.materialize_aux = True to compute them eagerly)")
            if shared is not None and self.operand_source.shared_operands(mask_feats, attn_feats) is None:
                raise RuntimeError("aux_outputs must be read before the next forward() of the decoder whose operand "
                                   "copies this call shared (resampler.operand_source)")
            with torch.cuda.device(dev):
                lg, m = head(j)
            return {"pred_logits": lg, "pred_masks": m}

        aux = LazyAuxOutputs(nl, compute_aux)
        if self.materialize_aux:
            list(aux)
        out["aux_outputs"] = aux
        return out


# ------------------------------------------------------------------------------------------------ BriVIS eval schedule
@torch.no_grad()
def brivis_video_inference(decoder, adapter, resampler, features, mask_features, clip_bk_feats, text_feats,
                           padded_size, image_size, height, width, api_exact=False, num_clips=1, num_classes=None,
                           to_host=True):
    """The part of ``BriVIS.forward``'s eval branch that lies on the hot path (openvis/brivis.py:157-190, 242-265), from the
    pixel decoder's outputs to the video result, composed from the drop-in pieces exactly as the reference composes its
    own: SAN frame decoder -> query matching -> TemporalInstanceResampler (heads through the CLIP side path) ->
    post_processing -> inference_video.

    decoder   SideAdapterFrameMultiScaleMaskedTransformerDecoder (sem_seg_head.predictor, brivis.py:160)
    adapter   object with post_encode_image / cal_sim_logits (ov_head.SideAdapterBlocks; self.clip_adapter)
    features  the three multi-scale maps (coarsest first) and mask_features [b*T, 256, Hp/4, Wp/4] of the clips' frames
    api_exact additionally computes what the reference computes but never reads in eval: the frame-level pred_logits
              (brivis.py:169-170) and reset_image_output_order (:174).
    num_clips b = len(batched_inputs) of the reference (1 in its eval loop).  The frame decoder, the matching and the
              resampler are all written for b clips of equal length (brivis.py:164-176 carry b through einops), and the
              query-side launch chain costs the same for 36 or 144 frames, so several clips per call is the
              throughput setting (BASELINE config 5).
    Returns (video_output as VideoMaskFormer.inference_video, resampler outputs, indices [b, T, Q]); video_output is a list
    of b dictionaries when num_clips > 1."""
    from .postprocess import inference_video
    image_outputs = decoder(features, mask_features)
    q, b = decoder.num_queries, int(num_clips)
    bt = mask_features.shape[0]
    if b < 1 or bt % b:
        raise ValueError(f"num_clips={b} does not divide the {bt} frames of this call")
    t = bt // b
    pred_embeds = image_outputs["pred_embeds"][0].view(b, t, q, -1)               # (1, bt, q, c) -> (b, t, q, c)
    if api_exact:
        biases = image_outputs["class_attn_biases"][0]                            # (1, bt, n, q, h, w) -> (bt, n, q, h, w)
        clip_feats = adapter.post_encode_image(clip_bk_feats, biases)
        lg = adapter.cal_sim_logits(text_feats, clip_feats)
        image_outputs["pred_logits"] = lg.view(b, t, q, -1)                       # (b t) q c -> b t q c
        pm = image_outputs["pred_masks"][0]                                       # q (b t) h w -> b q t h w
        image_outputs["pred_masks"] = pm[None] if b == 1 else pm.view(q, b, t, *pm.shape[-2:]).transpose(0, 1).contiguous()
    indices, frame_embeds = batch_video_match_via_embeds(pred_embeds)
    if api_exact:
        image_outputs = reset_image_output_order(image_outputs, indices)
    prev_source = resampler.operand_source
    resampler.operand_source = decoder              # for this call only: the decoder's fp16 operand copies are reused
    try:
        outputs = resampler(frame_embeds, image_outputs["mask_feats"], image_outputs["attn_feats"], adapter, clip_bk_feats,
                            text_feats)
    finally:
        resampler.operand_source = prev_source
    # post_processing (brivis.py:242-265): mean of the logits over the frames, softmax, drop the background column
    videos, scores = [], []
    for c in range(b):
        logits = outputs["pred_logits"][c].float().contiguous()                   # [t, q, K + 1]
        # brivis.py:246-249: softmax + drop the last column only when the logits carry the background column
        # (shape[-1] == num_classes + 1); num_classes=None keeps the shipped-config behaviour (text_feats has K + 1 rows)
        if num_classes is not None and logits.shape[-1] != num_classes + 1:
            mask_cls = logits.mean(0)
        else:
            with torch.cuda.device(logits.device):
                probs, _ = L.clip_aggregate(logits, torch.ones(t, q, dtype=torch.uint8, device=logits.device))
            mask_cls = probs[:, :-1].contiguous()
        scores.append(mask_cls)
        videos.append(inference_video(q, mask_cls.shape[1], mask_cls, outputs["pred_masks"][c], padded_size, image_size,
                                      height, width, to_host=to_host))
    # extension key: post_processing's scores [q, K] ([b, q, K] for several clips)
    outputs["mask_cls_result"] = scores[0] if b == 1 else torch.stack(scores)
    return (videos[0] if b == 1 else videos), outputs, indices


@torch.no_grad()
def san_online_video_inference(decoder, adapter, features, mask_features, clip_bk_feats, text_feats, padded_size, image_size,
                               height, width, num_clips=1, num_classes=None, to_host=True):
    """The hot-path part of ``SANOnline.forward``'s eval branch (openvis/san.py:226-283): SAN frame decoder ->
    post_encode_image(clip_bk_feats, class_attn_biases) -> cal_sim_logits -> MinVIS.post_processing (query matching
    through the embeddings, logits and masks brought into the matched order, minvis.py:320-338) -> mean over the frames,
    softmax, drop the background column -> inference_video.  Arguments as brivis_video_inference.
    Returns (video_output, outputs {pred_logits [b, t, q, K+1] matched, pred_masks matched, mask_cls_result}, indices)."""
    from .postprocess import inference_video
    outputs = decoder(features, mask_features)
    q, b = decoder.num_queries, int(num_clips)
    bt = mask_features.shape[0]
    if b < 1 or bt % b:
        raise ValueError(f"num_clips={b} does not divide the {bt} frames of this call")
    t = bt // b
    biases = outputs["class_attn_biases"]
    if hasattr(adapter, "post_encode_logits"):                                                  # fused tail
        logits = adapter.post_encode_logits(clip_bk_feats, biases.flatten(0, 1), text_feats).view(b, t, q, -1)
    else:
        clip_feats = adapter.post_encode_image(clip_bk_feats, biases.flatten(0, 1))             # san.py:230
        logits = adapter.cal_sim_logits(text_feats, clip_feats).view(b, t, q, -1)                # '(b t) q c -> b t q c'
    embeds = outputs["pred_embeds"][0].view(b, t, q, -1)
    indices, _ = batch_video_match_via_embeds(embeds)                                            # minvis.py:322-323
    pm = outputs["pred_masks"][0]                                                                # [q, (b t), h, w]
    with torch.cuda.device(pm.device):
        lg = L.reorder_queries(logits.float().contiguous(), indices, "btq")                      # minvis.py:325-330
        pm = L.reorder_queries(pm.view(q, b, t, *pm.shape[-2:]), indices, "qbt")                 # minvis.py:332-336
    videos, scores = [], []
    for c in range(b):
        lc = lg[c].contiguous()
        if num_classes is not None and lc.shape[-1] != num_classes + 1:                          # san.py:259-260
            mask_cls = lc.mean(0)
        else:
            with torch.cuda.device(lc.device):
                probs, _ = L.clip_aggregate(lc, torch.ones(t, q, dtype=torch.uint8, device=lc.device))
            mask_cls = probs[:, :-1].contiguous()
        scores.append(mask_cls)
        videos.append(inference_video(q, mask_cls.shape[1], mask_cls, pm[:, c], padded_size, image_size, height, width,
                                      to_host=to_host))
    out = {"pred_logits": lg, "pred_masks": pm.transpose(0, 1), "mask_cls_result": scores[0] if b == 1 else torch.stack(scores)}
    return (videos[0] if b == 1 else videos), out, indices
