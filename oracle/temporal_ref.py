"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's temporal association.

  match_via_embeds / batch_video_match_via_embeds   openvis/modeling/minvis.py:28-72
  reset_image_output_order                          openvis/brivis.py:231-240 (batch_index, openvis/utils/index.py:4-19)
  TemporalInstanceResampler.forward / heads         openvis/modeling/resampler.py:244-316

Same rules as oracle/decoder_ref.py: only tests/, smoke() and bench.py's CPU legs may import this; the product package
never does.  Pinned by tests/test_oracle_golden.py against tests/golden/temporal_match.npz and temporal_resampler.npz
(outputs of the reference's own functions / module, oracle/make_golden.py) and live against the reference when
/root/reference is mounted.  The assignment solver is scipy.optimize.linear_sum_assignment, the third-party routine the
reference itself calls (minvis.py:15,38; unpinned in requirements.txt:2, 1.x here).
"""
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

from .decoder_ref import layer_norm, mha, mlp


def match_cost(tgt_embeds, cur_embeds):
    """cost[cur, tgt] = 1 - cos (minvis.py:29-33)."""
    cur = cur_embeds / cur_embeds.norm(dim=1)[:, None]
    tgt = tgt_embeds / tgt_embeds.norm(dim=1)[:, None]
    return 1 - cur @ tgt.T


def match_via_embeds(tgt_embeds, cur_embeds):
    """Hungarian matching target x current; returns, for every target slot, the current query aligned to it
    (minvis.py:28-41)."""
    C = match_cost(tgt_embeds, cur_embeds)
    return linear_sum_assignment(C.T.numpy())[1].tolist()


def batch_video_match_via_embeds(orig_embeds):
    """Chain over the frames of each clip: frame i is matched to the re-ordered frame i-1 (frame 0 to itself)
    (minvis.py:44-72).  orig_embeds [b, t, q, c] -> indices [b, t, q] int64, re-ordered embeds [b, t, q, c]."""
    bs, t = orig_embeds.shape[:2]
    all_idx, all_emb = [], []
    for b in range(bs):
        last = orig_embeds[b, 0]
        idx_l, emb_l = [], []
        for i in range(t):
            idx = match_via_embeds(last, orig_embeds[b, i])
            last = orig_embeds[b, i][idx]
            idx_l.append(idx)
            emb_l.append(last)
        all_idx.append(idx_l)
        all_emb.append(torch.stack(emb_l))
    return torch.tensor(all_idx), torch.stack(all_emb)


def reset_image_output_order(pred_logits, pred_masks, indices):
    """brivis.py:231-240: pred_logits [b, t, q, k], pred_masks [b, q, t, h, w] gathered along q by indices [b, t, q]."""
    b, t, q = indices.shape
    bt = torch.arange(b * t)[:, None]
    fl = pred_logits.flatten(0, 1)[bt, indices.flatten(0, 1)]
    fm = pred_masks.transpose(2, 1).flatten(0, 1)[bt, indices.flatten(0, 1)]
    return fl.view(b, t, q, -1), fm.view(b, t, q, *fm.shape[-2:]).transpose(1, 2)


def resampler_heads(P, x_tq, mask_feats, attn_feats, post_encode_image, cal_sim_logits):
    """TemporalInstanceResampler.forward_prediction_heads (resampler.py:304-316) on x_tq [(b t), q, c] (the reference's
    `output.transpose(1, 0)` layout).  post_encode_image(attn_biases) / cal_sim_logits(clip_feats) are the adapter calls
    with clip_bk_feats / text_feats already bound."""
    d = layer_norm(x_tq, P["decode_norm.weight"], P["decode_norm.bias"])
    masks = torch.einsum("bqc,bchw->bqhw", mlp(P, "mask_embed", d, 3), mask_feats)
    biases = torch.einsum("bqc,bnchw->bnqhw", mlp(P, "attn_embed", d, 3), attn_feats)
    return cal_sim_logits(post_encode_image(biases)), masks


def resampler_layer(P, i, x, nheads=8):
    """One temporal layer (resampler.py:258-277) on x [(b q), t, c]: self-attention over the frames of each instance
    (no positional term), Conv1d(5) -> ReLU -> Conv1d(3) over t with replicate padding + residual, LayerNorm, FFN."""
    pre = f"long_aggregate_layers.{i}"
    a = mha(x, x, x, P[f"{pre}.self_attn.in_proj_weight"], P[f"{pre}.self_attn.in_proj_bias"],
            P[f"{pre}.self_attn.out_proj.weight"], P[f"{pre}.self_attn.out_proj.bias"], nheads)
    long_tgt = layer_norm(x + a, P[f"{pre}.norm.weight"], P[f"{pre}.norm.bias"])
    s = long_tgt.transpose(1, 2)                                                       # [(b q), c, t]
    h = F.conv1d(F.pad(s, (2, 2), mode="replicate"), P[f"short_aggregate_layers.{i}.0.weight"],
                 P[f"short_aggregate_layers.{i}.0.bias"]).relu()
    h = F.conv1d(F.pad(h, (1, 1), mode="replicate"), P[f"short_aggregate_layers.{i}.2.weight"],
                 P[f"short_aggregate_layers.{i}.2.bias"])
    y = layer_norm((h + s).transpose(1, 2), P[f"aggregate_norms.{i}.weight"], P[f"aggregate_norms.{i}.bias"])
    pre = f"transformer_ffn_layers.{i}"
    f = (y @ P[f"{pre}.linear1.weight"].T + P[f"{pre}.linear1.bias"]).relu() @ P[f"{pre}.linear2.weight"].T \
        + P[f"{pre}.linear2.bias"]
    return layer_norm(y + f, P[f"{pre}.norm.weight"], P[f"{pre}.norm.bias"])


def resampler_forward(P, frame_embeds, mask_feats, attn_feats, post_encode_image, cal_sim_logits, num_layers=6,
                      heads_at=None):
    """TemporalInstanceResampler.forward (resampler.py:244-302).  frame_embeds [b, t, q, c]; mask_feats [(b t), c, h, w];
    attn_feats [(b t), n, c, h', w'].  Returns pred_logits [b, t, q, K], pred_masks [b, q, t, h, w],
    pred_embeds [b, t, q, c] and aux = {head index: (logits, masks)} for the heads listed in `heads_at`
    (default: all seven; the last one is the main output)."""
    b, t, q, c = frame_embeds.shape
    heads_at = range(num_layers + 1) if heads_at is None else heads_at
    pack = lambda lg, m: (lg.view(b, t, q, -1), m.view(b, t, q, *m.shape[-2:]).transpose(1, 2))
    heads = {}
    if 0 in heads_at:
        heads[0] = pack(*resampler_heads(P, frame_embeds.reshape(b * t, q, c), mask_feats, attn_feats,
                                         post_encode_image, cal_sim_logits))
    x = frame_embeds.permute(0, 2, 1, 3).reshape(b * q, t, c)
    for i in range(num_layers):
        x = resampler_layer(P, i, x)
        if i + 1 in heads_at or i + 1 == num_layers:
            x_tq = x.view(b, q, t, c).transpose(1, 2).reshape(b * t, q, c)
            heads[i + 1] = pack(*resampler_heads(P, x_tq, mask_feats, attn_feats, post_encode_image, cal_sim_logits))
    emb = layer_norm(x, P["decode_norm.weight"], P["decode_norm.bias"]).view(b, q, t, c).transpose(1, 2)
    lg, m = heads[num_layers]
    return dict(pred_logits=lg, pred_masks=m, pred_embeds=emb, heads=heads)


def brivis_video_inference(dec_params, res_params, x, mask_features, post_encode_image, cal_sim_logits,
                           padded_size, image_size, out_hw, clip_heads=12):
    """The hot-path part of BriVIS.forward's eval branch (openvis/brivis.py:157-190, 242-265) composed from the pinned
    restatements: SAN frame decoder -> query matching -> resampler -> mean over frames / softmax / drop background
    -> VideoMaskFormer.inference_video.  One clip.  Returns a dict of intermediate and final results."""
    from . import decoder_ref as O
    dec = O.decoder_forward(dec_params, x, mask_features, kind="san_frame", clip_heads=clip_heads, return_attn_masks=False)
    indices, frame_embeds = batch_video_match_via_embeds(dec["pred_embeds"])
    res = resampler_forward(res_params, frame_embeds, dec["mask_feats"], dec["attn_feats"], post_encode_image,
                            cal_sim_logits, heads_at=())
    cls = res["pred_logits"].mean(dim=1)[0].softmax(-1)[:, :-1]
    sc, lab, qi, ent, masks, mlog = O.video_postprocess(cls, res["pred_masks"][0], padded_size, image_size, out_hw)
    return dict(decoder=dec, indices=indices, resampler=res, mask_cls=cls, scores=sc, labels=lab, queries=qi,
                entropys=ent, masks=masks, mask_logits=mlog)


def san_online_video_inference(dec_params, x, mask_features, post_encode_image, cal_sim_logits, padded_size, image_size,
                               out_hw, clip_heads=12):
    """The hot-path part of SANOnline.forward's eval branch (openvis/san.py:226-283) composed from the pinned restatements:
    SAN frame decoder -> post_encode_image(class_attn_biases.flatten(0, 1)) -> cal_sim_logits (san.py:230-231) ->
    MinVIS.post_processing (minvis.py:320-338: query matching, logits and masks gathered into the matched order; the gather
    is the same batch_index arithmetic as reset_image_output_order) -> mean over frames, softmax, drop background
    (san.py:255-260) -> VideoMaskFormer.inference_video.  One clip."""
    from . import decoder_ref as O
    dec = O.decoder_forward(dec_params, x, mask_features, kind="san_frame", clip_heads=clip_heads, return_attn_masks=False)
    biases = dec["class_attn_biases"]                                       # [1, t, n, q, h, w]
    t, q = biases.shape[1], biases.shape[3]
    logits = cal_sim_logits(post_encode_image(biases.flatten(0, 1))).view(1, t, q, -1)
    indices, _ = batch_video_match_via_embeds(dec["pred_embeds"])
    lg, pm = reset_image_output_order(logits, dec["pred_masks"], indices)
    cls = lg.mean(dim=1)[0].softmax(-1)[:, :-1]
    sc, lab, qi, ent, masks, mlog = O.video_postprocess(cls, pm[0], padded_size, image_size, out_hw)
    return dict(decoder=dec, indices=indices, pred_logits=lg, pred_masks=pm, mask_cls=cls, scores=sc, labels=lab, queries=qi,
                entropys=ent, masks=masks, mask_logits=mlog)
