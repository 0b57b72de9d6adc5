"""TEST INFRASTRUCTURE ONLY -- CPU (fp32/fp64 torch) restatement of the reference hot path.

This is the *oracle*: a functional, module-free restatement of the arithmetic of the reference's
masked transformer decoders, mask head and open-vocabulary tails.  It is what the CUDA path is
checked against on the GPU box (where ``/root/reference`` does not exist).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it; the product package ``openvis_b200`` never does.

Pinning: ``tests/test_oracle_golden.py`` checks it against the committed fixtures under
``tests/golden/`` (outputs of the reference's own, unmodified modules generated in the build container
by ``oracle/make_golden.py``) and, when ``/root/reference`` is present, directly against the live
reference modules at the full dimensions.  The reference's own test-suite holds no vectors for this
path, so those generated fixtures are the pin.

Every function cites the reference file:line it restates (paths relative to the reference root).
All functions take a flat ``params`` dict with the reference's state_dict names.
"""
import math

import torch
import torch.nn.functional as F

_DEC = "openvis/modeling/transformer_decoder/"


# --------------------------------------------------------------------------------------------
# position embeddings -- transformer_decoder/position_encoding.py
# --------------------------------------------------------------------------------------------
def _sincos(coord, npf, temperature):
    """interleaved sin(even)/cos(odd) of coord / T^(2*floor(i/2)/npf); position_encoding.py:89-100."""
    i = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / npf)
    ang = coord[..., None] / dim_t
    out = torch.empty_like(ang)
    out[..., 0::2] = ang[..., 0::2].sin()
    out[..., 1::2] = ang[..., 1::2].cos()
    return out


def sine_pos_2d(h, w, npf=128, temperature=10000.0):
    """PositionEmbeddingSine2D.forward with mask=None, normalize=True (position_encoding.py:78-103).
    Returns [2*npf, h, w] (identical for every batch element)."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32)
    x = torch.arange(1, w + 1, dtype=torch.float32)
    y = y / (y[-1] + eps) * scale
    x = x / (x[-1] + eps) * scale
    py = _sincos(y, npf, temperature)[:, None, :].expand(h, w, npf)
    px = _sincos(x, npf, temperature)[None, :, :].expand(h, w, npf)
    return torch.cat([py, px], dim=-1).permute(2, 0, 1).contiguous()


def sine_pos_3d(t, h, w, npf=128, temperature=10000.0):
    """PositionEmbeddingSine3D.forward (position_encoding.py:135-165): cat(pos_y,pos_x) + pos_z where
    pos_z spans 2*npf channels with its own frequency table.  Returns [t, 2*npf, h, w]."""
    eps, scale = 1e-6, 2 * math.pi
    z = torch.arange(1, t + 1, dtype=torch.float32)
    z = z / (z[-1] + eps) * scale
    pz = _sincos(z, 2 * npf, temperature)                      # [t, 2npf]
    p2 = sine_pos_2d(h, w, npf, temperature)                    # [2npf, h, w]
    return p2[None] + pz[:, :, None, None]


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def mlp(params, prefix, x, n_layers):
    """MLP.forward (video_mask2former_transformer_decoder.py:204-216): ReLU between layers."""
    for i in range(n_layers):
        x = x @ params[f"{prefix}.layers.{i}.weight"].T + params[f"{prefix}.layers.{i}.bias"]
        if i < n_layers - 1:
            x = x.relu()
    return x


def mha(q_in, k_in, v_in, in_w, in_b, out_w, out_b, nheads, blocked=None):
    """nn.MultiheadAttention as executed by the reference: packed in-proj, q scaled
    by d^-1/2 BEFORE the product, bool mask (True = blocked) -> -inf, softmax over keys, out-proj.
    q_in [G, Lq, C]; k_in, v_in [G, Lk, C]; blocked [G, Lq, Lk] bool shared by all heads."""
    C = q_in.shape[-1]
    d = C // nheads
    q = q_in @ in_w[:C].T + in_b[:C]
    k = k_in @ in_w[C:2 * C].T + in_b[C:2 * C]
    v = v_in @ in_w[2 * C:].T + in_b[2 * C:]
    G, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.view(G, Lq, nheads, d).transpose(1, 2) * (d ** -0.5)
    k = k.view(G, Lk, nheads, d).transpose(1, 2)
    v = v.view(G, Lk, nheads, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2)                                 # [G, h, Lq, Lk]
    if blocked is not None:
        s = s.masked_fill(blocked[:, None], float("-inf"))
    a = s.softmax(-1)
    o = (a @ v).transpose(1, 2).reshape(G, Lq, C)
    return o @ out_w.T + out_b


def unblock_full_rows(blocked):
    """attn_mask[where(attn_mask.sum(-1) == N)] = False (frame_...decoder.py:87, video_...:419)."""
    full = blocked.all(-1, keepdim=True)
    return blocked & ~full


def attn_mask_from_logits(mask_logits, target_hw):
    """forward_prediction_heads tail (frame_...decoder.py:148-152): bilinear resize, sigmoid, < 0.5.
    mask_logits [N, Q, H, W] -> bool [N, Q, h*w] (True = blocked).  Head replication is implicit."""
    m = F.interpolate(mask_logits, size=tuple(target_hw), mode="bilinear", align_corners=False)
    return m.sigmoid().flatten(2) < 0.5


def san_attn_features(params, mask_features, clip_heads):
    """attn_mlp(bilinear 1/4 (mask_features)) (side_adapter_frame_...decoder.py:67-71, ConvMLP 14-26).
    Returns [BT, clip_heads, C, h, w]."""
    bt, c = mask_features.shape[:2]
    a = F.interpolate(mask_features, scale_factor=0.25, mode="bilinear", align_corners=False)
    h, w = a.shape[-2:]
    for i in range(3):
        a = F.conv2d(a, params[f"attn_mlp.layers.{i}.weight"], params[f"attn_mlp.layers.{i}.bias"])
        if i < 2:
            a = a.relu()
    return a.reshape(bt, clip_heads, c, h, w)


# --------------------------------------------------------------------------------------------
# the decoder
# --------------------------------------------------------------------------------------------
def decoder_forward(params, x, mask_features, kind="frame", nheads=8, num_layers=None, clip_heads=12,
                    return_attn_masks=True, dtype=torch.float32):
    """Restates, for eval mode (bs = 1, t = bt):
      kind="video"     VideoMultiScaleMaskedTransformerDecoder.forward            (video_...:380-452)
      kind="frame"     FrameMultiScaleMaskedTransformerDecoder.forward            (frame_...:52-137)
      kind="san_frame" SideAdapterFrameMultiScaleMaskedTransformerDecoder.forward (side_adapter_frame...:57-149)
      kind="san_video" SideAdapterVideoMultiScaleMaskedTransformerDecoder.forward (side_adapter_video...:51-120)
      kind="zero_shot" ZeroShotMultiScaleMaskedTransformerDecoder.forward (zero_shot_...:172-265): a still-image decoder
                       (every image its own group, no frame axis in the outputs) whose class output is the decoder_norm
                       embedding itself plus a 2-way `object_embed` MLP
    x: list of 3 [T, C, h_l, w_l] (coarsest first); mask_features [T, C, H4, W4].
    Returns the reference's output dict (same keys / shapes) plus, when return_attn_masks,
    "attn_masks": list of L+1 bool tensors [G, Q, keys] (True = blocked, before the full-row fix).
    """
    P = {k: v.to(dtype) for k, v in params.items()}
    x = [t.to(dtype) for t in x]
    mask_features = mask_features.to(dtype)
    video = kind.endswith("video")          # (embedding_* / proposal_* kinds differ only in their class_embed parameters)
    san = kind in ("san_frame", "san_video")
    zs = kind == "zero_shot"
    T, C = mask_features.shape[:2]
    if num_layers is None:
        num_layers = 1 + max(int(k.split(".")[1]) for k in P if k.startswith("transformer_ffn_layers."))
    Qn = P["query_feat.weight"].shape[0]
    G = 1 if video else T                       # attention groups: one clip, or one per frame

    size_list, src, pos = [], [], []
    for l in range(3):
        h, w = x[l].shape[-2:]
        size_list.append((h, w))
        s = x[l].flatten(2) + P["level_embed.weight"][l][None, :, None]     # [T, C, hw]
        s = s.permute(0, 2, 1)                                              # [T, hw, C]
        if video:
            p = sine_pos_3d(T, h, w, C // 2).to(dtype).flatten(2).permute(0, 2, 1)   # [T, hw, C]
            src.append(s.reshape(1, T * h * w, C))
            pos.append(p.reshape(1, T * h * w, C))
        else:
            p = sine_pos_2d(h, w, C // 2).to(dtype).flatten(1).T                      # [hw, C]
            src.append(s)
            pos.append(p[None].expand(T, -1, -1))

    E = P["query_embed.weight"][None].expand(G, -1, -1)
    Z = P["query_feat.weight"][None].expand(G, -1, -1)

    attn_features = san_attn_features(P, mask_features, clip_heads) if san else None

    def heads(Z, target_hw):
        D = layer_norm(Z, P["decoder_norm.weight"], P["decoder_norm.bias"])         # [G, Q, C]
        cls = None
        if san:
            ae = mlp(P, "attn_embed", D, 3)
            if video:
                cls = torch.einsum("bqc,tnchw->btnqhw", ae, attn_features)           # b = 1
            else:
                cls = torch.einsum("bqc,bnchw->bnqhw", ae, attn_features)
        elif zs:
            cls = (mlp(P, "object_embed", D, 2), D)                                  # zero_shot_...:249, 265
        elif "class_embed.weight" in P:
            cls = D @ P["class_embed.weight"].T + P["class_embed.bias"]
        elif "class_embed.layers.0.weight" in P:
            cls = mlp(P, "class_embed", D, 2)
        me = mlp(P, "mask_embed", D, 3)
        if video:
            m = torch.einsum("bqc,tchw->bqthw", me, mask_features)                   # [1, Q, T, H, W]
            am = attn_mask_from_logits(m.flatten(0, 1), target_hw)                   # [Q, T, hw]
            blocked = am.reshape(1, Qn, -1)
        else:
            m = torch.einsum("bqc,bchw->bqhw", me, mask_features)                    # [T, Q, H, W]
            blocked = attn_mask_from_logits(m, target_hw)                            # [T, Q, hw]
        return cls, m, blocked, D

    pred_cls, pred_mask, attn_masks = [], [], []
    cls, m, blocked, D = heads(Z, size_list[0])
    pred_cls.append(cls), pred_mask.append(m), attn_masks.append(blocked)
    for i in range(num_layers):
        l = i % 3
        blocked = unblock_full_rows(blocked)
        pre = f"transformer_cross_attention_layers.{i}."
        a = mha(Z + E, src[l] + pos[l], src[l],
                P[pre + "multihead_attn.in_proj_weight"], P[pre + "multihead_attn.in_proj_bias"],
                P[pre + "multihead_attn.out_proj.weight"], P[pre + "multihead_attn.out_proj.bias"],
                nheads, blocked)
        Z = layer_norm(Z + a, P[pre + "norm.weight"], P[pre + "norm.bias"])
        pre = f"transformer_self_attention_layers.{i}."
        a = mha(Z + E, Z + E, Z,
                P[pre + "self_attn.in_proj_weight"], P[pre + "self_attn.in_proj_bias"],
                P[pre + "self_attn.out_proj.weight"], P[pre + "self_attn.out_proj.bias"], nheads)
        Z = layer_norm(Z + a, P[pre + "norm.weight"], P[pre + "norm.bias"])
        pre = f"transformer_ffn_layers.{i}."
        f = (Z @ P[pre + "linear1.weight"].T + P[pre + "linear1.bias"]).relu()
        f = f @ P[pre + "linear2.weight"].T + P[pre + "linear2.bias"]
        Z = layer_norm(Z + f, P[pre + "norm.weight"], P[pre + "norm.bias"])
        cls, m, blocked, D = heads(Z, size_list[(i + 1) % 3])
        pred_cls.append(cls), pred_mask.append(m), attn_masks.append(blocked)

    out = {}
    if zs:                                                         # zero_shot_...:236-243, 268-277
        out = {"pred_object_logits": pred_cls[-1][0], "pred_logits": pred_cls[-1][1], "pred_masks": pred_mask[-1],
               "pred_embeds": D,
               "aux_outputs": [{"pred_object_logits": a[0], "pred_logits": a[1], "pred_masks": b}
                               for a, b in zip(pred_cls[:-1], pred_mask[:-1])]}
        if return_attn_masks:
            out["attn_masks"] = attn_masks
        return out
    if not video:
        # '(b t) q h w -> b q t h w' with b = 1 (frame_...:113-121)
        pred_mask = [m.permute(1, 0, 2, 3)[None] for m in pred_mask]
        if san:
            pred_cls = [c[None] for c in pred_cls]                 # [1, T, n, Q, h, w]
        elif pred_cls[0] is not None:
            pred_cls = [c[None] for c in pred_cls]                 # [1, T, Q, cls]
        out["pred_embeds"] = D[None]                               # [1, T, Q, C]  (frame_...:123-124)
        out["mask_feats"] = mask_features
        out["ms_feats"] = [s.permute(1, 0, 2) for s in src]        # [hw, T, C]
        out["ms_pos"] = [p.permute(1, 0, 2) for p in pos]
        out["size_list"] = size_list
    if san:
        out["class_attn_biases"] = pred_cls[-1]
        if not video:
            out["attn_feats"] = attn_features
        out["aux_outputs"] = [{"class_attn_biases": a, "pred_masks": b}
                              for a, b in zip(pred_cls[:-1], pred_mask[:-1])]
    else:
        out["pred_logits"] = pred_cls[-1]
        if pred_cls[0] is not None:
            out["aux_outputs"] = [{"pred_logits": a, "pred_masks": b}
                                  for a, b in zip(pred_cls[:-1], pred_mask[:-1])]
        else:
            out["aux_outputs"] = [{"pred_masks": b} for b in pred_mask[:-1]]
    out["pred_masks"] = pred_mask[-1]
    if return_attn_masks:
        out["attn_masks"] = attn_masks
    return out


# --------------------------------------------------------------------------------------------
# open-vocabulary tails
# --------------------------------------------------------------------------------------------
def ov_cosine_logits(feats, text, scale=100.0, normalize=True):
    """ClipAdapter.normalize + cal_sim_logits (clip_adapter/adapter.py:118-119, 146-147):
    scale * (f / ||f||) @ text^T.  SideAdapter.cal_sim_logits (side_adapter.py:234-235) is the same with
    scale = exp(logit_scale) on already-normalised features (normalize=False)."""
    f = feats / feats.norm(dim=-1, keepdim=True) if normalize else feats
    return scale * f @ text.T


def ov2seg_logits(x, text, temperature=50.0):
    """OV2Seg ZeroShotClassifier.forward after its `linear` (openvis/ov2seg.py:519-526): a zero row is appended to the
    text matrix, x is L2-normalised and scaled by norm_temperature (50), logits = einsum('bqc,nc->bqn')."""
    zs = torch.cat([text, torch.zeros_like(text)[0:1]])
    return torch.einsum("bqc,nc->bqn", temperature * F.normalize(x, p=2, dim=-1), zs)


def openvis_clip_aggregate(clip_cls, valid_flag):
    """OpenVIS.open_vocabulary_inference tail (openvis/openvis.py:123-141).
    clip_cls [R, K]: logits of the valid (frame, query) regions in row-major order of valid_flag [T, Q].
    Returns (probs [Qv, K], valid_query_flag [Q]): per-query mean over its valid frames, softmax."""
    ids = torch.nonzero(valid_flag)
    vq = valid_flag.sum(0) > 0
    rows = [clip_cls[ids[:, 1] == q].mean(0) for q in torch.nonzero(vq)[:, 0]]
    return torch.stack(rows).softmax(-1), vq


def san_sos_tail(sos, ln_w, ln_b, proj, text, logit_scale_exp):
    """SideAdapter.post_encode_image tail + cal_sim_logits (side_adapter.py:201-207, 234-235):
    ln_post -> @ visual.proj -> F.normalize -> exp(logit_scale) * f @ text^T.
    sos [B, Q, W]; returns (clip_feats [B, Q, D], logits [B, Q, K+1])."""
    f = layer_norm(sos, ln_w, ln_b) @ proj
    f = F.normalize(f, dim=-1)
    return f, logit_scale_exp * f @ text.T


def adaptive_windows(n_in, n_out):
    """index windows of F.adaptive_max_pool2d: [floor(i*n_in/n_out), ceil((i+1)*n_in/n_out))."""
    return [((i * n_in) // n_out, -((-(i + 1) * n_in) // n_out)) for i in range(n_out)]


def san_pool_bias(attn_bias, grid_hw):
    """SideAdapter._build_attn_biases step 1 (side_adapter.py:241-250): adaptive max-pool of
    [B, n, Q, h, w] to the CLIP grid -> [B*n, Q, gh*gw].  Written with explicit windows."""
    b, n, q, h, w = attn_bias.shape
    gh, gw = grid_hw
    out = attn_bias.new_empty(b, n, q, gh, gw)
    for i, (y0, y1) in enumerate(adaptive_windows(h, gh)):
        for j, (x0, x1) in enumerate(adaptive_windows(w, gw)):
            out[..., i, j] = attn_bias[..., y0:y1, x0:x1].amax(dim=(-1, -2))
    return out.reshape(b * n, q, gh * gw)


def san_build_attn_bias(attn_bias, grid_hw):
    """SideAdapter._build_attn_biases (side_adapter.py:237-270) for one bias tensor with num_head == the
    CLIP head count: returns the additive [B*n, Q+1+L, Q+1+L] matrix.
    Columns :Q are -100 (nobody attends SOS tokens) except the SOS diagonal (0); SOS->CLS is -100;
    SOS->patch = pooled bias; everything else 0."""
    pooled = san_pool_bias(attn_bias, grid_hw)
    bn, q, L = pooled.shape
    n = q + 1 + L
    m = pooled.new_zeros(n, n)
    m[:, :q] = -100
    m[:q, q] = -100
    m[torch.arange(q), torch.arange(q)] = 0
    m = m[None].expand(bn, -1, -1).clone()
    m[:, :q, -L:] = pooled
    return m


def ms_deform_attn(value, spatial_shapes, sampling_locations, attention_weights):
    """Multi-scale deformable attention forward, restated tap by tap from the reference's CUDA kernel
    (ops/src/cuda/ms_deform_im2col_cuda.cuh:18-66, 243-305): h_im = y * H - 0.5, w_im = x * W - 0.5, samples outside
    (-1, H) x (-1, W) and corners outside the map contribute zero.  (The reference's own debug function,
    ms_deform_attn_core_pytorch, ops/functions/ms_deform_attn_func.py:55-77, does the same through F.grid_sample.)
    value [N, S, M, D], spatial_shapes [L, 2] (H, W), sampling_locations [N, Lq, M, L, P, 2] (x, y),
    attention_weights [N, Lq, M, L, P] -> [N, Lq, M*D]."""
    N, S, M, D = value.shape
    _, Lq, _, L_, P, _ = sampling_locations.shape
    out = value.new_zeros(N, Lq, M, D)
    start = 0
    bi = torch.arange(N)[:, None, None, None]
    mi = torch.arange(M)[None, None, :, None]
    for l in range(L_):
        H, W = int(spatial_shapes[l][0]), int(spatial_shapes[l][1])
        v = value[:, start:start + H * W]                                  # [N, H*W, M, D]
        start += H * W
        x = sampling_locations[:, :, :, l, :, 0] * W - 0.5                 # [N, Lq, M, P]
        y = sampling_locations[:, :, :, l, :, 1] * H - 0.5
        inside = (y > -1) & (x > -1) & (y < H) & (x < W)
        y0, x0 = torch.floor(y), torch.floor(x)
        ly, lx = y - y0, x - x0
        acc = value.new_zeros(N, Lq, M, P, D)
        for dy, dx, wgt in ((0, 0, (1 - ly) * (1 - lx)), (0, 1, (1 - ly) * lx), (1, 0, ly * (1 - lx)), (1, 1, ly * lx)):
            yy, xx = (y0 + dy).long(), (x0 + dx).long()
            ok = inside & (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
            idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1))            # [N, Lq, M, P]
            g = v[bi, idx, mi]                                             # [N, Lq, M, P, D]
            acc = acc + (wgt * ok)[..., None] * g
        out = out + (acc * attention_weights[:, :, :, l, :, None]).sum(3)
    return out.reshape(N, Lq, M * D)


def ms_deform_attn_module(P, query, reference_points, input_flatten, spatial_shapes, padding_mask=None, n_heads=8, n_points=4):
    """MSDeformAttn.forward (openvis/modeling/pixel_decoder/ops/modules/ms_deform_attn.py:83-125): P holds the module's
    state dict (sampling_offsets / attention_weights / value_proj / output_proj .weight / .bias); spatial_shapes [L, 2]."""
    N, Lq, C = query.shape
    L_ = spatial_shapes.shape[0]
    lin = lambda x, n: x @ P[n + ".weight"].T + P[n + ".bias"]
    value = lin(input_flatten, "value_proj")
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], 0.0)
    value = value.view(N, -1, n_heads, C // n_heads)
    off = lin(query, "sampling_offsets").view(N, Lq, n_heads, L_, n_points, 2)
    w = lin(query, "attention_weights").view(N, Lq, n_heads, L_ * n_points).softmax(-1).view(N, Lq, n_heads, L_, n_points)
    if reference_points.shape[-1] == 2:
        norm = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).float()
        loc = reference_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + off / n_points * reference_points[:, :, None, :, None, 2:] * 0.5
    return lin(ms_deform_attn(value, spatial_shapes, loc, w), "output_proj")


def video_postprocess(pred_cls, pred_masks, padded_size, img_size, out_hw, topk=10):
    """VideoMaskFormer.postprocess + inference_video (video_maskformer.py:215-229, 262-298) on pred_cls [Q, K] scores and
    stride-4 pred_masks [Q, T, h4, w4]: up-sample to the padded size, top-k over Q*K, crop, resize, > 0.
    Returns (scores [k], labels [k], query index [k], entropy [k], masks bool [k, T, H, W]) sorted by score."""
    Q, K = pred_cls.shape
    up = F.interpolate(pred_masks, size=tuple(padded_size), mode="bilinear", align_corners=False)
    labels = torch.arange(K).unsqueeze(0).repeat(Q, 1).flatten(0, 1)
    sc, idx = pred_cls.flatten(0, 1).topk(topk, sorted=True)
    lab = labels[idx]
    qi = idx // K
    ent = torch.sum(-pred_cls[qi] * torch.log(pred_cls[qi]), dim=-1)
    m = up[qi][:, :, : img_size[0], : img_size[1]]
    m = F.interpolate(m, size=tuple(out_hw), mode="bilinear", align_corners=False)
    return sc, lab, qi, ent, m > 0.0, m


def clip_block(P, i, x, attn_mask, nheads=12):
    """BiasedResidualAttentionBlock.forward (side_adapter.py:70-78) over mask_adapted_clip's ResidualAttentionBlock
    (model.py:237-268): x + MHA(ln_1(x), additive float mask) ; x + c_proj(QuickGELU(c_fc(ln_2(x)))).
    x [Lt, n, W] sequence-first like the reference; attn_mask [n*heads, Lt, Lt] additive fp32."""
    Lt, n, Wd = x.shape
    d = Wd // nheads
    y = layer_norm(x, P[f"{i}.ln_1.weight"], P[f"{i}.ln_1.bias"])
    qkv = y @ P[f"{i}.attn.in_proj_weight"].T + P[f"{i}.attn.in_proj_bias"]
    q, k, v = qkv.split(Wd, dim=-1)
    # [Lt, n, W] -> [n*heads, Lt, d]  (F.multi_head_attention_forward: batch index = b * heads + h)
    shp = lambda t: t.reshape(Lt, n * nheads, d).transpose(0, 1)
    q, k, v = shp(q) * d ** -0.5, shp(k), shp(v)
    s_ = q @ k.transpose(-1, -2)
    if attn_mask is not None:
        s_ = s_ + attn_mask
    a = (s_.softmax(-1) @ v).transpose(0, 1).reshape(Lt, n, Wd)
    x = x + (a @ P[f"{i}.attn.out_proj.weight"].T + P[f"{i}.attn.out_proj.bias"])
    y = layer_norm(x, P[f"{i}.ln_2.weight"], P[f"{i}.ln_2.bias"])
    h = y @ P[f"{i}.mlp.c_fc.weight"].T + P[f"{i}.mlp.c_fc.bias"]
    h = h * torch.sigmoid(1.702 * h)                                   # QuickGELU (model.py:232-234)
    return x + (h @ P[f"{i}.mlp.c_proj.weight"].T + P[f"{i}.mlp.c_proj.bias"])


def san_post_blocks(P, cls_token, pix_feat, attn_bias, num_queries, blocks=(9, 10, 11), nheads=12):
    """SideAdapter.post_encode_image up to (not including) ln_post (side_adapter.py:176-199):
    cls_token [1, n, W], pix_feat [n, W, h, w], attn_bias [n, heads, Q, H', W'].
    Token order [Q SOS copies of CLS | CLS | h*w patches]; the same additive bias matrix for every block.
    Returns the SOS tokens [n, Q, W]."""
    n, c, h, w = pix_feat.shape
    x = torch.cat([cls_token, pix_feat.reshape(n, c, -1).permute(2, 0, 1)])          # [1+L, n, W]
    sos = cls_token.repeat(num_queries, 1, 1)
    mask = san_build_attn_bias(attn_bias, (h, w)) if attn_bias is not None else None
    x = torch.cat([sos, x], dim=0)
    for i in blocks:
        x = clip_block(P, i, x, mask, nheads)
    return x[:num_queries].permute(1, 0, 2)


# --------------------------------------------------------------------------------------------
# seeded synthetic parameters / inputs: shared with bench.py, so they live in the (oracle-free) product package
# --------------------------------------------------------------------------------------------
from openvis_b200.synthetic import (clip_block_param_shapes, decoder_param_shapes, seeded_clip_block_params,  # noqa: E402,F401
                                    seeded_inputs, seeded_params)
