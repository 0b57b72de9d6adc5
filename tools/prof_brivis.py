"""BASELINE config 3 (BriVIS R50 online, LV-VIS vocabulary K = 1196 + background, 36 frames of 360x640 -> 384x640,
Q = 100) and config 4's shape (720x1280, Q = 200) through the whole device path: SAN frame decoder -> query matching ->
TemporalInstanceResampler (CLIP side path in its last head) -> post-processing.  Per-stage CUDA-event times for one clip."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openvis_b200 import _lib as L, decoder as D, temporal as T
from openvis_b200.ov_head import SideAdapterBlocks
from openvis_b200.synthetic import decoder_param_shapes, seeded_params, seeded_clip_block_params, seeded_resampler_params


def ev(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, Tn, Hp, Wp, Q, img, out_hw in (("cfg3_brivis_36x360x640_q100_k1197", 36, 384, 640, 100, (360, 640), (360, 640)),
                                         ("cfg4shape_36x720x1280_q200_k1197", 36, 736, 1280, 200, (720, 1280), (720, 1280))):
    K = 1197
    g = torch.Generator(device="cuda").manual_seed(1)
    kw = dict(in_channels=256, mask_classification=True, num_classes=1, hidden_dim=256, num_queries=Q, nheads=8,
              dim_feedforward=2048, dec_layers=9, pre_norm=False, mask_dim=256, enforce_input_project=False, num_frames=2,
              clip_heads=12)
    dec = D.SideAdapterFrameMultiScaleMaskedTransformerDecoder(**kw)
    dec.load_state_dict(seeded_params(decoder_param_shapes("san_frame", Q=Q), 0))
    dec = dec.cuda().eval()
    res = T.TemporalInstanceResampler().eval()
    res.load_state_dict(seeded_resampler_params(0))
    res = res.cuda()
    cg = torch.Generator().manual_seed(0)
    sd = {f"transformer.resblocks.{k}": v for k, v in seeded_clip_block_params(1).items()}
    sd.update({"ln_post.weight": torch.ones(768), "ln_post.bias": torch.zeros(768), "proj": torch.randn(768, 512, generator=cg) * 768 ** -0.5})
    ad = SideAdapterBlocks(num_queries=Q).load_clip_visual_state_dict(sd)
    x = [torch.randn(Tn, 256, Hp // 32 * 2 ** l, Wp // 32 * 2 ** l, generator=g, device="cuda") for l in range(3)]
    mf = torch.randn(Tn, 256, Hp // 4, Wp // 4, generator=g, device="cuda")
    bk = (torch.randn(1, Tn, 768, generator=g, device="cuda"), torch.randn(Tn, 768, 14, 14, generator=g, device="cuda"))
    text = torch.nn.functional.normalize(torch.randn(K, 512, generator=g, device="cuda"), dim=-1)
    run = lambda exact=False: T.brivis_video_inference(dec, ad, res, x, mf, bk, text, (Hp, Wp), img, out_hw[0], out_hw[1], api_exact=exact)
    n0 = L.launch_count()
    run()
    nl = L.launch_count() - n0
    if os.environ.get("OVIS_PROF_ONCE"):          # ncu launch list: two calls of config 3, nothing else
        run(); torch.cuda.synchronize(); break
    total = ev(run)
    total_exact = ev(lambda: run(True), n=3, warm=1)
    # stages
    out = dec(x, mf)
    t_dec = ev(lambda: dec(x, mf))
    emb = out["pred_embeds"][0][None]
    t_match = ev(lambda: T.batch_video_match_via_embeds(emb))
    idx, fe = T.batch_video_match_via_embeds(emb)
    out = dec(x, mf)
    res.operand_source = dec
    t_res = ev(lambda: res(fe, out["mask_feats"], out["attn_feats"], ad, bk, text))
    o = res(fe, out["mask_feats"], out["attn_feats"], ad, bk, text)
    from openvis_b200.postprocess import inference_video
    lg = o["pred_logits"][0].float().contiguous()
    ones = torch.ones(Tn, Q, dtype=torch.uint8, device="cuda")

    def post():
        probs, _ = L.clip_aggregate(lg, ones)
        return inference_video(Q, K - 1, probs[:, :-1].contiguous(), o["pred_masks"][0], (Hp, Wp), img, out_hw[0], out_hw[1])

    t_post = ev(post)
    # several clips per call (config 5's throughput setting): the launch chain costs the same for 36 or 144 frames
    multi = {}
    for nc in ((2, 4) if Q == 100 else (2,)):
        xs = [t.repeat(nc, 1, 1, 1) for t in x]
        mfs = mf.repeat(nc, 1, 1, 1)
        bks = (bk[0].repeat(1, nc, 1), bk[1].repeat(nc, 1, 1, 1))
        runm = lambda: T.brivis_video_inference(dec, ad, res, xs, mfs, bks, text, (Hp, Wp), img, out_hw[0], out_hw[1], num_clips=nc)
        ms = ev(runm, n=3, warm=2)
        multi[f"{nc}_clips_per_call"] = {"ms_per_call": round(ms, 3), "frames_per_s": round(nc * Tn / ms * 1e3)}
        del xs, mfs, bks
        torch.cuda.empty_cache()
    print(json.dumps({"workload": name, "clips_per_call": multi, "ms_per_clip": round(total, 3), "frames_per_s": round(Tn / total * 1e3),
                      "launches_per_clip": nl, "api_exact_ms_per_clip": round(total_exact, 3),
                      "stages_ms": {"san_frame_decoder": round(t_dec, 3), "query_matching": round(t_match, 3),
                                    "resampler_incl_clip_side_path": round(t_res, 3),
                                    "post_processing_incl_d2h": round(t_post, 3)}}))
    del dec, res, ad, x, mf, out, o
    torch.cuda.empty_cache()
