/* openvis_b200 -- C ABI of the B200-native (sm_100a) decoder + mask-head + open-vocabulary-head kernels.
 *
 * Drop-in boundary for the per-frame decoding hot path of clownrat6/OpenVIS.  The reference has no FFI for
 * this path (it is PyTorch: nn.MultiheadAttention / einsum / F.interpolate); each entry point below names the
 * reference call site (file:line, relative to the reference root) whose arithmetic it replaces.  The only
 * native interface in the reference, `ms_deform_attn_forward` (openvis/modeling/pixel_decoder/ops/src/vision.cpp:18-21),
 * is the model for the conventions used here: raw device pointers, explicit sizes, the caller's CUDA stream,
 * contiguity/arch checks up front, no CPU fallback (ops/src/cpu/ms_deform_attn_cpu.cpp:22-32 raises).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; "f16" buffers hold IEEE half;
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the device;
 *   - return value 0 = success, otherwise an OVIS_ERR_* code; ovis_last_error() gives the message;
 *   - all kernels require compute capability 10.x (B200); anything else returns OVIS_ERR_ARCH.
 */
#ifndef OPENVIS_B200_H
#define OPENVIS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define OVIS_OK 0
#define OVIS_ERR_ARG 1     /* bad shape / alignment / null pointer */
#define OVIS_ERR_ARCH 2    /* not an sm_100 device, or no CUDA device */
#define OVIS_ERR_CUDA 3    /* CUDA runtime / driver error (launch, tensor-map encode) */

int ovis_version(void);
const char* ovis_last_error(void);
/* 0 when the current device can run the kernels (sm_100). */
int ovis_device_check(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim). */
long long ovis_launch_count(void);
/* Launches replayed through a captured CUDA graph do not pass through the entry points: the host side adds them here. */
void ovis_add_launch_count(long long n);

/* ---- layout preparation ------------------------------------------------------------------------------
 * Multi-scale feature map x_l [B][C][N] fp32 (NCHW, N=h*w) -> token-major fp16 [B][N][C].
 * Replaces `src[-1].permute(2, 0, 1)` (frame_mask2former_transformer_decoder.py:65-69).  Optionally also writes
 * out_pos = fp16(x + pos[n] + pos_t[b]), the `memory + pos` key operand (video_...decoder.py:115-116), with
 * pos [N][C] = level_embed + 2-D sine embedding and pos_t [B][C] (or null) the frame term of the 3-D embedding. */
int ovis_nchw_to_tokens_f16(const float* in, void* out_f16, void* out_pos_f16, const float* pos, const float* pos_t,
                            int B, int C, int N, void* stream);
/* Same operation for [B][C][h][w] inputs through the TMA-fed kernel (bulk tensor loads / stores): the position table
 * is passed channel-major, pos_cn [C][h*w] (the input's own layout).  Requires C % 32 == 0, w % 4 == 0 and 16-byte
 * aligned pointers; bit-identical results to ovis_nchw_to_tokens_f16. */
int ovis_nchw_to_tokens_hw_f16(const float* in, void* out_f16, void* out_pos_f16, const float* pos_cn, const float* pos_t,
                               int B, int C, int h, int w, void* stream);
/* mask_features [B][C][H][W] fp32 -> ft [B][H*W][C] f16 and centre-2x2-pooled g0/g1/g2 [B][(H/s)(W/s)][C] f16
 * (s = 8, 4, 2).  Replaces the operand side of einsum("bqc,bchw->bqhw") + F.interpolate(bilinear)
 * (frame_...decoder.py:144-148).  Requires H % 8 == 0, W % 8 == 0, C % 32 == 0. */
int ovis_maskfeat_prep(const float* F, void* ft_f16, void* g0_f16, void* g1_f16, void* g2_f16,
                       int B, int C, int H, int W, void* stream);
int ovis_cast_f16(const float* in, void* out_f16, long long n, void* stream);
/* query_feat / query_embed broadcast + decoder_norm for the first prediction head
 * (frame_...decoder.py:73-74, 78, 140).  rows = G*Q, hidden 256. */
int ovis_init_queries(const float* query_feat, const float* query_embed, const float* dn_g, const float* dn_b,
                      float* z32, void* z16, void* ze16, float* d32, void* d16, int Q, int rows, void* stream);
/* Row-wise (mode&1) LayerNorm eps 1e-5 and/or (mode&2) L2 normalisation.  ClipAdapter.normalize
 * (clip_adapter/adapter.py:118-119); SideAdapter ln_post / F.normalize (clip_adapter/side_adapter.py:203-205). */
int ovis_rownorm(const float* in, const float* g, const float* b, float* out32, void* out16, int rows, int D,
                 int mode, void* stream);
/* Kernel 3 of the path, fused form ("mask-pool GEMM with the L2-normalise and text-cosine epilogue"): the logits GEMM
 * normalises in its epilogue, logits[r][k] = scale * (x[r] . text[k]) / max(||x[r]||, 1e-12), instead of a separate
 * normalisation pass (ClipAdapter.normalize + cal_sim_logits, clip_adapter/adapter.py:118-119, 146-147; F.normalize +
 * SideAdapter.cal_sim_logits, side_adapter.py:205, 234-235).
 * ovis_rowstats: out16 = fp16(x) (after an optional LayerNorm: ln_post, side_adapter.py:203) and ss[r] = sum of squares of
 *   that fp16 row; ss_zero (optional) is cleared (accumulator of a following row_ss_out).  group_rows > 0: input row r is
 *   row (r / group_rows) * group_stride + r % group_rows of `in` (the Q SOS tokens at the head of every frame's token block).
 * ovis_linear_rowscale_f16: out = (x @ w^T + bias) * scale [* rsqrt(row_ss_in[r])]; row_ss_out[r] += sum of squares of the
 *   output row (visual.proj GEMM feeding the normalised logits GEMM). */
int ovis_rowstats(const float* in, const float* ln_g, const float* ln_b, void* out_f16, float* ss, float* ss_zero, int rows,
                  int D, int layer_norm, int group_rows, int group_stride, void* stream);
int ovis_linear_rowscale_f16(const void* x_f16, long long rows, int K, int ldx, const void* w_f16, int N, const float* bias,
                             float scale, const float* row_ss_in, float* row_ss_out, void* out, int ldo, int out_f32,
                             void* stream);

/* ---- tcgen05 GEMM family: out = x[rows][K] * w[N][K]^T, fp16 operands, fp32 accumulation ---------------
 * nn.Linear / in_proj / out_proj / MLP layers (video_mask2former_transformer_decoder.py:57-58, 115-118, 176-178,
 * 204-216); cal_sim_logits (clip_adapter/adapter.py:146-147, side_adapter.py:234-235).
 * out = act((acc + bias) * scale); ldx/ldo in elements; K % 64 == 0; x, w 16-byte aligned. */
int ovis_linear_f16(const void* x_f16, long long rows, int K, int ldx, const void* w_f16, int N,
                    const float* bias, float scale, int relu, void* out, int ldo, int out_f32, void* stream);
/* The same with an activation selector and an optional fp32 residual added after it (residual connections of the
 * CLIP blocks of the SAN side path, mask_adapted_clip/model.py:265-268):
 * out = act((acc + bias) * scale) + resid;  act: 0 none, 1 ReLU, 2 QuickGELU x*sigmoid(1.702x) (model.py:232-234);
 * resid [rows][ldo] fp32 or null, may alias out. */
int ovis_linear_act_f16(const void* x_f16, long long rows, int K, int ldx, const void* w_f16, int N,
                        const float* bias, float scale, int act, const float* resid, void* out, int ldo,
                        int out_f32, void* stream);
/* Linear(K -> 256) + residual + LayerNorm [+ second LayerNorm], the post-norm tails of
 * CrossAttentionLayer/SelfAttentionLayer/FFNLayer.forward_post (video_...decoder.py:119-120, 59-60, 177-178)
 * fused with decoder_norm (frame_...decoder.py:140).  Any output pointer may be null.
 * ype16 = fp16(y + pe[row % pe_period]) is the "+ query_pos" operand of the next projection.
 * split_ws (optional, >= (K/256) * ceil128(rows) * 256 floats): scratch for the few-rows path (split-K partials
 * + a row-parallel LayerNorm kernel) taken when rows <= 16384; without it the fused single-pass epilogue is used. */
int ovis_linear_ln_f16(const void* x_f16, long long rows, int K, const void* w_f16, const float* bias,
                       const float* resid, const float* ln1_g, const float* ln1_b,
                       const float* ln2_g, const float* ln2_b, const float* pe, int pe_period,
                       float* y32, void* y16, void* ype16, float* d32, void* d16,
                       float* split_ws, long long split_ws_floats, void* stream);
/* Key/value projections of all decoder layers that read one feature level, in one launch: for n_tiles 256-wide
 * column tiles t, out[t] = (t even ? xk : xv) * w[t*256:(t+1)*256]^T + bias[t]; xk = memory + pos (+ level_embed),
 * xv = memory, as in `key=self.with_pos_embed(memory, pos), value=memory` (video_...decoder.py:115-118).
 * out / bias: HOST arrays of n_tiles device pointers (bias entries may be null). */
int ovis_kv_proj_f16(const void* xk_f16, const void* xv_f16, long long rows, const void* w_f16, int n_tiles,
                     void* const* out_host, const float* const* bias_host, void* stream);
/* Next-layer attention mask: bits[g][r/32][q] bit (r%32) = (g_l[g][r] . mask_embed[g][q] < 0) and
 * flags[g][q] = 1 when some key is unblocked.  Replaces einsum + F.interpolate + sigmoid() < 0.5 + repeat(heads)
 * (frame_...decoder.py:144-152) and feeds the all-masked-row rule (frame_...decoder.py:87).
 * flags must be zeroed by the caller. */
int ovis_mask_bits(const void* gt_f16, int groups, int rows_per_group, const void* me_f16, int Q,
                   unsigned int* bits, unsigned char* flags, int q_stride, void* stream);
/* Full-resolution mask logits out[g*t_group_stride + q*ldt + r] = ft[g][r] . mask_embed[g][q] (+ bias[q]).
 * einsum("bqc,bchw->bqhw") / ("bqc,btchw->bqthw") (frame_...:144, video_...:459); also the 1x1-conv output of the
 * SAN attention-bias branch when bias != null (side_adapter_frame_...decoder.py:67-71).
 * posflags (optional, zeroed by the caller) [frames][Q]: set to 1 when frame/query has a positive logit, i.e. a
 * non-empty mask -- the `valid` test of ClipAdapter._preprocess_image (clip_adapter/adapter.py:86-88). */
int ovis_mask_logits(const void* ft_f16, int groups, int rows_per_group, const void* me_f16, int me_group_stride,
                     int Q, const float* bias, float* out, long long t_group_stride, long long ldt,
                     unsigned char* posflags, int rows_per_frame, void* stream);
/* SAN per-head attention biases (side_adapter_frame_...decoder.py:157, einsum "bqc,bnchw->bnqhw"):
 * out[b][n][q][p] = af[b][p][n*256:(n+1)*256] . attn_embed[b][q];  af is the fp16 token-major copy
 * [B][P][heads*256] of attn_features. */
int ovis_san_bias_logits(const void* af_f16, int B, int P, int heads, const void* ae_f16, int Q, float* out,
                         void* stream);

/* ---- attention ---------------------------------------------------------------------------------------- */
/* Work-space sizing for ovis_xattn: chooses the key split count for (G, Q, keys) on the current device.
 * ml_part_floats includes, behind the (max, sum) partials, the tile-skip bitmap (512 words per group and 128-query tile). */
int ovis_xattn_plan(int G, int Q, int keys, int* splits, int* q_pad, long long* o_part_floats,
                    long long* ml_part_floats);
/* Masked multi-head cross-attention, 8 heads x 32 (CrossAttentionLayer -> nn.MultiheadAttention,
 * video_...decoder.py:110-122).  q [G*Q][256] f16 pre-scaled by 32^-1/2 * log2(e); k, v [G*keys][256] f16;
 * out [G*Q][256] f16 = concat of heads before out_proj.
 * 64-key tiles that every query of a 128-query tile blocks are skipped (no load, no MMA, no softmax step; rows under the
 * all-masked-row rule block nothing): xattn_skipmap_kernel marks them, each CTA takes an equal share of the surviving
 * tiles.  OVIS_XATTN_SKIP=0 disables it (A/B timing). */
int ovis_xattn(const void* q_f16, const void* k_f16, const void* v_f16, const unsigned int* bits,
               const unsigned char* flags, int G, int Q, int q_stride, int keys, int splits,
               float* o_part, float* ml_part, void* out_f16, void* stream);
/* Transposed-score variant of the same attention (xattn_tc3_kernel: thread = key, softmax reference and row sums folded
 * into the tensor-core products; csrc/xattn_tc3.cuh) for launches with long key runs per CTA.  It reads the mask in
 * KEY-major form, produced by ovis_mask_bits_t:
 *   bits_t   [G][keys][qw] u32, qw = 4 * ceil(Q / 128): bit q % 32 of word q / 32 = key blocked for query q (1 past Q)
 *   blockand [G][ceil(keys / 32)][qw]: AND of bits_t over each block of 32 keys (source of the tile-skip map)
 * ovis_xattn_plan_t: *use_t = 1 when this variant should be used for (G, Q, keys) on the current device (else use
 * ovis_xattn_plan / ovis_mask_bits / ovis_xattn); sizes as ovis_xattn_plan (ml_part_floats includes the skip map).
 * ovis_xattn_t: arguments as ovis_xattn; `splits` must come from ovis_xattn_plan_t; stats (optional, device int[2]):
 * [0] += CTAs, [1] += CTAs that re-ran their chunk with an exact softmax reference. */
int ovis_xattn_plan_t(int G, int Q, int keys, int* use_t, int* splits, int* q_pad, long long* o_part_floats,
                      long long* ml_part_floats);
int ovis_mask_bits_t(const void* gt_f16, int groups, int rows_per_group, const void* me_f16, int Q,
                     unsigned int* bits_t, unsigned int* blockand, unsigned char* flags, int q_stride, void* stream);
int ovis_xattn_t(const void* q_f16, const void* k_f16, const void* v_f16, const unsigned int* bits_t,
                 const unsigned int* blockand, const unsigned char* flags, int G, int Q, int q_stride, int keys,
                 int splits, float* o_part, float* ml_part, void* out_f16, int* stats, void* stream);
/* Query-side chain (csrc/chain.cuh): the post-attention part of a decoder layer -- out-projection + LayerNorm,
 * self-attention, FFN, decoder_norm, mask-embed MLP, the next layer's query projection (SelfAttentionLayer /
 * CrossAttentionLayer / FFNLayer.forward_post, MLP: video_...decoder.py:52-62, 110-122, 175-179, 204-216) -- as ONE launch:
 * a persistent CTA per group of Q <= 128 queries executes the phases in order on its own rows.  A chain is a handle holding
 * `nphases` phases for G groups; phases are set once (same arguments as ovis_linear_f16 / ovis_linear_ln_f16 /
 * ovis_self_attn, rows = G * Q implied; every operand must stay at its address), checked for completeness
 * (ovis_chain_upload), then runs of at most 12 consecutive phases are launched: ovis_chain_run(handle, first, count,
 * stream); the phases of a run travel as the launch's kernel parameter. */
int ovis_chain_create(int nphases, int G, int Q, void** handle);
/* wide = 1 (before any phase is set): every phase's 128-row tiles are spread over all CTAs of ONE cooperative launch and a
 * grid-wide barrier separates the phases -- for calls with many groups (Frame decoders: a group per frame), where the
 * one-CTA-per-group chain would serialise a layer on a few SMs.  Same phases, same results. */
int ovis_chain_set_wide(void* handle, int wide);
/* wide chains: fp32 workspace of >= (K_max / 256) * ceil128(G * Q) * 256 floats for the linear + LayerNorm phases (K slices ->
 * partial products -> row-parallel reduction + LayerNorm instead of the row-serial epilogue); set before those phases. */
int ovis_chain_set_scratch(void* handle, float* ws, long long floats);
/* wide chains: plain GEMM phase idx (already set) does not feed phase idx + 1, so no barrier separates them and their tiles go
 * to different CTAs (K/V-style sibling projections, the next layer's query projection beside the mask-embed MLP). */
int ovis_chain_set_parallel(void* handle, int idx, int parallel);
int ovis_chain_set_linear(void* handle, int idx, const void* x_f16, int K, int ldx, const void* w_f16, int N,
                          const float* bias, float scale, int relu, void* out, int ldo, int out_f32);
int ovis_chain_set_linear_ln(void* handle, int idx, const void* x_f16, int K, const void* w_f16, const float* bias,
                             const float* resid, const float* ln1_g, const float* ln1_b, const float* ln2_g,
                             const float* ln2_b, const float* pe, int pe_period, float* y32, void* y16, void* ype16,
                             float* d32, void* d16);
int ovis_chain_set_self_attn(void* handle, int idx, const void* qk_f16, const void* v_f16, void* out_f16);
int ovis_chain_upload(void* handle);
int ovis_chain_run(void* handle, int first, int count, void* stream);
/* profiling only: as ovis_chain_run; group 0's CTA also writes its SM cycle counter at the start of every phase and at the
 * end into trace[0 .. count] (device memory, int64) */
int ovis_chain_run_traced(void* handle, int first, int count, long long* trace, void* stream);
int ovis_chain_destroy(void* handle);
/* Unmasked self-attention over the Q queries (SelfAttentionLayer.forward_post, video_...decoder.py:52-62).
 * qk [G*Q][512] f16 (q | k, biased, unscaled), v [G*Q][256] f16 -> out [G*Q][256] f16.  Q <= 1536 rows per group
 * (the decoders use Q <= 256 queries; the temporal resampler attends over the frames of a clip, resampler.py:258-262). */
int ovis_self_attn(const void* qk_f16, const void* v_f16, void* out_f16, int G, int Q, void* stream);

/* ---- open-vocabulary head tails ----------------------------------------------------------------------- */
/* OpenVIS.open_vocabulary_inference aggregation (openvis/openvis.py:123-141): per-query mean of CLIP logits over
 * frames with a non-empty mask, softmax over K.  logits [T][Q][K], valid [T][Q] -> probs [Q][K], qvalid [Q]. */
int ovis_clip_aggregate(const float* logits, const unsigned char* valid, float* probs, unsigned char* qvalid,
                        int T, int Q, int K, void* stream);
/* ---- OpenVIS crop classifier front end --------------------------------------------
 * ClipAdapter._preprocess_image (openvis/modeling/clip_adapter/adapter.py:73-116) and the input half of encode_image
 * (adapter.py:140-143) + the token assembly of the CLIP visual tower (mask_adapted_clip model.py:327-342); the
 * transformer blocks themselves are the ovis_rownorm / ovis_linear_act_f16 / ovis_san_attn calls of the SAN side path.
 *  masks are addressed as masks[t * stride_t + n * stride_n + y * W + x] (elements): the [T][N][H][W] soft masks
 *  ClipAdapter.forward takes, or with logits = 1 the decoder's [N][T][H][W] mask logits, the sigmoid of openvis.py:118
 *  applied on load (no transposed fp32 copy).
 *  ovis_mask_boxes   valid[T * N] = any(mask > thresh), boxes[T * N][4] int32 = (x_min, y_min, x_max + 1, y_max + 1) of the
 *                    thresholded mask (detectron2 BitMasks.get_bounding_boxes; zeros when empty).  adapter.py:84-92
 *  ovis_crop_blend   for each of the M valid (frame, query) pairs `ids` [M][2] (row-major order of valid, adapter.py:103):
 *                    square box anchored at the top-left corner with side max(w, h) (:94-100), torchvision roi_align
 *                    (spatial_scale 1, sampling_ratio -1, aligned False) of the frame [T][3][H][W] and of the soft mask
 *                    [T][N][H][W] to R x R, regions = mask_region * frame_region, fp16 [M][3][R][R] (:106-113)
 *  ovis_clip_patchify regions -> (v / 255 - mean[c]) / std[c] as fp16 rows [M * (R/P)^2][3 * P * P] of the conv1 GEMM
 *  ovis_clip_embed   x [M * (1 + Lp)][width] fp32 = ln_pre([class_embedding | patch_tokens] + positional_embedding) */
int ovis_mask_boxes(const float* masks, int T, int N, long long stride_t, long long stride_n, int H, int W, int logits,
                    float thresh, int* boxes, unsigned char* valid, void* stream);
int ovis_crop_blend(const float* frames, const float* masks, long long stride_t, long long stride_n, int logits, const int* ids,
                    const int* boxes, int M, int T, int N, int H, int W, int R, void* regions_f16, void* stream);
int ovis_clip_patchify(const void* regions_f16, long long M, int R, int P, const float* mean, const float* std, void* out_f16,
                       void* stream);
int ovis_clip_embed(const float* patch_tokens, const float* class_embedding, const float* positional_embedding,
                    const float* ln_g, const float* ln_b, float* x, long long M, int Lp, int width, void* stream);
/* ---- multi-scale deformable attention, forward ---------------------------------
 * Replaces MSDA.ms_deform_attn_forward (openvis/modeling/pixel_decoder/ops/src/vision.cpp:18-21 ->
 * ms_deform_attn_cuda_forward, src/cuda/ms_deform_attn_cuda.cu:22-84 -> ms_deformable_im2col_gpu_kernel,
 * src/cuda/ms_deform_im2col_cuda.cuh:243-305), fp32:
 * value [N][S][M][D], spatial_shapes [L][2] int64 (H, W), level_start_index [L] int64, sampling_loc [N][Lq][M][L][P][2]
 * (x, y in [0,1]), attn_weight [N][Lq][M][L][P] -> out [N][Lq][M*D].  All device pointers.  No im2col_step: the whole
 * batch is one launch.  Backward (training) is out of scope. */
int ovis_ms_deform_attn_forward(const float* value, const long long* spatial_shapes, const long long* level_start_index,
                                const float* sampling_loc, const float* attn_weight, float* out, int N, int S, int M, int D,
                                int Lq, int L, int P, void* stream);
/* MSDeformAttn.forward between the query projections and the sampling op (ops/modules/ms_deform_attn.py:104-117): `proj`
 * [rows][M*L*P*3] fp32 is the output of ONE GEMM over the concatenated sampling_offsets | attention_weights weights (columns
 * [0, 2*M*L*P) offsets in (m, l, p, xy) order, then M*L*P attention logits); softmax over the L*P logits of a head and
 * sampling_locations = reference_points + offsets / (W_l, H_l) (ref_dim 2) or the reference-box form (ref_dim 4). */
int ovis_msda_prepare(const float* proj, const float* reference_points, const long long* spatial_shapes, long long rows, int M,
                      int L, int P, int ref_dim, float* sampling_loc, float* attn_weight, void* stream);
/* The same two steps fused (the pixel decoder's encoder path): softmax + sampling locations in registers, fp16 value map
 * [N][S][M][32] (the value projection's fp16 output), fp32 bilinear weights / accumulation, fp16 rows out [N*Lq][M*32] = the
 * A operand of output_proj.  proj as for ovis_msda_prepare; reference_points [N or 1][Lq][L][ref_dim] with ref_batch_stride
 * floats between samples (0 = one table for every sample).  Head width 32, L*P <= 16. */
int ovis_msda_fused_f16(const void* value_f16, const float* proj, const float* reference_points, long long ref_batch_stride,
                        const long long* spatial_shapes, const long long* level_start_index, void* out_f16, int N, int S, int M,
                        int Lq, int L, int P, int ref_dim, void* stream);
/* ---- device-side post-processing ----------------------------------------------
 * Top-k over the flattened [Q*K] scores with labels, query indices and per-query entropy:
 * scores.flatten(0, 1).topk(10), labels[topk], topk // num_classes, sum(-s log s)
 * (VideoMaskFormer.inference_video, video_maskformer.py:266-271).  k <= 32; outputs sorted by score. */
int ovis_topk_scores(const float* scores, int Q, int K, int k, float* out_scores, int* out_query, int* out_label,
                     float* out_entropy, void* stream);
/* For the selected queries: F.interpolate(pred_masks, padded size, bilinear) (video_maskformer.py:220-226 /
 * openvis.py:87-96), crop to the image, F.interpolate(output size, bilinear), `> 0` (:273-277), fused per output
 * pixel and bit-packed.  masks: frame t of query q at masks + q * q_stride + t * h4 * w4 (q_stride 0 = T*h4*w4; a
 * larger stride addresses one clip of a multi-clip call) -> bits [n_sel][T][out_h][ceil(out_w/32)] (bit x%32 of word x/32). */
int ovis_mask_postprocess(const float* masks, long long q_stride, const int* query, int n_sel, int T, int h4, int w4,
                          int pad_h, int pad_w, int img_h, int img_w, int out_h, int out_w, unsigned int* bits, void* stream);
/* SAN / BriVIS side path.  Adaptive max-pool of the per-head attention biases to the CLIP
 * grid only (step 1 of SideAdapter._build_attn_biases, side_adapter.py:241-250): bias [BN][Q][h][w] -> pooled [BN][Q][gh*gw]. */
int ovis_san_pool_bias(const float* bias, float* pooled, int BN, int Q, int h, int w, int gh, int gw, void* stream);
/* Attention of one post-split CLIP block (BiasedResidualAttentionBlock.attention, side_adapter.py:72-73) over
 * [Q SOS | CLS | L patches] tokens per frame, d = 64, with the additive bias matrix of _build_attn_biases
 * (side_adapter.py:252-266) applied from `pooled` without materialising it.
 * qkv [B*(Q+1+L)][3*heads*64] f16 (in_proj output, q|k|v), out [B*(Q+1+L)][heads*64] f16. */
int ovis_san_attn(const void* qkv_f16, const float* pooled, void* out_f16, int B, int Q, int L, int heads, void* stream);
/* SideAdapter._build_attn_biases (clip_adapter/side_adapter.py:237-270): adaptive max-pool to (gh, gw) fused with
 * the [Q+1+L]^2 additive-bias construction.  bias [B][n][Q][h][w] -> out [B*n][Q+1+L][Q+1+L]. */
int ovis_san_attn_bias(const float* bias, float* out, int BN, int Q, int h, int w, int gh, int gw, void* stream);

/* ---- temporal association --------------------------------------------------
 * Replicate-padded unfold along the frame axis: in [G][T][C] f16 -> out [G][T][taps*C] f16 with
 * out[g][t][k*C + c] = in[g][clamp(t + k - taps/2, 0, T-1)][c], so that Conv1d(C, C, taps, padding='same',
 * padding_mode='replicate') of the resampler's short-term aggregation (openvis/modeling/resampler.py:205-213, 262-264)
 * is ovis_linear_f16 / ovis_linear_ln_f16 with the weight laid out [C_out][taps*C_in].  taps odd, C % 8 == 0. */
int ovis_temporal_unfold_f16(const void* in_f16, void* out_f16, int G, int T, int C, int taps, void* stream);
/* Cosine-cost Hungarian matching of consecutive frames, all B*T problems at once (match_via_embeds,
 * openvis/modeling/minvis.py:28-41: 1 - cos, scipy linear_sum_assignment on target x current).
 * en [B][T][n][C] fp32 L2-normalised embeddings (ovis_rownorm mode 2).  Problem (b, i) matches frame i against frame
 * i-1 (frame 0 against itself): pi[b][i][a] = query of frame i assigned to query a of frame i-1 (exact optimum,
 * double-precision potentials).  cost [B][T][n][n] fp32 (row = target, column = current) receives the cost matrices;
 * it may be null when n*n floats fit shared memory (n <= ~215). */
int ovis_match_embeds(const float* en, int B, int T, int n, int C, float* cost, int* pi, void* stream);
/* The reference's frame-by-frame chain (batch_video_match_via_embeds, minvis.py:44-72) from the independent solves:
 * indices[b][i][j] = pi[b][i][indices[b][i-1][j]], indices[b][-1] = identity.  indices [B][T][n] int64. */
int ovis_match_compose(const int* pi, long long* indices, int B, int T, int n, void* stream);
/* out[b][t][q][:] = in[b][t][idx[b][t][q]][:]: batch_index (openvis/utils/index.py:4-11) as used on the embeddings
 * (minvis.py:57) and by BriVIS.reset_image_output_order (openvis/brivis.py:231-240).  Element strides of (b, t, q) are
 * explicit so [b,t,q,c] logits / embeddings and [b,q,t,h,w] masks are served in place of their transposes; `inner`
 * contiguous floats per (b, t, q).  in != out. */
int ovis_reorder_queries_f32(const float* in, const long long* idx, float* out, int B, int T, int n, long long inner,
                             long long stride_b, long long stride_t, long long stride_q, void* stream);

/* ---- pixel-decoder glue -----------------------------------------------------
 * GroupNorm(32, 256) on a token-major map (the `nn.GroupNorm(32, conv_dim)` after every projection / lateral / output
 * convolution, openvis/modeling/pixel_decoder/msdeformattn.py:227-236, 278-299).
 * ovis_gn_stats: x [B][S][256] fp32 -> stats [B][32][2] double (sum, sum of squares per sample and group; the caller zeroes
 * it).  ovis_gn_apply: y = (x - mean) * rstd * gamma + beta, then optionally `+ bilinear(add)` (F.interpolate(...,
 * mode="bilinear", align_corners=False) of an hs x ws map to the H x W positions, msdeformattn.py:342-343, 371; element
 * (b, c, p) of the added map at add[b*add_bs + c*add_cs + p*add_ps], so token-major and NCHW maps are both served), then
 * optionally ReLU; row (b, r) is written to row b*out_bs + out_off + r of out32 (fp32) and / or out16 (fp16). */
int ovis_gn_stats(const float* x, double* stats, int B, int S, void* stream);
int ovis_gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, float eps, int B, int H, int W,
                  int relu, const float* add, long long add_bs, long long add_cs, long long add_ps, int hs, int ws,
                  float* out32, void* out16, long long out_bs, long long out_off, void* stream);
/* Token-major rows in_off .. in_off+N of each of B blocks of in_bs rows of C floats -> out [B][C][N] fp32: the NCHW maps
 * forward_features returns (`z.transpose(1, 2).view(bs, -1, h, w)`, msdeformattn.py:358-359). */
int ovis_tokens_to_nchw_f32(const float* in, float* out, int B, int C, int N, long long in_bs, long long in_off, void* stream);
/* A operand of a 3x3 / padding-1 convolution as a GEMM (the FPN output convolution, msdeformattn.py:284-292):
 * in [B][H][W][C] f16 -> out [B*H*W][9*C] f16, out[(b,y,x)][(ky*3+kx)*C + c] = in[b][y+ky-1][x+kx-1][c] (zero outside);
 * the weight is laid out [C_out][(ky, kx, C_in)].  C % 8 == 0. */
int ovis_conv3x3_unfold_f16(const void* in_f16, void* out_f16, int B, int H, int W, int C, void* stream);

/* ---- token-major hand-off pixel decoder -> masked decoder (no NCHW fp32 round trip; 256 channels) -----------------------
 * ovis_tokens_pool_f16: ft [B][H][W][256] f16 -> g [B][H/s][W/s][256] f16 = fp32 mean of the centre 2x2 pixels of every s x s
 * block: the operand of the intermediate mask heads (F.interpolate(outputs_mask, size=attn_mask_target_size, "bilinear"),
 * frame_...decoder.py:146, for an integer factor s).
 * ovis_tokens_add_pos_f16: xp[b][n][:] = f16(xt[b][n][:] + pos[n][:] + pos_t[b][:]) -- the `memory + pos` key operand
 * (video_...decoder.py:115-116) from token-major f16 features; pos [N][256], pos_t [B][256] or null. */
int ovis_tokens_pool_f16(const void* ft_f16, void* g_f16, int B, int H, int W, int s, void* stream);
int ovis_tokens_add_pos_f16(const void* xt_f16, const float* pos, const float* pos_t, void* xp_f16, int B, int N, void* stream);

#ifdef __cplusplus
}
#endif
#endif
