/* far3d_b200.h - C ABI of libfar3d_sm100.so (hand-written sm_100a CUDA for Far3D's per-frame forward).
 *
 * Drop-in boundary for the operator layer of megvii-research/Far3D @ 5efb9d7.
 * The reference has no native code; its operator boundary is the Python call into third-party
 * libraries.  Each entry point below names the reference call it replaces (file:line relative to
 * projects/mmdet3d_plugin/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - tensors are dense, row-major in the order written in the comment; fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - no allocation; the entry points keep no state between calls and are thread-safe per stream.  The ONLY process-wide
 *     state is what the far3d_*_tune* / far3d_conv_umma_debug experiment hooks at the end of this header set (tools and tests
 *     use them to select kernel variants; a product caller never calls them and then every launch depends on its arguments only);
 *   - return 0 on success, <0 on error (FAR3D_E_*); never throws; far3d_last_error() gives a
 *     thread-local message for the last failure.
 */
#ifndef FAR3D_B200_H
#define FAR3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAR3D_OK 0
#define FAR3D_E_INVALID (-1)     /* bad argument (null pointer, non-positive size, misalignment) */
#define FAR3D_E_UNSUPPORTED (-2) /* shape outside what the kernels are built for */
#define FAR3D_E_CUDA (-3)        /* CUDA runtime / launch error */

#define FAR3D_MAX_LEVELS 8

const char* far3d_last_error(void);
int far3d_abi_version(void);
/* number of kernels this library has launched in this process */
int64_t far3d_launch_count(void);
/* account for kernels re-launched by a CUDA-graph replay captured from n of our launches */
void far3d_add_launches(int64_t n);

/* ---------------------------------------------------------------------------------------------
 * Perspective-aware deformable aggregation (fused).
 * Replaces DeformableFeatureAggregationCuda.feature_sampling, models/utils/detr3d_transformer.py:544-569
 * (projection :547-552, the 21 MB repeat :555, MultiScaleDeformableAttnFunction.apply :561-563, camera
 * sum :565-569) in ONE kernel; ABI modelled on the reference's dead Sparse4D wrapper
 * models/utils/deformable_aggregation.py:17-29.
 *   feat       [B*N, S, C]        channels-last multi-camera multi-level features (S = sum H_l*W_l)
 *   hw_host    [L][2] (H_l, W_l)  HOST int32;  start_host [L] HOST int32 (level start index)
 *   key_points [B, Nq, P, 3]      metres, lidar frame
 *   lidar2img  [B, N, 4, 4]
 *   weights    [B*N, Nq, G, L*P]  post-softmax, layout of `_get_weights` (:541-542)
 *   out        [B, Nq, C]
 * feat_dtype: 0 = fp32, 1 = bf16, 2 = fp16 (weights/points/out stay fp32).
 */
int far3d_deform_agg_fwd(const void* feat, int feat_dtype, const int32_t* hw_host, const int32_t* start_host,
                         const float* key_points, const float* lidar2img, const float* weights, float pad_h,
                         float pad_w, float* out, int B, int N, int S, int C, int G, int Nq, int L, int P,
                         void* stream);

/* Tools / tests: kernel variant of far3d_deform_agg_fwd - warps per CTA (4 = default: a query's 8 channel groups over two work
 * items; 8: one item per query; 2: four items) and wide (bit 0, default 1: 256-bit loads, two samples per warp instruction; 0:
 * 128-bit, one sample; bit 1: one resident wave of CTAs pulling work items from a device-side queue instead of one CTA per
 * item; bit 2: 4 instead of 8 two-sample loads in flight per lane; bit 3: far3d_dfa_prepare with 256-thread CTAs). */
void far3d_deform_agg_tune(int warps, int wide);

/* Tools / tests: far3d_mha_fwd[_masked] kernel form - bit 0 clear (default): warp-level tensor-core MMAs with split fp16 operands
 * (mha_mma_d32_kernel), set: the round-1 SIMT kernel (mha_d32_kernel, packed fp32 FMAs); bits 4-7: key groups per CTA of the
 * tensor-core kernel (1..4, 0 = default 3). */
void far3d_mha_tune(int simt);

/* Debug companion of the fused op: same projection + bounds arithmetic, dumps
 *   uv [B,N,Nq,P,2] fp32, idx [B,N,Nq,L,P,2] int32 (h_low,w_low), valid [B,N,Nq,L,P] uint8. */
int far3d_deform_agg_debug(const int32_t* hw_host, const float* key_points, const float* lidar2img, float pad_h,
                           float pad_w, float* uv, int32_t* idx, uint8_t* valid, int B, int N, int Nq, int L,
                           int P, void* stream);

/* Exact-layout replacement of mmcv MultiScaleDeformableAttnFunction.forward
 * (call site models/utils/detr3d_transformer.py:561-563):
 *   value [BN,S,G,D], spatial_shapes [L,2] int64 (device), level_start_index [L] int64 (device),
 *   sampling_locations [BN,Nq,G,L,P,2], attention_weights [BN,Nq,G,L*P] -> out [BN,Nq,G*D]. */
int far3d_msda_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                   const float* sampling_locations, const float* attention_weights, float* out, int BN, int S,
                   int G, int D, int Nq, int L, int P, void* stream);

/* Softmax of the aggregation weights, detr3d_transformer.py:539-542, using linearity of weights_fc:
 *   logits[b,q,n,j] = wq[b,q,j] + wc[b,n,j],  j = lp*G + g  (G fastest, :540)
 *   softmax over (n, lp) per (b,q,g)  ->  weights [B*N, Nq, G, LP]  (the layout :541-542 returns). */
int far3d_dfa_weights_softmax(const float* wq, const float* wc, float* weights, int B, int N, int Nq, int G,
                              int LP, void* stream);

/* Camera side of the aggregation logits for ALL decoder layers in one launch (detr3d_transformer.py:530-540):
 *   out[l, r, :] = W_fc[l] . LayerNorm(relu(W1[l] relu(W0[l] x_r + b0[l]) + b1[l])),  x_r = the first 12 values of lidar2img row r
 * lidar2img [rows, 16] (rows = B*N cameras); layer_ptrs: HOST array of 7 device pointers per layer (cam_embed.0.weight [H,12],
 * .0.bias, cam_embed.2.weight [E,H], .2.bias, cam_embed.4.weight, .4.bias, weights_fc.weight [J,E]); out [layers, rows, J]. */
#define FAR3D_MAX_CAM_LAYERS 8
int far3d_cam_logits(const float* lidar2img, const float* const* layer_ptrs, int layers, int rows, int E, int H, int J, float eps,
                     float* out, void* stream);

/* The same two steps (detr3d_transformer.py:539-542 softmax, :547-569 projection + sampling + camera sum) cut differently: the
 * softmax kernel, one CTA per query, also projects the query's key points and emits
 *   cnt [B*Nq] int32               in-view samples of the query (bounds rule of mmcv's ms_deformable_im2col)
 *   rec [B*Nq, N*L*P, 4] {u32, f32}  per in-view sample, (camera, point, level) order: the 4 bilinear corners as {row index in
 *                                    units of 4 channels, corner weight}; an out-of-map corner aliases an in-map one with weight 0
 *   wts [B*Nq, G, N*L*P] f32        the softmax weights of exactly those samples, by position
 * (only the first cnt entries of a row are written / read; `weights`, the full [B*N, Nq, G, L*P] tensor, is optional), and
 * far3d_deform_agg_gather does the feature gather from them - same sums in the same order as far3d_deform_agg_fwd, bit for bit.
 * far3d_dfa_prepare_supported: 1 when the shape is inside this path (32-channel groups, N*L*P <= 512, ...), else use the pair
 * far3d_dfa_weights_softmax + far3d_deform_agg_fwd. */
int far3d_dfa_prepare_supported(int N, int G, int L, int P, int C);
int far3d_dfa_prepare(const float* wq, const float* wc, const float* key_points, const float* lidar2img, const int32_t* hw_host,
                      const int32_t* start_host, float pad_h, float pad_w, int B, int N, int Nq, int G, int L, int P, int S, int C,
                      float* weights, int32_t* cnt, void* rec, float* wts, void* stream);
int far3d_deform_agg_gather(const void* feat, int feat_dtype, const int32_t* hw_host, const int32_t* start_host, const int32_t* cnt,
                            const void* rec, const float* wts, float* out, int B, int N, int S, int C, int G, int Nq, int L, int P,
                            void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense layers.  y[M,N] = act((x (+ x_add))[M,K] @ w[N,K]^T + bias) (+ residual[M,N]);  torch.nn.Linear semantics
 * (call sites: detr3d_transformer.py:503-512, farhead.py:228-282, mmcv FFN / MultiheadAttention).
 * x_add (optional, same shape/stride as x) fuses the `query + query_pos` additions of the attention blocks.
 * act: 0 none, 1 relu.  ldx/ldy/ldr = row strides in elements.  fp32 SIMT path (exact fp32 FMA). */
int far3d_linear_f32(const float* x, const float* x_add, int ldx, const float* w, const float* bias,
                     const float* residual, int ldr, float* y, int ldy, int M, int N, int K, int act, void* stream);

/* LayerNorm over the last dim (C <= 1024), y = LN(x (+ add)) * gamma + beta; optional relu_before
 * (applies ReLU to the input first, detr3d_transformer.py:506-512 cam_embed tail). */
int far3d_layernorm(const float* x, const float* add, const float* gamma, const float* beta, float* y, int M,
                    int C, float eps, int relu_before, int relu_after, void* stream);

/* Multi-head attention core (torch.nn.MultiheadAttention inside mmcv MultiheadAttention, far3d.py:112-116):
 *   q [B,Nq,H*Dh] (already projected, unscaled), k,v [B,Nk,H*Dh] -> o [B,Nq,H*Dh] = softmax(q k^T/sqrt(Dh)) v.
 *   Dh must be 32. ldq/ldk/ldv/ldo = row strides. */
int far3d_mha_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo,
                  int B, int Nq, int Nk, int H, int Dh, void* stream);
/* same with a hole in the key set: keys [key_skip[0], key_skip[0] + key_skip[1]) are ignored.  key_skip is a DEVICE int32[2]
 * (written by far3d_query2d_lift's caller): the padding rows of a bucketed adaptive-query count, which a replayed CUDA
 * graph must mask without a new capture (yolox_head.py:454-467 / farhead.py:585-602 make the count data dependent). */
int far3d_mha_fwd_masked(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* o, int ldo, int B,
                         int Nq, int Nk, int H, int Dh, const int32_t* key_skip, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 2D proposals -> adaptive 3D queries, on the device with fixed capacities (no boolean gathers, no host round trips).
 * far3d_roi_select replaces YOLOXHeadCustom.get_bboxes (models/dense_heads/yolox_head.py:355-489): score =
 *   sigmoid(obj) * sigmoid(max cls), 3x3 local-max peak pick, score > threshold, box decode (:491-501), per camera in the
 *   reference's order (level-major, row-major) into slots [N, cap_per_cam]; counts[n] may exceed the capacity (overflow).
 *   cls_host / reg_host: HOST arrays of L device pointers to NHWC fp32 maps [N,H_l,W_l,cls_cs] / [N,H_l,W_l,reg_cs]
 *   (reg channels 0-3 box, 4 objectness); hw_host [L][2], stride_host [L]: HOST int32; score_ws: N * sum(H_l*W_l) floats.
 * far3d_query2d_lift replaces FarHead.build_query2d_proposal (models/dense_heads/farhead.py:710-827) for depth-logit input:
 *   depth-bin softmax + top-k at the box centre, multi-depth duplicates (k-major after the primaries), un-projection with
 *   inverse(lidar2img) and pc_range normalisation -> ref2d [cap_total,3]; src_row = feat_flatten row of each query's peak,
 *   score_feat = (logit(score) - thr_logit) * depth-score ratio; meta = {queries, primaries, multi-depth sources, overflow};
 *   rows >= meta[0] are padding.  far3d_ctx_gather builds the context rows [cap_total, C+1] (farhead.py:585-590, :757-763). */
int far3d_roi_select(const void* const* cls_host, const void* const* reg_host, const int32_t* hw_host,
                     const int32_t* stride_host, int L, int N, int num_classes, int cls_cs, int reg_cs, float threshold,
                     float* score_ws, int cap_per_cam, int32_t* sel_pos, float* sel_score, float* sel_box, int32_t* counts,
                     void* stream);
int64_t far3d_query2d_lift_workspace_ints(int cap_total);
int far3d_query2d_lift(const int32_t* sel_pos, const float* sel_score, const float* sel_box, const int32_t* counts, int N,
                       int cap_per_cam, int S, const float* depth_logits, int Hd, int Wd, int D, int Dcs, int down, int topk,
                       int rmin_bin, float dmin, float bin_size, float thr_logit, const float* lidar2img,
                       const float* pc_range, int cap_total, float* ref2d, int32_t* src_row, float* score_feat,
                       int32_t* meta, int32_t* workspace, void* stream);
int far3d_ctx_gather(const float* feat_flatten, const int32_t* src_row, const float* score_feat, int C, int rows, float* ctx,
                     void* stream);

/* 3D position encoder input, models/utils/positional_encoding.py:13-25: pos [M,3] -> emb [M,3*F] (order y,x,z),
 * and the 1D (:27-36) / NeRF (:38-80) encodings used by farhead.py:284-313. */
int far3d_pos2posemb3d(const float* pos, float* emb, int M, int F, void* stream);
int far3d_pos2posemb1d(const float* pos, int ldp, float* emb, int M, int F, void* stream);
int far3d_nerf_posenc(const float* x, float* emb, int M, int Cin, int nfreq, void* stream);

/* MLN spatial alignment + flatten + level concat in one pass (farhead.py:553-567, misc.py:182-190):
 *   out[bn, start + hw, c] = gamma[bn,c] * x[bn,hw,c] + beta[bn,c];  x is NHWC (channels_last==1) or NCHW. */
int far3d_mln_flatten(const float* x, const float* gamma, const float* beta, float* out, int BN, int HW, int C,
                      int S, int start, int channels_last, void* stream);
/* generic MLN apply on tokens: out[m,c] = gamma[m,c]*x[m,c] + beta[m,c] with optional LN(no affine) of x first */
int far3d_mln_tokens(const float* x, const float* gamma, const float* beta, float* out, int M, int C, int use_ln,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backbone / neck (models/backbones/vovnet.py, mmdet FPN).  Activations are NHWC.
 *
 * far3d_conv2d_umma: implicit-GEMM convolution on tcgen05 tensor cores (fp16 operands, fp32 TMEM
 * accumulators, TMA-fed), kernel 1x1 or 3x3 (pad k/2), stride 1 or 2, fused bias (folded BN) + activation
 * (relu: 0 none, 1 ReLU, 2 Swish).
 * Replaces nn.Conv2d + BatchNorm2d(eval) + ReLU triples of vovnet.py:124-161 and FPN convs.
 *   x_hi/x_lo  fp16 NHWC [N,H,W,x_cs] read at channel offset x_co, Cin channels.  x_lo == NULL: plain fp16.
 *              x_lo != NULL: split-fp16 ("fp16x3") mode: value = hi + lo, products hi*hi + lo*hi + hi*lo
 *              give fp32-grade accuracy (2^-17) on fp16 tensor cores.
 *   w_hi/w_lo  fp16 [Cout, k*k, Cin]  (tap-major, Cin contiguous); w_lo required iff x_lo given.
 *   bias       fp32 [Cout] or NULL;  relu: 0/1
 *   outputs (any subset, each NHWC with its own channel stride/offset; image stride = Ho*Wo*cs unless y_f32_ns>0):
 *     y_f32 fp32, y_hi fp16, y_lo fp16 (residual y - fp16(y)).
 */
int far3d_conv2d_umma(const void* x_hi, const void* x_lo, int N, int H, int W, int x_cs, int x_co, int Cin,
                      const void* w_hi, const void* w_lo, const float* bias, int Cout, int ksize, int stride,
                      int relu, float* y_f32, int yf_cs, int yf_co, int64_t yf_ns, void* y_hi, void* y_lo,
                      int yb_cs, int yb_co, void* stream);

/* "fp16mx" operand format (2 tensor-pipe passes per MAC instead of the 3 of fp16x3; same reference calls replaced).
 * The lo plane of a tensor is replaced by an e4m3 CORRECTION plane of the same size (2 bytes per element):
 *     lo8 = e4m3((v - hi) * 2^(11+EA))   and   hi8 = e4m3(hi * 2^EA),    hi = fp16(v),
 * laid out per pixel row and 32-channel group g as 64 bytes [lo8 of channels 32g..32g+31 | hi8 of the same channels]
 * (channel counts, strides and offsets must be multiples of 32).  `lo_fmt` / `x_fmt` / `y_fmt` arguments name the format of a
 * lo plane: 0 = fp16 residual plane, FAR3D_LO_MX(EA) = this format with activation exponent EA.
 * Weights: w_c8 [Cout, k*k, Cin] in the same layout with w_hi8 = e4m3(w_hi * 2^w_exp) in the "lo8" slot and
 * w_lo8 = e4m3((w - w_hi) * 2^(w_exp+11)) in the "hi8" slot, so that MMA block 0 = a_lo8*w_hi8 and block 1 = a_hi8*w_lo8.
 * Both correction products carry the factor 2^(11+EA+w_exp); the fp16 weight plane `w_hi` of the *_mx entry points holds
 * fp16(w * 2^(11+EA+w_exp)) (w_hi above = that plane * 2^-(11+EA+w_exp)), so the conv accumulates hi*w_hi (kind::f16) and
 * both correction products (kind::f8f6f4, e4m3 x e4m3) into one fp32 TMEM accumulator at one scale and the epilogue
 * divides it out.  (Round 2 used kind::mxf8f6f4.block_scale with an unscaled w_hi plane.) */
#define FAR3D_LO_FP16 0
#define FAR3D_LO_MX(EA) (64 + (EA))
int far3d_conv2d_umma_mx(const void* x_hi, const void* x_c8, int x_fmt, int N, int H, int W, int x_cs, int x_co, int Cin,
                         const void* w_hi, const void* w_c8, int w_exp, const float* bias, int Cout, int ksize, int stride,
                         int relu, float* y_f32, int yf_cs, int yf_co, int64_t yf_ns, void* y_hi, void* y_lo, int y_fmt,
                         int yb_cs, int yb_co, void* stream);
int far3d_conv2d_umma_pool_mx(const void* x_hi, const void* x_c8, int x_fmt, int N, int H, int W, int x_cs, int x_co, int Cin,
                              const void* w_hi, const void* w_c8, int w_exp, const float* bias, int Cout, int relu,
                              float* y_f32, int yf_cs, int yf_co, float* workspace, float* mean, void* stream);

/* 1x1 far3d_conv2d_umma (the OSA concat conv, vovnet.py:230-232) that also returns the global average pool of its fp32
 * output, mean[N, Cout] (eSEModule's AdaptiveAvgPool2d(1), vovnet.py:173-185): the conv epilogue writes per-tile column
 * sums to `workspace` (>= far3d_conv_pool_workspace_floats(N, H, W, Cout) floats, deterministic: no atomics) and a
 * second small kernel folds them per image.  Saves one full HBM pass over the block output. */
int64_t far3d_conv_pool_workspace_floats(int N, int H, int W, int Cout);
int far3d_conv2d_umma_pool(const void* x_hi, const void* x_lo, int N, int H, int W, int x_cs, int x_co, int Cin,
                           const void* w_hi, const void* w_lo, const float* bias, int Cout, int relu, float* y_f32,
                           int yf_cs, int yf_co, float* workspace, float* mean, void* stream);

/* nn.Linear on the tensor cores (same kernel, a [rows,K] matrix is a 1 x M image): y = act(x @ w^T + bias) (+ residual).
 * x_hi/x_lo fp16 [M, ldx], w_hi/w_lo fp16 [N, K] (lo planes NULL = plain fp16); bias [N], residual [M, ldr], y [M, ldy] fp32.
 * Replaces the decoder's cuBLAS GEMMs (mmcv MultiheadAttention / FFN, detr3d_transformer.py:503-512, farhead.py:228-282). */
int far3d_linear_umma(const void* x_hi, const void* x_lo, int ldx, const void* w_hi, const void* w_lo, const float* bias,
                      const float* residual, int ldr, float* y, int ldy, int M, int N, int K, int act, void* stream);

/* The same nn.Linear for the decoder's token matrices (M ~ 1000 rows, K a multiple of 32) on warp-level tensor-core MMAs
 * (mma.sync.m16n8k16, three MMAs per product on fp16 hi / lo planes, fp32 accumulate: the fp16x3 result): x / x_add fp32 [M, ldx]
 * are split on the fly (no far3d_split_fp16 launch), w_hi / w_lo fp16 [N, K] as for far3d_linear_umma; any N. */
int far3d_linear_mma(const float* x, const float* x_add, int ldx, const void* w_hi, const void* w_lo, const float* bias,
                     const float* residual, int ldr, float* y, int ldy, int M, int N, int K, int act, void* stream);

/* fp32 SIMT implicit-GEMM convolution (exact fp32 FMA), same semantics, NHWC fp32 in/out; the correctness
 * anchor for the tensor-core path and the fallback for shapes the UMMA kernel does not take (Cin % 8 != 0). */
int far3d_conv2d_f32(const float* x, int N, int H, int W, int x_cs, int x_co, int Cin, const float* w /*[Cout,k*k,Cin]*/,
                     const float* bias, int Cout, int ksize, int stride, int relu, float* y, int y_cs, int y_co,
                     void* stream);

/* Camera images as they leave the decoder: uint8 [N,H,W,3] (cv2 channel order) -> normalised, zero-padded fp32 [N,3,Hp,Wp].
 * Replaces, on the device, NormalizeMultiviewImage (datasets/pipelines/transform_3d.py:74-101: mmcv.imnormalize =
 * (float32(x) - mean) * (1 / float64(std)), BGR->RGB swap first when to_rgb), AV2PadMultiViewImage (pad_val 0,
 * custom_pipeline.py:358-378) and the HWC->CHW transpose of the format bundle, so a frame crosses PCIe as 1 byte per
 * sample.  mean_host / std_host: HOST float[3], in the order of the OUTPUT channels (as mmcv applies them after the swap). */
int far3d_normalize_u8(const uint8_t* img_nhwc, int N, int H, int W, int Hp, int Wp, const float* mean_host,
                       const float* std_host, int to_rgb, float* out_nchw, void* stream);

/* Camera-frame resize + crop (+ horizontal flip), bit-exact with the Pillow calls of ResizeCropFlipRotImage._img_transform
 * (custom_pipeline.py:277-311: img.resize(resize_dims) with Pillow's default BICUBIC filter, img.crop(crop) with zero fill
 * outside the image, FLIP_LEFT_RIGHT) as AV2ResizeCropFlipRotImageV2.__call__ (custom_pipeline.py:48-149) applies it per view.
 * far3d_resample_ksize / far3d_resample_coeffs (HOST, no GPU needed): Pillow's 22-bit fixed-point tap tables of one axis -
 * bounds[out][2] = (first source index, tap count), k[out][ksize].  far3d_resize_crop_u8: src uint8 [H,W,3] on the device,
 * tables on the device, computes the window [crop_x0, crop_x0+out_w) x [crop_y0, crop_y0+out_h) of the new_w x new_h image into
 * dst (uint8 HWC, row stride dst_row_pixels pixels).  [y_first, y_first+rows) = source rows the window's vertical taps touch
 * (min / max over ybounds of the in-image window rows; rows == 0 when the window misses the image), tmp = rows*out_w*3 bytes. */
int far3d_resample_ksize(int in_size, int out_size);
int far3d_resample_coeffs(int in_size, int out_size, int* bounds_host, int* k_host);
int far3d_resize_crop_u8(const uint8_t* src_hwc, int H, int W, int new_w, int new_h, const int* xbounds, const int* xk, int xksize,
                         const int* ybounds, const int* yk, int yksize, int y_first, int rows, int crop_x0, int crop_y0,
                         int out_w, int out_h, int flip, uint8_t* tmp, uint8_t* dst_hwc, int dst_row_pixels, void* stream);

/* Stem conv 1 (vovnet.py:308): NCHW fp32 image -> NHWC, 3x3 stride 2 pad 1, Cin=3, fused BN+ReLU.
 * Outputs like far3d_conv2d_umma (fp32 and/or split fp16). w [Cout,3,3,3] as (Cout, ky, kx, cin). */
int far3d_stem_conv(const float* img_nchw, int N, int H, int W, const float* w, const float* bias, int Cout,
                    float* y_f32, void* y_hi, void* y_lo, int lo_fmt, void* stream);

/* MaxPool2d(3, stride 2, ceil_mode=True) NHWC (vovnet.py:249) on fp16 (hi[/lo]) or fp32 data.
 * Reads channels [x_co, x_co+C) of a tensor with channel stride x_cs, writes likewise. dtype: 0 fp32, 1 fp16.
 * For split data pool the recombined value and re-split (max is not linear). */
int far3d_maxpool3x3s2(const void* x_hi, const void* x_lo, int dtype, int N, int H, int W, int C, int x_cs, int x_co,
                       void* y_hi, void* y_lo, int y_cs, int y_co, int lo_fmt /* of x_lo and y_lo */, void* stream);

/* eSE (vovnet.py:173-185) in three steps: global average pool of xt (fp32 NHWC [N,HW,C]) -> mean [N,C];
 * gate [N,C] = relu6(fc(mean)+3)/6; then y = xt*gate (+identity), written as fp32 and/or split fp16. */
#define FAR3D_AVGPOOL_CHUNKS 64
/* workspace: N*FAR3D_AVGPOOL_CHUNKS*C floats (two-stage deterministic reduction) or NULL (single-stage) */
int far3d_global_avgpool(const float* x, float* mean, float* workspace, int N, int HW, int C, void* stream);
int far3d_ese_gate(const float* mean, const float* fc_w, const float* fc_b, float* gate, int N, int C, void* stream);
int far3d_ese_apply(const float* xt, const float* gate, const float* id_f32, const void* id_hi, const void* id_lo,
                    int id_cs, int id_co, int N, int HW, int C, float* y_f32, int yf_cs, int yf_co, void* y_hi,
                    void* y_lo, int yb_cs, int yb_co, int lo_fmt /* of id_lo and y_lo */, void* stream);

/* FPN top-down (mmdet FPN.forward): dst[n,h,w,c] += src[n, h*Hs/Hd, w*Ws/Wd, c] (nearest), fp32 NHWC in place,
 * also emits split fp16 copies of dst for the following 3x3 conv. */
int far3d_upsample_add(float* dst, const float* src, int N, int Hd, int Wd, int Hs, int Ws, int C, void* d_hi,
                       void* d_lo, int lo_fmt, void* stream);

/* GroupNorm over NHWC fp32 (+ optional ReLU), models/depth_predictor/depth_predictor.py:44-46; outputs fp32 and/or
 * split fp16. */
#define FAR3D_GN_CHUNKS 64
/* workspace: N*FAR3D_GN_CHUNKS*2*groups floats (coalesced two-pass path) or NULL (single-kernel path) */
int far3d_groupnorm_nhwc(const float* x, const float* gamma, const float* beta, float* workspace, int N, int HW, int C,
                         int groups, float eps, int relu, float* y_f32, void* y_hi, void* y_lo, int lo_fmt, void* stream);

/* fp32 -> split fp16 (hi, lo) and back; layout-preserving elementwise helpers. n = element count. */
int far3d_split_fp16(const float* x, const float* x_add /* optional, summed first */, void* hi, void* lo, int64_t n,
                     void* stream);
int far3d_merge_fp16(const void* hi, const void* lo, float* y, int64_t n, void* stream);
/* strided variants: rows x C with channel stride/offset on the fp16 side */
int far3d_merge_fp16_strided(const void* hi, const void* lo, int lo_fmt, int cs, int co, float* y, int64_t rows, int C,
                             void* stream);
/* dense rows x C fp32 -> hi plane + lo plane in either format (C %% 8 == 0) */
int far3d_split_planes(const float* x, void* hi, void* lo, int lo_fmt, int64_t rows, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused NMS-free box decode of one sample.  Replaces NMSFreeCoder.decode_single (core/bbox/coders/nms_free_coder.py:39-112:
 * sigmoid, top-max_num over queries x classes, label / query index, denormalize_bbox core/bbox/util.py:25-52,
 * post_center_range mask, optional score threshold) and, with bottom_center, the z shift of FarHead.get_bboxes
 * (models/dense_heads/farhead.py:1224-1245).  cls [Nq, C] logits, box [Nq, code] (code 8 or 10); outputs are fixed-size
 * [max_num, 7 | 9] / [max_num] arrays holding the *out_count surviving boxes first, in descending score order (ties: lowest
 * flat index first); post_center_range_host: HOST float[6]; score_threshold <= 0: none. */
int far3d_box_decode(const float* cls, const float* box, int Nq, int C, int code, int max_num,
                     const float* post_center_range_host, float score_threshold, int bottom_center, float* out_boxes,
                     float* out_scores, int32_t* out_labels, int32_t* out_query, int32_t* out_count, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Temporal memory bank of one stream (B = 1), models/dense_heads/farhead.py:446-508.
 * far3d_memory_post_update replaces post_update_memory (:479-508): top-K queries by max-class score of the last decoder layer
 *   (topk_idx [K] out, descending), their embedding / reference point / velocity pushed in front of the M old rows, every row
 *   moved by ego_pose (4x4 products for the poses, point transforms for the reference points), timestamps (fp64) minus the
 *   frame's.  cls_last [Nq,C] logits, box_last [Nq,code] (first 3 = centre in metres, last 2 = velocity), dec_last [Nq,E];
 *   old bank o_* with M rows; new bank n_* with K + M rows (emb [.,E], ref [.,3], ts [.] fp64, pose [.,16], velo [.,2]).
 * far3d_memory_pre_update replaces pre_update_memory (:446-477) on an existing bank: first n rows moved by ego_pose_inv,
 *   timestamps plus the frame's, everything times prev_exists, and (1 - prev_exists) * pseudo reference points [kprop,3] /
 *   identity poses added to the first kprop rows. */
int far3d_memory_post_update(const float* cls_last, const float* box_last, const float* dec_last, int Nq, int C, int code, int E,
                             int K, int M, const float* ego_pose, const double* timestamp, const float* o_emb,
                             const float* o_ref, const double* o_ts, const float* o_pose, const float* o_velo,
                             int32_t* topk_idx, float* n_emb, float* n_ref, double* n_ts, float* n_pose, float* n_velo,
                             void* stream);
int far3d_memory_pre_update(int n, int E, int kprop, const float* prev_exists, const float* ego_pose_inv,
                            const double* timestamp, const float* pseudo_points, const float* o_emb, const float* o_ref,
                            const double* o_ts, const float* o_pose, const float* o_velo, float* n_emb, float* n_ref,
                            double* n_ts, float* n_pose, float* n_velo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Experiment hooks (tools/, tests/): process-wide kernel-variant switches, NOT part of the reference-facing surface.
 * Defaults (never calling them) are the product configuration. */
void far3d_conv_umma_tune(int bn, int stages);          /* force the N tile / ring depth (0 = heuristic) */
void far3d_conv_umma_tune2(int grid, int halo);         /* persistent grid size (0 = one CTA per SM); halo -1 = generic mode only */
void far3d_conv_umma_tune4(int cta_group);              /* 0 heuristic, 1 single-CTA kernel, 2 CTA pairs wherever legal */
void far3d_conv_umma_tune6(float loss_per_mma);         /* accumulator-truncation compensation constant (0 = off) */
void far3d_conv_umma_tune7(int smem_reserve_bytes);     /* shared memory per SM the conv kernels leave free */
void far3d_conv_umma_tune8(int pdl);                    /* 1: conv launches carry the programmatic-dependent-launch attribute */
void far3d_conv_umma_debug(void* timestamps);           /* per-CTA timeline buffer (device pointer) or NULL */

#ifdef __cplusplus
}
#endif
#endif /* FAR3D_B200_H */
