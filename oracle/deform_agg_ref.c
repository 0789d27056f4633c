/* CPU oracle (TEST INFRASTRUCTURE) for the perspective-aware deformable aggregation.
 *
 * Scalar restatement, in plain C, of
 *   - the 3D->2D projection of key points,
 *       projects/mmdet3d_plugin/models/utils/detr3d_transformer.py:547-555
 *   - mmcv-full==1.6.2 `ms_deformable_im2col_gpu_kernel` (third-party, not under
 *     /root/reference; call site detr3d_transformer.py:561-563), i.e. bilinear
 *     sampling with zero padding, align_corners=False:
 *       h_im = loc_y*H - 0.5 ; w_im = loc_x*W - 0.5 (one fused multiply-add, as nvcc
 *       contracts it), sample used iff h_im>-1 && w_im>-1 && h_im<H && w_im<W,
 *       h_low=floor(h_im), w_low=floor(w_im), every corner individually zero outside.
 *   - the sum over cameras, detr3d_transformer.py:565-569.
 *
 * fp32 where the reference computes index/mask-deciding values (so the CUDA
 * kernels can be held to bit-exact floor indices and in-bounds masks), float64
 * accumulation for the sampled values (so it can serve as ground truth for the
 * fp32 kernels at rtol 1e-3 and far below).
 *
 * Built by oracle/build.py with `gcc -O2 -ffp-contract=off -shared -fPIC`; the
 * fused multiply-adds the reference arithmetic implies are written explicitly
 * with fmaf().  Pinned against the reference's feature_sampling() run on CPU (tests/test_ref_golden.py; see oracle/__init__.py).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* p = lidar2img(4x4,row-major) @ [x,y,z,1];  u = p0/max(p2,1e-5)/pad_w ; v = p1/max(p2,1e-5)/pad_h
 * detr3d_transformer.py:547-552.  Accumulation order (defined by this oracle, torch.matmul's
 * order on a GPU being unspecified): ((m0*x + m1*y) + m2*z) + m3 with fused multiply-adds. */
static void project_point(const float* m, float x, float y, float z, float pad_h, float pad_w,
                          float* u, float* v) {
    float p[3];
    for (int i = 0; i < 3; ++i) {
        float acc = m[4 * i + 0] * x;
        acc = fmaf(m[4 * i + 1], y, acc);
        acc = fmaf(m[4 * i + 2], z, acc);
        p[i] = acc + m[4 * i + 3];
    }
    float zc = p[2] > 1e-5f ? p[2] : 1e-5f;      /* torch.clamp(min=1e-5); NaN stays NaN */
    if (p[2] != p[2]) zc = p[2];
    *u = (p[0] / zc) / pad_w;
    *v = (p[1] / zc) / pad_h;
}

/* key_points [B,Nq,P,3], lidar2img [B,N,4,4] -> uv [B,N,Nq,P,2] (normalised x,y). */
void far3d_oracle_project(const float* key_points, const float* lidar2img, float pad_h, float pad_w,
                          int B, int N, int Nq, int P, float* uv) {
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n)
            for (int q = 0; q < Nq; ++q)
                for (int p = 0; p < P; ++p) {
                    const float* kp = key_points + (((size_t)b * Nq + q) * P + p) * 3;
                    float* o = uv + ((((size_t)b * N + n) * Nq + q) * P + p) * 2;
                    project_point(lidar2img + ((size_t)b * N + n) * 16, kp[0], kp[1], kp[2], pad_h, pad_w,
                                  o, o + 1);
                }
}

/* One bilinear sample of D channels, double accumulation into acc[D] scaled by aw.
 * value_l points at value[b, start_l, g, 0]; row stride (per pixel) = G*D floats.
 * Returns 1 if the sample passed the in-bounds test. h_low/w_low always written. */
static int sample_bilinear(const float* value_l, int H, int W, int pix_stride, int D, float loc_x,
                           float loc_y, double aw, double* acc, int32_t* h_low_o, int32_t* w_low_o) {
    float w_im = fmaf(loc_x, (float)W, -0.5f);
    float h_im = fmaf(loc_y, (float)H, -0.5f);
    /* clamp before the int conversion only to keep the (unused) index defined for huge/NaN coords */
    float hf = floorf(h_im), wf = floorf(w_im);
    int h_low = (hf >= -2147483000.f && hf <= 2147483000.f) ? (int)hf : 0;
    int w_low = (wf >= -2147483000.f && wf <= 2147483000.f) ? (int)wf : 0;
    *h_low_o = h_low; *w_low_o = w_low;
    if (!(h_im > -1 && w_im > -1 && h_im < (float)H && w_im < (float)W)) return 0;
    double lh = (double)h_im - h_low, lw = (double)w_im - w_low;
    double hh = 1 - lh, hw = 1 - lw;
    int ys[4] = {h_low, h_low, h_low + 1, h_low + 1};
    int xs[4] = {w_low, w_low + 1, w_low, w_low + 1};
    double cw[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
    for (int c = 0; c < 4; ++c) {
        if (ys[c] < 0 || ys[c] >= H || xs[c] < 0 || xs[c] >= W) continue;
        const float* px = value_l + ((size_t)ys[c] * W + xs[c]) * pix_stride;
        for (int d = 0; d < D; ++d) acc[d] += aw * cw[c] * (double)px[d];
    }
    return 1;
}

/* mmcv layout: value [BN, S, G, D]; shapes [L,2] (H,W); start [L]; loc [BN,Nq,G,L,P,2] (x,y);
 * w [BN,Nq,G,L*P] -> out [BN,Nq,G*D] (fp32, rounded once from double).
 * Optional debug outputs: idx [BN,Nq,G,L,P,2] int32 (h_low,w_low), valid [BN,Nq,G,L,P] uint8. */
void far3d_oracle_msda(const float* value, const int64_t* shapes, const int64_t* start, const float* loc,
                       const float* w, int BN, int S, int G, int D, int Nq, int L, int P, float* out,
                       int32_t* idx, uint8_t* valid) {
    double acc[1024];
    for (int b = 0; b < BN; ++b)
        for (int q = 0; q < Nq; ++q)
            for (int g = 0; g < G; ++g) {
                for (int d = 0; d < D; ++d) acc[d] = 0;
                for (int l = 0; l < L; ++l) {
                    int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
                    const float* vl = value + (((size_t)b * S + start[l]) * G + g) * D;
                    for (int p = 0; p < P; ++p) {
                        size_t si = ((((size_t)b * Nq + q) * G + g) * L + l) * P + p;
                        int32_t hl, wl;
                        int ok = sample_bilinear(vl, H, W, G * D, D, loc[2 * si], loc[2 * si + 1],
                                                 (double)w[si], acc, &hl, &wl);
                        if (idx) { idx[2 * si] = hl; idx[2 * si + 1] = wl; }
                        if (valid) valid[si] = (uint8_t)ok;
                    }
                }
                float* o = out + (((size_t)b * Nq + q) * G + g) * D;
                for (int d = 0; d < D; ++d) o[d] = (float)acc[d];
            }
}

/* Fused op (detr3d_transformer.py:544-569): feat [B*N, S, C] channels-last (C = G*D),
 * key_points [B,Nq,P,3], lidar2img [B,N,4,4], weights [B*N,Nq,G,L*P] (post-softmax, the layout
 * `_get_weights` returns, :541-542) -> out [B,Nq,C] = sum over cameras.
 * Optional debug outputs: uv [B,N,Nq,P,2], idx [B,N,Nq,L,P,2] int32, valid [B,N,Nq,L,P] uint8. */
void far3d_oracle_deform_agg(const float* feat, const int64_t* shapes, const int64_t* start,
                             const float* key_points, const float* lidar2img, const float* weights,
                             float pad_h, float pad_w, int B, int N, int S, int G, int D, int Nq, int L,
                             int P, float* out, float* uv_o, int32_t* idx, uint8_t* valid) {
    double acc[1024];
    for (int b = 0; b < B; ++b)
        for (int q = 0; q < Nq; ++q)
            for (int g = 0; g < G; ++g) {
                for (int d = 0; d < D; ++d) acc[d] = 0;
                for (int n = 0; n < N; ++n) {
                    int bn = b * N + n;
                    for (int p = 0; p < P; ++p) {
                        const float* kp = key_points + (((size_t)b * Nq + q) * P + p) * 3;
                        float u, v;
                        project_point(lidar2img + (size_t)bn * 16, kp[0], kp[1], kp[2], pad_h, pad_w, &u, &v);
                        if (uv_o && g == 0) {
                            float* o = uv_o + ((((size_t)b * N + n) * Nq + q) * P + p) * 2;
                            o[0] = u; o[1] = v;
                        }
                        for (int l = 0; l < L; ++l) {
                            int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
                            const float* vl = feat + ((size_t)bn * S + start[l]) * (G * D) + (size_t)g * D;
                            double aw = (double)weights[(((size_t)bn * Nq + q) * G + g) * (L * P) + l * P + p];
                            int32_t hl, wl;
                            int ok = sample_bilinear(vl, H, W, G * D, D, u, v, aw, acc, &hl, &wl);
                            if (g == 0) {
                                size_t si = ((((size_t)bn) * Nq + q) * L + l) * P + p;
                                if (idx) { idx[2 * si] = hl; idx[2 * si + 1] = wl; }
                                if (valid) valid[si] = (uint8_t)ok;
                            }
                        }
                    }
                }
                float* o = out + ((size_t)b * Nq + q) * (G * D) + (size_t)g * D;
                for (int d = 0; d < D; ++d) o[d] = (float)acc[d];
            }
}
