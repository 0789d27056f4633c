// Camera-frame resize + crop (+ horizontal flip) on the device, bit-exact with the Pillow calls of the reference's image
// augmentation: ResizeCropFlipRotImage._img_transform (custom_pipeline.py:277-311: img.resize(resize_dims) - Pillow's default
// BICUBIC filter - then img.crop(crop), FLIP_LEFT_RIGHT) as AV2ResizeCropFlipRotImageV2 (custom_pipeline.py:48-149) applies it
// to every view of a frame.  With it a frame crosses PCIe as the cameras' native uint8 pixels and the
// 960 x 640 network input is produced next to far3d_normalize_u8.
//
// Pillow resamples 8-bit images in two separable passes with 22-bit fixed-point coefficients and an 8-bit intermediate image
// (src/libImaging/Resample.c, third-party: restated from its published algorithm in oracle/preprocess.py, which is pinned
// bit-exact against Pillow itself).  The coefficient tables depend only on (input size, output size): the host entry point
// far3d_resample_coeffs builds them in double precision, the two kernels apply them in 32-bit integer arithmetic.  Only the
// cropped window is computed: horizontal pass over the source rows the window's vertical taps touch, vertical pass into the
// destination (a view of the batched [N, fH, fW, 3] uint8 tensor).  HBM-bound byte work: one thread per output pixel, the three
// channel accumulators in registers; a 2048 x 1550 view is ~9.5 MB read once.
#include <math.h>
#include "common.cuh"

namespace far3d {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t rs_clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// tmp[r][xo][c]: source row y_first + r resampled at column crop_x0 + xo of the resized image (zero outside it)
__global__ void __launch_bounds__(256)
resample_h_u8_kernel(const uint8_t* __restrict__ src, int W, int y_first, const int* __restrict__ xb, const int* __restrict__ xk,
                     int ksize, int new_w, int crop_x0, int out_w, uint8_t* __restrict__ tmp) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    if (xo >= out_w) return;
    const int r = blockIdx.y;
    uint8_t* t = tmp + ((size_t)r * out_w + xo) * 3;
    const int xr = crop_x0 + xo;
    if (xr < 0 || xr >= new_w) { t[0] = 0; t[1] = 0; t[2] = 0; return; }
    const int xmin = __ldg(xb + 2 * xr), n = __ldg(xb + 2 * xr + 1);
    const uint8_t* s = src + ((size_t)(y_first + r) * W + xmin) * 3;
    const int* k = xk + (size_t)xr * ksize;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int i = 0; i < n; ++i) {
        const int w = __ldg(k + i);
        s0 += (int)s[3 * i] * w; s1 += (int)s[3 * i + 1] * w; s2 += (int)s[3 * i + 2] * w;
    }
    t[0] = rs_clip8(s0); t[1] = rs_clip8(s1); t[2] = rs_clip8(s2);
}

// dst[yo][xo' ][c]: rows of tmp combined with the vertical taps of resized row crop_y0 + yo (zero outside the resized image)
__global__ void __launch_bounds__(256)
resample_v_u8_kernel(const uint8_t* __restrict__ tmp, int y_first, const int* __restrict__ yb, const int* __restrict__ yk, int ksize,
                     int new_h, int crop_y0, int out_w, int flip, uint8_t* __restrict__ dst, int dst_row_pixels) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    if (xo >= out_w) return;
    const int yo = blockIdx.y;
    uint8_t* d = dst + ((size_t)yo * dst_row_pixels + (flip ? out_w - 1 - xo : xo)) * 3;
    const int yr = crop_y0 + yo;
    if (yr < 0 || yr >= new_h) { d[0] = 0; d[1] = 0; d[2] = 0; return; }
    const int ymin = __ldg(yb + 2 * yr), n = __ldg(yb + 2 * yr + 1);
    const uint8_t* s = tmp + ((size_t)(ymin - y_first) * out_w + xo) * 3;
    const int* k = yk + (size_t)yr * ksize;
    const size_t pitch = (size_t)out_w * 3;
    int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int i = 0; i < n; ++i) {
        const int w = __ldg(k + i);
        s0 += (int)s[i * pitch] * w; s1 += (int)s[i * pitch + 1] * w; s2 += (int)s[i * pitch + 2] * w;
    }
    d[0] = rs_clip8(s0); d[1] = rs_clip8(s1); d[2] = rs_clip8(s2);
}

static double rs_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

}  // namespace far3d

using namespace far3d;

extern "C" int far3d_resample_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return 0;
    double fs = (double)in_size / out_size;
    if (fs < 1.0) fs = 1.0;
    return (int)ceil(2.0 * fs) * 2 + 1;
}

extern "C" int far3d_resample_coeffs(int in_size, int out_size, int* bounds_host, int* k_host) {
    FAR3D_REQUIRE(in_size > 0 && out_size > 0 && bounds_host && k_host, "bad argument");
    const double scale = (double)in_size / out_size;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * filterscale, ss = 1.0 / filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    double* kd = new double[ksize];
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            const double w = rs_bicubic((x + xmin - center + 0.5) * ss);
            kd[x] = w;
            ww += w;
        }
        int* k = k_host + (size_t)xx * ksize;
        for (int x = 0; x < ksize; ++x) {
            double v = x < xmax ? kd[x] : 0.0;
            if (x < xmax && ww != 0.0) v /= ww;
            k[x] = v < 0 ? (int)(-0.5 + v * (1 << RS_PRECISION_BITS)) : (int)(0.5 + v * (1 << RS_PRECISION_BITS));
        }
        bounds_host[2 * xx] = xmin;
        bounds_host[2 * xx + 1] = xmax;
    }
    delete[] kd;
    return FAR3D_OK;
}

extern "C" int far3d_resize_crop_u8(const uint8_t* src_hwc, int H, int W, int new_w, int new_h, const int* xbounds, const int* xk,
                                    int xksize, const int* ybounds, const int* yk, int yksize, int y_first, int rows, int crop_x0,
                                    int crop_y0, int out_w, int out_h, int flip, uint8_t* tmp, uint8_t* dst_hwc, int dst_row_pixels,
                                    void* stream) {
    FAR3D_REQUIRE(src_hwc && xbounds && xk && ybounds && yk && dst_hwc, "null pointer");
    FAR3D_REQUIRE(H > 0 && W > 0 && new_w > 0 && new_h > 0 && out_w > 0 && out_h > 0 && dst_row_pixels >= out_w, "bad size");
    FAR3D_REQUIRE(xksize == far3d_resample_ksize(W, new_w) && yksize == far3d_resample_ksize(H, new_h), "coefficient tables of another size pair");
    FAR3D_REQUIRE(rows >= 0 && y_first >= 0 && y_first + rows <= H && (rows == 0 || tmp), "source row window outside the image");
    FAR3D_REQUIRE(out_h <= 65535 && rows <= 65535, "window too tall");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 block(256);
    if (rows > 0) {
        resample_h_u8_kernel<<<dim3(cdiv(out_w, 256), rows), block, 0, st>>>(src_hwc, W, y_first, xbounds, xk, xksize, new_w, crop_x0,
                                                                            out_w, tmp);
        int rc = launched("resample_h_u8_kernel");
        if (rc) return rc;
    }
    // rows == 0: the crop window lies outside the resized image, the vertical kernel writes zeros without touching tmp
    resample_v_u8_kernel<<<dim3(cdiv(out_w, 256), out_h), block, 0, st>>>(tmp, y_first, ybounds, yk, yksize, new_h, crop_y0, out_w, flip,
                                                                         dst_hwc, dst_row_pixels);
    return launched("resample_v_u8_kernel");
}
