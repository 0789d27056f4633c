"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the image side of the reference's test pipeline that
`far3d_normalize_u8` replaces on the device.

  NormalizeMultiviewImage      projects/mmdet3d_plugin/datasets/pipelines/transform_3d.py:74-101  (config far3d.py:13-14,192)
  AV2PadMultiViewImage         projects/mmdet3d_plugin/datasets/pipelines/custom_pipeline.py:340-385  ('same2max', pad_val 0)
  HWC -> CHW stack             mmdet3d DefaultFormatBundle3D (third party), called from formating.py:36-60

`mmcv.imnormalize` is third-party (mmcv-full 1.6.2, mmcv/image/photometric.py: imnormalize_): img.astype(float32);
optional cv2 BGR->RGB; cv2.subtract(img, float64(mean)); cv2.multiply(img, 1 / float64(std)) - on a float32 image OpenCV
evaluates both in float32, which is what is restated here.  Parity unpinned against mmcv itself (not installable here): the
expression is one subtraction and one multiplication per sample; the GPU test bar is 1 ulp-level (1e-6 relative)."""
import numpy as np


def normalize_pad_u8(imgs_u8, mean, std, to_rgb=False, pad_hw=None):
    """imgs_u8: list of (H_i, W_i, 3) uint8 arrays or one (N, H, W, 3) array -> (N, 3, Hp, Wp) float32."""
    imgs = [np.asarray(i) for i in imgs_u8]
    mean = np.asarray(mean, dtype=np.float32)
    stdinv = (1.0 / np.asarray(std, dtype=np.float32).astype(np.float64)).astype(np.float32)
    out = []
    for img in imgs:
        x = img.astype(np.float32)
        if to_rgb:
            x = x[..., ::-1]
        out.append(((x - mean[None, None, :]) * stdinv[None, None, :]).astype(np.float32))
    if pad_hw is None:                                     # 'same2max': pad every view to the largest one
        pad_hw = (max(o.shape[0] for o in out), max(o.shape[1] for o in out))
    padded = np.zeros((len(out), pad_hw[0], pad_hw[1], 3), dtype=np.float32)
    for k, o in enumerate(out):
        padded[k, :o.shape[0], :o.shape[1]] = o
    return np.ascontiguousarray(padded.transpose(0, 3, 1, 2))


# ------------------------------------------------------------------------------------------------ resize + crop (+ flip)
# ResizeCropFlipRotImage._img_transform          custom_pipeline.py:277-311: img.resize(resize_dims); img.crop(crop); flip
# AV2ResizeCropFlipRotImageV2.__call__           custom_pipeline.py:48-149 (portrait views go through the transform twice)
#
# `PIL.Image.resize` is third-party (Pillow; 12.2.0 in this image, the algorithm is unchanged since 3.x): default filter for
# an RGB image = BICUBIC, implemented in src/libImaging/Resample.c as two separable passes on 8-bit data -
#   precompute_coeffs:   scale = in / out; filterscale = max(scale, 1); support = 2 * filterscale; per output index xx the taps
#                        xmin = int(center - support + 0.5) .. xmax = int(center + support + 0.5) (clamped to the image) with
#                        weights bicubic((x + xmin - center + 0.5) / filterscale), a = -0.5, normalised by their sum (double)
#   normalize_coeffs_8bpc: k_int = int(+-0.5 + k * 2^22)   (PRECISION_BITS = 32 - 8 - 2)
#   ImagingResampleHorizontal_8bpc, then ImagingResampleVertical_8bpc on its 8-bit result:
#                        out = clip8((2^21 + sum_x pixel[x + xmin] * k_int[x]) >> 22), 32-bit int accumulation
# `Image.crop` copies the box and fills what lies outside the image with zeros.  Pinned against Pillow itself:
# tests/test_cpu.py::test_resize_crop_oracle_is_pillow (bit-exact on random images, up- and down-scaling, out-of-image crops).
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    x = np.abs(x)
    a = -0.5
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def resample_coeffs(in_size, out_size):
    """(bounds int32 [out, 2] = (xmin, count), k int32 [out, ksize]) of Pillow's bicubic resampling of in_size -> out_size."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = 0.0 + (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    k = np.zeros((out_size, ksize), dtype=np.float64)
    ww = np.zeros(out_size, dtype=np.float64)
    for x in range(ksize):                                   # sequential accumulation of ww, as the C loop does
        live = x < xmax
        w = _bicubic(((x + xmin) - center + 0.5) * ss)
        w = np.where(live, w, 0.0)
        k[:, x] = w
        ww = ww + w
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz, None]
    ki = np.where(k < 0, np.trunc(-0.5 + k * (1 << PRECISION_BITS)), np.trunc(0.5 + k * (1 << PRECISION_BITS))).astype(np.int32)
    return np.stack([xmin, xmax], axis=1).astype(np.int32), ki


def _clip8(ss):
    return np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_u8(img, new_w, new_h):
    """img (H, W, C) uint8 -> (new_h, new_w, C) uint8: Image.resize((new_w, new_h)) with the default (bicubic) filter."""
    img = np.asarray(img, dtype=np.uint8)
    H, W, C = img.shape
    cur = img
    if new_w != W:                                           # horizontal pass
        b, k = resample_coeffs(W, new_w)
        acc = np.full((H, new_w, C), 1 << (PRECISION_BITS - 1), dtype=np.int32)
        for x in range(k.shape[1]):
            idx = np.minimum(b[:, 0] + x, W - 1)             # taps beyond the count carry zero weight
            acc += cur[:, idx, :].astype(np.int32) * k[None, :, x, None]
        cur = _clip8(acc)
    if new_h != H:                                           # vertical pass on the 8-bit result
        b, k = resample_coeffs(H, new_h)
        acc = np.full((new_h, cur.shape[1], C), 1 << (PRECISION_BITS - 1), dtype=np.int32)
        for x in range(k.shape[1]):
            idx = np.minimum(b[:, 0] + x, H - 1)
            acc += cur[idx, :, :].astype(np.int32) * k[:, None, x, None]
        cur = _clip8(acc)
    return cur


def resize_crop_flip_u8(img, resize_dims, crop, flip=False):
    """_img_transform of the reference on one uint8 view (rotation is asserted 0 by the AV2 pipeline, custom_pipeline.py:68)."""
    new_w, new_h = resize_dims
    r = pil_resize_u8(img, new_w, new_h)
    x0, y0, x1, y1 = crop
    out = np.zeros((y1 - y0, x1 - x0, r.shape[2]), dtype=np.uint8)
    sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, new_w), min(y1, new_h)
    if sx1 > sx0 and sy1 > sy0:
        out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = r[sy0:sy1, sx0:sx1]
    return out[:, ::-1].copy() if flip else out
