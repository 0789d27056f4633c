"""Image side of the reference's AV2 test pipeline on the device.

    LoadMultiViewImageFromFiles -> AV2ResizeCropFlipRotImageV2 -> NormalizeMultiviewImage -> AV2PadMultiViewImage
    (projects/configs/far3d.py:188-194)

`AV2ResizeCropFlipRotImageV2` below mirrors the reference class of the same name (custom_pipeline.py:48-149, helpers
:277-336) for the keys an inference pipeline carries (images, intrinsics / extrinsics -> lidar2img, ida_mat): same
augmentation sampling, same 3x3 post-homography matrices; the pixels come from `resize_crop_u8` = far3d_resize_crop_u8, which
is bit-exact with the `PIL.Image.resize / crop / transpose` calls of `_img_transform`.  Camera frames therefore go to the GPU
as the sensors' native uint8 pixels (2048 x 1550), the 960 x 640 crop is produced there and `ops.normalize_u8` finishes the
job.  Ground-truth boxes / depth maps (training keys) are not handled here.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream

_TABLES = {}


def resample_tables(in_size, out_size, device):
    """Pillow's bicubic tap tables of one axis (host copy of the bounds + both tables on the device), cached per size pair."""
    key = (int(in_size), int(out_size), str(device))
    t = _TABLES.get(key)
    if t is None:
        lib = _lib.load()
        ksize = int(lib.far3d_resample_ksize(int(in_size), int(out_size)))
        bounds = np.empty((out_size, 2), dtype=np.int32)
        k = np.empty((out_size, ksize), dtype=np.int32)
        _lib.call('far3d_resample_coeffs', int(in_size), int(out_size), bounds.ctypes.data_as(ctypes.c_void_p),
                  k.ctypes.data_as(ctypes.c_void_p))
        t = _TABLES[key] = (bounds, torch.from_numpy(bounds).to(device), torch.from_numpy(k).to(device), ksize)
    return t


def prefetch_tables(shape_hw, resize_lim, device):
    """Build (and upload) the tap tables of every resized size a view of this shape can draw from `resize_lim` - the reference's
    test pipeline draws a fresh resize factor per view and frame (custom_pipeline.py:313-317), i.e. ~290 size pairs for a
    2048 x 1550 camera at (0.47, 0.55), 12 MB on the device - so that a serving loop never builds one inside a frame."""
    H, W = int(shape_hw[0]), int(shape_hw[1])
    lo, hi = resize_lim
    for n_in in (W, H):
        for n_out in range(int(n_in * lo), int(n_in * hi) + 1):
            resample_tables(n_in, n_out, device)


def resize_crop_u8(src, resize_dims, crop, flip=False, out=None):
    """src uint8 CUDA [H, W, 3] -> uint8 [crop_h, crop_w, 3]: PIL `img.resize(resize_dims).crop(crop)` (+ FLIP_LEFT_RIGHT).
    `out`: destination view (last two dims contiguous, e.g. one camera of a [N, H, W, 3] batch)."""
    if not (src.is_cuda and src.dtype == torch.uint8 and src.dim() == 3 and src.shape[2] == 3 and src.is_contiguous()):
        raise _lib.Far3DNativeError('resize_crop_u8: contiguous uint8 CUDA [H, W, 3] image required')
    H, W = int(src.shape[0]), int(src.shape[1])
    new_w, new_h = int(resize_dims[0]), int(resize_dims[1])
    x0, y0, x1, y1 = (int(v) for v in crop)
    out_w, out_h = x1 - x0, y1 - y0
    if out is None:
        out = torch.empty(out_h, out_w, 3, device=src.device, dtype=torch.uint8)
    if not (out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (out_h, out_w, 3) and out.stride(2) == 1
            and out.stride(1) == 3 and out.stride(0) % 3 == 0):
        raise _lib.Far3DNativeError('resize_crop_u8: bad destination')
    _, xb, xk, xks = resample_tables(W, new_w, src.device)
    yb_host, yb, yk, yks = resample_tables(H, new_h, src.device)
    ya, yz = max(y0, 0), min(y1, new_h)                       # window rows inside the resized image
    if yz > ya:
        y_first = int(yb_host[ya:yz, 0].min())
        rows = int((yb_host[ya:yz, 0] + yb_host[ya:yz, 1]).max()) - y_first
    else:
        y_first, rows = 0, 0
    tmp = torch.empty(max(rows, 1) * out_w * 3, device=src.device, dtype=torch.uint8)
    _lib.call('far3d_resize_crop_u8', _ptr(src), H, W, new_w, new_h, _ptr(xb), _ptr(xk), xks, _ptr(yb), _ptr(yk), yks, y_first, rows,
              x0, y0, out_w, out_h, int(bool(flip)), _ptr(tmp), _ptr(out), out.stride(0) // 3, _stream())
    return out


class AV2ResizeCropFlipRotImageV2:
    """Device form of the reference transform (custom_pipeline.py:48-149).  results['img']: list of uint8 CUDA [H, W, 3] views
    (the reference holds numpy arrays); output images are uint8 CUDA [fH, fW, 3] (the reference casts to float32 after the
    Pillow calls - `ops.normalize_u8` does that cast)."""

    def __init__(self, data_aug_conf=None, multi_stamps=False):
        self.data_aug_conf = data_aug_conf
        self.min_size = 2.0
        self.multi_stamps = multi_stamps

    # ---- augmentation parameters: custom_pipeline.py:313-336 (np.random is consumed exactly as there)
    def _sample_augmentation(self, shape):
        H, W = shape[:2]
        fH, fW = self.data_aug_conf['final_dim']
        resize = np.random.uniform(*self.data_aug_conf['resize_lim'])
        resize_dims = (int(W * resize), int(H * resize))
        newW, newH = resize_dims
        crop_h = int((1 - np.random.uniform(*self.data_aug_conf['bot_pct_lim'])) * newH) - fH
        crop_w = int(np.random.uniform(0, max(0, newW - fW)))
        crop = (crop_w, crop_h, crop_w + fW, crop_h + fH)
        flip = False
        if self.data_aug_conf['rand_flip'] and np.random.choice([0, 1]):
            flip = True
        rotate = np.random.uniform(*self.data_aug_conf['rot_lim'])
        return resize, resize_dims, crop, flip, rotate

    @staticmethod
    def _sample_augmentation_f(shape):
        H, W = shape[:2]
        fH, fW = W, H
        resize = np.round(((H + 50) / W), 2)
        resize_dims = (int(W * resize), int(H * resize))
        newW, newH = resize_dims
        crop_h = int((newH - fH) / 2)
        crop_w = int((newW - fW) / 2)
        return resize, resize_dims, (crop_w, crop_h, crop_w + fW, crop_h + fH)

    @staticmethod
    def _ida_mat(resize, crop, flip=False, rotate=0):
        """post-homography matrix of _img_transform (custom_pipeline.py:293-311) as a float32 torch tensor.  The reference builds
        it from float32 2x2 products with the rotation matrix of `rotate` degrees; the AV2 pipeline asserts rotate == 0
        (custom_pipeline.py:68), where every factor is a signed identity and the products are exact, so the closed form below is
        the same float32 matrix: [[+-resize, 0, tx], [0, resize, ty], [0, 0, 1]]."""
        if rotate != 0:
            raise NotImplementedError('rotation is not supported by the AV2 pipeline (custom_pipeline.py:68)')
        r = np.float32(resize)
        m = np.zeros((3, 3), dtype=np.float32)
        m[0, 0] = -r if flip else r
        m[1, 1] = r
        m[2, 2] = 1
        m[0, 2] = np.float32(crop[0]) + np.float32(crop[2] - crop[0]) if flip else -np.float32(crop[0])
        m[1, 2] = -np.float32(crop[1])
        return torch.from_numpy(m)

    def plan(self, shapes, intrinsics, extrinsics):
        """HOST side of __call__ for views of the given (H, W[, 3]) shapes: draws the augmentation parameters (np.random consumed
        as the reference does), returns (steps per view = [(resize_dims, crop, flip), ...], intrinsics', lidar2img, ida_mats)."""
        assert self.data_aug_conf['rot_lim'] == (0.0, 0.0), 'Rotation is not currently supported'
        N = len(shapes) // 2 if self.multi_stamps else len(shapes)
        steps, ida_mats = [], []
        intrinsics = [np.asarray(k) for k in intrinsics]
        for i in range(N):
            H, W = shapes[i][:2]
            if H > W:                                            # portrait view (AV2 front centre): to landscape first
                resize, resize_dims, crop = self._sample_augmentation_f(shapes[i])
                ida_mat_f = self._ida_mat(resize, crop)
                mid = (crop[3] - crop[1], crop[2] - crop[0], 3)
                resize2, resize_dims2, crop2, flip2, rotate2 = self._sample_augmentation(mid)
                steps.append([(resize_dims, crop, False), (resize_dims2, crop2, flip2)])
                ida_mat = self._ida_mat(resize2, crop2, flip2, rotate2) @ ida_mat_f
            else:
                resize, resize_dims, crop, flip, rotate = self._sample_augmentation(shapes[i])
                steps.append([(resize_dims, crop, flip)])
                ida_mat = self._ida_mat(resize, crop, flip, rotate)
            intrinsics[i][:3, :3] = ida_mat.numpy() @ intrinsics[i][:3, :3]      # float32 @ float64, as torch.Tensor @ ndarray resolves
            ida_mats.append(ida_mat.numpy().copy())
        lidar2img = [intrinsics[i] @ np.asarray(extrinsics[i]) for i in range(len(extrinsics))]
        return steps, intrinsics, lidar2img, ida_mats

    @staticmethod
    def apply(views, steps, out=None):
        """DEVICE side: the planned resize / crop / flip steps on uint8 CUDA views; `out[i]` (optional) receives view i."""
        imgs = []
        for i, (v, st) in enumerate(zip(views, steps)):
            img = v
            for j, (dims, crop, flip) in enumerate(st):
                last = j == len(st) - 1
                img = resize_crop_u8(img, dims, crop, flip, out=out[i] if (out is not None and last) else None)
            imgs.append(img)
        return imgs

    def __call__(self, results):
        if 'depthmap' in results or len(results.get('gt_bboxes', [])) > 0:
            raise NotImplementedError('ground-truth boxes / depth maps are training keys: use the reference transform')
        imgs = results['img']
        steps, intrinsics, lidar2img, ida_mats = self.plan([tuple(i.shape) for i in imgs], results['intrinsics'], results['extrinsics'])
        new_imgs = self.apply(imgs[:len(steps)], steps)
        results['img'] = new_imgs
        results['intrinsics'] = intrinsics
        results['cam2img'] = results['intrinsics']
        results['lidar2img'] = lidar2img
        results['img_shape'] = [tuple(img.shape) for img in new_imgs]
        results['pad_shape'] = [tuple(img.shape) for img in new_imgs]
        results['ida_mat'] = ida_mats
        return results


def frame_from_cameras(views, intrinsics, extrinsics, transform, device, non_blocking=True):
    """Raw camera views (list of uint8 [H_i, W_i, 3], host - ideally pinned - or CUDA tensors / arrays) + 4x4 camera matrices ->
    the image-side keys of one frame of the detector's `**data` contract: img uint8 [1, N, fH, fW, 3] on `device` (the
    detector pipeline normalises and pads uint8 input with far3d_normalize_u8) and lidar2img / intrinsics / extrinsics
    [1, N, 4, 4] float32.  `transform`: an AV2ResizeCropFlipRotImageV2 of this module."""
    dev_views = []
    for v in views:
        t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))
        dev_views.append(t.to(device, non_blocking=non_blocking).contiguous())
    res = transform(dict(img=dev_views, intrinsics=[np.array(k, dtype=np.float64) for k in intrinsics],
                         extrinsics=[np.array(e, dtype=np.float64) for e in extrinsics]))
    shapes = {tuple(i.shape) for i in res['img']}
    if len(shapes) != 1:
        raise _lib.Far3DNativeError(f'views of different final sizes {sorted(shapes)}: pad with ops.normalize_u8 per view')
    to4 = lambda ms: torch.from_numpy(np.stack([np.asarray(m, dtype=np.float64) for m in ms])).float().unsqueeze(0).to(device)
    return dict(img=torch.stack(res['img']).unsqueeze(0), lidar2img=to4(res['lidar2img']), intrinsics=to4(res['intrinsics']),
                extrinsics=to4(res['extrinsics'])), res
