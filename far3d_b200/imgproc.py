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
        """post-homography matrix of _img_transform (custom_pipeline.py:293-311), float32 torch arithmetic as there"""
        ida_rot = torch.eye(2) * resize
        ida_tran = torch.zeros(2) - torch.Tensor(crop[:2])
        if flip:
            A = torch.Tensor([[-1, 0], [0, 1]])
            b = torch.Tensor([crop[2] - crop[0], 0])
            ida_rot = A.matmul(ida_rot)
            ida_tran = A.matmul(ida_tran) + b
        h = rotate / 180 * np.pi
        A = torch.Tensor([[np.cos(h), np.sin(h)], [-np.sin(h), np.cos(h)]])
        b = torch.Tensor([crop[2] - crop[0], crop[3] - crop[1]]) / 2
        b = A.matmul(-b) + b
        ida_rot = A.matmul(ida_rot)
        ida_tran = A.matmul(ida_tran) + b
        ida_mat = torch.eye(3)
        ida_mat[:2, :2] = ida_rot
        ida_mat[:2, 2] = ida_tran
        return ida_mat

    def __call__(self, results):
        if 'depthmap' in results or len(results.get('gt_bboxes', [])) > 0:
            raise NotImplementedError('ground-truth boxes / depth maps are training keys: use the reference transform')
        assert self.data_aug_conf['rot_lim'] == (0.0, 0.0), 'Rotation is not currently supported'
        imgs = results['img']
        N = len(imgs) // 2 if self.multi_stamps else len(imgs)
        new_imgs, ida_mats = [], []
        for i in range(N):
            H, W = imgs[i].shape[:2]
            if H > W:                                            # portrait view (AV2 front centre): to landscape first
                resize, resize_dims, crop = self._sample_augmentation_f(imgs[i].shape)
                img = resize_crop_u8(imgs[i], resize_dims, crop)
                ida_mat_f = self._ida_mat(resize, crop)
                resize, resize_dims, crop, flip, rotate = self._sample_augmentation(img.shape)
                img = resize_crop_u8(img, resize_dims, crop, flip)
                ida_mat = self._ida_mat(resize, crop, flip, rotate) @ ida_mat_f
            else:
                resize, resize_dims, crop, flip, rotate = self._sample_augmentation(imgs[i].shape)
                img = resize_crop_u8(imgs[i], resize_dims, crop, flip)
                ida_mat = self._ida_mat(resize, crop, flip, rotate)
            new_imgs.append(img)
            results['intrinsics'][i][:3, :3] = ida_mat @ results['intrinsics'][i][:3, :3]
            ida_mats.append(np.array(ida_mat))
        results['img'] = new_imgs
        results['cam2img'] = results['intrinsics']
        results['lidar2img'] = [results['intrinsics'][i] @ results['extrinsics'][i] for i in range(len(results['extrinsics']))]
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
