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
