"""Oracle for mmcv's MultiScaleDeformableAttnFunction.forward (TEST INFRASTRUCTURE).

The CUDA kernel itself (`ms_deformable_im2col_gpu_kernel`, mmcv-full==1.6.2,
py38.yaml:145) is third-party and absent from /root/reference.  Its published
algorithm is restated twice, independently:

* `msda_grid_sample`  - the `F.grid_sample(bilinear, zeros, align_corners=False)`
  form, which is exactly how the reference itself restates the sampling in
  `projects/mmdet3d_plugin/models/utils/sparse_blocks.py:234-255` (and how mmcv's
  own `multi_scale_deformable_attn_pytorch` CPU fallback does it).
* `msda_scalar`       - explicit per-sample loops (numpy, float64 or float32)
  following the im2col rule: `h_im = loc_y*H - 0.5`, `w_im = loc_x*W - 0.5`,
  sample used iff `h_im > -1 and w_im > -1 and h_im < H and w_im < W`, corners
  `floor`, each corner individually zero outside the map.  It also returns the
  floor indices and in-bounds masks so tests can demand bit-exact index/mask
  parity from the CUDA kernels.

Call site being replaced: `models/utils/detr3d_transformer.py:561-563`.
"""
import numpy as np
import torch
import torch.nn.functional as F


def fma32(a, b, c):
    """fp32 fused multiply-add emulated through float64 (the product of two fp32
    values is exact in fp64; one final rounding to fp32)."""
    return np.float32(np.float64(np.float32(a)) * np.float64(np.float32(b)) + np.float64(np.float32(c)))


def msda_grid_sample(value, spatial_shapes, level_start_index, sampling_locations,
                     attention_weights):
    """value (B, sumHW, G, D); spatial_shapes (L,2)[H,W]; sampling_locations
    (B, Nq, G, L, P, 2) normalised [x, y]; attention_weights (B, Nq, G, L*P).
    Returns (B, Nq, G*D).  Mirrors sparse_blocks.py:234-255 / mmcv's pytorch fallback."""
    B, _, G, D = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    value_list = value.split([h * w for h, w in shapes], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, (H, W) in enumerate(shapes):
        v = value_list[lvl].flatten(2).transpose(1, 2).reshape(B * G, D, H, W)
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)  # (B*G, Nq, P, 2)
        sampled.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros',
                                     align_corners=False))      # (B*G, D, Nq, P)
    w = attention_weights.reshape(B, Nq, G, L, P).permute(0, 2, 1, 3, 4).reshape(B * G, 1, Nq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * w).sum(-1).view(B, G * D, Nq)
    return out.transpose(1, 2).contiguous()


def msda_scalar(value, spatial_shapes, level_start_index, sampling_locations,
                attention_weights, dtype=np.float64):
    """Scalar-loop restatement of ms_deformable_im2col.  Small inputs only.
    Returns (out (B,Nq,G*D), floor_idx (B,Nq,G,L,P,2) int64 [h_low,w_low],
    valid (B,Nq,G,L,P) bool)."""
    value = np.asarray(value, dtype=np.float32)
    loc = np.asarray(sampling_locations, dtype=np.float32)
    wts = np.asarray(attention_weights, dtype=np.float32)
    shapes = [(int(h), int(w)) for h, w in np.asarray(spatial_shapes).tolist()]
    starts = [int(s) for s in np.asarray(level_start_index).tolist()]
    B, _, G, D = value.shape
    _, Nq, _, L, P, _ = loc.shape
    out = np.zeros((B, Nq, G, D), dtype=dtype)
    fidx = np.zeros((B, Nq, G, L, P, 2), dtype=np.int64)
    valid = np.zeros((B, Nq, G, L, P), dtype=bool)
    for b in range(B):
        for q in range(Nq):
            for g in range(G):
                acc = np.zeros(D, dtype=dtype)
                for l, (H, W) in enumerate(shapes):
                    for p in range(P):
                        # fp32 arithmetic exactly as the compiled kernel: nvcc contracts
                        # `loc * W - 0.5` into one fused multiply-add (default --fmad=true)
                        w_im = fma32(loc[b, q, g, l, p, 0], W, -0.5)
                        h_im = fma32(loc[b, q, g, l, p, 1], H, -0.5)
                        h_low = int(np.floor(h_im)) if np.isfinite(h_im) else 0
                        w_low = int(np.floor(w_im)) if np.isfinite(w_im) else 0
                        ok = bool(h_im > -1 and w_im > -1 and h_im < H and w_im < W)
                        fidx[b, q, g, l, p] = (h_low, w_low)
                        valid[b, q, g, l, p] = ok
                        if not ok:
                            continue
                        lh = dtype(h_im) - h_low
                        lw = dtype(w_im) - w_low
                        hh, hw = 1 - lh, 1 - lw
                        aw = dtype(wts[b, q, g, l * P + p])
                        base = starts[l]
                        for (yy, xx, cw) in ((h_low, w_low, hh * hw), (h_low, w_low + 1, hh * lw),
                                             (h_low + 1, w_low, lh * hw), (h_low + 1, w_low + 1, lh * lw)):
                            if 0 <= yy < H and 0 <= xx < W:
                                acc += aw * cw * value[b, base + yy * W + xx, g].astype(dtype)
                out[b, q, g] = acc
    return out.reshape(B, Nq, G * D), fidx, valid
