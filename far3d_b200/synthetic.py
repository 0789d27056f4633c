"""Synthetic frames and weights of the shapes BASELINE.json names: there is no dataset and no
checkpoint offline.  Geometry is realistic enough that 3D key points land inside 1-2 of the ring cameras."""
import math

import numpy as np
import torch

CONFIGS = {
    # name: (num_cams, H, W)
    'cfg1': (1, 256, 256),
    'cfg2': (7, 640, 960),
    'cfg4': (6, 640, 1600),
    'cfg5': (7, 1024, 1536),
    'tiny': (2, 128, 192),
}


def camera_ring(num_cams, H, W, rng):
    """intrinsics (N,4,4) 'viewpad' (argoverse2_dataset_t.py:207-211) and lidar->camera extrinsics (N,4,4): cameras on a
    ring, yaw 360*i/N, 1.5 m off the origin, z forward / x right / y down."""
    intr = np.tile(np.eye(4, dtype=np.float64), (num_cams, 1, 1))
    extr = np.tile(np.eye(4, dtype=np.float64), (num_cams, 1, 1))
    for i in range(num_cams):
        f = 700.0 * (W / 960.0) * rng.uniform(0.95, 1.05)
        intr[i, 0, 0] = intr[i, 1, 1] = f
        intr[i, 0, 2], intr[i, 1, 2] = W / 2.0, H / 2.0
        yaw = 2 * math.pi * i / num_cams
        fwd = np.array([math.cos(yaw), math.sin(yaw), 0.0])
        right = np.array([math.sin(yaw), -math.cos(yaw), 0.0])
        down = np.array([0.0, 0.0, -1.0])
        R = np.stack([right, down, fwd])
        pos = np.array([1.5 * math.cos(yaw), 1.5 * math.sin(yaw), 1.5])
        extr[i, :3, :3] = R
        extr[i, :3, 3] = -R @ pos
    return intr, extr


def make_frame(config='cfg2', frame_idx=0, seed=0, device='cpu', scene='s0'):
    """One frame of the `**data` contract at Far3D.forward(return_loss=False) after forward_test unwrapping
    (SURVEY.md App. C).  Returns (img_metas, data)."""
    N, H, W = CONFIGS[config] if isinstance(config, str) else config
    rng = np.random.RandomState(seed)
    intr, extr = camera_ring(N, H, W, rng)
    g = torch.Generator().manual_seed(seed * 1000 + frame_idx)
    img = torch.randn(1, N, 3, H, W, generator=g)
    yaw = math.radians(rng.uniform(-1, 1)) * frame_idx
    pose = np.eye(4)
    pose[:2, :2] = [[math.cos(yaw), -math.sin(yaw)], [math.sin(yaw), math.cos(yaw)]]
    pose[0, 3] = 1.0 * frame_idx
    data = dict(
        img=img,
        lidar2img=torch.from_numpy(intr @ extr).float().unsqueeze(0),
        intrinsics=torch.from_numpy(intr).float().unsqueeze(0),
        extrinsics=torch.from_numpy(extr).float().unsqueeze(0),
        timestamp=torch.tensor([float(frame_idx)], dtype=torch.float64),
        img_timestamp=torch.zeros(1, N, dtype=torch.float64),
        ego_pose=torch.from_numpy(pose).float().unsqueeze(0),
        ego_pose_inv=torch.from_numpy(np.linalg.inv(pose)).float().unsqueeze(0),
    )
    data = {k: v.to(device) for k, v in data.items()}
    img_metas = [dict(pad_shape=[(H, W, 3)] * N, scene_token=scene, box_type_3d=None)]
    return img_metas, data


@torch.no_grad()
def cold_2d_head_(model, scale=0.01, reg_scale=None):
    """Put the 2D proposal head at the reference's initial operating point (yolox_head.py:232-236: cls / obj biases at
    logit(0.01)) with near-zero predictor weights, so obj*cls ~ 1e-4 << threshold_score and the number of adaptive
    queries is 0 - the state SURVEY.md section 8 predicts for random-init weights - instead of thousands of random peaks.
    `scale=0.05` ("tepid") lets ~150 peaks through at cfg-2, which exercises the adaptive-query path at full size."""
    h = getattr(model, 'img_roi_head', None)
    if h is None:
        return model
    b = float(-math.log((1 - 0.01) / 0.01))
    for c, o in zip(h.multi_level_conv_cls, h.multi_level_conv_obj):
        c.weight.mul_(scale); o.weight.mul_(scale)
        c.bias.fill_(b); o.bias.fill_(b)
    if reg_scale is not None:                  # keep exp(box size logits) finite: random regressors overflow to inf -> NaN centres
        for r in h.multi_level_conv_reg:
            r.weight.mul_(reg_scale); r.bias.mul_(reg_scale)
    if hasattr(h, 'invalidate'):
        h.invalidate()
    return model


@torch.no_grad()
def randomize_(model, seed=0):
    """Deterministic non-degenerate weights for any module tree with the reference's parameter names: variance-preserving
    conv / linear weights, non-trivial BatchNorm statistics, non-zero `weights_fc` (the reference zero-initialises it,
    detr3d_transformer.py:518, which would hide weight-indexing bugs behind a uniform softmax)."""
    g = torch.Generator().manual_seed(seed)

    def normal(t, std):
        t.copy_(torch.randn(t.shape, generator=g) * std)

    def uniform(t, a, b):
        t.copy_(torch.rand(t.shape, generator=g) * (b - a) + a)

    seen = set()
    for name, p in list(model.named_parameters()) + list(model.named_buffers()):
        if id(p) in seen or not p.dtype.is_floating_point:
            continue
        seen.add(id(p))
        leaf = name.split('.')[-1]
        if name.endswith(('code_weights', 'match_costs', 'pc_range')) or leaf == 'num_batches_tracked':
            continue
        if 'reference_points' in name:
            uniform(p, 0.05, 0.95)
        elif leaf == 'running_mean':
            normal(p, 0.1)
        elif leaf == 'running_var':
            uniform(p, 0.5, 1.5)
        elif p.dim() == 4:                                   # conv weight
            fan_in = p.shape[1] * p.shape[2] * p.shape[3]
            normal(p, math.sqrt(2.0 / fan_in))
        elif p.dim() == 2:                                   # linear / in_proj weight
            normal(p, math.sqrt(1.0 / p.shape[1]))
        elif p.dim() == 1:
            is_norm_w = leaf == 'weight'
            if 'learnable_fc.bias' in name:
                uniform(p, -2.0, 2.0)
            elif is_norm_w:
                uniform(p, 0.7, 1.3)
            else:
                normal(p, 0.05)
    return model
