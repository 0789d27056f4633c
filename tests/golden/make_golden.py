"""Generates the golden fixtures under tests/golden/ from the CPU oracle (python tests/golden/make_golden.py).

These two fixtures are outputs of the ORACLE (regression pins for it and, through the GPU parity tests, for the CUDA
path).  The fixtures that come from the reference's own code are made by make_ref_golden.py next to this file.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from far3d_b200 import synthetic  # noqa: E402
from helpers import build_oracle, model_cfg  # noqa: E402
from oracle import cref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def deform_agg_small():
    g = torch.Generator().manual_seed(7)
    B, N, Nq, G, P, D = 1, 3, 24, 2, 5, 32
    shapes = np.array([[16, 24], [8, 12], [4, 6]], dtype=np.int64)
    start = np.concatenate([[0], np.cumsum(shapes.prod(1))[:-1]]).astype(np.int64)
    S, C, L = int(shapes.prod(1).sum()), G * D, 3
    _, data = synthetic.make_frame((N, 128, 192), 0)
    feat = torch.randn(B * N, S, C, generator=g).numpy()
    kp = (torch.randn(B, Nq, P, 3, generator=g) * torch.tensor([20., 20., 2.])).numpy()
    w = torch.softmax(torch.randn(B, Nq, G, N * L * P, generator=g), -1).view(B, Nq, G, N, L * P).permute(0, 3, 1, 2, 4) \
        .reshape(B * N, Nq, G, L * P).contiguous().numpy()
    l2i = data['lidar2img'].numpy()
    out, uv, idx, valid = cref.deform_agg(feat, shapes, start, kp, l2i, w, 128, 192, G, debug=True)
    np.savez_compressed(os.path.join(HERE, 'deform_agg_small.npz'), feat=feat, shapes=shapes, start=start, key_points=kp,
                        lidar2img=l2i, weights=w, pad_hw=np.array([128, 192]), num_groups=G, out=out, uv=uv, idx=idx,
                        valid=valid)
    print('deform_agg_small: valid fraction', valid.mean())


def tiny_model():
    o = build_oracle(model_cfg(), seed=1)
    z = {}
    for f in range(2):
        metas, data = synthetic.make_frame('tiny', f)
        res, outs = o.simple_test(metas, **data)
        z[f'cls{f}'] = outs['all_cls_scores'].numpy()
        z[f'box{f}'] = outs['all_bbox_preds'].numpy()
        print('frame', f, 'queries', outs['all_cls_scores'].shape[2])
    np.savez_compressed(os.path.join(HERE, 'tiny_model.npz'), **z)


if __name__ == '__main__':
    cref.build(force=True)
    deform_agg_small()
    tiny_model()
