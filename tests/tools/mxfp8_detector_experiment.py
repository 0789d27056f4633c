"""Detector-level check of the round-2 conv operand scheme (see mxfp8_correction_experiment.py): the CPU oracle at cfg-2 FULL SIZE
(7 x 960 x 640, V-99, 6 decoder layers, ~150 adaptive queries) with every image-branch convolution emulated as fp16 main term +
block-scaled-FP8 correction stream, against the plain fp32 oracle - the row-error statistics the GPU parity tests apply
(tests/test_gpu_ref_golden.py::test_cfg2_full_size_vs_reference: >= 99 % of the logit rows within 2e-3, every row within 4e-2,
top-300 scores within 1e-3, labels identical).
    python tests/tools/mxfp8_detector_experiment.py [mx8|mx6]      (CPU only, a few minutes)"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), HERE]
import mxfp8_correction_experiment as E  # noqa: E402
import ref_cases as C  # noqa: E402
from far3d_b200 import api, synthetic  # noqa: E402
from helpers import build_oracle  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else 'mx8'
    torch.set_num_threads(8)
    mc = api.load_model_cfg(num_cams=7)
    o = build_oracle(mc, seed=0)
    synthetic.cold_2d_head_(o, C.CFG2_HEAD_SCALE, C.CFG2_HEAD_SCALE)
    metas, data = synthetic.make_frame('cfg2', 0)
    outs = {}
    for m in ('exact', mode):
        E.MODE['m'] = m
        F.conv2d = E.conv_hook if m != 'exact' else E.orig
        torch.nn.functional.conv2d = F.conv2d
        o.prev_scene_token = None
        res, out = o.simple_test(metas, **data)
        outs[m] = (res, out)
        print(m, 'queries', out['all_cls_scores'].shape[2], flush=True)
    F.conv2d = torch.nn.functional.conv2d = E.orig
    (r0, a), (r1, b) = outs['exact'], outs[mode]
    if a['all_cls_scores'].shape != b['all_cls_scores'].shape:
        print('adaptive-query count differs:', a['all_cls_scores'].shape, b['all_cls_scores'].shape)
        return
    ff = ((b['feat_flatten'] - a['feat_flatten']).norm() / a['feat_flatten'].norm()).item()
    print(f'feat_flatten rel-L2 {ff:.2e}')
    for key in ('all_cls_scores', 'all_bbox_preds'):
        x, y = b[key][-1][0].double(), a[key][-1][0].double()
        d = (x - y).abs().amax(1) / y.abs().max()
        q = torch.quantile(d, torch.tensor([0.5, 0.9, 0.99, 0.999], dtype=torch.double)).tolist()
        print(f'{key} rows: q50 {q[0]:.1e} q90 {q[1]:.1e} q99 {q[2]:.1e} q99.9 {q[3]:.1e} max {d.max().item():.1e} | '
              f'<1e-3: {(d < 1e-3).double().mean().item():.4f} <2e-3: {(d < 2e-3).double().mean().item():.4f}')
    s0, s1 = r0[0]['pts_bbox']['scores_3d'], r1[0]['pts_bbox']['scores_3d']
    print('top-300 scores rel', ((s1 - s0).abs().max() / s0.abs().max()).item(),
          'labels equal', float((r0[0]['pts_bbox']['labels_3d'] == r1[0]['pts_bbox']['labels_3d']).float().mean()))


if __name__ == '__main__':
    main()
