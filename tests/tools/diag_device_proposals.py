"""Diagnostic: device proposal kernels vs torch glue, per frame / per output row-error statistics (tiny config)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import build_oracle, build_product, model_cfg, rel_err, to_dev  # noqa: E402
from far3d_b200 import synthetic  # noqa: E402

cuda = torch.device('cuda:0')
mc = model_cfg()
o = build_oracle(mc, seed=1)
outs = {}
for kernels in (False, True):
    p = build_product(mc, o.state_dict(), cuda, 'fp16x3')
    p.pts_bbox_head.proposal_kernels = kernels
    res = []
    for f in range(2):
        metas, data = synthetic.make_frame('tiny', f)
        p.simple_test(metas, **to_dev(data, cuda))
        lo = p.last_outs
        res.append((lo['reference_points2d'].float().cpu(), lo['all_cls_scores'].float().cpu(), lo['all_bbox_preds'].float().cpu()))
    outs[kernels] = res
for f in range(2):
    a, b = outs[False][f], outs[True][f]
    print('frame', f, 'ref2d', rel_err(b[0], a[0]), 'cls', rel_err(b[1], a[1]), 'box', rel_err(b[2], a[2]))
    d = (b[1][-1, 0] - a[1][-1, 0]).abs().max(-1).values
    print('   last layer cls rows: max', float(d.max()), 'rows > 1e-4:', int((d > 1e-4).sum()), 'of', d.numel(), 'worst rows', d.topk(5).indices.tolist())
    for l in range(a[1].shape[0]):
        dl = (b[1][l, 0] - a[1][l, 0]).abs().max(-1).values
        print('   layer', l, 'max', float(dl.max()), 'rows > 1e-4:', int((dl > 1e-4).sum()))
