"""Row-error distribution of the full-size cfg-2 frames against the reference-generated fixture, per precision mode.
    python tests/tools/diag_cfg2_parity.py            (GPU box)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import ref_cases as C  # noqa: E402
from far3d_b200 import api, synthetic  # noqa: E402
from helpers import GOLDEN, build_oracle, build_product, rel_err, rel_l2, to_dev  # noqa: E402


def rows(a, b, match=False):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    s = b.abs().max()
    d = (torch.cdist(a, b, p=float('inf')).min(1).values if match else (a - b).abs().amax(1)) / s
    q = torch.quantile(d, torch.tensor([0.5, 0.9, 0.99, 0.999], dtype=torch.double)).tolist()
    return 'q50 %.1e q90 %.1e q99 %.1e q99.9 %.1e max %.1e | <1e-3: %.4f <2e-3: %.4f' % (*q, d.max().item(), (d < 1e-3).double().mean().item(), (d < 2e-3).double().mean().item())


def main():
    dev = torch.device('cuda:0')
    z = np.load(os.path.join(GOLDEN, 'ref_cfg2_frames.npz'))
    mc = api.load_model_cfg(num_cams=7)
    o = build_oracle(mc, seed=0)
    synthetic.cold_2d_head_(o, C.CFG2_HEAD_SCALE, C.CFG2_HEAD_SCALE)
    sd = o.state_dict()
    del o
    for prec in (sys.argv[1:] or ('fp32', 'fp16x3', 'fp16mx', 'fp16')):
        p = build_product(mc, sd, dev, prec)
        for f in range(C.CFG2_FRAMES):
            metas, data = synthetic.make_frame('cfg2', f)
            res = p.simple_test(metas, **to_dev(data, dev))
            outs = p.last_outs
            nq = outs['all_cls_scores'].shape[2]
            print(f'[{prec}] frame {f}: queries {nq} (fixture {z[f"cls{f}"].shape[1]})')
            if nq != z[f'cls{f}'].shape[1]:
                continue
            nfix = p.pts_bbox_head.num_query + outs['reference_points2d'].shape[1]
            ff = torch.from_numpy(C.sample(outs['feat_flatten'].float().cpu()))
            print('   feat_flatten sample rel_l2 %.2e max %.2e' % (rel_l2(ff, torch.from_numpy(z[f'feat_flatten{f}'])), rel_err(ff, torch.from_numpy(z[f'feat_flatten{f}']))))
            print('   ref2d rel', rel_err(outs['reference_points2d'], torch.from_numpy(z[f'ref2d{f}'])))
            print('   cls rows (fixed part)   ', rows(outs['all_cls_scores'][-1][0, :nfix], z[f'cls{f}'][0, :nfix]))
            print('   box rows (fixed part)   ', rows(outs['all_bbox_preds'][-1][0, :nfix], z[f'box{f}'][0, :nfix]))
            print('   cls rows (all, matched) ', rows(outs['all_cls_scores'][-1][0], z[f'cls{f}'][0], True))
            od = torch.from_numpy(C.sample(outs['outs_dec'].float().cpu()))
            print('   outs_dec sample rel_l2 %.2e max %.2e' % (rel_l2(od, torch.from_numpy(z[f'outs_dec{f}'])), rel_err(od, torch.from_numpy(z[f'outs_dec{f}']))))
            print('   top-300 scores rel', rel_err(res[0]['pts_bbox']['scores_3d'], torch.from_numpy(z[f'scores3d{f}'])),
                  'labels equal', float((res[0]['pts_bbox']['labels_3d'].cpu().numpy() == z[f'labels3d{f}']).mean()))
        del p
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
