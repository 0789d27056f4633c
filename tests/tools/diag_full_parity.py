"""Row-error distribution of the full-size cfg3 / cfg4 / cfg5 frames against the reference-generated fixtures, per precision.
    python tests/tools/diag_full_parity.py [cfg3 cfg4 cfg5] [--precision fp16x3 fp32]      (GPU box)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import ref_cases as C  # noqa: E402
from far3d_b200 import synthetic  # noqa: E402
from helpers import GOLDEN, build_oracle, build_product, rel_err, rel_l2, to_dev  # noqa: E402
from diag_cfg2_parity import rows  # noqa: E402


def main():
    args = sys.argv[1:]
    precs = [a for a in args if a.startswith('fp')] or ['fp16x3', 'fp32']
    names = [a for a in args if a.startswith('cfg')] or ['cfg3', 'cfg4', 'cfg5']
    dev = torch.device('cuda:0')
    for name in names:
        z = np.load(os.path.join(GOLDEN, f'ref_{name}_frames.npz'))
        case = C.FULL_CASES[name]
        mc = C.full_model_cfg(name)
        o = build_oracle(mc, seed=0)
        synthetic.cold_2d_head_(o, C.CFG2_HEAD_SCALE, C.CFG2_HEAD_SCALE)
        sd = o.state_dict()
        del o
        for prec in precs:
            p = build_product(mc, sd, dev, prec)
            for f in range(case['frames']):
                metas, data = synthetic.make_frame(case['rig'], f)
                res = p.simple_test(metas, **to_dev(data, dev))
                outs = p.last_outs
                nq = outs['all_cls_scores'].shape[2]
                print(f'[{name} {prec}] frame {f}: queries {nq} (fixture {z[f"cls{f}"].shape[1]})')
                ff = torch.from_numpy(C.sample(outs['feat_flatten'].float().cpu()))
                print('   feat_flatten sample rel_l2 %.2e max %.2e' % (rel_l2(ff, torch.from_numpy(z[f'feat_flatten{f}'])), rel_err(ff, torch.from_numpy(z[f'feat_flatten{f}']))))
                print('   cls rows (all, matched) ', rows(outs['all_cls_scores'][-1][0], z[f'cls{f}'][0], True))
                print('   box rows (all, matched) ', rows(outs['all_bbox_preds'][-1][0], z[f'box{f}'][0], True))
                if nq == z[f'cls{f}'].shape[1] and f == 0:
                    nfix = p.pts_bbox_head.num_query + outs['reference_points2d'].shape[1]
                    print('   cls rows (fixed part)   ', rows(outs['all_cls_scores'][-1][0, :nfix], z[f'cls{f}'][0, :nfix]))
                b = res[0]['pts_bbox']
                s, zs = torch.as_tensor(b['scores_3d']).float().cpu(), torch.from_numpy(z[f'scores3d{f}'])
                n = min(len(s), len(zs))
                print('   top-300 scores rel', rel_err(s[:n], zs[:n]), 'labels equal',
                      float((torch.as_tensor(b['labels_3d']).cpu().numpy()[:n] == z[f'labels3d{f}'][:n]).mean()),
                      'boxes rows', rows(torch.as_tensor(b['boxes_3d']).float(), z[f'boxes3d{f}'], True))
            del p
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
