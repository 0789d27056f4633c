"""Prints where the CUDA detector departs from the oracle on the tiny config (debug aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import build_oracle, build_product, model_cfg, rel_err, to_dev
from far3d_b200 import synthetic, ops

prec = sys.argv[1] if len(sys.argv) > 1 else 'fp16x3'
if len(sys.argv) > 2:
    ops.LINEAR_MODE = sys.argv[2]
dev = torch.device('cuda:0')
mc = model_cfg()
o = build_oracle(mc, seed=1)
p = build_product(mc, o.state_dict(), dev, prec)
for f in range(2):
    metas, data = synthetic.make_frame('tiny', f)
    d = to_dev(data, dev)
    with torch.no_grad():
        fo = o.extract_img_feat(data['img'])
        fp = p.extract_img_feat(d['img'])
        print('frame', f, 'fpn', [f'{rel_err(a, b):.1e}' for a, b in zip(fp, fo)])
        ro = o.img_roi_head(None, img_feats=fo)
        rp = p.img_roi_head(None, img_feats=fp)
        for k in ('enc_cls_scores', 'enc_bbox_preds', 'objectnesses'):
            print('  ', k, [f'{rel_err(a, b):.1e}' for a, b in zip(rp[k], ro[k])])
        print('   depth_logit', f"{rel_err(rp['depth_logit'], ro['depth_logit']):.1e}",
              'argmax equal', bool((rp['depth_logit'].argmax(1).cpu() == ro['depth_logit'].argmax(1)).all()))
        bo, bp = o.img_roi_head.get_bboxes(ro), p.img_roi_head.get_bboxes(rp)
        print('   valid equal', bool((bo['valid_indices'] == bp['valid_indices'].cpu()).all()), 'M', int(bo['valid_indices'].sum()))
    res_o, outs_o = o.simple_test(metas, **data)
    res_p = p.simple_test(metas, **d)
    op = p.last_outs
    r2o, r2p = outs_o['reference_points2d'], op['reference_points2d']
    print('   ref2d', None if r2o is None else f'{rel_err(r2p, r2o):.1e}')
    for l in range(outs_o['outs_dec'].shape[0]):
        print('   outs_dec layer', l, f"{rel_err(op['outs_dec'][l], outs_o['outs_dec'][l]):.1e}")
    print('   cls', f"{rel_err(op['all_cls_scores'], outs_o['all_cls_scores']):.1e}", 'box', f"{rel_err(op['all_bbox_preds'], outs_o['all_bbox_preds']):.1e}")
    e = (op['outs_dec'][-1].cpu() - outs_o['outs_dec'][-1]).abs().amax(-1)[0]
    print('   worst queries', e.topk(5).indices.tolist(), [f'{v:.1e}' for v in e.topk(5).values.tolist()], 'of', e.numel())
