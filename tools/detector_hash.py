"""Bit-level fingerprint of one cfg-2 frame through the detector: run it under two builds of the library (FAR3D_LIB_PATH) or two
settings of a launch knob and diff the printed lines.  Also runs the frame twice in one process (run-to-run determinism).

    python tools/detector_hash.py [--precision fp16mx] [--pdl 0|1] [--eager]
"""
import argparse
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from far3d_b200 import api, ops, synthetic  # noqa: E402


def h(t):
    return hashlib.sha1(t.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='fp16mx')
    ap.add_argument('--config', default='cfg2')
    ap.add_argument('--pdl', type=int, default=-1)
    ap.add_argument('--eager', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    if args.pdl >= 0:
        ops.conv_umma_tune8(args.pdl)
    N, H, W = synthetic.CONFIGS[args.config]
    mc = api.load_model_cfg(num_cams=N)
    pipe = api.Far3DPipeline(mc, device=dev, precision=args.precision, seed=0)
    if args.eager:
        pipe.model.use_cuda_graph = False
        pipe.model.pts_bbox_head.use_cuda_graph = False
    metas, d = synthetic.make_frame(args.config, 0, seed=0)
    d = {k: v.to(dev) for k, v in d.items()}
    for rep in range(3):
        pipe.infer_device([dict(metas[0], scene_token=f'scene{rep}')], **dict(d))
        torch.cuda.synchronize()
        lo = pipe.model.last_outs
        print(f'rep {rep}: cls {h(lo["all_cls_scores"])} box {h(lo["all_bbox_preds"])}')


if __name__ == '__main__':
    main()
