"""The `streaming_adaptive` operating point of bench.py alone (one streaming scene, ~150 adaptive queries per frame, temporal memory
bank live), pipelined and one frame at a time, for A/B runs of library builds (FAR3D_LIB_PATH) and launch knobs.

    python tools/adaptive_bench.py [--steps 10] [--reserve BYTES] [--pdl 0|1]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from far3d_b200 import api, ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--reserve', type=int, default=-1)
    ap.add_argument('--pdl', type=int, default=-1)
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    if args.reserve >= 0:
        ops.conv_umma_tune7(args.reserve)
    if args.pdl >= 0:
        ops.conv_umma_tune8(args.pdl)
    N, H, W = synthetic.CONFIGS['cfg2']
    mc = api.load_model_cfg(num_cams=N)
    apipe = api.Far3DPipeline(mc, device=dev, precision='fp16mx', seed=0)
    h_ = apipe.model.img_roi_head
    for c_, o_ in zip(h_.multi_level_conv_cls, h_.multi_level_conv_obj):
        c_.weight.data.mul_(5.0); o_.weight.data.mul_(5.0)
    for r_ in h_.multi_level_conv_reg:
        r_.weight.data.mul_(0.05); r_.bias.data.mul_(0.05)
    h_.invalidate()
    FA, K = 8, args.steps
    adev = []
    for i in range(FA):
        m, d = synthetic.make_frame('cfg2', i, seed=0)
        adev.append((m, {k: v.to(dev) for k, v in d.items()}))

    def step(i, pipelined):
        metas, d = adev[i % FA]
        m = [dict(metas[0], scene_token='stream')]
        if not pipelined:
            return apipe.infer_device(m, **dict(d))
        apipe.submit(m, **dict(d))
        return apipe.collect() if apipe.pending() > 1 else None

    def flush():
        while apipe.pending():
            apipe.collect()

    for pipelined in (True, False, True):
        for i in range(FA + 2):
            step(i, pipelined)
        flush()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(FA + 2 + i, pipelined)
        flush()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        nq = apipe.model.last_outs['all_cls_scores'].shape[2]
        print(f'{"pipelined" if pipelined else "one frame at a time"}: {ms:.3f} ms per frame ({1e3 / ms:.1f} frames/s), {nq} queries, '
              f'decoder graphs {len(apipe.model.pts_bbox_head.__dict__.get("_graphs", {}))}, image graphs {len(apipe.model.__dict__.get("_img_graphs", {}))}')


if __name__ == '__main__':
    main()
