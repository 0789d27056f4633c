"""Per-launch breakdown of the convolution kernels of one cfg-2 frame (eager launches, CUDA events per launch): which layer
classes the image branch's time goes to.    python tools/conv_frame_breakdown.py [--precision fp16mx] [--config cfg2]"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from far3d_b200 import api, ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--precision', default='fp16mx')
    ap.add_argument('--config', default='cfg2')
    a = ap.parse_args()
    N, H, W = synthetic.CONFIGS[a.config]
    pipe = api.Far3DPipeline(api.load_model_cfg(num_cams=N), precision=a.precision, seed=0)
    metas, data = synthetic.make_frame(a.config, 0)
    img = data['img'].cuda()
    pipe.model.use_cuda_graph = False
    for _ in range(2):
        pipe.model.image_branch(img)
    torch.cuda.synchronize()
    acc = collections.OrderedDict()
    reps = 5
    for _ in range(reps):
        ops.PROFILE, ops.PROFILE_TAGS = [], []
        pipe.model.image_branch(img)
        torch.cuda.synchronize()
        for (name, work, e0, e1), tag in zip(ops.PROFILE, ops.PROFILE_TAGS):
            if name != 'conv_umma':
                continue
            t = acc.setdefault(tag, [0, 0.0, 0.0])
            t[0] += 1; t[1] += e0.elapsed_time(e1); t[2] += work
    ops.PROFILE = ops.PROFILE_TAGS = None
    tot = sum(v[1] for v in acc.values()) / reps
    print(f'{a.config} {a.precision}: {sum(v[0] for v in acc.values()) // reps} conv launches, {tot:.3f} ms per frame, '
          f'{sum(v[2] for v in acc.values()) / reps / tot / 1e9:.0f} TFLOP/s algorithmic')
    print(f'{"layer class":40s} {"n":>3s} {"us each":>8s} {"ms/frame":>9s} {"share":>6s} {"TFLOP/s":>8s}')
    for tag, (n, ms, work) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f'{tag:40s} {n // reps:3d} {ms / n * 1e3:8.1f} {ms / reps:9.3f} {ms / reps / tot * 100:5.1f}% {work / ms / 1e9:8.0f}')


if __name__ == '__main__':
    main()
