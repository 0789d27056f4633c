"""Algorithmic (compulsory) bytes of the image-branch convolutions of one cfg-2 frame: every conv reads its input tensor once,
writes its output once, reads its weights once, activations at 4 bytes per element (fp32, or the fp16 hi+lo plane pair) - the
denominator next to the ncu DRAM-traffic figure in bench.py's `roofline.traffic`.  Runs the CPU oracle's image branch once
with torch.nn.functional.conv2d hooked (build container or GPU box; ~10 s).
    python tests/tools/conv_algorithmic_bytes.py [cfg2]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from far3d_b200 import api, synthetic  # noqa: E402
from helpers import build_oracle  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    N, H, W = synthetic.CONFIGS[cfg]
    o = build_oracle(api.load_model_cfg(num_cams=N), seed=0)
    acc = dict(n=0, act=0, w=0, flop=0)
    orig = F.conv2d

    def hook(x, w, b=None, *a, **k):
        y = orig(x, w, b, *a, **k)
        if x.shape[-1] > 1:                                  # skip the eSE 1x1 convs on pooled (N,C,1,1) vectors
            acc['n'] += 1
            acc['act'] += 4 * (x.numel() + y.numel())
            acc['w'] += 4 * w.numel()
            acc['flop'] += 2 * y.numel() * w.shape[1] * w.shape[2] * w.shape[3]
        return y
    F.conv2d = hook
    torch.nn.functional.conv2d = hook
    _, data = synthetic.make_frame(cfg, 0)
    with torch.no_grad():
        feats = o.extract_img_feat(data['img'])
        o.img_roi_head(None, img_feats=feats) if hasattr(o, 'img_roi_head') and o.img_roi_head is not None else None
    print(f"{cfg}: {acc['n']} convolutions, {acc['flop'] / 1e12:.3f} TFLOP, activations in+out {acc['act'] / 1e9:.3f} GB, "
          f"weights {acc['w'] / 1e6:.1f} MB, algorithmic bytes {(acc['act'] + acc['w']) / 1e9:.3f} GB per frame")


if __name__ == '__main__':
    main()
