"""Where the aggregation kernel's time goes: per-CTA cycle counts of its phases (build with FAR3D_NVCC_EXTRA=-DFAR3D_DA_PHASES).

    FAR3D_NVCC_EXTRA=-DFAR3D_DA_PHASES python -m far3d_b200.build --force && python tools/agg_phases.py [--nq 1047] [--work-queue] [--u4]
"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from far3d_b200 import _lib, ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nq', type=int, default=1047)
    ap.add_argument('--warps', type=int, default=4)
    ap.add_argument('--work-queue', action='store_true')
    ap.add_argument('--u4', action='store_true')
    ap.add_argument('--ghz', type=float, default=1.92)
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    N, H, W = synthetic.CONFIGS['cfg2']
    shapes = [(H // s, W // s) for s in (8, 16, 32, 64)]
    starts, S = [], 0
    for h, w in shapes:
        starts.append(S); S += h * w
    g = torch.Generator().manual_seed(0)
    Nq, G, P, L, C = a.nq, 8, 13, 4, 256
    _, data = synthetic.make_frame('cfg2', 0)
    feat = torch.randn(N, S, C, device=dev)
    ref = torch.rand(1, Nq, 1, 3, generator=g) * torch.tensor([304.8, 304.8, 10.0]) - torch.tensor([152.4, 152.4, 5.0])
    kp = (ref + torch.rand(1, Nq, P, 3, generator=g) * 4 - 2).contiguous().to(dev)
    w = torch.softmax(torch.randn(1, Nq, G, N * L * P, generator=g), -1).view(1, Nq, G, N, L * P).permute(0, 3, 1, 2, 4) \
        .reshape(N, Nq, G, L * P).contiguous().to(dev)
    l2i = data['lidar2img'].to(dev)
    lib = ctypes.CDLL(_lib.load()._name)
    buf = torch.zeros(8192, 8, dtype=torch.int64, device=dev)
    ops.deform_agg_tune(a.warps, True, a.work_queue, a.u4)
    fn = lambda: ops.deform_agg(feat, shapes, starts, kp, l2i, w, H, W, G)
    fn(); torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush.zero_()
    assert lib.far3d_deform_agg_phases(ctypes.c_void_p(buf.data_ptr())) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.far3d_deform_agg_phases(None)
    d = buf.cpu().double()
    d = d[d[:, 5] > 0]
    us = d[:, :5].sum(1) / a.ghz / 1e3
    names = ['phase A (project + scan)', 'records', 'softmax-weight gather', 'feature gather', 'fold + store', 'items', 'in-view samples', 'queue pull']
    print(f'Nq {Nq} warps {a.warps} work queue {a.work_queue} u4 {a.u4}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us (cold L2), {d.shape[0]} CTAs with work, '
          f'items per CTA mean {float(d[:, 5].mean()):.2f} max {int(d[:, 5].max())}; busy per CTA mean {float(us.mean()):.1f} us max {float(us.max()):.1f} us')
    for i, n in enumerate(names):
        if i in (5, 6):
            continue
        v = d[:, i] / a.ghz / 1e3
        print(f'   {n:28s} per CTA mean {float(v.mean()):7.2f} us  per item {float(d[:, i].sum() / d[:, 5].sum()) / a.ghz / 1e3:7.2f} us')
    spi = d[:, 6] / d[:, 5]
    print(f'   in-view samples per item: mean {float(spi.mean()):.1f}  p90 {float(spi.quantile(0.9)):.1f}  max {float(spi.max()):.1f}')


if __name__ == '__main__':
    main()
