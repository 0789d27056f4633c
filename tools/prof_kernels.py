"""Isolated launches of the two roofline kernels at cfg-2 shapes, for ncu captures and quick timing.

    python tools/prof_kernels.py conv  [--shape s2|s3|s4|s5|c3|c4] [--precision fp16x3|fp16] [--iters 5]
    python tools/prof_kernels.py agg   [--iters 5]
Prints CUDA-event time per launch and the achieved algorithmic TFLOP/s or GB/s.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from far3d_b200 import ops, synthetic  # noqa: E402

SHAPES = {  # name: (N, H, W, Cin, Cout, k, stride)   cfg-2 backbone work-list classes
    'stem2': (7, 320, 480, 64, 64, 3, 1),
    's2': (7, 160, 240, 128, 128, 3, 1),
    'c2': (7, 160, 240, 768, 256, 1, 1),
    's3': (7, 80, 120, 160, 160, 3, 1),
    'c3': (7, 80, 120, 1312, 512, 1, 1),
    's4': (7, 40, 60, 192, 192, 3, 1),
    's4b': (7, 40, 60, 768, 192, 3, 1),
    'c4': (7, 40, 60, 1728, 768, 1, 1),
    's5': (7, 20, 30, 224, 224, 3, 1),
    'c5': (7, 20, 30, 2144, 1024, 1, 1),
    'fpn': (7, 80, 120, 256, 256, 3, 1),
    'fpn40': (7, 40, 60, 256, 256, 3, 1),
    'lat3': (7, 80, 120, 512, 256, 1, 1),
    'lat4': (7, 40, 60, 768, 256, 1, 1),
}


def time_it(fn, iters, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts), min(ts)


def conv(args):
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    names = list(SHAPES) if args.shape == 'all' else [args.shape]
    for name in names:
        N, H, W, Cin, Cout, k, s = SHAPES[name]
        split = args.precision in ('fp16x3', 'fp16mx')
        mx = args.precision == 'fp16mx'
        x = torch.randn(N, H, W, Cin, device=dev)
        w = torch.randn(Cout, k * k, Cin, device=dev) / (Cin * k * k) ** 0.5
        b = torch.randn(Cout, device=dev)
        fmt, w_exp = (ops.lo_mx(), 0) if mx else (0, 0)
        if mx:
            x_hi, x_lo = ops.split_planes(x, lo_fmt=fmt)
            w_hi, w_lo, w_exp = ops.pack_weight_mx(w)
        else:
            x_hi, x_lo = ops.split_fp16(x, want_lo=split)
            w_hi, w_lo = ops.split_fp16(w, want_lo=split)
        Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
        yh = torch.empty(N, Ho, Wo, Cout, device=dev, dtype=torch.float16)
        yl = torch.empty_like(yh) if split else None
        fn = lambda: ops.conv2d_umma(x_hi, x_lo, N, H, W, Cin, 0, Cin, w_hi, w_lo, b, Cout, k, s, 1, y_hi=yh, y_lo=yl,
                                     yb_cs=Cout, yb_co=0, x_fmt=fmt, w_exp=w_exp, y_fmt=fmt)
        avg, best = time_it(fn, args.iters, flush)
        fl = 2.0 * N * Ho * Wo * Cout * Cin * k * k
        print(f'conv {name:6s} {args.precision:7s} N{N} {H}x{W} {Cin}->{Cout} k{k} s{s}: avg {avg * 1e3:8.1f} us  best {best * 1e3:8.1f} us  '
              f'{fl / (avg * 1e-3) / 1e12:7.1f} TFLOP/s algorithmic')


def agg(args):
    dev = torch.device('cuda:0')
    N, H, W = synthetic.CONFIGS['cfg2']
    shapes = [(H // s, W // s) for s in (8, 16, 32, 64)]
    starts, S = [], 0
    for h, w in shapes:
        starts.append(S); S += h * w
    g = torch.Generator().manual_seed(0)
    Nq, G, P, L, C = args.nq, 8, 13, 4, 256
    _, data = synthetic.make_frame('cfg2', 0)
    feat = torch.randn(N, S, C, device=dev)
    if args.fp16:
        feat = feat.bfloat16()
    ref = torch.rand(1, Nq, 1, 3, generator=g) * torch.tensor([304.8, 304.8, 10.0]) - torch.tensor([152.4, 152.4, 5.0])
    kp = (ref + torch.rand(1, Nq, P, 3, generator=g) * 4 - 2).contiguous().to(dev)      # learnable_fc bias U(-2,2) m offsets
    w = torch.softmax(torch.randn(1, Nq, G, N * L * P, generator=g), -1).view(1, Nq, G, N, L * P).permute(0, 3, 1, 2, 4) \
        .reshape(N, Nq, G, L * P).contiguous().to(dev)
    l2i = data['lidar2img'].to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ops.deform_agg_tune(args.warps, not args.narrow, args.work_queue, args.u4, args.prepare256)
    fn = lambda: ops.deform_agg(feat, shapes, starts, kp, l2i, w, H, W, G)
    if args.prepared:
        import ctypes
        from far3d_b200 import _lib
        wq = torch.randn(1, Nq, L * P * G, generator=g).to(dev)
        wc = torch.randn(1, N, L * P * G, generator=g).to(dev)
        ops.deform_agg_prepared(feat, shapes, starts, kp, l2i, wq, wc, H, W, G)
        cnt, rec, wts = next(iter(ops._AGG_WS.values()))
        out = torch.empty(1, Nq, C, device=dev)
        hw, hw_p = ops._host_i32(shapes); st, st_p = ops._host_i32(starts)
        pre = lambda: ops.call('far3d_dfa_prepare', ops._ptr(wq), ops._ptr(wc), ops._ptr(kp), ops._ptr(l2i), hw_p, st_p, float(H), float(W),
                               1, N, Nq, G, L, P, S, C, None, ops._ptr(cnt), ops._ptr(rec), ops._ptr(wts), ops._stream())
        fn = lambda: ops.call('far3d_deform_agg_gather', ops._ptr(feat), 0, hw_p, st_p, ops._ptr(cnt), ops._ptr(rec), ops._ptr(wts),
                              ops._ptr(out), 1, N, S, C, G, Nq, L, P, ops._stream())
        pa, pb = time_it(pre, args.iters, flush)
        sm = lambda: ops.dfa_weights_softmax(wq, wc, G)
        sa, sb = time_it(sm, args.iters, flush)
        print(f'dfa_prepare Nq={Nq}: avg {pa * 1e3:.1f} us best {pb * 1e3:.1f} us   (dfa_weights_softmax alone: avg {sa * 1e3:.1f} us best {sb * 1e3:.1f} us)')
    avg, best = time_it(fn, args.iters, flush)
    by = N * S * C * feat.element_size() + N * Nq * G * L * P * 4 + Nq * P * 12 + N * 64 + Nq * C * 4
    _, _, valid = ops.deform_agg_debug(shapes, kp, l2i, H, W)
    warm = time_it(fn, args.iters, None)
    print(f'deform_agg warps={args.warps} narrow={args.narrow} warm L2 (back-to-back launches): avg {warm[0] * 1e3:.1f} us best {warm[1] * 1e3:.1f} us')
    print(f'deform_agg Nq={Nq} feat={feat.dtype}: avg {avg * 1e3:.1f} us best {best * 1e3:.1f} us  {by / (avg * 1e-3) / 1e9:.0f} GB/s algorithmic '
          f'({by / 1e6:.1f} MB), in-bounds samples {float(valid.float().mean()) * 100:.1f}% of cam x level x point grid')


def misc(args):
    """the memory-bound image-branch kernels at cfg-2 shapes: stem conv 1, FPN upsample-add (algorithmic bytes / time)"""
    dev = torch.device('cuda:0')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    N, H, W = synthetic.CONFIGS['cfg2']
    img = torch.randn(N, 3, H, W, device=dev)
    w, b = torch.randn(64, 3, 3, 3, device=dev) * 0.2, torch.randn(64, device=dev)
    Ho, Wo = H // 2, W // 2
    yh = torch.empty(N, Ho, Wo, 64, device=dev, dtype=torch.float16); yl = torch.empty_like(yh)
    fn = lambda: ops.stem_conv(img, w, b, 64, y_hi=yh, y_lo=yl)
    avg, best = time_it(fn, args.iters, flush)
    by = img.numel() * 4 + 2 * yh.numel() * 2
    print(f'stem_conv 3->64 s2 {N}x{H}x{W}: avg {avg * 1e3:.1f} us best {best * 1e3:.1f} us  {by / (avg * 1e-3) / 1e9:.0f} GB/s ({by / 1e6:.0f} MB)')
    for (hd, wd) in ((80, 120), (40, 60)):
        d = torch.randn(N, hd, wd, 256, device=dev); s_ = torch.randn(N, hd // 2, wd // 2, 256, device=dev)
        dh = torch.empty(N, hd, wd, 256, device=dev, dtype=torch.float16); dl = torch.empty_like(dh)
        fn = lambda: ops.upsample_add(d, s_, N, hd, wd, hd // 2, wd // 2, 256, dh, dl)
        avg, best = time_it(fn, args.iters, flush)
        by = d.numel() * 8 + s_.numel() * 4 + 2 * dh.numel() * 2
        print(f'upsample_add {N}x{hd}x{wd}x256: avg {avg * 1e3:.1f} us best {best * 1e3:.1f} us  {by / (avg * 1e-3) / 1e9:.0f} GB/s ({by / 1e6:.0f} MB)')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['conv', 'agg', 'misc'])
    ap.add_argument('--warps', type=int, default=4)
    ap.add_argument('--prepare256', action='store_true')
    ap.add_argument('--u4', action='store_true', help='aggregation: 4 instead of 8 two-sample loads in flight per lane')
    ap.add_argument('--work-queue', action='store_true', help='aggregation: resident wave of CTAs pulling work items')
    ap.add_argument('--prepared', action='store_true', help='aggregation: far3d_dfa_prepare + far3d_deform_agg_gather (each timed)')
    ap.add_argument('--narrow', action='store_true', help='aggregation: 128-bit one-sample-per-warp-load form')
    ap.add_argument('--shape', default='all')
    ap.add_argument('--precision', default='fp16x3')
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--nq', type=int, default=900)
    ap.add_argument('--fp16', action='store_true')
    ap.add_argument('--bn', type=int, default=0)
    ap.add_argument('--stages', type=int, default=0)
    ap.add_argument('--grid', type=int, default=0)
    ap.add_argument('--halo', type=int, default=0)
    ap.add_argument('--exp', type=int, default=0, help='experiment mask: 1 no epilogue, 2 no TMA, 4 no MMA')
    ap.add_argument('--cg', type=int, default=0, help='0 heuristic, 1 single CTA, 2 CTA pair')
    a = ap.parse_args()
    ops.conv_umma_tune(a.bn, a.stages)
    ops.conv_umma_tune2(a.grid, a.halo)
    ops.conv_umma_tune4(a.cg)
    dict(conv=conv, agg=agg, misc=misc)[a.what](a)
