"""Per-CTA timeline of the conv kernels from in-kernel %globaltimer stamps (debug hook far3d_conv_umma_debug).

    python tools/conv_timeline.py --shape s2 --precision fp16x3 [--cm 1 --halo -1 --stages 3 --bn 0]
Stamps per CTA (ns): 0 start (after setup), 1 first operands landed, 2 last MMA issued, 3 accumulator ready,
4 epilogue done, 5 all loads issued.
"""
import argparse
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))

from far3d_b200 import _lib, ops  # noqa: E402
from prof_kernels import SHAPES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='s2')
    ap.add_argument('--precision', default='fp16x3')
    ap.add_argument('--bn', type=int, default=0)
    ap.add_argument('--stages', type=int, default=0)
    ap.add_argument('--grid', type=int, default=0)
    ap.add_argument('--halo', type=int, default=0)
    ap.add_argument('--exp', type=int, default=0)
    ap.add_argument('--cg', type=int, default=0, help='0 heuristic, 1 single CTA, 2 CTA pair')
    ap.add_argument('--ghz', default='1.9', help='SM clock assumed when converting the cycle counters')
    a = ap.parse_args()
    ops.conv_umma_tune(a.bn, a.stages)
    ops.conv_umma_tune2(a.grid, a.halo)
    ops.conv_umma_tune4(a.cg)
    dev = torch.device('cuda:0')
    N, H, W, Cin, Cout, k, s = SHAPES[a.shape]
    split = a.precision in ('fp16x3', 'fp16mx')
    mx = a.precision == 'fp16mx'
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
    fn = lambda: ops.conv2d_umma(x_hi, x_lo, N, H, W, Cin, 0, Cin, w_hi, w_lo, b, Cout, k, s, 1, y_hi=yh, y_lo=yl, yb_cs=Cout,
                                 x_fmt=fmt, w_exp=w_exp, y_fmt=fmt)
    fn(); torch.cuda.synchronize()
    dbg = torch.zeros(1 << 16, 16, dtype=torch.int64, device=dev)
    _lib.load().far3d_conv_umma_debug(ctypes.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    _lib.load().far3d_conv_umma_debug(None)
    dall = dbg.cpu().double()
    d = dall[dall[:, 0] > 1e15]
    t0 = d[:, 0].min()
    q = lambda v: ' '.join(f'{float(v.quantile(p_)):8.2f}' for p_ in (0.1, 0.5, 0.9))
    us = lambda col: (d[:, col] - d[:, 0]) / 1e3
    print(f'{a.shape} {a.precision} bn={a.bn} stages={a.stages} grid={a.grid} halo={a.halo} cg={a.cg}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us '
          f'(CUDA events, includes the host launch gap), {d.shape[0]} CTAs, first CTA start -> last CTA end {float((d[:, 4].max() - t0) / 1e3):.1f} us')
    print('  per-CTA (us after the CTA\'s own start) p10 p50 p90:')
    print('   CTA start after the first CTA      :', q((d[:, 0] - t0) / 1e3))
    lead = d[d[:, 2] > 1e15]
    print('   all TMA loads issued               :', q((d[:, 5] - d[:, 0])[d[:, 5] > 1e15] / 1e3))
    if (d[:, 7] > 1e15).any():
        print('   kernel entry -> start (setup)      :', q((d[:, 0] - d[:, 7]) / 1e3))
    if (lead[:, 1] > 1e15).any():
        print('   first tile: MMAs issued (leaders)  :', q((lead[:, 1] - lead[:, 0])[lead[:, 1] > 1e15] / 1e3))
    if (d[:, 6] > 1e15).any():
        print('   first tile: accumulator ready      :', q((d[:, 6] - d[:, 0])[d[:, 6] > 1e15] / 1e3))
    print('   last MMA issued (leader CTAs)      :', q((lead[:, 2] - lead[:, 0]) / 1e3))
    print('   epilogue done                      :', q(us(3)))
    print('   CTA end                            :', q(us(4)))
    if lead[:, 15].max() > 0:
        # build with FAR3D_NVCC_EXTRA=-DFAR3D_CONV_WAITSTATS: cycles each role spent blocked on its barriers, at the measured clock
        ghz = float(a.ghz)
        tiles = lead[:, 15]
        w = lambda t, col: q(t[:, col] / ghz / 1e3)
        print(f'  wait accounting (us per CTA at {ghz} GHz; tiles per worker p10 p50 p90: {q(tiles)}):')
        print('   MMA warp blocked on acc_empty (epilogue behind) :', w(lead, 8))
        print('   MMA warp blocked on a_full  (A patch not landed):', w(lead, 9))
        print('   MMA warp blocked on b_full  (stage not landed)  :', w(lead, 10))
        print('   producer 0 blocked on b_empty (ring full)       :', w(d, 11))
        print('   producer 0 blocked on a_empty                   :', w(d, 12))
        print('   epilogue warp blocked on acc_full (MMA behind)  :', w(d, 13))
        print('   epilogue warp busy (TMEM -> global)             :', w(d, 14))
        print('   epilogue busy per tile                          :', q(lead[:, 14] / tiles.clamp(min=1) / ghz / 1e3))


if __name__ == '__main__':
    main()
