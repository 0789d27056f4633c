"""CUDA-event timing of the decoder-side kernels at cfg-2 sizes (900 queries, 1924 keys, 7 cameras).
    python tools/time_decoder_ops.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from far3d_b200 import ops  # noqa: E402


def timeit(name, fn, iters=20):
    """GPU time per call: `iters` calls captured in one CUDA graph (the Python / ctypes launch path costs ~15 us per call,
    more than most of these kernels), replayed and timed with events."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f'{name:44s} {1e3 * e0.elapsed_time(e1) / iters:8.1f} us')


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--cg', type=int, default=0, help='conv / linear kernel: 0 heuristic, 1 single CTAs, 2 CTA pairs')
    ap.add_argument('--bn', type=int, default=0)
    ap.add_argument('--umma', action='store_true', help='decoder GEMMs on far3d_linear_umma instead of far3d_linear_mma')
    a = ap.parse_args()
    ops.conv_umma_tune4(a.cg)
    ops.conv_umma_tune(a.bn, 0)
    ops.LINEAR_MMA = not a.umma
    print(f'cg {a.cg} bn {a.bn} linear_mma {ops.LINEAR_MMA}')
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    q, k, v = r(1, 900, 256), r(1, 1924, 256), r(1, 1924, 256)
    q, k, v = r(1, 1047, 256), r(1, 1924, 256), r(1, 1924, 256)
    for kg in (1, 2, 3, 4):
        ops.mha_tune(False, kg)
        timeit(f'mha 1047 x 1924, 8 heads, tensor cores, {kg} key group(s)', lambda: ops.mha(q, k, v, 8))
    ops.mha_tune(True)
    timeit('mha 1047 x 1924, 8 heads, SIMT', lambda: ops.mha(q, k, v, 8))
    ops.mha_tune(False)
    x, a = r(1047, 256), r(1047, 256)
    x4 = r(1047, 1024)
    w4, b4 = r(256, 1024) / 32, r(256)
    timeit('linear 1047x1024 -> 256 (FFN 2)', lambda: ops.linear(x4, w4, b4))
    xk, ak = r(1815, 256), r(1815, 256)
    wk, bk = r(512, 256) / 16, r(512)
    timeit('linear 1815x256 -> 512 (Q|K)', lambda: ops.linear(xk, wk, bk, x_add=ak))
    for N in (39, 416, 256, 1024):
        w, b = r(N, 256) / 16, r(N)
        timeit(f'linear 900x256 -> {N} (default mode)', lambda: ops.linear(x, w, b, x_add=a))
    c = r(7, 256)
    for N, K in ((256, 12), (256, 256), (416, 256)):
        w, b = r(N, K), r(N)
        xi = r(7, K)
        timeit(f'linear 7x{K} -> {N}', lambda: ops.linear(xi, w, b, act=1))
    gam, bet = r(256), r(256)
    timeit('layernorm 900x256', lambda: ops.layernorm(x, gam, bet, 1e-5, add=a))
    wq, wc = r(1, 900, 416), r(1, 7, 416)
    timeit('dfa_weights_softmax 900 x 7 x 416', lambda: ops.dfa_weights_softmax(wq, wc, 8))


if __name__ == '__main__':
    main()
