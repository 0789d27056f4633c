"""Build + ctypes binding of oracle/deform_agg_ref.c (TEST INFRASTRUCTURE)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libfar3d_oracle.so')
_SRC = os.path.join(_HERE, 'deform_agg_ref.c')


def build(force=False):
    """gcc the C restatement into oracle/_build/libfar3d_oracle.so (git-ignored, travels via gpurun)."""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', _SO, _SRC, '-lm'])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def project(key_points, lidar2img, pad_h, pad_w):
    kp = np.ascontiguousarray(key_points, dtype=np.float32)
    m = np.ascontiguousarray(lidar2img, dtype=np.float32)
    B, Nq, P, _ = kp.shape
    N = m.shape[1]
    uv = np.empty((B, N, Nq, P, 2), np.float32)
    lib().far3d_oracle_project(_p(kp, ctypes.c_float), _p(m, ctypes.c_float), ctypes.c_float(pad_h),
                               ctypes.c_float(pad_w), B, N, Nq, P, _p(uv, ctypes.c_float))
    return uv


def msda(value, shapes, start, loc, w, debug=False):
    value = np.ascontiguousarray(value, dtype=np.float32)
    loc = np.ascontiguousarray(loc, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    start = np.ascontiguousarray(start, dtype=np.int64)
    BN, S, G, D = value.shape
    _, Nq, _, L, P, _ = loc.shape
    out = np.empty((BN, Nq, G * D), np.float32)
    idx = np.empty((BN, Nq, G, L, P, 2), np.int32) if debug else None
    valid = np.empty((BN, Nq, G, L, P), np.uint8) if debug else None
    lib().far3d_oracle_msda(_p(value, ctypes.c_float), _p(shapes, ctypes.c_int64), _p(start, ctypes.c_int64),
                            _p(loc, ctypes.c_float), _p(w, ctypes.c_float), BN, S, G, D, Nq, L, P,
                            _p(out, ctypes.c_float), _p(idx, ctypes.c_int32), _p(valid, ctypes.c_uint8))
    return (out, idx, valid) if debug else out


def deform_agg(feat, shapes, start, key_points, lidar2img, weights, pad_h, pad_w, num_groups, debug=False):
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    kp = np.ascontiguousarray(key_points, dtype=np.float32)
    m = np.ascontiguousarray(lidar2img, dtype=np.float32)
    w = np.ascontiguousarray(weights, dtype=np.float32)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    start = np.ascontiguousarray(start, dtype=np.int64)
    BN, S, C = feat.shape
    B, Nq, P, _ = kp.shape
    N = m.shape[1]
    L = shapes.shape[0]
    G = num_groups
    D = C // G
    assert BN == B * N and w.shape == (BN, Nq, G, L * P)
    out = np.empty((B, Nq, C), np.float32)
    uv = np.empty((B, N, Nq, P, 2), np.float32) if debug else None
    idx = np.empty((B, N, Nq, L, P, 2), np.int32) if debug else None
    valid = np.empty((B, N, Nq, L, P), np.uint8) if debug else None
    lib().far3d_oracle_deform_agg(_p(feat, ctypes.c_float), _p(shapes, ctypes.c_int64), _p(start, ctypes.c_int64),
                                  _p(kp, ctypes.c_float), _p(m, ctypes.c_float), _p(w, ctypes.c_float),
                                  ctypes.c_float(pad_h), ctypes.c_float(pad_w), B, N, S, G, D, Nq, L, P,
                                  _p(out, ctypes.c_float), _p(uv, ctypes.c_float), _p(idx, ctypes.c_int32),
                                  _p(valid, ctypes.c_uint8))
    return (out, uv, idx, valid) if debug else out


if __name__ == '__main__':
    print(build(force=True))
