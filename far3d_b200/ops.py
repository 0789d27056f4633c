"""Tensor-level wrappers over the C ABI (include/far3d_b200.h).  torch is used for device memory and streams
only; every op runs in libfar3d_sm100.so and raises if the tensor is not on a CUDA device."""
import ctypes

import numpy as np
import os

import torch

from . import _lib

call = _lib.call

# optional per-launch CUDA-event profiler (bench.py turns it on): list of (name, work, start_event, end_event)
PROFILE = None
PROFILE_TAGS = None     # optional parallel list of per-launch tags (layer shapes) for tools/conv_frame_breakdown.py


class _Timed:
    def __init__(self, name, work, tag=None):
        self.name, self.work, self.tag = name, work, tag

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, self.work, self.e0, self.e1))
            if PROFILE_TAGS is not None:
                PROFILE_TAGS.append(self.tag)
        return False


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype=torch.float32, name='tensor'):
    if not t.is_cuda:
        raise _lib.Far3DNativeError(f'{name} must be a CUDA tensor: far3d_b200 has no CPU path')
    if t.dtype != dtype:
        raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be contiguous')
    return t


def _host_i32(a):
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return arr, arr.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------------------------------ aggregation
def deform_agg(feat, spatial_shapes, level_start_index, key_points, lidar2img, weights, pad_h, pad_w, num_groups):
    """feat [B*N,S,C] fp32|bf16, key_points [B,Nq,P,3], lidar2img [B,N,4,4], weights [B*N,Nq,G,L*P] -> [B,Nq,C].
    spatial_shapes / level_start_index: host sequences."""
    dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[feat.dtype]
    _chk(feat, feat.dtype, 'feat')
    _chk(key_points, name='key_points'); _chk(lidar2img, name='lidar2img'); _chk(weights, name='weights')
    B, Nq, P, _ = key_points.shape
    N = lidar2img.shape[1]
    BN, S, C = feat.shape
    hw, hwp = _host_i32(spatial_shapes)
    st, stp = _host_i32(level_start_index)
    L = hw.shape[0]
    assert BN == B * N and tuple(weights.shape) == (BN, Nq, num_groups, L * P), (feat.shape, weights.shape)
    out = torch.empty(B, Nq, C, device=feat.device, dtype=torch.float32)
    # algorithmic (compulsory) bytes, SURVEY.md section 8d: features once + weights + key points + matrices + output
    work = BN * S * C * feat.element_size() + BN * Nq * num_groups * L * P * 4 + B * Nq * P * 12 + BN * 64 + B * Nq * C * 4
    with _Timed('deform_agg', work):
        call('far3d_deform_agg_fwd', _ptr(feat), dt, hwp, stp, _ptr(key_points), _ptr(lidar2img), _ptr(weights),
             float(pad_h), float(pad_w), _ptr(out), B, N, S, C, num_groups, Nq, L, P, _stream())
    return out


def deform_agg_tune(warps=4, wide=True, work_queue=False, u4=False, prepare256=False):
    """tools / tests: aggregation kernel variant - warps per CTA (4: a query's 8 channel groups over two work items; 8: one item
    per query; 2: four items), wide = 256-bit loads covering two samples per warp instruction (default) or the 128-bit
    one-sample form, work_queue = one resident wave of CTAs pulling items from a device-side queue instead of one CTA per item,
    u4 = 4 instead of 8 two-sample loads in flight per lane, prepare256 = far3d_dfa_prepare with 256- instead of 128-thread CTAs"""
    _lib.load().far3d_deform_agg_tune(int(warps), int(bool(wide)) | (2 if work_queue else 0) | (4 if u4 else 0) | (8 if prepare256 else 0))


def deform_agg_debug(spatial_shapes, key_points, lidar2img, pad_h, pad_w):
    _chk(key_points); _chk(lidar2img)
    B, Nq, P, _ = key_points.shape
    N = lidar2img.shape[1]
    hw, hwp = _host_i32(spatial_shapes)
    L = hw.shape[0]
    dev = key_points.device
    uv = torch.empty(B, N, Nq, P, 2, device=dev)
    idx = torch.empty(B, N, Nq, L, P, 2, device=dev, dtype=torch.int32)
    valid = torch.empty(B, N, Nq, L, P, device=dev, dtype=torch.uint8)
    call('far3d_deform_agg_debug', hwp, _ptr(key_points), _ptr(lidar2img), float(pad_h), float(pad_w), _ptr(uv),
         _ptr(idx), _ptr(valid), B, N, Nq, L, P, _stream())
    return uv, idx, valid


def msda(value, spatial_shapes, level_start_index, sampling_locations, attention_weights):
    """mmcv MultiScaleDeformableAttnFunction.forward layout (detr3d_transformer.py:561-563)."""
    _chk(value); _chk(sampling_locations); _chk(attention_weights)
    _chk(spatial_shapes, torch.int64); _chk(level_start_index, torch.int64)
    BN, S, G, D = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    out = torch.empty(BN, Nq, G * D, device=value.device)
    call('far3d_msda_fwd', _ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(sampling_locations),
         _ptr(attention_weights), _ptr(out), BN, S, G, D, Nq, L, P, _stream())
    return out


def dfa_weights_softmax(wq, wc, num_groups):
    """wq [B,Nq,LP*G], wc [B,N,LP*G] -> weights [B*N,Nq,G,LP] (softmax over cams x levels x points)."""
    _chk(wq); _chk(wc)
    B, Nq, J = wq.shape
    N = wc.shape[1]
    LP = J // num_groups
    out = torch.empty(B * N, Nq, num_groups, LP, device=wq.device)
    call('far3d_dfa_weights_softmax', _ptr(wq), _ptr(wc), _ptr(out), B, N, Nq, num_groups, LP, _stream())
    return out


def cam_logits(lidar2img, layers):
    """Camera side of the aggregation logits of several decoder layers in one launch.  lidar2img [B, N, 4, 4]; layers: sequence of
    (w0 [H,12], b0, w1 [E,H], b1, ln_weight, ln_bias, wfc [J,E]) fp32 CUDA tensors, same shapes and LayerNorm eps for every
    layer -> [len(layers), B, N, J]."""
    _chk(lidar2img, name='lidar2img')
    B, N = lidar2img.shape[:2]
    H, E, J = layers[0][0].shape[0], layers[0][2].shape[0], layers[0][6].shape[0]
    ptrs = (ctypes.c_void_p * (7 * len(layers)))()
    for i, t in enumerate(layers):
        assert t[0].shape == (H, 12) and t[2].shape == (E, H) and t[6].shape == (J, E), [tuple(x.shape) for x in t]
        for k, x in enumerate(t):
            _chk(x, name='cam layer tensor')
            ptrs[7 * i + k] = x.data_ptr()
    out = torch.empty(len(layers), B, N, J, device=lidar2img.device)
    call('far3d_cam_logits', _ptr(lidar2img), ctypes.cast(ptrs, ctypes.c_void_p), len(layers), B * N, E, H, J, 1e-5, _ptr(out), _stream())
    return out


def dfa_prepare_supported(N, G, L, P, C):
    return bool(_lib.load().far3d_dfa_prepare_supported(int(N), int(G), int(L), int(P), int(C)))


_AGG_WS = {}


def deform_agg_prepared(feat, spatial_shapes, level_start_index, key_points, lidar2img, wq, wc, pad_h, pad_w, num_groups,
                        want_weights=False):
    """The aggregation of deform_agg(..., weights=dfa_weights_softmax(wq, wc)) in its two-kernel form: far3d_dfa_prepare (softmax
    + projection + corner records + weights compacted to the in-view samples) -> far3d_deform_agg_gather.  Same sums in the same
    order (bit-identical output).  Returns out [B, Nq, C] (and the full weight tensor when want_weights)."""
    dt = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[feat.dtype]
    _chk(feat, feat.dtype, 'feat')
    _chk(key_points, name='key_points'); _chk(lidar2img, name='lidar2img'); _chk(wq, name='wq'); _chk(wc, name='wc')
    B, Nq, P, _ = key_points.shape
    N = lidar2img.shape[1]
    BN, S, C = feat.shape
    hw, hw_p = _host_i32(spatial_shapes)
    st, st_p = _host_i32(level_start_index)
    L = hw.shape[0]
    E = N * L * P
    G = num_groups
    assert BN == B * N and tuple(wq.shape) == (B, Nq, L * P * G) and tuple(wc.shape) == (B, N, L * P * G), (feat.shape, wq.shape, wc.shape)
    key = (feat.device, torch.cuda.current_stream().cuda_stream, B * Nq, E, G)          # (one workspace per stream: not shared by concurrent calls)
    ws = _AGG_WS.get(key)
    if ws is None:            # persistent workspace: stable addresses for captured graphs; rows are rewritten by every call
        ws = _AGG_WS[key] = (torch.zeros(B * Nq, dtype=torch.int32, device=feat.device),
                             torch.zeros(B * Nq, E, 4, 2, dtype=torch.int32, device=feat.device),
                             torch.zeros(B * Nq, G, E, device=feat.device))
    cnt, rec, wts = ws
    weights = torch.empty(BN, Nq, G, L * P, device=feat.device) if want_weights else None
    out = torch.empty(B, Nq, C, device=feat.device)
    with _Timed('dfa_prepare', 0.0):
        call('far3d_dfa_prepare', _ptr(wq), _ptr(wc), _ptr(key_points), _ptr(lidar2img), hw_p, st_p, float(pad_h), float(pad_w),
             B, N, Nq, G, L, P, S, C, _ptr(weights), _ptr(cnt), _ptr(rec), _ptr(wts), _stream())
    work = BN * S * C * feat.element_size() + BN * Nq * G * L * P * 4 + B * Nq * P * 12 + BN * 64 + B * Nq * C * 4
    with _Timed('deform_agg', work):
        call('far3d_deform_agg_gather', _ptr(feat), dt, hw_p, st_p, _ptr(cnt), _ptr(rec), _ptr(wts), _ptr(out),
             B, N, S, C, G, Nq, L, P, _stream())
    return (out, weights) if want_weights else out


# ------------------------------------------------------------------------------------------ dense
def linear(x, weight, bias=None, act=0, residual=None, out=None, x_add=None):
    """y = act((x + x_add) @ weight.T + bias) (+ residual).  x [..., K] (last dim contiguous), weight [N, K]."""
    K = x.shape[-1]
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    a2 = None
    if x_add is not None:
        a2 = x_add.reshape(-1, K)
        if a2.stride() != x2.stride():
            a2 = a2.contiguous(); x2 = x2.contiguous()
    if not x2.is_cuda:
        raise _lib.Far3DNativeError('linear: CUDA tensor required')
    M, N = x2.shape[0], weight.shape[0]
    _chk(weight, name='weight')
    y = out if out is not None else torch.empty(M, N, device=x.device)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, N)
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    ldr = r2.stride(0) if r2 is not None else 0
    if (LINEAR_MODE == 'fp16x3' and LINEAR_MMA and M >= 16 and N >= 16 and K % 32 == 0 and K <= LINEAR_MMA_MAX_K
            and x2.stride(0) % 2 == 0 and x2.data_ptr() % 8 == 0 and (a2 is None or a2.data_ptr() % 8 == 0)):
        # token matrices of the decoder: warp-level tensor-core MMAs, activations split on the fly, 20 KB / 128-thread CTAs
        w_hi, w_lo = _packed_weight(weight, True)
        call('far3d_linear_mma', _ptr(x2), _ptr(a2), x2.stride(0), _ptr(w_hi), _ptr(w_lo), _ptr(bias), _ptr(r2), ldr, _ptr(y),
             y.stride(0), M, N, K, int(act), _stream())
    elif (LINEAR_MODE != 'fp32' and M >= 64 and K % 16 == 0 and N % 4 == 0 and y.stride(0) % 4 == 0 and ldr % 4 == 0
            and x2.is_contiguous() and (a2 is None or a2.is_contiguous())):
        # tensor-core path: split-fp16 operands (fp16x3 = fp32-grade), weights split once and cached
        split = LINEAR_MODE == 'fp16x3'
        w_hi, w_lo = _packed_weight(weight, split)
        x_hi = torch.empty(M, K, device=x.device, dtype=torch.float16)
        x_lo = torch.empty_like(x_hi) if split else None
        call('far3d_split_fp16', _ptr(x2), _ptr(a2), _ptr(x_hi), _ptr(x_lo), M * K, _stream())
        call('far3d_linear_umma', _ptr(x_hi), _ptr(x_lo), K, _ptr(w_hi), _ptr(w_lo), _ptr(bias), _ptr(r2), ldr, _ptr(y),
             y.stride(0), M, N, K, int(act), _stream())
    else:
        call('far3d_linear_f32', _ptr(x2), _ptr(a2), x2.stride(0), _ptr(weight), _ptr(bias), _ptr(r2), ldr,
             _ptr(y), y.stride(0), M, N, K, int(act), _stream())
    return y.view(*x.shape[:-1], N)


# 'fp16x3' (default): nn.Linear layers with M >= 64 rows run on tcgen05 with split-fp16 operands (2^-17 relative);
# 'fp16': plain fp16 operands; 'fp32': exact fp32 SIMT kernel everywhere.
LINEAR_MODE = 'fp16x3'
# fp16x3 mode: GEMMs with K <= LINEAR_MMA_MAX_K go to far3d_linear_mma (mma.sync, no split launch, small CTAs that fit next to the
# persistent conv CTAs of the other frame in flight); longer K (the second FFN layer) and LINEAR_MMA = False: far3d_linear_umma
LINEAR_MMA = True
LINEAR_MMA_MAX_K = 512


PACK_STATS = [0, 0, 0]       # packed-weight cache hits / misses / misses during graph capture (diagnostics)


def _packed_weight(weight, split):
    """split-fp16 copy of a weight (or of a row-slice view of one), cached ON the owning parameter object so the cache
    lives and dies with the model and is invalidated by in-place updates (`_version`)."""
    base = weight._base if weight._base is not None else weight
    cache = getattr(base, '_far3d_packed', None)
    # `p.data = other` (what .to() / .half() do) keeps the Parameter object and its _version: the storage address, device
    # and dtype are part of the stamp
    stamp = (base._version, base.data_ptr(), str(base.device), base.dtype)
    if cache is None or cache[0] != stamp:
        cache = (stamp, {})
        try:
            base._far3d_packed = cache
        except Exception:
            pass
    key = (weight.storage_offset(), tuple(weight.shape), split)
    hit = cache[1].get(key)
    PACK_STATS[0 if hit is not None else 1] += 1
    if hit is None and torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        PACK_STATS[2] += 1               # a weight packed inside a graph capture would be re-packed at every replay
    if hit is None:
        hi = torch.empty(weight.shape, device=weight.device, dtype=torch.float16)
        lo = torch.empty_like(hi) if split else None
        call('far3d_split_fp16', _ptr(weight), None, _ptr(hi), _ptr(lo), weight.numel(), _stream())
        hit = (hi, lo)
        cache[1][key] = hit
    return hit


def layernorm(x, gamma, beta, eps=1e-5, add=None, relu_before=False, relu_after=False):
    _chk(x)
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty_like(x)
    if add is not None:
        _chk(add)
    call('far3d_layernorm', _ptr(x), _ptr(add), _ptr(gamma), _ptr(beta), _ptr(y), M, C, float(eps), int(relu_before),
         int(relu_after), _stream())
    return y


def mha_tune(simt=False, key_groups=0):
    """tools / tests: attention core on warp-level tensor-core MMAs (default) or the SIMT kernel; key_groups: warps groups per CTA
    of the tensor-core kernel that split the keys (1..4, 0 = default)"""
    _lib.load().far3d_mha_tune(int(bool(simt)) | (int(key_groups) << 4))


# device int32[2] {first masked key, number of masked keys} or None: the padding rows of a bucketed adaptive-query count
# (FarHead sets it around the decoder; the self-attention of every layer reads it)
MHA_KEY_SKIP = None


def mha(q, k, v, num_heads):
    """q [B,Nq,E], k/v [B,Nk,E] (last dim contiguous; row strides free) -> [B,Nq,E]."""
    B, Nq, E = q.shape
    Nk = k.shape[1]
    for t in (q, k, v):
        assert t.is_cuda and t.stride(-1) == 1 and (B == 1 or t.stride(0) == t.stride(1) * t.shape[1])
    o = torch.empty(B, Nq, E, device=q.device)
    if MHA_KEY_SKIP is not None:
        assert B == 1 and MHA_KEY_SKIP.dtype == torch.int32 and MHA_KEY_SKIP.numel() == 2
        call('far3d_mha_fwd_masked', _ptr(q), q.stride(1), _ptr(k), k.stride(1), _ptr(v), v.stride(1), _ptr(o), E, B, Nq, Nk,
             num_heads, E // num_heads, _ptr(MHA_KEY_SKIP), _stream())
    else:
        call('far3d_mha_fwd', _ptr(q), q.stride(1), _ptr(k), k.stride(1), _ptr(v), v.stride(1), _ptr(o), E, B, Nq, Nk,
             num_heads, E // num_heads, _stream())
    return o


# ------------------------------------------------------------------------------------------ temporal memory bank
def memory_post_update(cls_last, box_last, dec_last, K, ego_pose, timestamp, bank):
    """farhead.py:479-508 for one stream.  bank: dict emb [M,E], ref [M,3], ts [M] fp64, pose [M,4,4], velo [M,2] (contiguous).
    Returns (new bank with K + M rows, topk indices int32 [K])."""
    for t in (cls_last, box_last, dec_last, ego_pose):
        _chk(t)
    _chk(timestamp, torch.float64, 'timestamp'); _chk(bank['ts'], torch.float64, 'memory_timestamp')
    Nq, C = cls_last.shape
    code, E, M = box_last.shape[1], dec_last.shape[1], bank['emb'].shape[0]
    dev = cls_last.device
    idx = torch.empty(K, device=dev, dtype=torch.int32)
    new = dict(emb=torch.empty(K + M, E, device=dev), ref=torch.empty(K + M, 3, device=dev),
               ts=torch.empty(K + M, device=dev, dtype=torch.float64), pose=torch.empty(K + M, 4, 4, device=dev),
               velo=torch.empty(K + M, 2, device=dev))
    call('far3d_memory_post_update', _ptr(cls_last), _ptr(box_last), _ptr(dec_last), Nq, C, code, E, int(K), M, _ptr(ego_pose),
         _ptr(timestamp), _ptr(bank['emb']), _ptr(bank['ref']), _ptr(bank['ts']), _ptr(bank['pose']), _ptr(bank['velo']), _ptr(idx),
         _ptr(new['emb']), _ptr(new['ref']), _ptr(new['ts']), _ptr(new['pose']), _ptr(new['velo']), _stream())
    return new, idx


def memory_pre_update(bank, n, kprop, prev_exists, ego_pose_inv, timestamp, pseudo_points):
    """farhead.py:446-477 on an existing bank (>= n rows) -> new bank with n rows."""
    _chk(prev_exists); _chk(ego_pose_inv); _chk(timestamp, torch.float64, 'timestamp'); _chk(bank['ts'], torch.float64, 'memory_timestamp')
    E = bank['emb'].shape[1]
    dev = bank['emb'].device
    assert bank['emb'].shape[0] >= n
    new = dict(emb=torch.empty(n, E, device=dev), ref=torch.empty(n, 3, device=dev), ts=torch.empty(n, device=dev, dtype=torch.float64),
               pose=torch.empty(n, 4, 4, device=dev), velo=torch.empty(n, 2, device=dev))
    call('far3d_memory_pre_update', int(n), E, int(kprop), _ptr(prev_exists), _ptr(ego_pose_inv), _ptr(timestamp),
         _ptr(pseudo_points), _ptr(bank['emb']), _ptr(bank['ref']), _ptr(bank['ts']), _ptr(bank['pose']), _ptr(bank['velo']),
         _ptr(new['emb']), _ptr(new['ref']), _ptr(new['ts']), _ptr(new['pose']), _ptr(new['velo']), _stream())
    return new


# ------------------------------------------------------------------------------------------ box decode
def box_decode(cls, box, max_num, post_center_range, score_threshold=None, bottom_center=False):
    """cls [Nq,C] logits, box [Nq,code] -> (boxes [K,7|9], scores [K], labels int32 [K], query int32 [K], count int32 [1]): the
    surviving boxes first, in descending score order (nms_free_coder.py:39-112 [+ farhead.py:1237 with bottom_center])."""
    _chk(cls, name='cls'); _chk(box, name='box')
    Nq, C = cls.shape
    code = box.shape[1]
    K = int(max_num)
    dev = cls.device
    W = 9 if code > 8 else 7
    boxes = torch.zeros(K, W, device=dev)
    scores = torch.zeros(K, device=dev)
    labels = torch.zeros(K, device=dev, dtype=torch.int32)
    query = torch.zeros(K, device=dev, dtype=torch.int32)
    count = torch.zeros(1, device=dev, dtype=torch.int32)
    r = np.ascontiguousarray(np.asarray(post_center_range, dtype=np.float32))
    assert r.shape == (6,)
    call('far3d_box_decode', _ptr(cls), _ptr(box), Nq, C, code, K, r.ctypes.data_as(ctypes.c_void_p),
         float(score_threshold or 0.0), int(bool(bottom_center)), _ptr(boxes), _ptr(scores), _ptr(labels), _ptr(query), _ptr(count),
         _stream())
    return boxes, scores, labels, query, count


# ------------------------------------------------------------------------------------------ 2D proposals -> 3D queries
def roi_select(cls_maps, reg_maps, strides, num_classes, threshold, cap_per_cam):
    """cls_maps / reg_maps: per level NHWC fp32 [N,H,W,Ccls] / [N,H,W,8] -> dict of per-camera slots (yolox_head.py:355-489)."""
    N = cls_maps[0].shape[0]
    L = len(cls_maps)
    for c, r in zip(cls_maps, reg_maps):
        _chk(c, name='cls map'); _chk(r, name='reg map')
    dev = cls_maps[0].device
    hw = np.ascontiguousarray(np.asarray([[c.shape[1], c.shape[2]] for c in cls_maps], dtype=np.int32))
    st = np.ascontiguousarray(np.asarray(strides, dtype=np.int32))
    cp = (ctypes.c_void_p * L)(*[c.data_ptr() for c in cls_maps])
    rp = (ctypes.c_void_p * L)(*[r.data_ptr() for r in reg_maps])
    S2 = int(sum(c.shape[1] * c.shape[2] for c in cls_maps))
    ws = torch.empty(N * S2, device=dev)
    sel = dict(pos=torch.empty(N, cap_per_cam, device=dev, dtype=torch.int32), score=torch.empty(N, cap_per_cam, device=dev),
               box=torch.empty(N, cap_per_cam, 4, device=dev), counts=torch.empty(N, device=dev, dtype=torch.int32),
               N=N, cap=cap_per_cam, S2=S2, score_map=ws)
    call('far3d_roi_select', cp, rp, hw.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p), L, N, int(num_classes),
         cls_maps[0].shape[3], reg_maps[0].shape[3], float(threshold), _ptr(ws), int(cap_per_cam), _ptr(sel['pos']),
         _ptr(sel['score']), _ptr(sel['box']), _ptr(sel['counts']), _stream())
    return sel


def query2d_lift(sel, depth_logits, num_bins, down, topk, rmin_bin, dmin, bin_size, thr_logit, lidar2img, pc_range, S, cap_total):
    """slots of roi_select + depth logits NHWC [N,Hd,Wd,Dcs] -> ref2d [cap_total,3], src_row, score_feat, meta int32[4]
    {queries, primaries, multi-depth sources, overflow}  (farhead.py:710-827)."""
    _chk(depth_logits, name='depth logits'); _chk(lidar2img, name='lidar2img'); _chk(pc_range, name='pc_range')
    dev = depth_logits.device
    N, Hd, Wd, Dcs = depth_logits.shape
    assert N == sel['N'] and lidar2img.numel() == N * 16
    ref2d = torch.empty(cap_total, 3, device=dev)
    src_row = torch.empty(cap_total, device=dev, dtype=torch.int32)
    score_feat = torch.empty(cap_total, device=dev)
    meta = torch.empty(4, device=dev, dtype=torch.int32)
    ws = torch.empty(int(_lib.load().far3d_query2d_lift_workspace_ints(int(cap_total))), device=dev, dtype=torch.int32)
    call('far3d_query2d_lift', _ptr(sel['pos']), _ptr(sel['score']), _ptr(sel['box']), _ptr(sel['counts']), N, sel['cap'], int(S),
         _ptr(depth_logits), Hd, Wd, int(num_bins), Dcs, int(down), int(topk), int(rmin_bin), float(dmin), float(bin_size),
         float(thr_logit), _ptr(lidar2img), _ptr(pc_range), int(cap_total), _ptr(ref2d), _ptr(src_row), _ptr(score_feat),
         _ptr(meta), _ptr(ws), _stream())
    return ref2d, src_row, score_feat, meta


def ctx_gather(feat_flatten, src_row, score_feat, rows):
    """[rows, C + 1]: feat_flatten row of every query's 2D peak ++ its score channel (zeros for padding rows)."""
    _chk(feat_flatten, name='feat_flatten')
    C = feat_flatten.shape[-1]
    ctx = torch.empty(rows, C + 1, device=feat_flatten.device)
    call('far3d_ctx_gather', _ptr(feat_flatten), _ptr(src_row), _ptr(score_feat), C, int(rows), _ptr(ctx), _stream())
    return ctx


def pos2posemb3d(pos, num_pos_feats=128):
    _chk(pos)
    M = pos.numel() // 3
    emb = torch.empty(*pos.shape[:-1], 3 * num_pos_feats, device=pos.device)
    call('far3d_pos2posemb3d', _ptr(pos), _ptr(emb), M, num_pos_feats, _stream())
    return emb


def pos2posemb1d(pos, num_pos_feats=256):
    """pos [..., 1] fp32."""
    _chk(pos)
    M = pos.numel() // pos.shape[-1]
    emb = torch.empty(*pos.shape[:-1], num_pos_feats, device=pos.device)
    call('far3d_pos2posemb1d', _ptr(pos), pos.shape[-1], _ptr(emb), M, num_pos_feats, _stream())
    return emb


def nerf_posenc(x, nfreq=6):
    _chk(x)
    Cin = x.shape[-1]
    M = x.numel() // Cin
    emb = torch.empty(*x.shape[:-1], Cin * 2 * nfreq, device=x.device)
    call('far3d_nerf_posenc', _ptr(x), _ptr(emb), M, Cin, nfreq, _stream())
    return emb


def mln_flatten(x, gamma, beta, out, start, channels_last):
    """x [BN,HW,C] (channels_last) or [BN,C,HW]; writes out[:, start:start+HW, :] of out [BN,S,C]."""
    _chk(x); _chk(gamma); _chk(beta); _chk(out)
    BN, S, C = out.shape
    HW = x.shape[1] if channels_last else x.shape[2]
    call('far3d_mln_flatten', _ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), BN, HW, C, S, int(start), int(channels_last),
         _stream())


def mln_tokens(x, gamma, beta, use_ln):
    _chk(x); _chk(gamma); _chk(beta)
    C = x.shape[-1]
    out = torch.empty_like(x)
    call('far3d_mln_tokens', _ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), x.numel() // C, C, int(use_ln), _stream())
    return out


# ------------------------------------------------------------------------------------------ backbone
# lo-plane formats (include/far3d_b200.h): 0 = fp16 residual plane, lo_mx(EA) = e4m3 correction plane ("fp16mx" operands)
LO_FP16 = 0
MX_EA = -1         # activation pre-scale exponent of the e4m3 correction planes: full 4-bit precision for |v| in [2^-6, 448] * 2^-EA


def lo_mx(ea=None):
    return 64 + (MX_EA if ea is None else int(ea))


def conv2d_umma(x_hi, x_lo, N, H, W, x_cs, x_co, Cin, w_hi, w_lo, bias, Cout, ksize, stride, relu,
                y_f32=None, yf_cs=0, yf_co=0, yf_ns=0, y_hi=None, y_lo=None, yb_cs=0, yb_co=0, x_fmt=0, w_exp=0, y_fmt=0):
    """x_fmt / y_fmt: format of the x_lo (and w_lo) / y_lo planes; x_fmt != 0 = fp16mx operands (w_lo is then the weights'
    e4m3 correction plane and w_exp their pre-scale exponent)."""
    pad = ksize // 2
    Ho, Wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    with _Timed('conv_umma', 2.0 * N * Ho * Wo * Cout * Cin * ksize * ksize,      # algorithmic FLOPs (2*MAC)
                f'{N}x{H}x{W} {Cin}->{Cout} k{ksize} s{stride}'):
        if x_fmt == 0 and y_fmt == 0:
            call('far3d_conv2d_umma', _ptr(x_hi), _ptr(x_lo), N, H, W, x_cs, x_co, Cin, _ptr(w_hi), _ptr(w_lo), _ptr(bias), Cout,
                 ksize, stride, int(relu), _ptr(y_f32), yf_cs, yf_co, int(yf_ns), _ptr(y_hi), _ptr(y_lo), yb_cs, yb_co, _stream())
        else:
            call('far3d_conv2d_umma_mx', _ptr(x_hi), _ptr(x_lo), int(x_fmt), N, H, W, x_cs, x_co, Cin, _ptr(w_hi), _ptr(w_lo),
                 int(w_exp), _ptr(bias), Cout, ksize, stride, int(relu), _ptr(y_f32), yf_cs, yf_co, int(yf_ns), _ptr(y_hi),
                 _ptr(y_lo), int(y_fmt), yb_cs, yb_co, _stream())


def conv_pool_workspace_floats(N, H, W, Cout):
    return int(_lib.load().far3d_conv_pool_workspace_floats(int(N), int(H), int(W), int(Cout)))


def conv2d_umma_pool(x_hi, x_lo, N, H, W, x_cs, x_co, Cin, w_hi, w_lo, bias, Cout, relu, y_f32, yf_cs, yf_co, workspace, mean,
                     x_fmt=0, w_exp=0):
    """1x1 conv + global average pool of its fp32 output in one pass (OSA concat conv + eSE pooling)."""
    with _Timed('conv_umma', 2.0 * N * H * W * Cout * Cin, f'{N}x{H}x{W} {Cin}->{Cout} k1 s1 +pool'):
        if x_fmt == 0:
            call('far3d_conv2d_umma_pool', _ptr(x_hi), _ptr(x_lo), N, H, W, x_cs, x_co, Cin, _ptr(w_hi), _ptr(w_lo), _ptr(bias),
                 Cout, int(relu), _ptr(y_f32), yf_cs, yf_co, _ptr(workspace), _ptr(mean), _stream())
        else:
            call('far3d_conv2d_umma_pool_mx', _ptr(x_hi), _ptr(x_lo), int(x_fmt), N, H, W, x_cs, x_co, Cin, _ptr(w_hi), _ptr(w_lo),
                 int(w_exp), _ptr(bias), Cout, int(relu), _ptr(y_f32), yf_cs, yf_co, _ptr(workspace), _ptr(mean), _stream())


# max|w| * 2^w_exp lies in (2^(MX_W_TOP-1), 2^MX_W_TOP]: the e4m3 weight bytes keep full 4-bit precision down to
# max|w| * 2^-(MX_W_TOP + 6), and the pre-scaled fp16 plane tops out at 2^(11 + EA + MX_W_TOP) (2^15 at EA = -1: below fp16's 65504)
MX_W_TOP = int(os.environ.get('FAR3D_MX_W_TOP', '5'))


def pack_weight_mx(wk, ea=None):
    """wk fp32 [Cout, taps, Cin] (Cin % 32 == 0) -> (w_hi fp16, w_c8 e4m3 correction plane viewed as fp16 [Cout, taps, Cin], w_exp).
    Static weights, packed once: torch does the byte shuffling (plumbing); the arithmetic is the header's definition -
    w_hi8 = e4m3(w_hi * 2^w_exp) in the first 32 bytes of every 32-channel group, w_lo8 = e4m3((w - w_hi) * 2^(w_exp + 11)) in
    the second, and the fp16 plane holds w * 2^(11 + EA + w_exp) - the factor both correction products carry - rounded to fp16;
    w_hi = that plane * 2^-(11 + EA + w_exp).  w_exp: see MX_W_TOP (lowered further if the fp16 plane would pass 2^15)."""
    import math
    Cout, taps, Cin = wk.shape
    assert Cin % 32 == 0
    ea = MX_EA if ea is None else int(ea)
    amax = float(wk.abs().max())
    w_exp = 0
    if amax > 0 and math.isfinite(amax):
        lg = int(math.ceil(math.log2(amax)))
        w_exp = min(MX_W_TOP, 15 - (11 + ea)) - lg
    w_exp = max(-40, min(40, w_exp))
    q = 11 + ea + w_exp
    w_hi_s = (wk * (2.0 ** q)).to(torch.float16)              # the plane the kernel reads
    w_hi = w_hi_s.float() * (2.0 ** -q)                       # the value it stands for
    hi8 = (w_hi * (2.0 ** w_exp)).clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    lo8 = ((wk - w_hi) * (2.0 ** (w_exp + 11))).clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)
    c8 = torch.stack([hi8.view(Cout, taps, Cin // 32, 32), lo8.view(Cout, taps, Cin // 32, 32)], dim=3)   # [.., g, 2, 32]
    return w_hi_s, c8.contiguous().view(Cout, taps, Cin * 2).view(torch.float16), w_exp


def conv2d_f32(x, N, H, W, x_cs, x_co, Cin, w, bias, Cout, ksize, stride, relu, y, y_cs, y_co):
    call('far3d_conv2d_f32', _ptr(x), N, H, W, x_cs, x_co, Cin, _ptr(w), _ptr(bias), Cout, ksize, stride, int(relu),
         _ptr(y), y_cs, y_co, _stream())


def stem_conv(img, w, bias, Cout, y_f32=None, y_hi=None, y_lo=None, lo_fmt=0):
    _chk(img)
    N, _, H, W = img.shape
    call('far3d_stem_conv', _ptr(img), N, H, W, _ptr(w), _ptr(bias), Cout, _ptr(y_f32), _ptr(y_hi), _ptr(y_lo), int(lo_fmt),
         _stream())


def normalize_u8(img_u8, mean, std, to_rgb=False, pad_hw=None, out=None):
    """uint8 camera images [N,H,W,3] (or [1,N,H,W,3]) -> normalised, zero-padded fp32 [N,3,Hp,Wp] (same leading dims):
    NormalizeMultiviewImage + AV2PadMultiViewImage + HWC->CHW of the reference's test pipeline, on the device."""
    if not img_u8.is_cuda or img_u8.dtype != torch.uint8 or not img_u8.is_contiguous():
        raise _lib.Far3DNativeError('normalize_u8: img must be a contiguous CUDA uint8 tensor (no CPU path)')
    lead = img_u8.shape[:-3]
    H, W, c3 = img_u8.shape[-3:]
    assert c3 == 3, img_u8.shape
    N = int(np.prod(lead)) if len(lead) else 1
    Hp, Wp = (H, W) if pad_hw is None else (int(pad_hw[0]), int(pad_hw[1]))
    if out is None:
        out = torch.empty(*lead, 3, Hp, Wp, device=img_u8.device, dtype=torch.float32)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == N * 3 * Hp * Wp, out.shape
    m = np.ascontiguousarray(np.asarray(mean, dtype=np.float32)); sd = np.ascontiguousarray(np.asarray(std, dtype=np.float32))
    assert m.shape == (3,) and sd.shape == (3,)
    call('far3d_normalize_u8', _ptr(img_u8), N, H, W, Hp, Wp, m.ctypes.data_as(ctypes.c_void_p), sd.ctypes.data_as(ctypes.c_void_p),
         int(bool(to_rgb)), _ptr(out), _stream())
    return out


def maxpool3x3s2(x_hi, x_lo, dtype, N, H, W, C, x_cs, x_co, y_hi, y_lo, y_cs, y_co, lo_fmt=0):
    call('far3d_maxpool3x3s2', _ptr(x_hi), _ptr(x_lo), dtype, N, H, W, C, x_cs, x_co, _ptr(y_hi), _ptr(y_lo), y_cs, y_co,
         int(lo_fmt), _stream())


def global_avgpool(x, mean, workspace, N, HW, C):
    call('far3d_global_avgpool', _ptr(x), _ptr(mean), _ptr(workspace), N, HW, C, _stream())


def ese_gate(mean, fc_w, fc_b, gate, N, C):
    call('far3d_ese_gate', _ptr(mean), _ptr(fc_w), _ptr(fc_b), _ptr(gate), N, C, _stream())


def ese_apply(xt, gate, id_f32, id_hi, id_lo, id_cs, id_co, N, HW, C, y_f32, yf_cs, yf_co, y_hi, y_lo, yb_cs, yb_co, lo_fmt=0):
    call('far3d_ese_apply', _ptr(xt), _ptr(gate), _ptr(id_f32), _ptr(id_hi), _ptr(id_lo), id_cs, id_co, N, HW, C,
         _ptr(y_f32), yf_cs, yf_co, _ptr(y_hi), _ptr(y_lo), yb_cs, yb_co, int(lo_fmt), _stream())


def upsample_add(dst, src, N, Hd, Wd, Hs, Ws, C, d_hi=None, d_lo=None, lo_fmt=0):
    call('far3d_upsample_add', _ptr(dst), _ptr(src), N, Hd, Wd, Hs, Ws, C, _ptr(d_hi), _ptr(d_lo), int(lo_fmt), _stream())


def groupnorm_nhwc(x, gamma, beta, N, HW, C, groups, eps, relu, y_f32=None, y_hi=None, y_lo=None, lo_fmt=0):
    ws = torch.empty(N * 64 * 2 * groups, device=x.device)
    call('far3d_groupnorm_nhwc', _ptr(x), _ptr(gamma), _ptr(beta), _ptr(ws), N, HW, C, groups, float(eps), int(relu),
         _ptr(y_f32), _ptr(y_hi), _ptr(y_lo), int(lo_fmt), _stream())


def split_fp16(x, want_lo=True):
    _chk(x)
    hi = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    lo = torch.empty_like(hi) if want_lo else None
    call('far3d_split_fp16', _ptr(x), None, _ptr(hi), _ptr(lo), x.numel(), _stream())
    return hi, lo


def merge_fp16(hi, lo=None):
    y = torch.empty(hi.shape, device=hi.device, dtype=torch.float32)
    call('far3d_merge_fp16', _ptr(hi), _ptr(lo), _ptr(y), hi.numel(), _stream())
    return y


def merge_fp16_strided(hi, lo, cs, co, rows, C, lo_fmt=0):
    y = torch.empty(rows, C, device=hi.device, dtype=torch.float32)
    call('far3d_merge_fp16_strided', _ptr(hi), _ptr(lo), int(lo_fmt), cs, co, _ptr(y), rows, C, _stream())
    return y


def split_planes(x, lo_fmt=0, want_lo=True):
    """x fp32 [..., C] dense -> (hi, lo) planes of the same logical shape; lo in the given format (C % 32 == 0 for e4m3)."""
    _chk(x)
    C = x.shape[-1]
    hi = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    lo = torch.empty_like(hi) if want_lo else None
    call('far3d_split_planes', _ptr(x), _ptr(hi), _ptr(lo), int(lo_fmt), x.numel() // C, C, _stream())
    return hi, lo


def conv_umma_tune(bn=0, stages=0):
    _lib.load().far3d_conv_umma_tune(int(bn), int(stages))


def conv_umma_tune4(cta_group=0):
    """experiment knob: 0 = heuristic, 1 = single-CTA kernel only, 2 = CTA-pair (cta_group::2) kernel wherever legal."""
    _lib.load().far3d_conv_umma_tune4(int(cta_group))


def conv_umma_tune2(grid=0, halo=0):
    """experiment knobs: persistent grid size (0 = one CTA per SM) and halo mode switch (-1 = force the generic mode)."""
    _lib.load().far3d_conv_umma_tune2(int(grid), int(halo))


def conv_umma_tune6(loss_per_mma=1.6e-8):
    """tools / tests: expected truncation loss per accumulating tcgen05.mma compensated in the conv epilogue (0 = off)."""
    _lib.load().far3d_conv_umma_tune6(float(loss_per_mma))


def conv_umma_tune7(smem_reserve_bytes=0):
    """bytes of shared memory per SM the persistent conv kernels leave free (co-residency of the other frame's head kernels
    in the two-deep frame pipeline); takes effect for launches (and graph captures) made afterwards"""
    _lib.load().far3d_conv_umma_tune7(int(smem_reserve_bytes))


def conv_umma_tune8(pdl=1):
    """1: conv launches carry cudaLaunchAttributeProgrammaticStreamSerialization (the next conv's prologue - barrier init, TMEM
    allocation, scale-factor fill - runs behind the previous kernel's tail; the kernel waits with griddepcontrol.wait before it
    touches global memory); takes effect for launches (and graph captures) made afterwards"""
    _lib.load().far3d_conv_umma_tune8(int(pdl))
