"""Multi-GPU plan for the per-frame path: one process per GPU, `torch.distributed` (NCCL over
NVLink 5 / NVSwitch on the box, gloo in CPU tests).

Two ways the path shards, both used by bench.py:
  * streams / frames are independent (each owns a memory bank) - the reference's own test-time sharding
    (datasets/samplers/distributed_sampler.py:41-44): ranks run whole frames, no data-path collective ("replicas").
    This is the throughput mode (`bench.py --shard streams`, the default).
  * cameras are independent through backbone -> FPN -> 2D-head convolutions -> MLN/flatten (everything before
    farhead.py:565); the decoder is cross-camera (softmax over cameras x levels x points, detr3d_transformer.py:540).
    `CameraShardedFar3D` runs the image branch on a contiguous camera slice per rank, then ONE all-gather of the flattened
    channels-last feature maps (`feat_flatten`, 13.06 MB per camera at cfg-2) plus one small all-gather of the per-camera
    dense 2D-head maps (3.5 MB per camera), and the decoder runs replicated: every rank ends the frame with identical
    boxes and an identical memory bank.  This is the latency mode for ONE camera rig (`bench.py --shard cameras`).
"""
import torch
import torch.distributed as dist


def shard_frames(num_frames, world_size, rank):
    """The frame indices rank `rank` evaluates: the reference's test-time DistributedSampler with shuffle=False
    (datasets/samplers/distributed_sampler.py:29-48) - the index list is repeated up to a multiple of the world size and cut
    into CONTIGUOUS chunks, one per rank, so that a rank walks whole stretches of a camera-rig stream in order and its
    temporal memory bank stays meaningful (a strided split would interleave the streams frame by frame)."""
    if num_frames <= 0:
        return []
    per = -(-num_frames // world_size)
    total = per * world_size
    idx = (list(range(num_frames)) * -(-total // num_frames))[:total]
    return idx[rank * per:(rank + 1) * per]


def assign_streams(stream_lengths, world_size):
    """Whole camera-rig streams (scenes) to ranks for the serving case, where - unlike the sampler above - a stream is never
    cut: longest stream first onto the least loaded rank (LPT), ties to the lower rank.  Returns (per-rank lists of stream
    ids in the order they were assigned, per-rank frame counts).  No rank idles while there are at least `world_size`
    streams; with fewer streams than ranks the rest of the box is better used by `CameraShardedFar3D`."""
    order = sorted(range(len(stream_lengths)), key=lambda i: (-stream_lengths[i], i))
    ranks, load = [[] for _ in range(world_size)], [0] * world_size
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        ranks[r].append(i)
        load[r] += stream_lengths[i]
    return ranks, load


def interleave_streams(streams):
    """Round-robin schedule of one rank's streams: [(stream_id, frame_index)] taking one frame of every unfinished stream per
    turn (what `Far3DPipeline(..., stream_id=)` is for: each stream keeps its own resident memory bank, so frames of
    different rigs can alternate and the two-deep frame pipeline never drains at a scene boundary).
    `streams`: {stream_id: number of frames}."""
    left = {k: 0 for k in streams}
    out = []
    while left:
        for k in list(left):
            out.append((k, left[k]))
            left[k] += 1
            if left[k] >= streams[k]:
                del left[k]
    return out


def shard_cameras(num_cams, world_size):
    """Contiguous [begin, end) camera ranges per rank; ranks beyond the camera count get empty ranges."""
    per = -(-num_cams // world_size)
    return [(min(r * per, num_cams), min((r + 1) * per, num_cams)) for r in range(world_size)]


def all_gather_cameras(local, num_cams, group=None, send=None, recv=None):
    """local [n_local, ...] (this rank's cameras, may be empty) -> [num_cams, ...] on every rank, a contiguous prefix view of
    the receive buffer.  One collective; ragged shards are padded to the largest shard so a single fixed-size all-gather
    suffices (rank r owns rows [r*per, r*per + n_local), so the cameras come out in order and only the tail is padding).
    `send` [per, ...] / `recv` [world*per, ...]: optional persistent buffers (stable addresses for captured graphs)."""
    world = dist.get_world_size(group)
    per = -(-num_cams // world)
    tail = tuple(local.shape[1:])
    if send is None:
        send = local if local.shape[0] == per and local.is_contiguous() else local.new_zeros((per,) + tail)
    if send.data_ptr() != local.data_ptr() and local.shape[0]:        # (a caller may have filled `send` in place already)
        send[:local.shape[0]].copy_(local)
    if recv is None:
        recv = local.new_empty((world * per,) + tail)
    if dist.get_backend(group) == 'nccl':
        dist.all_gather_into_tensor(recv, send, group=group)
    else:                                   # gloo (CPU tests; CUDA tensors are staged through the host - tests only)
        s = send.cpu() if send.is_cuda else send
        chunks = [torch.empty_like(s) for _ in range(world)]
        dist.all_gather(chunks, s, group=group)
        recv.view((world, per) + tail).copy_(torch.stack(chunks))
    return recv[:num_cams]


def level_shapes(H, W, strides):
    """feature-map sizes of the FPN levels for a padded H x W image (every down-sampling step on the path rounds up)."""
    return [(-(-H // s), -(-W // s)) for s in strides]


def roi_layout(shapes, cls_channels, reg_channels, depth_channels, depth_level):
    """Packing plan of one camera's dense 2D-head outputs as one fp32 row, in the predictors' own NHWC layouts (channel counts
    are the padded buffer widths: 28 class channels for 26 classes, 8 for box + objectness + centre offsets, 52 depth logits):
    [(key, level, H, W, C, offset)], row length.  NHWC so that the gathered maps feed far3d_roi_select directly."""
    plan, off = [], 0
    for l, (h, w) in enumerate(shapes):
        for key, c in (('cls', cls_channels), ('reg', reg_channels)):
            plan.append((key, l, h, w, c, off))
            off += c * h * w
    if depth_channels:
        h, w = shapes[depth_level]
        plan.append(('depth', depth_level, h, w, depth_channels, off))
        off += depth_channels * h * w
    return plan, off


def pack_roi(roi, plan, row):
    """roi: the 2D head's `out` dict of this rank (NHWC buffers `_cls_nhwc`, `_reg_nhwc`, `_depth_logit_nhwc`, leading dim
    n_local) -> `row` [>= n_local, total] filled in place."""
    for key, l, h, w, c, off in plan:
        t = roi['_depth_logit_nhwc'] if key == 'depth' else roi['_cls_nhwc' if key == 'cls' else '_reg_nhwc'][l]
        n = t.shape[0]
        assert tuple(t.shape[1:]) == (h, w, c), (key, l, tuple(t.shape), (h, w, c))
        row[:n, off:off + c * h * w].view(n, h, w, c).copy_(t)
    return row


def unpack_roi(rows, plan, num_classes, depth_bins=0):
    """rows [N, total] -> the 2D head's `out` dict over all N cameras: dense NHWC copies (what the proposal kernels read) and
    the reference-layout NCHW views of them (yolox_head.py:260-341)."""
    n = rows.shape[0]
    out = dict(enc_cls_scores=[], enc_bbox_preds=[], objectnesses=[], pred_centers2d_offset=[], topk_indexes=None,
               _cls_nhwc=[], _reg_nhwc=[])
    for key, l, h, w, c, off in plan:
        v = rows[:, off:off + c * h * w].reshape(n, h, w, c).contiguous()       # dense [N,H,W,C] (a row holds other maps too)
        nchw = v.permute(0, 3, 1, 2)
        if key == 'cls':
            out['_cls_nhwc'].append(v)
            out['enc_cls_scores'].append(nchw[:, :num_classes])
        elif key == 'reg':
            out['_reg_nhwc'].append(v)
            out['enc_bbox_preds'].append(nchw[:, 0:4]); out['objectnesses'].append(nchw[:, 4:5])
            out['pred_centers2d_offset'].append(nchw[:, 5:7])
        else:
            logit = nchw[:, :depth_bins]
            out.update(_depth_logit_nhwc=v, _depth_bins=depth_bins, depth_logit=logit, pred_depth=logit.softmax(dim=1))
    return out


class CameraShardedFar3D:
    """One camera-rig stream over all ranks of `group`: image branch on this rank's camera slice, all-gather, replicated
    decoder.  Same call surface as `Far3D.simple_test`; every rank passes the same frame (only its camera slice of `img` is
    read) and returns the same result."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._bufs = {}
        self.last_gather_bytes = 0

    def _buf(self, name, shape, like):
        b = self._bufs.get(name)
        if b is None or tuple(b.shape) != tuple(shape) or b.device != like.device:
            b = self._bufs[name] = torch.zeros(shape, device=like.device, dtype=torch.float32)
        return b

    def camera_range(self, num_cams):
        return shard_cameras(num_cams, self.world)[self.rank]

    @torch.no_grad()
    def simple_test(self, img_metas, **data):
        m = self.model
        head, roi_head = m.pts_bbox_head, m.img_roi_head
        img = data['img']
        assert img.dim() == 5 and img.shape[0] == 1, 'one sample per frame (farhead.py:813-816 raises for B > 1 too)'
        N = data['lidar2img'].shape[1]
        a, b = self.camera_range(N)
        # `img` may hold all N cameras or only this rank's slice (a host caller uploads just the slice)
        img_local = img[:, a:b] if img.shape[1] == N else img
        assert img_local.shape[1] == b - a, (img.shape, a, b)
        per = -(-N // self.world)
        pad_h, pad_w = img_metas[0]['pad_shape'][0][:2]
        shapes = level_shapes(pad_h, pad_w, m.stride)
        S, C = sum(h * w for h, w in shapes), head.embed_dims
        m._mark('start')
        roi = None
        send = self._buf('feat_send', (per, S, C), img)
        if b > a:
            feats, roi = m.image_branch(img_local.contiguous())
            assert [tuple(f.shape[-2:]) for f in feats] == shapes, ([tuple(f.shape[-2:]) for f in feats], shapes)
            if roi is None and m.with_img_roi_head:
                roi = roi_head(None, img_feats=feats)
            local = dict(intrinsics=data['intrinsics'][:, a:b], extrinsics=data['extrinsics'][:, a:b])
            flat, _, _ = head.flatten_features(feats, local)                 # [n_local, S, C], MLN applied per camera
            send[:b - a].copy_(flat)
        # ---- the exchange: per-view feature maps (+ the small dense 2D-head maps)
        feat_flatten = all_gather_cameras(send[:b - a], N, self.group, send=send,
                                          recv=self._buf('feat_recv', (self.world * per, S, C), img))
        self.last_gather_bytes = send.numel() * 4 * self.world
        outs_roi = None
        if m.with_img_roi_head:
            nd = (int(roi_head.depthnet_config['num_depth_bins']) + 1) if roi_head.pred_with_depth else 0
            ridx = ['p3', 'p4', 'p5'].index(roi_head.reg_depth_level) if nd else 0
            pad4 = lambda c: (c + 3) // 4 * 4                  # the predictors' buffer widths (roi_head.py forward)
            plan, total = roi_layout(shapes, pad4(roi_head.num_classes), 8, pad4(nd), ridx)
            rsend = self._buf('roi_send', (per, total), img)
            if roi is not None:
                pack_roi(roi, plan, rsend)
            rows = all_gather_cameras(rsend[:b - a], N, self.group, send=rsend,
                                      recv=self._buf('roi_recv', (self.world * per, total), img))
            self.last_gather_bytes += rsend.numel() * 4 * self.world
            outs_roi = unpack_roi(rows, plan, roi_head.num_classes, nd)
        m._mark('image_branch_and_gather')
        dev = feat_flatten.device
        starts, s = [], 0
        for h, w in shapes:
            starts.append(s)
            s += h * w
        head._levels_host = (tuple(shapes), tuple(starts))
        pre = (feat_flatten, torch.as_tensor(shapes, dtype=torch.long, device=dev),
               torch.as_tensor(starts, dtype=torch.long, device=dev))
        data = dict(data, img=img_local, img_feats=None, _feat_flatten=pre, _outs_roi_dense=outs_roi)
        bbox_pts, _ = m.simple_test_pts(img_metas, **data)
        return [dict(pts_bbox=p) for p in bbox_pts]

    @torch.no_grad()
    def infer(self, img_metas, **host_data):
        """host entry: uploads ONLY this rank's camera slice of `img` (plus the small per-frame tensors), runs the frame and
        reads the boxes back.  Returns (result on host, h2d bytes, d2h bytes)."""
        dev = self.model.pts_bbox_head.pc_range.device
        N = host_data['lidar2img'].shape[1]
        a, b = self.camera_range(N)
        data, h2d = {}, 0
        for k, v in host_data.items():
            if not torch.is_tensor(v):
                data[k] = v
                continue
            if k == 'img':
                v = v[:, a:b]
            if not v.is_pinned():
                p = self._bufs.get('pin_' + k)
                if p is None or p.shape != v.shape or p.dtype != v.dtype:
                    p = self._bufs['pin_' + k] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                p.copy_(v)
                v = p
            data[k] = v.to(dev, non_blocking=True)
            h2d += v.numel() * v.element_size()
        res = self.simple_test(img_metas, **data)
        out, d2h = [], 0
        for r in res:
            cpu = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in r['pts_bbox'].items()}
            d2h += sum(v.numel() * v.element_size() for v in cpu.values() if torch.is_tensor(v))
            out.append(dict(pts_bbox=cpu))
        return out, h2d, d2h
