"""GPU parity tests, module / detector level: the CUDA product modules against the CPU oracle on identical weights and
inputs.  Bar (BASELINE.json): 1e-3 relative fp32 for the parity precision (`fp16x3`); the reduced-precision `fp16` mode
is checked against its own documented budget."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, build_oracle, build_product, model_cfg, rel_err, rel_l2, to_dev, variant_cfg

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope='module')
def tiny(cuda, lib_built):
    mc = model_cfg()
    o = build_oracle(mc, seed=1)
    return mc, o


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-4), ('fp16x3', 1e-3), ('fp16mx', 1e-3), ('fp16', 6e-2)])
def test_backbone_fpn_vs_oracle(tiny, cuda, precision, tol):
    from far3d_b200 import synthetic
    mc, o = tiny
    p = build_product(mc, o.state_dict(), cuda, precision)
    _, data = synthetic.make_frame('tiny', 0)
    img = data['img'][0]
    with torch.no_grad():
        ref_b = o.img_backbone(img)
        ref_f = o.img_neck(ref_b)
        out_b = p.img_backbone(img.to(cuda))
        out_f = p.img_neck(out_b)
    for r, t in zip(ref_b, out_b):
        assert tuple(r.shape) == tuple(t.shape)
        assert rel_l2(t, r) < tol, (precision, rel_l2(t, r))
    for r, t in zip(ref_f, out_f):
        assert tuple(r.shape) == tuple(t.shape)
        assert rel_l2(t, r) < tol and rel_err(t, r) < 5 * tol, (precision, rel_l2(t, r), rel_err(t, r))


def test_v99_backbone_vs_oracle(cuda, lib_built):
    """the real 99-layer backbone (random non-degenerate weights) at 2 x 3 x 192 x 256, parity precision."""
    from far3d_b200 import synthetic
    from far3d_b200.plugin import VoVNet
    from oracle import model as O
    o = O.VoVNet('V-99-eSE').eval()
    synthetic.randomize_(o, 3)
    p = VoVNet('V-99-eSE', out_features=('stage2', 'stage3', 'stage4', 'stage5')).eval()
    p.load_state_dict(o.state_dict()); p.to(cuda)
    x = torch.randn(2, 3, 192, 256, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = o(x)
        out = p(x.to(cuda))
    for r, t in zip(ref, out):
        assert rel_l2(t, r) < TOL and rel_err(t, r) < 5 * TOL, (rel_l2(t, r), rel_err(t, r))


def test_decoder_layer_vs_oracle(tiny, cuda):
    """one Detr3DTemporalDecoderLayer (self-attn, fused aggregation, FFN, 3 LayerNorms) on random tokens/features."""
    from far3d_b200 import synthetic
    mc, o = tiny
    p = build_product(mc, o.state_dict(), cuda)
    lo = o.pts_bbox_head.transformer.decoder.layers[0]
    lp = p.pts_bbox_head.transformer.decoder.layers[0]
    g = torch.Generator().manual_seed(1)
    metas, data = synthetic.make_frame('tiny', 0)
    shapes = [(16, 24), (8, 12), (4, 6), (2, 3)]
    S = sum(h * w for h, w in shapes)
    feat = torch.randn(2, S, 256, generator=g)
    q, qp = torch.randn(1, 90, 256, generator=g), torch.randn(1, 90, 256, generator=g)
    mem, mp = torch.randn(1, 40, 256, generator=g), torch.randn(1, 40, 256, generator=g)
    ref_pts = torch.rand(1, 90, 3, generator=g) * 0.2 + 0.4
    sp = torch.tensor(shapes); st = torch.cat((sp.new_zeros(1), sp.prod(1).cumsum(0)[:-1]))
    pr = o.pts_bbox_head.pc_range
    with torch.no_grad():
        ref = lo(q, qp, feat, mem, mp, ref_pts, sp, st, pr, data['lidar2img'], metas)
        out = lp(q.to(cuda), qp.to(cuda), feat.to(cuda), mem.to(cuda), mp.to(cuda), ref_pts.to(cuda), sp.to(cuda), st.to(cuda),
                 pr.to(cuda), data['lidar2img'].to(cuda), metas)
    assert rel_err(out, ref) < TOL, rel_err(out, ref)


def _rowset_err(a, b):
    """max over rows of a of the distance to the nearest row of b, relative to max|b| (order-invariant comparison)."""
    a, b = a.double().cpu(), b.double().cpu()
    d = torch.cdist(a, b, p=float('inf')).min(dim=1).values
    return (d.max() / b.abs().max()).item()


@pytest.mark.parametrize('precision,tol', [('fp16x3', 1e-3), ('fp16mx', 1e-3), ('fp32', 1e-3)])
def test_detector_two_frames_vs_oracle_and_golden(tiny, cuda, precision, tol):
    """full per-frame path (backbone, FPN, 2D head, adaptive queries, memory bank, 2 decoder layers, box decode) streamed
    over two frames: product == oracle == committed golden outputs."""
    from far3d_b200 import synthetic
    mc, o = tiny
    o.prev_scene_token = None
    p = build_product(mc, o.state_dict(), cuda, precision)
    z = np.load(os.path.join(GOLDEN, 'tiny_model.npz'))
    for f in range(2):
        metas, data = synthetic.make_frame('tiny', f)
        res_o, outs_o = o.simple_test(metas, **data)
        res_p = p.simple_test(metas, **to_dev(data, cuda))
        outs_p = p.last_outs
        assert outs_p['all_cls_scores'].shape == outs_o['all_cls_scores'].shape      # same number of adaptive queries
        assert rel_err(outs_p['feat_flatten'], outs_o['feat_flatten']) < tol
        # The rows of the propagated-query block come from the previous frame's top-256; scores tied within fp32 noise may
        # permute that block (seen on this tiny random-weight model), so rows are matched to their nearest oracle row.
        assert _rowset_err(outs_p['outs_dec'][-1][0], outs_o['outs_dec'][-1][0]) < 2 * tol
        assert _rowset_err(outs_p['all_cls_scores'][-1][0], outs_o['all_cls_scores'][-1][0]) < 2 * tol
        assert _rowset_err(outs_p['all_bbox_preds'][-1][0], outs_o['all_bbox_preds'][-1][0]) < 2 * tol
        assert _rowset_err(outs_p['all_cls_scores'][-1][0], torch.from_numpy(z[f'cls{f}'])[-1][0]) < 2 * tol
        nfix = o.pts_bbox_head.num_query              # learned + adaptive queries keep their order
        assert rel_err(outs_p['outs_dec'][:, :, :nfix], outs_o['outs_dec'][:, :, :nfix]) < 2 * tol
        bo, bp = res_o[0]['pts_bbox'], res_p[0]['pts_bbox']
        assert bo['boxes_3d'].shape == bp['boxes_3d'].shape
        assert rel_err(bp['scores_3d'], bo['scores_3d']) < 2 * tol
        # integer outputs: identical wherever the top-k order is not decided by a sub-tolerance score gap
        so = bo['scores_3d']
        gap = torch.minimum((so[:-1] - so[1:]).abs(), torch.cat([so.new_ones(1), (so[:-2] - so[1:-1]).abs()]))
        clear = torch.cat([gap > 10 * tol * so[:-1].abs(), torch.tensor([False])])
        assert torch.equal(bo['labels_3d'][clear], bp['labels_3d'].cpu()[clear])
        # away from the rank-300 cut the label multiset is identical (ties may permute, never change, the set)
        above = so > so[-1] * (1 + 10 * tol)
        assert torch.equal(bo['labels_3d'][above].sort().values, bp['labels_3d'].cpu()[above].sort().values)
    # memory bank after two frames
    ho, hp = o.pts_bbox_head, p.pts_bbox_head
    # (tiny random-weight model: many scores tie within fp32 noise, so the top-k ORDER may permute; the selected set and
    # the rows stored for each selected query must agree)
    io, ip = ho.last_topk_indexes.flatten(), hp.last_topk_indexes.flatten().cpu()
    common = sorted(set(io.tolist()) & set(ip.tolist()))
    assert len(common) >= 0.95 * io.numel()
    # an index inside the propagated block may name a different (permuted) query on the two sides, so the stored rows are
    # matched to their nearest oracle row instead of position by position
    assert _rowset_err(hp.memory_embedding[0], ho.memory_embedding[0]) < 2 * tol
    assert _rowset_err(hp.memory_reference_point[0], ho.memory_reference_point[0]) < 2 * tol


@pytest.mark.parametrize('kind,ncam,nq', [('nus', 3, 60), ('longrange', 2, 120)])
def test_detector_other_conventions_vs_oracle(cuda, lib_built, kind, ncam, nq):
    """BASELINE.json configs[3] / configs[4] at test size: 10-wide box code with velocity channels and a +-51.2 m range (`nus`),
    +-150 m range and a larger learned-query set (`longrange`); two streamed frames, product against the oracle."""
    from far3d_b200 import synthetic
    mc = variant_cfg(kind, num_cams=ncam, num_query=nq)
    o = build_oracle(mc, seed=2)
    o.prev_scene_token = None
    p = build_product(mc, o.state_dict(), cuda)
    tol = 1e-3
    for f in range(2):
        metas, data = synthetic.make_frame((ncam, 128, 192), f)
        res_o, outs_o = o.simple_test(metas, **data)
        res_p = p.simple_test(metas, **to_dev(data, cuda))
        outs_p = p.last_outs
        assert outs_p['all_cls_scores'].shape == outs_o['all_cls_scores'].shape
        assert outs_p['all_bbox_preds'].shape == outs_o['all_bbox_preds'].shape
        assert outs_p['all_bbox_preds'].shape[-1] == (10 if kind == 'nus' else 8)
        assert rel_err(outs_p['feat_flatten'], outs_o['feat_flatten']) < tol
        assert _rowset_err(outs_p['all_cls_scores'][-1][0], outs_o['all_cls_scores'][-1][0]) < 2 * tol
        assert _rowset_err(outs_p['all_bbox_preds'][-1][0], outs_o['all_bbox_preds'][-1][0]) < 2 * tol
        bo, bp = res_o[0]['pts_bbox'], res_p[0]['pts_bbox']
        assert torch.as_tensor(bo['boxes_3d']).shape == torch.as_tensor(bp['boxes_3d']).shape
        assert torch.as_tensor(bp['boxes_3d']).shape[-1] == (9 if kind == 'nus' else 7)
        assert rel_err(bp['scores_3d'], bo['scores_3d']) < 2 * tol
        assert _rowset_err(torch.as_tensor(bp['boxes_3d']).float(), torch.as_tensor(bo['boxes_3d']).float()) < 2 * tol


def test_detector_without_2d_proposals_vs_oracle(cuda, lib_built):
    """the state SURVEY.md section 8 predicts for random-init weights and the one bench.py runs in: the 2D head at its initial
    operating point lets no peak through, `build_query2d_proposal` returns (None, None) (farhead.py:727-728) and the
    adaptive-query branch is skipped - learned + propagated queries only.  Two streamed frames, product against the oracle."""
    from far3d_b200 import synthetic
    mc = model_cfg()
    o = build_oracle(mc, seed=1)
    synthetic.cold_2d_head_(o)
    o.prev_scene_token = None
    p = build_product(mc, o.state_dict(), cuda)
    h = o.pts_bbox_head
    tol = 1e-3
    for f in range(2):
        metas, data = synthetic.make_frame('tiny', f)
        res_o, outs_o = o.simple_test(metas, **data)
        res_p = p.simple_test(metas, **to_dev(data, cuda))
        outs_p = p.last_outs
        assert outs_o['reference_points2d'] is None and outs_p['reference_points2d'] is None
        assert outs_p['all_cls_scores'].shape == outs_o['all_cls_scores'].shape
        assert outs_p['all_cls_scores'].shape[2] == h.num_query + h.num_propagated
        assert rel_err(outs_p['feat_flatten'], outs_o['feat_flatten']) < tol
        assert _rowset_err(outs_p['all_cls_scores'][-1][0], outs_o['all_cls_scores'][-1][0]) < 2 * tol
        assert _rowset_err(outs_p['all_bbox_preds'][-1][0], outs_o['all_bbox_preds'][-1][0]) < 2 * tol
        assert rel_err(res_p[0]['pts_bbox']['scores_3d'], res_o[0]['pts_bbox']['scores_3d']) < 2 * tol


def test_module_signatures_match_reference(tiny, cuda):
    """drop-in surface: positional forward of the aggregation module with device int64 level tensors (as the reference
    passes them, detr3d_transformer.py:403-413) and `.embed_dims` / `init_weight` attributes."""
    from far3d_b200 import synthetic
    from far3d_b200.plugin import DeformableFeatureAggregationCuda
    from oracle import model as O
    torch.manual_seed(0)
    o = O.DeformableFeatureAggregationCuda(256, 8, 4, 2, 13).eval()
    synthetic.randomize_(o, 2)
    m = DeformableFeatureAggregationCuda(embed_dims=256, num_groups=8, num_levels=4, num_cams=2, dropout=0.1, num_pts=13,
                                         bias=2., batch_first=True).eval()
    assert m.embed_dims == 256 and hasattr(m, 'init_weight')
    m.load_state_dict(o.state_dict()); m.to(cuda)
    metas, data = synthetic.make_frame('tiny', 0)
    shapes = [(16, 24), (8, 12), (4, 6), (2, 3)]
    sp = torch.tensor(shapes); st = torch.cat((sp.new_zeros(1), sp.prod(1).cumsum(0)[:-1]))
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(2, int(sp.prod(1).sum()), 256, generator=g)
    x, qp = torch.randn(1, 33, 256, generator=g), torch.randn(1, 33, 256, generator=g)
    ref_pts = torch.rand(1, 33, 3, generator=g) * 0.2 + 0.4
    pr = torch.tensor([-152.4, -152.4, -5.0, 152.4, 152.4, 5.0])
    with torch.no_grad():
        ref = o(x, qp, feat, ref_pts, sp, st, pr, data['lidar2img'], metas)
        out = m(x.to(cuda), qp.to(cuda), feat.to(cuda), ref_pts.to(cuda), sp.to(cuda), st.to(cuda), pr.to(cuda),
                data['lidar2img'].to(cuda), metas)
    assert rel_err(out, ref) < TOL


def test_image_branch_graph_equals_eager(tiny, cuda):
    """the CUDA-graph replay of backbone + FPN + 2D-head convolutions is bit-identical to the eager launches, stays
    correct when the frame changes, and is rebuilt after load_state_dict."""
    from far3d_b200 import synthetic
    mc, o = tiny
    p = build_product(mc, o.state_dict(), cuda)
    res = {}
    for mode in (False, True, True):
        p.use_cuda_graph = mode
        p.pts_bbox_head.reset_memory() if hasattr(p.pts_bbox_head, 'reset_memory') else None
        p.prev_scene_token = None
        outs = []
        for f in range(2):
            metas, data = synthetic.make_frame('tiny', f)
            p.simple_test(metas, **to_dev(data, cuda))
            outs.append({k: p.last_outs[k].clone() for k in ('feat_flatten', 'all_cls_scores', 'all_bbox_preds')})
        res.setdefault(mode, []).append(outs)
    for other in res[True]:
        for a, b in zip(res[False][0], other):
            for k in a:
                assert torch.equal(a[k], b[k]), k
    assert p.__dict__.get('_img_graphs')
    p.load_state_dict(o.state_dict())
    assert not p.__dict__.get('_img_graphs')


def test_frame_pipeline_equals_sequential(tiny, cuda):
    """Far3DPipeline.stream (image branch of frame i+1 on a second stream while frame i's head runs, two graph instances
    with their own buffers) returns bit-identical results to one-frame-at-a-time inference, across scene changes and with
    the temporal memory bank carried between frames of a scene; device-resident and host (pinned upload) inputs."""
    from far3d_b200 import synthetic
    from far3d_b200.api import Far3DPipeline
    mc, o = tiny
    p = build_product(mc, o.state_dict(), cuda)
    pipe = Far3DPipeline.wrap(p, cuda)
    frames = []
    for f in range(5):
        metas, data = synthetic.make_frame('tiny', f % 3)
        metas = [dict(metas[0], scene_token='sceneA' if f < 3 else 'sceneB')]
        frames.append((metas, data))

    def run_sequential():
        p.prev_scene_token = None
        return [pipe.infer_device(m, **to_dev(d, cuda)) for m, d in frames]

    def flat(res):
        return [r[0]['pts_bbox'][k].float().cpu() if k != 'boxes_3d' else torch.as_tensor(r[0]['pts_bbox'][k]).float().cpu()
                for r in res for k in ('scores_3d', 'labels_3d', 'boxes_3d')]

    ref = flat(run_sequential())
    p.prev_scene_token = None
    out_dev = flat(list(pipe.stream(((m, to_dev(d, cuda)) for m, d in frames))))
    p.prev_scene_token = None
    out_host = flat(list(pipe.stream(frames, host=True, to_host=True)))
    for a, b, c in zip(ref, out_dev, out_host):
        assert torch.equal(a, b) and torch.equal(a, c)
    # at most two frames in flight: a third submit without a collect is refused, not silently overwritten
    pipe.submit(*frames[0][:1], **to_dev(frames[0][1], cuda))
    pipe.submit(*frames[1][:1], **to_dev(frames[1][1], cuda))
    with pytest.raises(RuntimeError, match='two frames are in flight'):
        pipe.submit(*frames[2][:1], **to_dev(frames[2][1], cuda))
    while pipe.pending():
        pipe.collect()


def test_interleaved_streams_keep_their_own_memory_bank(tiny, cuda):
    """two camera-rig streams served by ONE pipeline, frames alternating (far3d_b200.parallel.interleave_streams), each with its
    resident memory bank (`stream_id`): bit-identical to serving each stream alone - sequential and pipelined."""
    from far3d_b200 import synthetic
    from far3d_b200.api import Far3DPipeline
    from far3d_b200.parallel import interleave_streams
    mc, o = tiny
    p = build_product(mc, o.state_dict(), cuda)
    pipe = Far3DPipeline.wrap(p, cuda)

    def frame(stream, f):
        metas, data = synthetic.make_frame('tiny', f, seed=11 if stream == 'A' else 23)
        return [dict(metas[0], scene_token='scene' + stream)], to_dev(data, cuda)

    def flat(r):
        return [torch.as_tensor(r[0]['pts_bbox'][k]).float().cpu() for k in ('scores_3d', 'labels_3d', 'boxes_3d')]

    alone = {}
    for s_, n in (('A', 3), ('B', 2)):
        p.prev_scene_token = None
        p.pts_bbox_head.reset_memory()
        alone[s_] = [flat(pipe.infer_device(*frame(s_, f)[:1], **frame(s_, f)[1])) for f in range(n)]
    sched = interleave_streams({'A': 3, 'B': 2})
    for mode in ('sequential', 'pipelined'):
        for s_ in ('A', 'B'):
            pipe.drop_stream(s_)
        p.prev_scene_token = None
        p.pts_bbox_head.reset_memory()
        if mode == 'sequential':
            got = [flat(pipe.infer_device(*frame(s_, f)[:1], stream_id=s_, **frame(s_, f)[1])) for s_, f in sched]
        else:
            got = [flat(r) for r in pipe.stream([frame(s_, f) + (s_,) for s_, f in sched])]
        for (s_, f), g in zip(sched, got):
            for a, b in zip(alone[s_][f], g):
                assert torch.equal(a, b), (mode, s_, f)


def test_raw_camera_frames_through_the_pipeline(tiny, cuda):
    """the cameras' native uint8 frames in (one portrait view, one landscape view, ~2x the network input), resize / crop /
    normalise on the device inside the two-deep frame pipeline (Far3DPipeline.submit_cameras): identical results to feeding the
    fp32 images and camera matrices that the oracle's restatement of the reference's CPU pipeline produces from the same views
    with the same np.random seed."""
    from far3d_b200 import imgproc, synthetic
    from far3d_b200.api import Far3DPipeline
    from oracle import preprocess as P
    mc, o = tiny
    p = build_product(mc, o.state_dict(), cuda)
    pipe = Far3DPipeline.wrap(p, cuda)
    N, H, W = synthetic.CONFIGS['tiny']
    conf = dict(resize_lim=(0.47, 0.55), final_dim=(H, W), bot_pct_lim=(0.0, 0.0), rot_lim=(0.0, 0.0), rand_flip=False)
    T = imgproc.AV2ResizeCropFlipRotImageV2(data_aug_conf=conf)
    rng = np.random.default_rng(9)
    raw_shapes = [(410, 310), (310, 410)]                       # (H, W): portrait (through the transform twice), landscape
    frames = []
    for f in range(3):
        metas, data = synthetic.make_frame('tiny', f)
        views = [rng.integers(0, 256, size=hw + (3,), dtype=np.uint8) for hw in raw_shapes]
        intr, extr = synthetic.camera_ring(N, 310, 410, np.random.RandomState(f))
        frames.append(([dict(metas[0], scene_token='s')], data, views, intr, extr))
    c = pipe.img_norm_cfg

    def flat(res):
        return [torch.as_tensor(r[0]['pts_bbox'][k]).float().cpu() for r in res for k in ('scores_3d', 'labels_3d', 'boxes_3d')]

    # reference path: oracle pixels + the transform's host plan, fed as fp32 images
    np.random.seed(21)
    ref_frames = []
    for metas, data, views, intr, extr in frames:
        steps, k2, l2i, _ = T.plan([v.shape for v in views], [k.copy() for k in intr], extr)
        imgs = []
        for v, st in zip(views, steps):
            for dims, crop, flip in st:
                v = P.resize_crop_flip_u8(v, dims, crop, flip)
            imgs.append(v)
        f32 = torch.from_numpy(P.normalize_pad_u8(imgs, c['mean'], c['std'], c['to_rgb'], pad_hw=(H, W)))[None]
        to4 = lambda ms: torch.from_numpy(np.stack([np.asarray(m, dtype=np.float64) for m in ms])).float().unsqueeze(0)
        ref_frames.append((metas, dict(data, img=f32, lidar2img=to4(l2i), intrinsics=to4(k2), extrinsics=to4(extr))))
    p.prev_scene_token = None
    ref = flat([pipe.infer_device(m, **to_dev(d, cuda)) for m, d in ref_frames])
    # device path, two frames in flight
    np.random.seed(21)
    p.prev_scene_token = None
    got = []
    small_keys = ('timestamp', 'img_timestamp', 'ego_pose', 'ego_pose_inv')
    for metas, data, views, intr, extr in frames:
        pinned = [torch.from_numpy(v).pin_memory() for v in views]
        pipe.submit_cameras(metas, pinned, intr, extr, T, **{k: data[k] for k in small_keys})
        if pipe.pending() > 1:
            got.append(pipe.collect(to_host=True))
    while pipe.pending():
        got.append(pipe.collect(to_host=True))
    got = flat(got)
    assert len(ref) == len(got)
    for r, x in zip(ref, got):
        assert torch.equal(r, x)
    assert pipe.last_h2d_bytes >= sum(v.size for v in frames[0][2])


def test_uint8_frames_through_the_pipeline(tiny, cuda):
    """camera bytes in (uint8 HWC, 1 byte per sample over PCIe), normalised on the device: identical results to feeding the
    fp32 images the reference's CPU pipeline would have produced (oracle/preprocess.py restates it), on all three entry points."""
    from far3d_b200 import synthetic
    from far3d_b200.api import Far3DPipeline
    from oracle import preprocess as P
    mc, o = tiny
    p = build_product(mc, o.state_dict(), cuda)
    pipe = Far3DPipeline.wrap(p, cuda)
    g = torch.Generator().manual_seed(5)
    frames_u8, frames_f32 = [], []
    for f in range(3):
        metas, data = synthetic.make_frame('tiny', f)
        N, _, H, W = data['img'].shape[1:]
        u8 = torch.randint(0, 256, (1, N, H, W, 3), generator=g, dtype=torch.uint8)
        c = pipe.img_norm_cfg
        f32 = torch.from_numpy(P.normalize_pad_u8(list(u8[0].numpy()), c['mean'], c['std'], c['to_rgb'], pad_hw=(H, W)))[None]
        metas = [dict(metas[0], scene_token='s')]
        frames_u8.append((metas, dict(data, img=u8)))
        frames_f32.append((metas, dict(data, img=f32)))

    def flat(res):
        return [torch.as_tensor(r[0]['pts_bbox'][k]).float().cpu() for r in res for k in ('scores_3d', 'labels_3d', 'boxes_3d')]

    p.prev_scene_token = None
    ref = flat([pipe.infer_device(m, **to_dev(d, cuda)) for m, d in frames_f32])
    p.prev_scene_token = None
    a = flat([pipe.infer_device(m, **to_dev(d, cuda)) for m, d in frames_u8])
    p.prev_scene_token = None
    b = flat(list(pipe.stream(frames_u8, host=True, to_host=True)))
    p.prev_scene_token = None
    c_ = flat([pipe.infer(m, **d) for m, d in frames_u8])
    for r, x, y, z in zip(ref, a, b, c_):
        assert torch.equal(r, x) and torch.equal(r, y) and torch.equal(r, z)
    assert pipe.last_h2d_bytes < 0.3 * sum(v.numel() * v.element_size() for v in frames_f32[0][1].values() if torch.is_tensor(v))


@pytest.mark.parametrize('precision', ['fp16x3', 'fp16mx'])
def test_device_proposals_match_torch_glue(tiny, cuda, precision):
    """the sync-free proposal kernels (far3d_roi_select / far3d_query2d_lift / far3d_ctx_gather) + the
    bucketed, key-masked decoder against the boolean-gather torch glue they replace (which the reference fixtures pin): same
    adaptive queries in the same order, same logits / boxes, over two streamed frames; the padded rows never leak."""
    from far3d_b200 import synthetic
    mc, o = tiny
    outs = {}
    for kernels in (False, True):
        p = build_product(mc, o.state_dict(), cuda, precision)
        p.pts_bbox_head.proposal_kernels = kernels
        res = []
        for f in range(2):
            metas, data = synthetic.make_frame('tiny', f)
            r = p.simple_test(metas, **to_dev(data, cuda))
            lo = p.last_outs
            res.append((lo['reference_points2d'].float().cpu(), lo['all_cls_scores'].float().cpu(), lo['all_bbox_preds'].float().cpu(),
                        torch.as_tensor(r[0]['pts_bbox']['scores_3d']).float().cpu()))
        outs[kernels] = res
        if kernels:
            h = p.pts_bbox_head
            assert h.proposal_bucket == 64 and '_prop_host' in h.__dict__
            keys = list(h.__dict__.get('_graphs', {}))
            assert keys and all(k[0][1] % 1 == 0 for k in keys)
            m = res[0][0].shape[1]
            assert any(k[0][1] == h.num_query + h.num_propagated + (-(-m // 64) * 64) for k in keys), (m, [k[0] for k in keys])
    def matched(x, y):
        """largest distance from a row of one set to its nearest row in the other (both directions): the propagated queries of a
        streamed frame are a top-k of the previous frame's scores, and two near-tied scores may swap places between the two forms
        (their arithmetic differs in the last bits: e.g. a linear over 41 real rows runs a different kernel than over 64 padded)"""
        d = (x[:, None, :] - y[None, :, :]).abs().amax(-1)
        return float(max(d.min(1).values.max(), d.min(0).values.max())) / float(y.abs().max())

    for f in range(2):
        a, b = outs[False][f], outs[True][f]
        assert a[0].shape == b[0].shape and a[0].shape[1] > 20, a[0].shape            # adaptive queries present, same count
        assert rel_err(b[0], a[0]) < 1e-5                                             # same points, same order
        assert a[1].shape == b[1].shape and a[2].shape == b[2].shape
        if f == 0:
            assert rel_err(b[1], a[1]) < 2e-4 and rel_err(b[2], a[2]) < 2e-4
        else:
            for l in range(a[1].shape[0]):
                assert matched(b[1][l, 0], a[1][l, 0]) < 2e-4 and matched(b[2][l, 0], a[2][l, 0]) < 2e-4, (f, l)
        assert rel_err(b[3].sort().values, a[3].sort().values) < 2e-4


def test_memory_bank_kernels_match_torch_glue(tiny, cuda):
    """far3d_memory_pre_update / far3d_memory_post_update (bitonic top-k of the propagated queries, 4x4 ego-pose products, fp64
    timestamps, in-place bank) against the torch statement of farhead.py:330-420 they replace, over three streamed frames with
    a scene cut in the middle: same top-k indices, same bank, same detector outputs."""
    from far3d_b200 import synthetic
    mc, o = tiny
    outs = {}
    for kernels in (False, True):
        p = build_product(mc, o.state_dict(), cuda, 'fp16x3')
        h = p.pts_bbox_head
        h.memory_kernels = kernels
        res = []
        for f, scene in ((0, 'a'), (1, 'a'), (0, 'b'), (1, 'b')):
            metas, data = synthetic.make_frame('tiny', f, scene=scene)
            p.simple_test(metas, **to_dev(data, cuda))
            res.append((h.last_topk_indexes.flatten().cpu(), h.memory_embedding.float().cpu(), h.memory_reference_point.float().cpu(),
                        h.memory_timestamp.double().cpu(), h.memory_egopose.float().cpu(), h.memory_velo.float().cpu(),
                        p.last_outs['all_cls_scores'].float().cpu()))
        outs[kernels] = res
    for a, b in zip(outs[False], outs[True]):
        assert torch.equal(a[0], b[0])                                                 # the propagated queries, in order
        for x, y in zip(a[1:], b[1:]):
            assert x.shape == y.shape and rel_err(y, x) < 1e-5, (x.shape, rel_err(y, x))


def test_decoder_hoisted_invariants_equal_per_layer_form(tiny, cuda):
    """Detr3DTransformerDecoder computes the layer-invariant operands once (key position rows, metric reference rows, the camera
    logits of all layers through far3d_cam_logits): same detector outputs as the per-layer form over two streamed frames."""
    from far3d_b200 import synthetic
    mc, o = tiny
    outs = {}
    for hoist in (False, True):
        p = build_product(mc, o.state_dict(), cuda, 'fp16x3')
        p.pts_bbox_head.transformer.decoder.hoist = hoist
        res = []
        for f in range(2):
            metas, data = synthetic.make_frame('tiny', f)
            p.simple_test(metas, **to_dev(data, cuda))
            res.append((p.last_outs['all_cls_scores'].float().cpu(), p.last_outs['all_bbox_preds'].float().cpu()))
        outs[hoist] = res
    for a, b in zip(outs[False], outs[True]):
        assert rel_err(b[0], a[0]) < 2e-5 and rel_err(b[1], a[1]) < 2e-5, (rel_err(b[0], a[0]), rel_err(b[1], a[1]))
