"""Golden vectors produced by the REFERENCE ITSELF (python tests/golden/make_ref_golden.py; build container only).

The reference's own, unmodified modules under /root/reference/projects/mmdet3d_plugin (VoVNet, Far3D, FarHead,
YOLOXHeadCustom, DepthPredictor, Detr3DTransformer / Decoder / TemporalDecoderLayer, DeformableFeatureAggregationCuda, MLN,
positional encodings, NMSFreeCoder, ...) are imported by file and executed on CPU in fp32.  The third-party packages they
import (mmcv-full 1.6.2, mmdet 2.28.2, mmdet3d) are absent offline and are replaced by the restatements in ref_shims.py
(ConvModule, mmcv MultiheadAttention / FFN, mmdet FPN, MlvlPointGenerator, and mmcv's own pure-PyTorch statement of the
multi-scale deformable attention op).  Inputs are the seeded synthetic frames / tensors of tests/ref_cases.py; weights are
`synthetic.randomize_` applied to the reference modules themselves (same parameter names => same values everywhere).

Outputs (committed): tests/golden/ref_tiny_model.npz, ref_modules.npz, ref_cfg2_frames.npz, ref_cfg3/4/5_frames.npz, ref_preprocess.npz, ref_av2_export.feather, ref_state_dict_full.json.  They pin the oracle
(`-m "not gpu"` tests) and the CUDA path (`-m gpu` tests); /root/reference is not needed to run either.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, 'tests'), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import ref_cases as C  # noqa: E402
import ref_shims as R  # noqa: E402
from far3d_b200 import synthetic  # noqa: E402


def like_oracle(ref_module, oracle_module, seed):
    """Weights: the oracle-side module (same parameter names) is randomised exactly as the tests do it and its state_dict is
    loaded STRICTLY into the reference module - which also proves the names / shapes are the reference's."""
    synthetic.randomize_(oracle_module, seed)
    ref_module.load_state_dict(oracle_module.state_dict(), strict=True)
    ref_module.eval()                              # (the reference's VoVNet.train() returns None, vovnet.py:375)
    return ref_module


def build_reference_detector(model_cfg, seed, prepare=None):
    """the reference's `Far3D` from a model dict, with the weights tests/helpers.build_oracle(model_cfg, seed) has."""
    from helpers import build_oracle
    mc = dict(model_cfg)
    mc['train_cfg'] = None                         # assigners are training-only (out of scope)
    m = R.build_from_cfg(R.to_config(mc), R.DETECTORS).eval()
    o = build_oracle(model_cfg, seed)
    if prepare is not None:
        prepare(o)
    m.load_state_dict(o.state_dict(), strict=True)
    return m


@torch.no_grad()
def tiny_model(mods):
    from helpers import model_cfg
    ref = build_reference_detector(model_cfg(), seed=1)
    cap = {}
    ref.img_backbone.register_forward_hook(lambda m, i, o: cap.__setitem__('backbone', o))
    ref.img_neck.register_forward_hook(lambda m, i, o: cap.__setitem__('fpn', o))
    ref.pts_bbox_head.transformer.register_forward_hook(
        lambda m, i, o: cap.update(feat_flatten=i[2], tgt=i[0], query_pos=i[1], reference_points=i[8], outs_dec=o))
    ref.pts_bbox_head.register_forward_hook(lambda m, i, o: cap.__setitem__('outs', o))
    z = {}
    for f in range(C.TINY_FRAMES):
        metas, data = synthetic.make_frame('tiny', f)
        metas[0]['box_type_3d'] = R.Boxes3D
        res = ref.simple_test(metas, **data)
        outs = cap['outs']
        z[f'cls{f}'] = outs['all_cls_scores'].numpy()
        z[f'box{f}'] = outs['all_bbox_preds'].numpy()
        z[f'ref2d{f}'] = outs['reference_points2d'].numpy()
        z[f'reference_points{f}'] = cap['reference_points'].numpy()
        z[f'outs_dec_last{f}'] = cap['outs_dec'][-1].numpy()
        for k in ('tgt', 'query_pos'):
            z[f'{k}{f}'] = C.sample(cap[k])
            z[f'{k}{f}_norm'] = C.norm(cap[k])
        z[f'feat_flatten{f}'] = C.sample(cap['feat_flatten'])
        z[f'feat_flatten{f}_norm'] = C.norm(cap['feat_flatten'])
        for i, t in enumerate(cap['backbone']):
            z[f'backbone{f}_{i}'], z[f'backbone{f}_{i}_norm'], z[f'backbone{f}_{i}_shape'] = C.sample(t), C.norm(t), np.array(t.shape)
        for i, t in enumerate(cap['fpn']):
            z[f'fpn{f}_{i}'], z[f'fpn{f}_{i}_norm'], z[f'fpn{f}_{i}_shape'] = C.sample(t), C.norm(t), np.array(t.shape)
        b = res[0]['pts_bbox']
        z[f'boxes3d{f}'], z[f'scores3d{f}'], z[f'labels3d{f}'] = b['boxes_3d'].tensor.numpy(), b['scores_3d'].numpy(), b['labels_3d'].numpy()
        print(f'tiny frame {f}: {outs["all_cls_scores"].shape[2]} queries ({outs["reference_points2d"].shape[1]} adaptive), '
              f'{len(b["scores_3d"])} boxes')
    h = ref.pts_bbox_head
    z['memory_embedding'] = h.memory_embedding[0, :C.MEM_ROWS].numpy()
    z['memory_reference_point'] = h.memory_reference_point[0, :C.MEM_ROWS].numpy()
    z['memory_timestamp'] = h.memory_timestamp[0, :C.MEM_ROWS].numpy()
    z['memory_egopose'] = h.memory_egopose[0, :C.MEM_ROWS].numpy()
    z['memory_velo'] = h.memory_velo[0, :C.MEM_ROWS].numpy()
    np.savez_compressed(os.path.join(HERE, 'ref_tiny_model.npz'), **z)


@torch.no_grad()
def modules(mods):
    from oracle import model as O
    z = {}
    pe = mods['models/utils/positional_encoding.py']
    misc = mods['models/utils/misc.py']
    tr = mods['models/utils/detr3d_transformer.py']
    vov = mods['models/backbones/vovnet.py']
    util = mods['core/bbox/util.py']
    coder = mods['core/bbox/coders/nms_free_coder.py']

    # -- positional encoders (positional_encoding.py:13-80)
    x3, x1, xn = C.posenc_inputs()
    z['pos2posemb3d'] = pe.pos2posemb3d(x3).numpy()
    z['pos2posemb1d'] = pe.pos2posemb1d(x1).numpy()
    z['nerf_posenc'] = pe.nerf_positional_encoding(xn).numpy()

    # -- MLN (misc.py:153-190), both flavours used by FarHead
    for name, (c_dim, use_ln) in C.MLN_CASES.items():
        m = like_oracle(misc.MLN(c_dim, use_ln=use_ln), O.MLN(c_dim, use_ln=use_ln), 5)
        x, c = C.mln_inputs(c_dim)
        z[f'mln_{name}'] = m(x, c).numpy()

    # -- misc helpers
    pts, pose = C.transform_inputs()
    z['transform_reference_points'] = misc.transform_reference_points(pts, pose, reverse=False).numpy()
    z['locations'] = misc.locations(torch.zeros(1, 1, 5, 7), 8, 40, 56).numpy()
    z['inverse_sigmoid'] = R.inverse_sigmoid(x3).numpy()

    # -- box decode (nms_free_coder.py:39-112, util.py:25-52)
    cls, box = C.coder_inputs()
    bc = coder.NMSFreeCoder(**C.CODER_CFG)
    d = bc.decode({'all_cls_scores': cls, 'all_bbox_preds': box})[0]
    z['coder_bboxes'], z['coder_scores'], z['coder_labels'] = d['bboxes'].numpy(), d['scores'].numpy(), d['labels'].numpy()
    z['denormalize_bbox'] = util.denormalize_bbox(box[-1, 0], None).numpy()

    # -- VoVNet-99 (vovnet.py), the real 99-layer spec
    bb = like_oracle(vov.VoVNet('V-99-eSE', input_ch=3, out_features=('stage2', 'stage3', 'stage4', 'stage5')),
                     O.VoVNet('V-99-eSE'), 3)
    outs = bb(C.v99_input())
    for i, t in enumerate(outs):
        z[f'v99_{i}'], z[f'v99_{i}_norm'], z[f'v99_{i}_shape'] = C.sample(t), C.norm(t), np.array(t.shape)
    print('V-99 outputs', [tuple(t.shape) for t in outs])

    # -- DeformableFeatureAggregationCuda (detr3d_transformer.py:483-569): module forward + the arguments it hands to MSDA
    m = like_oracle(tr.DeformableFeatureAggregationCuda(**C.DFA_CFG), O.DeformableFeatureAggregationCuda(**C.DFA_CFG), 2)
    a = C.dfa_inputs()
    seen = {}
    orig = tr.MultiScaleDeformableAttnFunction

    class Spy:
        @staticmethod
        def apply(value, shapes, start, loc, w, step):
            seen.update(loc=loc, w=w)
            return orig.apply(value, shapes, start, loc, w, step)
    tr.MultiScaleDeformableAttnFunction = Spy
    try:
        out = m(a['x'], a['query_pos'], a['feat'], a['reference_points'], a['spatial'], a['start'], a['pc_range'], a['lidar2img'],
                a['metas'])
    finally:
        tr.MultiScaleDeformableAttnFunction = orig
    z['dfa_out'] = out.numpy()
    z['dfa_loc'] = seen['loc'][:, :, 0, 0].numpy()               # (N, Nq, P, 2): identical over groups and levels (:555)
    z['dfa_weights'] = C.sample(seen['w'])
    z['dfa_weights_norm'] = C.norm(seen['w'])
    z['dfa_weights_shape'] = np.array(seen['w'].shape)
    kp = tr.get_global_pos(a['reference_points'], a['pc_range']).unsqueeze(-2) + m.learnable_fc(a['x']).reshape(1, -1, C.DFA_CFG['num_pts'], 3)
    z['dfa_key_points'] = kp.numpy()
    # the fused op's operands and result (projection -> MSDA -> camera sum), before output_proj
    feats = m.feature_sampling(a['feat'], a['spatial'], a['start'], kp, seen['w'], a['lidar2img'], a['metas'])
    z['dfa_features'] = feats.numpy()
    inb = ((seen['loc'][:, :, 0, 0] > 0) & (seen['loc'][:, :, 0, 0] < 1)).all(-1)
    print('DFA: in-view fraction of (cam, query, point)', inb.float().mean().item())
    np.savez_compressed(os.path.join(HERE, 'ref_modules.npz'), **z)


@torch.no_grad()
def cfg2_frames():
    """BASELINE.json configs[1] at FULL size through the reference detector: 7 x 960x640, V-99, 644 + 256 + ~150 adaptive
    queries, 6 decoder layers, two streamed frames (the second one reads the memory bank)."""
    from far3d_b200 import api
    mc = api.load_model_cfg(num_cams=7)
    assert mc == R.reference_model_cfg()
    ref = build_reference_detector(mc, seed=0, prepare=lambda o: synthetic.cold_2d_head_(o, C.CFG2_HEAD_SCALE, C.CFG2_HEAD_SCALE))
    cap = {}
    ref.pts_bbox_head.transformer.register_forward_hook(lambda m, i, o: cap.update(feat_flatten=i[2], outs_dec=o))
    ref.pts_bbox_head.register_forward_hook(lambda m, i, o: cap.__setitem__('outs', o))
    z = {}
    for f in range(C.CFG2_FRAMES):
        metas, data = synthetic.make_frame('cfg2', f)
        metas[0]['box_type_3d'] = R.Boxes3D
        res = ref.simple_test(metas, **data)
        outs = cap['outs']
        z[f'cls{f}'] = outs['all_cls_scores'][-1].numpy()
        z[f'box{f}'] = outs['all_bbox_preds'][-1].numpy()
        z[f'cls_all{f}'], z[f'cls_all{f}_norm'] = C.sample(outs['all_cls_scores']), C.norm(outs['all_cls_scores'])
        z[f'box_all{f}'], z[f'box_all{f}_norm'] = C.sample(outs['all_bbox_preds']), C.norm(outs['all_bbox_preds'])
        z[f'ref2d{f}'] = outs['reference_points2d'].numpy()
        z[f'feat_flatten{f}'], z[f'feat_flatten{f}_norm'] = C.sample(cap['feat_flatten']), C.norm(cap['feat_flatten'])
        z[f'outs_dec{f}'], z[f'outs_dec{f}_norm'] = C.sample(cap['outs_dec']), C.norm(cap['outs_dec'])
        b = res[0]['pts_bbox']
        z[f'boxes3d{f}'], z[f'scores3d{f}'], z[f'labels3d{f}'] = b['boxes_3d'].tensor.numpy(), b['scores_3d'].numpy(), b['labels_3d'].numpy()
        print(f'cfg2 frame {f}: {outs["all_cls_scores"].shape[2]} queries ({outs["reference_points2d"].shape[1]} adaptive)')
    np.savez_compressed(os.path.join(HERE, 'ref_cfg2_frames.npz'), **z)


@torch.no_grad()
def full_frames(name):
    """BASELINE.json configs[2] (cfg3: the cfg-2 rig streamed over 8 frames - the memory bank turns over), configs[3] (cfg4:
    6 x 1600x640, nuScenes conventions: 10-wide box code, +-51.2 m) and configs[4] (cfg5: 7 x 1536x1024, 2000 learned queries,
    +-150 m) at FULL size through the reference detector.  Stored per frame: the last decoder layer's class logits and box codes
    (complete), what the detector returns, and seeded samples + norms of the big tensors."""
    case = C.FULL_CASES[name]
    mc = C.full_model_cfg(name)
    ref = build_reference_detector(mc, seed=0, prepare=lambda o: synthetic.cold_2d_head_(o, C.CFG2_HEAD_SCALE, C.CFG2_HEAD_SCALE))
    cap = {}
    ref.pts_bbox_head.transformer.register_forward_hook(lambda m, i, o: cap.update(feat_flatten=i[2], outs_dec=o))
    ref.pts_bbox_head.register_forward_hook(lambda m, i, o: cap.__setitem__('outs', o))
    z = {}
    import time
    for f in range(case['frames']):
        t0 = time.time()
        metas, data = synthetic.make_frame(case['rig'], f)
        metas[0]['box_type_3d'] = R.Boxes3D
        res = ref.simple_test(metas, **data)
        outs = cap['outs']
        z[f'cls{f}'] = outs['all_cls_scores'][-1].numpy().astype(np.float32)
        z[f'box{f}'] = outs['all_bbox_preds'][-1].numpy().astype(np.float32)
        z[f'ref2d{f}'] = outs['reference_points2d'].numpy()
        z[f'feat_flatten{f}'], z[f'feat_flatten{f}_norm'] = C.sample(cap['feat_flatten']), C.norm(cap['feat_flatten'])
        z[f'outs_dec{f}'], z[f'outs_dec{f}_norm'] = C.sample(cap['outs_dec']), C.norm(cap['outs_dec'])
        b = res[0]['pts_bbox']
        z[f'boxes3d{f}'], z[f'scores3d{f}'], z[f'labels3d{f}'] = b['boxes_3d'].tensor.numpy(), b['scores_3d'].numpy(), b['labels_3d'].numpy()
        print(f'{name} frame {f}: {outs["all_cls_scores"].shape[2]} queries ({outs["reference_points2d"].shape[1]} adaptive), '
              f'{len(b["scores_3d"])} boxes, {time.time() - t0:.0f} s', flush=True)
    h = ref.pts_bbox_head
    for k in ('memory_embedding', 'memory_reference_point', 'memory_timestamp', 'memory_egopose', 'memory_velo'):
        t = getattr(h, k)[0, :C.MEM_ROWS]
        z[k] = t.numpy().astype(np.float32) if k != 'memory_embedding' else C.sample(t)
    z['memory_embedding_norm'] = C.norm(h.memory_embedding[0, :C.MEM_ROWS])
    np.savez_compressed(os.path.join(HERE, f'ref_{name}_frames.npz'), **z)


def preprocess():
    """uint8 camera views of three sizes through the reference's own
    NormalizeMultiviewImage (config far3d.py:13-14) and AV2PadMultiViewImage('same2max') classes, stacked CHW as the format
    bundle does.  Also the to_rgb=True variant."""
    pl = R.load_reference_pipelines()
    ns = {}
    with open('/root/reference/projects/configs/far3d.py') as f:
        exec(compile(f.read(), 'far3d.py', 'exec'), ns)
    cfg = ns['img_norm_cfg']
    rng = np.random.default_rng(0)
    views = [rng.integers(0, 256, size=hw + (3,), dtype=np.uint8) for hw in ((40, 64), (48, 64), (40, 56))]   # the lexicographic max shape (custom_pipeline.py:361) must cover every view
    z = {f'view{i}': v for i, v in enumerate(views)}
    z['mean'], z['std'] = np.asarray(cfg['mean'], np.float32), np.asarray(cfg['std'], np.float32)
    for tag, to_rgb in (('', cfg['to_rgb']), ('_rgb', True)):
        res = dict(img=[v.astype(np.float32) for v in views])       # AV2LoadMultiViewImageFromFiles: imread(...).astype(float32)
        res = pl['transform_3d.py'].NormalizeMultiviewImage(mean=cfg['mean'], std=cfg['std'], to_rgb=to_rgb)(res)
        res = pl['custom_pipeline.py'].AV2PadMultiViewImage(size='same2max')(res)
        z['out' + tag] = np.ascontiguousarray(np.stack([i.transpose(2, 0, 1) for i in res['img']], axis=0))
        z['pad_shape' + tag] = np.asarray(res['pad_shape'])
    assert cfg['to_rgb'] is False
    np.savez_compressed(os.path.join(HERE, 'ref_preprocess.npz'), **z)
    print('preprocess:', z['out'].shape, z['pad_shape'].tolist())


RESIZE_CROP_CONF = dict(resize_lim=(0.47, 0.55), final_dim=(64, 96), final_dim_f=(64, 72), bot_pct_lim=(0.0, 0.0), rot_lim=(0.0, 0.0),
                        rand_flip=False)            # far3d.py:167-174 with final_dim scaled 1/10 (small fixture)
RESIZE_CROP_VIEWS = ((155, 205), (205, 155), (155, 205), (160, 200))    # (H, W): landscape, portrait (goes through the transform twice), ...
RESIZE_CROP_SEED = 11


def resize_crop_case():
    """seeded uint8 views + camera matrices of the resize / crop fixture (shared by the generator and the tests)"""
    rng = np.random.default_rng(5)
    views = [rng.integers(0, 256, size=hw + (3,), dtype=np.uint8) for hw in RESIZE_CROP_VIEWS]
    intr = [np.eye(4) + 0.01 * rng.standard_normal((4, 4)) for _ in views]
    for k in intr:
        k[0, 0], k[1, 1], k[0, 2], k[1, 2] = 180.0, 181.0, 100.0, 77.0
    extr = [np.eye(4) + 0.1 * rng.standard_normal((4, 4)) for _ in views]
    return views, intr, extr


def resize_crop():
    """the reference's own AV2ResizeCropFlipRotImageV2 (custom_pipeline.py:48-149: PIL resize / crop per view, portrait views
    twice, ida_mat, intrinsics, lidar2img) on small seeded views; also the flip branch of _img_transform on one view."""
    pl = R.load_reference_pipelines()
    T = pl['custom_pipeline.py'].AV2ResizeCropFlipRotImageV2(data_aug_conf=dict(RESIZE_CROP_CONF))
    views, intr, extr = resize_crop_case()
    np.random.seed(RESIZE_CROP_SEED)
    res = T(dict(img=[v.astype(np.float32) for v in views], intrinsics=[k.copy() for k in intr], extrinsics=[e.copy() for e in extr]))
    z = {}
    for i, im in enumerate(res['img']):
        assert im.dtype == np.float32 and np.array_equal(im, np.round(im)) and im.min() >= 0 and im.max() <= 255
        z[f'img{i}'] = im.astype(np.uint8)
    z['intrinsics'] = np.stack([np.asarray(k, dtype=np.float64) for k in res['intrinsics']])
    z['lidar2img'] = np.stack([np.asarray(k, dtype=np.float64) for k in res['lidar2img']])
    z['ida_mat'] = np.stack([np.asarray(k, dtype=np.float64) for k in res['ida_mat']])
    # _img_transform with a flip and a crop window that leaves the resized image (zero fill)
    from PIL import Image
    img, ida, _ = T._img_transform(Image.fromarray(views[0]), resize=0.5, resize_dims=(102, 77), crop=(-6, 10, 110, 90), flip=True, rotate=0)
    z['flip_img'], z['flip_ida'] = np.array(img), np.asarray(ida, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'ref_resize_crop.npz'), **z)
    print('resize_crop:', [z[f'img{i}'].shape for i in range(len(views))], z['flip_img'].shape)


def av2_export_case():
    """seeded detections of three frames from two logs (inputs of the export fixture / test)"""
    g = torch.Generator().manual_seed(0)
    outs, infos = [], []
    for f in range(3):
        K = 5 + f
        b = torch.randn(K, 7, generator=g) * torch.tensor([50, 50, 2, 1, 1, 1, 3.0])
        b[:, 3:6] = b[:, 3:6].abs() + 0.5
        outs.append(dict(boxes_3d=b, scores_3d=torch.rand(K, generator=g), labels_3d=torch.randint(0, 26, (K,), generator=g)))
        infos.append(dict(scene_id=f'log{f % 2}', lidar_timestamp_ns=315969904359876000 + 100000000 * f))
    return outs, infos


def av2_export():
    """the reference's own Argoverse2Dataset.format_results + box_to_av2 + yaw_to_quat on seeded detections -> the feather file
    the AV2 evaluation reads"""
    ref = R.load_reference_av2_export()
    ds = ref['argoverse2_dataset.py'].Argoverse2Dataset
    outs, infos = av2_export_case()

    class Self:
        data_infos, CLASSES = infos, ds.CLASSES

        def box_to_av2(self, b):
            return ds.box_to_av2(self, b)
    wrapped = [dict(pts_bbox=dict(o, boxes_3d=ref['LiDARInstance3DBoxes'](o['boxes_3d']))) for o in outs]
    dts = ds.format_results(Self(), wrapped)
    dts.reset_index().to_feather(os.path.join(HERE, 'ref_av2_export.feather'))
    assert [c for c in ds.CLASSES] == ref['class_names']
    print('av2 export:', dts.shape)


def state_dict_full():
    """parameter / buffer names and shapes of the reference detector built from ITS OWN config file."""
    mc = R.reference_model_cfg()
    m = build_reference_detector(mc, seed=0)
    sd = m.state_dict()
    json.dump({k: list(v.shape) for k, v in sd.items()}, open(os.path.join(HERE, 'ref_state_dict_full.json'), 'w'), indent=0)
    print('full config:', len(sd), 'state_dict entries,', sum(v.numel() for v in sd.values()) / 1e6, 'M values')


if __name__ == '__main__':
    torch.set_num_threads(8)
    mods = R.load_reference()
    only = sys.argv[1:] or ['modules', 'tiny', 'state_dict', 'cfg2', 'preprocess', 'av2_export', 'resize_crop']
    if 'preprocess' in only:
        preprocess()
    if 'av2_export' in only:
        av2_export()
    if 'resize_crop' in only:
        resize_crop()
    if 'modules' in only:
        modules(mods)
    if 'tiny' in only:
        tiny_model(mods)
    if 'state_dict' in only:
        state_dict_full()
    if 'cfg2' in only:
        cfg2_frames()
    for name in ('cfg3', 'cfg4', 'cfg5'):
        if name in only:
            full_frames(name)
