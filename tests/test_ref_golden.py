"""The oracle against golden vectors produced by the REFERENCE'S OWN modules (tests/golden/make_ref_golden.py ran the
unmodified files of /root/reference/projects/mmdet3d_plugin on CPU, third-party mmcv/mmdet pieces restated in
tests/golden/ref_shims.py).  These are the tests that pin the oracle; the `-m gpu` twins in
test_gpu_ref_golden.py hold the CUDA path to the same vectors.  Nothing here needs /root/reference, except the last test,
which re-runs the reference live when it is present and checks that the committed fixtures are what it produces."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import ref_cases as C
from helpers import GOLDEN, ROOT, build_oracle, model_cfg, rel_err, rowset_err

from oracle import cref
from oracle import model as O

TOL = 2e-5        # oracle vs reference, both CPU fp32: only summation-order noise is allowed


@pytest.fixture(scope='module')
def zm():
    return np.load(os.path.join(GOLDEN, 'ref_modules.npz'))


@pytest.fixture(scope='module')
def zt():
    return np.load(os.path.join(GOLDEN, 'ref_tiny_model.npz'))


def close(a, b, tol=TOL):
    a, b = torch.as_tensor(np.asarray(a)), torch.as_tensor(np.asarray(b))
    assert a.shape == b.shape, (a.shape, b.shape)
    e = rel_err(a, b)
    assert e < tol, e


def close_sampled(t, z, key, tol=TOL):
    """large tensors are stored as values at seeded positions + the L2 norm of the whole tensor."""
    if key + '_shape' in z.files:
        assert tuple(t.shape) == tuple(z[key + '_shape'])
    close(C.sample(t), z[key], tol)
    assert abs(C.norm(t) / float(z[key + '_norm']) - 1) < tol


def test_positional_encoders(zm):
    x3, x1, xn = C.posenc_inputs()
    close(O.pos2posemb3d(x3), zm['pos2posemb3d'])
    close(O.pos2posemb1d(x1), zm['pos2posemb1d'])
    close(O.nerf_positional_encoding(xn), zm['nerf_posenc'])
    close(O.inverse_sigmoid(x3), zm['inverse_sigmoid'])


def test_mln_and_helpers(zm):
    from far3d_b200 import synthetic
    for name, (c_dim, use_ln) in C.MLN_CASES.items():
        m = O.MLN(c_dim, use_ln=use_ln).eval()
        synthetic.randomize_(m, 5)
        x, c = C.mln_inputs(c_dim)
        with torch.no_grad():
            close(m(x, c), zm[f'mln_{name}'])
    pts, pose = C.transform_inputs()
    close(O.transform_reference_points(pts, pose), zm['transform_reference_points'])


def test_box_coder(zm):
    cls, box = C.coder_inputs()
    d = O.NMSFreeCoder(**C.CODER_CFG).decode({'all_cls_scores': cls, 'all_bbox_preds': box})[0]
    close(d['bboxes'], zm['coder_bboxes'])
    close(d['scores'], zm['coder_scores'])
    assert np.array_equal(d['labels'].numpy(), zm['coder_labels'])
    close(O.denormalize_bbox(box[-1, 0]), zm['denormalize_bbox'])


def test_vovnet99(zm):
    from far3d_b200 import synthetic
    o = O.VoVNet('V-99-eSE').eval()
    synthetic.randomize_(o, 3)
    with torch.no_grad():
        outs = o(C.v99_input())
    for i, t in enumerate(outs):
        close_sampled(t, zm, f'v99_{i}', 1e-4)


def test_deformable_aggregation_module(zm):
    """module forward, key points, sampling locations, softmax weights - and the C oracle of the FUSED op (projection ->
    bilinear gather -> camera sum) against the reference's feature_sampling() output."""
    from far3d_b200 import synthetic
    m = O.DeformableFeatureAggregationCuda(**C.DFA_CFG).eval()
    synthetic.randomize_(m, 2)
    a = C.dfa_inputs()
    with torch.no_grad():
        out = m(a['x'], a['query_pos'], a['feat'], a['reference_points'], a['spatial'], a['start'], a['pc_range'], a['lidar2img'],
                a['metas'])
        kp = m.key_points(a['x'], a['reference_points'], a['pc_range'])
        w = m.weights(a['x'], a['query_pos'], a['lidar2img'])
        loc = m.sampling_locations(kp, a['lidar2img'], C.DFA_PAD_HW)
    close(out, zm['dfa_out'], 1e-4)
    close(kp, zm['dfa_key_points'])
    close_sampled(w, zm, 'dfa_weights', 1e-4)
    # points far behind a camera are divided by clamp(z, 1e-5) (:550) and blow up to ~1e7: compare where it matters
    ref_loc = torch.from_numpy(zm['dfa_loc'])
    mine = loc[:, :, 0, 0]
    near = ref_loc.abs().amax(-1) < 4
    assert near.float().mean() > 0.2
    assert (mine[near] - ref_loc[near]).abs().max() < 1e-5
    assert torch.equal(mine.abs().amax(-1) < 4, near)
    # fused C oracle on the reference's own operands
    feats, uv, idx, valid = cref.deform_agg(a['feat'].numpy(), np.array(C.DFA_SHAPES), a['start'].numpy(), zm['dfa_key_points'],
                                            a['lidar2img'].numpy(), w.numpy(), *C.DFA_PAD_HW, C.DFA_CFG['num_groups'], debug=True)
    close(feats, zm['dfa_features'], 1e-4)
    # in-bounds mask of the C oracle == the reference's sampling locations pushed through mmcv's bounds rule, away from
    # the borders' rounding neighbourhood
    u, v = ref_loc[..., 0], ref_loc[..., 1]                     # (N, Nq, P)
    for l, (H, W) in enumerate(C.DFA_SHAPES):
        h_im, w_im = v * H - 0.5, u * W - 0.5
        inb = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        edge = ((h_im + 1).abs() < 1e-3) | ((w_im + 1).abs() < 1e-3) | ((h_im - H).abs() < 1e-3) | ((w_im - W).abs() < 1e-3)
        got = torch.from_numpy(valid.reshape(1, C.DFA_CFG['num_cams'], C.DFA_NQ, len(C.DFA_SHAPES), -1)[0, :, :, l].astype(bool))
        assert torch.equal(got[~edge], inb[~edge])


def test_detector_two_frames(zt):
    """the whole per-frame path, two streamed frames (2D head with adaptive queries, memory bank, box decode)."""
    from far3d_b200 import synthetic
    o = build_oracle(model_cfg(), seed=1)
    for f in range(C.TINY_FRAMES):
        metas, data = synthetic.make_frame('tiny', f)
        with torch.no_grad():
            feats = o.img_neck(o.img_backbone(data['img'][0]))
            bb = o.img_backbone(data['img'][0])
        for i, t in enumerate(bb):
            close_sampled(t, zt, f'backbone{f}_{i}', 1e-4)
        for i, t in enumerate(feats):
            close_sampled(t, zt, f'fpn{f}_{i}', 1e-4)
        res, outs = o.simple_test(metas, **data)
        assert outs['all_cls_scores'].shape == zt[f'cls{f}'].shape            # same number of adaptive queries
        close(outs['reference_points2d'], zt[f'ref2d{f}'], 1e-4)
        close_sampled(outs['feat_flatten'], zt, f'feat_flatten{f}', 1e-4)
        nfix = o.pts_bbox_head.num_query + outs['reference_points2d'].shape[1]   # learned + adaptive keep their order
        close(outs['all_cls_scores'][:, :, :nfix], zt[f'cls{f}'][:, :, :nfix], 2e-4)
        close(outs['all_bbox_preds'][:, :, :nfix], zt[f'box{f}'][:, :, :nfix], 2e-4)
        # propagated block: rows come from the previous frame's top-k, whose order may permute among tied scores
        assert rowset_err(outs['all_cls_scores'][-1][0], torch.from_numpy(zt[f'cls{f}'])[-1][0]) < 2e-4
        assert rowset_err(outs['all_bbox_preds'][-1][0], torch.from_numpy(zt[f'box{f}'])[-1][0]) < 2e-4
        assert rowset_err(outs['outs_dec'][-1][0], torch.from_numpy(zt[f'outs_dec_last{f}'])[0]) < 2e-4
        b = res[0]['pts_bbox']
        close(b['scores_3d'], zt[f'scores3d{f}'], 2e-4)
        assert rowset_err(torch.as_tensor(b['boxes_3d']), torch.from_numpy(zt[f'boxes3d{f}'])) < 2e-4
        same = (b['labels_3d'].numpy() == zt[f'labels3d{f}']).mean()
        assert same > 0.97, same                  # (ties within fp32 noise may swap neighbours in the top-300)
    h = o.pts_bbox_head
    n = C.MEM_ROWS
    assert rowset_err(h.memory_embedding[0, :n], torch.from_numpy(zt['memory_embedding'])) < 2e-4
    assert rowset_err(h.memory_reference_point[0, :n], torch.from_numpy(zt['memory_reference_point'])) < 2e-4
    close(h.memory_timestamp[0, :n], zt['memory_timestamp'])
    close(h.memory_egopose[0, :n], zt['memory_egopose'], 1e-4)


def test_state_dict_names_and_shapes_are_the_references():
    """every parameter / buffer of the reference detector built from ITS OWN config file (1065 entries) exists under the
    same name with the same shape in the oracle and in the CUDA product."""
    import far3d_b200.plugin  # noqa: F401
    from far3d_b200 import api
    from far3d_b200.compat import DETECTORS, build_from_cfg
    want = {k: tuple(v) for k, v in json.load(open(os.path.join(GOLDEN, 'ref_state_dict_full.json'))).items()}
    assert len(want) == 1065
    mc = api.load_model_cfg(num_cams=7)
    got_p = {k: tuple(v.shape) for k, v in build_from_cfg(mc, DETECTORS).state_dict().items()}
    mo = dict(mc); mo.pop('type')
    got_o = {k: tuple(v.shape) for k, v in O.Far3D(**mo).state_dict().items()}
    assert got_p == want
    assert got_o == want


def test_av2_result_export_vs_reference():
    """far3d_b200.export.results_to_av2 against the table the reference's own Argoverse2Dataset.format_results produced for the
    same seeded detections (tests/golden/make_ref_golden.py av2_export): identical values, columns, order and index; the
    feather file written is the score-sorted table the AV2 tools read."""
    import pandas as pd
    sys.path.insert(0, GOLDEN)
    import tempfile
    from make_ref_golden import av2_export_case
    from far3d_b200 import api, export
    outs, infos = av2_export_case()
    names = api.Config.fromfile(api.DEFAULT_CONFIG).class_names
    assert len(names) == 26
    ref = pd.read_feather(os.path.join(GOLDEN, 'ref_av2_export.feather')).set_index(['log_id', 'timestamp_ns']).sort_index()
    with tempfile.TemporaryDirectory() as d:
        got = export.results_to_av2([dict(pts_bbox=o) for o in outs], infos, names, feather_path=os.path.join(d, 'dts'))
        on_disk = pd.read_feather(os.path.join(d, 'dts.feather'))
    pd.testing.assert_frame_equal(got, ref, check_exact=True)
    assert list(on_disk.columns[:2]) == ['log_id', 'timestamp_ns'] and on_disk['score'].is_monotonic_decreasing
    assert len(on_disk) == sum(len(o['scores_3d']) for o in outs)
    q = got[['qw', 'qx', 'qy', 'qz']].to_numpy()
    np.testing.assert_allclose((q ** 2).sum(1), 1.0, atol=1e-6)


def test_image_preprocessing_oracle_vs_reference_pipeline():
    """oracle/preprocess.py against the output of the reference's own NormalizeMultiviewImage + AV2PadMultiViewImage classes
    (tests/golden/make_ref_golden.py preprocess): three uint8 views of different sizes, 'same2max' padding, both channel
    orders.  Bit-exact: one float32 subtraction and one float32 multiplication per sample."""
    from oracle import preprocess as P
    z = np.load(os.path.join(GOLDEN, 'ref_preprocess.npz'))
    views = [z[f'view{i}'] for i in range(3)]
    np.testing.assert_array_equal(P.normalize_pad_u8(views, z['mean'], z['std'], to_rgb=False), z['out'])
    np.testing.assert_array_equal(P.normalize_pad_u8(views, z['mean'], z['std'], to_rgb=True), z['out_rgb'])
    assert z['out'].shape == (3, 3, 48, 64) and z['pad_shape'].tolist() == [[48, 64, 3]] * 3
    assert float(np.abs(z['out'][0, :, 40:, :]).max()) == 0.0            # pad_val 0 AFTER normalisation


def _resize_crop_params(T, shapes, seed):
    """the augmentation parameters AV2ResizeCropFlipRotImageV2 draws for these views (np.random consumed in its order)"""
    np.random.seed(seed)
    out = []
    for hw in shapes:
        if hw[0] > hw[1]:
            r, dims, crop = T._sample_augmentation_f(hw + (3,))
            first = (dims, crop, False)
            r2, dims2, crop2, flip2, _ = T._sample_augmentation((crop[3] - crop[1], crop[2] - crop[0], 3))
            out.append([first, (dims2, crop2, flip2)])
        else:
            r, dims, crop, flip, _ = T._sample_augmentation(hw + (3,))
            out.append([(dims, crop, flip)])
    return out


def test_resize_crop_oracle_vs_reference_pipeline():
    """oracle/preprocess.py (numpy restatement of Pillow's bicubic resize + crop + flip) against the images the reference's own
    AV2ResizeCropFlipRotImageV2 class produced (tests/golden/make_ref_golden.py resize_crop: landscape and portrait views, the
    portrait one through the transform twice) - bit-exact; and the host logic of the product's device transform (sampling, 3x3
    post-homography, intrinsics, lidar2img) against the same fixture."""
    sys.path.insert(0, GOLDEN)
    from make_ref_golden import RESIZE_CROP_CONF, RESIZE_CROP_SEED, RESIZE_CROP_VIEWS, resize_crop_case
    from far3d_b200 import imgproc
    from oracle import preprocess as P
    z = np.load(os.path.join(GOLDEN, 'ref_resize_crop.npz'))
    views, intr, extr = resize_crop_case()
    T = imgproc.AV2ResizeCropFlipRotImageV2(data_aug_conf=dict(RESIZE_CROP_CONF))
    for i, steps in enumerate(_resize_crop_params(T, RESIZE_CROP_VIEWS, RESIZE_CROP_SEED)):
        img = views[i]
        for dims, crop, flip in steps:
            img = P.resize_crop_flip_u8(img, dims, crop, flip)
        np.testing.assert_array_equal(img, z[f'img{i}'])
    np.testing.assert_array_equal(P.resize_crop_flip_u8(views[0], (102, 77), (-6, 10, 110, 90), flip=True), z['flip_img'])
    np.testing.assert_allclose(np.asarray(T._ida_mat(0.5, (-6, 10, 110, 90), True, 0), dtype=np.float64), z['flip_ida'], rtol=0, atol=0)

    # host side of the device transform with the pixel op replaced by the oracle: every non-pixel key of the fixture
    class HostOnly(imgproc.AV2ResizeCropFlipRotImageV2):
        pass
    real = imgproc.resize_crop_u8
    imgproc.resize_crop_u8 = lambda src, dims, crop, flip=False, out=None: P.resize_crop_flip_u8(np.asarray(src), dims, crop, flip)
    try:
        np.random.seed(RESIZE_CROP_SEED)
        res = HostOnly(data_aug_conf=dict(RESIZE_CROP_CONF))(dict(img=list(views), intrinsics=[k.copy() for k in intr],
                                                                  extrinsics=[e.copy() for e in extr]))
    finally:
        imgproc.resize_crop_u8 = real
    for i in range(len(views)):
        np.testing.assert_array_equal(np.asarray(res['img'][i]), z[f'img{i}'])
    np.testing.assert_array_equal(np.stack([np.asarray(k, dtype=np.float64) for k in res['intrinsics']]), z['intrinsics'])
    np.testing.assert_array_equal(np.stack([np.asarray(k, dtype=np.float64) for k in res['lidar2img']]), z['lidar2img'])
    np.testing.assert_array_equal(np.stack([np.asarray(k, dtype=np.float64) for k in res['ida_mat']]), z['ida_mat'])


def test_resize_oracle_is_pillow_and_coefficient_tables():
    """the oracle's resize against Pillow itself (third-party dependency of the reference; present in this image) on random
    sizes - up- and down-scaling, identity along one axis - and the product's HOST coefficient builder (far3d_resample_coeffs,
    no GPU needed) against the oracle's tables, including the AV2 size pairs."""
    import ctypes
    from PIL import Image
    from far3d_b200 import _lib
    from oracle import preprocess as P
    rng = np.random.default_rng(3)
    for H, W, nw, nh in ((155, 205, 96, 73), (64, 48, 100, 90), (50, 50, 50, 20), (33, 77, 10, 77), (97, 131, 131, 97), (310, 410, 196, 148)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        np.testing.assert_array_equal(P.pil_resize_u8(img, nw, nh), np.array(Image.fromarray(img).resize((nw, nh))))
    for _ in range(40):                                      # random small shapes incl. extreme ratios and 1-pixel axes
        H, W, nh, nw = (int(v) for v in rng.integers(1, 90, 4))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        np.testing.assert_array_equal(P.pil_resize_u8(img, nw, nh), np.array(Image.fromarray(img).resize((nw, nh))))
    lib = _lib.load()
    av2 = [(2048, o) for o in range(962, 1127)] + [(1550, o) for o in range(728, 853)]     # every size the (0.47, 0.55) resize range reaches
    for n_in, n_out in av2 + [(1550, 2092), (2048, 2764), (64, 100), (7, 3), (5, 5), (3, 40), (1, 9), (9, 1)]:
        b, k = P.resample_coeffs(n_in, n_out)
        ks = lib.far3d_resample_ksize(n_in, n_out)
        assert ks == k.shape[1]
        bb, kk = np.empty((n_out, 2), np.int32), np.empty((n_out, ks), np.int32)
        _lib.call('far3d_resample_coeffs', n_in, n_out, bb.ctypes.data_as(ctypes.c_void_p), kk.ctypes.data_as(ctypes.c_void_p))
        np.testing.assert_array_equal(bb, b)
        np.testing.assert_array_equal(kk, k)


@pytest.mark.skipif(not os.path.isdir('/root/reference/projects/mmdet3d_plugin'), reason='reference tree not present')
def test_fixtures_are_what_the_reference_produces_live():
    """build container only: run the reference's own modules again and compare with the committed fixtures."""
    sys.path.insert(0, GOLDEN)
    import ref_shims as R
    mods = R.load_reference()
    pe = mods['models/utils/positional_encoding.py']
    tr = mods['models/utils/detr3d_transformer.py']
    z = np.load(os.path.join(GOLDEN, 'ref_modules.npz'))
    x3, x1, xn = C.posenc_inputs()
    np.testing.assert_array_equal(pe.pos2posemb3d(x3).numpy(), z['pos2posemb3d'])
    from far3d_b200 import synthetic
    oracle_m = O.DeformableFeatureAggregationCuda(**C.DFA_CFG)
    synthetic.randomize_(oracle_m, 2)
    m = tr.DeformableFeatureAggregationCuda(**C.DFA_CFG)
    m.load_state_dict(oracle_m.state_dict(), strict=True)
    m.eval()
    a = C.dfa_inputs()
    with torch.no_grad():
        out = m(a['x'], a['query_pos'], a['feat'], a['reference_points'], a['spatial'], a['start'], a['pc_range'], a['lidar2img'],
                a['metas'])
    close(out, z['dfa_out'], 1e-6)
    # the reference's config file is the one the product ships a restatement of
    from far3d_b200 import api
    assert R.reference_model_cfg() == api.load_model_cfg(num_cams=7)
    # image-side fixture: the reference's own NormalizeMultiviewImage / AV2PadMultiViewImage classes again
    pl = R.load_reference_pipelines()
    zp = np.load(os.path.join(GOLDEN, 'ref_preprocess.npz'))
    res = dict(img=[zp[f'view{i}'].astype(np.float32) for i in range(3)])
    res = pl['transform_3d.py'].NormalizeMultiviewImage(mean=zp['mean'].tolist(), std=zp['std'].tolist(), to_rgb=False)(res)
    res = pl['custom_pipeline.py'].AV2PadMultiViewImage(size='same2max')(res)
    np.testing.assert_array_equal(np.stack([i.transpose(2, 0, 1) for i in res['img']]), zp['out'])
    # resize / crop fixture: the reference's own AV2ResizeCropFlipRotImageV2 again
    from make_ref_golden import RESIZE_CROP_CONF, RESIZE_CROP_SEED, resize_crop_case
    zr = np.load(os.path.join(GOLDEN, 'ref_resize_crop.npz'))
    views, intr, extr = resize_crop_case()
    np.random.seed(RESIZE_CROP_SEED)
    res = pl['custom_pipeline.py'].AV2ResizeCropFlipRotImageV2(data_aug_conf=dict(RESIZE_CROP_CONF))(
        dict(img=[v.astype(np.float32) for v in views], intrinsics=[k.copy() for k in intr], extrinsics=[e.copy() for e in extr]))
    for i, im in enumerate(res['img']):
        np.testing.assert_array_equal(im, zr[f'img{i}'].astype(np.float32))
    np.testing.assert_array_equal(np.stack([np.asarray(k, dtype=np.float64) for k in res['lidar2img']]), zr['lidar2img'])
    # closed-form post-homography matrix of the device transform against the reference's torch arithmetic (_img_transform), random
    # resize / crop / flip parameters, and both sampling functions under the same np.random stream
    from PIL import Image
    from far3d_b200 import imgproc
    Tref = pl['custom_pipeline.py'].AV2ResizeCropFlipRotImageV2(data_aug_conf=dict(RESIZE_CROP_CONF, rand_flip=True))
    Tdev = imgproc.AV2ResizeCropFlipRotImageV2(data_aug_conf=dict(RESIZE_CROP_CONF, rand_flip=True))
    prng = np.random.default_rng(8)
    small = Image.fromarray(views[0][:40, :50])
    for _ in range(40):
        rs = float(prng.uniform(0.3, 1.7))
        dims = (int(50 * rs), int(40 * rs))
        x0, y0 = int(prng.integers(-5, 20)), int(prng.integers(-5, 20))
        crop = (x0, y0, x0 + int(prng.integers(8, 40)), y0 + int(prng.integers(8, 40)))
        flip = bool(prng.integers(0, 2))
        _, ida_ref, _ = Tref._img_transform(small, resize=rs, resize_dims=dims, crop=crop, flip=flip, rotate=0)
        np.testing.assert_array_equal(np.asarray(ida_ref), Tdev._ida_mat(rs, crop, flip, 0).numpy())
    for seed in range(5):
        np.random.seed(seed); a_ = Tref._sample_augmentation(views[0].astype(np.float32))
        np.random.seed(seed); b_ = Tdev._sample_augmentation(views[0].shape)
        assert a_ == b_
        assert Tref._sample_augmentation_f(views[1].astype(np.float32)) == Tdev._sample_augmentation_f(views[1].shape)
    # the shipped config restates the reference's class list and normalisation constants
    ns = {}
    with open('/root/reference/projects/configs/far3d.py') as f:
        exec(compile(f.read(), 'far3d.py', 'exec'), ns)
    cfg = api.Config.fromfile(api.DEFAULT_CONFIG)
    assert list(cfg.class_names) == ns['class_names'] and dict(cfg.img_norm_cfg) == ns['img_norm_cfg'] == api.DEFAULT_IMG_NORM_CFG
    # result export: the reference's own format_results again
    import pandas as pd
    from make_ref_golden import av2_export_case
    ref = R.load_reference_av2_export()
    ds = ref['argoverse2_dataset.py'].Argoverse2Dataset
    outs, infos = av2_export_case()

    class Self:
        data_infos, CLASSES = infos, ds.CLASSES

        def box_to_av2(self, b):
            return ds.box_to_av2(self, b)
    live = ds.format_results(Self(), [dict(pts_bbox=dict(o, boxes_3d=ref['LiDARInstance3DBoxes'](o['boxes_3d']))) for o in outs])
    pd.testing.assert_frame_equal(live, pd.read_feather(os.path.join(GOLDEN, 'ref_av2_export.feather'))
                                  .set_index(['log_id', 'timestamp_ns']).sort_index(), check_exact=True)


def test_cfg2_full_size_two_frames():
    """BASELINE.json configs[1] at full size (7 x 960x640, V-99, 6 layers, ~1047 queries incl. ~147 adaptive, second frame
    reading the memory bank): the oracle against the reference detector's own outputs."""
    from far3d_b200 import api, synthetic
    from helpers import rows_close
    z = np.load(os.path.join(GOLDEN, 'ref_cfg2_frames.npz'))
    o = build_oracle(api.load_model_cfg(num_cams=7), seed=0)
    synthetic.cold_2d_head_(o, C.CFG2_HEAD_SCALE, C.CFG2_HEAD_SCALE)
    for f in range(C.CFG2_FRAMES):
        metas, data = synthetic.make_frame('cfg2', f)
        res, outs = o.simple_test(metas, **data)
        assert outs['all_cls_scores'].shape[2] == z[f'cls{f}'].shape[1]
        close(outs['reference_points2d'], z[f'ref2d{f}'], 1e-4)
        close_sampled(outs['feat_flatten'], z, f'feat_flatten{f}', 1e-4)
        nfix = o.pts_bbox_head.num_query + outs['reference_points2d'].shape[1]
        rows_close(outs['all_cls_scores'][-1][0, :nfix], z[f'cls{f}'][0, :nfix], 3e-4)
        rows_close(outs['all_bbox_preds'][-1][0, :nfix], z[f'box{f}'][0, :nfix], 3e-4)
        rows_close(outs['all_cls_scores'][-1][0], z[f'cls{f}'][0], 3e-4, match_rows=True)
        rows_close(outs['all_bbox_preds'][-1][0], z[f'box{f}'][0], 3e-4, match_rows=True)
        close(res[0]['pts_bbox']['scores_3d'], z[f'scores3d{f}'], 2e-4)
