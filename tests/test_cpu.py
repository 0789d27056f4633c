"""CPU suite (`-m "not gpu"`): oracle self-consistency + golden vectors, host logic, C-ABI surface."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT, build_oracle, model_cfg

REF_CFG = '/root/reference/projects/configs/far3d.py'


# ------------------------------------------------------------------------------------------------ oracle
def _rand_msda(seed, BN=2, G=2, D=4, Nq=5, L=2, P=3, shapes=((6, 9), (3, 5))):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(BN, S, G, D, generator=g)
    loc = torch.rand(BN, Nq, G, L, P, 2, generator=g) * 1.4 - 0.2        # some samples fall outside
    w = torch.rand(BN, Nq, G, L * P, generator=g)
    sp = torch.tensor(shapes, dtype=torch.long)
    st = torch.cat((sp.new_zeros(1), sp.prod(1).cumsum(0)[:-1]))
    return value, sp, st, loc, w


def test_msda_three_restatements_agree():
    """grid_sample form (the reference's own restatement, sparse_blocks.py:234-255) == scalar im2col rule == C oracle."""
    from oracle import cref, msda
    value, sp, st, loc, w = _rand_msda(0)
    a = msda.msda_grid_sample(value, sp, st, loc, w).numpy()
    b, idx_b, val_b = msda.msda_scalar(value.numpy(), sp.numpy(), st.numpy(), loc.numpy(), w.numpy())
    c, idx_c, val_c = cref.msda(value.numpy(), sp.numpy(), st.numpy(), loc.numpy(), w.numpy(), debug=True)
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(c, b, rtol=1e-6, atol=1e-6)
    assert np.array_equal(val_b, val_c.astype(bool))
    assert np.array_equal(idx_b[val_b], idx_c.astype(np.int64)[val_b])
    assert 0.2 < val_b.mean() < 0.95          # both branches of the bounds test are exercised


def test_oracle_restatements_agree_on_random_shapes():
    """hypothesis sweep over ragged shapes (1-pixel levels, single cameras, group widths that are not 32, points straddling the
    camera plane): the fused C oracle == the module-level torch oracle (projection + grid_sample MSDA + camera sum), and the C
    MSDA == the scalar numpy statement of mmcv's im2col rule, masks and floor indices exactly."""
    from hypothesis import given, settings, strategies as hs
    from far3d_b200 import synthetic
    from oracle import cref, msda
    from oracle import model as O

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(seed=hs.integers(0, 10 ** 6), N=hs.integers(1, 4), Nq=hs.integers(1, 9), G=hs.sampled_from([1, 2, 4]),
           D=hs.sampled_from([1, 3, 8]), P=hs.integers(1, 5),
           shapes=hs.lists(hs.tuples(hs.integers(1, 9), hs.integers(1, 9)), min_size=1, max_size=3),
           spread=hs.sampled_from([0.5, 5.0, 40.0]))
    def run(seed, N, Nq, G, D, P, shapes, spread):
        g = torch.Generator().manual_seed(seed)
        L, C = len(shapes), G * D
        sp = torch.tensor(shapes)
        st = torch.cat((sp.new_zeros(1), sp.prod(1).cumsum(0)[:-1]))
        S = int(sp.prod(1).sum())
        _, data = synthetic.make_frame((N, 64, 96), 0, seed=seed % 97)
        feat = torch.randn(N, S, C, generator=g)
        kp = torch.randn(1, Nq, P, 3, generator=g) * spread
        w = torch.rand(N, Nq, G, L * P, generator=g)
        m = O.DeformableFeatureAggregationCuda(embed_dims=C, num_groups=G, num_levels=L, num_cams=N, num_pts=P)
        loc = m.sampling_locations(kp, data['lidar2img'], (64, 96))
        ref = msda.msda_grid_sample(feat.view(N, S, G, D), sp, st, loc, w).view(1, N, Nq, C).sum(1)
        out = cref.deform_agg(feat.numpy(), sp.numpy(), st.numpy(), kp.numpy(), data['lidar2img'].numpy(), w.numpy(), 64, 96, G)
        scale = max(float(ref.abs().max()), 1e-3)
        assert float(np.abs(out - ref.numpy()).max()) <= 2e-4 * scale + 1e-5
        b, idx_b, val_b = msda.msda_scalar(feat.view(N, S, G, D).numpy(), sp.numpy(), st.numpy(), loc.numpy(), w.numpy())
        c, idx_c, val_c = cref.msda(feat.view(N, S, G, D).numpy(), sp.numpy(), st.numpy(), loc.numpy(), w.numpy(), debug=True)
        np.testing.assert_allclose(c, b, rtol=1e-5, atol=1e-5)
        assert np.array_equal(val_b, val_c.astype(bool))
        assert np.array_equal(idx_b[val_b], idx_c.astype(np.int64)[val_b])

    run()


def test_msda_edge_cases():
    """exact borders: loc 0 / 1 (h_im = -0.5 / H-0.5) are sampled with zero padding; loc far outside contributes 0."""
    from oracle import cref, msda
    value = torch.ones(1, 12, 1, 2)
    sp = torch.tensor([[3, 4]]); st = torch.tensor([0])
    loc = torch.tensor([[0., 0.], [1., 1.], [0.5, 0.5], [-0.3, 0.5], [5., 5.], [1.0 + 0.49 / 4, 0.5]]).view(1, 6, 1, 1, 1, 2)
    w = torch.ones(1, 6, 1, 1)
    out = msda.msda_grid_sample(value, sp, st, loc, w)[0, :, 0]
    ref, _, valid = cref.msda(value.numpy(), sp.numpy(), st.numpy(), loc.numpy(), w.numpy(), debug=True)
    np.testing.assert_allclose(out.numpy(), ref[0, :, 0], atol=1e-6)
    np.testing.assert_allclose(ref[0, :, 0], [0.25, 0.25, 1.0, 0.0, 0.0, 0.01], atol=1e-6)
    assert valid.reshape(-1).tolist() == [1, 1, 1, 0, 0, 1]


def test_fused_oracle_matches_module_oracle():
    """C fused deform-agg == torch module oracle's feature_sampling path (projection + MSDA + camera sum)."""
    from oracle import cref
    from oracle import model as O
    g = torch.Generator().manual_seed(3)
    B, N, Nq, G, P, L = 1, 3, 7, 2, 4, 2
    shapes = [(8, 12), (4, 6)]
    S = sum(h * w for h, w in shapes)
    C = G * 4
    m = O.DeformableFeatureAggregationCuda(embed_dims=C, num_groups=G, num_levels=L, num_cams=N, num_pts=P)
    from far3d_b200 import synthetic
    _, data = synthetic.make_frame((N, 64, 96), 0)
    feat = torch.randn(B * N, S, C, generator=g)
    kp = torch.randn(B, Nq, P, 3, generator=g) * 15
    w = torch.rand(B * N, Nq, G, L * P, generator=g)
    loc = m.sampling_locations(kp, data['lidar2img'], (64, 96))
    sp = torch.tensor(shapes); st = torch.tensor([0, shapes[0][0] * shapes[0][1]])
    from oracle.msda import msda_grid_sample
    ref = msda_grid_sample(feat.view(B * N, S, G, -1), sp, st, loc, w).view(B, N, Nq, C).sum(1)
    out, uv, idx, valid = cref.deform_agg(feat.numpy(), sp.numpy(), st.numpy(), kp.numpy(), data['lidar2img'].numpy(),
                                          w.numpy(), 64, 96, G, debug=True)
    np.testing.assert_allclose(out, ref.numpy(), rtol=2e-4, atol=2e-4)
    assert 0.02 < valid.mean() < 0.9


def test_cfg1_resnet18_plumbing_on_cpu():
    """BASELINE.json configs[0]: 1 camera, 256 x 256, a (torchvision) ResNet-18 backbone - not in the reference, plumbing only -
    50 learned queries, 1 decoder layer, on the CPU: the oracle's neck / 2D head / FarHead / memory bank / box decode run
    end to end behind a different backbone over two streamed frames, and the layer's aggregation equals the fused C oracle."""
    import torch.nn as nn
    import torchvision
    from far3d_b200 import synthetic
    from oracle import cref
    from oracle import model as O

    class ResNet18Stages(nn.Module):
        def __init__(self):
            super().__init__()
            r = torchvision.models.resnet18(weights=None)
            self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
            self.layers = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])

        def forward(self, x):
            x = self.stem(x)
            outs = []
            for l in self.layers:
                x = l(x)
                outs.append(x)
            return outs                                   # strides 4, 8, 16, 32; channels 64, 128, 256, 512

    mc = model_cfg(num_cams=1, num_query=50, num_layers=1)
    mc['img_neck'] = dict(mc['img_neck'], in_channels=[64, 128, 256, 512])
    mc['img_backbone'] = ResNet18Stages()
    mc.pop('type', None)
    o = O.Far3D(**mc).eval()
    synthetic.randomize_(o, 4)
    synthetic.cold_2d_head_(o, 0.05, 0.05)               # finite box-size logits (random regressors overflow exp())
    assert len(o.pts_bbox_head.transformer.decoder.layers) == 1
    seen = {}
    layer = o.pts_bbox_head.transformer.decoder.layers[0]
    dfa = [a for a in layer.attentions if isinstance(a, O.DeformableFeatureAggregationCuda)][0]
    hook = dfa.register_forward_hook(lambda m, args, out: seen.update(args=args, out=out))
    for f in range(2):
        metas, data = synthetic.make_frame('cfg1', f)
        res, outs = o.simple_test(metas, **data)
        nq = outs['all_cls_scores'].shape[2]
        m2d = 0 if outs['reference_points2d'] is None else outs['reference_points2d'].shape[1]
        assert outs['all_cls_scores'].shape == (1, 1, nq, 26) and nq == 50 + 256 + m2d
        assert outs['all_bbox_preds'].shape == (1, 1, nq, 8)
        assert torch.isfinite(outs['all_cls_scores']).all() and torch.isfinite(outs['all_bbox_preds']).all()
        b = res[0]['pts_bbox']
        assert b['boxes_3d'].shape[1] == 7 and len(b['scores_3d']) == len(b['labels_3d']) <= 300
        assert (b['scores_3d'][:-1] >= b['scores_3d'][1:]).all() and int(b['labels_3d'].max()) < 26
        assert o.pts_bbox_head.memory_embedding.shape == (1, 256 + 1024, 256)     # top-256 prepended (farhead.py:479-508), trimmed next frame
    hook.remove()
    # the decoder layer's aggregation once more through the fused C oracle (one camera, 4 levels of a 256 x 256 image)
    x, qpos, feat, ref_pts, sp, st, pr, l2i, metas_ = seen['args']
    kp = dfa.key_points(x, ref_pts, pr)
    w = dfa.weights(x, qpos, l2i)
    fused = cref.deform_agg(feat.numpy(), sp.numpy(), st.numpy(), kp.detach().numpy(), l2i.numpy(), w.detach().numpy(), 256, 256, 8)
    expect = (seen['out'] - x)                              # output_proj(features) = module output - residual
    got = dfa.output_proj(torch.from_numpy(fused))
    np.testing.assert_allclose(got.detach().numpy(), expect.detach().numpy(), rtol=1e-3, atol=1e-4)


def test_golden_deform_agg():
    """the committed golden vector (tests/golden/make_golden.py) still comes out of the oracle."""
    from oracle import cref
    z = np.load(os.path.join(GOLDEN, 'deform_agg_small.npz'))
    out, uv, idx, valid = cref.deform_agg(z['feat'], z['shapes'], z['start'], z['key_points'], z['lidar2img'], z['weights'],
                                          float(z['pad_hw'][0]), float(z['pad_hw'][1]), int(z['num_groups']), debug=True)
    np.testing.assert_allclose(out, z['out'], rtol=1e-6, atol=1e-6)
    assert np.array_equal(valid, z['valid'])
    assert np.array_equal(idx[valid.astype(bool)], z['idx'][valid.astype(bool)])
    np.testing.assert_array_equal(uv, z['uv'])


def test_golden_tiny_model():
    """two streamed frames through the full oracle detector reproduce the committed outputs."""
    from far3d_b200 import synthetic
    z = np.load(os.path.join(GOLDEN, 'tiny_model.npz'))
    o = build_oracle(model_cfg(), seed=1)
    for f in range(2):
        metas, data = synthetic.make_frame('tiny', f)
        res, outs = o.simple_test(metas, **data)
        np.testing.assert_allclose(outs['all_cls_scores'].numpy(), z[f'cls{f}'], rtol=2e-3, atol=2e-4)
        np.testing.assert_allclose(outs['all_bbox_preds'].numpy(), z[f'box{f}'], rtol=2e-3, atol=2e-3)


def test_oracle_matches_torch_reference_pieces():
    """third-party pieces restated in the oracle behave like their torch definitions."""
    from oracle import model as O
    x = torch.rand(2, 5, 3)
    e = O.pos2posemb3d(x)
    assert e.shape == (2, 5, 384)
    # (y, x, z) order, interleaved sin/cos (positional_encoding.py:13-25)
    assert torch.allclose(e[..., 0], torch.sin(x[..., 1] * 2 * np.pi), atol=1e-6)
    assert torch.allclose(e[..., 129], torch.cos(x[..., 0] * 2 * np.pi), atol=1e-6)
    n = O.nerf_positional_encoding(torch.rand(4, 15))
    assert n.shape == (4, 180)
    fpn = O.FPN([256, 512, 768, 1024], 256, 4, start_level=1, add_extra_convs='on_output', relu_before_extra_convs=True)
    outs = fpn([torch.randn(1, c, s, s + 2) for c, s in ((256, 32), (512, 16), (768, 8), (1024, 4))])
    assert [tuple(o.shape[2:]) for o in outs] == [(16, 18), (8, 10), (4, 6), (2, 3)]


# ------------------------------------------------------------------------------------------------ host logic
def test_reference_config_builds_unchanged():
    import far3d_b200.plugin  # noqa: F401
    from far3d_b200.compat import DETECTORS, Config, build_from_cfg
    mine = Config.fromfile(os.path.join(ROOT, 'configs', 'far3d_av2.py'))
    m = build_from_cfg(mine.model, DETECTORS)
    sd = m.state_dict()
    # key names the released checkpoint uses
    for k in ('img_backbone.stem.stem_1/conv.weight', 'img_backbone.stage3.OSA3_2.layers.0.OSA3_2_0/conv.weight',
              'img_backbone.stage3.OSA3_2.concat.OSA3_2_concat/conv.weight', 'img_backbone.stage3.OSA3_2.ese.fc.weight',
              'img_neck.lateral_convs.0.conv.weight', 'img_neck.fpn_convs.3.conv.bias',
              'pts_bbox_head.transformer.decoder.layers.5.attentions.0.attn.in_proj_weight',
              'pts_bbox_head.transformer.decoder.layers.0.attentions.1.cam_embed.4.weight',
              'pts_bbox_head.transformer.decoder.layers.0.ffns.0.layers.0.0.weight',
              'pts_bbox_head.cls_branches.5.6.weight', 'pts_bbox_head.spatial_alignment.reduce.0.weight',
              'pts_bbox_head.ego_pose_memory.gamma.bias', 'pts_bbox_head.pseudo_reference_points.weight',
              'img_roi_head.multi_level_cls_convs.0.0.bn.running_mean', 'img_roi_head.depthnet.depth_head.1.1.weight'):
        assert k in sd, k
    assert tuple(sd['pts_bbox_head.transformer.decoder.layers.0.ffns.0.layers.0.0.weight'].shape) == (1024, 256)
    o = build_oracle(mine.model)
    assert set(o.state_dict().keys()) == set(sd.keys())
    if os.path.exists(REF_CFG):
        ref = Config.fromfile(REF_CFG)
        assert ref.model == mine.model
        build_from_cfg(ref.model, DETECTORS)


def test_product_has_no_cpu_path(lib_built):
    """ops refuse CPU tensors; the package never imports the oracle."""
    from far3d_b200 import _lib, ops
    with pytest.raises(_lib.Far3DNativeError):
        ops.layernorm(torch.zeros(4, 256), torch.ones(256), torch.zeros(256))
    for dp, _, files in os.walk(os.path.join(ROOT, 'far3d_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f'{f} imports the oracle'


def test_abi_exports_everything_the_header_declares(lib_built):
    hdr = open(os.path.join(ROOT, 'include', 'far3d_b200.h')).read()
    declared = set(re.findall(r'\b(far3d_[a-z0-9_]+)\s*\(', hdr))
    lib = ctypes.CDLL(lib_built)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    from far3d_b200 import _lib
    assert declared <= set(_lib.SIGNATURES)
    l = _lib.load()
    assert l.far3d_abi_version() == 1
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert l.far3d_layernorm(None, None, None, None, None, 4, 256, 1e-5, 0, 0, None) == -1
    assert b'null pointer' in l.far3d_last_error()


def test_frame_and_stream_sharding():
    """shard_frames == the reference's DistributedSampler (shuffle=False) for every rank; stream assignment never cuts a
    stream, covers each exactly once and is balanced; the round-robin interleave keeps every stream's frames in order."""
    from far3d_b200.parallel import assign_streams, interleave_streams, shard_frames
    import math
    for n, world in ((10, 4), (7, 8), (24, 8), (1, 2), (150, 8)):
        total = math.ceil(n / world) * world              # distributed_sampler.py / torch DistributedSampler: total_size
        ref = (list(range(n)) * math.ceil(total / n))[:total]
        per = total // world
        for r in range(world):
            assert shard_frames(n, world, r) == ref[r * per:(r + 1) * per]
    lengths = [156, 157, 150, 30, 160, 155, 90, 156, 12, 140]
    ranks, load = assign_streams(lengths, 4)
    assert sorted(i for r in ranks for i in r) == list(range(len(lengths)))
    assert load == [sum(lengths[i] for i in r) for r in ranks]
    assert max(load) - min(load) <= max(lengths)
    assert assign_streams([5, 5], 4)[1] == [5, 5, 0, 0]
    sched = interleave_streams({'a': 3, 'b': 1, 'c': 2})
    assert sched == [('a', 0), ('b', 0), ('c', 0), ('a', 1), ('c', 1), ('a', 2)]


@pytest.mark.skipif(not os.path.isdir('/root/reference/projects/mmdet3d_plugin'), reason='reference tree not present')
def test_shard_frames_is_the_references_sampler():
    """build container only: the reference's own DistributedSampler class (datasets/samplers/distributed_sampler.py, loaded by
    file; its registry import is the one third-party symbol stubbed) yields exactly `shard_frames` for every rank."""
    import importlib.util
    import types
    from far3d_b200.parallel import shard_frames
    pkg = types.ModuleType('far3d_ref_samplers'); pkg.__path__ = []
    reg = types.ModuleType('far3d_ref_samplers.sampler')

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c
    reg.SAMPLER = _Reg()
    sys.modules['far3d_ref_samplers'], sys.modules['far3d_ref_samplers.sampler'] = pkg, reg
    spec = importlib.util.spec_from_file_location(
        'far3d_ref_samplers.distributed_sampler',
        '/root/reference/projects/mmdet3d_plugin/datasets/samplers/distributed_sampler.py')
    m = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = m
    spec.loader.exec_module(m)
    for n, world in ((10, 4), (7, 8), (24, 8), (150, 8), (3, 2)):
        for r in range(world):
            ref = list(m.DistributedSampler(dataset=list(range(n)), num_replicas=world, rank=r, shuffle=False))
            assert ref == shard_frames(n, world, r), (n, world, r)


def test_per_stream_memory_banks_host_logic():
    """Far3DPipeline's bank switching without a GPU: a stand-in detector whose 'memory' is a counter per stream."""
    from far3d_b200.api import Far3DPipeline

    class Head:
        MEMORY_KEYS = ('memory_embedding',)
        memory_embedding = None

        def export_memory(self):
            return {'memory_embedding': self.memory_embedding}

        def import_memory(self, st):
            self.memory_embedding = None if st is None else st['memory_embedding']

        def reset_memory(self):
            self.memory_embedding = None

    class Model:
        def __init__(self):
            self.pts_bbox_head, self.prev_scene_token = Head(), None

    mdl = Model()
    pipe = Far3DPipeline.wrap(mdl, 'cpu')

    def frame(stream):                       # what a frame's head does to the live bank
        pipe._swap_in(stream)
        h = mdl.pts_bbox_head
        h.memory_embedding = (h.memory_embedding or 0) + 1
        mdl.prev_scene_token = f'scene-{stream}'
        return h.memory_embedding

    assert [frame(s) for s in 'ABABBA'] == [1, 1, 2, 2, 3, 3]
    assert mdl.prev_scene_token == 'scene-A'
    pipe._swap_in('B')
    assert mdl.prev_scene_token == 'scene-B' and mdl.pts_bbox_head.memory_embedding == 3
    pipe.drop_stream('B')                    # live stream dropped: bank gone, detector back to a clean state
    assert mdl.pts_bbox_head.memory_embedding is None and mdl.prev_scene_token is None
    assert frame('B') == 1 and frame('A') == 4
    pipe._swap_in(None)                      # frames without a stream id use whatever bank is live (the reference's behaviour)
    assert mdl.pts_bbox_head.memory_embedding == 4


def test_abi_rejects_bad_arguments_before_touching_the_gpu(lib_built):
    """error behaviour of the C ABI (include/far3d_b200.h): null pointers, non-positive sizes and unsupported dtypes come back as
    FAR3D_E_INVALID with a message naming the entry point - argument checks run before any CUDA call, so this needs no GPU -
    and the Python wrappers refuse CPU tensors instead of computing on the host."""
    import ctypes
    import numpy as np
    import torch
    from far3d_b200 import _lib, ops
    lib = _lib.load()
    hw = np.array([[4, 6]], dtype=np.int32); st = np.array([0], dtype=np.int32)
    hp, sp = hw.ctypes.data_as(ctypes.c_void_p), st.ctypes.data_as(ctypes.c_void_p)
    buf = ctypes.c_void_p(ctypes.addressof(ctypes.create_string_buffer(4096)))       # a non-null (host) address: never dereferenced
    n0 = _lib.launch_count()
    rc = lib.far3d_deform_agg_fwd(None, 0, hp, sp, buf, buf, buf, 64.0, 96.0, buf, 1, 1, 24, 256, 8, 4, 1, 13, None)
    assert rc == -1 and b'far3d_deform_agg_fwd' in lib.far3d_last_error() and b'null pointer' in lib.far3d_last_error()
    rc = lib.far3d_deform_agg_fwd(buf, 0, hp, sp, buf, buf, buf, 64.0, 96.0, buf, 1, 0, 24, 256, 8, 4, 1, 13, None)
    assert rc == -1 and b'non-positive size' in lib.far3d_last_error()
    rc = lib.far3d_deform_agg_fwd(buf, 7, hp, sp, buf, buf, buf, 64.0, 96.0, buf, 1, 1, 24, 256, 8, 4, 1, 13, None)
    assert rc == -1 and b'feat_dtype' in lib.far3d_last_error()
    rc = lib.far3d_deform_agg_fwd(buf, 0, hp, sp, buf, buf, buf, 64.0, 96.0, buf, 1, 1, 24, 250, 8, 4, 1, 13, None)
    assert rc == -1 and b'divisible' in lib.far3d_last_error()
    rc = lib.far3d_msda_fwd(buf, None, buf, buf, buf, buf, 1, 24, 8, 32, 4, 1, 13, None)
    assert rc == -1 and b'far3d_msda_fwd' in lib.far3d_last_error()
    mean = np.zeros(3, np.float32); std = np.array([1, 0, 1], np.float32)
    rc = lib.far3d_normalize_u8(buf, 1, 4, 8, 4, 8, mean.ctypes.data_as(ctypes.c_void_p), std.ctypes.data_as(ctypes.c_void_p), 0, buf, None)
    assert rc == -1 and b'std must be non-zero' in lib.far3d_last_error()
    rc = lib.far3d_normalize_u8(buf, 1, 4, 8, 2, 8, mean.ctypes.data_as(ctypes.c_void_p), mean.ctypes.data_as(ctypes.c_void_p), 0, buf, None)
    assert rc == -1 and b'padded size must cover' in lib.far3d_last_error()
    assert _lib.launch_count() == n0                                             # nothing was launched
    with pytest.raises(_lib.Far3DNativeError, match='no CPU path'):
        ops.normalize_u8(torch.zeros(1, 4, 8, 3, dtype=torch.uint8), [0, 0, 0], [1, 1, 1])


def test_camera_shard_plan():
    from far3d_b200.parallel import shard_cameras
    assert shard_cameras(7, 1) == [(0, 7)]
    assert shard_cameras(7, 2) == [(0, 4), (4, 7)]
    p = shard_cameras(7, 8)
    assert sum(b - a for a, b in p) == 7 and len(p) == 8 and p[-1] == (7, 7)


def test_roi_pack_roundtrip_and_level_shapes():
    """the dense 2D-head maps of a camera travel as one fp32 row; unpacking yields views with the head's `out` layout."""
    from far3d_b200.parallel import level_shapes, pack_roi, roi_layout, unpack_roi
    assert level_shapes(640, 960, [8, 16, 32, 64]) == [(80, 120), (40, 60), (20, 30), (10, 15)]
    assert level_shapes(128, 192, [8, 16, 32, 64]) == [(16, 24), (8, 12), (4, 6), (2, 3)]
    shapes = [(4, 6), (2, 3)]
    plan, total = roi_layout(shapes, 8, 8, 12, 0)                    # padded buffer widths: 8 class, 8 box/obj/centre, 12 depth
    assert total == sum(h * w for h, w in shapes) * (8 + 8) + 12 * 24
    g = torch.Generator().manual_seed(0)
    n = 3
    roi = dict(_cls_nhwc=[torch.randn(n, h, w, 8, generator=g) for h, w in shapes],
               _reg_nhwc=[torch.randn(n, h, w, 8, generator=g) for h, w in shapes],
               _depth_logit_nhwc=torch.randn(n, 4, 6, 12, generator=g))
    rows = pack_roi(roi, plan, torch.zeros(4, total))
    back = unpack_roi(rows[:n], plan, num_classes=5, depth_bins=11)
    for l in range(2):
        assert torch.equal(back['_cls_nhwc'][l], roi['_cls_nhwc'][l]) and back['_cls_nhwc'][l].is_contiguous()
        assert torch.equal(back['enc_cls_scores'][l], roi['_cls_nhwc'][l].permute(0, 3, 1, 2)[:, :5])
        assert torch.equal(back['enc_bbox_preds'][l], roi['_reg_nhwc'][l].permute(0, 3, 1, 2)[:, :4])
        assert torch.equal(back['objectnesses'][l], roi['_reg_nhwc'][l].permute(0, 3, 1, 2)[:, 4:5])
    assert torch.equal(back['depth_logit'], roi['_depth_logit_nhwc'].permute(0, 3, 1, 2)[:, :11]) and back['topk_indexes'] is None
    assert torch.allclose(back['pred_depth'].sum(1), torch.ones(n, 4, 6))
    assert rows[3].abs().sum() == 0


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from far3d_b200.parallel import all_gather_cameras, shard_cameras
    torch.manual_seed(0)
    full = torch.randn(5, 11, 8)                        # [cams, S, C] identical on every rank
    a, b = shard_cameras(5, world)[rank]
    got = all_gather_cameras(full[a:b].contiguous(), 5)
    # persistent send / receive buffers, shard already written in place (the CameraShardedFar3D calling pattern)
    per = -(-5 // world)
    send, recv = torch.zeros(per, 11, 8), torch.zeros(world * per, 11, 8)
    send[:b - a] = full[a:b]
    got2 = all_gather_cameras(send[:b - a], 5, send=send, recv=recv)
    q.put((rank, bool(torch.equal(got, full)) and bool(torch.equal(got2, full)) and got2.data_ptr() == recv.data_ptr()))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4, 8])
def test_all_gather_cameras_gloo(world):
    """the one collective on the path (camera shards -> full feat_flatten) on CPU/gloo: 5 cameras over 2 ranks (3 + 2), over 4
    (2 + 2 + 1 + 0) and over 8 (five ranks with one camera, three with none - the 7-cameras-on-8-GPUs shape of the box)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    [p.join(180) for p in ps]
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert res == [(r, True) for r in range(world)]


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the oracle timed on host cores) prints one JSON line."""
    import json
    env = dict(os.environ, FAR3D_BENCH_CPU_CONFIG='tiny')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'


def test_fp16mx_weight_packing_host_side():
    """ops.pack_weight_mx (host-side packing of static weights for the fp16mx conv; torch plumbing, runs on the CPU): the fp16
    plane is the weight times the power of two both correction products carry (11 + EA + w_exp), never overflows fp16 whatever
    the weights' scale, and plane + e4m3 residual reproduce the weight to ~2^-15 of the largest one; the emulation the GPU
    tests compare the kernel with (tests/mx_emulation.py) reads the same planes."""
    import torch
    from far3d_b200 import ops
    from mx_emulation import mx_decode
    g = torch.Generator().manual_seed(3)
    for scale in (1e-6, 3e-3, 0.07, 1.0, 37.0, 2.5e4):
        w = torch.randn(48, 9, 64, generator=g) * scale
        w[0, 0, 0] = 0.0
        w_hi_s, c8, w_exp = ops.pack_weight_mx(w)
        q = 11 + ops.MX_EA + w_exp
        assert w_hi_s.dtype == torch.float16 and bool(torch.isfinite(w_hi_s.float()).all())
        top = float(w_hi_s.float().abs().max())
        assert 2.0 ** 13 < top <= 2.0 ** 15 * 1.001, (scale, top)            # MX_W_TOP = 5 at EA = -1: max lands in (2^14, 2^15]
        amax = float(w.abs().max())
        assert 2.0 ** (ops.MX_W_TOP - 1) * 0.999 <= amax * 2.0 ** w_exp <= 2.0 ** ops.MX_W_TOP * 1.001
        w_hi = w_hi_s.float() * 2.0 ** -q
        hi8, lo8 = mx_decode(c8.view(torch.uint8).view(48, 9, 128))           # weights: [w_hi8 | w_lo8] per 32-channel group
        assert float((hi8 * 2.0 ** -w_exp - w_hi).abs().max()) <= amax * 2.0 ** -4      # 4-bit copy of the fp16 plane
        rec = w_hi + lo8 * 2.0 ** -(w_exp + 11)
        assert float((rec - w).abs().max()) <= amax * 2.0 ** -15, (scale, float((rec - w).abs().max()) / amax)
    z_hi, z_c8, z_exp = ops.pack_weight_mx(torch.zeros(16, 1, 32))
    assert z_exp == 0 and float(z_hi.float().abs().max()) == 0 and int(z_c8.view(torch.uint8).max()) == 0
