"""Seeded inputs and sampling helpers shared by tests/golden/make_ref_golden.py (which runs the REFERENCE's own modules on
them in the build container) and by the parity tests that compare the oracle / the CUDA path with the committed outputs."""
import numpy as np
import torch

TINY_FRAMES = 2
CFG2_FRAMES = 2
CFG2_HEAD_SCALE = 0.05        # synthetic.cold_2d_head_(scale): ~150 adaptive queries at cfg-2
MEM_ROWS = 320                 # rows of the memory bank kept in the fixture (256 fresh + the head of the old bank)
SAMPLE = 4096                  # large tensors are stored as SAMPLE seeded positions + their L2 norm

# BASELINE.json configs[2..4] at full size (tests/golden/make_ref_golden.py full_frames -> ref_<name>_frames.npz)
FULL_CASES = {
    'cfg3': dict(rig=(7, 640, 960), frames=8, variant=None, num_query=644),           # Argoverse2 far3d.py, streaming temporal
    'cfg4': dict(rig=(6, 640, 1600), frames=2, variant='nus', num_query=644),         # nuScenes 6-cam 1600x640, 10-wide box code
    'cfg5': dict(rig=(7, 1024, 1536), frames=2, variant='longrange', num_query=2000),  # long-range stress, 2000 queries, 150 m
}


def full_model_cfg(name):
    """the model dict of a FULL_CASES entry: the reference's far3d.py config (V-99, 6 decoder layers) with the case's camera
    count, query count and range / box-code conventions"""
    from helpers import model_cfg, variant_cfg
    case = FULL_CASES[name]
    kw = dict(spec='V-99-eSE', num_cams=case['rig'][0], num_query=case['num_query'], num_layers=6)
    return variant_cfg(case['variant'], **kw) if case['variant'] else model_cfg(**kw)


MLN_CASES = {'spatial14': (14, False), 'egopose180': (180, True)}
CODER_CFG = dict(pc_range=[-152.4, -152.4, -5.0, 152.4, 152.4, 5.0], post_center_range=[-152.4, -152.4, -5.0, 152.4, 152.4, 5.0],
                 max_num=300, voxel_size=[0.2, 0.2, 8], num_classes=26)
DFA_CFG = dict(embed_dims=256, num_groups=8, num_levels=4, num_cams=7, dropout=0.1, num_pts=13, bias=2., batch_first=True)
DFA_SHAPES = [(20, 30), (10, 15), (5, 8), (3, 4)]
DFA_PAD_HW = (160, 240)
DFA_NQ = 96
PC_RANGE = [-152.4, -152.4, -5.0, 152.4, 152.4, 5.0]


def _g(seed):
    return torch.Generator().manual_seed(seed)


def sample_index(numel, n=SAMPLE):
    return np.random.RandomState(numel % 65521).choice(numel, size=min(n, numel), replace=False)


def sample(t):
    """values of tensor `t` at a position set that depends only on its size."""
    t = torch.as_tensor(t).detach().cpu().contiguous().reshape(-1)
    return t[torch.from_numpy(sample_index(t.numel()))].numpy()


def norm(t):
    return np.float64(torch.as_tensor(t).detach().double().cpu().norm().item())


def posenc_inputs():
    g = _g(11)
    return torch.rand(2, 5, 3, generator=g), torch.rand(2, 7, 1, generator=g) * 4 - 2, torch.randn(3, 4, 15, generator=g)


def mln_inputs(c_dim):
    g = _g(12 + c_dim)
    return torch.randn(2, 9, 256, generator=g), torch.randn(2, 1 if c_dim == 14 else 9, c_dim, generator=g)


def transform_inputs():
    g = _g(13)
    pts = torch.randn(1, 6, 3, generator=g) * 30
    yaw = 0.3
    pose = torch.tensor([[np.cos(yaw), -np.sin(yaw), 0, 2.5], [np.sin(yaw), np.cos(yaw), 0, -1.0], [0, 0, 1, 0.2], [0, 0, 0, 1]],
                        dtype=torch.float32).unsqueeze(0)
    return pts, pose


def coder_inputs():
    g = _g(14)
    cls = torch.randn(2, 1, 120, 26, generator=g) * 2
    box = torch.randn(2, 1, 120, 8, generator=g)
    box[..., 0:2] *= 100
    box[..., 2] *= 3
    return cls, box


def v99_input():
    return torch.randn(2, 3, 128, 192, generator=_g(15))


def dfa_inputs():
    from far3d_b200 import synthetic
    g = _g(16)
    N, (H, W) = DFA_CFG['num_cams'], DFA_PAD_HW
    metas, data = synthetic.make_frame((N, H, W), 0)
    sp = torch.tensor(DFA_SHAPES)
    st = torch.cat((sp.new_zeros(1), sp.prod(1).cumsum(0)[:-1]))
    S = int(sp.prod(1).sum())
    ref = torch.rand(1, DFA_NQ, 3, generator=g)
    ref[..., :2] = ref[..., :2] * 0.3 + 0.35          # within ~45 m of the rig so that most points land in some camera
    return dict(x=torch.randn(1, DFA_NQ, 256, generator=g), query_pos=torch.randn(1, DFA_NQ, 256, generator=g),
                feat=torch.randn(N, S, 256, generator=g), reference_points=ref, spatial=sp, start=st,
                pc_range=torch.tensor(PC_RANGE), lidar2img=data['lidar2img'], metas=metas)
