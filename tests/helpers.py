"""Shared test helpers: small-config builders for the oracle and the CUDA product, error metrics."""
import copy
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def model_cfg(spec='V-19-eSE', num_cams=2, num_query=50, num_layers=2, roi_head=True, memory_len=1024, num_propagated=256,
              topk_proposals=256):
    """The reference's far3d.py model dict (re-stated in configs/far3d_av2.py) shrunk for fast tests."""
    from far3d_b200.compat import Config
    cfg = Config.fromfile(os.path.join(ROOT, 'configs', 'far3d_av2.py'))
    mc = copy.deepcopy(dict(cfg.model))
    mc['img_backbone']['spec_name'] = spec
    h = mc['pts_bbox_head']
    h['num_query'], h['memory_len'], h['num_propagated'], h['topk_proposals'] = num_query, memory_len, num_propagated, topk_proposals
    dec = h['transformer']['decoder']
    dec['num_layers'] = num_layers
    for a in dec['transformerlayers']['attn_cfgs']:
        if a['type'] == 'DeformableFeatureAggregationCuda':
            a['num_cams'] = num_cams
    if not roi_head:
        mc['img_roi_head'] = None
        h['add_query_from_2d'] = False
    return mc


def variant_cfg(kind, **kw):
    """tiny stand-ins for BASELINE.json configs[3] / configs[4]: the conventions that differ from the Argoverse2 config.
    `nus`: 6-camera rig conventions of the StreamPETR lineage the reference descends from - code_size 10 (velocity channels
    feed `denormalize_bbox`'s 10-wide branch, core/bbox/util.py:44-50), pc_range +-51.2 m, z in [-5, 3];
    `longrange`: pc_range +-150 m and a larger learned-query set (2000 at full size)."""
    import copy as _copy
    mc = _copy.deepcopy(model_cfg(**kw))
    h = mc['pts_bbox_head']
    if kind == 'nus':
        rng = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
        h['code_size'] = 10
        h['code_weights'] = [1.0] * 8 + [0.2, 0.2]
        post = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    elif kind == 'longrange':
        rng = [-150.0, -150.0, -5.0, 150.0, 150.0, 5.0]
        post = rng
    else:
        raise KeyError(kind)
    h['bbox_coder']['pc_range'] = rng
    h['bbox_coder']['post_center_range'] = post
    return mc


def build_oracle(mc, seed=1):
    from oracle import model as O
    from far3d_b200 import synthetic
    m = dict(mc)
    m.pop('type', None)
    o = O.Far3D(**m).eval()
    synthetic.randomize_(o, seed)
    return o


def build_product(mc, state_dict, device, precision='fp16x3'):
    import far3d_b200.plugin  # noqa: F401  (registers the classes)
    from far3d_b200.compat import DETECTORS, build_from_cfg
    m = build_from_cfg(mc, DETECTORS).eval()
    m.load_state_dict(state_dict)
    m.to(device)
    m.set_precision(precision)
    return m


def rel_err(a, b):
    """max |a-b| / max |b|  (the 1e-3 'rel fp32' bar of BASELINE.json is applied to this and to rel_l2)."""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def to_dev(data, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in data.items()}


def rowset_err(a, b):
    """max over rows of a of the distance to the nearest row of b, relative to max|b| (order-invariant comparison)."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    d = torch.cdist(a, b, p=float('inf')).min(dim=1).values
    return (d.max() / b.abs().max()).item()


def rows_close(a, b, tol, frac=0.995, hard=None, match_rows=False):
    """Row-wise comparison that tolerates isolated discontinuity rows: a query whose sample lands within rounding distance of
    an in-bounds border or of a pixel boundary may take the other branch of the bilinear rule on the two sides (the reference
    itself has this discontinuity).  At least `frac` of the rows must agree within `tol` and every row within `hard`
    (default 20 x tol), relative to max|b|.  `match_rows`: compare each row of a with its nearest row of b (order-free)."""
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    a, b = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    scale = b.abs().max().clamp_min(1e-30)
    if match_rows:
        d = torch.cdist(a, b, p=float('inf')).min(dim=1).values / scale
    else:
        assert a.shape == b.shape, (a.shape, b.shape)
        d = (a - b).abs().amax(dim=1) / scale
    ok = (d < tol).double().mean().item()
    worst = d.max().item()
    assert ok >= frac, f'only {ok:.4f} of rows within {tol} (worst {worst:.3e})'
    assert worst < (hard if hard is not None else 20 * tol), f'worst row {worst:.3e}'
    return ok, worst
