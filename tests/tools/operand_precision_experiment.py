"""CPU emulation: error at the backbone / FPN outputs when the conv operands are rounded to fp16 (activations, weights, both)
or split into the three-term fp16x3 form, against the fp32 oracle.  DESIGN.md section 8 item 1.
    python tests/tools/operand_precision_experiment.py"""
import sys, os, torch, torch.nn as nn, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from helpers import build_oracle
from far3d_b200 import api, synthetic
torch.set_num_threads(8)
mc = api.load_model_cfg(num_cams=2)
o = build_oracle(mc, seed=0)
bb, neck = o.img_backbone, o.img_neck
MODE = {'m': 'exact'}
def r16(t):
    return t.half().float()
def r22(t):
    h = t.half().float(); return h + (t - h).half().float()
orig = F.conv2d
def conv_hook(x, w, b=None, *a, **k):
    m = MODE['m']
    if m == 'exact': return orig(x, w, b, *a, **k)
    if m == 'a16': return orig(r16(x), r22(w), b, *a, **k)
    if m == 'w16': return orig(r22(x), r16(w), b, *a, **k)
    if m == 'both16': return orig(r16(x), r16(w), b, *a, **k)
    if m == 'x3': # hi*hi + lo*hi + hi*lo
        xh = r16(x); xl = (x - xh).half().float(); wh = r16(w); wl = (w - wh).half().float()
        return orig(xh, wh, b, *a, **k) + orig(xl, wh, None, *a, **k) + orig(xh, wl, None, *a, **k)
F.conv2d = conv_hook
torch.nn.functional.conv2d = conv_hook
g = torch.Generator().manual_seed(0)
x = torch.randn(2, 3, 320, 480, generator=g)
def run():
    with torch.no_grad():
        feats = bb(x)
        outs = neck(feats)
    return feats, outs
MODE['m'] = 'exact'
ref_f, ref_o = run()
# float64 reference
for m in ['a16', 'w16', 'both16', 'x3']:
    MODE['m'] = m
    f, o_ = run()
    e_bb = [((a - b).norm() / b.norm()).item() for a, b in zip(f, ref_f)]
    e_fp = [((a - b).norm() / b.norm()).item() for a, b in zip(o_, ref_o)]
    e_mx = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(o_, ref_o)]
    print(m, 'backbone rel_l2', ['%.2e' % e for e in e_bb], 'fpn rel_l2', ['%.2e' % e for e in e_fp], 'fpn relmax', ['%.2e' % e for e in e_mx])
