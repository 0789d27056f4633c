"""VoVNet-V2 backbone and FPN neck on the sm_100a kernels (registry names `VoVNet`, `FPN`).

Reference: projects/mmdet3d_plugin/models/backbones/vovnet.py (spec :79-87, OSA module :188-238, stage :241-273,
forward :349-360, eval-mode BN :375-384) and mmdet FPN (config projects/configs/far3d.py:50-57).

B200 design (inference only)
  * activations live NHWC in HBM; BatchNorm (eval) is folded into the conv weights once, ReLU fused in the conv epilogue;
  * an OSA block owns ONE concat buffer [N,H,W, Cin + 5*Cmid]: each 3x3 conv reads a channel slice and writes the
    next slice in place, so the reference's `torch.cat` (vovnet.py:230) never happens and the 1x1 concat conv reads
    the buffer directly;
  * convs run on tcgen05 tensor cores (far3d_conv2d_umma).  precision:
      'fp16x3'           split-fp16 operands, three fp16 MMAs per k-step -> fp32-grade results (2^-17);
      'fp16mx'           fp16 main term + both correction terms as ONE e4m3 (kind::f8f6f4) MMA stream: two tensor-pipe
                         passes per MAC instead of three, ~2^-15 per operand (csrc/common.cuh, tools/mma_mx.cu);
      'fp16'             plain fp16 operands, fp32 accumulate -> fastest, ~1e-2 relative at the backbone output;
      'fp32'             exact fp32 SIMT kernels (far3d_conv2d_f32), the anchor the tensor-core path is checked against;
  * eSE = global-avg-pool + tiny fc + hsigmoid gate applied together with the identity add in one pass that also emits
    the split-fp16 operand of the next block.
Returned feature maps are fp32 tensors of logical shape (N,C,H,W) with channels-last strides (zero-copy views of the
NHWC buffers), so reference-style callers (`img_feats[i].size()`, `.view(B, N, C, H, W)`) keep working.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import ops
from ..compat import BACKBONES, NECKS

_SPECS = {
    'V-19-eSE': dict(stem=[64, 64, 128], conv_ch=[128, 160, 192, 224], out_ch=[256, 512, 768, 1024], layers=3,
                     blocks=[1, 1, 1, 1]),
    'V-39-eSE': dict(stem=[64, 64, 128], conv_ch=[128, 160, 192, 224], out_ch=[256, 512, 768, 1024], layers=5,
                     blocks=[1, 1, 2, 2]),
    'V-57-eSE': dict(stem=[64, 64, 128], conv_ch=[128, 160, 192, 224], out_ch=[256, 512, 768, 1024], layers=5,
                     blocks=[1, 1, 4, 3]),
    'V-99-eSE': dict(stem=[64, 64, 128], conv_ch=[128, 160, 192, 224], out_ch=[256, 512, 768, 1024], layers=5,
                     blocks=[1, 3, 9, 3]),
}

PRECISIONS = ('fp16x3', 'fp16mx', 'fp16', 'fp32')


class Buf:
    """NHWC activation buffer [N,H,W,C] with the planes the chosen precision needs."""

    def __init__(self, N, H, W, C, device, precision, f32=False, lowp=True):
        self.N, self.H, self.W, self.C = N, H, W, C
        self.f32 = self.hi = self.lo = None
        # format of the lo plane: fp16 residual (0) or, in 'fp16mx', the e4m3 correction plane of the same size (needs whole
        # 32-channel groups; other widths keep the fp16 residual and their convs run the three-pass form)
        self.fmt = ops.lo_mx() if (precision == 'fp16mx' and C % 32 == 0) else ops.LO_FP16
        if precision == 'fp32' or f32:
            self.f32 = torch.empty(N, H, W, C, device=device, dtype=torch.float32)
        if precision != 'fp32' and lowp:
            self.hi = torch.empty(N, H, W, C, device=device, dtype=torch.float16)
            if precision in ('fp16x3', 'fp16mx'):
                self.lo = torch.empty(N, H, W, C, device=device, dtype=torch.float16)

    def nchw(self):
        t = self.f32.permute(0, 3, 1, 2)
        t._far3d_buf = self
        return t


def fold_bn(conv_w, bn):
    """eval-mode BatchNorm folded into the preceding bias-free conv (vovnet.py:124-161, :380-384)."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return conv_w * scale.view(-1, 1, 1, 1), bn.bias - bn.running_mean * scale


class PackedConv:
    """Weights of one conv in the kernel's layout [Cout, ky*kx, Cin] (+ split-fp16 copies), bias fp32."""

    def __init__(self, w, b, precision, stride=1):
        Cout, Cin, kh, kw = w.shape
        self.Cout, self.Cin, self.k, self.stride = Cout, Cin, kh, stride
        wk = w.detach().float().permute(0, 2, 3, 1).contiguous().view(Cout, kh * kw, Cin)
        self.bias = None if b is None else b.detach().float().contiguous()
        self.w_f32 = self.w_hi = self.w_lo = self.w_c8 = self.w_hi_mx = None
        self.w_exp = 0
        if precision == 'fp32':
            self.w_f32 = wk
        else:
            self.w_hi = wk.to(torch.float16)
            if precision in ('fp16x3', 'fp16mx'):
                self.w_lo = (wk - self.w_hi.float()).to(torch.float16)
            if precision == 'fp16mx' and Cin % 32 == 0:
                self.w_hi_mx, self.w_c8, self.w_exp = ops.pack_weight_mx(wk)

    def hi_for(self, x_fmt):
        """the weights' fp16 plane: pre-scaled by the correction products' common power of two in the fp16mx operand format"""
        return self.w_hi_mx if x_fmt != 0 else self.w_hi

    def lo_for(self, x_fmt):
        """the weights' second plane in the format of the activations' (fp16 residual / e4m3 correction)"""
        return self.w_c8 if x_fmt != 0 else self.w_lo


def run_conv(pc, precision, src, src_co, dst_f32=None, dst_f32_co=0, dst_b=None, dst_b_co=0, relu=True, f32_ns=0,
             f32_cs=None, f32_ptr=None):
    """Launch one conv.  src: Buf (reads channels [src_co, src_co+Cin)); dst_f32 / dst_b: Buf to receive fp32 / fp16 planes
    at the given channel offsets.  f32_ptr/f32_cs/f32_ns: raw fp32 destination with its own strides (flatten buffer)."""
    N, H, W = src.N, src.H, src.W
    if precision == 'fp32':
        y = f32_ptr if f32_ptr is not None else dst_f32.f32
        cs = f32_cs if f32_cs is not None else dst_f32.C
        assert f32_ns == 0
        ops.conv2d_f32(src.f32, N, H, W, src.C, src_co, pc.Cin, pc.w_f32, pc.bias, pc.Cout, pc.k, pc.stride, relu,
                       y, cs, dst_f32_co)
        return
    yf = f32_ptr if f32_ptr is not None else (dst_f32.f32 if dst_f32 is not None else None)
    yf_cs = f32_cs if f32_cs is not None else (dst_f32.C if dst_f32 is not None else 0)
    x_fmt = conv_in_fmt(pc, src, src_co)
    ops.conv2d_umma(src.hi, src.lo, N, H, W, src.C, src_co, pc.Cin, pc.hi_for(x_fmt), pc.lo_for(x_fmt), pc.bias, pc.Cout, pc.k,
                    pc.stride, relu, y_f32=yf, yf_cs=yf_cs, yf_co=dst_f32_co, yf_ns=f32_ns,
                    y_hi=dst_b.hi if dst_b is not None else None, y_lo=dst_b.lo if dst_b is not None else None,
                    yb_cs=dst_b.C if dst_b is not None else 0, yb_co=dst_b_co, x_fmt=x_fmt, w_exp=pc.w_exp,
                    y_fmt=dst_b.fmt if (dst_b is not None and dst_b.lo is not None) else 0)


def conv_in_fmt(pc, src, src_co):
    """operand format a conv reads `src` in: the buffer's e4m3 correction plane when the weights have one too"""
    if getattr(src, 'fmt', 0) != 0 and src.lo is not None:
        if pc.w_c8 is None or src_co % 32:
            raise RuntimeError('fp16mx: conv reads an e4m3 correction plane but its Cin / channel offset is not a multiple of 32')
        return src.fmt
    return 0


def _cbr(cin, cout, name, k, stride=1):
    return [(f'{name}/conv', nn.Conv2d(cin, cout, k, stride, k // 2, bias=False)),
            (f'{name}/norm', nn.BatchNorm2d(cout)), (f'{name}/relu', nn.ReLU(inplace=True))]


class PackedMixin:
    """Packed weights (`_packed`) and activation plans (`_plan`) hold raw device addresses of one device / dtype: anything
    that re-homes the parameters (`.to()`, `.cuda()`, `.half()`, `.float()`: all go through `_apply`) or reloads them drops
    both, and `_weights_key()` lets a caller detect edits that bypass `_apply` (`p.data = ...`, in-place updates)."""

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        self._plan = None
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._packed = None
        return super()._load_from_state_dict(*args, **kwargs)

    def invalidate(self):
        self._packed = None

    def _weights_key(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _check_packed(self):
        """drop the packs when any parameter / buffer was replaced or edited in place since they were made"""
        if self._packed is not None:
            k = self._weights_key()
            if self.__dict__.get('_packed_key') != k:
                self._packed = None
        if self._packed is None:
            self.__dict__['_packed_key'] = self._weights_key()


class eSEModule(nn.Module):            # parameter container, vovnet.py:173-185
    def __init__(self, c):
        super().__init__()
        self.fc = nn.Conv2d(c, c, 1)


class _OSA_module(nn.Module):          # parameter container, vovnet.py:188-238
    def __init__(self, cin, cmid, cout, nlayers, name, identity):
        super().__init__()
        self.identity, self.cin, self.cmid, self.cout, self.nlayers = identity, cin, cmid, cout, nlayers
        self.layers = nn.ModuleList()
        c = cin
        for i in range(nlayers):
            self.layers.append(nn.Sequential(OrderedDict(_cbr(c, cmid, f'{name}_{i}', 3))))
            c = cmid
        self.concat = nn.Sequential(OrderedDict(_cbr(cin + nlayers * cmid, cout, f'{name}_concat', 1)))
        self.ese = eSEModule(cout)


@BACKBONES.register_module()
class VoVNet(PackedMixin, nn.Module):
    """Same constructor / forward signature as vovnet.py:276-360.  `precision` is a far3d_b200 extension."""

    def __init__(self, spec_name, input_ch=3, out_features=None, frozen_stages=-1, norm_eval=True, pretrained=None,
                 init_cfg=None, precision='fp16x3'):
        super().__init__()
        assert precision in PRECISIONS
        sp = _SPECS[spec_name]
        self.spec, self.precision = sp, precision
        self.frozen_stages, self.norm_eval = frozen_stages, norm_eval
        self._out_features = tuple(out_features) if out_features is not None else ()
        st = sp['stem']
        self.stem = nn.Sequential(OrderedDict(_cbr(input_ch, st[0], 'stem_1', 3, 2) + _cbr(st[0], st[1], 'stem_2', 3, 1)
                                              + _cbr(st[1], st[2], 'stem_3', 3, 2)))
        cin = [st[2]] + sp['out_ch'][:-1]
        self.stage_names = []
        for i in range(4):
            stage = nn.Sequential()
            if i > 0:
                stage.add_module('Pooling', nn.MaxPool2d(3, 2, ceil_mode=True))
            for b in range(sp['blocks'][i]):
                name = f'OSA{i + 2}_{b + 1}'
                stage.add_module(name, _OSA_module(cin[i] if b == 0 else sp['out_ch'][i], sp['conv_ch'][i],
                                                   sp['out_ch'][i], sp['layers'], name, identity=b > 0))
            self.add_module(f'stage{i + 2}', stage)
            self.stage_names.append(f'stage{i + 2}')
        self._packed = None
        self._plan = None

    # -- weight packing (once per weight version: PackedMixin)
    def set_precision(self, precision):
        assert precision in PRECISIONS
        if precision != self.precision:
            self.precision, self._packed, self._plan = precision, None, None

    @torch.no_grad()
    def _pack(self):
        pr = self.precision
        pk = {}
        seq = self.stem
        w1, b1 = fold_bn(getattr(seq, 'stem_1/conv').weight, getattr(seq, 'stem_1/norm'))
        pk['stem1_w'] = w1.float().permute(0, 2, 3, 1).contiguous()           # (Cout, ky, kx, cin)
        pk['stem1_b'] = b1.float().contiguous()
        pk['stem2'] = PackedConv(*fold_bn(getattr(seq, 'stem_2/conv').weight, getattr(seq, 'stem_2/norm')), pr, 1)
        pk['stem3'] = PackedConv(*fold_bn(getattr(seq, 'stem_3/conv').weight, getattr(seq, 'stem_3/norm')), pr, 2)
        for sn in self.stage_names:
            for bname, blk in getattr(self, sn).named_children():
                if not isinstance(blk, _OSA_module):
                    continue
                convs = []
                for i, layer in enumerate(blk.layers):
                    convs.append(PackedConv(*fold_bn(getattr(layer, f'{bname}_{i}/conv').weight,
                                                     getattr(layer, f'{bname}_{i}/norm')), pr))
                cc = PackedConv(*fold_bn(getattr(blk.concat, f'{bname}_concat/conv').weight,
                                         getattr(blk.concat, f'{bname}_concat/norm')), pr)
                pk[bname] = dict(convs=convs, concat=cc,
                                 fc_w=blk.ese.fc.weight.detach().float().reshape(blk.cout, blk.cout).contiguous(),
                                 fc_b=blk.ese.fc.bias.detach().float().contiguous())
        self._packed = pk

    # -- buffer plan (once per input shape)
    def _make_plan(self, N, H, W, device):
        pr, sp = self.precision, self.spec
        plan = dict(key=(N, H, W, str(device), pr))
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        H4, W4 = (H2 + 1) // 2, (W2 + 1) // 2
        plan['s1'] = Buf(N, H2, W2, sp['stem'][0], device, pr)
        plan['s2'] = Buf(N, H2, W2, sp['stem'][1], device, pr)
        stages = []
        h, w = H4, W4
        cin = sp['stem'][2]
        for i in range(4):
            if i > 0:
                h, w = _pool_out(h), _pool_out(w)
            cmid, cout, nb, nl = sp['conv_ch'][i], sp['out_ch'][i], sp['blocks'][i], sp['layers']
            first = Buf(N, h, w, cin + nl * cmid, device, pr)
            rest = [Buf(N, h, w, cout + nl * cmid, device, pr) for _ in range(min(2, nb - 1))]
            xt = Buf(N, h, w, cout, device, 'fp32')
            out = Buf(N, h, w, cout, device, pr, f32=True, lowp=(i < 3) or True)
            mean = torch.empty(N, cout, device=device)
            gate = torch.empty(N, cout, device=device)
            ws = torch.empty(max(N * 64 * cout, ops.conv_pool_workspace_floats(N, h, w, cout)), device=device)
            stages.append(dict(h=h, w=w, cin=cin, first=first, rest=rest, xt=xt, out=out, mean=mean, gate=gate, ws=ws))
            cin = cout
        plan['stages'] = stages
        self._plan = plan

    @torch.no_grad()
    def forward(self, x):
        if self.training:
            raise RuntimeError('far3d_b200 VoVNet implements the inference forward only; call .eval()')
        if not x.is_cuda:
            raise RuntimeError('far3d_b200 VoVNet needs CUDA tensors (no CPU path)')
        x = x.contiguous().float()
        N, _, H, W = x.shape
        self._check_packed()
        if self._packed is None:
            self._pack()
        if self._plan is None or self._plan['key'] != (N, H, W, str(x.device), self.precision):
            self._make_plan(N, H, W, x.device)
        pr, pk, plan, sp = self.precision, self._packed, self._plan, self.spec
        s1, s2 = plan['s1'], plan['s2']
        ops.stem_conv(x, pk['stem1_w'], pk['stem1_b'], sp['stem'][0], y_f32=s1.f32, y_hi=s1.hi, y_lo=s1.lo, lo_fmt=s1.fmt)
        run_conv(pk['stem2'], pr, s1, 0, dst_f32=s2 if pr == 'fp32' else None, dst_b=s2 if pr != 'fp32' else None)
        st0 = plan['stages'][0]
        run_conv(pk['stem3'], pr, s2, 0, dst_f32=st0['first'] if pr == 'fp32' else None,
                 dst_b=st0['first'] if pr != 'fp32' else None)
        outs = []
        prev_out = None
        for i, sn in enumerate(self.stage_names):
            st = plan['stages'][i]
            cur = st['first']
            if i > 0:       # MaxPool2d(3, 2, ceil_mode=True), vovnet.py:249
                po = prev_out
                if pr == 'fp32':
                    ops.maxpool3x3s2(po.f32, None, 0, po.N, po.H, po.W, po.C, po.C, 0, cur.f32, None, cur.C, 0)
                else:
                    assert po.fmt == cur.fmt
                    ops.maxpool3x3s2(po.hi, po.lo, 1, po.N, po.H, po.W, po.C, po.C, 0, cur.hi, cur.lo, cur.C, 0, lo_fmt=cur.fmt)
            blocks = [(n, m) for n, m in getattr(self, sn).named_children() if isinstance(m, _OSA_module)]
            for bi, (bname, blk) in enumerate(blocks):
                pb = pk[bname]
                cin_b = blk.cin
                for li, pc in enumerate(pb['convs']):
                    src_co = 0 if li == 0 else cin_b + (li - 1) * blk.cmid
                    dst_co = cin_b + li * blk.cmid
                    run_conv(pc, pr, cur, src_co, dst_f32=cur if pr == 'fp32' else None, dst_f32_co=dst_co,
                             dst_b=cur if pr != 'fp32' else None, dst_b_co=dst_co)
                xt = st['xt']
                HW = st['h'] * st['w']
                pcc = pb['concat']
                if pr != 'fp32' and blk.cout % 8 == 0:
                    # concat conv with the eSE average pool folded into its epilogue (one HBM pass less)
                    x_fmt = conv_in_fmt(pcc, cur, 0)
                    ops.conv2d_umma_pool(cur.hi, cur.lo, N, cur.H, cur.W, cur.C, 0, pcc.Cin, pcc.hi_for(x_fmt), pcc.lo_for(x_fmt), pcc.bias,
                                         pcc.Cout, True, xt.f32, xt.C, 0, st['ws'], st['mean'], x_fmt=x_fmt, w_exp=pcc.w_exp)
                else:
                    run_conv(pcc, pr, cur, 0, dst_f32=xt)
                    ops.global_avgpool(xt.f32, st['mean'], st['ws'], N, HW, blk.cout)
                ops.ese_gate(st['mean'], pb['fc_w'], pb['fc_b'], st['gate'], N, blk.cout)
                last = bi == len(blocks) - 1
                nxt = st['out'] if last else st['rest'][bi % 2]
                ident = cur if blk.identity else None
                assert ident is None or pr == 'fp32' or ident.fmt == nxt.fmt
                ops.ese_apply(xt.f32, st['gate'],
                              ident.f32 if (ident is not None and pr == 'fp32') else None,
                              ident.hi if (ident is not None and pr != 'fp32') else None,
                              ident.lo if (ident is not None and pr != 'fp32') else None,
                              ident.C if ident is not None else 0, 0, N, HW, blk.cout,
                              nxt.f32, nxt.C, 0, nxt.hi, nxt.lo, nxt.C, 0, lo_fmt=nxt.fmt)
                cur = nxt
            prev_out = st['out']
            if sn in self._out_features:
                outs.append(st['out'].nchw())
        return outs


def _pool_out(h):
    o = (h - 3 + 1) // 2 + 1
    if (o - 1) * 2 >= h:
        o -= 1
    return o


class _ConvHolder(nn.Module):
    """mmcv ConvModule without norm / activation: a `.conv` child, keys `*.conv.{weight,bias}`."""

    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, k // 2)


@NECKS.register_module()
class FPN(PackedMixin, nn.Module):
    """mmdet FPN (2.28.2) for the configuration family far3d.py:50-57 uses: start_level, `add_extra_convs='on_output'`,
    nearest top-down upsampling, no norm, no activation."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode='nearest'), init_cfg=None, precision='fp16x3'):
        super().__init__()
        assert norm_cfg is None and act_cfg is None and conv_cfg is None
        assert end_level in (-1, len(in_channels) - 1)
        if add_extra_convs is True:
            add_extra_convs = 'on_input'
        assert add_extra_convs in (False, 'on_output'), 'only on_output extra convs are implemented'
        assert upsample_cfg.get('mode', 'nearest') == 'nearest'
        self.in_channels, self.out_channels, self.num_outs = list(in_channels), out_channels, num_outs
        self.start_level, self.add_extra_convs = start_level, add_extra_convs
        self.relu_before_extra_convs = relu_before_extra_convs
        self.nlvl = len(in_channels) - start_level
        assert num_outs >= self.nlvl
        self.lateral_convs = nn.ModuleList(_ConvHolder(c, out_channels, 1) for c in in_channels[start_level:])
        self.fpn_convs = nn.ModuleList(_ConvHolder(out_channels, out_channels, 3) for _ in range(self.nlvl))
        n_extra = num_outs - self.nlvl
        assert n_extra == 0 or add_extra_convs == 'on_output'
        assert n_extra <= 1, 'one extra level (far3d.py:57 num_outs=4) is implemented'
        for _ in range(n_extra):
            self.fpn_convs.append(_ConvHolder(out_channels, out_channels, 3, 2))
        self.precision = precision
        self._packed = None
        self._plan = None

    def set_precision(self, precision):
        assert precision in PRECISIONS
        if precision != self.precision:
            self.precision, self._packed, self._plan = precision, None, None

    @torch.no_grad()
    def forward(self, inputs):
        if self.training:
            raise RuntimeError('far3d_b200 FPN implements the inference forward only; call .eval()')
        pr = self.precision
        self._check_packed()
        if self._packed is None:
            self._packed = dict(
                lat=[PackedConv(m.conv.weight, m.conv.bias, pr) for m in self.lateral_convs],
                out=[PackedConv(m.conv.weight, m.conv.bias, pr, stride=m.conv.stride[0]) for m in self.fpn_convs])
        srcs = [_as_buf(inputs[i + self.start_level], pr) for i in range(self.nlvl)]
        N, dev = srcs[0].N, srcs[0].f32.device if srcs[0].f32 is not None else srcs[0].hi.device
        # the OUTPUT maps are per `plan_slot`: a pipelined caller replays slot 1's graph while slot 0's outputs are still
        # being read by the previous frame's head (the laterals and everything upstream are internal to the image branch).
        # Every slot's buffers stay referenced here: captured graphs hold raw addresses of them.
        key = (tuple((s.N, s.H, s.W, s.C) for s in srcs), str(dev), pr)
        slot = getattr(self, 'plan_slot', 0)
        if self._plan is None or self._plan['key'] != key:
            C = self.out_channels
            self._plan = dict(key=key, lat=[Buf(s.N, s.H, s.W, C, dev, pr, f32=True) for s in srcs], outs={})
        if slot not in self._plan['outs']:
            C = self.out_channels
            # outputs keep their low-precision planes too: the 2D head's towers and the extra conv consume them
            outs = [Buf(s.N, s.H, s.W, C, dev, pr, f32=True) for s in srcs]
            if self.num_outs > self.nlvl:
                s = srcs[-1]
                outs.append(Buf(s.N, (s.H + 1) // 2, (s.W + 1) // 2, C, dev, pr, f32=True))
            self._plan['outs'][slot] = outs
        lat, outs, pk = self._plan['lat'], self._plan['outs'][slot], self._packed
        for i in range(self.nlvl):
            # the top lateral is consumed as-is by its 3x3 conv, so it also needs its low-precision planes now; the
            # others get theirs from the top-down add
            top = i == self.nlvl - 1
            run_conv(pk['lat'][i], pr, srcs[i], 0, dst_f32=lat[i], dst_b=lat[i] if (top and pr != 'fp32') else None,
                     relu=False)
        for i in range(self.nlvl - 1, 0, -1):
            d, s = lat[i - 1], lat[i]
            ops.upsample_add(d.f32, s.f32, d.N, d.H, d.W, s.H, s.W, d.C, d.hi, d.lo, lo_fmt=d.fmt)
        for i in range(self.nlvl):
            o = outs[i]
            run_conv(pk['out'][i], pr, lat[i], 0, dst_f32=o, dst_b=o if (o.hi is not None and pr != 'fp32') else None,
                     relu=False)
        if self.num_outs > self.nlvl:
            o = outs[self.nlvl]
            run_conv(pk['out'][self.nlvl], pr, outs[self.nlvl - 1], 0, dst_f32=o, dst_b=o if pr != 'fp32' else None,
                     relu=False)
        return tuple(o.nchw() for o in outs)


def _as_buf(t, precision):
    """Accept a tensor produced by VoVNet (carries its NHWC Buf) or any (N,C,H,W) fp32 CUDA tensor."""
    b = getattr(t, '_far3d_buf', None)
    want_fmt = ops.lo_mx() if (precision == 'fp16mx' and t.shape[1] % 32 == 0) else ops.LO_FP16
    if b is not None and (precision == 'fp32' or (b.hi is not None and (precision == 'fp16' or
                                                                        (b.lo is not None and getattr(b, 'fmt', 0) == want_fmt)))):
        return b
    if not t.is_cuda:
        raise RuntimeError('far3d_b200 FPN needs CUDA tensors (no CPU path)')
    N, C, H, W = t.shape
    nb = Buf.__new__(Buf)
    nb.N, nb.H, nb.W, nb.C = N, H, W, C
    nb.f32 = t.permute(0, 2, 3, 1).contiguous().float()
    nb.hi = nb.lo = None
    nb.fmt = want_fmt
    if precision != 'fp32':
        if C % 8 == 0:
            nb.hi, nb.lo = ops.split_planes(nb.f32, lo_fmt=want_fmt, want_lo=precision in ('fp16x3', 'fp16mx'))
        else:
            nb.hi, nb.lo = ops.split_fp16(nb.f32, want_lo=precision in ('fp16x3', 'fp16mx'))
    return nb
