"""GPU parity tests of the "fp16mx" operand format (round 2): fp16 main term + both correction terms as one e4m3
(tcgen05 kind::f8f6f4; the fp16 weight plane pre-scaled by the products' common power of two) MMA stream, 2 tensor-pipe passes
per MAC instead of the 3 of fp16x3.

Two bars per convolution:
  * EXACTNESS of the implementation: against an fp64 convolution of the SAME rounded operands (the format's definition in
    csrc/common.cuh restated with torch's float8_e4m3fn on the CPU: tests/mx_emulation.py) - 2e-5, i.e. only accumulation order;
    this is the bar that catches a wrong byte layout or a wrong power of two;
  * ACCURACY of the format: against the fp64 convolution of the unrounded operands - 2e-4 per layer (operands carry
    ~2^-15 each instead of fp16x3's 2^-22).
The planes themselves are checked byte for byte against the emulation."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err, rel_l2
from mx_emulation import conv_mx_reference, mx_planes, mx_unpack, planes_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(cuda, lib_built):
    from far3d_b200 import ops as _ops
    return _ops


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _pack_w(w):
    Cout, Cin, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).contiguous().view(Cout, kh * kw, Cin)


@pytest.fixture(params=[1, 2], ids=['cta', 'ctapair'])
def cta_group(request, ops):
    ops.conv_umma_tune4(request.param)
    yield request.param
    ops.conv_umma_tune4(0)


@pytest.mark.parametrize('ea', [0, 2, -3])
def test_mx_planes_byte_exact_and_roundtrip(ops, cuda, ea):
    g = torch.Generator().manual_seed(5 + ea)
    x = torch.randn(3, 7, 5, 96, generator=g) * torch.logspace(-3, 2, 96)        # channel magnitudes from 1e-3 to 1e2
    x[0, 0, 0, :8] = torch.tensor([0.0, -0.0, 500.0, -1000.0, 1e-6, 65000.0, 2 ** -9, 3.1415926])
    fmt = ops.lo_mx(ea)
    hi, c8 = ops.split_planes(x.to(cuda), lo_fmt=fmt)
    ref_hi, ref_c8 = mx_planes(x, ea)
    assert torch.equal(hi.cpu(), ref_hi)
    assert planes_equal(c8.cpu().view(torch.uint8), ref_c8), 'correction plane differs from the format definition'
    back = ops.merge_fp16_strided(hi, c8, 96, 0, 3 * 7 * 5, 96, lo_fmt=fmt).view(3, 7, 5, 96)
    lo, _ = mx_unpack(ref_c8, ea)
    assert torch.equal(back.cpu(), ref_hi.float() + lo)
    ok = (x.abs() * 2.0 ** ea < 400) & (x.abs() * 2.0 ** ea > 2.0 ** -5)
    assert float(((back.cpu() - x).abs() / x.abs().clamp_min(1e-30))[ok].max()) < 2.0 ** -15


MX_CASES = [
    # N, H, W, Cin, Cout, k, stride, extra channels in the input buffer, channel offset of the slice read
    (2, 24, 40, 128, 128, 3, 1, 0, 0),         # halo kernel
    (2, 32, 48, 160, 160, 3, 1, 96, 32),       # halo, slice at a 32-channel offset, K tail of half a chunk
    (1, 20, 30, 224, 224, 3, 1, 0, 0),         # 224-column N tile
    (1, 40, 60, 768, 192, 3, 1, 0, 0),         # 12 chunks: both rings wrap
    (1, 16, 24, 1056, 512, 1, 1, 32, 0),       # concat 1x1: N tiles of 256 (split further: few work items), half-chunk K tail, padded row stride
    (2, 20, 30, 256, 256, 1, 1, 0, 0),         # Cout 256: one 256-column accumulator stage (round 3: no scale-factor columns)
    (7, 40, 60, 256, 256, 3, 1, 0, 0),         # halo kernel, N tile 256, 140 M tiles: both accumulator stages of 256 columns in use
    (4, 80, 120, 512, 512, 1, 1, 0, 0),        # 1x1, two N tiles of 256 per M tile, several tiles per worker
    (2, 160, 240, 64, 64, 3, 1, 0, 0),         # stem conv 2 shape: one chunk, weights resident in the B ring over 4 tiles per worker
    (1, 32, 48, 64, 128, 3, 2, 0, 0),          # stride 2 (5-D tensor map)
    (2, 20, 30, 256, 256, 3, 2, 0, 0),         # FPN extra conv
    (3, 24, 40, 192, 96, 3, 1, 0, 0),          # Cout 96: one N tile, three 32-channel groups
    (1, 8, 12, 256, 26, 1, 1, 0, 0),           # predictor: fp32 output only, Cout not a multiple of 8
]


@pytest.mark.parametrize('case', MX_CASES)
def test_conv_umma_mx_vs_emulation(ops, cuda, case, cta_group):
    N, H, W, Cin, Cout, k, s, extra, co = case
    g = torch.Generator().manual_seed(Cin + Cout + 7)
    cs = Cin + extra
    xfull = torch.randn(N, cs, H, W, generator=g).relu() * 1.7 + 0.05 * torch.randn(N, cs, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    exact = _nhwc(F.relu(F.conv2d(xfull[:, co:co + Cin].double(), w.double(), b.double(), stride=s, padding=k // 2)))
    fmt = ops.lo_mx()
    x_hi, x_c8 = ops.split_planes(_nhwc(xfull).to(cuda), lo_fmt=fmt)
    w_hi, w_c8, w_exp = ops.pack_weight_mx(_pack_w(w).to(cuda))
    emul = _nhwc(F.relu(conv_mx_reference(xfull[:, co:co + Cin], w, b, s, ops.MX_EA, w_exp)))
    Ho, Wo = exact.shape[1:3]
    planes = Cout % 32 == 0
    fcs = (Cout + 3) // 4 * 4 + 8
    bcs = Cout + 64
    yf = torch.zeros(N, Ho, Wo, fcs, device=cuda)
    yh = torch.zeros(N, Ho, Wo, bcs, device=cuda, dtype=torch.float16) if planes else None
    yc = torch.zeros_like(yh) if planes else None
    ops.conv2d_umma(x_hi, x_c8, N, H, W, cs, co, Cin, w_hi, w_c8, b.to(cuda), Cout, k, s, 1, y_f32=yf, yf_cs=fcs, yf_co=4,
                    y_hi=yh, y_lo=yc, yb_cs=bcs, yb_co=32, x_fmt=fmt, w_exp=w_exp, y_fmt=fmt if planes else 0)
    torch.cuda.synchronize()
    out = yf[..., 4:4 + Cout]
    assert rel_err(out, emul) < 2e-5, ('kernel vs same-operand emulation', rel_err(out, emul))
    assert rel_err(out, exact) < 2e-4 and rel_l2(out, exact) < 1e-4, ('format accuracy', rel_err(out, exact), rel_l2(out, exact))
    assert float(yf[..., :4].abs().max()) == 0 and float(yf[..., 4 + Cout:].abs().max()) == 0
    if planes:
        # the planes the epilogue wrote are the format's planes of the fp32 output it wrote
        ref_hi, ref_c8 = mx_planes(out.cpu(), ops.MX_EA)
        assert torch.equal(yh[..., 32:32 + Cout].cpu(), ref_hi)
        got = yc.cpu().view(torch.uint8).view(N, Ho, Wo, bcs * 2)[..., 64:64 + 2 * Cout]
        assert planes_equal(got, ref_c8)
        assert float(yh[..., :32].float().abs().max()) == 0 and float(yh[..., 32 + Cout:].float().abs().max()) == 0
        assert int(yc.cpu().view(torch.uint8).view(N, Ho, Wo, bcs * 2)[..., :64].max()) == 0


def test_conv_fp16x3_input_writes_mx_planes(ops, cuda):
    """a conv that READS fp16x3 operands may WRITE an e4m3 correction plane (buffers with odd widths feed fp16mx buffers)"""
    g = torch.Generator().manual_seed(3)
    N, H, W, Cin, Cout = 1, 16, 24, 48, 64
    x, w, b = torch.randn(N, Cin, H, W, generator=g), torch.randn(Cout, Cin, 3, 3, generator=g) / 20, torch.randn(Cout, generator=g)
    ref = _nhwc(F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1)))
    x_hi, x_lo = ops.split_fp16(_nhwc(x).to(cuda))
    w_hi, w_lo = ops.split_fp16(_pack_w(w).to(cuda))
    fmt = ops.lo_mx()
    yf = torch.zeros(N, H, W, Cout, device=cuda)
    yh = torch.zeros(N, H, W, Cout, device=cuda, dtype=torch.float16); yc = torch.zeros_like(yh)
    ops.conv2d_umma(x_hi, x_lo, N, H, W, Cin, 0, Cin, w_hi, w_lo, b.to(cuda), Cout, 3, 1, 1, y_f32=yf, yf_cs=Cout, y_hi=yh,
                    y_lo=yc, yb_cs=Cout, y_fmt=fmt)
    assert rel_err(yf, ref) < 2e-5
    ref_hi, ref_c8 = mx_planes(yf.cpu(), ops.MX_EA)
    assert torch.equal(yh.cpu(), ref_hi) and planes_equal(yc.cpu().view(torch.uint8).view(N, H, W, 2 * Cout), ref_c8)


def test_mx_helper_kernels(ops, cuda):
    """max-pool, eSE apply, FPN top-down add, GroupNorm and the stem conv read / write e4m3 correction planes"""
    g = torch.Generator().manual_seed(11)
    fmt, ea = ops.lo_mx(), ops.MX_EA
    # stem
    img = torch.randn(2, 3, 32, 48, generator=g)
    w, b = torch.randn(64, 3, 3, 3, generator=g) / 5, torch.randn(64, generator=g)
    yf = torch.empty(2, 16, 24, 64, device=cuda); yh = torch.empty_like(yf, dtype=torch.float16); yc = torch.empty_like(yh)
    ops.stem_conv(img.to(cuda), w.permute(0, 2, 3, 1).contiguous().to(cuda), b.to(cuda), 64, yf, yh, yc, lo_fmt=fmt)
    rh, rc = mx_planes(yf.cpu(), ea)
    assert torch.equal(yh.cpu(), rh) and planes_equal(yc.cpu().view(torch.uint8).view(2, 16, 24, 128), rc)
    # max-pool on a 32-channel-offset slice into a 32-channel-offset slice
    x = torch.randn(2, 21, 31, 96, generator=g)
    xh, xc = ops.split_planes(x.to(cuda), lo_fmt=fmt)
    xv = mx_unpack(mx_planes(x, ea)[1], ea)[0] + mx_planes(x, ea)[0].float()            # the values the planes carry
    refp = F.max_pool2d(xv[..., 32:96].permute(0, 3, 1, 2), 3, 2, ceil_mode=True).permute(0, 2, 3, 1)
    Ho, Wo = refp.shape[1:3]
    ph = torch.zeros(2, Ho, Wo, 128, device=cuda, dtype=torch.float16); pc = torch.zeros_like(ph)
    ops.maxpool3x3s2(xh, xc, 1, 2, 21, 31, 64, 96, 32, ph, pc, 128, 64, lo_fmt=fmt)
    got = ops.merge_fp16_strided(ph, pc, 128, 64, 2 * Ho * Wo, 64, lo_fmt=fmt).view(2, Ho, Wo, 64)
    assert torch.equal(got.cpu(), refp.contiguous())
    assert int(pc.cpu().view(torch.uint8).view(2, Ho, Wo, 256)[..., :128].max()) == 0
    # eSE apply with an identity read from correction planes
    xt = torch.randn(2, 96, 256, generator=g); gate = torch.rand(2, 256, generator=g); ident = torch.randn(2, 96, 256, generator=g)
    ih, ic = ops.split_planes(ident.to(cuda), lo_fmt=fmt)
    idv = mx_planes(ident, ea)[0].float() + mx_unpack(mx_planes(ident, ea)[1], ea)[0]
    y_ref = xt * gate[:, None] + idv
    yf = torch.empty(2, 96, 256, device=cuda); yh = torch.empty(2, 96, 256, device=cuda, dtype=torch.float16); yc = torch.empty_like(yh)
    ops.ese_apply(xt.to(cuda), gate.to(cuda), None, ih, ic, 256, 0, 2, 96, 256, yf, 256, 0, yh, yc, 256, 0, lo_fmt=fmt)
    assert rel_err(yf, y_ref) < 1e-6
    rh, rc = mx_planes(yf.cpu(), ea)
    assert torch.equal(yh.cpu(), rh) and planes_equal(yc.cpu().view(torch.uint8).view(2, 96, 512), rc)
    # FPN top-down
    d, s = torch.randn(2, 8, 12, 64, generator=g), torch.randn(2, 4, 6, 64, generator=g)
    dd = d.to(cuda); dh = torch.empty(2, 8, 12, 64, device=cuda, dtype=torch.float16); dc = torch.empty_like(dh)
    ops.upsample_add(dd, s.to(cuda), 2, 8, 12, 4, 6, 64, dh, dc, lo_fmt=fmt)
    rh, rc = mx_planes(dd.cpu(), ea)
    assert torch.equal(dh.cpu(), rh) and planes_equal(dc.cpu().view(torch.uint8).view(2, 8, 12, 128), rc)
    # GroupNorm
    x = torch.randn(2, 50, 256, generator=g) * 2 + 1
    gw, gb = torch.rand(256, generator=g) + 0.5, torch.randn(256, generator=g)
    y = torch.empty(2, 50, 256, device=cuda); yh = torch.empty(2, 50, 256, device=cuda, dtype=torch.float16); yc = torch.empty_like(yh)
    ops.groupnorm_nhwc(x.to(cuda), gw.to(cuda), gb.to(cuda), 2, 50, 256, 32, 1e-5, True, y_f32=y, y_hi=yh, y_lo=yc, lo_fmt=fmt)
    rh, rc = mx_planes(y.cpu(), ea)
    assert torch.equal(yh.cpu(), rh) and planes_equal(yc.cpu().view(torch.uint8).view(2, 50, 512), rc)


def test_mx_rejects_unaligned(ops, cuda):
    from far3d_b200 import _lib
    x = torch.zeros(1, 8, 8, 48, device=cuda)
    with pytest.raises(_lib.Far3DNativeError, match='32'):
        ops.split_planes(x, lo_fmt=ops.lo_mx())
