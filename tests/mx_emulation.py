"""CPU restatement of the "fp16mx" operand format (far3d_b200/csrc/common.cuh, include/far3d_b200.h) with torch's
float8_e4m3fn (round to nearest even; clamped to +-448 first = the hardware's satfinite): test infrastructure only."""
import torch
import torch.nn.functional as F


def e4m3(x):
    return x.float().clamp(-448.0, 448.0).to(torch.float8_e4m3fn)


def mx_planes(x, ea):
    """x fp32 [..., C], C % 32 == 0 -> (hi fp16 [..., C], correction plane uint8 [..., 2C]: per 32 channels [lo8 x32 | hi8 x32])"""
    C = x.shape[-1]
    assert C % 32 == 0
    hi = x.half()
    lo8 = e4m3((x - hi.float()) * 2.0 ** (11 + ea)).view(torch.uint8)
    hi8 = e4m3(hi.float() * 2.0 ** ea).view(torch.uint8)
    lead = x.shape[:-1]
    c8 = torch.stack([lo8.view(*lead, C // 32, 32), hi8.view(*lead, C // 32, 32)], dim=-2)
    return hi, c8.reshape(*lead, 2 * C)


def mx_decode(c8):
    """correction plane bytes [..., 2C] -> raw e4m3 values (lo8 [..., C], hi8 [..., C]) as fp32"""
    C2 = c8.shape[-1]
    lead = c8.shape[:-1]
    v = c8.contiguous().view(torch.float8_e4m3fn).float().view(*lead, C2 // 64, 2, 32)
    return v[..., 0, :].reshape(*lead, C2 // 2), v[..., 1, :].reshape(*lead, C2 // 2)


def mx_unpack(c8, ea):
    """-> (residual values lo8 * 2^-(11+ea), coarse values hi8 * 2^-ea)"""
    lo8, hi8 = mx_decode(c8)
    return lo8 * 2.0 ** -(11 + ea), hi8 * 2.0 ** -ea


def planes_equal(c8_a, c8_b):
    """same correction planes up to the sign of zero"""
    la, ha = mx_decode(c8_a)
    lb, hb = mx_decode(c8_b)
    return torch.equal(la, lb) and torch.equal(ha, hb)


def conv_mx_reference(x, w, b, stride, ea, w_exp):
    """what the fp16mx conv computes, in fp64: hi*w_hi + lo8*w_hi8 + hi8*w_lo8 with the operands rounded as the format says.
    x [N,Cin,H,W] fp32, w [Cout,Cin,k,k] fp32."""
    k = w.shape[-1]
    a_hi = x.half().float()
    a_lo = e4m3((x - a_hi) * 2.0 ** (11 + ea)).float() * 2.0 ** -(11 + ea)
    a_h8 = e4m3(a_hi * 2.0 ** ea).float() * 2.0 ** -ea
    q = 11 + ea + w_exp                                    # the fp16 weight plane is stored pre-multiplied by 2^q (ops.pack_weight_mx)
    w_hi = (w * 2.0 ** q).half().float() * 2.0 ** -q
    w_h8 = e4m3(w_hi * 2.0 ** w_exp).float() * 2.0 ** -w_exp
    w_l8 = e4m3((w - w_hi) * 2.0 ** (w_exp + 11)).float() * 2.0 ** -(w_exp + 11)
    conv = lambda a, ww: F.conv2d(a.double(), ww.double(), None, stride=stride, padding=k // 2)
    y = conv(a_hi, w_hi) + conv(a_lo, w_h8) + conv(a_h8, w_l8)
    return y + b.double().view(1, -1, 1, 1) if b is not None else y
