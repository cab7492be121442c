"""Host-side engine: weight packing (BN folding, tap lists, sub-pixel phases), workspace buffers and the
layer schedules that drive the C-ABI kernels.  PyTorch is used for device memory and one-off weight
repacking only; all per-step arithmetic runs in rpnet_b200/csrc kernels."""
import math

import torch

from . import ops

BN_EPS = 1e-5
# bumped by every optimizer step that updates the parameters through raw pointers (rpnet_b200.train): packed-weight
# caches key on it because such updates do not touch torch's tensor version counters
WEIGHTS_EPOCH = 0
TAPS_3X3 = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]

# Arithmetic of the encoder convs (yaml key `b200_precision`, default from the environment variable RPNET_PRECISION):
#   'split8' (default): x.w = hi.Wh + 2^-15 (lo8.Wh8 + x8.Wl8) - the main term on fp16 operands, the two first-order corrections on
#                      e4m3 operands at twice the MMA rate (include/rpnet_b200.h "fp8 corrections"): every encoder activation
#                      travels as an fp16 plane + a c8 plane.  Two units of tensor time per conv; logits within 1e-3 (rel-Linf)
#                      of the reference's fp32 forward, the same error as 'split' (a correction is 2^-11 of the product, its
#                      e4m3 rounding 2^-15);
#   'split'          : split-fp16 - hi + lo fp16 planes and Wh | Wl packs, hi.Wh + lo.Wh + hi.Wl in fp32 (three units);
#   'fp16'           : single-term fp16 operands (the 11-bit significand of a TF32 cuDNN conv): one unit of tensor time,
#                      logits 2e-3 .. 5e-3 from the fp32 reference on unsaturated fixtures.
PRECISIONS = ('split8', 'split', 'fp16')
# weight-pack levels (the `w_split` of ConvPack / ops.conv_split / the pack kernels): fp16 weights, Wh | Wl, Wh | fp8 corrections
W_FP16, W_SPLIT, W_C8 = 0, 1, 2


def is_split(p):
    """Activations of precision `p` travel as (hi, lo) plane pairs."""
    return p in ('split8', 'split', 'split-a')


def w_level(p):
    return {'split8': W_C8, 'split': W_SPLIT}.get(p, W_FP16)


def default_precision():
    import os
    p = os.environ.get('RPNET_PRECISION', 'split8')
    if p not in PRECISIONS:
        raise ValueError('RPNET_PRECISION=%r (expected one of %r)' % (p, PRECISIONS))
    return p


def precision_of(cfg):
    p = (cfg or {}).get('b200_precision') or default_precision()
    if p not in PRECISIONS:
        raise ValueError('b200_precision=%r (expected one of %r)' % (p, PRECISIONS))
    return p


def decoder_precision(p):
    """Arithmetic of the decoder half of the U-Net (Up5, Up_conv5, Up4, Up_conv4 — 59 % of the encoder FLOPs) under precision
    `p`.  RPNET_SPLIT_DECODER=2 keeps the split ACTIVATIONS there (hi + lo planes in and out) but uses plain fp16 weights: two
    tensor-core passes (hi.W + lo.W) instead of three.  Measured on B200 (cfg3 train step): 48.0 instead of 49.9 ms, logits
    5.0e-4 .. 5.6e-4 instead of 3.5e-4 .. 4.9e-4 from the fp32 oracle on the 8 x 256 x 256 eval fixture (margin error 7.6e-4
    against the 1e-3 gate) — the default stays the full three-term product everywhere (margin over speed)."""
    import os
    if p != 'split':                                   # 'split8' runs the same (two-unit) arithmetic everywhere
        return p
    return 'split-a' if os.environ.get('RPNET_SPLIT_DECODER', '3') == '2' else 'split'


class Workspace:
    """Named, shape-keyed persistent device buffers (stable pointers: CUDA-graph friendly).  A buffer set exists per distinct
    activation shape; owners bound the number of shapes they keep with ShapeBudget and call clear() when it is exceeded."""

    def __init__(self):
        self._bufs = {}

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self._bufs.values())

    def get(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=device)
            self._bufs[key] = t
        return t

    def clear(self):
        self._bufs.clear()


class ConvPack:
    """One tap-list conv ready for rpnet_conv_igemm_f16 (split: rpnet_conv_split_f16, wpack = Wh | Wl along cin)."""
    __slots__ = ('wpack', 'taps', 'scale', 'shift', 'relu', 'cout', 'cin', 'split', 'w_split')

    def __init__(self, wpack, taps, scale, shift, relu, split=False, w_split=None):
        # split: activations travel as hi + lo planes; w_split: weight-pack level (W_FP16 / W_SPLIT: Wl as well, three-term product /
        # W_C8: fp8 corrections, the lo planes are c8 planes)
        self.wpack, self.taps, self.scale, self.shift, self.relu, self.split = wpack, taps, scale, shift, relu, split
        self.w_split = int(split if w_split is None else w_split)
        self.cout, self.cin = wpack.shape[1], wpack.shape[2] // (2 if self.w_split else 1)


def split_f16(x):
    """fp32 tensor -> (hi, lo) fp16 pair with hi + lo == x to ~2^-22 relative."""
    hi = x.to(torch.float16)
    return hi, (x - hi.float()).to(torch.float16)


C8_LO_SCALE, C8_X_SCALE, C8_WH_SCALE, C8_WL_SCALE = 2.0 ** 9, 2.0 ** -2, 2.0 ** 6, 2.0 ** 17      # csrc/common.cuh, "c8"


def _e4m3(x):
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)


def _c8_interleave(a8, b8):
    """two uint8 tensors [..., C] -> [..., 2 * C]: per 64 channels, 64 bytes of a8 then 64 bytes of b8."""
    c = a8.shape[-1]
    assert c % 64 == 0, 'c8 planes / packs need a multiple of 64 channels (got %d)' % c
    lead = a8.shape[:-1]
    return torch.stack((a8.reshape(*lead, c // 64, 64), b8.reshape(*lead, c // 64, 64)), dim=-2).reshape(*lead, 2 * c).contiguous()


def split_c8(x):
    """fp32 NHWC tensor -> (hi fp16 [..., C], c8 plane uint8 [..., 2 * C]): per 64 channels lo8 = e4m3((x - hi) * 2^9) | x8 = e4m3(x / 4)."""
    hi = x.to(torch.float16)
    return hi, _c8_interleave(_e4m3((x - hi.float()) * C8_LO_SCALE), _e4m3(x * C8_X_SCALE))


def split_planes(x, level):
    return split_c8(x) if level == W_C8 else split_f16(x)


def c8_lo(plane):
    """fp32 residual x - hi held by a c8 plane [..., 2 * C] (to the 2^-4 relative precision of its e4m3 entries)."""
    c = plane.shape[-1] // 2
    lo8 = plane.reshape(*plane.shape[:-1], c // 64, 2, 64)[..., 0, :].reshape(*plane.shape[:-1], c).contiguous()
    return lo8.view(torch.float8_e4m3fn).float() / C8_LO_SCALE


def join_planes(hi, lo):
    """fp32 value of a (hi, lo) activation pair, whichever format the lo plane has."""
    if lo is None:
        return hi.float()
    return hi.float() + (c8_lo(lo) if lo.dtype == torch.uint8 else lo.float())


def lo_buffer(ws, name, shape, device, level):
    """Workspace buffer of the lo plane of an activation [n, h, w, C]: fp16 residual plane, or the c8 plane (uint8, 2 * C)."""
    if level == W_C8:
        return ws.get(name, tuple(shape[:-1]) + (2 * shape[-1],), torch.uint8, device)
    return ws.get(name, shape, torch.float16, device)


def split_weight(w32, level):
    """fp32 [..., cin] (K innermost) -> fp16 pack row [..., cin] / [..., 2 * cin] of weight-pack level `level`."""
    if level == W_FP16:
        return w32.to(torch.float16).contiguous()
    hi, lo = split_f16(w32)
    if level == W_SPLIT:
        return torch.cat((hi, lo), dim=-1).contiguous()
    corr = _c8_interleave(_e4m3(hi.float() * C8_WH_SCALE), _e4m3((w32 - hi.float()) * C8_WL_SCALE))
    return torch.cat((hi, corr.view(torch.float16)), dim=-1).contiguous()


def hi_of(a):
    """hi plane of an activation that is either a tensor or a (hi, lo) pair."""
    return a[0] if isinstance(a, tuple) else a


def lo_of(a):
    return a[1] if isinstance(a, tuple) else None


def fold_bn(conv_bias, bn_weight=None, bn_bias=None, running_mean=None, running_var=None, eps=BN_EPS):
    """Conv bias + eval-mode BatchNorm2d -> per-channel (scale, shift):  y = conv_nobias * scale + shift.
    nn.BatchNorm2d eval: y = (x - rm) / sqrt(rv + eps) * gamma + beta  (SURVEY Appendix B)."""
    bias = conv_bias.detach().float()
    if bn_weight is None:
        return torch.ones_like(bias).contiguous(), bias.clone().contiguous()
    scale = bn_weight.detach().float() / torch.sqrt(running_var.detach().float() + eps)
    shift = (bias - running_mean.detach().float()) * scale + bn_bias.detach().float()
    return scale.contiguous(), shift.contiguous()


def pack_weight_taps(weight, dilation=1, split=False):
    """[cout, cin, k, k] fp32 -> fp16 [k*k, cout, cin] (split = W_SPLIT: [k*k, cout, 2*cin] = Wh | Wl; W_C8: Wh | fp8 corrections)
    + tap offsets (cross-correlation, 'same' padding)."""
    cout, cin, kh, kw = weight.shape
    w32 = weight.detach().float().permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    w = split_weight(w32, int(split))
    taps = [((ky - kh // 2) * dilation, (kx - kw // 2) * dilation) for ky in range(kh) for kx in range(kw)]
    return w, taps


def pack_upsample_phases(weight, split=False):
    """nn.Upsample(scale_factor=2, nearest) followed by a 3x3/pad-1 conv == four 2x2 convs on the
    low-resolution input, one per output parity (py, px), with row/column-summed weights:
      py = 0: rows {-1: w[0], 0: w[1] + w[2]};   py = 1: rows {0: w[0] + w[1], +1: w[2]}   (same for columns).
    Returns [(wpack fp16 [4, cout, cin], taps, (py, px))] — 2.25x fewer MACs than the materialised form
    (net/modules.py:65-68)."""
    w = weight.detach().float()
    rows = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}
    out = []
    for py in (0, 1):
        for px in (0, 1):
            taps, mats = [], []
            for dy, kys in rows[py]:
                for dx, kxs in rows[px]:
                    m = sum(w[:, :, ky, kx] for ky in kys for kx in kxs)
                    taps.append((dy, dx))
                    mats.append(m)
            m = torch.stack(mats, dim=0)
            wp = split_weight(m, int(split))
            out.append((wp, taps, (py, px)))
    return out


def conv_bn_pack(conv, bn, relu=True, dilation=1, split=False, w_split=None, pad_cin_to=None):
    """pad_cin_to: zero-pad the input channels of the pack (a 65- / 129-channel conv of the mask_feature_map variants runs
    as 128 / 192 packed channels: the extra source carries the mask in its first channel)."""
    scale, shift = fold_bn(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps) if bn is not None \
        else fold_bn(conv.bias)
    w_split = split if w_split is None else w_split
    w = conv.weight.detach()
    if pad_cin_to is not None and pad_cin_to > w.shape[1]:
        wpad = torch.zeros(w.shape[0], pad_cin_to, w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
        wpad[:, :w.shape[1]] = w
        w = wpad
    wp, taps = pack_weight_taps(w, dilation, w_split)
    return ConvPack(wp, taps, scale, shift, relu, split, w_split)


def run_conv(pack, src0, ws, name, src1=None, want_out=True, want_pool=False, out_f32=False):
    """Run one packed conv.  src0 / src1: fp16 NHWC tensors or (hi, lo) pairs.  Returns (out, pooled) - fp16 NHWC tensors, or
    (hi, lo) pairs for a split pack (None when not requested) - or the fp32 output."""
    n, h, w, _ = hi_of(src0).shape
    dev = hi_of(src0).device
    f16 = torch.float16
    out = ws.get(name, (n, h, w, pack.cout), f16, dev) if (want_out and not out_f32) else None
    pool = ws.get(name + '.pool', (n, h // 2, w // 2, pack.cout), f16, dev) if want_pool else None
    o32 = ws.get(name + '.f32', (n, h, w, pack.cout), torch.float32, dev) if out_f32 else None
    if pack.split:
        out_lo = lo_buffer(ws, name + '.lo', (n, h, w, pack.cout), dev, pack.w_split) if out is not None else None
        pool_lo = lo_buffer(ws, name + '.pool.lo', (n, h // 2, w // 2, pack.cout), dev, pack.w_split) if want_pool else None
        ops.conv_split(hi_of(src0), pack.wpack, pack.taps, pack.scale, pack.shift, pack.relu, src0_lo=lo_of(src0),
                       src1=None if src1 is None else hi_of(src1), src1_lo=None if src1 is None else lo_of(src1), w_split=pack.w_split,
                       out=out, out_lo=out_lo, out_pool=pool, out_pool_lo=pool_lo, out_f32=o32)
        if out_f32:
            return o32
        return (None if out is None else (out, out_lo)), (None if pool is None else (pool, pool_lo))
    ops.conv_igemm(hi_of(src0), pack.wpack, pack.taps, pack.scale, pack.shift, pack.relu, src1=None if src1 is None else hi_of(src1),
                   out=out, out_pool=pool, out_f32=o32)
    if out_f32:
        return o32
    return out, pool


def run_upconv(phases, scale, shift, src, ws, name, split=False, w_split=None):
    """Sub-pixel form of up_conv: four phase convs scatter into the 2x resolution output."""
    n, h, w, _ = hi_of(src).shape
    cout = phases[0][0].shape[1]
    dev = hi_of(src).device
    out = ws.get(name, (n, 2 * h, 2 * w, cout), torch.float16, dev)
    if split:
        w_split = int(split if w_split is None else w_split)
        out_lo = lo_buffer(ws, name + '.lo', (n, 2 * h, 2 * w, cout), dev, w_split)
        for wp, taps, (py, px) in phases:
            ops.conv_split(hi_of(src), wp, taps, scale, shift, True, src0_lo=lo_of(src), w_split=w_split,
                           out=out, out_lo=out_lo, out_map=(2, py, 2, px))
        return out, out_lo
    for wp, taps, (py, px) in phases:
        ops.conv_igemm(hi_of(src), wp, taps, scale, shift, True, out=out, out_map=(2, py, 2, px))
    return out


class ShapeBudget:
    """Bounds the device memory held for past input shapes: every distinct shape key allocates its own workspace buffers (and
    CUDA-graph capture), so a long run over heterogeneous volumes / batch sizes would grow without bound.  note(key) returns True
    when `key` is new and the number of shapes seen since the last reset exceeds `limit` — the owner then drops its workspaces and
    graphs (they are re-created on demand)."""

    def __init__(self, limit=8):
        self.limit, self.seen = limit, set()

    def note(self, key):
        if key in self.seen:
            return False
        self.seen.add(key)
        if len(self.seen) > self.limit:
            self.seen = {key}
            return True
        return False


def nchw_f32_to_nhwc_f16(x):
    return x.detach().permute(0, 2, 3, 1).contiguous().to(torch.float16)


def nhwc_to_nchw_f32(x):
    if isinstance(x, tuple):                       # split pair (fp16 residual or c8 plane)
        x = join_planes(x[0], x[1])
    return x.permute(0, 3, 1, 2).float().contiguous()
