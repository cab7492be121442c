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


class Workspace:
    """Named, shape-keyed persistent device buffers (stable pointers: CUDA-graph friendly)."""

    def __init__(self):
        self._bufs = {}

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
    """One tap-list conv ready for rpnet_conv_igemm_f16."""
    __slots__ = ('wpack', 'taps', 'scale', 'shift', 'relu', 'cout', 'cin')

    def __init__(self, wpack, taps, scale, shift, relu):
        self.wpack, self.taps, self.scale, self.shift, self.relu = wpack, taps, scale, shift, relu
        self.cout, self.cin = wpack.shape[1], wpack.shape[2]


def fold_bn(conv_bias, bn_weight=None, bn_bias=None, running_mean=None, running_var=None, eps=BN_EPS):
    """Conv bias + eval-mode BatchNorm2d -> per-channel (scale, shift):  y = conv_nobias * scale + shift.
    nn.BatchNorm2d eval: y = (x - rm) / sqrt(rv + eps) * gamma + beta  (SURVEY Appendix B)."""
    bias = conv_bias.detach().float()
    if bn_weight is None:
        return torch.ones_like(bias).contiguous(), bias.clone().contiguous()
    scale = bn_weight.detach().float() / torch.sqrt(running_var.detach().float() + eps)
    shift = (bias - running_mean.detach().float()) * scale + bn_bias.detach().float()
    return scale.contiguous(), shift.contiguous()


def pack_weight_taps(weight, dilation=1):
    """[cout, cin, k, k] fp32 -> fp16 [k*k, cout, cin] + tap offsets (cross-correlation, 'same' padding)."""
    cout, cin, kh, kw = weight.shape
    w = weight.detach().permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).to(torch.float16).contiguous()
    taps = [((ky - kh // 2) * dilation, (kx - kw // 2) * dilation) for ky in range(kh) for kx in range(kw)]
    return w, taps


def pack_upsample_phases(weight):
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
            wp = torch.stack(mats, dim=0).to(torch.float16).contiguous()
            out.append((wp, taps, (py, px)))
    return out


def conv_bn_pack(conv, bn, relu=True, dilation=1):
    scale, shift = fold_bn(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps) if bn is not None \
        else fold_bn(conv.bias)
    wp, taps = pack_weight_taps(conv.weight, dilation)
    return ConvPack(wp, taps, scale, shift, relu)


def run_conv(pack, src0, ws, name, src1=None, want_out=True, want_pool=False, out_f32=False):
    """Run one packed conv.  Returns (out, pooled) fp16 NHWC tensors (None when not requested) or the fp32 output."""
    n, h, w, _ = src0.shape
    dev = src0.device
    out = ws.get(name, (n, h, w, pack.cout), torch.float16, dev) if (want_out and not out_f32) else None
    pool = ws.get(name + '.pool', (n, h // 2, w // 2, pack.cout), torch.float16, dev) if want_pool else None
    o32 = ws.get(name + '.f32', (n, h, w, pack.cout), torch.float32, dev) if out_f32 else None
    ops.conv_igemm(src0, pack.wpack, pack.taps, pack.scale, pack.shift, pack.relu, src1=src1, out=out, out_pool=pool,
                   out_f32=o32)
    if out_f32:
        return o32
    return out, pool


def run_upconv(phases, scale, shift, src, ws, name):
    """Sub-pixel form of up_conv: four phase convs scatter into the 2x resolution output."""
    n, h, w, _ = src.shape
    cout = phases[0][0].shape[1]
    out = ws.get(name, (n, 2 * h, 2 * w, cout), torch.float16, src.device)
    for wp, taps, (py, px) in phases:
        ops.conv_igemm(src, wp, taps, scale, shift, True, out=out, out_map=(2, py, 2, px))
    return out


def nchw_f32_to_nhwc_f16(x):
    return x.detach().permute(0, 2, 3, 1).contiguous().to(torch.float16)


def nhwc_to_nchw_f32(x):
    return x.permute(0, 3, 1, 2).float().contiguous()
