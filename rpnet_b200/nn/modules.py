"""conv_block / up_conv with the reference's constructor signatures and state_dict layout
(net/modules.py:42-75).  The nn.Sequential children are parameter containers only: forward() never calls
them — it runs the tcgen05 implicit-GEMM kernel through the C ABI (rpnet_b200.engine)."""
import torch
import torch.nn as nn

from .. import engine, ops

bn_momentum = 0.1
affine = True


def _check_eval(mod):
    if mod.training:
        raise NotImplementedError('%s.forward on its own runs in eval mode only: the train-mode kernels (batch-statistics BatchNorm + '
                                  'backward) are scheduled for the whole model by RP_Net.forward / rpnet_b200.train.TrainStep; call '
                                  '.eval() for a standalone block — the B200 path has no PyTorch fallback' % type(mod).__name__)


def _check_norm(normalization_type):
    """'BatchNorm2d' (yamls/example.yml:40) or 'InstanceNorm2d' — nn.InstanceNorm2d(C): no parameters, per-image statistics in
    train and eval mode; the U-Net then runs the two-pass schedule of rpnet_b200.train.EncoderEngine with one call group per image."""
    if normalization_type not in ('BatchNorm2d', 'InstanceNorm2d'):
        raise NotImplementedError("unet_normalize_type=%r: 'BatchNorm2d' and 'InstanceNorm2d' have B200 kernels" % (normalization_type,))


class _PackedModule(nn.Module):
    """Caches packed fp16 weights / folded BN; invalidated when any parameter or buffer changes."""

    def _signature(self):
        return (engine.WEIGHTS_EPOCH,) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _packs(self):
        sig = self._signature()
        if getattr(self, '_pack_sig', None) != sig:
            self._pack_cache = self._build_packs()
            self._pack_sig = sig
        return self._pack_cache

    def _workspace(self):
        if not hasattr(self, '_ws'):
            self._ws = engine.Workspace()
        return self._ws


def run_conv_any(conv, bn, x_nhwc_or_img, ws, name, pack, relu=True, **kw):
    """Dispatch one conv(+BN+ReLU): Cin in {1, 3} fp32 NCHW images go to the streaming first-conv kernel,
    everything else to the tensor-core implicit GEMM."""
    if pack is None or isinstance(pack, tuple):        # first conv, fp32 NCHW image input
        if isinstance(pack, tuple):
            scale, shift = pack                          # folded once per weight version (conv_block._build_packs)
        else:
            scale, shift = engine.fold_bn(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps) \
                if bn is not None else engine.fold_bn(conv.bias)
        n, _, h, w = x_nhwc_or_img.shape
        out = ws.get(name, (n, h, w, conv.out_channels), torch.float16, x_nhwc_or_img.device)
        out_lo = engine.lo_buffer(ws, name + '.lo', (n, h, w, conv.out_channels), x_nhwc_or_img.device, kw.get('w_level', engine.W_SPLIT)) \
            if kw.get('split') else None
        ops.conv3x3_first(x_nhwc_or_img.float().contiguous(), conv.weight.detach().float().contiguous(), scale, shift, relu, out,
                          out_lo=out_lo)
        return ((out, out_lo) if out_lo is not None else out), None
    return engine.run_conv(pack, x_nhwc_or_img, ws, name, **kw)


class conv_block(_PackedModule):
    def __init__(self, ch_in, ch_out, normalization_type, kernel=3, padding=1, precision=None):
        super(conv_block, self).__init__()
        _check_norm(normalization_type)
        pr = precision or engine.default_precision()                           # engine.PRECISIONS (+ 'split-a': fp16 weights)
        self.split, self.w_split = engine.is_split(pr), engine.w_level(pr)
        if kernel != 3 or padding != 1:
            raise NotImplementedError('conv_block: only kernel=3, padding=1 is used by RP-Net')
        self.inorm = normalization_type == 'InstanceNorm2d'
        self.ch_in = ch_in
        self.ch_out = ch_out
        self.conv = nn.Sequential(
            nn.Conv2d(ch_in, ch_out, kernel_size=kernel, stride=1, padding=padding, bias=True),
            getattr(nn, normalization_type)(ch_out),
            nn.ReLU(inplace=True),
            nn.Conv2d(ch_out, ch_out, kernel_size=kernel, stride=1, padding=padding, bias=True),
            getattr(nn, normalization_type)(ch_out),
            nn.ReLU(inplace=True)
        )

    def _build_packs(self):
        if self.inorm:
            raise NotImplementedError('conv_block with InstanceNorm2d runs inside U_Net (per-image statistics need the two-pass schedule)')
        if self.ch_in in (1, 2, 3):                                    # 2: image + mask channel (mask_feature_map: x)
            if self.ch_out != 64:
                raise NotImplementedError('conv_block: a %d-channel image input needs ch_out == 64' % self.ch_in)
            c, b = self.conv[0], self.conv[1]
            first = engine.fold_bn(c.bias, b.weight, b.bias, b.running_mean, b.running_var, b.eps)    # (scale, shift) tuple
        else:
            # 65 / 129 input channels (mask_feature_map: x2 / x3): packed as 128 / 192, the mask travels in a 64-channel extra source
            pad = (self.ch_in + 63) // 64 * 64 if self.ch_in % 64 else None
            first = engine.conv_bn_pack(self.conv[0], self.conv[1], split=self.split, w_split=self.w_split, pad_cin_to=pad)
        return first, engine.conv_bn_pack(self.conv[3], self.conv[4], split=self.split, w_split=self.w_split)

    def run_nhwc(self, x, ws, name, x1=None, want_out=True, want_pool=False):
        """x: fp16 NHWC (or the fp32 NCHW image for the first block); x1: optional second source (channel concat).
        Returns (out, pooled)."""
        _check_eval(self)
        p0, p1 = self._packs()
        if isinstance(p0, tuple):
            a, _ = run_conv_any(self.conv[0], self.conv[1], x, ws, name + '.0', p0, split=self.split, w_level=self.w_split)
        else:
            a, _ = engine.run_conv(p0, x, ws, name + '.0', src1=x1)
        return engine.run_conv(p1, a, ws, name + '.3', want_out=want_out, want_pool=want_pool)

    def forward(self, x):
        """NCHW fp32 in / out like the reference module (layout conversion at the boundary only)."""
        ws = self._workspace()
        xin = x if self.ch_in in (1, 3) else engine.nchw_f32_to_nhwc_f16(x)
        out, _ = self.run_nhwc(xin, ws, 'cb')
        return engine.nhwc_to_nchw_f32(out)


class up_conv(_PackedModule):
    def __init__(self, ch_in, ch_out, normalization_type, kernel=3, padding=1, precision=None):
        super(up_conv, self).__init__()
        _check_norm(normalization_type)
        self.inorm = normalization_type == 'InstanceNorm2d'
        pr = precision or engine.default_precision()
        self.split, self.w_split = engine.is_split(pr), engine.w_level(pr)
        if kernel != 3 or padding != 1:
            raise NotImplementedError('up_conv: only kernel=3, padding=1 is used by RP-Net')
        self.ch_in = ch_in
        self.ch_out = ch_out
        self.up = nn.Sequential(
            nn.Upsample(scale_factor=2),
            nn.Conv2d(ch_in, ch_out, kernel_size=kernel, stride=1, padding=padding, bias=True),
            getattr(nn, normalization_type)(ch_out),
            nn.ReLU(inplace=True)
        )

    def _build_packs(self):
        if self.inorm:
            raise NotImplementedError('up_conv with InstanceNorm2d runs inside U_Net (per-image statistics need the two-pass schedule)')
        conv, bn = self.up[1], self.up[2]
        scale, shift = engine.fold_bn(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
        return engine.pack_upsample_phases(conv.weight, self.w_split), scale, shift

    def run_nhwc(self, x, ws, name):
        _check_eval(self)
        phases, scale, shift = self._packs()
        return engine.run_upconv(phases, scale, shift, x, ws, name, self.split, self.w_split)

    def forward(self, x):
        out = self.run_nhwc(engine.nchw_f32_to_nhwc_f16(x), self._workspace(), 'uc')
        return engine.nhwc_to_nchw_f32(out)
