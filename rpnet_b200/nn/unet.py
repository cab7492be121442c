"""Truncated U-Net encoder with the reference's constructor / forward signature and state_dict keys
(net/unet.py:351-467: Unet_2D base + U_Net).  16 convs -> 'd4' (256 channels at H/4)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import engine
from .modules import conv_block, up_conv, _PackedModule


class Unet_2D(nn.Module):
    """net/unet.py:351-390 (constructor state only; the loss helpers there are unused by RP-Net)."""

    def __init__(self, cfg, img_ch=5, output_ch=6, t=2, pretrained=True, resnet_type='resnet18'):
        super(Unet_2D, self).__init__()
        self.cfg = cfg
        self.img_ch = img_ch
        self.t = t
        self.pretrained = pretrained
        self.resnet_type = resnet_type
        self.final_activation = cfg['final_activation']
        self.output_ch = output_ch

    def set_mode(self, mode):
        assert mode in ['train', 'valid', 'eval', 'test']
        self.mode = mode
        if mode in ['train']:
            self.train()
        else:
            self.eval()


class U_Net(Unet_2D):
    def __init__(self, cfg, img_ch=1, output_ch=6, resnet_type=None):
        super().__init__(cfg, img_ch, output_ch)
        # yamls/example.yml:103 `mask_feature_map: no` -> False; 'x' / 'x2' / 'x3' concatenate the (avg-pooled) support
        # mask to the input of Conv1 / Conv2 / Conv3 (net/unet.py:401-414, 437-449).  'x4' / 'x5' exist in the reference constructor
        # only: its forward never concatenates there and fails on the channel count.
        self.mfm = cfg['mask_feature_map'] or False
        if self.mfm not in (False, 'x', 'x2', 'x3'):
            raise NotImplementedError('mask_feature_map=%r: the reference forward handles x, x2, x3 (net/unet.py:437-449)' % (self.mfm,))
        self.Maxpool = nn.MaxPool2d(kernel_size=2, stride=2)     # fused into the preceding conv's epilogue
        num_feats = [64, 128, 256, 512, 1024]
        norm = cfg['unet_normalize_type']
        self.inorm = norm == 'InstanceNorm2d'
        pr = engine.precision_of(cfg)          # 'split8' (fp32-class, default), 'split' or 'fp16' (engine.PRECISIONS)
        pd = engine.decoder_precision(pr)      # decoder half: the same, or fp16 weights with RPNET_SPLIT_DECODER=2 (engine.decoder_precision)
        self.Conv1 = conv_block(ch_in=self.img_ch + (1 if self.mfm == 'x' else 0), ch_out=num_feats[0], normalization_type=norm, precision=pr)
        self.Conv2 = conv_block(ch_in=num_feats[0] + (1 if self.mfm == 'x2' else 0), ch_out=num_feats[1], normalization_type=norm, precision=pr)
        self.Conv3 = conv_block(ch_in=num_feats[1] + (1 if self.mfm == 'x3' else 0), ch_out=num_feats[2], normalization_type=norm, precision=pr)
        self.Conv4 = conv_block(ch_in=num_feats[2], ch_out=num_feats[3], normalization_type=norm, precision=pr)
        self.Conv5 = conv_block(ch_in=num_feats[3], ch_out=num_feats[4], normalization_type=norm, precision=pr)
        self.Up5 = up_conv(ch_in=num_feats[4], ch_out=num_feats[3], normalization_type=norm, precision=pd)
        self.Up_conv5 = conv_block(ch_in=num_feats[3] * 2, ch_out=num_feats[3], normalization_type=norm, precision=pd)
        self.Up4 = up_conv(ch_in=num_feats[3], ch_out=num_feats[2], normalization_type=norm, precision=pd)
        self.Up_conv4 = conv_block(ch_in=num_feats[2] * 2, ch_out=num_feats[2], normalization_type=norm, precision=pd)
        self._ws = engine.Workspace()

    def _mask_source(self, mask, pool, ws, name, split):
        """The extra conv source of mask_feature_map x2 / x3: avg_pool2d(mask, pool) in channel 0 of a 64-channel fp16 NHWC tensor
        (the packed weights are zero for the 63 padding channels).  Quarter / sixteenth steps are exact in fp16: the residual is 0
        (a c8 plane still carries the e4m3 copy of the mask for the x8 . Wl8 correction)."""
        n, _, H, W = mask.shape
        m = ws.get(name, (n, H // pool, W // pool, 64), torch.float16, mask.device)
        m.zero_()
        m[..., 0] = F.avg_pool2d(mask.float(), pool)[:, 0].to(torch.float16)
        if not split:
            return m
        lo = engine.lo_buffer(ws, name + '.lo', tuple(m.shape), mask.device, self.Conv1.w_split)
        lo.copy_(engine.split_planes(m.float(), self.Conv1.w_split)[1])
        return m, lo

    def encode_nhwc(self, x, tag='enc', mask=None):
        """x: fp32 NCHW image batch [n, 1, H, W] (H, W multiples of 16) -> d4 fp16 NHWC [n, H/4, W/4, 256].
        Schedule (net/unet.py:435-467): the 2x2 max-pools are fused into the epilogue of the conv that feeds
        them (x1 / x2 are never written at full resolution), nearest-upsample + conv runs in sub-pixel form and
        the skip concatenations are two-source K loops.  mask [n, 1, H, W]: only read by the mask_feature_map variants."""
        n, c, H, W = x.shape
        if H % 16 or W % 16:
            raise ValueError('U_Net input must be a multiple of 16 pixels (got %d x %d)' % (H, W))
        if self.mfm and (mask is None or mask.shape[0] != n):
            raise ValueError('mask_feature_map=%r needs one mask per image (got %s for %d images)'
                             % (self.mfm, None if mask is None else tuple(mask.shape), n))
        if self.inorm:
            from ..train import EncoderEngine
            eng = self.__dict__.get('_in_engine')
            if eng is None or not eng.attached():
                eng = EncoderEngine(self, norm='instance')
                self.__dict__['_in_engine'] = eng
            return eng.encode(x.float().contiguous(), mask if self.mfm else None)
        ws = self._ws
        split = self.Conv1.split
        if self.mfm == 'x':
            x = torch.cat([x, mask.float()], dim=1).contiguous()                     # net/unet.py:437-438
        _, p1 = self.Conv1.run_nhwc(x, ws, tag + '.c1', want_out=False, want_pool=True)
        m2 = self._mask_source(mask, 2, ws, tag + '.m2', split) if self.mfm == 'x2' else None     # :442-443
        _, p2 = self.Conv2.run_nhwc(p1, ws, tag + '.c2', x1=m2, want_out=False, want_pool=True)
        m3 = self._mask_source(mask, 4, ws, tag + '.m3', split) if self.mfm == 'x3' else None     # :446-447
        x3, p3 = self.Conv3.run_nhwc(p2, ws, tag + '.c3', x1=m3, want_pool=True)
        x4, p4 = self.Conv4.run_nhwc(p3, ws, tag + '.c4', want_pool=True)
        x5, _ = self.Conv5.run_nhwc(p4, ws, tag + '.c5')
        u5 = self.Up5.run_nhwc(x5, ws, tag + '.u5')
        d5, _ = self.Up_conv5.run_nhwc(x4, ws, tag + '.uc5', x1=u5)
        u4 = self.Up4.run_nhwc(d5, ws, tag + '.u4')
        d4, _ = self.Up_conv4.run_nhwc(x3, ws, tag + '.uc4', x1=u4)
        return engine.hi_of(d4)                # the context-relation encoder consumes plain fp16 features

    def forward(self, x, mask, do_last_conv=True):
        """Reference signature (net/unet.py:435); `mask` is only read by the unsupported mask_feature_map variants."""
        return {'d4': engine.nhwc_to_nchw_f32(self.encode_nhwc(x.float().contiguous(), 'fwd', mask=mask))}
