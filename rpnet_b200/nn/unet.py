"""Truncated U-Net encoder with the reference's constructor / forward signature and state_dict keys
(net/unet.py:351-467: Unet_2D base + U_Net).  16 convs -> 'd4' (256 channels at H/4)."""
import torch.nn as nn

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
        if cfg['mask_feature_map']:
            # yamls/example.yml:103 `mask_feature_map: no` -> False; the 'x'/'x2'/'x3' variants add a
            # mask channel (Cin 2 / 65 / 129) that has no B200 kernel yet.
            raise NotImplementedError('mask_feature_map=%r is not supported by the B200 path' % (cfg['mask_feature_map'],))
        self.Maxpool = nn.MaxPool2d(kernel_size=2, stride=2)     # fused into the preceding conv's epilogue
        num_feats = [64, 128, 256, 512, 1024]
        norm = cfg['unet_normalize_type']
        pr = engine.precision_of(cfg)          # 'split' (fp32-class, default) or 'fp16' (engine.PRECISIONS)
        pd = engine.decoder_precision(pr)      # decoder half: the same, or fp16 weights with RPNET_SPLIT_DECODER=2 (engine.decoder_precision)
        self.Conv1 = conv_block(ch_in=self.img_ch, ch_out=num_feats[0], normalization_type=norm, precision=pr)
        self.Conv2 = conv_block(ch_in=num_feats[0], ch_out=num_feats[1], normalization_type=norm, precision=pr)
        self.Conv3 = conv_block(ch_in=num_feats[1], ch_out=num_feats[2], normalization_type=norm, precision=pr)
        self.Conv4 = conv_block(ch_in=num_feats[2], ch_out=num_feats[3], normalization_type=norm, precision=pr)
        self.Conv5 = conv_block(ch_in=num_feats[3], ch_out=num_feats[4], normalization_type=norm, precision=pr)
        self.Up5 = up_conv(ch_in=num_feats[4], ch_out=num_feats[3], normalization_type=norm, precision=pd)
        self.Up_conv5 = conv_block(ch_in=num_feats[3] * 2, ch_out=num_feats[3], normalization_type=norm, precision=pd)
        self.Up4 = up_conv(ch_in=num_feats[3], ch_out=num_feats[2], normalization_type=norm, precision=pd)
        self.Up_conv4 = conv_block(ch_in=num_feats[2] * 2, ch_out=num_feats[2], normalization_type=norm, precision=pd)
        self._ws = engine.Workspace()

    def encode_nhwc(self, x, tag='enc'):
        """x: fp32 NCHW image batch [n, 1, H, W] (H, W multiples of 16) -> d4 fp16 NHWC [n, H/4, W/4, 256].
        Schedule (net/unet.py:435-467): the 2x2 max-pools are fused into the epilogue of the conv that feeds
        them (x1 / x2 are never written at full resolution), nearest-upsample + conv runs in sub-pixel form and
        the skip concatenations are two-source K loops."""
        n, c, H, W = x.shape
        if H % 16 or W % 16:
            raise ValueError('U_Net input must be a multiple of 16 pixels (got %d x %d)' % (H, W))
        ws = self._ws
        _, p1 = self.Conv1.run_nhwc(x, ws, tag + '.c1', want_out=False, want_pool=True)
        _, p2 = self.Conv2.run_nhwc(p1, ws, tag + '.c2', want_out=False, want_pool=True)
        x3, p3 = self.Conv3.run_nhwc(p2, ws, tag + '.c3', want_pool=True)
        x4, p4 = self.Conv4.run_nhwc(p3, ws, tag + '.c4', want_pool=True)
        x5, _ = self.Conv5.run_nhwc(p4, ws, tag + '.c5')
        u5 = self.Up5.run_nhwc(x5, ws, tag + '.u5')
        d5, _ = self.Up_conv5.run_nhwc(x4, ws, tag + '.uc5', x1=u5)
        u4 = self.Up4.run_nhwc(d5, ws, tag + '.u4')
        d4, _ = self.Up_conv4.run_nhwc(x3, ws, tag + '.uc4', x1=u4)
        return engine.hi_of(d4)                # the context-relation encoder consumes plain fp16 features

    def forward(self, x, mask, do_last_conv=True):
        """Reference signature (net/unet.py:435); `mask` is only read by the unsupported mask_feature_map variants."""
        return {'d4': engine.nhwc_to_nchw_f32(self.encode_nhwc(x.float().contiguous(), 'fwd'))}
