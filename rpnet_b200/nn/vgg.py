"""VGG-16 style encoder with the reference's signature and state_dict keys (net/vgg.py:8-74):
13 convs (+ReLU, none after the last), MaxPool k3 p1 (stride 2, 2, 2, 1), last block dilation 2 -> 512 ch at H/8.
The reference cannot run this backbone through RP_Net; it is parity-checked standalone."""
import torch
import torch.nn as nn

from .. import engine, ops
from .modules import _PackedModule


class Encoder(_PackedModule):
    def __init__(self, in_channels=3, pretrained_path=None):
        super().__init__()
        self.pretrained_path = pretrained_path
        self.features = nn.Sequential(
            self._make_layer(2, in_channels, 64),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            self._make_layer(2, 64, 128),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            self._make_layer(3, 128, 256),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            self._make_layer(3, 256, 512),
            nn.MaxPool2d(kernel_size=3, stride=1, padding=1),
            self._make_layer(3, 512, 512, dilation=2, lastRelu=False),
        )
        self._init_weights()
        self._ws = engine.Workspace()
        # arithmetic of the 12 tensor-core convs: 'split8' (fp32-class, default), 'split' or 'fp16' — engine.PRECISIONS; the
        # reference's constructor has no config argument, so the choice comes from RPNET_PRECISION / `encoder.split = False`
        # (+ `encoder.w_level`, the weight-pack level engine.W_SPLIT / engine.W_C8)
        self.split = engine.is_split(engine.default_precision())
        self.w_level = engine.w_level(engine.default_precision())

    def _make_layer(self, n_convs, in_channels, out_channels, dilation=1, lastRelu=True):
        layer = []
        for i in range(n_convs):
            layer.append(nn.Conv2d(in_channels, out_channels, kernel_size=3, dilation=dilation, padding=dilation))
            if i != n_convs - 1 or lastRelu:
                layer.append(nn.ReLU(inplace=True))
            in_channels = out_channels
        return nn.Sequential(*layer)

    def _init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight, nonlinearity='relu')
        if self.pretrained_path is not None:        # net/vgg.py:65-74: first 26 tensors, positional
            dic = torch.load(self.pretrained_path, map_location='cpu')
            keys = list(dic.keys())
            new_dic = self.state_dict()
            new_keys = list(new_dic.keys())
            for i in range(26):
                new_dic[new_keys[i]] = dic[keys[i]]
            self.load_state_dict(new_dic)

    def _signature(self):
        return super()._signature() + (self.split, self.w_level)

    def _build_packs(self):
        plan = []       # ('conv', conv, pack-or-None, relu) | ('pool', k, s, p)
        for blk in self.features:
            if isinstance(blk, nn.MaxPool2d):
                plan.append(('pool', blk.kernel_size, blk.stride, blk.padding))
                continue
            mods = list(blk)
            for i, m in enumerate(mods):
                if isinstance(m, nn.Conv2d):
                    relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                    if m.in_channels in (1, 3):
                        plan.append(('first', m, None, relu))
                    else:
                        scale, shift = engine.fold_bn(m.bias)
                        lvl = self.w_level if self.split else engine.W_FP16
                        wp, taps = engine.pack_weight_taps(m.weight, m.dilation[0], split=lvl)
                        plan.append(('conv', m, engine.ConvPack(wp, taps, scale, shift, relu, split=self.split, w_split=lvl), relu))
        return plan

    def encode_nhwc(self, x, tag='vgg', trace=None):
        """x fp32 NCHW [n, 3, H, W] -> fp16 NHWC [n, H/8, W/8, 512].  trace: optional list that receives one record per layer
        (what the train-step backward needs: ('first' | 'conv', conv module, relu, input, output) / ('pool', (k, s, p), input, output,
        argmax positions)) — the train-mode forward of this stack IS its eval forward (no normalisation layers)."""
        ws = self._ws
        cur = None
        for i, step in enumerate(self._packs()):
            name = '%s.%d' % (tag, i)
            if step[0] == 'first':
                conv = step[1]
                scale, shift = engine.fold_bn(conv.bias)
                n, _, h, w = x.shape
                cur = ws.get(name, (n, h, w, 64), torch.float16, x.device)
                lo = engine.lo_buffer(ws, name + '.lo', (n, h, w, 64), x.device, self.w_level) if self.split else None
                ops.conv3x3_first(x.float().contiguous(), conv.weight.detach().float().contiguous(), scale, shift, step[3], cur, out_lo=lo)
                cur = (cur, lo) if self.split else cur
                if trace is not None:
                    trace.append(('first', conv, step[3], x, cur))
            elif step[0] == 'conv':
                prev = cur
                cur, _ = engine.run_conv(step[2], cur, ws, name)
                if trace is not None:
                    trace.append(('conv', step[1], step[3], prev, cur))
            else:
                _, k, s, p = step
                n, h, w, c = engine.hi_of(cur).shape
                shp = (n, (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1, c)
                out = ws.get(name, shp, torch.float16, x.device)
                idx = ws.get(name + '.idx', shp, torch.uint8, x.device) if trace is not None else None
                prev = cur
                if self.split:
                    out_lo = engine.lo_buffer(ws, name + '.lo', shp, x.device, self.w_level)
                    ops.maxpool(cur[0], k, s, p, out, x_lo=cur[1], out_lo=out_lo, idx=idx)
                    cur = (out, out_lo)
                else:
                    ops.maxpool(cur, k, s, p, out, idx=idx)
                    cur = out
                if trace is not None:
                    trace.append(('pool', (k, s, p), prev, cur, idx))
        return cur

    def forward(self, x, mask=None):
        return engine.nhwc_to_nchw_f32(self.encode_nhwc(x))      # split mode: hi + lo planes joined in fp32
