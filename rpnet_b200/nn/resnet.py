"""ResNet18 backbone wrapper with the reference's structure and state_dict keys (net/rp_net.py:19-42): torchvision's
resnet18 stem + layer1 (stride 4) followed by three custom stride-1 BasicBlock stages -> 512 channels at H/4.
"Next" row N3 of SURVEY §8(f).  This module runs the eval forward; the train step (batch-statistics BatchNorm, residual backward,
stem weight gradient) is scheduled for the whole model by rpnet_b200.train.ResNetTrainEngine (RP_Net.forward in train mode / TrainStep).

The torchvision modules are parameter containers: forward() never calls them — it runs the stem kernel, the max-pool kernel
and the tcgen05 implicit-GEMM conv with the BasicBlock's residual add + ReLU fused into its epilogue."""
import torch
import torch.nn as nn

from .. import engine, ops
from .modules import _PackedModule


class ResNet18(_PackedModule):
    def __init__(self, use_pretrained=False):
        super().__init__()
        import torchvision
        from torchvision.models.resnet import BasicBlock
        if use_pretrained:
            raise NotImplementedError('pretrained torchvision weights cannot be downloaded here; load a state_dict instead')
        resnet_net = torchvision.models.resnet18()
        modules = list(resnet_net.children())[:-5]                      # conv1, bn1, relu, maxpool, layer1
        for cin, cout in ((64, 128), (128, 256), (256, 512)):           # net/rp_net.py:24-35
            modules.append(nn.Sequential(
                BasicBlock(cin, cout, downsample=nn.Sequential(nn.Conv2d(cin, cout, 1), nn.BatchNorm2d(cout))),
                BasicBlock(cout, cout)))
        self.backbone = nn.Sequential(*modules)
        self.backbone.out_channels = 512
        # arithmetic of the 15 tensor-core convs: 'split8' (fp32-class, default), 'split' or 'fp16' — engine.PRECISIONS; the
        # reference's constructor has no config argument, so the choice comes from RPNET_PRECISION / `encoder.split = False`
        # (+ `encoder.w_level`), like nn/vgg.py
        self.split = engine.is_split(engine.default_precision())
        self.w_level = engine.w_level(engine.default_precision())
        self._ws = engine.Workspace()

    def _signature(self):
        return super()._signature() + (self.split, self.w_level)

    def _blocks(self):
        return [blk for stage in list(self.backbone)[4:] for blk in stage]

    def _build_packs(self):
        conv1, bn1 = self.backbone[0], self.backbone[1]
        zero = torch.zeros(64, device=conv1.weight.device)
        stem = engine.fold_bn(zero, bn1.weight, bn1.bias, bn1.running_mean, bn1.running_var, bn1.eps)
        packs = []
        for blk in self._blocks():
            def cb(conv, bn, relu):
                bias = conv.bias if conv.bias is not None else torch.zeros(conv.out_channels, device=conv.weight.device)
                scale, shift = engine.fold_bn(bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
                lvl = self.w_level if self.split else engine.W_FP16
                wp, taps = engine.pack_weight_taps(conv.weight, split=lvl)
                return engine.ConvPack(wp, taps, scale, shift, relu, split=self.split, w_split=lvl)
            down = cb(blk.downsample[0], blk.downsample[1], False) if blk.downsample is not None else None
            packs.append((cb(blk.conv1, blk.bn1, True), cb(blk.conv2, blk.bn2, True), down))
        return stem, packs

    def encode_nhwc(self, x, tag='res'):
        """x fp32 NCHW [n, 3, H, W] -> fp16 NHWC [n, H/4, W/4, 512] ((hi, lo) planes in split precision)."""
        if self.training:
            raise NotImplementedError('ResNet18.forward on its own runs in eval mode only: the train-mode kernels are scheduled for the whole '
                                      'model by RP_Net.forward / rpnet_b200.train.TrainStep (ResNetTrainEngine); call .eval() for the encoder alone')
        ws, dev = self._ws, x.device
        (s_scale, s_shift), packs = self._packs()
        n, _, H, W = x.shape
        h2, w2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        f16, sp = torch.float16, self.split

        def buf(name, shape):                                            # hi plane (+ residual plane in split precision)
            return ws.get(name, shape, f16, dev), (engine.lo_buffer(ws, name + '.lo', shape, dev, self.w_level) if sp else None)

        def conv(pk, src, dst, res=(None, None)):
            if sp:
                ops.conv_split(src[0], pk.wpack, pk.taps, pk.scale, pk.shift, pk.relu, src0_lo=src[1], w_split=pk.w_split, out=dst[0],
                               out_lo=dst[1], res=res[0], res_lo=res[1])
            else:
                ops.conv_res(src[0], pk.wpack, pk.taps, pk.scale, pk.shift, dst[0], res=res[0], relu=pk.relu)

        stem = buf(tag + '.stem', (n, h2, w2, 64))
        ops.conv7x7s2_stem(x.float().contiguous(), self.backbone[0].weight.detach().float().contiguous(), s_scale, s_shift, stem[0],
                           out_lo=stem[1])
        h4, w4 = (h2 + 2 - 3) // 2 + 1, (w2 + 2 - 3) // 2 + 1
        cur = buf(tag + '.pool', (n, h4, w4, 64))
        ops.maxpool(stem[0], 3, 2, 1, cur[0], x_lo=stem[1], out_lo=cur[1])   # nn.MaxPool2d(3, 2, 1)
        for i, (p1, p2, down) in enumerate(packs):                       # BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + identity)
            a = buf('%s.b%d.a' % (tag, i), (n, h4, w4, p1.cout))
            conv(p1, cur, a)
            identity = cur
            if down is not None:
                identity = buf('%s.b%d.d' % (tag, i), (n, h4, w4, down.cout))
                conv(down, cur, identity)
            out = buf('%s.b%d.o' % (tag, i), (n, h4, w4, p2.cout))
            conv(p2, a, out, res=identity)
            cur = out
        return cur if sp else cur[0]

    def forward(self, x, mask=None):
        """Reference signature (net/rp_net.py:39-42): returns {'d4': NCHW fp32}."""
        return {'d4': engine.nhwc_to_nchw_f32(self.encode_nhwc(x.float().contiguous(), 'fwd'))}
