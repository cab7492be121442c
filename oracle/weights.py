"""Reference-identical random initialisation, restated.  TEST INFRASTRUCTURE.

Builds the RP_Net state_dict with the same torch RNG consumption order as the reference
constructors (net/rp_net.py:195-224 -> net/unet.py:393-430 -> net/modules.py:42-75, then
net/rp_net.py:45-74), so that `torch.manual_seed(s)` + this function yields bit-identical
tensors to `torch.manual_seed(s); RP_Net(...)` (test_rpnet.py:8-10).  Checked against
checksums recorded from the real reference in tests/golden/init_checksums.npz.
"""
from collections import OrderedDict

import torch
import torch.nn as nn


def _conv_bn(sd, key_conv, key_bn, cin, cout, k):
    conv = nn.Conv2d(cin, cout, k, padding=k // 2)
    bn = nn.BatchNorm2d(cout)
    for n, t in conv.state_dict().items():
        sd[key_conv + '.' + n] = t
    for n, t in bn.state_dict().items():
        sd[key_bn + '.' + n] = t


def unet_rpnet_state_dict(seed=0, radius=5):
    torch.manual_seed(seed)
    sd = OrderedDict()
    f = [64, 128, 256, 512, 1024]

    def block(name, cin, cout):     # conv_block: net/modules.py:46-54
        _conv_bn(sd, 'encoder.%s.conv.0' % name, 'encoder.%s.conv.1' % name, cin, cout, 3)
        _conv_bn(sd, 'encoder.%s.conv.3' % name, 'encoder.%s.conv.4' % name, cout, cout, 3)

    def up(name, cin, cout):        # up_conv: net/modules.py:65-71
        _conv_bn(sd, 'encoder.%s.up.1' % name, 'encoder.%s.up.2' % name, cin, cout, 3)

    block('Conv1', 1, f[0]); block('Conv2', f[0], f[1]); block('Conv3', f[1], f[2])
    block('Conv4', f[2], f[3]); block('Conv5', f[3], f[4])
    up('Up5', f[4], f[3]); block('Up_conv5', 2 * f[3], f[3])
    up('Up4', f[3], f[2]); block('Up_conv4', 2 * f[2], f[2])
    c = 256                         # ContextCorrelationEncoder: net/rp_net.py:49-74
    _conv_bn(sd, 'cre.w_k.0', 'cre.w_k.1', c, c, 3)
    _conv_bn(sd, 'cre.w_q.0', 'cre.w_q.1', c, c, 3)
    _conv_bn(sd, 'cre.w_context.0', 'cre.w_context.1', 2 * c, c, 1)
    _conv_bn(sd, 'cre.q.0', 'cre.q.1', c + (2 * radius + 1) ** 2, 64, 1)
    _conv_bn(sd, 'cre.out.0', 'cre.out.1', 2 * c, 64, 1)
    return sd


def resnet_rpnet_state_dict(seed=0, radius=5):
    """RP_Net(backbone='resnet') state_dict with the reference's RNG consumption order (net/rp_net.py:214-216 ->
    ResNet18.__init__ :20-37: a full torchvision resnet18 is constructed first, then the three custom stages; then the cre
    with 512 input channels, :221)."""
    import torchvision
    from torchvision.models.resnet import BasicBlock
    torch.manual_seed(seed)
    sd = OrderedDict()
    net = torchvision.models.resnet18()
    modules = list(net.children())[:-5]
    for cin, cout in ((64, 128), (128, 256), (256, 512)):
        modules.append(nn.Sequential(BasicBlock(cin, cout, downsample=nn.Sequential(nn.Conv2d(cin, cout, 1), nn.BatchNorm2d(cout))),
                                     BasicBlock(cout, cout)))
    for k, v in nn.Sequential(*modules).state_dict().items():
        sd['encoder.backbone.' + k] = v
    c = 512
    _conv_bn(sd, 'cre.w_k.0', 'cre.w_k.1', c, c, 3)
    _conv_bn(sd, 'cre.w_q.0', 'cre.w_q.1', c, c, 3)
    _conv_bn(sd, 'cre.w_context.0', 'cre.w_context.1', 2 * c, c, 1)
    _conv_bn(sd, 'cre.q.0', 'cre.q.1', c + (2 * radius + 1) ** 2, 64, 1)
    _conv_bn(sd, 'cre.out.0', 'cre.out.1', 2 * c, 64, 1)
    return sd


def vgg_state_dict(seed=0, in_channels=3):
    """net/vgg.py:22-63: 13 convs then kaiming_normal_ on every conv weight in module order."""
    torch.manual_seed(seed)
    sd = OrderedDict()
    plan = [(2, in_channels, 64), (2, 64, 128), (3, 128, 256), (3, 256, 512), (3, 512, 512)]
    convs = []
    for b, (n, cin, cout) in enumerate(plan):
        for i in range(n):
            conv = nn.Conv2d(cin, cout, 3, padding=1)
            convs.append(('features.%d.%d' % (2 * b, 2 * i), conv))
            cin = cout
    for key, conv in convs:
        torch.nn.init.kaiming_normal_(conv.weight, nonlinearity='relu')
        sd[key + '.weight'] = conv.weight.detach()
        sd[key + '.bias'] = conv.bias.detach()
    return sd


def checksums(sd):
    """name -> (sum, abs-sum, first element) in float64, for init-equivalence checks."""
    out = {}
    for k, v in sd.items():
        v = v.detach().double().reshape(-1)
        out[k] = (v.sum().item(), v.abs().sum().item(), v[0].item() if v.numel() else 0.0)
    return out
