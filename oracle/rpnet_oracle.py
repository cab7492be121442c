"""CPU oracle for the RP-Net hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain fp32 CPU restatement of the reference algorithm
(uci-cbcl/RP-Net @169a0268) written functionally over a ``state_dict``.  It is
NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and
only as the checker / the CPU baseline.  The shipped forward
(``rpnet_b200``) never imports anything from ``oracle/``.

Parity pinning: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference modules themselves,
executed in the build container by ``tests/golden/make_golden.py`` and committed
under ``tests/golden/*.npz`` (checked by ``tests/test_oracle_golden.py``).

All arithmetic is torch-2.11 ATen CPU fp32 — the same third-party numeric
backend the reference calls (it pins no version).  Every function cites the
reference ``file:line`` it follows (paths relative to the reference root).

Generalisation ("oracle-ext"): the shipped reference forward
only runs 1-way 1-shot.  ``forward`` here accepts Wa ways x Sh shots with the
minimal extension
  * ``cre`` is applied per (way, shot) with that shot's own pooled fore mask
    (reference: ``net/rp_net.py:274-275`` does it for [0][0] only);
  * the recurrent mask is ``sum_{w>=1} softmax(logits)[:, w] > 0.5``
    (reference: channel 1 only, ``net/rp_net.py:308-311``) which for Wa == 1 is
    the identical expression;
and reduces bit-for-bit to the reference for Wa == Sh == 1 (tested).
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default, net/modules.py:49
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default

# ----------------------------------------------------------------------------
# storage emulation (tests only)
# ----------------------------------------------------------------------------
# STORAGE = None      : the reference's arithmetic, fp32 everywhere (the oracle proper).
# STORAGE = 'b200'    : same algorithm, but every tensor the B200 path keeps in HBM as fp16 (conv weights, the pre-BN conv
#                       output z, the activation y, masked inputs, the correlation volume) is rounded to fp16 at that point
#                       (straight-through gradient), and — with GRAD_STORAGE = 'bf16' — activation gradients are rounded to
#                       bf16 where the backward kernels store them.  It separates "what fp16/bf16 storage does to this
#                       network" (the fp32 reference itself moves by that much under a 5e-4 relative perturbation of its
#                       activations) from kernel errors.  Parity against the fp32 oracle is always reported as well.
STORAGE = None
GRAD_STORAGE = None


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.half().float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _q(x):
    """fp16 storage point of the B200 path (identity for the oracle proper)."""
    return _RoundFwd.apply(x) if STORAGE == 'b200' else x


def _qg(x):
    """bf16 storage point of an activation gradient in the B200 backward (identity for the oracle proper)."""
    return _RoundGrad.apply(x) if (STORAGE == 'b200' and GRAD_STORAGE == 'bf16' and x.requires_grad) else x


# ----------------------------------------------------------------------------
# conv / norm building blocks
# ----------------------------------------------------------------------------
def conv_bn_relu(x, sd, conv, bn, training=False, padding=1, dilation=1, relu=True):
    """Conv2d(bias) -> BatchNorm2d -> ReLU.  net/modules.py:47-50, net/rp_net.py:50-54.

    ``conv`` / ``bn`` are state_dict prefixes (e.g. 'encoder.Conv1.conv.0').
    In training mode the running statistics held in ``sd`` are updated in place
    exactly like nn.BatchNorm2d does (momentum 0.1, unbiased running_var) and
    ``num_batches_tracked`` is incremented.
    """
    if STORAGE == 'b200' and training and bn is not None:
        # B200 train path: fp16 weights (fp32 for the Cin=1 first conv), bias dropped in front of batch-statistics BN (it
        # re-enters the running mean), z and y stored as fp16 (the fp32 head output of cre.q is not rounded)
        w = sd[conv + '.weight']
        first = x.shape[1] < 64
        z = F.conv2d(x if first else _qg(x), w if first else _q(w), None, padding=padding, dilation=dilation)
        z = _qg(_q(z))
        sd[bn + '.num_batches_tracked'] += 1
        with torch.no_grad():       # (1-m) * (rm + m/(1-m) * bias) + m * mean(z) == (1-m) * rm + m * (mean(z) + bias)
            sd[bn + '.running_mean'].data.add_(BN_MOMENTUM / (1 - BN_MOMENTUM) * sd[conv + '.bias'].data)
        y = F.batch_norm(z, sd[bn + '.running_mean'], sd[bn + '.running_var'], sd[bn + '.weight'], sd[bn + '.bias'], True,
                         BN_MOMENTUM, BN_EPS)
        y = F.relu(y) if relu else y
        return y if conv.endswith('cre.q.0') else _q(y)
    y = F.conv2d(x, sd[conv + '.weight'], sd[conv + '.bias'], padding=padding, dilation=dilation)
    if bn is not None and (bn + '.running_mean') not in sd:
        # unet_normalize_type: InstanceNorm2d — nn.InstanceNorm2d(C) (net/modules.py:49,52,69 through getattr(nn, ...)): no affine
        # parameters, no running statistics: per-(image, channel) statistics in train AND eval mode, eps 1e-5
        y = F.instance_norm(y, eps=BN_EPS)
        return F.relu(y) if relu else y
    if bn is not None:
        if training and (bn + '.num_batches_tracked') in sd:
            sd[bn + '.num_batches_tracked'] += 1
        y = F.batch_norm(y, sd[bn + '.running_mean'], sd[bn + '.running_var'], sd[bn + '.weight'],
                         sd[bn + '.bias'], training, BN_MOMENTUM, BN_EPS)
    return F.relu(y) if relu else y


def conv_block(x, sd, p, training=False):
    """(3x3 conv + BN + ReLU) x 2.  net/modules.py:42-58."""
    x = conv_bn_relu(x, sd, p + '.conv.0', p + '.conv.1', training)
    return conv_bn_relu(x, sd, p + '.conv.3', p + '.conv.4', training)


def up_conv(x, sd, p, training=False):
    """nearest x2 upsample + 3x3 conv + BN + ReLU.  net/modules.py:61-75."""
    x = F.interpolate(x, scale_factor=2, mode='nearest')
    return conv_bn_relu(x, sd, p + '.up.1', p + '.up.2', training)


def unet_encoder(x, sd, prefix='encoder.', training=False, want=None, mask=None, mask_feature_map=False):
    """Truncated U-Net, returns d4 (256 ch @ H/4).  net/unet.py:435-467.  mask_feature_map False (yamls/example.yml:103
    parses to False) or 'x' / 'x2' / 'x3': ``mask`` [n, 1, H, W] (avg-pooled by 2 / 4) is concatenated to the
    input of Conv1 / Conv2 / Conv3 (:437-449).  ``want`` may be a dict that receives the intermediate maps (tests)."""
    p = prefix
    if mask_feature_map == 'x':
        x = torch.cat([x, mask], dim=1)
    x1 = conv_block(x, sd, p + 'Conv1', training)
    x2 = F.max_pool2d(x1, 2, 2)
    if mask_feature_map == 'x2':
        x2 = torch.cat([x2, F.avg_pool2d(mask, 2)], dim=1)
    x2 = conv_block(x2, sd, p + 'Conv2', training)
    x3 = F.max_pool2d(x2, 2, 2)
    if mask_feature_map == 'x3':
        x3 = torch.cat([x3, F.avg_pool2d(mask, 4)], dim=1)
    x3 = conv_block(x3, sd, p + 'Conv3', training)
    x4 = conv_block(F.max_pool2d(x3, 2, 2), sd, p + 'Conv4', training)
    x5 = conv_block(F.max_pool2d(x4, 2, 2), sd, p + 'Conv5', training)
    d5 = up_conv(x5, sd, p + 'Up5', training)
    d5 = conv_block(torch.cat((x4, d5), dim=1), sd, p + 'Up_conv5', training)
    d4 = up_conv(d5, sd, p + 'Up4', training)
    d4 = conv_block(torch.cat((x3, d4), dim=1), sd, p + 'Up_conv4', training)
    if want is not None:
        want.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, d5=d5, d4=d4)
    return d4


# (n_convs, cin, cout, dilation, last_relu) per block and the pool that follows; net/vgg.py:22-32
_VGG_BLOCKS = [(2, None, 64, 1, True, (3, 2, 1)), (2, 64, 128, 1, True, (3, 2, 1)),
               (3, 128, 256, 1, True, (3, 2, 1)), (3, 256, 512, 1, True, (3, 1, 1)),
               (3, 512, 512, 2, False, None)]


def vgg_encoder(x, sd, prefix='encoder.'):
    """VGG-16 style conv stack -> 512 ch @ H/8.  net/vgg.py:22-58 (conv+ReLU, MaxPool k3 p1,
    last block dilation 2 and no final ReLU).  state_dict keys follow
    nn.Sequential nesting: features.<2*blk>.<2*i>.{weight,bias}."""
    for b, (n, _, _, dil, last_relu, pool) in enumerate(_VGG_BLOCKS):
        for i in range(n):
            key = '%sfeatures.%d.%d' % (prefix, 2 * b, 2 * i)
            x = F.conv2d(x, sd[key + '.weight'], sd[key + '.bias'], padding=dil, dilation=dil)
            if i != n - 1 or last_relu:
                x = F.relu(x)
        if pool is not None:
            x = F.max_pool2d(x, pool[0], pool[1], pool[2])
    return x


def _bn_eval(x, sd, bn):
    return F.batch_norm(x, sd[bn + '.running_mean'], sd[bn + '.running_var'], sd[bn + '.weight'], sd[bn + '.bias'], False, 0.0, BN_EPS)


def _bn(x, sd, bn, training):
    """nn.BatchNorm2d: eval mode on the running statistics, train mode on the batch statistics of this call (running statistics
    and num_batches_tracked updated in place, momentum 0.1 — SURVEY D14)."""
    if not training:
        return _bn_eval(x, sd, bn)
    if (bn + '.num_batches_tracked') in sd:
        sd[bn + '.num_batches_tracked'] += 1
    return F.batch_norm(x, sd[bn + '.running_mean'], sd[bn + '.running_var'], sd[bn + '.weight'], sd[bn + '.bias'], True, BN_MOMENTUM,
                        BN_EPS)


def resnet_encoder(x, sd, prefix='encoder.', training=False):
    """ResNet18 wrapper, net/rp_net.py:19-42: torchvision resnet18 conv1 (7x7 s2 p3, no bias) + bn1 + relu +
    maxpool(3, 2, 1) + layer1, then three stride-1 stages of [BasicBlock(cin, cout, downsample = 1x1 conv(bias) + BN),
    BasicBlock(cout, cout)] -> 512 ch @ H/4.  BasicBlock (torchvision): relu(bn2(conv2(relu(bn1(conv1(x))))) + identity)."""
    p = prefix + 'backbone.'
    x = F.relu(_bn(F.conv2d(x, sd[p + '0.weight'], None, stride=2, padding=3), sd, p + '1', training))
    x = F.max_pool2d(x, 3, 2, 1)
    for stage in (4, 5, 6, 7):
        for b in (0, 1):
            q = '%s%d.%d.' % (p, stage, b)
            out = F.relu(_bn(F.conv2d(x, sd[q + 'conv1.weight'], None, padding=1), sd, q + 'bn1', training))
            out = _bn(F.conv2d(out, sd[q + 'conv2.weight'], None, padding=1), sd, q + 'bn2', training)
            identity = x
            if (q + 'downsample.0.weight') in sd:
                identity = _bn(F.conv2d(x, sd[q + 'downsample.0.weight'], sd[q + 'downsample.0.bias']), sd, q + 'downsample.1', training)
            x = F.relu(out + identity)
    return x


# ----------------------------------------------------------------------------
# context-relation encoder
# ----------------------------------------------------------------------------
def correlation_allpairs(fmap1, fmap2, r=3):
    """Faithful restatement of the reference data flow, net/rp_net.py:153-181 (+130-150):
    all-pairs (H'W')x(H'W') matmul, / sqrt(C), then a (2r+1)^2 window gathered with
    grid_sample(align_corners=True, zeros) at integer pixel coordinates.
    Used to pin ``correlation_local`` and as the timed CPU baseline (it is what the
    reference executes)."""
    b, c, h, w = fmap1.shape
    corr = torch.matmul(fmap1.reshape(b, c, h * w).transpose(1, 2), fmap2.reshape(b, c, h * w))
    corr = corr / torch.sqrt(torch.tensor(c).float())
    corr = corr.reshape(b * h * w, 1, h, w)
    dev = fmap1.device                                                   # (the reference builds these on fmap1.device too, :131,171)
    ys, xs = torch.meshgrid(torch.arange(h, device=dev), torch.arange(w, device=dev), indexing='ij')
    centre = torch.stack([xs, ys], dim=-1).float()                       # (x, y) per pixel, :130-133
    centre = centre[None].repeat(b, 1, 1, 1).reshape(b * h * w, 1, 1, 2)
    d = torch.linspace(-r, r, 2 * r + 1, device=dev)
    delta = torch.stack(torch.meshgrid(d, d, indexing='ij'), dim=-1)     # [a, b'] -> (d[a], d[b']) added to (x, y): D8
    coords = centre + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
    gx = 2 * coords[..., 0:1] / (w - 1) - 1                               # :139-141
    gy = 2 * coords[..., 1:2] / (h - 1) - 1
    out = F.grid_sample(corr, torch.cat([gx, gy], dim=-1), align_corners=True)
    return out.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous().float()


def correlation_local(fmap1, fmap2, r=3):
    """Algebraically identical local form:
    out[b, a*(2r+1)+b', i, j] = 1/sqrt(C) * sum_c fmap1[b,c,i,j] * fmap2[b,c,i+(b'-r),j+(a-r)],
    zero outside the map.  Channel index: a -> column (x) offset, b' -> row (y) offset."""
    b, c, h, w = fmap1.shape
    k = 2 * r + 1
    pad = F.pad(fmap2, (r, r, r, r))
    out = fmap1.new_empty(b, k * k, h, w)
    scale = 1.0 / math.sqrt(c)
    for a in range(k):          # x offset a - r
        for bb in range(k):     # y offset bb - r
            out[:, a * k + bb] = (fmap1 * pad[:, :, bb:bb + h, a:a + w]).sum(1) * scale
    return out


def cre(x_fg, x_bg, sd, radius, prefix='cre.', training=False, allpairs=False, want=None):
    """ContextCorrelationEncoder.forward, net/rp_net.py:77-84.  (w_context / out are
    never used by the reference forward: SURVEY D4.)"""
    fm1 = conv_bn_relu(_q(x_fg), sd, prefix + 'w_k.0', prefix + 'w_k.1', training)
    fm2 = conv_bn_relu(_q(x_bg), sd, prefix + 'w_q.0', prefix + 'w_q.1', training)
    corr = _q((correlation_allpairs if allpairs else correlation_local)(fm1, fm2, r=radius))
    out = conv_bn_relu(torch.cat([corr, fm1], dim=1), sd, prefix + 'q.0', prefix + 'q.1', training, padding=0)
    if want is not None:
        want.update(fm1=fm1, fm2=fm2, corr=corr, out=out)
    return out


# ----------------------------------------------------------------------------
# prototype matching
# ----------------------------------------------------------------------------
def get_features(fts, mask):
    """Masked average pooling, net/rp_net.py:366-376.  fts 1xCxH'xW', mask 1xHxW -> 1xC."""
    fts = F.interpolate(fts, size=mask.shape[-2:], mode='bilinear')
    return torch.sum(fts * mask[None, ...], dim=(2, 3)) / (mask[None, ...].sum(dim=(2, 3)) + 1e-5)


def get_prototype(fg_fts, bg_fts):
    """net/rp_net.py:379-391: fg = mean over shots per way; bg = mean over shots, then ways."""
    n_ways, n_shots = len(fg_fts), len(fg_fts[0])
    fg = [sum(way) / n_shots for way in fg_fts]
    bg = sum([sum(way) / n_shots for way in bg_fts]) / n_ways
    return fg, bg


def cal_dist(fts, prototype, scaler=20):
    """net/rp_net.py:353-363: 20 * cosine similarity over channels."""
    return F.cosine_similarity(fts, prototype[..., None, None], dim=1) * scaler


def prototypes_for(supp_fts, fore_mask, back_mask, epi):
    """Prototype list [bg, fg_1..fg_Wa] for episode ``epi``; net/rp_net.py:288-300.
    supp_fts Wa x Sh x B x C x H' x W'; masks Wa x Sh x B x H x W."""
    n_ways, n_shots = supp_fts.shape[:2]
    fg = [[get_features(supp_fts[w, s, [epi]], fore_mask[w, s, [epi]]) for s in range(n_shots)] for w in range(n_ways)]
    bg = [[get_features(supp_fts[w, s, [epi]], back_mask[w, s, [epi]]) for s in range(n_shots)] for w in range(n_ways)]
    fg_p, bg_p = get_prototype(fg, bg)
    return [bg_p] + fg_p


def align_loss(qry_fts, pred, supp_fts, fore_mask, back_mask):
    """Prototype alignment loss for one episode, net/rp_net.py:394-440.
    qry_fts NxCxH'xW', pred Nx(1+Wa)xH'xW', supp_fts WaxShxCxH'xW', masks WaxShxHxW."""
    n_ways, n_shots = len(fore_mask), len(fore_mask[0])
    pred_mask = pred.argmax(dim=1, keepdim=True)
    binary_masks = [pred_mask == i for i in range(1 + n_ways)]
    skip_ways = [i for i in range(n_ways) if binary_masks[i + 1].sum() == 0]
    pred_mask = torch.stack(binary_masks, dim=1).float()
    qry_prototypes = torch.sum(qry_fts.unsqueeze(1) * pred_mask, dim=(0, 3, 4))
    qry_prototypes = qry_prototypes / (pred_mask.sum((0, 3, 4)) + 1e-5)
    loss = 0
    for way in range(n_ways):
        if way in skip_ways:
            continue
        prototypes = [qry_prototypes[[0]], qry_prototypes[[way + 1]]]
        for shot in range(n_shots):
            img_fts = supp_fts[way, [shot]]
            supp_pred = torch.stack([cal_dist(img_fts, p) for p in prototypes], dim=1)
            supp_pred = F.interpolate(supp_pred, size=fore_mask.shape[-2:], mode='bilinear')
            supp_label = torch.full_like(fore_mask[way, shot], 255).long()
            supp_label[fore_mask[way, shot] == 1] = 1
            supp_label[back_mask[way, shot] == 1] = 0
            loss = loss + F.cross_entropy(supp_pred, supp_label[None, ...], ignore_index=255) / n_shots / n_ways
    return loss


# ----------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------
def dice_loss_softmax(logits, true, eps=1e-7):
    """net/rp_net.py:87-120, multi-class branch (num_classes >= 2).  ``eye`` is built on
    logits.device (the reference indexes a CPU eye: SURVEY D10; same values)."""
    num_classes = logits.shape[1]
    one_hot = torch.eye(num_classes, device=logits.device)[true].permute(0, 3, 1, 2).float()
    probas = F.softmax(logits, dim=1)
    dims = (0, 2, 3)
    intersection = torch.sum(probas * one_hot, dims)
    cardinality = torch.sum(probas + one_hot, dims)
    return 1 - (2. * intersection / (cardinality + eps)).mean()


def dice_ce(logits, true, eps=1e-7):
    """net/rp_net.py:123-127."""
    return dice_loss_softmax(logits, true, eps) + F.cross_entropy(logits, true)


# ----------------------------------------------------------------------------
# the forward
# ----------------------------------------------------------------------------
def recurrent_mask(logits, cfg, scale=4):
    """The mask update of net/rp_net.py:308-311 (generalised to Wa ways as in the module docstring): logits B x (1+Wa) x H x W ->
    the pooled mask B x 1 x H/scale x W/scale that the next refinement iteration consumes."""
    n_ways = logits.shape[1] - 1
    prob = logits.softmax(dim=1)[:, 1:, ...].sum(dim=1) if n_ways > 1 else logits.softmax(dim=1)[:, 1, ...]
    if cfg['soft_mask'] == False:  # noqa: E712  (same test as the reference)
        prob = (prob > 0.5).float()
    return F.avg_pool2d(prob.unsqueeze(1), scale)


def forward(sd, cfg, supp_imgs, fore_mask, back_mask, qry_imgs, appr_query_labels,
            training=False, align=True, backbone='UNet', allpairs=False, want=None, mask_override=None):
    """RP_Net.forward, net/rp_net.py:226-350, generalised to Wa x Sh (module docstring).

    sd   : state_dict (name -> tensor); BN buffers are updated in place when training.
    cfg  : the flat yaml dict (keys n_iter_refinement, mask_refinement_correlation_radius,
           soft_mask, optional scale) — net/rp_net.py:200-202, :48, :309.
    Returns {'output', 'align_loss', 'refinement'} like the reference.

    mask_override (tests only): {i: pooled mask B x 1 x H' x W'} replaces the recurrent mask that iteration i >= 1 consumes
    ("teacher forcing").  The hard mask (:310) makes iteration i+1 a discontinuous function of iteration i's logits: one
    near-tie pixel that another implementation thresholds the other way changes the next iteration's input by 1/16 at one
    feature pixel.  Feeding the oracle the masks the implementation under test derived from ITS logits compares every iteration
    on identical inputs; the flipped pixels themselves are counted separately.
    """
    n_ways, n_shots = len(supp_imgs), len(supp_imgs[0])
    n_queries = len(qry_imgs)
    assert n_queries == 1, 'the reference indexes qry_fts[0] only (net/rp_net.py:283)'
    batch = supp_imgs[0][0].shape[0]
    img_size = qry_imgs[0].shape[-2:]
    scale = cfg.get('scale', 4)
    T = cfg['n_iter_refinement']
    radius = cfg['mask_refinement_correlation_radius']

    def encode(x):
        if backbone == 'UNet':
            # net/rp_net.py:248,257: BOTH encoder passes receive fore_mask[0][0] (only read by the mask_feature_map variants)
            return unet_encoder(x, sd, 'encoder.', training, mask=fore_mask[0][0].unsqueeze(1),
                                mask_feature_map=cfg.get('mask_feature_map', False))
        if backbone == 'vgg':      # wiring per SURVEY D1: wrap as d4, caller passes scale=8
            return vgg_encoder(x.expand(-1, 3, -1, -1), sd, 'encoder.')
        if backbone == 'resnet':   # net/rp_net.py:246-249
            return resnet_encoder(x.expand(-1, 3, -1, -1), sd, 'encoder.', training)
        raise NotImplementedError(backbone)

    # :245-262 two separate encoder passes (BN batch statistics are per pass: D14)
    s_fts = encode(torch.cat([torch.cat(way, dim=0) for way in supp_imgs], dim=0))
    fts_size = s_fts.shape[-2:]
    s_fts = s_fts.view(n_ways, n_shots, batch, -1, *fts_size)
    q_fts = encode(torch.cat(qry_imgs, dim=0)).view(n_queries, batch, -1, *fts_size)

    fore = torch.stack([torch.stack(way, dim=0) for way in fore_mask], dim=0)   # Wa x Sh x B x H x W
    back = torch.stack([torch.stack(way, dim=0) for way in back_mask], dim=0)

    qry_mask = F.avg_pool2d(appr_query_labels.unsqueeze(1), scale)               # :269-270
    # :271-275 (per (way, shot) in the extension; [0][0] only in the reference)
    supp = []
    for w in range(n_ways):
        row = []
        for s in range(n_shots):
            m = F.avg_pool2d(fore[w, s].unsqueeze(1), scale)
            row.append(cre(s_fts[w, s] * m, s_fts[w, s] * (1 - m), sd, radius, 'cre.', training, allpairs))
        supp.append(torch.stack(row, dim=0))
    supp = torch.stack(supp, dim=0)                                              # Wa x Sh x B x 64 x H' x W'

    def match(inter):   # :287-305 (prototypes are loop invariant: D6 — recomputed here as the reference does)
        outs = []
        preds = []
        for epi in range(batch):
            protos = prototypes_for(supp, fore, back, epi)
            pred = torch.stack([cal_dist(inter[:, epi], p) for p in protos], dim=1)
            preds.append(pred)
            outs.append(F.interpolate(pred, size=img_size, mode='bilinear'))
        outs = torch.stack(outs, dim=1)
        return outs.view(-1, *outs.shape[2:]), preds

    refinement = {}
    inter = q_fts
    for i in range(T):                                                            # :280-312
        if mask_override is not None and i in mask_override:
            qry_mask = mask_override[i]
        inter = cre(q_fts[0] * qry_mask, q_fts[0] * (1 - qry_mask), sd, radius, 'cre.', training, allpairs,
                    want if (want is not None and i == 0) else None)[None]
        logits, _ = match(inter)
        qry_mask = recurrent_mask(logits, cfg, scale)
        refinement[i] = logits

    # :314-346 the final block recomputes the last iteration (D5) and the align loss
    output, preds = match(inter)
    loss = 0
    if align and training:
        for epi in range(batch):
            loss = loss + align_loss(inter[:, epi], preds[epi], supp[:, :, epi], fore[:, :, epi], back[:, :, epi])
    if want is not None:
        want.update(supp_fts=s_fts, qry_fts=q_fts, supp_cre=supp, inter=inter)
    return {'output': output, 'align_loss': loss / batch, 'refinement': refinement}


def train_loss(out, query_labels, align_loss_scaler=1.0):
    """Reconstructed training loss (the reference ships no train script: SURVEY D9/§3.5):
    sum_i dice_ce(refinement[i], labels) + align_loss_scaler * align_loss
    (yamls/example.yml:94 `align_loss_scaler: 1`, :115 `loss: dice_ce`)."""
    loss = 0
    for i in sorted(out['refinement']):
        loss = loss + dice_ce(out['refinement'][i], query_labels)
    return loss + align_loss_scaler * out['align_loss']
