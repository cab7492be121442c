"""Train step of the RP-Net hot path on the C-ABI kernels: train-mode forward (batch-statistics BatchNorm per reference
call), the hand-scheduled backward of every op on the path, the data-parallel gradient exchange and Adam.

The reference trains through torch autograd over `RP_Net.forward` (net/rp_net.py:226-350) and ships no train script
(SURVEY D9); the step reconstructed in SURVEY §3.5 is
    loss = sum_i dice_ce(out['refinement'][i], query_labels) + align_loss_scaler * out['align_loss']
    Adam(lr 1e-5, weight_decay 1e-4)                                  (yamls/example.yml:64-67,94,115)
`TrainEngine.forward()/backward()` restate exactly the graph autograd would build for that forward; `TrainStep` adds
the loss kernel, the bucketed NCCL all-reduce of the flat gradient buffer and the fused Adam kernel.

Layout: activations fp16 NHWC, activation gradients bf16 NHWC, everything else fp32.  Parameters,
gradients and Adam moments live in flat fp32 buffers; the nn.Parameters of the model are re-pointed into the flat
parameter buffer (state_dict / load_state_dict keep working) and their `.grad` are views of the flat gradient buffer.

With `soft_mask: False` the thresholded mask cuts the graph between refinement iterations (net/rp_net.py:309-311), so
the T query `cre` calls are independent in the backward and are processed as ONE batched launch per kernel
(T call groups).  With `soft_mask: True` the mask m_{i+1} = avg_pool2d(p_fg(logits_i)) stays in the graph: the backward then
walks the iterations from the last to the first, sending each iteration's pre-mask gradient back into the previous
iteration's logits (rpnet_premask_mask_bwd, rpnet_soft_mask_bwd_f32).
"""
import torch
import torch.nn as nn

from . import engine, ops

bf16 = torch.bfloat16
f16 = torch.float16
f32 = torch.float32

UNUSED_PREFIXES = ('cre.w_context.', 'cre.out.')        # never used by the reference forward: grad None


def used_parameters(net):
    return [(n, p) for n, p in net.named_parameters() if not n.startswith(UNUSED_PREFIXES)]


class FlatParams:
    """Flat fp32 parameter / gradient / Adam-moment buffers; model parameters become views of `param`."""

    def __init__(self, net):
        named = used_parameters(net)
        dev = named[0][1].device
        self.names = [n for n, _ in named]
        self.offsets = {}
        off = 0
        for n, p in named:
            self.offsets[n] = (off, p.numel())
            off += (p.numel() + 3) // 4 * 4                      # keep every tensor 16-byte aligned
        self.numel = off
        self.param = torch.zeros(off, dtype=f32, device=dev)
        self.grad = torch.zeros(off, dtype=f32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=f32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=f32, device=dev)
        for n, p in named:
            o, k = self.offsets[n]
            self.param[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.param[o:o + k].view(p.shape)
        self._named = named
        self._first = named[0][1]
        self._by_id = {id(p): n for n, p in named}

    def grad_of(self, p):
        """View of the flat gradient buffer that belongs to parameter `p`."""
        o, k = self.offsets[self._by_id[id(p)]]
        return self.grad[o:o + k].view(p.shape)

    def alias_grads(self):
        """p.grad = view of the flat gradient buffer (TrainStep: what loss.backward() would leave behind, without copies)."""
        for n, p in self._named:
            if p.grad is None or p.grad.data_ptr() != self.grad_of(p).data_ptr():
                p.grad = self.grad_of(p)

    def unalias_grads(self):
        for n, p in self._named:
            if p.grad is not None and p.grad.data_ptr() == self.grad_of(p).data_ptr():
                p.grad = None

    def attached(self):
        return self._first.data_ptr() == self.param.data_ptr()

    def span(self, prefixes):
        """[lo, hi) of the flat range covering every parameter whose name starts with one of `prefixes`."""
        lo, hi = self.numel, 0
        for n in self.names:
            if n.startswith(tuple(prefixes)):
                o, k = self.offsets[n]
                lo, hi = min(lo, o), max(hi, (o + k + 3) // 4 * 4)
        return lo, hi


# Gradient buckets in the order the backward finishes them (reverse of the forward / flat order).
BUCKET_GROUPS = [('cre.',), ('encoder.Up_conv4.', 'encoder.Up4.'), ('encoder.Up_conv5.', 'encoder.Up5.'),
                 ('encoder.Conv5.',), ('encoder.Conv4.', 'encoder.Conv3.', 'encoder.Conv2.', 'encoder.Conv1.')]


class GradBuckets:
    """Bucketed sum all-reduce of the flat gradient buffer (the one exchange step of the path).
    `ready(i)` is called when bucket i's last gradient kernel has been enqueued; on CUDA the collective runs on a side
    stream behind an event so that it overlaps the rest of the backward; `finish()` joins.  The 1/world averaging is
    folded into the Adam kernel's grad_scale."""

    def __init__(self, flat, world_size, group=None, groups=None):
        self.flat, self.world, self.group = flat, world_size, group
        self.ranges = [flat.span(g) for g in (groups or BUCKET_GROUPS)]
        covered = sorted(self.ranges)
        assert covered[0][0] == 0 and covered[-1][1] == flat.numel and all(a[1] == b[0] for a, b in zip(covered, covered[1:])), \
            'gradient buckets must tile the flat buffer exactly once: %r' % (covered,)
        self.cuda = flat.grad.is_cuda
        self.stream = torch.cuda.Stream() if (self.cuda and world_size > 1) else None
        self.pending = []

    def ready(self, i):
        if self.world <= 1:
            return
        import torch.distributed as dist
        lo, hi = self.ranges[i]
        view = self.flat.grad[lo:hi]
        if self.stream is None:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            self.pending.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        for w in self.pending:
            w.wait()
        self.pending = []
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)


def shard_range(total, rank, world):
    """Contiguous slice [lo, hi) of `total` independent slices owned by `rank`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _NoAffineNorm:
    """What the kernels see of an nn.InstanceNorm2d(C) (unet_normalize_type: InstanceNorm2d): gamma = 1, beta = 0, no running
    statistics; the per-image statistics come from one BatchNorm call group per image."""

    def __init__(self, c, dev, eps=1e-5):
        from types import SimpleNamespace
        self.weight = SimpleNamespace(data=torch.ones(c, dtype=f32, device=dev))
        self.bias = SimpleNamespace(data=torch.zeros(c, dtype=f32, device=dev))
        self.running_mean = self.running_var = self.num_batches_tracked = None
        self.eps, self.momentum = eps, 0.1


class _ConvBN:
    """One conv (bias dropped: it cancels in batch- / instance-statistics normalisation) + BatchNorm(batch stats) + ReLU of the path."""

    def __init__(self, name, conv, bn, flat, first=False, hole=(0, 0), up=False, split=False, w_split=None, relu=True, bwd_relu=None):
        # split: the forward runs on (hi, lo) plane pairs; w_split = weight-pack level: engine.W_C8 (fp8 corrections, two units of
        # tensor time, the default), engine.W_SPLIT (Wh | Wl, three fp16 passes), engine.W_FP16 (hi.W + lo.W); all fp32-class but the
        # last.  The backward is unchanged (it reads the hi planes).
        self.name, self.conv, self.bn, self.first, self.hole, self.up, self.split = name, conv, bn, first, hole, up, split
        self.w_split = int(split if w_split is None else w_split)
        # relu: the activation of the forward apply pass; bwd_relu: whether the backward masks by bn(z) > 0 — False for a BasicBlock's
        # second conv, whose ReLU follows the residual add and is applied to the incoming gradient instead (ResNetTrainEngine)
        self.relu, self.bwd_relu = relu, relu if bwd_relu is None else bwd_relu
        if not isinstance(bn, nn.BatchNorm2d):                 # nn.InstanceNorm2d: no parameters, no buffers
            self.bn = bn = _NoAffineNorm(conv.out_channels, conv.weight.device, getattr(bn, 'eps', 1e-5))
        affine = isinstance(bn, nn.BatchNorm2d)
        self.gw = flat.grad_of(conv.weight) if flat is not None else None
        self.ggamma = flat.grad_of(bn.weight) if (flat is not None and affine) else None
        self.gbeta = flat.grad_of(bn.bias) if (flat is not None and affine) else None
        k = conv.kernel_size[0]
        d = conv.dilation[0]
        self.taps = [((ky - k // 2) * d, (kx - k // 2) * d) for ky in range(k) for kx in range(k)]
        self.cout = conv.out_channels
        self.cin_pack = conv.in_channels + hole[1]
        dev = conv.weight.device
        ks = 2 if self.w_split else 1
        if not first:
            self.wf = torch.empty(len(self.taps), self.cout, self.cin_pack * ks, dtype=f16, device=dev)
            self.wd = torch.empty(len(self.taps), self.cin_pack, self.cout, dtype=bf16, device=dev)
        if up:      # up_conv in sub-pixel form: phase packs (forward) and the 4x4 stride-2 pack (data gradient)
            self.wf_up = torch.empty(4, 4, self.cout, conv.in_channels * ks, dtype=f16, device=dev)
            self.w16_up = torch.empty(16, conv.in_channels, self.cout, dtype=bf16, device=dev)
        self.ones = torch.ones(self.cout, dtype=f32, device=dev)
        self.zeros = torch.zeros(self.cout, dtype=f32, device=dev)

    def pack(self):
        if not self.first:
            ops.pack_conv_weight(self.conv.weight.data, self.wf, self.wd, hole=self.hole, split=self.w_split)
        if self.up:
            ops.pack_upconv_weight(self.conv.weight.data, self.wf_up, self.w16_up, split=self.w_split)

    def fwd_up(self, x_low, z, gs, sums, stats, y):
        """z = conv3x3(upsample2x(x_low)) in sub-pixel form (four phase convs, statistics in their epilogues) + BN + ReLU.
        x_low, z, y: (hi, lo) plane pairs (lo None outside split mode)."""
        n, H, W, c = z[0].shape
        bn = self.bn
        if self.split:
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                taps = [((-1 if py == 0 else 0) + (t >> 1), (-1 if px == 0 else 0) + (t & 1)) for t in range(4)]
                ops.conv_split(x_low[0], self.wf_up[ph], taps, self.ones, self.zeros, False, src0_lo=x_low[1], w_split=self.w_split,
                               out=z[0], out_lo=z[1], out_map=(2, py, 2, px), group_start=gs, sums=sums, keep_sums=ph > 0)
        else:
            ops.upconv_fwd_bnstats(x_low[0], self.wf_up, self.ones, self.zeros, z[0], gs, sums)
        ops.bn_finalize(sums, gs, c, H * W, bn.weight.data, bn.bias.data, self.conv.bias.data, bn.running_mean, bn.running_var,
                        bn.num_batches_tracked, stats, eps=bn.eps, momentum=bn.momentum)
        ops.bn_apply(z[0], stats, gs, True, y=y[0], z_lo=z[1], y_lo=y[1])

    def bwd_up(self, eng, x_low, z, stats, gs, dx_low, **src):
        """Backward of fwd_up: BN+ReLU backward -> dz (high resolution); weight gradient from the four phase GEMMs; data
        gradient w.r.t. the LOW-resolution input (4x4 stride-2 conv of dz)."""
        n, H, W, c = z.shape
        dz = eng.scratch('dz', n * H * W * c, bf16).view(n, H, W, c)
        ops.bn_bwd(z, stats, gs, dz, eng.scratch('bn_bwd', (len(gs) - 1) * c * 6, f32), True, dgamma=self.ggamma, dbeta=self.gbeta, **src)
        h, w, cin = x_low.shape[1:]
        nb = ops.upconv_wgrad_workspace_bytes(cin, n, h, w, self.cout)
        ops.upconv_wgrad(x_low, dz, self.gw, eng.scratch('wgrad', nb // 4, f32))
        ops.upconv_dgrad(dz, self.w16_up, dx_low)
        return dx_low

    def fwd(self, x0, x1, z, gs, sums, stats, y=None, pool=None, y32=None, res=None):
        """z = conv(x0 | x1) and its batch statistics (fused into the conv epilogue), then BatchNorm(train) (+ res) + ReLU.
        x0, x1, z, y, pool, res: (hi, lo) plane pairs (lo None outside split mode; x0 = the fp32 image for the first conv)."""
        n, h, w, c = z[0].shape
        bn = self.bn
        y = y or (None, None)
        pool = pool or (None, None)
        res = res or (None, None)
        bias = self.conv.bias.data if self.conv.bias is not None else None
        if self.first:
            ops.conv3x3_first(x0, self.conv.weight.data, self.ones, self.zeros, False, z[0], out_lo=z[1])
            ops.bn_stats(z[0], gs, sums, z_lo=z[1])
        elif self.split:
            ops.conv_split(x0[0], self.wf, self.taps, self.ones, self.zeros, False, src0_lo=x0[1], src1=None if x1 is None else x1[0],
                           src1_lo=None if x1 is None else x1[1], w_split=self.w_split, out=z[0], out_lo=z[1], group_start=gs, sums=sums)
        else:
            ops.conv_bnstats(x0[0], self.wf, self.taps, self.ones, self.zeros, z[0], gs, sums, src1=None if x1 is None else x1[0])
        ops.bn_finalize(sums, gs, c, h * w, bn.weight.data, bn.bias.data, bias, bn.running_mean, bn.running_var,
                        bn.num_batches_tracked, stats, eps=bn.eps, momentum=bn.momentum)
        ops.bn_apply(z[0], stats, gs, self.relu, y=y[0], y_pool=pool[0], y_f32=y32, z_lo=z[1], y_lo=y[1], y_pool_lo=pool[1], res=res[0],
                     res_lo=res[1])

    def bwd(self, eng, x0, x1, z, stats, gs, dx=None, accumulate=False, **src):
        """BN+ReLU backward -> dz; weight gradient (= or += with `accumulate`); data gradient into `dx` (bf16 [n,h,w,cin_pack])
        when given."""
        n, h, w, c = z.shape
        dz = eng.scratch('dz', n * h * w * c, bf16).view(n, h, w, c)
        ops.bn_bwd(z, stats, gs, dz, eng.scratch('bn_bwd', (len(gs) - 1) * c * 6, f32), self.bwd_relu,
                   dgamma=self.ggamma, dbeta=self.gbeta, **src)
        if self.first:
            ops.conv3x3_first_wgrad(x0, dz, self.gw)
        else:
            c0 = x0.shape[3]
            c1 = 0 if x1 is None else x1.shape[3]
            nb = ops.conv_wgrad_workspace_bytes(c0, c1, n, h, w, len(self.taps), self.cout)
            ops.conv_wgrad(x0, dz, self.taps, self.gw, eng.scratch('wgrad', nb // 4, f32), x1=x1, hole=self.hole,
                           accumulate=accumulate)
            if dx is not None:
                ops.conv_dgrad(dz, self.wd, self.taps, dx)
        return dx


class EncoderEngine:
    """The two-pass (statistics, then apply) schedule of the U-Net encoder on the C-ABI kernels: conv (split-fp16 or fp16) with the
    normalisation statistics in its epilogue -> finalize -> apply (+ReLU, + fused 2x2 max-pool).  TrainEngine builds on it
    (train-mode BatchNorm: one call group per reference call); on its own it is the eval forward of `unet_normalize_type:
    InstanceNorm2d` (one call group per image, no running statistics)."""
    MAX_GROUPS = 64                                    # kMaxGroups of the BatchNorm kernels

    def __init__(self, encoder, norm='batch', flat=None, cfg=None):
        dev = next(encoder.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('rpnet_b200 runs on CUDA (sm_100a) only; there is no CPU fallback')
        self.mfm = getattr(encoder, 'mfm', False)
        self.encoder, self.dev, self.norm, self.flat = encoder, dev, norm, flat
        # encoder forward in split-fp16 (fp32-class, the default) or plain fp16 (`b200_precision: fp16`, TF32-class: faster,
        # train-mode logits 2e-3 .. 5e-3 from the fp32 reference)
        pr = engine.precision_of(cfg if cfg is not None else encoder.cfg)
        self.split = sp = engine.is_split(pr)
        lv = engine.w_level(pr)
        self.lo_level = engine.W_C8 if lv == engine.W_C8 else engine.W_SPLIT      # format of the activations' lo planes
        wd = engine.w_level(engine.decoder_precision(pr))  # decoder half: RPNET_SPLIT_DECODER=2 -> 'split' activations, fp16 weights
        e, L = encoder, {}
        for nm, blk in (('c1', e.Conv1), ('c2', e.Conv2), ('c3', e.Conv3), ('c4', e.Conv4), ('c5', e.Conv5),
                        ('uc5', e.Up_conv5), ('uc4', e.Up_conv4)):
            ws = wd if nm.startswith('uc') else lv
            # mask_feature_map x2 / x3 (net/unet.py:401-424): the 65- / 129-channel conv runs as 128 / 192 packed channels, the mask
            # travels in channel 0 of a 64-channel extra source and the 63 padding channels are a hole of the packs
            hole = (blk.conv[0].in_channels, 63) if (nm, self.mfm) in (('c2', 'x2'), ('c3', 'x3')) else (0, 0)
            L[nm + 'a'] = _ConvBN(nm + 'a', blk.conv[0], blk.conv[1], flat, first=(nm == 'c1'), hole=hole, split=sp, w_split=ws)
            L[nm + 'b'] = _ConvBN(nm + 'b', blk.conv[3], blk.conv[4], flat, split=sp, w_split=ws)
        L['up5'] = _ConvBN('up5', e.Up5.up[1], e.Up5.up[2], flat, up=True, split=sp, w_split=wd)
        L['up4'] = _ConvBN('up4', e.Up4.up[1], e.Up4.up[2], flat, up=True, split=sp, w_split=wd)
        self.L = L
        self.ws = engine.Workspace()
        self._scratch = {}
        self.act = {}
        self._first = e.Conv1.conv[0].weight
        self._first_ptr = self._first.data_ptr()
        self._pack_sig = None

    def attached(self):
        """False once the encoder's parameters were moved (net.to(other_device)): build a new engine."""
        return self._first.data_ptr() == self._first_ptr and self._first.device == self.dev

    def groups_for(self, calls):
        """BatchNorm2d: the call groups given (one per reference call); InstanceNorm2d: one group per image."""
        if self.norm == 'instance':
            n = calls[-1]
            if n > self.MAX_GROUPS:
                raise ValueError('InstanceNorm2d: at most %d images per encoder pass (got %d)' % (self.MAX_GROUPS, n))
            return list(range(n + 1))
        return calls

    def encode(self, imgs, mask=None):
        """Eval forward (InstanceNorm2d): fp32 NCHW images -> d4 fp16 NHWC, in passes of at most MAX_GROUPS images (mask: one per
        image, only read by the mask_feature_map variants)."""
        sig = (engine.WEIGHTS_EPOCH,) + tuple((p.data_ptr(), p._version) for p in self.encoder.parameters())
        if sig != self._pack_sig:
            self.pack_weights()
            self._pack_sig = sig
        n, _, H, W = imgs.shape
        out = self.buf('encode.d4', (n, H // 4, W // 4, self.L['uc4b'].cout), f16)
        for lo in range(0, n, self.MAX_GROUPS):
            hi = min(n, lo + self.MAX_GROUPS)
            self.act = {}
            d4 = self._encoder_fwd(imgs[lo:hi].contiguous(), self.groups_for([0, hi - lo]),
                                   None if mask is None else mask[lo:hi].contiguous())[0]
            out[lo:hi].copy_(d4)
        self.act = {}
        return out

    # ------------------------------------------------------------------ buffers
    def buf(self, name, shape, dtype):
        return self.ws.get(name, shape, dtype, self.dev)

    def scratch(self, name, numel, dtype):
        """Grow-only scratch (contents dead after the call that uses it)."""
        t = self._scratch.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(int(numel), dtype=dtype, device=self.dev)
            self._scratch[name] = t
        return t[:numel]

    def pack_weights(self):
        """fp32 parameters -> the fp16 (split: Wh | Wl) forward and bf16 data-gradient packs, after every optimizer step: all
        tap-list convs in one launch, the two sub-pixel up-conv packs in one launch each."""
        if getattr(self, '_pack_table', None) is None:
            layers = [(l.conv.weight.data, l.wf, l.wd, l.hole, l.w_split) for l in self.L.values() if not l.first]
            self._pack_table = ops.pack_conv_weights(layers)
        ops.run_pack_conv_weights(*self._pack_table)
        for l in self.L.values():
            if l.up:
                ops.pack_upconv_weight(l.conv.weight.data, l.wf_up, l.w16_up, split=l.w_split)

    # ------------------------------------------------------------------ encoder
    def pair(self, name, shape, lo, z=False):
        """(hi, lo) planes of one activation; lo is None outside split mode, the fp16 residual plane for a pre-normalisation conv
        output `z` (|mean| >> std: it needs the full residual) and in 'split' precision, the c8 plane in 'split8'."""
        if not lo:
            return self.buf(name, shape, f16), None
        if z or self.lo_level != engine.W_C8:
            return self.buf(name, shape, f16), self.buf(name + '.lo', shape, f16)
        return self.buf(name, shape, f16), self.buf(name + '.lo', tuple(shape[:-1]) + (2 * shape[-1],), torch.uint8)

    def _layer_fwd(self, key, x0, x1, gs, want_y=True, want_pool=False):
        """x0 / x1: (hi, lo) pairs (x0 = the fp32 image batch for the first conv).  Returns the (hi, lo) pairs of y and of
        its 2x2 max-pooled copy.  The backward only keeps the hi planes."""
        l = self.L[key]
        n, h, w = (x0.shape[0], x0.shape[2], x0.shape[3]) if l.first else x0[0].shape[:3]
        c = l.cout
        z = self.pair(key + '.z', (n, h, w, c), l.split, z=True)
        stats = self.buf(key + '.stats', (len(gs) - 1, c, 4), f32)
        y = self.pair(key + '.y', (n, h, w, c), l.split) if want_y else None
        pool = self.pair(key + '.pool', (n, h // 2, w // 2, c), l.split) if want_pool else None
        l.fwd(x0, x1, z, gs, self.scratch('bn_sums', (len(gs) - 1) * c * 2, torch.float64), stats, y=y, pool=pool)
        self.act[key] = dict(x0=x0 if l.first else x0[0], x1=None if x1 is None else x1[0], z=z[0], stats=stats, gs=gs)
        return y, pool

    def _encoder_fwd(self, imgs, gs, mask=None):
        """net/unet.py:435-467 in train mode; `gs` = BatchNorm call groups (support pass, query pass: net/rp_net.py:248,257);
        mask [n, 1, H, W]: only read by the mask_feature_map variants (net/unet.py:437-449)."""
        f = self._layer_fwd
        m2 = m3 = None
        if self.mfm:
            if mask is None or mask.shape[0] != imgs.shape[0]:
                raise ValueError('mask_feature_map=%r needs one mask per image' % (self.mfm,))
            src = lambda pool, nm: (lambda m: m if isinstance(m, tuple) else (m, None))(
                self.encoder._mask_source(mask, pool, self.ws, nm, self.split))
            if self.mfm == 'x':
                imgs = torch.cat([imgs, mask.float()], dim=1).contiguous()            # net/unet.py:437-438
            m2 = src(2, 'mfm.m2') if self.mfm == 'x2' else None                        # :442-443
            m3 = src(4, 'mfm.m3') if self.mfm == 'x3' else None                        # :446-447
        a, _ = f('c1a', imgs, None, gs)
        _, p1 = f('c1b', a, None, gs, want_y=False, want_pool=True)
        a, _ = f('c2a', p1, m2, gs)
        _, p2 = f('c2b', a, None, gs, want_y=False, want_pool=True)
        a, _ = f('c3a', p2, m3, gs)
        x3, p3 = f('c3b', a, None, gs, want_pool=True)
        a, _ = f('c4a', p3, None, gs)
        x4, p4 = f('c4b', a, None, gs, want_pool=True)
        a, _ = f('c5a', p4, None, gs)
        x5, _ = f('c5b', a, None, gs)
        u5 = self._up_fwd('up5', x5, gs)                             # nn.Upsample(x2) + conv, net/modules.py:61-75
        a, _ = f('uc5a', x4, u5, gs)                                  # torch.cat((x4, d5), dim=1)  net/unet.py:460
        d5, _ = f('uc5b', a, None, gs)
        u4 = self._up_fwd('up4', d5, gs)
        a, _ = f('uc4a', x3, u4, gs)                                  # torch.cat((x3, d4), dim=1)  net/unet.py:464
        d4, _ = f('uc4b', a, None, gs)
        return d4

    def _up_fwd(self, key, x_low, gs):
        """up_conv: sub-pixel form on the low-resolution input when the maps are at least one pixel tile large, else the
        materialised nearest-x2 map + 3x3 conv."""
        l = self.L[key]
        n, h, w, cin = x_low[0].shape
        if ops.upconv_fusable(h, w):
            c = l.cout
            z = self.pair(key + '.z', (n, 2 * h, 2 * w, c), l.split, z=True)
            stats = self.buf(key + '.stats', (len(gs) - 1, c, 4), f32)
            y = self.pair(key + '.y', (n, 2 * h, 2 * w, c), l.split)
            l.fwd_up(x_low, z, gs, self.scratch('bn_sums', (len(gs) - 1) * c * 2, torch.float64), stats, y)
            self.act[key] = dict(x0=x_low[0], x1=None, z=z[0], stats=stats, gs=gs, sub=True)
            return y
        u = self.pair(key + '.in', (n, 2 * h, 2 * w, cin), l.split)
        ops.upsample2x(x_low[0], u[0])                                # net/modules.py:67
        if l.split:                                                   # a c8 plane is copied as the fp16 plane of the same bytes
            ops.upsample2x(x_low[1].view(f16), u[1].view(f16))
        y, _ = self._layer_fwd(key, u, None, gs)
        return y

    def _up_bwd(self, key, below, **src):
        """Backward of an up_conv layer followed by the BN backward of the layer `below` that produced its input: returns the
        gradient w.r.t. the input of `below`'s conv."""
        a = self.act[key]
        if a.get('sub'):
            dx_low = self.buf(key + '.dx', tuple(a['x0'].shape), bf16)
            self.L[key].bwd_up(self, a['x0'], a['z'], a['stats'], a['gs'], dx_low, **src)
            return self._layer_bwd(below, tuple(self.act[below]['x0'].shape), direct=dx_low)
        du = self._layer_bwd(key, tuple(a['x0'].shape), **src)
        return self._layer_bwd(below, tuple(self.act[below]['x0'].shape), up=du)

    def _layer_bwd(self, key, dx_shape=None, **src):
        a = self.act[key]
        dx = self.buf(key + '.dx', dx_shape, bf16) if dx_shape is not None else None
        return self.L[key].bwd(self, a['x0'], a['x1'], a['z'], a['stats'], a['gs'], dx=dx, **src)

    def _encoder_bwd(self, g_d4, buckets=None):
        b = self._layer_bwd
        A = self.act
        shp = lambda key: tuple(A[key]['x0'].shape)
        cat = lambda key: tuple(A[key]['x0'].shape[:3]) + (A[key]['x0'].shape[3] + A[key]['x1'].shape[3],)
        g = b('uc4b', shp('uc4b'), direct=g_d4)
        dcat4 = b('uc4a', cat('uc4a'), direct=g)
        c3 = A['uc4a']['x0'].shape[3]
        g = self._up_bwd('up4', 'uc5b', direct=dcat4, d_off=c3)
        if buckets:
            buckets.ready(1)
        dcat5 = b('uc5a', cat('uc5a'), direct=g)
        c4 = A['uc5a']['x0'].shape[3]
        g = self._up_bwd('up5', 'c5b', direct=dcat5, d_off=c4)
        if buckets:
            buckets.ready(2)
        dp4 = b('c5a', shp('c5a'), direct=g)
        if buckets:
            buckets.ready(3)
        g = b('c4b', shp('c4b'), direct=dcat5, d_off=0, pooled=dp4)
        dp3 = b('c4a', shp('c4a'), direct=g)
        g = b('c3b', shp('c3b'), direct=dcat4, d_off=0, pooled=dp3)
        full = lambda key: cat(key) if A[key]['x1'] is not None else shp(key)     # mask_feature_map x2 / x3: gradient of the packed
        dp2 = b('c3a', full('c3a'), direct=g)                                      # input; the consumer reads its first 64 / 128 channels
        g = b('c2b', shp('c2b'), pooled=dp2)
        dp1 = b('c2a', full('c2a'), direct=g)
        n, _, H, W = A['c1a']['x0'].shape
        g = b('c1b', (n, H, W, 64), pooled=dp1)
        b('c1a', None, direct=g)
        if buckets:
            buckets.ready(4)

class TrainEngine(EncoderEngine):
    """Train-mode forward + backward of RP_Net (U-Net backbone) on the C-ABI kernels."""

    def __init__(self, net):
        from .nn.unet import U_Net
        if not isinstance(net.encoder, U_Net):
            raise NotImplementedError("TrainEngine schedules backbone 'UNet'; 'vgg' / 'resnet' have VggTrainEngine / ResNetTrainEngine "
                                      "(TrainEngine.of(net) picks the right one)")
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('rpnet_b200 trains on CUDA (sm_100a) only; there is no CPU fallback')
        self.net = net
        flat = FlatParams(net)
        # the cre convs are single-term fp16 in both precisions (their rounding moves the logits by ~1e-4) and
        # always BatchNorm2d (net/rp_net.py:50-69)
        super().__init__(net.encoder, norm='instance' if net.encoder.inorm else 'batch', flat=flat, cfg=net.backbone_cfg)
        c, L = net.cre, self.L
        k = (2 * c.radius + 1) ** 2
        self.kcorr, self.corr_c = k, c.corr_channels
        L['wk'] = _ConvBN('wk', c.w_k[0], c.w_k[1], self.flat)
        L['wq'] = _ConvBN('wq', c.w_q[0], c.w_q[1], self.flat)
        L['q'] = _ConvBN('q', c.q[0], c.q[1], self.flat, hole=(k, self.corr_c - k))
        self.saved = None

    @staticmethod
    def of(net):
        """The (cached) engine of `net`; rebuilt when the parameters were moved (e.g. net.to(other_device))."""
        eng = net.__dict__.get('_b200_train_engine')
        if eng is None or not eng.flat.attached():
            from .nn.resnet import ResNet18
            from .nn.vgg import Encoder
            eng = VggTrainEngine(net) if isinstance(net.encoder, Encoder) else \
                (ResNetTrainEngine(net) if isinstance(net.encoder, ResNet18) else TrainEngine(net))
            net.__dict__['_b200_train_engine'] = eng
        return eng

    # ------------------------------------------------------------------ context-relation encoder
    def _cre_fwd(self, d4, mask, lo, hi, g0, gs):
        """ContextCorrelationEncoder.forward (net/rp_net.py:77-84) in train mode on images [lo, hi) of the batched cre
        buffers; `gs` = call groups relative to lo, stored from stats row g0."""
        S = self.cre
        L = self.L
        n = hi - lo
        xfg, xbg = S['xfg'][lo:hi], S['xbg'][lo:hi]
        ops.premask(d4, mask, xfg, xbg)
        G = len(gs) - 1
        sums = self.scratch('bn_sums', G * max(256, d4.shape[-1]) * 2, torch.float64)
        P = lambda t: (t[lo:hi], None)                                # the cre convs are single-term fp16: no residual planes
        L['wk'].fwd((xfg, None), None, P(S['z1']), gs, sums, S['st1'][g0:g0 + G], y=P(S['fm1']))
        L['wq'].fwd((xbg, None), None, P(S['z2']), gs, sums, S['st2'][g0:g0 + G], y=P(S['fm2']))
        ops.local_corr(S['fm1'][lo:hi], S['fm2'][lo:hi], self.net.cre.radius, S['corr'][lo:hi])
        L['q'].fwd(P(S['corr']), P(S['fm1']), P(S['z3']), gs, sums, S['st3'][g0:g0 + G], y32=S['feat'][lo:hi])
        return S['feat'][lo:hi]

    # ------------------------------------------------------------------ forward
    def forward(self, d):
        """d: dict with the RP_Net.forward arguments (supp_imgs, fore_mask, back_mask, qry_imgs, appr_query_labels) on the
        device.  Returns (logits [T, B, 1+Wa, H, W] fp32 = out['refinement'][i], align_loss [1] fp32 tensor)."""
        net = self.net
        if not self.flat.attached():
            raise RuntimeError('model parameters were moved after TrainEngine was created; build a new TrainEngine')
        supp_imgs, qry_imgs = d['supp_imgs'], d['qry_imgs']
        Wa, Sh = len(supp_imgs), len(supp_imgs[0])
        if len(qry_imgs) != 1:
            raise NotImplementedError('the reference forward only consumes qry_imgs[0] (net/rp_net.py:283)')
        B = supp_imgs[0][0].shape[0]
        H, W = qry_imgs[0].shape[-2:]
        S, T, P = net.scale, net.num_iter, 1 + Wa
        n_supp = Wa * Sh * B
        n_img = n_supp + B
        self.act = {}
        engine.WEIGHTS_EPOCH += 1          # BN running statistics are updated through raw pointers below
        budget = self.__dict__.setdefault('_shapes', engine.ShapeBudget())
        if budget.note((Wa, Sh, B, H, W)):                # varying batch / crop sizes must not grow device memory without bound
            torch.cuda.synchronize()
            self.ws.clear()
            self._scratch.clear()
            self.saved = None

        imgs = torch.cat([torch.cat(way, dim=0) for way in supp_imgs] + [qry_imgs[0]], dim=0).float().contiguous()
        fore = torch.stack([torch.stack(way, dim=0) for way in d['fore_mask']], dim=0).float().reshape(n_supp, H, W).contiguous()
        back = torch.stack([torch.stack(way, dim=0) for way in d['back_mask']], dim=0).float().reshape(n_supp, H, W).contiguous()
        emask = net._encoder_mask(fore, Wa, Sh, B)                     # fore_mask[0][0] for both passes (mask_feature_map variants only)
        d4 = self._encoder_fwd(imgs, self.groups_for([0, n_supp, n_img]), emask)[0]   # two BN calls: support, query pass (D14); hi plane
        h, w, C = d4.shape[1:]
        if h * S != H or w * S != W:
            raise ValueError('scale=%d does not match the encoder stride' % S)

        n_tot = n_supp + T * B
        G_tot = Wa * Sh + T
        buf = self.buf
        self.cre = dict(
            xfg=buf('cre.xfg', (n_tot, h, w, C), f16), xbg=buf('cre.xbg', (n_tot, h, w, C), f16),
            z1=buf('cre.z1', (n_tot, h, w, C), f16), z2=buf('cre.z2', (n_tot, h, w, C), f16),
            fm1=buf('cre.fm1', (n_tot, h, w, C), f16), fm2=buf('cre.fm2', (n_tot, h, w, C), f16),
            corr=buf('cre.corr', (n_tot, h, w, self.corr_c), f16), z3=buf('cre.z3', (n_tot, h, w, 64), f16),
            feat=buf('cre.feat', (n_tot, h, w, 64), f32),
            st1=buf('cre.st1', (G_tot, C, 4), f32), st2=buf('cre.st2', (G_tot, C, 4), f32), st3=buf('cre.st3', (G_tot, 64, 4), f32))

        # support branch: one cre CALL per (way, shot) (net/rp_net.py:271-275; oracle-ext), batched as Wa*Sh call groups
        supp_m = buf('supp_m', (n_supp, h, w), f32)
        ops.avgpool_mask(fore, S, supp_m)
        supp_feat = self._cre_fwd(d4[:n_supp], supp_m, 0, n_supp, 0, [i * B for i in range(Wa * Sh + 1)])

        # prototypes (net/rp_net.py:288-299, 366-391), hoisted out of the T loop (D6)
        wf, wb = buf('wmap_f', (n_supp, h, w), f32), buf('wmap_b', (n_supp, h, w), f32)
        sf, sb = buf('msum_f', (n_supp,), f32), buf('msum_b', (n_supp,), f32)
        ops.bilinear_adjoint(fore, wf, sf)
        ops.bilinear_adjoint(back, wb, sb)
        raw = buf('proto_raw', (Wa, Sh, B, 2, 64), f32)
        ops.weighted_pool(supp_feat, wf, wb, sf, sb, raw.view(n_supp, 2, 64))
        protos = buf('protos', (B, P, 64), f32)
        ops.proto_finalize(raw, protos)

        # recurrent refinement (net/rp_net.py:280-312)
        soft = bool(net.backbone_cfg['soft_mask'])                      # net/rp_net.py:309: `soft_mask == False` thresholds
        qm = buf('qry_m', (T + 1, B, h, w), f32)
        ops.avgpool_mask(d['appr_query_labels'].reshape(B, H, W).float().contiguous(), S, qm[0])
        pred = buf('pred', (T, B, P, h, w), f32)
        logits = buf('logits', (T, B, P, H, W), f32)
        qd4 = d4[n_supp:]
        for i in range(T):
            lo = n_supp + i * B
            qfeat = self._cre_fwd(qd4, qm[i], lo, lo + B, Wa * Sh + i, [0, B])
            ops.cos_sim(qfeat, protos, pred[i], 20.0)
            ops.upsample_tail(pred[i], logits[i], qm[i + 1], S, soft)

        # alignLoss (net/rp_net.py:340-343, 394-440) on the last iteration's features / prediction (D5)
        align = buf('align', (1,), f32)
        use_align = bool(net.config.get('align', False))
        if use_align:
            lo = n_supp + (T - 1) * B
            qproto, counts = buf('al.qproto', (B, P, 64), f32), buf('al.counts', (B, P), f32)
            amax = buf('al.amax', (B, h, w), torch.int32)
            ops.class_pool(self.cre['feat'][lo:lo + B], pred[T - 1], qproto, counts, amax)
            ps, wgt = buf('al.ps', (n_supp, 2, 64), f32), buf('al.w', (n_supp,), f32)
            ops.align_gather(qproto, counts, Wa, Sh, 1.0, ps, wgt)
            pred_s = buf('al.pred_s', (n_supp, 2, h, w), f32)
            ops.cos_sim(supp_feat, ps, pred_s, 20.0)
            lg = buf('al.lg', (n_supp, 2, H, W), f32)
            ops.bilinear_up(pred_s.view(n_supp * 2, h, w), lg.view(n_supp * 2, H, W))
            self._align_args = (lg, fore, back, wgt)
            ops.ce_mask(lg, fore, back, wgt, self.scratch('al.sums', n_supp * 2, torch.float64), align)
        else:
            align.zero_()
        self.saved = dict(Wa=Wa, Sh=Sh, B=B, H=H, W=W, h=h, w=w, C=C, T=T, P=P, n_supp=n_supp, n_tot=n_tot, d4=d4, supp_m=supp_m,
                          wf=wf, wb=wb, sf=sf, sb=sb, protos=protos, qm=qm, pred=pred, logits=logits, use_align=use_align, soft=soft)
        return logits, align

    # ------------------------------------------------------------------ backward
    def backward(self, dlogits, dalign=1.0, buckets=None):
        """dlogits: fp32 [T, B, P, H, W] = d loss / d out['refinement'][i]; dalign = d loss / d out['align_loss'] (float).
        Fills the flat gradient buffer (+=: zero it first with zero_grad())."""
        s = self.saved
        if s is None:
            raise RuntimeError('backward() without forward()')
        Wa, Sh, B, H, W, h, w, C, T, P = (s[k] for k in ('Wa', 'Sh', 'B', 'H', 'W', 'h', 'w', 'C', 'T', 'P'))
        n_supp, n_tot = s['n_supp'], s['n_tot']
        S, L, buf = self.cre, self.L, self.buf
        G_tot = Wa * Sh + T
        gs_all = [i * B for i in range(G_tot + 1)]

        dpred = buf('b.dpred', (T, B, P, h, w), f32)
        dfeat = buf('b.dfeat', (n_tot, h, w, 64), f32)
        dprotos = buf('b.dprotos', (B, P, 64), f32)
        dq = buf('b.dq', (n_tot, h, w, self.corr_c + C), bf16)
        df1, df2 = buf('b.df1', (n_tot, h, w, C), bf16), buf('b.df2', (n_tot, h, w, C), bf16)
        dxfg, dxbg = buf('b.dxfg', (n_tot, h, w, C), bf16), buf('b.dxbg', (n_tot, h, w, C), bf16)
        r = self.net.cre.radius

        def cre_bwd(lo, hi, g0, gs, accumulate):
            """Backward of the cre calls on images [lo, hi) of the batched buffers (call groups `gs`, statistics rows from g0)."""
            sl, G = slice(lo, hi), len(gs) - 1
            L['q'].bwd(self, S['corr'][sl], S['fm1'][sl], S['z3'][sl], S['st3'][g0:g0 + G], gs, dx=dq[sl], accumulate=accumulate, direct=dfeat[sl])
            ops.local_corr_bwd(S['fm1'][sl], S['fm2'][sl], dq[sl], self.corr_c, r, df1[sl], df2[sl],
                               workspace=self.scratch('corr_bwd', ops.local_corr_bwd_workspace_bytes(hi - lo, h, w, r) // 2, bf16))
            L['wk'].bwd(self, S['xfg'][sl], None, S['z1'][sl], S['st1'][g0:g0 + G], gs, dx=dxfg[sl], accumulate=accumulate, direct=df1[sl])
            L['wq'].bwd(self, S['xbg'][sl], None, S['z2'][sl], S['st2'][g0:g0 + G], gs, dx=dxbg[sl], accumulate=accumulate, direct=df2[sl])

        def align_bwd():
            """alignLoss backward (net/rp_net.py:394-440): into the support features and the last iteration's query features."""
            lg, fore, back, wgt = self._align_args
            dlg = self.scratch('b.dlg', lg.numel(), f32).view(lg.shape)
            ops.ce_mask(lg, fore, back, wgt, self.scratch('al.sums', n_supp * 2, torch.float64), buf('b.align', (1,), f32), dlg, float(dalign))
            dpred_s = buf('b.dpred_s', (n_supp, 2, h, w), f32)
            ops.bilinear_adjoint(dlg.view(n_supp * 2, H, W), dpred_s.view(n_supp * 2, h, w))
            dps = buf('b.dps', (n_supp, 2, 64), f32)
            ops.cos_sim_bwd(S['feat'][:n_supp], buf('al.ps', (n_supp, 2, 64), f32), dpred_s, dfeat[:n_supp], dps, 20.0)
            dqp = buf('b.dqp', (B, P, 64), f32)
            ops.align_scatter(dps, Wa, Sh, dqp)
            lo = n_supp + (T - 1) * B
            ops.class_pool_bwd(dqp, buf('al.counts', (B, P), f32), buf('al.amax', (B, h, w), torch.int32), dfeat[lo:lo + B])
        acc = bool(s['use_align'] and dalign != 0.0)

        if not s['soft']:
            # logits -> pred (adjoint of the bilinear upsample, net/rp_net.py:303) -> query features / prototypes (calDist)
            ops.bilinear_adjoint(dlogits.view(T * B * P, H, W), dpred.view(T * B * P, h, w))
            ops.cos_sim_bwd(S['feat'][n_supp:], s['protos'], dpred.view(T * B, P, h, w), dfeat[n_supp:], dprotos, 20.0)
            if acc:
                align_bwd()
            draw = buf('b.draw', (Wa, Sh, B, 2, 64), f32)
            ops.proto_finalize_bwd(dprotos, draw)
            ops.weighted_pool_bwd(draw.view(n_supp, 2, 64), s['wf'], s['wb'], s['sf'], s['sb'], dfeat[:n_supp], accumulate=acc)
            # cre backward, all Wa*Sh + T calls as one batched launch per kernel
            cre_bwd(0, n_tot, 0, gs_all, False)
        else:
            # soft mask: iteration i's logits also feed iteration i + 1 through m_{i+1} = avg_pool2d(p_fg(logits_i)) — walk the
            # iterations backwards, folding each pre-mask gradient into the previous iteration's dlogits before it is consumed
            dlogits = dlogits.clone()
            dprotos_t = buf('b.dprotos_t', (T, B, P, 64), f32)
            dm = buf('b.dm', (B, h, w), f32)
            qd4 = s['d4'][n_supp:]
            for i in range(T - 1, -1, -1):
                lo = n_supp + i * B
                ops.bilinear_adjoint(dlogits[i].view(B * P, H, W), dpred[i].view(B * P, h, w))
                ops.cos_sim_bwd(S['feat'][lo:lo + B], s['protos'], dpred[i], dfeat[lo:lo + B], dprotos_t[i], 20.0)
                if acc and i == T - 1:
                    align_bwd()
                cre_bwd(lo, lo + B, Wa * Sh + i, [0, B], True)
                if i >= 1:
                    ops.premask_mask_bwd(dxfg[lo:lo + B], dxbg[lo:lo + B], qd4, dm)
                    ops.soft_mask_bwd(s['logits'][i - 1], dm, self.net.scale, dlogits[i - 1])
            torch.sum(dprotos_t, dim=0, out=dprotos)          # [B, P, 64]: plumbing-sized
            draw = buf('b.draw', (Wa, Sh, B, 2, 64), f32)
            ops.proto_finalize_bwd(dprotos, draw)
            ops.weighted_pool_bwd(draw.view(n_supp, 2, 64), s['wf'], s['wb'], s['sf'], s['sb'], dfeat[:n_supp], accumulate=acc)
            cre_bwd(0, n_supp, 0, [i * B for i in range(Wa * Sh + 1)], True)
        if buckets:
            buckets.ready(0)
        g_d4 = buf('b.g_d4', tuple(s['d4'].shape), bf16)
        ops.premask_bwd(dxfg[:n_supp], dxbg[:n_supp], s['supp_m'], g_d4[:n_supp], iters=1)
        ops.premask_bwd(dxfg[n_supp:], dxbg[n_supp:], s['qm'][:T], g_d4[n_supp:], iters=T)
        self._encoder_bwd(g_d4, buckets)

    def zero_grad(self):
        self.flat.grad.zero_()


class VggTrainEngine(TrainEngine):
    """TrainEngine for `backbone: vgg` (`scale: 8`; net/vgg.py:22-58 — the reference raises TypeError when this backbone goes
    through RP_Net; here it runs under the stated generalisation and trains).  The stack has no normalisation layers, so
    its train-mode forward is the eval forward (conv + bias + ReLU epilogues, split precision, MaxPool2d(3, stride, 1)) with the
    layer inputs / outputs kept; the backward walks the trace: ReLU mask + bias gradient (rpnet_relu_bias_bwd), weight gradient
    and data gradient on the tcgen05 kernels, max-pool routing through the recorded argmax positions."""
    bucket_groups = [('cre.',), ('encoder.',)]

    def __init__(self, net):
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('rpnet_b200 trains on CUDA (sm_100a) only; there is no CPU fallback')
        self.net, self.dev, self.norm = net, dev, 'batch'
        self.encoder = net.encoder
        self.flat = FlatParams(net)
        self.split = net.encoder.split
        c, L = net.cre, {}
        k = (2 * c.radius + 1) ** 2
        self.kcorr, self.corr_c = k, c.corr_channels
        L['wk'] = _ConvBN('wk', c.w_k[0], c.w_k[1], self.flat)
        L['wq'] = _ConvBN('wq', c.w_q[0], c.w_q[1], self.flat)
        L['q'] = _ConvBN('q', c.q[0], c.q[1], self.flat, hole=(k, self.corr_c - k))
        self.L = L
        self.ws = engine.Workspace()
        self._scratch = {}
        self.act = {}
        self.saved = None
        # bf16 [taps, cin, cout] data-gradient packs of the 12 tensor-core convs (the forward packs are the module's own)
        self.wd = {}
        for m in net.encoder.modules():
            if isinstance(m, nn.Conv2d) and m.in_channels >= 64:
                self.wd[m] = torch.empty(9, m.in_channels, m.out_channels, dtype=bf16, device=dev)
        self._wd_table = ops.pack_conv_weights([(m.weight.data, None, wd, (0, 0), False) for m, wd in self.wd.items()])

    def pack_weights(self):
        super().pack_weights()                                  # the cre layers
        ops.run_pack_conv_weights(*self._wd_table)

    def _encoder_fwd(self, imgs, gs, mask=None):
        if imgs.shape[1] == 1:
            imgs = imgs.expand(-1, 3, -1, -1).contiguous()       # net/rp_net.py:246-247
        self.trace = []
        self.imgs3 = imgs
        out = self.encoder.encode_nhwc(imgs, 'train', trace=self.trace)
        return (engine.hi_of(out), engine.lo_of(out))

    def _encoder_bwd(self, g_d4, buckets=None):
        g = g_d4                                                # bf16, gradient w.r.t. the output of the current layer
        for rec in reversed(self.trace):
            if rec[0] == 'pool':
                _, (k, s, p), x, y, idx = rec
                dx = self.buf('vgg.dpool.%d' % id(idx), tuple(engine.hi_of(x).shape), bf16)
                ops.maxpool_bwd(g, idx, k, s, p, dx)
                g = dx
                continue
            kind, conv, relu, x, y = rec
            y_hi = engine.hi_of(y)
            gz = self.scratch('vgg.gz', y_hi.numel(), bf16).view(y_hi.shape)
            ops.relu_bias_bwd(g.contiguous(), y_hi if relu else None, gz, self.flat.grad_of(conv.bias))
            gw = self.flat.grad_of(conv.weight)
            if kind == 'first':
                ops.conv3x3_first_wgrad(self.imgs3.float().contiguous(), gz, gw)
                break
            x_hi = engine.hi_of(x)
            n, h, w, cin = x_hi.shape
            d = conv.dilation[0]
            taps = [((ky - 1) * d, (kx - 1) * d) for ky in range(3) for kx in range(3)]
            nb = ops.conv_wgrad_workspace_bytes(cin, 0, n, h, w, 9, conv.out_channels)
            ops.conv_wgrad(x_hi, gz, taps, gw, self.scratch('wgrad', nb // 4, f32), accumulate=False)
            dx = self.buf('vgg.dx.%d' % id(conv), (n, h, w, cin), bf16)
            ops.conv_dgrad(gz, self.wd[conv], taps, dx)
            g = dx
        if buckets:
            buckets.ready(1)


class ResNetTrainEngine(TrainEngine):
    """TrainEngine for `backbone: resnet` (net/rp_net.py:19-42: torchvision resnet18 stem + layer1, then three stride-1 stages of
    BasicBlocks, 512 channels at H/4; all BatchNorm2d in train mode: batch statistics per encoder call, running statistics updated).
    Forward: stem conv (CUDA cores) -> statistics -> apply (+ReLU) -> MaxPool2d(3, 2, 1) with recorded argmax positions; every
    BasicBlock conv on the tcgen05 kernel with its statistics in the epilogue; the block tail y = relu(bn2(z2) + identity) is one
    apply pass with the identity planes as a third input.  Backward per block: ReLU mask of y on the incoming gradient
    (rpnet_add_relu_mask_bf16), BatchNorm backward + weight / data gradients of conv2, conv1 (and the 1x1 downsample branch), sum of
    the two branches' input gradients; then max-pool routing, the stem's BatchNorm backward and its weight gradient
    (rpnet_conv7x7s2_stem_wgrad)."""
    bucket_groups = [('cre.',), ('encoder.',)]

    def __init__(self, net):
        dev = next(net.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('rpnet_b200 trains on CUDA (sm_100a) only; there is no CPU fallback')
        self.net, self.dev, self.norm = net, dev, 'batch'
        self.encoder = enc = net.encoder
        self.flat = flat = FlatParams(net)
        self.split = sp = bool(enc.split)
        lv = int(enc.w_level) if sp else engine.W_FP16
        self.lo_level = engine.W_C8 if lv == engine.W_C8 else engine.W_SPLIT
        c, L = net.cre, {}
        k = (2 * c.radius + 1) ** 2
        self.kcorr, self.corr_c = k, c.corr_channels
        L['wk'] = _ConvBN('wk', c.w_k[0], c.w_k[1], flat)
        L['wq'] = _ConvBN('wq', c.w_q[0], c.w_q[1], flat)
        L['q'] = _ConvBN('q', c.q[0], c.q[1], flat, hole=(k, self.corr_c - k))
        bb = enc.backbone
        self.stem_conv, self.stem_bn = bb[0], bb[1]
        self.blocks = []
        for si in (4, 5, 6, 7):
            for bi, blk in enumerate(bb[si]):
                nm = 'r%d%d' % (si, bi)
                L[nm + 'a'] = _ConvBN(nm + 'a', blk.conv1, blk.bn1, flat, split=sp, w_split=lv)
                L[nm + 'b'] = _ConvBN(nm + 'b', blk.conv2, blk.bn2, flat, split=sp, w_split=lv, relu=True, bwd_relu=False)
                if blk.downsample is not None:
                    L[nm + 'd'] = _ConvBN(nm + 'd', blk.downsample[0], blk.downsample[1], flat, split=sp, w_split=lv, relu=False)
                self.blocks.append((nm, blk.downsample is not None))
        self.L = L
        self.ws = engine.Workspace()
        self._scratch = {}
        self.act = {}
        self.saved = None
        self.ones64 = torch.ones(64, dtype=f32, device=dev)
        self.zeros64 = torch.zeros(64, dtype=f32, device=dev)

    def _encoder_fwd(self, imgs, gs, mask=None):
        if imgs.shape[1] == 1:
            imgs = imgs.expand(-1, 3, -1, -1)                     # net/rp_net.py:246-247
        imgs = imgs.float().contiguous()
        self.imgs3, self.gs = imgs, gs
        n, _, H, W = imgs.shape
        G, sp = len(gs) - 1, self.split
        h2, w2 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        h4, w4 = (h2 + 2 - 3) // 2 + 1, (w2 + 2 - 3) // 2 + 1
        sums = lambda c: self.scratch('bn_sums', G * c * 2, torch.float64)
        bn = self.stem_bn
        z = self.pair('stem.z', (n, h2, w2, 64), sp, z=True)
        ops.conv7x7s2_stem(imgs, self.stem_conv.weight.data, self.ones64, self.zeros64, z[0], relu=False, out_lo=z[1])
        st = self.buf('stem.stats', (G, 64, 4), f32)
        ops.bn_stats(z[0], gs, sums(64), z_lo=z[1])
        ops.bn_finalize(sums(64), gs, 64, h2 * w2, bn.weight.data, bn.bias.data, None, bn.running_mean, bn.running_var,
                        bn.num_batches_tracked, st, eps=bn.eps, momentum=bn.momentum)
        y = self.pair('stem.y', (n, h2, w2, 64), sp)
        ops.bn_apply(z[0], st, gs, True, y=y[0], z_lo=z[1], y_lo=y[1])
        cur = self.pair('stem.p', (n, h4, w4, 64), sp)
        idx = self.buf('stem.idx', (n, h4, w4, 64), torch.uint8)
        ops.maxpool(y[0], 3, 2, 1, cur[0], x_lo=y[1], out_lo=cur[1], idx=idx)          # nn.MaxPool2d(3, 2, 1)
        self.act = {'stem': dict(z=z[0], stats=st, idx=idx, y_shape=tuple(y[0].shape))}
        for nm, down in self.blocks:
            la, lb = self.L[nm + 'a'], self.L[nm + 'b']
            shp = lambda c: (n, h4, w4, c)
            za, sa = self.pair(nm + 'a.z', shp(la.cout), sp, z=True), self.buf(nm + 'a.stats', (G, la.cout, 4), f32)
            a = self.pair(nm + 'a.y', shp(la.cout), sp)
            la.fwd(cur, None, za, gs, sums(la.cout), sa, y=a)
            rec = dict(x=cur[0], za=za[0], sa=sa, a=a[0])
            idn = cur
            if down:
                ld = self.L[nm + 'd']
                zd, sd = self.pair(nm + 'd.z', shp(ld.cout), sp, z=True), self.buf(nm + 'd.stats', (G, ld.cout, 4), f32)
                idn = self.pair(nm + 'd.y', shp(ld.cout), sp)
                ld.fwd(cur, None, zd, gs, sums(ld.cout), sd, y=idn)
                rec.update(zd=zd[0], sd=sd)
            zb, sb = self.pair(nm + 'b.z', shp(lb.cout), sp, z=True), self.buf(nm + 'b.stats', (G, lb.cout, 4), f32)
            out = self.pair(nm + '.y', shp(lb.cout), sp)
            lb.fwd(a, None, zb, gs, sums(lb.cout), sb, y=out, res=idn)             # relu(bn2(conv2(a)) + identity)
            rec.update(zb=zb[0], sb=sb, y=out[0])
            self.act[nm] = rec
            cur = out
        return cur

    def _encoder_bwd(self, g_d4, buckets=None):
        gs, g = self.gs, g_d4                                     # g: bf16 gradient w.r.t. the output of the current block
        for nm, down in reversed(self.blocks):
            A = self.act[nm]
            la, lb = self.L[nm + 'a'], self.L[nm + 'b']
            g1 = self.buf(nm + '.g1', tuple(g.shape), bf16)
            ops.add_relu_mask(g.contiguous(), g1, y=A['y'])       # through the block's final ReLU: feeds bn2 and the identity branch
            da = self.buf(nm + '.da', tuple(A['a'].shape), bf16)
            lb.bwd(self, A['a'], None, A['zb'], A['sb'], gs, dx=da, direct=g1)
            dx = self.buf(nm + '.dx', tuple(A['x'].shape), bf16)
            la.bwd(self, A['x'], None, A['za'], A['sa'], gs, dx=dx, direct=da)
            other = g1
            if down:
                other = self.buf(nm + '.dxd', tuple(A['x'].shape), bf16)
                self.L[nm + 'd'].bwd(self, A['x'], None, A['zd'], A['sd'], gs, dx=other, direct=g1)
            g = self.buf(nm + '.gin', tuple(A['x'].shape), bf16)
            ops.add_relu_mask(dx, g, b=other)                     # the previous block's ReLU is applied when that block is visited
        S = self.act['stem']
        dy = self.buf('stem.dy', S['y_shape'], bf16)
        ops.maxpool_bwd(g, S['idx'], 3, 2, 1, dy)
        bn = self.stem_bn
        dz = self.scratch('dz', dy.numel(), bf16).view(dy.shape)
        ops.bn_bwd(S['z'], S['stats'], gs, dz, self.scratch('bn_bwd', (len(gs) - 1) * 64 * 6, f32), True,
                   dgamma=self.flat.grad_of(bn.weight), dbeta=self.flat.grad_of(bn.bias), direct=dy)
        ops.conv7x7s2_stem_wgrad(self.imgs3, dz, self.flat.grad_of(self.stem_conv.weight))
        if buckets:
            buckets.ready(1)


class TrainStep:
    """One optimisation step: pack weights -> forward -> dice_ce (+grad) -> backward -> all-reduce -> Adam."""

    def __init__(self, net, world_size=1, lr=1e-5, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, align_loss_scaler=1.0,
                 process_group=None):
        self.eng = TrainEngine.of(net)
        self.eng.flat.alias_grads()
        self.net, self.world = net, world_size
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.align_scaler = float(align_loss_scaler)
        self.buckets = GradBuckets(self.eng.flat, world_size, process_group, getattr(self.eng, 'bucket_groups', None))
        self.t = 0
        self.last = {}
        if world_size > 1:
            self.sync_from_rank0(process_group)

    def sync_from_rank0(self, group=None):
        """Rank 0's parameters, Adam moments and BatchNorm buffers on every rank (what torch DDP does at construction).
        The BatchNorm running statistics then evolve rank-locally, like the reference without SyncBN; call this again (or
        average them) before writing a checkpoint that should not depend on the rank that saves it."""
        import torch.distributed as dist
        f = self.eng.flat
        for t in (f.param, f.exp_avg, f.exp_avg_sq):
            dist.broadcast(t, src=0, group=group)
        for b in self.net.buffers():
            dist.broadcast(b, src=0, group=group)
        engine.WEIGHTS_EPOCH += 1

    def forward_backward(self, d):
        """Loss + gradients (flat buffer / p.grad) without the optimizer: what loss.backward() leaves behind."""
        with torch.cuda.device(self.eng.dev):      # kernels launch on the current device's stream
            return self._forward_backward(d)

    def _forward_backward(self, d):
        eng = self.eng
        eng.flat.alias_grads()
        eng.zero_grad()
        eng.pack_weights()
        logits, align = eng.forward(d)
        T, B, P, H, W = logits.shape
        dlogits = eng.scratch('dlogits', logits.numel(), f32).view(logits.shape)
        losses = eng.buf('losses', (T,), f32)
        ops.dice_ce(logits, d['query_labels'].contiguous(), eng.scratch('dice.sums', T * (2 * P + 1), torch.float64), losses, dlogits, 1.0)
        eng.backward(dlogits, self.align_scaler, self.buckets)
        self.buckets.finish()
        loss = losses.sum() + self.align_scaler * align[0]
        self.last = {'loss': loss, 'dice_ce': losses, 'align_loss': align, 'logits': logits}
        return loss

    def step(self, d):
        loss = self.forward_backward(d)
        self.t += 1
        f = self.eng.flat
        with torch.cuda.device(self.eng.dev):
            ops.adam(f.param, f.grad, f.exp_avg, f.exp_avg_sq, self.t, self.lr, self.betas, self.eps, self.wd, 1.0 / self.world)
        engine.WEIGHTS_EPOCH += 1
        return loss


class _TrainForward(torch.autograd.Function):
    """Autograd bridge for `RP_Net.forward` in train mode: the forward runs TrainEngine.forward, the backward runs the
    hand-scheduled kernel backward and hands the per-parameter gradients to autograd (which accumulates them into
    `p.grad` exactly as it does for the reference's graph).  `TrainStep` bypasses this bridge (no gradient copies)."""

    @staticmethod
    def forward(ctx, eng, d, *params):
        eng.pack_weights()
        logits, align = eng.forward(d)
        ctx.eng = eng
        ctx.token = eng.saved
        return logits.clone(), align.clone()

    @staticmethod
    def backward(ctx, dlogits, dalign):
        eng = ctx.eng
        with torch.cuda.device(eng.dev):
            return _TrainForward._backward(ctx, eng, dlogits, dalign)

    @staticmethod
    def _backward(ctx, eng, dlogits, dalign):
        if eng.saved is not ctx.token:
            raise RuntimeError('rpnet_b200: backward through a train-mode forward that is not the most recent one of this '
                               'model (activations are kept for one forward at a time)')
        eng.flat.unalias_grads()
        eng.zero_grad()
        dl = torch.zeros_like(eng.saved['logits']) if dlogits is None else dlogits.contiguous().float()
        da = 0.0 if dalign is None else float(dalign.reshape(-1)[0].item())
        eng.backward(dl, da)
        return (None, None) + tuple(eng.flat.grad_of(p).clone() for _, p in eng.flat._named)


def train_forward(net, d):
    """out['refinement'] logits [T, B, 1+Wa, H, W] and out['align_loss'] [1] of a train-mode RP_Net.forward, attached to
    the autograd graph when gradients are enabled."""
    eng = TrainEngine.of(net)
    if torch.is_grad_enabled():
        return _TrainForward.apply(eng, d, *[p for _, p in eng.flat._named])
    eng.pack_weights()
    logits, align = eng.forward(d)
    return logits.clone(), align.clone()
