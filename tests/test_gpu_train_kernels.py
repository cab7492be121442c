"""Per-kernel parity of the TRAINING entry points (include/rpnet_b200.h "Training path") against torch autograd on
the CPU (the oracle's arithmetic: the reference trains through autograd over the same ATen ops).  Tensor-core kernels
are fed pre-rounded fp16 / bf16 operands so that the comparison isolates the kernel (fp32 accumulation); tolerances are
stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import _lib
    _lib.load()
    return torch.device('cuda:0')


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _nhwc(x, dtype, dev):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype).to(dev)


def _nchw(y):
    return y.float().permute(0, 3, 1, 2).contiguous().cpu()


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------- wgrad / dgrad
WG_CASES = [
    # n, c0, c1, cout, h, w, k, dil
    (2, 64, 0, 64, 16, 16, 3, 1),        # odd X-box count (9): duplicated half tile
    (1, 128, 0, 256, 32, 32, 3, 1),      # BN = 256
    (3, 64, 64, 128, 8, 8, 3, 1),        # two-source concat, batch folded into pixel boxes
    (2, 256, 0, 64, 24, 40, 3, 1),       # ragged H, W
    (2, 192, 0, 64, 16, 16, 1, 1),       # 1x1
    (5, 64, 0, 64, 4, 4, 3, 1),          # tiny maps, ragged batch box
    (4, 128, 0, 128, 64, 64, 3, 1),      # many pixel tiles -> split-K
]


@pytest.mark.parametrize('case', WG_CASES)
def test_conv_wgrad_vs_autograd(dev, case):
    from rpnet_b200 import ops
    n, c0, c1, cout, h, w, k, dil = case
    g = _gen(hash(case) % 997)
    cin = c0 + c1
    x = torch.randn(n, cin, h, w, generator=g).to(bf16).float()      # exactly representable in fp16 AND bf16
    dz = (torch.randn(n, cout, h, w, generator=g) * 0.1).to(bf16).float()
    wt = torch.zeros(cout, cin, k, k, requires_grad=True)
    F.conv2d(x, wt, None, padding=dil * (k // 2), dilation=dil).backward(dz)
    taps = [((ky - k // 2) * dil, (kx - k // 2) * dil) for ky in range(k) for kx in range(k)]
    x0 = _nhwc(x[:, :c0], torch.float16, dev)
    x1 = _nhwc(x[:, c0:], torch.float16, dev) if c1 else None
    grad = torch.full((cout, cin, k, k), 0.5, device=dev)
    ws = torch.empty(ops.conv_wgrad_workspace_bytes(c0, c1, n, h, w, k * k, cout) // 4, device=dev)
    ops.conv_wgrad(x0, _nhwc(dz, bf16, dev), taps, grad, ws, x1=x1, accumulate=True)
    torch.cuda.synchronize()
    got = grad.cpu() - 0.5
    assert _rel(got, wt.grad) < 2e-5, _rel(got, wt.grad)
    ops.conv_wgrad(x0, _nhwc(dz, bf16, dev), taps, grad, ws, x1=x1, accumulate=False)
    torch.cuda.synchronize()
    assert _rel(grad.cpu(), wt.grad) < 2e-5


def test_conv_wgrad_hole(dev):
    """cre.q: packed input = [corr(121) + 7 zero pad | fm1(256)]; the gradient skips the padding channels."""
    from rpnet_b200 import ops
    g = _gen(3)
    n, h, w, cout = 2, 16, 16, 64
    corr = torch.randn(n, 121, h, w, generator=g).to(bf16).float()
    fm1 = torch.randn(n, 256, h, w, generator=g).to(bf16).float()
    dz = (torch.randn(n, cout, h, w, generator=g) * 0.1).to(bf16).float()
    wt = torch.zeros(cout, 377, 1, 1, requires_grad=True)
    F.conv2d(torch.cat([corr, fm1], 1), wt).backward(dz)
    corr_p = torch.cat([corr, torch.zeros(n, 7, h, w)], 1)
    grad = torch.zeros(cout, 377, 1, 1, device=dev)
    ws = torch.empty(ops.conv_wgrad_workspace_bytes(128, 256, n, h, w, 1, cout) // 4, device=dev)
    ops.conv_wgrad(_nhwc(corr_p, torch.float16, dev), _nhwc(dz, bf16, dev), [(0, 0)], grad, ws, x1=_nhwc(fm1, torch.float16, dev),
                   hole=(121, 7), accumulate=True)
    torch.cuda.synchronize()
    assert _rel(grad.cpu(), wt.grad) < 2e-5


@pytest.mark.parametrize('case', [(2, 64, 128, 16, 16, 3, 1), (1, 256, 64, 24, 40, 3, 1), (2, 384, 64, 16, 16, 1, 1)])
def test_conv_dgrad_and_pack(dev, case):
    """pack_conv_weight (+ padding hole) and the bf16 data-gradient conv against autograd."""
    from rpnet_b200 import ops
    n, cin, cout, h, w, k, dil = case
    g = _gen(hash(case) % 991)
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k))
    x = torch.zeros(n, cin, h, w, requires_grad=True)
    dz = (torch.randn(n, cout, h, w, generator=g)).to(bf16).float()
    F.conv2d(x, wt.to(bf16).float(), None, padding=dil * (k // 2), dilation=dil).backward(dz)
    wf = torch.empty(k * k, cout, cin, dtype=torch.float16, device=dev)
    wd = torch.empty(k * k, cin, cout, dtype=bf16, device=dev)
    ops.pack_conv_weight(wt.to(dev), wf, wd)
    torch.cuda.synchronize()
    assert torch.equal(wf.cpu(), wt.permute(2, 3, 0, 1).reshape(k * k, cout, cin).half())
    assert torch.equal(wd.cpu(), wt.permute(2, 3, 1, 0).reshape(k * k, cin, cout).to(bf16))
    taps = [((ky - k // 2) * dil, (kx - k // 2) * dil) for ky in range(k) for kx in range(k)]
    dx = torch.empty(n, h, w, cin, dtype=bf16, device=dev)
    ops.conv_dgrad(_nhwc(dz, bf16, dev), wd, taps, dx)
    torch.cuda.synchronize()
    assert _rel(_nchw(dx), x.grad) < 4e-3          # bf16 output rounding


def test_pack_conv_weight_hole(dev):
    from rpnet_b200 import ops
    wt = torch.randn(64, 377, 1, 1, generator=_gen(1))
    wf = torch.empty(1, 64, 384, dtype=torch.float16, device=dev)
    wd = torch.empty(1, 384, 64, dtype=bf16, device=dev)
    ops.pack_conv_weight(wt.to(dev), wf, wd, hole=(121, 7))
    torch.cuda.synchronize()
    want = torch.cat([wt[:, :121, 0, 0], torch.zeros(64, 7), wt[:, 121:, 0, 0]], 1)
    assert torch.equal(wf.cpu()[0], want.half()) and torch.equal(wd.cpu()[0], want.t().to(bf16))


def test_first_conv_wgrad(dev):
    from rpnet_b200 import ops
    g = _gen(5)
    n, h, w = 3, 32, 48
    img = torch.randn(n, 1, h, w, generator=g)
    dz = (torch.randn(n, 64, h, w, generator=g) * 0.1).to(bf16).float()
    wt = torch.zeros(64, 1, 3, 3, requires_grad=True)
    F.conv2d(img, wt, None, padding=1).backward(dz)
    grad = torch.zeros(64, 1, 3, 3, device=dev)
    ops.conv3x3_first_wgrad(img.to(dev), _nhwc(dz, bf16, dev), grad)
    torch.cuda.synchronize()
    assert _rel(grad.cpu(), wt.grad) < 1e-5


def test_stem_wgrad_vs_autograd(dev):
    """rpnet_conv7x7s2_stem_wgrad: weight gradient of torchvision resnet18.conv1 (7x7, stride 2, padding 3; net/rp_net.py:19-23),
    ragged tiles included; accumulates (+=) and is deterministic."""
    from rpnet_b200 import ops
    g = _gen(15)
    n, h, w = 3, 52, 76
    img = torch.randn(n, 3, h, w, generator=g)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    dz = (torch.randn(n, 64, ho, wo, generator=g) * 0.1).to(bf16).float()
    wt = torch.zeros(64, 3, 7, 7, requires_grad=True)
    F.conv2d(img, wt, None, stride=2, padding=3).backward(dz)
    grad = torch.ones(64, 3, 7, 7, device=dev)
    ops.conv7x7s2_stem_wgrad(img.to(dev), _nhwc(dz, bf16, dev), grad)
    g2 = torch.ones(64, 3, 7, 7, device=dev)
    ops.conv7x7s2_stem_wgrad(img.to(dev), _nhwc(dz, bf16, dev), g2)
    torch.cuda.synchronize()
    assert _rel(grad.cpu() - 1.0, wt.grad) < 1e-5
    assert torch.equal(grad, g2)


def test_add_relu_mask(dev):
    from rpnet_b200 import ops
    g = _gen(16)
    a = torch.randn(2, 8, 8, 64, generator=g).to(bf16)
    b = torch.randn(2, 8, 8, 64, generator=g).to(bf16)
    y = torch.randn(2, 8, 8, 64, generator=g).half()
    out = torch.empty(2, 8, 8, 64, dtype=bf16, device=dev)
    ops.add_relu_mask(a.to(dev), out, b=b.to(dev), y=y.to(dev))
    torch.cuda.synchronize()
    want = ((a.float() + b.float()) * (y.float() > 0)).to(bf16)
    assert torch.equal(out.cpu(), want)
    ops.add_relu_mask(a.to(dev), out, y=y.to(dev))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), (a.float() * (y.float() > 0)).to(bf16))
    ops.add_relu_mask(a.to(dev), out, b=b.to(dev))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), (a.float() + b.float()).to(bf16))


@pytest.mark.parametrize('level', [1, 2])
def test_bn_apply_with_residual(dev, level):
    """rpnet_bn_apply_res_f16: y = relu(bn(z) + identity), identity as hi + lo planes (BasicBlock tail in train mode)."""
    from rpnet_b200 import engine, ops
    g = _gen(17)
    n, c, h, w, gs = 4, 128, 8, 8, [0, 3, 4]
    z = torch.randn(n, c, h, w, generator=g) * 1.5 + 0.3
    res = torch.randn(n, c, h, w, generator=g)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    want = torch.cat([F.relu(F.batch_norm(z[a:b].double(), None, None, gamma.double(), beta.double(), True, 0.0, 1e-5) + res[a:b].double())
                      for a, b in ((0, 3), (3, 4))]).float()
    zh, zl = (t.to(dev) for t in engine.split_f16(z.permute(0, 2, 3, 1).contiguous()))
    rh, rl = (t.to(dev) for t in engine.split_planes(res.permute(0, 2, 3, 1).contiguous(), level))
    G = len(gs) - 1
    sums = torch.empty(G, c, 2, device=dev, dtype=torch.float64); stats = torch.empty(G, c, 4, device=dev)
    rm, rv, nbt = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    ops.bn_stats(zh, gs, sums, z_lo=zl)
    ops.bn_finalize(sums, gs, c, h * w, gamma.to(dev), beta.to(dev), None, rm, rv, nbt, stats)
    y = torch.empty(n, h, w, c, dtype=torch.float16, device=dev)
    y_lo = torch.empty(n, h, w, 2 * c, dtype=torch.uint8, device=dev) if level == 2 else torch.empty_like(y)
    ops.bn_apply(zh, stats, gs, True, y=y, z_lo=zl, y_lo=y_lo, res=rh, res_lo=rl)
    torch.cuda.synchronize()
    got = engine.join_planes(y, y_lo).permute(0, 3, 1, 2).cpu()
    assert (got - want).abs().max().item() / want.abs().max().item() < (1e-5 if level == 1 else 1e-4)


# ------------------------------------------------------------------------------------------- batch norm
@pytest.mark.parametrize('case', [(6, 64, 16, 16, [0, 4, 6], True), (4, 256, 8, 8, [0, 4], False), (3, 1024, 4, 4, [0, 1, 2, 3], False)])
def test_bn_train_forward(dev, case):
    """bn_stats + bn_finalize + bn_apply (+ fused 2x2 max-pool) == nn.BatchNorm2d(train) + ReLU + MaxPool2d per call group,
    incl. the running-statistics updates applied once per group in order."""
    from rpnet_b200 import ops
    n, c, h, w, gs, pool = case
    g = _gen(n * c)
    z = (torch.randn(n, c, h, w, generator=g) * 1.5 + 0.3).half().float()
    gamma, beta, bias = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1, torch.randn(c, generator=g) * 0.1
    rm, rv = torch.randn(c, generator=g) * 0.1, torch.rand(c, generator=g) + 0.5
    bn = torch.nn.BatchNorm2d(c)
    bn.weight.data.copy_(gamma); bn.bias.data.copy_(beta); bn.running_mean.copy_(rm); bn.running_var.copy_(rv)
    bn.train()
    ys = [F.relu(bn(z[gs[i]:gs[i + 1]] + bias[None, :, None, None])) for i in range(len(gs) - 1)]
    want = torch.cat(ys).detach()
    zd = _nhwc(z, torch.float16, dev)
    G = len(gs) - 1
    sums = torch.empty(G, c, 2, device=dev, dtype=torch.float64); stats = torch.empty(G, c, 4, device=dev)
    rmd, rvd, nbt = rm.to(dev), rv.to(dev), torch.zeros((), dtype=torch.int64, device=dev)
    ops.bn_stats(zd, gs, sums)
    ops.bn_finalize(sums, gs, c, h * w, gamma.to(dev), beta.to(dev), bias.to(dev), rmd, rvd, nbt, stats)
    y = torch.empty(n, h, w, c, dtype=torch.float16, device=dev)
    y32 = torch.empty(n, h, w, c, dtype=torch.float32, device=dev)
    yp = torch.empty(n, h // 2, w // 2, c, dtype=torch.float16, device=dev) if pool else None
    ops.bn_apply(zd, stats, gs, True, y=y, y_pool=yp, y_f32=y32)
    torch.cuda.synchronize()
    torch.testing.assert_close(_nchw(y32), want, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(_nchw(y), want, rtol=2e-3, atol=2e-3)
    if pool:
        torch.testing.assert_close(_nchw(yp), F.max_pool2d(want, 2, 2), rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(rmd.cpu(), bn.running_mean, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rvd.cpu(), bn.running_var, rtol=1e-4, atol=1e-5)
    assert nbt.item() == G == bn.num_batches_tracked.item()


@pytest.mark.parametrize('mode', ['direct', 'direct_f32', 'pool+direct', 'up', 'pool'])
def test_bn_backward(dev, mode):
    """dz, dgamma, dbeta of BatchNorm(train)+ReLU with the activation gradient arriving directly (channel slice of a wider
    tensor), through the 2x2 max-pool and / or through the nearest x2 upsample."""
    from rpnet_b200 import ops
    g = _gen(len(mode))
    n, c, h, w, gs = 4, 64, 8, 12, [0, 3, 4]
    z = (torch.randn(n, c, h, w, generator=g) * 1.2 + 0.2).half().float().requires_grad_(True)
    gamma = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(c, generator=g) * 0.1).requires_grad_(True)
    ys = []
    for i in range(len(gs) - 1):
        ys.append(F.relu(F.batch_norm(z[gs[i]:gs[i + 1]], None, None, gamma, beta, True, 0.1, 1e-5)))
    y = torch.cat(ys)
    kw = {}
    loss = 0
    if 'direct' in mode:
        wide = (torch.randn(n, 2 * c, h, w, generator=g)).to(bf16).float()
        loss = loss + (y * wide[:, c:]).sum()
        kw.update(direct=_nhwc(wide, torch.float32 if mode == 'direct_f32' else bf16, dev), d_off=c)
    if 'pool' in mode:
        gp = torch.randn(n, c, h // 2, w // 2, generator=g).to(bf16).float()
        loss = loss + (F.max_pool2d(y, 2, 2) * gp).sum()
        kw.update(pooled=_nhwc(gp, bf16, dev))
    if mode == 'up':
        gu = torch.randn(n, c, 2 * h, 2 * w, generator=g).to(bf16).float()
        loss = loss + (F.interpolate(y, scale_factor=2, mode='nearest') * gu).sum()
        kw.update(up=_nhwc(gu, bf16, dev))
    loss.backward()
    zd = _nhwc(z.detach(), torch.float16, dev)
    G = len(gs) - 1
    sums = torch.empty(G, c, 2, device=dev, dtype=torch.float64); stats = torch.empty(G, c, 4, device=dev)
    ops.bn_stats(zd, gs, sums)
    ops.bn_finalize(sums, gs, c, h * w, gamma.detach().to(dev), beta.detach().to(dev), None, None, None, None, stats)
    dz = torch.empty(n, h, w, c, dtype=bf16, device=dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    ops.bn_bwd(zd, stats, gs, dz, torch.empty(G, c, 6, device=dev), True, dgamma=dg, dbeta=db, **kw)
    torch.cuda.synchronize()
    assert _rel(_nchw(dz), z.grad) < 5e-3, _rel(_nchw(dz), z.grad)        # bf16 output rounding
    assert _rel(dg.cpu(), gamma.grad) < 1e-4 and _rel(db.cpu(), beta.grad) < 1e-4


def test_upsample2x_and_premask_bwd(dev):
    from rpnet_b200 import ops
    g = _gen(9)
    x = torch.randn(2, 64, 6, 10, generator=g).half()
    y = torch.empty(2, 12, 20, 64, dtype=torch.float16, device=dev)
    ops.upsample2x(_nhwc(x, torch.float16, dev), y)
    torch.cuda.synchronize()
    assert torch.equal(_nchw(y), F.interpolate(x.float(), scale_factor=2, mode='nearest'))
    iters, n, c, h, w = 3, 2, 64, 8, 8
    dfg = torch.randn(iters, n, h, w, c, generator=g).to(bf16)
    dbg = torch.randn(iters, n, h, w, c, generator=g).to(bf16)
    m = torch.rand(iters, n, h, w, generator=g)
    dx = torch.empty(n, h, w, c, dtype=bf16, device=dev)
    ops.premask_bwd(dfg.to(dev), dbg.to(dev), m.to(dev), dx, iters=iters)
    torch.cuda.synchronize()
    want = (dfg.float() * m[..., None] + dbg.float() * (1 - m[..., None])).sum(0)
    assert _rel(dx.float().cpu(), want) < 4e-3


# ------------------------------------------------------------------------------------------- tail backward
@pytest.mark.parametrize('r,h,w', [(5, 16, 16), (2, 12, 20), (5, 8, 24)])
def test_local_corr_bwd(dev, r, h, w):
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    g = _gen(r * h)
    n, c, k = 2, 64, (2 * r + 1) ** 2
    f1 = torch.randn(n, c, h, w, generator=g).half().float().requires_grad_(True)
    f2 = torch.randn(n, c, h, w, generator=g).half().float().requires_grad_(True)
    ld, add_off = 128 + c, 128
    dq = torch.randn(n, ld, h, w, generator=g).to(bf16).float()
    corr = O.correlation_local(f1, f2, r)
    ((corr * dq[:, :k]).sum() + (f1 * dq[:, add_off:add_off + c]).sum()).backward()
    df1 = torch.empty(n, h, w, c, dtype=bf16, device=dev); df2 = torch.empty_like(df1)
    ops.local_corr_bwd(_nhwc(f1.detach(), torch.float16, dev), _nhwc(f2.detach(), torch.float16, dev), _nhwc(dq, bf16, dev), add_off, r,
                       df1, df2)
    torch.cuda.synchronize()
    assert _rel(_nchw(df1), f1.grad) < 4e-3 and _rel(_nchw(df2), f2.grad) < 4e-3


def test_cos_sim_bwd(dev):
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    g = _gen(11)
    sets, reps, P, h, w = 3, 2, 3, 8, 12
    n = sets * reps
    feat = torch.randn(n, 64, h, w, generator=g).relu()
    feat[0, :, 2, 3] = 0
    feat.requires_grad_(True)
    protos = torch.randn(sets, P, 64, generator=g).requires_grad_(True)
    dpred = torch.randn(n, P, h, w, generator=g)
    pred = torch.stack([torch.stack([O.cal_dist(feat[i:i + 1], protos[i % sets, p][None])[0] for p in range(P)]) for i in range(n)])
    (pred * dpred).sum().backward()
    fd = feat.detach().permute(0, 2, 3, 1).contiguous().to(dev)
    dfeat = torch.full((n, h, w, 64), 1.0, device=dev)
    dprotos = torch.zeros(sets, P, 64, device=dev)
    ops.cos_sim_bwd(fd, protos.detach().to(dev), dpred.to(dev), dfeat, dprotos, accumulate=True)
    torch.cuda.synchronize()
    got = dfeat.cpu().permute(0, 3, 1, 2) - 1.0
    mask = torch.ones(n, 1, h, w, dtype=torch.bool); mask[0, :, 2, 3] = False       # zero vector: eps-clamped branch
    assert _rel(got * mask, feat.grad * mask) < 1e-4
    assert _rel(dprotos.cpu(), protos.grad) < 1e-4


def test_masked_pool_adjoint_fwd_bwd(dev, golden):
    """bilinear_adjoint + weighted_pool == getFeatures (reference golden) and its backward == autograd."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    gz = golden('proto_loss')
    g = _gen(13)
    n, h, w, S = 3, 16, 16, 4
    fts = torch.cat([torch.from_numpy(gz['fts']), torch.randn(n - 1, 64, h, w, generator=g).relu()]).requires_grad_(True)
    m0 = torch.cat([torch.from_numpy(gz['mask']), (torch.rand(n - 1, h * S, w * S, generator=g) > 0.6).float()])
    m1 = 1 - m0
    m1[2] = 0                                         # empty mask -> ~0 prototype
    want = torch.stack([torch.cat([O.get_features(fts[i:i + 1], m0[i:i + 1]), O.get_features(fts[i:i + 1], m1[i:i + 1])]) for i in range(n)])
    dout = torch.randn(n, 2, 64, generator=g)
    (want * dout).sum().backward()
    fd = fts.detach().permute(0, 2, 3, 1).contiguous().to(dev)
    w0 = torch.empty(n, h, w, device=dev); w1 = torch.empty(n, h, w, device=dev)
    s0 = torch.empty(n, device=dev); s1 = torch.empty(n, device=dev)
    ops.bilinear_adjoint(m0.to(dev), w0, s0)
    ops.bilinear_adjoint(m1.to(dev), w1, s1)
    out = torch.empty(n, 2, 64, device=dev)
    ops.weighted_pool(fd, w0, w1, s0, s1, out)
    dfeat = torch.empty(n, h, w, 64, device=dev)
    ops.weighted_pool_bwd(dout.to(dev), w0, w1, s0, s1, dfeat)
    torch.cuda.synchronize()
    torch.testing.assert_close(s0.cpu(), m0.sum((1, 2)))
    torch.testing.assert_close(out.cpu()[0, 0], torch.from_numpy(gz['proto'])[0], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out.cpu(), want.detach(), rtol=1e-4, atol=1e-6)
    assert _rel(dfeat.cpu().permute(0, 3, 1, 2), fts.grad) < 1e-5


def test_upsample_backward_is_the_adjoint(dev):
    from rpnet_b200 import ops
    g = _gen(17)
    x = torch.randn(4, 1, 8, 12, generator=g, requires_grad=True)
    dl = torch.randn(4, 1, 32, 48, generator=g)
    (F.interpolate(x, size=(32, 48), mode='bilinear') * dl).sum().backward()
    out = torch.empty(4, 8, 12, device=dev)
    ops.bilinear_adjoint(dl[:, 0].contiguous().to(dev), out)
    up = torch.empty(4, 32, 48, device=dev)
    ops.bilinear_up(x.detach()[:, 0].contiguous().to(dev), up)
    torch.cuda.synchronize()
    assert _rel(out.cpu(), x.grad[:, 0]) < 1e-5
    torch.testing.assert_close(up.cpu(), F.interpolate(x.detach(), size=(32, 48), mode='bilinear')[:, 0], rtol=1e-5, atol=1e-6)


def test_proto_finalize_bwd(dev):
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    g = _gen(19)
    Wa, Sh, B, C = 3, 2, 4, 64
    raw = torch.randn(Wa, Sh, B, 2, C, generator=g, requires_grad=True)
    protos = []
    for b in range(B):
        fg, bg = O.get_prototype([[raw[w, s, b, 0][None] for s in range(Sh)] for w in range(Wa)],
                                 [[raw[w, s, b, 1][None] for s in range(Sh)] for w in range(Wa)])
        protos.append(torch.cat([bg] + fg))
    protos = torch.stack(protos)
    dp = torch.randn(B, 1 + Wa, C, generator=g)
    (protos * dp).sum().backward()
    draw = torch.empty(Wa, Sh, B, 2, C, device=dev)
    ops.proto_finalize_bwd(dp.to(dev), draw)
    torch.cuda.synchronize()
    torch.testing.assert_close(draw.cpu(), raw.grad, rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------- losses
@pytest.mark.parametrize('P', [2, 5])
def test_dice_ce(dev, golden, P):
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    gz = golden('proto_loss')
    lg = torch.from_numpy(gz['logits' if P == 2 else 'logits5'])
    lab = torch.from_numpy(gz['labels' if P == 2 else 'labels5'])
    g = _gen(P)
    logits = torch.stack([lg, lg * 0.5 + torch.randn(lg.shape, generator=g)]).requires_grad_(True)      # G = 2 tensors, same labels
    losses = torch.stack([O.dice_ce(logits[i], lab) for i in range(2)])
    (losses[0] * 1.0 + losses[1] * 1.0).backward()
    ld = logits.detach().to(dev)
    sums = torch.empty(2, 2 * P + 1, device=dev, dtype=torch.float64); loss = torch.empty(2, device=dev); dl = torch.empty_like(ld)
    ops.dice_ce(ld, lab.to(dev), sums, loss, dl)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss.cpu()[0], torch.from_numpy(gz['dice_ce' if P == 2 else 'dice_ce5']), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(loss.cpu(), losses.detach(), rtol=1e-5, atol=1e-6)
    assert _rel(dl.cpu(), logits.grad) < 1e-4


def test_align_loss_pipeline(dev, golden):
    """class_pool -> align_gather -> cos_sim -> bilinear_up -> ce_mask (+ the whole backward chain) == alignLoss
    (net/rp_net.py:394-440) on the reference's golden case (Wa=2, Sh=2) and its autograd gradients."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    gz = golden('proto_loss')
    a_q = torch.from_numpy(gz['a_q']).requires_grad_(True)          # 1 x 64 x 16 x 16
    a_pred = torch.from_numpy(gz['a_pred'])                          # 1 x 3 x 16 x 16
    a_s = torch.from_numpy(gz['a_s']).requires_grad_(True)          # Wa x Sh x 64 x 16 x 16
    a_f, a_b = torch.from_numpy(gz['a_f']), torch.from_numpy(gz['a_b'])
    ref = O.align_loss(a_q, a_pred, a_s, a_f, a_b)
    torch.testing.assert_close(ref.detach(), torch.from_numpy(gz['align']), rtol=1e-5, atol=1e-6)
    ref.backward()
    Wa, Sh, B, h, w, S = 2, 2, 1, 16, 16, 4
    P, n = 1 + Wa, Wa * Sh * B
    qf = a_q.detach().permute(0, 2, 3, 1).contiguous().to(dev)
    sf = a_s.detach().reshape(n, 64, h, w).permute(0, 2, 3, 1).contiguous().to(dev)
    fore, back = a_f.reshape(n, h * S, w * S).to(dev), a_b.reshape(n, h * S, w * S).contiguous().to(dev)
    qproto = torch.empty(B, P, 64, device=dev); counts = torch.empty(B, P, device=dev)
    amax = torch.empty(B, h, w, dtype=torch.int32, device=dev)
    ops.class_pool(qf, a_pred.to(dev), qproto, counts, amax)
    ps = torch.empty(n, 2, 64, device=dev); wgt = torch.empty(n, device=dev)
    ops.align_gather(qproto, counts, Wa, Sh, 1.0, ps, wgt)
    pred_s = torch.empty(n, 2, h, w, device=dev)
    ops.cos_sim(sf, ps, pred_s, 20.0)
    lg = torch.empty(n, 2, h * S, w * S, device=dev)
    ops.bilinear_up(pred_s.view(n * 2, h, w), lg.view(n * 2, h * S, w * S))
    sums = torch.empty(n, 2, device=dev, dtype=torch.float64); loss = torch.empty(1, device=dev); dlg = torch.empty_like(lg)
    ops.ce_mask(lg, fore, back, wgt, sums, loss, dlg)
    dpred_s = torch.empty(n, 2, h, w, device=dev)
    ops.bilinear_adjoint(dlg.view(n * 2, h * S, w * S), dpred_s.view(n * 2, h, w))
    dsf = torch.empty_like(sf); dps = torch.zeros(n, 2, 64, device=dev)
    ops.cos_sim_bwd(sf, ps, dpred_s, dsf, dps)
    dqp = torch.empty(B, P, 64, device=dev)
    ops.align_scatter(dps, Wa, Sh, dqp)
    dqf = torch.zeros_like(qf)
    ops.class_pool_bwd(dqp, counts, amax, dqf)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss.cpu()[0], ref.detach(), rtol=1e-5, atol=1e-6)
    assert _rel(dsf.cpu().permute(0, 3, 1, 2).reshape(a_s.shape), a_s.grad) < 1e-4
    assert _rel(dqf.cpu().permute(0, 3, 1, 2), a_q.grad) < 1e-4


def test_adam_matches_torch(dev):
    from rpnet_b200 import ops
    g = _gen(23)
    p = torch.randn(1000, generator=g)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, weight_decay=1e-4)
    pd, m, v = p.to(dev), torch.zeros(1000, device=dev), torch.zeros(1000, device=dev)
    for step in range(1, 4):
        gr = torch.randn(1000, generator=g)
        ref.grad = gr.clone() * 0.5
        opt.step()
        ops.adam(pd, gr.to(dev), m, v, step, 1e-3, weight_decay=1e-4, grad_scale=0.5)
    torch.cuda.synchronize()
    torch.testing.assert_close(pd.cpu(), ref.detach(), rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('case', [
    # n, c0, c1, cout, h, w, k, group_start
    (4, 64, 0, 64, 16, 32, 3, [0, 3, 4]),          # one image per tile: statistics fused into the conv epilogue
    (3, 128, 64, 256, 24, 40, 3, [0, 1, 3]),       # ragged tiles (out-of-range pixels must not enter the statistics), BN=256
    (6, 64, 0, 128, 8, 8, 1, [0, 2, 6]),           # two images per tile, boundaries aligned: still fused
    (6, 64, 0, 64, 4, 4, 3, [0, 3, 6]),            # eight images per tile, boundary inside a tile: separate statistics pass
    (2, 512, 0, 1024, 16, 16, 3, [0, 1, 2]),       # cout = 1024 (largest per-CTA accumulator)
])
def test_conv_bnstats(dev, case):
    """rpnet_conv_bnstats_f16: z == conv (fp16-rounded) and sums == per-call-group {sum z, sum z^2} of the fp32 conv output."""
    from rpnet_b200 import ops
    n, c0, c1, cout, h, w, k, gs = case
    g = _gen(sum(case[:7]))
    cin = c0 + c1
    x = torch.randn(n, cin, h, w, generator=g).half().float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).half().float()
    want = F.conv2d(x, wt, None, padding=k // 2)
    taps = [(ky - k // 2, kx - k // 2) for ky in range(k) for kx in range(k)]
    wp = wt.permute(2, 3, 0, 1).reshape(k * k, cout, cin).half().contiguous().to(dev)
    z = torch.empty(n, h, w, cout, dtype=torch.float16, device=dev)
    G = len(gs) - 1
    sums = torch.full((G, cout, 2), 7.0, device=dev, dtype=torch.float64)
    ops.conv_bnstats(_nhwc(x[:, :c0], torch.float16, dev), wp, taps, torch.ones(cout, device=dev), torch.zeros(cout, device=dev), z, gs, sums,
                     src1=_nhwc(x[:, c0:], torch.float16, dev) if c1 else None)
    torch.cuda.synchronize()
    torch.testing.assert_close(_nchw(z), want, rtol=2e-3, atol=2e-3)
    for i in range(G):
        blk = want[gs[i]:gs[i + 1]]
        s1, s2 = blk.sum((0, 2, 3)), (blk * blk).sum((0, 2, 3))
        torch.testing.assert_close(sums[i, :, 0].float().cpu(), s1, rtol=2e-3, atol=2e-3 * blk[0, 0].numel() ** 0.5)
        torch.testing.assert_close(sums[i, :, 1].float().cpu(), s2, rtol=2e-3, atol=1e-3)


@pytest.mark.parametrize('case', [(2, 256, 64, 64, 5), (1, 64, 26, 18, 5), (2, 128, 37, 29, 5), (1, 64, 40, 24, 3)])
def test_local_corr_bwd_tensor_core_path(dev, case):
    """With a workspace and maps of at least one halo window the backward runs as two tcgen05 band GEMMs
    (local_corr_tc.cu); same contract as the CUDA-core kernels.  Gradients span several orders of magnitude on purpose (the
    band operand is fp16 with a per-tile power-of-two scale)."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    n, c, h, w, r = case
    g = _gen(sum(case))
    k = (2 * r + 1) ** 2
    f1 = torch.randn(n, c, h, w, generator=g).half().float().requires_grad_(True)
    f2 = torch.randn(n, c, h, w, generator=g).half().float().requires_grad_(True)
    ld, add_off = 128 + c, 128
    dq = torch.randn(n, ld, h, w, generator=g)
    dq[:, :k] *= 1e-6 * torch.exp(3 * torch.randn(n, 1, h, w, generator=g))        # tiny, pixel-dependent magnitudes
    dq = dq.to(bf16).float()
    corr = O.correlation_local(f1, f2, r)
    ((corr * dq[:, :k]).sum() + (f1 * dq[:, add_off:add_off + c]).sum()).backward()
    df1 = torch.empty(n, h, w, c, dtype=bf16, device=dev); df2 = torch.empty_like(df1)
    ws = torch.empty(ops.local_corr_bwd_workspace_bytes(n, h, w, r), dtype=torch.uint8, device=dev)
    ops.local_corr_bwd(_nhwc(f1.detach(), torch.float16, dev), _nhwc(f2.detach(), torch.float16, dev), _nhwc(dq, bf16, dev), add_off, r,
                       df1, df2, workspace=ws)
    torch.cuda.synchronize()
    assert _rel(_nchw(df1), f1.grad) < 4e-3, _rel(_nchw(df1), f1.grad)
    assert _rel(_nchw(df2), f2.grad) < 6e-3, _rel(_nchw(df2), f2.grad)


@pytest.mark.parametrize('case', [(3, 128, 64, 16, 16, [0, 2, 3]), (2, 64, 128, 16, 32, [0, 2]), (1, 256, 256, 32, 32, [0, 1])])
def test_upconv_subpixel_train_path(dev, case):
    """up_conv (nn.Upsample(x2) + 3x3 conv, net/modules.py:61-75) in sub-pixel form: forward z + BatchNorm statistics from
    four phase convs on the low-resolution input, data gradient as a 4x4 stride-2 conv of dZ, weight gradient from four
    4-tap GEMMs + combine — all against torch autograd of the materialised form."""
    from rpnet_b200 import ops
    n, cin, cout, h, w, gs = case
    g = _gen(sum(case[:5]))
    x = torch.randn(n, cin, h, w, generator=g).to(bf16).float().requires_grad_(True)       # exact in fp16 and bf16
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).requires_grad_(True)
    dz = (torch.randn(n, cout, 2 * h, 2 * w, generator=g) * 0.1).to(bf16).float()
    z = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), wt, None, padding=1)
    z.backward(dz)
    wf = torch.empty(4, 4, cout, cin, dtype=torch.float16, device=dev)
    w16 = torch.empty(16, cin, cout, dtype=bf16, device=dev)
    ops.pack_upconv_weight(wt.detach().to(dev), wf, w16)
    zd = torch.full((n, 2 * h, 2 * w, cout), 9.0, dtype=torch.float16, device=dev)
    G = len(gs) - 1
    sums = torch.full((G, cout, 2), 5.0, dtype=torch.float64, device=dev)
    assert ops.upconv_fusable(h, w)
    ops.upconv_fwd_bnstats(_nhwc(x.detach(), torch.float16, dev), wf, torch.ones(cout, device=dev), torch.zeros(cout, device=dev), zd, gs, sums)
    torch.cuda.synchronize()
    # forward: phase weights are sums of fp32 taps rounded to fp16 once (the dense form rounds every tap): 2e-3 of the scale
    assert _rel(_nchw(zd), z.detach()) < 3e-3
    for i in range(G):
        blk = z.detach()[gs[i]:gs[i + 1]]
        torch.testing.assert_close(sums[i, :, 0].float().cpu(), blk.sum((0, 2, 3)), rtol=5e-3, atol=5e-3 * blk[0, 0].numel() ** 0.5)
        torch.testing.assert_close(sums[i, :, 1].float().cpu(), (blk * blk).sum((0, 2, 3)), rtol=5e-3, atol=1e-2)
    dzd = _nhwc(dz, bf16, dev)
    dx = torch.empty(n, h, w, cin, dtype=bf16, device=dev)
    ops.upconv_dgrad(dzd, w16, dx)
    grad = torch.full((cout, cin, 3, 3), 0.25, device=dev)
    ws = torch.empty(ops.upconv_wgrad_workspace_bytes(cin, n, h, w, cout) // 4, device=dev)
    ops.upconv_wgrad(_nhwc(x.detach(), torch.float16, dev), dzd, grad, ws, accumulate=True)
    torch.cuda.synchronize()
    assert _rel(_nchw(dx), x.grad) < 6e-3, _rel(_nchw(dx), x.grad)            # bf16 weights (summed taps) + bf16 output
    assert _rel(grad.cpu() - 0.25, wt.grad) < 2e-5, _rel(grad.cpu() - 0.25, wt.grad)
