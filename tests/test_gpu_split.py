"""Split ("fp32-class") convolution paths (include/rpnet_b200.h, rpnet_conv_split_f16 and friends) against torch fp32/fp64
on UNROUNDED operands: the point of the paths is that (hi, lo) planes and split weight packs reproduce the reference's fp32
nn.Conv2d (net/modules.py:47-54) without the 2^-11 operand rounding of a single-term tensor-core product.
Two weight-pack levels (engine.W_SPLIT / engine.W_C8, the `level` parameter of the tests):
  1  split-fp16: hi.Wh + lo.Wh + hi.Wl on fp16 operands.  Tolerance 2e-5 relative to the output scale (fp32 accumulation order
     + the dropped lo.Wl term), ~50x tighter than the single-term kernel reaches on the same inputs (asserted below);
  2  fp8 corrections (the default): hi.Wh + 2^-15 (lo8.Wh8 + x8.Wl8), c8 planes.  The kernel must agree with the same sum
     evaluated in fp64 on the identical rounded operands to fp32 accumulation error (5e-6: this pins the byte layouts, the
     scales and the scale-input-d step), and lands within 1e-4 of the unrounded fp64 conv (the e4m3 rounding of the
     corrections: 2^-15 per product), still an order of magnitude inside the single-term error."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import _lib
    _lib.load()
    return torch.device('cuda:0')


def _gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.fixture(params=[1, 2], ids=['split16', 'c8'])
def level(request):
    return request.param


TOL = {1: 2e-5, 2: 1e-4}           # vs the unrounded fp64 result, relative to the output scale


def _pair(x, dev, level=1):
    """NCHW fp32 -> (hi, lo) NHWC planes on the device (lo: fp16 residual plane, or the c8 plane for level 2)."""
    from rpnet_b200 import engine
    hi, lo = engine.split_planes(x.permute(0, 2, 3, 1).contiguous(), level)
    return hi.to(dev), lo.to(dev)


def _lo_like(hi, level=1):
    """An empty lo plane for the hi plane `hi`."""
    if level == 2:
        return torch.empty(tuple(hi.shape[:-1]) + (2 * hi.shape[-1],), dtype=torch.uint8, device=hi.device)
    return torch.empty_like(hi)


def _join(hi, lo):
    from rpnet_b200 import engine
    return engine.join_planes(hi, lo).permute(0, 3, 1, 2).contiguous().cpu()


def _c8_conv_fp64(x, wt, **kw):
    """The fp8-correction sum hi.Wh + 2^-15 (lo8.Wh8 + x8.Wl8) in fp64 on the operands as the kernel rounds them."""
    e4 = lambda t: t.clamp(-448, 448).to(torch.float8_e4m3fn).double()
    xh, wh = x.half().float(), wt.half().float()
    main = F.conv2d(xh.double(), wh.double(), None, **kw)
    corr = F.conv2d(e4((x - xh) * 2.0 ** 9), e4(wh * 2.0 ** 6), None, **kw) + F.conv2d(e4(x * 0.25), e4((wt - wh) * 2.0 ** 17), None, **kw)
    return main + corr * 2.0 ** -15


CASES = [
    # n, c0, c1, cout, h, w, k
    (2, 64, 0, 64, 32, 32, 3),
    (3, 128, 0, 256, 8, 8, 3),         # CTA-pair tiles, ragged batch tile
    (2, 128, 64, 128, 24, 40, 3),      # channel concat of two split sources, ragged H / W
    (1, 256, 256, 256, 16, 16, 3),     # the Up_conv4 shape class
    (2, 192, 0, 64, 16, 16, 1),        # 1x1
]


@pytest.mark.parametrize('case', [(2, 64, 128, 16, 24, 3), (3, 128, 128, 8, 8, 3), (2, 64, 256, 16, 16, 1)])
@pytest.mark.parametrize('res_split', [True, False])
def test_conv_split_residual_vs_fp64(dev, case, res_split, level):
    """rpnet_conv_split_res_f16: relu(affine(conv(x)) + identity), the BasicBlock tail (net/rp_net.py:24-35), identity as hi + lo
    planes or a plain fp16 tensor."""
    from rpnet_b200 import engine, ops
    n, cin, cout, h, w, k = case
    g = _gen(sum(case) + 7)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(n, cout, h, w, generator=g)
    r = _pair(res, dev, level)
    if not res_split:
        r = (r[0], None)
        res = r[0].float().permute(0, 3, 1, 2).cpu()
    ref = F.relu(F.conv2d(x.double(), wt.double(), None, padding=k // 2) * scale[None, :, None, None].double()
                 + shift[None, :, None, None].double() + res.double()).float()
    wp, taps = engine.pack_weight_taps(wt.to(dev), split=level)
    a = _pair(x, dev, level)
    out = torch.empty(n, h, w, cout, dtype=torch.float16, device=dev)
    out_lo = _lo_like(out, level)
    ops.conv_split(a[0], wp, taps, scale.to(dev), shift.to(dev), True, src0_lo=a[1], w_split=level, out=out, out_lo=out_lo, res=r[0],
                   res_lo=r[1])
    torch.cuda.synchronize()
    err = (_join(out, out_lo) - ref).abs().max().item() / ref.abs().max().item()
    assert err < TOL[level], err


def test_stem_and_maxpool_split_vs_fp64(dev, level):
    """ResNet18 stem in split precision: 7x7/s2 conv + BN + ReLU writing hi + lo planes, then MaxPool2d(3, 2, 1) on the pair."""
    from rpnet_b200 import ops
    g = _gen(91)
    n, H, W = 2, 64, 96
    img = torch.randn(n, 3, H, W, generator=g)
    wt = torch.randn(64, 3, 7, 7, generator=g) / math.sqrt(147)
    scale = torch.rand(64, generator=g) + 0.5
    shift = torch.randn(64, generator=g) * 0.1
    ref = F.relu(F.conv2d(img.double(), wt.double(), None, stride=2, padding=3) * scale[None, :, None, None].double()
                 + shift[None, :, None, None].double())
    ref_pool = F.max_pool2d(ref, 3, 2, 1).float()
    ref = ref.float()
    h2, w2 = H // 2, W // 2
    out = torch.empty(n, h2, w2, 64, dtype=torch.float16, device=dev)
    out_lo = _lo_like(out, level)
    ops.conv7x7s2_stem(img.to(dev), wt.to(dev), scale.to(dev), shift.to(dev), out, out_lo=out_lo)
    pool = torch.empty(n, h2 // 2, w2 // 2, 64, dtype=torch.float16, device=dev)
    pool_lo = _lo_like(pool, level)
    ops.maxpool(out, 3, 2, 1, pool, x_lo=out_lo, out_lo=pool_lo)
    torch.cuda.synchronize()
    s = ref.abs().max().item()
    tol = 2e-6 if level == 1 else 2.0 ** -15          # a c8 plane holds the residual to 2^-4: hi + lo8 * 2^-11 is x to 2^-15
    assert (_join(out, out_lo) - ref).abs().max().item() / s < tol
    assert (_join(pool, pool_lo) - ref_pool).abs().max().item() / s < 2 * tol
    if level == 2:
        # the c8 plane byte for byte; the lo8 half follows the last bits of the kernel's fp32 sum (the reference sums in fp64),
        # the x8 half only differs at e4m3 rounding ties
        from rpnet_b200 import engine
        want = engine.split_c8(ref.permute(0, 2, 3, 1).contiguous())[1]
        diff = (out_lo.cpu() != want).view(n, h2, w2, 2, 64).float()
        assert diff[..., 0, :].mean().item() < 5e-2 and diff[..., 1, :].mean().item() < 1e-3


@pytest.mark.parametrize('case', CASES)
def test_conv_split_vs_fp64(dev, case, level):
    from rpnet_b200 import engine, ops
    n, c0, c1, cout, h, w, k = case
    g = _gen(sum(case))
    x = torch.randn(n, c0 + c1, h, w, generator=g)
    wt = torch.randn(cout, c0 + c1, k, k, generator=g) / math.sqrt((c0 + c1) * k * k)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    ref = (F.conv2d(x.double(), wt.double(), None, padding=k // 2) * scale[None, :, None, None].double()
           + shift[None, :, None, None].double())
    ref = F.relu(ref).float()
    wp, taps = engine.pack_weight_taps(wt.to(dev), split=level)
    assert tuple(wp.shape) == (k * k, cout, 2 * (c0 + c1))
    a = _pair(x[:, :c0], dev, level)
    b = _pair(x[:, c0:], dev, level) if c1 else (None, None)
    out = torch.empty(n, h, w, cout, dtype=torch.float16, device=dev)
    out_lo = _lo_like(out, level)
    out32 = torch.empty(n, h, w, cout, dtype=torch.float32, device=dev)
    pool = torch.empty(n, h // 2, w // 2, cout, dtype=torch.float16, device=dev)
    pool_lo = _lo_like(pool, level)
    ops.conv_split(a[0], wp, taps, scale.to(dev), shift.to(dev), True, src0_lo=a[1], src1=b[0], src1_lo=b[1], w_split=level, out=out,
                   out_lo=out_lo, out_pool=pool, out_pool_lo=pool_lo, out_f32=out32)
    torch.cuda.synchronize()
    s = ref.abs().max().item()
    err32 = (out32.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() / s
    err_pair = (_join(out, out_lo) - ref).abs().max().item() / s
    assert err32 < TOL[level], err32
    assert err_pair < TOL[level] + (2.0 ** -15 if level == 2 else 0), err_pair      # the output planes carry the fp32 result
    assert torch.equal(out, out32.half())                 # hi plane = the plain fp16 rounding
    got_pool = _join(pool, pool_lo)
    assert (got_pool - F.max_pool2d(ref, 2, 2)).abs().max().item() / s < TOL[level] + (2.0 ** -15 if level == 2 else 0)
    if level == 2:
        # the kernel against the same fp8-correction sum in fp64: only the fp32 accumulation order is left
        emu = F.relu(_c8_conv_fp64(x, wt, padding=k // 2) * scale[None, :, None, None].double() + shift[None, :, None, None].double()).float()
        err_emu = (out32.permute(0, 3, 1, 2).cpu() - emu).abs().max().item() / s
        assert err_emu < 5e-6, err_emu
    # the single-term kernel on the same (rounded) operands is two orders of magnitude further away: the test has teeth
    wp1, _ = engine.pack_weight_taps(wt.to(dev))
    o1 = torch.empty_like(out32)
    ops.conv_igemm(a[0], wp1, taps, scale.to(dev), shift.to(dev), True, src1=b[0], out_f32=o1)
    torch.cuda.synchronize()
    err1 = (o1.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() / s
    assert err1 > (10 if level == 1 else 4) * err32, (err1, err32)


def test_conv_c8_activation_range(dev):
    """The fixed e4m3 scales of the c8 planes keep the corrections exact for |x| < 1792 (an un-normalised VGG stack reaches a few
    hundred): activations of magnitude ~1000 stay within the fp32-class bound; ten times larger ones saturate the corrections and fall
    back to single-term accuracy (still finite, never worse than the fp16 conv)."""
    from rpnet_b200 import engine, ops
    g = _gen(77)
    n, cin, cout, h, w = 2, 128, 128, 16, 16
    x0 = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    one, zero = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    wp, taps = engine.pack_weight_taps(wt.to(dev), split=2)
    wp1, _ = engine.pack_weight_taps(wt.to(dev))
    errs = {}
    for amp in (300.0, 3000.0):
        x = x0 * amp                                            # max |x| ~ 4.5 * amp
        ref = F.conv2d(x.double(), wt.double(), None, padding=1).float()
        a = _pair(x, dev, 2)
        o = torch.empty(n, h, w, cout, dtype=torch.float32, device=dev)
        ops.conv_split(a[0], wp, taps, one, zero, False, src0_lo=a[1], w_split=2, out_f32=o)
        o1 = torch.empty_like(o)
        ops.conv_igemm(a[0], wp1, taps, one, zero, False, out_f32=o1)
        torch.cuda.synchronize()
        s = ref.abs().max().item()
        errs[amp] = ((o.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() / s, (o1.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() / s)
        assert torch.isfinite(o).all()
    assert errs[300.0][0] < TOL[2], errs
    assert errs[3000.0][0] < 1.5 * errs[3000.0][1], errs        # saturated corrections: no worse than the single-term conv


def test_conv_split_plain_sources_and_partial_split(dev):
    """lo planes / Wl are optional: with neither the entry point is the single-term conv, bit for bit."""
    from rpnet_b200 import engine, ops
    g = _gen(5)
    n, cin, cout, h, w = 2, 128, 128, 16, 16
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    one, zero = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    a = _pair(x, dev)
    wp1, taps = engine.pack_weight_taps(wt.to(dev))
    o_ref = torch.empty(n, h, w, cout, dtype=torch.float32, device=dev)
    ops.conv_igemm(a[0], wp1, taps, one, zero, False, out_f32=o_ref)
    o = torch.empty_like(o_ref)
    ops.conv_split(a[0], wp1, taps, one, zero, False, w_split=False, out_f32=o)
    torch.cuda.synchronize()
    assert torch.equal(o, o_ref)
    # activations split, weights plain: exact activations x fp16 weights
    ops.conv_split(a[0], wp1, taps, one, zero, False, src0_lo=a[1], w_split=False, out_f32=o)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), wt.half().double(), None, padding=1).float()
    assert (o.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() / ref.abs().max().item() < 2e-5


def test_conv_split_bnstats_and_bn_apply(dev, level):
    """Train-mode split conv: z = hi + lo planes + BatchNorm statistics of the fp32 accumulators (two call groups), then
    bn_apply on z_hi + z_lo writing y and the pooled y as hi / lo planes.  nn.BatchNorm2d batch statistics, net/modules.py:49."""
    from rpnet_b200 import engine, ops
    g = _gen(9)
    n, cin, cout, h, w = 4, 64, 128, 16, 16
    x = torch.randn(n, cin, h, w, generator=g) + 0.5
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    gs = [0, 3, 4]
    a = _pair(x, dev, level)
    wf = torch.empty(9, cout, 2 * cin, dtype=torch.float16, device=dev)
    ops.pack_conv_weight(wt.to(dev), wf, None, split=level)
    wp, taps = engine.pack_weight_taps(wt.to(dev), split=level)
    assert torch.equal(wf.view(torch.uint8), wp.view(torch.uint8))   # the device pack kernel and the torch pack agree bit for bit
    one, zero = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    z, z_lo = (torch.empty(n, h, w, cout, dtype=torch.float16, device=dev) for _ in range(2))
    sums = torch.zeros(2 * cout * 2, dtype=torch.float64, device=dev)
    ops.conv_split(a[0], wf, taps, one, zero, False, src0_lo=a[1], w_split=level, out=z, out_lo=z_lo, group_start=gs, sums=sums)
    stats = torch.empty(2, cout, 4, device=dev)
    rm, rv, nbt = torch.zeros(cout, device=dev), torch.ones(cout, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    ops.bn_finalize(sums, gs, cout, h * w, gamma.to(dev), beta.to(dev), zero, rm, rv, nbt, stats)
    y, p = torch.empty_like(z), torch.empty(n, h // 2, w // 2, cout, dtype=torch.float16, device=dev)
    y_lo, p_lo = _lo_like(y, level), _lo_like(p, level)        # z keeps its fp16 residual plane, y / pool leave as c8 planes
    ops.bn_apply(z, stats, gs, True, y=y, y_pool=p, z_lo=z_lo, y_lo=y_lo, y_pool_lo=p_lo)
    torch.cuda.synchronize()
    zr = F.conv2d(x.double(), wt.double(), None, padding=1)
    ref = torch.cat([F.relu(F.batch_norm(zr[lo:hi], None, None, gamma.double(), beta.double(), True, 0.0, 1e-5))
                     for lo, hi in ((0, 3), (3, 4))]).float()
    assert (_join(z, z_lo) - zr.float()).abs().max().item() / zr.abs().max().item() < TOL[level]
    assert (_join(y, y_lo) - ref).abs().max().item() / ref.abs().max().item() < 2.5 * TOL[level]
    assert (_join(p, p_lo) - F.max_pool2d(ref, 2, 2)).abs().max().item() / ref.abs().max().item() < 2.5 * TOL[level]
    assert int(nbt) == 2
    # bn_stats on the planes == the fused statistics
    s2 = torch.zeros_like(sums)
    ops.bn_stats(z, gs, s2, z_lo=z_lo)
    torch.cuda.synchronize()
    torch.testing.assert_close(s2, sums, rtol=2e-5, atol=2e-5)          # fp32 partial sums grouped differently


@pytest.mark.parametrize('cin', [1, 3])
def test_conv3x3_first_split(dev, cin, level):
    from rpnet_b200 import ops
    g = _gen(3 + cin)
    n, h, w = 2, 40, 24
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(64, cin, 3, 3, generator=g) / 3
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    ref = F.relu(F.conv2d(x.double(), wt.double(), None, padding=1) * scale[None, :, None, None].double() + shift[None, :, None, None].double()).float()
    out = torch.empty(n, h, w, 64, dtype=torch.float16, device=dev)
    lo = _lo_like(out, level)
    ops.conv3x3_first(x.to(dev), wt.to(dev), scale.to(dev), shift.to(dev), True, out, out_lo=lo)
    torch.cuda.synchronize()
    assert (_join(out, lo) - ref).abs().max().item() / ref.abs().max().item() < (1e-5 if level == 1 else 2.0 ** -15)


def test_upconv_split_phases(dev, level):
    """nn.Upsample(x2, nearest) + 3x3 conv (net/modules.py:65-68) as four split-fp16 phase convs."""
    from rpnet_b200 import engine, ops
    g = _gen(11)
    n, cin, cout, h, w = 2, 128, 64, 16, 8
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(F.interpolate(x.double(), scale_factor=2, mode='nearest'), wt.double(), None, padding=1)
                 * scale[None, :, None, None].double() + shift[None, :, None, None].double()).float()
    phases = engine.pack_upsample_phases(wt.to(dev), split=level)
    out = engine.run_upconv(phases, scale.to(dev), shift.to(dev), _pair(x, dev, level), engine.Workspace(), 't', split=True, w_split=level)
    torch.cuda.synchronize()
    assert (_join(*out) - ref).abs().max().item() / ref.abs().max().item() < TOL[level] + (2.0 ** -15 if level == 2 else 0)
    # the device pack of the train path
    wf = torch.empty(4, 4, cout, 2 * cin, dtype=torch.float16, device=dev)
    w16 = torch.empty(16, cin, cout, dtype=torch.bfloat16, device=dev)
    ops.pack_upconv_weight(wt.to(dev), wf, w16, split=level)
    torch.cuda.synchronize()
    for ph, (wp, taps, _) in enumerate(phases):
        torch.testing.assert_close(wf[ph][..., :cin].float(), wp[..., :cin].float(), rtol=0, atol=2e-7)   # sums of taps in a different order
        if level == 1:
            torch.testing.assert_close(wf[ph].float(), wp.float(), rtol=0, atol=2e-7)
        else:                                          # fp8 corrections: the same bytes unless the two sums round apart
            assert (wf[ph].view(torch.uint8) != wp.view(torch.uint8)).float().mean().item() < 2e-3
