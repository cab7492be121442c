"""Per-kernel parity: every C-ABI entry point against the CPU oracle (oracle/rpnet_oracle.py — torch ATen fp32,
the same arithmetic the reference calls) on the same seeded inputs.  Tensor-core kernels take fp16 operands:
the oracle is fed the SAME fp16-rounded values, so the comparison isolates the kernel's arithmetic
(fp32 accumulation) from the documented operand rounding.  Tolerances are stated per test."""
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
    _lib.load()                       # fails loudly if the extension is missing
    return torch.device('cuda:0')


def _h(x):
    """fp16 rounding as fp32 (what the tensor-core operands see)."""
    return x.half().float()


def _nhwc16(x, dev):
    return x.permute(0, 2, 3, 1).contiguous().half().to(dev)


def _nchw32(y):
    return y.float().permute(0, 3, 1, 2).contiguous().cpu()


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------- conv_igemm
CONV_CASES = [
    # n, cin, cout, h, w, k, dil
    (1, 64, 64, 16, 16, 3, 1),
    (2, 64, 128, 32, 32, 3, 1),
    (3, 128, 256, 8, 8, 3, 1),        # batch folded into one pixel tile + ragged n
    (1, 256, 512, 16, 16, 3, 1),      # BN = 256 tiles, long K
    (2, 128, 64, 24, 40, 3, 1),       # ragged H, W (partial tiles)
    (1, 64, 64, 20, 12, 3, 2),        # dilation 2 (VGG last block), w < 16
    (2, 192, 64, 16, 16, 1, 1),       # 1x1
    (5, 64, 64, 4, 4, 3, 1),          # tiny maps, ragged batch tile
    (1, 1024, 128, 8, 8, 3, 1),       # K = 9216
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_igemm_vs_oracle(dev, case):
    from rpnet_b200 import engine, ops
    n, cin, cout, h, w, k, dil = case
    g = _gen(hash(case) % 1000)
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    ref = F.conv2d(_h(x), _h(wt), None, padding=dil * (k // 2), dilation=dil) * scale[None, :, None, None] \
        + shift[None, :, None, None]
    for relu in (True, False):
        want = F.relu(ref) if relu else ref
        wp, taps = engine.pack_weight_taps(wt.to(dev), dil)
        out16 = torch.empty(n, h, w, cout, dtype=torch.float16, device=dev)
        out32 = torch.empty(n, h, w, cout, dtype=torch.float32, device=dev)
        ops.conv_igemm(_nhwc16(x, dev), wp, taps, scale.to(dev), shift.to(dev), relu, out=out16, out_f32=out32)
        torch.cuda.synchronize()
        # fp32 accumulate of identical operands: only summation order differs
        torch.testing.assert_close(_nchw32(out32), want, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(_nchw32(out16), want, rtol=2e-3, atol=2e-3)      # + fp16 output rounding


def test_conv_igemm_fused_pool_and_concat(dev):
    """2x2 max-pool fused in the epilogue (net/unet.py:442) and the two-source K loop for torch.cat (net/unet.py:460)."""
    from rpnet_b200 import engine, ops
    g = _gen(7)
    n, c0, c1, cout, h, w = 2, 128, 64, 128, 24, 32
    a, b = torch.randn(n, c0, h, w, generator=g), torch.randn(n, c1, h, w, generator=g)
    wt = torch.randn(cout, c0 + c1, 3, 3, generator=g) / math.sqrt((c0 + c1) * 9)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(_h(torch.cat([a, b], 1)), _h(wt), None, padding=1) * scale[None, :, None, None]
                 + shift[None, :, None, None])
    wp, taps = engine.pack_weight_taps(wt.to(dev))
    out = torch.empty(n, h, w, cout, dtype=torch.float16, device=dev)
    pool = torch.empty(n, h // 2, w // 2, cout, dtype=torch.float16, device=dev)
    ops.conv_igemm(_nhwc16(a, dev), wp, taps, scale.to(dev), shift.to(dev), True, src1=_nhwc16(b, dev), out=out, out_pool=pool)
    torch.cuda.synchronize()
    torch.testing.assert_close(_nchw32(out), ref, rtol=2e-3, atol=2e-3)
    # pooling is exact on the fp16-rounded activations (max commutes with monotone rounding)
    assert torch.equal(_nchw32(pool), F.max_pool2d(_nchw32(out), 2, 2))
    # pool-only mode (x1 / x2 are never written at full resolution)
    pool2 = torch.zeros_like(pool)
    ops.conv_igemm(_nhwc16(a, dev), wp, taps, scale.to(dev), shift.to(dev), True, src1=_nhwc16(b, dev), out_pool=pool2)
    torch.cuda.synchronize()
    assert torch.equal(pool2, pool)


def test_upconv_subpixel_phases(dev):
    """nn.Upsample(x2, nearest) + 3x3 conv (net/modules.py:65-68) as four 2x2 phase convs."""
    from rpnet_b200 import engine
    g = _gen(11)
    n, cin, cout, h, w = 2, 128, 64, 16, 8
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(F.interpolate(_h(x), scale_factor=2, mode='nearest'), wt, None, padding=1)
                 * scale[None, :, None, None] + shift[None, :, None, None])
    phases = engine.pack_upsample_phases(wt.to(dev))
    out = engine.run_upconv(phases, scale.to(dev), shift.to(dev), _nhwc16(x, dev), engine.Workspace(), 't')
    torch.cuda.synchronize()
    # phase weights are sums of up to four fp32 taps rounded once to fp16: 1e-3-level relative error per weight
    torch.testing.assert_close(_nchw32(out), ref, rtol=5e-3, atol=5e-3)


@pytest.mark.parametrize('cin', [1, 3])
def test_conv3x3_first(dev, cin):
    from rpnet_b200 import ops
    g = _gen(3 + cin)
    n, h, w = 2, 40, 24
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(64, cin, 3, 3, generator=g) / 3
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    ref = F.relu(F.conv2d(x, wt, None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None])
    out = torch.empty(n, h, w, 64, dtype=torch.float16, device=dev)
    ops.conv3x3_first(x.to(dev), wt.to(dev), scale.to(dev), shift.to(dev), True, out)
    torch.cuda.synchronize()
    torch.testing.assert_close(_nchw32(out), ref, rtol=1e-3, atol=1e-3)      # fp32 math, fp16 store


# ------------------------------------------------------------------------------------------- streaming kernels
def test_avgpool_and_premask(dev):
    from rpnet_b200 import ops
    g = _gen(5)
    m = (torch.rand(3, 64, 48, generator=g) > 0.5).float()
    for s in (4, 8):
        out = torch.empty(3, 64 // s, 48 // s, device=dev)
        ops.avgpool_mask(m.to(dev), s, out)
        assert torch.equal(out.cpu(), F.avg_pool2d(m[:, None], s)[:, 0])       # sums of {0,1}: exact
    x = torch.randn(3, 16, 12, 64, generator=g).half()
    pm = torch.rand(3, 16, 12, generator=g)
    fg, bg = torch.empty_like(x, device=dev), torch.empty_like(x, device=dev)
    ops.premask(x.to(dev), pm.to(dev), fg, bg)
    assert torch.equal(fg.cpu(), (x.float() * pm[..., None]).half())
    assert torch.equal(bg.cpu(), (x.float() * (1 - pm[..., None])).half())


def test_local_corr_golden_and_oracle(dev, golden):
    """Correlation (net/rp_net.py:153-181): golden outputs of the reference + oracle at the model's shape."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    g = _gen(9)
    cases = [(torch.randn(2, 256, 16, 24, generator=g), torch.randn(2, 256, 16, 24, generator=g), 5),
             (torch.randn(1, 64, 9, 7, generator=g), torch.randn(1, 64, 9, 7, generator=g), 3),
             (torch.randn(1, 32, 20, 33, generator=g), torch.randn(1, 32, 20, 33, generator=g), 1)]
    for f1, f2, r in cases:
        k = (2 * r + 1) ** 2
        oc = (k + 63) // 64 * 64
        out = torch.full((f1.shape[0], f1.shape[2], f1.shape[3], oc), 7.0, dtype=torch.float16, device=dev)
        ops.local_corr(_nhwc16(f1, dev), _nhwc16(f2, dev), r, out)
        torch.cuda.synchronize()
        want = O.correlation_local(_h(f1), _h(f2), r)
        got = _nchw32(out)
        torch.testing.assert_close(got[:, :k], want, rtol=2e-3, atol=2e-3)
        assert torch.count_nonzero(got[:, k:]) == 0                               # K padding is zero filled
        if f1.shape[2] * f1.shape[3] <= 1024:
            torch.testing.assert_close(got[:, :k], O.correlation_allpairs(_h(f1), _h(f2), r), rtol=2e-3, atol=2e-3)


def test_correlation_function_vs_reference_golden(dev, golden):
    """The reference-signature Correlation() against outputs recorded from the reference itself (channels % 32 cases
    are not in the fixture, so inputs are zero-padded on the channel axis: the correlation scale uses the true C)."""
    from rpnet_b200.nn.rp_net import Correlation
    g = golden('correlation')
    for i in range(int(g['n'])):
        f1, f2, r = torch.from_numpy(g['f1_%d' % i]), torch.from_numpy(g['f2_%d' % i]), int(g['r_%d' % i])
        ref = torch.from_numpy(g['out_%d' % i])
        c = f1.shape[1]
        cp = 32
        pad = lambda t: F.pad(t, (0, 0, 0, 0, 0, cp - c))
        got = Correlation(pad(f1).to(dev), pad(f2).to(dev), r).cpu() * math.sqrt(cp / c)
        torch.testing.assert_close(got, ref, rtol=5e-3, atol=5e-3)               # fp16 operands + fp16 store


def test_masked_avg_pool_and_prototypes(dev, golden):
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    g = golden('proto_loss')
    fts, mask = torch.from_numpy(g['fts']), torch.from_numpy(g['mask'])
    feat = fts.permute(0, 2, 3, 1).contiguous().to(dev)
    out = torch.empty(1, 2, 64, device=dev)
    ops.masked_avg_pool(feat, mask.to(dev), torch.zeros_like(mask).to(dev), out)
    torch.cuda.synchronize()
    # adjoint form re-associates the fp32 sums: 1e-5 relative
    torch.testing.assert_close(out[:, 0].cpu(), torch.from_numpy(g['proto']), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(out[:, 1].cpu(), torch.from_numpy(g['proto_empty']), rtol=0, atol=1e-6)   # empty mask
    # other scales / ragged masks vs the oracle
    gg = _gen(13)
    for (n, h, w, s) in [(3, 16, 16, 4), (2, 8, 12, 8), (1, 32, 32, 4)]:
        f = torch.randn(n, 64, h, w, generator=gg)
        m0 = (torch.rand(n, h * s, w * s, generator=gg) > 0.7).float()
        m1 = 1 - m0
        out = torch.empty(n, 2, 64, device=dev)
        ops.masked_avg_pool(f.permute(0, 2, 3, 1).contiguous().to(dev), m0.to(dev), m1.to(dev), out)
        for i in range(n):
            torch.testing.assert_close(out[i, 0].cpu(), O.get_features(f[[i]], m0[[i]])[0], rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(out[i, 1].cpu(), O.get_features(f[[i]], m1[[i]])[0], rtol=1e-4, atol=1e-5)
    # getPrototype averaging (net/rp_net.py:379-391)
    ways, shots, B = 3, 2, 4
    raw = torch.randn(ways, shots, B, 2, 64, generator=gg)
    protos = torch.empty(B, 1 + ways, 64, device=dev)
    ops.proto_finalize(raw.to(dev), protos)
    for b in range(B):
        fg = [[raw[w, s, b, 0][None] for s in range(shots)] for w in range(ways)]
        bg = [[raw[w, s, b, 1][None] for s in range(shots)] for w in range(ways)]
        fgp, bgp = O.get_prototype(fg, bg)
        torch.testing.assert_close(protos[b, 0].cpu(), bgp[0], rtol=1e-6, atol=1e-7)
        for w in range(ways):
            torch.testing.assert_close(protos[b, 1 + w].cpu(), fgp[w][0], rtol=1e-6, atol=1e-7)


def test_cos_sim(dev, golden):
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    g = golden('proto_loss')
    qf, proto = torch.from_numpy(g['qf']), torch.from_numpy(g['proto'])
    n = qf.shape[0]
    pred = torch.empty(n, 1, 16, 16, device=dev)
    ops.cos_sim(qf.permute(0, 2, 3, 1).contiguous().to(dev), proto.expand(n, 1, 64).contiguous().to(dev), pred)
    torch.testing.assert_close(pred[:, 0].cpu(), torch.from_numpy(g['dist']), rtol=1e-5, atol=1e-5)
    assert pred[0, 0, 3, 4].item() == 0                 # all-zero feature vector -> cosine 0 (eps clamp)
    gg = _gen(17)
    f = torch.randn(3, 64, 9, 11, generator=gg)
    p = torch.randn(3, 5, 64, generator=gg)
    p[1, 2] = 0                                         # zero prototype (empty mask)
    pred = torch.empty(3, 5, 9, 11, device=dev)
    ops.cos_sim(f.permute(0, 2, 3, 1).contiguous().to(dev), p.to(dev), pred)
    for b in range(3):
        for k in range(5):
            torch.testing.assert_close(pred[b, k].cpu(), O.cal_dist(f[[b]], p[b, [k]])[0], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('soft', [False, True])
@pytest.mark.parametrize('shape', [(2, 2, 16, 16, 4), (1, 5, 8, 12, 4), (2, 2, 8, 8, 8)])
def test_upsample_tail(dev, soft, shape):
    """bilinear upsample -> softmax fg prob -> threshold -> avg_pool (net/rp_net.py:303-311)."""
    from rpnet_b200 import ops
    b, p, h, w, s = shape
    pred = torch.randn(b, p, h, w, generator=_gen(b * 100 + p)) * 5
    logits = torch.empty(b, p, h * s, w * s, device=dev)
    m = torch.empty(b, h, w, device=dev)
    ops.upsample_tail(pred.to(dev), logits, m, s, soft)
    ref = F.interpolate(pred, size=(h * s, w * s), mode='bilinear')
    torch.testing.assert_close(logits.cpu(), ref, rtol=1e-6, atol=2e-6)
    # recompute the mask from the kernel's own logits so near-tie pixels cannot flip the comparison
    prob = logits.cpu().softmax(1)[:, 1:].sum(1)
    want = F.avg_pool2d((prob if soft else (prob > 0.5).float())[:, None], s)[:, 0]
    if soft:
        torch.testing.assert_close(m.cpu(), want, rtol=1e-5, atol=1e-6)
    else:
        near_tie = F.avg_pool2d(((prob - 0.5).abs() < 1e-6).float()[:, None], s)[:, 0] > 0
        assert torch.equal(m.cpu()[~near_tie], want[~near_tie])


def test_maxpool_vgg(dev):
    from rpnet_b200 import ops
    x = torch.randn(2, 64, 17, 20, generator=_gen(23)).half()
    for (k, s, p) in [(3, 2, 1), (3, 1, 1), (2, 2, 0)]:
        ho, wo = (17 + 2 * p - k) // s + 1, (20 + 2 * p - k) // s + 1
        out = torch.empty(2, ho, wo, 64, dtype=torch.float16, device=dev)
        ops.maxpool(x.permute(0, 2, 3, 1).contiguous().to(dev), k, s, p, out)
        assert torch.equal(_nchw32(out), F.max_pool2d(x.float(), k, s, p))


def test_error_convention(dev):
    """C ABI: bad arguments return a negative code + message, never crash (include/rpnet_b200.h)."""
    from rpnet_b200 import _lib, ops
    x = torch.zeros(1, 8, 8, 48, dtype=torch.float16, device=dev)          # 48 channels: not a multiple of 64
    w = torch.zeros(1, 64, 48, dtype=torch.float16, device=dev)
    sc = torch.ones(64, device=dev)
    with pytest.raises(_lib.RpnetError, match='multiples of 64'):
        ops.conv_igemm(x, w, [(0, 0)], sc, sc, out=torch.empty(1, 8, 8, 64, dtype=torch.float16, device=dev))
    with pytest.raises(_lib.RpnetError):
        ops.conv_igemm(x.cpu(), w, [(0, 0)], sc, sc, out=torch.empty(1, 8, 8, 64, dtype=torch.float16, device=dev))


@pytest.mark.parametrize('case', [(2, 256, 64, 64, 5), (1, 64, 26, 18, 5), (3, 128, 37, 29, 5), (1, 64, 40, 24, 3), (2, 64, 18, 10, 1),
                                  (1, 128, 33, 50, 2), (1, 64, 30, 31, 4)])
def test_local_corr_tensor_core_path(dev, case):
    """Maps at least one halo window large (w >= 8 + 2r, h >= 16 + 2r, c % 64 == 0) take the tcgen05 banded-GEMM kernel
    (local_corr_tc.cu): same contract as the CUDA-core kernel, incl. ragged tiles, zero padding at the map border, the RAFT
    x/y channel order (D8) and zero-filled padding channels."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    n, c, h, w, r = case
    g = _gen(sum(case))
    f1, f2 = torch.randn(n, c, h, w, generator=g), torch.randn(n, c, h, w, generator=g)
    k = (2 * r + 1) ** 2
    oc = (k + 7) // 8 * 8 if r < 5 else 128
    out = torch.full((n, h, w, oc), 7.0, dtype=torch.float16, device=dev)
    ops.local_corr(_nhwc16(f1, dev), _nhwc16(f2, dev), r, out)
    torch.cuda.synchronize()
    want = O.correlation_local(_h(f1), _h(f2), r)
    got = _nchw32(out)
    torch.testing.assert_close(got[:, :k], want, rtol=2e-3, atol=2e-3)
    assert torch.count_nonzero(got[:, k:]) == 0


@pytest.mark.parametrize('case', [(3, 128, 256, 16, 24, 1, 2, 3), (4, 64, 0, 32, 32, 3, 5, 2), (2, 128, 64, 8, 8, 1, 3, 1)])
def test_conv_cos_fused_epilogue(dev, case):
    """rpnet_conv_cos_f16: 64-channel conv (+affine+ReLU) with calDist (net/rp_net.py:353-363) in the epilogue == the conv
    followed by F.cosine_similarity * 20 against the per-image prototype set (image i -> set i % sets), incl. an all-zero
    feature vector (cosine 0) and the optional fp32 feature output."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    n, c0, c1, h, w, k, P, sets = case
    g = _gen(sum(case))
    cin = c0 + c1
    x = torch.randn(n, cin, h, w, generator=g).half().float()
    wt = (torch.randn(64, cin, k, k, generator=g) / (cin * k * k) ** 0.5).half().float()
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.2
    shift[:] -= 0.1
    protos = torch.randn(sets, P, 64, generator=g)
    y = torch.relu(F.conv2d(x, wt, None, padding=k // 2) * scale[None, :, None, None] + shift[None, :, None, None])
    want = torch.stack([torch.stack([O.cal_dist(y[i:i + 1], protos[i % sets, p][None])[0] for p in range(P)]) for i in range(n)])
    taps = [(ky - k // 2, kx - k // 2) for ky in range(k) for kx in range(k)]
    wp = wt.permute(2, 3, 0, 1).reshape(k * k, 64, cin).half().contiguous().to(dev)
    pred = torch.full((n, P, h, w), 7.0, device=dev)
    feat = torch.empty(n, h, w, 64, device=dev)
    ops.conv_cos(_nhwc16(x[:, :c0], dev), wp, taps, scale.to(dev), shift.to(dev), protos.to(dev), pred, relu=True,
                 src1=_nhwc16(x[:, c0:], dev) if c1 else None, scaler=20.0, out_f32=feat)
    torch.cuda.synchronize()
    torch.testing.assert_close(feat.cpu().permute(0, 3, 1, 2), y, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(pred.cpu(), want, rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize('case', [(2, 256, 64, 64, 2, 2), (3, 128, 37, 29, 3, 1), (1, 64, 26, 18, 5, 1)])
def test_relation_head_fused(dev, case):
    """rpnet_relation_head_f16 == Correlation -> cat([corr, fm1]) -> 1x1 conv + affine + ReLU -> calDist (net/rp_net.py:79-84,
    287-303), incl. ragged tiles and per-image prototype sets."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import ops
    n, c, h, w, P, sets = case
    r, k = 5, 121
    g = _gen(sum(case))
    f1 = torch.randn(n, c, h, w, generator=g).relu().half().float()
    f2 = torch.randn(n, c, h, w, generator=g).relu().half().float()
    wq = (torch.randn(64, k + c, 1, 1, generator=g) / (k + c) ** 0.5).half().float()
    scale, shift = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.2
    protos = torch.randn(sets, P, 64, generator=g)
    corr = O.correlation_local(f1, f2, r).half().float()                   # the unfused path stores corr as fp16 too
    y = torch.relu(F.conv2d(torch.cat([corr, f1], 1), wq) * scale[None, :, None, None] + shift[None, :, None, None])
    want = torch.stack([torch.stack([O.cal_dist(y[i:i + 1], protos[i % sets, p][None])[0] for p in range(P)]) for i in range(n)])
    wpad = torch.zeros(64, 128 + c)
    wpad[:, :k] = wq[:, :k, 0, 0]
    wpad[:, 128:] = wq[:, k:, 0, 0]
    pred = torch.full((n, P, h, w), 7.0, device=dev)
    assert ops.relation_head_supported(_nhwc16(f1, dev), r)
    ops.relation_head(_nhwc16(f1, dev), _nhwc16(f2, dev), wpad.half().reshape(1, 64, 128 + c).contiguous().to(dev), scale.to(dev),
                      shift.to(dev), protos.to(dev), pred, r, 20.0)
    torch.cuda.synchronize()
    torch.testing.assert_close(pred.cpu(), want, rtol=2e-3, atol=3e-2)


_VARIANT_CHILD = r'''
import hashlib, os, sys
sys.path.insert(0, os.getcwd())
import torch
from rpnet_b200 import ops
dev = torch.device('cuda:0')
taps = [(dy, dx) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
# (n, h, w, c0, c1, cout): CTA-pair shapes (cout % 256 == 0; odd tile counts, ragged maps), weights-stationary shape (64 -> 64, many tiles)
for (n, h, w, c0, c1, cout) in [(5, 64, 64, 256, 256, 256), (3, 24, 40, 64, 0, 512), (1, 16, 8, 128, 0, 256), (20, 256, 256, 64, 0, 64)]:
    torch.manual_seed(1)
    x0 = torch.randn(n, h, w, c0, device=dev).half()
    x1 = torch.randn(n, h, w, c1, device=dev).half() if c1 else None
    wf = (torch.randn(9, cout, c0 + c1, device=dev) * 0.05).half()
    sc, sh = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
    out = torch.empty(n, h, w, cout, device=dev, dtype=torch.float16)
    pool = torch.empty(n, h // 2, w // 2, cout, device=dev, dtype=torch.float16)
    ops.conv_igemm(x0, wf, taps, sc, sh, True, src1=x1, out=out, out_pool=pool)
    torch.cuda.synchronize()
    print(hashlib.md5(out.cpu().numpy().tobytes()).hexdigest(), hashlib.md5(pool.cpu().numpy().tobytes()).hexdigest())
'''


def test_conv_variants_bit_identical(dev):
    """The CTA-pair (cta_group::2) and weights-stationary variants of conv_igemm produce the same bits as the plain
    single-CTA kernel (the variant is chosen per process: RPNET_CONV_2CTA / RPNET_CONV_NO_WS)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for env_extra in ({}, {'RPNET_CONV_2CTA': '0', 'RPNET_CONV_NO_WS': '1'}):
        r = subprocess.run([sys.executable, '-c', _VARIANT_CHILD], cwd=root, env=dict(os.environ, **env_extra), capture_output=True,
                           text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines())
    assert len(outs[0]) == 4 and outs[0] == outs[1], (outs[0], outs[1])
