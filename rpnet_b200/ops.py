"""Thin torch-tensor wrappers over the C ABI (include/rpnet_b200.h).  Pointers + sizes only cross the
boundary; kernels run asynchronously on torch's current CUDA stream.  No fallbacks."""
import ctypes

import torch

from . import _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- launch accounting (bench.py): number of kernels launched and, optionally, per-kernel CUDA-event timing on
# the launching stream.  _PROFILE maps kernel name -> list of (start_event, end_event, work) tuples.
LAUNCHES = 0
_PROFILE = None


def set_profiler(store):
    """store: dict to fill (or None to switch off)."""
    global _PROFILE
    _PROFILE = store


class _Timed:
    __slots__ = ('name', 'work', 'n', 'e0')

    def __init__(self, name, work=0.0, n=1):
        self.name, self.work, self.n = name, work, n

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += self.n
        if _PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _PROFILE.setdefault(self.name, []).append((self.e0, e1, self.work))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _req(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise _lib.RpnetError('%s must be a contiguous CUDA %s tensor (got %s %s contiguous=%s)'
                              % (name, dtype, t.device, t.dtype, t.is_contiguous()))
    if t.device.index != torch.cuda.current_device():
        # kernels, TMA descriptors and function attributes belong to the CURRENT device and launch on its current stream
        raise _lib.RpnetError('%s lives on %s but the current CUDA device is cuda:%d: run the call under `with torch.cuda.device(%d):` '
                              '(RP_Net.forward / TrainEngine / segment_volume do this for the model\'s device)'
                              % (name, t.device, torch.cuda.current_device(), t.device.index))
    return t


def _lo_fmt(pairs):
    """Format of the lo planes of a call (include/rpnet_b200.h `lo_fmt`): every (lo, hi, name) with lo not None must be either the
    fp16 residual plane of hi (same shape; returns 0) or its c8 plane (uint8, last dim doubled; returns 1) — all the same."""
    fmt = None
    for t, ref, nm in pairs:
        if t is None:
            continue
        if ref is None:
            raise _lib.RpnetError('%s given without its main plane' % nm)
        if t.dtype == torch.uint8:
            _req(t, torch.uint8, nm)
            ok, f = tuple(t.shape) == tuple(ref.shape[:-1]) + (2 * ref.shape[-1],) and ref.shape[-1] % 64 == 0, 1
        else:
            _req(t, torch.float16, nm)
            ok, f = t.shape == ref.shape, 0
        if not ok:
            raise _lib.RpnetError('%s %s does not match its main plane %s' % (nm, tuple(t.shape), tuple(ref.shape)))
        if fmt is not None and fmt != f:
            raise _lib.RpnetError('%s: fp16 residual planes and c8 planes mixed in one call' % nm)
        fmt = f
    return fmt or 0


def conv_igemm(src0, wpack, taps, scale, shift, relu=True, src1=None, out=None, out_pool=None, out_f32=None,
               out_map=None, out_coff=0):
    """src0/src1: fp16 NHWC [n,h,w,c]; wpack fp16 [ntaps,cout,cin]; taps: list of (dy, dx).
    out: fp16 NHWC [n,oh,ow,oc]; out_map = (oy_mul, oy_off, ox_mul, ox_off) (default identity)."""
    lib = _lib.load()
    _req(src0, torch.float16, 'src0'); _req(wpack, torch.float16, 'wpack')
    _req(scale, torch.float32, 'scale'); _req(shift, torch.float32, 'shift')
    n, h, w, c0 = src0.shape
    c1 = 0
    if src1 is not None:
        _req(src1, torch.float16, 'src1')
        assert src1.shape[:3] == src0.shape[:3]
        c1 = src1.shape[3]
    ntaps, cout, cin = wpack.shape
    if cin != c0 + c1 or ntaps != len(taps) or scale.numel() != cout or shift.numel() != cout:
        raise _lib.RpnetError('conv_igemm: weight pack %s does not match sources (%d + %d channels, %d taps)'
                              % (tuple(wpack.shape), c0, c1, len(taps)))
    dy = (ctypes.c_int * ntaps)(*[int(t[0]) for t in taps])
    dx = (ctypes.c_int * ntaps)(*[int(t[1]) for t in taps])
    oh = ow = oc = 0
    om = out_map or (1, 0, 1, 0)
    if out is not None:
        _req(out, torch.float16, 'out')
        assert out.shape[0] == n
        oh, ow, oc = out.shape[1:]
    if out_pool is not None:
        _req(out_pool, torch.float16, 'out_pool')
        assert tuple(out_pool.shape) == (n, h // 2, w // 2, cout), (tuple(out_pool.shape), (n, h // 2, w // 2, cout))
    if out_f32 is not None:
        _req(out_f32, torch.float32, 'out_f32')
        assert tuple(out_f32.shape) == (n, h, w, cout)
    with _Timed('conv_igemm', 2.0 * n * h * w * cout * cin * ntaps):
        rc = lib.rpnet_conv_igemm_f16(_ptr(src0), c0, _ptr(src1), c1, n, h, w, _ptr(wpack), ntaps, dy, dx, cout, _ptr(scale),
                                      _ptr(shift), int(bool(relu)), _ptr(out), oh, ow, oc, out_coff, om[0], om[1], om[2], om[3],
                                      _ptr(out_pool), _ptr(out_f32), _stream())
    _lib.check(rc, 'rpnet_conv_igemm_f16')


def conv_cos(src0, wpack, taps, scale, shift, protos, pred, relu=True, src1=None, scaler=20.0, out_f32=None):
    """64-channel conv + fused calDist epilogue: pred fp32 [n, P, h, w] from protos fp32 [sets, P, 64] (image i -> set i % sets)."""
    lib = _lib.load()
    _req(src0, torch.float16, 'src0'); _req(wpack, torch.float16, 'wpack'); _req(protos, torch.float32, 'protos'); _req(pred, torch.float32, 'pred')
    n, h, w, c0 = src0.shape
    c1 = 0
    if src1 is not None:
        _req(src1, torch.float16, 'src1')
        c1 = src1.shape[3]
    ntaps, cout, cin = wpack.shape
    sets, P = protos.shape[0], protos.shape[1]
    if cout != 64 or cin != c0 + c1 or ntaps != len(taps) or tuple(pred.shape) != (n, P, h, w) or protos.shape[2] != 64:
        raise _lib.RpnetError('conv_cos: shapes do not match (wpack %s, pred %s, protos %s)' % (tuple(wpack.shape), tuple(pred.shape), tuple(protos.shape)))
    dy, dx = _taps(taps)
    with _Timed('conv_cos', 2.0 * n * h * w * cout * cin * ntaps):
        rc = lib.rpnet_conv_cos_f16(_ptr(src0), c0, _ptr(src1), c1, n, h, w, _ptr(wpack), ntaps, dy, dx, _ptr(scale), _ptr(shift),
                                    int(bool(relu)), _ptr(protos), P, sets, float(scaler), _ptr(pred), _ptr(out_f32), _stream())
    _lib.check(rc, 'rpnet_conv_cos_f16')


def conv3x3_first(img, weight, scale, shift, relu, out, out_lo=None):
    """out_lo: optional lo plane of the split representation — the fp16 residual fp16(y - fp16(y)), or (a uint8 tensor with the last
    dim doubled) the c8 plane of include/rpnet_b200.h "fp8 corrections"."""
    lib = _lib.load()
    _req(img, torch.float32, 'img'); _req(weight, torch.float32, 'weight'); _req(out, torch.float16, 'out')
    n, cin, h, w = img.shape
    assert tuple(weight.shape) == (64, cin, 3, 3) and tuple(out.shape) == (n, h, w, 64)
    lo_fmt = _lo_fmt([(out_lo, out, 'out_lo')])
    with _Timed('conv3x3_first', float(img.numel() * 4 + out.numel() * (2 if out_lo is None else 4))):
        rc = lib.rpnet_conv3x3_first_split_f16(_ptr(img), n, cin, h, w, _ptr(weight), _ptr(_req(scale, torch.float32, 'scale')),
                                               _ptr(_req(shift, torch.float32, 'shift')), int(bool(relu)), _ptr(out), _ptr(out_lo),
                                               int(lo_fmt), _stream())
    _lib.check(rc, 'rpnet_conv3x3_first_split_f16')


def conv_split(src0, wpack, taps, scale, shift, relu=True, src0_lo=None, src1=None, src1_lo=None, w_split=True, out=None,
               out_lo=None, out_pool=None, out_pool_lo=None, out_f32=None, out_map=None, out_coff=0, group_start=None, sums=None,
               keep_sums=False, res=None, res_lo=None):
    """Split-fp16 tap-list conv (include/rpnet_b200.h, rpnet_conv_split_f16): sources and outputs as hi / lo fp16 planes,
    wpack fp16 [ntaps, cout, (c0 + c1) * (2 if w_split else 1)] = Wh | Wl.  w_split = 2: fp8 corrections — the lo planes (sources,
    residual: uint8, last dim doubled) are c8 planes and the second half of the pack holds Wh8 | Wl8; the OUTPUT lo planes may be c8
    planes or fp16 residuals (the pre-BatchNorm z of the train path).  sums (fp64) + group_start: train-mode BatchNorm
    statistics of the fp32 accumulators (scale = 1, shift = 0, relu = False).  res / res_lo: residual planes [n, h, w, cout] added
    after the affine and before the ReLU (BasicBlock)."""
    lib = _lib.load()
    _req(src0, torch.float16, 'src0'); _req(wpack, torch.float16, 'wpack')
    _req(scale, torch.float32, 'scale'); _req(shift, torch.float32, 'shift')
    n, h, w, c0 = src0.shape
    c1 = 0
    in_fmt = _lo_fmt([(src0_lo, src0, 'src0_lo'), (src1_lo, src1, 'src1_lo'), (res_lo, res, 'res_lo')])
    if (int(w_split) == 2) != (in_fmt == 1):
        raise _lib.RpnetError('conv_split: w_split = %d needs %s source lo planes' % (int(w_split), 'c8' if int(w_split) == 2 else 'fp16'))
    if src1 is not None:
        _req(src1, torch.float16, 'src1')
        assert src1.shape[:3] == src0.shape[:3]
        c1 = src1.shape[3]
        assert (src0_lo is None) == (src1_lo is None)
    ntaps, cout, cin_pack = wpack.shape
    cin = c0 + c1
    if cin_pack != cin * (2 if w_split else 1) or ntaps != len(taps) or scale.numel() != cout or shift.numel() != cout:
        raise _lib.RpnetError('conv_split: weight pack %s does not match sources (%d + %d channels, %d taps, w_split=%s)'
                              % (tuple(wpack.shape), c0, c1, len(taps), w_split))
    dy, dx = _taps(taps)
    oh = ow = oc = 0
    om = out_map or (1, 0, 1, 0)
    if out is not None:
        _req(out, torch.float16, 'out')
        assert out.shape[0] == n
        oh, ow, oc = out.shape[1:]
    if out_pool is not None:
        _req(out_pool, torch.float16, 'out_pool')
        assert tuple(out_pool.shape) == (n, h // 2, w // 2, cout)
    out_fmt = _lo_fmt([(out_lo, out, 'out_lo'), (out_pool_lo, out_pool, 'out_pool_lo')])
    if out_fmt == 1 and int(w_split) != 2:
        raise _lib.RpnetError('conv_split: c8 output planes need w_split = 2')
    have_lo_out = out_lo is not None or out_pool_lo is not None
    w_mode = (3 if (have_lo_out and out_fmt == 0) else 2) if int(w_split) == 2 else int(bool(w_split))
    if out_f32 is not None:
        _req(out_f32, torch.float32, 'out_f32')
        assert tuple(out_f32.shape) == (n, h, w, cout)
    gs, g = (None, 0)
    if sums is not None:
        _req(sums, torch.float64, 'sums')
        gs, g = _groups(group_start)
        assert sums.numel() >= g * cout * 2
    # `work` stays the reference's (algorithmic) FLOPs of the conv; the kernel executes 1 + (lo planes) + (Wl) passes of them
    with _Timed('conv_igemm', 2.0 * n * h * w * cout * cin * ntaps):
        if res is not None:
            _req(res, torch.float16, 'res')
            assert tuple(res.shape) == (n, h, w, cout)
        rc = lib.rpnet_conv_split_res_f16(_ptr(src0), _ptr(src0_lo), c0, _ptr(src1), _ptr(src1_lo), c1, n, h, w, _ptr(wpack),
                                          w_mode, ntaps, dy, dx, cout, _ptr(scale), _ptr(shift), _ptr(res), _ptr(res_lo),
                                          int(bool(relu)), _ptr(out), _ptr(out_lo), oh, ow, oc, out_coff, om[0], om[1], om[2], om[3],
                                          _ptr(out_pool), _ptr(out_pool_lo), _ptr(out_f32), gs, g, _ptr(sums), int(bool(keep_sums)),
                                          _stream())
    _lib.check(rc, 'rpnet_conv_split_res_f16')


def avgpool_mask(mask, s, out):
    lib = _lib.load()
    _req(mask, torch.float32, 'mask'); _req(out, torch.float32, 'out')
    n, h, w = mask.shape
    assert tuple(out.shape) == (n, h // s, w // s)
    with _Timed('avgpool_mask', float(mask.numel() * 4 + out.numel() * 4)):
        _lib.check(lib.rpnet_avgpool_mask_f32(_ptr(mask), _ptr(out), n, h, w, s, _stream()), 'rpnet_avgpool_mask_f32')


def premask(x, m, x_fg, x_bg):
    lib = _lib.load()
    _req(x, torch.float16, 'x'); _req(m, torch.float32, 'm'); _req(x_fg, torch.float16, 'x_fg'); _req(x_bg, torch.float16, 'x_bg')
    c = x.shape[-1]
    pixels = x.numel() // c
    assert m.numel() == pixels and x_fg.shape == x.shape and x_bg.shape == x.shape
    with _Timed('premask', float(x.numel() * 6 + m.numel() * 4)):
        _lib.check(lib.rpnet_premask_f16(_ptr(x), _ptr(m), _ptr(x_fg), _ptr(x_bg), pixels, c, _stream()), 'rpnet_premask_f16')


def local_corr(f1, f2, radius, out):
    lib = _lib.load()
    _req(f1, torch.float16, 'f1'); _req(f2, torch.float16, 'f2'); _req(out, torch.float16, 'out')
    n, h, w, c = f1.shape
    assert f2.shape == f1.shape and tuple(out.shape[:3]) == (n, h, w)
    with _Timed('local_corr', float(f1.numel() * 4 + out.numel() * 2)):
        _lib.check(lib.rpnet_local_corr_f16(_ptr(f1), _ptr(f2), _ptr(out), n, h, w, c, radius, out.shape[3], _stream()),
                   'rpnet_local_corr_f16')


def relation_head_supported(f1, radius):
    n, h, w, c = f1.shape
    return c % 64 == 0 and c <= 256 and radius in (3, 5) and w >= 8 + 2 * radius and h >= 16 + 2 * radius


def relation_head(f1, f2, wq_pack, scale, shift, protos, pred, radius, scaler=20.0):
    """Fused eval tail: corr(f1, f2) -> [corr | f1] -> 1x1 conv + affine + ReLU -> cosine vs protos -> pred fp32 [n, P, h, w]."""
    lib = _lib.load()
    _req(f1, torch.float16, 'f1'); _req(f2, torch.float16, 'f2'); _req(wq_pack, torch.float16, 'wq_pack')
    _req(protos, torch.float32, 'protos'); _req(pred, torch.float32, 'pred')
    n, h, w, c = f1.shape
    sets, P = protos.shape[0], protos.shape[1]
    if tuple(wq_pack.shape) != (1, 64, 128 + c) or tuple(pred.shape) != (n, P, h, w) or f2.shape != f1.shape:
        raise _lib.RpnetError('relation_head: shapes do not match (wq %s, pred %s)' % (tuple(wq_pack.shape), tuple(pred.shape)))
    with _Timed('relation_head', float(f1.numel() * 4 + pred.numel() * 4)):
        rc = lib.rpnet_relation_head_f16(_ptr(f1), _ptr(f2), _ptr(wq_pack), _ptr(scale), _ptr(shift), _ptr(protos), P, sets, float(scaler),
                                         _ptr(pred), n, h, w, c, radius, _stream())
    _lib.check(rc, 'rpnet_relation_head_f16')


def masked_avg_pool(feat, mask0, mask1, out):
    lib = _lib.load()
    _req(feat, torch.float32, 'feat'); _req(mask0, torch.float32, 'mask0'); _req(mask1, torch.float32, 'mask1')
    _req(out, torch.float32, 'out')
    n, h, w, c = feat.shape
    assert mask0.shape == mask1.shape and mask0.shape[0] == n and tuple(out.shape) == (n, 2, c)
    with _Timed('masked_avg_pool', float(feat.numel() * 8 + mask0.numel() * 8)):
        _lib.check(lib.rpnet_masked_avg_pool_f32(_ptr(feat), _ptr(mask0), _ptr(mask1), _ptr(out), n, h, w, c, mask0.shape[1],
                                                 mask0.shape[2], _stream()), 'rpnet_masked_avg_pool_f32')


def proto_finalize(raw, protos):
    lib = _lib.load()
    _req(raw, torch.float32, 'raw'); _req(protos, torch.float32, 'protos')
    ways, shots, batch, two, c = raw.shape
    assert two == 2 and tuple(protos.shape) == (batch, 1 + ways, c)
    with _Timed('proto_finalize', float(raw.numel() * 4)):
        _lib.check(lib.rpnet_proto_finalize_f32(_ptr(raw), _ptr(protos), ways, shots, batch, c, _stream()), 'rpnet_proto_finalize_f32')


def cos_sim(feat, protos, pred, scaler=20.0):
    lib = _lib.load()
    _req(feat, torch.float32, 'feat'); _req(protos, torch.float32, 'protos'); _req(pred, torch.float32, 'pred')
    b, h, w, c = feat.shape
    p = protos.shape[1]
    assert tuple(protos.shape) == (b, p, c) and tuple(pred.shape) == (b, p, h, w)
    with _Timed('cos_sim', float(feat.numel() * 4 + pred.numel() * 4)):
        _lib.check(lib.rpnet_cos_sim_f32(_ptr(feat), _ptr(protos), _ptr(pred), b, h * w, c, p, float(scaler), _stream()), 'rpnet_cos_sim_f32')


def upsample_tail(pred, logits, mask_out, scale, soft_mask):
    lib = _lib.load()
    _req(pred, torch.float32, 'pred'); _req(logits, torch.float32, 'logits'); _req(mask_out, torch.float32, 'mask_out')
    b, p, h, w = pred.shape
    assert tuple(logits.shape) == (b, p, h * scale, w * scale) and mask_out.numel() == b * h * w
    with _Timed('upsample_tail', float(pred.numel() * 4 + logits.numel() * 4 + mask_out.numel() * 4)):
        _lib.check(lib.rpnet_upsample_tail_f32(_ptr(pred), _ptr(logits), _ptr(mask_out), b, p, h, w, scale, int(bool(soft_mask)),
                                               _stream()), 'rpnet_upsample_tail_f32')


def maxpool(x, k, stride, pad, out, x_lo=None, out_lo=None, idx=None):
    """x_lo / out_lo: lo planes of a split activation, fp16 residuals or c8 planes (max of hi + lo, re-split); idx: optional uint8 tensor like `out`
    that receives the window position of the first maximum (for maxpool_bwd)."""
    lib = _lib.load()
    _req(x, torch.float16, 'x'); _req(out, torch.float16, 'out')
    n, h, w, c = x.shape
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    assert tuple(out.shape) == (n, ho, wo, c) and (x_lo is None) == (out_lo is None)
    lo_fmt = _lo_fmt([(x_lo, x, 'x_lo'), (out_lo, out, 'out_lo')])
    if idx is not None:
        _req(idx, torch.uint8, 'idx')
        assert idx.shape == out.shape
    with _Timed('maxpool', float((x.numel() * 2 + out.numel() * 2) * (1 if x_lo is None else 2))):
        _lib.check(lib.rpnet_maxpool_idx_f16(_ptr(x), _ptr(x_lo), _ptr(out), _ptr(out_lo), int(lo_fmt), _ptr(idx), n, h, w, c, k, stride, pad,
                                             _stream()),
                   'rpnet_maxpool_idx_f16')
    

# =====================================================================================================
# training entry points (include/rpnet_b200.h, "Training path")
# =====================================================================================================
bf16 = torch.bfloat16


def _groups(group_start):
    g = [int(v) for v in group_start]
    return (ctypes.c_int * len(g))(*g), len(g) - 1


def _taps(taps):
    n = len(taps)
    return (ctypes.c_int * n)(*[int(t[0]) for t in taps]), (ctypes.c_int * n)(*[int(t[1]) for t in taps])


def conv_dgrad(dz, wpack_t, taps, out, out_coff=0):
    """Data gradient of a tap-list conv: dz bf16 NHWC [n,h,w,cout], wpack_t bf16 [ntaps, cin, cout] (transposed pack),
    taps = the FORWARD tap list (negated here); out bf16 NHWC [n,h,w,>=cin]."""
    lib = _lib.load()
    _req(dz, bf16, 'dz'); _req(wpack_t, bf16, 'wpack_t'); _req(out, bf16, 'out')
    n, h, w, cout = dz.shape
    ntaps, cin, cout2 = wpack_t.shape
    assert cout2 == cout and ntaps == len(taps) and out.shape[:3] == dz.shape[:3]
    dy, dx = _taps([(-t[0], -t[1]) for t in taps])
    one, zero = _const_vec(cin, 1.0, dz.device), _const_vec(cin, 0.0, dz.device)
    with _Timed('conv_igemm', 2.0 * n * h * w * cout * cin * ntaps):
        rc = lib.rpnet_conv_igemm_bf16(_ptr(dz), cout, None, 0, n, h, w, _ptr(wpack_t), ntaps, dy, dx, cin, _ptr(one), _ptr(zero), 0,
                                       _ptr(out), h, w, out.shape[3], out_coff, 1, 0, 1, 0, None, None, _stream())
    _lib.check(rc, 'rpnet_conv_igemm_bf16')


_CONST = {}


def _const_vec(n, val, device):
    key = (n, val, str(device))
    t = _CONST.get(key)
    if t is None:
        t = torch.full((n,), val, dtype=torch.float32, device=device)
        _CONST[key] = t
    return t


def conv_wgrad_workspace_bytes(c0, c1, n, h, w, ntaps, cout):
    r = _lib.load().rpnet_conv_wgrad_workspace_bytes(c0, c1, n, h, w, ntaps, cout)
    if r < 0:
        _lib.check(int(r), 'rpnet_conv_wgrad_workspace_bytes')
    return int(r)


def cvt_f16_to_bf16(x, out):
    lib = _lib.load()
    _req(x, torch.float16, 'x'); _req(out, bf16, 'out')
    assert out.numel() >= x.numel()
    with _Timed('cvt_f16_to_bf16', float(x.numel() * 4)):
        _lib.check(lib.rpnet_cvt_f16_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), 'rpnet_cvt_f16_to_bf16')


_CVT = {}


def _as_bf16(x, slot):
    """fp16 activations -> bf16 copy in a persistent scratch buffer (one per operand slot, grown on demand)."""
    if x is None or x.dtype == bf16:
        return x
    key = (slot, str(x.device))
    buf = _CVT.get(key)
    if buf is None or buf.numel() < x.numel():
        buf = torch.empty(x.numel(), dtype=bf16, device=x.device)
        _CVT[key] = buf
    out = buf[:x.numel()].view(x.shape)
    cvt_f16_to_bf16(x, out)
    return out


def conv_wgrad(x0, dz, taps, grad, workspace, x1=None, hole=(0, 0), accumulate=True):
    """grad fp32 [cout, cin_real, kh, kw] (+)= sum dz * shifted x.  x0/x1 fp16 or bf16 NHWC (fp16 is converted to a bf16
    scratch copy first: the tensor-core GEMM needs one operand format), dz bf16 NHWC."""
    lib = _lib.load()
    _req(dz, bf16, 'dz'); _req(grad, torch.float32, 'grad')
    if x0.dtype not in (torch.float16, bf16) or not x0.is_contiguous():
        raise _lib.RpnetError('conv_wgrad: x0 must be contiguous fp16/bf16')
    if x1 is not None and x1.dtype != x0.dtype:
        x0, x1 = _as_bf16(x0, 0), _as_bf16(x1, 1)        # mixed formats: fall back to explicit bf16 copies
    n, h, w, c0 = x0.shape
    c1 = 0 if x1 is None else x1.shape[3]
    cout = dz.shape[3]
    assert tuple(dz.shape[:3]) == (n, h, w) and grad.numel() == cout * (c0 + c1 - hole[1]) * len(taps)
    dy, dx = _taps(taps)
    with _Timed('conv_wgrad', 2.0 * n * h * w * cout * (c0 + c1) * len(taps), n=2):
        rc = lib.rpnet_conv_wgrad(_ptr(x0), c0, _ptr(x1), c1, int(x0.dtype == bf16), _ptr(dz), n, h, w, len(taps), dy, dx, cout,
                                  _ptr(grad), hole[0], hole[1], int(bool(accumulate)), _ptr(workspace),
                                  workspace.numel() * workspace.element_size(), _stream())
    _lib.check(rc, 'rpnet_conv_wgrad')


def _scratch64(n, device, slot):
    """Small persistent fp64 scratch (reduction accumulators; contents dead after the call)."""
    key = ('f64', slot, str(device))
    t = _CONST.get(key)
    if t is None or t.numel() < n:
        t = torch.empty(max(int(n), 1024), dtype=torch.float64, device=device)
        _CONST[key] = t
    return t


def conv3x3_first_wgrad(img, dz, grad):
    """grad fp32 [64, cin, 3, 3] += sum dz * shifted img; img fp32 NCHW [n, cin, h, w] (cin <= 4), dz bf16 NHWC [n, h, w, 64]."""
    lib = _lib.load()
    _req(img, torch.float32, 'img'); _req(dz, bf16, 'dz'); _req(grad, torch.float32, 'grad')
    n, cin, h, w = img.shape
    assert 1 <= cin <= 4 and tuple(dz.shape) == (n, h, w, 64) and grad.numel() == 64 * cin * 9
    with _Timed('conv3x3_first_wgrad', float(img.numel() * 4 + dz.numel() * 2 * cin), n=2 * cin):
        _lib.check(lib.rpnet_conv3x3_first_wgrad_cin(_ptr(img), cin, _ptr(dz), n, h, w, _ptr(grad),
                                                     _ptr(_scratch64(576, img.device, 'first_wgrad')), _stream()), 'rpnet_conv3x3_first_wgrad_cin')


def relu_bias_bwd(dy, y, g, dbias):
    """g bf16 = dy * [y > 0] (y fp16 NHWC or None = no ReLU); dbias fp32 [c] += sum over pixels of g."""
    lib = _lib.load()
    _req(dy, bf16, 'dy'); _req(g, bf16, 'g'); _req(dbias, torch.float32, 'dbias')
    c = dy.shape[-1]
    pixels = dy.numel() // c
    assert g.shape == dy.shape and dbias.numel() == c
    if y is not None:
        _req(y, torch.float16, 'y')
        assert y.shape == dy.shape
    with _Timed('relu_bias_bwd', float(dy.numel() * (6 if y is not None else 4)), n=2):
        _lib.check(lib.rpnet_relu_bias_bwd(_ptr(dy), _ptr(y), pixels, c, _ptr(g), _ptr(dbias), _ptr(_scratch64(c, dy.device, 'relu_bias')),
                                           _stream()), 'rpnet_relu_bias_bwd')


def add_relu_mask(a, out, b=None, y=None):
    """out = (a + b) where y > 0 else 0 (b, y optional): a, b, out bf16, y fp16, same shapes."""
    lib = _lib.load()
    _req(a, bf16, 'a'); _req(out, bf16, 'out')
    assert out.shape == a.shape and a.numel() % 8 == 0
    if b is not None:
        _req(b, bf16, 'b'); assert b.shape == a.shape
    if y is not None:
        _req(y, torch.float16, 'y'); assert y.shape == a.shape
    with _Timed('add_relu_mask', float(a.numel() * (4 + (2 if b is not None else 0) + (2 if y is not None else 0)))):
        _lib.check(lib.rpnet_add_relu_mask_bf16(_ptr(a), _ptr(b), _ptr(y), _ptr(out), a.numel(), _stream()), 'rpnet_add_relu_mask_bf16')


def conv7x7s2_stem_wgrad(img, dz, grad):
    """grad fp32 [64, 3, 7, 7] += weight gradient of the stem conv from img fp32 [n, 3, H, W] and dz bf16 [n, H/2, W/2, 64]."""
    lib = _lib.load()
    _req(img, torch.float32, 'img'); _req(dz, bf16, 'dz'); _req(grad, torch.float32, 'grad')
    n, c, h, w = img.shape
    assert c == 3 and tuple(dz.shape) == (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, 64) and grad.numel() == 64 * 147
    with _Timed('conv7x7s2_stem_wgrad', float(img.numel() * 4 + dz.numel() * 2)):
        _lib.check(lib.rpnet_conv7x7s2_stem_wgrad(_ptr(img), _ptr(dz), n, h, w, _ptr(grad), _ptr(_scratch64(64 * 147, img.device, 'stem_wgrad')),
                                                  _stream()), 'rpnet_conv7x7s2_stem_wgrad')


def maxpool_bwd(dy, idx, k, stride, pad, dx):
    """dx bf16 [n, h, w, c] from dy bf16 [n, ho, wo, c] and the argmax positions idx uint8 [n, ho, wo, c] of maxpool(..., idx=)."""
    lib = _lib.load()
    _req(dy, bf16, 'dy'); _req(idx, torch.uint8, 'idx'); _req(dx, bf16, 'dx')
    n, h, w, c = dx.shape
    assert idx.shape == dy.shape and tuple(dy.shape) == (n, (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1, c)
    with _Timed('maxpool_bwd', float(dy.numel() * 3 + dx.numel() * 2)):
        _lib.check(lib.rpnet_maxpool_bwd_bf16(_ptr(dy), _ptr(idx), _ptr(dx), n, h, w, c, k, stride, pad, _stream()), 'rpnet_maxpool_bwd_bf16')


def pack_conv_weight(w, w_fwd=None, w_dgrad=None, hole=(0, 0), split=False):
    """w fp32 [cout, cin_real, kh, kw] -> w_fwd fp16 [taps, cout, cin] (split 1: [taps, cout, 2 * cin] = Wh | Wl; split 2: Wh | fp8
    corrections, same shape) / w_dgrad bf16 [taps, cin, cout]."""
    lib = _lib.load()
    _req(w, torch.float32, 'w')
    cout, cin_real = w.shape[:2]
    ntaps = w.shape[2] * w.shape[3]
    if w_fwd is not None:
        assert w_fwd.numel() == ntaps * cout * (cin_real + hole[1]) * (2 if split else 1)
    with _Timed('pack_conv_weight', float(w.numel() * 8)):
        _lib.check(lib.rpnet_pack_conv_weight_split(_ptr(w), cout, cin_real, ntaps, hole[0], hole[1], _ptr(w_fwd), int(split),
                                                    _ptr(w_dgrad), _stream()), 'rpnet_pack_conv_weight_split')


class _PackDesc(ctypes.Structure):          # rpnet_pack_desc of include/rpnet_b200.h
    _fields_ = [('w', ctypes.c_void_p), ('w_fwd_f16', ctypes.c_void_p), ('w_dgrad_bf16', ctypes.c_void_p), ('cout', ctypes.c_int),
                ('cin_real', ctypes.c_int), ('ntaps', ctypes.c_int), ('hole_start', ctypes.c_int), ('hole_len', ctypes.c_int),
                ('split', ctypes.c_int)]


def pack_conv_weights(layers):
    """layers: list of (w fp32 [cout, cin_real, kh, kw], w_fwd fp16 or None, w_dgrad bf16 or None, hole (start, len), split) -> every
    pack in one launch (rpnet_pack_conv_weights).  Returns the descriptor array (cache it: the pointers do not change)."""
    arr = (_PackDesc * len(layers))()
    nbytes = 0
    for d, (w, wf, wd, hole, split) in zip(arr, layers):
        _req(w, torch.float32, 'w')
        d.w, d.w_fwd_f16, d.w_dgrad_bf16 = w.data_ptr(), (wf.data_ptr() if wf is not None else None), (wd.data_ptr() if wd is not None else None)
        d.cout, d.cin_real, d.ntaps = w.shape[0], w.shape[1], w.shape[2] * w.shape[3]
        d.hole_start, d.hole_len, d.split = int(hole[0]), int(hole[1]), int(split)
        nbytes += w.numel() * 8
    return arr, nbytes


def run_pack_conv_weights(arr, nbytes):
    lib = _lib.load()
    with _Timed('pack_conv_weight', float(nbytes)):
        _lib.check(lib.rpnet_pack_conv_weights(ctypes.cast(arr, ctypes.c_void_p), len(arr), _stream()), 'rpnet_pack_conv_weights')


def conv_bnstats(src0, wpack, taps, ones, zeros, z, group_start, sums, src1=None):
    """Train-mode conv (no bias) -> z fp16 NHWC + BatchNorm statistics sums[g][cout][2] in the same launch."""
    lib = _lib.load()
    _req(src0, torch.float16, 'src0'); _req(wpack, torch.float16, 'wpack'); _req(z, torch.float16, 'z'); _req(sums, torch.float64, 'sums')
    n, h, w, c0 = src0.shape
    c1 = 0
    if src1 is not None:
        _req(src1, torch.float16, 'src1')
        assert src1.shape[:3] == src0.shape[:3]
        c1 = src1.shape[3]
    ntaps, cout, cin = wpack.shape
    if cin != c0 + c1 or ntaps != len(taps) or tuple(z.shape) != (n, h, w, cout):
        raise _lib.RpnetError('conv_bnstats: weight pack %s / output %s do not match the sources (%d + %d channels, %d taps)'
                              % (tuple(wpack.shape), tuple(z.shape), c0, c1, len(taps)))
    dy, dx = _taps(taps)
    gs, g = _groups(group_start)
    assert sums.numel() >= g * cout * 2
    with _Timed('conv_igemm', 2.0 * n * h * w * cout * cin * ntaps):
        rc = lib.rpnet_conv_bnstats_f16(_ptr(src0), c0, _ptr(src1), c1, n, h, w, _ptr(wpack), ntaps, dy, dx, cout, _ptr(ones), _ptr(zeros),
                                        _ptr(z), gs, g, _ptr(sums), _stream())
    _lib.check(rc, 'rpnet_conv_bnstats_f16')


def bn_stats(z, group_start, sums, z_lo=None):
    lib = _lib.load()
    _req(z, torch.float16, 'z'); _req(sums, torch.float64, 'sums')
    n, h, w, c = z.shape
    gs, g = _groups(group_start)
    assert sums.numel() >= g * c * 2
    if z_lo is not None:
        _req(z_lo, torch.float16, 'z_lo')
        assert z_lo.shape == z.shape
    with _Timed('bn_stats', float(z.numel() * (2 if z_lo is None else 4))):
        _lib.check(lib.rpnet_bn_stats_split_f16(_ptr(z), _ptr(z_lo), n, h, w, c, gs, g, _ptr(sums), _stream()), 'rpnet_bn_stats_split_f16')


def bn_finalize(sums, group_start, c, hw, gamma, beta, conv_bias, running_mean, running_var, nbt, stats, eps=1e-5, momentum=0.1):
    lib = _lib.load()
    gs, g = _groups(group_start)
    assert stats.numel() >= g * c * 4 and stats.dtype == torch.float32 and sums.dtype == torch.float64
    with _Timed('bn_finalize', float(g * c * 24)):
        _lib.check(lib.rpnet_bn_finalize_f32(_ptr(sums), gs, g, c, hw, _ptr(gamma), _ptr(beta), _ptr(conv_bias), float(eps),
                                             float(momentum), _ptr(running_mean), _ptr(running_var), _ptr(nbt), _ptr(stats), _stream()),
                   'rpnet_bn_finalize_f32')


def bn_apply(z, stats, group_start, relu=True, y=None, y_pool=None, y_f32=None, z_lo=None, y_lo=None, y_pool_lo=None, res=None,
             res_lo=None):
    """z_lo: fp16 residual plane of z (z = z + z_lo); y_lo / y_pool_lo: lo planes of the outputs — fp16 residual planes, or c8 planes
    (uint8, last dim doubled; include/rpnet_b200.h "fp8 corrections")."""
    lib = _lib.load()
    _req(z, torch.float16, 'z')
    n, h, w, c = z.shape
    gs, g = _groups(group_start)
    _lo_fmt([(z_lo, z, 'z_lo')])
    if z_lo is not None and z_lo.dtype != torch.float16:
        raise _lib.RpnetError('bn_apply: z_lo is the fp16 residual plane of the pre-BatchNorm conv output')
    lo_fmt = _lo_fmt([(y_lo, y, 'y_lo'), (y_pool_lo, y_pool, 'y_pool_lo'), (res_lo, res, 'res_lo')])
    if res is not None:                      # BasicBlock: y = relu(bn(z) + identity)
        _req(res, torch.float16, 'res')
        assert res.shape == z.shape and y_pool is None
    nb = sum(t.numel() * t.element_size() for t in (z, z_lo, y, y_lo, y_pool, y_pool_lo, y_f32, res, res_lo) if t is not None)
    with _Timed('bn_apply', float(nb)):
        _lib.check(lib.rpnet_bn_apply_res_f16(_ptr(z), _ptr(z_lo), _ptr(stats), n, h, w, c, gs, g, int(bool(relu)), _ptr(res), _ptr(res_lo),
                                              _ptr(y), _ptr(y_lo), _ptr(y_pool), _ptr(y_pool_lo), _ptr(y_f32), int(lo_fmt), _stream()),
                   'rpnet_bn_apply_res_f16')


def bn_bwd(z, stats, group_start, dz, scratch, relu=True, direct=None, d_off=0, pooled=None, p_off=0, up=None, u_off=0,
           dgamma=None, dbeta=None):
    """direct: bf16/fp32 NHWC tensor [n,h,w,ld]; pooled: bf16 [n,h/2,w/2,ld]; up: bf16 [n,2h,2w,ld]."""
    lib = _lib.load()
    _req(z, torch.float16, 'z'); _req(dz, bf16, 'dz')
    n, h, w, c = z.shape
    gs, g = _groups(group_start)
    assert scratch.dtype == torch.float32 and scratch.numel() >= g * c * 6 and scratch.data_ptr() % 8 == 0
    d_ld = direct.shape[-1] if direct is not None else 0
    p_ld = pooled.shape[-1] if pooled is not None else 0
    u_ld = up.shape[-1] if up is not None else 0
    if direct is not None:
        assert direct.dtype in (bf16, torch.float32) and direct.is_contiguous() and direct.numel() == n * h * w * d_ld
    if pooled is not None:
        assert pooled.dtype == bf16 and pooled.is_contiguous() and pooled.numel() == n * (h // 2) * (w // 2) * p_ld
    if up is not None:
        assert up.dtype == bf16 and up.is_contiguous() and up.numel() == n * 4 * h * w * u_ld
    with _Timed('bn_bwd', float(z.numel() * 2 * 5), n=3):
        rc = lib.rpnet_bn_bwd(_ptr(z), _ptr(stats), n, h, w, c, gs, g, int(bool(relu)), _ptr(direct), d_ld, d_off,
                              int(direct is not None and direct.dtype == torch.float32), _ptr(pooled), p_ld, p_off, _ptr(up), u_ld,
                              u_off, _ptr(dgamma), _ptr(dbeta), _ptr(scratch), _ptr(dz), _stream())
    _lib.check(rc, 'rpnet_bn_bwd')


def upsample2x(x, y):
    lib = _lib.load()
    _req(x, torch.float16, 'x'); _req(y, torch.float16, 'y')
    n, h, w, c = x.shape
    assert tuple(y.shape) == (n, 2 * h, 2 * w, c)
    with _Timed('upsample2x', float(x.numel() * 2 + y.numel() * 2)):
        _lib.check(lib.rpnet_upsample2x_f16(_ptr(x), _ptr(y), n, h, w, c, _stream()), 'rpnet_upsample2x_f16')


def premask_bwd(dxfg, dxbg, mask, dx, iters=1):
    lib = _lib.load()
    _req(dxfg, bf16, 'dxfg'); _req(dxbg, bf16, 'dxbg'); _req(mask, torch.float32, 'mask'); _req(dx, bf16, 'dx')
    c = dx.shape[-1]
    pixels = dx.numel() // c
    assert dxfg.numel() == iters * pixels * c and dxbg.numel() == dxfg.numel() and mask.numel() == iters * pixels
    with _Timed('premask_bwd', float(dxfg.numel() * 4 + dx.numel() * 2)):
        _lib.check(lib.rpnet_premask_bwd_bf16(_ptr(dxfg), _ptr(dxbg), _ptr(mask), iters, pixels, c, _ptr(dx), _stream()), 'rpnet_premask_bwd_bf16')


def premask_mask_bwd(dxfg, dxbg, x, dmask):
    """dmask fp32 [n, h, w] = sum_c (dxfg - dxbg) * x  (gradient of the pre-mask w.r.t. a soft recurrent mask)."""
    lib = _lib.load()
    _req(dxfg, bf16, 'dxfg'); _req(dxbg, bf16, 'dxbg'); _req(x, torch.float16, 'x'); _req(dmask, torch.float32, 'dmask')
    c = x.shape[-1]
    pixels = x.numel() // c
    assert dxfg.shape == x.shape and dxbg.shape == x.shape and dmask.numel() == pixels
    with _Timed('premask_mask_bwd', float(x.numel() * 6)):
        _lib.check(lib.rpnet_premask_mask_bwd(_ptr(dxfg), _ptr(dxbg), _ptr(x), pixels, c, _ptr(dmask), _stream()), 'rpnet_premask_mask_bwd')


def soft_mask_bwd(logits, dmask, scale, dlogits):
    """dlogits [b, P, H, W] += the gradient of avg_pool2d(p_fg(logits), scale) against dmask [b, H/scale, W/scale]."""
    lib = _lib.load()
    _req(logits, torch.float32, 'logits'); _req(dmask, torch.float32, 'dmask'); _req(dlogits, torch.float32, 'dlogits')
    b, p, H, W = logits.shape
    assert dlogits.shape == logits.shape and dmask.numel() == b * (H // scale) * (W // scale)
    with _Timed('soft_mask_bwd', float(logits.numel() * 12)):
        _lib.check(lib.rpnet_soft_mask_bwd_f32(_ptr(logits), _ptr(dmask), b, p, H // scale, W // scale, scale, _ptr(dlogits), _stream()),
                   'rpnet_soft_mask_bwd_f32')


def local_corr_bwd(f1, f2, dq, add_off, radius, df1, df2, workspace=None):
    """workspace: optional uint8/any tensor of >= local_corr_bwd_workspace_bytes(...) bytes (enables the tensor-core path)."""
    lib = _lib.load()
    _req(f1, torch.float16, 'f1'); _req(f2, torch.float16, 'f2'); _req(dq, bf16, 'dq'); _req(df1, bf16, 'df1'); _req(df2, bf16, 'df2')
    n, h, w, c = f1.shape
    assert f2.shape == f1.shape and df1.shape == f1.shape and df2.shape == f1.shape and tuple(dq.shape[:3]) == (n, h, w)
    with _Timed('local_corr_bwd', float(f1.numel() * 8 + dq.numel() * 2), n=2):
        _lib.check(lib.rpnet_local_corr_bwd(_ptr(f1), _ptr(f2), _ptr(dq), dq.shape[3], add_off, _ptr(df1), _ptr(df2), n, h, w, c, radius,
                                            _ptr(workspace), 0 if workspace is None else workspace.numel() * workspace.element_size(),
                                            _stream()), 'rpnet_local_corr_bwd')


def local_corr_bwd_workspace_bytes(n, h, w, radius):
    return int(_lib.load().rpnet_local_corr_bwd_workspace_bytes(n, h, w, radius))


def cos_sim_bwd(feat, protos, dpred, dfeat, dprotos=None, scaler=20.0, accumulate=False):
    lib = _lib.load()
    _req(feat, torch.float32, 'feat'); _req(protos, torch.float32, 'protos'); _req(dpred, torch.float32, 'dpred'); _req(dfeat, torch.float32, 'dfeat')
    n, h, w, c = feat.shape
    sets, p = protos.shape[0], protos.shape[1]
    assert tuple(dpred.shape) == (n, p, h, w) and dfeat.shape == feat.shape
    acc = _scratch64(sets * p * c, feat.device, 'cos_sim_bwd') if dprotos is not None else None
    with _Timed('cos_sim_bwd', float(feat.numel() * 8 + dpred.numel() * 4), n=1 if dprotos is None else 2):
        _lib.check(lib.rpnet_cos_sim_bwd_f32(_ptr(feat), _ptr(protos), _ptr(dpred), n, h * w, c, p, sets, float(scaler), _ptr(dfeat),
                                             int(bool(accumulate)), _ptr(dprotos), _ptr(acc), _stream()), 'rpnet_cos_sim_bwd_f32')


def bilinear_adjoint(x, out, sums=None):
    """x fp32 [n,H,W] -> out fp32 [n,h,w] = U^T x; sums [n] (optional) = x.sum((1,2))."""
    lib = _lib.load()
    _req(x, torch.float32, 'x'); _req(out, torch.float32, 'out')
    n, H, W = x.shape
    assert out.shape[0] == n
    with _Timed('bilinear_adjoint', float(x.numel() * 4 + out.numel() * 4)):
        _lib.check(lib.rpnet_bilinear_adjoint_f32(_ptr(x), _ptr(out), _ptr(sums), n, H, W, out.shape[1], out.shape[2], _stream()),
                   'rpnet_bilinear_adjoint_f32')


def weighted_pool(feat, wmap0, wmap1, msum0, msum1, out):
    lib = _lib.load()
    _req(feat, torch.float32, 'feat'); _req(out, torch.float32, 'out')
    n, h, w, c = feat.shape
    assert wmap0.numel() == n * h * w and wmap1.numel() == n * h * w and tuple(out.shape) == (n, 2, c)
    with _Timed('weighted_pool', float(feat.numel() * 8)):
        _lib.check(lib.rpnet_weighted_pool_f32(_ptr(feat), _ptr(wmap0), _ptr(wmap1), _ptr(msum0), _ptr(msum1), _ptr(out), n, h * w, c,
                                               _stream()), 'rpnet_weighted_pool_f32')


def weighted_pool_bwd(dout, wmap0, wmap1, msum0, msum1, dfeat, accumulate=False):
    lib = _lib.load()
    _req(dout, torch.float32, 'dout'); _req(dfeat, torch.float32, 'dfeat')
    n, h, w, c = dfeat.shape
    with _Timed('weighted_pool_bwd', float(dfeat.numel() * 4)):
        _lib.check(lib.rpnet_weighted_pool_bwd_f32(_ptr(dout), _ptr(wmap0), _ptr(wmap1), _ptr(msum0), _ptr(msum1), _ptr(dfeat),
                                                   int(bool(accumulate)), n, h * w, c, _stream()), 'rpnet_weighted_pool_bwd_f32')


def proto_finalize_bwd(dprotos, draw):
    lib = _lib.load()
    _req(dprotos, torch.float32, 'dprotos'); _req(draw, torch.float32, 'draw')
    ways, shots, batch, two, c = draw.shape
    assert tuple(dprotos.shape) == (batch, 1 + ways, c)
    with _Timed('proto_finalize_bwd', float(draw.numel() * 4)):
        _lib.check(lib.rpnet_proto_finalize_bwd_f32(_ptr(dprotos), _ptr(draw), ways, shots, batch, c, _stream()), 'rpnet_proto_finalize_bwd_f32')


def dice_ce(logits, labels, sums, loss, dlogits=None, grad_scale=1.0, eps=1e-7):
    """logits fp32 [G,B,P,H,W]; labels int64 [B,H,W]; loss fp32 [G]; dlogits like logits (optional); sums: fp64 scratch."""
    lib = _lib.load()
    _req(logits, torch.float32, 'logits'); _req(labels, torch.int64, 'labels'); _req(loss, torch.float32, 'loss')
    _req(sums, torch.float64, 'sums')
    g, b, p, h, w = logits.shape
    assert tuple(labels.shape) == (b, h, w) and sums.numel() >= g * (2 * p + 1) and loss.numel() >= g
    with _Timed('dice_ce', float(logits.numel() * (12 if dlogits is not None else 4)), n=2):
        _lib.check(lib.rpnet_dice_ce_f32(_ptr(logits), _ptr(labels), g, b, p, h * w, float(eps), float(grad_scale), _ptr(sums),
                                         _ptr(dlogits), _ptr(loss), _stream()), 'rpnet_dice_ce_f32')


def class_pool(feat, pred, qproto, counts, amax):
    lib = _lib.load()
    _req(feat, torch.float32, 'feat'); _req(pred, torch.float32, 'pred'); _req(amax, torch.int32, 'amax')
    b, h, w, c = feat.shape
    p = pred.shape[1]
    assert tuple(qproto.shape) == (b, p, c) and counts.numel() == b * p and amax.numel() == b * h * w
    with _Timed('class_pool', float(feat.numel() * 4)):
        _lib.check(lib.rpnet_class_pool_f32(_ptr(feat), _ptr(pred), b, h * w, c, p, _ptr(qproto), _ptr(counts), _ptr(amax), _stream()),
                   'rpnet_class_pool_f32')


def class_pool_bwd(dqproto, counts, amax, dfeat):
    lib = _lib.load()
    _req(dqproto, torch.float32, 'dqproto'); _req(dfeat, torch.float32, 'dfeat')
    b, h, w, c = dfeat.shape
    with _Timed('class_pool_bwd', float(dfeat.numel() * 8)):
        _lib.check(lib.rpnet_class_pool_bwd_f32(_ptr(dqproto), _ptr(counts), _ptr(amax), b, h * w, c, dqproto.shape[1], _ptr(dfeat),
                                                _stream()), 'rpnet_class_pool_bwd_f32')


def align_gather(qproto, counts, ways, shots, scaler, protos_s, weight):
    lib = _lib.load()
    b = qproto.shape[0]
    assert protos_s.numel() == ways * shots * b * 128 and weight.numel() == ways * shots * b
    with _Timed('align_gather', float(protos_s.numel() * 8)):
        _lib.check(lib.rpnet_align_gather_f32(_ptr(qproto), _ptr(counts), ways, shots, b, float(scaler), _ptr(protos_s), _ptr(weight),
                                              _stream()), 'rpnet_align_gather_f32')


def align_scatter(dprotos_s, ways, shots, dqproto):
    lib = _lib.load()
    b = dqproto.shape[0]
    with _Timed('align_scatter', float(dprotos_s.numel() * 4)):
        _lib.check(lib.rpnet_align_scatter_f32(_ptr(dprotos_s), ways, shots, b, _ptr(dqproto), _stream()), 'rpnet_align_scatter_f32')


def ce_mask(logits, fore, back, weight, sums, loss, dlogits=None, grad_scale=1.0):
    lib = _lib.load()
    _req(logits, torch.float32, 'logits'); _req(fore, torch.float32, 'fore'); _req(back, torch.float32, 'back')
    _req(sums, torch.float64, 'sums')
    n, two, h, w = logits.shape
    assert two == 2 and fore.numel() == n * h * w and back.numel() == n * h * w and weight.numel() == n and sums.numel() >= 2 * n
    with _Timed('ce_mask', float(logits.numel() * (12 if dlogits is not None else 4)), n=2):
        _lib.check(lib.rpnet_ce_mask_f32(_ptr(logits), _ptr(fore), _ptr(back), _ptr(weight), n, h * w, float(grad_scale), _ptr(sums),
                                         _ptr(dlogits), _ptr(loss), _stream()), 'rpnet_ce_mask_f32')


def bilinear_up(x, out):
    lib = _lib.load()
    _req(x, torch.float32, 'x'); _req(out, torch.float32, 'out')
    n, h, w = x.shape
    assert out.shape[0] == n
    with _Timed('bilinear_up', float(x.numel() * 4 + out.numel() * 4)):
        _lib.check(lib.rpnet_bilinear_up_f32(_ptr(x), _ptr(out), n, h, w, out.shape[1], out.shape[2], _stream()), 'rpnet_bilinear_up_f32')


def adam(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    lib = _lib.load()
    for t, nm in ((param, 'param'), (grad, 'grad'), (exp_avg, 'exp_avg'), (exp_avg_sq, 'exp_avg_sq')):
        _req(t, torch.float32, nm)
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    with _Timed('adam', float(n * 28)):
        _lib.check(lib.rpnet_adam_f32(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), n, float(lr), float(betas[0]),
                                      float(betas[1]), float(eps), float(weight_decay), int(step), float(grad_scale), _stream()),
                   'rpnet_adam_f32')


# =====================================================================================================
# "next" row N1: batched affine registration (include/rpnet_b200.h, last section)
# =====================================================================================================
def affine_register(moving, fixed, theta, iters=50, lr=0.01, betas=(0.9, 0.999), eps=1e-8, loss_curve=None):
    """moving / fixed fp32 [n, h, w]; theta fp32 [n, 2, 3] (out); loss_curve fp32 [n, iters] (optional)."""
    lib = _lib.load()
    _req(moving, torch.float32, 'moving'); _req(fixed, torch.float32, 'fixed'); _req(theta, torch.float32, 'theta')
    n, h, w = moving.shape
    assert fixed.shape == moving.shape and tuple(theta.shape) == (n, 2, 3)
    if loss_curve is not None:
        _req(loss_curve, torch.float32, 'loss_curve')
        assert tuple(loss_curve.shape) == (n, iters)
    with _Timed('affine_register', float(moving.numel() * 8 * iters)):
        _lib.check(lib.rpnet_affine_register_f32(_ptr(moving), _ptr(fixed), n, h, w, int(iters), float(lr), float(betas[0]), float(betas[1]),
                                                 float(eps), _ptr(theta), _ptr(loss_curve), _stream()), 'rpnet_affine_register_f32')


def affine_warp(x, theta, out):
    """x, out fp32 [n, c, h, w]; theta fp32 [n, 2, 3]: F.grid_sample(x, F.affine_grid(theta, x.size()))."""
    lib = _lib.load()
    _req(x, torch.float32, 'x'); _req(theta, torch.float32, 'theta'); _req(out, torch.float32, 'out')
    n, c, h, w = x.shape
    assert out.shape == x.shape and tuple(theta.shape) == (n, 2, 3)
    with _Timed('affine_warp', float(x.numel() * 8)):
        _lib.check(lib.rpnet_affine_warp_f32(_ptr(x), _ptr(theta), _ptr(out), n, c, h, w, _stream()), 'rpnet_affine_warp_f32')


def ncc(moving, fixed):
    """NCC(moving, fixed) of net/registration.py:157-160 as a device scalar tensor [1]."""
    lib = _lib.load()
    _req(moving, torch.float32, 'moving'); _req(fixed, torch.float32, 'fixed')
    assert moving.numel() == fixed.numel()
    scratch = torch.empty(5, dtype=torch.float64, device=moving.device)
    out = torch.empty(1, dtype=torch.float32, device=moving.device)
    with _Timed('ncc', float(moving.numel() * 8), n=2):
        _lib.check(lib.rpnet_ncc_f32(_ptr(moving), _ptr(fixed), moving.numel(), _ptr(scratch), _ptr(out), _stream()), 'rpnet_ncc_f32')
    return out


# =====================================================================================================
# up_conv in sub-pixel form (train path)
# =====================================================================================================
def upconv_fusable(h, w):
    """True when a 128-pixel tile of the low-resolution map never spans two images (the BatchNorm statistics of the phase
    convs can then be accumulated in their epilogues)."""
    def p2(v):
        r = 1
        while r * 2 <= v:
            r *= 2
        return r
    bw = p2(min(w, 16))
    bh = p2(min(h, 128 // bw))
    return bw * bh == 128


def pack_upconv_weight(w, wf, w16, split=False):
    """w fp32 [cout, cin, 3, 3] -> wf fp16 [4, 4, cout, cin] (phase forward; split: [4, 4, cout, 2 * cin] = Wh | Wl) and
    w16 bf16 [16, cin, cout] (data gradient)."""
    lib = _lib.load()
    _req(w, torch.float32, 'w'); _req(wf, torch.float16, 'wf'); _req(w16, bf16, 'w16')
    cout, cin = w.shape[:2]
    assert tuple(wf.shape) == (4, 4, cout, cin * (2 if split else 1)) and tuple(w16.shape) == (16, cin, cout)
    with _Timed('pack_conv_weight', float(w.numel() * 12)):
        _lib.check(lib.rpnet_pack_upconv_weight_split(_ptr(w), cout, cin, _ptr(wf), int(split), _ptr(w16), _stream()),
                   'rpnet_pack_upconv_weight_split')


def upconv_fwd_bnstats(x_low, wf, ones, zeros, z, group_start, sums):
    """z [n, 2h, 2w, cout] = conv3x3(upsample2x(x_low)) as four phase convs + BatchNorm statistics of z."""
    lib = _lib.load()
    _req(x_low, torch.float16, 'x_low'); _req(wf, torch.float16, 'wf'); _req(z, torch.float16, 'z'); _req(sums, torch.float64, 'sums')
    n, h, w, cin = x_low.shape
    cout = wf.shape[2]
    assert tuple(z.shape) == (n, 2 * h, 2 * w, cout) and tuple(wf.shape) == (4, 4, cout, cin)
    gs, g = _groups(group_start)
    for ph in range(4):
        with _Timed('conv_igemm', 2.0 * n * h * w * cout * cin * 4):
            rc = lib.rpnet_upconv_phase_bnstats_f16(_ptr(x_low), cin, n, h, w, _ptr(wf[ph]), ph >> 1, ph & 1, cout, _ptr(ones), _ptr(zeros),
                                                    _ptr(z), gs, g, _ptr(sums), int(ph > 0), _stream())
        _lib.check(rc, 'rpnet_upconv_phase_bnstats_f16')


def upconv_dgrad(dz, w16, out, out_coff=0):
    """dx_low bf16 [n, h, w, >= cin] from dz bf16 [n, 2h, 2w, cout] (4x4 stride-2 conv over the four parity views of dz)."""
    lib = _lib.load()
    _req(dz, bf16, 'dz'); _req(w16, bf16, 'w16'); _req(out, bf16, 'out')
    n, H, W, cout = dz.shape
    h, w = H // 2, W // 2
    cin = w16.shape[1]
    assert tuple(w16.shape) == (16, cin, cout) and tuple(out.shape[:3]) == (n, h, w)
    one, zero = _const_vec(cin, 1.0, dz.device), _const_vec(cin, 0.0, dz.device)
    with _Timed('conv_igemm', 2.0 * n * h * w * cout * cin * 16):
        rc = lib.rpnet_upconv_dgrad_bf16(_ptr(dz), cout, n, h, w, _ptr(w16), cin, _ptr(out), out.shape[3], out_coff, _ptr(one), _ptr(zero),
                                         _stream())
    _lib.check(rc, 'rpnet_upconv_dgrad_bf16')


def upconv_wgrad_workspace_bytes(cin, n, h, w, cout):
    r = _lib.load().rpnet_upconv_wgrad_workspace_bytes(cin, n, h, w, cout)
    if r < 0:
        _lib.check(int(r), 'rpnet_upconv_wgrad_workspace_bytes')
    return int(r)


def upconv_wgrad(x_low, dz, grad, workspace, accumulate=False):
    """grad fp32 [cout, cin, 3, 3] from x_low fp16/bf16 [n, h, w, cin] and dz bf16 [n, 2h, 2w, cout]."""
    lib = _lib.load()
    _req(dz, bf16, 'dz'); _req(grad, torch.float32, 'grad')
    n, h, w, cin = x_low.shape
    cout = dz.shape[3]
    assert tuple(dz.shape) == (n, 2 * h, 2 * w, cout) and grad.numel() == cout * cin * 9
    with _Timed('conv_wgrad', 2.0 * n * h * w * cout * cin * 16, n=9):
        rc = lib.rpnet_upconv_wgrad(_ptr(x_low), int(x_low.dtype == bf16), _ptr(dz), n, h, w, cin, cout, _ptr(grad), int(bool(accumulate)),
                                    _ptr(workspace), workspace.numel() * workspace.element_size(), _stream())
    _lib.check(rc, 'rpnet_upconv_wgrad')


# =====================================================================================================
# "next" row N3: ResNet18 backbone pieces
# =====================================================================================================
def conv_res(src, wpack, taps, scale, shift, out, res=None, relu=True):
    """out = act(scale * conv(src) + shift (+ res)); src / res / out fp16 NHWC."""
    lib = _lib.load()
    _req(src, torch.float16, 'src'); _req(wpack, torch.float16, 'wpack'); _req(out, torch.float16, 'out')
    n, h, w, cin = src.shape
    ntaps, cout, cin2 = wpack.shape
    if cin2 != cin or ntaps != len(taps) or tuple(out.shape) != (n, h, w, cout):
        raise _lib.RpnetError('conv_res: shapes do not match')
    if res is not None:
        _req(res, torch.float16, 'res')
        assert res.shape == out.shape
    dy, dx = _taps(taps)
    with _Timed('conv_igemm', 2.0 * n * h * w * cout * cin * ntaps):
        rc = lib.rpnet_conv_res_f16(_ptr(src), cin, n, h, w, _ptr(wpack), ntaps, dy, dx, cout, _ptr(scale), _ptr(shift), _ptr(res),
                                    int(bool(relu)), _ptr(out), _stream())
    _lib.check(rc, 'rpnet_conv_res_f16')


def conv7x7s2_stem(img, weight, scale, shift, out, relu=True, out_lo=None):
    """img fp32 NCHW [n, 3, H, W]; weight fp32 [64, 3, 7, 7]; out fp16 NHWC [n, (H-1)//2+1, (W-1)//2+1, 64] (+ residual plane)."""
    lib = _lib.load()
    _req(img, torch.float32, 'img'); _req(weight, torch.float32, 'weight'); _req(out, torch.float16, 'out')
    n, c, h, w = img.shape
    assert c == 3 and tuple(weight.shape) == (64, 3, 7, 7) and tuple(out.shape) == (n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, 64)
    with _Timed('conv7x7s2_stem', float(img.numel() * 4 + out.numel() * 2)):
        lo_fmt = _lo_fmt([(out_lo, out, 'out_lo')])
        _lib.check(lib.rpnet_conv7x7s2_stem_split_f16(_ptr(img), n, h, w, _ptr(weight), _ptr(scale), _ptr(shift), int(bool(relu)),
                                                      _ptr(out), _ptr(out_lo), int(lo_fmt), _stream()), 'rpnet_conv7x7s2_stem_split_f16')


# =====================================================================================================
# "next" row N1, deformable half: batched demons registration (include/rpnet_b200.h, last section)
# =====================================================================================================
def demons_register(moving, fixed, gauss, iters=50, lr=0.01, betas=(0.9, 0.999), eps=1e-8, scaling=10, loss_curve=None):
    """moving / fixed fp32 [n, h, w] in [0, 1]; gauss: CPU fp32 tensor [kh, kw] (the regulariser's kernel).
    Returns (flow [n, 2, h, w], disp = exp(flow) [n, 2, h, w])."""
    lib = _lib.load()
    _req(moving, torch.float32, 'moving'); _req(fixed, torch.float32, 'fixed')
    n, h, w = moving.shape
    assert fixed.shape == moving.shape and gauss.dim() == 2 and not gauss.is_cuda
    flow = torch.empty(n, 2, h, w, dtype=torch.float32, device=moving.device)
    disp = torch.empty_like(flow)
    if loss_curve is not None:
        _req(loss_curve, torch.float32, 'loss_curve')
        assert tuple(loss_curve.shape) == (n, iters)
    nb = int(lib.rpnet_demons_workspace_bytes(n, h, w, scaling))
    if nb < 0:
        _lib.check(nb, 'rpnet_demons_workspace_bytes')
    ws = torch.empty(nb // 4, dtype=torch.float32, device=moving.device)
    g = gauss.float().contiguous()
    garr = (ctypes.c_float * g.numel())(*g.flatten().tolist())
    with _Timed('demons_register', float(moving.numel() * 4 * 60 * max(iters, 1))):
        _lib.check(lib.rpnet_demons_register_f32(_ptr(moving), _ptr(fixed), n, h, w, int(iters), float(lr), float(betas[0]), float(betas[1]),
                                                 float(eps), int(scaling), ctypes.cast(garr, ctypes.c_void_p), g.shape[0], g.shape[1],
                                                 _ptr(flow), _ptr(disp), _ptr(loss_curve), _ptr(ws), nb, _stream()), 'rpnet_demons_register_f32')
    return flow, disp


def demons_warp(x, disp, out):
    """x, out fp32 [n, c, h, w]; disp fp32 [n, 2, h, w]: F.grid_sample(x, compute_grid + disp) (DemonsRegistration.forward)."""
    lib = _lib.load()
    _req(x, torch.float32, 'x'); _req(disp, torch.float32, 'disp'); _req(out, torch.float32, 'out')
    n, c, h, w = x.shape
    assert out.shape == x.shape and tuple(disp.shape) == (n, 2, h, w)
    with _Timed('demons_warp', float(x.numel() * 8 + disp.numel() * 4)):
        _lib.check(lib.rpnet_demons_warp_f32(_ptr(x), _ptr(disp), _ptr(out), n, c, h, w, _stream()), 'rpnet_demons_warp_f32')
