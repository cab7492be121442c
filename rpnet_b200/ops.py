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
    return t


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


def conv3x3_first(img, weight, scale, shift, relu, out):
    lib = _lib.load()
    _req(img, torch.float32, 'img'); _req(weight, torch.float32, 'weight'); _req(out, torch.float16, 'out')
    n, cin, h, w = img.shape
    assert tuple(weight.shape) == (64, cin, 3, 3) and tuple(out.shape) == (n, h, w, 64)
    with _Timed('conv3x3_first', float(img.numel() * 4 + out.numel() * 2)):
        rc = lib.rpnet_conv3x3_first_f16(_ptr(img), n, cin, h, w, _ptr(weight), _ptr(_req(scale, torch.float32, 'scale')),
                                         _ptr(_req(shift, torch.float32, 'shift')), int(bool(relu)), _ptr(out), _stream())
    _lib.check(rc, 'rpnet_conv3x3_first_f16')


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


def maxpool(x, k, stride, pad, out):
    lib = _lib.load()
    _req(x, torch.float16, 'x'); _req(out, torch.float16, 'out')
    n, h, w, c = x.shape
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    assert tuple(out.shape) == (n, ho, wo, c)
    with _Timed('maxpool', float(x.numel() * 2 + out.numel() * 2)):
        _lib.check(lib.rpnet_maxpool_f16(_ptr(x), _ptr(out), n, h, w, c, k, stride, pad, _stream()), 'rpnet_maxpool_f16')
    