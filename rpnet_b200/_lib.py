"""ctypes binding of librpnet_sm100.so (the C ABI in include/rpnet_b200.h).  Fails loudly when absent."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'librpnet_sm100.so')
ABI_VERSION = 1

_c_int, _c_ll, _c_f, _vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
_ip = ctypes.POINTER(ctypes.c_int)

# name -> argtypes; mirrors include/rpnet_b200.h (tests check the two stay in sync)
SIGNATURES = {
    'rpnet_abi_version': [],
    'rpnet_conv_igemm_f16': [_vp, _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _vp, _c_int, _ip, _ip, _c_int, _vp, _vp,
                             _c_int, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp],
    'rpnet_conv3x3_first_f16': [_vp, _c_int, _c_int, _c_int, _c_int, _vp, _vp, _vp, _c_int, _vp, _vp],
    'rpnet_avgpool_mask_f32': [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp],
    'rpnet_premask_f16': [_vp, _vp, _vp, _vp, _c_ll, _c_int, _vp],
    'rpnet_local_corr_f16': [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp],
    'rpnet_masked_avg_pool_f32': [_vp, _vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp],
    'rpnet_proto_finalize_f32': [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _vp],
    'rpnet_cos_sim_f32': [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_f, _vp],
    'rpnet_maxpool_f16': [_vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp],
    'rpnet_upsample_tail_f32': [_vp, _vp, _vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _vp],
}

_lib = None


class RpnetError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises RpnetError if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RpnetError('%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                         '(nvcc, sm_100a). rpnet_b200 has no CPU or PyTorch fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.rpnet_last_error.restype = ctypes.c_char_p
    lib.rpnet_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library out of sync
        fn.restype = _c_int
        fn.argtypes = argtypes
    if lib.rpnet_abi_version() != ABI_VERSION:
        raise RpnetError('ABI mismatch: library %d, binding %d' % (lib.rpnet_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().rpnet_last_error().decode('utf-8', 'replace')
        raise RpnetError('%s failed (%d): %s' % (what, rc, msg))
