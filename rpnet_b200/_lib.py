"""ctypes binding of librpnet_sm100.so (the C ABI in include/rpnet_b200.h).  Fails loudly when absent."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'librpnet_sm100.so')
ABI_VERSION = 8

_c_int, _c_ll, _c_f, _vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
_ip = ctypes.POINTER(ctypes.c_int)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'rpnet_b200.h')


def _parse_header(path=HEADER_PATH):
    """include/rpnet_b200.h is the single source of truth for the ABI: name -> (restype, argtypes).
    Pointers map to void* (device pointers travel as integers) except `const int*` (HOST int arrays: tap lists,
    BatchNorm call-group boundaries)."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    sigs = {}
    for m in re.finditer(r'(long long|int|const char\s*\*)\s+(rpnet_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        args = [a.strip() for a in args.replace('\n', ' ').split(',')]
        argtypes = []
        for a in ([] if args == ['void'] else args):
            if '*' in a:
                argtypes.append(_ip if re.match(r'const\s+int\s*\*', a) else _vp)
            elif a.startswith('long long'):
                argtypes.append(_c_ll)
            elif a.startswith('float'):
                argtypes.append(_c_f)
            elif a.startswith('int'):
                argtypes.append(_c_int)
            else:
                raise RuntimeError('rpnet_b200.h: cannot map argument %r of %s' % (a, name))
        restype = ctypes.c_char_p if '*' in ret else (_c_ll if ret == 'long long' else _c_int)
        sigs[name] = (restype, argtypes)
    return sigs


_SIGS = _parse_header()
# name -> argtypes of every int-returning entry point (tests check the library exports each of them)
SIGNATURES = {k: v[1] for k, v in _SIGS.items() if k != 'rpnet_last_error'}

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
    for name, (restype, argtypes) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError here == header / library out of sync
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.rpnet_abi_version() != ABI_VERSION:
        raise RpnetError('ABI mismatch: library %d, binding %d' % (lib.rpnet_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().rpnet_last_error().decode('utf-8', 'replace')
        raise RpnetError('%s failed (%d): %s' % (what, rc, msg))
