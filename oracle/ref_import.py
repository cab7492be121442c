"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by tests/golden/make_golden.py to produce the committed golden
vectors and by tests that are skipped when /root/reference is absent (the GPU box).
Recipe: SURVEY.md Appendix A — stub the heavy optional imports of utils/util.py:4,15-20
and load the reference's `net` package under the private name `_rpnet_ref_net` so that it
cannot shadow this repository's own `net` shim package.
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get('RPNET_REFERENCE_ROOT', '/root/reference')


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        m = _Stub(self.__name__ + '.' + k)
        setattr(self, k, m)
        return m


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'net', 'rp_net.py'))


def load():
    """Returns a namespace with the reference's RP_Net, Correlation, dice_ce, Encoder, U_Net, NCC, MSE."""
    if '_rpnet_ref' in sys.modules:
        return sys.modules['_rpnet_ref']
    for name in ['pydicom', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'SimpleITK', 'skimage',
                 'skimage.measure', 'nrrd', 'nibabel', 'torchviz']:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = _Stub(name)
                m.__spec__ = importlib.machinery.ModuleSpec(name, None)
                m.__path__ = []
                sys.modules[name] = m
    # Temporarily make `net` / `utils` resolve to the reference tree, import, then rename.
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ('net', 'utils', 'dataset') or k.startswith(('net.', 'utils.', 'dataset.'))}
    sys.path.insert(0, REF_ROOT)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            rp = importlib.import_module('net.rp_net')
            vgg = importlib.import_module('net.vgg')
            unet = importlib.import_module('net.unet')
            reg = importlib.import_module('net.registration')
            util = importlib.import_module('utils.util')
    finally:
        sys.path.remove(REF_ROOT)
        for k in list(sys.modules):
            if k in ('net', 'utils', 'dataset') or k.startswith(('net.', 'utils.', 'dataset.')):
                sys.modules['_rpnet_ref_' + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    ns = types.SimpleNamespace(RP_Net=rp.RP_Net, Correlation=rp.Correlation, dice_ce=rp.dice_ce,
                               dice_loss_softmax=rp.dice_loss_softmax, Encoder=vgg.Encoder, U_Net=unet.U_Net,
                               NCC=reg.NCC, MSE=reg.MSE, dice_score_seperate=util.dice_score_seperate,
                               normalize=util.normalize, rp_net=rp)
    sys.modules['_rpnet_ref'] = ns
    return ns
