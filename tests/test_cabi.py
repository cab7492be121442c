"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/rpnet_b200.h
declares (no compute calls — there is no GPU here), the ctypes binding matches the header, and the reference-shaped
Python surface (model_factory, state_dict keys, constructor/forward signatures) is intact."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    src = open(os.path.join(ROOT, 'include', 'rpnet_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    decls = {}
    for m in re.finditer(r'(?:long long|int|const char\s*\*)\s+(rpnet_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
        args = [a.strip() for a in m.group(2).replace('\n', ' ').split(',')]
        decls[m.group(1)] = [] if args == ['void'] else args
    return decls


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from rpnet_b200 import _lib
    return ctypes.CDLL(_lib.LIB_PATH)


def test_library_exports_every_declared_symbol(lib):
    decls = _header_decls()
    assert len(decls) >= 12
    for name in decls:
        assert hasattr(lib, name), 'include/rpnet_b200.h declares %s but the library does not export it' % name
    lib.rpnet_abi_version.restype = ctypes.c_int
    from rpnet_b200 import _lib
    assert lib.rpnet_abi_version() == _lib.ABI_VERSION
    lib.rpnet_last_error.restype = ctypes.c_char_p
    assert lib.rpnet_last_error() == b''


def test_ctypes_binding_matches_header():
    from rpnet_b200 import _lib
    decls = _header_decls()
    bound = dict(_lib.SIGNATURES)
    bound['rpnet_last_error'] = []
    assert set(bound) == set(decls), set(bound) ^ set(decls)
    for name, argtypes in bound.items():
        assert len(argtypes) == len(decls[name]), '%s: binding has %d args, header %d' % (name, len(argtypes), len(decls[name]))
        for ct, decl in zip(argtypes, decls[name]):
            is_ptr = '*' in decl
            assert is_ptr == (ct in (ctypes.c_void_p, ctypes.POINTER(ctypes.c_int))), (name, decl, ct)


def test_bad_arguments_return_error_codes_without_a_gpu(lib):
    """Argument validation happens before any CUDA call: exercised here without a device."""
    lib.rpnet_last_error.restype = ctypes.c_char_p
    rc = lib.rpnet_avgpool_mask_f32(None, None, 1, 8, 8, 4, None)
    assert rc == -2 and b'null pointer' in lib.rpnet_last_error()
    rc = lib.rpnet_local_corr_f16(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 8, 8, 48, 5, 128, None)
    assert rc == -2 and b'c % 32' in lib.rpnet_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from rpnet_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/librpnet_sm100.so')
    with pytest.raises(_lib.RpnetError, match='no CPU or PyTorch fallback'):
        _lib.load()


def _cfg(T):
    return dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
                n_iter_refinement=T, soft_mask=False, mask_refinement_correlation_radius=5)


def test_state_dict_surface():
    """147 state_dict keys with the reference's names and shapes."""
    from oracle import weights
    from net.model import model_factory
    sd = weights.unet_rpnet_state_dict(0)
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=_cfg(4))
    mine = net.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    assert all(mine[k].shape == sd[k].shape for k in sd)
    net.load_state_dict(sd)
    assert net.scale == 4 and net.num_iter == 4


def test_signatures_match_reference():
    """Constructor / forward argument names of the reference (net/rp_net.py:195,226; net/unet.py:393,435; net/vgg.py:16,36)."""
    from net.rp_net import RP_Net, Correlation, dice_ce
    from net.unet import U_Net
    from net.vgg import Encoder
    assert list(inspect.signature(RP_Net.__init__).parameters) == ['self', 'in_channels', 'pretrained_path', 'cfg', 'backbone_cfg']
    assert list(inspect.signature(RP_Net.forward).parameters) == [
        'self', 'supp_imgs', 'fore_mask', 'back_mask', 'qry_imgs', 'registration_field', 'grid', 'query_labels',
        'appr_query_labels']
    assert list(inspect.signature(U_Net.forward).parameters) == ['self', 'x', 'mask', 'do_last_conv']
    assert list(inspect.signature(Encoder.forward).parameters) == ['self', 'x', 'mask']
    assert list(inspect.signature(Correlation).parameters) == ['fmap1', 'fmap2', 'r']
    assert list(inspect.signature(dice_ce).parameters) == ['logits', 'true', 'eps']
    for m in ('calDist', 'getFeatures', 'getPrototype', 'alignLoss'):
        assert hasattr(RP_Net, m)


def test_cpu_tensors_are_rejected_not_silently_computed():
    """No CPU fallback: a forward on CPU tensors raises instead of running PyTorch ops."""
    from net.model import model_factory
    from rpnet_b200.synthetic import make_episode
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=_cfg(1)).eval()
    ep = make_episode(1, size=64)
    with pytest.raises(RuntimeError, match='CUDA'):
        net(ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], appr_query_labels=ep['appr_query_labels'])


def test_volume_helpers_cpu():
    """Host-side pieces of the slice-sharded volume loop (no kernels): Dice as utils/util.py:379-390, synthetic item shapes."""
    import torch
    from rpnet_b200 import volume as V
    p = torch.tensor([[1, 1, 0, 0]]); t = torch.tensor([[1, 0, 1, 0]])
    assert V.dice_from_sums(V.dice_sums(p, t)) == 0.5
    assert V.dice_from_sums(V.dice_sums(torch.zeros(2, 2), torch.zeros(2, 2))) is None
    item = V.make_synthetic_volume(6, 32, ways=1, shots=2, seed=1)
    assert item['query_images'].shape == (6, 1, 32, 32) and item['query_labels'].shape == (6, 32, 32)
    assert len(item['support_images']) == 1 and len(item['support_images'][0]) == 2
    assert torch.equal(item['support_bg'][0][0], 1 - item['support_fg'][0][0])
