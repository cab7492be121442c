"""Golden vectors for the episode builder (row N4), from the UNMODIFIED reference Dataset classes
(dataset/few_shot_reader.py) run on the synthetic on-disk dataset of rpnet_b200.dataset.synthetic_abd.
The reference's `nrrd` import (pynrrd, absent here) is bound to rpnet_b200.dataset.nrrd_io.read, so the golden covers
everything above the container format.  Run in the build container only:  python tests/golden/make_golden_dataset.py"""
import importlib
import os
import random
import sys
import tempfile
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
from oracle import ref_import                                   # noqa: E402
from rpnet_b200.dataset import nrrd_io                          # noqa: E402
from rpnet_b200.dataset.synthetic_abd import make_synthetic_dataset   # noqa: E402

ref_import.load()
saved = {k: sys.modules.pop(k) for k in list(sys.modules)
         if k in ('net', 'utils', 'dataset') or k.startswith(('net.', 'utils.', 'dataset.'))}
for k in list(sys.modules):                                     # let `dataset.*` import the already loaded reference net/utils
    if k.startswith('_rpnet_ref_'):
        sys.modules[k[len('_rpnet_ref_'):]] = sys.modules[k]
sys.path.insert(0, ref_import.REF_ROOT)
fsr = importlib.import_module('dataset.few_shot_reader')
fsr.nrrd.read = nrrd_io.read
# the train branch passes `fillcolor=None` to transforms.RandomAffine (few_shot_reader.py:29,52), a keyword torchvision removed:
# map it to today's spelling of the same request (constant fill 0) so that the reference's own code runs
_RA = fsr.transforms.RandomAffine


def _random_affine(*a, fillcolor=None, **kw):
    return _RA(*a, fill=0 if fillcolor is None else fillcolor, **kw)


fsr.transforms.RandomAffine = _random_affine


def digest(t):
    a = t.detach().cpu().double().numpy() if torch.is_tensor(t) else np.asarray(t, dtype=np.float64)
    idx = np.arange(a.size, dtype=np.float64).reshape(a.shape)
    return np.array([a.sum(), (a * np.cos(idx)).sum(), float(a.size)] + list(a.shape) + [0] * (5 - a.ndim))


out = {}
with tempfile.TemporaryDirectory() as tmp:
    data_dir, set_name, cfg = make_synthetic_dataset(tmp)
    # (1) volume + slice readers without registration: exact
    cfg_plain = dict(cfg, use_registration_loss=False)
    ds = fsr.FewshotSliceReader(data_dir, set_name, cfg_plain, mode='eval')
    out['n_items'] = len(ds)
    for i in range(len(ds)):
        random.seed(100 + i)
        it = ds[i]
        out['plain%d_supp_pid' % i] = np.array(it['supp_pids'][0])
        out['plain%d_k' % i] = ds.k
        for key in ('support_images', 'support_labels'):
            out['plain%d_%s' % (i, key)] = digest(it[key][0][0])
        for key in ('query_images', 'query_labels'):
            out['plain%d_%s' % (i, key)] = digest(it[key])
        out['plain%d_q3d' % i] = digest(it['query_images_3D'][0][0])
        if i == 0:                                              # one item in full (small): slices subsampled
            out['plain0_query_full'] = it['query_images'][:, 0, ::4, ::4].numpy().astype(np.float32)
            out['plain0_support_full'] = it['support_images'][0][0][:, 0, ::4, ::4].numpy().astype(np.float32)
            out['plain0_labels_full'] = np.packbits(it['query_labels'].numpy().astype(np.uint8))
            out['plain0_support_labels_full'] = np.packbits(it['support_labels'][0][0].numpy().astype(np.uint8))
    # (2) the full FewshotRegReader item with the reference's CPU registration (do_deformable False)
    ds = fsr.FewshotRegReader(data_dir, set_name, cfg, mode='eval')
    for i in (0, 2):
        random.seed(100 + i)
        torch.manual_seed(0)
        it = ds[i]
        out['reg%d_appr' % i] = np.packbits(it['appr_query_labels'].numpy().astype(np.uint8))
        out['reg%d_support_labels' % i] = np.packbits(it['support_labels'][0][0].numpy().astype(np.uint8))
        out['reg%d_support_images' % i] = it['support_images'][0][0][:, 0, ::2, ::2].numpy().astype(np.float32)
        out['reg%d_query_images' % i] = it['query_images'][:, 0, ::2, ::2].numpy().astype(np.float32)
        out['reg%d_warped_supp' % i] = it['warped_supp'][:, ::2, ::2].numpy().astype(np.float32)
        out['reg%d_shapes' % i] = np.array(list(it['support_images'][0][0].shape) + list(it['query_images'].shape)
                                           + list(it['appr_query_labels'].shape) + list(it['grid'].shape))
        out['reg%d_grid_digest' % i] = digest(it['grid'])
        out['reg%d_theta' % i] = np.stack([r.affine_reg.theta.detach().numpy()[0] for r, _ in it['registration_field']])
    # (3) mode='train': random slice per block + gamma + random affine + shuffle (elastic off: the reference seeds it from OS entropy)
    cfg_train = dict(cfg, use_registration_loss=False, train_classes=cfg['eval_classes'], do_elastic=False, do_intaug=True)
    ds = fsr.FewshotSliceReader(data_dir, set_name, cfg_train, mode='train')
    for i in (0, 1, 3):
        random.seed(200 + i); np.random.seed(300 + i); torch.manual_seed(400 + i)
        it = ds[i]
        out['train%d_k' % i] = ds.k
        out['train%d_supp_pid' % i] = np.array(it['supp_pids'][0])
        out['train%d_query_images' % i] = it['query_images'][:, 0, ::2, ::2].numpy().astype(np.float32)
        out['train%d_query_labels' % i] = np.packbits(it['query_labels'].numpy().astype(np.uint8))
        out['train%d_support_images' % i] = it['support_images'][0][0][:, 0, ::2, ::2].numpy().astype(np.float32)
        out['train%d_support_labels' % i] = np.packbits(it['support_labels'][0][0].numpy().astype(np.uint8))
        out['train%d_shapes' % i] = np.array(list(it['support_images'][0][0].shape) + list(it['query_images'].shape)
                                             + list(it['query_labels'].shape))
        out['train%d_dtypes' % i] = np.array([str(it['query_images'].dtype), str(it['query_labels'].dtype),
                                              str(it['support_images'][0][0].dtype)])
    # (3b) the full train item: FewshotRegReader(mode='train') with the reference's CPU registration of the k augmented pairs
    ds = fsr.FewshotRegReader(data_dir, set_name, dict(cfg, train_classes=cfg['eval_classes'], do_elastic=False, do_intaug=True), mode='train')
    random.seed(210); np.random.seed(310); torch.manual_seed(410)
    it = ds[1]
    out['regtrain_appr'] = np.packbits(it['appr_query_labels'].numpy().astype(np.uint8))
    out['regtrain_support_labels'] = np.packbits(it['support_labels'][0][0].numpy().astype(np.uint8))
    out['regtrain_query_images'] = it['query_images'][:, 0, ::2, ::2].numpy().astype(np.float32)
    out['regtrain_query_labels'] = np.packbits(it['query_labels'].numpy().astype(np.uint8))
    out['regtrain_support_images'] = it['support_images'][0][0][:, 0, ::2, ::2].numpy().astype(np.float32)
    out['regtrain_shapes'] = np.array(list(it['support_images'][0][0].shape) + list(it['query_images'].shape)
                                      + list(it['appr_query_labels'].shape) + list(it['grid'].shape))
    out['regtrain_theta'] = np.stack([r.affine_reg.theta.detach().numpy()[0] for r, _ in it['registration_field']])
    # (4) the elastic deformation with an explicit generator (dataset/brain_reader.py:248-293)
    br = importlib.import_module('dataset.brain_reader')
    rs = np.random.RandomState(11)
    vol = (rs.rand(1, 3, 96, 128).astype(np.float32) * 2 - 1)
    msk = np.zeros((2, 3, 96, 128), np.float32); msk[0, :, 30:70, 40:90] = 1; msk[1, 1, 25:45, 30:60] = 1
    ei, em = br.elastic_transform(vol, msk, alpha=100, sigma=6, alpha_affine=3.0, random_state=np.random.RandomState(5))
    out['elastic_image'] = ei.astype(np.float32)
    # the reference warps the masks with BORDER_TRANSPARENT into an uninitialised destination: only the interior is defined
    out['elastic_mask_interior'] = np.packbits(em[:, :, 20:-20, 20:-20].astype(np.uint8))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dataset.npz')
np.savez_compressed(path, **out)
print('dataset.npz', os.path.getsize(path), 'bytes;', {k: v for k, v in out.items() if k.endswith('_k') or k == 'n_items'})
