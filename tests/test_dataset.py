"""Row N4: NRRD container reader and the eval episode builder (dataset/few_shot_reader.py) against hand-assembled NRRD
bytes and the golden vectors of the reference's own Dataset classes (tests/golden/make_golden_dataset.py)."""
import gzip
import os
import random

import numpy as np
import pytest
import torch

from rpnet_b200.dataset import FewshotRegReader, FewshotSliceReader, nrrd_io
from rpnet_b200.dataset.synthetic_abd import make_synthetic_dataset

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'dataset.npz')


def _digest(t):
    a = t.detach().cpu().double().numpy() if torch.is_tensor(t) else np.asarray(t, dtype=np.float64)
    idx = np.arange(a.size, dtype=np.float64).reshape(a.shape)
    return np.array([a.sum(), (a * np.cos(idx)).sum(), float(a.size)] + list(a.shape) + [0] * (5 - a.ndim))


# ---------------------------------------------------------------------------------------------------------------------
def test_nrrd_reader_hand_assembled_files(tmp_path):
    vals = np.arange(2 * 3 * 4, dtype=np.int16) - 5                  # file order: first axis of `sizes` is fastest
    want = vals.reshape(4, 3, 2).transpose(2, 1, 0)                  # sizes: 2 3 4  -> data[i, j, k] = vals[i + 2j + 6k]
    head = b'NRRD0004\n# a comment\ntype: short\ndimension: 3\nsizes: 2 3 4\nspace: left-posterior-superior\n' \
           b'space directions: (1,0,0) (0,1,0) (0,0,2.5)\nspace origin: (0,0,0)\nkinds: domain domain domain\n' \
           b'endian: %s\nencoding: %s\nmy key:=a value\n\n'
    cases = {'raw_le.nrrd': head % (b'little', b'raw') + vals.astype('<i2').tobytes(),
             'raw_be.nrrd': head % (b'big', b'raw') + vals.astype('>i2').tobytes(),
             'gz.nrrd': head % (b'little', b'gzip') + gzip.compress(vals.astype('<i2').tobytes()),
             'txt.nrrd': head % (b'little', b'ascii') + ' '.join(str(v) for v in vals).encode() + b'\n'}
    for name, blob in cases.items():
        p = tmp_path / name
        p.write_bytes(blob)
        data, hdr = nrrd_io.read(str(p))
        assert data.shape == (2, 3, 4) and data.dtype.kind == 'i' and data.dtype.itemsize == 2, name
        np.testing.assert_array_equal(data, want, err_msg=name)
        assert hdr['dimension'] == 3 and list(hdr['sizes']) == [2, 3, 4] and hdr['my key'] == 'a value'
        assert hdr['space directions'].shape == (3, 3) and hdr['space directions'][2, 2] == 2.5
        assert data.flags.writeable
    # detached header: payload = the tail of another file (byte skip -1), float type, CRLF line ends
    f32 = np.linspace(-1, 1, 12, dtype=np.float32)
    (tmp_path / 'blob.raw').write_bytes(b'JUNKJUNK' + f32.astype('<f4').tobytes())
    (tmp_path / 'det.nhdr').write_bytes(b'NRRD0005\r\ntype: float\r\ndimension: 2\r\nsizes: 3 4\r\nendian: little\r\n'
                                        b'encoding: raw\r\nbyteskip: -1\r\ndatafile: blob.raw\r\n')
    data, _ = nrrd_io.read(str(tmp_path / 'det.nhdr'))
    np.testing.assert_array_equal(data, f32.reshape(4, 3).T)
    # byte skip / line skip in an attached raw payload
    (tmp_path / 'skip.nrrd').write_bytes(b'NRRD0004\ntype: uchar\ndimension: 1\nsizes: 4\nencoding: raw\nline skip: 1\n'
                                         b'byte skip: 2\n\nskipped line\nXY\x01\x02\x03\x04')
    np.testing.assert_array_equal(nrrd_io.read(str(tmp_path / 'skip.nrrd'))[0], np.array([1, 2, 3, 4], dtype=np.uint8))


def test_nrrd_errors(tmp_path):
    def check(blob, match):
        p = tmp_path / 'bad.nrrd'
        p.write_bytes(blob)
        with pytest.raises(nrrd_io.NrrdError, match=match):
            nrrd_io.read(str(p))
    check(b'NOPE\n\n', 'not an NRRD')
    check(b'NRRD0004\ntype: short\ndimension: 1\nencoding: raw\n\n', 'sizes')
    check(b'NRRD0004\ntype: short\ndimension: 2\nsizes: 4\nencoding: raw\nendian: little\n\n', 'dimension')
    check(b'NRRD0004\ntype: short\ndimension: 1\nsizes: 4\nencoding: raw\n\n' + b'\0' * 8, 'endian')
    check(b'NRRD0004\ntype: short\ndimension: 1\nsizes: 4\nencoding: raw\nendian: little\n\n\0\0', 'bytes')
    check(b'NRRD0004\ntype: quaternion\ndimension: 1\nsizes: 1\nencoding: raw\n\n\0', 'unknown NRRD type')
    check(b'NRRD0004\ntype: uchar\ndimension: 1\nsizes: 1\nencoding: zstd\n\n\0', 'encoding')
    check(b'NRRD0004\ntype: uchar\ndimension: 1\nsizes: 4\nencoding: gzip\n\nnot gzip', 'decompress')
    check(b'NRRD0004\ntype: uchar\ntype: uchar\ndimension: 1\nsizes: 1\nencoding: raw\n\n\0', 'duplicate')


@pytest.mark.parametrize('dtype', ['uint8', 'int8', 'int16', 'uint16', 'int32', 'uint32', 'int64', 'float32', 'float64'])
@pytest.mark.parametrize('encoding', ['raw', 'gzip', 'bzip2'])
def test_nrrd_round_trip(tmp_path, dtype, encoding):
    rs = np.random.RandomState(3)
    a = (rs.randn(5, 7, 3) * 50).astype(dtype)
    p = str(tmp_path / 'a.nrrd')
    nrrd_io.write(p, a, header={'space origin': [0.0, 1.0, 2.0], 'note': 'x\ny'}, encoding=encoding)
    b, hdr = nrrd_io.read(p)
    assert b.dtype == np.dtype(dtype) and b.shape == a.shape
    np.testing.assert_array_equal(a, b)
    assert hdr['note'] == 'x\ny' and list(hdr['space origin']) == [0.0, 1.0, 2.0]
    # empty leading axis and a C-ordered view survive as well
    nrrd_io.write(p, a[:, ::2].transpose(1, 0, 2), encoding=encoding)
    np.testing.assert_array_equal(nrrd_io.read(p)[0], a[:, ::2].transpose(1, 0, 2))


# ---------------------------------------------------------------------------------------------------------------------
def test_slice_reader_matches_reference_golden(tmp_path):
    """FewshotVolumeReader + FewshotSliceReader (eval, no registration) == the reference classes, item by item."""
    g = np.load(GOLD)
    data_dir, set_name, cfg = make_synthetic_dataset(str(tmp_path))
    ds = FewshotSliceReader(data_dir, set_name, dict(cfg, use_registration_loss=False), mode='eval')
    assert len(ds) == int(g['n_items'])
    for i in range(len(ds)):
        random.seed(100 + i)
        it = ds[i]
        np.testing.assert_array_equal(np.array(it['supp_pids'][0]), g['plain%d_supp_pid' % i])
        assert ds.k == int(g['plain%d_k' % i])
        for key, val in (('support_images', it['support_images'][0][0]), ('support_labels', it['support_labels'][0][0]),
                         ('query_images', it['query_images']), ('query_labels', it['query_labels']),
                         ('q3d', it['query_images_3D'][0][0])):
            np.testing.assert_allclose(_digest(val), g['plain%d_%s' % (i, key)], rtol=1e-11, atol=1e-9, err_msg='%d %s' % (i, key))
        assert it['query_images'].shape[1] == 3 and it['support_images'][0][0].shape == it['query_images'].shape
        assert it['registration_field'] is None and it['warped_supp_label'] is None
        if i == 0:
            np.testing.assert_array_equal(it['query_images'][:, 0, ::4, ::4].numpy(), g['plain0_query_full'])
            np.testing.assert_array_equal(it['support_images'][0][0][:, 0, ::4, ::4].numpy(), g['plain0_support_full'])
            np.testing.assert_array_equal(np.packbits(it['query_labels'].numpy().astype(np.uint8)), g['plain0_labels_full'])
            np.testing.assert_array_equal(np.packbits(it['support_labels'][0][0].numpy().astype(np.uint8)),
                                          g['plain0_support_labels_full'])


def test_slice_blocks_edge_cases():
    s, q = FewshotSliceReader.slice_blocks([12, 12], 12)               # k == number of slices: identity matching
    np.testing.assert_array_equal(s[0], np.arange(12))
    np.testing.assert_array_equal(q, np.arange(13))
    s, q = FewshotSliceReader.slice_blocks([17, 5, 23], 4)             # ragged supports, non-divisible query
    assert [len(v) for v in s] == [4, 4] and q[0] == 0 and q[-1] == 23 and (np.diff(q) > 0).all()
    assert s[0].max() < 17 and s[1].max() < 5


def test_train_mode_items_match_reference_golden(tmp_path):
    """`mode='train'` (dataset/few_shot_reader.py:482-515): random query slice per block, gamma augmentation, one random affine per
    slice (image and label together), shuffle of the k pairs — bit for bit the items of the reference classes under the same
    `random` / `np.random` / torch seeds (elastic off: the reference seeds that one from OS entropy)."""
    g = np.load(GOLD)
    data_dir, set_name, cfg = make_synthetic_dataset(str(tmp_path))
    cfg = dict(cfg, use_registration_loss=False, train_classes=cfg['eval_classes'], do_elastic=False, do_intaug=True)
    ds = FewshotSliceReader(data_dir, set_name, cfg, mode='train')
    for i in (0, 1, 3):
        random.seed(200 + i); np.random.seed(300 + i); torch.manual_seed(400 + i)
        it = ds[i]
        assert ds.k == int(g['train%d_k' % i])
        np.testing.assert_array_equal(np.array(it['supp_pids'][0]), g['train%d_supp_pid' % i])
        shapes = list(it['support_images'][0][0].shape) + list(it['query_images'].shape) + list(it['query_labels'].shape)
        np.testing.assert_array_equal(np.array(shapes), g['train%d_shapes' % i])
        dtypes = [str(it['query_images'].dtype), str(it['query_labels'].dtype), str(it['support_images'][0][0].dtype)]
        assert dtypes == [str(d) for d in g['train%d_dtypes' % i]]
        np.testing.assert_array_equal(it['query_images'][:, 0, ::2, ::2].numpy().astype(np.float32), g['train%d_query_images' % i])
        np.testing.assert_array_equal(np.packbits(it['query_labels'].numpy().astype(np.uint8)), g['train%d_query_labels' % i])
        np.testing.assert_array_equal(it['support_images'][0][0][:, 0, ::2, ::2].numpy().astype(np.float32), g['train%d_support_images' % i])
        np.testing.assert_array_equal(np.packbits(it['support_labels'][0][0].numpy().astype(np.uint8)), g['train%d_support_labels' % i])
        assert it['query_images'].shape[1] == 3 and it['support_images'][0][0].shape == it['query_images'].shape


def test_elastic_transform_matches_reference_golden():
    """dataset/brain_reader.py:248-293 with an explicit generator (the reference's own call seeds from OS entropy, so the train
    readers' elastic branch is checked for shape / label preservation only)."""
    from rpnet_b200.dataset import augment
    g = np.load(GOLD)
    rs = np.random.RandomState(11)
    vol = (rs.rand(1, 3, 96, 128).astype(np.float32) * 2 - 1)
    msk = np.zeros((2, 3, 96, 128), np.float32); msk[0, :, 30:70, 40:90] = 1; msk[1, 1, 25:45, 30:60] = 1
    ei, em = augment.elastic_transform(vol, msk, alpha=100, sigma=6, alpha_affine=3.0, random_state=np.random.RandomState(5))
    np.testing.assert_array_equal(ei.astype(np.float32), g['elastic_image'])
    # masks: interior only — the reference warps them with BORDER_TRANSPARENT into an uninitialised destination (garbage in the
    # strip the warp uncovers); here that strip is zero
    np.testing.assert_array_equal(np.packbits(em[:, :, 20:-20, 20:-20].astype(np.uint8)), g['elastic_mask_interior'])
    assert set(np.unique(em)) <= {0.0, 1.0}
    ei2, em2 = augment.elastic_transform_all(vol, msk)                  # OS-entropy seeded, like the reference
    assert ei2.shape == vol.shape and em2.shape == msk.shape and set(np.unique(em2)) <= {0.0, 1.0}
    assert abs(float(em2[0].sum()) / float(msk[0].sum()) - 1.0) < 0.5


def test_train_mode_with_elastic_runs(tmp_path):
    data_dir, set_name, cfg = make_synthetic_dataset(str(tmp_path), n_patients=2)
    cfg = dict(cfg, use_registration_loss=False, train_classes=cfg['eval_classes'], do_elastic=True, do_intaug=True)
    ds = FewshotSliceReader(data_dir, set_name, cfg, mode='train')
    np.random.seed(1)                                                   # the coin of :306 comes up 1 for this seed
    it = ds[0]
    assert it['query_images'].shape[1] == 3 and it['query_images'].shape[0] == ds.k
    assert torch.isfinite(it['query_images']).all()


def test_reg_reader_requires_registration(tmp_path):
    data_dir, set_name, cfg = make_synthetic_dataset(str(tmp_path), n_patients=2)
    with pytest.raises(NotImplementedError):
        FewshotRegReader(data_dir, set_name, cfg, mode='test')
    ds = FewshotRegReader(data_dir, set_name, dict(cfg, use_registration_loss=False), mode='eval')
    with pytest.raises(TypeError):
        ds[0]


@pytest.mark.gpu
def test_reg_reader_deformable_item(tmp_path):
    """`do_deformable: True` (the default of get_registration_field, dataset/few_shot_reader.py:109): same item contract, the demons
    stage moves the warped support towards the query (NCC improves over the affine-only warp) and the fields carry the flow."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import ops
    data_dir, set_name, cfg = make_synthetic_dataset(str(tmp_path), n_patients=2)
    random.seed(1)
    a = FewshotRegReader(data_dir, set_name, dict(cfg, do_deformable=False), mode='eval')[0]
    random.seed(1)
    d = FewshotRegReader(data_dir, set_name, dict(cfg, do_deformable=True), mode='eval')[0]
    assert set(a.keys()) == set(d.keys())
    for k in ('support_images', 'support_labels'):
        assert d[k][0][0].shape == a[k][0][0].shape
    assert d['warped_supp'].shape == a['warped_supp'].shape and d['appr_query_labels'].shape == a['appr_query_labels'].shape
    assert len(d['registration_field'][0]) == 3 and d['registration_field'][0][2].shape[0] == 2          # [theta, grid, flow]
    q = d['query_images'].float().cuda().contiguous()
    ncc = lambda w: ops.ncc(q, w.float().cuda().reshape(q.shape).contiguous()).item()
    assert ncc(d['warped_supp']) < ncc(a['warped_supp'])                                                   # NCC is negated: lower is better
    assert set(torch.unique(d['appr_query_labels']).tolist()) <= {0.0, 1.0}


@pytest.mark.gpu
def test_reg_reader_item_matches_reference_golden(tmp_path):
    """The item test_rpnet.py consumes: registration on the device (one launch for the volume) vs the reference's per-slice
    CPU optimisation.  Tolerances as tests/test_registration.py: theta 2e-3, label maps differ on edge pixels only."""
    g = np.load(GOLD)
    data_dir, set_name, cfg = make_synthetic_dataset(str(tmp_path))
    ds = FewshotRegReader(data_dir, set_name, cfg, mode='eval')
    for i in (0, 2):
        random.seed(100 + i)
        it = ds[i]
        S = it['query_images'].shape[0]
        shapes = list(it['support_images'][0][0].shape) + list(it['query_images'].shape) + list(it['appr_query_labels'].shape) \
            + list(it['grid'].shape)
        np.testing.assert_array_equal(np.array(shapes), g['reg%d_shapes' % i])
        assert it['appr_query_labels'].is_cuda and it['support_images'][0][0].is_cuda
        theta = torch.stack([t for t, _ in it['registration_field']]).cpu().numpy()
        np.testing.assert_allclose(theta, g['reg%d_theta' % i], rtol=0, atol=2e-3)
        np.testing.assert_allclose(_digest(it['grid']), g['reg%d_grid_digest' % i], rtol=1e-5, atol=1e-3)
        np.testing.assert_array_equal(it['query_images'][:, 0, ::2, ::2].cpu().numpy(), g['reg%d_query_images' % i])
        np.testing.assert_allclose(it['support_images'][0][0][:, 0, ::2, ::2].cpu().numpy(), g['reg%d_support_images' % i], atol=2e-2)
        np.testing.assert_allclose(it['warped_supp'][:, ::2, ::2].cpu().numpy(), g['reg%d_warped_supp' % i], atol=2e-2)
        H, W = it['appr_query_labels'].shape[1:]
        for key, val in (('appr', it['appr_query_labels']), ('support_labels', it['support_labels'][0][0])):
            ref = np.unpackbits(g['reg%d_%s' % (i, key)])[:S * H * W].reshape(S, H, W)
            got = val.cpu().numpy().astype(np.uint8)
            assert set(np.unique(got)) <= {0, 1}
            assert (got != ref).mean() < 2e-3, (key, (got != ref).mean())
        # the network runs on the item as the eval loop passes it (test_rpnet.py:163-215)
    from rpnet_b200.nn.rp_net import RP_Net
    net = RP_Net(cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=dict(
        unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=2, soft_mask=False,
        mask_refinement_correlation_radius=5)).cuda().eval()
    with torch.no_grad():
        fg = it['support_labels'][0][0].float().cuda()
        out = net([[it['support_images'][0][0].float().cuda()]], [[fg]], [[1 - fg]], [it['query_images'].float().cuda()],
                  grid=it['grid'].cuda(), query_labels=it['query_labels'].long().cuda(),
                  appr_query_labels=it['appr_query_labels'].cuda())
    assert out['output'].shape == (S, 2, H, W) and torch.isfinite(out['output']).all()


@pytest.mark.gpu
def test_reg_reader_train_item_feeds_a_train_step(tmp_path):
    """FewshotRegReader(mode='train'): the k augmented (support, query) pairs, registered on the device, against the reference's item
    (CPU registration; same tolerances as the eval item), then one train step of the network on it (what a train loop would do —
    the reference ships none)."""
    g = np.load(GOLD)
    data_dir, set_name, cfg = make_synthetic_dataset(str(tmp_path))
    ds = FewshotRegReader(data_dir, set_name, dict(cfg, train_classes=cfg['eval_classes'], do_elastic=False, do_intaug=True), mode='train')
    random.seed(210); np.random.seed(310); torch.manual_seed(410)
    it = ds[1]
    S = it['query_images'].shape[0]
    shapes = list(it['support_images'][0][0].shape) + list(it['query_images'].shape) + list(it['appr_query_labels'].shape) + list(it['grid'].shape)
    np.testing.assert_array_equal(np.array(shapes), g['regtrain_shapes'])
    np.testing.assert_array_equal(it['query_images'][:, 0, ::2, ::2].cpu().numpy().astype(np.float32), g['regtrain_query_images'])
    np.testing.assert_array_equal(np.packbits(it['query_labels'].cpu().numpy().astype(np.uint8)), g['regtrain_query_labels'])
    theta = torch.stack([t for t, _ in it['registration_field']]).cpu().numpy()
    np.testing.assert_allclose(theta, g['regtrain_theta'], rtol=0, atol=5e-3)
    np.testing.assert_allclose(it['support_images'][0][0][:, 0, ::2, ::2].cpu().numpy(), g['regtrain_support_images'], atol=5e-2)
    H, W = it['appr_query_labels'].shape[1:]
    for key, val in (('appr', it['appr_query_labels']), ('support_labels', it['support_labels'][0][0])):
        ref = np.unpackbits(g['regtrain_%s' % key])[:S * H * W].reshape(S, H, W)
        got = val.cpu().numpy().astype(np.uint8)
        assert (got != ref).mean() < 5e-3, (key, (got != ref).mean())
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.train import TrainStep
    net = RP_Net(cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=dict(
        unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=2, soft_mask=False,
        mask_refinement_correlation_radius=5)).cuda().train()
    fg = it['support_labels'][0][0].float().cuda()
    d = {'supp_imgs': [[it['support_images'][0][0].float().cuda()]], 'fore_mask': [[fg]], 'back_mask': [[1 - fg]],
         'qry_imgs': [it['query_images'].float().cuda()], 'query_labels': it['query_labels'].long().cuda(),
         'appr_query_labels': it['appr_query_labels'].float().cuda()}
    loss = TrainStep(net, lr=1e-4).step(d)
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all()
