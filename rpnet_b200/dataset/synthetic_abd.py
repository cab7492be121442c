"""A tiny ABD-110-shaped dataset on disk (the real one is private): per patient `<pid>_clean.nrrd` (CT in HU, int16,
[D, H, W]) and `<pid>_<ROI>.nrrd` (uint8 mask), `<class_csv_dir>/<ROI>.csv` with the pid / z_start / z_end columns of
split/abd_110_classes/*.csv, and a split list like split/abd_110_test.csv.  Deterministic (numpy RandomState), so tests and
tools/eval_synthetic.py can rebuild the same files anywhere."""
import os

import numpy as np

from . import nrrd_io


def _volume(seed, depth, size, roi_scale):
    rs = np.random.RandomState(seed)
    lin = np.linspace(-1.0, 1.0, size, dtype=np.float32)
    yy, xx = np.meshgrid(lin, lin, indexing='ij')
    z_lo, z_hi = 2 + rs.randint(0, 3), depth - 2 - rs.randint(0, 3)           # annotated range [z_lo, z_hi]
    cx, cy = 0.25 * (rs.rand() - 0.5), 0.25 * (rs.rand() - 0.5)
    rx, ry = roi_scale * (0.30 + 0.1 * rs.rand()), roi_scale * (0.22 + 0.1 * rs.rand())
    ct = np.full((depth, size, size), -1000.0, dtype=np.float32)
    mask = np.zeros((depth, size, size), dtype=np.uint8)
    body = ((xx / 0.85) ** 2 + (yy / 0.65) ** 2) < 1
    zc, zr = 0.5 * (z_lo + z_hi), 0.5 * (z_hi - z_lo) + 0.5
    for z in range(depth):
        ct[z][body] = 40.0
        t = 1.0 - ((z - zc) / zr) ** 2
        if z_lo <= z <= z_hi and t > 0:
            s = np.sqrt(t)
            organ = (((xx - cx) / (rx * s + 1e-3)) ** 2 + ((yy - cy) / (ry * s + 1e-3)) ** 2) < 1
            if not organ.any():                                                # keep every slice of the range annotated
                organ[size // 2, size // 2] = True
            ct[z][organ] = 110.0
            mask[z][organ] = 1
    ct += rs.randn(depth, size, size).astype(np.float32) * 20.0
    return np.clip(np.rint(ct), -1024, 3071).astype(np.int16), mask, z_lo, z_hi


def make_synthetic_dataset(root, n_patients=4, size=72, depths=(22, 26, 24, 28), roi='Liver', seed=0, encoding='gzip'):
    """Writes the dataset under `root` and returns (data_dir, set_name, config-fragment) for the readers."""
    data_dir, csv_dir = os.path.join(root, 'preprocessed'), os.path.join(root, 'classes')
    os.makedirs(data_dir, exist_ok=True)
    os.makedirs(csv_dir, exist_ok=True)
    rows, pids = [], []
    for p in range(n_patients):
        pid = 'PA%03d' % (p + 1)
        ct, mask, z_lo, z_hi = _volume(seed * 1000 + p, depths[p % len(depths)], size, 1.0 + 0.1 * (p % 3))
        nrrd_io.write(os.path.join(data_dir, '%s_clean.nrrd' % pid), ct, encoding=encoding)
        nrrd_io.write(os.path.join(data_dir, '%s_%s.nrrd' % (pid, roi)), mask, encoding=encoding)
        rows.append('%s,%d,%d,%s' % (pid, z_lo, z_hi, os.path.join(data_dir, '%s_%s.nrrd' % (pid, roi))))
        pids.append(pid)
    with open(os.path.join(csv_dir, '%s.csv' % roi), 'w') as fh:
        fh.write('pid,z_start,z_end,path\n' + '\n'.join(rows) + '\n')
    set_name = os.path.join(root, 'test.csv')
    with open(set_name, 'w') as fh:
        fh.write('\n'.join(pids) + '\n')
    cfg = {'class_csv_dir': csv_dir, 'eval_classes': [roi], 'train_classes': [roi], 'n_shot': 1, 'n_way': 1, 'k': 4,
           'num_slice': 280, 'num_x': 272, 'num_y': 272, 'pad_value': -1024, 'HU_range': [-1024, 3072], 'crop_size': [64, 64],
           'use_registration_loss': True, 'use_registration_mask': True, 'do_deformable': False}
    return data_dir, set_name, cfg
