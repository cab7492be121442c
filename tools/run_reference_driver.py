#!/usr/bin/env python
"""Run the reference's OWN evaluation driver, unmodified, on top of this repository (BASELINE.json north_star: "so test_rpnet.py
drops in unchanged"), and compare what it prints with rpnet_b200.evaluate.eval_volumes on the same data.

    python tools/run_reference_driver.py --stage     # build container: copy /root/reference/test_rpnet.py to oracle/_ref/ (git-ignored)
    python tools/run_reference_driver.py [--volumes 4] [--size 128] [--T 4]      # GPU box, >= 2 GPUs

The driver hard-codes CUDA_VISIBLE_DEVICES='1' (test_rpnet.py:3), hence the 2-GPU lease.  It is executed as
`python oracle/_ref/test_rpnet.py --yaml <generated yaml>` with this repository first on PYTHONPATH, so that its
`from net.model import model_factory`, `from dataset.few_shot_reader import FewshotRegReader`, `from utils.util import ...` and
`from net.registration import NCC, MSE` resolve to the shims at the repository root.  Data: the synthetic ABD-110-shaped NRRD set
of rpnet_b200.dataset.synthetic_abd (the real data is private); weights: random initialisation (the reference ships no checkpoint),
seeded by the driver itself (test_rpnet.py:8-10).  The yaml carries the keys of yamls/example.yml that the driver and the readers use."""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAGED = os.path.join(ROOT, 'oracle', '_ref', 'test_rpnet.py')


def stage():
    src = '/root/reference/test_rpnet.py'
    os.makedirs(os.path.dirname(STAGED), exist_ok=True)
    shutil.copyfile(src, STAGED)
    print('staged', src, '->', STAGED)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--stage', action='store_true')
    ap.add_argument('--volumes', type=int, default=4)
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--T', type=int, default=4)
    ap.add_argument('--log', default=os.path.join(ROOT, 'gpurun_out', 'reference_driver.log'))
    args = ap.parse_args()
    if args.stage:
        return stage()
    if not os.path.isfile(STAGED):
        raise SystemExit('%s is missing: run `python tools/run_reference_driver.py --stage` in the build container first' % STAGED)
    import yaml
    from rpnet_b200.dataset.synthetic_abd import make_synthetic_dataset
    with tempfile.TemporaryDirectory() as tmp:
        data_dir, set_name, dcfg = make_synthetic_dataset(tmp, n_patients=args.volumes, size=args.size + 8, depths=(20, 24, 22, 26))
        cfg = dict(dcfg)
        cfg.update(crop_size=[args.size, args.size], k=4, net='RP_Net', ckpt=None, optimizer='Adam', eval_set_name=set_name, data_dir=data_dir,
                   out_dir=os.path.join(tmp, 'results') + '/', pretrained_path=None, backbone='UNet', unet_normalize_type='BatchNorm2d',
                   final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=4, n_test_iter_refinement=args.T, soft_mask=False,
                   mask_refinement_correlation_radius=5, n_runs=1, align_loss_scaler=1, loss='dice_ce')
        ypath = os.path.join(tmp, 'synthetic.yml')
        with open(ypath, 'w') as f:
            yaml.safe_dump(cfg, f)
        env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
        r = subprocess.run([sys.executable, STAGED, '--yaml', ypath], cwd=tmp, env=env, capture_output=True, text=True)
        ref_out = r.stdout
        os.makedirs(os.path.dirname(args.log), exist_ok=True)
        with open(args.log, 'w') as f:
            f.write('# python oracle/_ref/test_rpnet.py --yaml synthetic.yml   (unmodified reference driver, rc=%d)\n' % r.returncode)
            f.write(ref_out)
            f.write('\n# ---- stderr (tail)\n' + r.stderr[-3000:])
        if r.returncode != 0:
            print(r.stderr[-3000:])
            raise SystemExit('the reference driver failed (rc=%d); log: %s' % (r.returncode, args.log))
        # the same evaluation through this repository's own driver loop, in the same process layout (device 0 of what is visible)
        import random
        import numpy as np
        import torch
        np.random.seed(0); random.seed(0); torch.manual_seed(0)                     # test_rpnet.py:7-10
        from net.model import model_factory
        from dataset.few_shot_reader import FewshotRegReader
        from rpnet_b200 import evaluate
        config = yaml.safe_load(open(ypath))
        config['n_iter_refinement'] = config['n_test_iter_refinement']                # test_rpnet.py:51
        ds = FewshotRegReader(data_dir, set_name, config, mode='eval')
        net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=config).cuda()
        items = []
        for j in range(len(ds)):
            it = ds[j]
            c, si = it['supp_pids'][0]
            it['supp_pid'] = ds.fewshot_reader.fewshot_volume_reader.data_info[c][si]['pid']
            items.append(it)
        ours = []
        evaluate.eval_volumes(net, items, config['eval_classes'], batch_size=16, out=ours.append)
    vol = re.compile(r'^\d+ \S+ \S+ affine \(')
    # per-volume lines (:231-242) and the per-class summary of eval() (:246-251, the one with 'voxel morph'); the closing
    # "Average performance" block of main() (:128-141) averages over n_runs and is not part of the loop being compared
    ref_lines = [l.rstrip() for l in ref_out.splitlines() if vol.match(l) or (l.startswith(tuple(c + ', affine' for c in config['eval_classes']))
                                                                               and 'voxel morph' in l)]
    our_lines = [l.rstrip() for l in ours]
    same = ref_lines == our_lines
    with open(args.log, 'a') as f:
        f.write('\n# ---- rpnet_b200.evaluate.eval_volumes on the same data (batches of 16 instead of 2)\n' + '\n'.join(our_lines) + '\n')
        f.write('# ---- per-volume and per-class lines identical: %s\n' % same)
    print('\n'.join(ref_lines))
    print('reference driver lines == eval_volumes lines:', same)
    if not same:
        for a, b in zip(ref_lines, our_lines):
            if a != b:
                print('REF :', a)
                print('OURS:', b)
        raise SystemExit(1)


if __name__ == '__main__':
    main()
