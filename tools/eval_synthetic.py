#!/usr/bin/env python
"""End-to-end evaluation on synthetic volumes: batched affine registration (row N1) -> RP_Net eval forward (the hot path,
CUDA-graph replay) -> on-device Dice / NCC with the reference driver's printed lines (row N2).

    python tools/eval_synthetic.py [--volumes 4] [--slices 96] [--size 256] [--T 4]
    python tools/eval_synthetic.py --on-disk [--volumes 4] [--size 256]      # NRRD files -> FewshotRegReader (row N4) -> same loop

Random-init weights (no checkpoint ships with the reference), so the Dice values are those of an untrained model; the
script exists to exercise and time the whole pipeline the reference runs in `python test_rpnet.py --yaml ...`."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--volumes', type=int, default=4)
    ap.add_argument('--slices', type=int, default=96)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--T', type=int, default=4)
    ap.add_argument('--on-disk', action='store_true', help='write a synthetic ABD-110-shaped NRRD dataset and read it back through the episode builder')
    args = ap.parse_args()
    import torch
    from net.model import model_factory
    from rpnet_b200 import evaluate, registration, volume
    from rpnet_b200.synthetic import perturb_bn_stats
    cfg = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
               n_iter_refinement=args.T, soft_mask=False, mask_refinement_correlation_radius=5)
    torch.manual_seed(0)
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    perturb_bn_stats(net.state_dict())
    net = net.cuda().eval()
    net.enable_cuda_graph(True)
    dev = torch.device('cuda:0')
    items = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.on_disk:
        import random
        import tempfile
        from rpnet_b200.dataset import FewshotRegReader
        from rpnet_b200.dataset.synthetic_abd import make_synthetic_dataset
        with tempfile.TemporaryDirectory() as tmp:
            data_dir, set_name, dcfg = make_synthetic_dataset(tmp, n_patients=args.volumes, size=args.size + 8,
                                                              depths=(40, 48, 44, 52))
            dcfg.update(crop_size=[args.size, args.size], k=12)
            t0 = time.perf_counter()
            ds = FewshotRegReader(data_dir, set_name, dcfg, mode='eval')           # test_rpnet.py:70
            random.seed(0)
            for j in range(len(ds)):
                it = ds[j]
                c, si = it['supp_pids'][0]                                          # test_rpnet.py:182-183
                it['supp_pid'] = ds.fewshot_reader.fewshot_volume_reader.data_info[c][si]['pid']
                items.append(it)
        args.slices = sum(it['query_images'].shape[0] for it in items) // max(1, len(items))
    for v in range(0 if args.on_disk else args.volumes):
        raw = volume.make_synthetic_volume(args.slices, args.size, 1, 1, seed=100 * v)
        q = raw['query_images'].to(dev)
        s, l = [[raw['support_images'][0][0].to(dev)]], [[raw['support_fg'][0][0].to(dev)]]
        theta, warped_label, warped_src = registration.get_affine_registration(q, s, l)       # row N1
        items.append({'support_images': [[warped_src[:, None]]], 'support_labels': [[warped_label[:, 0]]],    # few_shot_reader.py:604-606
                      'query_images': q, 'query_labels': raw['query_labels'], 'appr_query_labels': (warped_label[:, 0] > 0.5).float(),
                      'warped_supp': warped_src, 'class_id': 0, 'pid': 'syn%03d' % v, 'supp_pid': 'syn_supp%03d' % v})
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    evaluate.eval_volumes(net, items, ['organ'], batch_size=16)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    n = args.volumes * args.slices
    print('registration %.1f ms (%d slices, incl. synthetic data generation on the host), evaluation %.1f ms (%.0f slices/s)'
          % ((t1 - t0) * 1e3, n, (t2 - t1) * 1e3, n / (t2 - t1)))


if __name__ == '__main__':
    main()
