#!/usr/bin/env python
"""Throughput of the batched affine registration ("next" row N1) next to the reference's own slice-by-slice torch path.

    python tools/bench_register.py [--slices 96] [--size 256] [--iters 50]

Prints one JSON line: slices/s of `rpnet_b200.registration.get_affine_registration` (one launch for all slices) with the
inputs resident in HBM, the same through pinned host buffers (`e2e`), and `cpu_baseline` = the oracle restatement of
get_registration_field's affine part (dataset/few_shot_reader.py:109-198) on the host cores for a bounded sample."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--slices', type=int, default=96)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--steps', type=int, default=10)
    args = ap.parse_args()
    import torch
    from oracle import registration_oracle as R
    from rpnet_b200 import registration as RG
    from rpnet_b200 import volume as V
    item = V.make_synthetic_volume(args.slices, args.size, 1, 1, seed=0)
    q, s, l = item['query_images'], item['support_images'], item['support_fg']
    dev = torch.device('cuda:0')
    qd, sd, ld = q.to(dev), [[s[0][0].to(dev)]], [[l[0][0].to(dev)]]
    for _ in range(3):
        RG.get_affine_registration(qd, sd, ld, args.iters)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        RG.get_affine_registration(qd, sd, ld, args.iters)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    qh, sh, lh = q.pin_memory(), s[0][0].pin_memory(), l[0][0].pin_memory()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        th, wl, ws = RG.get_affine_registration(qh.to(dev, non_blocking=True), [[sh.to(dev, non_blocking=True)]],
                                                [[lh.to(dev, non_blocking=True)]], args.iters)
        wl.cpu(); ws.cpu()
    ms_e2e = (time.perf_counter() - t0) / args.steps * 1e3
    n_cpu = 4
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    R.get_affine_registration(q[:n_cpu], [[s[0][0][:n_cpu]]], [[l[0][0][:n_cpu]]], args.iters)
    cpu_s = time.perf_counter() - t0
    # algorithmic traffic: every iteration reads the moving and the fixed slice once (fp32)
    bytes_iter = args.slices * args.size * args.size * 8
    print(json.dumps({'metric': 'slices/sec affine registration (%d Adam iterations, %dx%d)' % (args.iters, args.size, args.size),
                      'value': args.slices / (ms * 1e-3), 'unit': 'slices/s', 'ms_per_volume': ms, 'slices': args.slices,
                      'e2e': {'value': args.slices / (ms_e2e * 1e-3), 'unit': 'slices/s'},
                      'effective_gbs': bytes_iter * args.iters / (ms * 1e-3) / 1e9,
                      'cpu_baseline': {'value': n_cpu / cpu_s, 'unit': 'slices/s', 'cores': os.cpu_count(), 'kind': 'port',
                                       'sample': '%d slices, slice by slice like the reference' % n_cpu}}))


if __name__ == '__main__':
    main()
