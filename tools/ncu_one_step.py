#!/usr/bin/env python
"""One step of a bench.py workload for ncu captures (run under gpurun, never timed):

    ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s <launches of the warm-up steps> -c <launches per step> \
        -o gpurun_out/prof python tools/ncu_one_step.py --workload train --warmup 1

Without ncu it prints (and writes to gpurun_out/one_step.json) how many launches of each kernel family one step makes and the
ALGORITHMIC bytes of every conv_igemm launch (its operand and result tensors, each counted once), which is what
tools/summarize_ncu.py compares the measured DRAM bytes with."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='train')
    ap.add_argument('--warmup', type=int, default=1)
    args = ap.parse_args()
    import bench
    from rpnet_b200 import engine, ops
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.synthetic import make_episode, to_device
    wl = bench.WORKLOADS[args.workload]
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    cfg = bench.model_cfg(wl['T'])
    net = RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg).to(dev)
    d = to_device(make_episode(wl['batch'], wl['ways'], wl['shots'], wl['size'], seed=0), dev)
    if wl['train']:
        from rpnet_b200.train import TrainStep
        net.train()
        ts = TrainStep(net)
        step = lambda: ts.step(d)
    else:
        net.eval()

        def step():
            with torch.no_grad():
                net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
    # algorithmic bytes of the conv_igemm launches: every tensor argument of the wrappers that end in conv_igemm_kernel
    conv_bytes = []

    def wrap(name):
        fn = getattr(ops, name)

        def inner(*a, **kw):
            ts_ = [t for t in list(a) + list(kw.values()) if torch.is_tensor(t) and t.dtype in (torch.float16, torch.bfloat16, torch.float32, torch.uint8)
                   and t.numel() > 4096]
            conv_bytes.append(sum(t.numel() * t.element_size() for t in ts_))
            return fn(*a, **kw)
        setattr(ops, name, inner)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    for nm in ('conv_igemm', 'conv_split', 'conv_bnstats', 'conv_dgrad', 'conv_cos', 'upconv_dgrad'):
        wrap(nm)
    l0 = ops.LAUNCHES
    prof = {}
    ops.set_profiler(prof)
    step()
    torch.cuda.synchronize()
    ops.set_profiler(None)
    # upconv_fwd_bnstats launches four phase convs from inside ops (not wrapped): count them from the profiler store
    n_conv = len(prof.get('conv_igemm', []))
    out = {'workload': args.workload, 'precision': engine.precision_of(cfg), 'launches_per_step': ops.LAUNCHES - l0,
           'conv_igemm_launches_per_step': n_conv, 'conv_wrapped_calls': len(conv_bytes), 'conv_algorithmic_bytes_per_step': sum(conv_bytes),
           'kernels': {k: len(v) for k, v in prof.items()}}
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'one_step_%s.json' % args.workload), 'w') as f:
        json.dump(out, f)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
