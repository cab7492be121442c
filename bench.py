#!/usr/bin/env python
"""bench.py — slices/sec of the RP-Net hot path on N B200s (one process per GPU), with roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer|train|cfg4|volume] [--impl reference|library]

A "step" is one pass of the hot path (RP_Net.forward [+ backward + Adam for the train workload]) over one batch of
synthetic CT-like slices per rank.  `value` times it with the inputs already resident in HBM; `e2e` times the same
call from pinned HOST buffers (H2D of every input + D2H of the result inside the timed region).  See DESIGN.md §4.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'slices/sec fwd+bwd @256x256 5-shot T=4; Dice parity vs ref'
# algorithmic conv FLOPs (2*MAC) per encoder image / per cre call at 256 x 256
E_FLOPS = 82_216_747_008
R_FLOPS = 10_115_088_384
FIRST_CONV_FLOPS = 2 * 256 * 256 * 64 * 9           # Conv1.conv.0 (Cin=1): runs in the streaming kernel
CORR_FLOPS = 253_755_392                             # local correlation: runs in its own kernel

WORKLOADS = {
    # BASELINE.json configs[1]
    'infer': dict(name='cfg2: 1-shot 1-way, batch 8x256x256, T=4, forward only', ways=1, shots=1, batch=8, size=256, T=4,
                  train=False),
    # BASELINE.json configs[2] (the configuration the metric is quoted on)
    'train': dict(name='cfg3: 5-shot 1-way, batch 16x256x256, T=4, train step (fwd+bwd+Adam)', ways=1, shots=5, batch=16,
                  size=256, T=4, train=True),
    # BASELINE.json configs[3]: global batch 32 = 4 slices per GPU at 8 GPUs (weak scaling: 4 per rank at every N)
    'cfg4': dict(name='cfg4: 5-shot 4-way, 4x256x256 per GPU (global 32 at 8 GPUs), T=6, train step (fwd+bwd+all-reduce+Adam)', ways=4,
                 shots=5, batch=4, size=256, T=6, train=True),
    # BASELINE.json configs[4]: one 96-slice volume, slices sharded over the ranks (strong scaling), eval loop of test_rpnet.py
    'volume': dict(name='cfg5: 3D volume 96x256x256, sliding-window inference in batches of 16 slices, slice-sharded, T=4', ways=1,
                   shots=1, batch=16, slices=96, size=256, T=4, train=False, volume=True),
}


def model_cfg(T):
    return dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
                n_iter_refinement=T, soft_mask=False, mask_refinement_correlation_radius=5)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()                         # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); power.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------------------------------------------- reference arm
def cpu_reference(wl, steps, warmup, sample_batch):
    """The reference's own CPU data flow (oracle restatement of net/rp_net.py:226-350 incl. the all-pairs bmm
    correlation it executes) on the box's host cores.  /root/reference is pure Python and cannot travel to the GPU
    box, so this is the 'port' kind.  Each step = `sample_batch` slices of the workload."""
    import torch
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = model_cfg(wl['T'])
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(0))
    ep = make_episode(sample_batch, wl['ways'], wl['shots'], wl['size'], seed=0)
    params = None
    if wl['train']:
        params = [k for k, v in sd.items() if v.is_floating_point() and 'running' not in k]
        for k in params:
            sd[k] = sd[k].clone().requires_grad_(True)
        opt = torch.optim.Adam([sd[k] for k in params], lr=1e-5, weight_decay=1e-4)

    def step():
        if wl['train']:
            opt.zero_grad(set_to_none=True)
            out = O.forward(sd, cfg, ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'],
                            training=True, allpairs=True)
            O.train_loss(out, ep['query_labels']).backward()
            opt.step()
        else:
            with torch.no_grad():
                O.forward(sd, cfg, ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'],
                          allpairs=True)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample_batch * steps / dt, dt / steps * 1e3, cores


def cpu_sample(wl):
    """Slices per CPU step: the workload's own per-step batch when the host can hold it.  The 5-shot train workloads keep
    (Wa*Sh + 1) encoder passes + (Wa*Sh + T) all-pairs correlation volumes per slice alive for autograd (~3.5 GB per slice of
    cfg3, ~12 GB per slice of cfg4), so their CPU step is bounded to 4 / 1 slices; throughput is per slice either way."""
    if not wl['train']:
        return min(wl['batch'], 8)
    return 4 if wl['ways'] == 1 else 1


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = cpu_sample(wl)
    steps, warmup = min(args.steps, 3), min(args.warmup, 1)
    val, ms, cores = cpu_reference(wl, steps, warmup, sample)
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'slices/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'ways': wl['ways'], 'shots': wl['shots'], 'size': wl['size'], 'T': wl['T'],
                       'backbone': 'UNet', 'sample': '%d slices per step (B200 arm: %d per GPU)' % (sample, wl['batch'])},
            'cpu_baseline': {'value': val, 'unit': 'slices/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d steps x %d slices of the workload (oracle restatement of the reference CPU '
                                       'data flow incl. all-pairs correlation), torch %d threads' % (steps, sample, cores)},
            'e2e': {'value': val, 'unit': 'slices/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- library arm
def library_baseline(wl, dev, steps=5, warmup=2, batch=None):
    """The "GPU library" baseline of SURVEY §8(d) / BASELINE.md §3: the reference's data flow (the oracle's functional torch
    graph: nn.functional convs / batch_norm / bmm all-pairs correlation / grid_sample, torch autograd and torch.optim.Adam for
    the train workloads) on the SAME B200 with stock PyTorch — cuDNN / cuBLAS kernels — on the same synthetic workload and
    per-step batch.  Modes: 'tf32' = torch's defaults (cuDNN convs may use TF32, matmul fp32), 'tf32_channels_last' = the same
    with channels-last images / weights, 'fp32' = TF32 off everywhere (the arithmetic the CPU reference performs).  Also
    reports how far the TF32 modes' first-iteration logits are from the fp32 mode's (rel-Linf, first step, identical weights): the
    library's own default arithmetic is TF32-class."""
    import torch
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats, to_device
    B = batch or wl['batch']
    cfg = model_cfg(wl['T'])
    ep = to_device(make_episode(B, wl['ways'], wl['shots'], wl['size'], seed=0), dev)
    res, logits = {}, {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for mode in ('tf32', 'tf32_channels_last', 'fp32'):
            torch.backends.cudnn.allow_tf32 = mode != 'fp32'
            torch.backends.cuda.matmul.allow_tf32 = False
            sd = weights.unet_rpnet_state_dict(0)
            if not wl['train']:
                sd = perturb_bn_stats(sd)
            sd = {k: v.to(dev) for k, v in sd.items()}
            d = ep
            if mode == 'tf32_channels_last':
                cl = lambda t: t.contiguous(memory_format=torch.channels_last)
                sd = {k: (cl(v) if v.dim() == 4 else v) for k, v in sd.items()}
                d = dict(ep, supp_imgs=[[cl(t) for t in way] for way in ep['supp_imgs']], qry_imgs=[cl(t) for t in ep['qry_imgs']])
            opt = None
            if wl['train']:
                params = [k for k, v in sd.items() if v.is_floating_point() and 'running' not in k]
                for k in params:
                    sd[k] = sd[k].clone().requires_grad_(True)
                opt = torch.optim.Adam([sd[k] for k in params], lr=1e-5, weight_decay=1e-4)
            last = {}

            def step():
                if wl['train']:
                    opt.zero_grad(set_to_none=True)
                    out = O.forward(sd, cfg, d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], d['appr_query_labels'],
                                    training=True, allpairs=True)
                    O.train_loss(out, d['query_labels']).backward()
                    opt.step()
                else:
                    with torch.no_grad():
                        out = O.forward(sd, cfg, d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], d['appr_query_labels'],
                                        allpairs=True)
                last['logits'] = out['refinement'][0].detach()        # iteration 0: before any hard-mask feedback
            try:
                step()
                logits[mode] = last['logits'].float().clone()        # first step: identical weights in every mode
                for _ in range(warmup - 1):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                res[mode] = {'value': B / (ms * 1e-3), 'unit': 'slices/s', 'ms_per_step': ms, 'batch': B}
            except RuntimeError as e:                                  # e.g. out of memory in one mode: report, keep going
                res[mode] = {'unavailable': str(e).splitlines()[0][:200]}
            del sd, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    if 'fp32' in logits:
        ref = logits['fp32']
        for mode in ('tf32', 'tf32_channels_last'):
            if mode in logits and 'value' in res.get(mode, {}):
                res[mode]['logits_rel_linf_vs_fp32_mode'] = ((logits[mode] - ref).abs().max() / ref.abs().max()).item()
    res['what'] = ('oracle functional graph (the reference data flow incl. all-pairs bmm correlation%s) with stock PyTorch %s on this GPU, '
                   '%d timed steps after %d warm-up, CUDA events' % (', torch autograd + torch.optim.Adam' if wl['train'] else '',
                                                                      torch.__version__, steps, warmup))
    return res


def run_library(args, wl):
    """`--impl library`: one JSON line with the stock-PyTorch-on-B200 numbers (rank 0 only)."""
    import torch
    if int(os.environ.get('RANK', '0')) != 0:
        return
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(dev)
    lib = library_baseline(wl, dev, steps=min(args.steps, 5), warmup=max(min(args.warmup, 2), 1))
    best = max((v for v in lib.values() if isinstance(v, dict) and 'value' in v), key=lambda v: v['value'], default=None)
    line = {'impl': 'library', 'metric': METRIC, 'value': best['value'] if best else None, 'unit': 'slices/s', 'n_gpus': 1,
            'ms_per_step': best['ms_per_step'] if best else None, 'higher_is_better': True, 'data': 'synthetic',
            'dtype': 'fp32 tensors; cuDNN convs TF32 (torch default) or fp32',
            'config': {'workload': wl['name'], 'ways': wl['ways'], 'shots': wl['shots'], 'batch_per_gpu': wl['batch'], 'size': wl['size'],
                       'T': wl['T'], 'backbone': 'UNet'}, 'library_baseline': lib}
    print(json.dumps(line), flush=True)


def parity_field(sd, wl, dev, precision, cfg_extra=None):
    """1-2 slices of the workload's shape through the CPU oracle and through the B200 path (outside every timed region) with
    the state_dict the timed run ended with: rel-Linf / margin error / argmax mismatch / Dice-vs-reference of the last
    refinement iteration (rpnet_b200/parity.py)."""
    import torch
    from oracle import rpnet_oracle as O
    from rpnet_b200 import parity
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.synthetic import make_episode, to_device
    B = 1 if wl['ways'] > 1 else 2
    cfg = model_cfg(wl['T'])
    cfg.update(cfg_extra or {})
    ep = make_episode(B, wl['ways'], wl['shots'], wl['size'], seed=4242)
    net = RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    net.load_state_dict(sd)
    net = net.to(dev)
    d = to_device(ep, dev)
    a = (ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'])
    T = wl['T']
    # iterations i >= 1 of the oracle consume the recurrent masks derived from OUR logits of iteration i - 1 (teacher forcing): the
    # hard threshold (net/rp_net.py:310) is discontinuous, a single near-tie pixel would otherwise change the next input
    over = lambda lg: {i: O.recurrent_mask(lg[i - 1].float().cpu(), cfg) for i in range(1, T)}
    loss = None
    if wl['train']:
        from rpnet_b200.train import TrainStep
        net.train()
        ts = TrainStep(net)
        loss = ts.forward_backward(d)
        torch.cuda.synchronize()
        ours = [ts.last['logits'][i].cpu() for i in range(T)]
        with torch.no_grad():
            out = O.forward({k: v.clone() for k, v in sd.items()}, cfg, *a, training=True, mask_override=over(ours))
        ref_loss = O.train_loss(out, ep['query_labels']).item()
    else:
        net.eval()
        with torch.no_grad():
            o = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
            ours = [o['refinement'][i].cpu() for i in range(T)]
            out = O.forward({k: v.clone() for k, v in sd.items()}, cfg, *a, mask_override=over(ours))
    per_iter = [parity.compare_logits(ours[i], out['refinement'][i]) for i in range(T)]
    r = {k: max(p[k] for p in per_iter) for k in ('rel_linf', 'margin_rel_err', 'argmax_mismatch')}
    r['dice_vs_ref'] = min(p['dice_vs_ref'] for p in per_iter)
    r['margin_median'] = per_iter[-1]['margin_median']
    if loss is not None:
        r['loss'] = {'b200': loss.item(), 'oracle': ref_loss}
    got = ours[T - 1]
    tgt = ep['query_labels'] > 0
    dice = lambda m: (2.0 * (m & tgt).sum().item() / max(m.sum().item() + tgt.sum().item(), 1))
    r['dice_vs_ground_truth'] = {'b200': dice(parity.fg_mask(got)), 'oracle': dice(parity.fg_mask(out['refinement'][T - 1]))}
    r['what'] = ('%d slice(s) of the workload shape (%s mode), worst of the %d refinement iterations (oracle teacher-forced with the masks '
                 'of the B200 path), B200 path (%s) vs the fp32 CPU oracle on the state_dict the timed run ended with'
                 % (B, 'train' if wl['train'] else 'eval', T, precision))
    r['tolerance'] = {'rel_linf': 1e-3, 'margin_rel_err': 1e-3}
    return r


def ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of the current build
    (profiles/r02_ncu_conv_igemm_traffic.json, written by tools/summarize_ncu.py).  None when no capture is committed."""
    path = os.path.join(ROOT, 'profiles', 'r02_ncu_conv_igemm_traffic.json')
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


# ------------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'library'])
    ap.add_argument('--workload', default=os.environ.get('RPNET_BENCH_WORKLOAD', 'train'), choices=sorted(WORKLOADS),
                    help="'train' = BASELINE.json configs[2], the configuration the metric is quoted on (default); 'infer' = configs[1]")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-library-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-fast-mode', action='store_true', help="skip the extra timing of `b200_precision: fp16` (TF32-class arithmetic)")
    ap.add_argument('--no-cuda-graph', action='store_true', help='inference workloads: launch every kernel from Python')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        return run_reference(args, wl)
    if args.impl == 'library':
        return run_library(args, wl)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (B200); there is no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()

    from rpnet_b200 import engine, ops
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.synthetic import fitted_state_dict, make_episode, to_device

    cfg = model_cfg(wl['T'])
    torch.manual_seed(0)                            # random-init weights of the reference architecture (test_rpnet.py:8-10)
    net = RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    net = net.to(dev)
    precision = engine.precision_of(cfg)
    if not wl['train']:
        # eval: BatchNorm running statistics fitted to the synthetic data (lr = 0: weights stay at their random initialisation), so
        # that BN folding is exercised and the logits are not pinned at the 20 * cos cap (the parity field below means something)
        net.load_state_dict(fitted_state_dict(net, lambda i: to_device(make_episode(2, wl['ways'], wl['shots'], wl['size'], seed=9000 + i), dev),
                                              steps=30, lr=0.0))
        net = net.to(dev)
    B = wl['batch']
    if wl.get('volume'):
        # one volume, this rank's contiguous slice range (strong scaling); a "step" = one pass over the rank's slices
        from rpnet_b200 import volume as rp_volume
        from rpnet_b200.train import shard_range
        item = rp_volume.make_synthetic_volume(wl['slices'], wl['size'], wl['ways'], wl['shots'], seed=0)
        lo, hi = shard_range(wl['slices'], rank, world)
        cutv = lambda t: t[lo:hi].contiguous()
        ep = {'supp_imgs': [[cutv(t) for t in way] for way in item['support_images']],
              'fore_mask': [[cutv(t) for t in way] for way in item['support_fg']],
              'back_mask': [[cutv(t) for t in way] for way in item['support_bg']], 'qry_imgs': [cutv(item['query_images'])],
              'query_labels': cutv(item['query_labels']), 'appr_query_labels': cutv(item['appr_query_labels'])}
        B = hi - lo                                   # slices this rank processes per step
    else:
        ep = make_episode(B, wl['ways'], wl['shots'], wl['size'], seed=1000 * rank)     # per-rank shard (weak scaling)

    # pinned host buffers (e2e) and resident device copies (value)
    def pin(t):
        return t.contiguous().pin_memory()
    host = {'supp_imgs': [[pin(t) for t in way] for way in ep['supp_imgs']],
            'fore_mask': [[pin(t) for t in way] for way in ep['fore_mask']],
            'back_mask': [[pin(t) for t in way] for way in ep['back_mask']],
            'qry_imgs': [pin(t) for t in ep['qry_imgs']], 'appr_query_labels': pin(ep['appr_query_labels']),
            'query_labels': pin(ep['query_labels'])}

    def upload(h):
        mv = lambda t: t.to(dev, non_blocking=True)
        return {'supp_imgs': [[mv(t) for t in way] for way in h['supp_imgs']],
                'fore_mask': [[mv(t) for t in way] for way in h['fore_mask']],
                'back_mask': [[mv(t) for t in way] for way in h['back_mask']],
                'qry_imgs': [mv(t) for t in h['qry_imgs']], 'appr_query_labels': mv(h['appr_query_labels']),
                'query_labels': mv(h['query_labels'])}

    def nbytes(x):
        if isinstance(x, (list, tuple)):
            return sum(nbytes(t) for t in x)
        return x.numel() * x.element_size()
    h2d = sum(nbytes(v) for v in host.values())
    resident = upload(host)
    out_host = (torch.empty(B, wl['size'], wl['size'], dtype=torch.uint8) if wl.get('volume') else
                torch.empty(B, 1 + wl['ways'], wl['size'], wl['size'], dtype=torch.float32)).pin_memory()

    if wl['train']:
        from rpnet_b200 import train as rp_train
        stepper = rp_train.TrainStep(net, world_size=world)
        net.train()

        def step(d):
            return stepper.step(d)                   # returns the loss tensor (device)
    elif wl.get('volume'):
        net.eval()
        net.enable_cuda_graph(not args.no_cuda_graph)

        def step(d):
            r = rp_volume.segment_volume(net, d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'][0], d['appr_query_labels'],
                                         batch_size=wl['batch'], rank=0, world=1)       # d already holds this rank's slices
            return r['mask']
    else:
        net.eval()
        net.enable_cuda_graph(not args.no_cuda_graph)

        def step(d):
            with torch.no_grad():
                return net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], query_labels=d['query_labels'],
                           appr_query_labels=d['appr_query_labels'])['output']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=None):
        barrier()
        ops.set_profiler(profile)
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ops.set_profiler(None)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), ops.LAUNCHES - l0

    for _ in range(args.warmup):
        step(resident)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    prof = {}
    graphed = (not wl['train']) and not args.no_cuda_graph
    if graphed:
        # the timed region replays CUDA graphs (no per-kernel events possible); the per-kernel CUDA-event timing for the
        # roofline comes from a second pass of the same K steps with every kernel launched from Python
        ms_total, _ = timed(lambda: step(resident), args.steps)
        clk = clocks.stop() if rank == 0 else None
        net.enable_cuda_graph(False)
        for _ in range(2):
            step(resident)
        _, launches = timed(lambda: step(resident), args.steps, prof)
        net.enable_cuda_graph(True)
        step(resident)
    else:
        ms_total, launches = timed(lambda: step(resident), args.steps, prof)
        clk = clocks.stop() if rank == 0 else None

    # ---- e2e: pinned host inputs -> H2D -> step -> D2H of the result, every step
    def e2e_step():
        d = upload(host)
        r = step(d)
        if wl['train']:
            r.cpu()
        else:
            out_host.copy_(r, non_blocking=True)
            torch.cuda.current_stream().synchronize()
    for _ in range(2):
        e2e_step()
    ms_e2e, _ = timed(e2e_step, args.steps)
    d2h = 4 if wl['train'] else out_host.numel() * out_host.element_size()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (tcgen05 implicit-GEMM conv), from CUDA events recorded on the launching
    # stream around every launch inside the timed region
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        peaks = {}                                   # fallback peaks of B200_PROFILING.md ("of fallback")
    kern = {}
    for name, evs in prof.items():
        kern[name] = {'launches': len(evs) // args.steps, 'ms_per_step': sum(a.elapsed_time(b) for a, b, _ in evs) / args.steps,
                      'work_per_step': sum(w for _, _, w in evs) / args.steps}
    # dominant kernel: conv_igemm_kernel (forward convs, and in training also every data-gradient conv = the same kernel on
    # bf16 operands).  Algorithmic FLOPs: forward n_img*(E-first) + n_cre*(R-corr); training runs that twice
    # through this kernel (forward + dgrad) and once through conv_wgrad_kernel (reported in `kernels`).
    n_img = (wl['ways'] * wl['shots'] + 1) * B
    n_cre = (wl['ways'] * wl['shots'] + wl['T']) * B
    scale = (wl['size'] / 256.0) ** 2
    fwd_algo = (n_img * (E_FLOPS - FIRST_CONV_FLOPS) + n_cre * (R_FLOPS - CORR_FLOPS)) * scale
    conv_ms = kern['conv_igemm']['ms_per_step']
    algo = fwd_algo * (2 if wl['train'] else 1)
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    achieved = algo / (conv_ms * 1e-3) / 1e12
    # executed tensor-core FLOPs: the split-fp16 encoder forward runs three fp16 passes (hi.Wh + lo.Wh + hi.Wl) of its algorithmic
    # FLOPs; 'split8' one fp16 pass + two e4m3 passes (lo8.Wh8 + x8.Wl8), the e4m3 ones at twice the MMA rate: `executed_flops`
    # counts all of them, `executed_fp16_equiv` weighs an e4m3 FLOP as half (= units of fp16 tensor time, what the bf16 peak prices)
    enc_fwd = n_img * (E_FLOPS - FIRST_CONV_FLOPS) * scale
    executed = algo + (2 * enc_fwd if precision in ('split', 'split8') else 0)
    executed_eq = algo + (2 * enc_fwd if precision == 'split' else (enc_fwd if precision == 'split8' else 0))
    burst_tf = peaks.get('bf16_tflops', 1590.0)
    roofline = {'bound': 'tensor', 'kernel': 'conv_igemm_kernel (tcgen05 implicit GEMM: forward%s)' % (' + data-gradient convs' if wl['train'] else ''),
                'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': None,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step), of measured' if peaks else 'fallback 1400 TF/s, of fallback',
                'frac_of_burst_peak': achieved / burst_tf,
                'algorithmic_flops_per_step': algo, 'launches_per_step': kern['conv_igemm']['launches'], 'kernel_ms_per_step': conv_ms,
                'executed_flops_per_step': executed, 'executed_tflops': executed / (conv_ms * 1e-3) / 1e12,
                'executed_fp16_equiv_flops_per_step': executed_eq, 'executed_fp16_equiv_tflops': executed_eq / (conv_ms * 1e-3) / 1e12,
                'executed_frac': executed_eq / (conv_ms * 1e-3) / 1e12 / peak_tf,
                'executed_frac_of_burst_peak': executed_eq / (conv_ms * 1e-3) / 1e12 / burst_tf,
                'precision': precision + {'split': ': split-fp16 encoder forward = 3 fp16 tensor-core passes per algorithmic FLOP (fp32-class result); '
                                                   '`achieved` counts the algorithmic FLOPs once',
                                          'split8': ': encoder forward = 1 fp16 pass + 2 e4m3 correction passes at twice the MMA rate = 2 units of fp16 '
                                                    'tensor time per algorithmic FLOP (fp32-class result); `achieved` counts the algorithmic FLOPs '
                                                    'once, `executed_frac*` the fp16-equivalent tensor time against the bf16 peak'}.get(precision, ''),
                'kernel_share_of_step': conv_ms / (ms_total / args.steps)}
    # DRAM traffic per launch of the kernel: `dram__bytes_read.sum + dram__bytes_write.sum` from the committed `ncu --set full`
    # capture of the current build over the launches of one step (profiles/r02_ncu_conv_igemm_traffic.json)
    tr = ncu_traffic()
    if tr and tr.get('workload') == args.workload and tr.get('precision') == precision:
        roofline['traffic'] = tr['dram_bytes_per_launch']
        roofline['traffic_capture'] = {k: tr[k] for k in ('launches', 'dram_bytes_per_step', 'algorithmic_bytes_per_step', 'source') if k in tr}
    if graphed:
        roofline['note'] = 'kernel times from a second pass without CUDA-graph replay; value / ms_per_step from graph replay'
    if 'conv_wgrad' in kern:
        wg_ms = kern['conv_wgrad']['ms_per_step']
        roofline['wgrad'] = {'kernel': 'conv_wgrad_kernel (tcgen05, MN-major operands, split-K)', 'achieved': fwd_algo / (wg_ms * 1e-3) / 1e12,
                             'frac': fwd_algo / (wg_ms * 1e-3) / 1e12 / peak_tf, 'kernel_ms_per_step': wg_ms,
                             'kernel_share_of_step': wg_ms / (ms_total / args.steps)}
    hbm = peaks.get('hbm_gbs', 6650.0)
    stream_kernels = {k: {'launches': v['launches'], 'ms_per_step': v['ms_per_step'],
                          'gbs': v['work_per_step'] / (v['ms_per_step'] * 1e-3) / 1e9 if k not in ('conv_igemm', 'conv_wgrad') else None,
                          'frac_hbm': v['work_per_step'] / (v['ms_per_step'] * 1e-3) / 1e9 / hbm if k not in ('conv_igemm', 'conv_wgrad') else None}
                      for k, v in kern.items()}

    cpu = None
    if not args.no_cpu_baseline and world == 1:          # reported at N=1 only (rank 0's host cores)
        sample = cpu_sample(wl)
        val, ms, cores = cpu_reference(wl, 2, 1, sample)
        cpu = {'value': val, 'unit': 'slices/s', 'cores': cores, 'kind': 'port',
               'sample': '2 steps x %d slices of the workload after 1 warm-up (oracle restatement of the reference CPU data '
                         'flow incl. all-pairs correlation), %.0f ms/step' % (sample, ms)}
    lib = None
    if not args.no_library_baseline and world == 1 and not wl.get('volume'):
        lib = library_baseline(wl, dev)
    par = None
    if not args.no_parity and world == 1:
        par = parity_field({k: v.detach().cpu().clone() for k, v in net.state_dict().items()}, wl, dev, precision)
    fast = None
    if not args.no_fast_mode and world == 1 and precision in ('split', 'split8') and wl['train']:
        # the same workload with `b200_precision: fp16`: single-term fp16 operands = the 11-bit significand of the library's
        # default TF32 convs.  Reported beside the headline (which is the fp32-class split mode): what relaxing the arithmetic to
        # the library's own default precision buys, and what it costs in parity.
        cfg16 = dict(cfg, b200_precision='fp16')
        torch.manual_seed(0)
        net16 = RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg16).to(dev).train()
        st16 = rp_train.TrainStep(net16, world_size=1)
        for _ in range(args.warmup):
            st16.step(resident)
        ms16, _ = timed(lambda: st16.step(resident), args.steps)
        fast = {'b200_precision': 'fp16', 'value': B / (ms16 / args.steps * 1e-3), 'unit': 'slices/s', 'ms_per_step': ms16 / args.steps,
                'what': 'same workload, single-term fp16 conv operands (TF32-class, like the library default); device-timed, inputs resident'}
        if not args.no_parity:
            p16 = parity_field({k: v.detach().cpu().clone() for k, v in net16.state_dict().items()}, dict(wl), dev, 'fp16', cfg_extra={'b200_precision': 'fp16'})
            fast['parity'] = {k: p16[k] for k in ('rel_linf', 'margin_rel_err', 'argmax_mismatch', 'dice_vs_ref')}
        del net16, st16

    ms_step = ms_total / args.steps
    total_units = wl['slices'] if wl.get('volume') else world * B
    line = {'metric': METRIC, 'value': total_units / (ms_step * 1e-3), 'unit': 'slices/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong' if wl.get('volume') else 'weak', 'vs_baseline': None,
            'dtype': {'split': 'split-fp16 (hi + lo planes, 3 tensor-core passes) encoder forward, fp16 cre convs',
                      'split8': 'fp16 main term + e4m3 first-order corrections (hi.Wh + 2^-15 (lo8.Wh8 + x8.Wl8): fp16 + c8 planes, 1 fp16 + 2 e4m3 '
                                'tensor-core passes) encoder forward, fp16 cre convs'}.get(precision, 'fp16 forward operands')
                     + '; fp32 accumulate; activations stored fp16 (hi [+ lo]), activation gradients bf16, dgrad / wgrad operands bf16, '
                       'BatchNorm statistics fp64, parameters / weight gradients / Adam / losses fp32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'ways': wl['ways'], 'shots': wl['shots'], 'batch_per_gpu': B,
                       'global_batch': total_units, 'size': wl['size'], 'T': wl['T'], 'backbone': 'UNet',
                       'parallelism': 'dp%d (slices sharded, no data-path collective%s)' % (
                           world, '; NCCL grad all-reduce' if wl['train'] else ''),
                       'l2': 'per-step activation traffic (>1 GB) exceeds the 126 MB L2; no explicit flush'},
            'clocks': clk,
            'e2e': {'value': total_units / (ms_e2e / args.steps * 1e-3), 'unit': 'slices/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'library_baseline': lib, 'parity': par,
            'precision': precision, 'tf32_class_mode': fast, 'kernels': stream_kernels}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
