#!/usr/bin/env python
"""bench.py — slices/sec of the RP-Net hot path on N B200s (one process per GPU), with roofline + CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer|train] [--impl reference]

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


def run_reference(args, wl):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = 2 if not wl['train'] else 1
    steps, warmup = min(args.steps, 5), min(args.warmup, 1)
    val, ms, cores = cpu_reference(wl, steps, warmup, sample)
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'slices/s', 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'sample': '%d slices per step' % sample},
            'cpu_baseline': {'value': val, 'unit': 'slices/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d steps x %d slices of the workload (oracle restatement of the reference CPU '
                                       'data flow incl. all-pairs correlation), torch %d threads' % (steps, sample, cores)},
            'e2e': {'value': val, 'unit': 'slices/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=os.environ.get('RPNET_BENCH_WORKLOAD', 'train'), choices=sorted(WORKLOADS),
                    help="'train' = BASELINE.json configs[2], the configuration the metric is quoted on (default); 'infer' = configs[1]")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cuda-graph', action='store_true', help='inference workloads: launch every kernel from Python')
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == 'reference':
        return run_reference(args, wl)
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

    from rpnet_b200 import ops
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats

    cfg = model_cfg(wl['T'])
    torch.manual_seed(0)                            # random-init weights of the reference architecture (test_rpnet.py:8-10)
    net = RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    if not wl['train']:
        perturb_bn_stats(net.state_dict())          # eval: non-trivial running statistics, so that BN folding is exercised
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
    roofline = {'bound': 'tensor', 'kernel': 'conv_igemm_kernel (tcgen05 implicit GEMM: forward%s)' % (' + data-gradient convs' if wl['train'] else ''),
                'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': None,
                'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step), of measured' if peaks else 'fallback 1400 TF/s, of fallback',
                'algorithmic_flops_per_step': algo, 'launches_per_step': kern['conv_igemm']['launches'], 'kernel_ms_per_step': conv_ms,
                'executed_tflops': kern['conv_igemm']['work_per_step'] / (conv_ms * 1e-3) / 1e12,
                'kernel_share_of_step': conv_ms / (ms_total / args.steps)}
    # DRAM traffic of the kernel from the committed `ncu --set full` capture (profiles/r01_ncu_conv_pair_uc4a.csv): one
    # representative launch, Up_conv4.conv.0 forward at the cfg3 shape (96 x 64 x 64, 512 -> 256 channels), CTA-pair kernel
    roofline['traffic'] = 405.076224e6 + 171.506176e6
    roofline['traffic_launch'] = {'layer': 'Up_conv4.conv.0 forward, 96x64x64, 512->256', 'dram_bytes': 405.076224e6 + 171.506176e6,
                                  'algorithmic_bytes': 96 * 64 * 64 * (512 + 256) * 2 + 9 * 512 * 256 * 2, 'duration_us': 558.7,
                                  'tensor_pipe_active_pct': 97.2, 'source': 'profiles/r01_ncu_conv_pair_uc4a.csv'}
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
        sample = 2 if not wl['train'] else 1
        val, ms, cores = cpu_reference(wl, 3, 1, sample)
        cpu = {'value': val, 'unit': 'slices/s', 'cores': cores, 'kind': 'port',
               'sample': '3 steps x %d slices of the workload after 1 warm-up (oracle restatement of the reference CPU data '
                         'flow incl. all-pairs correlation), %.0f ms/step' % (sample, ms)}

    ms_step = ms_total / args.steps
    total_units = wl['slices'] if wl.get('volume') else world * B
    line = {'metric': METRIC, 'value': total_units / (ms_step * 1e-3), 'unit': 'slices/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong' if wl.get('volume') else 'weak', 'vs_baseline': None,
            'dtype': 'fp16 operands, fp32 accumulate (tensor-core convs); fp32 elsewhere', 'data': 'synthetic',
            'config': {'workload': wl['name'], 'ways': wl['ways'], 'shots': wl['shots'], 'batch_per_gpu': B,
                       'global_batch': total_units, 'size': wl['size'], 'T': wl['T'], 'backbone': 'UNet',
                       'parallelism': 'dp%d (slices sharded, no data-path collective%s)' % (
                           world, '; NCCL grad all-reduce' if wl['train'] else ''),
                       'l2': 'per-step activation traffic (>1 GB) exceeds the 126 MB L2; no explicit flush'},
            'clocks': clk,
            'e2e': {'value': total_units / (ms_e2e / args.steps * 1e-3), 'unit': 'slices/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'kernels': stream_kernels}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
