#!/usr/bin/env python
"""N-rank data-parallel parity of the train step: run under torchrun on N GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py

Every rank owns a contiguous shard of the global batch (rpnet_b200.train.shard_range).  Checks, on every rank:
  (1) the bucketed NCCL all-reduce of the flat gradient buffer (side stream, overlapped with the backward) equals the sum of
      the per-rank gradients computed WITHOUT communication (gathered with all_gather) — i.e. the N-rank gradient is the
      reference run independently on each rank's shard and summed; the 1/N average is folded into Adam;
  (2) after one Adam step all ranks hold bit-identical parameters;
  (3) BatchNorm running statistics stay rank-local (the reference has no SyncBN)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.synthetic import make_episode, to_device
    from rpnet_b200.train import TrainStep, shard_range
    cfg = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=2,
               soft_mask=False, mask_refinement_correlation_radius=5)
    B_global, ways, shots, size = 4 * world, 1, 2, 128
    lo, hi = shard_range(B_global, rank, world)
    ep = make_episode(B_global, ways, shots, size, seed=3)
    cut = lambda t: t[lo:hi].contiguous()
    shard = {'supp_imgs': [[cut(t) for t in way] for way in ep['supp_imgs']], 'fore_mask': [[cut(t) for t in way] for way in ep['fore_mask']],
             'back_mask': [[cut(t) for t in way] for way in ep['back_mask']], 'qry_imgs': [cut(t) for t in ep['qry_imgs']],
             'query_labels': cut(ep['query_labels']), 'appr_query_labels': cut(ep['appr_query_labels'])}
    d = to_device(shard, dev)

    def build():
        torch.manual_seed(0)                          # the same initial weights on every rank and for every replica
        net = RP_Net(cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
        return net.to(dev).train()

    # (a) local gradients, no communication
    net_a = build()
    ts_a = TrainStep(net_a, world_size=1)
    ts_a.forward_backward(d)
    g_local = ts_a.eng.flat.grad.clone()
    gathered = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    g_sum = torch.stack(gathered).sum(0)
    # (b) the data-parallel step
    net_b = build()
    ts_b = TrainStep(net_b, world_size=world, lr=1e-4)
    ts_b.forward_backward(d)
    g_ddp = ts_b.eng.flat.grad.clone()
    rel = ((g_ddp - g_sum).norm() / g_sum.norm()).item()
    # (a) and (b) are two executions of the same backward; it is deterministic (fp32 partials accumulated in fp64, see
    # include/rpnet_b200.h), so the only difference left is the order in which NCCL adds the N ranks' fp32 gradients
    assert rel < 1e-5, 'rank %d: all-reduced gradient differs from the sum of per-rank gradients: %.3e' % (rank, rel)
    # the collective wiring itself (bucket ranges, side stream, event ordering) is also checked exactly on known data
    flat, buckets = ts_b.eng.flat, ts_b.buckets
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    flat.grad.copy_(torch.randint(-1000, 1000, (flat.numel,), generator=gen, device=dev).float())
    for i in range(len(buckets.ranges)):
        buckets.ready(i)
    buckets.finish()
    want = sum(torch.randint(-1000, 1000, (flat.numel,), generator=torch.Generator(device=dev).manual_seed(1234 + r), device=dev).float()
               for r in range(world))
    assert torch.equal(flat.grad, want), 'rank %d: bucketed all-reduce is not the exact sum over ranks' % rank
    ts_b.step(d)
    torch.cuda.synchronize()
    p = ts_b.eng.flat.param
    ref = p.clone()
    dist.broadcast(ref, src=0)
    # all ranks applied the same averaged gradient to the same parameters
    same = torch.equal(p, ref)
    rm = net_b.encoder.Conv1.conv[1].running_mean.clone()
    rms = [torch.empty_like(rm) for _ in range(world)]
    dist.all_gather(rms, rm)
    local_bn = world == 1 or not torch.equal(rms[0], rms[-1])
    ok = torch.tensor([int(same), int(local_bn)], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print('ddp_check world=%d: grad rel diff %.2e, identical params after Adam: %s, BN stats rank-local: %s'
              % (world, rel, bool(ok[0].item()), bool(ok[1].item())), flush=True)
    assert ok[0].item() == 1, 'parameters diverged across ranks'
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
