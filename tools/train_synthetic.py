#!/usr/bin/env python
"""Train -> evaluate on synthetic episodes, end to end through the reference-shaped API: `TrainStep` (forward + backward + Adam,
the reconstructed loss of SURVEY §3.5) for a number of steps, then `net.eval()` and the eval forward on held-out episodes with
Dice against the synthetic ground truth.  The reference ships no train script; this is the loop a user would write.

    python tools/train_synthetic.py [--steps 200] [--batch 8] [--shots 1] [--size 128] [--T 2] [--lr 1e-4]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/train_synthetic.py     # data parallel
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--batch', type=int, default=8, help='query slices per GPU and step')
    ap.add_argument('--shots', type=int, default=1)
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--T', type=int, default=2)
    ap.add_argument('--lr', type=float, default=1e-4)
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from net.model import model_factory
    from rpnet_b200.synthetic import make_episode, to_device
    from rpnet_b200.train import TrainStep
    from rpnet_b200.volume import dice_from_sums, dice_sums
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cfg = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=args.T,
               soft_mask=False, mask_refinement_correlation_radius=5)
    torch.manual_seed(0)                                            # identical initial weights on every rank
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg).to(dev)

    def evaluate(tag):
        net.eval()
        sums = torch.zeros(3, device=dev)
        with torch.no_grad():
            for k in range(4):
                d = to_device(make_episode(args.batch, 1, args.shots, args.size, seed=100000 + 17 * k), dev)
                out = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
                sums += dice_sums(out['output'].argmax(1) > 0, d['query_labels'] > 0)
        if rank == 0:
            print('%s: Dice vs synthetic ground truth on held-out episodes = %s' % (tag, dice_from_sums(sums.cpu())))

    evaluate('before training')
    net.train()
    ts = TrainStep(net, world_size=world, lr=args.lr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(args.steps):
        d = to_device(make_episode(args.batch, 1, args.shots, args.size, seed=1 + step * world + rank), dev)   # a new episode per rank and step
        loss = ts.step(d)
        if rank == 0 and (step % 20 == 0 or step == args.steps - 1):
            print('step %4d  loss %.4f' % (step, loss.item()))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print('%d steps in %.1f s: %.0f query slices/s incl. synthetic data generation on the host'
              % (args.steps, dt, args.steps * args.batch * world / dt))
    evaluate('after training')
    if args.out and rank == 0:
        torch.save(net.state_dict(), args.out)                     # loads into the reference's RP_Net as well (same keys)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
