#!/usr/bin/env python
"""Parity report of the B200 path against the CPU oracle on one synthetic episode (run on a GPU box):

    python tools/parity_report.py --ways 1 --shots 5 --batch 2 --size 256 --T 4 [--train] [--fp64] [--pretrain 100]

Prints, per refinement iteration, the rel-Linf of the logits, the error of the decision margin relative to the margin range,
the argmax mismatch fraction and Dice(pred mask, oracle mask); with --train also the per-parameter gradient rel-L2 against the
fp32 (and optionally fp64) oracle + torch autograd.  --pretrain N first runs N Adam steps (lr 1e-3) on other episodes, so that
the weights are not at their random initialisation (unsaturated logits, better-conditioned gradients).  One JSON line at the end."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def model_cfg(T, precision=None):
    cfg = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=T,
               soft_mask=False, mask_refinement_correlation_radius=5)
    if precision:
        cfg['b200_precision'] = precision
    return cfg


def logits_report(got, ref):
    """got / ref: lists of [B, P, H, W] logits.  See rpnet_b200/parity.py."""
    from rpnet_b200 import parity
    return [parity.compare_logits(g, r) for g, r in zip(got, ref)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ways', type=int, default=1)
    ap.add_argument('--shots', type=int, default=1)
    ap.add_argument('--batch', type=int, default=2)
    ap.add_argument('--size', type=int, default=128)
    ap.add_argument('--T', type=int, default=2)
    ap.add_argument('--seed', type=int, default=7)
    ap.add_argument('--train', action='store_true')
    ap.add_argument('--fp64', action='store_true', help='also run the oracle in fp64 (conditioning of the fixture)')
    ap.add_argument('--pretrain', type=int, default=0)
    ap.add_argument('--precision', default=None)
    ap.add_argument('--calibrate', action='store_true', help='eval: BatchNorm running statistics = batch statistics of another episode')
    args = ap.parse_args()
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200.nn.rp_net import RP_Net
    from rpnet_b200.synthetic import make_episode, to_device
    from rpnet_b200.train import TrainStep
    dev = torch.device('cuda:0')
    cfg = model_cfg(args.T, args.precision)
    sd = weights.unet_rpnet_state_dict(0)
    net = RP_Net(cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    net.load_state_dict(sd)
    net = net.to(dev).train()
    if args.pretrain or args.calibrate:
        ts = TrainStep(net, lr=1e-3 if args.pretrain else 0.0, weight_decay=0.0)
        for i in range(max(args.pretrain, 20 if args.calibrate else 0)):
            ts.step(to_device(make_episode(args.batch, args.ways, args.shots, args.size, seed=1000 + i), dev))
        torch.cuda.synchronize()
        sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
        net = RP_Net(cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
        net.load_state_dict(sd)
        net = net.to(dev).train()
    ep = make_episode(args.batch, args.ways, args.shots, args.size, seed=args.seed)
    d = to_device(ep, dev)
    a = (ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'])
    out = {'config': vars(args)}
    if not args.train:
        net.eval()
        with torch.no_grad():
            got = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
            over = {i: O.recurrent_mask(got['refinement'][i - 1].float().cpu(), cfg) for i in range(1, args.T)}      # teacher forcing
            ref = O.forward({k: v.clone() for k, v in sd.items()}, cfg, *a, mask_override=over)
        out['logits'] = logits_report([got['refinement'][i].cpu() for i in range(args.T)], [ref['refinement'][i] for i in range(args.T)])
    else:
        ts = TrainStep(net)
        loss = ts.forward_backward(d)
        torch.cuda.synchronize()

        def oracle(dtype):
            sdo, params = {}, {}
            for k, v in sd.items():
                v = v.clone().to(dtype) if v.is_floating_point() else v.clone()
                if v.is_floating_point() and 'running' not in k:
                    v.requires_grad_(True)
                    params[k] = v
                sdo[k] = v
            cast = lambda x: [[t.to(dtype) for t in way] for way in x]
            over = {i: O.recurrent_mask(ts.last['logits'][i - 1].float().cpu(), cfg).to(dtype) for i in range(1, args.T)}   # teacher forcing
            o = O.forward(sdo, cfg, cast(ep['supp_imgs']), cast(ep['fore_mask']), cast(ep['back_mask']), [t.to(dtype) for t in ep['qry_imgs']],
                          ep['appr_query_labels'].to(dtype), training=True, mask_override=over)
            ls = O.train_loss(o, ep['query_labels'])
            ls.backward()
            return o, ls.detach(), params
        o32, l32, p32 = oracle(torch.float32)
        out['logits'] = logits_report([ts.last['logits'][i].cpu() for i in range(args.T)], [o32['refinement'][i].detach() for i in range(args.T)])
        out['loss'] = {'got': loss.item(), 'oracle': l32.item()}
        p64 = None
        if args.fp64:
            _, _, p64 = oracle(torch.float64)
        grads = {}
        for name, p in net.named_parameters():
            rg = p32[name].grad
            if rg is None or p.grad is None:
                continue
            g = p.grad.float().cpu()
            row = {'rel_l2': ((g - rg).norm() / rg.norm().clamp_min(1e-30)).item(), 'norm': rg.norm().item()}
            if p64 is not None:
                r64 = p64[name].grad
                row['oracle32_vs_64'] = ((rg.double() - r64).norm() / r64.norm().clamp_min(1e-300)).item()
                row['rel_l2_vs_64'] = ((g.double() - r64).norm() / r64.norm().clamp_min(1e-300)).item()
            grads[name] = row
        out['grads'] = grads
        for k, v in grads.items():
            if k.endswith('.bias') and v['norm'] < 1e-4:
                continue
            print('%-36s rel-L2 %.3e  |g| %.3e %s' % (k, v['rel_l2'], v['norm'], ('  oracle fp32 vs fp64 %.2e, ours vs fp64 %.3e' % (v['oracle32_vs_64'], v['rel_l2_vs_64'])) if 'oracle32_vs_64' in v else ''))
    for i, r in enumerate(out['logits']):
        print('iter %d: %s' % (i, json.dumps(r)))
    print(json.dumps(out))


if __name__ == '__main__':
    main()
