#!/usr/bin/env python
"""The other two backbones (net/vgg.py:22-58, `scale: 8`, where the reference raises TypeError; row N3,
net/rp_net.py:19-42, ResNet18): throughput of the standalone encoder, of the RP_Net eval forward and of the train step on one B200,
next to stock PyTorch (cuDNN) running the oracle's functional graph on the same GPU, with the parity of the timed eval path against
the fp32 CPU oracle.

    python tools/bench_vgg.py [--backbone vgg|resnet] [--batch 8] [--size 256] [--T 4] [--steps 20]

One JSON line: images/s of the encoder (roofline against E_vgg = 50.96 / E_resnet18 = 91.7 GF per 256 x 256 image), slices/s of the
forward and of the train step (forward + backward + Adam)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
E_VGG = 50_960_793_600


def resnet18_flops(size=256):
    """2 * MACs of the convs of the reference's ResNet18 wrapper (stem 7x7/2, layer1, three stride-1 stages with 1x1 downsamples)."""
    h2, h4 = size // 2, size // 4
    f = 2 * h2 * h2 * 64 * 147
    f += 4 * 2 * h4 * h4 * 64 * 64 * 9
    for cin, cout in ((64, 128), (128, 256), (256, 512)):
        f += 2 * h4 * h4 * (cin * cout * 9 + 3 * cout * cout * 9 + cin * cout)
    return f



def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--backbone', default='vgg', choices=['vgg', 'resnet'])
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--size', type=int, default=256)
    ap.add_argument('--T', type=int, default=4)
    ap.add_argument('--steps', type=int, default=20)
    args = ap.parse_args()
    from net.rp_net import RP_Net
    from oracle import rpnet_oracle as O
    from rpnet_b200 import engine, parity
    from rpnet_b200.synthetic import make_episode, to_device
    dev = torch.device('cuda:0')
    cfg = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=args.T,
               soft_mask=False, mask_refinement_correlation_radius=5)
    bb = args.backbone
    scale = 8 if bb == 'vgg' else 4
    if bb == 'vgg':
        cfg['scale'] = 8
    enc_oracle = (lambda x, sd_: O.vgg_encoder(x, sd_, 'encoder.')) if bb == 'vgg' else (lambda x, sd_: O.resnet_encoder(x, sd_, 'encoder.'))
    torch.manual_seed(0)
    net = RP_Net(in_channels=3, cfg={'align': True, 'backbone': bb}, backbone_cfg=cfg)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(dev).eval()
    B = args.batch
    ep = make_episode(B, 1, 1, args.size, seed=0)
    d = to_device(ep, dev)
    imgs = torch.cat([d['supp_imgs'][0][0], d['qry_imgs'][0]]).expand(-1, 3, -1, -1).contiguous()
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    peak = peaks.get('bf16_tflops_sustained', 1400.0)

    with torch.no_grad():
        ms_enc = timed(lambda: net.encoder.encode_nhwc(imgs, 'bench'), args.steps)
        fwd = lambda: net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
        ms_fwd = timed(fwd, args.steps)
        net.enable_cuda_graph(True)
        ms_graph = timed(fwd, args.steps)
        net.enable_cuda_graph(False)
        out = fwd()
        # parity: 2 slices through the CPU oracle, teacher-forced (the hard mask is discontinuous)
        ep2 = make_episode(2, 1, 1, args.size, seed=4242)
        d2 = to_device(ep2, dev)
        got = net(d2['supp_imgs'], d2['fore_mask'], d2['back_mask'], d2['qry_imgs'], appr_query_labels=d2['appr_query_labels'])
        over = {i: O.recurrent_mask(got['refinement'][i - 1].float().cpu(), cfg, scale) for i in range(1, args.T)}
        ref = O.forward(sd, cfg, ep2['supp_imgs'], ep2['fore_mask'], ep2['back_mask'], ep2['qry_imgs'], ep2['appr_query_labels'],
                        backbone=bb, mask_override=over)
        per = [parity.compare_logits(got['refinement'][i].cpu(), ref['refinement'][i]) for i in range(args.T)]
        # stock PyTorch on the same GPU: the oracle's functional VGG graph (cuDNN, TF32 default / fp32)
        sdg = {k: v.to(dev) for k, v in sd.items()}
        lib = {}
        for mode in ('tf32', 'fp32'):
            torch.backends.cudnn.allow_tf32 = mode == 'tf32'
            torch.backends.cudnn.benchmark = True
            lib[mode] = {'encoder_images_per_s': imgs.shape[0] / (timed(lambda: enc_oracle(imgs, sdg), 10) * 1e-3),
                         'forward_slices_per_s': B / (timed(lambda: O.forward(sdg, cfg, d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'],
                                                                               d['appr_query_labels'], backbone=bb, allpairs=True), 5) * 1e-3)}
        torch.backends.cudnn.allow_tf32 = True
    # train step (forward + backward + Adam) on the same episode: VggTrainEngine / ResNetTrainEngine
    from rpnet_b200.train import TrainStep
    net.train()
    ts = TrainStep(net)
    ms_train = timed(lambda: ts.step(d), max(5, args.steps // 2))
    n_img = imgs.shape[0]
    algo = n_img * (E_VGG * (args.size / 256.0) ** 2 if bb == 'vgg' else resnet18_flops(args.size))
    units = {'split': 3, 'split8': 2}.get(engine.default_precision(), 1)      # fp16-equivalent tensor passes per conv
    line = {'metric': '%s backbone: encoder images/s, RP_Net(backbone=%s, scale=%d) eval forward and train step slices/s'
                      % ({'vgg': 'VGG (net/vgg.py:22-58)', 'resnet': 'ResNet18 (net/rp_net.py:19-42)'}[bb], bb, scale), 'n_gpus': 1,
            'precision': engine.default_precision(),
            'encoder': {'images_per_s': n_img / (ms_enc * 1e-3), 'ms': ms_enc, 'images': n_img,
                        'roofline': {'bound': 'tensor', 'achieved': algo / (ms_enc * 1e-3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                                     'frac': algo / (ms_enc * 1e-3) / 1e12 / peak,
                                     'executed_fp16_equiv_tflops': algo * units / (ms_enc * 1e-3) / 1e12,
                                     'note': 'algorithmic FLOPs of all convs (first conv on CUDA cores included); split executes the tensor-core convs three times, split8 once in fp16 plus two e4m3 passes at twice the rate (2 units)'}},
            'forward': {'slices_per_s': B / (ms_fwd * 1e-3), 'ms': ms_fwd, 'cuda_graph_slices_per_s': B / (ms_graph * 1e-3), 'batch': B, 'size': args.size, 'T': args.T},
            'train_step': {'slices_per_s': B / (ms_train * 1e-3), 'ms': ms_train, 'batch': B, 'what': '1-shot 1-way, forward + backward + Adam'},
            'parity': {'rel_linf': max(p['rel_linf'] for p in per), 'margin_rel_err': max(p['margin_rel_err'] for p in per),
                       'argmax_mismatch': max(p['argmax_mismatch'] for p in per), 'dice_vs_ref': min(p['dice_vs_ref'] for p in per),
                       'what': '2 slices, worst of T iterations, teacher-forced, vs the fp32 CPU oracle (random init)'},
            'library_baseline': lib, 'logits_shape': list(out['output'].shape)}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
