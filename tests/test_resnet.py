"""Backbone 'resnet' ("next" row N3, net/rp_net.py:19-42): oracle restatement and reference-identical init against golden
vectors from the reference (CPU), and the B200 eval forward against the same golden vectors (GPU)."""
import numpy as np
import pytest
import torch


def _cfg(T):
    return dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
                n_iter_refinement=T, soft_mask=False, mask_refinement_correlation_radius=5)


def test_resnet_oracle_and_init_vs_reference_golden(golden):
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    g = golden('resnet')
    sd = weights.resnet_rpnet_state_dict(int(g['w_seed']))
    cs = weights.checksums(sd)
    assert list(g['keys']) == list(cs.keys())                                   # same 158 state_dict keys, same order
    # same tensors (float64 checksums; the summation order of .sum() depends on the host's vector width: 1e-12)
    np.testing.assert_allclose(np.array(list(cs.values())), g['vals'], rtol=1e-12, atol=1e-12)
    perturb_bn_stats(sd, int(g['bn_seed']))
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    with torch.no_grad():
        out = O.forward(sd, _cfg(int(g['T'])), ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'],
                        backbone='resnet', allpairs=True)
        d4 = O.resnet_encoder(ep['qry_imgs'][0].expand(-1, 3, -1, -1), sd)
    assert torch.equal(d4[:, ::16, ::2, ::2], torch.from_numpy(g['d4_qry']))
    for i in range(int(g['T'])):
        assert torch.equal(out['refinement'][i][:, :, ::2, ::2], torch.from_numpy(g['ref%d' % i]))
    assert torch.equal(out['output'][:, :, ::2, ::2], torch.from_numpy(g['output']))


def test_resnet_oracle_train_mode_vs_torchvision_modules():
    """Train-mode restatement (batch-statistics BatchNorm, running statistics, num_batches_tracked) and its gradients against the
    torch modules themselves: rpnet_b200.nn.resnet.ResNet18.backbone is the nn.Sequential of torchvision conv / bn / BasicBlock
    modules the reference builds (net/rp_net.py:19-36); forward() of the wrapper never calls it, here it is the independent check."""
    from oracle import rpnet_oracle as O
    from rpnet_b200.nn.resnet import ResNet18
    torch.manual_seed(3)
    enc = ResNet18().train()
    x = torch.randn(3, 3, 64, 64)
    sd = {'encoder.' + k: v.clone() for k, v in enc.state_dict().items()}
    params = {}
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            sd[k] = v.clone().requires_grad_(True)
            params[k] = sd[k]
    want = enc.backbone(x)
    got = O.resnet_encoder(x, sd, 'encoder.', training=True)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    w = torch.randn_like(want)
    (want * w).sum().backward()
    (got * w).sum().backward()
    for k, p in enc.named_parameters():
        g, rg = params['encoder.' + k].grad, p.grad
        assert ((g - rg).norm() / rg.norm().clamp_min(1e-12)).item() < 1e-4, k
    for k, b in enc.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            torch.testing.assert_close(sd['encoder.' + k].float(), b.float(), rtol=1e-5, atol=1e-6, msg=k)


def test_resnet_module_state_dict_keys(golden):
    from rpnet_b200.nn.rp_net import RP_Net
    g = golden('resnet')
    net = RP_Net(cfg={'align': True, 'backbone': 'resnet'}, backbone_cfg=_cfg(2))
    assert list(net.state_dict().keys()) == [str(k) for k in g['keys']]


@pytest.mark.gpu
def test_resnet_eval_forward_vs_reference_golden(golden):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from net.model import model_factory
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats, to_device
    g = golden('resnet')
    dev = torch.device('cuda:0')
    sd = perturb_bn_stats(weights.resnet_rpnet_state_dict(int(g['w_seed'])), int(g['bn_seed']))
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'resnet'}, backbone_cfg=_cfg(int(g['T'])))
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    d = to_device(ep, dev)
    with torch.no_grad():
        out = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
        d4 = net.encoder(d['qry_imgs'][0].expand(-1, 3, -1, -1).contiguous(), None)['d4'].cpu()
    want = torch.from_numpy(g['d4_qry'])
    assert ((d4[:, ::16, ::2, ::2] - want).abs().max() / want.abs().max()).item() < 1e-3           # split-fp16 (fp32-class) convs
    for i in range(int(g['T'])):
        ref = torch.from_numpy(g['ref%d' % i])
        rel = ((out['refinement'][i].cpu()[:, :, ::2, ::2] - ref).abs().max() / ref.abs().max()).item()
        assert rel < 1e-3, (i, rel)
    assert torch.equal(out['output'], out['refinement'][int(g['T']) - 1])
