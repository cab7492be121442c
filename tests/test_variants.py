"""yaml variants of the U-Net encoder: `unet_normalize_type: InstanceNorm2d` (net/modules.py:49,52,69) and
`mask_feature_map: x | x2 | x3` (net/unet.py:401-424, 437-449), pinned to outputs of the UNMODIFIED reference
(tests/golden/variants.npz, made by tests/golden/make_golden_variants.py).  The reference's weights are rebuilt with this
repository's constructors under the same seed and checked against the recorded checksums (same RNG consumption order, same keys)."""
import numpy as np
import pytest
import torch

VARIANTS = {'inorm': dict(unet_normalize_type='InstanceNorm2d'), 'mfm_x': dict(mask_feature_map='x'),
            'mfm_x2': dict(mask_feature_map='x2'), 'mfm_x3': dict(mask_feature_map='x3')}


def _cfg(T, **extra):
    d = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=T,
             soft_mask=False, mask_refinement_correlation_radius=5)
    d.update(extra)
    return d


def _build(name, g):
    """Our RP_Net for the variant with the golden's seeds; returns (net on CPU, cfg, episode)."""
    from net.model import model_factory
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    cfg = _cfg(int(g['T']), **VARIANTS[name])
    torch.manual_seed(int(g['w_seed']))
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    perturb_bn_stats(net.state_dict(), seed=int(g['bn_seed']))
    cs = weights.checksums(net.state_dict())
    assert list(cs.keys()) == [str(k) for k in g[name + '_keys']]                     # the reference's state_dict keys, in order
    np.testing.assert_allclose(np.array(list(cs.values())), g[name + '_cs'], rtol=1e-12, atol=1e-12)
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    return net, cfg, ep


@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_oracle_variants_vs_reference_golden(golden, name):
    from oracle import rpnet_oracle as O
    g = golden('variants')
    net, cfg, ep = _build(name, g)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    with torch.no_grad():
        out = O.forward(sd, cfg, ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'])
        d4 = O.unet_encoder(ep['qry_imgs'][0], sd, mask=ep['fore_mask'][0][0].unsqueeze(1), mask_feature_map=cfg['mask_feature_map'])
    for i in range(int(g['T'])):
        np.testing.assert_allclose(out['refinement'][i].numpy(), g['%s_ref%d' % (name, i)], rtol=0, atol=2e-5)
    np.testing.assert_allclose(d4[:, ::8, ::2, ::2].numpy(), g[name + '_d4'], rtol=0, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_variants_vs_reference_golden(golden, name):
    """Eval forward of the variant on the B200 path (split precision) against the reference's own logits and d4 features: 1e-3."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200.synthetic import to_device
    g = golden('variants')
    net, cfg, ep = _build(name, g)
    dev = torch.device('cuda:0')
    net = net.to(dev).eval()
    d = to_device(ep, dev)
    with torch.no_grad():
        out = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
        d4 = net.encoder(d['qry_imgs'][0], d['fore_mask'][0][0].unsqueeze(1))['d4'].cpu()
    torch.cuda.synchronize()
    want = torch.from_numpy(g[name + '_d4'])
    assert ((d4[:, ::8, ::2, ::2] - want).abs().max() / want.abs().max()).item() < 1e-3
    for i in range(int(g['T'])):
        ref = torch.from_numpy(g['%s_ref%d' % (name, i)])
        got = out['refinement'][i].cpu()
        rel = ((got - ref).abs().max() / ref.abs().max()).item()
        assert rel < 1e-3, (name, i, rel)
        if (got.argmax(1) != ref.argmax(1)).any():          # a near-tie flip changes the next iteration's input
            break
    if name.startswith('mfm'):                              # more than one way / shot: the reference's own shape error
        with pytest.raises(ValueError):
            net([[d['supp_imgs'][0][0], d['supp_imgs'][0][0]]], [[d['fore_mask'][0][0]] * 2], [[d['back_mask'][0][0]] * 2], d['qry_imgs'],
                appr_query_labels=d['appr_query_labels'])


@pytest.mark.gpu
def test_instance_norm_train_step_vs_oracle():
    """`unet_normalize_type: InstanceNorm2d` through the train step (one statistics group per image, no affine parameters, no running
    statistics in the encoder; `cre` keeps BatchNorm2d): logits 1e-3, loss 1e-3, head gradients against the oracle's autograd."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from net.model import model_factory
    from oracle import rpnet_oracle as O
    from rpnet_b200.synthetic import make_episode, to_device
    from rpnet_b200.train import TrainStep
    from rpnet_b200 import parity
    T = 2
    cfg = _cfg(T, unet_normalize_type='InstanceNorm2d')
    torch.manual_seed(0)
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    assert not any('encoder' in k and 'running' in k for k in sd)
    dev = torch.device('cuda:0')
    net = net.to(dev).train()
    ep = make_episode(2, 1, 2, 64, seed=7)
    ts = TrainStep(net)
    loss = ts.forward_backward(to_device(ep, dev))
    torch.cuda.synchronize()
    params = {}
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            sd[k] = v.clone().requires_grad_(True)
            params[k] = sd[k]
    over = {i: O.recurrent_mask(ts.last['logits'][i - 1].float().cpu(), cfg) for i in range(1, T)}
    out = O.forward(sd, cfg, ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'], training=True,
                    mask_override=over)
    ref_loss = O.train_loss(out, ep['query_labels'])
    ref_loss.backward()
    for i in range(T):
        r = parity.compare_logits(ts.last['logits'][i].cpu(), out['refinement'][i].detach())
        assert r['rel_linf'] < 1e-3 and r['margin_rel_err'] < 2e-3, (i, r)
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    named = dict(net.named_parameters())
    for name in ('cre.q.0.weight', 'cre.q.1.weight', 'cre.w_k.0.weight', 'encoder.Up_conv4.conv.3.weight', 'encoder.Conv3.conv.0.weight'):
        g, rg = named[name].grad.float().cpu(), params[name].grad
        rel = ((g - rg).norm() / rg.norm()).item()
        assert rel < (5e-2 if name.startswith('cre') else 0.15), (name, rel)


@pytest.mark.gpu
@pytest.mark.parametrize('mfm', ['x', 'x3'])
def test_instance_norm_with_mask_feature_map_vs_oracle(mfm):
    """Both yaml variants together (`unet_normalize_type: InstanceNorm2d` + `mask_feature_map`), eval forward and train step against
    the oracle (which restates both branches, each pinned to the reference on its own above)."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from net.model import model_factory
    from oracle import rpnet_oracle as O
    from rpnet_b200 import parity
    from rpnet_b200.synthetic import make_episode, to_device
    from rpnet_b200.train import TrainStep
    T = 2
    cfg = _cfg(T, unet_normalize_type='InstanceNorm2d', mask_feature_map=mfm)
    torch.manual_seed(2)
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    dev = torch.device('cuda:0')
    ep = make_episode(2, 1, 1, 64, seed=5)
    d = to_device(ep, dev)
    a = (ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'])
    net = net.to(dev).eval()
    with torch.no_grad():
        out = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
        ref = O.forward({k: v.clone() for k, v in sd.items()}, cfg, *a)
    r = parity.compare_logits(out['refinement'][0].cpu(), ref['refinement'][0])
    assert r['rel_linf'] < 1e-3, r
    net.train()
    ts = TrainStep(net)
    loss = ts.forward_backward(d)
    torch.cuda.synchronize()
    params = {}
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            sd[k] = v.clone().requires_grad_(True)
            params[k] = sd[k]
    over = {i: O.recurrent_mask(ts.last['logits'][i - 1].float().cpu(), cfg) for i in range(1, T)}
    tr = O.forward(sd, cfg, *a, training=True, mask_override=over)
    ref_loss = O.train_loss(tr, ep['query_labels'])
    ref_loss.backward()
    for i in range(T):
        r = parity.compare_logits(ts.last['logits'][i].cpu(), tr['refinement'][i].detach())
        assert r['rel_linf'] < 1e-3 and r['margin_rel_err'] < 2e-3, (i, r)
    assert abs(loss.item() - ref_loss.item()) < 1e-3 * abs(ref_loss.item())
    named = dict(net.named_parameters())
    for name in ('cre.q.0.weight', 'encoder.Conv3.conv.0.weight', 'encoder.Conv1.conv.0.weight'):
        g, rg = named[name].grad.float().cpu(), params[name].grad
        assert ((g - rg).norm() / rg.norm()).item() < (5e-2 if name.startswith('cre') else 0.15), name
