"""End-to-end parity of the B200 forward (through the reference-shaped nn.Module API and the C ABI) against
(a) golden outputs recorded from the unmodified reference (tests/golden/*.npz) and (b) the CPU oracle on the same
seeded inputs.  Tolerance from BASELINE.json north_star: 1e-3 relative (L-inf, relative to max |logit|) on logits;
argmax masks must match except at near-tie pixels (|logit1 - logit0| below the logit tolerance), which are counted."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL = 1e-3


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import _lib
    _lib.load()
    return torch.device('cuda:0')


def _cfg(T, soft=False, radius=5, **kw):
    d = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
             n_iter_refinement=T, soft_mask=soft, mask_refinement_correlation_radius=radius)
    d.update(kw)
    return d


def _model(sd, cfg, dev, backbone='UNet'):
    from net.model import model_factory          # the reference import path (test_rpnet.py:11,74)
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': backbone}, backbone_cfg=cfg)
    net.load_state_dict(sd)
    return net.to(dev).eval()


def _run(net, ep, dev):
    from rpnet_b200.synthetic import to_device
    d = to_device(ep, dev)
    with torch.no_grad():
        out = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], grid=None,
                  query_labels=d['query_labels'], appr_query_labels=d['appr_query_labels'])
    torch.cuda.synchronize()
    return out


def _check_logits(got, ref, what):
    got = got.float().cpu()
    rel = ((got - ref).abs().max() / ref.abs().max()).item()
    assert rel < REL_TOL, '%s: rel-Linf %.3e' % (what, rel)
    # argmax parity: mismatches are only allowed where the reference itself is within tolerance of a tie
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    mism = got.argmax(1) != ref.argmax(1)
    assert not (mism & (margin > 2 * REL_TOL * ref.abs().max())).any(), '%s: argmax differs away from ties' % what
    return rel, mism.float().mean().item()


def test_cfg1_eval_vs_reference_golden(dev, golden):
    """BASELINE.json configs[0] (1-shot 1-way, 2 x 128 x 128, T=1) against the reference's own output."""
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    g = golden('cfg1_eval')
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(int(g['w_seed'])), int(g['bn_seed']))
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    net = _model(sd, _cfg(1), dev)
    out = _run(net, ep, dev)
    ref = torch.from_numpy(g['output'])
    rel, mism = _check_logits(out['output'], ref, 'cfg1 output')
    assert torch.equal(out['output'], out['refinement'][0])
    assert out['output'].shape == (2, 2, 128, 128) and out['output'].dtype == torch.float32
    assert set(out.keys()) == {'output', 'align_loss', 'refinement'}
    # encoder features at the fixture's sub-sampling (fp16 activations: 1e-2 of the feature scale)
    d4 = net.encoder(ep['qry_imgs'][0].to(dev), None)['d4'].cpu()
    want = torch.from_numpy(g['d4_qry'])
    err = (d4[:, ::8, ::2, ::2] - want).abs().max() / want.abs().max()
    assert err < 1e-3, err


@pytest.mark.parametrize('name', ['cfg1_T3', 'cfg1_T2_soft'])
def test_recurrent_vs_reference_golden(dev, golden, name):
    """T > 1 with hard and soft masks against the reference's own refinement outputs and packed argmax masks."""
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    g = golden(name)
    Tn, soft = int(g['T']), bool(g['soft'])
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(int(g['w_seed'])), int(g['bn_seed']))
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    out = _run(_model(sd, _cfg(Tn, soft), dev), ep, dev)
    for i in range(Tn):
        got = out['refinement'][i].cpu()
        ref = torch.from_numpy(g['ref%d' % i])
        rel = ((got[:, :, ::2, ::2] - ref).abs().max() / ref.abs().max()).item()
        assert rel < REL_TOL, (name, i, rel)
        mask = np.unpackbits(g['mask%d' % i])[:got.shape[0] * got.shape[2] * got.shape[3]].reshape(got.argmax(1).shape)
        mism = (got.argmax(1).numpy().astype(np.uint8) != mask).mean()
        assert mism < 1e-3, (name, i, mism)          # near-tie pixels only (checked exactly in the oracle tests below)


@pytest.mark.parametrize('ways,shots,B,size,T', [(1, 1, 3, 64, 2), (1, 5, 2, 64, 2), (4, 2, 2, 64, 2), (1, 1, 2, 256, 2)])
def test_forward_vs_oracle(dev, ways, shots, B, size, T):
    """Oracle parity incl. the Wa x Sh generalisation (configs 3-4; 'oracle-ext') and 256 x 256."""
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(0))
    cfg = _cfg(T)
    ep = make_episode(B, ways, shots, size, seed=3)
    with torch.no_grad():
        ref = O.forward(sd, cfg, ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'])
    out = _run(_model(sd, cfg, dev), ep, dev)
    assert out['output'].shape == (B, 1 + ways, size, size)
    for i in range(T):
        _check_logits(out['refinement'][i], ref['refinement'][i], 'refinement[%d]' % i)
    _check_logits(out['output'], ref['output'], 'output')


def test_vgg_encoder_vs_reference_golden(dev, golden):
    """net.vgg.Encoder standalone against the reference's own output."""
    from net.vgg import Encoder
    from oracle import weights
    from rpnet_b200.synthetic import make_episode
    g = golden('vgg')
    enc = Encoder(3)
    enc.load_state_dict(weights.vgg_state_dict(int(g['w_seed'])))
    enc = enc.to(dev).eval()
    x = make_episode(1, size=int(g['size']), seed=int(g['ep_seed']))['qry_imgs'][0].expand(-1, 3, -1, -1)
    with torch.no_grad():
        y = enc(x.to(dev).contiguous()).cpu()
    ref = torch.from_numpy(g['out'])
    # split-fp16 convs (fp32-class): north_star's 1e-3 of the output scale holds for the standalone encoder too
    err = ((y - ref).abs().max() / ref.abs().max()).item()
    assert y.shape == ref.shape and err < 1e-3, err


def test_vgg_backbone_through_rpnet(dev):
    """`backbone: vgg` + `scale: 8` wiring (the reference raises TypeError here: SURVEY D1) vs the oracle."""
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200.synthetic import make_episode
    from net.rp_net import RP_Net
    cfg = _cfg(2, scale=8)
    torch.manual_seed(0)
    net = RP_Net(in_channels=3, cfg={'align': True, 'backbone': 'vgg'}, backbone_cfg=cfg)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    ep = make_episode(2, 1, 1, 128, seed=5)
    with torch.no_grad():
        ref = O.forward(sd, cfg, ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'],
                        backbone='vgg')
    out = _run(net.to(dev).eval(), ep, dev)
    for i in range(2):
        _check_logits(out['refinement'][i], ref['refinement'][i], 'vgg refinement[%d]' % i)


def test_cuda_graph_replay_equals_eager(dev):
    """enable_cuda_graph(): the eval schedule replayed as one CUDA graph gives bit-identical logits, follows new inputs,
    and is re-captured after the weights change."""
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats, to_device
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(0))
    net = _model(sd, _cfg(2), dev)
    eps = [to_device(make_episode(2, 1, 2, 64, seed=s), dev) for s in (1, 2)]

    def run(d):
        with torch.no_grad():
            return net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
    eager = [run(d) for d in eps]
    net.enable_cuda_graph(True)
    for rep in range(2):
        for d, ref in zip(eps, eager):
            out = run(d)
            assert torch.equal(out['output'], ref['output'])
            assert all(torch.equal(out['refinement'][i], ref['refinement'][i]) for i in range(2))
    first = run(eps[0])['output'].clone()
    run(eps[1])
    assert torch.equal(first, eager[0]['output'])                      # returned tensors are not aliased by later replays
    with torch.no_grad():
        net.cre.q[0].weight.mul_(1.5)                                   # weight change -> stale capture must not be replayed
    changed = run(eps[0])['output']
    net.enable_cuda_graph(False)
    assert torch.equal(changed, run(eps[0])['output']) and not torch.equal(changed, eager[0]['output'])


def test_forward_edge_cases_vs_oracle(dev):
    """Empty support mask on one slice, empty appr_query_labels on another (getFeatures' +1e-5 and the cosine eps clamp carry
    the path, net/rp_net.py:362,373-376), a one-slice batch, and the reference's error behaviour."""
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(0))
    cfg = _cfg(2)
    ep = make_episode(3, 1, 1, 64, seed=11)
    ep['fore_mask'][0][0][0].zero_()                     # slice 0: no support foreground at all
    ep['back_mask'][0][0][0].fill_(1)
    ep['appr_query_labels'][1].zero_()                   # slice 1: empty registered label -> first query mask is empty
    ep['fore_mask'][0][0][2].fill_(1)                    # slice 2: support foreground everywhere, background empty
    ep['back_mask'][0][0][2].zero_()
    net = _model(sd, cfg, dev)
    with torch.no_grad():
        ref = O.forward(sd, cfg, ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'])
    out = _run(net, ep, dev)
    assert torch.isfinite(out['output']).all()
    for i in range(2):
        _check_logits(out['refinement'][i], ref['refinement'][i], 'edge refinement[%d]' % i)
    # B = 1
    ep1 = make_episode(1, 1, 1, 64, seed=12)
    with torch.no_grad():
        ref1 = O.forward(sd, cfg, ep1['supp_imgs'], ep1['fore_mask'], ep1['back_mask'], ep1['qry_imgs'], ep1['appr_query_labels'])
    _check_logits(_run(net, ep1, dev)['output'], ref1['output'], 'B=1 output')
    # error behaviour of the reference: missing appr_query_labels -> AttributeError (net/rp_net.py:269); unknown backbone ->
    # NotImplementedError (:218-219)
    from rpnet_b200.synthetic import to_device
    d = to_device(ep1, dev)
    with pytest.raises(AttributeError):
        net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'])
    from net.model import model_factory
    with pytest.raises(NotImplementedError):
        model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'densenet'}, backbone_cfg=cfg)


def test_workspace_growth_is_bounded(dev):
    """Every distinct input shape owns a workspace buffer set (and a CUDA-graph capture); engine.ShapeBudget drops them all once more
    than `limit` shapes have been seen, so a long eval over heterogeneous volumes cannot grow device memory without bound — and the
    results after a release are the results before it."""
    from oracle import weights
    from rpnet_b200 import engine
    from rpnet_b200.synthetic import make_episode, perturb_bn_stats
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(0))
    net = _model(sd, _cfg(2), dev)
    net.enable_cuda_graph(True)
    first = _run(net, make_episode(1, 1, 1, 64, seed=3), dev)['output'].clone()
    peak = 0
    for b in range(1, 14):                                           # 13 distinct shapes > the budget of 8
        _run(net, make_episode(b, 1, 1, 64, seed=3), dev)
        held = net._ws.nbytes() + net.encoder._ws.nbytes()
        peak = max(peak, held)
        assert len(net._graphs) <= engine.ShapeBudget().limit
    one = net._ws.nbytes() + net.encoder._ws.nbytes()                 # what the shapes since the last release hold
    assert peak < 9 * 13 * (one // max(1, len(net._shapes.seen))) and len(net._shapes.seen) <= 8
    again = _run(net, make_episode(1, 1, 1, 64, seed=3), dev)['output']
    assert torch.equal(first, again)
