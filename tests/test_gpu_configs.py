"""Parity at the sizes of BASELINE.json's configs against the fp32 CPU oracle, on UNSATURATED fixtures.

At random initialisation with arbitrary running statistics every logit of the eval forward sits within 1.5 % of the
20 * cos cap and rel-Linf says little; the fixtures here first run a few Adam steps of the B200 train step on other synthetic
episodes (rpnet_b200.synthetic.fitted_state_dict), so that the BatchNorm statistics describe the data and the decision
margins span most of [-40, 40].  Gates (rpnet_b200/parity.py), per refinement iteration:
  rel_linf        <= 1e-3   BASELINE.json north_star ("within 1e-3 rel fp32"), relative to max |logit|
  margin_rel_err  <= 1e-3   error of the decision margin relative to the range of the margin
  argmax          no mismatch away from reference near-ties (|margin| <= 2e-3 * max |logit|), mismatch fraction <= 3e-4
  dice_vs_ref     >= 0.999  Dice(our foreground mask, the oracle's)
Iterations i >= 1 are compared with the oracle consuming the recurrent masks derived from OUR logits of iteration i - 1
(oracle.forward(mask_override=...)): the hard threshold (net/rp_net.py:310) makes the next iteration's input discontinuous in
the logits, so a single near-tie pixel thresholded the other way would otherwise be compared through different inputs.  The
flipped pixels themselves are bounded by the argmax gate of every iteration.
Train-mode cases also check the loss (rel 1e-3) and per-parameter gradients of the head.
The CPU oracle needs 5 - 60 s per case on the GPU box's host cores."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import _lib
    _lib.load()
    return torch.device('cuda:0')


def _cfg(T):
    return dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=T,
                soft_mask=False, mask_refinement_correlation_radius=5)


def _net(sd, T, dev):
    from net.model import model_factory
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=_cfg(T))
    net.load_state_dict(sd)
    return net.to(dev)


def _fitted(ways, shots, size, T, dev, steps=30, batch=2):
    from oracle import weights
    from rpnet_b200.synthetic import fitted_state_dict, make_episode, to_device
    net = _net(weights.unet_rpnet_state_dict(0), T, dev)
    return fitted_state_dict(net, lambda i: to_device(make_episode(batch, ways, shots, size, seed=500 + 7 * i), dev), steps=steps)


def _gate(got, ref, what):
    from rpnet_b200 import parity
    r = parity.compare_logits(got.float().cpu(), ref)
    assert r['rel_linf'] <= 1e-3, (what, r)
    assert r['margin_rel_err'] <= 1e-3, (what, r)
    assert r['argmax_mismatch'] <= 3e-4 and r['dice_vs_ref'] >= 0.999, (what, r)      # every mismatch must also be a near-tie (below)
    top2 = ref.topk(2, dim=1).values
    far = (top2[:, 0] - top2[:, 1]) > 2e-3 * ref.abs().max()
    assert not ((got.float().cpu().argmax(1) != ref.argmax(1)) & far).any(), (what, 'argmax differs away from ties')
    assert r['margin_median'] < 19.0, (what, 'fixture is saturated', r)
    return r


def _overrides(logits, T):
    """{i: pooled mask from OUR logits of iteration i - 1} for the oracle's teacher-forced run."""
    from oracle import rpnet_oracle as O
    cfg = _cfg(T)
    return {i: O.recurrent_mask(logits[i - 1].float().cpu(), cfg) for i in range(1, T)}


def _eval_case(dev, ways, shots, B, size, T, seed):
    from oracle import rpnet_oracle as O
    from rpnet_b200.synthetic import make_episode, to_device
    sd = _fitted(ways, shots, size, T, dev)
    ep = make_episode(B, ways, shots, size, seed=seed)
    net = _net(sd, T, dev).eval()
    d = to_device(ep, dev)
    with torch.no_grad():
        out = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])
        torch.cuda.synchronize()
        ref = O.forward({k: v.clone() for k, v in sd.items()}, _cfg(T), ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'],
                        ep['appr_query_labels'], mask_override=_overrides(out['refinement'], T))
    return [_gate(out['refinement'][i], ref['refinement'][i], 'refinement[%d]' % i) for i in range(T)]


def test_precisions_on_one_fixture(dev):
    """`b200_precision` (engine.PRECISIONS) on one unsaturated fixture, eval and train mode, iteration 0: 'split8' (default: fp16
    main term + e4m3 corrections) and 'split' (three fp16 passes) meet the 1e-3 gate with the same error; 'fp16' (single-term,
    TF32-class) is measurably further away (on the 256 x 256 BASELINE shapes it misses the gate: DESIGN.md §2)."""
    from net.model import model_factory
    from oracle import rpnet_oracle as O
    from rpnet_b200 import parity
    from rpnet_b200.synthetic import make_episode, to_device
    from rpnet_b200.train import TrainStep
    T, B, size = 1, 4, 128
    sd = _fitted(1, 1, size, T, dev)
    ep = make_episode(B, 1, 1, size, seed=17)
    d = to_device(ep, dev)
    a = (ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'])
    with torch.no_grad():
        ref_eval = O.forward({k: v.clone() for k, v in sd.items()}, _cfg(T), *a)['refinement'][0]
        ref_train = O.forward({k: v.clone() for k, v in sd.items()}, _cfg(T), *a, training=True)['refinement'][0]
    res = {}
    for pr in ('split8', 'split', 'fp16'):
        net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=dict(_cfg(T), b200_precision=pr))
        net.load_state_dict(sd)
        net = net.to(dev).eval()
        with torch.no_grad():
            ev = net(d['supp_imgs'], d['fore_mask'], d['back_mask'], d['qry_imgs'], appr_query_labels=d['appr_query_labels'])['refinement'][0]
        net.train()
        ts = TrainStep(net)
        ts.forward_backward(d)
        torch.cuda.synchronize()
        res[pr] = (parity.compare_logits(ev.float().cpu(), ref_eval), parity.compare_logits(ts.last['logits'][0].float().cpu(), ref_train))
    for pr in ('split8', 'split'):
        for r in res[pr]:
            assert r['rel_linf'] <= 1e-3 and r['margin_rel_err'] <= 1e-3, (pr, res)
    # the e4m3 corrections cost nothing measurable against the three-pass form ...
    assert res['split8'][0]['rel_linf'] < 2 * res['split'][0]['rel_linf'] + 1e-4, res
    assert res['split8'][1]['rel_linf'] < 2 * res['split'][1]['rel_linf'] + 1e-4, res
    # ... and the single-term arithmetic is several times further away
    assert res['fp16'][1]['rel_linf'] > 1.5 * res['split8'][1]['rel_linf'], res           # measured 2.9e-4 vs 1.3e-4 (eval 1.5e-4)


def test_cfg2_eval_8x256_T4(dev):
    """BASELINE.json configs[1]: 1-shot 1-way, batch 8 x 256 x 256, T = 4, forward only."""
    _eval_case(dev, 1, 1, 8, 256, 4, seed=11)


def test_cfg4_shape_eval_4way_5shot_T6(dev):
    """BASELINE.json configs[3] shape, forward: 5-shot 4-way, T = 6, 256 x 256 (one query slice + 20 support slices)."""
    _eval_case(dev, 4, 5, 1, 256, 6, seed=13)


def test_cfg5_volume_32_slices(dev):
    """BASELINE.json configs[4] shape: a 32-slice 256 x 256 volume through the slice-batched eval loop (batches of 16) vs the
    oracle on every slice: logits, masks and the volume Dice."""
    from oracle import rpnet_oracle as O
    from rpnet_b200 import volume as V
    T = 4
    sd = _fitted(1, 1, 256, T, dev)
    item = V.make_synthetic_volume(32, 256, 1, 1, seed=3)
    net = _net(sd, T, dev).eval()
    mv = lambda t: t.to(dev)
    res = V.segment_volume(net, [[mv(t) for t in way] for way in item['support_images']], [[mv(t) for t in way] for way in item['support_fg']],
                           [[mv(t) for t in way] for way in item['support_bg']], mv(item['query_images']), mv(item['appr_query_labels']),
                           batch_size=16, keep_logits=True)
    torch.cuda.synchronize()
    import torch.nn.functional as F
    over = {i: F.avg_pool2d(res['masks_per_iter'][i - 1].cpu().float().unsqueeze(1), 4) for i in range(1, T)}
    with torch.no_grad():
        ref = O.forward({k: v.clone() for k, v in sd.items()}, _cfg(T), item['support_images'], item['support_fg'], item['support_bg'],
                        [item['query_images']], item['appr_query_labels'], mask_override=over)
    _gate(res['logits'], ref['output'], 'volume output')
    ref_mask = (ref['output'][:, 1] > ref['output'][:, 0])
    assert (res['mask'].cpu().bool() != ref_mask).float().mean().item() <= 1e-4
    tgt = item['query_labels'] > 0
    d_ours, _ = V.volume_dice(res, item['query_labels'].to(dev))
    d_ref = V.dice_from_sums(V.dice_sums(ref_mask, tgt), float(tgt.sum()))
    assert abs(d_ours - d_ref) <= 2e-4, (d_ours, d_ref)


@pytest.mark.parametrize('ways,shots,B,T', [(1, 5, 2, 4), (4, 5, 1, 6)])
def test_train_step_at_config_shapes(dev, ways, shots, B, T):
    """BASELINE.json configs[2] (5-shot 1-way, T = 4) and configs[3] (5-shot 4-way, T = 6) shapes at 256 x 256, train step:
    train-mode logits of every iteration, loss and the head gradients against the fp32 oracle + torch autograd."""
    from oracle import rpnet_oracle as O
    from rpnet_b200.synthetic import make_episode, to_device
    from rpnet_b200.train import TrainStep
    size = 256
    sd = _fitted(ways, shots, size, T, dev, steps=20, batch=1 if ways > 1 else 2)
    ep = make_episode(B, ways, shots, size, seed=17)
    net = _net({k: v.clone() for k, v in sd.items()}, T, dev).train()
    ts = TrainStep(net)
    loss = ts.forward_backward(to_device(ep, dev))
    torch.cuda.synchronize()
    params = {}
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            sd[k] = v.clone().requires_grad_(True)
            params[k] = sd[k]
    out = O.forward(sd, _cfg(T), ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], ep['appr_query_labels'], training=True,
                    mask_override=_overrides(ts.last['logits'], T))
    ref_loss = O.train_loss(out, ep['query_labels'])
    ref_loss.backward()
    for i in range(T):
        _gate(ts.last['logits'][i], out['refinement'][i].detach(), 'train refinement[%d]' % i)
    assert abs(loss.item() - ref_loss.item()) <= 1e-3 * abs(ref_loss.item()), (loss.item(), ref_loss.item())
    for name in ('cre.q.0.weight', 'cre.q.1.weight', 'cre.w_k.0.weight', 'cre.w_q.0.weight', 'encoder.Up_conv4.conv.3.weight'):
        g, rg = dict(net.named_parameters())[name].grad.float().cpu(), params[name].grad
        rel = ((g - rg).norm() / rg.norm()).item()
        assert rel <= 1e-2, (name, rel)
