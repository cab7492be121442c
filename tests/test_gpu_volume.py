"""Slice-sharded volume inference (rpnet_b200.volume: the eval loop of test_rpnet.py:151-258 on device) against the CPU
oracle run slice-batch by slice-batch like the reference driver does (batches of 2, test_rpnet.py:164), incl. Dice as
utils/util.py:379-390 computes it, and shard-invariance (BASELINE.json configs[4]: slices are independent units)."""
import numpy as np
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
    return dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False,
                n_iter_refinement=T, soft_mask=False, mask_refinement_correlation_radius=5)


def _dice_ref(t, p):
    """utils/util.py:379-390 (dice_score_seperate for one class)."""
    t, p = t.astype(np.float64), p.astype(np.float64)
    if t.sum() + p.sum() == 0:
        return None
    return round(2 * (t * p).sum() / (t.sum() + p.sum()), 4)


def test_volume_inference_vs_oracle_and_sharding(dev):
    from net.model import model_factory
    from oracle import rpnet_oracle as O
    from oracle import weights
    from rpnet_b200 import volume as V
    from rpnet_b200.synthetic import perturb_bn_stats
    T, S, size = 2, 10, 64
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(0))
    item = V.make_synthetic_volume(S, size, 1, 1, seed=4)
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=_cfg(T))
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    mv = lambda t: t.to(dev)
    args = ([[mv(t) for t in way] for way in item['support_images']], [[mv(t) for t in way] for way in item['support_fg']],
            [[mv(t) for t in way] for way in item['support_bg']], mv(item['query_images']), mv(item['appr_query_labels']))
    full = V.segment_volume(net, *args, batch_size=4, keep_logits=True)
    assert full['range'] == (0, S) and full['mask'].shape == (S, size, size) and full['mask'].dtype == torch.uint8
    # reference driver: batches of 2 slices through the oracle (eval-mode BN: batch composition does not matter)
    ref_masks, ref_logits = [], []
    with torch.no_grad():
        for b0 in range(0, S, 2):
            cut = lambda t: t[b0:b0 + 2]
            out = O.forward(sd, _cfg(T), [[cut(t) for t in w] for w in item['support_images']], [[cut(t) for t in w] for w in item['support_fg']],
                            [[cut(t) for t in w] for w in item['support_bg']], [cut(item['query_images'])], cut(item['appr_query_labels']))
            ref_logits.append(out['output'])
            ref_masks.append((out['output'].softmax(dim=1)[:, 1] > 0.5))                       # test_rpnet.py:219,224
    ref_logits, ref_masks = torch.cat(ref_logits), torch.cat(ref_masks)
    rel = ((full['logits'].cpu() - ref_logits).abs().max() / ref_logits.abs().max()).item()
    assert rel < 1e-3, rel
    mism = (full['mask'].cpu().bool() != ref_masks).float().mean().item()
    assert mism < 1e-3, mism
    tgt = (item['query_labels'] > 0).numpy()
    d_ref = _dice_ref(tgt, ref_masks.numpy())
    d_got, per_iter = V.volume_dice(full, mv(item['query_labels']))
    assert d_got == _dice_ref(tgt, full['mask'].cpu().numpy())                                  # same formula, same rounding
    assert abs(d_got - d_ref) < 5e-3 and len(per_iter) == T and per_iter[-1] == d_got          # D5: output == refinement[T-1]
    # Dice between our masks and the reference's masks (parity metric of SURVEY §8d)
    assert _dice_ref(ref_masks.numpy(), full['mask'].cpu().numpy()) > 0.995
    # shard invariance: 3 ranks' slices concatenated == the unsharded run, bit for bit
    parts = [V.segment_volume(net, *args, batch_size=4, rank=r, world=3) for r in range(3)]
    assert [p['range'] for p in parts] == [(0, 4), (4, 7), (7, 10)]
    assert torch.equal(torch.cat([p['mask'] for p in parts]), full['mask'])


def test_eval_driver_lines_and_metrics(dev):
    """rpnet_b200.evaluate.eval_volumes (the loop of test_rpnet.py:151-258): Dice / NCC values against numpy restatements of
    utils/util.py:379-390 and net/registration.py:157-160, and the reference's printed line format."""
    from net.model import model_factory
    from oracle import weights
    from rpnet_b200 import evaluate, registration
    from rpnet_b200 import volume as V
    from rpnet_b200.synthetic import perturb_bn_stats
    T, S, size = 2, 6, 64
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(0))
    net = model_factory['RP_Net'](pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=_cfg(T))
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    raw = V.make_synthetic_volume(S, size, 1, 1, seed=2)
    q = raw['query_images'].to(dev)
    theta, wl, ws = registration.get_affine_registration(q, [[raw['support_images'][0][0].to(dev)]], [[raw['support_fg'][0][0].to(dev)]], iters=10)
    item = {'support_images': [[ws[:, None]]], 'support_labels': [[wl[:, 0]]], 'query_images': q, 'query_labels': raw['query_labels'],
            'appr_query_labels': (wl[:, 0] > 0.5).float(), 'warped_supp': ws, 'class_id': 0, 'pid': 'p0', 'supp_pid': 's0'}
    lines = []
    aff, few, ref = evaluate.eval_volumes(net, [item], ['Liver'], batch_size=4, out=lines.append)
    assert len(lines) == 2 and lines[0].startswith('0 p0 s0 affine (') and ' fewshot ' in lines[0] and 'ref 1 ' in lines[0]
    assert lines[1].startswith('Liver, affine ')
    tgt = (raw['query_labels'] > 0).numpy()
    assert aff['Liver'][0] == _dice_ref(tgt, (item['appr_query_labels'] > 0).cpu().numpy())
    res = V.segment_volume(net, item['support_images'], item['support_labels'], [[1 - wl[:, 0]]], q, item['appr_query_labels'], batch_size=4)
    assert few['Liver'][0] == _dice_ref(tgt, res['mask'].cpu().numpy()) and ref['Liver'][T - 1][0] == few['Liver'][0]
    # NCC against the reference formula in float64
    f, m = q.cpu().double().numpy(), ws.cpu().double().numpy().reshape(q.shape)
    want = -((f - f.mean()) * (m - m.mean())).sum() / np.sqrt(((f - f.mean()) ** 2).sum() * ((m - m.mean()) ** 2).sum() + 1e-10)
    got = float(lines[0].split('affine (')[1].split(',')[0])
    assert abs(got - want) < 1e-5
