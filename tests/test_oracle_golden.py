"""Pin the CPU oracle (oracle/rpnet_oracle.py) against golden vectors produced by executing the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import rpnet_oracle as O
from oracle import weights
from rpnet_b200.synthetic import make_episode, perturb_bn_stats

T = torch.from_numpy


def cfg_of(Tn, soft=False, radius=5):
    return dict(n_iter_refinement=Tn, soft_mask=soft, mask_refinement_correlation_radius=radius)


def test_init_matches_reference_rng_order(golden):
    g = golden('init_checksums')
    sd = weights.unet_rpnet_state_dict(0)
    assert list(g['keys']) == list(sd.keys())          # 147 state_dict keys, SURVEY §8(b)
    assert len(sd) == 147
    cs = weights.checksums(sd)
    np.testing.assert_array_equal(np.array(list(cs.values())), g['vals'])
    g = golden('init_checksums_vgg')
    sd = weights.vgg_state_dict(0)
    assert list(g['keys']) == list(sd.keys())
    np.testing.assert_array_equal(np.array(list(weights.checksums(sd).values())), g['vals'])


def test_correlation_local_and_allpairs(golden):
    g = golden('correlation')
    for i in range(int(g['n'])):
        f1, f2, r = T(g['f1_%d' % i]), T(g['f2_%d' % i]), int(g['r_%d' % i])
        ref = T(g['out_%d' % i])
        assert torch.equal(O.correlation_allpairs(f1, f2, r), ref)       # faithful restatement: bit-equal
        loc = O.correlation_local(f1, f2, r)
        assert loc.shape == ref.shape
        torch.testing.assert_close(loc, ref, rtol=0, atol=2e-5)        # D7: local form == all-pairs + gather


def test_prototype_and_losses(golden):
    g = golden('proto_loss')
    assert torch.equal(O.get_features(T(g['fts']), T(g['mask'])), T(g['proto']))
    assert torch.equal(O.get_features(T(g['fts']), torch.zeros(1, 64, 64)), T(g['proto_empty']))
    d = O.cal_dist(T(g['qf']), T(g['proto']))
    assert torch.equal(d, T(g['dist']))
    assert d[0, 3, 4] == 0                                               # all-zero feature -> cosine 0
    fg = [[x for x in w] for w in T(g['fg_l'])]
    bg = [[x for x in w] for w in T(g['bg_l'])]
    fgp, bgp = O.get_prototype(fg, bg)
    assert torch.equal(torch.stack(fgp), T(g['fgp'])) and torch.equal(bgp, T(g['bgp']))
    assert torch.equal(O.dice_ce(T(g['logits']), T(g['labels'])), T(g['dice_ce']))
    assert torch.equal(O.dice_ce(T(g['logits5']), T(g['labels5'])), T(g['dice_ce5']))
    al = O.align_loss(T(g['a_q']), T(g['a_pred']), T(g['a_s']), T(g['a_f']), T(g['a_b']))
    assert torch.equal(al, T(g['align']))


def test_forward_cfg1_eval(golden):
    """BASELINE.json configs[0]: 1-shot 1-way, 2x128x128, T=1, CPU forward."""
    g = golden('cfg1_eval')
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(int(g['w_seed'])), int(g['bn_seed']))
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    want = {}
    with torch.no_grad():
        out = O.forward(sd, cfg_of(1), ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'],
                        ep['appr_query_labels'], want=want)
    ref = T(g['output'])
    assert torch.equal(out['output'], out['refinement'][0])             # D5
    # local correlation reorders the fp32 sums: not bit-equal, but far inside 1e-3 rel and same masks
    rel = ((out['output'] - ref).abs().max() / ref.abs().max()).item()
    assert rel < 1e-5, rel
    assert torch.equal(out['output'].argmax(1), ref.argmax(1))
    torch.testing.assert_close(want['supp_fts'][0, 0][:, ::8, ::2, ::2], T(g['d4_supp']), rtol=0, atol=0)
    torch.testing.assert_close(want['qry_fts'][0][:, ::8, ::2, ::2], T(g['d4_qry']), rtol=0, atol=0)
    torch.testing.assert_close(want['supp_cre'][0, 0][:, ::4, ::2, ::2], T(g['cre_supp']), rtol=0, atol=1e-5)
    # the all-pairs data flow reproduces the reference bit for bit
    with torch.no_grad():
        out2 = O.forward(sd, cfg_of(1), ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'],
                         ep['appr_query_labels'], allpairs=True)
    assert torch.equal(out2['output'], ref)


@pytest.mark.parametrize('name', ['cfg1_T3', 'cfg1_T2_soft'])
def test_forward_recurrent(golden, name):
    g = golden(name)
    Tn, soft = int(g['T']), bool(g['soft'])
    sd = perturb_bn_stats(weights.unet_rpnet_state_dict(int(g['w_seed'])), int(g['bn_seed']))
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    with torch.no_grad():
        out = O.forward(sd, cfg_of(Tn, soft), ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'],
                        ep['appr_query_labels'], allpairs=True)
    for i in range(Tn):
        assert torch.equal(out['refinement'][i][:, :, ::2, ::2], T(g['ref%d' % i]))
        mask = np.packbits(out['refinement'][i].argmax(1).numpy().astype(np.uint8))
        np.testing.assert_array_equal(mask, g['mask%d' % i])


def test_vgg_encoder(golden):
    g = golden('vgg')
    sd = {'encoder.' + k: v for k, v in weights.vgg_state_dict(int(g['w_seed'])).items()}
    x = make_episode(1, size=int(g['size']), seed=int(g['ep_seed']))['qry_imgs'][0].expand(-1, 3, -1, -1)
    with torch.no_grad():
        y = O.vgg_encoder(x, sd)
    assert torch.equal(y, T(g['out']))


def test_train_step(golden):
    """Reconstructed train step: loss, per-parameter grad norms, BN running stats (D14)."""
    g = golden('train_step')
    sd = weights.unet_rpnet_state_dict(int(g['w_seed']))
    params = {}
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            sd[k] = v.clone().requires_grad_(True)
            params[k] = sd[k]
    ep = make_episode(int(g['B']), size=int(g['size']), seed=int(g['ep_seed']))
    out = O.forward(sd, cfg_of(int(g['T'])), ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'],
                    ep['appr_query_labels'], training=True, allpairs=True)
    loss = O.train_loss(out, ep['query_labels'])
    loss.backward()
    assert torch.equal(out['refinement'][0], T(g['out0'])) and torch.equal(out['refinement'][1], T(g['out1']))
    assert torch.equal(loss.detach(), T(g['loss']))
    assert list(g['names']) == list(params.keys())
    for name, n_ref, head in zip(g['names'], g['grad_norm'], g['grad_head']):
        p = params[str(name)]
        if n_ref < 0:                                  # cre.w_context / cre.out: unused (D4)
            assert p.grad is None, name
        else:
            np.testing.assert_allclose(p.grad.norm().item(), n_ref, rtol=1e-6, err_msg=str(name))
    bn = {k: v for k, v in sd.items() if 'running' in k or 'num_batches' in k}
    assert list(g['bn_keys']) == list(bn.keys())
    np.testing.assert_allclose(np.array([v.double().sum().item() for v in bn.values()]), g['bn_sums'], rtol=1e-12)
    assert int(sd['encoder.Conv1.conv.1.num_batches_tracked']) == 2      # two encoder passes (D14)
    assert int(sd['cre.w_k.1.num_batches_tracked']) == 3                 # 1 + T cre calls
