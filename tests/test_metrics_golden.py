"""Eval-driver metrics ("next" row N2) pinned to the reference: tests/golden/metrics.npz holds inputs, the values of the reference's
own utils.util.dice_score_seperate / net.registration.NCC and the lines test_rpnet.py:231-251 prints (made by
tests/golden/make_golden_metrics.py from the unmodified reference).  CPU part: the line formatting and the Dice rule on plain
tensors; GPU part: the device reductions (rpnet_b200.evaluate._dice, rpnet_ncc_f32)."""
import numpy as np
import pytest
import torch


def _vals(g, j):
    return [None if np.isnan(x) else float(x) for x in g['v%d_vals' % j]]


def test_printed_lines_match_reference(golden):
    from rpnet_b200 import evaluate as E
    g = golden('metrics')
    T = int(g['T'])
    aff, few, refs = [], [], {k: [] for k in range(T)}
    for j in range(3):
        d, d2, dsc_affine, dsc_fewshot, *r = _vals(g, j)
        line = E.format_volume_line(j, 'p%03d' % j, 's%03d' % j, d, d2, dsc_affine, dsc_fewshot, r)
        assert line + '\n' == str(g['lines'][j]), (line, str(g['lines'][j]))
        if j < 2:
            aff.append(dsc_affine); few.append(dsc_fewshot)
            for k in range(T):
                refs[k].append(r[k])
    assert E.format_class_line('Liver', aff, few, refs) + '\n' == str(g['class_line'])


def test_dice_rule_matches_reference_on_cpu(golden):
    """volume.dice_sums / dice_from_sums == dice_score_seperate incl. raw (fractional) prediction values and the None rule."""
    from rpnet_b200 import volume as V
    g = golden('metrics')
    T = int(g['T'])
    for j in range(3):
        d, d2, dsc_affine, dsc_fewshot, *r = _vals(g, j)
        lab = torch.from_numpy(g['v%d_lab' % j].astype(np.int64))
        assert V.dice_from_sums(V.dice_sums(torch.from_numpy(g['v%d_appr' % j]), lab)) == dsc_affine
        preds = torch.from_numpy(g['v%d_preds' % j])
        for k in range(T):
            assert V.dice_from_sums(V.dice_sums(preds[k], lab)) == r[k]
        assert V.dice_from_sums(V.dice_sums(preds[T - 1], lab)) == dsc_fewshot
    # empty target, non-empty prediction: None (utils/util.py:384-388), not 0.0
    assert V.dice_from_sums(V.dice_sums(torch.ones(2, 2), torch.zeros(2, 2))) is None


@pytest.mark.gpu
def test_device_metrics_match_reference(golden):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import evaluate as E
    from rpnet_b200 import ops
    g = golden('metrics')
    T = int(g['T'])
    dev = torch.device('cuda:0')
    for j in range(3):
        d, d2, dsc_affine, dsc_fewshot, *r = _vals(g, j)
        lab = torch.from_numpy(g['v%d_lab' % j].astype(np.int64)).to(dev)
        assert E._dice(torch.from_numpy(g['v%d_appr' % j]).to(dev), lab) == dsc_affine
        preds = torch.from_numpy(g['v%d_preds' % j]).to(dev)
        for k in range(T):
            assert E._dice(preds[k], lab) == r[k]
        img = torch.from_numpy(g['v%d_img' % j]).to(dev)
        got = ops.ncc(img, torch.from_numpy(g['v%d_warped' % j]).to(dev)).item()
        got2 = ops.ncc(img, torch.from_numpy(g['v%d_supp' % j]).to(dev)).item()
        assert abs(got - d) <= 1e-5 * abs(d) and abs(got2 - d2) <= 1e-5 * abs(d2), (got, d, got2, d2)


def test_utils_shim_matches_reference(golden, tmp_path):
    """`from utils.util import Logger, load_yaml, dice_score_seperate` (test_rpnet.py:15,27,29) resolves to this repository and
    dice_score_seperate reproduces the reference's values (golden) on numpy inputs."""
    from utils.util import Logger, dice_score_seperate, load_yaml
    g = golden('metrics')
    T = int(g['T'])
    for j in range(3):
        d, d2, dsc_affine, dsc_fewshot, *r = _vals(g, j)
        lab = g['v%d_lab' % j].astype(np.int64)
        assert dice_score_seperate(g['v%d_appr' % j][None], lab[None], num_class=1)[0] == dsc_affine
        for k in range(T):
            assert dice_score_seperate(g['v%d_preds' % j][k].astype(np.int32)[None], lab[None], num_class=1)[0] == r[k]
    y = tmp_path / 'c.yml'
    y.write_text('net: RP_Net\nmask_feature_map: no\nn_iter_refinement: 4\n')
    d, a = load_yaml(str(y))
    assert d['mask_feature_map'] is False and a.net == 'RP_Net' and d['n_iter_refinement'] == 4
    import sys
    log = Logger(str(tmp_path / 'log'))
    log.write('x')
    log.flush()
    assert log.terminal is sys.stdout
    from net.registration import MSE, NCC  # noqa: F401  (test_rpnet.py:32)
    assert float(MSE(torch.ones(3), torch.zeros(3))) == 1.0
