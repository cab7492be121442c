"""Golden vectors for the eval-driver metrics ("next" row N2), from the UNMODIFIED reference functions:
utils.util.dice_score_seperate (utils/util.py:379-390), net.registration.NCC (net/registration.py:157-160) and the lines
test_rpnet.py:231-251 prints (their f-strings are inline in eval(); they are re-typed below character for character and
evaluated on the reference's own metric values).  Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_metrics.py"""
import io
import os
import sys
from collections import defaultdict
from contextlib import redirect_stdout

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import                                 # noqa: E402
from rpnet_b200.synthetic import _slice                       # noqa: E402

ref = ref_import.load()
g = torch.Generator().manual_seed(3)
S, size, T = 4, 32, 3
cases = []
for v in range(3):                                            # three "volumes": normal, soft appr label, empty target
    lab = torch.stack([_slice(200 + 10 * v + s, size, 1)[1] for s in range(S)])
    if v == 2:
        lab = torch.zeros_like(lab)
    img = torch.stack([_slice(200 + 10 * v + s, size, 1)[0] for s in range(S)])[:, None]
    warped = torch.roll(img, shifts=(2, -1), dims=(2, 3)) + 0.01 * torch.randn(img.shape, generator=g)
    supp = torch.stack([_slice(300 + 10 * v + s, size, 1)[0] for s in range(S)])[:, None]
    appr = torch.roll((lab > 0).float(), shifts=(3, -2), dims=(1, 2))
    if v == 1:
        appr = torch.nn.functional.avg_pool2d(appr[:, None], 3, 1, 1)[:, 0]           # fractional values, like a bilinear warp
    preds = [(torch.roll((lab > 0), shifts=(k, 0), dims=(1, 2)) | (torch.rand(lab.shape, generator=g) > 0.995)).to(torch.uint8) for k in range(T)]
    cases.append(dict(lab=lab, img=img, warped=warped, supp=supp, appr=appr, preds=preds))

out = {}
lines = []
eval_classes = ['Liver']
dsc_list = defaultdict(list)
dsc_affine_list, dsc_fewshot_list = defaultdict(list), defaultdict(list)
dsc_refinement_list = defaultdict(lambda: defaultdict(list))
for j, c in enumerate(cases):
    query_labels, appr_query_labels = c['lab'].long(), c['appr']
    fewshot_pred = c['preds'][T - 1][None].numpy().astype(np.float32)
    # test_rpnet.py:226-230
    dsc_affine = ref.dice_score_seperate(appr_query_labels.cpu().data.numpy()[None, ...], query_labels.cpu().data.numpy()[None, ...], num_class=1)[0]
    dsc_fewshot = ref.dice_score_seperate(fewshot_pred, query_labels.cpu().data.numpy()[None, ...], num_class=1)[0]
    d = ref.NCC(c['img'], c['warped']).item()
    d2 = ref.NCC(c['img'], c['supp']).item()
    pid, supp_pid = 'p%03d' % j, 's%03d' % j
    buf = io.StringIO()
    with redirect_stdout(buf):
        print(f'{j} {pid} {supp_pid} affine ({d}, {d2}) {dsc_affine}, fewshot {dsc_fewshot}', end=' ')          # :231
        dsc_affine_list['Liver'].append(dsc_affine)
        dsc_fewshot_list['Liver'].append(dsc_fewshot)
        refs = []
        for k in range(T):
            s = ref.dice_score_seperate(c['preds'][k].numpy().astype(np.int32)[None, ...], query_labels.cpu().data.numpy()[None, ...], num_class=1)[0]   # :238
            dsc_refinement_list['Liver'][k].append(s)
            refs.append(s)
            print(f'ref {k} {s}, ', end=' ')                                                                        # :240
        print()                                                                                                    # :242
    lines.append(buf.getvalue())
    for name, val in (('lab', c['lab'].numpy().astype(np.int8)), ('img', c['img'].numpy()), ('warped', c['warped'].numpy()),
                      ('supp', c['supp'].numpy()), ('appr', c['appr'].numpy()),
                      ('preds', torch.stack(c['preds']).numpy())):
        out['v%d_%s' % (j, name)] = val
    out['v%d_vals' % j] = np.array([np.nan if x is None else x for x in [d, d2, dsc_affine, dsc_fewshot] + refs], dtype=np.float64)
# the per-class summary is only printable when no volume has an empty target (np.average over None raises in the reference too)
keep = [0, 1]
buf = io.StringIO()
with redirect_stdout(buf):
    for k in eval_classes:                                                                                         # :246-251
        v = dsc_list[k]
        a = [dsc_affine_list[k][i] for i in keep]
        f = [dsc_fewshot_list[k][i] for i in keep]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            print(f'{k}, affine {np.average(a)}, voxel morph {np.average(v)}, {np.std(v)}, fewshot {np.average(f)}', end=' ')
        for r, l in dsc_refinement_list[k].items():
            print(f'ref {r} {np.average([l[i] for i in keep])}, ', end=' ')
        print()
out['lines'] = np.array(lines)
out['class_line'] = np.array(buf.getvalue())
out['T'] = np.array(T)
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'metrics.npz'), **out)
print(''.join(lines) + buf.getvalue())
