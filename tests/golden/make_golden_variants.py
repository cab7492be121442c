"""Golden vectors for the yaml variants of the U-Net encoder, from the UNMODIFIED reference modules:
`unet_normalize_type: InstanceNorm2d` (net/modules.py:49,52,69 through getattr(nn, ...)) and `mask_feature_map: x | x2 | x3`
(net/unet.py:401-424, 437-449).  Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_variants.py
Each fixture: eval forward of RP_Net (1-way 1-shot — the only setting the reference forward runs), 2 x 64 x 64, T = 2,
weights from torch.manual_seed(0) (the test rebuilds them with this repository's constructors and checks the checksums), BatchNorm
running statistics perturbed like the other goldens."""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import, weights                     # noqa: E402
from rpnet_b200.synthetic import make_episode, perturb_bn_stats  # noqa: E402

warnings.filterwarnings('ignore')
ref = ref_import.load()
OUT = os.path.dirname(os.path.abspath(__file__))
T, B, size = 2, 2, 64
out = {'T': T, 'B': B, 'size': size, 'ep_seed': 20, 'w_seed': 0, 'bn_seed': 1}
for name, extra in (('inorm', dict(unet_normalize_type='InstanceNorm2d')), ('mfm_x', dict(mask_feature_map='x')),
                    ('mfm_x2', dict(mask_feature_map='x2')), ('mfm_x3', dict(mask_feature_map='x3'))):
    cfg = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=T,
               soft_mask=False, mask_refinement_correlation_radius=5)
    cfg.update(extra)
    torch.manual_seed(0)
    net = ref.RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'UNet'}, backbone_cfg=cfg).eval()
    perturb_bn_stats(net.state_dict(), seed=1)
    cs = weights.checksums(net.state_dict())
    ep = make_episode(B, size=size, seed=20)
    with torch.no_grad():
        o = net(ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], query_labels=ep['query_labels'],
                appr_query_labels=ep['appr_query_labels'])
        d4 = net.encoder(ep['qry_imgs'][0], ep['fore_mask'][0][0].unsqueeze(1))['d4']
    out[name + '_keys'] = np.array(list(cs.keys()))
    out[name + '_cs'] = np.array(list(cs.values()))
    for i in range(T):
        out['%s_ref%d' % (name, i)] = o['refinement'][i].numpy()
    out[name + '_d4'] = d4[:, ::8, ::2, ::2].numpy()
    print(name, len(cs), 'state_dict tensors; logits range', o['output'].min().item(), o['output'].max().item())
path = os.path.join(OUT, 'variants.npz')
np.savez_compressed(path, **out)
print('variants.npz %.1f KB' % (os.path.getsize(path) / 1024))
