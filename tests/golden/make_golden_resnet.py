"""Golden vectors for backbone 'resnet' (row N3) from the UNMODIFIED reference (needs /root/reference):
    python tests/golden/make_golden_resnet.py"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import, weights                             # noqa: E402
from rpnet_b200.synthetic import make_episode, perturb_bn_stats    # noqa: E402

warnings.filterwarnings('ignore')
ref = ref_import.load()
T = 2
cfg = dict(unet_normalize_type='BatchNorm2d', final_activation='sigmoid', mask_feature_map=False, n_iter_refinement=T,
           soft_mask=False, mask_refinement_correlation_radius=5)
torch.manual_seed(0)
net = ref.RP_Net(pretrained_path=None, cfg={'align': True, 'backbone': 'resnet'}, backbone_cfg=cfg).eval()
cs = weights.checksums(net.state_dict())
perturb_bn_stats(net.state_dict(), seed=1)
feats = []
net.encoder.register_forward_hook(lambda m, i, o: feats.append(o['d4']))
ep = make_episode(2, size=128, seed=4)
with torch.no_grad():
    out = net(ep['supp_imgs'], ep['fore_mask'], ep['back_mask'], ep['qry_imgs'], query_labels=ep['query_labels'],
              appr_query_labels=ep['appr_query_labels'])
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'resnet.npz')
np.savez_compressed(path, B=2, size=128, T=T, ep_seed=4, w_seed=0, bn_seed=1, keys=np.array(list(cs.keys())), vals=np.array(list(cs.values())),
                    d4_qry=feats[1][:, ::16, ::2, ::2].numpy(), ref0=out['refinement'][0][:, :, ::2, ::2].numpy(),
                    ref1=out['refinement'][1][:, :, ::2, ::2].numpy(), output=out['output'][:, :, ::2, ::2].numpy())
print('resnet.npz', os.path.getsize(path), 'bytes', len(cs), 'tensors')
