"""Golden vectors for the deformable-registration oracle (the `do_deformable: True` half of row N1), from the UNMODIFIED
reference classes run on the CPU.  Run in the build container only:  python tests/golden/make_golden_demons.py"""
import importlib
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
from oracle import ref_import                                 # noqa: E402
from rpnet_b200.synthetic import _slice                       # noqa: E402

ref_import.load()
reg = sys.modules.get('_rpnet_ref_net.registration') or importlib.import_module('net.registration')

size, iters = 48, 50
src = _slice(300, size, 1)[0]
lab = (_slice(300, size, 1)[1] > 0).float()
yy, xx = torch.meshgrid(torch.linspace(-1, 1, size), torch.linspace(-1, 1, size), indexing='ij')
bump = 0.08 * torch.exp(-((xx - 0.1) ** 2 + (yy + 0.05) ** 2) / 0.15)            # a smooth local deformation of the same slice
g = torch.stack([xx + bump, yy - 0.5 * bump], dim=-1)[None]
dst = F.grid_sample(src[None, None], g, padding_mode='border', align_corners=False)[0, 0]
src01, dst01 = ((src + 1) / 2)[None, None], ((dst + 1) / 2)[None, None]

def run(n_demons):
    torch.manual_seed(0)
    r = reg.AffineDemonsRegistration((size, size), use_diffeomorphic=True, use_GPU=False, stop_shear=False)
    opt_a = torch.optim.Adam(r.affine_reg.parameters(), lr=0.01)                  # dataset/few_shot_reader.py:147-148
    opt_d = torch.optim.Adam(r.demons.parameters(), lr=0.01)
    regul = reg.GaussianRegulariser([1, 1], sigma=[2, 2], dtype=torch.float32, device='cpu')
    r.train_registraion(src01, dst01, [opt_a, opt_d], regulariser=regul, iters=[iters, n_demons], regularise_displacement=False, verbose=False)
    return r, regul


# the optimisation is an Adam descent with a Gaussian smoothing after every step: two implementations that agree to 1e-7
# for ~25 iterations can part by a few 1e-4 later (one sign flip of a near-zero update, spread by the smoothing), so the
# flow is recorded half-way (tight check) and at the end (loose check)
flow_half = run(iters // 2)[0].demons.flow.detach().numpy()
r, regul = run(iters)
grid = reg.compute_grid((size, size))
with torch.no_grad():
    affined = r.affine_reg(src01)
    warped = r(src01, grid)
    warped_label = (r(lab[None, None], grid) > 0.1).float()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'demons.npz')
np.savez_compressed(out, size=size, iters=iters, dst=dst.numpy(), theta=r.affine_reg.theta.detach().numpy(), affined=affined[0, 0].numpy(),
                    flow=r.demons.flow.detach().numpy(), flow_half=flow_half, warped=warped[0, 0].numpy(),
                    warped_label=np.packbits(warped_label[0, 0].numpy().astype(np.uint8)), kernel=regul._kernel[0, 0].numpy(),
                    grid=grid.numpy())
print('demons.npz', os.path.getsize(out), 'bytes; |flow| max', r.demons.flow.detach().abs().max().item())
