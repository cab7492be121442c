"""Golden vectors for the affine registration oracle, from the UNMODIFIED reference classes.
Run in the build container only (needs /root/reference):  python tests/golden/make_golden_registration.py"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import                                 # noqa: E402
from rpnet_b200.synthetic import _slice                       # noqa: E402

warnings.filterwarnings('ignore')
ref_import.load()
import importlib                                              # noqa: E402
reg = importlib.import_module('net.registration') if 'net.registration' in sys.modules else None
if reg is None:
    sys.path.insert(0, '/root/reference')
    reg = importlib.import_module('net.registration')

size, S, iters = 64, 3, 50
src = torch.stack([_slice(100 + s, size, 1)[0] for s in range(S)])
lab = torch.stack([(_slice(100 + s, size, 1)[1] > 0).float() for s in range(S)])
# fixed = an affinely deformed, re-noised copy of another slice of the same "anatomy": the optimisation has real work to do
import torch.nn.functional as F                               # noqa: E402
TRUE = torch.tensor([[[1.08, 0.05, 0.12], [-0.04, 0.95, -0.08]], [[0.93, -0.06, -0.10], [0.07, 1.05, 0.06]],
                     [[1.00, 0.10, 0.05], [-0.10, 1.00, 0.15]]])
base = torch.stack([_slice(100 + s, size, 1)[0] for s in range(S)])[:, None]
g = torch.Generator().manual_seed(7)
dst = (F.grid_sample(base, F.affine_grid(TRUE, base.size(), align_corners=False), padding_mode='border', align_corners=False)[:, 0]
       + 0.01 * torch.randn(S, size, size, generator=g))
src01, dst01 = (src + 1) / 2, (dst + 1) / 2
thetas, curves, wl, ws = [], [], [], []
for s in range(S):
    torch.manual_seed(0)
    a = reg.AffineRegistration((size, size))
    opt = torch.optim.Adam(a.parameters(), lr=0.01)            # dataset/few_shot_reader.py:147
    losses = []
    for i in range(iters):                                     # AffineRegistration.train_registraion, unrolled to record the loss
        opt.zero_grad()
        loss = reg.MSE(a(src01[s][None, None]), dst01[s][None, None], mask=None)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    with torch.no_grad():
        thetas.append(a.theta.detach()[0].clone())
        wl.append((a(lab[s][None, None]) > 0.1).float()[0, 0])
        ws.append(a(src01[s][None, None])[0, 0] * 2 - 1)
    curves.append(losses)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'registration.npz')
np.savez_compressed(out, size=size, S=S, iters=iters, src_seed=100, dst=dst.numpy(), theta=torch.stack(thetas).numpy(),
                    loss=np.array(curves), warped_label=np.packbits(torch.stack(wl).numpy().astype(np.uint8)),
                    warped_src=torch.stack(ws).numpy()[:, ::2, ::2])
print('registration.npz', os.path.getsize(out), 'bytes; final thetas:\n', torch.stack(thetas))
