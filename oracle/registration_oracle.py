"""CPU oracle for the affine registration in front of the hot path ("next" row N1).  TEST INFRASTRUCTURE ONLY.

Plain torch restatement of AffineRegistration (net/registration.py:316-357: theta = identity, forward = grid_sample o
affine_grid with torch's defaults, train_registraion = `iters` x {zero_grad, MSE, backward, optimizer.step}) as it is
driven by get_registration_field (dataset/few_shot_reader.py:109-198: Adam(lr=0.01), 50 iterations, images mapped to
[0, 1], warped label thresholded at 0.1).  Pinned against the reference classes themselves in
tests/golden/registration.npz (tests/golden/make_golden_registration.py)."""
import torch
import torch.nn.functional as F


def affine_forward(x, theta):
    """AffineRegistration.forward, net/registration.py:337-344 (stop_shear False)."""
    return F.grid_sample(x, F.affine_grid(theta, x.size(), align_corners=False), align_corners=False)


def affine_register(moving, fixed, iters=50, lr=0.01):
    """moving, fixed [1, 1, H, W] -> (theta [1, 2, 3], loss curve).  net/registration.py:347-357 + few_shot_reader.py:147."""
    theta = torch.zeros(1, 2, 3)
    theta[0, 0, 0] = 1
    theta[0, 1, 1] = 1
    theta.requires_grad_(True)
    opt = torch.optim.Adam([theta], lr=lr)
    curve = []
    for _ in range(iters):
        opt.zero_grad()
        loss = torch.mean((fixed - affine_forward(moving, theta)) ** 2)          # MSE, net/registration.py:147-154
        loss.backward()
        opt.step()
        curve.append(loss.item())
    return theta.detach(), curve


def get_affine_registration(query_images, support_images, support_labels, iters=50):
    """Affine outputs of get_registration_field (few_shot_reader.py:109-198), slice by slice like the reference."""
    src_all = (support_images[0][0][:, 0] + 1) / 2.0
    dst_all = (query_images[:, 0] + 1) / 2.0
    thetas, labels, srcs = [], [], []
    for s in range(dst_all.shape[0]):
        src, dst = src_all[s][None, None], dst_all[s][None, None]
        theta, _ = affine_register(src, dst, iters)
        with torch.no_grad():
            labels.append((affine_forward(support_labels[0][0][s][None, None].float(), theta) > 0.1).float()[0])
            srcs.append(affine_forward(src, theta)[0, 0] * 2 - 1)
        thetas.append(theta[0])
    return torch.stack(thetas), torch.stack(labels), torch.stack(srcs)


# ---------------------------------------------------------------------------------------------------------------------
# Deformable half of get_registration_field (`do_deformable: True`): DemonsRegistration + Diffeomorphic + NCC +
# GaussianRegulariser (net/registration.py:16-160,190-313).  Oracle only — the CUDA path for it is not built yet
# (rpnet_b200.registration raises NotImplementedError); pinned against the reference classes in tests/golden/demons.npz.
# ---------------------------------------------------------------------------------------------------------------------
import math                                                                               # noqa: E402

import numpy as np                                                                        # noqa: E402


def compute_grid(size):
    """net/registration.py:171-186 — identity grid [1, 2, H, W] (x first), corner-aligned normalisation 2 i / (n - 1) - 1."""
    h, w = int(size[0]), int(size[1])
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    return torch.stack([2 * (xs / (w - 1) - 0.5), 2 * (ys / (h - 1) - 0.5)])[None]


def gaussian_kernel_2d(sigma):
    """net/registration.py:16-51 — separable Gaussian, size 2 * ceil(2 sigma) + 1 per axis, each factor and the product normalised."""
    def k1(s):
        size = int(2 * np.ceil(s * 2) + 1)
        x = np.linspace(-(size - 1) // 2, (size - 1) // 2, num=size)
        k = 1.0 / (s * np.sqrt(2 * np.pi)) * np.exp(-(x ** 2) / (2 * s ** 2))
        return k / np.sum(k)
    k = np.tensordot(k1(sigma[0]), k1(sigma[1]), 0)
    return torch.tensor(k / np.sum(k), dtype=torch.float32)


def diffeomorphic_2d(displacement, grid, scaling=10):
    """Diffeomorphic.diffeomorphic_2D, net/registration.py:201-211: scaling and squaring, `scaling` compositions."""
    g = grid.permute(0, 2, 3, 1).contiguous()
    d = displacement / (2 ** scaling)
    for _ in range(scaling):
        d = d + F.grid_sample(d, d.permute(0, 2, 3, 1) + g, align_corners=False)
    return d


def demons_forward(x, flow, grid, scaling=10):
    """DemonsRegistration.forward, net/registration.py:244-258 (use_diffeomorphic=True): warp x by grid + exp(flow)."""
    locs = (grid + diffeomorphic_2d(flow, grid, scaling)).permute(0, 2, 3, 1)
    return F.grid_sample(x, locs, align_corners=False)


def ncc(moving, fixed):
    """NCC, net/registration.py:157-160 (global, negated, 1e-10 inside the square root)."""
    fm, mm = fixed - fixed.mean(), moving - moving.mean()
    return -1.0 * torch.sum(fm * mm) / torch.sqrt(torch.sum(fm ** 2) * torch.sum(mm ** 2) + 1e-10)


def demons_register(moving, fixed, iters=50, lr=0.01, sigma=(2, 2)):
    """DemonsRegistration.train_registraion (net/registration.py:290-313) as get_registration_field drives it
    (few_shot_reader.py:148-163): Adam(lr 0.01) on the flow, NCC loss, Gaussian smoothing of the flow after every step.
    moving (already affinely warped), fixed: [1, 1, H, W] in [0, 1].  Returns (flow [1, 2, H, W], loss curve)."""
    h, w = moving.shape[2:]
    flow = torch.zeros(1, 2, h, w, requires_grad=True)
    opt = torch.optim.Adam([flow], lr=lr)
    k = gaussian_kernel_2d(sigma)
    pad = [(k.shape[0] - 1) // 2, (k.shape[1] - 1) // 2]
    kern = k[None, None].expand(2, -1, -1, -1).contiguous()
    curve = []
    for _ in range(iters):
        grid = compute_grid((h, w))
        opt.zero_grad()
        loss = ncc(demons_forward(moving, flow, grid), fixed)
        loss.backward()
        opt.step()
        with torch.no_grad():                                                             # GaussianRegulariser._regularise_2d
            flow.data = F.conv2d(flow.data, kern, padding=pad, groups=2)
        curve.append(loss.item())
    return flow.detach(), curve
