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
