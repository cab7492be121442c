"""Batched affine registration of support slices onto query slices ("next" row N1 of SURVEY §8f): the affine half of
`get_registration_field` (dataset/few_shot_reader.py:109-198) and `AffineRegistration` (net/registration.py:316-357).

The reference registers one slice at a time inside `Dataset.__getitem__`: per slice 50 Adam iterations, each a handful of
tiny launches (affine_grid, grid_sample, MSE, backward, Adam).  Here all slices of a volume are registered by ONE kernel
launch (one CTA per slice runs all iterations, ops.affine_register) and warped by one more (ops.affine_warp).

The deformable refinement (DemonsRegistration, net/registration.py:221-313: dense flow field + NCC + scaling-and-squaring)
is not built: this module reproduces the `do_deformable: False` behaviour, whose outputs are the affine ones."""
import torch

from . import ops


def affine_register(moving, fixed, iters=50, lr=0.01, return_loss=False):
    """moving, fixed: [n, 1, h, w] or [n, h, w] CUDA fp32 in the reference's [0, 1] range -> theta [n, 2, 3]
    (AffineRegistration.train_registraion with torch.optim.Adam(lr=0.01), MSE loss: few_shot_reader.py:147,153-163)."""
    m = moving.reshape(moving.shape[0], moving.shape[-2], moving.shape[-1]).float().contiguous()
    f = fixed.reshape(m.shape).float().contiguous()
    theta = torch.empty(m.shape[0], 2, 3, dtype=torch.float32, device=m.device)
    curve = torch.empty(m.shape[0], iters, dtype=torch.float32, device=m.device) if return_loss else None
    ops.affine_register(m, f, theta, iters=iters, lr=lr, loss_curve=curve)
    return (theta, curve) if return_loss else theta


def affine_warp(x, theta):
    """AffineRegistration.forward (net/registration.py:337-344) for a batch: x [n, c, h, w] fp32, theta [n, 2, 3]."""
    x = x.float().contiguous()
    out = torch.empty_like(x)
    ops.affine_warp(x, theta.contiguous(), out)
    return out


def get_affine_registration(query_images, support_images, support_labels, iters=50):
    """The affine outputs of get_registration_field (few_shot_reader.py:109-198), batched on the device:
    query_images S x 1 x H x W and support_images [[S x 1 x H x W]] in [-1, 1], support_labels [[S x H x W]].
    Returns (theta [S,2,3], affine_warped_label [S,1,H,W] in {0,1}, affine_warped_src [S,H,W] in [-1,1])
    — `py_affine_reg_pred` and `affine_warped_src_list` of the reference (:170-172,178-179,193-196)."""
    src = (support_images[0][0][:, 0].float() + 1) / 2.0                     # :111-116 intensities to [0, 1]
    dst = (query_images[:, 0].float() + 1) / 2.0
    theta = affine_register(src, dst, iters=iters, lr=0.01)
    lab = support_labels[0][0].float()[:, None]
    warped_label = (affine_warp(lab, theta) > 0.1).float()                    # :171-172
    warped_src = affine_warp(src[:, None], theta)[:, 0] * 2 - 1               # :178-179, :196
    return theta, warped_label, warped_src
