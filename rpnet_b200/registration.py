"""Batched affine registration of support slices onto query slices ("next" row N1 of SURVEY §8f): the affine half of
`get_registration_field` (dataset/few_shot_reader.py:109-198) and `AffineRegistration` (net/registration.py:316-357).

The reference registers one slice at a time inside `Dataset.__getitem__`: per slice 50 Adam iterations, each a handful of
tiny launches (affine_grid, grid_sample, MSE, backward, Adam).  Here all slices of a volume are registered by ONE kernel
launch (one CTA per slice runs all iterations, ops.affine_register) and warped by one more (ops.affine_warp).

The deformable refinement (DemonsRegistration, net/registration.py:221-313: dense flow field + NCC + scaling-and-squaring +
Gaussian regulariser) is one more launch for all slices (ops.demons_register: one CTA per slice runs the 50 iterations).  With
`do_deformable: False` (yamls/example.yml) the demons stage runs zero iterations but is still *applied* — with a zero flow it
resamples by n/(n-1) (`demons_identity_theta`), which is what separates `warped_supp_label` / `appr_query_labels` from the
affine outputs."""
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


def compute_grid(img_size, device=None):
    """net/registration.py:171-186: the identity sampling grid [1, 2, H, W] (x first), normalised as 2 * (i / (n - 1) - 0.5)."""
    h, w = int(img_size[0]), int(img_size[1])
    # built on the host like the reference's (CUDA divides by a scalar through its reciprocal: 1-ulp differences), then moved
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing='ij')
    grid = torch.stack([2 * (xs / (w - 1) - 0.5), 2 * (ys / (h - 1) - 0.5)])[None]
    return grid if device is None else grid.to(device)


def demons_identity_theta(n, h, w, device):
    """What an untrained DemonsRegistration does to its input (net/registration.py:244-258 with flow == 0): it samples with
    `compute_grid` — corner-aligned coordinates 2i/(n-1) - 1 — through F.grid_sample's default align_corners=False, i.e. a
    zoom by n/(n-1) about the centre with zero padding.  In affine_grid/grid_sample(align_corners=False) terms that is
    theta = diag(w/(w-1), h/(h-1))."""
    t = torch.zeros(n, 2, 3, dtype=torch.float32, device=device)
    t[:, 0, 0] = w / (w - 1.0)
    t[:, 1, 1] = h / (h - 1.0)
    return t


def gaussian_kernel_2d(sigma=(2, 2)):
    """The regulariser's kernel (net/registration.py:16-51): per axis size 2 * ceil(2 sigma) + 1 over linspace(-(size-1)//2, (size-1)//2),
    each 1-D factor and the outer product normalised to sum 1.  CPU fp32 tensor [ky, kx]."""
    import numpy as np

    def k1(s):
        size = int(2 * np.ceil(s * 2) + 1)
        x = np.linspace(-(size - 1) // 2, (size - 1) // 2, num=size)
        k = 1.0 / (s * np.sqrt(2 * np.pi)) * np.exp(-(x ** 2) / (2 * s ** 2))
        return k / np.sum(k)
    k = np.tensordot(k1(sigma[0]), k1(sigma[1]), 0)
    return torch.tensor(k / np.sum(k), dtype=torch.float32)


def demons_register(moving, fixed, iters=50, lr=0.01, sigma=(2, 2), return_loss=False):
    """DemonsRegistration(use_diffeomorphic=True).train_registraion as get_registration_field drives it (few_shot_reader.py:137-163:
    NCC loss, Adam(lr 0.01), GaussianRegulariser(sigma 2)) for a batch of slices: moving (affinely warped already), fixed
    [n, h, w] or [n, 1, h, w] CUDA fp32 in [0, 1] -> (flow [n, 2, h, w], disp = exp(flow) [n, 2, h, w])."""
    m = moving.reshape(moving.shape[0], moving.shape[-2], moving.shape[-1]).float().contiguous()
    f = fixed.reshape(m.shape).float().contiguous()
    curve = torch.empty(m.shape[0], iters, dtype=torch.float32, device=m.device) if return_loss else None
    flow, disp = ops.demons_register(m, f, gaussian_kernel_2d(sigma), iters=iters, lr=lr, loss_curve=curve)
    return (flow, disp, curve) if return_loss else (flow, disp)


def demons_warp(x, disp):
    """DemonsRegistration.forward (net/registration.py:244-258) for a batch: x [n, c, h, w], disp = exp(flow) [n, 2, h, w]."""
    x = x.float().contiguous()
    out = torch.empty_like(x)
    ops.demons_warp(x, disp.contiguous(), out)
    return out


def get_registration_field(query_images, support_images, support_labels, do_deformable=True):
    """get_registration_field (dataset/few_shot_reader.py:109-198), all slices in one launch per stage.
    Same return order as the reference:
      registration_field        list of [theta_i [2, 3], grid [1, 2, H, W]] per slice (the reference keeps its nn.Module there;
                                RP_Net.forward ignores the argument, net/rp_net.py:226)
      py_reg_pred               [S, 1, H, W] {0, 1}: (demons(affine(label)) > 0.1) with the demons flow still zero (:168-170)
      warped_src_list           [S, H, W] in [-1, 1]: demons(affine(src)) * 2 - 1 (:175-176, :191)
      py_affine_reg_pred        [S, 1, H, W] {0, 1}: affine(label) > 0.1 (:171-173)
      affine_warped_src_list    [S, H, W] in [-1, 1] (:178-179, :196)
    The inputs may live on the host; the outputs stay on the CUDA device."""
    dev = torch.device('cuda', torch.cuda.current_device())
    src = (support_images[0][0][:, 0].to(dev, torch.float32) + 1) / 2.0
    dst = (query_images[:, 0].to(dev, torch.float32) + 1) / 2.0
    lab = support_labels[0][0].to(dev, torch.float32)[:, None]
    n, h, w = src.shape
    theta = affine_register(src, dst, iters=50, lr=0.01)
    aff_lab, aff_src = affine_warp(lab, theta), affine_warp(src[:, None], theta)
    grid = compute_grid((h, w), dev)
    if do_deformable:
        # :137-163 — 50 demons iterations on the affinely warped source; the trained module is then applied to label and source
        flow, disp = demons_register(aff_src[:, 0], dst, iters=50, lr=0.01)
        reg_pred = (demons_warp(aff_lab, disp) > 0.1).float()
        warped_src = demons_warp(aff_src, disp)[:, 0] * 2 - 1
        field = [[theta[i], grid, flow[i]] for i in range(n)]
    else:
        zoom = demons_identity_theta(n, h, w, dev)
        reg_pred = (affine_warp(aff_lab, zoom) > 0.1).float()
        warped_src = affine_warp(aff_src, zoom)[:, 0] * 2 - 1
        field = [[theta[i], grid] for i in range(n)]
    return field, reg_pred, warped_src, (aff_lab > 0.1).float(), aff_src[:, 0] * 2 - 1


# ---------------------------------------------------------------------------------------------------------------------
# The two similarity measures test_rpnet.py imports from net.registration (:32) with the reference's signatures.
# ---------------------------------------------------------------------------------------------------------------------
def NCC(moving_image_valid, fixed_image_valid, mask=None):
    """net/registration.py:157-160: minus the normalised cross-correlation of the two tensors (`mask` is ignored there too).
    CUDA inputs go through rpnet_ncc_f32 (one fused reduction); the result is a 0-dim tensor like the reference's."""
    m, f = moving_image_valid, fixed_image_valid
    if not (m.is_cuda and f.is_cuda):
        raise RuntimeError('rpnet_b200 runs on CUDA only: move the images to the GPU (test_rpnet.py:229-230 passes CUDA tensors)')
    with torch.cuda.device(m.device):
        return ops.ncc(m.float().contiguous(), f.float().contiguous().reshape(m.shape))[0]


def MSE(y_pred, y_true, mask=None):
    """net/registration.py:147-154 (plain tensor arithmetic: loss plumbing of the registration, not on the hot path)."""
    value = torch.mean((y_true - y_pred) ** 2)
    if mask is not None:
        return torch.masked_select(value, mask).mean()
    return value
