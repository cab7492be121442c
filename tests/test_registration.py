"""Affine registration in front of the hot path ("next" row N1).
CPU: the oracle restatement against golden vectors recorded from the reference's AffineRegistration
(tests/golden/registration.npz).  GPU: the one-launch batched kernel against the same golden vectors and the oracle."""
import numpy as np
import pytest
import torch


def _inputs(g):
    from rpnet_b200.synthetic import _slice
    S, size = int(g['S']), int(g['size'])
    src = torch.stack([_slice(int(g['src_seed']) + s, size, 1)[0] for s in range(S)])
    lab = torch.stack([(_slice(int(g['src_seed']) + s, size, 1)[1] > 0).float() for s in range(S)])
    return src, lab, torch.from_numpy(g['dst'])


def test_oracle_affine_registration_vs_reference_golden(golden):
    from oracle import registration_oracle as R
    g = golden('registration')
    src, lab, dst = _inputs(g)
    for s in range(int(g['S'])):
        theta, curve = R.affine_register(((src[s] + 1) / 2)[None, None], ((dst[s] + 1) / 2)[None, None], int(g['iters']))
        np.testing.assert_allclose(theta[0].numpy(), g['theta'][s], rtol=0, atol=1e-6)
        np.testing.assert_allclose(np.array(curve), g['loss'][s], rtol=1e-5, atol=1e-9)
    th, wl, ws = R.get_affine_registration(dst[:, None], [[src[:, None]]], [[lab]], int(g['iters']))
    n = wl.numel()
    assert np.array_equal(np.packbits(wl[:, 0].numpy().astype(np.uint8)), g['warped_label'])
    np.testing.assert_allclose(ws[:, ::2, ::2].numpy(), g['warped_src'], atol=1e-6)


@pytest.mark.gpu
def test_affine_registration_kernel_vs_reference_golden(golden):
    """One launch registers all slices (one CTA per slice, every Adam iteration inside the kernel).  The optimisation is
    smooth (6 parameters, MSE): the only difference to the reference is the summation order of the gradient (1e-6 relative),
    which Adam's normalised steps keep small: theta within 2e-3 after 50 iterations, loss curve within 1e-3 relative."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import registration as RG
    g = golden('registration')
    src, lab, dst = _inputs(g)
    dev = torch.device('cuda:0')
    iters = int(g['iters'])
    theta, curve = RG.affine_register(((src + 1) / 2).to(dev), ((dst + 1) / 2).to(dev), iters=iters, return_loss=True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(curve.cpu().numpy()[:, 0], g['loss'][:, 0], rtol=1e-5)          # identity warp: same loss
    np.testing.assert_allclose(curve.cpu().numpy(), g['loss'], rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(theta.cpu().numpy(), g['theta'], rtol=0, atol=2e-3)
    th, wl, ws = RG.get_affine_registration(dst[:, None].to(dev), [[src[:, None].to(dev)]], [[lab.to(dev)]], iters)
    ref_wl = np.unpackbits(g['warped_label'])[:wl.numel()].reshape(wl[:, 0].shape)
    assert (wl[:, 0].cpu().numpy().astype(np.uint8) != ref_wl).mean() < 2e-3                      # label edge pixels only
    np.testing.assert_allclose(ws.cpu().numpy()[:, ::2, ::2], g['warped_src'], atol=2e-2)
    # the warp kernel alone, with the reference's theta: exact up to fp32 rounding
    tref = torch.from_numpy(g['theta']).to(dev)
    w2 = RG.affine_warp(((src + 1) / 2)[:, None].to(dev), tref)[:, 0] * 2 - 1
    np.testing.assert_allclose(w2.cpu().numpy()[:, ::2, ::2], g['warped_src'], atol=1e-5)


@pytest.mark.gpu
def test_affine_registration_batched_equals_per_slice():
    """Slices are independent: registering a batch == registering each slice on its own (bit for bit), any image size."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import registration as RG
    gen = torch.Generator().manual_seed(3)
    mov = torch.rand(5, 48, 80, generator=gen).cuda()
    fix = torch.roll(mov, shifts=(2, -3), dims=(1, 2)) * 0.9 + 0.05
    th = RG.affine_register(mov, fix, iters=20)
    for s in range(5):
        assert torch.equal(RG.affine_register(mov[s:s + 1], fix[s:s + 1], iters=20)[0], th[s])
    assert th.shape == (5, 2, 3) and torch.isfinite(th).all()


def test_demons_oracle_vs_reference_golden():
    """Deformable half of get_registration_field (`do_deformable: True`): the oracle restatement of DemonsRegistration +
    Diffeomorphic + NCC + GaussianRegulariser reproduces the reference classes (tests/golden/make_golden_demons.py).
    Oracle only: the CUDA path is not built yet (rpnet_b200.registration raises NotImplementedError)."""
    import os
    import torch
    from oracle import registration_oracle as R
    from rpnet_b200.synthetic import _slice
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'demons.npz'))
    size, iters = int(g['size']), int(g['iters'])
    np.testing.assert_allclose(R.gaussian_kernel_2d((2, 2)).numpy(), g['kernel'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(R.compute_grid((size, size)).numpy(), g['grid'], rtol=0, atol=0)
    src = _slice(300, size, 1)[0]
    lab = (_slice(300, size, 1)[1] > 0).float()
    src01, dst01 = ((src + 1) / 2)[None, None], ((torch.from_numpy(g['dst']) + 1) / 2)[None, None]
    theta, _ = R.affine_register(src01, dst01, iters)
    np.testing.assert_allclose(theta.numpy(), g['theta'], rtol=0, atol=1e-6)
    with torch.no_grad():
        affined = R.affine_forward(src01, theta)
    np.testing.assert_allclose(affined[0, 0].numpy(), g['affined'], atol=1e-6)
    # tight half-way, loose at the end: an Adam descent with smoothing after every step lets two implementations that agree
    # to 1e-7 part by a few 1e-4 late in the run (one sign flip of a near-zero update, spread by the Gaussian)
    flow_half, _ = R.demons_register(affined, dst01, iters // 2)
    np.testing.assert_allclose(flow_half.numpy(), g['flow_half'], rtol=0, atol=2e-6)
    flow, curve = R.demons_register(affined, dst01, iters)
    assert curve[-1] < curve[0]                                                  # NCC improves
    assert np.abs(flow.numpy() - g['flow']).max() < 0.01 * np.abs(g['flow']).max()
    grid = R.compute_grid((size, size))
    with torch.no_grad():
        warped = R.demons_forward(affined, flow, grid)
        wl = (R.demons_forward(R.affine_forward(lab[None, None], theta), flow, grid) > 0.1).float()
    np.testing.assert_allclose(warped[0, 0].numpy(), g['warped'], atol=2e-3)
    ref_wl = np.unpackbits(g['warped_label'])[:size * size].reshape(size, size)
    assert (wl[0, 0].numpy().astype(np.uint8) != ref_wl).mean() < 2e-3


@pytest.mark.gpu
def test_demons_kernel_vs_reference_golden():
    """The one-launch demons registration (rpnet_demons_register_f32: exp(flow) by scaling and squaring, NCC, hand-derived
    backward, Adam, Gaussian smoothing — net/registration.py:190-313 driven as in few_shot_reader.py:137-170) against the flow,
    the warped image and the warped label recorded from the UNMODIFIED reference classes (tests/golden/demons.npz).  Same
    tolerances as the oracle's own pin: tight half-way through the descent, 1 % of the flow scale at the end."""
    import os
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from rpnet_b200 import registration as RG
    from rpnet_b200.synthetic import _slice
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'demons.npz'))
    size, iters = int(g['size']), int(g['iters'])
    dev = torch.device('cuda:0')
    np.testing.assert_allclose(RG.gaussian_kernel_2d((2, 2)).numpy(), g['kernel'], rtol=0, atol=1e-9)
    src = _slice(300, size, 1)[0]
    lab = (_slice(300, size, 1)[1] > 0).float()
    src01, dst01 = ((src + 1) / 2)[None].to(dev), ((torch.from_numpy(g['dst']) + 1) / 2)[None].to(dev)
    theta = torch.from_numpy(g['theta']).to(dev)
    affined = RG.affine_warp(src01[:, None], theta)                      # the affine stage is pinned in test_affine_* above
    np.testing.assert_allclose(affined[0, 0].cpu().numpy(), g['affined'], atol=1e-5)
    flow_half, _ = RG.demons_register(affined[:, 0], dst01, iters=iters // 2)
    np.testing.assert_allclose(flow_half.cpu().numpy(), g['flow_half'], rtol=0, atol=5e-6)
    flow, disp, curve = RG.demons_register(affined[:, 0], dst01, iters=iters, return_loss=True)
    torch.cuda.synchronize()
    assert curve[0, -1] < curve[0, 0]                                      # NCC improves
    assert np.abs(flow.cpu().numpy() - g['flow']).max() < 0.01 * np.abs(g['flow']).max()
    warped = RG.demons_warp(affined, disp)
    np.testing.assert_allclose(warped[0, 0].cpu().numpy(), g['warped'], atol=2e-3)
    wl = (RG.demons_warp(RG.affine_warp(lab[None, None].to(dev), theta), disp) > 0.1).float()
    ref_wl = np.unpackbits(g['warped_label'])[:size * size].reshape(size, size)
    assert (wl[0, 0].cpu().numpy().astype(np.uint8) != ref_wl).mean() < 2e-3
    # zero iterations == the untrained module: a pure n / (n - 1) resampling (do_deformable: False, demons_identity_theta)
    f0, d0 = RG.demons_register(affined[:, 0], dst01, iters=0)
    assert float(f0.abs().max()) == 0.0 and float(d0.abs().max()) == 0.0
    z = RG.affine_warp(affined, RG.demons_identity_theta(1, size, size, dev))
    np.testing.assert_allclose(RG.demons_warp(affined, d0).cpu().numpy(), z.cpu().numpy(), atol=1e-6)


@pytest.mark.gpu
def test_demons_batch_vs_oracle_rectangular():
    """Slices are independent (batch == per-slice results) and non-square images work: three 40 x 56 slices against the CPU
    oracle (oracle/registration_oracle.py:demons_register) after 12 iterations."""
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from oracle import registration_oracle as R
    from rpnet_b200 import registration as RG
    gen = torch.Generator().manual_seed(5)
    h, w, n, iters = 40, 56, 3, 12
    base = torch.rand(n, 1, h // 4, w // 4, generator=gen)
    mov = torch.nn.functional.interpolate(base, size=(h, w), mode='bilinear', align_corners=False)[:, 0]
    fix = torch.roll(mov, shifts=(1, -2), dims=(1, 2)) * 0.95 + 0.02
    flow, disp = RG.demons_register(mov.cuda(), fix.cuda(), iters=iters)
    for s in range(n):
        ref, _ = R.demons_register(mov[s][None, None], fix[s][None, None], iters)
        assert (flow[s].cpu() - ref[0]).abs().max().item() < 2e-5 + 1e-3 * ref.abs().max().item(), s
        one, _ = RG.demons_register(mov[s:s + 1].cuda(), fix[s:s + 1].cuda(), iters=iters)
        assert (one[0] - flow[s]).abs().max().item() < 1e-6          # float atomics in the scatter: not bit-exact run to run
