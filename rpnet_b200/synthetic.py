"""Deterministic CT-like phantom episodes (synthetic stand-in for the private ABD-110 data).

Mimics the dataset item contract of the reference episode builder
(dataset/few_shot_reader.py:629-647): support slices + {0,1} support labels per way/shot,
query slices, int64 query labels and `appr_query_labels` (the registered support label,
few_shot_reader.py:608 — emulated by the true query mask shifted by (+6, -5) px).
Intensities follow the reference normalisation rule clip(HU, -1024, 3072) -> [-1, 1]
(utils/util.py:455-467, yamls/example.yml:29-31).  Spec: SURVEY.md §8(d).
"""
import torch

_ORGAN_SLOTS = [(-0.30, -0.10), (0.30, -0.10), (-0.25, 0.22), (0.28, 0.20)]   # organ centres per way


def _slice(seed, size, ways):
    """One slice: returns (image HxW in [-1,1], label HxW int64 with 0=bg, w+1=organ w)."""
    g = torch.Generator().manual_seed(int(seed))
    lin = torch.linspace(-1.0, 1.0, size)
    yy, xx = torch.meshgrid(lin, lin, indexing='ij')
    hu = torch.full((size, size), -1000.0)
    hu[((xx / 0.85) ** 2 + (yy / 0.65) ** 2) < 1] = 40.0
    label = torch.zeros(size, size, dtype=torch.int64)
    jit = (torch.rand(ways, 4, generator=g) - 0.5)
    for w in range(ways):
        cx, cy = _ORGAN_SLOTS[w % len(_ORGAN_SLOTS)] if ways > 1 else (0.0, 0.0)
        cx = cx + 0.3 * jit[w, 0].item() * (0.5 if ways > 1 else 1.0)
        cy = cy + 0.3 * jit[w, 1].item() * (0.5 if ways > 1 else 1.0)
        rx = (0.24 if ways == 1 else 0.13) * (1 + 0.3 * jit[w, 2].item())
        ry = (0.17 if ways == 1 else 0.10) * (1 + 0.3 * jit[w, 3].item())
        organ = (((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2) < 1
        hu[organ] = 60.0 + 60.0 * (w + 1) / ways
        label[organ] = w + 1
    hu = hu + torch.randn(size, size, generator=g) * 20.0
    img = (hu.clamp(-1024.0, 3072.0) + 1024.0) / 4096.0 * 2.0 - 1.0
    return img, label


def make_episode(batch, ways=1, shots=1, size=256, seed=0, dtype=torch.float32):
    """Returns a dict with the exact argument structure RP_Net.forward takes
    (net/rp_net.py:226-236): supp_imgs / fore_mask / back_mask are way x shot lists of
    B x 1 x H x W / B x H x W tensors, qry_imgs is a 1-list of B x 1 x H x W."""
    supp_imgs, fore_mask, back_mask = [], [], []
    for w in range(ways):
        si, fm, bm = [], [], []
        for s in range(shots):
            imgs, fgs = [], []
            for b in range(batch):
                img, lab = _slice(seed + 1 + b + 97 * s + 389 * w, size, ways)
                imgs.append(img)
                fgs.append((lab == w + 1).float())
            fg = torch.stack(fgs)
            si.append(torch.stack(imgs)[:, None].to(dtype))
            fm.append(fg.to(dtype))
            bm.append((1 - fg).to(dtype))       # test_rpnet.py:170 back = 1 - fore
        supp_imgs.append(si)
        fore_mask.append(fm)
        back_mask.append(bm)
    q, ql = zip(*[_slice(seed + 1001 + b, size, ways) for b in range(batch)])
    qry = torch.stack(q)[:, None].to(dtype)
    query_labels = torch.stack(ql)
    appr = torch.roll((query_labels > 0).to(dtype), shifts=(6, -5), dims=(1, 2))
    return {'supp_imgs': supp_imgs, 'fore_mask': fore_mask, 'back_mask': back_mask, 'qry_imgs': [qry],
            'query_labels': query_labels, 'appr_query_labels': appr}


def perturb_bn_stats(state_dict, seed=1):
    """Give every BatchNorm non-trivial running statistics so that BN folding is actually
    exercised by eval-mode parity tests: mean ~ N(0, 0.1), var ~ U(0.5, 1.5)."""
    g = torch.Generator().manual_seed(seed)
    for k, v in state_dict.items():
        if k.endswith('running_mean'):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith('running_var'):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    return state_dict


def to_device(ep, device):
    mv = lambda t: t.to(device)
    return {'supp_imgs': [[mv(t) for t in way] for way in ep['supp_imgs']],
            'fore_mask': [[mv(t) for t in way] for way in ep['fore_mask']],
            'back_mask': [[mv(t) for t in way] for way in ep['back_mask']],
            'qry_imgs': [mv(t) for t in ep['qry_imgs']],
            'query_labels': mv(ep['query_labels']), 'appr_query_labels': mv(ep['appr_query_labels'])}


def fitted_state_dict(net, make_batch, steps=30, lr=1e-3):
    """Unsaturated, better-conditioned fixture for parity tests and bench.py's parity field: `steps` Adam steps of the
    B200 train step (rpnet_b200.train.TrainStep, weight_decay 0) on `make_batch(i)` episodes.  After them the BatchNorm running
    statistics describe the data (eval-mode logits are no longer pinned at the 20 * cos cap, where every implementation agrees)
    and the weights have left the random initialisation where the reference's own gradient is ill-conditioned.  Returns a CPU
    state_dict; lr = 0 only calibrates the running statistics."""
    import torch as _t
    from .train import TrainStep
    was_training = net.training
    net.train()
    ts = TrainStep(net, lr=lr, weight_decay=0.0)
    for i in range(steps):
        ts.step(make_batch(i))
    _t.cuda.synchronize()
    ts.eng.flat.unalias_grads()
    net.train(was_training)
    return {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
