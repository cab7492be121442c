"""Slice-sharded inference over a 3-D volume: the eval loop of the reference driver (test_rpnet.py:151-258) restated on
device.  The reference walks a query volume in batches of 2 slices (test_rpnet.py:164,189), copies every batch's
probabilities to the host (:219) and computes Dice in numpy (utils/util.py:379-390); here the slices of the volume are
independent units: every rank takes a contiguous range of slices, runs `RP_Net.forward` in batches of
`batch_size`, keeps the thresholded masks on the device and reduces the Dice sums with one small all-reduce.

Dataset item contract (dataset/few_shot_reader.py:629-647): support_images [[S x 1 x H x W]], support_labels
[[S x H x W]] (slice-matched to the query by the k-block lookup), query_images S x 1 x H x W, query_labels S x H x W,
appr_query_labels S x H x W."""
import torch

from .train import shard_range


def dice_sums(pred, target):
    """(2 * sum(t * p), sum(t) + sum(p), sum(t)) on the device — the sums of dice_score_seperate (utils/util.py:379-390)."""
    p, t = pred.float(), target.float()
    return torch.stack([2.0 * (p * t).sum(), p.sum() + t.sum(), t.sum()])


def dice_from_sums(sums, target_sum=None):
    """utils/util.py:379-390: None when the TARGET is empty (whatever the prediction), else 2*sum(t*p) / (sum(t)+sum(p))
    rounded to 4 decimals.  `sums` = dice_sums(...) (the target sum is its third entry unless given explicitly)."""
    num, den = float(sums[0]), float(sums[1])
    tsum = float(sums[2]) if target_sum is None else float(target_sum)
    return None if tsum == 0 else round(num / den, 4)


@torch.no_grad()
def segment_volume(net, support_images, support_fg, support_bg, query_images, appr_query_labels, batch_size=16, rank=0, world=1,
                   keep_logits=False):
    """Run the eval forward over this rank's slices of one query volume.

    support_images / support_fg / support_bg: way x shot lists of S x 1 x H x W / S x H x W device tensors (slice-matched);
    query_images S x 1 x H x W; appr_query_labels S x H x W.  Returns a dict:
      'range'  (lo, hi) slice range owned by this rank,
      'mask'   uint8 [hi-lo, H, W]: softmax(output)[:, 1] > 0.5  (test_rpnet.py:219,224; for Wa > 1: any foreground class),
      'masks_per_iter' list of T uint8 tensors from out['refinement'][k] (test_rpnet.py:237-241),
      'logits' (optional) fp32 [hi-lo, 1+Wa, H, W]."""
    if net.training:
        raise RuntimeError('segment_volume is the eval loop: call net.eval() first')
    S = query_images.shape[0]
    lo, hi = shard_range(S, rank, world)
    T = net.num_iter
    masks, per_iter, logits = [], [[] for _ in range(T)], []
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        nb = b1 - b0
        # the last partial batch is padded to `batch_size` by repeating its last slice (slices are independent in the eval
        # forward): one batch shape per volume = one set of workspace buffers and one captured CUDA graph, whatever the slice count
        pad = batch_size - nb if (hi - lo) > batch_size else 0

        def cut(t):
            c = t[b0:b1]
            if pad:
                c = torch.cat([c, c[-1:].expand(pad, *c.shape[1:])], dim=0)
            return c.contiguous()
        out = net([[cut(t) for t in way] for way in support_images], [[cut(t) for t in way] for way in support_fg],
                  [[cut(t) for t in way] for way in support_bg], [cut(query_images)], appr_query_labels=cut(appr_query_labels))

        def to_mask(lg):                       # softmax(dim=1)[:, 1] > 0.5  <=>  fg logit(s) win: no softmax pass needed
            if lg.shape[1] == 2:
                return (lg[:, 1] > lg[:, 0]).to(torch.uint8)
            p = lg.softmax(dim=1)
            return (p[:, 1:].sum(1) > 0.5).to(torch.uint8)
        masks.append(to_mask(out['output'][:nb]))
        for k in range(T):
            per_iter[k].append(to_mask(out['refinement'][k][:nb]))
        if keep_logits:
            logits.append(out['output'][:nb].clone())
    res = {'range': (lo, hi), 'mask': torch.cat(masks) if masks else None,
           'masks_per_iter': [torch.cat(m) if m else None for m in per_iter]}
    if keep_logits:
        res['logits'] = torch.cat(logits) if logits else None
    return res


def volume_dice(result, query_labels, world=1, group=None):
    """Dice of the whole volume from per-rank partial sums (one 2 x (1+T) all-reduce when world > 1).
    Returns (dice_final, [dice per refinement iteration])."""
    lo, hi = result['range']
    tgt = (query_labels[lo:hi] > 0)
    sums = [dice_sums(result['mask'], tgt)] + [dice_sums(m, tgt) for m in result['masks_per_iter']]
    sums = torch.stack(sums)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    sums = sums.cpu()
    return dice_from_sums(sums[0]), [dice_from_sums(s) for s in sums[1:]]


def make_synthetic_volume(slices=96, size=256, ways=1, shots=1, seed=0):
    """cfg5-shaped synthetic item: an ellipsoid organ through `slices` slices; support = a different 'patient'
    (seed) with one support slice per k-block of 8 query slices, replicated over the block
    (dataset/few_shot_reader.py:516-545).  Returns the dataset-item dict on the CPU."""
    from .synthetic import _slice
    qs, ql = [], []
    for z in range(slices):
        img, lab = _slice(seed + 5000 + z // 4, size, ways)            # neighbouring slices share anatomy
        qs.append(img)
        ql.append(lab)
    query = torch.stack(qs)[:, None]
    labels = torch.stack(ql)
    kblock = 8
    supp_imgs, supp_fg, supp_bg = [], [], []
    for w in range(ways):
        si, sf, sb = [], [], []
        for s in range(shots):
            imgs, fgs = [], []
            for z in range(slices):
                img, lab = _slice(seed + 9000 + 31 * s + (z // kblock), size, ways)
                imgs.append(img)
                fgs.append((lab == w + 1).float())
            fg = torch.stack(fgs)
            si.append(torch.stack(imgs)[:, None]); sf.append(fg); sb.append(1 - fg)
        supp_imgs.append(si); supp_fg.append(sf); supp_bg.append(sb)
    appr = torch.roll((labels > 0).float(), shifts=(6, -5), dims=(1, 2))
    return {'support_images': supp_imgs, 'support_fg': supp_fg, 'support_bg': supp_bg, 'query_images': query,
            'query_labels': labels, 'appr_query_labels': appr}
