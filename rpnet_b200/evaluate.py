"""The evaluation driver loop of the reference (test_rpnet.py:151-258, "next" row N2) on the device.

Per query volume the reference copies every 2-slice batch's probabilities to the host and computes Dice in numpy; here the
volume goes through `volume.segment_volume` in large batches, the masks never leave the device, Dice / NCC are small device
reductions, and the printed lines keep the reference's format:

    {j} {pid} {supp_pid} affine ({ncc_warped}, {ncc_support}) {dsc_affine}, fewshot {dsc_fewshot} ref 0 {..},  ref 1 {..}, ...
    {class}, affine {mean}, voxel morph nan, nan, fewshot {mean} ref 0 {mean},  ...

Items follow the FewshotRegReader contract (dataset/few_shot_reader.py:599-647); `tools/eval_synthetic.py` builds them from
synthetic volumes with the batched affine registration of rpnet_b200.registration in place of get_registration_field."""
from collections import defaultdict

import numpy as np
import torch

from . import ops, volume


def _dice(pred, target):
    """dice_score_seperate(y_pred[None], y_true[None], num_class=1)[0] (utils/util.py:379-390) from device tensors: the RAW
    values of both arguments enter the sums (test_rpnet.py:226-227 passes the int labels and the registered soft label as they are);
    None when the target is empty."""
    return volume.dice_from_sums(volume.dice_sums(pred, target).cpu())


def format_volume_line(j, pid, supp_pid, d, d2, dsc_affine, dsc_fewshot, refs):
    """The line test_rpnet.py:231-241 prints per query volume (its print(..., end=' ') pieces joined): refs = [Dice of
    refinement iteration k]."""
    line = '%d %s %s affine (%s, %s) %s, fewshot %s ' % (j, pid, supp_pid, d, d2, dsc_affine, dsc_fewshot)
    for k, s in enumerate(refs):
        line += 'ref %d %s,  ' % (k, s)
    return line


def format_class_line(cls, affine, fewshot, refs):
    """The per-class summary line of test_rpnet.py:246-251 (dsc_list is never filled there: 'voxel morph nan, nan')."""
    line = '%s, affine %s, voxel morph nan, nan, fewshot %s ' % (cls, np.average(affine), np.average(fewshot))
    for ref, l in refs.items():
        line += 'ref %s %s,  ' % (ref, np.average(l))
    return line


@torch.no_grad()
def eval_volumes(net, items, eval_classes, batch_size=16, device='cuda', out=print):
    """items: iterable of dataset items (dict with support_images, support_labels, query_images, query_labels,
    appr_query_labels, warped_supp, class_id, pid, supp_pid).  Returns (dsc_affine_list, dsc_fewshot_list,
    dsc_refinement_list) like test_rpnet.eval (:258)."""
    net.eval()
    dsc_affine_list, dsc_fewshot_list = defaultdict(list), defaultdict(list)
    dsc_refinement_list = defaultdict(lambda: defaultdict(list))
    for j, item in enumerate(items):
        dev = torch.device(device)
        support_images = [[shot.float().to(dev) for shot in way] for way in item['support_images']]        # :166-171
        support_fg = [[shot.float().to(dev) for shot in way] for way in item['support_labels']]
        support_bg = [[1 - shot for shot in way] for way in support_fg]
        query_images = item['query_images'].float().to(dev)
        query_labels = item['query_labels'].long().to(dev)
        appr = item['appr_query_labels'].float().to(dev)
        warped_supp = item['warped_supp'].float().to(dev).reshape(query_images.shape)
        res = volume.segment_volume(net, support_images, support_fg, support_bg, query_images, appr, batch_size=batch_size)
        tgt = query_labels
        with torch.cuda.device(dev):
            dsc_affine = _dice(appr, tgt)                                                                    # :226
            dsc_fewshot = _dice(res['mask'], tgt)                                                            # :227
            d = ops.ncc(query_images.contiguous(), warped_supp.contiguous()).item()                          # :229 NCC(query, warped)
            d2 = ops.ncc(query_images.contiguous(), support_images[0][0].contiguous()).item()                # :230
            refs = [_dice(m, tgt) for m in res['masks_per_iter']]                                            # :237-241
        cls = eval_classes[item['class_id']]
        dsc_affine_list[cls].append(dsc_affine)
        dsc_fewshot_list[cls].append(dsc_fewshot)
        for k, s in enumerate(refs):
            dsc_refinement_list[cls][k].append(s)
        out(format_volume_line(j, item.get('pid'), item.get('supp_pid'), d, d2, dsc_affine, dsc_fewshot, refs))
    for k in eval_classes:                                                                                    # :246-251
        if k not in dsc_fewshot_list:
            continue
        out(format_class_line(k, dsc_affine_list[k], dsc_fewshot_list[k], dsc_refinement_list[k]))
    return dsc_affine_list, dsc_fewshot_list, dsc_refinement_list
