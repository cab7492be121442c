"""Parity measures between two logit tensors of the hot path (ours vs the reference / oracle), shared by the tests, bench.py
and tools/parity_report.py.  Pure tensor arithmetic on whatever device the inputs live on; no kernels."""
import torch


def fg_mask(logits):
    """The mask the recurrent loop / the eval driver derive from logits [B, P, H, W]: foreground probability > 0.5
    (net/rp_net.py:308-311; for P == 2 this is logit_1 > logit_0, test_rpnet.py:216-218 uses argmax)."""
    if logits.shape[1] == 2:
        return logits[:, 1] > logits[:, 0]
    return logits.argmax(1) > 0


def compare_logits(got, ref):
    """dict of
      rel_linf        : max |got - ref| / max |ref|                                  (north_star's 1e-3 gate)
      margin_rel_err  : max error of the decision margin (top-1 minus top-2 logit of the reference's ranking) relative to the
                        range of that margin — the gate that stays meaningful when all logits are close to the 20 * cos cap
      argmax_mismatch : fraction of pixels whose argmax differs
      dice_vs_ref     : Dice(our foreground mask, the reference's foreground mask)
    """
    got, ref = got.float(), ref.float()
    rel = ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
    order = ref.argsort(dim=1, descending=True)
    top = lambda x: torch.gather(x, 1, order[:, :2])
    mg, mr = top(got), top(ref)
    margin_g, margin_r = mg[:, 0] - mg[:, 1], mr[:, 0] - mr[:, 1]
    rng = (margin_r.max() - margin_r.min()).clamp_min(1e-30)
    a, b = fg_mask(got), fg_mask(ref)
    inter, card = (a & b).sum().item(), a.sum().item() + b.sum().item()
    return {'rel_linf': rel, 'margin_rel_err': ((margin_g - margin_r).abs().max() / rng).item(),
            'argmax_mismatch': (got.argmax(1) != ref.argmax(1)).float().mean().item(),
            'dice_vs_ref': (2.0 * inter / card) if card else 1.0, 'margin_median': margin_r.median().item()}
