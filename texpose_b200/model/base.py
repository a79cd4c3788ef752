"""Drop-in for the one hot-path method of model/base.py: `Model.summarize_loss` (reference :145-157).

The reference weighs every term with 10**loss_weight, sums them into `loss.all`, and asserts finiteness with
`torch.isinf` / `torch.isnan` on each scalar -- two host syncs per term and step.  Here the weighted sum comes out of the
fused loss pass (Graph.compute_loss leaves it in `var.loss_all_fused`), and the finiteness asserts run only under
`opt.b200.nan_guard`.  The reference's own method also works on the dict Graph.compute_loss returns.
"""
from __future__ import annotations

import torch


def summarize_loss(opt, var, loss):
    assert "all" not in loss
    b = opt.get("b200") if hasattr(opt, "get") else None
    guard = bool(b.get("nan_guard", False)) if b else False
    for key in loss:
        assert key in opt.loss_weight
        assert loss[key].shape == ()
        if guard and opt.loss_weight[key] is not None:
            assert not torch.isinf(loss[key]), "loss {} is Inf".format(key)
            assert not torch.isnan(loss[key]), "loss {} is NaN".format(key)
    fused = var.get("loss_all_fused") if hasattr(var, "get") else None
    if fused is not None:
        loss_all = fused
    else:       # terms that did not come from the fused pass: the reference's sum
        loss_all = 0.
        for key in loss:
            if opt.loss_weight[key] is not None:
                loss_all = loss_all + 10 ** float(opt.loss_weight[key]) * loss[key]
    loss.update(all=loss_all)
    return loss
