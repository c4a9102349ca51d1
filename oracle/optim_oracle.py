"""TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's loss / metric / optimizer step.
Nothing under pathtracker_models_b200/ may import this module.

* `bce_scores` follows mainclean.py:156,189 (BCEWithLogitsLoss, mean) and utils/misc_functions.py:12-45
  (`metric_scores`, `acc_scores`): the prediction is `logit > 0.5` (:40), P is the batch size (:19), tp counts
  correct positives (:14), precision = tp / max(pred.sum(), 1e-6) (:20-24), recall = tp / P (:23),
  bal_acc = correct / P * 100 (:27,45), f1 = 2 tp / (P + tpfp) (:28).
* `adam_reference` is torch.optim.Adam with the reference's arguments (mainclean.py:157: lr only), the
  third-party algorithm the reference calls (torch/optim/adam.py); parity is pinned on torch itself."""
import numpy as np
import torch


def bce_scores(logits, target):
    z = np.asarray(logits, dtype=np.float64).reshape(-1)
    y = np.asarray(target, dtype=np.float64).reshape(-1)
    loss = np.mean(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z))))
    pred = (z > 0.5).astype(np.int64)
    tgt = y.astype(np.uint8).astype(np.int64)
    correct = (pred == tgt)
    P = float(len(z))
    tp = float(correct[tgt == 1].sum())
    tpfp = float(pred.sum()) or 1e-6
    dz = (1.0 / (1.0 + np.exp(-z)) - y) / P
    return dict(loss=loss, bal_acc=100.0 * correct.sum() / P, precision=tp / tpfp, recall=tp / P,
                f1=2 * tp / (P + tpfp), dlogits=dz)


def adam_reference(params, grads_per_step, lr):
    """params: list of tensors; grads_per_step: list (steps) of lists of tensors (or None). Returns updated copies."""
    ps = [torch.nn.Parameter(p.clone().double()) for p in params]
    opt = torch.optim.Adam(ps, lr=lr)
    for grads in grads_per_step:
        for p, g in zip(ps, grads):
            p.grad = None if g is None else g.clone().double()
        opt.step()
        opt.zero_grad()
    return [p.detach() for p in ps]
