"""The per-batch body of the reference's ``train()`` loop (/root/reference/train_triplet.py:208-299) on the B200 engine.

``train_step`` restates both branches of the loop with the drop-in classes and the repo's kernels:

* branch A (``epoch > min_softmax_epoch``, :217-224): triplet loss over all triplets, backward, optimizer step;
* branch B (:251-291): margin mask -> hard-triplet indices, triplet loss on the detached selected embeddings,
  second train-mode forward of the selected inputs through ``forward_classifier``, cross-entropy over
  ``cat[cls_a, cls_p, cls_n]`` vs ``cat[label_p, label_p, label_n]``, ``loss = CE + loss_ratio * triplet``.

What differs from the reference is only where things run: the mask, the ascending index list and every gather stay on
the device (``dsk_margin_select`` / ``dsk_gather_rows``); the reference makes six device->host->device round trips
through numpy (:253-274).  ONE host synchronisation remains - reading the number of selected triplets k, which sizes the
second forward (and implements ``if len(hard_triplets[0]) == 0: continue``, :263-264).

The second forward is NOT replaced by re-using the first one's activations: in train mode its BatchNorm layers normalise
with the statistics of the k SELECTED utterances (and update the running statistics three more times), so its logits are
a different function of the parameters than anything the first forward computed.

Data parallelism: pass ``bucket`` (``parallel.GradBucket``) or a ``FusedAdagrad`` optimizer; branch A
averages gradients over ranks, branch B weights each rank's mean gradient by its own k (``k_r / sum k``) through the
same single allreduce.
"""
from __future__ import annotations

import torch

from . import engine as _engine
from .head import CrossEntropyLoss
from .model import PairwiseDistance, TripletMarginLoss, select_hard_triplets
from .optim import FusedAdagrad

_l2 = PairwiseDistance(2)   # train_triplet.py:119


def _reduce_and_step(optimizer, bucket, weight):
    """backward has filled the gradients: the step's one collective, then optimizer.step() (:224, :291)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        weight = None      # one process: the local mean IS the global mean (and the update stays bit-identical to torch's)
    if isinstance(optimizer, FusedAdagrad):
        if weight is not None:
            optimizer.flat_grad.mul_(weight)     # gradients of k_r * (local mean loss); step() divides by sum_r k_r
        optimizer.allreduce(weight=weight)
    elif bucket is not None:
        if weight is None:
            bucket.allreduce_mean()
        else:
            bucket.allreduce_weighted_mean(weight)
    optimizer.step()


def train_step(model, optimizer, data_a, data_p, data_n, label_p, label_n, *, margin, epoch, min_softmax_epoch=2,
               loss_ratio=2.0, bucket=None):
    """One batch of train_triplet.py:208-299.  Returns a dict of device scalars (``loss``, ``triplet``, ``ce``),
    ``selected`` (python int: triplets that entered the loss) and the bookkeeping distances ``d_p`` / ``d_n``
    (:238-245, :251-252); ``None`` when branch B selects nothing (the reference's ``continue``)."""
    if not model.training:
        raise RuntimeError("train_step needs model.train() (train_triplet.py:203)")
    out_a, out_p, out_n = model.forward_triplet(data_a, data_p, data_n)                 # :215
    crit = TripletMarginLoss(margin)
    if epoch > min_softmax_epoch:
        triplet = crit.forward(out_a, out_p, out_n)                                     # :218
        loss = triplet                                                                  # :219
        optimizer.zero_grad()                                                           # :221
        loss.backward()                                                                 # :222
        _reduce_and_step(optimizer, bucket, None)                                       # :223
        with torch.no_grad():
            d_n = _l2.forward(out_a.detach(), out_n.detach())                           # :237
            d_p = _l2.forward(out_a.detach(), out_p.detach())                           # :242
        return {"loss": loss.detach(), "triplet": triplet.detach(), "ce": None, "selected": int(out_a.shape[0]),
                "d_p": d_p, "d_n": d_n}
    # ---- choose the hard negatives (:250-274) -----------------------------------------------------------------------
    with torch.no_grad():
        d_p = _l2.forward(out_a.detach(), out_p.detach())                               # :251
        d_n = _l2.forward(out_a.detach(), out_n.detach())                               # :252
        idx, cnt = select_hard_triplets(d_p, d_n, margin)                               # :253-262 (device)
    k = int(cnt.item())                                                                 # the branch's one host sync
    if k == 0:
        return None                                                                     # :263-264
    with torch.no_grad():
        g = lambda t: _engine.gather_rows(t, idx, cnt)[:k]                              # :265-274 (device gathers)
        sel_a, sel_p, sel_n = g(out_a), g(out_p), g(out_n)
        xa, xp, xn = g(data_a), g(data_p), g(data_n)
        hard = idx[:k]
        true = torch.cat([label_p.to(hard.device)[hard], label_p.to(hard.device)[hard], label_n.to(hard.device)[hard]])  # :283
    triplet = crit.forward(sel_a, sel_p, sel_n)                                         # :275 (constant w.r.t. the parameters)
    cls_a = model.forward_classifier(xa)                                                # :277
    cls_p = model.forward_classifier(xp)                                                # :278
    cls_n = model.forward_classifier(xn)                                                # :279
    ce = CrossEntropyLoss()(torch.cat([cls_a, cls_p, cls_n]), true)                     # :281-285
    loss = ce + triplet * loss_ratio                                                    # :287
    optimizer.zero_grad()                                                               # :289
    loss.backward()                                                                     # :290
    _reduce_and_step(optimizer, bucket, cnt.to(torch.float32).reshape(()))              # :291 (k_r-weighted under DP)
    return {"loss": loss.detach(), "triplet": triplet.detach(), "ce": ce.detach(), "selected": k, "d_p": d_p, "d_n": d_n,
            "hard": hard}
