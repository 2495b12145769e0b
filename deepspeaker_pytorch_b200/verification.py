"""Verification loop of the reference's ``test()`` (/root/reference/train_triplet.py:330-366) on the B200 engine, plus
the derived equal error rate (the reference sweeps thresholds for best accuracy,
/root/reference/eval_metrics.py:5-50, and has no EER function).

Distances come from the CUDA kernels (eval forward + PairwiseDistance); the threshold sweep is a CPU metric over a few
thousand scalars and stays in numpy, as in the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from .model import PairwiseDistance


@torch.no_grad()
def verification_distances(model, data_a: torch.Tensor, data_p: torch.Tensor) -> torch.Tensor:
    """data_* (P, crops, T, 64): `crops` random crops per file (train_triplet.py:339-340 resizes them to
    (crops*P, 1, T, 64)); returns the per-pair distance averaged over crops ("length normalization", :348-350)."""
    if model.training:
        raise RuntimeError("verification runs in eval mode (train_triplet.py:332)")
    P, crops, T, F = data_a.shape
    a = data_a.reshape(P * crops, 1, T, F)
    p = data_p.reshape(P * crops, 1, T, F)
    out_a, out_p = model(a), model(p)
    d = PairwiseDistance(2).forward(out_a, out_p)
    return d.reshape(P, crops).mean(dim=1)


def sweep(distances, labels, thresholds=None):
    """(best-threshold accuracy, EER).  Accuracy follows eval_metrics.py:16-50 (thresholds 0..30 step 0.01, first
    argmax); EER is the FAR == FRR crossing of the same sweep."""
    d = np.asarray(distances, dtype=np.float64)
    same = np.asarray(labels).astype(bool)
    if thresholds is None:
        thresholds = np.arange(0, 30, 0.01)
    below = d[None, :] < thresholds[:, None]
    tp = (below & same[None, :]).sum(1)
    fp = (below & ~same[None, :]).sum(1)
    tn = (~below & ~same[None, :]).sum(1)
    fn = (~below & same[None, :]).sum(1)
    acc = (tp + tn) / d.size
    far = fp / max(1, int((~same).sum()))
    frr = fn / max(1, int(same.sum()))
    diff = far - frr
    i = int(np.argmax(diff >= 0))
    if i == 0:
        eer = float((far[0] + frr[0]) / 2)
    else:
        w = -diff[i - 1] / (diff[i] - diff[i - 1]) if diff[i] != diff[i - 1] else 0.0
        eer = float((far[i - 1] + w * (far[i] - far[i - 1]) + frr[i - 1] + w * (frr[i] - frr[i - 1])) / 2)
    return float(acc[int(np.argmax(acc))]), eer
