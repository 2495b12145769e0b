"""Verification metric: restated sweep vs the reference's own eval_metrics (golden), and held-out EER parity of the
B200 engine vs the oracle on synthetic speakers (north star: EER within 0.1 % absolute)."""
import os

import numpy as np
import pytest
import torch

from deepspeaker_pytorch_b200 import verification as V
from oracle import rescnn_oracle as O
from oracle import verification_oracle as VO


def synthetic_pairs(P, crops, T, seed):
    """Same-speaker pairs share a base spectrogram (+ per-crop noise), different-speaker pairs do not."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(P, 1, T, 64, generator=g) * 4.0
    other = torch.randn(P, 1, T, 64, generator=g) * 4.0
    same = torch.arange(P) % 2 == 0
    a = base + 2.5 * torch.randn(P, crops, T, 64, generator=g)
    p = torch.where(same.view(P, 1, 1, 1), base, other) + 2.5 * torch.randn(P, crops, T, 64, generator=g)
    return a, p, same.numpy()


def test_sweep_matches_reference_eval_metrics(golden_dir):
    g = np.load(os.path.join(golden_dir, "verification.npz"))
    acc, eer = V.sweep(g["distances"], g["labels"])
    assert abs(acc - float(g["ref_accuracy"])) < 1e-12          # reference eval_metrics.evaluate (best-threshold accuracy)
    tpr, fpr, oacc = VO.evaluate_accuracy(g["distances"], g["labels"])
    assert abs(oacc - float(g["ref_accuracy"])) < 1e-12 and abs(tpr - float(g["ref_tpr"])) < 1e-12
    assert abs(eer - VO.equal_error_rate(g["distances"], g["labels"])) < 1e-12
    assert 0.0 < eer < 0.5


@pytest.mark.gpu
def test_heldout_eer_within_a_tenth_of_a_percent(cuda_dev):
    import deepspeaker_pytorch_b200 as dsk

    sd = O.make_state_dict(0, 16)
    m = dsk.DeepSpeakerModel(512, 16).to(cuda_dev).eval()
    m.load_state_dict(sd)
    P, crops, T = 96, 8, 32                       # 8 crops per file, 32-frame crops: the reference's own test setup
    a, p, same = synthetic_pairs(P, crops, T, seed=11)
    d_gpu = V.verification_distances(m, a.cuda(), p.cuda()).cpu().numpy()
    with torch.no_grad():
        ea = O.forward(sd, a.reshape(P * crops, 1, T, 64))
        ep = O.forward(sd, p.reshape(P * crops, 1, T, 64))
        d_ref = VO.crop_mean_distances(O.pairwise_distance(ea, ep).numpy(), P, crops)
    assert np.allclose(d_gpu, d_ref, rtol=2e-3, atol=2e-3)
    acc_g, eer_g = V.sweep(d_gpu, same)
    _, _, acc_r = VO.evaluate_accuracy(d_ref, same)
    eer_r = VO.equal_error_rate(d_ref, same)
    assert 0.0 <= eer_r < 0.5
    assert abs(eer_g - eer_r) <= 1e-3, (eer_g, eer_r)          # north star: within 0.1 % absolute
    assert abs(acc_g - acc_r) <= 1.0 / P + 1e-9
