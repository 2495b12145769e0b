"""Train-mode forward (batch-statistics BN) and backward of the B200 engine vs the oracle and the
reference's golden branch-A step (train_triplet.py:215-224)."""
import os

import numpy as np
import pytest
import torch

import deepspeaker_pytorch_b200 as dsk
from oracle import rescnn_oracle as O
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu

# tolerances: forward 1e-3 (north star); gradients rel-L2 per tensor (<= 1e-2 for 16-bit operands)
GRAD_TOL = {"fp16": 1e-2, "bf16": 6e-2}


def make_model(sd, dt, dev):
    m = dsk.DeepSpeakerModel(512, 16, operand_dtype=dt).to(dev)
    m.load_state_dict(sd)
    return m.train()


def run_step(m, xa, xp, xn, margin=0.1):
    out_a, out_p, out_n = m(xa), m(xp), m(xn)                      # train_triplet.py:215
    loss = dsk.TripletMarginLoss(margin).forward(out_a, out_p, out_n)   # :219
    m.zero_grad()
    loss.backward()                                                # :223
    return loss, out_a, out_p, out_n


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
def test_branch_a_step_matches_reference_golden(cuda_dev, golden_dir, dt):
    g = np.load(os.path.join(golden_dir, "train_step.npz"))
    B, T, s0, s1, s2, scale = g["cfg"]
    sd = O.make_state_dict(0, 16)
    m = make_model(sd, dt, cuda_dev)
    xa, xp, xn = (O.make_input(int(B), int(T), int(s), float(scale)).cuda() for s in (s0, s1, s2))
    loss, oa, op, on = run_step(m, xa, xp, xn)
    ftol = 1e-3 if dt == "fp16" else 8e-3
    for got, key in ((oa, "out_a"), (op, "out_p"), (on, "out_n")):
        ref = torch.from_numpy(g[key])
        assert ((got.detach().cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max().item() < ftol, key
    # the hinge is a difference of two O(10) distances: its error is bounded relative to the distance scale
    d_scale = (oa - op).detach().norm(dim=1).mean().item()
    assert abs(loss.item() - float(g["loss"])) <= (2e-3 if dt == "fp16" else 1.5e-2) * d_scale
    # running statistics after three train-mode forwards
    for k, v in m.state_dict().items():
        if "running" in k:
            assert np.allclose(v.cpu().numpy(), g["stat/" + k], rtol=5e-3, atol=5e-4), k
        if "num_batches_tracked" in k:
            assert int(v.item()) == 3
    # gradients: per-tensor norm and sampled entries
    checked = 0
    for k, p in m.named_parameters():
        if "gnorm/" + k not in g:
            assert p.grad is None or "classifier" in k
            continue
        ref_norm = float(g["gnorm/" + k])
        gr = p.grad.detach().cpu()
        assert abs(gr.double().norm().item() - ref_norm) <= GRAD_TOL[dt] * ref_norm + 1e-9, (k, gr.norm().item(), ref_norm)
        ix = torch.from_numpy(g["gidx/" + k])
        err = (gr.flatten()[ix] - torch.from_numpy(g["gval/" + k])).abs().max().item()
        assert err <= 6 * GRAD_TOL[dt] * ref_norm / np.sqrt(gr.numel()) + 1e-9, (k, err)
        checked += 1
    assert checked == 38


@pytest.mark.parametrize("B,T", [(6, 160), (5, 32), (16, 48)])
def test_train_step_matches_oracle_full_gradients(cuda_dev, B, T):
    sd = O.make_state_dict(1, 16)
    m = make_model(sd, "fp16", cuda_dev)
    xa, xp, xn = (O.make_input(B, T, s, 3.0) for s in (20, 21, 22))
    loss, oa, _op, _ = run_step(m, xa.cuda(), xp.cuda(), xn.cuda())
    stats = {}
    oloss, grads, ooa, _, _ = O.triplet_step_branch_a(sd, xa, xp, xn, 0.1, stats)
    assert ((oa.detach().cpu() - ooa).norm(dim=1) / ooa.norm(dim=1)).max().item() < 1e-3
    d_scale = (oa - _op).detach().norm(dim=1).mean().item()
    assert abs(loss.item() - oloss.item()) <= 2e-3 * d_scale
    worst = 0.0
    for k, p in m.named_parameters():
        if grads.get(k) is None:
            continue
        r = rel_l2(p.grad.detach().cpu(), grads[k])
        worst = max(worst, r)
        assert r < 2e-2, (k, r)
    for k, v in m.state_dict().items():
        if "running" in k:
            assert torch.allclose(v.cpu(), stats[k], rtol=5e-3, atol=5e-4), k
    print("worst grad rel-L2", worst)


def test_train_mode_without_grad_and_eval_after_train(cuda_dev):
    sd = O.make_state_dict(2, 16)
    m = make_model(sd, "fp16", cuda_dev)
    x = O.make_input(4, 32, 1, 2.0)
    with torch.no_grad():
        e = m(x.cuda())                       # train-mode BN, no graph
    st = {}
    ref = O.forward(sd, x, True, st)
    assert ((e.cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max().item() < 1e-3
    m.eval()                                  # eval fold must pick up the updated running stats
    sd2 = dict(sd)
    sd2.update(st)
    with torch.no_grad():
        e2 = m(x.cuda()).cpu()
        ref2 = O.forward(sd2, x)
    assert ((e2 - ref2).norm(dim=1) / ref2.norm(dim=1)).max().item() < 1e-3


def test_optimizer_step_is_picked_up(cuda_dev):
    """Adagrad step as in train_triplet.py:378-382, then a second forward must use the new weights."""
    sd = O.make_state_dict(3, 16)
    m = make_model(sd, "fp16", cuda_dev)
    opt = torch.optim.Adagrad(m.parameters(), lr=0.01, lr_decay=1e-4, weight_decay=0.0)
    xa, xp, xn = (O.make_input(4, 32, s, 3.0).cuda() for s in (1, 2, 3))
    l0, *_ = run_step(m, xa, xp, xn)
    opt.step()
    l1, *_ = run_step(m, xa, xp, xn)
    assert torch.isfinite(l0) and torch.isfinite(l1) and abs(l0.item() - l1.item()) > 0
    assert m.model.classifier.weight.grad is None
