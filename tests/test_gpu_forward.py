"""DeepSpeakerModel.forward on the B200 engine vs the oracle and the reference's golden embeddings."""
import os

import numpy as np
import pytest
import torch

import deepspeaker_pytorch_b200 as dsk
from oracle import rescnn_oracle as O
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp16": 1e-3, "bf16": 6e-3}   # north star: 1e-3 relative fp32 (met with fp16 operands; bf16 documented at ~3e-3)


@pytest.fixture(scope="module")
def models(cuda_dev):
    sd = O.make_state_dict(0, 16)
    out = {}
    for dt in ("fp16", "bf16"):
        m = dsk.DeepSpeakerModel(512, 16, operand_dtype=dt).to(cuda_dev).eval()
        m.load_state_dict(sd)
        out[dt] = m
    return sd, out


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
def test_eval_forward_matches_reference_golden(models, golden_dir, dt):
    sd, ms = models
    g = np.load(os.path.join(golden_dir, "eval_forward.npz"))
    for name in ("a", "b", "c"):
        B, T, seed, scale = g[f"{name}_cfg"]
        x = O.make_input(int(B), int(T), int(seed), float(scale)).cuda()
        with torch.no_grad():
            e = ms[dt](x)
        ref = torch.from_numpy(g[f"{name}_emb"])
        rel = ((e.cpu() - ref).norm(dim=1) / ref.norm(dim=1)).max().item()
        assert rel < TOL[dt], (name, rel)
        assert torch.allclose(e.norm(dim=1).cpu(), torch.full((int(B),), 10.0), atol=1e-3)
        assert ms[dt].features is e          # side-effect attribute, model.py:210,213


@pytest.mark.parametrize("B,T", [(1, 160), (7, 160), (16, 48), (33, 32), (64, 160)])
def test_eval_forward_matches_oracle(models, B, T):
    sd, ms = models
    x = O.make_input(B, T, seed=100 + B, scale=5.0)
    with torch.no_grad():
        ref = O.forward(sd, x)
        e = ms["fp16"](x.cuda()).cpu()
    rel = ((e - ref).norm(dim=1) / ref.norm(dim=1))
    assert rel.max().item() < 1e-3
    # per-component gate: absolute error below 0.3 % of the RMS component (10/sqrt(512) = 0.44); the
    # floor-relative maximum of SURVEY §8d is a reported figure, not a gate (a near-zero component inflates it)
    assert (e - ref).abs().max().item() < 3e-3 * 10 / 512 ** 0.5


def test_full_size_properties(models):
    """BASELINE configs[1] size: determinism, unit-10 norms, batch-composition invariance (eval BN)."""
    sd, ms = models
    m = ms["fp16"]
    x = O.make_input(64, 160, seed=7, scale=8.0).cuda()
    with torch.no_grad():
        e1 = m(x).clone()
        e2 = m(x).clone()
        assert torch.equal(e1, e2)                                    # idempotent / deterministic
        assert torch.allclose(e1.norm(dim=1), torch.full((64,), 10.0, device=x.device), atol=1e-3)
        perm = torch.randperm(64, device=x.device)
        ep = m(x[perm].contiguous())
        assert torch.allclose(ep, e1[perm], atol=2e-5)                # each utterance is independent of its batch
        e_small = m(x[5:8].contiguous())
        assert torch.allclose(e_small, e1[5:8], atol=2e-5)
        assert torch.isfinite(m(torch.full_like(x, 1e4))).all()       # clip at 20 keeps everything finite


def test_weights_follow_parameter_updates(models):
    sd, ms = models
    m = dsk.DeepSpeakerModel(512, 16).cuda().eval()
    m.load_state_dict(sd)
    x = O.make_input(2, 32, seed=3).cuda()
    with torch.no_grad():
        e0 = m(x).clone()
        m.model.layer2[0].conv1.weight.mul_(1.5)          # in-place update bumps ._version -> repack
        sd2 = {k: v.clone() for k, v in sd.items()}
        sd2["model.layer2.0.conv1.weight"] = sd2["model.layer2.0.conv1.weight"] * 1.5
        e1 = m(x).cpu()
        ref = O.forward(sd2, x.cpu())
    assert rel_l2(e1, ref) < 1e-3 and rel_l2(e0.cpu(), ref) > 1e-3


def test_rejects_bad_input(models):
    sd, ms = models
    m = ms["fp16"]
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 1, 64, 160, device="cuda"))       # transposed layout
    with pytest.raises(RuntimeError, match="multiple of 16"):
        m(torch.zeros(2, 1, 100, 64, device="cuda"))
