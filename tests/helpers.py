"""Shared helpers for the test-suite (oracle-side data builders)."""
import numpy as np
import torch


def make_triplet_embeddings(B=64, D=512, seed=5):
    """Same recipe as tools/make_golden.py::make_triplet_embeddings (a, p close; n at varying distance)."""
    g = torch.Generator().manual_seed(seed)
    nrm = lambda t: 10.0 * t / t.norm(dim=1, keepdim=True)
    a = nrm(torch.randn(B, D, generator=g))
    p = nrm(a + 0.25 * torch.randn(B, D, generator=g))
    sig = torch.linspace(0.18, 0.34, B).view(B, 1)[torch.randperm(B, generator=g)]
    n = nrm(a + sig * torch.randn(B, D, generator=g))
    return a, p, n


def sample_idx(numel, n=16, seed=123):
    g = np.random.RandomState(seed)
    return g.randint(0, numel, size=n).astype(np.int64)


def rel_l2(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
