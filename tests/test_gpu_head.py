"""Classifier head + cross-entropy (model.py:167,220-223; train_triplet.py:281-285) and the fused Adagrad step
(train_triplet.py:369-383) on repo kernels vs plain PyTorch."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import deepspeaker_pytorch_b200 as dsk
from deepspeaker_pytorch_b200 import head

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K", [(6, 16, 512), (384, 1211, 512), (1, 1211, 512), (130, 67, 96)])
def test_linear_matches_fp64(cuda_dev, M, N, K):
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    gy = torch.randn(M, N, generator=g)
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    F.linear(xd, wd, bd).backward(gy.double())
    xc, wc, bc = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = head.LinearFn.apply(xc, wc, bc)
    y.backward(gy.cuda())
    ref = F.linear(x.double(), w.double(), b.double())
    assert torch.allclose(y.detach().cpu().double(), ref, rtol=1e-5, atol=1e-5)
    for got, want in ((xc.grad, xd.grad), (wc.grad, wd.grad), (bc.grad, bd.grad)):
        assert torch.allclose(got.cpu().double(), want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))


@pytest.mark.parametrize("M,C", [(6, 16), (384, 1211), (1, 5)])
def test_cross_entropy_matches_torch(cuda_dev, M, C):
    g = torch.Generator().manual_seed(C)
    logits = 3.0 * torch.randn(M, C, generator=g)
    labels = torch.randint(0, C, (M,), generator=g)
    ld = logits.double().requires_grad_(True)
    ref = F.cross_entropy(ld, labels)
    (2.5 * ref).backward()
    lc = logits.cuda().requires_grad_(True)
    loss = dsk.CrossEntropyLoss()(lc, labels.cuda())
    (2.5 * loss).backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 * max(1.0, abs(ref.item()))
    assert torch.allclose(lc.grad.cpu().double(), ld.grad, rtol=1e-5, atol=1e-8)


def test_cross_entropy_is_deterministic_and_rejects_cpu(cuda_dev):
    logits = torch.randn(64, 1211).cuda()
    labels = torch.randint(0, 1211, (64,)).cuda()
    a, b = dsk.CrossEntropyLoss()(logits, labels), dsk.CrossEntropyLoss()(logits, labels)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        dsk.CrossEntropyLoss()(logits.cpu(), labels.cpu())


def _mk_params(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 1, 5, 5), (64,), (64,), (128, 64, 3, 3), (7,), (513, 3), (1211, 512)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(dev)) for s in shapes]


@pytest.mark.parametrize("wd", [0.0, 1e-3])
def test_fused_adagrad_is_bit_identical_to_torch(cuda_dev, wd):
    """Same bits as torch.optim.Adagrad (foreach path) after 5 steps with lr_decay, the reference's hyper-parameters
    (train_triplet.py:70-77,378-382)."""
    ours, theirs = _mk_params(cuda_dev), _mk_params(cuda_dev)
    fo = dsk.FusedAdagrad(ours, lr=0.1, lr_decay=1e-4, weight_decay=wd)
    to = torch.optim.Adagrad(theirs, lr=0.1, lr_decay=1e-4, weight_decay=wd)
    g = torch.Generator().manual_seed(9)
    for it in range(5):
        fo.zero_grad()
        to.zero_grad()
        for p, q in zip(ours, theirs):
            gr = torch.randn(p.shape, generator=g).to(cuda_dev) * 10.0 ** (-(it % 3))
            p.grad.copy_(gr)
            q.grad = gr.clone()
        fo.step()
        to.step()
        for i, (p, q) in enumerate(zip(ours, theirs)):
            assert torch.equal(p.data, q.data), (it, i, (p.data - q.data).abs().max().item())
    sd_o, sd_t = fo.state_dict(), to.state_dict()
    for i in sd_t["state"]:
        assert torch.equal(sd_o["state"][i]["sum"], sd_t["state"][i]["sum"]), i
        assert float(sd_o["state"][i]["step"]) == float(sd_t["state"][i]["step"])
    # checkpoints interoperate (train_triplet.py:177-186,325-327): torch state -> fused, fused state -> torch
    fo2 = dsk.FusedAdagrad(_mk_params(cuda_dev), lr=0.1, lr_decay=1e-4, weight_decay=wd)
    fo2.load_state_dict(sd_t)
    assert fo2.step_count == 5 and torch.equal(fo2.state_dict()["state"][3]["sum"], sd_t["state"][3]["sum"])
    to.load_state_dict(sd_o)


def test_fused_adagrad_matches_cpu_golden(cuda_dev, golden_dir):
    """Pinned against torch.optim.Adagrad run on the CPU in the build container (tools/make_golden.py)."""
    import os

    g = np.load(os.path.join(golden_dir, "adagrad.npz"))
    p = torch.nn.Parameter(torch.from_numpy(g["p0"]).to(cuda_dev))
    opt = dsk.FusedAdagrad([p], lr=float(g["lr"]), lr_decay=float(g["lr_decay"]), weight_decay=0.0)
    for it in range(g["grads"].shape[0]):
        opt.zero_grad()
        p.grad.copy_(torch.from_numpy(g["grads"][it]).to(cuda_dev))
        opt.step()
    # torch's CUDA foreach Adagrad (which the kernel mirrors bit for bit, previous test) fuses sum += g*g into one FMA
    # where the CPU rounds the product first: a few ulp over the 6 steps, absolute where p_final cancels towards zero
    pmax = float(np.abs(g["p0"]).max())
    assert np.allclose(p.detach().cpu().numpy(), g["p_final"], rtol=2e-6, atol=2e-6 * pmax)
    assert np.allclose(opt.flat_sum[:p.numel()].cpu().numpy(), g["sum_final"].ravel(), rtol=1e-6, atol=0)


def test_engine_sees_fused_optimizer_updates(cuda_dev):
    """The fused step writes parameters through raw pointers: the engine must repack before the next forward."""
    from oracle import rescnn_oracle as O

    m = dsk.DeepSpeakerModel(512, 16).to(cuda_dev)
    m.load_state_dict(O.make_state_dict(0, 16))
    m.train()
    opt = dsk.FusedAdagrad(m.parameters(), lr=0.01, lr_decay=1e-4)
    xs = [O.make_input(4, 32, s, 3.0).cuda() for s in (1, 2, 3)]
    losses = []
    for _ in range(2):
        loss = dsk.TripletMarginLoss(5.0).forward(*[m(x) for x in xs])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[0] != losses[1]
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    m.eval()
    with torch.no_grad():
        e = m(xs[0]).cpu()
        ref = O.forward(sd, xs[0].cpu())
    assert ((e - ref).norm(dim=1) / ref.norm(dim=1)).max().item() < 2e-3
