"""Host-side mirror of the reference interface: names, state_dict keys, error behaviour (CPU only)."""
import json
import os
import re

import pytest
import torch

import deepspeaker_pytorch_b200 as dsk


def test_state_dict_keys_and_shapes_match_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    m = dsk.DeepSpeakerModel(512, 16)
    got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    assert got == ref


def test_reference_checkpoint_roundtrip():
    from oracle import rescnn_oracle as O

    sd = O.make_state_dict(0, 16)
    m = dsk.DeepSpeakerModel(512, 16)
    m.load_state_dict(sd)                       # strict: every reference key is consumed
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    assert sum(p.numel() for p in dsk.DeepSpeakerModel(512, 1211).parameters()) == 12245371


def test_init_follows_reference():
    torch.manual_seed(0)
    m = dsk.DeepSpeakerModel(512, 16)
    w = m.model.layer3[0].conv1.weight
    assert abs(w.std().item() - (2.0 / (9 * 256)) ** 0.5) < 2e-3      # model.py:114-117
    assert torch.all(m.model.bn2.weight == 1) and torch.all(m.model.bn2.bias == 0)   # :118-120
    assert m.model.fc.weight.shape == (512, 2048) and m.model.classifier.weight.shape == (16, 512)


def test_no_cpu_fallback():
    m = dsk.DeepSpeakerModel(512, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 1, 160, 64))
    with pytest.raises(RuntimeError):
        dsk.TripletMarginLoss(0.1).forward(torch.zeros(2, 4), torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(RuntimeError):
        dsk.PairwiseDistance(2).forward(torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(ValueError):
        dsk.DeepSpeakerModel(512, 16, feature_dim=40)
    with pytest.raises(ValueError):
        dsk.PairwiseDistance(1)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "deepspeaker_pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert not re.search(r"#include.*oracle|libdsk_oracle|c_oracle", src), f


def test_conv1_split_operand_arithmetic():
    """csrc/conv1_umma.cuh runs conv1 (model.py:93, Cin = 1) on the tensor cores with 16-bit operand halves:
    x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo with fp32 accumulation.  Emulated here in numpy: the dropped x_lo*w_lo
    term and the rounding of the low halves must leave the result at fp32-level accuracy (fp16 halves) and far below
    the 16-bit rounding of the stored activation (bf16 halves)."""
    import numpy as np

    rng = np.random.default_rng(0)
    x = (rng.standard_normal((4096, 25)) * 3.0).astype(np.float32)          # im2col rows of a normalised fbank patch
    w = (rng.standard_normal((25, 64)) * (2.0 / (25 * 64)) ** 0.5).astype(np.float32)   # model.py:114-117 init scale
    ref = x.astype(np.float64) @ w.astype(np.float64)

    def split(a, to16):
        hi = to16(a)
        return hi, to16(a - hi)

    def f16(a):
        return a.astype(np.float16).astype(np.float32)

    def bf16(a):                                                             # round-to-nearest-even on the top 16 bits
        u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
        u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        return u.astype(np.uint32).view(np.float32)

    scale = np.abs(ref).max()
    for to16, bound in ((f16, 2e-6), (bf16, 1e-4)):
        xh, xl = split(x, to16)
        wh, wl = split(w, to16)
        got = (xh + xl).astype(np.float64) @ wh.astype(np.float64) + xh.astype(np.float64) @ wl.astype(np.float64)
        assert np.abs(got - ref).max() / scale < bound
        plain = xh.astype(np.float64) @ wh.astype(np.float64)                # what a single 16-bit product would give
        assert np.abs(plain - ref).max() / scale > 20 * bound


def test_bucket_accumulation_is_only_taken_when_autograd_accumulates():
    """train._engine_accumulates_into decides, inside a backward, whether the gradients of the path's parameters may be added
    straight into the optimizer bucket: yes under loss.backward() (the AccumulateGrad nodes of the leaves will run), no under
    torch.autograd.grad (gradients are captured and must be returned)."""
    from deepspeaker_pytorch_b200 import train as T

    seen = []

    class Scale(torch.autograd.Function):
        @staticmethod
        def forward(ctx, tag, x, w, b):          # a non-tensor argument first, like TripletForwardFn(engine, k, ...)
            ctx.params = (w, b)
            return x * w + b

        @staticmethod
        def backward(ctx, g):
            seen.append(T._engine_accumulates_into(ctx, ctx.params))
            return None, None, g, g

    w = torch.ones(3, requires_grad=True)
    b = torch.zeros(3, requires_grad=True)
    x = torch.arange(3.0)
    Scale.apply("t", x, w, b).sum().backward()
    torch.autograd.grad(Scale.apply("t", x, w, b).sum(), [w, b])
    Scale.apply("t", x, w, b).sum().backward(inputs=[w, b])
    other = torch.ones(3, requires_grad=True)
    seen.append(T._engine_accumulates_into(Scale.apply("t", x, w, b).grad_fn, (w, other)))   # not a parameter of the node
    assert seen == [True, False, True, False]
