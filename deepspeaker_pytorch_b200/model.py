"""Drop-in mirror of /root/reference/model.py for the hot path, backed by libdsk.so (sm_100a CUDA).

Same names, constructor arguments, submodule tree and ``state_dict`` keys as the reference:

* ``DeepSpeakerModel(embedding_size, num_classes, feature_dim=64)`` — model.py:153-223
* ``TripletMarginLoss(margin).forward(anchor, positive, negative)`` — model.py:19-33
* ``PairwiseDistance(p).forward(x1, x2)`` — model.py:8-18

The ``nn`` submodules below only *hold* parameters/buffers (so ``.cuda()``, ``parameters()``,
``state_dict()``, ``load_state_dict()``, optimizers and checkpoints work exactly as with the
reference, train_triplet.py:168-186,325,372-382); the arithmetic runs in the CUDA engine.  There is no
PyTorch/CPU fallback: non-CUDA inputs or a missing extension raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import engine as _engine
from . import head as _head


class ReLU(nn.Hardtanh):
    """Clipped ReLU, Hardtanh(0, 20) — model.py:36-44 (parameter-free; fused into the conv epilogues)."""

    def __init__(self, inplace=False):
        super().__init__(0, 20, inplace)

    def __repr__(self):
        return self.__class__.__name__ + " (" + ("inplace" if self.inplace else "") + ")"


def conv3x3(in_planes, out_planes, stride=1):
    """3x3 convolution with padding — model.py:47-50."""
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    """Parameter holder with the reference's attribute names — model.py:53-82."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3(inplanes, planes, stride)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = ReLU(inplace=True)
        self.conv2 = conv3x3(planes, planes)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class myResNet(nn.Module):
    """Parameter holder for the 4-stage ResCNN — model.py:85-147 (layers=[1,1,1,1], :162)."""

    def __init__(self, block=BasicBlock, layers=(1, 1, 1, 1), num_classes=1000):
        super().__init__()
        self.relu = ReLU(inplace=True)
        chans = (64, 128, 256, 512)
        cin = 1
        for s, ch in enumerate(chans):
            setattr(self, f"conv{s + 1}", nn.Conv2d(cin, ch, kernel_size=5, stride=2, padding=2, bias=False))
            setattr(self, f"bn{s + 1}", nn.BatchNorm2d(ch))
            setattr(self, f"layer{s + 1}", nn.Sequential(*[block(ch, ch) for _ in range(layers[s])]))
            cin = ch
        self.avgpool = nn.AdaptiveAvgPool2d((1, None))
        self.fc = nn.Linear(512 * block.expansion, num_classes)
        for m in self.modules():  # model.py:114-120
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()


class DeepSpeakerModel(nn.Module):
    """ResCNN speaker-embedding network — model.py:153-223 — running on the B200 engine.

    Extra keyword ``operand_dtype`` ("fp16" default, or "bf16") selects the 16-bit tensor-core operand
    format; accumulation, BatchNorm, pooling, fc and the L2-norm are fp32 either way.
    """

    def __init__(self, embedding_size, num_classes, feature_dim=64, operand_dtype="fp16"):
        super().__init__()
        if feature_dim != 64:
            # model.py:165-166: the feature_dim==40 branch does not match the 4-stage net
            raise ValueError("only feature_dim=64 is supported")
        if operand_dtype not in ("fp16", "bf16"):
            raise ValueError("operand_dtype must be 'fp16' or 'bf16'")
        self.embedding_size = embedding_size
        self.operand_dtype = operand_dtype
        self.model = myResNet(BasicBlock, [1, 1, 1, 1])
        self.model.fc = nn.Linear(512 * 4, self.embedding_size)          # model.py:163-164
        self.model.classifier = nn.Linear(self.embedding_size, num_classes)  # :167
        self._engine = None

    # -- engine plumbing -------------------------------------------------------------------------
    def _get_engine(self, device):
        if self._engine is None or self._engine.device != device:
            self._engine = _engine.Engine(self, device, self.operand_dtype)
        return self._engine

    def refresh_weights(self):
        """Tell the engine(s) that parameters / BatchNorm buffers were modified through ``.data`` (which PyTorch's
        version counter does not see): the packed 16-bit weights and the folded BN affine are rebuilt at the next
        forward.  In-place updates of the parameters themselves (optimizers, ``load_state_dict``) are detected
        automatically."""
        if self._engine is not None:
            self._engine.invalidate()

    def l2_norm(self, input):
        """model.py:172-183 (kept for API parity; the engine fuses it into the tail kernel)."""
        input_size = input.size()
        buffer = torch.pow(input, 2)
        normp = torch.sum(buffer, 1).add_(1e-10)
        norm = torch.sqrt(normp)
        _output = torch.div(input, norm.view(-1, 1).expand_as(input))
        return _output.view(input_size)

    def forward(self, x):
        """model.py:185-218: x (B,1,T,64) float CUDA tensor -> (B, embedding_size), L2 norm 10."""
        if not x.is_cuda:
            raise RuntimeError("DeepSpeakerModel (B200 engine) needs CUDA tensors; there is no CPU fallback")
        if x.dim() != 4 or x.size(1) != 1 or x.size(3) != 64:
            raise RuntimeError(f"expected input (B,1,T,64), got {tuple(x.shape)}")
        eng = self._get_engine(x.device)
        self.features = eng.forward(x, self.training)
        return self.features

    def forward_triplet(self, anchor, positive, negative):
        """The three forwards of a triplet step, ``model(data_a), model(data_p), model(data_n)``
        (/root/reference/train_triplet.py:215), issued together: same three results (bit-identical embeddings, batch
        statistics per call, running statistics updated in the order a, p, n), but in train mode the calls run on three
        streams and overlap - as do their backwards.  In eval mode it is simply three forwards.  ``self.features`` is
        left at the negative's embeddings, as after the reference's third call."""
        xs = (anchor, positive, negative)
        for x in xs:
            if not x.is_cuda:
                raise RuntimeError("DeepSpeakerModel (B200 engine) needs CUDA tensors; there is no CPU fallback")
            if x.dim() != 4 or x.size(1) != 1 or x.size(3) != 64:
                raise RuntimeError(f"expected input (B,1,T,64), got {tuple(x.shape)}")
        if not self.training:
            outs = [self.forward(x) for x in xs]
        else:
            from . import train as _train

            eng = self._get_engine(anchor.device)
            with torch.cuda.device(anchor.device):
                outs = _train.forward_train_many(eng, list(xs))
        self.features = outs[-1]
        return tuple(outs)

    def forward_classifier(self, x):
        """model.py:220-223: embeddings -> ``model.classifier`` logits (B, num_classes), on the repo's fp32 GEMM
        kernels (``dsk_linear_forward/backward``); ``model.classifier`` only holds the parameters."""
        features = self.forward(x)
        c = self.model.classifier
        return _head.LinearFn.apply(features, c.weight, c.bias)


class PairwiseDistance:
    """model.py:8-18 — ``.forward(x1, x2)`` is called directly (train_triplet.py:119,238,251,...)."""

    def __init__(self, p):
        if p != 2:
            raise ValueError("only the L2 distance (p=2) used by the reference hot path is implemented")
        self.norm = p

    def forward(self, x1, x2):
        assert x1.size() == x2.size()
        return _engine.PairwiseDistanceFn.apply(x1, x2)

    __call__ = forward


class TripletMarginLoss:
    """model.py:19-33 — ``TripletMarginLoss(margin).forward(a, p, n)`` (train_triplet.py:219,275)."""

    def __init__(self, margin):
        self.margin = margin
        self.pdist = PairwiseDistance(2)

    def forward(self, anchor, positive, negative):
        return _engine.TripletLossFn.apply(anchor, positive, negative, float(self.margin))

    __call__ = forward


def select_hard_triplets(d_p, d_n, margin):
    """Device-side restatement of train_triplet.py:251-262: returns (idx int64 (B,), count int32 (1,)) on the
    GPU; idx[:count] equals np.where((d_n - d_p < margin) == 1)[0].  No host synchronisation."""
    return _engine.margin_select(d_p, d_n, margin)


def allpairs_topk(E, labels, k, exact_cuda_cores=False):
    """BASELINE config 4: per-row k nearest different-label embeddings (idx int64 (N,k), dist fp32 (N,k)).
    Default: tcgen05 Gram GEMM + exact fp32 refinement (bit-identical to the all-fp32 path, `exact_cuda_cores=True`)."""
    return _engine.allpairs_topk(E, labels, k, exact_cuda_cores)
