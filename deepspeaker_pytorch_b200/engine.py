"""Host-side engine: owns the libdsk handle of one DeepSpeakerModel, keeps the repacked weights in sync
with the nn.Parameters, and wraps the C-ABI calls in torch.autograd.Functions.

PyTorch is plumbing here (device memory, streams, autograd graph); all arithmetic is in libdsk.so.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L

_CONV_ORDER = None


def conv_bn_modules(model):
    """The 12 (conv, bn) pairs in C-ABI order: i = 3*stage + {0: convK/bnK, 1: layerK.0.conv1/bn1, 2: conv2/bn2}."""
    m = model.model
    out = []
    for s in range(1, 5):
        blk = getattr(m, f"layer{s}")[0]
        out += [(getattr(m, f"conv{s}"), getattr(m, f"bn{s}")), (blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
    return out


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError("libdsk needs contiguous float32 CUDA tensors")
    return t


class Engine:
    def __init__(self, module, device, operand_dtype, share_from=None):
        """``share_from``: another Engine of the same module whose packed weights this one borrows (inference lanes of
        ``EmbeddingPipeline``: one weight image in L2 for all forwards in flight)."""
        if device.type != "cuda":
            raise RuntimeError("the B200 engine runs on CUDA devices only")
        self.lib = L.load()
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        self.module_ref = module  # plain reference; Engine lifetime == module lifetime
        self.handle = ctypes.c_void_p()
        L.check(self.lib.dsk_create(ctypes.byref(self.handle), self.index,
                                    L.DSK_BF16 if operand_dtype == "bf16" else L.DSK_F16), "dsk_create")
        self._versions = None
        self._wstruct = None
        self.train_calls = 0  # train-mode forwards update BN running stats through raw pointers
        self.side_streams = []  # forward_train_many: one stream per forward in flight
        self._gscratch, self._gslot = [], 0
        self.bucket_accumulations = 0  # backwards that added their gradients straight into an optimizer bucket
        self.share_from = share_from
        if share_from is not None:
            L.check(self.lib.dsk_share_weights(self.handle, share_from.handle), "dsk_share_weights")

    def set_loss_scale(self, scale: float):
        """fp16 gradient scale used inside the backward (0 = automatic, see include/dsk.h)."""
        L.check(self.lib.dsk_set_loss_scale(self.handle, float(scale)), "dsk_set_loss_scale")

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.dsk_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    # -- parameters ---------------------------------------------------------------------------------
    def _param_versions(self, eval_mode):
        m = self.module_ref
        vs = []
        for conv, bn in conv_bn_modules(m):
            vs += [conv.weight._version, conv.weight.data_ptr(), bn.weight._version, bn.bias._version]
            if eval_mode:
                vs += [bn.running_mean._version, bn.running_var._version]
        fc = m.model.fc
        vs += [fc.weight._version, fc.bias._version, fc.weight.data_ptr(), int(eval_mode),
               self.train_calls if eval_mode else 0]
        return tuple(vs)

    def invalidate(self):
        """Force a repack / BN re-fold at the next forward.  Change detection relies on ``tensor._version`` (bumped by
        every in-place op on the parameter itself, which is what optimizers and ``load_state_dict`` do) and on
        ``data_ptr``; writes through ``param.data`` / ``buffer.data`` (``p.data.fill_()``, ``p.data.copy_()``, the
        reference's init idiom, model.py:114-120) do NOT bump the version: call this (or
        ``DeepSpeakerModel.refresh_weights()``) after such a write."""
        self._versions = None

    def sync_weights(self, eval_mode=True):
        """Repack/fold when any parameter (or, in eval, BN buffer) changed since the last call (see ``invalidate``)."""
        if self.share_from is not None:
            if not eval_mode:
                raise RuntimeError("an engine that borrows its weights is inference-only")
            self.share_from.sync_weights(True)
            return
        vs = self._param_versions(eval_mode)
        if vs == self._versions:
            return
        m = self.module_ref
        w = L.DskWeights()
        for i, (conv, bn) in enumerate(conv_bn_modules(m)):
            for t in (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var):
                if not t.is_cuda or t.device != self.device:
                    raise RuntimeError("model parameters must live on the engine's CUDA device (call model.cuda())")
            w.conv_w[i] = _f32c(conv.weight.data).data_ptr()
            w.bn_gamma[i] = _f32c(bn.weight.data).data_ptr()
            w.bn_beta[i] = _f32c(bn.bias.data).data_ptr()
            w.bn_running_mean[i] = _f32c(bn.running_mean).data_ptr()
            w.bn_running_var[i] = _f32c(bn.running_var).data_ptr()
        w.fc_w = _f32c(m.model.fc.weight.data).data_ptr()
        w.fc_b = _f32c(m.model.fc.bias.data).data_ptr()
        w.embedding_size = m.embedding_size
        self._wstruct = w
        if eval_mode:
            L.check(self.lib.dsk_load_weights(self.handle, ctypes.byref(w), L.cur_stream()), "dsk_load_weights")
        else:   # once per training step: only the operand images the training path reads, in one launch
            L.check(self.lib.dsk_load_weights_train(self.handle, ctypes.byref(w), L.cur_stream()), "dsk_load_weights_train")
        self._versions = vs

    def grad_scratch(self, params, slots: int = 6, with_flat: bool = False):
        """Per-parameter gradient tensors for one backward, as views of a flat scratch buffer that is allocated once and
        rotated over ``slots`` buffers (a backward allocated 38 tensors per call: ~0.2 ms of host time per context, and
        the training step is host-sensitive - ~450 kernel launches per 7 ms of GPU work).  Safe to recycle: a slot is
        reused ``slots`` backwards later, and autograd has consumed a backward's gradients (added them into ``p.grad`` on
        the stream the next step is ordered after) long before that."""
        key = tuple((p.data_ptr(), p.numel()) for p in params)
        if not self._gscratch or self._gscratch[0][0] != key:
            total = sum((p.numel() + 3) // 4 * 4 for p in params)
            self._gscratch = []
            for _ in range(slots):
                flat = torch.empty(total, device=self.device, dtype=torch.float32)
                views, off = [], 0
                for p in params:
                    views.append(flat[off:off + p.numel()].view_as(p))
                    off += (p.numel() + 3) // 4 * 4
                self._gscratch.append((key, flat, views))
            self._gslot = 0
        _, flat, views = self._gscratch[self._gslot]
        self._gslot = (self._gslot + 1) % len(self._gscratch)
        return (flat, views) if with_flat else views

    # -- forward ------------------------------------------------------------------------------------
    def forward(self, x, training):
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        B, _, T, _ = x.shape
        with torch.cuda.device(self.device):
            if training:
                from . import train as _train  # batch-statistics BN + autograd path

                return _train.forward_train(self, x)
            self.sync_weights(eval_mode=True)
            emb = torch.empty(B, self.module_ref.embedding_size, device=x.device, dtype=torch.float32)
            L.check(self.lib.dsk_rescnn_forward(self.handle, x.data_ptr(), B, T, emb.data_ptr(), L.DSK_EVAL,
                                                L.cur_stream()), "dsk_rescnn_forward")
        return emb


# ---------------------------------------------------------------------------------------------------
# distances / loss / selection
# ---------------------------------------------------------------------------------------------------
def _check2d(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("libdsk distance/loss ops need CUDA tensors; there is no CPU fallback")
        if t.dim() != 2:
            raise RuntimeError("expected (B, D) tensors")


class PairwiseDistanceFn(torch.autograd.Function):
    """PairwiseDistance(2).forward — /root/reference/model.py:13-18."""

    @staticmethod
    def forward(ctx, x1, x2):
        _check2d(x1, x2)
        x1c, x2c = x1.detach().float().contiguous(), x2.detach().float().contiguous()
        B, D = x1c.shape
        out = torch.empty(B, device=x1.device, dtype=torch.float32)
        with torch.cuda.device(x1.device):
            L.check(L.load().dsk_pairwise_distance(x1c.data_ptr(), x2c.data_ptr(), B, D, out.data_ptr(), L.cur_stream()),
                    "dsk_pairwise_distance")
        ctx.save_for_backward(x1c, x2c, out)
        return out

    @staticmethod
    def backward(ctx, go):
        x1, x2, dist = ctx.saved_tensors
        B, D = x1.shape
        go = go.float().contiguous()
        g1 = torch.empty_like(x1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(x2) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x1.device):
            L.check(L.load().dsk_pairwise_distance_bwd(x1.data_ptr(), x2.data_ptr(), dist.data_ptr(), go.data_ptr(), B, D,
                                                       L.ptr(g1), L.ptr(g2), L.cur_stream()), "dsk_pairwise_distance_bwd")
        return g1, g2


class TripletLossFn(torch.autograd.Function):
    """TripletMarginLoss(margin).forward — /root/reference/model.py:27-33 (loss is a device scalar)."""

    @staticmethod
    def forward(ctx, a, p, n, margin):
        _check2d(a, p, n)
        ac, pc, nc = (t.detach().float().contiguous() for t in (a, p, n))
        B, D = ac.shape
        loss = torch.empty(1, device=a.device, dtype=torch.float32)
        d_p = torch.empty(B, device=a.device, dtype=torch.float32)
        d_n = torch.empty(B, device=a.device, dtype=torch.float32)
        with torch.cuda.device(a.device):
            L.check(L.load().dsk_triplet_loss(ac.data_ptr(), pc.data_ptr(), nc.data_ptr(), B, D, margin, loss.data_ptr(),
                                              d_p.data_ptr(), d_n.data_ptr(), L.cur_stream()), "dsk_triplet_loss")
        ctx.save_for_backward(ac, pc, nc, d_p, d_n)
        ctx.margin = margin
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gl):
        a, p, n, d_p, d_n = ctx.saved_tensors
        B, D = a.shape
        gl = gl.float().reshape(1).contiguous()
        ga, gp, gn = torch.empty_like(a), torch.empty_like(p), torch.empty_like(n)
        with torch.cuda.device(a.device):
            L.check(L.load().dsk_triplet_loss_bwd(a.data_ptr(), p.data_ptr(), n.data_ptr(), d_p.data_ptr(), d_n.data_ptr(),
                                                  gl.data_ptr(), B, D, ctx.margin, ga.data_ptr(), gp.data_ptr(),
                                                  gn.data_ptr(), L.cur_stream()), "dsk_triplet_loss_bwd")
        return ga, gp, gn, None


def margin_select(d_p, d_n, margin):
    _ = [t for t in (d_p, d_n) if not t.is_cuda and (_ for _ in ()).throw(RuntimeError("CUDA tensors required"))]
    d_p, d_n = d_p.detach().float().contiguous(), d_n.detach().float().contiguous()
    B = d_p.numel()
    idx = torch.empty(B, device=d_p.device, dtype=torch.int64)
    count = torch.empty(1, device=d_p.device, dtype=torch.int32)
    with torch.cuda.device(d_p.device):
        L.check(L.load().dsk_margin_select(d_p.data_ptr(), d_n.data_ptr(), B, float(margin), idx.data_ptr(),
                                           count.data_ptr(), L.cur_stream()), "dsk_margin_select")
    return idx, count


def gather_rows(src, idx, count):
    """out[j] = src[idx[j]] for j < count (train_triplet.py:265-274), rows beyond count are left untouched."""
    src = src.detach().float().contiguous()
    rows = src.shape[0]
    row_elems = src[0].numel()
    out = torch.zeros_like(src)
    with torch.cuda.device(src.device):
        L.check(L.load().dsk_gather_rows(src.data_ptr(), idx.data_ptr(), count.data_ptr(), rows, row_elems,
                                         out.data_ptr(), L.cur_stream()), "dsk_gather_rows")
    return out


_AP_HANDLES = {}


def _allpairs_handle(device):
    """Module-level fp16 engine handle for the tensor-core all-pairs path (one per device)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _AP_HANDLES:
        h = ctypes.c_void_p()
        L.check(L.load().dsk_create(ctypes.byref(h), key, L.DSK_F16), "dsk_create")
        _AP_HANDLES[key] = h
    return _AP_HANDLES[key]


def allpairs_topk(E, labels, k, exact_cuda_cores: bool = False):
    """exact_cuda_cores=True forces the all-fp32 CUDA-core path; the default tensor-core path returns the same bits."""
    if not E.is_cuda:
        raise RuntimeError("CUDA tensors required")
    E = E.detach().float().contiguous()
    labels = labels.to(device=E.device, dtype=torch.int64).contiguous()
    N, D = E.shape
    idx = torch.empty(N, k, device=E.device, dtype=torch.int64)
    val = torch.empty(N, k, device=E.device, dtype=torch.float32)
    with torch.cuda.device(E.device):
        if exact_cuda_cores:
            L.check(L.load().dsk_allpairs_topk(E.data_ptr(), labels.data_ptr(), N, D, k, idx.data_ptr(), val.data_ptr(),
                                               L.cur_stream()), "dsk_allpairs_topk")
        else:
            L.check(L.load().dsk_allpairs_topk_tc(_allpairs_handle(E.device), E.data_ptr(), labels.data_ptr(), N, D, k,
                                                  idx.data_ptr(), val.data_ptr(), L.cur_stream()), "dsk_allpairs_topk_tc")
    return idx, val
