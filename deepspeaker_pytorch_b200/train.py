"""Train-mode forward/backward of DeepSpeakerModel on the B200 engine.

Mirrors what autograd does for the reference when the module is in train mode
(/root/reference/train_triplet.py:203,215-224): BatchNorm uses the batch statistics of each call, running
statistics are updated in place, and ``loss.backward()`` produces gradients for the 12 conv weights, the 12
BatchNorm affine pairs and fc (the classifier is outside this path and gets no gradient).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from .engine import conv_bn_modules


def _train_params(module):
    """The 38 parameters the path differentiates, in a fixed order: 12 x (conv.weight, bn.weight, bn.bias), fc.weight, fc.bias."""
    ps = []
    for conv, bn in conv_bn_modules(module):
        ps += [conv.weight, bn.weight, bn.bias]
    ps += [module.model.fc.weight, module.model.fc.bias]
    return ps


class _CtxGuard:
    """Returns the library-side context to the pool if the autograd graph is dropped without a backward."""

    def __init__(self, engine, tctx):
        self.engine, self.tctx, self.live = engine, tctx, True

    def consume(self):
        self.live = False

    def __del__(self):
        try:
            if self.live and self.engine.handle.value:
                self.engine.lib.dsk_train_ctx_release(self.engine.handle, self.tctx)
        except Exception:
            pass


class TrainForwardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, engine, *params):
        B, _, T, _ = x.shape
        emb = torch.empty(B, engine.module_ref.embedding_size, device=x.device, dtype=torch.float32)
        tctx = ctypes.c_void_p()
        L.check(engine.lib.dsk_rescnn_forward_train(engine.handle, x.data_ptr(), B, T, emb.data_ptr(), ctypes.byref(tctx),
                                                    L.cur_stream()), "dsk_rescnn_forward_train")
        ctx.engine = engine
        ctx.guard = _CtxGuard(engine, tctx)
        ctx.save_for_backward(x, *params)  # x must outlive the backward (conv1's weight gradient reads it)
        return emb

    @staticmethod
    def backward(ctx, grad_emb):
        engine = ctx.engine
        saved = ctx.saved_tensors
        params = saved[1:]
        grads = engine.grad_scratch(params)
        g = L.DskGrads()
        for i in range(L.NUM_CONV):
            g.conv_w[i] = grads[3 * i].data_ptr()
            g.bn_gamma[i] = grads[3 * i + 1].data_ptr()
            g.bn_beta[i] = grads[3 * i + 2].data_ptr()
        g.fc_w = grads[-2].data_ptr()
        g.fc_b = grads[-1].data_ptr()
        ge = grad_emb.float().contiguous()
        with torch.cuda.device(ge.device):
            L.check(engine.lib.dsk_rescnn_backward(engine.handle, ctx.guard.tctx, ge.data_ptr(), ctypes.byref(g),
                                                   L.cur_stream()), "dsk_rescnn_backward")
        ctx.guard.consume()
        return (None, None) + tuple(grads)


def _forward_one(engine, x, params, need_grad):
    """One train-mode forward on the CURRENT stream.  Returns (emb, library context or None if already released)."""
    if need_grad:
        emb = TrainForwardFn.apply(x, engine, *params)
        return emb, emb.grad_fn.guard.tctx
    B, _, T, _ = x.shape
    emb = torch.empty(B, engine.module_ref.embedding_size, device=x.device, dtype=torch.float32)
    tctx = ctypes.c_void_p()
    L.check(engine.lib.dsk_rescnn_forward_train(engine.handle, x.data_ptr(), B, T, emb.data_ptr(), ctypes.byref(tctx),
                                                L.cur_stream()), "dsk_rescnn_forward_train")
    return emb, tctx


def forward_train(engine, x):
    module = engine.module_ref
    engine.sync_weights(eval_mode=False)
    engine.train_calls += 1  # running statistics are about to change: invalidates the eval-mode BN fold
    params = _train_params(module)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    emb, tctx = _forward_one(engine, x, params, need_grad)
    if not need_grad:
        L.check(engine.lib.dsk_train_ctx_release(engine.handle, tctx), "dsk_train_ctx_release")
    # nn.BatchNorm2d bookkeeping in train mode
    torch._foreach_add_([bn.num_batches_tracked for _, bn in conv_bn_modules(module)], 1)
    return emb


class TripletForwardFn(torch.autograd.Function):
    """The K train-mode forwards of one step as ONE autograd node.  Forward: call k on side stream k.  Backward: the K
    backward chains on the same K streams, each into its own flat gradient buffer, then ONE ordered sum of the flat
    buffers on the caller's stream.  K separate nodes gave the same numbers, but autograd then summed the three
    gradients of each of the 38 parameters itself: 114 small ``add`` launches at the end of every step, serial, after the
    last backward kernel (profiles/r02_train_launches_ncu.md).  The sum here runs in the order autograd used (last
    call first: (g_n + g_p) + g_a), so the gradients are the same bits as those of K sequential calls."""

    @staticmethod
    def forward(ctx, engine, k, *args):
        xs, params = args[:k], args[k:]
        outs, tctxs = _launch_many(engine, xs)
        ctx.engine, ctx.k, ctx.params = engine, k, params
        ctx.guards = [_CtxGuard(engine, t) for t in tctxs]
        ctx.out_shape = outs[0].shape
        ctx.save_for_backward(*xs, *params)  # the inputs must outlive the backward (conv1's weight gradient reads them)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_embs):
        engine, k, params = ctx.engine, ctx.k, ctx.params
        _ = ctx.saved_tensors                      # raises if a parameter was modified in place since the forward
        dev = engine.device
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            flats, views = [], None
            for ge, st, guard in zip(grad_embs, engine.side_streams, ctx.guards):
                ge = (torch.zeros(ctx.out_shape, device=dev, dtype=torch.float32) if ge is None
                      else ge.float().contiguous())
                flat, views = engine.grad_scratch(params, with_flat=True)
                g = L.DskGrads()
                for i in range(L.NUM_CONV):
                    g.conv_w[i] = views[3 * i].data_ptr()
                    g.bn_gamma[i] = views[3 * i + 1].data_ptr()
                    g.bn_beta[i] = views[3 * i + 2].data_ptr()
                g.fc_w = views[-2].data_ptr()
                g.fc_b = views[-1].data_ptr()
                st.wait_stream(cur)                # grad_emb was produced on the caller's stream
                with torch.cuda.stream(st):
                    L.check(engine.lib.dsk_rescnn_backward(engine.handle, guard.tctx, ge.data_ptr(), ctypes.byref(g),
                                                           L.cur_stream()), "dsk_rescnn_backward")
                guard.consume()
                flats.append(flat)
            for st in engine.side_streams[:k]:
                cur.wait_stream(st)                # also keeps every grad_emb alive long enough: it is freed on `cur`
            acc = flats[-1]                        # `views` are the views of this buffer
            for f in reversed(flats[:-1]):
                acc.add_(f)
            # parameters whose .grad is a view of an optimizer / data-parallel bucket (FusedAdagrad, GradBucket):
            # accumulate into the bucket with one multi-tensor add instead of 38 AccumulateGrad nodes
            if (all(p.grad is not None and p.grad is getattr(p, "_dsk_bucket_grad", None) for p in params)
                    and _engine_accumulates_into(ctx, params)):
                torch._foreach_add_([p.grad for p in params], list(views))
                engine.bucket_accumulations += 1
                return (None, None) + (None,) * k + (None,) * len(params)
        return (None, None) + (None,) * k + tuple(views)


def _engine_accumulates_into(node, params):
    """True when the running backward will accumulate this node's parameter gradients into ``p.grad`` (``loss.backward()``),
    False when it captures them instead (``torch.autograd.grad``: the AccumulateGrad nodes are not executed and the query
    raises) - then the gradients must be returned to autograd, not added into the bucket."""
    try:
        acc = {id(fn.variable): fn for fn, _ in node.next_functions if fn is not None and hasattr(fn, "variable")}
        return all(id(p) in acc and torch._C._will_engine_execute_node(acc[id(p)]) for p in params)
    except Exception:
        return False


def _launch_many(engine, xs):
    """Forward k of ``xs`` on side stream k, running-statistics updates deferred; joins the side streams.  Returns
    (embeddings, library contexts)."""
    dev = engine.device
    cur = torch.cuda.current_stream(dev)
    while len(engine.side_streams) < len(xs):
        engine.side_streams.append(torch.cuda.Stream(dev))
    L.check(engine.lib.dsk_set_defer_running_stats(engine.handle, 1), "dsk_set_defer_running_stats")
    outs, ctxs = [], []
    try:
        for x, st in zip(xs, engine.side_streams):
            st.wait_stream(cur)                      # inputs (and the parameters) were produced on the caller's stream
            with torch.cuda.stream(st):
                emb, tctx = _forward_one(engine, x, None, False)
            x.record_stream(st)
            outs.append(emb)
            ctxs.append(tctx)
    finally:
        L.check(engine.lib.dsk_set_defer_running_stats(engine.handle, 0), "dsk_set_defer_running_stats")
    for st, emb in zip(engine.side_streams, outs):
        cur.wait_stream(st)
        emb.record_stream(cur)
    return outs, ctxs


def forward_train_many(engine, xs):
    """Several independent train-mode forwards of one step (the anchor / positive / negative calls of
    train_triplet.py:215) IN FLIGHT TOGETHER: forward k runs on its own side stream, so the HBM-bound BatchNorm passes
    of one call overlap the tensor-core convs of another, and so do the three backwards (``TripletForwardFn``).
    Results are those of the sequential calls, bit for bit: batch statistics are per call anyway, the running-statistics
    updates are recorded per call and committed afterwards in call order (``dsk_train_ctx_commit_stats``), and the
    gradients are summed in autograd's order.  All side streams are joined before returning."""
    module = engine.module_ref
    engine.sync_weights(eval_mode=False)
    engine.train_calls += 1
    params = _train_params(module)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    xs = [x.contiguous().float() for x in xs]
    if need_grad:
        outs = list(TripletForwardFn.apply(engine, len(xs), *xs, *params))
        ctxs = [g.tctx for g in outs[0].grad_fn.guards]
    else:
        outs, ctxs = _launch_many(engine, xs)
    for tctx in ctxs:                                # momentum updates in call order, on the caller's stream
        L.check(engine.lib.dsk_train_ctx_commit_stats(engine.handle, tctx, L.cur_stream()), "dsk_train_ctx_commit_stats")
        if not need_grad:
            L.check(engine.lib.dsk_train_ctx_release(engine.handle, tctx), "dsk_train_ctx_release")
    torch._foreach_add_([bn.num_batches_tracked for _, bn in conv_bn_modules(module)], len(xs))
    return outs
