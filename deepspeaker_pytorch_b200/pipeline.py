"""Host-to-host (or device-to-device) embedding extraction with copy/compute overlap and several forwards in flight.

The reference's ``test()`` loop (/root/reference/train_triplet.py:337-350) moves every batch to the GPU, runs the
model and pulls the distances back, all serialised on one stream.  ``EmbeddingPipeline`` keeps the same per-batch
call but

* runs the H2D copy of batch i+1 and the D2H copy of batch i-1 on their own streams while batch i is in the ResCNN
  kernels (PCIe is full duplex; 2.6 MB in, 128 KB out per 64 utterances), and
* alternates batches between ``lanes`` compute streams, each with its own engine handle and activation workspace:
  the late ResCNN stages of a 64-utterance batch have only 80-190 tiles for 148 SMs, so a second forward in flight
  fills the SMs the first one leaves idle.
"""
from __future__ import annotations

import torch

from . import engine as _engine


class EmbeddingPipeline:
    def __init__(self, model, lanes: int = 3, depth: int = 4):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("EmbeddingPipeline needs the model on a CUDA device")
        if model.training:
            raise RuntimeError("EmbeddingPipeline is an inference helper: call model.eval() first")
        self.model = model
        self.device = p.device
        self.depth = depth
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.lanes = [torch.cuda.Stream(self.device) for _ in range(lanes)]
        # lane 0 uses the module's own engine; further lanes get engines with their own activation workspace that
        # borrow lane 0's packed weights (one 21 MB weight image in L2 for all forwards in flight)
        e0 = model._get_engine(self.device)
        self.engines = [e0] + [_engine.Engine(model, self.device, model.operand_dtype, share_from=e0)
                               for _ in range(lanes - 1)]
        self._slots = {}
        self._i = 0

    def _slot(self, shape, k):
        key = (tuple(shape), k)
        if key not in self._slots:
            s = {"x": torch.empty(shape, device=self.device, dtype=torch.float32),
                 "h2d_done": torch.cuda.Event(), "free": torch.cuda.Event(), "emb_ready": torch.cuda.Event()}
            s["free"].record(torch.cuda.current_stream(self.device))
            self._slots[key] = s
        return self._slots[key]

    @torch.no_grad()
    def embed_device(self, x: torch.Tensor) -> torch.Tensor:
        """Queue one device-resident batch on the next compute lane; returns the (asynchronously produced) embeddings.

        Lifetimes: ``x`` may be dropped by the caller right after this call (it is recorded on the lane stream, so the
        caching allocator does not recycle its block before the forward has read it).  The returned ``emb`` is
        produced on the lane stream: order the consuming stream with ``wait_lanes()`` / ``synchronize()`` before
        reading it, and call ``emb.record_stream(consumer)`` if it is consumed on another stream and then dropped."""
        lane = self._i % len(self.lanes)
        self._i += 1
        st = self.lanes[lane]
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            emb = self.engines[lane].forward(x, False)
        x.record_stream(st)
        return emb

    def wait_lanes(self, stream=None):
        """Make ``stream`` (default: the current stream) wait for everything queued on the compute lanes."""
        stream = stream or torch.cuda.current_stream(self.device)
        for st in self.lanes:
            stream.wait_stream(st)

    @torch.no_grad()
    def embed(self, x_host: torch.Tensor, out_host: torch.Tensor) -> torch.cuda.Event:
        """Queue one batch: ``x_host`` (B,1,T,64) pinned fp32 -> ``out_host`` (B,E) pinned fp32.  Asynchronous:
        returns the event that marks ``out_host`` as complete (or call ``synchronize()``)."""
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise RuntimeError("EmbeddingPipeline.embed needs pinned host tensors (asynchronous copies)")
        lane = self._i % len(self.lanes)
        s = self._slot(x_host.shape, self._i % (self.depth * len(self.lanes)))
        self._i += 1
        st = self.lanes[lane]
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(s["free"])               # the previous forward that read this slot has finished
            s["x"].copy_(x_host, non_blocking=True)
            s["h2d_done"].record(self.h2d)
        with torch.cuda.stream(st):
            st.wait_event(s["h2d_done"])
            emb = self.engines[lane].forward(s["x"], False)
            s["free"].record(st)
            s["emb_ready"].record(st)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s["emb_ready"])
            out_host.copy_(emb, non_blocking=True)
            emb.record_stream(self.d2h)
            done = torch.cuda.Event()
            done.record(self.d2h)
        return done

    def synchronize(self):
        self.h2d.synchronize()
        for st in self.lanes:
            st.synchronize()
        self.d2h.synchronize()
