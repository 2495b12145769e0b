"""Host-to-host embedding extraction with copy/compute overlap.

The reference's ``test()`` loop (/root/reference/train_triplet.py:337-350) moves every batch to the GPU, runs the
model and pulls the distances back, all serialised on one stream.  ``EmbeddingPipeline`` keeps the same per-batch
call (``embed(x_host) -> embeddings on the host``) but runs the H2D copy of batch i+1 and the D2H copy of batch
i-1 on their own streams while batch i is in the ResCNN kernels (PCIe is full duplex; the copies are ~2.7 MB in,
128 KB out per 64 utterances).
"""
from __future__ import annotations

import torch


class EmbeddingPipeline:
    def __init__(self, model, depth: int = 2):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("EmbeddingPipeline needs the model on a CUDA device")
        self.model = model
        self.device = p.device
        self.depth = depth
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.compute = torch.cuda.Stream(self.device)
        self._slots = {}
        self._i = 0

    def _slot(self, shape, k):
        key = (tuple(shape), k)
        if key not in self._slots:
            self._slots[key] = {
                "x": torch.empty(shape, device=self.device, dtype=torch.float32),
                "h2d_done": torch.cuda.Event(), "free": torch.cuda.Event(), "emb_ready": torch.cuda.Event(),
            }
            self._slots[key]["free"].record(self.compute)
        return self._slots[key]

    @torch.no_grad()
    def embed(self, x_host: torch.Tensor, out_host: torch.Tensor) -> torch.cuda.Event:
        """Queue one batch: ``x_host`` (B,1,T,64) pinned fp32 -> ``out_host`` (B,E) pinned fp32.  Asynchronous:
        returns the event that marks ``out_host`` as complete (or call ``synchronize()``)."""
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise RuntimeError("EmbeddingPipeline.embed needs pinned host tensors (asynchronous copies)")
        s = self._slot(x_host.shape, self._i % self.depth)
        self._i += 1
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_event(s["free"])               # the previous forward that read this slot has finished
            s["x"].copy_(x_host, non_blocking=True)
            s["h2d_done"].record(self.h2d)
        with torch.cuda.stream(self.compute):
            self.compute.wait_event(s["h2d_done"])
            emb = self.model(s["x"])
            s["free"].record(self.compute)
            s["emb_ready"].record(self.compute)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s["emb_ready"])
            out_host.copy_(emb, non_blocking=True)
            emb.record_stream(self.d2h)
            done = torch.cuda.Event()
            done.record(self.d2h)
        return done

    def synchronize(self):
        self.h2d.synchronize()
        self.compute.synchronize()
        self.d2h.synchronize()
