"""EmbeddingPipeline (host-to-host serving path: H2D / compute lanes / D2H streams, CUDA-graph forwards, lanes that
borrow lane 0's packed weights through dsk_share_weights) must return exactly what DeepSpeakerModel.forward returns."""
import pytest
import torch

import deepspeaker_pytorch_b200 as dsk
from oracle import rescnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(cuda_dev):
    m = dsk.DeepSpeakerModel(512, 16).to(cuda_dev).eval()
    m.load_state_dict(O.make_state_dict(0, 16))
    return m


@pytest.mark.parametrize("lanes", [1, 3])
def test_pipeline_matches_direct_forward(model, lanes):
    pipe = dsk.EmbeddingPipeline(model, lanes=lanes)
    n = 4 * lanes + 3                                  # every lane runs plain, capturing and re-targeted graph launches
    xs = [O.make_input(6, 48, seed=200 + i, scale=4.0) for i in range(n)]
    xh = [x.pin_memory() for x in xs]
    oh = [torch.empty(6, 512).pin_memory() for _ in range(n)]
    for i in range(n):
        pipe.embed(xh[i], oh[i])
    pipe.synchronize()
    with torch.no_grad():
        for i in range(n):
            assert torch.equal(oh[i], model(xs[i].cuda()).cpu()), i
    # device-resident entry point
    with torch.no_grad():
        outs = [pipe.embed_device(x.cuda()) for x in xs[:lanes + 1]]
        pipe.synchronize()
        torch.cuda.synchronize()
        for i, e in enumerate(outs):
            assert torch.equal(e, model(xs[i].cuda())), i


def test_borrowing_lanes_follow_weight_updates(model):
    pipe = dsk.EmbeddingPipeline(model, lanes=3)
    x = O.make_input(4, 32, seed=300, scale=3.0)
    xh, outs = x.pin_memory(), [torch.empty(4, 512).pin_memory() for _ in range(6)]
    for k in range(3):
        pipe.embed(xh, outs[k])
    pipe.synchronize()
    with torch.no_grad():
        model.model.conv3.weight.mul_(1.25)            # parameter version bump -> lane 0 repacks, lanes 1-2 re-adopt
    try:
        for k in range(3, 6):
            pipe.embed(xh, outs[k])
        pipe.synchronize()
        with torch.no_grad():
            ref = model(x.cuda()).cpu()
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
        assert not torch.equal(outs[0], outs[3])
        for k in range(3, 6):
            assert torch.equal(outs[k], ref), k
    finally:
        with torch.no_grad():
            model.model.conv3.weight.div_(1.25)


def test_borrower_rejects_own_weights_and_training(model):
    import ctypes

    from deepspeaker_pytorch_b200 import _lib as L
    from deepspeaker_pytorch_b200 import engine as E

    e0 = model._get_engine(next(model.parameters()).device)
    e1 = E.Engine(model, e0.device, model.operand_dtype, share_from=e0)
    with pytest.raises(RuntimeError):
        e1.sync_weights(eval_mode=False)
    e0.sync_weights(True)
    rc = e1.lib.dsk_load_weights(e1.handle, ctypes.byref(e0._wstruct), L.cur_stream())
    assert rc != 0 and b"borrows" in e1.lib.dsk_last_error()


def test_shape_change_mid_stream_keeps_earlier_batches_intact(model):
    """A ragged last batch (or any new (B, T)) re-zeroes the lane's padded workspace: that memset must be ordered after
    the lane's forwards still in flight and before the new shape's kernels (it used to run on the NULL
    stream, unordered against the non-blocking lane streams).  Many full batches are queued ahead of the short one so
    that the host runs well ahead of the GPU, then shapes alternate."""
    pipe = dsk.EmbeddingPipeline(model, lanes=2, depth=4)
    full = [O.make_input(48, 160, seed=400 + i, scale=4.0) for i in range(10)]
    short = [O.make_input(5, 160, seed=500 + i, scale=4.0) for i in range(3)]
    seq = full[:8] + [short[0]] + full[8:] + [short[1], short[2], full[0]]
    xh = [x.pin_memory() for x in seq]
    oh = [torch.empty(x.shape[0], 512).pin_memory() for x in seq]
    for i in range(len(seq)):
        pipe.embed(xh[i], oh[i])
    pipe.synchronize()
    with torch.no_grad():
        for i, x in enumerate(seq):
            assert torch.equal(oh[i], model(x.cuda()).cpu()), i
    # device entry point: the caller drops its input right away (record_stream keeps the block alive for the lane)
    with torch.no_grad():
        outs = [pipe.embed_device(x.cuda()) for x in seq]
        pipe.synchronize()
        torch.cuda.synchronize()
        for i, x in enumerate(seq):
            assert torch.equal(outs[i], model(x.cuda())), i


def test_tickets_and_frozen_mode(model):
    """embed() returns a ticket; wait(ticket) blocks the host until that batch's pinned output is complete (later
    batches may still be in flight).  check_every=0 freezes the packed weights until refresh()."""
    pipe = dsk.EmbeddingPipeline(model, lanes=2, depth=2, check_every=0)
    xs = [O.make_input(5, 32, seed=400 + i, scale=3.0) for i in range(9)]     # 9 batches > lanes * depth slots
    xh = [x.pin_memory() for x in xs]
    oh = [torch.zeros(5, 512).pin_memory() for _ in xs]
    tickets = [pipe.embed(a, b) for a, b in zip(xh, oh)]
    assert tickets == list(range(tickets[0], tickets[0] + 9))
    pipe.wait(tickets[2])
    with torch.no_grad():
        assert torch.equal(oh[2], model(xs[2].cuda()).cpu())
    pipe.wait(tickets[0])                  # a ticket whose slot was reused since: already complete
    pipe.synchronize()
    with torch.no_grad():
        for i in (0, 8):
            assert torch.equal(oh[i], model(xs[i].cuda()).cpu()), i
        model.model.conv3.weight.mul_(1.25)        # a PACKED parameter (the fc bias is read in place, not packed)
        try:
            stale = torch.zeros(5, 512).pin_memory()
            pipe.wait(pipe.embed(xh[0], stale))
            assert torch.equal(stale, oh[0])                       # frozen: still the old weights
            pipe.refresh()
            fresh = torch.zeros(5, 512).pin_memory()
            pipe.wait(pipe.embed(xh[0], fresh))
            assert torch.equal(fresh, model(xs[0].cuda()).cpu()) and not torch.equal(fresh, stale)
        finally:
            model.model.conv3.weight.div_(1.25)
    with pytest.raises(RuntimeError):
        pipe.embed(xs[0], oh[0])                                   # not pinned
    with pytest.raises(RuntimeError):
        pipe.wait(10 ** 9)
