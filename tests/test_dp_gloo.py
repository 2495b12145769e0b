"""World-size-2 gloo test (CPU) of the data-parallel host logic: one flat bucket, one allreduce per step, result
equal to the single-process gradient of the mean loss over the global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepspeaker_pytorch_b200 import DeepSpeakerModel
from deepspeaker_pytorch_b200.parallel import GradBucket, broadcast_parameters, path_parameters, shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy_loss(w, b, x):
    return ((x @ w + b) ** 2).mean()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                      # deliberately different initial parameters per rank
    w = torch.nn.Parameter(torch.randn(6, 3))
    b = torch.nn.Parameter(torch.randn(3))
    holder = torch.nn.ParameterList([w, b])
    broadcast_parameters(holder, src=0)
    bucket = GradBucket([w, b])
    g = torch.Generator().manual_seed(7)
    x_global = torch.randn(8, 6, generator=g)
    bucket.zero()
    _toy_loss(w, b, shard(x_global, rank, world)).backward()       # accumulates into the flat views
    work = bucket.allreduce_mean(async_op=True)
    work.wait()
    out[rank] = (w.detach().clone(), bucket.flat.clone(), bucket.collectives, w.grad.data_ptr() == bucket.flat.data_ptr())
    dist.destroy_process_group()


def test_single_allreduce_matches_global_batch_gradient():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    w0, flat0, ncoll0, view0 = out[0]
    w1, flat1, ncoll1, view1 = out[1]
    assert torch.equal(w0, w1)                          # broadcast from rank 0
    assert torch.allclose(flat0, flat1)                 # identical averaged gradients on both ranks
    assert ncoll0 == 1 and ncoll1 == 1                  # exactly one collective per step
    assert view0 and view1                              # p.grad aliases the bucket: no copy before the collective
    # single-process reference: gradient of the mean loss over the global batch
    w = w0.clone().requires_grad_(True)
    torch.manual_seed(100)
    _ = torch.randn(6, 3)
    b = torch.randn(3).requires_grad_(True)
    x_global = torch.randn(8, 6, generator=torch.Generator().manual_seed(7))
    _toy_loss(w, b, x_global).backward()
    assert torch.allclose(flat0, torch.cat([w.grad.flatten(), b.grad.flatten()]), atol=1e-6)


def _worker_weighted(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = torch.nn.Parameter(torch.ones(6, 3))
    b = torch.nn.Parameter(torch.zeros(3))
    bucket = GradBucket([w, b])
    x_global = torch.randn(8, 6, generator=torch.Generator().manual_seed(7))
    rows = (x_global[:5], x_global[5:])[rank]           # rank 0 selected 5 "hard triplets", rank 1 selected 3
    bucket.zero()
    _toy_loss(w, b, rows).backward()                    # local MEAN over the rank's own selection
    total = bucket.allreduce_weighted_mean(float(rows.shape[0]))
    out[rank] = (bucket.flat.clone(), float(total), bucket.collectives)
    dist.destroy_process_group()


def test_weighted_allreduce_is_the_mean_over_the_global_selection():
    """Branch B under data parallelism: k_r differs per rank; one collective carries gradients and weights."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_weighted, args=(world, port, out), nprocs=world, join=True)
    (flat0, tot0, n0), (flat1, tot1, n1) = out[0], out[1]
    assert tot0 == 8.0 and tot1 == 8.0 and n0 == 1 and n1 == 1
    assert torch.allclose(flat0, flat1)
    w = torch.ones(6, 3, requires_grad=True)
    b = torch.zeros(3, requires_grad=True)
    x_global = torch.randn(8, 6, generator=torch.Generator().manual_seed(7))
    _toy_loss(w, b, x_global).backward()                # mean over all 8 selected rows
    assert torch.allclose(flat0, torch.cat([w.grad.flatten(), b.grad.flatten()]), atol=1e-6)


def test_bucket_covers_the_triplet_path_parameters():
    m = DeepSpeakerModel(512, 1211)
    ps = path_parameters(m)
    assert len(ps) == 38
    b = GradBucket(ps)
    assert b.numel == 11624128                          # 46.5 MB fp32 allreduce payload
    assert m.model.classifier.weight.grad is None
    assert m.model.conv1.weight.grad.data_ptr() == b.flat.data_ptr()
    b.flat.fill_(1.0)
    assert float(m.model.fc.bias.grad.sum()) == 512.0
    b.zero()
    assert float(m.model.layer4[0].conv2.weight.grad.abs().sum()) == 0.0
    with pytest.raises(ValueError):
        shard(torch.zeros(7, 2), 0, 2)
