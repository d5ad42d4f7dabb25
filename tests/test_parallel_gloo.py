"""CPU, world_size 2 over gloo: the flat gradient bucket and its single all-reduce (DDP-mean semantics)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audiolm_pytorch_b200.parallel import FlatGradBucket

    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(8, 16), nn.Tanh(), nn.Linear(16, 4))
    bucket = FlatGradBucket(model.parameters())
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    x = torch.full((3, 8), float(rank + 1))
    bucket.zero_()
    model(x).sum().backward()
    local = bucket.flat.clone()
    bucket.all_reduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = torch.stack(gathered).mean(0)
    ok = torch.allclose(bucket.flat, expect, atol=1e-6)
    # grads are still views of the bucket after autograd accumulated into them
    ok = ok and all(torch.equal(p.grad.view(-1), bucket.flat[o:o + p.numel()])
                    for p, o in zip(bucket.params, _offsets(bucket.params)))
    # overlapped exchange: one slice early (as Transformer.grad_ready_hook would), the rest in finish()
    bucket.flat.copy_(local)
    lo, hi = bucket.range_of(list(model[2].parameters()))
    assert (lo, hi) == (8 * 16 + 16, bucket.numel)
    bucket.reduce_range_async(lo, hi)
    bucket.finish()
    ok = ok and torch.allclose(bucket.flat, expect, atol=1e-6)
    norm = bucket.clip_grad_norm_(0.5)
    ok = ok and bucket.flat.norm() <= 0.5 + 1e-4 and norm > 0
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def _offsets(params):
    off = 0
    for p in params:
        yield off
        off += p.numel()


def test_flat_bucket_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r, False) for r in range(world))


def test_flat_bucket_single_process():
    from audiolm_pytorch_b200.parallel import FlatGradBucket

    m = nn.Linear(4, 4)
    b = FlatGradBucket(m.parameters())
    m(torch.ones(2, 4)).sum().backward()
    assert b.flat.abs().sum() > 0
    assert b.all_reduce_mean() is None  # no process group: no-op
    b.zero_()
    assert m.weight.grad.abs().sum() == 0
