#!/usr/bin/env python
"""All-reduce of the flat gradient bucket (65.6 M fp32 = 262 MB) timed alone: SUM + div vs AVG, fp32 vs bf16, chunked.
torchrun --nproc-per-node N tools/allreduce_probe.py"""
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 65_628_969
flat = torch.randn(n, device=dev)
half = flat.to(torch.bfloat16)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def report(name, ms, nbytes):
    if rank == 0:
        print(f"{name:44s} {ms:7.3f} ms  busbw {2 * (world - 1) / world * nbytes / (ms * 1e-3) / 1e9:7.1f} GB/s", flush=True)


report("fp32 SUM + div_", timed(lambda: (dist.all_reduce(flat), flat.div_(world))), n * 4)
report("fp32 AVG", timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG)), n * 4)
report("bf16 AVG", timed(lambda: dist.all_reduce(half, op=dist.ReduceOp.AVG)), n * 2)
for chunks in (2, 4, 8):
    parts = flat.chunk(chunks)
    report(f"fp32 AVG in {chunks} chunks", timed(lambda: [dist.all_reduce(p, op=dist.ReduceOp.AVG) for p in parts]), n * 4)
dist.destroy_process_group()
