#!/usr/bin/env python
"""torchrun --nproc-per-node 2 tools/check_overlap.py : overlapped (per-layer) gradient exchange == single all-reduce."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from audiolm_pytorch_b200.audiolm import CoarseTransformer  # noqa: E402
from audiolm_pytorch_b200.heads import cross_entropy  # noqa: E402
from audiolm_pytorch_b200.parallel import FlatGradBucket  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = CoarseTransformer(num_semantic_tokens=50, codebook_size=64, num_coarse_quantizers=3, dim=256, depth=3, heads=4,
                      flash_attn=True).to(dev).train()
bucket = FlatGradBucket(m.parameters()).attach(m)
torch.manual_seed(100 + rank)
sem, coarse = torch.randint(0, 50, (4, 60), device=dev), torch.randint(0, 64, (4, 130), device=dev)


def grads(overlap):
    ranges = [bucket.range_of(list(layer.parameters())) for layer in m.transformer.layers]
    m.transformer.grad_ready_hook = (lambda i: bucket.reduce_range_async(*ranges[i])) if overlap else None
    bucket.zero_()
    sl, cl = m(semantic_token_ids=sem, coarse_token_ids=coarse)
    (cross_entropy(sl, sem) + cross_entropy(cl, torch.cat((coarse, coarse[:, :1]), 1))).backward()
    if overlap:
        bucket.finish()
    else:
        bucket.all_reduce_mean()
    torch.cuda.synchronize()
    return bucket.flat.clone()


a, b = grads(False), grads(True)
err = (a - b).abs().max().item() / a.abs().max().item()
print(f"rank {rank}: overlapped vs single all-reduce max rel diff {err:.3e}", flush=True)
assert err < 1e-5, err   # split-K atomics make the local gradients run-to-run ~1e-7 different
dist.destroy_process_group()
