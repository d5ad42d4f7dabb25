#!/usr/bin/env python
"""Exactly N C3 training steps (bench.py's step) with no other GPU work - the target of ncu launch lists."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from audiolm_pytorch_b200.audiolm import CoarseTransformer  # noqa: E402
from audiolm_pytorch_b200.heads import cross_entropy  # noqa: E402
from audiolm_pytorch_b200.parallel import FlatGradBucket  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
flash = "--no-flash" not in sys.argv
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = CoarseTransformer(**{**bench.CFG, "flash_attn": flash}).to(dev).train()
bucket = FlatGradBucket(model.parameters()).attach(model)
sem, coarse = (t.to(dev) for t in bench.synth_ids(bench.BATCH, 0))
eos = torch.full((bench.BATCH, 1), bench.CFG["codebook_size"], device=dev)
for _ in range(n):
    bucket.zero_()
    model.transformer.invalidate_weight_cache()
    sl, cl = model(semantic_token_ids=sem, coarse_token_ids=coarse)
    loss = (cross_entropy(sl, sem) * sl.shape[1] + cross_entropy(cl, torch.cat((coarse, eos), 1)) * cl.shape[1]) / (
        sl.shape[1] + cl.shape[1])
    loss.backward()
torch.cuda.synchronize()
print("loss", loss.item())
