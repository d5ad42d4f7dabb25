#!/usr/bin/env python
"""Per-graph replay time of the decode engine (which of the captured steps is slow?)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from audiolm_pytorch_b200.audiolm import (CoarseTransformer, CoarseTransformerWrapper, FineTransformer,  # noqa: E402
                                          FineTransformerWrapper)

dev = "cuda"
torch.manual_seed(0)
kw = dict(dim=1024, depth=6, heads=8, flash_attn=True)


class _Codec:
    rq_groups = 1
    num_quantizers = 8


fine = FineTransformerWrapper(transformer=FineTransformer(num_coarse_quantizers=3, num_fine_quantizers=5,
                                                          codebook_size=1024, **kw).to(dev), codec=_Codec())
coarse = CoarseTransformerWrapper(transformer=CoarseTransformer(num_semantic_tokens=500, codebook_size=1024,
                                                                num_coarse_quantizers=3, **kw).to(dev),
                                  codec=_Codec(), unique_consecutive=False)
fine.generate(coarse_token_ids=torch.randint(0, 1024, (1, 60, 3), device=dev))
coarse.generate(semantic_token_ids=torch.randint(0, 500, (1, 500), device=dev), max_time_steps=60)
for name, w in (("fine", fine), ("coarse", coarse)):
    dec = w._engine[1]
    print(name, "graphs", sorted(dec._graphs), "cache len", int(dec.stack.len.item()), "max_len", dec.stack.max_len)
    for key in sorted(dec._graphs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.inference_mode():
            for _ in range(100):
                dec.advance(key)
        torch.cuda.synchronize()
        print(f"  key {key}: {(time.perf_counter() - t0) * 10:.3f} ms per replay, graph nodes unknown")
# full generate timings
for name, fn, n in (("coarse", lambda: coarse.generate(semantic_token_ids=torch.randint(0, 500, (1, 500), device=dev), max_time_steps=100), 300),
                    ("fine", lambda: fine.generate(coarse_token_ids=torch.randint(0, 1024, (1, 60, 3), device=dev)), 300)):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    print(name, "generate", (time.perf_counter() - t0) * 1e3 / n, "ms/token")
