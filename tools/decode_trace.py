#!/usr/bin/env python
"""Phase timeline of the one-kernel decode step (CTA 0), C5 shapes.  Needs libalm_b200.so built with
ALM_EXTRA_NVCC_FLAGS=-DALM_DSTEP_TRACE:  ALM_EXTRA_NVCC_FLAGS=-DALM_DSTEP_TRACE python -m audiolm_pytorch_b200.build --force
    python tools/decode_trace.py [cache_len] [rows]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from audiolm_pytorch_b200 import _lib, decode  # noqa: E402
from audiolm_pytorch_b200.transformer import Transformer  # noqa: E402

n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 600
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = "cuda"
torch.manual_seed(0)
tr = Transformer(dim=1024, depth=6, heads=8, flash_attn=True).to(dev).eval()
dec = decode.StackDecoder(tr, b, 2048)
dec.load_cache(torch.randn(6, 2, b, n0, 64, device=dev))
x = torch.randn(b, 1024, device=dev)
for _ in range(5):
    dec.step(x)
torch.cuda.synchronize()
inner = tr.layers[0][2].branch.inner
off = int(_lib.load().alm_decode_stack_trace_offset(b, 1024, 8, inner))
tr_buf = dec._fused[2][off:off + 64 * 16 * 8].view(torch.int64).view(64, 16)[:6].cpu()
names = ["A hc", "A gemv q/kv", "A prefetch+barrier", "B attention", "B barrier", "C merge", "C gemv out", "C barrier",
         "D hc", "D gemv W1", "D barrier", "E geglu+ln", "E gemv W2", "E barrier"]
tot = torch.zeros(14)
for layer in range(6):
    d = (tr_buf[layer, 1:15] - tr_buf[layer, 0:14]).float()
    tot += d
    print(f"layer {layer}: " + " ".join(f"{v:6.0f}" for v in d.tolist()))
print("\nmean clk per layer (CTA 0; ~1.9 clk/ns):")
for n, v in zip(names, (tot / 6).tolist()):
    print(f"  {n:20s} {v:8.0f} clk  {v / 1.9e3:6.2f} us")
print(f"  layer total {tot.sum().item() / 6:8.0f} clk = {tot.sum().item() / 6 / 1.9e3:.1f} us; "
      f"kernel (first to last stamp) {(tr_buf[5, 14] - tr_buf[0, 0]).item() / 1.9e3:.1f} us")

sub = dec._fused[2][off + 32 * 16 * 8:off + 64 * 16 * 8].view(torch.int64).cpu()[:6 * 4 * 4].view(6, 4, 4)
dsub = (sub[:, :, 1:] - sub[:, :, :-1]).float().mean(0)
print("\ngemv_phase sub-stamps (mean clk over layers): wait-for-copy | columns | issue next copy + split-K sums")
for name, row in zip(("A q/kv", "C out", "D W1", "E W2"), dsub.tolist()):
    print(f"  {name:8s} " + " ".join(f"{v:8.0f}" for v in row))
