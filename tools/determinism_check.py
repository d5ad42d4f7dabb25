#!/usr/bin/env python
"""Run-to-run bitwise determinism of the forward kernels at C3 sizes (races show up as mismatches)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from audiolm_pytorch_b200 import ops  # noqa: E402

dev = "cuda"
bf16 = torch.bfloat16
torch.manual_seed(0)
M, d = 16 * 2048, 1024


def rnd(*s, dt=bf16, k=1.0):
    return (torch.randn(*s, device=dev) * k).to(dt)


def check(name, fn, n=6):
    ref = fn()
    ref = [t.clone() for t in (ref if isinstance(ref, (tuple, list)) else [ref]) if torch.is_tensor(t)]
    bad = 0
    for _ in range(n):
        out = fn()
        out = [t for t in (out if isinstance(out, (tuple, list)) else [out]) if torch.is_tensor(t)]
        for a, b in zip(ref, out):
            if not torch.equal(a, b):
                bad += 1
                diff = (a.float() - b.float()).abs()
                print(f"   {name}: mismatch, max abs {diff.max().item():.4g}, count {(diff > 0).sum().item()} of {a.numel()}")
                break
    print(f"{name}: {'DETERMINISTIC' if bad == 0 else f'{bad}/{n} runs differ'}")


a, w1 = rnd(M, d), rnd(5472, d, k=0.03)
check("gemm W1 fwd (tma store)", lambda: ops.gemm(a, w1))
w2, gn = rnd(d, 2736, k=0.03), rnd(M, 2736)
check("gemm W2 fwd", lambda: ops.gemm(gn, w2))
dy = rnd(M, d)
check("gemm dgrad (kmn)", lambda: ops.gemm(dy, w2, b_mn=True))
wh = rnd(1032, d, k=0.03)[:1025]
check("gemm fp32 out", lambda: ops.gemm(a[:8928], wh, out_dtype=torch.float32))
w2d = rnd(2736, d, k=0.03)
check("gemm N=2736 (tma store, tail)", lambda: ops.gemm(a, w2d))
out = torch.zeros(1024, 512, device=dev)
q, k, v = rnd(16, 2048, 512), rnd(16, 2048, 64), rnd(16, 2048, 64)
mask = torch.rand(16, 2048, device=dev) > 0.15
mask[:, 0] = True
check("attn fwd", lambda: ops.mqa_attn_fwd(q, k, v, heads=8))
check("attn fwd masked", lambda: ops.mqa_attn_fwd(q, k, v, heads=8, key_mask=mask))
o, lse = ops.mqa_attn_fwd(q, k, v, heads=8)
do = rnd(16, 2048, 512)
check("attn bwd", lambda: ops.mqa_attn_bwd(q, k, v, o, do, lse, heads=8))
q1 = rnd(16, 1, 512)
check("attn fwd decode n_q=1", lambda: ops.mqa_attn_fwd(q1, k, v, heads=8)[0])  # lse pad rows are unwritten
hc = dict(gamma=rnd(d, dt=torch.float32, k=0.1), dyn_alpha=rnd(d, 5, dt=torch.float32, k=0.05),
          dyn_beta=rnd(d, dt=torch.float32, k=0.05), static_alpha=rnd(4, 5, dt=torch.float32),
          static_beta=rnd(4, dt=torch.float32), alpha_scale=torch.tensor(0.3, device=dev),
          beta_scale=torch.tensor(0.3, device=dev))
lng = rnd(d, dt=torch.float32)
R, Y, bp = rnd(M, 4, d), rnd(M, d), rnd(M, 4, dt=torch.float32)
check("hc fwd", lambda: ops.hc_pre_fwd(hc, lng, R_in=R, Y=Y, beta_prev=bp, M=M, d=d))
R_out, bin_, xn, beta, aux = ops.hc_pre_fwd(hc, lng, R_in=R, Y=Y, beta_prev=bp, M=M, d=d)
dR, dxn, dbe, dbin = rnd(M, 4, d), rnd(M, d), rnd(M, 4, dt=torch.float32), rnd(M, d)


def hcb():
    grads = {k_: torch.zeros_like(v_) for k_, v_ in hc.items()}
    return ops.hc_pre_bwd(hc, lng, grads, torch.zeros_like(lng), aux, dR, dxn, dbe, dbin_extra=dbin, R_in=R, Y=Y,
                          beta_prev=bp, M=M, d=d)


check("hc bwd (data outputs)", hcb)
h = rnd(M, 5472)
g = rnd(2730, dt=torch.float32)
check("geglu fwd", lambda: ops.geglu_ln_fwd(h, g, inner=2730, inner_pad=2736))
gno, st = ops.geglu_ln_fwd(h, g, inner=2730, inner_pad=2736)
dgn = rnd(M, 2736)
check("geglu bwd dh", lambda: ops.geglu_ln_bwd(h, g, st, dgn, torch.zeros_like(g), inner=2730, inner_pad=2736))
