#!/usr/bin/env python
"""Run a few launches of ONE hot kernel at C3 shapes (for `ncu --set full -k regex:<name> -c 1`).

    python tools/ncu_one.py gemm_ffn | gemm_wgrad | attn_fwd | attn_bwd | hc_fwd | hc_bwd | geglu | rvq | conv
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from audiolm_pytorch_b200 import ops  # noqa: E402

dev = "cuda"
bf16 = torch.bfloat16
which = sys.argv[1] if len(sys.argv) > 1 else "gemm_ffn"
M, d = 16 * 2048, 1024
torch.manual_seed(0)


def rnd(*s, dt=bf16, k=1.0):
    return (torch.randn(*s, device=dev) * k).to(dt)


for _ in range(3):
    if which == "gemm_ffn":
        ops.gemm(rnd(M, d), rnd(5472, d, k=0.03))
    elif which == "gemm_wgrad":
        out = torch.zeros(2730, d, device=dev)
        dh = rnd(M, 5472)
        ops.gemm(dh[:, :2730], rnd(M, d), a_mn=True, b_mn=True, out=out, acc_mode=2, split_k=5)
    elif which in ("attn_fwd", "attn_bwd"):
        q, k, v = rnd(16, 2048, 512), rnd(16, 2048, 64), rnd(16, 2048, 64)
        o, lse = ops.mqa_attn_fwd(q, k, v, heads=8)
        if which == "attn_bwd":
            ops.mqa_attn_bwd(q, k, v, o, rnd(16, 2048, 512), lse, heads=8)
    elif which in ("hc_fwd", "hc_bwd"):
        hc = dict(gamma=rnd(d, dt=torch.float32, k=0.1), dyn_alpha=rnd(d, 5, dt=torch.float32, k=0.05),
                  dyn_beta=rnd(d, dt=torch.float32, k=0.05), static_alpha=rnd(4, 5, dt=torch.float32),
                  static_beta=rnd(4, dt=torch.float32), alpha_scale=torch.tensor(0.3, device=dev),
                  beta_scale=torch.tensor(0.3, device=dev))
        lng = rnd(d, dt=torch.float32)
        R, Y, bp = rnd(M, 4, d), rnd(M, d), rnd(M, 4, dt=torch.float32)
        R_out, bin_, xn, beta, aux = ops.hc_pre_fwd(hc, lng, R_in=R, Y=Y, beta_prev=bp, M=M, d=d)
        if which == "hc_bwd":
            grads = {k_: torch.zeros_like(v_) for k_, v_ in hc.items()}
            ops.hc_pre_bwd(hc, lng, grads, torch.zeros_like(lng), aux, rnd(M, 4, d), rnd(M, d),
                           rnd(M, 4, dt=torch.float32), dbin_extra=rnd(M, d), R_in=R, Y=Y, beta_prev=bp, M=M, d=d)
    elif which == "geglu":
        h = rnd(M, 5472)
        g = rnd(2730, dt=torch.float32)
        gn, st = ops.geglu_ln_fwd(h, g, inner=2730, inner_pad=2736)
        ops.geglu_ln_bwd(h, g, st, rnd(M, 2736), torch.zeros_like(g), inner=2730, inner_pad=2736)
    elif which == "rvq":
        ops.rvq_encode(rnd(4800, 512, dt=torch.float32), rnd(8, 1024, 512, dt=torch.float32))
    elif which == "conv":
        x = rnd(32, 64, 24000, dt=torch.float32)
        w = rnd(64, 64, 7, dt=torch.float32, k=0.05)
        ops.causal_conv1d(x, w, rnd(64, dt=torch.float32), dilation=9, elu=True,
                          weight_packed=w.permute(1, 2, 0).contiguous())
torch.cuda.synchronize()
print("done", which)
