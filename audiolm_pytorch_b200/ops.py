"""Python faces of the C-ABI kernels (raw ops; autograd wiring lives in `functional.py`).

Every function here launches hand-written sm_100a kernels from libalm_b200.so on the current CUDA
stream.  No function has a CPU or stock-PyTorch implementation.
"""
from __future__ import annotations

import torch

from . import _lib

bf16 = torch.bfloat16
f32 = torch.float32


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.AlmError("audiolm_pytorch_b200 ops need CUDA tensors (no CPU fallback)")


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=bf16, alpha=1.0, bias=None,
         acc_mode=0, split_k=1):
    """out[b,m,n] (op)= alpha * sum_k A(m,k) B(n,k) (+bias[n]).   bf16 operands, fp32 accumulate.

    a: [(batch,) M, K] (a_mn=False) or [(batch,) K, M] (a_mn=True); row stride must be a multiple of 8.
    b: [(batch,) N, K] (b_mn=False) or [(batch,) K, N] (b_mn=True).
    Reference counterpart: nn.Linear / einsum calls listed in include/alm_b200.h.
    """
    _check_cuda(a, b, out, bias)
    assert a.dtype == bf16 and b.dtype == bf16, "gemm operands must be bf16"
    batched = a.dim() == 3
    if not batched:
        a3, b3 = a.unsqueeze(0), b.unsqueeze(0)
    else:
        a3, b3 = a, b
    assert a3.stride(-1) == 1 and b3.stride(-1) == 1
    nb = a3.shape[0]
    assert b3.shape[0] == nb
    if a_mn:
        K, M = a3.shape[1], a3.shape[2]
    else:
        M, K = a3.shape[1], a3.shape[2]
    if b_mn:
        Kb, N = b3.shape[1], b3.shape[2]
    else:
        N, Kb = b3.shape[1], b3.shape[2]
    assert K == Kb, f"K mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((nb, M, N) if batched else (M, N), device=a.device, dtype=out_dtype)
        assert acc_mode == 0
    o3 = out.unsqueeze(0) if out.dim() == 2 else out
    assert o3.shape == (nb, M, N) and o3.stride(-1) == 1
    assert o3.dtype in (bf16, f32)
    if bias is not None:
        assert bias.dtype == f32 and bias.numel() == N and bias.is_contiguous()
    _lib.call(
        "alm_gemm_bf16",
        a3, int(a_mn), a3.stride(1), a3.stride(0) if nb > 1 else 0,
        b3, int(b_mn), b3.stride(1), b3.stride(0) if nb > 1 else 0,
        o3, int(o3.dtype == f32), o3.stride(1), o3.stride(0) if nb > 1 else 0,
        M, N, K, nb, float(alpha), bias, int(acc_mode), int(split_k),
    )
    return out


def mqa_attn_fwd(q, k, v, *, heads, key_mask=None, causal=True, scale=None, return_lse=True):
    """Multi-query attention forward (attend.py:69-146).

    q: [b, n_q, heads*64] bf16 (last dim contiguous; may be a column slice of a wider buffer)
    k, v: [b, n_k, 64] bf16 (one shared head);  key_mask: [b, n_k] bool/uint8 (True = attend) or None.
    Queries are right-aligned against keys (query i sees keys <= i + n_k - n_q) when causal.
    Returns o [b, n_q, heads*64] bf16 and lse [b, heads, n_q] fp32.
    """
    _check_cuda(q, k, v, key_mask)
    assert q.dtype == bf16 and k.dtype == bf16 and v.dtype == bf16
    b, n_q, hd = q.shape
    n_k = k.shape[1]
    assert hd == heads * 64 and k.shape[-1] == 64 and v.shape[-1] == 64, "dim_head must be 64"
    assert q.stride(-1) == 1 and k.stride(-1) == 1 and v.stride(-1) == 1
    assert q.stride(0) == n_q * q.stride(1)
    o = torch.empty(b, n_q, hd, device=q.device, dtype=bf16)
    lse = torch.empty(b, heads, n_q, device=q.device, dtype=f32) if return_lse else None
    if key_mask is not None:
        key_mask = key_mask.to(torch.uint8).contiguous()
        assert key_mask.shape == (b, n_k)
    if scale is None:
        scale = 64 ** -0.5
    _lib.call(
        "alm_mqa_attn_fwd",
        q, q.stride(1), k, k.stride(1), k.stride(0), v, v.stride(1), v.stride(0), key_mask,
        o, o.stride(1), lse, b, heads, n_q, n_k, int(causal), float(scale),
    )
    return o, lse
