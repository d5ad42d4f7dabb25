"""Logit heads, cross entropy and sampling for the three transformers (autograd glue over C-ABI kernels).

Reference: audiolm_pytorch.py:621, 798 (Linear heads), :965-983, 1325-1361 (grouped per-quantizer heads),
:1561-1565, 1836-1854, 2119-2137 (cross entropy), :98-126 (top-k / gumbel / eos masking).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops
from .transformer import _PackedWeights, _pad8, best_split_k, bf16, f32


def pack_head(w):
    """[V, d] fp32 -> bf16 [pad8(V), d]; the zero rows let dgrad contract over the padded dlogits."""
    V, d = w.shape
    assert d % 8 == 0
    out = torch.zeros(_pad8(V), d, device=w.device, dtype=bf16)
    ops.cast_pad(w.detach(), out=out[:V])
    return out


# training losses run the head GEMM and the cross entropy as ONE fused pair of kernels (alm_gemm_head_ce): the [b, n, V]
# fp32 logits, their strided copies and the separate CE launch disappear.  Off = logits materialised + alm_ce_fwd_bwd.
FUSED_HEAD_CE = True


class LazyLogits:
    """`head(tokens)` that has NOT been computed: what the transformers hand to the wrappers' loss path
    (`_defer_heads=True`).  cross_entropy() consumes it through the fused head + CE kernels; anything else can call
    materialize().  `grouped`: position p uses weight[p mod Q] (per-quantizer heads)."""

    def __init__(self, cache, tokens, weight, bias, key, grouped):
        self.cache, self.tokens, self.weight, self.bias, self.key, self.grouped = cache, tokens, weight, bias, key, grouped
        self.shape = (tokens.shape[0], tokens.shape[1], weight.shape[-2])

    def materialize(self):
        b, n, d = self.tokens.shape
        if self.grouped:
            return self.cache.grouped(self.tokens, self.weight, self.key)
        return self.cache.linear(self.tokens.reshape(-1, d), self.weight, self.bias, self.key).view(b, n, -1)

    def ce_sum(self, labels, ignore_index):
        """sum over positions of the cross entropy against labels [b, n] (ignored positions contribute 0)"""
        b, n, d = self.tokens.shape
        if not self.grouped:
            return self.cache.linear_ce_sum(self.tokens.reshape(-1, d), self.weight, self.bias, self.key,
                                            labels.reshape(-1), ignore_index)
        Q = self.weight.shape[0]
        total = None
        for q in range(min(Q, n)):
            part = self.cache.linear_ce_sum(self.tokens[:, q::Q].reshape(-1, d), self.weight[q], None, (self.key, q),
                                            labels[:, q::Q].reshape(-1), ignore_index)
            total = part if total is None else total + part
        return total


class HeadCache:
    """bf16 operand copies of head weights, refreshed when the parameter version changes."""

    def __init__(self):
        self._pk = _PackedWeights()

    def clear(self):
        self._pk.clear()

    def linear(self, x2d, weight, bias, key):
        packed = self._pk.get(key, [weight], lambda: pack_head(weight))
        V = weight.shape[0]
        # forward uses the first V rows; backward (dgrad) uses all pad8(V) rows against the padded dlogits
        return _LinearPacked.apply(x2d, weight, bias, packed, V)

    def linear_ce_sum(self, x2d, weight, bias, key, labels, ignore_index):
        """sum_r CE(x2d[r] @ weight^T + bias, labels[r]) without materialising the logits (fused head + CE)"""
        packed = self._pk.get(key, [weight], lambda: pack_head(weight))
        return _HeadCESum.apply(x2d, weight, bias, packed, weight.shape[0], labels, ignore_index)

    @torch.no_grad()
    def linear_decode(self, x2d, weight, bias, key):
        """inference-only logits for a handful of rows (decode step): weight-read-bound GEMV, no autograd."""
        packed = self._pk.get(key, [weight], lambda: pack_head(weight))
        V = weight.shape[0]
        x = x2d.to(bf16).contiguous()
        if x.shape[0] <= 8:
            return ops.gemv(x, packed[:V], out_dtype=f32, bias=None if bias is None else bias.detach())
        return ops.gemm(x, packed[:V], out_dtype=f32, bias=None if bias is None else bias.detach().float().contiguous())

    def grouped(self, tokens, weights, key):
        """position p of tokens [b, n, d] uses weights[p mod Q]  ->  logits [b, n, V] fp32."""
        b, n, d = tokens.shape
        Q, V, _ = weights.shape
        logits = torch.empty(b, n, V, device=tokens.device, dtype=f32)
        for q in range(min(Q, n)):
            tq = tokens[:, q::Q]
            lq = self.linear(tq.reshape(-1, d), weights[q], None, (key, q))
            logits[:, q::Q] = lq.view(b, -1, V)
        return logits


class _LinearPacked(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, w_packed, V):
        x = x.to(bf16).contiguous()
        y = ops.gemm(x, w_packed[:V], out_dtype=f32,
                     bias=None if bias is None else bias.detach().float().contiguous())
        ctx.save_for_backward(x)
        # plain attribute, not save_for_backward: the packed copy may have been built under generate()'s
        # torch.inference_mode() and inference tensors cannot be saved for backward
        ctx.w_packed = w_packed
        ctx.has_bias = bias is not None
        ctx.V = V
        ctx.d = weight.shape[1]
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        w_packed = ctx.w_packed
        V, d = ctx.V, ctx.d
        dyp = ops.cast_pad(dy.contiguous(), _pad8(V))                    # bf16 [R, pad8(V)], zero tail
        dx = ops.gemm(dyp, w_packed, b_mn=True)[:, :d]                   # contraction over pad8(V) rows
        dw = torch.zeros(V, d, device=dy.device, dtype=f32)
        s = best_split_k(V, d, x.shape[0])
        ops.gemm(dyp[:, :V], x[:, :d], a_mn=True, b_mn=True, out=dw, acc_mode=2 if s > 1 else 1, split_k=s)
        db = dy.sum(0) if ctx.has_bias else None
        return dx, dw, db, None, None


class _HeadCESum(torch.autograd.Function):
    """sum of the row cross entropies of (x @ W^T + bias): alm_gemm_head_ce mode 1 + alm_ce_finish forward, mode 2
    (recompute, emit d logits as bf16) + the usual dgrad / wgrad GEMMs backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, w_packed, V, labels, ignore_index):
        x = x.to(bf16).contiguous()
        labels = labels.contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        lse, rows = ops.head_ce_fwd(x, w_packed[:V], b32, labels, ignore_index)
        ctx.save_for_backward(x, lse, labels)
        ctx.w_packed, ctx.b32 = w_packed, b32      # plain attributes: see _LinearPacked
        ctx.V, ctx.d, ctx.ignore = V, weight.shape[1], ignore_index
        return rows.sum()

    @staticmethod
    def backward(ctx, g):
        x, lse, labels = ctx.saved_tensors
        V, d, w_packed = ctx.V, ctx.d, ctx.w_packed
        dlog = torch.zeros(x.shape[0], _pad8(V), device=x.device, dtype=bf16)       # zero padding columns for the dgrad
        one = torch.ones(1, device=x.device, dtype=f32)
        ops.head_ce_bwd(x, w_packed[:V], ctx.b32, labels, ctx.ignore, lse, g.reshape(1).to(f32).contiguous(), one, dlog)
        dx = ops.gemm(dlog, w_packed, b_mn=True)[:, :d]
        dw = torch.zeros(V, d, device=x.device, dtype=f32)
        s = best_split_k(V, d, x.shape[0])
        ops.gemm(dlog[:, :V], x[:, :d], a_mn=True, b_mn=True, out=dw, acc_mode=2 if s > 1 else 1, split_k=s)
        db = dlog[:, :V].float().sum(0) if ctx.b32 is not None else None
        return dx, dw, db, None, None, None, None


class CrossEntropyFn(torch.autograd.Function):
    """mean cross entropy over rows whose label != ignore_index (fused forward + d logits kernel)."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        R, V = logits.shape
        den = (labels != ignore_index).sum().to(f32).clamp(min=1.0)
        one = torch.ones((), device=logits.device, dtype=f32)
        rows, dlog = ops.ce_fwd_bwd(logits, labels.contiguous(), ignore_index=ignore_index, scale_num=one,
                                    scale_den=den, want_grad=True)
        ctx.save_for_backward(dlog)
        ctx.V = V
        return rows.sum() / den

    @staticmethod
    def backward(ctx, g):
        (dlog,) = ctx.saved_tensors
        return dlog[:, :ctx.V].float() * g, None, None


def cross_entropy(logits, labels, ignore_index=-1):
    """F.cross_entropy(rearrange(logits, 'b n c -> b c n'), labels, ignore_index=...) for logits [b, n, c]."""
    if isinstance(logits, LazyLogits):
        den = (labels != ignore_index).sum().to(f32).clamp(min=1.0)
        return logits.ce_sum(labels, ignore_index) / den
    V = logits.shape[-1]
    return CrossEntropyFn.apply(logits.reshape(-1, V), labels.reshape(-1), ignore_index)


# ---- sampling helpers (audiolm_pytorch.py:98-126) -------------------------------------------------
def top_k(logits, thres=0.5):
    k = max(int((1 - thres) * logits.shape[-1]), 1)
    val, ind = torch.topk(logits, k)
    return torch.full_like(logits, float("-inf")).scatter_(1, ind, val)


def gumbel_sample(t, temperature=1.0, dim=-1):
    noise = torch.zeros_like(t).uniform_(0, 1)
    g = -torch.log(-torch.log(noise + 1e-20) + 1e-20)
    return (t / temperature + g).argmax(dim=dim)


def mask_out_after_eos_id(t, eos_id, mask_value=-1, keep_eos=True):
    eos = (t == eos_id).float()
    if keep_eos:
        eos = F.pad(eos, (1, -1))
    return t.masked_fill(eos.cumsum(dim=-1) > 0, mask_value)


def generate_mask_with_prob(shape, mask_prob, device):
    """forgetful causal mask (audiolm_pytorch.py:82-89): exactly int(n*p) keys dropped, position 0 kept."""
    n = shape[-1]
    r = torch.randn(shape, device=device)
    r[:, 0] = -torch.finfo(r.dtype).max
    num = min(int(n * mask_prob), n - 1)
    idx = r.topk(num, dim=-1).indices
    return ~torch.zeros(shape, device=device).scatter(1, idx, 1.0).bool()
