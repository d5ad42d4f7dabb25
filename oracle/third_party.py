"""Oracle restatements of un-vendored dependencies (TEST INFRASTRUCTURE — never imported by the product).

PARITY UNPINNED: `vector-quantize-pytorch>=1.19.3` and `hyper-connections>=0.1.8` are declared in
/root/reference/setup.py:30,40 but their sources are absent and not installable offline.  These
classes restate the published algorithms with the same constructor kwargs the
reference passes (soundstream.py:592-607, audiolm_pytorch.py:446-454) and the same state_dict
keys, so the reference's own files can run on top of them (oracle/ref_import.py).
"""
from __future__ import annotations

import random
from functools import partial

import torch
import torch.nn.functional as F
from torch import nn
from torch.utils._pytree import tree_flatten, tree_unflatten


# ----------------------------------------------------------------------------------------------
# hyper-connections: get_init_and_expand_reduce_stream_functions / HyperConnections / Residual
# ----------------------------------------------------------------------------------------------
class StreamRMSNorm(nn.Module):
    """F.normalize(x) * sqrt(d) * (gamma + 1), gamma initialised to 0."""

    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        return F.normalize(x, dim=-1) * self.scale * (self.gamma + 1)


def hc_width(residuals, num_streams, norm, dyn_alpha_fn, dyn_alpha_scale, static_alpha, dyn_beta_fn,
             dyn_beta_scale, static_beta):
    """'(b s) n d' residual streams -> branch input [b n d], mixed residuals [b n s d], beta [b n s]."""
    bs, n, d = residuals.shape
    r = residuals.reshape(bs // num_streams, num_streams, n, d).permute(0, 2, 1, 3)  # b n s d
    normed = norm(r)
    alpha = torch.tanh(normed @ dyn_alpha_fn) * dyn_alpha_scale + static_alpha  # b n s (s+1)
    beta = torch.tanh(normed @ dyn_beta_fn) * dyn_beta_scale + static_beta      # b n s
    mix = torch.einsum("bnst,bnsd->bntd", alpha, r)
    return mix[..., 0, :], mix[..., 1:, :], beta


def hc_depth(branch_out, mixed, beta):
    """residual' = mixed + beta (x) branch_out, back to '(b s) n d'."""
    out = mixed + torch.einsum("bnd,bns->bnsd", branch_out, beta)
    b, n, s, d = out.shape
    return out.permute(0, 2, 1, 3).reshape(b * s, n, d)


class HyperConnections(nn.Module):
    def __init__(self, num_residual_streams, *, dim, branch=None, layer_index=None):
        super().__init__()
        self.branch = branch
        self.num_residual_streams = s = num_residual_streams
        self.norm = StreamRMSNorm(dim)
        init_index = (layer_index if layer_index is not None else random.randrange(s)) % s
        self.static_beta = nn.Parameter(torch.ones(s))
        alpha0 = torch.zeros(s, 1)
        alpha0[init_index, 0] = 1.0
        self.static_alpha = nn.Parameter(torch.cat((alpha0, torch.eye(s)), dim=1))
        self.dynamic_alpha_fn = nn.Parameter(torch.zeros(dim, s + 1))
        self.dynamic_alpha_scale = nn.Parameter(torch.ones(()) * 1e-2)
        self.dynamic_beta_fn = nn.Parameter(torch.zeros(dim))
        self.dynamic_beta_scale = nn.Parameter(torch.ones(()) * 1e-2)

    def forward(self, residuals, *args, **kwargs):
        branch_in, mixed, beta = hc_width(
            residuals, self.num_residual_streams, self.norm, self.dynamic_alpha_fn, self.dynamic_alpha_scale,
            self.static_alpha, self.dynamic_beta_fn, self.dynamic_beta_scale, self.static_beta)
        out = self.branch(branch_in, *args, **kwargs)
        leaves, spec = tree_flatten(out)
        leaves[0] = hc_depth(leaves[0], mixed, beta)
        return tree_unflatten(leaves, spec)


class PlainResidual(nn.Module):
    """num_residual_streams == 1: x + first leaf of branch(x)."""

    def __init__(self, *, branch, dim=None, **_):
        super().__init__()
        self.branch = branch

    def forward(self, x, *args, **kwargs):
        out = self.branch(x, *args, **kwargs)
        leaves, spec = tree_flatten(out)
        leaves[0] = leaves[0] + x
        return tree_unflatten(leaves, spec)


def get_init_and_expand_reduce_stream_functions(num_streams, disable=False):
    if disable:
        return PlainResidual, nn.Identity(), nn.Identity()

    def expand(x):  # 'b ... -> (b s) ...'
        return x.repeat_interleave(num_streams, dim=0)

    def reduce(x):  # '(b s) ... -> b ...' by sum
        return x.reshape(x.shape[0] // num_streams, num_streams, *x.shape[1:]).sum(dim=1)

    return partial(HyperConnections, num_streams), expand, reduce


# ----------------------------------------------------------------------------------------------
# vector-quantize-pytorch: GroupedResidualVQ -> ResidualVQ -> VectorQuantize -> EuclideanCodebook
# (eval / inference path only: no EMA, no k-means init, no dead-code expiry, no quantize dropout)
# ----------------------------------------------------------------------------------------------
def euclid_nearest(x, embed):
    """x [n, d], embed [c, d] fp32 -> index of nearest code; ties -> lowest index.

    dist = -sqrt(clamp(|x|^2 + |e|^2 - 2 x.e, 0)); idx = argmax(dist)   (EuclideanCodebook.forward)
    """
    x = x.float()
    embed = embed.float()
    x2 = (x ** 2).sum(-1)
    e2 = (embed ** 2).sum(-1)
    xy = (x @ embed.t()) * -2
    d = (x2[:, None] + e2[None, :] + xy).clamp(min=0).sqrt()
    return (-d).argmax(dim=-1)


class _Codebook(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        self.register_buffer("initted", torch.Tensor([False]))
        self.register_buffer("cluster_size", torch.ones(1, codebook_size))
        self.register_buffer("embed_avg", torch.zeros(1, codebook_size, dim))
        self.register_buffer("embed", torch.zeros(1, codebook_size, dim))


class _VQLayer(nn.Module):
    def __init__(self, dim, codebook_size):
        super().__init__()
        self._codebook = _Codebook(dim, codebook_size)


class ResidualVQ(nn.Module):
    def __init__(self, *, dim, num_quantizers, codebook_size, **_):
        super().__init__()
        self.layers = nn.ModuleList([_VQLayer(dim, codebook_size) for _ in range(num_quantizers)])

    @property
    def codebooks(self):
        return torch.stack([l._codebook.embed[0] for l in self.layers])  # q c d

    def forward(self, x):
        assert not self.training, "oracle RVQ restates the eval path only"
        b, n, d = x.shape
        residual = x.float().reshape(b * n, d)
        out = torch.zeros_like(residual)
        idxs = []
        for cb in self.codebooks:
            idx = euclid_nearest(residual, cb)
            quant = cb[idx]
            residual = residual - quant
            out = out + quant
            idxs.append(idx)
        indices = torch.stack(idxs, dim=-1).reshape(b, n, -1)
        losses = torch.zeros(1, len(idxs), device=x.device)
        return out.reshape(b, n, d), indices, losses

    def get_output_from_indices(self, indices):
        cbs = self.codebooks
        out = 0
        for q in range(indices.shape[-1]):
            idx = indices[..., q]
            sel = cbs[q][idx.clamp(min=0)]
            out = out + sel.masked_fill((idx < 0)[..., None], 0.0)
        return out


class GroupedResidualVQ(nn.Module):
    def __init__(self, *, dim, groups=1, **kwargs):
        super().__init__()
        assert dim % groups == 0
        self.groups = groups
        self.kwargs = dict(dim=dim, groups=groups, **kwargs)
        self.rvqs = nn.ModuleList([ResidualVQ(dim=dim // groups, **kwargs) for _ in range(groups)])

    def forward(self, x):
        chunks = x.chunk(self.groups, dim=-1)
        outs = [rvq(c) for rvq, c in zip(self.rvqs, chunks)]
        quantized = torch.cat([o[0] for o in outs], dim=-1)
        indices = torch.stack([o[1] for o in outs])   # g b n q
        losses = torch.stack([o[2] for o in outs])    # g 1 q
        return quantized, indices, losses

    def get_output_from_indices(self, indices):  # g b n q
        return torch.cat([rvq.get_output_from_indices(i) for rvq, i in zip(self.rvqs, indices)], dim=-1)


def seed_codebooks(rq: GroupedResidualVQ, seed=7, std=1.0):
    """Inject random codebooks and mark them initialised (the reference would k-means-init otherwise)."""
    g = torch.Generator().manual_seed(seed)
    for rvq in rq.rvqs:
        for layer in rvq.layers:
            cb = layer._codebook
            cb.embed.copy_(torch.randn(cb.embed.shape, generator=g) * std)
            cb.embed_avg.copy_(cb.embed)
            cb.initted.fill_(True)


# ----------------------------------------------------------------------------------------------
# local-attention (>= 1.9.0): LocalMHA / LocalAttention / SinusoidalEmbeddings (xpos) / FeedForward
# PARITY UNPINNED: the package is declared in /root/reference/setup.py:32 but its source is absent offline.  Restated
# from the published implementation with the constructor kwargs the reference passes (soundstream.py:414-428, 533-543:
# causal, prenorm, qk_rmsnorm, use_xpos, use_rotary_pos_emb, gate_values_per_head, window_size, xpos_scale_base) and the
# upstream state_dict keys (norm.*, to_qkv.weight, q_scale, k_scale, to_v_gate.0.*, to_out.weight,
# attn_fn.rel_pos.inv_freq).  Eval path only (dropout 0).
# ----------------------------------------------------------------------------------------------
def _l2norm(t):
    return F.normalize(t, dim=-1)


def _look_around(x, backward=1, forward=0, pad_value=-1, dim=2):
    """concat each window with its `backward` predecessors / `forward` successors along `dim` (windows at dim 1)."""
    t = x.shape[1]
    dims = (len(x.shape) - dim) * (0, 0)
    padded = F.pad(x, (*dims, backward, forward), value=pad_value)
    return torch.cat([padded[:, ind:(ind + t), ...] for ind in range(forward + backward + 1)], dim=dim)


class SinusoidalEmbeddings(nn.Module):
    def __init__(self, dim, scale_base=None, use_xpos=False, theta=10000):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim)))
        self.use_xpos = use_xpos
        self.scale_base = scale_base
        assert not (use_xpos and scale_base is None)
        self.register_buffer("scale", (torch.arange(0, dim, 2) + 0.4 * dim) / (1.4 * dim), persistent=False)

    def forward(self, x):
        seq_len = x.shape[-2]
        t = torch.arange(seq_len, device=x.device).type_as(self.inv_freq)
        freqs = torch.einsum("i,j->ij", t, self.inv_freq)
        freqs = torch.cat((freqs, freqs), dim=-1)
        if not self.use_xpos:
            return freqs, torch.ones(1, device=x.device)
        power = (t - (seq_len // 2)) / self.scale_base
        scale = self.scale ** power[:, None]
        return freqs, torch.cat((scale, scale), dim=-1)


def _rotate_half(x):
    x1, x2 = x.reshape(*x.shape[:-1], 2, x.shape[-1] // 2).unbind(dim=-2)
    return torch.cat((-x2, x1), dim=-1)


def _apply_rotary(q, k, freqs, scale=1):
    q_len = q.shape[-2]
    q_freqs = freqs[..., -q_len:, :]
    inv_scale = scale ** -1
    if torch.is_tensor(scale) and scale.ndim == 2:
        scale = scale[-q_len:, :]
    q = (q * q_freqs.cos() * scale) + (_rotate_half(q) * q_freqs.sin() * scale)
    k = (k * freqs.cos() * inv_scale) + (_rotate_half(k) * freqs.sin() * inv_scale)
    return q, k


class LocalAttention(nn.Module):
    def __init__(self, *, dim, window_size, causal=False, look_backward=1, look_forward=None, autopad=False,
                 exact_windowsize=False, scale=None, use_rotary_pos_emb=True, use_xpos=False, xpos_scale_base=None,
                 **_):
        super().__init__()
        look_forward = (0 if causal else 1) if look_forward is None else look_forward
        assert not (causal and look_forward > 0)
        self.scale, self.window_size, self.causal = scale, window_size, causal
        self.autopad, self.exact_windowsize = autopad, exact_windowsize
        self.look_backward, self.look_forward = look_backward, look_forward
        self.rel_pos = None
        if use_rotary_pos_emb:
            self.rel_pos = SinusoidalEmbeddings(dim, use_xpos=use_xpos,
                                                scale_base=window_size // 2 if xpos_scale_base is None else xpos_scale_base)

    def forward(self, q, k, v, mask=None, attn_bias=None):
        assert mask is None and attn_bias is None, "restated for the reference's call (no mask, no dynamic bias)"
        lead = q.shape[:-2]
        q, k, v = (t.reshape(-1, *t.shape[-2:]) for t in (q, k, v))
        ws = self.window_size
        orig_n = q.shape[1]
        if self.autopad:
            pad = (-orig_n) % ws
            q, k, v = (F.pad(t, (0, 0, 0, pad)) for t in (q, k, v))
        b, n, dh = q.shape
        scale = dh ** -0.5 if self.scale is None else self.scale
        windows = n // ws
        b_t = torch.arange(n, device=q.device).reshape(1, windows, ws)
        bq, bk, bv = (t.reshape(b, windows, ws, dh) for t in (q, k, v))
        bq = bq * scale
        la = dict(backward=self.look_backward, forward=self.look_forward, pad_value=-1)
        bk, bv = _look_around(bk, **la), _look_around(bv, **la)
        if self.rel_pos is not None:
            pos_emb, xpos_scale = self.rel_pos(bk)
            bq, bk = _apply_rotary(bq, bk, pos_emb, scale=xpos_scale)
        bq_t = b_t[..., :, None]
        bq_k = _look_around(b_t, **la)[..., None, :]
        pad_mask = bq_k == -1
        sim = torch.einsum("bhie,bhje->bhij", bq, bk)
        neg = -torch.finfo(sim.dtype).max
        if self.causal:
            causal_mask = bq_t < bq_k
            if self.exact_windowsize:
                causal_mask = causal_mask | (bq_t > (bq_k + ws * self.look_backward))
            sim = sim.masked_fill(causal_mask, neg)
        sim = sim.masked_fill(pad_mask, neg)
        attn = sim.softmax(dim=-1)
        out = torch.einsum("bhij,bhje->bhie", attn, bv).reshape(b, n, dh)
        return out[:, :orig_n].reshape(*lead, orig_n, dh)


class LocalMHA(nn.Module):
    def __init__(self, *, dim, window_size, dim_head=64, heads=8, dropout=0.0, causal=False, prenorm=False,
                 qk_rmsnorm=False, qk_scale=8, use_xpos=False, xpos_scale_base=None, exact_windowsize=None,
                 gate_values_per_head=False, **kwargs):
        super().__init__()
        inner = dim_head * heads
        self.norm = nn.LayerNorm(dim) if prenorm else None
        self.heads = heads
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.qk_rmsnorm = qk_rmsnorm
        if qk_rmsnorm:
            self.q_scale = nn.Parameter(torch.ones(dim_head))
            self.k_scale = nn.Parameter(torch.ones(dim_head))
        self.attn_fn = LocalAttention(dim=dim_head, window_size=window_size, causal=causal, autopad=True,
                                      scale=(qk_scale if qk_rmsnorm else None),
                                      exact_windowsize=True if exact_windowsize is None else exact_windowsize,
                                      use_xpos=use_xpos, xpos_scale_base=xpos_scale_base, **kwargs)
        self.to_v_gate = nn.Sequential(nn.Linear(dim, heads)) if gate_values_per_head else None
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, mask=None, attn_bias=None):
        if self.norm is not None:
            x = self.norm(x)
        b, n, _ = x.shape
        q, k, v = (t.reshape(b, n, self.heads, -1).transpose(1, 2) for t in self.to_qkv(x).chunk(3, dim=-1))
        if self.qk_rmsnorm:
            q, k = _l2norm(q) * self.q_scale, _l2norm(k) * self.k_scale
        out = self.attn_fn(q, k, v, mask=mask, attn_bias=attn_bias)
        if self.to_v_gate is not None:
            out = out * self.to_v_gate(x).transpose(1, 2)[..., None].sigmoid()
        return self.to_out(out.transpose(1, 2).reshape(b, n, -1))


class _GEGLU(nn.Module):
    def forward(self, x):
        x, gate = x.chunk(2, dim=-1)
        return x * F.gelu(gate)


def LocalFeedForward(dim, mult=4, dropout=0.0):
    inner = int(dim * mult * 2 / 3)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner * 2, bias=False), _GEGLU(), nn.Dropout(dropout),
                         nn.Linear(inner, dim, bias=False))
