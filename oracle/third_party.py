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
