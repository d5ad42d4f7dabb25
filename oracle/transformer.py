"""Functional fp32 CPU oracle of the transformer hot path — TEST INFRASTRUCTURE ONLY.

Every function takes a flat ``state`` dict (the reference's state_dict, same keys) and plain tensors.
Each docstring cites the reference lines it restates (/root/reference/audiolm_pytorch/...).  Pinned
against the real reference by oracle/make_golden.py (fixtures under tests/golden/), except for the
Hyper-Connections arithmetic, which follows oracle/third_party.py (PARITY UNPINNED upstream).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

NEG = lambda dt: -torch.finfo(dt).max  # noqa: E731  masked_fill value (attend.py:128,135)


def sub(state: dict, prefix: str) -> dict:
    """view of `state` restricted to keys under `prefix.` (prefix stripped)."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in state.items() if k.startswith(p)}


# ---------------------------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------------------------
def layer_norm(x, gamma, eps=1e-5):
    """LayerNorm with learned gamma and a zero beta buffer (audiolm_pytorch.py:191-198)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * gamma


def attend(q, k, v, mask=None, attn_bias=None, causal=True):
    """Attend.forward math path (attend.py:98-146): q [b h i d], k/v [b j d] (one shared head)."""
    scale = q.shape[-1] ** -0.5
    sim = torch.einsum("bhid,bjd->bhij", q, k) * scale
    if attn_bias is not None:
        sim = sim + attn_bias
    if mask is not None:
        sim = sim.masked_fill(~mask[:, None, None, :], NEG(sim.dtype))
    if causal:
        i, j = sim.shape[-2:]
        future = torch.ones(i, j, dtype=torch.bool, device=q.device).triu(j - i + 1)
        sim = sim.masked_fill(future, NEG(sim.dtype))
    return torch.einsum("bhij,bjd->bhid", sim.softmax(-1), v)


def attention(st, x, heads, mask=None, attn_bias=None, kv_cache=None, value_residual=None):
    """Attention.forward, self-attention case (audiolm_pytorch.py:307-406).

    st keys: norm.gamma, to_q.weight [h*dh, d], to_kv.weight [2*dh, d], to_out.0.weight [d, h*dh].
    Returns out, stacked (k, v) cache [2 b n dh], and the un-mixed values (for value-residual).
    """
    xn = layer_norm(x, st["norm.gamma"])
    q = xn @ st["to_q.weight"].t()
    # NB: the reference binds kv_input = x BEFORE the pre-norm (:325 vs :347), so keys/values are
    # projected from the UN-normalised branch input while queries use the normalised one (:351).
    kv = x @ st["to_kv.weight"].t()
    k, v = kv.chunk(2, dim=-1)
    orig_v = v
    if value_residual is not None:
        v = 0.5 * (v + value_residual)
    if kv_cache is not None:
        k = torch.cat((kv_cache[0], k), dim=-2)
        v = torch.cat((kv_cache[1], v), dim=-2)
    new_cache = torch.stack((k, v))
    b, n, _ = q.shape
    qh = q.reshape(b, n, heads, -1).permute(0, 2, 1, 3)
    o = attend(qh, k, v, mask=mask, attn_bias=attn_bias, causal=True)
    o = o.permute(0, 2, 1, 3).reshape(b, n, -1)
    return o @ st["to_out.0.weight"].t(), new_cache, orig_v


def feed_forward(st, x):
    """FeedForward = LN -> Linear(d, 2*inner) -> GEGLU -> LN(inner) -> Linear(inner, d) (:246-260)."""
    h = layer_norm(x, st["0.gamma"]) @ st["1.weight"].t()
    a, gate = h.chunk(2, dim=-1)
    g = F.gelu(gate) * a
    return layer_norm(g, st["3.gamma"]) @ st["5.weight"].t()


def hyper_width(st, residuals, s):
    """Hyper-Connections width connection (third-party; see oracle/third_party.py): residuals [(b s) n d]."""
    bs, n, d = residuals.shape
    r = residuals.reshape(bs // s, s, n, d).permute(0, 2, 1, 3)
    normed = F.normalize(r, dim=-1) * (d ** 0.5) * (st["norm.gamma"] + 1)
    alpha = torch.tanh(normed @ st["dynamic_alpha_fn"]) * st["dynamic_alpha_scale"] + st["static_alpha"]
    beta = torch.tanh(normed @ st["dynamic_beta_fn"]) * st["dynamic_beta_scale"] + st["static_beta"]
    mix = torch.einsum("bnst,bnsd->bntd", alpha, r)
    return mix[..., 0, :], mix[..., 1:, :], beta


def hyper_depth(branch_out, mixed, beta):
    out = mixed + torch.einsum("bnd,bns->bnsd", branch_out, beta)
    b, n, s, d = out.shape
    return out.permute(0, 2, 1, 3).reshape(b * s, n, d)


def rel_pos_bias(st, i, j):
    """RelativePositionBias.forward (audiolm_pytorch.py:225-242): MLP over the offsets -(j-1)..(j-1), gathered
    to [h, i, j] with queries right-aligned.  st holds net.{k}.0.{weight,bias} (SiLU layers) and the last Linear."""
    n_layers = max(int(k.split(".")[1]) for k in st if k.startswith("net."))
    x = torch.arange(-j + 1, j, device=st[f"net.{n_layers}.weight"].device).float()[:, None]
    for k in range(n_layers):
        x = F.silu(x @ st[f"net.{k}.0.weight"].t() + st[f"net.{k}.0.bias"])
    x = x @ st[f"net.{n_layers}.weight"].t() + st[f"net.{n_layers}.bias"]
    i_pos = torch.arange(i, device=x.device) + (j - i)
    j_pos = torch.arange(j, device=x.device)
    rel = i_pos[:, None] - j_pos[None, :] + (j - 1)
    return x[rel].permute(2, 0, 1)


def transformer(st, x, *, heads, depth, num_streams=4, self_attn_mask=None, attn_bias=None, kv_cache=None,
                add_value_residual=True):
    """Transformer.forward without cross attention (audiolm_pytorch.py:461-560).

    x [b n d] (already the full sequence; with kv_cache only x[:, cache_len:] is processed).
    Returns normed output [b n' d] and the new kv cache [depth 2 b n dh].
    (grad_shrink is identity in forward, :93-94.)
    """
    cache_len = 0 if kv_cache is None else kv_cache.shape[-2]
    n_full = x.shape[1]
    x = x[:, cache_len:]
    if attn_bias is None and "rel_pos_bias.net.0.0.weight" in st:  # flash_attn=False models (:442, 503)
        attn_bias = rel_pos_bias(sub(st, "rel_pos_bias"), n_full, n_full)
    if attn_bias is not None:
        attn_bias = attn_bias[..., cache_len:, :]
    if num_streams > 1:
        x = x.repeat_interleave(num_streams, dim=0)
    value_res = None
    caches = []
    for i in range(depth):
        a_st = sub(st, f"layers.{i}.0")
        f_st = sub(st, f"layers.{i}.2")
        layer_cache = None if kv_cache is None else kv_cache[i]
        if num_streams > 1:
            xin, mixed, beta = hyper_width(a_st, x, num_streams)
        else:
            xin = x
        out, kv, values = attention(sub(a_st, "branch"), xin, heads, mask=self_attn_mask, attn_bias=attn_bias,
                                    kv_cache=layer_cache, value_residual=value_res)
        x = hyper_depth(out, mixed, beta) if num_streams > 1 else x + out
        if add_value_residual and value_res is None:
            value_res = values
        caches.append(kv)
        if num_streams > 1:
            xin, mixed, beta = hyper_width(f_st, x, num_streams)
        else:
            xin = x
        out = feed_forward(sub(f_st, "branch"), xin)
        x = hyper_depth(out, mixed, beta) if num_streams > 1 else x + out
    if num_streams > 1:
        x = x.reshape(x.shape[0] // num_streams, num_streams, *x.shape[1:]).sum(dim=1)
    return layer_norm(x, st["norm.gamma"]), torch.stack(caches)


# ---------------------------------------------------------------------------------------------
# the three transformers (flash_attn=True / rel_pos_bias=False configuration, no text conditioning)
# ---------------------------------------------------------------------------------------------
def grouped_logits(weights, tokens):
    """position p of `tokens` [b n d] uses weights[p mod q] (audiolm_pytorch.py:965-983, 1343-1361)."""
    q = weights.shape[0]
    n = tokens.shape[1]
    idx = torch.arange(n, device=tokens.device) % q
    return torch.einsum("ncd,bnd->bnc", weights[idx], tokens)


def _offsets(n, q, step, device):
    return (torch.arange(n, device=device) % q) * step


def semantic_forward(st, ids, *, heads, depth, num_streams=4, self_attn_mask=None, kv_cache=None):
    """SemanticTransformer.forward (audiolm_pytorch.py:671-724), ids [b n] int64 -> logits [b n+1 V+1]."""
    tok = F.embedding(ids, st["semantic_embedding.weight"])
    start = st["start_token"].expand(ids.shape[0], 1, -1)
    x = torch.cat((start, tok), dim=1)
    if self_attn_mask is not None:
        self_attn_mask = F.pad(self_attn_mask, (1, 0), value=True)
    h, cache = transformer(sub(st, "transformer"), x, heads=heads, depth=depth, num_streams=num_streams,
                           self_attn_mask=self_attn_mask, kv_cache=kv_cache)
    return h @ st["to_logits.weight"].t() + st["to_logits.bias"], cache


def coarse_forward(st, semantic_ids, coarse_ids, *, heads, depth, codebook_size, num_coarse_quantizers,
                   num_streams=4, self_attn_mask=None, kv_cache=None, embed_cache=None,
                   return_only_coarse_logits=False):
    """CoarseTransformer.forward (audiolm_pytorch.py:858-990). Returns (semantic_logits, coarse_logits), caches."""
    dev = semantic_ids.device
    q = num_coarse_quantizers
    b = semantic_ids.shape[0]
    nc = coarse_ids.shape[-1]
    coarse_ids = coarse_ids + _offsets(nc, q, codebook_size, dev)  # NB: offset step is codebook_size (:896)
    sem = F.embedding(semantic_ids, st["semantic_embedding.weight"])
    coarse = F.embedding(coarse_ids, st["coarse_embedding.weight"])
    coarse = coarse + st["coarse_quantize_embedding.weight"][torch.arange(nc, device=dev) % q]
    S = sem.shape[1]
    x = torch.cat((st["semantic_start_token"].expand(b, 1, -1), sem,
                   st["coarse_start_token"].expand(b, 1, -1), coarse), dim=1)
    attn_bias = None
    if "cross_attn_bias" in st:  # :920-936: one learned scalar per head between the two segments
        n_all = x.shape[1]
        attn_bias = rel_pos_bias(sub(st, "transformer.rel_pos_bias"), n_all, n_all)
        is_sem = torch.arange(n_all, device=dev) < S + 1
        attn_bias = torch.where(is_sem[:, None] ^ is_sem[None, :], st["cross_attn_bias"], attn_bias)
    h, cache = transformer(sub(st, "transformer"), x, heads=heads, depth=depth, num_streams=num_streams,
                           self_attn_mask=self_attn_mask, attn_bias=attn_bias, kv_cache=kv_cache)
    if embed_cache is not None:
        h = torch.cat((embed_cache, h), dim=-2)
    pred_sem, pred_coarse = h[:, :S], h[:, S + 1:]
    sem_logits = None
    if not return_only_coarse_logits and "to_semantic_logits.weight" in st:
        sem_logits = pred_sem @ st["to_semantic_logits.weight"].t() + st["to_semantic_logits.bias"]
    return (sem_logits, grouped_logits(st["coarse_logit_weights"], pred_coarse)), (cache, h)


def fine_pos_bias(st, n, nf, qc, qf, dev):
    """the engineered coarse/fine attention bias (audiolm_pytorch.py:1229-1298) -> [h, L, L], L = n + nf + 2."""
    cs, fs = -(-n // qc), -(-nf // qf)
    max_seq = max(cs, fs)
    pos = torch.cat((torch.tensor([-1], device=dev), torch.arange(cs, device=dev).repeat_interleave(qc)[:n],
                     torch.tensor([-1], device=dev), torch.arange(fs, device=dev).repeat_interleave(qf)[:nf]))
    off = torch.cat((torch.tensor([0], device=dev), torch.arange(qc, device=dev).repeat(cs)[:n],
                     torch.tensor([0], device=dev), torch.arange(qf, device=dev).repeat(fs)[:nf] + qc))
    num_off = qc + qf
    rel_seq, rel_off = 2 * max_seq - 1, 2 * num_off - 1
    inp = torch.stack((torch.arange(rel_seq, device=dev).repeat_interleave(rel_off),
                       torch.arange(rel_off, device=dev).repeat(rel_seq)), dim=-1).float()
    t = F.silu(inp @ st["pos_bias_mlp.0.weight"].t() + st["pos_bias_mlp.0.bias"])
    t = F.silu(t @ st["pos_bias_mlp.2.weight"].t() + st["pos_bias_mlp.2.bias"])
    t = t @ st["pos_bias_mlp.4.weight"].t() + st["pos_bias_mlp.4.bias"]
    pc = pos.clamp(min=0)
    d_pos = pc[:, None] - pc[None, :] + max_seq - 1
    d_off = off[:, None] - off[None, :] + num_off - 1
    bias = t[d_pos * rel_off + d_off].permute(2, 0, 1)
    start = pos == -1
    return torch.where(start[:, None] | start[None, :], st["null_pos_bias"], bias)


def fine_forward(st, coarse_ids, fine_ids, *, heads, depth, codebook_size, num_coarse_quantizers,
                 num_fine_quantizers, num_streams=4, self_attn_mask=None, kv_cache=None, embed_cache=None,
                 return_only_fine_logits=False, pad_id=-1):
    """FineTransformer.forward (audiolm_pytorch.py:1136-1368)."""
    dev = coarse_ids.device
    b, n = coarse_ids.shape
    eos_id = codebook_size
    keep = (coarse_ids != pad_id) & (coarse_ids != eos_id)
    coarse_ids = coarse_ids.masked_fill(~keep, 0)
    nf = fine_ids.shape[-1]
    keep = F.pad(keep, (1, nf + 1), value=True)
    self_attn_mask = keep if self_attn_mask is None else (self_attn_mask & keep)
    qc, qf = num_coarse_quantizers, num_fine_quantizers
    coarse = F.embedding(coarse_ids + _offsets(n, qc, codebook_size, dev), st["coarse_embedding.weight"])
    fine = F.embedding(fine_ids + _offsets(nf, qf, codebook_size, dev), st["fine_embedding.weight"])
    coarse = coarse + st["coarse_quantize_embedding.weight"][torch.arange(n, device=dev) % qc]
    fine = fine + st["fine_quantize_embedding.weight"][torch.arange(nf, device=dev) % qf]
    x = torch.cat((st["coarse_start_token"].expand(b, 1, -1), coarse,
                   st["fine_start_token"].expand(b, 1, -1), fine), dim=1)
    attn_bias = None
    if "pos_bias_mlp.0.weight" in st:
        attn_bias = fine_pos_bias(st, n, nf, qc, qf, dev)
    h, cache = transformer(sub(st, "transformer"), x, heads=heads, depth=depth, num_streams=num_streams,
                           self_attn_mask=self_attn_mask, attn_bias=attn_bias, kv_cache=kv_cache)
    if embed_cache is not None:
        h = torch.cat((embed_cache, h), dim=-2)
    pred_coarse, pred_fine = h[:, :n], h[:, n + 1:]
    coarse_logits = None
    if not return_only_fine_logits and "coarse_logit_weights" in st:
        coarse_logits = grouped_logits(st["coarse_logit_weights"], pred_coarse)
    return (coarse_logits, grouped_logits(st["fine_logit_weights"], pred_fine)), (cache, h)


# ---------------------------------------------------------------------------------------------
# losses / sampling helpers used by the wrappers
# ---------------------------------------------------------------------------------------------
def cross_entropy(logits, labels, ignore_index=-1):
    """F.cross_entropy(rearrange(logits,'b n c -> b c n'), labels, ignore_index) (:1561-1565 etc.)."""
    return F.cross_entropy(logits.transpose(1, 2), labels, ignore_index=ignore_index)


def coarse_wrapper_loss(sem_logits, coarse_logits, sem_labels, coarse_labels, sem_weight=1.0, pad_id=-1):
    """loss mix of CoarseTransformerWrapper.forward with unique_consecutive=False (:1826-1854)."""
    n_sem, n_coarse = sem_logits.shape[1], coarse_logits.shape[1]
    ls = cross_entropy(sem_logits, sem_labels, pad_id)
    lc = cross_entropy(coarse_logits, coarse_labels, pad_id)
    return (ls * n_sem * sem_weight + lc * n_coarse) / (n_sem + n_coarse)


def top_k_filter(logits, thres=0.9):
    """top_k (audiolm_pytorch.py:111-117): keep the k = max(int((1-thres)*V),1) largest, others -> -inf."""
    k = max(int((1 - thres) * logits.shape[-1]), 1)
    val, ind = torch.topk(logits, k)
    out = torch.full_like(logits, float("-inf"))
    return out.scatter(1, ind, val)


def gumbel_argmax(logits, uniform, temperature=1.0):
    """gumbel_sample with the uniform noise passed in (audiolm_pytorch.py:98-109)."""
    g = -torch.log(-torch.log(uniform + 1e-20) + 1e-20)
    return (logits / temperature + g).argmax(dim=-1)


def fcm_mask(shape, mask_prob, generator=None):
    """generate_mask_with_prob (audiolm_pytorch.py:82-89): exactly int(n*p) keys masked, position 0 kept."""
    n = shape[-1]
    r = torch.randn(shape, generator=generator)
    r[:, 0] = -torch.finfo(r.dtype).max
    num = min(int(n * mask_prob), n - 1)
    idx = r.topk(num, dim=-1).indices
    return ~torch.zeros(shape).scatter(1, idx, 1.0).bool()

def fixed_fcm(shape, mask_prob, device=None):
    """deterministic stand-in for generate_mask_with_prob (audiolm_pytorch.py:82-89) used on BOTH sides of the wrapper
    goldens: same construction (exactly int(n*p) keys dropped, position 0 kept), seeded by the shape."""
    shape = tuple(shape)
    return fcm_mask(shape, mask_prob, torch.Generator().manual_seed(1000 + shape[0] * 131 + shape[-1]))
