"""Raw CUDA ops (C ABI) vs fp32 torch restatements + autograd: Hyper-Connections, GEGLU+LN, CE, attention bwd."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"
bf16 = torch.bfloat16


def rel_err(a, b):
    a, b = a.float(), b.float()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


def make_hc(d, S=4, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s, k=1.0: (torch.randn(*s, generator=g) * k).to(DEV)  # noqa: E731
    hc = dict(gamma=r(d, k=0.1), dyn_alpha=r(d, S + 1, k=0.05), dyn_beta=r(d, k=0.05),
              static_alpha=torch.cat((torch.eye(S)[:, :1], torch.eye(S)), 1).to(DEV) + r(S, S + 1, k=0.1),
              static_beta=torch.ones(S, device=DEV) + r(S, k=0.1),
              alpha_scale=torch.tensor(0.3, device=DEV), beta_scale=torch.tensor(0.3, device=DEV))
    ln_gamma = 1 + r(d, k=0.1)
    return hc, ln_gamma


def hc_ref(hc, ln_gamma, R, d):
    normed = F.normalize(R, dim=-1) * math.sqrt(d) * (hc["gamma"] + 1)
    alpha = torch.tanh(normed @ hc["dyn_alpha"]) * hc["alpha_scale"] + hc["static_alpha"]
    beta = torch.tanh(normed @ hc["dyn_beta"]) * hc["beta_scale"] + hc["static_beta"]
    mix = torch.einsum("mst,msd->mtd", alpha, R)
    bin_ = mix[:, 0]
    xn = F.layer_norm(bin_, (d,)) * ln_gamma
    return mix[:, 1:], bin_, xn, beta


# d = 1536 / 2048 run the first-generation kernels of hyper_conn.cu (the only ones that take d > 1024)
@pytest.mark.parametrize("M,d,expand", [(37, 64, False), (300, 1024, False), (64, 128, True), (200, 1024, True),
                                        (130, 1536, False), (66, 2048, True)])
def test_hc_pre_fwd_bwd(M, d, expand):
    from audiolm_pytorch_b200 import ops

    torch.manual_seed(M + d)
    S = 4
    hc, ln_gamma = make_hc(d, seed=d)
    hc_leaf = {k: v.clone().requires_grad_(True) for k, v in hc.items()}
    lng_leaf = ln_gamma.clone().requires_grad_(True)
    if expand:
        x = torch.randn(M, d, device=DEV)
        x_leaf = x.clone().requires_grad_(True)
        R = x_leaf[:, None, :].expand(M, S, d)
        outs = ops.hc_pre_fwd(hc, ln_gamma, x_expand=x, M=M, d=d)
    else:
        R_in = torch.randn(M, S, d, device=DEV).to(bf16)
        Y = torch.randn(M, d, device=DEV).to(bf16)
        bp = 1 + 0.2 * torch.randn(M, S, device=DEV)
        Ri, Yl, bpl = (t.float().clone().requires_grad_(True) for t in (R_in, Y, bp))
        R = Ri + bpl[..., None] * Yl[:, None, :]
        outs = ops.hc_pre_fwd(hc, ln_gamma, R_in=R_in, Y=Y, beta_prev=bp, M=M, d=d)
    R_out, bin_, xn, beta, aux = outs
    r_out, r_bin, r_xn, r_beta = hc_ref(hc_leaf, lng_leaf, R, d)
    assert rel_err(R_out, r_out) < 1e-2
    assert rel_err(bin_, r_bin) < 1e-2
    assert rel_err(xn, r_xn) < 1.5e-2
    assert rel_err(beta, r_beta) < 1e-3

    w1 = torch.randn(M, S, d, device=DEV).to(bf16)
    w2 = torch.randn(M, d, device=DEV).to(bf16)
    w3 = torch.randn(M, d, device=DEV).to(bf16)
    w4 = torch.randn(M, S, device=DEV)
    loss = (r_out * w1.float()).sum() + (r_xn * w2.float()).sum() + (r_bin * w3.float()).sum() + (r_beta * w4).sum()
    loss.backward()
    grads = {k: torch.zeros_like(v) for k, v in hc.items()}
    g_ln = torch.zeros_like(ln_gamma)
    if expand:
        dx = ops.hc_pre_bwd(hc, ln_gamma, grads, g_ln, aux, w1, w2, w4, dbin_extra=w3, x_expand=x, dx_scale=0.1,
                            M=M, d=d)
        assert rel_err(dx, 0.1 * x_leaf.grad) < 2e-2
    else:
        dR_in, dY, dbp = ops.hc_pre_bwd(hc, ln_gamma, grads, g_ln, aux, w1, w2, w4, dbin_extra=w3, R_in=R_in, Y=Y,
                                        beta_prev=bp, M=M, d=d)
        assert rel_err(dR_in, Ri.grad) < 2e-2
        assert rel_err(dY, Yl.grad) < 2e-2
        assert rel_err(dbp, bpl.grad) < 2e-2
    torch.cuda.synchronize()
    for k in hc:
        assert rel_err(grads[k], hc_leaf[k].grad) < 3e-2, k
    assert rel_err(g_ln, lng_leaf.grad) < 3e-2


@pytest.mark.parametrize("M,d", [(50, 64), (300, 1024), (70, 1536)])
def test_hc_post_fwd_bwd(M, d):
    from audiolm_pytorch_b200 import ops

    torch.manual_seed(M)
    S = 4
    R_in = torch.randn(M, S, d, device=DEV).to(bf16)
    Y = torch.randn(M, d, device=DEV).to(bf16)
    bp = 1 + 0.2 * torch.randn(M, S, device=DEV)
    lng = (1 + 0.1 * torch.randn(d, device=DEV))
    Ri, Yl, bpl, gl = (t.float().clone().requires_grad_(True) for t in (R_in, Y, bp, lng))
    xs = (Ri + bpl[..., None] * Yl[:, None, :]).sum(1)
    ref = F.layer_norm(xs, (d,)) * gl
    out, stats = ops.hc_post_fwd(R_in, Y, bp, lng, M=M, d=d)
    assert rel_err(out, ref) < 1.5e-2
    w = torch.randn(M, d, device=DEV).to(bf16)
    (ref * w.float()).sum().backward()
    g_ln = torch.zeros_like(lng)
    dR, dY, dbp = ops.hc_post_bwd(R_in, Y, bp, lng, stats, w, g_ln, M=M, d=d)
    assert rel_err(dR, Ri.grad) < 2e-2
    assert rel_err(dY, Yl.grad) < 2e-2
    assert rel_err(dbp, bpl.grad) < 2e-2
    assert rel_err(g_ln, gl.grad) < 2e-2


@pytest.mark.parametrize("M,inner", [(33, 170), (256, 2730), (64, 512)])
def test_geglu_ln(M, inner):
    from audiolm_pytorch_b200 import ops

    torch.manual_seed(inner)
    ip = (inner + 7) // 8 * 8
    h = torch.randn(M, 2 * ip, device=DEV).to(bf16)
    gamma = 1 + 0.1 * torch.randn(inner, device=DEV)
    hl = h.float().clone().requires_grad_(True)
    gml = gamma.clone().requires_grad_(True)
    a, gate = hl[:, :inner], hl[:, ip:ip + inner]
    ref = F.layer_norm(F.gelu(gate) * a, (inner,)) * gml
    gn, stats = ops.geglu_ln_fwd(h, gamma, inner=inner, inner_pad=ip)
    assert rel_err(gn[:, :inner], ref) < 1.5e-2
    assert (gn[:, inner:] == 0).all()
    w = torch.randn(M, ip, device=DEV).to(bf16)
    (ref * w[:, :inner].float()).sum().backward()
    g_gamma = torch.zeros_like(gamma)
    dh = ops.geglu_ln_bwd(h, gamma, stats, w, g_gamma, inner=inner, inner_pad=ip)
    assert rel_err(dh[:, :inner], hl.grad[:, :inner]) < 2e-2
    assert rel_err(dh[:, ip:ip + inner], hl.grad[:, ip:ip + inner]) < 2e-2
    assert rel_err(g_gamma, gml.grad) < 2e-2


@pytest.mark.parametrize("R,V", [(40, 65), (1000, 1025), (300, 501)])
def test_cross_entropy(R, V):
    from audiolm_pytorch_b200 import ops

    torch.manual_seed(V)
    logits = (torch.randn(R, V, device=DEV) * 3)
    labels = torch.randint(0, V, (R,), device=DEV)
    labels[::7] = -1
    ll = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(ll, labels, ignore_index=-1)
    ref.backward()
    num = torch.ones((), device=DEV)
    den = (labels != -1).sum().float()
    rows, dlog = ops.ce_fwd_bwd(logits, labels, scale_num=num, scale_den=den)
    loss = rows.sum() / den
    assert abs(loss.item() - ref.item()) < 1e-4 * max(1, abs(ref.item()))
    assert rel_err(dlog[:, :V], ll.grad) < 1e-2
    assert (dlog[:, V:] == 0).all()


def attend_ref(q, k, v, mask, causal):
    scale = q.shape[-1] ** -0.5
    sim = torch.einsum("bhid,bjd->bhij", q, k) * scale
    neg = -torch.finfo(sim.dtype).max
    if mask is not None:
        sim = sim.masked_fill(~mask[:, None, None, :], neg)
    if causal:
        i, j = sim.shape[-2:]
        sim = sim.masked_fill(torch.ones(i, j, dtype=torch.bool, device=q.device).triu(j - i + 1), neg)
    return torch.einsum("bhij,bjd->bhid", sim.softmax(-1), v)


@pytest.mark.parametrize("b,h,n_q,n_k,masked,causal", [
    (1, 1, 128, 128, False, True),
    (2, 8, 256, 256, False, True),
    (2, 4, 300, 300, True, True),
    (1, 8, 1024, 1024, True, True),
    (2, 2, 200, 200, True, False),
])
def test_attn_bwd(b, h, n_q, n_k, masked, causal):
    from audiolm_pytorch_b200 import ops

    torch.manual_seed(n_q * 3 + h)
    q = torch.randn(b, n_q, h * 64, device=DEV).to(bf16)
    k = torch.randn(b, n_k, 64, device=DEV).to(bf16)
    v = torch.randn(b, n_k, 64, device=DEV).to(bf16)
    d_o = torch.randn(b, n_q, h * 64, device=DEV).to(bf16)
    mask = None
    if masked:
        mask = torch.rand(b, n_k, device=DEV) > 0.15
        mask[:, 0] = True
    o, lse = ops.mqa_attn_fwd(q, k, v, heads=h, key_mask=mask, causal=causal)
    dq, dk, dv = ops.mqa_attn_bwd(q, k, v, o, d_o, lse, heads=h, key_mask=mask, causal=causal)
    torch.cuda.synchronize()
    ql, kl, vl = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    qh = ql.reshape(b, n_q, h, 64).permute(0, 2, 1, 3)
    ref = attend_ref(qh, kl, vl, mask, causal).permute(0, 2, 1, 3).reshape(b, n_q, h * 64)
    (ref * d_o.float()).sum().backward()
    assert rel_err(dq, ql.grad) < 2e-2
    assert rel_err(dk, kl.grad) < 2e-2
    assert rel_err(dv, vl.grad) < 2e-2


def test_axpby_cast():
    from audiolm_pytorch_b200 import ops

    x = torch.randn(100, 64, device=DEV).to(bf16)
    y = torch.randn(100, 128, device=DEV).to(bf16)[:, 64:]
    out = ops.axpby(x, 0.5, y, 0.5)
    assert rel_err(out, 0.5 * (x.float() + y.float())) < 1e-2
    w = torch.randn(50, 2730, device=DEV)
    p = ops.cast_pad(w, 2736)
    assert torch.equal(p[:, :2730], w.to(bf16)) and (p[:, 2730:] == 0).all()


def test_embed_gather_scatter_vs_torch():
    """alm_embed_gather / alm_embed_scatter vs index_select + index_add on the source-list encoding"""
    from audiolm_pytorch_b200 import ops

    torch.manual_seed(3)
    d = 256
    tables = [torch.randn(r, d, device=DEV) for r in (1, 501, 1, 3075, 3)]
    M = 4000
    tid0 = torch.randint(0, 4, (M,), device=DEV)
    rows0 = torch.stack([torch.randint(0, tables[t].shape[0], (1,), device=DEV)[0] for t in tid0.tolist()])
    src0 = (tid0 << 24) | rows0
    src0[::37] = -1                                       # padding positions
    src1 = torch.where(torch.rand(M, device=DEV) < 0.6, (4 << 24) | torch.randint(0, 3, (M,), device=DEV),
                       torch.full((M,), -1, device=DEV))
    src = torch.stack((src0, src1), -1).to(torch.int32).contiguous()
    out = ops.embed_gather(src, tables, d)
    ref = torch.zeros(M, d, device=DEV)
    for col in (src0, src1):
        ok = col >= 0
        for t in range(5):
            sel = ok & ((col >> 24) == t)
            ref[sel] += tables[t][(col[sel] & 0xFFFFFF)]
    assert torch.equal(out, ref) or (out - ref).abs().max().item() < 1e-6
    dout = torch.randn(M, d, device=DEV)
    grads = [torch.zeros_like(t) for t in tables]
    ops.embed_scatter(src, grads, dout)
    for t in range(5):
        g_ref = torch.zeros_like(tables[t])
        for col in (src0, src1):
            sel = (col >= 0) & ((col >> 24) == t)
            g_ref.index_add_(0, col[sel] & 0xFFFFFF, dout[sel])
        assert rel_err(grads[t], g_ref) < 1e-5
