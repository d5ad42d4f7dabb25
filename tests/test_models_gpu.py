"""B200: the product transformers (CUDA kernels through the C ABI) vs goldens produced by the real reference."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"
DEV = "cuda"


def load(name):
    return torch.load(G / name, map_location="cpu", weights_only=False)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


TOL = 3e-2   # bf16 activations / fp32 accumulation vs the fp32 reference ("within 1e-2 rel" is checked as RMS below)


def rms_rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp(min=1e-12)).item()


def check_grads(named, golden_grads, noise=None):
    """every parameter gradient vs the reference's fp32 gradient.  Tensors: RMS-relative error < max(7e-2, 2 x the
    deviation the REFERENCE itself shows under bf16 autocast on the same problem) — `noise` is the golden's
    `bf16_noise` dict (oracle/make_golden.py::bf16_noise); the CUDA path computes in bf16 like an autocast step, so
    its distance to the fp32 gradients is bounded by the reference's own bf16 spread, not by fp32 round-off.  The tiny hyper-connection
    tensors (static_alpha/beta [4,5]/[4], the two scalars) are sums over every token of strongly cancelling
    bf16-noisy terms, so their error is measured against the largest gradient of the same kind across layers."""
    kind_scale = {}
    for k, gr in golden_grads.items():
        if gr.numel() <= 20:
            kind = k.split(".")[-1]
            kind_scale[kind] = max(kind_scale.get(kind, 0.0), gr.float().pow(2).mean().sqrt().item())
    errs = {}
    for k, gr in golden_grads.items():
        assert named[k].grad is not None, k
        if gr.numel() <= 20:
            err = (named[k].grad.float().cpu() - gr.float()).pow(2).mean().sqrt().item()
            errs[k] = (err / kind_scale[k.split(".")[-1]], 0.35)  # 40-68 tokens only; tighter check: *_more_tokens
        else:
            errs[k] = (rms_rel(named[k].grad, gr), max(7e-2, 2.0 * (noise or {}).get(k, 0.0)))
    for k, (e, tol) in sorted(errs.items(), key=lambda kv: -kv[1][0] / kv[1][1])[:5]:
        print(f"  grad err {e:.4f} (tol {tol}) {k}")
    bad = {k: v for k, v in errs.items() if v[0] >= v[1]}
    assert not bad, bad
    return max(v[0] for v in errs.values())


def build(cls, g):
    m = cls(**g["kwargs"])
    m.load_state_dict(g["state"], strict=True)
    return m.to(DEV)


def test_semantic_forward_cache_loss_grads():
    from audiolm_pytorch_b200.audiolm import SemanticTransformer, SemanticTransformerWrapper

    g = load("semantic.pt")
    m = build(SemanticTransformer, g).eval()
    ids = g["ids"].to(DEV)
    with torch.no_grad():
        logits = m(ids=ids)
        masked = m(ids=ids, self_attn_mask=g["mask"].to(DEV))
        l12, cache = m(ids=ids[:, :12], return_kv_cache=True)
        inc, _ = m(ids=ids, kv_cache=cache, return_kv_cache=True)
    assert rms_rel(logits, g["logits"]) < 1e-2 and rel(logits, g["logits"]) < TOL
    assert rms_rel(masked, g["logits_masked"]) < 1e-2
    assert rel(cache, g["cache12"]) < TOL
    assert rms_rel(inc, g["logits_inc"]) < 1.5e-2
    w = SemanticTransformerWrapper(transformer=m, unique_consecutive=False, mask_prob=0.0).train()
    loss = w(semantic_token_ids=ids, return_loss=True)
    assert abs(loss.item() - g["loss"].item()) < 2e-2 * abs(g["loss"].item())
    loss.backward()
    worst = check_grads(dict(m.named_parameters()), g["grads"], g.get("bf16_noise"))
    print("semantic worst grad rms rel err", worst)


def test_coarse_forward_cache_loss_grads():
    from audiolm_pytorch_b200.audiolm import CoarseTransformer, CoarseTransformerWrapper

    g = load("coarse.pt")
    m = build(CoarseTransformer, g).eval()
    sem, coarse = g["sem"].to(DEV), g["coarse"].to(DEV)
    with torch.no_grad():
        sl, cl = m(semantic_token_ids=sem, coarse_token_ids=coarse)
        slm, clm = m(semantic_token_ids=sem, coarse_token_ids=coarse, self_attn_mask=g["mask"].to(DEV))
        (_, _), (kv_a, emb_a) = m(semantic_token_ids=sem, coarse_token_ids=coarse[:, :9], return_cache=True,
                                   return_only_coarse_logits=True)
        (_, cb), _ = m(semantic_token_ids=sem, coarse_token_ids=coarse[:, :10], return_cache=True, kv_cache=kv_a,
                       embed_cache=emb_a, return_only_coarse_logits=True)
    print("coarse flash logits err", rms_rel(sl, g["sem_logits"]), rms_rel(cl, g["coarse_logits"]),
          "reference under bf16 autocast:", g["logits_bf16_noise"])
    # 1e-2 RMS-relative (north star), or the reference's OWN bf16-autocast deviation on this fixture if that is larger
    tol_s, tol_c = (max(1e-2, n_) for n_ in g["logits_bf16_noise"])
    assert rms_rel(sl, g["sem_logits"]) < tol_s and rms_rel(cl, g["coarse_logits"]) < tol_c
    assert rms_rel(slm, g["sem_logits_masked"]) < 1e-2 and rms_rel(clm, g["coarse_logits_masked"]) < 1e-2
    assert rel(kv_a, g["kv_a"]) < TOL and rms_rel(emb_a, g["emb_a"]) < 1e-2
    assert rms_rel(cb, g["coarse_logits_b"]) < 1.5e-2

    class _Codec:  # the wrapper constructor only reads rq_groups
        rq_groups = 1

    w = CoarseTransformerWrapper(transformer=m, codec=_Codec(), unique_consecutive=False, mask_prob=0.0).train()
    loss = w(semantic_token_ids=sem, coarse_token_ids=coarse[:, :21], return_loss=True)
    assert abs(loss.item() - g["loss"].item()) < 2e-2 * abs(g["loss"].item())
    loss.backward()
    worst = check_grads(dict(m.named_parameters()), g["grads"], g.get("bf16_noise"))
    print("coarse worst grad rms rel err", worst)


def test_fine_forward():
    from audiolm_pytorch_b200.audiolm import FineTransformer

    g = load("fine.pt")
    m = build(FineTransformer, g).eval()
    with torch.no_grad():
        cl, fl = m(coarse_token_ids=g["coarse"].to(DEV), fine_token_ids=g["fine"].to(DEV))
    assert rms_rel(cl, g["coarse_logits"]) < 1e-2 and rms_rel(fl, g["fine_logits"]) < 1e-2


def test_matches_oracle_at_larger_size():
    """seeded random weights, seq 384 / dim 256 / depth 3: CUDA path vs the CPU oracle restatement."""
    from audiolm_pytorch_b200.audiolm import CoarseTransformer
    from oracle import transformer as ot

    torch.manual_seed(5)
    kw = dict(num_semantic_tokens=100, codebook_size=128, num_coarse_quantizers=3, dim=256, depth=3, heads=4,
              flash_attn=True)
    m = CoarseTransformer(**kw)
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "dynamic_alpha_fn" in n_ or "dynamic_beta_fn" in n_:
                p.normal_(0, 0.02)
            if "logit_weights" in n_:
                p.mul_(0.1)
    st = {k: v.detach().clone() for k, v in m.state_dict().items()}
    sem = torch.randint(0, 100, (2, 90))
    coarse = torch.randint(0, 128, (2, 292))
    (osl, ocl), _ = ot.coarse_forward(st, sem, coarse, heads=4, depth=3, codebook_size=128, num_coarse_quantizers=3)
    m = m.to(DEV).eval()
    with torch.no_grad():
        sl, cl = m(semantic_token_ids=sem.to(DEV), coarse_token_ids=coarse.to(DEV))
    assert rms_rel(sl, osl) < 1e-2 and rms_rel(cl, ocl) < 1e-2


def test_sampling_helpers_bit_exact():
    from audiolm_pytorch_b200 import heads

    g = load("sampling.pt")
    filt = heads.top_k(g["logits"].to(DEV), thres=0.9)
    assert torch.equal(filt.cpu(), g["filtered"])
    noise = g["uniform"].to(DEV)
    gum = -torch.log(-torch.log(noise + 1e-20) + 1e-20)
    assert torch.equal((filt + gum).argmax(-1).cpu(), g["ids"])
    assert torch.equal(heads.mask_out_after_eos_id(g["seq"].to(DEV), 64, keep_eos=False).cpu(), g["seq_masked"])


def test_fused_sampler_kernel_bit_exact():
    """alm_topk_gumbel_sample vs the reference's top_k + gumbel_sample under identical uniform noise."""
    from audiolm_pytorch_b200 import ops

    g = load("sampling.pt")
    ids = ops.topk_gumbel_sample(g["logits"].to(DEV), g["uniform"].to(DEV), k=max(int(0.1 * 65), 1))
    assert torch.equal(ids.cpu(), g["ids"])
    torch.manual_seed(9)
    for V in (501, 1025, 1024):
        logits = torch.randn(64, V, device=DEV) * 4
        u = torch.rand(64, V, device=DEV)
        k = max(int(0.1 * V), 1)
        val, ind = torch.topk(logits, k)
        filt = torch.full_like(logits, float("-inf")).scatter_(1, ind, val)
        ref = (filt / 0.8 + (-torch.log(-torch.log(u + 1e-20) + 1e-20))).argmax(-1)
        got = ops.topk_gumbel_sample(logits, u, k=k, temperature=0.8)
        assert torch.equal(got, ref)


@pytest.mark.parametrize("V,ties", [(2049, False), (4097, False), (32000, False), (5000, True)])
def test_fused_sampler_large_vocab(V, ties):
    """V > 2048 takes the radix-select threshold path of alm_topk_gumbel_sample; ties at the threshold keep the lowest
    indices (stable descending sort), mixed-sign logits and -inf entries included."""
    from audiolm_pytorch_b200 import ops

    torch.manual_seed(V)
    logits = torch.randn(16, V, device=DEV) * 4
    if ties:
        logits = (logits * 2).round() / 2
    logits[:, 7] = float("-inf")
    u = torch.rand(16, V, device=DEV)
    k = max(int(0.1 * V), 1)
    order = torch.sort(logits, dim=-1, descending=True, stable=True).indices[:, :k]
    filt = torch.full_like(logits, float("-inf")).scatter_(1, order, logits.gather(1, order))
    ref = (filt / 0.9 + (-torch.log(-torch.log(u + 1e-20) + 1e-20))).argmax(-1)
    got = ops.topk_gumbel_sample(logits, u, k=k, temperature=0.9)
    assert torch.equal(got, ref)


def test_generate_paths_end_to_end():
    """Semantic/Coarse/Fine .generate() with KV cache + codec decode (AudioLM.forward plumbing, tiny models).
    Also: KV-cache generation == no-cache generation under the same noise (teacher-forcing free check)."""
    from audiolm_pytorch_b200.audiolm import (AudioLM, CoarseTransformer, CoarseTransformerWrapper, FineTransformer,
                                              SemanticTransformer)
    from audiolm_pytorch_b200.soundstream import SoundStream

    torch.manual_seed(3)
    kw = dict(dim=64, depth=2, heads=2, flash_attn=True)
    sem = SemanticTransformer(num_semantic_tokens=50, **kw).to(DEV)
    coarse = CoarseTransformer(num_semantic_tokens=50, codebook_size=64, num_coarse_quantizers=2, **kw).to(DEV)
    fine = FineTransformer(num_coarse_quantizers=2, num_fine_quantizers=2, codebook_size=64, **kw).to(DEV)
    codec = SoundStream(codebook_size=64, rq_num_quantizers=4, channels=4, codebook_dim=32, use_local_attn=False)
    for layer in codec.rq.rvqs[0].layers:
        layer._codebook.embed.normal_(0, 0.5)
        layer._codebook.initted.fill_(True)
    codec = codec.to(DEV).eval()
    cw = CoarseTransformerWrapper(transformer=coarse, codec=codec, unique_consecutive=False)
    sem_ids = torch.randint(0, 50, (2, 12), device=DEV)
    from audiolm_pytorch_b200 import audiolm
    audiolm.USE_DECODE_GRAPHS = False   # eager engine steps draw the sampling noise in the same order as the slow path
    try:
        torch.manual_seed(5)
        a = cw.generate(semantic_token_ids=sem_ids, max_time_steps=6, use_kv_cache=True)
    finally:
        audiolm.USE_DECODE_GRAPHS = True
        cw._engine = None
    torch.manual_seed(5)
    b = cw.generate(semantic_token_ids=sem_ids, max_time_steps=6, use_kv_cache=False)
    assert a.shape == (2, 6, 2) and a.max() <= 64
    assert (a == b).float().mean() > 0.9  # identical noise; bf16 logits may flip a rare near-tie
    lm = AudioLM(wav2vec=None, codec=codec, semantic_transformer=sem, coarse_transformer=coarse, fine_transformer=fine)
    lm.coarse.generate.__func__  # noqa: B018  (bound method exists)
    # shorten the hard-coded 512 coarse steps for the test by calling the three stages directly
    s_ids = lm.semantic.generate(batch_size=1, max_length=10)
    s_ids = s_ids.clamp(min=0)
    c_ids = lm.coarse.generate(semantic_token_ids=s_ids, max_time_steps=8)
    c_ids = c_ids.clamp(min=0)
    wav = lm.fine.generate(coarse_token_ids=c_ids, reconstruct_wave=True)
    wav = wav if torch.is_tensor(wav) else wav[0]
    assert torch.isfinite(wav).all() and wav.shape[-1] == 8 * 320


def test_grads_vs_oracle_autograd_more_tokens():
    """1024 tokens (b4 x n256), dim 128, depth 2: every gradient of the hand-written backward vs torch autograd of the
    oracle restatement (run in fp32 on the GPU).  Tolerances reflect bf16 activations vs an fp32 oracle: 7 % RMS for
    tensors; 30 % of the largest same-kind gradient for the 4..20-element hyper-connection tensors, whose entries
    (~1e-3) are sums over all tokens of strongly cancelling terms (measured 19 % on the worst one, same sign/size)."""
    from audiolm_pytorch_b200.audiolm import SemanticTransformer
    from audiolm_pytorch_b200.heads import cross_entropy
    from oracle import transformer as ot

    torch.manual_seed(17)
    m = SemanticTransformer(num_semantic_tokens=100, dim=128, depth=2, heads=2, flash_attn=True)
    with torch.no_grad():
        for n_, p in m.named_parameters():
            if "dynamic_alpha_fn" in n_ or "dynamic_beta_fn" in n_:
                p.normal_(0, 0.05)
            if "dynamic_alpha_scale" in n_ or "dynamic_beta_scale" in n_:
                p.fill_(0.3)
    m = m.to(DEV).train()
    ids = torch.randint(0, 100, (4, 255), device=DEV)
    labels = torch.cat((ids, torch.full((4, 1), 100, device=DEV)), 1)
    st = {k: v.detach().clone().float().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}
    ol, _ = ot.semantic_forward(st, ids, heads=2, depth=2)
    oloss = ot.cross_entropy(ol, labels)
    oloss.backward()
    logits = m(ids=ids)
    loss = cross_entropy(logits, labels)
    loss.backward()
    assert abs(loss.item() - oloss.item()) < 1e-2 * abs(oloss.item())
    golden = {}
    for k, p in m.named_parameters():
        if p.grad is None:
            continue
        g = st[k].grad
        # grad_shrink (audiolm_pytorch.py:93-94): what reaches the embeddings is scaled by 0.1
        golden[k] = (g * 0.1 if k in ("start_token", "semantic_embedding.weight") else g).cpu()
    kind_scale = {}
    for k, gr in golden.items():
        if gr.numel() <= 20:
            kind = k.split(".")[-1]
            kind_scale[kind] = max(kind_scale.get(kind, 0.0), gr.pow(2).mean().sqrt().item())
    named = dict(m.named_parameters())
    errs = {}
    for k, gr in golden.items():
        got = named[k].grad.float().cpu()
        if gr.numel() <= 20:
            errs[k] = ((got - gr).pow(2).mean().sqrt().item() / kind_scale[k.split(".")[-1]], 0.30)
        else:
            errs[k] = (rms_rel(got, gr), 7e-2)
    for k, (e, tol) in sorted(errs.items(), key=lambda kv: -kv[1][0] / kv[1][1])[:12]:
        print(f"  grad err {e:.4f} (tol {tol}) {k}  ref={golden[k].flatten()[:4].tolist()} got={named[k].grad.flatten()[:4].tolist()}")
    bad = {k: v for k, v in errs.items() if v[0] >= v[1]}
    assert not bad, bad


def test_single_residual_stream_path():
    """num_residual_streams=1 (plain Residual wrappers): logits, loss and every gradient vs the reference golden."""
    from audiolm_pytorch_b200.audiolm import SemanticTransformer, SemanticTransformerWrapper

    g = load("semantic_plain.pt")
    m = build(SemanticTransformer, g).eval()
    ids = g["ids"].to(DEV)
    with torch.no_grad():
        logits = m(ids=ids)
        masked = m(ids=ids, self_attn_mask=g["mask"].to(DEV))
        l12, cache = m(ids=ids[:, :12], return_kv_cache=True)
        inc, _ = m(ids=ids, kv_cache=cache, return_kv_cache=True)
    assert rms_rel(logits, g["logits"]) < 1e-2 and rms_rel(masked, g["logits_masked"]) < 1e-2
    assert rms_rel(inc, logits[:, 13:]) < 1.5e-2
    w = SemanticTransformerWrapper(transformer=m, unique_consecutive=False, mask_prob=0.0).train()
    loss = w(semantic_token_ids=ids, return_loss=True)
    assert abs(loss.item() - g["loss"].item()) < 2e-2 * abs(g["loss"].item())
    loss.backward()
    check_grads(dict(m.named_parameters()), g["grads"], g.get("bf16_noise"))
