"""B200: SoundStream codec kernels (causal convs, convT, RVQ) vs goldens from the real reference + the oracle."""
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
G = Path(__file__).parent / "golden"
DEV = "cuda"


def load(name):
    return torch.load(G / name, map_location="cpu", weights_only=False)


def err(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item()


@pytest.mark.parametrize("name", ["k7", "k7d3", "k7d9", "k1", "s2", "s4", "s5", "s8", "k3"])
@pytest.mark.parametrize("mode", ["reflect", "constant"])
def test_causal_conv_golden(name, mode):
    from audiolm_pytorch_b200 import ops

    c = load("soundstream.pt")["convs"][f"{name}/{mode}"]
    y = ops.causal_conv1d(c["x"].to(DEV), c["w"].to(DEV), c["b"].to(DEV), stride=c["stride"],
                          dilation=c["dilation"], pad_mode=c["mode"])
    assert y.shape == c["y"].shape and err(y, c["y"]) < 1e-4


@pytest.mark.parametrize("s", [2, 4, 5, 8])
def test_conv_transpose_golden(s):
    from audiolm_pytorch_b200 import ops

    c = load("soundstream.pt")["convs"][f"convT{s}"]
    y = ops.causal_conv_transpose1d(c["x"].to(DEV), c["w"].to(DEV), c["b"].to(DEV), stride=s)
    assert y.shape == c["y"].shape and err(y, c["y"]) < 1e-4


def test_conv_large_vs_oracle():
    """encoder-sized layers (64 -> 128, k8 s4 and a dilated residual unit) at T = 24000 against the oracle."""
    from audiolm_pytorch_b200 import ops
    from oracle import codec as oc

    torch.manual_seed(0)
    x = torch.randn(2, 64, 6000)
    w = torch.randn(128, 64, 8) * 0.05
    b = torch.randn(128) * 0.1
    y = ops.causal_conv1d(x.to(DEV), w.to(DEV), b.to(DEV), stride=4)
    assert err(y, oc.causal_conv1d(x, w, b, stride=4)) < 2e-4
    w7, b7 = torch.randn(64, 64, 7) * 0.05, torch.randn(64) * 0.1
    w1, b1 = torch.randn(64, 64, 1) * 0.1, torch.randn(64) * 0.1
    ref = x + torch.nn.functional.elu(oc.causal_conv1d(torch.nn.functional.elu(oc.causal_conv1d(x, w7, b7, dilation=9)),
                                                       w1, b1))
    h = ops.causal_conv1d(x.to(DEV), w7.to(DEV), b7.to(DEV), dilation=9, elu=True)
    out = ops.causal_conv1d(h, w1.to(DEV), b1.to(DEV), elu=True, residual=x.to(DEV))
    assert err(out, ref) < 2e-4


def test_soundstream_golden_end_to_end():
    from audiolm_pytorch_b200.soundstream import SoundStream

    g = load("soundstream.pt")
    ss = SoundStream(**g["kwargs"])
    missing = ss.load_state_dict(g["state"], strict=True)
    ss = ss.to(DEV).eval()
    wave = g["wave"].to(DEV)
    with torch.no_grad():
        enc = ss.encoder(wave[:, None, :])
        quant, idx, _ = ss(wave, return_encoded=True)
        codes = ss.tokenize(wave)
        recon = ss(wave, return_recons_only=True)
        recon_idx = ss.decode_from_codebook_indices(idx)
    assert err(enc, g["enc"]) < 1e-4
    assert torch.equal(idx.cpu(), g["idx"]), "RVQ indices must be bit-exact"
    assert torch.equal(codes.cpu(), g["codes"])
    assert err(quant, g["quant"]) < 1e-5
    assert err(recon, g["recon"]) < 1e-4
    assert err(recon_idx, recon) < 1e-5  # README.md:100-113 round trip


@pytest.mark.parametrize("impl", ["fp32_cuda_cores", "tensor_cores"])
def test_rvq_bit_exact_at_config_size(impl):
    """C1 sizes: 8 stages x 1024 codes x 512 dims.  Rows whose best/second-best gap exceeds fp32 noise must
    match the oracle exactly; the flip rate on the rest is reported."""
    from audiolm_pytorch_b200 import ops
    from oracle import codec as oc

    torch.manual_seed(7)
    cb = torch.randn(8, 1024, 512)
    x = torch.randn(600, 512) * 3
    q_ref, i_ref = oc.rvq_encode(x, cb)
    margin = oc.rvq_margin(x, cb)
    if impl == "tensor_cores":   # distance GEMM on tcgen05 + exact fp32 re-rank of the candidates (csrc/rvq_tc.cu)
        q, i = ops.rvq_encode_tc(x.to(DEV), ops.rvq_pack_codebooks(cb.to(DEV)))
    else:
        q, i = ops.rvq_encode(x.to(DEV), cb.to(DEV))
    safe = margin > 1e-3
    assert safe.float().mean() > 0.9
    assert torch.equal(i.cpu()[safe], i_ref[safe])
    flips = (i.cpu() != i_ref).any(-1).float().mean().item()
    print(f"rvq rows differing from the oracle: {flips:.4%} (margin-safe rows: {safe.float().mean().item():.2%})")
    same = (i.cpu() == i_ref).all(-1)
    assert err(q[same.to(DEV)], q_ref[same]) < 1e-5
    dec = ops.rvq_decode(i, cb.to(DEV))
    assert err(dec, oc.rvq_decode(i.cpu(), cb)) < 1e-5
    # dropped quantizers (-1) contribute nothing
    i2 = i.clone()
    i2[:, 5:] = -1
    assert err(ops.rvq_decode(i2, cb.to(DEV)), oc.rvq_decode(i2.cpu(), cb)) < 1e-5


# ---- the kernels the C1 bench times, at C1 shapes --------------------------------------------
C1_LAYERS = [(32, 48000), (64, 24000), (128, 6000), (256, 1200)]


def _ru_weights(C, seed):
    g = torch.Generator().manual_seed(seed)
    w7 = torch.randn(C, C, 7, generator=g) * (0.7 / (7 * C) ** 0.5)
    b7 = torch.randn(C, generator=g) * 0.1
    w1 = torch.randn(C, C, 1, generator=g) * (0.7 / C ** 0.5)
    b1 = torch.randn(C, generator=g) * 0.1
    return w7, b7, w1, b1


@pytest.mark.parametrize("C,T", C1_LAYERS)
@pytest.mark.parametrize("d", [1, 3, 9])
def test_residual_unit_product_path_vs_oracle(C, T, d):
    """ResidualUnit.forward (the module path SoundStream.encoder takes: fused kernel at these widths) against the
    oracle restatement of soundstream.py:362-369, and against the two-launch path built from ops.causal_conv1d."""
    import torch.nn.functional as F

    from audiolm_pytorch_b200 import ops
    from audiolm_pytorch_b200 import soundstream as ss_mod
    from oracle import codec as oc

    w7, b7, w1, b1 = _ru_weights(C, 100 + C + d)
    x = torch.randn(2, C, T, generator=torch.Generator().manual_seed(7 + d))
    ref = x + F.elu(oc.causal_conv1d(F.elu(oc.causal_conv1d(x, w7, b7, dilation=d)), w1, b1))
    ru = ss_mod.ResidualUnit(C, C, d)
    with torch.no_grad():
        getattr(ru.fn, "0").conv.weight.copy_(w7)
        getattr(ru.fn, "0").conv.bias.copy_(b7)
        getattr(ru.fn, "2").conv.weight.copy_(w1)
        getattr(ru.fn, "2").conv.bias.copy_(b1)
    ru = ru.to(DEV).eval()
    xd = x.to(DEV)
    with torch.no_grad():
        y = ru(xd)
        h = ops.causal_conv1d(xd, w7.to(DEV), b7.to(DEV), dilation=d, elu=True)
        y2 = ops.causal_conv1d(h, w1.to(DEV), b1.to(DEV), elu=True, residual=xd)
    scale = ref.abs().max().item()
    assert err(y, ref) < 2e-4 * max(1.0, scale), (err(y, ref), scale)
    assert err(y2, ref) < 2e-4 * max(1.0, scale)
    assert err(y, y2) < 2e-4 * max(1.0, scale)
    # first samples: the reflect halo (x[1..pad] mirrored) is the edge case of the in-kernel padding
    assert err(y[..., :64], ref[..., :64]) < 2e-4 * max(1.0, scale)


@pytest.mark.parametrize("cin,cout,k,s,T", [(1, 32, 7, 1, 48000), (32, 64, 4, 2, 48000), (64, 128, 8, 4, 24000),
                                            (128, 256, 10, 5, 6000), (256, 512, 16, 8, 1200), (512, 512, 3, 1, 150)])
def test_encoder_convs_product_path_vs_oracle(cin, cout, k, s, T):
    """the non-residual convs of the C1 encoder through CausalConv1d.forward (packed-weight tiled kernels)"""
    from audiolm_pytorch_b200 import soundstream as ss_mod
    from oracle import codec as oc

    g = torch.Generator().manual_seed(cin + k)
    conv = ss_mod.CausalConv1d(cin, cout, k, stride=s)
    with torch.no_grad():
        conv.conv.weight.copy_(torch.randn(cout, cin, k, generator=g) * (0.7 / (cin * k) ** 0.5))
        conv.conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    x = torch.randn(2, cin, T, generator=g)
    ref = oc.causal_conv1d(x, conv.conv.weight.detach(), conv.conv.bias.detach(), stride=s)
    conv = conv.to(DEV).eval()
    with torch.no_grad():
        y = conv(x.to(DEV))
    assert y.shape == ref.shape and err(y, ref) < 2e-4 * max(1.0, ref.abs().max().item())


def test_c1_encoder_and_rvq_indices_vs_oracle():
    """full C1 encoder (32 channels, strides 2/4/5/8, 48 000 samples -> 150 frames x 512) + 8-stage RVQ: encoder output
    vs the oracle, code indices bit-exact on every frame whose best/second-best gap exceeds the encoder's own
    fp32-accumulation-order noise; flip rate on the rest is printed."""
    from audiolm_pytorch_b200.soundstream import SoundStream
    from oracle import codec as oc
    from oracle.transformer import sub

    torch.manual_seed(12)
    ss = SoundStream(codebook_size=1024, rq_num_quantizers=8, target_sample_hz=24000, use_local_attn=False)
    g = torch.Generator().manual_seed(7)
    for layer in ss.rq.rvqs[0].layers:
        layer._codebook.embed.copy_(torch.randn(1, 1024, 512, generator=g) * 0.05)
        layer._codebook.initted.fill_(True)
    st = {k: v.detach().clone() for k, v in ss.state_dict().items()}
    wave = torch.randn(2, 48000, generator=g)
    torch.set_num_threads(min(32, torch.get_num_threads()))
    enc_ref = oc.encoder(sub(st, "encoder"), wave[:, None, :]).transpose(1, 2)          # b n c
    cbs = oc.codebooks_of(st)
    flat = enc_ref.reshape(-1, 512)
    q_ref, i_ref = oc.rvq_encode(flat, cbs)
    margin = oc.rvq_margin(flat, cbs)
    ss = ss.to(DEV).eval()
    with torch.no_grad():
        enc = ss.encoder(wave.to(DEV)[:, None, :]).transpose(1, 2)
        quant, idx, _ = ss(wave.to(DEV), return_encoded=True)
    e = err(enc, enc_ref)
    scale = enc_ref.abs().max().item()
    print(f"C1 encoder max abs err {e:.3e} (scale {scale:.2f})")
    assert e < 2e-4 * max(1.0, scale)
    idx = idx.reshape(-1, 8).cpu()
    safe = margin > max(20 * e, 1e-4)
    print(f"margin-safe frames {safe.float().mean().item():.2%}, frames with any differing index "
          f"{(idx != i_ref).any(-1).float().mean().item():.2%}")
    assert safe.float().mean() > 0.5
    assert torch.equal(idx[safe], i_ref[safe]), "RVQ indices must be bit-exact on margin-safe frames"
    same = (idx == i_ref).all(-1)
    assert err(quant.reshape(-1, 512)[same.to(DEV)], q_ref[same]) < 1e-4


# ---- tensor-core encoder kernels (csrc/codec_tc.cu): split-bf16 implicit GEMM in the C8S layout --------------------
def test_codec_first_conv_tc():
    from audiolm_pytorch_b200 import ops
    from oracle import codec as oc

    g = torch.Generator().manual_seed(3)
    w, b = torch.randn(32, 1, 7, generator=g) * 0.3, torch.randn(32, generator=g) * 0.1
    x = torch.randn(3, 1, 5000, generator=g)
    for mode in ("reflect", "constant"):
        ref = oc.causal_conv1d(x, w, b, pad_mode=mode)
        y = ops.c8s_unpack(ops.codec_first_conv(x[:, 0].to(DEV), w.to(DEV), b.to(DEV), pad_mode=mode))
        assert err(y, ref) < 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("C,T", [(32, 5000), (64, 3000), (128, 1500), (256, 700)])
@pytest.mark.parametrize("d,phases,mode", [(1, 1, "reflect"), (3, 1, "constant"), (9, 4, "reflect"), (9, 5, "reflect")])
def test_residual_unit_tc_vs_oracle(C, T, d, phases, mode):
    """alm_codec_ru_tc (tcgen05, A operand of the 1x1 conv in tensor memory) vs soundstream.py:362-369 restated; ragged
    last tile (T % 128 != 0), reflect / constant halo, phase-split output as fed to the strided convs."""
    import torch.nn.functional as F

    from audiolm_pytorch_b200 import ops
    from oracle import codec as oc

    w7, b7, w1, b1 = _ru_weights(C, 200 + C + d)
    x = torch.randn(2, C, T, generator=torch.Generator().manual_seed(17 + d))
    ref = x + F.elu(oc.causal_conv1d(F.elu(oc.causal_conv1d(x, w7, b7, dilation=d, pad_mode=mode)), w1, b1))
    xc = ops.c8s_pack(x.to(DEV))
    wu = ops.pack_ru_weights(w7.to(DEV), w1.to(DEV))
    y = ops.codec_ru_tc(xc, wu, b7.to(DEV), b1.to(DEV), dilation=d, pad_mode=mode, out_phases=phases)
    assert y.shape == (2, 2 * C // 8, phases, T // phases, 8)
    got = ops.c8s_unpack(y)
    scale = max(1.0, ref.abs().max().item())
    e = err(got, ref)
    print(f"RU tc C={C} d={d}: max abs err {e:.2e} (scale {scale:.2f})")
    assert e < 1e-4 * scale
    assert err(got[..., :64], ref[..., :64]) < 1e-4 * scale


@pytest.mark.parametrize("cin,cout,k,s,T", [(32, 64, 4, 2, 4000), (64, 128, 8, 4, 2000), (128, 256, 10, 5, 1500),
                                            (256, 512, 16, 8, 1200), (512, 512, 3, 1, 150)])
@pytest.mark.parametrize("mode", ["reflect", "constant"])
def test_conv_tc_vs_oracle(cin, cout, k, s, T, mode):
    from audiolm_pytorch_b200 import ops
    from oracle import codec as oc

    g = torch.Generator().manual_seed(cin + k)
    w = torch.randn(cout, cin, k, generator=g) * (0.7 / (cin * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(2, cin, T, generator=g)
    ref = oc.causal_conv1d(x, w, b, stride=s, pad_mode=mode)
    xc = ops.c8s_pack(x.to(DEV), phases=s)
    wu = ops.pack_conv_weights(w.to(DEV))
    y = ops.codec_conv_tc(xc, wu, b.to(DEV), cout=cout, kernel_size=k, stride=s, pad_mode=mode)
    scale = max(1.0, ref.abs().max().item())
    assert err(ops.c8s_unpack(y), ref) < 1e-4 * scale
    y32 = ops.codec_conv_tc(xc, wu, b.to(DEV), cout=cout, kernel_size=k, stride=s, pad_mode=mode, out_fp32=True)
    assert err(y32.transpose(1, 2), ref) < 1e-4 * scale


def test_kmeans_nearest_centroid_vs_cdist():
    """HubertWithKmeans cluster assignment (hubert_kmeans.py:114-116) at HuBERT-base sizes: 768-d features, 500 clusters"""
    from audiolm_pytorch_b200 import ops

    g = torch.Generator().manual_seed(21)
    centers = torch.randn(500, 768, generator=g)
    x = torch.randn(3000, 768, generator=g) * 1.5
    d = torch.cdist(x.double(), centers.double())
    ref = (-d).argmax(dim=-1)
    top2 = d.topk(2, dim=-1, largest=False).values
    safe = (top2[:, 1] - top2[:, 0]) > 1e-4
    ids = ops.nearest_centroid(x.to(DEV), ops.rvq_pack_codebooks(centers[None].to(DEV))).cpu()
    assert safe.float().mean() > 0.99 and torch.equal(ids[safe], ref[safe])


@pytest.mark.parametrize("cin,cout,s,n", [(512, 256, 8, 150), (256, 128, 5, 1200), (128, 64, 4, 700), (64, 32, 2, 3000)])
def test_conv_transpose_tc_vs_oracle(cin, cout, s, n):
    """CausalConvTranspose1d as the 2-tap tensor-core conv with s * Cout columns scattered to s time steps"""
    from audiolm_pytorch_b200 import ops
    from oracle import codec as oc

    g = torch.Generator().manual_seed(cin + s)
    w = torch.randn(cin, cout, 2 * s, generator=g) * (0.7 / (2 * cin) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(2, cin, n, generator=g)
    ref = oc.causal_conv_transpose1d(x, w, b, s)
    xc = ops.c8s_pack(x.to(DEV))
    y = ops.codec_conv_tc(xc, ops.pack_convT_weights(w.to(DEV), s), b.to(DEV).repeat(s).contiguous(), cout=s * cout,
                          kernel_size=2, stride=1, pad_mode="constant", upsample=s)
    got = ops.c8s_unpack(y)
    assert got.shape == ref.shape and err(got, ref) < 1e-4 * max(1.0, ref.abs().max().item())


def test_decoder_edge_kernels_tc():
    """fp32 channels-last -> C8S packing and the last CausalConv1d(32, 1, 7) -> fp32 wave"""
    from audiolm_pytorch_b200 import ops
    from oracle import codec as oc

    g = torch.Generator().manual_seed(4)
    q = torch.randn(2, 150, 512, generator=g)
    packed = ops.codec_pack_c8s(q.to(DEV))
    assert err(ops.c8s_unpack(packed), q.transpose(1, 2)) < 2e-5 * q.abs().max().item()
    x = torch.randn(2, 32, 5000, generator=g)
    w, b = torch.randn(1, 32, 7, generator=g) * 0.1, torch.randn(1, generator=g)
    for mode in ("reflect", "constant"):
        ref = oc.causal_conv1d(x, w, b, pad_mode=mode)
        y = ops.codec_last_conv(ops.c8s_pack(x.to(DEV)), w.to(DEV), b.to(DEV), pad_mode=mode)
        assert y.shape == ref.shape and err(y, ref) < 5e-5 * max(1.0, ref.abs().max().item())


def test_c1_decoder_vs_oracle():
    """full C1 decoder (512 x 150 frames -> 48 000 samples) on the tensor-core path vs the oracle"""
    from audiolm_pytorch_b200.soundstream import SoundStream
    from oracle import codec as oc
    from oracle.transformer import sub

    torch.manual_seed(13)
    ss = SoundStream(codebook_size=1024, rq_num_quantizers=8, target_sample_hz=24000, use_local_attn=False)
    st = {k: v.detach().clone() for k, v in ss.state_dict().items()}
    q = torch.randn(2, 150, 512) * 0.5
    torch.set_num_threads(min(32, torch.get_num_threads()))
    ref = oc.decoder(sub(st, "decoder"), q.transpose(1, 2))
    ss = ss.to(DEV).eval()
    assert ss._tc_plan_dec() is not None
    with torch.no_grad():
        got = ss.decode(q.to(DEV))
    e, scale = err(got, ref), ref.abs().max().item()
    print(f"C1 decoder max abs err {e:.3e} (scale {scale:.3f})")
    assert got.shape == ref.shape == (2, 1, 48000) and e < 2e-4 * max(1.0, scale)
