"""SoundStream codec inference path on libalm_b200 (sm_100a): causal conv encoder -> residual VQ -> decoder.

Drop-in surface of /root/reference/audiolm_pytorch/soundstream.py:314-395, 451-866 for the calls the AudioLM
hot path makes: `forward(x, return_encoded=True | return_codes_only=True | return_recons_only=True)`,
`tokenize`, `decode_from_codebook_indices`, `decode`, with the reference's constructor kwargs and
state_dict keys for `encoder.*`, `decoder.*`, `rq.*`.  GAN / mel training losses, LFQ / FSQ quantizers and
FiLM denoising are outside this build; the local-attention bottleneck lives in local_attn.py.
"""
from __future__ import annotations

import functools
import pickle
from itertools import cycle
from pathlib import Path

import torch
from torch import nn

from . import ops
from .local_attn import LocalTransformer

f32 = torch.float32


FUSE_RESIDUAL_UNITS = True  # False: conv7 + conv1 as two launches (A/B tests)
RVQ_ON_TENSOR_CORES = True      # False: fp32 CUDA-core search kernel (A/B tests)
ENCODER_ON_TENSOR_CORES = True  # False: fp32 CUDA-core conv kernels for the whole encoder (A/B tests)


def exists(v):
    return v is not None


class CausalConv1d(nn.Module):
    """soundstream.py:332-345; parameters live in `.conv` (nn.Conv1d) for state_dict compatibility."""

    def __init__(self, chan_in, chan_out, kernel_size, pad_mode="reflect", **kwargs):
        super().__init__()
        self.dilation = kwargs.get("dilation", 1)
        self.stride = kwargs.get("stride", 1)
        self.pad_mode = pad_mode
        self.causal_padding = self.dilation * (kernel_size - 1) + (1 - self.stride)
        self.conv = nn.Conv1d(chan_in, chan_out, kernel_size, **kwargs)
        self._packed = (None, None)  # (weight version key, [Cin, K, Cout] copy for the register-tiled kernel)

    def _packed_weight(self):
        w = self.conv.weight
        key = (w.data_ptr(), w._version, w.device)
        if self._packed[0] != key:
            with torch.no_grad():
                self._packed = (key, w.detach().permute(1, 2, 0).contiguous())
        return self._packed[1]

    def forward(self, x, elu=False, residual=None):
        k = self.conv.kernel_size[0]
        wp = self._packed_weight() if (k, self.stride, self.dilation) in ops.CONV_TILED_SHAPES else None
        return ops.causal_conv1d(x, self.conv.weight, self.conv.bias, stride=self.stride, dilation=self.dilation,
                                 pad_mode=self.pad_mode, elu=elu, residual=residual, weight_packed=wp)


class CausalConvTranspose1d(nn.Module):
    """soundstream.py:347-360."""

    def __init__(self, chan_in, chan_out, kernel_size, stride, **kwargs):
        super().__init__()
        assert kernel_size == 2 * stride
        self.upsample_factor = stride
        self.conv = nn.ConvTranspose1d(chan_in, chan_out, kernel_size, stride, **kwargs)

    def forward(self, x):
        return ops.causal_conv_transpose1d(x, self.conv.weight, self.conv.bias, stride=self.upsample_factor)


class _RUBody(nn.Module):
    """holds the two convs under the reference's Sequential indices 0 and 2 (1, 3 are ELUs)."""

    def __init__(self, chan_in, chan_out, dilation, kernel_size, pad_mode):
        super().__init__()
        self.add_module("0", CausalConv1d(chan_in, chan_out, kernel_size, dilation=dilation, pad_mode=pad_mode))
        self.add_module("2", CausalConv1d(chan_out, chan_out, 1, pad_mode=pad_mode))


class ResidualUnit(nn.Module):
    """x + ELU(conv1(ELU(conv7_dil(x)))) (soundstream.py:362-369) in two fused launches; keys `fn.{0,2}.conv.*`."""

    def __init__(self, chan_in, chan_out, dilation, kernel_size=7, squeeze_excite=False, pad_mode="reflect"):
        super().__init__()
        if squeeze_excite:
            raise NotImplementedError("squeeze_excite=True is not built")
        self.fn = _RUBody(chan_in, chan_out, dilation, kernel_size, pad_mode)

    def forward(self, x):
        c7, c1 = getattr(self.fn, "0"), getattr(self.fn, "2")
        C = x.shape[1]
        if (FUSE_RESIDUAL_UNITS and C in ops.RU_FUSED_CHANNELS and c7.dilation in ops.RU_FUSED_DILATIONS
                and c7.conv.kernel_size[0] == 7 and c1.conv.kernel_size[0] == 1 and c7.conv.out_channels == C
                and x.shape[-1] > 6 * c7.dilation):
            return ops.residual_unit(x, c7._packed_weight(), c7.conv.bias, c1._packed_weight(), c1.conv.bias,
                                     dilation=c7.dilation, pad_mode=c7.pad_mode)
        h = c7(x, elu=True)
        return c1(h, elu=True, residual=x)


def EncoderBlock(chan_in, chan_out, stride, cycle_dilations=(1, 3, 9), squeeze_excite=False, pad_mode="reflect"):
    it = cycle(cycle_dilations)
    return nn.Sequential(*[ResidualUnit(chan_in, chan_in, next(it), squeeze_excite=squeeze_excite, pad_mode=pad_mode)
                           for _ in range(3)],
                         CausalConv1d(chan_in, chan_out, 2 * stride, stride=stride))


def DecoderBlock(chan_in, chan_out, stride, cycle_dilations=(1, 3, 9), squeeze_excite=False, pad_mode="reflect"):
    it = cycle(cycle_dilations)
    return nn.Sequential(CausalConvTranspose1d(chan_in, chan_out, 2 * stride, stride=stride),
                         *[ResidualUnit(chan_out, chan_out, next(it), squeeze_excite=squeeze_excite, pad_mode=pad_mode)
                           for _ in range(3)])


# ---- residual VQ containers (state_dict keys of vector-quantize-pytorch) ---------------------------
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

    def codebooks(self):
        return torch.stack([l._codebook.embed[0] for l in self.layers]).to(f32)

    def forward(self, x):
        if self.training:
            raise NotImplementedError("RVQ training (EMA / k-means / commitment loss) is outside this build")
        if not all(bool(l._codebook.initted.item()) for l in self.layers):
            raise RuntimeError("codebooks are not initialised (load a checkpoint; k-means init is not built)")
        b, n, d = x.shape
        flat = x.reshape(b * n, d).to(f32).contiguous()
        if RVQ_ON_TENSOR_CORES and d % 8 == 0:
            embeds = [l._codebook.embed for l in self.layers]
            key = tuple((e.data_ptr(), e._version) for e in embeds)
            if self.__dict__.get("_tc_key") != key:
                with torch.inference_mode(False), torch.no_grad():
                    self.__dict__["_tc_pack"] = ops.rvq_pack_codebooks(self.codebooks())
                self.__dict__["_tc_key"] = key
            quant, idx = ops.rvq_encode_tc(flat, self.__dict__["_tc_pack"])
        else:
            quant, idx = ops.rvq_encode(flat, self.codebooks())
        return quant.view(b, n, d), idx.view(b, n, -1), torch.zeros(1, len(self.layers), device=x.device)

    def get_output_from_indices(self, indices):
        b, n, q = indices.shape
        return ops.rvq_decode(indices.reshape(b * n, q), self.codebooks()).view(b, n, -1)


class GroupedResidualVQ(nn.Module):
    """channels split into `groups`, one ResidualVQ each (soundstream.py:592-607)."""

    def __init__(self, *, dim, groups=1, **kwargs):
        super().__init__()
        assert dim % groups == 0
        self.groups = groups
        self.rvqs = nn.ModuleList([ResidualVQ(dim=dim // groups, **kwargs) for _ in range(groups)])

    def forward(self, x):
        outs = [rvq(c) for rvq, c in zip(self.rvqs, x.chunk(self.groups, dim=-1))]
        return (torch.cat([o[0] for o in outs], dim=-1), torch.stack([o[1] for o in outs]),
                torch.stack([o[2] for o in outs]))

    def get_output_from_indices(self, indices):  # g b n q
        return torch.cat([rvq.get_output_from_indices(i) for rvq, i in zip(self.rvqs, indices)], dim=-1)


class SoundStream(nn.Module):
    """soundstream.py:451-866 (inference path)."""

    def __init__(self, *, channels=32, strides=(2, 4, 5, 8), channel_mults=(2, 4, 8, 16), codebook_dim=512,
                 codebook_size=None, finite_scalar_quantizer_levels=None, rq_num_quantizers=8,
                 rq_commitment_weight=1.0, rq_ema_decay=0.95, rq_quantize_dropout_multiple_of=1, rq_groups=1,
                 rq_stochastic_sample_codes=False, rq_rotation_trick=True, rq_kwargs: dict = {},
                 use_lookup_free_quantizer=False, use_finite_scalar_quantizer=False, input_channels=1,
                 discr_multi_scales=(1, 0.5, 0.25), stft_normalized=False, enc_cycle_dilations=(1, 3, 9),
                 dec_cycle_dilations=(1, 3, 9), multi_spectral_window_powers_of_two=tuple(range(6, 12)),
                 multi_spectral_n_ffts=512, multi_spectral_n_mels=64, recon_loss_weight=1.0,
                 multi_spectral_recon_loss_weight=1e-5, adversarial_loss_weight=1.0, feature_loss_weight=100,
                 quantize_dropout_cutoff_index=1, target_sample_hz=16000, use_local_attn=True, attn_window_size=128,
                 attn_dim_head=64, attn_heads=8, attn_depth=1, attn_xpos_scale_base=None,
                 attn_dynamic_pos_bias=False, use_gate_loop_layers=False, squeeze_excite=False,
                 complex_stft_discr_logits_abs=True, pad_mode="reflect", stft_discriminator=None,
                 complex_stft_discr_kwargs: dict = dict()):
        super().__init__()
        cfg = dict(locals())
        cfg.pop("self", None)
        cfg.pop("__class__", None)
        self._configs = pickle.dumps(cfg)
        if use_lookup_free_quantizer or use_finite_scalar_quantizer or use_gate_loop_layers:
            raise NotImplementedError("LFQ / FSQ / gate-loop variants are outside this build")
        assert exists(codebook_size)
        self.target_sample_hz = target_sample_hz
        self.single_channel = input_channels == 1
        self.strides = strides
        layer_channels = (channels, *[m * channels for m in channel_mults])
        pairs = tuple(zip(layer_channels[:-1], layer_channels[1:]))
        self.encoder = nn.Sequential(
            CausalConv1d(input_channels, channels, 7, pad_mode=pad_mode),
            *[EncoderBlock(ci, co, s, enc_cycle_dilations, squeeze_excite, pad_mode) for (ci, co), s in zip(pairs, strides)],
            CausalConv1d(layer_channels[-1], codebook_dim, 3, pad_mode=pad_mode))
        attn_kwargs = dict(dim=codebook_dim, dim_head=attn_dim_head, heads=attn_heads, depth=attn_depth,
                           window_size=attn_window_size, xpos_scale_base=attn_xpos_scale_base,
                           dynamic_pos_bias=attn_dynamic_pos_bias, prenorm=True, causal=True)
        # windowed causal attention bottleneck on both sides of the quantizer (soundstream.py:533-545, 613)
        self.encoder_attn = LocalTransformer(**attn_kwargs) if use_local_attn else None
        self.decoder_attn = LocalTransformer(**attn_kwargs) if use_local_attn else None
        self.num_quantizers = rq_num_quantizers
        self.codebook_dim = codebook_dim
        self.codebook_size = codebook_size
        self.rq_groups = rq_groups
        self.use_lookup_free_quantizer = False
        self.use_finite_scalar_quantizer = False
        self.rq = GroupedResidualVQ(dim=codebook_dim, num_quantizers=rq_num_quantizers, codebook_size=codebook_size,
                                    groups=rq_groups)
        self.decoder = nn.Sequential(
            CausalConv1d(codebook_dim, layer_channels[-1], 7, pad_mode=pad_mode),
            *[DecoderBlock(co, ci, s, dec_cycle_dilations, squeeze_excite, pad_mode)
              for (ci, co), s in zip(reversed(pairs), reversed(strides))],
            CausalConv1d(channels, input_channels, 7, pad_mode=pad_mode))
        self.register_buffer("zero", torch.tensor(0.0), persistent=False)

    # ---- bookkeeping -----------------------------------------------------------------------------
    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def configs(self):
        return pickle.loads(self._configs)

    @property
    def seq_len_multiple_of(self):
        return functools.reduce(lambda a, b: a * b, self.strides)

    @property
    def downsample_factor(self):
        return self.seq_len_multiple_of

    def save(self, path):
        torch.save(dict(model=self.state_dict(), config=self._configs, version="2.4.0"), str(Path(path)))

    def load(self, path, strict=False):
        """loads encoder / decoder / rq weights; keys of sub-modules outside this build (discriminators,
        FiLM, mel transforms, local attention) are ignored unless strict=True."""
        pkg = torch.load(str(Path(path)), map_location="cpu", weights_only=False)
        sd = pkg["ema_model"] if "ema_model" in pkg else pkg["model"]
        if "ema_model" in pkg:
            sd = {k[len("ema_model."):]: v for k, v in sd.items() if k.startswith("ema_model.")}
        if not strict:
            mine = self.state_dict()
            sd = {k: v for k, v in sd.items() if k in mine}
        self.load_state_dict(sd, strict=strict)

    @classmethod
    def init_and_load_from(cls, path, strict=False):
        pkg = torch.load(str(Path(path)), map_location="cpu", weights_only=False)
        assert "config" in pkg, "model configs were not found in this saved checkpoint"
        m = cls(**pickle.loads(pkg["config"]))
        m.load(path, strict=strict)
        return m.eval()

    # ---- encoder on the tensor cores (csrc/codec_tc.cu) ------------------------------------------------
    def _tc_plan(self):
        """layer list for the split-bf16 tensor-core encoder, or None when this configuration is outside what those
        kernels are built for (then the fp32 CUDA-core kernels run).  Structure follows soundstream.py:519-531."""
        if not (ENCODER_ON_TENSOR_CORES and self.single_channel):
            return None
        enc = list(self.encoder)
        first, blocks, last = enc[0], enc[1:-1], enc[-1]
        ok = (isinstance(first, CausalConv1d) and first.conv.kernel_size[0] <= 8 and first.conv.out_channels in (32, 64)
              and first.stride == 1 and first.dilation == 1 and isinstance(last, CausalConv1d) and last.dilation == 1
              and last.conv.in_channels % 16 == 0 and last.conv.out_channels % 64 == 0)
        plan = []
        for blk in blocks if ok else ():
            *rus, down = list(blk)
            for ru in rus:
                c7, c1 = getattr(ru.fn, "0"), getattr(ru.fn, "2")
                C = c7.conv.in_channels
                ok = ok and (C in (32, 64, 128, 256) and c7.conv.out_channels == C and c7.conv.kernel_size[0] == 7
                             and c1.conv.kernel_size[0] == 1 and 6 * c7.dilation <= 54 and c7.stride == 1)
            ok = ok and down.dilation == 1 and down.conv.in_channels % 16 == 0 and down.conv.out_channels % 64 == 0
            plan.append((rus, down))
        return (first, plan, last) if ok else None

    @staticmethod
    def _cached(mod, name, params, build):
        key = tuple((p_.data_ptr(), p_._version) for p_ in params)
        hit = mod.__dict__.get(name)
        if hit is None or hit[0] != key:
            with torch.inference_mode(False), torch.no_grad():
                hit = (key, build())
            mod.__dict__[name] = hit
        return hit[1]

    def _encode_tc(self, wave, plan):
        """wave fp32 [B, T] -> encoder output fp32 [B, n, codebook_dim] (channels-last, what the RVQ consumes).
        Activations stay in the C8S split-bf16 layout between layers; every layer is one kernel launch."""
        first, blocks, last = plan
        h = ops.codec_first_conv(wave, first.conv.weight, first.conv.bias, pad_mode=first.pad_mode)
        for rus, down in blocks:
            for i, ru in enumerate(rus):
                c7, c1 = getattr(ru.fn, "0"), getattr(ru.fn, "2")
                wu = self._cached(ru, "_tc_units", [c7.conv.weight, c1.conv.weight],
                                  lambda c7=c7, c1=c1: ops.pack_ru_weights(c7.conv.weight, c1.conv.weight))
                h = ops.codec_ru_tc(h, wu, c7.conv.bias, c1.conv.bias, dilation=c7.dilation, pad_mode=c7.pad_mode,
                                    out_phases=down.stride if i == len(rus) - 1 else 1)
            if not rus:
                raise NotImplementedError  # (guarded by _tc_plan: every block has residual units)
            wu = self._cached(down, "_tc_units", [down.conv.weight], lambda d=down: ops.pack_conv_weights(d.conv.weight))
            h = ops.codec_conv_tc(h, wu, down.conv.bias, cout=down.conv.out_channels,
                                  kernel_size=down.conv.kernel_size[0], stride=down.stride, pad_mode=down.pad_mode)
        wu = self._cached(last, "_tc_units", [last.conv.weight], lambda: ops.pack_conv_weights(last.conv.weight))
        return ops.codec_conv_tc(h, wu, last.conv.bias, cout=last.conv.out_channels,
                                 kernel_size=last.conv.kernel_size[0], stride=1, pad_mode=last.pad_mode, out_fp32=True)

    def _tc_plan_dec(self):
        """layer list for the tensor-core decoder (soundstream.py:615-627) or None"""
        if not (ENCODER_ON_TENSOR_CORES and self.single_channel):
            return None
        dec = list(self.decoder)
        first, blocks, last = dec[0], dec[1:-1], dec[-1]
        ok = (isinstance(first, CausalConv1d) and first.stride == 1 and first.dilation == 1
              and first.conv.in_channels % 16 == 0 and first.conv.out_channels % 64 == 0
              and isinstance(last, CausalConv1d) and last.conv.out_channels == 1 and last.conv.in_channels in (32, 64)
              and last.conv.kernel_size[0] <= 8 and last.stride == 1 and last.dilation == 1)
        plan = []
        for blk in blocks if ok else ():
            up, *rus = list(blk)
            ok = ok and isinstance(up, CausalConvTranspose1d) and up.conv.in_channels % 16 == 0 \
                and up.conv.out_channels % 16 == 0 and (up.upsample_factor * up.conv.out_channels) % 64 == 0
            for ru in rus:
                c7, c1 = getattr(ru.fn, "0"), getattr(ru.fn, "2")
                C = c7.conv.in_channels
                ok = ok and (C in (32, 64, 128, 256) and c7.conv.out_channels == C and c7.conv.kernel_size[0] == 7
                             and c1.conv.kernel_size[0] == 1 and 6 * c7.dilation <= 54 and c7.stride == 1)
            plan.append((up, rus))
        return (first, plan, last) if ok else None

    def _decode_tc(self, x, plan):
        """x fp32 [B, n, codebook_dim] channels-last -> wave [B, 1, n * prod(strides)]; C8S between layers"""
        first, blocks, last = plan
        h = ops.codec_pack_c8s(x)
        wu = self._cached(first, "_tc_units", [first.conv.weight], lambda: ops.pack_conv_weights(first.conv.weight))
        h = ops.codec_conv_tc(h, wu, first.conv.bias, cout=first.conv.out_channels,
                              kernel_size=first.conv.kernel_size[0], stride=1, pad_mode=first.pad_mode)
        for up, rus in blocks:
            s_ = up.upsample_factor
            wu = self._cached(up, "_tc_units", [up.conv.weight],
                              lambda up=up, s_=s_: ops.pack_convT_weights(up.conv.weight, s_))
            bias_up = self._cached(up, "_tc_bias", [up.conv.bias],
                                   lambda up=up, s_=s_: up.conv.bias.detach().float().repeat(s_).contiguous())
            h = ops.codec_conv_tc(h, wu, bias_up, cout=s_ * up.conv.out_channels, kernel_size=2, stride=1,
                                  pad_mode="constant", upsample=s_)
            for ru in rus:
                c7, c1 = getattr(ru.fn, "0"), getattr(ru.fn, "2")
                wr = self._cached(ru, "_tc_units", [c7.conv.weight, c1.conv.weight],
                                  lambda c7=c7, c1=c1: ops.pack_ru_weights(c7.conv.weight, c1.conv.weight))
                h = ops.codec_ru_tc(h, wr, c7.conv.bias, c1.conv.bias, dilation=c7.dilation, pad_mode=c7.pad_mode)
        return ops.codec_last_conv(h, last.conv.weight, last.conv.bias, pad_mode=last.pad_mode)

    def decode_frames(self, x):
        """x [B, n, codebook_dim] channels-last (quantized) -> wave [B, 1, T] (soundstream.py:859-861)"""
        plan = self._tc_plan_dec()
        n = x.shape[1]
        # the first residual units see n x first upsampling factor samples and need more than their reflect halo
        if plan is not None and x.is_cuda and n >= 8 and n * plan[1][0][0].upsample_factor > 54:
            return self._decode_tc(x, plan)
        return self.decoder(x.transpose(1, 2).contiguous())

    def encode_frames(self, x):
        """x [B, 1, T] fp32 -> encoder output [B, n, codebook_dim] channels-last (soundstream.py:827-836)."""
        plan = self._tc_plan()
        T = x.shape[-1]
        frames = T // self.seq_len_multiple_of
        # every residual unit needs more samples than its reflect halo (6 x dilation <= 54); the shortest ones see
        # frames x last stride samples
        if plan is not None and x.is_cuda and T % self.seq_len_multiple_of == 0 and frames >= 4 \
                and frames * self.strides[-1] > 54:
            return self._encode_tc(x[:, 0].to(f32).contiguous(), plan)
        return self.encoder(x.to(f32)).transpose(1, 2).contiguous()

    # ---- hot path ----------------------------------------------------------------------------------
    def process_input(self, x, input_sample_hz=None, curtail_from_left=False):
        lead = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1])  # the reference packs every leading dim ('* n', soundstream.py:785)
        if exists(input_sample_hz) and input_sample_hz != self.target_sample_hz:
            from torchaudio.functional import resample
            x = resample(x, input_sample_hz, self.target_sample_hz)
        mult = self.seq_len_multiple_of
        keep = x.shape[-1] // mult * mult
        x = x[..., -keep:] if curtail_from_left else x[..., :keep]
        return x[:, None, :], lead

    def decode_from_codebook_indices(self, quantized_indices):
        assert quantized_indices.dtype in (torch.long, torch.int32)
        if quantized_indices.ndim == 3:
            b, n, gq = quantized_indices.shape
            quantized_indices = quantized_indices.reshape(b, n, self.rq_groups, -1).permute(2, 0, 1, 3)
        return self.decode(self.rq.get_output_from_indices(quantized_indices.long()))

    def decode(self, x, quantize=False):
        if quantize:
            x, *_ = self.rq(x)
        if exists(self.decoder_attn):
            x = self.decoder_attn(x)
        return self.decode_frames(x)

    @torch.no_grad()
    def tokenize(self, audio):
        self.eval()
        return self.forward(audio, return_codes_only=True)

    def forward(self, x, target=None, is_denoising=None, return_encoded=False, return_codes_only=False,
                return_discr_loss=False, return_discr_losses_separately=False, return_loss_breakdown=False,
                return_recons_only=False, input_sample_hz=None, apply_grad_penalty=False, curtail_from_left=False):
        if exists(is_denoising) or exists(target):
            raise NotImplementedError("FiLM denoising / target losses are outside this build")
        x, lead = self.process_input(x, input_sample_hz=input_sample_hz, curtail_from_left=curtail_from_left)
        h = self.encode_frames(x)                                      # b n c
        if exists(self.encoder_attn):
            h = self.encoder_attn(h)
        quantized, indices, commit_loss = self.rq(h)
        if return_codes_only:
            return indices
        if return_encoded:
            g, b, n, q = indices.shape
            return quantized, indices.permute(1, 2, 0, 3).reshape(b, n, g * q), commit_loss
        if exists(self.decoder_attn):
            quantized = self.decoder_attn(quantized)
        recon = self.decode_frames(quantized)
        if return_recons_only:
            return recon.reshape(*lead, *recon.shape[-2:]) if len(lead) != 1 else recon
        raise NotImplementedError("SoundStream training losses (GAN / mel / feature matching) are outside this build")


def AudioLMSoundStream(strides=(2, 4, 5, 8), target_sample_hz=16000, rq_num_quantizers=12, **kwargs):
    return SoundStream(strides=strides, target_sample_hz=target_sample_hz, rq_num_quantizers=rq_num_quantizers, **kwargs)


def MusicLMSoundStream(strides=(3, 4, 5, 8), target_sample_hz=24000, rq_num_quantizers=12, **kwargs):
    return SoundStream(strides=strides, target_sample_hz=target_sample_hz, rq_num_quantizers=rq_num_quantizers, **kwargs)
