"""Drop-in for REF/model/audio_encoder.py: same constructor, same forward signature, same state_dict keys
(SURVEY.md appendix B, 424 tensors for HuBERT-large + pool) -- but forward() is one call into the sm_100a
library (b2s_hubert_forward): conv feature extractor, positional conv, transformer stack, final LayerNorm fused
with AvgPool1d, projector.

The module tree below exists only to own parameters under the reference's names so checkpoints written by the
reference trainer (REF/trainer.py:516-528) and read by its inference script (REF/inference.py:24-26) load
unchanged; both weight-norm spellings of the positional conv are accepted.
There is no CPU path: forward() on a non-CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.nn as nn

from .. import _lib, ops, packing
from ..config import EncoderArch, encoder_arch_from_config


class _Ptr:
    """A raw device address with the one method the ctypes marshalling code calls."""

    def __init__(self, addr: int):
        self.addr = addr

    def data_ptr(self) -> int:
        return self.addr


class _Params(nn.Module):
    """A bag of named parameters (so state_dict keys read `<name>.weight` / `<name>.bias`)."""

    def __init__(self, **shapes):
        super().__init__()
        for name, shape in shapes.items():
            self.register_parameter(name, nn.Parameter(torch.zeros(*shape)))


class _ConvLayer(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = _Params(weight=(cout, cin, k), bias=(cout,))
        self.layer_norm = _Params(weight=(cout,), bias=(cout,))


class _FeatureExtractor(nn.Module):
    def __init__(self, arch: EncoderArch):
        super().__init__()
        cin = 1
        layers = []
        for co, k in zip(arch.conv_dim, arch.conv_kernel):
            layers.append(_ConvLayer(cin, co, k))
            cin = co
        self.conv_layers = nn.ModuleList(layers)


class _FeatureProjection(nn.Module):
    def __init__(self, arch: EncoderArch):
        super().__init__()
        self.layer_norm = _Params(weight=(arch.conv_dim[-1],), bias=(arch.conv_dim[-1],))
        self.projection = _Params(weight=(arch.hidden, arch.conv_dim[-1]), bias=(arch.hidden,))


class _WeightNormParams(nn.Module):
    def __init__(self, arch: EncoderArch):
        super().__init__()
        self.weight = _Params(original0=(1, 1, arch.pos_k), original1=(arch.hidden, arch.hidden // arch.pos_groups,
                                                                     arch.pos_k))


class _PosConvInner(nn.Module):
    def __init__(self, arch: EncoderArch):
        super().__init__()
        self.register_parameter("bias", nn.Parameter(torch.zeros(arch.hidden)))
        self.parametrizations = _WeightNormParams(arch)


class _PosConv(nn.Module):
    def __init__(self, arch: EncoderArch):
        super().__init__()
        self.conv = _PosConvInner(arch)


class _Attention(nn.Module):
    def __init__(self, H):
        super().__init__()
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            setattr(self, n, _Params(weight=(H, H), bias=(H,)))


class _FeedForward(nn.Module):
    def __init__(self, H, Fd):
        super().__init__()
        self.intermediate_dense = _Params(weight=(Fd, H), bias=(Fd,))
        self.output_dense = _Params(weight=(H, Fd), bias=(H,))


class _EncoderLayer(nn.Module):
    def __init__(self, H, Fd):
        super().__init__()
        self.attention = _Attention(H)
        self.layer_norm = _Params(weight=(H,), bias=(H,))
        self.feed_forward = _FeedForward(H, Fd)
        self.final_layer_norm = _Params(weight=(H,), bias=(H,))


class _Encoder(nn.Module):
    def __init__(self, arch: EncoderArch):
        super().__init__()
        self.pos_conv_embed = _PosConv(arch)
        self.layer_norm = _Params(weight=(arch.hidden,), bias=(arch.hidden,))
        self.layers = nn.ModuleList([_EncoderLayer(arch.hidden, arch.ffn) for _ in range(arch.layers)])


class _EncoderConfigView:
    """`self.encoder.config.hidden_size` is read by the reference's constructor (REF/model/audio_encoder.py:40)."""

    def __init__(self, arch: EncoderArch):
        self.hidden_size = arch.hidden
        self.num_hidden_layers = arch.layers
        self.num_attention_heads = arch.heads
        self.intermediate_size = arch.ffn


class HubertBackbone(nn.Module):
    """Parameter container with HF HubertModel's names (TF/models/hubert/modeling_hubert.py)."""

    def __init__(self, arch: EncoderArch):
        super().__init__()
        self.arch = arch
        self.config = _EncoderConfigView(arch)
        self.register_parameter("masked_spec_embed", nn.Parameter(torch.zeros(arch.hidden)))
        self.feature_extractor = _FeatureExtractor(arch)
        self.feature_projection = _FeatureProjection(arch)
        self.encoder = _Encoder(arch)


class _WhisperAttn(nn.Module):
    def __init__(self, H):
        super().__init__()
        self.k_proj = _Params(weight=(H, H))  # no bias (TF/models/whisper/modeling_whisper.py:279)
        self.v_proj = _Params(weight=(H, H), bias=(H,))
        self.q_proj = _Params(weight=(H, H), bias=(H,))
        self.out_proj = _Params(weight=(H, H), bias=(H,))


class _WhisperLayer(nn.Module):
    def __init__(self, H, Fd):
        super().__init__()
        self.self_attn = _WhisperAttn(H)
        self.self_attn_layer_norm = _Params(weight=(H,), bias=(H,))
        self.fc1 = _Params(weight=(Fd, H), bias=(Fd,))
        self.fc2 = _Params(weight=(H, Fd), bias=(H,))
        self.final_layer_norm = _Params(weight=(H,), bias=(H,))


class WhisperBackbone(nn.Module):
    """Parameter container with HF WhisperEncoder's names (TF/models/whisper/modeling_whisper.py:541-647)."""

    def __init__(self, arch):
        super().__init__()
        self.arch = arch
        self.config = _EncoderConfigView(arch)
        self.conv1 = _Params(weight=(arch.hidden, arch.mel_bins, 3), bias=(arch.hidden,))
        self.conv2 = _Params(weight=(arch.hidden, arch.hidden, 3), bias=(arch.hidden,))
        self.embed_positions = _Params(weight=(arch.max_positions, arch.hidden))
        self.embed_positions.weight.requires_grad_(False)
        self.layers = nn.ModuleList([_WhisperLayer(arch.hidden, arch.ffn) for _ in range(arch.layers)])
        self.layer_norm = _Params(weight=(arch.hidden,), bias=(arch.hidden,))


def _random_init_requested(config) -> bool:
    return bool(getattr(config.model.audio_encoder, "random_init", False))


@torch.no_grad()
def init_hf_default_(backbone: nn.Module, seed: int = 1234) -> nn.Module:
    """HF's `_init_weights` distributions for a from-scratch backbone (TF/models/hubert/modeling_hubert.py:640-673,
    TF/models/whisper/modeling_whisper.py `_init_weights`): Linear / Whisper conv N(0, 0.02), HuBERT feature-extractor
    convs kaiming-normal, LayerNorm 1 / 0, biases 0, positional conv N(0, 2 sqrt(1 / (k C_in))) under weight-norm,
    masked_spec_embed uniform, Whisper's frozen sinusoid table. Only used when the config says
    `model.audio_encoder.random_init: true` (benchmarks, tests): the default is the pretrained checkpoint."""
    import math
    g = torch.Generator().manual_seed(int(seed))
    v_param = None
    for name, p in backbone.named_parameters():
        if "layer_norm.weight" in name:
            p.fill_(1.0)
        elif name.endswith("bias") or name.endswith(".bias"):
            p.zero_()
        elif name.endswith("masked_spec_embed"):
            p.copy_(torch.rand(p.shape, generator=g))
        elif name.endswith("parametrizations.weight.original1"):
            cout, cin_g, k = p.shape
            p.copy_(torch.randn(p.shape, generator=g) * (2.0 * math.sqrt(1.0 / (k * cout))))
            v_param = p
        elif name.endswith("parametrizations.weight.original0"):
            pass  # = ||v|| per tap, set below
        elif name.endswith("embed_positions.weight"):
            length, channels = p.shape
            inc = math.log(10000.0) / (channels // 2 - 1)
            inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float32))
            t = torch.arange(length, dtype=torch.float32)[:, None] * inv[None, :]
            p.copy_(torch.cat([t.sin(), t.cos()], dim=1))
        elif "feature_extractor.conv_layers" in name and name.endswith("conv.weight"):
            fan_in = p.shape[1] * p.shape[2]
            p.copy_(torch.randn(p.shape, generator=g) * math.sqrt(2.0 / fan_in))
        else:
            p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    if v_param is not None:
        for name, p in backbone.named_parameters():
            if name.endswith("parametrizations.weight.original0"):
                p.copy_(v_param.pow(2).sum(dim=(0, 1), keepdim=True).sqrt())
    return backbone


def _load_pretrained_into(backbone: nn.Module, hub_id: str, pick=lambda m: m) -> None:
    """REF/model/audio_encoder.py:6-13: `AutoModel.from_pretrained(config.model.audio_encoder.type)`; the tensors are
    re-homed in the parameter container (same names as HF's module). Raises when the checkpoint cannot be read -- the
    module must never be left at its all-zero construction state."""
    try:
        from transformers import AutoModel
        hf = pick(AutoModel.from_pretrained(hub_id))
    except Exception as e:  # no network / no local cache / transformers missing
        raise RuntimeError(
            f"could not load the pretrained audio encoder {hub_id!r} ({type(e).__name__}: {e}). The reference downloads "
            "it in AudioEncoder.__init__ (REF/model/audio_encoder.py:6-13); offline, point HF_HOME at a local copy, or "
            "set `model.audio_encoder.random_init: true` in the config to start from HF's default random "
            "initialisation (benchmarks / tests).") from e
    sd = dict(hf.state_dict())
    for k in list(sd):  # torch-2.0 weight-norm spelling
        for old, new in (("weight_g", "parametrizations.weight.original0"), ("weight_v", "parametrizations.weight.original1")):
            if k.endswith("pos_conv_embed.conv." + old):
                sd[k[:-len(old)] + new] = sd.pop(k)
    own = backbone.state_dict()
    missing = [k for k in own if k not in sd]
    if missing:
        raise RuntimeError(f"pretrained checkpoint {hub_id!r} lacks {len(missing)} tensors, e.g. {missing[:3]}")
    backbone.load_state_dict({k: sd[k] for k in own}, strict=True)


def load_whisper_encoder(config):
    """REF/model/audio_encoder.py:10-13: openai/whisper-medium's encoder + its feature extractor. The log-mel extractor
    is transformers' WhisperFeatureExtractor with its default (= whisper) parameters, which needs no download."""
    from ..config import whisper_arch_from_config
    feature_extractor = None
    try:
        from transformers import WhisperFeatureExtractor
        feature_extractor = WhisperFeatureExtractor()
    except Exception:  # transformers missing: the collate-side extractor is outside the hot path anyway
        pass
    backbone = WhisperBackbone(whisper_arch_from_config(config))
    if _random_init_requested(config):
        init_hf_default_(backbone, getattr(config, "seed_everything", 1234))
    else:
        _load_pretrained_into(backbone, config.model.audio_encoder.type, pick=lambda m: m.encoder)
    return backbone, feature_extractor


def load_hubert_encoder(config):
    """REF/model/audio_encoder.py:6-7: facebook/hubert-large-ls960-ft (`config.model.audio_encoder.type`). The
    architecture comes from the config (defaults = HuBERT-large); the weights from the hub checkpoint unless the config
    asks for `model.audio_encoder.random_init: true`."""
    backbone = HubertBackbone(encoder_arch_from_config(config))
    if _random_init_requested(config):
        init_hf_default_(backbone, getattr(config, "seed_everything", 1234))
    else:
        _load_pretrained_into(backbone, config.model.audio_encoder.type)
    return backbone


class AudioEncoder(nn.Module):
    def __init__(self, config, device):
        super(AudioEncoder, self).__init__()
        self.config = config
        self.device = device

        if self.config.model.audio_encoder.base == "hubert":
            self.encoder_base = "hubert"
            self.encoder = load_hubert_encoder(self.config)
        elif self.config.model.audio_encoder.base == "whisper":
            self.encoder_base = "whisper"
            self.encoder, self.feature_extractor = load_whisper_encoder(self.config)
        else:
            raise Exception("Unexpected encoder type in config.")

        self.downsample_method = self.config.model.audio_encoder.downsample_method
        self.downsample_factor = self.config.model.audio_encoder.downsample_factor
        H = self.encoder.config.hidden_size
        if self.downsample_method == "pool":
            self.pool_kernel = int(self.config.model.audio_encoder.pooling.kernel_size)
            self.pool_stride = int(self.config.model.audio_encoder.pooling.stride)
            self.embed_projection = nn.Linear(H, self.config.model.llm_embedding_channels)
        elif self.downsample_method in ("stack", "ctc_pool"):
            # unused by every shipped yaml; kept out of the hot path on purpose
            raise NotImplementedError(f"downsample_method={self.downsample_method!r} is outside the B200 hot path")
        else:
            raise Exception("Invalid downsampling method for audio encoder.")

        # 16-bit format of the GEMM / attention operands (weights copies, activations, gradients). fp16 is what the
        # reference computes in (torch.autocast(dtype=torch.float16), REF/trainer.py:270, REF/inference.py:99) and what
        # keeps the projected embeddings accurate enough for the 2e-2 logit tolerance downstream; the fp32 master
        # parameters are untouched. `model.audio_encoder.compute_dtype: bfloat16` in the yaml (or assigning
        # `operand_dtype`) selects bf16.
        cd = getattr(self.config.model.audio_encoder, "compute_dtype", "float16")
        self.operand_dtype = {"float16": torch.float16, "fp16": torch.float16, "half": torch.float16,
                              "bfloat16": torch.bfloat16, "bf16": torch.bfloat16}[str(cd).replace("torch.", "")]
        self._packed = None
        self._packed_key = None
        self._pos_w_packed = None
        self._pos_w_dgrad = None
        self._grads = None
        self._train_ctx = None
        # train-mode regularisers (dropout / LayerDrop / SpecAugment, regularizers.RegularizerConfig). None keeps
        # forward_train deterministic; EncoderTrainer(regularize=True) or the caller sets it. HuBERT only: Whisper-medium
        # has dropout = encoder_layerdrop = 0.
        self.regularizers = None
        self._register_load_state_dict_pre_hook(self._rename_legacy_weight_norm)

    # -- checkpoints written by torch 2.0 spell the weight-norm parameters weight_g / weight_v
    @staticmethod
    def _rename_legacy_weight_norm(state_dict, prefix, *args):
        base = prefix + "encoder.encoder.pos_conv_embed.conv."
        for old, new in (("weight_g", "parametrizations.weight.original0"),
                         ("weight_v", "parametrizations.weight.original1")):
            if base + old in state_dict:
                state_dict[base + new] = state_dict.pop(base + old)

    # ---------------------------------------------------------------------------------------------
    def _weights_key(self):
        return (self.operand_dtype,) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def pack_weights(self, force: bool = False):
        """bf16 / K-major copies of the parameters in the layouts the kernels consume, rebuilt whenever a
        parameter changes (optimizer step, load_state_dict)."""
        key = self._weights_key()
        if not force and self._packed is not None and self._packed_key == key:
            return self._packed
        if self.encoder_base == "whisper":
            self._packed = self._pack_whisper()
            self._packed_key = key
            return self._packed
        enc = self.encoder
        arch: EncoderArch = enc.arch
        dev = enc.masked_spec_embed.device
        if dev.type != "cuda":
            raise RuntimeError("AudioEncoder (B200 path) needs its parameters on a CUDA device; there is no CPU path")
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        bf = lambda t: t.detach().to(self.operand_dtype).contiguous()
        keep: List[torch.Tensor] = []

        def K(t):
            keep.append(t)
            return t.data_ptr()

        w = _lib.HubertWeights()
        c0 = enc.feature_extractor.conv_layers[0]
        w.conv0_w = K(f32(c0.conv.weight).reshape(arch.conv_dim[0], arch.conv_kernel[0]).contiguous())
        w.conv0_b, w.conv0_ln_g, w.conv0_ln_b = K(f32(c0.conv.bias)), K(f32(c0.layer_norm.weight)), K(
            f32(c0.layer_norm.bias))
        for i in range(6):
            cl = enc.feature_extractor.conv_layers[i + 1]
            w.conv_w[i] = K(bf(packing.pack_conv(cl.conv.weight.detach())))
            w.conv_b[i], w.conv_ln_g[i], w.conv_ln_b[i] = K(f32(cl.conv.bias)), K(f32(cl.layer_norm.weight)), K(
                f32(cl.layer_norm.bias))
            w.conv_k[i], w.conv_stride[i] = arch.conv_kernel[i + 1], arch.conv_stride[i + 1]
        fp = enc.feature_projection
        w.fp_ln_g, w.fp_ln_b = K(f32(fp.layer_norm.weight)), K(f32(fp.layer_norm.bias))
        w.fp_w, w.fp_b = K(bf(fp.projection.weight)), K(f32(fp.projection.bias))
        pc = enc.encoder.pos_conv_embed.conv
        self._pos_w_packed = ops.posconv_weight_pack(f32(pc.parametrizations.weight.original0).reshape(-1),
                                                     f32(pc.parametrizations.weight.original1),
                                                     out_dtype=self.operand_dtype)
        self._pos_w_dgrad = None
        w.pos_w = K(self._pos_w_packed)
        w.pos_b, w.pos_k, w.pos_groups = K(f32(pc.bias)), arch.pos_k, arch.pos_groups
        layers = (_lib.EncoderLayer * arch.layers)()
        for l, lay in enumerate(enc.encoder.layers):
            a = lay.attention
            L = layers[l]
            L.ln1_g, L.ln1_b = K(f32(lay.layer_norm.weight)), K(f32(lay.layer_norm.bias))
            L.wqkv = K(bf(packing.pack_qkv(a.q_proj.weight.detach(), a.k_proj.weight.detach(),
                                           a.v_proj.weight.detach())))
            L.bqkv = K(f32(torch.cat([a.q_proj.bias.detach(), a.k_proj.bias.detach(), a.v_proj.bias.detach()])))
            L.wo, L.bo = K(bf(a.out_proj.weight)), K(f32(a.out_proj.bias))
            L.ln2_g, L.ln2_b = K(f32(lay.final_layer_norm.weight)), K(f32(lay.final_layer_norm.bias))
            L.w1, L.b1 = K(bf(lay.feed_forward.intermediate_dense.weight)), K(
                f32(lay.feed_forward.intermediate_dense.bias))
            L.w2, L.b2 = K(bf(lay.feed_forward.output_dense.weight)), K(f32(lay.feed_forward.output_dense.bias))
        w.layers = C.cast(layers, C.POINTER(_lib.EncoderLayer))
        w.num_layers, w.hidden, w.heads, w.ffn = arch.layers, arch.hidden, arch.heads, arch.ffn
        w.final_ln_g, w.final_ln_b = K(f32(enc.encoder.layer_norm.weight)), K(f32(enc.encoder.layer_norm.bias))
        w.ln_eps = arch.ln_eps
        w.pool_kernel, w.pool_stride = self.pool_kernel, self.pool_stride
        w.proj_w, w.proj_b = K(bf(self.embed_projection.weight)), K(f32(self.embed_projection.bias))
        w.llm_dim = self.embed_projection.out_features
        w.fmt = ops.fmt_of(self.operand_dtype)
        self._packed = (w, layers, keep)
        self._packed_key = key
        return self._packed

    def _pack_whisper(self):
        enc = self.encoder
        arch = enc.arch
        dev = enc.conv1.weight.device
        if dev.type != "cuda":
            raise RuntimeError("AudioEncoder (B200 path) needs its parameters on a CUDA device; there is no CPU path")
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        bf = lambda t: t.detach().to(self.operand_dtype).contiguous()
        keep: List[torch.Tensor] = []

        def K(t):
            keep.append(t)
            return t.data_ptr()

        w = _lib.WhisperWeights()
        w.conv1_w, w.conv1_b = K(bf(packing.pack_conv(enc.conv1.weight.detach()))), K(f32(enc.conv1.bias))
        w.conv2_w, w.conv2_b = K(bf(packing.pack_conv(enc.conv2.weight.detach()))), K(f32(enc.conv2.bias))
        w.pos_emb = K(f32(enc.embed_positions.weight))
        layers = (_lib.EncoderLayer * arch.layers)()
        for l, lay in enumerate(enc.layers):
            a = lay.self_attn
            L = layers[l]
            L.ln1_g, L.ln1_b = K(f32(lay.self_attn_layer_norm.weight)), K(f32(lay.self_attn_layer_norm.bias))
            L.wqkv = K(bf(packing.pack_qkv(a.q_proj.weight.detach(), a.k_proj.weight.detach(),
                                           a.v_proj.weight.detach())))
            L.bqkv = K(f32(torch.cat([a.q_proj.bias.detach(), torch.zeros_like(a.q_proj.bias), a.v_proj.bias.detach()])))
            L.wo, L.bo = K(bf(a.out_proj.weight)), K(f32(a.out_proj.bias))
            L.ln2_g, L.ln2_b = K(f32(lay.final_layer_norm.weight)), K(f32(lay.final_layer_norm.bias))
            L.w1, L.b1 = K(bf(lay.fc1.weight)), K(f32(lay.fc1.bias))
            L.w2, L.b2 = K(bf(lay.fc2.weight)), K(f32(lay.fc2.bias))
        w.layers = C.cast(layers, C.POINTER(_lib.EncoderLayer))
        w.num_layers, w.hidden, w.heads, w.ffn = arch.layers, arch.hidden, arch.heads, arch.ffn
        w.mel_bins, w.max_positions = arch.mel_bins, arch.max_positions
        w.final_ln_g, w.final_ln_b = K(f32(enc.layer_norm.weight)), K(f32(enc.layer_norm.bias))
        w.ln_eps = arch.ln_eps
        w.pool_kernel, w.pool_stride = self.pool_kernel, self.pool_stride
        w.proj_w, w.proj_b = K(bf(self.embed_projection.weight)), K(f32(self.embed_projection.bias))
        w.llm_dim = self.embed_projection.out_features
        w.fmt = ops.fmt_of(self.operand_dtype)
        return (w, layers, keep)

    def _whisper_forward_fp32(self, input: torch.Tensor, return_last_hidden: bool = False):
        w = self.pack_weights()[0]
        mel = input.to(torch.float32).contiguous()
        if mel.dim() != 3 or mel.shape[1] != w.mel_bins:
            raise ValueError(f"expected (B, {w.mel_bins}, T) log-mel features")
        B, _, T = mel.shape
        if T != 2 * w.max_positions:  # same check and message as WhisperEncoder.forward
            raise ValueError(f"Whisper expects the mel input features to be of length {2 * w.max_positions}, but "
                             f"found {T}. Make sure to pad the input mel features to {2 * w.max_positions}.")
        pooled = (w.max_positions - w.pool_kernel) // w.pool_stride + 1
        lib = _lib.load()
        nbytes = lib.b2s_whisper_workspace_bytes(C.byref(w), B)
        ws = torch.empty(nbytes, device=mel.device, dtype=torch.uint8)
        out = torch.empty(B, pooled, w.llm_dim, device=mel.device, dtype=torch.float32)
        last = (torch.empty(B, w.max_positions, w.hidden, device=mel.device, dtype=torch.float32)
                if return_last_hidden else None)
        _lib.check(lib.b2s_whisper_forward(C.byref(w), mel.data_ptr(), B, T, ws.data_ptr(), nbytes, out.data_ptr(),
                                           None if last is None else last.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "whisper_forward")
        return (out, last) if return_last_hidden else out

    def extract_features(self, waves: torch.Tensor) -> torch.Tensor:
        """Whisper collate step on the GPU (REF/trainer.py:178-182): (B, n_samples) fp32 CUDA waveforms, already padded
        / truncated to the extractor's 30 s window, -> (B, 80, 3000) log-mel `input_features`. Same numbers as
        `self.feature_extractor(..., return_tensors="pt").input_features` without the CPU STFT."""
        if self.encoder_base != "whisper":
            raise Exception("Unexpected encoder type in config.")
        if not waves.is_cuda:
            raise RuntimeError("AudioEncoder.extract_features (B200 path) needs a CUDA input; there is no CPU path")
        if self.feature_extractor is None:
            raise RuntimeError("transformers' WhisperFeatureExtractor (mel filter bank) is unavailable")
        if getattr(self, "_mel_filters", None) is None or self._mel_filters.device != waves.device:
            self._mel_filters = torch.as_tensor(self.feature_extractor.mel_filters, dtype=torch.float32).contiguous().to(
                waves.device)
        w = waves.to(torch.float32)
        if w.stride(-1) != 1:
            w = w.contiguous()
        return ops.whisper_log_mel(w, self._mel_filters)

    def num_frames(self, samples: int):
        w = self.pack_weights()[0]
        frames, pooled = C.c_int32(), C.c_int32()
        _lib.check(_lib.load().b2s_hubert_num_frames(C.byref(w), samples, C.byref(frames), C.byref(pooled)),
                   "hubert_num_frames")
        return frames.value, pooled.value

    def forward_fp32(self, input: torch.Tensor, return_last_hidden: bool = False):
        """(B, T0) waveform -> fp32 (B, A, llm_dim) projected audio embeddings (and optionally the fp32
        pre-final-norm residual stream (B, N, H) for parity tests)."""
        if not input.is_cuda:
            raise RuntimeError("AudioEncoder.forward (B200 path) needs a CUDA input; there is no CPU path")
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError(
                "autograd does not flow through the B200 forward: use forward_train() / backward() / flush_grads() "
                "(the explicit training path), or call .eval() / torch.no_grad() for inference")
        if self.encoder_base == "whisper":
            return self._whisper_forward_fp32(input, return_last_hidden)
        w = self.pack_weights()[0]
        wave = input.to(torch.float32)
        if wave.dim() != 2:
            raise ValueError("expected a (B, T0) waveform batch")
        if wave.stride(1) != 1:
            wave = wave.contiguous()
        B, T0 = wave.shape
        frames, pooled = self.num_frames(T0)
        lib = _lib.load()
        nbytes = lib.b2s_hubert_workspace_bytes(C.byref(w), B, T0)
        ws = torch.empty(nbytes, device=wave.device, dtype=torch.uint8)
        out = torch.empty(B, pooled, w.llm_dim, device=wave.device, dtype=torch.float32)
        last = torch.empty(B, frames, w.hidden, device=wave.device, dtype=torch.float32) if return_last_hidden else None
        _lib.check(lib.b2s_hubert_forward(C.byref(w), wave.data_ptr(), wave.stride(0), B, T0, ws.data_ptr(), nbytes,
                                          out.data_ptr(), None if last is None else last.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "hubert_forward")
        return (out, last) if return_last_hidden else out

    # ------------------------------------------------------------------------------------------ training
    def mark_weights_changed(self) -> None:
        """The optimizer updates the parameters in place with its own kernel (no autograd version bump): drop the
        packed bf16 copies so the next forward rebuilds them."""
        self._packed = None
        self._packed_key = None
        self._pos_w_dgrad = None

    def transformer_layers(self):
        return self.encoder.encoder.layers if self.encoder_base == "hubert" else self.encoder.layers

    def flat_param_order(self):
        """Parameter order for a flat optimizer buffer: the transformer layers first, each layer's parameters in ONE
        contiguous block (= one all-reduce bucket, exchanged while the layers below it are still in their backward),
        with q|k|v weights (and biases) adjacent inside it so the fused-QKV gradient [3H, H] the kernels produce IS the
        concatenation of the three `.grad` views; everything else (conv stack, projections, final norm) follows."""
        order, seen = [], set()

        def push(p):
            if p is not None and id(p) not in seen:
                seen.add(id(p))
                order.append(p)

        hubert = self.encoder_base == "hubert"
        for lay in self.transformer_layers():
            a = lay.attention if hubert else lay.self_attn
            for proj in (a.q_proj, a.k_proj, a.v_proj):
                push(proj.weight)
            if hubert:  # Whisper's k_proj has no bias: its fused bias gradient goes through scratch
                for proj in (a.q_proj, a.k_proj, a.v_proj):
                    push(proj.bias)
            for p in lay.parameters():
                push(p)
        for p in self.parameters():
            push(p)
        return order

    def layer_param_groups(self):
        """Parameters of each transformer layer (index = layer): the all-reduce buckets of the training step."""
        return [list(lay.parameters()) for lay in self.transformer_layers()]

    def _grad_spec(self):
        """(buffer name, packed shape, parameters that tile the buffer row-wise) for every accumulator whose packed
        layout equals the parameters' own layout (everything except the conv weights and the weight-normed pos conv)."""
        if self.encoder_base == "whisper":
            return self._grad_spec_whisper()
        enc = self.encoder
        arch = enc.arch
        H, F_ = arch.hidden, arch.ffn
        fe, fp, pc = enc.feature_extractor.conv_layers, enc.feature_projection, enc.encoder.pos_conv_embed.conv
        C_ = self.embed_projection.out_features
        spec = [("masked_spec_embed", (H,), [enc.masked_spec_embed]),
                ("conv0_w", (arch.conv_dim[0], arch.conv_kernel[0]), [fe[0].conv.weight]),
                ("conv0_b", (512,), [fe[0].conv.bias]), ("conv0_ln_g", (512,), [fe[0].layer_norm.weight]),
                ("conv0_ln_b", (512,), [fe[0].layer_norm.bias]),
                ("fp_ln_g", (512,), [fp.layer_norm.weight]), ("fp_ln_b", (512,), [fp.layer_norm.bias]),
                ("fp_w", (H, 512), [fp.projection.weight]), ("fp_b", (H,), [fp.projection.bias]),
                ("pos_b", (H,), [pc.bias]), ("final_ln_g", (H,), [enc.encoder.layer_norm.weight]),
                ("final_ln_b", (H,), [enc.encoder.layer_norm.bias]), ("proj_w", (C_, H), [self.embed_projection.weight]),
                ("proj_b", (C_,), [self.embed_projection.bias])]
        for i in range(6):
            spec += [(f"conv_b{i}", (512,), [fe[i + 1].conv.bias]), (f"conv_ln_g{i}", (512,), [fe[i + 1].layer_norm.weight]),
                     (f"conv_ln_b{i}", (512,), [fe[i + 1].layer_norm.bias])]
        for l, lay in enumerate(enc.encoder.layers):
            a, ff = lay.attention, lay.feed_forward
            spec += [(f"l{l}.ln1_g", (H,), [lay.layer_norm.weight]), (f"l{l}.ln1_b", (H,), [lay.layer_norm.bias]),
                     (f"l{l}.wqkv", (3 * H, H), [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]),
                     (f"l{l}.bqkv", (3 * H,), [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias]),
                     (f"l{l}.wo", (H, H), [a.out_proj.weight]), (f"l{l}.bo", (H,), [a.out_proj.bias]),
                     (f"l{l}.ln2_g", (H,), [lay.final_layer_norm.weight]), (f"l{l}.ln2_b", (H,), [lay.final_layer_norm.bias]),
                     (f"l{l}.w1", (F_, H), [ff.intermediate_dense.weight]), (f"l{l}.b1", (F_,), [ff.intermediate_dense.bias]),
                     (f"l{l}.w2", (H, F_), [ff.output_dense.weight]), (f"l{l}.b2", (H,), [ff.output_dense.bias])]
        return spec

    def _grad_spec_whisper(self):
        enc = self.encoder
        arch = enc.arch
        H, F_ = arch.hidden, arch.ffn
        C_ = self.embed_projection.out_features
        spec = [("conv1_b", (H,), [enc.conv1.bias]), ("conv2_b", (H,), [enc.conv2.bias]),
                ("final_ln_g", (H,), [enc.layer_norm.weight]), ("final_ln_b", (H,), [enc.layer_norm.bias]),
                ("proj_w", (C_, H), [self.embed_projection.weight]), ("proj_b", (C_,), [self.embed_projection.bias])]
        for l, lay in enumerate(enc.layers):
            a = lay.self_attn
            spec += [(f"l{l}.ln1_g", (H,), [lay.self_attn_layer_norm.weight]),
                     (f"l{l}.ln1_b", (H,), [lay.self_attn_layer_norm.bias]),
                     (f"l{l}.wqkv", (3 * H, H), [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight]),
                     (f"l{l}.wo", (H, H), [a.out_proj.weight]), (f"l{l}.bo", (H,), [a.out_proj.bias]),
                     (f"l{l}.ln2_g", (H,), [lay.final_layer_norm.weight]), (f"l{l}.ln2_b", (H,), [lay.final_layer_norm.bias]),
                     (f"l{l}.w1", (F_, H), [lay.fc1.weight]), (f"l{l}.b1", (F_,), [lay.fc1.bias]),
                     (f"l{l}.w2", (H, F_), [lay.fc2.weight]), (f"l{l}.b2", (H,), [lay.fc2.bias])]
        return spec

    def _grad_buffers(self):
        """fp32 gradient accumulators in the kernels' packed layouts (include/b2s.h: b2s_hubert_grads). When the
        parameters already own fp32 `.grad` tensors laid out like the packed buffer (a flat optimizer buffer in
        `flat_param_order`), the kernels accumulate straight into them; otherwise into scratch that `flush_grads`
        adds to `.grad`."""
        hubert = self.encoder_base == "hubert"
        key = tuple(0 if p.grad is None else p.grad.data_ptr() for p in self.parameters())
        if self._grads is not None and self._grads[3] == key:
            return self._grads
        if self._grads is not None and any(float(t.abs().max()) != 0.0 for t in self._grads[2].values()):
            raise RuntimeError("AudioEncoder: .grad tensors were replaced while un-flushed gradients are pending")
        arch = self.encoder.arch
        dev = self.embed_projection.weight.device
        H, L = arch.hidden, arch.layers
        z = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
        scratch, ptr, pending = {}, {}, []

        def direct(ps):
            addr = None
            for q in ps:
                gq = q.grad
                if gq is None or gq.dtype != torch.float32 or not gq.is_contiguous() or gq.shape != q.shape:
                    return None
                if addr is not None and gq.data_ptr() != addr:
                    return None
                addr = gq.data_ptr() + gq.numel() * 4
            return ps[0].grad.data_ptr()

        for name, shape, ps in self._grad_spec():
            d = direct(ps)
            if d is None:
                scratch[name] = z(*shape)
                pending.append((name, ps))
                d = scratch[name].data_ptr()
            ptr[name] = d
        if hubert:
            scratch["pos_w"] = z(H, arch.pos_k * (H // arch.pos_groups))
            for i in range(6):
                scratch[f"conv_w{i}"] = z(512, arch.conv_kernel[i + 1] * 512)
            g = _lib.HubertGrads()
            names = ("conv0_w", "conv0_b", "conv0_ln_g", "conv0_ln_b", "fp_ln_g", "fp_ln_b", "fp_w", "fp_b", "pos_w",
                     "pos_b", "final_ln_g", "final_ln_b", "proj_w", "proj_b")
        else:
            scratch["conv1_w"] = z(H, 3 * arch.mel_bins)
            scratch["conv2_w"] = z(H, 3 * H)
            for l in range(L):  # q | (no k bias) | v
                scratch[f"l{l}.bqkv"] = z(3 * H)
            g = _lib.WhisperGrads()
            names = ("conv1_w", "conv1_b", "conv2_w", "conv2_b", "final_ln_g", "final_ln_b", "proj_w", "proj_b")
        for k, v in scratch.items():
            ptr.setdefault(k, v.data_ptr())
        for k in names:
            setattr(g, k, ptr[k])
        if hubert:
            for i in range(6):
                g.conv_w[i], g.conv_b[i] = ptr[f"conv_w{i}"], ptr[f"conv_b{i}"]
                g.conv_ln_g[i], g.conv_ln_b[i] = ptr[f"conv_ln_g{i}"], ptr[f"conv_ln_b{i}"]
        layers = (_lib.EncoderLayerGrads * L)()
        for l in range(L):
            for k in ("ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2"):
                setattr(layers[l], k, ptr[f"l{l}.{k}"])
        g.layers = C.cast(layers, C.POINTER(_lib.EncoderLayerGrads))
        self._grads = (g, layers, scratch, key, pending, _Ptr(ptr["masked_spec_embed"]) if hubert else None)
        return self._grads

    def num_audio_embeds(self, samples: int) -> int:
        """Pooled frames (= projected audio embeddings) an utterance of `samples` samples produces on its own."""
        return self.num_frames(int(samples))[1]

    def forward_train(self, input: torch.Tensor, generator=None, draw=None, lengths=None) -> torch.Tensor:
        """Training forward (REF/trainer.py:278): (B, T0) waveform -> fp32 (B, A, llm_dim), keeping the activations
        for `backward`. With `self.regularizers` set and the module in train mode, HF's train-mode dropout / LayerDrop /
        SpecAugment are applied (host randomness from `generator`, or an explicit `draw`); otherwise deterministic.
        `lengths` (samples per utterance, HuBERT only): a ragged batch, zero-padded on the right to T0 like the
        reference's collate (REF/trainer.py:146-149). Every utterance gets the numbers it would get alone; rows
        >= num_audio_embeds(lengths[b]) of output b are padding (REF/trainer.py:280-291 crops them)."""
        if not input.is_cuda:
            raise RuntimeError("AudioEncoder.forward_train (B200 path) needs a CUDA input; there is no CPU path")
        if self.encoder_base == "whisper":
            if lengths is not None:
                raise NotImplementedError("ragged batches are a HuBERT feature: Whisper inputs are fixed 30 s windows")
            return self._whisper_forward_train(input)
        w = self.pack_weights()[0]
        wave = input.to(torch.float32)
        if wave.dim() != 2:
            raise ValueError("expected a (B, T0) waveform batch")
        if wave.stride(1) != 1:
            wave = wave.contiguous()
        B, T0 = wave.shape
        frames, pooled = self.num_frames(T0)
        lib = _lib.load()
        nbytes = lib.b2s_hubert_saved_bytes(C.byref(w), B, T0)
        ctx = self._train_ctx
        if ctx is None or ctx["saved"].numel() < nbytes:
            ctx = {"saved": torch.empty(nbytes, device=wave.device, dtype=torch.uint8)}
        out = torch.empty(B, pooled, w.llm_dim, device=wave.device, dtype=torch.float32)
        if draw is None and self.regularizers is not None and self.training:
            from ..regularizers import draw as _draw
            vf = None
            if lengths is not None and len(set(int(n) for n in lengths)) > 1:
                vf = [self.num_frames(int(n))[0] for n in lengths]  # SpecAugment spans inside each utterance's own length
            draw = _draw(self.regularizers, B, frames, w.num_layers, wave.device, generator, valid_frames=vf)
        reg = None
        if draw is not None:
            mse = self.encoder.masked_spec_embed.detach()
            if mse.dtype != torch.float32 or not mse.is_contiguous():
                mse = mse.float().contiguous()
            ctx["mse"] = mse
            reg = C.byref(draw.c_struct(mse, None))
        c_len, n_valid = None, None
        if lengths is not None:
            lengths = [int(n) for n in lengths]
            if len(lengths) != B or max(lengths) > T0 or min(lengths) <= 0:
                raise ValueError("lengths must give 0 < samples <= T0 for each of the B utterances")
            if any(n != T0 for n in lengths):
                c_len = (C.c_int32 * B)(*lengths)
                n_valid = [self.num_audio_embeds(n) for n in lengths]
        _lib.check(lib.b2s_hubert_forward_train(C.byref(w), wave.data_ptr(), wave.stride(0), B, T0, c_len,
                                                ctx["saved"].data_ptr(), ctx["saved"].numel(), out.data_ptr(), reg,
                                                torch.cuda.current_stream().cuda_stream), "hubert_forward_train")
        ctx.update(wave=wave, B=B, T0=T0, pooled=pooled, draw=draw, c_len=c_len, n_valid=n_valid)
        self._train_ctx = ctx
        return out

    def _whisper_forward_train(self, input: torch.Tensor) -> torch.Tensor:
        """(B, mel_bins, 2*max_positions) log-mel features -> fp32 (B, pooled, llm_dim), activations kept."""
        w = self.pack_weights()[0]
        mel = input.to(torch.float32).contiguous()
        if mel.dim() != 3 or mel.shape[1] != w.mel_bins:
            raise ValueError(f"expected (B, {w.mel_bins}, T) log-mel features")
        B, _, T = mel.shape
        if T != 2 * w.max_positions:
            raise ValueError(f"Whisper expects the mel input features to be of length {2 * w.max_positions}, but "
                             f"found {T}. Make sure to pad the input mel features to {2 * w.max_positions}.")
        pooled = (w.max_positions - w.pool_kernel) // w.pool_stride + 1
        lib = _lib.load()
        nbytes = lib.b2s_whisper_saved_bytes(C.byref(w), B)
        ctx = self._train_ctx
        if ctx is None or ctx["saved"].numel() < nbytes:
            ctx = {"saved": torch.empty(nbytes, device=mel.device, dtype=torch.uint8)}
        out = torch.empty(B, pooled, w.llm_dim, device=mel.device, dtype=torch.float32)
        _lib.check(lib.b2s_whisper_forward_train(C.byref(w), mel.data_ptr(), B, T, ctx["saved"].data_ptr(),
                                                 ctx["saved"].numel(), out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream), "whisper_forward_train")
        ctx.update(wave=mel, B=B, T0=T, pooled=pooled)
        self._train_ctx = ctx
        return out

    def backward(self, d_audio_embeds: torch.Tensor, layer_events=None) -> None:
        """Accumulate d(loss)/d(parameters) for the last `forward_train` batch given d(loss)/d(audio_embeds)
        (fp32 (B, A, llm_dim)). Gradients stay in packed accumulators until `flush_grads`.
        layer_events (optional list of torch.cuda.Event, one per transformer layer): event l is recorded on the current
        stream once every gradient of layer l has been enqueued."""
        ev_arr = None
        if layer_events is not None:
            assert len(layer_events) == self.encoder.arch.layers
            for e in layer_events:  # a torch event only owns a CUDA event once it has been recorded
                if not getattr(e, "_b2s_primed", False):
                    e.record()
                    e._b2s_primed = True
            ev_arr = (C.c_void_p * len(layer_events))(*[e.cuda_event for e in layer_events])
        ctx = self._train_ctx
        if ctx is None or "wave" not in ctx:
            raise RuntimeError("AudioEncoder.backward called without a preceding forward_train")
        w = self.pack_weights()[0]
        arch = self.encoder.arch
        g = self._grad_buffers()[0]
        d = d_audio_embeds.to(torch.float32).contiguous()
        assert d.shape == (ctx["B"], ctx["pooled"], w.llm_dim), "d_audio_embeds shape mismatch"
        if self.encoder_base == "whisper":
            lib = _lib.load()
            nbytes = lib.b2s_whisper_backward_workspace_bytes(C.byref(w), ctx["B"])
            if ctx.get("bws") is None or ctx["bws"].numel() < nbytes:
                ctx["bws"] = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
            _lib.check(lib.b2s_whisper_backward(C.byref(w), C.byref(g), ctx["B"], ctx["saved"].data_ptr(),
                                                ctx["saved"].numel(), d.data_ptr(), ctx["bws"].data_ptr(),
                                                ctx["bws"].numel(), ev_arr, torch.cuda.current_stream().cuda_stream),
                       "whisper_backward")
            del ctx["wave"]
            return
        if self._pos_w_dgrad is None:
            G, K_, cg = arch.pos_groups, arch.pos_k, arch.hidden // arch.pos_groups
            # conv transpose: taps reversed, each (out, in) block transposed: [g*cg+o][j][i] -> [g*cg+i][K-1-j][o]
            self._pos_w_dgrad = (self._pos_w_packed.view(G, cg, K_, cg).flip(2).permute(0, 3, 2, 1)
                                 .contiguous().view(arch.hidden, K_ * cg))
        lib = _lib.load()
        nbytes = lib.b2s_hubert_backward_workspace_bytes(C.byref(w), ctx["B"], ctx["T0"])
        if ctx.get("bws") is None or ctx["bws"].numel() < nbytes:
            ctx["bws"] = torch.empty(nbytes, device=d.device, dtype=torch.uint8)
        wave = ctx["wave"]
        reg = None
        if ctx.get("draw") is not None:
            reg = C.byref(ctx["draw"].c_struct(ctx["mse"], self._grad_buffers()[5]))
        if ctx.get("n_valid") is not None:  # ragged batch: the padding rows of the embeddings carry no gradient
            d = d.clone() if d.data_ptr() == d_audio_embeds.data_ptr() else d
            for b_, nv in enumerate(ctx["n_valid"]):
                d[b_, nv:].zero_()
        _lib.check(lib.b2s_hubert_backward(C.byref(w), self._pos_w_dgrad.data_ptr(), C.byref(g), wave.data_ptr(),
                                           wave.stride(0), ctx["B"], ctx["T0"], ctx.get("c_len"),
                                           ctx["saved"].data_ptr(),
                                           ctx["saved"].numel(), d.data_ptr(), ctx["bws"].data_ptr(),
                                           ctx["bws"].numel(), reg, ev_arr, torch.cuda.current_stream().cuda_stream),
                   "hubert_backward")
        del ctx["wave"]

    @torch.no_grad()
    def flush_grads(self) -> None:
        """Scratch accumulators -> `.grad` of the parameters (HF layouts; += like autograd), then zero the scratch.
        Pure re-indexing plus the weight-norm chain rule of the positional conv (once per optimizer step); buffers
        that alias `.grad` directly need nothing."""
        _, _, t, _, pending, _ = self._grad_buffers()
        enc = self.encoder
        arch = enc.arch
        H = arch.hidden

        def add(p, g):
            g = g.reshape(p.shape).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.add_(g)

        for name, ps in pending:
            row = 0
            for p in ps:
                n = p.shape[0]
                add(p, t[name][row:row + n])
                row += n
        if self.encoder_base == "whisper":
            add(enc.conv1.weight, t["conv1_w"].view(H, 3, arch.mel_bins).permute(0, 2, 1))
            add(enc.conv2.weight, t["conv2_w"].view(H, 3, H).permute(0, 2, 1))
            for l, lay in enumerate(enc.layers):  # fused bias gradient: q | (k_proj has no bias) | v
                add(lay.self_attn.q_proj.bias, t[f"l{l}.bqkv"][:H])
                add(lay.self_attn.v_proj.bias, t[f"l{l}.bqkv"][2 * H:])
            for buf in t.values():
                buf.zero_()
            return
        fe = enc.feature_extractor.conv_layers
        for i in range(6):
            k = arch.conv_kernel[i + 1]
            add(fe[i + 1].conv.weight, t[f"conv_w{i}"].view(512, k, 512).permute(0, 2, 1))
        pc = enc.encoder.pos_conv_embed.conv
        cg = H // arch.pos_groups
        dW = t["pos_w"].view(H, arch.pos_k, cg).permute(0, 2, 1)  # [H, cg, K] like original1
        g0 = pc.parametrizations.weight.original0.detach().float()  # [1, 1, K]
        v = pc.parametrizations.weight.original1.detach().float()   # [H, cg, K]
        nrm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
        dot = (dW * v).sum(dim=(0, 1), keepdim=True)
        add(pc.parametrizations.weight.original0, dot / nrm)
        add(pc.parametrizations.weight.original1, (g0 / nrm) * (dW - dot / (nrm * nrm) * v))
        for buf in t.values():
            buf.zero_()

    def forward(self, input, ctc_pool_ranges=None):
        """Same contract as REF/model/audio_encoder.py:56-88 (`pool` branch): (B, T0) -> (B, A, llm_dim).
        The result is returned in the encoder's 16-bit operand dtype -- fp16 by default, like the reference under
        autocast."""
        if self.downsample_method != "pool":
            raise Exception("Invalid downsampling method for audio encoder.")
        return self.forward_fp32(input).to(self.operand_dtype)
