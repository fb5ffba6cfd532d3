"""Architectures and deterministic synthetic weights / inputs for the oracle and tests.

TEST INFRASTRUCTURE (see oracle/__init__.py). Key names follow the reference's checkpoint layout:
AudioEncoder.state_dict() (`encoder.*` HF HubertModel names + `embed_projection.*`) and
HF LlamaForCausalLM names for the LLM (`model.embed_tokens.weight`, `model.layers.N.*`, `model.norm.weight`,
`lm_head.weight`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional, Tuple

import torch


@dataclass
class EncoderCfg:
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    ffn: int = 4096
    conv_dim: Tuple[int, ...] = (512,) * 7
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    pos_k: int = 128
    pos_groups: int = 16
    ln_eps: float = 1e-5
    pool_kernel: int = 8
    pool_stride: int = 4
    llm_dim: int = 3072


@dataclass
class LlmCfg:
    vocab: int = 128256
    hidden: int = 3072
    ffn: int = 8192
    layers: int = 28
    heads: int = 24
    kv_heads: int = 8
    head_dim: int = 128
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = field(default_factory=lambda: dict(
        rope_type="llama3", factor=32.0, high_freq_factor=4.0, low_freq_factor=1.0,
        original_max_position_embeddings=8192))
    tie_embeddings: bool = True
    max_pos: int = 131072
    bos: int = 128000
    llm_type: str = "meta-llama/Llama-3.2-3B-Instruct"


HUBERT_LARGE = EncoderCfg()
LLAMA32_3B = LlmCfg()
MINICHAT_3B = LlmCfg(vocab=49216, hidden=3072, ffn=8192, layers=24, heads=24, kv_heads=24, head_dim=128,
                     rope_theta=10000.0, rope_scaling=None, tie_embeddings=False, max_pos=4096, bos=1,
                     llm_type="GeneZC/MiniChat-2-3B")

# small shapes that still satisfy the CUDA kernels' structural constraints (conv width 512, head_dim 64 / 128,
# 64 channels per positional-conv group, widths multiple of 256)
TINY_ENCODER = EncoderCfg(hidden=256, layers=2, heads=4, ffn=512, pos_k=16, pos_groups=4, llm_dim=256)
TINY_LLAMA = LlmCfg(vocab=1024, hidden=256, ffn=512, layers=3, heads=2, kv_heads=1, head_dim=128, max_pos=16384,
                    bos=1000)
TINY_MINICHAT = LlmCfg(vocab=1000, hidden=256, ffn=512, layers=2, heads=2, kv_heads=2, head_dim=128,
                       rope_theta=10000.0, rope_scaling=None, tie_embeddings=False, max_pos=2048, bos=1,
                       llm_type="GeneZC/MiniChat-2-3B")


@dataclass
class WhisperCfg:
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    ffn: int = 4096
    mel_bins: int = 80
    max_positions: int = 1500
    ln_eps: float = 1e-5
    pool_kernel: int = 8
    pool_stride: int = 4
    llm_dim: int = 3072


WHISPER_MEDIUM = WhisperCfg()
TINY_WHISPER = WhisperCfg(hidden=256, layers=2, heads=4, ffn=512, max_positions=100, llm_dim=256)


def make_whisper_state_dict(cfg: WhisperCfg, seed: int = 2468) -> Dict[str, torch.Tensor]:
    """Synthetic AudioEncoder(base=whisper) weights under the reference's checkpoint names (`encoder.` = HF
    WhisperEncoder, TF/models/whisper/modeling_whisper.py:541-647; k_proj has no bias)."""
    g = torch.Generator().manual_seed(seed)
    H = cfg.hidden
    sd: Dict[str, torch.Tensor] = {}
    sd["encoder.conv1.weight"] = _randn(g, H, cfg.mel_bins, 3, std=math.sqrt(1.0 / (3 * cfg.mel_bins)))
    sd["encoder.conv1.bias"] = _randn(g, H, std=0.05)
    sd["encoder.conv2.weight"] = _randn(g, H, H, 3, std=math.sqrt(1.0 / (3 * H)))
    sd["encoder.conv2.bias"] = _randn(g, H, std=0.05)
    sd["encoder.embed_positions.weight"] = _randn(g, cfg.max_positions, H, std=0.3)
    for l in range(cfg.layers):
        p = f"encoder.layers.{l}."
        sd[p + "self_attn.k_proj.weight"] = _randn(g, H, H, std=0.02)
        for n in ("v_proj", "q_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = _randn(g, H, H, std=0.02)
            sd[p + f"self_attn.{n}.bias"] = _randn(g, H, std=0.02)
        sd[p + "self_attn_layer_norm.weight"] = _randn(g, H, std=0.1, mean=1.0)
        sd[p + "self_attn_layer_norm.bias"] = _randn(g, H, std=0.1)
        sd[p + "fc1.weight"] = _randn(g, cfg.ffn, H, std=0.02)
        sd[p + "fc1.bias"] = _randn(g, cfg.ffn, std=0.02)
        sd[p + "fc2.weight"] = _randn(g, H, cfg.ffn, std=0.02)
        sd[p + "fc2.bias"] = _randn(g, H, std=0.02)
        sd[p + "final_layer_norm.weight"] = _randn(g, H, std=0.1, mean=1.0)
        sd[p + "final_layer_norm.bias"] = _randn(g, H, std=0.1)
    sd["encoder.layer_norm.weight"] = _randn(g, H, std=0.1, mean=1.0)
    sd["encoder.layer_norm.bias"] = _randn(g, H, std=0.1)
    sd["embed_projection.weight"] = _randn(g, cfg.llm_dim, H, std=0.02)
    sd["embed_projection.bias"] = _randn(g, cfg.llm_dim, std=0.02)
    return sd


def synthetic_log_mel(cfg: WhisperCfg, index: int, batch: int = 1) -> torch.Tensor:
    """Synthetic log-mel in the extractor's output range: randn*0.5 clamped to [-1, 1.5]."""
    g = torch.Generator().manual_seed(4321 + index)
    return (torch.randn(batch, cfg.mel_bins, 2 * cfg.max_positions, generator=g) * 0.5).clamp(-1.0, 1.5)


def _randn(gen: torch.Generator, *shape, std: float = 1.0, mean: float = 0.0) -> torch.Tensor:
    return torch.randn(*shape, generator=gen, dtype=torch.float32) * std + mean


def make_encoder_state_dict(cfg: EncoderCfg, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Synthetic AudioEncoder weights. Distributions follow HF's HubertPreTrainedModel._init_weights
    (TF/models/hubert/modeling_hubert.py:640-673: Linear N(0, 0.02), Conv1d kaiming-normal, positional conv
    N(0, 2*sqrt(1/(k*C_in)))) but biases and LayerNorm affine terms are perturbed so every term is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    H = cfg.hidden
    sd["encoder.masked_spec_embed"] = torch.rand(H, generator=g)
    cin = 1
    for i, (co, k) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
        p = f"encoder.feature_extractor.conv_layers.{i}."
        sd[p + "conv.weight"] = _randn(g, co, cin, k, std=math.sqrt(2.0 / (cin * k)))
        sd[p + "conv.bias"] = _randn(g, co, std=0.05)
        sd[p + "layer_norm.weight"] = _randn(g, co, std=0.1, mean=1.0)
        sd[p + "layer_norm.bias"] = _randn(g, co, std=0.1)
        cin = co
    sd["encoder.feature_projection.layer_norm.weight"] = _randn(g, cin, std=0.1, mean=1.0)
    sd["encoder.feature_projection.layer_norm.bias"] = _randn(g, cin, std=0.1)
    sd["encoder.feature_projection.projection.weight"] = _randn(g, H, cin, std=0.02)
    sd["encoder.feature_projection.projection.bias"] = _randn(g, H, std=0.02)
    cg = H // cfg.pos_groups
    pc = "encoder.encoder.pos_conv_embed.conv."
    sd[pc + "bias"] = _randn(g, H, std=0.02)
    v = _randn(g, H, cg, cfg.pos_k, std=2 * math.sqrt(1.0 / (cfg.pos_k * H)))
    sd[pc + "parametrizations.weight.original0"] = v.norm(dim=(0, 1), keepdim=True) * (
        1.0 + 0.1 * torch.randn(1, 1, cfg.pos_k, generator=g))
    sd[pc + "parametrizations.weight.original1"] = v
    sd["encoder.encoder.layer_norm.weight"] = _randn(g, H, std=0.1, mean=1.0)
    sd["encoder.encoder.layer_norm.bias"] = _randn(g, H, std=0.1)
    for l in range(cfg.layers):
        p = f"encoder.encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"attention.{n}.weight"] = _randn(g, H, H, std=0.02)
            sd[p + f"attention.{n}.bias"] = _randn(g, H, std=0.02)
        sd[p + "layer_norm.weight"] = _randn(g, H, std=0.1, mean=1.0)
        sd[p + "layer_norm.bias"] = _randn(g, H, std=0.1)
        sd[p + "feed_forward.intermediate_dense.weight"] = _randn(g, cfg.ffn, H, std=0.02)
        sd[p + "feed_forward.intermediate_dense.bias"] = _randn(g, cfg.ffn, std=0.02)
        sd[p + "feed_forward.output_dense.weight"] = _randn(g, H, cfg.ffn, std=0.02)
        sd[p + "feed_forward.output_dense.bias"] = _randn(g, H, std=0.02)
        sd[p + "final_layer_norm.weight"] = _randn(g, H, std=0.1, mean=1.0)
        sd[p + "final_layer_norm.bias"] = _randn(g, H, std=0.1)
    sd["embed_projection.weight"] = _randn(g, cfg.llm_dim, H, std=0.02)
    sd["embed_projection.bias"] = _randn(g, cfg.llm_dim, std=0.02)
    return sd


def make_llm_state_dict(cfg: LlmCfg, seed: int = 4321, dtype: torch.dtype = torch.float32,
                        resid_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Synthetic LlamaForCausalLM weights, N(0, 0.02) like HF's default init; RMSNorm weights perturbed.
    resid_scale multiplies the std of the residual-output projections (o_proj, down_proj): 1.0 = HF default init;
    1/sqrt(2*layers) = the GPT-2 / Megatron depth-scaled init, which keeps the residual stream well conditioned
    (HF-default init at hidden 3072 has per-branch gain > 1 and amplifies ANY rounding difference chaotically)."""
    g = torch.Generator().manual_seed(seed)
    H, D = cfg.hidden, cfg.head_dim
    sd: Dict[str, torch.Tensor] = {}
    sd["model.embed_tokens.weight"] = _randn(g, cfg.vocab, H, std=0.02).to(dtype)
    for l in range(cfg.layers):
        p = f"model.layers.{l}."
        sd[p + "input_layernorm.weight"] = _randn(g, H, std=0.1, mean=1.0).to(dtype)
        sd[p + "self_attn.q_proj.weight"] = _randn(g, cfg.heads * D, H, std=0.02).to(dtype)
        sd[p + "self_attn.k_proj.weight"] = _randn(g, cfg.kv_heads * D, H, std=0.02).to(dtype)
        sd[p + "self_attn.v_proj.weight"] = _randn(g, cfg.kv_heads * D, H, std=0.02).to(dtype)
        sd[p + "self_attn.o_proj.weight"] = _randn(g, H, cfg.heads * D, std=0.02 * resid_scale).to(dtype)
        sd[p + "post_attention_layernorm.weight"] = _randn(g, H, std=0.1, mean=1.0).to(dtype)
        sd[p + "mlp.gate_proj.weight"] = _randn(g, cfg.ffn, H, std=0.02).to(dtype)
        sd[p + "mlp.up_proj.weight"] = _randn(g, cfg.ffn, H, std=0.02).to(dtype)
        sd[p + "mlp.down_proj.weight"] = _randn(g, H, cfg.ffn, std=0.02 * resid_scale).to(dtype)
    sd["model.norm.weight"] = _randn(g, H, std=0.1, mean=1.0).to(dtype)
    if cfg.tie_embeddings:
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    else:
        sd["lm_head.weight"] = _randn(g, cfg.vocab, H, std=0.02).to(dtype)
    return sd


@dataclass
class StubTokenizer:
    """Fixed-id tokenizer standing in for AutoTokenizer (tokenizer files are unreachable offline, SURVEY.md
    appendix A). Called like the reference calls it: tokenizer(text, return_tensors="pt").input_ids."""
    prefix_ids: List[int]
    suffix_ids: List[int]
    prefix_text: str
    suffix_text: str
    extra: Dict[str, List[int]] = field(default_factory=dict)

    class _Out:
        def __init__(self, ids):
            self.input_ids = torch.tensor([ids], dtype=torch.long)

    def __call__(self, text, return_tensors="pt"):
        if text == self.prefix_text:
            return self._Out(self.prefix_ids)
        if text == self.suffix_text:
            return self._Out(self.suffix_ids)
        if text in self.extra:
            return self._Out(self.extra[text])
        raise KeyError(f"StubTokenizer has no ids for {text!r}")


def stub_tokenizer(cfg: LlmCfg, seed: int = 99) -> StubTokenizer:
    """P = 9 / S = 6 ids for the Llama-3 template, P = 6 / S = 8 for MiniChat, first id BOS."""
    from .reference_math import prompt_strings
    g = torch.Generator().manual_seed(seed)
    P, S = (9, 6) if "llama" in cfg.llm_type.lower() else (6, 8)
    hi = max(2, cfg.vocab - 256)
    prefix = [cfg.bos % cfg.vocab] + torch.randint(0, hi, (P - 1,), generator=g).tolist()
    suffix = [cfg.bos % cfg.vocab] + torch.randint(0, hi, (S - 1,), generator=g).tolist()
    pre, suf = prompt_strings(cfg.llm_type)
    return StubTokenizer(prefix, suffix, pre, suf)


def synthetic_utterance(cfg: LlmCfg, index: int, samples: int, T: int = 40, R: int = 64):
    """Seeded synthetic inputs for utterance `index`: raw audio randn*0.1 (no normalisation, the reference feeds
    raw floats: REF/trainer.py:146-149), transcript ids (T) and response ids (R), both WITHOUT the leading BOS the
    collate already stripped (REF/trainer.py:155-156)."""
    g = torch.Generator().manual_seed(1234 + index)
    audio = torch.randn(samples, generator=g) * 0.1
    hi = max(2, cfg.vocab - 256)
    text_ids = torch.randint(0, hi, (T,), generator=g)
    resp_ids = torch.randint(0, hi, (R,), generator=g)
    return audio, text_ids, resp_ids


def cfg_dict(cfg) -> dict:
    return asdict(cfg)
