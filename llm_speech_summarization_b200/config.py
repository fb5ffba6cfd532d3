"""Config plumbing: the reference's yaml files (REF/config/*.yaml, same key names) parsed with pyyaml into an
attribute namespace (the reference uses OmegaConf: REF/train.py:21, REF/inference.py:159), plus the model
architectures the reference pulls from the HF hub, which are not reachable offline.

Optional, non-reference keys (all default to the published architectures):
    model.audio_encoder.arch: {hidden, layers, heads, ffn, pos_k, pos_groups}
    model.llm_arch: {vocab, hidden, ffn, layers, heads, kv_heads, head_dim, rope_theta, rope_scaling, ...}
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional, Tuple


class Namespace(SimpleNamespace):
    """Attribute access like OmegaConf's DictConfig, plus dict-style `in` / get."""

    def __contains__(self, k):
        return k in self.__dict__

    def get(self, k, default=None):
        return self.__dict__.get(k, default)


def to_namespace(obj):
    if isinstance(obj, dict):
        return Namespace(**{k: to_namespace(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_namespace(v) for v in obj]
    return obj


def _yaml_loader():
    """SafeLoader with YAML 1.2 float resolution: PyYAML (YAML 1.1) reads `lr: 5e-5` -- the spelling of every shipped
    reference yaml (REF/config/*.yaml, train.optimizer.lr) -- as the STRING '5e-5' because it has no dot; OmegaConf, which
    the reference uses (REF/train.py:21), reads a float."""
    import re
    import yaml

    class Loader(yaml.SafeLoader):
        pass

    Loader.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                       |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                       |\.[0-9_]+(?:[eE][-+]?[0-9]+)?
                       |[-+]?\.(?:inf|Inf|INF)
                       |\.(?:nan|NaN|NAN))$""", re.X),
        list("-+0123456789."))
    return Loader


def load_config(path: str) -> Namespace:
    import yaml
    with open(path) as f:
        return to_namespace(yaml.load(f, Loader=_yaml_loader()))


@dataclass
class EncoderArch:
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
    # train-mode regularisers of facebook/hubert-large-ls960-ft; only read when the training
    # step runs with regularisers enabled (llm_speech_summarization_b200/regularizers.py)
    feat_proj_dropout: float = 0.1
    hidden_dropout: float = 0.1
    attention_dropout: float = 0.1
    activation_dropout: float = 0.1
    layerdrop: float = 0.1
    apply_spec_augment: bool = True
    mask_time_prob: float = 0.05
    mask_time_length: int = 10
    mask_time_min_masks: int = 2


@dataclass
class WhisperArch:
    """openai/whisper-medium encoder."""
    hidden: int = 1024
    layers: int = 24
    heads: int = 16
    ffn: int = 4096
    mel_bins: int = 80
    max_positions: int = 1500
    ln_eps: float = 1e-5


@dataclass
class LlmArch:
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
    max_pos: int = 8192  # RoPE table rows built up front (prompts on this path are < 1k tokens)
    bos: int = 128000
    eos: Tuple[int, ...] = (128001, 128008, 128009)


KNOWN_LLMS = {
    "meta-llama/Llama-3.2-3B-Instruct": LlmArch(),
    "GeneZC/MiniChat-2-3B": LlmArch(vocab=49216, hidden=3072, ffn=8192, layers=24, heads=24, kv_heads=24,
                                    head_dim=128, rope_theta=10000.0, rope_scaling=None, tie_embeddings=False,
                                    max_pos=4096, bos=1, eos=(2,)),
}


def _override(arch, overrides):
    if overrides is None:
        return arch
    d = overrides.__dict__ if isinstance(overrides, SimpleNamespace) else dict(overrides)
    for k, v in d.items():
        if not hasattr(arch, k):
            raise KeyError(f"unknown architecture key {k!r}")
        if isinstance(v, SimpleNamespace):
            v = dict(v.__dict__)
        if isinstance(getattr(arch, k), tuple) and isinstance(v, list):
            v = tuple(v)
        setattr(arch, k, v)
    return arch


def encoder_arch_from_config(config) -> EncoderArch:
    ae = config.model.audio_encoder
    return _override(EncoderArch(), getattr(ae, "arch", None))


def whisper_arch_from_config(config) -> WhisperArch:
    ae = config.model.audio_encoder
    return _override(WhisperArch(), getattr(ae, "arch", None))


def llm_arch_from_config(config) -> LlmArch:
    import copy
    llm_type = config.model.llm_type
    if llm_type not in KNOWN_LLMS:
        raise Exception("Unknown LLM type.")  # REF/utils.py:57,102
    return _override(copy.deepcopy(KNOWN_LLMS[llm_type]), getattr(config.model, "llm_arch", None))


def arch_from_oracle_cfg(cfg, cls):
    """Build an EncoderArch / LlmArch from the oracle's dataclass of the same field names (tests only)."""
    fields = cls.__dataclass_fields__
    return cls(**{k: getattr(cfg, k) for k in fields if hasattr(cfg, k)})
