"""Shared builders for the tests: product modules (CUDA path) and oracle inputs from the same seeded weights."""
import os
import sys
from types import SimpleNamespace as NS

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def ns_config(enc_cfg, llm_cfg):
    """A config namespace with the reference's yaml keys (REF/config/llama3_hubert.yaml) + architecture overrides."""
    return NS(
        seed_everything=1234,
        model=NS(
            audio_encoder=NS(base="hubert", type="facebook/hubert-large-ls960-ft", random_init=True,
                             downsample_method="pool",
                             downsample_factor=4, pooling=NS(kernel_size=enc_cfg.pool_kernel, stride=enc_cfg.pool_stride),
                             arch=NS(hidden=enc_cfg.hidden, layers=enc_cfg.layers, heads=enc_cfg.heads, ffn=enc_cfg.ffn,
                                     pos_k=enc_cfg.pos_k, pos_groups=enc_cfg.pos_groups)),
            llm_type=llm_cfg.llm_type, llm_embedding_channels=enc_cfg.llm_dim,
            llm_arch=NS(vocab=llm_cfg.vocab, hidden=llm_cfg.hidden, ffn=llm_cfg.ffn, layers=llm_cfg.layers,
                        heads=llm_cfg.heads, kv_heads=llm_cfg.kv_heads, head_dim=llm_cfg.head_dim,
                        rope_theta=llm_cfg.rope_theta, rope_scaling=llm_cfg.rope_scaling,
                        tie_embeddings=llm_cfg.tie_embeddings, max_pos=min(llm_cfg.max_pos, 4096), bos=llm_cfg.bos)),
        audio=NS(sampling_rate=16000),
        train=NS(use_ld_loss=True, use_fd_loss=True, ntp_loss_weight=0.5, ld_loss_weight=0.5, fd_loss_weight=1.0,
                 fd_loss_connector_layers=[0, 5, 11, 17, 23], grad_accum_interval=16, batch_size=1),
    )


def build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, device, dtype=None):
    """AudioEncoder + AudioLlamaForCausalLM of the CUDA path, loaded from reference-layout state dicts.
    dtype: 16-bit operand format of both modules (None = their default, fp16 like the reference's autocast)."""
    from llm_speech_summarization_b200.config import llm_arch_from_config
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    from llm_speech_summarization_b200.model.audio_llama import AudioLlamaForCausalLM
    cfg = ns_config(enc_cfg, llm_cfg)
    enc = AudioEncoder(cfg, device)
    if dtype is not None:
        enc.operand_dtype = dtype
    enc.load_state_dict(enc_sd, strict=True)
    enc.eval().to(device)
    llm = AudioLlamaForCausalLM(llm_arch_from_config(cfg), **({} if dtype is None else {"dtype": dtype}))
    # the synthetic weights are bf16-representable (oracle/configs.py); the copy into fp16 parameters is exact
    llm.load_state_dict({k: v.to(torch.bfloat16) for k, v in llm_sd.items()}, strict=True)
    llm.eval().to(device)
    return cfg, enc, llm


def bf16_round_sd(sd):
    """The reference holds the frozen LLM in half precision (REF/trainer.py:58-62); round the synthetic fp32 weights
    the same way so both sides see identical parameters."""
    return {k: v.to(torch.bfloat16).to(torch.float32) for k, v in sd.items()}


def ns_config_whisper(cfg, llm_type="meta-llama/Llama-3.2-3B-Instruct"):
    return NS(model=NS(audio_encoder=NS(base="whisper", type="openai/whisper-medium", random_init=True,
                                        downsample_method="pool",
                                        downsample_factor=4, pooling=NS(kernel_size=cfg.pool_kernel, stride=cfg.pool_stride),
                                        arch=NS(hidden=cfg.hidden, layers=cfg.layers, heads=cfg.heads, ffn=cfg.ffn,
                                                mel_bins=cfg.mel_bins, max_positions=cfg.max_positions)),
                       llm_type=llm_type, llm_embedding_channels=cfg.llm_dim),
              audio=NS(sampling_rate=16000))
