"""One-time weight repacking (load time, torch tensor ops; never on the hot path).

Layouts consumed by the kernels:
  * every Linear weight stays in nn.Linear layout [out, in] (K-major), cast to bf16;
  * q/k/v projections are concatenated row-wise into one fused QKV weight;
  * gate/up projections are interleaved in 64-row blocks (64 gate rows | 64 up rows) so one GEMM tile holds
    both halves of SwiGLU for the same output columns;
  * Conv1d(C->C', k, stride) weights [C', C, k] become [C', k*C] (column = tap*C + c_in), the B operand of the
    strided-window implicit GEMM;
  * the positional conv is weight-normalised and repacked on device (ops.posconv_weight_pack).
"""
from __future__ import annotations

import torch


def pack_gate_up(w_gate: torch.Tensor, w_up: torch.Tensor, block: int = 64) -> torch.Tensor:
    F_, H = w_gate.shape
    assert w_up.shape == (F_, H) and F_ % block == 0
    g = w_gate.reshape(F_ // block, block, H)
    u = w_up.reshape(F_ // block, block, H)
    return torch.cat([g, u], dim=1).reshape(2 * F_, H).contiguous()


def pack_qkv(wq: torch.Tensor, wk: torch.Tensor, wv: torch.Tensor) -> torch.Tensor:
    return torch.cat([wq, wk, wv], dim=0).contiguous()


def pack_conv(w: torch.Tensor) -> torch.Tensor:
    """[C_out, C_in, k] -> [C_out, k*C_in]."""
    co, ci, k = w.shape
    return w.permute(0, 2, 1).reshape(co, k * ci).contiguous()


def rope_table(head_dim: int, max_pos: int, theta: float, scaling: dict | None = None,
               device: torch.device | str = "cpu") -> torch.Tensor:
    """fp32 [max_pos, head_dim] = cos[0:D/2] | sin[0:D/2].

    Default RoPE and the llama3 frequency scaling (TF/modeling_rope_utils.py:550-625): wavelengths above
    original_max/low_freq_factor are divided by `factor`, those below original_max/high_freq_factor are kept,
    and the band in between is interpolated smoothly.
    """
    import math
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    if scaling is not None and scaling.get("rope_type", scaling.get("type")) == "llama3":
        factor = float(scaling["factor"])
        low = float(scaling["low_freq_factor"])
        high = float(scaling["high_freq_factor"])
        old_ctx = float(scaling["original_max_position_embeddings"])
        low_wavelen = old_ctx / low
        high_wavelen = old_ctx / high
        wavelen = 2 * math.pi / inv_freq
        inv_llama = torch.where(wavelen > low_wavelen, inv_freq / factor, inv_freq)
        smooth = (old_ctx / wavelen - low) / (high - low)
        smoothed = (1 - smooth) * inv_llama / factor + smooth * inv_llama
        is_medium = ~(wavelen < high_wavelen) * ~(wavelen > low_wavelen)
        inv_freq = torch.where(is_medium, smoothed, inv_llama)
    pos = torch.arange(max_pos, dtype=torch.float32)
    ang = pos[:, None] * inv_freq[None, :].float()
    return torch.cat([ang.cos(), ang.sin()], dim=1).contiguous().to(device)
