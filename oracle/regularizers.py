"""Train-mode regularisers for the CPU oracle: a numpy restatement of the counter-based keep/drop generator of
llm_speech_summarization_b200/csrc/rng.cuh, so the oracle can run HF's train-mode HuBERT math
(TF/models/hubert/modeling_hubert.py:223-230,254,351-368,383-393,585-599,842-886) under EXACTLY the masks the CUDA path
uses. TEST INFRASTRUCTURE (see oracle/__init__.py).

    mix(x, k)   : x ^= x >> 16; x = x * 0x7feb352d + k; x ^= x >> 15; x *= 0x846ca68b; x ^= x >> 16      (uint32)
    stream key  : t = mix(site ^ seed_lo, seed_hi); t = mix(t ^ a, 0x9E3779B9); t = mix(t ^ b, 0x85EBCA6B);
                  k1 = t; k2 = mix(t ^ seed_hi, seed_lo)
    kept  <=>  mix(elem ^ k1, k2) >= floor(p * 2^32)
    elem        : row * N + col (elementwise sites), (query << 16) | key (attention, a = sequence, b = head)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

SITE_FEAT_PROJ = 1
SITE_POS_ADD = 2
SITE_LAYER0 = 16


def site_attn_out(l): return SITE_LAYER0 + 4 * l
def site_ff_act(l): return SITE_LAYER0 + 4 * l + 1
def site_ff_out(l): return SITE_LAYER0 + 4 * l + 2
def site_attn_prob(l): return SITE_LAYER0 + 4 * l + 3


_M32 = np.uint64(0xFFFFFFFF)


def mix(x, k):
    """uint32 arithmetic carried in uint64 lanes (numpy would warn on uint32 overflow)."""
    x = np.asarray(x, dtype=np.uint64) & _M32
    k = np.uint64(int(k) & 0xFFFFFFFF)
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(0x7FEB352D) + k) & _M32
    x = x ^ (x >> np.uint64(15))
    x = (x * np.uint64(0x846CA68B)) & _M32
    x = x ^ (x >> np.uint64(16))
    return x


def stream_key(seed: int, site: int, a: int = 0, b: int = 0):
    s0, s1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    t = int(mix((site ^ s0) & 0xFFFFFFFF, s1))
    t = int(mix(t ^ a, 0x9E3779B9))
    t = int(mix(t ^ b, 0x85EBCA6B))
    return t, int(mix(t ^ s1, s0))


def threshold(p: float) -> int:
    if not p > 0:
        return 0
    return min(int(float(np.float32(p)) * 4294967296.0), 4294967295)


def keep(elems: np.ndarray, seed: int, site: int, a: int, b: int, p: float) -> np.ndarray:
    k1, k2 = stream_key(seed, site, a, b)
    h = mix(np.asarray(elems, dtype=np.uint64) ^ np.uint64(k1), k2)
    return h >= np.uint64(threshold(p))


def elementwise_multiplier(seed: int, site: int, p: float, rows: int, cols: int) -> torch.Tensor:
    """[rows, cols] float: 1/(1-p) where kept, 0 where dropped (element index = row * cols + col)."""
    if not p > 0:
        return torch.ones(rows, cols)
    e = np.arange(rows * cols, dtype=np.uint64)
    m = keep(e, seed, site, 0, 0, p).reshape(rows, cols)
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy(m.astype(np.float32) * inv)


def attention_multiplier(seed: int, layer: int, p: float, batch: int, heads: int, n: int) -> torch.Tensor:
    """[batch, heads, n, n] float multiplier of the softmax probabilities (query-major)."""
    if not p > 0:
        return torch.ones(batch, heads, n, n)
    q = np.arange(n, dtype=np.uint64)[:, None]
    k = np.arange(n, dtype=np.uint64)[None, :]
    e = (q << np.uint64(16)) | k
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    out = np.empty((batch, heads, n, n), dtype=np.float32)
    for b in range(batch):
        for h in range(heads):
            out[b, h] = keep(e, seed, site_attn_prob(layer), b, h, p).astype(np.float32) * inv
    return torch.from_numpy(out)


@dataclass
class OracleRegularizers:
    seed: int
    p_feat_proj: float = 0.1
    p_hidden: float = 0.1
    p_attention: float = 0.1
    p_activation: float = 0.1
    layer_skip: Optional[np.ndarray] = None   # [layers] 0/1
    time_mask: Optional[np.ndarray] = None    # [batch, frames] bool

    def skipped(self, l: int) -> bool:
        return self.layer_skip is not None and bool(self.layer_skip[l])
