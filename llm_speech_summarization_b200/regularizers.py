"""Host side of the train-mode regularisers of the HuBERT encoder.

The reference trains with `self.audio_encoder.train()` (REF/trainer.py:258), which switches on, inside HF's
HubertModel: nn.Dropout at five kinds of sites, LayerDrop (one `torch.rand([])` per layer on the host,
TF/models/hubert/modeling_hubert.py:596-599) and SpecAugment time masking (`_compute_mask_indices`, numpy on the host,
:842-886). Here the host draws exactly what the reference draws on the host -- the per-layer skip decisions and the
frame mask -- plus ONE 64-bit seed per micro-batch; every elementwise keep/drop decision is then a pure function of
(seed, site, element index) evaluated inside the CUDA kernels (csrc/rng.cuh), identically in forward and backward.

Whisper-medium has dropout = 0 and encoder_layerdrop = 0: nothing to do there.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib


@dataclass
class RegularizerConfig:
    feat_proj_dropout: float = 0.1
    hidden_dropout: float = 0.1
    attention_dropout: float = 0.1
    activation_dropout: float = 0.1
    layerdrop: float = 0.1
    apply_spec_augment: bool = True
    mask_time_prob: float = 0.05
    mask_time_length: int = 10
    mask_time_min_masks: int = 2

    @classmethod
    def from_arch(cls, arch) -> "RegularizerConfig":
        return cls(**{k: getattr(arch, k) for k in cls.__dataclass_fields__ if hasattr(arch, k)})


def compute_time_mask(batch: int, frames: int, mask_prob: float, mask_length: int, min_masks: int,
                      rng: np.random.Generator) -> np.ndarray:
    """SpecAugment span mask over time, (batch, frames) bool: the sampling scheme of HF `_compute_mask_indices`
    (TF/models/hubert/modeling_hubert.py: number of spans = int(mask_prob * frames / mask_length + eps) with one
    uniform eps per call, at least `min_masks`, span starts drawn without replacement among the positions where a
    whole span fits, spans may overlap). No padding mask: the reference feeds un-padded batch-1 audio."""
    if mask_length < 1:
        raise ValueError("`mask_length` has to be bigger than 0.")
    if mask_length > frames:
        raise ValueError(f"`mask_length` has to be smaller than `sequence_length`, but got `mask_length`: "
                         f"{mask_length} and `sequence_length`: {frames}`")
    eps = float(rng.random())
    spans = max(int(mask_prob * frames / mask_length + eps), min_masks)
    spans = min(spans, frames // mask_length)            # never more masked frames than frames
    spans = max(min(spans, frames - (mask_length - 1)), 0)  # distinct starts must exist
    mask = np.zeros((batch, frames), dtype=bool)
    if spans == 0:
        return mask
    offsets = np.arange(mask_length)
    for b in range(batch):
        starts = rng.choice(frames - (mask_length - 1), size=spans, replace=False)
        mask[b, (starts[:, None] + offsets[None, :]).reshape(-1)] = True
    return mask


@dataclass
class RegularizerDraw:
    """What one micro-batch's forward AND backward share."""
    seed: int
    layer_skip: np.ndarray             # uint8 [layers], host
    time_mask: Optional[torch.Tensor]  # uint8 [batch * frames], device (None = no SpecAugment)
    cfg: RegularizerConfig

    def c_struct(self, masked_spec_embed: torch.Tensor, g_masked_spec_embed: Optional[torch.Tensor]):
        r = _lib.EncoderRegularizers()
        r.seed = self.seed
        r.p_feat_proj = self.cfg.feat_proj_dropout
        r.p_hidden = self.cfg.hidden_dropout
        r.p_attention = self.cfg.attention_dropout
        r.p_activation = self.cfg.activation_dropout
        r.layer_skip = self.layer_skip.ctypes.data_as(C.c_void_p)
        r.time_mask = None if self.time_mask is None else self.time_mask.data_ptr()
        r.masked_spec_embed = masked_spec_embed.data_ptr()
        r.g_masked_spec_embed = None if g_masked_spec_embed is None else g_masked_spec_embed.data_ptr()
        return r


def draw(cfg: RegularizerConfig, batch: int, frames: int, layers: int, device,
         generator: Optional[torch.Generator] = None, valid_frames=None) -> RegularizerDraw:
    """Draw the host-side randomness of one micro-batch from a torch CPU generator (None = torch's global one, which
    is what the reference's LayerDrop consumes). valid_frames (ragged batch): frames of each utterance on its own --
    its SpecAugment spans are counted and placed inside ITS length, like HF's per-`input_length` loop, instead of
    landing in the padding."""
    seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator).item())
    skip = (torch.rand(layers, generator=generator) < cfg.layerdrop).to(torch.uint8).numpy().copy()
    time_mask = None
    if cfg.apply_spec_augment and cfg.mask_time_prob > 0:
        np_seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator).item())
        rng = np.random.default_rng(np_seed)
        if valid_frames is None:
            m = compute_time_mask(batch, frames, cfg.mask_time_prob, cfg.mask_time_length, cfg.mask_time_min_masks, rng)
        else:
            m = np.zeros((batch, frames), dtype=bool)
            for b, n in enumerate(valid_frames):
                n = int(n)
                if n >= cfg.mask_time_length:
                    m[b, :n] = compute_time_mask(1, n, cfg.mask_time_prob, cfg.mask_time_length,
                                                 cfg.mask_time_min_masks, rng)[0]
        time_mask = torch.from_numpy(m.reshape(-1).astype(np.uint8)).to(device)
    return RegularizerDraw(seed=seed, layer_skip=skip, time_mask=time_mask, cfg=cfg)
