"""CPU restatement of the reference's audio-prompt forward / loss path in plain torch.

TEST INFRASTRUCTURE (see oracle/__init__.py) -- checker and timed CPU baseline only.

REF/ = /root/reference, TF/ = the transformers package the reference delegates its arithmetic to
(pinned 4.47.0, REF/requirements.txt:15). Every function names the lines it follows. Everything runs in the dtype
of the tensors it is given (fp32 for the baseline, fp64 for tight checks); eval mode by default, HF's train mode
(dropout sites, LayerDrop, SpecAugment) when a `reg` block with explicit masks is passed to the HuBERT functions.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .configs import EncoderCfg, LlmCfg

# ----------------------------------------------------------------------------------------------------------
# prompt templates (REF/utils.py:6-10)
SYSTEM_PROMPT = ""
MINICHAT_PROMPT_PREFIX = f"{SYSTEM_PROMPT}[|User|]"
MINICHAT_PROMPT_SUFFIX = "</s>[|Assistant|]"
LLAMA_PROMPT_PREFIX = (f"<|start_header_id|>system<|end_header_id|>{SYSTEM_PROMPT}<|eot_id|>"
                       "<|start_header_id|>user<|end_header_id|>\n\n")
LLAMA_PROMPT_SUFFIX = "<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n"


def prompt_strings(llm_type: str) -> Tuple[str, str]:
    """REF/utils.py:50-57,95-102."""
    if llm_type == "GeneZC/MiniChat-2-3B":
        return MINICHAT_PROMPT_PREFIX, MINICHAT_PROMPT_SUFFIX
    if llm_type == "meta-llama/Llama-3.2-3B-Instruct":
        return LLAMA_PROMPT_PREFIX, LLAMA_PROMPT_SUFFIX
    raise Exception("Unknown LLM type.")


# ----------------------------------------------------------------------------------------------------------
# HuBERT (third-party: transformers HubertModel)
def hubert_feature_extractor(sd: Dict[str, torch.Tensor], wave: torch.Tensor, cfg: EncoderCfg) -> torch.Tensor:
    """7 x [Conv1d -> LayerNorm over channels -> erf-GELU]; (B, T0) -> (B, C, N).
    TF/models/hubert/modeling_hubert.py:127-151 (HubertLayerNormConvLayer), :178-213 (HubertFeatureEncoder)."""
    h = wave[:, None, :]
    for i, s in enumerate(cfg.conv_stride):
        p = f"encoder.feature_extractor.conv_layers.{i}."
        h = F.conv1d(h, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=s)
        h = h.transpose(-2, -1)
        h = F.layer_norm(h, (h.shape[-1],), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)
        h = h.transpose(-2, -1)
        h = F.gelu(h)
    return h


def _pos_conv_weight(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """weight_norm(dim=2): w = g * v / ||v|| over dims (0,1), one norm per tap
    (TF/models/hubert/modeling_hubert.py:60-78). Accepts both checkpoint spellings."""
    pc = "encoder.encoder.pos_conv_embed.conv."
    if pc + "parametrizations.weight.original0" in sd:
        g, v = sd[pc + "parametrizations.weight.original0"], sd[pc + "parametrizations.weight.original1"]
    else:
        g, v = sd[pc + "weight_g"], sd[pc + "weight_v"]
    return g * v / v.norm(dim=(0, 1), keepdim=True)


def hubert_encoder(sd: Dict[str, torch.Tensor], feats: torch.Tensor, cfg: EncoderCfg,
                   return_pre_norm: bool = False, reg=None) -> torch.Tensor:
    """Feature projection, positional conv embedding and the stable-layer-norm transformer stack; (B, N, 512) ->
    (B, N, H). TF/models/hubert/modeling_hubert.py:216-231 (projection), :45-103 (pos conv + same-pad),
    :262-345 (attention, scaling head_dim**-0.5, no mask), :348-369 (FFN), :505-548 (layer), :563-624 (stack).
    `reg` (oracle.regularizers.OracleRegularizers) switches on the train-mode regularisers with explicit masks:
    dropout after the projection (:230), SpecAugment (:842-886), dropout after the positional add (:587), LayerDrop
    (:596-599), attention-probability dropout (:254), dropout after the attention block (:536), after the FFN
    activation (:365) and after the FFN output (:368)."""
    H, nh = cfg.hidden, cfg.heads
    hd = H // nh
    Bn, Nn = feats.shape[0], feats.shape[1]
    if reg is not None:
        from . import regularizers as rg

        def drop(t, site, p):
            return t * rg.elementwise_multiplier(reg.seed, site, p, Bn * Nn, t.shape[-1]).view(t.shape)
    else:
        def drop(t, site, p):
            return t
    x = F.layer_norm(feats, (feats.shape[-1],), sd["encoder.feature_projection.layer_norm.weight"],
                     sd["encoder.feature_projection.layer_norm.bias"], cfg.ln_eps)
    x = F.linear(x, sd["encoder.feature_projection.projection.weight"],
                 sd["encoder.feature_projection.projection.bias"])
    if reg is not None:
        x = drop(x, rg.SITE_FEAT_PROJ, reg.p_feat_proj)
        if reg.time_mask is not None:
            tm = torch.from_numpy(reg.time_mask.astype(bool)).view(Bn, Nn, 1)
            x = torch.where(tm, sd["encoder.masked_spec_embed"].view(1, 1, -1).to(x.dtype), x)
    pos = F.conv1d(x.transpose(1, 2), _pos_conv_weight(sd), sd["encoder.encoder.pos_conv_embed.conv.bias"],
                   padding=cfg.pos_k // 2, groups=cfg.pos_groups)
    if cfg.pos_k % 2 == 0:
        pos = pos[:, :, :-1]
    x = x + F.gelu(pos).transpose(1, 2)
    if reg is not None:
        x = drop(x, rg.SITE_POS_ADD, reg.p_hidden)
    B, N, _ = x.shape
    for l in range(cfg.layers):
        if reg is not None and reg.skipped(l):
            continue
        p = f"encoder.encoder.layers.{l}."
        y = F.layer_norm(x, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], cfg.ln_eps)
        q = F.linear(y, sd[p + "attention.q_proj.weight"], sd[p + "attention.q_proj.bias"])
        k = F.linear(y, sd[p + "attention.k_proj.weight"], sd[p + "attention.k_proj.bias"])
        v = F.linear(y, sd[p + "attention.v_proj.weight"], sd[p + "attention.v_proj.bias"])
        q = q.view(B, N, nh, hd).transpose(1, 2)
        k = k.view(B, N, nh, hd).transpose(1, 2)
        v = v.view(B, N, nh, hd).transpose(1, 2)
        a = torch.softmax(torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5), dim=-1)
        if reg is not None:
            a = a * rg.attention_multiplier(reg.seed, l, reg.p_attention, B, nh, N)
        a = torch.matmul(a, v).transpose(1, 2).reshape(B, N, H)
        o = F.linear(a, sd[p + "attention.out_proj.weight"], sd[p + "attention.out_proj.bias"])
        if reg is not None:
            o = drop(o, rg.site_attn_out(l), reg.p_hidden)
        x = x + o
        y = F.layer_norm(x, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], cfg.ln_eps)
        y = F.gelu(F.linear(y, sd[p + "feed_forward.intermediate_dense.weight"],
                            sd[p + "feed_forward.intermediate_dense.bias"]))
        if reg is not None:
            y = drop(y, rg.site_ff_act(l), reg.p_activation)
        y = F.linear(y, sd[p + "feed_forward.output_dense.weight"], sd[p + "feed_forward.output_dense.bias"])
        if reg is not None:
            y = drop(y, rg.site_ff_out(l), reg.p_hidden)
        x = x + y
    if return_pre_norm:
        return x
    return F.layer_norm(x, (H,), sd["encoder.encoder.layer_norm.weight"], sd["encoder.encoder.layer_norm.bias"],
                        cfg.ln_eps)


def hubert_last_hidden_state(sd, wave, cfg: EncoderCfg, reg=None) -> torch.Tensor:
    """HubertModel.forward(...).last_hidden_state (TF/models/hubert/modeling_hubert.py:889-958): eval mode, or train
    mode under the explicit masks of `reg`."""
    feats = hubert_feature_extractor(sd, wave, cfg).transpose(1, 2)
    return hubert_encoder(sd, feats, cfg, reg=reg)


# ----------------------------------------------------------------------------------------------------------
# AudioEncoder.forward (the reference's own code)
def audio_encoder_forward(sd, wave: torch.Tensor, cfg: EncoderCfg, reg=None) -> torch.Tensor:
    """REF/model/audio_encoder.py:56-88, `pool` branch: last_hidden_state -> AvgPool1d(kernel, stride) over time
    (:59-63) -> embed_projection Linear (:87). (B, T0) -> (B, A, llm_dim)."""
    enc = hubert_last_hidden_state(sd, wave, cfg, reg=reg)
    pooled = F.avg_pool1d(enc.transpose(1, 2), kernel_size=cfg.pool_kernel, stride=cfg.pool_stride).transpose(1, 2)
    return F.linear(pooled, sd["embed_projection.weight"], sd["embed_projection.bias"])


def whisper_last_hidden_state(sd, mel: torch.Tensor, cfg) -> torch.Tensor:
    """WhisperEncoder.forward(...).last_hidden_state, eval mode (TF/models/whisper/modeling_whisper.py:593-647; layer
    :361-414; attention :262-345 with q scaled by head_dim**-0.5 and a bias-free k_proj :279). (B, mel, 2*P) -> (B, P, H)."""
    H, nh = cfg.hidden, cfg.heads
    hd = H // nh
    x = F.gelu(F.conv1d(mel, sd["encoder.conv1.weight"], sd["encoder.conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, sd["encoder.conv2.weight"], sd["encoder.conv2.bias"], stride=2, padding=1))
    x = x.permute(0, 2, 1) + sd["encoder.embed_positions.weight"]
    B, N, _ = x.shape
    for l in range(cfg.layers):
        p = f"encoder.layers.{l}."
        y = F.layer_norm(x, (H,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], cfg.ln_eps)
        q = F.linear(y, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"])
        k = F.linear(y, sd[p + "self_attn.k_proj.weight"])
        v = F.linear(y, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        q = q.view(B, N, nh, hd).transpose(1, 2)
        k = k.view(B, N, nh, hd).transpose(1, 2)
        v = v.view(B, N, nh, hd).transpose(1, 2)
        a = torch.softmax(torch.matmul(q, k.transpose(2, 3)) * (hd ** -0.5), dim=-1)
        a = torch.matmul(a, v).transpose(1, 2).reshape(B, N, H)
        x = x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        y = F.layer_norm(x, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], cfg.ln_eps)
        y = F.gelu(F.linear(y, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
        x = x + F.linear(y, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
    return F.layer_norm(x, (H,), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], cfg.ln_eps)


def audio_encoder_forward_whisper(sd, mel: torch.Tensor, cfg) -> torch.Tensor:
    """REF/model/audio_encoder.py:56-88 with base == "whisper": encoder -> AvgPool1d -> embed_projection.
    The crop to compute_num_audio_embeds happens in the trainer (REF/trainer.py:280-291), not here."""
    enc = whisper_last_hidden_state(sd, mel, cfg)
    pooled = F.avg_pool1d(enc.transpose(1, 2), kernel_size=cfg.pool_kernel, stride=cfg.pool_stride).transpose(1, 2)
    return F.linear(pooled, sd["embed_projection.weight"], sd["embed_projection.bias"])


def compute_num_audio_embeds(audio_samples, sr=16000):
    """REF/utils.py:13-24 (float floor-division, then int)."""
    num_embeds = (audio_samples - (sr * 0.01)) // (sr * 0.02)
    return int(num_embeds // 4 - 1)


# ----------------------------------------------------------------------------------------------------------
# Llama (third-party: transformers LlamaModel)
def rope_inv_freq(cfg: LlmCfg) -> torch.Tensor:
    """Default RoPE frequencies and the llama3 scaling (TF/modeling_rope_utils.py:550-625)."""
    D = cfg.head_dim
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
    sc = cfg.rope_scaling
    if sc is not None and sc.get("rope_type") == "llama3":
        factor, lo, hi = sc["factor"], sc["low_freq_factor"], sc["high_freq_factor"]
        old = sc["original_max_position_embeddings"]
        low_wl, high_wl = old / lo, old / hi
        wl = 2 * math.pi / inv
        inv_l = torch.where(wl > low_wl, inv / factor, inv)
        smooth = (old / wl - lo) / (hi - lo)
        smoothed = (1 - smooth) * inv_l / factor + smooth * inv_l
        medium = ~(wl < high_wl) * ~(wl > low_wl)
        inv = torch.where(medium, smoothed, inv_l)
    return inv


def _rotate_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def llama_model_forward(sd: Dict[str, torch.Tensor], inputs_embeds: torch.Tensor,
                        attention_mask: Optional[torch.Tensor], cfg: LlmCfg,
                        output_hidden_states: bool = False):
    """LlamaModel.forward on inputs_embeds with a {0,1} left-padding mask, eval mode
    (TF/models/llama/modeling_llama.py:375-425; layer :292-332; attention :225-289 with GQA repeat_kv :187-197;
    MLP :171-184; RMSNorm :53-67 (fp32 statistics); rotary :73-168). position_ids = arange(L) for every sample
    (the reference passes none, so padded samples keep absolute positions: REF/trainer.py:317-322).
    Returns (last_hidden_state after the final norm, tuple of hidden states like HF: [0] = inputs_embeds,
    [l] = input of layer l, [-1] = post-norm output)."""
    B, L, H = inputs_embeds.shape
    dt = inputs_embeds.dtype
    nh, nkv, D = cfg.heads, cfg.kv_heads, cfg.head_dim
    pos = torch.arange(L, dtype=torch.float32)
    ang = pos[:, None] * rope_inv_freq(cfg)[None, :].float()
    emb = torch.cat((ang, ang), dim=-1)
    cos, sin = emb.cos().to(dt)[None, None], emb.sin().to(dt)[None, None]
    neg = torch.finfo(dt).min
    causal = torch.ones(L, L, dtype=torch.bool).tril()
    if attention_mask is None:
        attention_mask = torch.ones(B, L, dtype=torch.long)
    allowed = causal[None, None, :, :] & attention_mask[:, None, None, :].bool()
    bias = torch.zeros(B, 1, L, L, dtype=dt).masked_fill(~allowed, neg)

    def rms(x, w):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return w * (x.float() * torch.rsqrt(v + cfg.rms_eps)).to(dt)

    x = inputs_embeds
    hs = []
    for l in range(cfg.layers):
        if output_hidden_states:
            hs.append(x)
        p = f"model.layers.{l}."
        y = rms(x, sd[p + "input_layernorm.weight"])
        q = F.linear(y, sd[p + "self_attn.q_proj.weight"]).view(B, L, nh, D).transpose(1, 2)
        k = F.linear(y, sd[p + "self_attn.k_proj.weight"]).view(B, L, nkv, D).transpose(1, 2)
        v = F.linear(y, sd[p + "self_attn.v_proj.weight"]).view(B, L, nkv, D).transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        k = k.repeat_interleave(nh // nkv, dim=1)
        v = v.repeat_interleave(nh // nkv, dim=1)
        w = torch.matmul(q, k.transpose(2, 3)) * (D ** -0.5) + bias
        w = torch.softmax(w.float(), dim=-1).to(dt)
        a = torch.matmul(w, v).transpose(1, 2).reshape(B, L, nh * D)
        x = x + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        y = rms(x, sd[p + "post_attention_layernorm.weight"])
        y = F.silu(F.linear(y, sd[p + "mlp.gate_proj.weight"])) * F.linear(y, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(y, sd[p + "mlp.down_proj.weight"])
    x = rms(x, sd["model.norm.weight"])
    if output_hidden_states:
        hs.append(x)
    return x, tuple(hs)


def audio_llama_forward(sd, inputs_embeds, attention_mask, labels, cfg: LlmCfg, output_hidden_states=False,
                        num_logits_to_keep: int = 0):
    """AudioLlamaForCausalLM.forward (REF/model/audio_llama.py:22-113): logits = lm_head(h[:, -k:, :]) with k = 0
    meaning ALL rows (:67), then a per-sample CE over logits[-R:-1] vs labels[1:] averaged over samples (:72-101).
    `labels` is a list of 1-D tensors or a (B, R) tensor. Returns (loss|None, logits, hidden_states)."""
    h, hs = llama_model_forward(sd, inputs_embeds, attention_mask, cfg, output_hidden_states)
    logits = F.linear(h[:, -num_logits_to_keep:, :], sd["lm_head.weight"])
    loss = None
    if labels is not None:
        loss = 0.0
        for sample_logits, sample_labels in zip(logits, labels):
            R = sample_labels.shape[0]
            shift_logits = sample_logits[-R:-1, :]
            shift_labels = sample_labels[1:]
            loss = loss + F.cross_entropy(shift_logits.float().reshape(-1, cfg.vocab), shift_labels.reshape(-1))
        loss = loss / logits.shape[0]
    return loss, logits, hs


# ----------------------------------------------------------------------------------------------------------
# utils.py (the reference's own code)
def merge_prompt_response_tokens(prefix_input_ids, suffix_input_ids, inputs_embeds, response_input_ids, embed):
    """REF/utils.py:27-46: prefix | prompt | suffix[1:] | response[1:] along time."""
    return torch.cat([embed(prefix_input_ids), inputs_embeds, embed(suffix_input_ids)[:, 1:, :],
                      embed(response_input_ids)[:, 1:, :]], dim=1)


def merge_prompt_tokens(inputs_embeds, tokenizer, embed, llm_type):
    """REF/utils.py:49-73: prefix | prompt | suffix[1:]."""
    pre, suf = prompt_strings(llm_type)
    prefix_ids = tokenizer(pre, return_tensors="pt").input_ids
    suffix_ids = tokenizer(suf, return_tensors="pt").input_ids
    return torch.cat([embed(prefix_ids), inputs_embeds, embed(suffix_ids)[:, 1:, :]], dim=1)


def construct_attention_mask(seq_lens: Sequence[int]) -> torch.Tensor:
    """REF/utils.py:76-82: left-padded {0,1} int64 mask."""
    max_len = max(seq_lens)
    return torch.stack([F.pad(torch.ones(n), (max_len - n, 0)) for n in seq_lens]).long()


def batch_full_embed_sequence(all_audio_embeds, all_text_input_ids, all_response_input_ids, tokenizer, embed,
                              llm_type, process_text=False):
    """REF/utils.py:85-164: per sample merge, then LEFT zero-pad to the batch max and build masks."""
    pre, suf = prompt_strings(llm_type)
    prefix_ids = tokenizer(pre, return_tensors="pt").input_ids
    suffix_ids = tokenizer(suf, return_tensors="pt").input_ids
    audio_seqs, text_seqs = [], []
    for audio_embeds, text_ids, resp_ids in zip(all_audio_embeds, all_text_input_ids, all_response_input_ids):
        audio_seqs.append(merge_prompt_response_tokens(prefix_ids, suffix_ids, audio_embeds.unsqueeze(0),
                                                       resp_ids.unsqueeze(0), embed))
        if process_text:
            text_seqs.append(merge_prompt_response_tokens(prefix_ids, suffix_ids, embed(text_ids.unsqueeze(0)),
                                                          resp_ids.unsqueeze(0), embed))

    def pad(seqs):
        lens = [s.shape[1] for s in seqs]
        m = max(lens)
        return torch.cat([F.pad(s, (0, 0, m - s.shape[1], 0)) for s in seqs]), construct_attention_mask(lens)

    a_seq, a_mask = pad(audio_seqs)
    if process_text:
        t_seq, t_mask = pad(text_seqs)
    else:
        t_seq, t_mask = None, None
    return a_seq, a_mask, t_seq, t_mask


def soft_cross_entropy(input, target, reduction="mean"):
    """REF/utils.py:167-178."""
    ce = -torch.sum(F.softmax(target, dim=-1) * F.log_softmax(input, dim=-1), dim=-1)
    return ce.mean() if reduction == "mean" else ce


# ----------------------------------------------------------------------------------------------------------
# the train step's forward + losses (REF/trainer.py:270-374), batch size 1 like the reference
def train_step_losses(enc_sd, llm_sd, enc_cfg: EncoderCfg, llm_cfg: LlmCfg, tokenizer, audio: torch.Tensor,
                      text_ids: torch.Tensor, resp_ids: torch.Tensor, *, use_ld=True, use_fd=True,
                      w_ntp=0.5, w_ld=0.5, w_fd=1.0, fd_layers=(0, 5, 11, 17, 23), keep: bool = False):
    """One utterance: encoder forward (REF/trainer.py:278), splice (:299-313), student forward with labels (:317-322),
    teacher forward under no_grad (:337-344), KD on the last R rows (:349-352), FD MSE over the tapped hidden states
    (:358-370), total = w_ntp*ntp + w_ld*ld + w_fd*fd (:325-370). `text_ids` / `resp_ids` are the collate outputs
    (leading BOS already stripped once, REF/trainer.py:155-156)."""
    audio_embeds = audio_encoder_forward(enc_sd, audio[None, :], enc_cfg)  # (1, A, llm_dim)
    return losses_from_audio_embeds(audio_embeds, llm_sd, llm_cfg, tokenizer, text_ids, resp_ids, use_ld=use_ld,
                                    use_fd=use_fd, w_ntp=w_ntp, w_ld=w_ld, w_fd=w_fd, fd_layers=fd_layers, keep=keep)


def losses_from_audio_embeds(audio_embeds, llm_sd, llm_cfg: LlmCfg, tokenizer, text_ids, resp_ids, *, use_ld=True,
                             use_fd=True, w_ntp=0.5, w_ld=0.5, w_fd=1.0, fd_layers=(0, 5, 11, 17, 23), keep=False):
    """The LLM half of the step (REF/trainer.py:299-370) as a differentiable function of the projected audio
    embeddings (1, A, llm_dim): autograd through this is the oracle for the LLM backward."""
    embed = lambda ids: F.embedding(ids, llm_sd["model.embed_tokens.weight"])
    a_seq, a_mask, t_seq, t_mask = batch_full_embed_sequence(
        audio_embeds, [text_ids], [resp_ids], tokenizer, embed, llm_cfg.llm_type, process_text=(use_ld or use_fd))
    out = {}
    ntp, s_logits, s_hs = audio_llama_forward(llm_sd, a_seq, a_mask, [resp_ids], llm_cfg, output_hidden_states=True)
    total = w_ntp * ntp
    out["ntp_loss"] = ntp
    R = resp_ids.shape[0]
    if use_ld or use_fd:
        with torch.no_grad():
            _, t_logits, t_hs = audio_llama_forward(llm_sd, t_seq, t_mask, [resp_ids], llm_cfg,
                                                    output_hidden_states=True)
        if use_ld:
            ld = soft_cross_entropy(s_logits[:, -R:, :].float(), t_logits[:, -R:, :].float())
            total = total + w_ld * ld
            out["ld_loss"] = ld
        if use_fd:
            fd = 0.0
            for l in fd_layers:
                fd = fd + F.mse_loss(s_hs[l][:, -R:, :], t_hs[l][:, -R:, :])
            total = total + w_fd * fd
            out["fd_loss"] = fd
    out["total_loss"] = total
    if keep:
        out["audio_embeds"] = audio_embeds
        out["student_logits"] = s_logits[:, -R:, :]
        out["teacher_logits"] = t_logits[:, -R:, :] if (use_ld or use_fd) else None
        out["L_audio"], out["L_text"] = a_seq.shape[1], (t_seq.shape[1] if t_seq is not None else 0)
    return out


def audio_prompt_prefill(enc_sd, llm_sd, enc_cfg, llm_cfg, tokenizer, audio: torch.Tensor,
                         additional_text_ids: Optional[torch.Tensor] = None):
    """generate_audio_response up to and including the first LLM forward (REF/inference.py:95-135 -> :55-74):
    encoder -> optional text prompt (ids already stripped of BOS, :116-118) concatenated BEFORE the audio (:121-122)
    -> merge_prompt_tokens (:128-134) -> prefill. fp32 input instead of .half().
    Returns (audio_embeds, prompt_embeds, last-row logits)."""
    embed = lambda ids: F.embedding(ids, llm_sd["model.embed_tokens.weight"])
    audio_embeds = audio_encoder_forward(enc_sd, audio[None, :], enc_cfg)
    combined = audio_embeds
    if additional_text_ids is not None and additional_text_ids.numel() > 0:
        combined = torch.cat([embed(additional_text_ids[None, :]), audio_embeds], dim=1)
    prompt = merge_prompt_tokens(combined, tokenizer, embed, llm_cfg.llm_type)
    _, logits, _ = audio_llama_forward(llm_sd, prompt, None, None, llm_cfg, num_logits_to_keep=1)
    return audio_embeds, prompt, logits[:, -1, :]


def whisper_log_mel(wave, mel_filters, n_fft: int = 400, hop: int = 160):
    """WhisperFeatureExtractor._np_extract_fbank_features (TF/models/whisper/feature_extraction_whisper.py:105-133 over
    TF/audio_utils.py spectrogram): reflect-padded (center) STFT with a periodic Hann window, power spectrum, mel
    filter bank [201, 80], log10(max(., 1e-10)), last frame dropped, clamp to max - 8, (x + 4) / 4.
    wave: 1-D float array of n samples (already padded to the 30 s window) -> float64 [80, n // hop]."""
    import numpy as np
    x = np.asarray(wave, dtype=np.float64)
    xp = np.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    n_frames = 1 + (len(xp) - n_fft) // hop
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n_fft) / n_fft)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    spec = np.abs(np.fft.rfft(xp[idx] * win[None, :], axis=1)) ** 2          # [frames, 201]
    mel = np.maximum(np.asarray(mel_filters, dtype=np.float64).T @ spec.T, 1e-10)  # [80, frames]
    log_spec = np.log10(mel)[:, :-1]
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0
