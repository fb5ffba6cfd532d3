"""Drop-in for REF/utils.py (same names, arguments and error behaviour).

The prompt/response splice is done by one gather kernel (b2s_embed_splice_fwd) that writes the whole padded
batch at once instead of the reference's per-sample embedding lookups + cat + pad (REF/utils.py:27-46,85-164);
soft_cross_entropy runs the fused streaming KD kernel (b2s_kd_ce_loss_fwd) instead of materialising
log_softmax/softmax (REF/utils.py:167-178). Everything here needs CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import ops

SYSTEM_PROMPT = ""
MINICHAT_PROMPT_PREFIX = f"{SYSTEM_PROMPT}[|User|]"
MINICHAT_PROMPT_SUFFIX = "</s>[|Assistant|]"
LLAMA_PROMPT_PREFIX = f"<|start_header_id|>system<|end_header_id|>{SYSTEM_PROMPT}<|eot_id|><|start_header_id|>user<|end_header_id|>\n\n"
LLAMA_PROMPT_SUFFIX = "<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n"

PAD_ROW = -2 ** 31  # row_src value that makes the splice kernel write zeros (left padding)


def compute_num_audio_embeds(audio_samples, sr=16000):
    """REF/utils.py:13-24: embeddings every 20 ms, then pooled by 4, minus one (float floor-division)."""
    num_embeds = (audio_samples - (sr * 0.01)) // (sr * 0.02)
    num_pooled_embeds = int(num_embeds // 4 - 1)
    return num_pooled_embeds


def _prompt_strings(llm_type):
    if llm_type == "GeneZC/MiniChat-2-3B":
        return MINICHAT_PROMPT_PREFIX, MINICHAT_PROMPT_SUFFIX
    elif llm_type == "meta-llama/Llama-3.2-3B-Instruct":
        return LLAMA_PROMPT_PREFIX, LLAMA_PROMPT_SUFFIX
    else:
        raise Exception("Unknown LLM type.")


_ID_CACHE = {}


def prompt_ids(tokenizer, llm_type):
    """Prefix / suffix token ids, tokenised once per (tokenizer, llm_type) instead of twice per step on the CPU
    (REF/utils.py:105-106). Returned as python lists, BOS included."""
    key = (id(tokenizer), llm_type)
    if key not in _ID_CACHE:
        pre, suf = _prompt_strings(llm_type)
        p = tokenizer(pre, return_tensors="pt").input_ids.reshape(-1).tolist()
        s = tokenizer(suf, return_tensors="pt").input_ids.reshape(-1).tolist()
        _ID_CACHE[key] = (p, s)
    return _ID_CACHE[key]


def _table(embed_tokens) -> torch.Tensor:
    w = getattr(embed_tokens, "weight", None)
    if w is None or w.dtype not in (torch.float16, torch.bfloat16) or not w.is_cuda:
        raise RuntimeError("embed_tokens must be the 16-bit CUDA embedding table of AudioLlamaForCausalLM "
                           "(llm.model.embed_tokens); there is no CPU path")
    return w


def _ids_list(t) -> List[int]:
    return [int(x) for x in (t.reshape(-1).tolist() if torch.is_tensor(t) else t)]


def _splice(embed_tokens, segments_per_sample, audio_list, pad_to: Optional[int]):
    """segments_per_sample[i] = list of ("ids", [..]) / ("audio", k) entries; audio_list = the (A_k, C) tensors.
    Builds row_src for the whole (left-padded when pad_to is given) batch and runs the gather kernel once."""
    table = _table(embed_tokens)
    dev = table.device
    audio_offsets, off = [], 0
    for a in audio_list:
        audio_offsets.append(off)
        off += a.shape[0]
    audio_cat = None
    if audio_list:
        audio_cat = torch.cat([a.to(torch.float32) for a in audio_list], dim=0).contiguous()
    rows_src: List[int] = []
    lens = []
    for segs in segments_per_sample:
        src: List[int] = []
        for kind, val in segs:
            if kind == "ids":
                src.extend(val)
            else:
                n = audio_list[val].shape[0]
                src.extend(-(audio_offsets[val] + r) - 1 for r in range(n))
        lens.append(len(src))
        rows_src.append(src)
    width = max(lens) if pad_to is None else pad_to
    flat: List[int] = []
    for src in rows_src:
        flat.extend([PAD_ROW] * (width - len(src)))
        flat.extend(src)
    row_src = torch.tensor(flat, dtype=torch.int32, device=dev)
    out = ops.embed_splice(table, audio_cat, row_src)
    return out.view(len(rows_src), width, table.shape[1]).to(table.dtype), lens


def merge_prompt_response_tokens(prefix_input_ids, suffix_input_ids, inputs_embeds, response_input_ids, embed_tokens):
    """REF/utils.py:27-46: prefix | inputs_embeds | suffix[1:] | response[1:] (batch 1)."""
    segs = [("ids", _ids_list(prefix_input_ids)), ("audio", 0), ("ids", _ids_list(suffix_input_ids)[1:]),
            ("ids", _ids_list(response_input_ids)[1:])]
    out, _ = _splice(embed_tokens, [segs], [inputs_embeds.reshape(-1, inputs_embeds.shape[-1])], None)
    return out


def merge_prompt_tokens(inputs_embeds, tokenizer, embed_tokens, llm_type, device):
    """REF/utils.py:49-73: prefix | inputs_embeds | suffix[1:], ready for generation."""
    prefix, suffix = prompt_ids(tokenizer, llm_type)
    segs = [("ids", prefix), ("audio", 0), ("ids", suffix[1:])]
    out, _ = _splice(embed_tokens, [segs], [inputs_embeds.reshape(-1, inputs_embeds.shape[-1])], None)
    return out


def construct_attention_mask(seq_lens):
    """REF/utils.py:76-82: left-padded {0,1} int64 mask."""
    max_len = max(seq_lens)
    mask = torch.zeros(len(seq_lens), max_len, dtype=torch.long)
    for i, n in enumerate(seq_lens):
        mask[i, max_len - n:] = 1
    return mask


def batch_full_embed_sequence(all_audio_embeds, all_text_input_ids, all_response_input_ids, tokenizer, embed_tokens,
                              llm_type, device, process_text=False):
    """REF/utils.py:85-164: per-sample prefix | audio (or transcript) | suffix[1:] | response[1:], left zero-padded
    to the batch max, plus the {0,1} masks. One gather launch per (audio, text) batch."""
    prefix, suffix = prompt_ids(tokenizer, llm_type)
    audio_list = [a for a in all_audio_embeds]
    a_segs, t_segs = [], []
    for i, (text_ids, resp_ids) in enumerate(zip(all_text_input_ids, all_response_input_ids)):
        resp = _ids_list(resp_ids)[1:]
        a_segs.append([("ids", prefix), ("audio", i), ("ids", suffix[1:]), ("ids", resp)])
        if process_text:
            t_segs.append([("ids", prefix), ("ids", _ids_list(text_ids)), ("ids", suffix[1:]), ("ids", resp)])
    padded_audio_sequences, a_lens = _splice(embed_tokens, a_segs, audio_list, None)
    audio_attention_mask = construct_attention_mask(a_lens)
    if process_text:
        padded_text_sequences, t_lens = _splice(embed_tokens, t_segs, [], None)
        text_attention_mask = construct_attention_mask(t_lens)
    else:
        padded_text_sequences = None
        text_attention_mask = None
    return padded_audio_sequences, audio_attention_mask, padded_text_sequences, text_attention_mask


def soft_cross_entropy(input, target, reduction="mean"):
    """REF/utils.py:167-178: -sum softmax(target) * log_softmax(input) over the last dim; mean over all rows."""
    if not (input.is_cuda and target.is_cuda):
        raise RuntimeError("soft_cross_entropy (B200 path) needs CUDA tensors; there is no CPU path")
    V = input.shape[-1]
    s = input.reshape(-1, V).to(torch.bfloat16).contiguous()
    t = target.reshape(-1, V).to(torch.bfloat16).contiguous()
    rows = s.shape[0]
    labels = torch.full((rows,), -1, dtype=torch.int32, device=s.device)
    if reduction == "mean":
        offs = torch.tensor([0, rows], dtype=torch.int32, device=s.device)
        return ops.kd_ce_loss(s, t, labels, offs).loss_ld[0]
    offs = torch.arange(0, rows + 1, dtype=torch.int32, device=s.device)  # one segment per row: no reduction
    return ops.kd_ce_loss(s, t, labels, offs).loss_ld.reshape(input.shape[:-1])
