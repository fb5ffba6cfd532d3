"""Drop-in for REF/model/audio_llama.py: `AudioLlamaForCausalLM` with the reference's forward signature, output
object (.loss / .logits / .hidden_states), `.model.embed_tokens` and a greedy `.generate(inputs_embeds=...)`,
backed by the sm_100a prefill (b2s_llama_prefill) instead of transformers' LlamaModel.

Parameters are held frozen in ONE 16-bit dtype -- fp16 by default, what the reference loads (`torch_dtype=torch.float16`,
REF/trainer.py:57-61, REF/inference.py:46-51) and what meets the 2e-2 logit tolerance at Llama-3.2-3B size; bf16 on
request (`dtype=torch.bfloat16`) -- under HF's names (`model.embed_tokens.weight`, `model.layers.N.*`,
`model.norm.weight`, `lm_head.weight`), so an HF Llama-3.2-3B / MiniChat-2-3B state_dict loads unchanged; the
fused-QKV and gate|up-interleaved copies the kernels consume are built once (the LLM is frozen on this path,
REF/trainer.py:62-64). No CPU path exists.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import _lib, ops, packing
from ..config import LlmArch


@dataclass
class CausalLMOutputWithPast:
    """Field-compatible with transformers.modeling_outputs.CausalLMOutputWithPast (REF/model/audio_llama.py:107-113)."""
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    past_key_values: Optional[object] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None

    def __getitem__(self, i):
        return tuple(v for v in (self.loss, self.logits, self.past_key_values, self.hidden_states, self.attentions)
                     if v is not None)[i]


class _W(nn.Module):
    def __init__(self, *shape, dtype=torch.float16):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(*shape, dtype=dtype), requires_grad=False)


class EmbedTokens(nn.Module):
    """`llm.model.embed_tokens(ids)` (REF/utils.py:33-35,61-62,117; REF/inference.py:121): gathers through the
    splice kernel, returns the LLM's 16-bit dtype."""

    def __init__(self, vocab, hidden, dtype=torch.float16):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(vocab, hidden, dtype=dtype), requires_grad=False)

    def forward(self, input_ids: torch.Tensor) -> torch.Tensor:
        ids = input_ids.to(device=self.weight.device, dtype=torch.int32).contiguous()
        out = ops.embed_splice(self.weight, None, ids.reshape(-1))
        return out.to(self.weight.dtype).reshape(*input_ids.shape, self.weight.shape[1])


class _Attn(nn.Module):
    def __init__(self, a: LlmArch, dt):
        super().__init__()
        self.q_proj = _W(a.heads * a.head_dim, a.hidden, dtype=dt)
        self.k_proj = _W(a.kv_heads * a.head_dim, a.hidden, dtype=dt)
        self.v_proj = _W(a.kv_heads * a.head_dim, a.hidden, dtype=dt)
        self.o_proj = _W(a.hidden, a.heads * a.head_dim, dtype=dt)


class _Mlp(nn.Module):
    def __init__(self, a: LlmArch, dt):
        super().__init__()
        self.gate_proj = _W(a.ffn, a.hidden, dtype=dt)
        self.up_proj = _W(a.ffn, a.hidden, dtype=dt)
        self.down_proj = _W(a.hidden, a.ffn, dtype=dt)


class _Layer(nn.Module):
    def __init__(self, a: LlmArch, dt):
        super().__init__()
        self.self_attn = _Attn(a, dt)
        self.mlp = _Mlp(a, dt)
        self.input_layernorm = _W(a.hidden, dtype=dt)
        self.post_attention_layernorm = _W(a.hidden, dtype=dt)


class _Model(nn.Module):
    def __init__(self, a: LlmArch, dt):
        super().__init__()
        self.embed_tokens = EmbedTokens(a.vocab, a.hidden, dtype=dt)
        self.layers = nn.ModuleList([_Layer(a, dt) for _ in range(a.layers)])
        self.norm = _W(a.hidden, dtype=dt)


class _ConfigView:
    def __init__(self, a: LlmArch):
        self.vocab_size = a.vocab
        self.hidden_size = a.hidden
        self.num_hidden_layers = a.layers
        self.output_attentions = False
        self.output_hidden_states = False
        self.use_return_dict = True
        self.tie_word_embeddings = a.tie_embeddings
        self.eos_token_id = list(a.eos)
        self.bos_token_id = a.bos


class AudioLlamaForCausalLM(nn.Module):
    def __init__(self, config: LlmArch, dtype: torch.dtype = torch.float16):
        super().__init__()
        ops.fmt_of(dtype)  # fp16 or bf16 only
        self.arch = config
        self.config = _ConfigView(config)
        self.model = _Model(config, dtype)
        self.lm_head = _W(config.vocab, config.hidden, dtype=dtype)
        if config.tie_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight
        self._packed = None
        self._packed_t = None
        self._pw = None

    # ------------------------------------------------------------------------------------------ loading
    @classmethod
    def from_pretrained(cls, llm_type: str, use_cache: bool = True, torch_dtype=None, **kw):
        """Same call as REF/trainer.py:58-62 / REF/inference.py:46-51. Reads the HF checkpoint's tensors through
        transformers (needs local files or network) and re-homes them in this module."""
        from ..config import KNOWN_LLMS
        if llm_type not in KNOWN_LLMS:
            raise Exception("Unknown LLM type.")
        from transformers import AutoModelForCausalLM  # only to read the checkpoint
        dtype = torch_dtype if torch_dtype in (torch.float16, torch.bfloat16) else torch.float16
        hf = AutoModelForCausalLM.from_pretrained(llm_type, torch_dtype=dtype)
        self = cls(KNOWN_LLMS[llm_type], dtype=dtype)
        self.load_state_dict(hf.state_dict(), strict=False)
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        sd = dict(state_dict)
        if self.arch.tie_embeddings:
            sd.setdefault("lm_head.weight", sd["model.embed_tokens.weight"])
        sd = {k: v for k, v in sd.items() if not k.endswith("rotary_emb.inv_freq")}
        self._packed = self._packed_t = self._pw = None
        return super().load_state_dict(sd, strict=strict, assign=assign)

    @property
    def device(self):
        return self.model.embed_tokens.weight.device

    @property
    def dtype(self) -> torch.dtype:
        """The 16-bit format of the weights and of every 16-bit activation / gradient buffer of the LLM."""
        return self.model.embed_tokens.weight.dtype

    def _apply(self, fn, *a, **k):
        self._packed = self._packed_t = self._pw = None
        out = super()._apply(fn, *a, **k)
        if self.arch.tie_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight
        return out

    def packed(self):
        """LlamaWeights struct (+ keep-alive list) in the kernels' layouts; built once (frozen LLM)."""
        if self._packed is not None:
            return self._packed
        a = self.arch
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("AudioLlamaForCausalLM (B200 path) needs its weights on a CUDA device; no CPU path")
        keep: List[torch.Tensor] = []

        def K(t):
            keep.append(t)
            return t.data_ptr()

        dt = self.dtype
        bf = lambda t: t.detach().to(dt).contiguous()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        layers = (_lib.LlamaLayer * a.layers)()
        self._pw = []
        for l, lay in enumerate(self.model.layers):
            L = layers[l]
            sa, mlp = lay.self_attn, lay.mlp
            pw = dict(wqkv=bf(packing.pack_qkv(sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight)),
                      wo=bf(sa.o_proj.weight),
                      wgu=bf(packing.pack_gate_up(mlp.gate_proj.weight.detach(), mlp.up_proj.weight.detach())),
                      wd=bf(mlp.down_proj.weight))
            self._pw.append(pw)
            L.ln1_w = K(f32(lay.input_layernorm.weight))
            L.wqkv = K(pw["wqkv"])
            L.wo = K(pw["wo"])
            L.ln2_w = K(f32(lay.post_attention_layernorm.weight))
            L.wgu = K(pw["wgu"])
            L.wd = K(pw["wd"])
        w = _lib.LlamaWeights()
        w.layers = C.cast(layers, C.POINTER(_lib.LlamaLayer))
        w.num_layers, w.hidden, w.heads, w.kv_heads = a.layers, a.hidden, a.heads, a.kv_heads
        w.head_dim, w.ffn, w.vocab, w.rms_eps = a.head_dim, a.ffn, a.vocab, a.rms_eps
        w.final_norm_w = K(f32(self.model.norm.weight))
        w.lm_head = K(bf(self.lm_head.weight))
        w.rope_cs = K(packing.rope_table(a.head_dim, a.max_pos, a.rope_theta, a.rope_scaling, device=dev))
        w.max_pos = a.max_pos
        w.fmt = ops.fmt_of(dt)
        self._packed = (w, layers, keep)
        return self._packed

    def packed_t(self):
        """Transposed copies of the packed weights: the B operands of the dgrad GEMMs (dX = dY . W needs W^T in the
        kernel's [N, K] K-major layout). Frozen LLM -> built once, only when a backward pass is requested."""
        if self._packed_t is not None:
            return self._packed_t
        self.packed()
        a = self.arch
        keep: List[torch.Tensor] = []

        def K(t):
            keep.append(t)
            return t.data_ptr()

        layers = (_lib.LlamaLayerT * a.layers)()
        for l, pw in enumerate(self._pw):
            T = layers[l]
            T.wqkv_t = K(pw["wqkv"].t().contiguous())
            T.wo_t = K(pw["wo"].t().contiguous())
            T.wgu_t = K(pw["wgu"].t().contiguous())
            T.wd_t = K(pw["wd"].t().contiguous())
        wt = _lib.LlamaWeightsT()
        wt.layers = C.cast(layers, C.POINTER(_lib.LlamaLayerT))
        wt.lm_head_t = K(self.lm_head.weight.detach().to(self.dtype).t().contiguous())
        self._packed_t = (wt, layers, keep)
        return self._packed_t

    def alloc_saved(self, rows: int, device):
        """Per-layer activation buffers of the training forward (include/b2s.h: b2s_llama_saved)."""
        a = self.arch
        L, H, F_ = a.layers, a.hidden, a.ffn
        qkv_cols = (a.heads + 2 * a.kv_heads) * a.head_dim
        t = dict(h=torch.empty(L + 1, rows, H, device=device, dtype=torch.float32),
                 h_mid=torch.empty(L, rows, H, device=device, dtype=torch.float32),
                 qkv=torch.empty(L, rows, qkv_cols, device=device, dtype=self.dtype),
                 ao=torch.empty(L, rows, a.heads * a.head_dim, device=device, dtype=self.dtype),
                 lse=torch.empty(L, rows, a.heads, device=device, dtype=torch.float32),
                 gu=torch.empty(L, rows, 2 * F_, device=device, dtype=self.dtype))
        sv = _lib.LlamaSaved()
        for k, v in t.items():
            setattr(sv, k, v.data_ptr())
        return sv, t

    def forward_train_packed(self, saved, saved_tensors, cu_seqlens, max_seqlen, positions, logit_rows,
                             tap_layers: Sequence[int] = (), tap_rows_a=None, tap_rows_b=None):
        """Training forward over packed sequences; saved_tensors["h"][0] must hold the spliced input."""
        w = self.packed()[0]
        a = self.arch
        lib = _lib.load()
        rows = saved_tensors["h"].shape[1]
        dev = saved_tensors["h"].device
        n_log = logit_rows.numel()
        logits = torch.empty(n_log, a.vocab, device=dev, dtype=torch.bfloat16)
        taps = [int(t) for t in tap_layers if 0 < int(t) < a.layers]
        pairs = 0 if tap_rows_a is None else tap_rows_a.numel()
        fd = torch.zeros(max(1, len(taps)), max(1, pairs), device=dev, dtype=torch.float32)
        tap_arr = (C.c_int32 * max(1, len(taps)))(*taps)
        nbytes = lib.b2s_llama_train_workspace_bytes(C.byref(w), rows, n_log)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _lib.check(lib.b2s_llama_forward_train(
            C.byref(w), C.byref(saved), rows, cu_seqlens.data_ptr(), cu_seqlens.numel() - 1, int(max_seqlen),
            positions.data_ptr(), logit_rows.data_ptr(), n_log, logits.data_ptr(), tap_arr, len(taps) if pairs else 0,
            None if not pairs else tap_rows_a.data_ptr(), None if not pairs else tap_rows_b.data_ptr(), pairs,
            fd.data_ptr(), ws.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream), "llama_forward_train")
        return logits, (fd if pairs and taps else None), taps

    def backward_packed(self, saved, saved_tensors, rows_bwd: int, cu_seqlens, num_seqs_bwd: int, max_seqlen: int,
                        d_logits, dl_rows, taps: Sequence[int], tap_rows_a, tap_rows_b, tap_coef,
                        loss_scale: Optional[torch.Tensor] = None):
        """dL/d(input rows [0, rows_bwd)) given d_logits (dtype = self.dtype, already carrying the loss scale) on rows
        dl_rows and the FD coefficients (multiplied by the device scalar loss_scale here)."""
        assert d_logits.dtype == self.dtype, "gradients travel in the LLM's own 16-bit format"
        w = self.packed()[0]
        wt = self.packed_t()[0]
        lib = _lib.load()
        a = self.arch
        rows = saved_tensors["h"].shape[1]
        dev = d_logits.device
        dh = torch.empty(rows_bwd, a.hidden, device=dev, dtype=torch.float32)
        n_dl = dl_rows.numel()
        pairs = 0 if tap_rows_a is None or not len(taps) else tap_rows_a.numel()
        tap_arr = (C.c_int32 * max(1, len(taps)))(*[int(t) for t in taps])
        nbytes = lib.b2s_llama_backward_workspace_bytes(C.byref(w), rows_bwd, n_dl)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _lib.check(lib.b2s_llama_backward(
            C.byref(w), C.byref(wt), C.byref(saved), rows, rows_bwd, cu_seqlens.data_ptr(), num_seqs_bwd,
            int(max_seqlen), d_logits.data_ptr(), dl_rows.data_ptr(), n_dl, tap_arr, len(taps) if pairs else 0,
            None if not pairs else tap_rows_a.data_ptr(), None if not pairs else tap_rows_b.data_ptr(),
            None if not pairs else tap_coef.data_ptr(), None if loss_scale is None else loss_scale.data_ptr(), pairs,
            dh.data_ptr(), ws.data_ptr(), nbytes,
            torch.cuda.current_stream().cuda_stream), "llama_backward")
        return dh

    # ------------------------------------------------------------------------------------------ packed prefill
    def prefill_packed(self, h: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int, positions: torch.Tensor,
                       logit_rows: Optional[torch.Tensor], *, tap_layers: Sequence[int] = (),
                       tap_rows_a: Optional[torch.Tensor] = None, tap_rows_b: Optional[torch.Tensor] = None,
                       all_hidden: bool = False, logits_out: Optional[torch.Tensor] = None,
                       shared_prefix_len: int = 0):
        """Run the LLM over packed sequences. h: fp32 [rows, H] (overwritten with the last residual stream).
        Returns (logits bf16 [n_logit_rows, V] | None, fd_sq fp32 [taps, pairs] | None, all_hidden | None).
        shared_prefix_len > 0: sequence 0 is the prompt prefix every other sequence also attends to (step.build_plan's
        shared-prefix layout)."""
        w = self.packed()[0]
        a = self.arch
        lib = _lib.load()
        rows = h.shape[0]
        assert h.dtype == torch.float32 and h.is_contiguous() and h.shape[1] == a.hidden
        if int(max_seqlen) > a.max_pos:
            raise ValueError(f"sequence of {max_seqlen} tokens exceeds the RoPE table ({a.max_pos})")
        n_log = 0 if logit_rows is None else logit_rows.numel()
        logits = None
        if n_log:
            logits = logits_out if logits_out is not None else torch.empty(n_log, a.vocab, device=h.device,
                                                                           dtype=torch.bfloat16)
        taps = [int(t) for t in tap_layers if 0 < int(t) < a.layers]  # tap 0 = identical embeddings -> 0 (skipped)
        pairs = 0 if tap_rows_a is None else tap_rows_a.numel()
        fd = torch.zeros(max(1, len(taps)), max(1, pairs), device=h.device, dtype=torch.float32)
        tap_arr = (C.c_int32 * max(1, len(taps)))(*taps)
        hid = torch.empty(a.layers + 1, rows, a.hidden, device=h.device, dtype=torch.float32) if all_hidden else None
        nbytes = lib.b2s_llama_workspace_bytes(C.byref(w), rows, n_log)
        ws = torch.empty(nbytes, device=h.device, dtype=torch.uint8)
        args = (C.byref(w), h.data_ptr(), rows, cu_seqlens.data_ptr(), cu_seqlens.numel() - 1, int(max_seqlen),
                positions.data_ptr(), None if not n_log else logit_rows.data_ptr(), n_log,
                None if logits is None else logits.data_ptr(), tap_arr, len(taps) if pairs else 0,
                None if not pairs else tap_rows_a.data_ptr(), None if not pairs else tap_rows_b.data_ptr(), pairs,
                fd.data_ptr(), None if hid is None else hid.data_ptr(), ws.data_ptr(), nbytes)
        stream = torch.cuda.current_stream().cuda_stream
        if shared_prefix_len > 0:
            _lib.check(lib.b2s_llama_prefill_prefix(*args, int(shared_prefix_len), stream), "llama_prefill (shared prefix)")
        else:
            _lib.check(lib.b2s_llama_prefill(*args, stream), "llama_prefill")
        return logits, (fd if pairs and taps else None), hid

    # ------------------------------------------------------------------------------------------ reference API
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None,
                output_hidden_states=None, return_dict=None, cache_position=None, num_logits_to_keep: int = 0,
                **kwargs):
        """REF/model/audio_llama.py:22-113. Batched inputs use the reference's left-padding + {0,1} mask
        (REF/utils.py:136-146); internally the valid rows are packed and each sample gets its own causal mask."""
        if output_attentions:
            raise NotImplementedError("output_attentions is not available on the fused attention path")
        if past_key_values is not None:
            raise NotImplementedError("KV-cache continuation is the decode loop (SURVEY.md 8/f1), not built yet")
        if inputs_embeds is None:
            if input_ids is None:
                raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
            inputs_embeds = self.model.embed_tokens(input_ids)
        if not inputs_embeds.is_cuda:
            raise RuntimeError("AudioLlamaForCausalLM.forward (B200 path) needs CUDA inputs; there is no CPU path")
        dev = inputs_embeds.device
        B, L, H = inputs_embeds.shape
        if attention_mask is None:
            lens = [L] * B
        else:
            lens = attention_mask.to("cpu").long().sum(dim=1).tolist()
        # pack: sample b occupies padded columns [L - lens[b], L)
        cu = [0]
        for n in lens:
            cu.append(cu[-1] + int(n))
        rows = cu[-1]
        if all(n == L for n in lens):
            h = inputs_embeds.reshape(B * L, H).to(torch.float32).contiguous()
        else:
            h = torch.cat([inputs_embeds[b, L - lens[b]:, :] for b in range(B)], dim=0).to(torch.float32).contiguous()
        cu_t = torch.tensor(cu, dtype=torch.int32, device=dev)
        pos_t = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).to(dev)
        keep = [n if num_logits_to_keep == 0 else min(int(num_logits_to_keep), n) for n in lens]
        log_rows = torch.cat([torch.arange(cu[b + 1] - keep[b], cu[b + 1], dtype=torch.int32) for b in range(B)]).to(dev)
        want_hidden = bool(output_hidden_states)
        logits_p, _, hid = self.prefill_packed(h, cu_t, max(lens), pos_t, log_rows, all_hidden=want_hidden)

        Lk = max(keep)
        if all(k == Lk for k in keep):
            logits = logits_p.view(B, Lk, -1)
        else:
            logits = logits_p.new_zeros(B, Lk, logits_p.shape[-1])
            o = 0
            for b in range(B):
                logits[b, Lk - keep[b]:] = logits_p[o:o + keep[b]]
                o += keep[b]

        loss = None
        if labels is not None:
            # per-sample CE over logits[-R:-1] vs labels[1:], mean over samples (REF/model/audio_llama.py:72-101)
            lab_rows, lab_vals, offs = [], [], [0]
            o = 0
            for b in range(B):
                lb = labels[b].to(torch.int32).reshape(-1)
                R = lb.numel()
                start = o + keep[b] - R
                lab_rows.append(torch.arange(start, start + R, dtype=torch.long))
                lab_vals.append(torch.cat([lb[1:].cpu(), torch.tensor([-1], dtype=torch.int32)]))
                offs.append(offs[-1] + R)
                o += keep[b]
            sel = logits_p[torch.cat(lab_rows).to(dev)]
            res = ops.kd_ce_loss(sel, sel, torch.cat(lab_vals).to(dev), torch.tensor(offs, dtype=torch.int32, device=dev))
            loss = res.loss_ntp.sum() / B

        hidden_states = None
        if want_hidden:
            if all(n == L for n in lens):
                hidden_states = tuple(hid[l].view(B, L, H) for l in range(hid.shape[0]))
            else:
                outs = []
                for l in range(hid.shape[0]):
                    t = hid.new_zeros(B, L, H)
                    for b in range(B):
                        t[b, L - lens[b]:] = hid[l, cu[b]:cu[b + 1]]
                    outs.append(t)
                hidden_states = tuple(outs)
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=hidden_states,
                                      attentions=None)

    # ------------------------------------------------------------------------------------------ KV-cache decode
    @torch.no_grad()
    def prefill_with_cache(self, embeds: Sequence[torch.Tensor], max_new_tokens: int):
        """Prefill B prompts (list of (L_b, H) embeddings) and leave their K/V in a fresh cache with room for
        max_new_tokens more tokens each. Returns (last-row logits bf16 [B, V], cache state dict)."""
        w = self.packed()[0]
        a = self.arch
        lib = _lib.load()
        dev = self.device
        lens = [int(e.shape[0]) for e in embeds]
        B = len(lens)
        cap = [L + int(max_new_tokens) for L in lens]
        starts = [sum(cap[:b]) for b in range(B)]
        slots = sum(cap)
        if max(cap) > a.max_pos:
            raise ValueError(f"prompt + max_new_tokens = {max(cap)} exceeds the RoPE table ({a.max_pos})")
        i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)
        cu = i32([0] + [sum(lens[:b + 1]) for b in range(B)])
        pos = i32([t for L in lens for t in range(L)])
        slot_of_row = i32([starts[b] + t for b, L in enumerate(lens) for t in range(L)])
        last = i32([sum(lens[:b + 1]) - 1 for b in range(B)])
        h = torch.cat([e.to(dev, torch.float32) for e in embeds], dim=0).contiguous()
        rows = h.shape[0]
        cache = torch.empty(lib.b2s_llama_kv_cache_bytes(C.byref(w), slots), device=dev, dtype=torch.uint8)
        logits = torch.empty(B, a.vocab, device=dev, dtype=torch.bfloat16)
        nbytes = lib.b2s_llama_workspace_bytes(C.byref(w), rows, B)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _lib.check(lib.b2s_llama_prefill_kv(C.byref(w), h.data_ptr(), rows, cu.data_ptr(), B, max(lens), pos.data_ptr(),
                                            last.data_ptr(), B, logits.data_ptr(), cache.data_ptr(), slots,
                                            slot_of_row.data_ptr(), ws.data_ptr(), nbytes,
                                            torch.cuda.current_stream().cuda_stream), "llama_prefill_kv")
        dbytes = lib.b2s_llama_decode_workspace_bytes(C.byref(w), B)
        state = dict(cache=cache, slots=slots, seq_start=i32(starts), seq_len=i32(lens), B=B,
                     ws=torch.empty(dbytes, device=dev, dtype=torch.uint8), cap=cap, lens=list(lens))
        return logits, state

    @torch.no_grad()
    def decode_step(self, token_ids: torch.Tensor, state, use_graph: bool = True) -> torch.Tensor:
        """Append one token per sequence (int [B], device) and return the next-token logits bf16 [B, V] (a buffer
        owned by `state`, overwritten by the next call). A step is ~230 short launches, so after one eager call the
        whole step (kernels + the seq_len increment) is captured into a CUDA graph and replayed: every buffer the
        kernels touch is persistent in `state`, positions / cache lengths are read from device memory."""
        w = self.packed()[0]
        lib = _lib.load()
        B = state["B"]
        if max(l + 1 for l in state["lens"]) > max(state["cap"]):
            raise RuntimeError("KV cache is full")
        if "tok" not in state:
            state["tok"] = torch.zeros(B, device=self.device, dtype=torch.int32)
            state["logits"] = torch.empty(B, self.arch.vocab, device=self.device, dtype=torch.bfloat16)
            state["calls"], state["graph"] = 0, None
        state["tok"].copy_(token_ids.to(device=self.device, dtype=torch.int32).reshape(B))

        def run():
            _lib.check(lib.b2s_llama_decode_step(
                C.byref(w), self.model.embed_tokens.weight.data_ptr(), state["tok"].data_ptr(), B,
                state["cache"].data_ptr(), state["slots"], state["seq_start"].data_ptr(), state["seq_len"].data_ptr(),
                state["logits"].data_ptr(), state["ws"].data_ptr(), state["ws"].numel(),
                torch.cuda.current_stream().cuda_stream), "llama_decode_step")
            state["seq_len"].add_(1)

        if state["graph"] is not None:
            state["graph"].replay()
        elif use_graph and state["calls"] >= 1:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            state["graph"] = graph
            graph.replay()
        else:
            run()
        state["calls"] += 1
        state["lens"] = [l + 1 for l in state["lens"]]
        return state["logits"]

    @torch.no_grad()
    def generate(self, input_ids=None, inputs_embeds=None, max_new_tokens: int = 256, use_kv_cache: bool = True,
                 **kwargs):
        """Greedy decoding from a prompt given as embeddings (REF/inference.py:55-66, REF/trainer.py:530-545);
        returns only the new token ids, like HF generate with inputs_embeds. The prompt is prefilled once into a KV
        cache and every new token is one `b2s_llama_decode_step` (use_kv_cache=False re-runs the prefill per token:
        the O(n^2) cross-check used by the tests)."""
        if inputs_embeds is None:
            inputs_embeds = self.model.embed_tokens(input_ids)
        if inputs_embeds.shape[0] != 1:
            raise NotImplementedError("generate() is batch-1 like the reference")
        eos = set(int(e) for e in self.arch.eos)
        out: List[int] = []
        if not use_kv_cache:
            seq = inputs_embeds.to(self.dtype)
            for _ in range(int(max_new_tokens)):
                logits = self.forward(inputs_embeds=seq, num_logits_to_keep=1).logits
                nxt = int(logits[0, -1].float().argmax())
                out.append(nxt)
                if nxt in eos:
                    break
                tok = torch.tensor([[nxt]], device=seq.device)
                seq = torch.cat([seq, self.model.embed_tokens(tok)], dim=1)
            return torch.tensor([out], dtype=torch.long, device=seq.device)
        if int(max_new_tokens) <= 0:
            return torch.zeros(1, 0, dtype=torch.long, device=self.device)
        logits, state = self.prefill_with_cache([inputs_embeds[0]], int(max_new_tokens))
        for i in range(int(max_new_tokens)):
            nxt_t = logits[0].float().argmax().to(torch.int32).reshape(1)
            nxt = int(nxt_t)
            out.append(nxt)
            if nxt in eos or i + 1 == int(max_new_tokens):
                break
            logits = self.decode_step(nxt_t, state)
        return torch.tensor([out], dtype=torch.long, device=self.device)
