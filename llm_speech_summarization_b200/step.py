"""The fused audio-prompt step: the body of the reference's training loop up to the losses
(REF/trainer.py:270-370) for a whole micro-batch of utterances in a handful of launches.

    encoder (b2s_hubert_forward, batched) -> one splice launch that builds EVERY student (audio-prompt) and teacher
    (text-prompt) sequence packed back to back -> ONE LLM prefill over all of them (student and teacher share the
    frozen weights; independent causal masks via cu_seqlens) with the FD taps fused in and the LM head computed
    only on the consumed rows (last R of each sequence, REF/trainer.py:334,350-351) -> fused CE+KD loss.

Per-utterance semantics are exactly the reference's batch-1 step: every utterance gets its own means
(ntp over R-1 rows, ld over R rows, fd over R x hidden per tapped layer), and
total_u = w_ntp*ntp_u + w_ld*ld_u + w_fd*fd_u (REF/trainer.py:325-370).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .utils import prompt_ids


@dataclass
class StepPlan:
    """Host-built index arrays for one micro-batch (everything the kernels need besides the tensors)."""
    row_src: torch.Tensor       # int32 [rows]   splice sources
    cu_seqlens: torch.Tensor    # int32 [2B+1]
    positions: torch.Tensor     # int32 [rows]
    logit_rows: torch.Tensor    # int32 [2*sumR] student rows then teacher rows
    labels: torch.Tensor        # int32 [sumR]
    row_offsets: torch.Tensor   # int32 [B+1]
    max_seqlen: int
    rows: int
    sum_r: int
    resp_lens: List[int]
    L_audio: List[int]
    L_text: List[int]
    seg: Optional[torch.Tensor] = None        # int64 [sumR] utterance index of each response row
    resp_len_f: Optional[torch.Tensor] = None  # fp32 [B]
    audio_rows: Optional[torch.Tensor] = None  # int32 [B*A] packed row of every audio embedding (student sequences)
    student_rows_total: int = 0                # rows of all student sequences (they are packed first)
    shared_prefix_len: int = 0                 # > 0: sequence 0 is the prompt prefix, stored once (forward-only layout)


def build_plan(prefix: Sequence[int], suffix: Sequence[int], n_audio, text_ids: Sequence[Sequence[int]],
               resp_ids: Sequence[Sequence[int]], with_teacher: bool = True,
               audio_stride: Optional[int] = None) -> Dict[str, object]:
    """Pure-python index construction (no torch, unit-tested on CPU).
    Sequence layout (REF/utils.py:27-46 + the double BOS strip):
        student_i = prefix | audio rows of utterance i | suffix[1:] | response_i[1:]
        teacher_i = prefix | transcript_i              | suffix[1:] | response_i[1:]
    Packed order: student_0..student_{B-1}, teacher_0..teacher_{B-1}. Logits are produced for the last R_i rows of
    each; CE labels for row j < R_i - 1 of utterance i are response_i[j+1], the last row has none (-1).
    `n_audio` is one count for all utterances or a per-utterance list (ragged batch); the audio embeddings sit in a
    [B, audio_stride, C] tensor (audio_stride = the largest count by default) and only the first n_audio[i] rows of
    utterance i enter its sequence (REF/trainer.py:280-291)."""
    B = len(resp_ids)
    n_each = [int(n_audio)] * B if isinstance(n_audio, int) else [int(n) for n in n_audio]
    A = max(n_each) if audio_stride is None else int(audio_stride)
    assert len(n_each) == B and max(n_each) <= A
    row_src: List[int] = []
    cu = [0]
    positions: List[int] = []
    L_audio, L_text = [], []
    suf = list(suffix[1:])
    for i in range(B):
        resp = list(resp_ids[i])[1:]
        seq = list(prefix) + [-(i * A + r) - 1 for r in range(n_each[i])] + suf + resp
        row_src.extend(seq)
        positions.extend(range(len(seq)))
        cu.append(cu[-1] + len(seq))
        L_audio.append(len(seq))
    if with_teacher:
        for i in range(B):
            resp = list(resp_ids[i])[1:]
            seq = list(prefix) + list(text_ids[i]) + suf + resp
            row_src.extend(seq)
            positions.extend(range(len(seq)))
            cu.append(cu[-1] + len(seq))
            L_text.append(len(seq))
    resp_lens = [len(r) for r in resp_ids]
    s_rows: List[int] = []
    t_rows: List[int] = []
    labels: List[int] = []
    offs = [0]
    for i in range(B):
        R = resp_lens[i]
        if R > L_audio[i]:
            raise ValueError("response longer than its sequence")
        s_rows.extend(range(cu[i + 1] - R, cu[i + 1]))
        if with_teacher:
            t_rows.extend(range(cu[B + i + 1] - R, cu[B + i + 1]))
        labels.extend(list(resp_ids[i])[1:] + [-1])
        offs.append(offs[-1] + R)
    P = len(prefix)
    # packed row of audio embedding (i, r) in the [B, A] layout; -1 = padding row of a ragged batch (zero gradient)
    audio_rows = [cu[i] + P + r if r < n_each[i] else -1 for i in range(B) for r in range(A)]
    return dict(row_src=row_src, cu_seqlens=cu, positions=positions, student_rows=s_rows, teacher_rows=t_rows,
                labels=labels, row_offsets=offs, max_seqlen=max(L_audio + L_text), rows=cu[-1], sum_r=offs[-1],
                resp_lens=resp_lens, L_audio=L_audio, L_text=L_text, audio_rows=audio_rows,
                student_rows_total=cu[B])


def build_plan_arrays(prefix: Sequence[int], suffix: Sequence[int], n_audio, text_ids, resp_ids,
                      with_teacher: bool = True, audio_stride: Optional[int] = None,
                      shared_prefix: bool = False) -> Dict[str, object]:
    """`build_plan` with numpy (same keys, int32 arrays instead of lists): what the step calls per micro-batch -- the
    list version costs ~7 ms of Python per 32 utterances, this one well under 1 ms. Equality with `build_plan` is a
    CPU unit test.

    shared_prefix=True is the forward-only SHARED-PREFIX layout: every sequence the reference builds starts with the same
    prompt-prefix tokens, and under a causal mask those rows are identical in all of them at every layer. They are packed
    ONCE, as sequence 0; sequences 1 .. 2B hold only their own rows (positions continue after the prefix) and the
    attention kernel lets them see sequence 0's keys (b2s_llama_prefill_prefix). Every row the losses consume has the
    value it has in the reference's layout; (2B - 1) * len(prefix) rows are not computed."""
    import numpy as np
    B = len(resp_ids)
    n_each = [int(n_audio)] * B if isinstance(n_audio, int) else [int(n) for n in n_audio]
    A = max(n_each) if audio_stride is None else int(audio_stride)
    assert len(n_each) == B and max(n_each) <= A
    i32 = lambda x: np.asarray(x, dtype=np.int32).reshape(-1)
    pre, suf = i32(prefix), i32(suffix)[1:]
    resp = [i32(r) for r in resp_ids]
    text = [i32(t) for t in text_ids] if with_teacher else []
    P = len(pre)
    head = np.zeros(0, dtype=np.int32) if shared_prefix else pre  # what every sequence starts with in the packed layout
    seqs = [np.concatenate([head, -(i * A + np.arange(n_each[i], dtype=np.int32)) - 1, suf, resp[i][1:]]) for i in range(B)]
    extra = P if shared_prefix else 0  # rows of each sequence that live in sequence 0
    L_audio = [len(q) + extra for q in seqs]
    L_text = []
    if with_teacher:
        t_seqs = [np.concatenate([head, text[i], suf, resp[i][1:]]) for i in range(B)]
        L_text = [len(q) + extra for q in t_seqs]
        seqs = seqs + t_seqs
    if shared_prefix:
        seqs = [pre] + seqs
    s0 = 1 if shared_prefix else 0  # index of the first student sequence
    lens = np.asarray([len(q) for q in seqs], dtype=np.int64)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    row_src = np.concatenate(seqs).astype(np.int32)
    first_pos = np.full(len(seqs), extra, dtype=np.int64)
    if shared_prefix:
        first_pos[0] = 0
    positions = (np.arange(int(cu[-1]), dtype=np.int64) - np.repeat(cu[:-1].astype(np.int64), lens) +
                 np.repeat(first_pos, lens)).astype(np.int32)
    resp_lens = [len(r) for r in resp]
    for i in range(B):
        if resp_lens[i] > L_audio[i]:
            raise ValueError("response longer than its sequence")
    R = np.asarray(resp_lens, dtype=np.int64)
    within = np.arange(int(R.sum()), dtype=np.int64) - np.repeat(np.concatenate([[0], np.cumsum(R)[:-1]]), R)
    s_rows = (np.repeat(cu[s0 + 1:s0 + B + 1].astype(np.int64) - R, R) + within).astype(np.int32)
    t_rows = ((np.repeat(cu[s0 + B + 1:s0 + 2 * B + 1].astype(np.int64) - R, R) + within).astype(np.int32) if with_teacher
              else np.zeros(0, dtype=np.int32))
    labels = np.concatenate([np.concatenate([resp[i][1:], [-1]]) for i in range(B)]).astype(np.int32) if B else i32([])
    offs = np.concatenate([[0], np.cumsum(R)]).astype(np.int32)
    r_idx = np.arange(A, dtype=np.int64)[None, :]
    audio_rows = np.where(r_idx < np.asarray(n_each, dtype=np.int64)[:, None],
                          cu[s0:s0 + B].astype(np.int64)[:, None] + (P - extra) + r_idx, -1).astype(np.int32).reshape(-1)
    seg = np.repeat(np.arange(B, dtype=np.int32), R)
    # max_seqlen bounds both the query blocks of the longest sequence and the largest position + 1 (RoPE table check)
    return dict(row_src=row_src, cu_seqlens=cu, positions=positions, student_rows=s_rows, teacher_rows=t_rows,
                labels=labels, row_offsets=offs, max_seqlen=int(lens.max()) + extra, rows=int(cu[-1]), sum_r=int(R.sum()),
                resp_lens=resp_lens, L_audio=L_audio, L_text=L_text, audio_rows=audio_rows,
                student_rows_total=int(cu[s0 + B]), seg=seg, shared_prefix_len=extra)


class PendingStep:
    """A submitted micro-batch (`AudioPromptStep.submit` / `EncoderTrainer.submit`): the per-utterance losses travel
    to pinned host memory asynchronously; `result()` waits for exactly that copy."""

    def __init__(self, keys, host_buf: torch.Tensor, event: torch.cuda.Event, extra=None):
        self.keys, self.host_buf, self.event, self.extra = keys, host_buf, event, extra or {}

    def result(self) -> Dict[str, torch.Tensor]:
        self.event.synchronize()
        out = {k: self.host_buf[i] for i, k in enumerate(self.keys)}
        out.update(self.extra)
        return out


def _stage_to_device(host: torch.Tensor, device, copy_stream: torch.cuda.Stream) -> torch.Tensor:
    """Pinned host tensor -> device on a side stream, so the copy of batch i+1 overlaps the compute of batch i; the
    compute stream waits on the copy's event and the allocator is told about the cross-stream use."""
    main = torch.cuda.current_stream(device)
    with torch.cuda.stream(copy_stream):
        dev = host.to(device, non_blocking=True)
        done = torch.cuda.Event()
        done.record(copy_stream)
    main.wait_event(done)
    dev.record_stream(main)
    return dev


def _losses_to_host(out: Dict[str, torch.Tensor]) -> PendingStep:
    keys = [k for k in ("ntp_loss", "ld_loss", "fd_loss", "total_loss") if k in out]
    stacked = torch.stack([out[k] for k in keys])
    host = torch.empty(stacked.shape, dtype=stacked.dtype, pin_memory=True)
    host.copy_(stacked, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return PendingStep(keys, host, ev)


class AudioPromptStep:
    def __init__(self, audio_encoder, llm, tokenizer, llm_type: str, *, use_ld_loss: bool = True,
                 use_fd_loss: bool = True, ntp_loss_weight: float = 0.5, ld_loss_weight: float = 0.5,
                 fd_loss_weight: float = 1.0, fd_loss_connector_layers: Sequence[int] = (0, 5, 11, 17, 23),
                 share_prefix: bool = True):
        # share_prefix: the forward-only paths (forward_losses, prefill_prompts) compute the prompt-prefix rows once per
        # step instead of once per sequence (build_plan_arrays). The training step keeps the reference's layout: its
        # backward kernels have no shared-prefix form.
        self.share_prefix = bool(share_prefix)
        self.audio_encoder = audio_encoder
        self.llm = llm
        self.llm_type = llm_type
        self.prefix, self.suffix = prompt_ids(tokenizer, llm_type)
        self.use_ld, self.use_fd = use_ld_loss, use_fd_loss
        self.w_ntp, self.w_ld, self.w_fd = ntp_loss_weight, ld_loss_weight, fd_loss_weight
        self.fd_layers = [int(l) for l in fd_loss_connector_layers]
        for l in self.fd_layers:
            if not 0 <= l < llm.arch.layers:
                raise ValueError(f"fd_loss_connector_layers entry {l} outside the LLM's {llm.arch.layers} layers")

    @classmethod
    def from_config(cls, config, audio_encoder, llm, tokenizer):
        t = config.train
        return cls(audio_encoder, llm, tokenizer, config.model.llm_type, use_ld_loss=t.use_ld_loss,
                   use_fd_loss=t.use_fd_loss, ntp_loss_weight=t.ntp_loss_weight, ld_loss_weight=t.ld_loss_weight,
                   fd_loss_weight=t.fd_loss_weight, fd_loss_connector_layers=t.fd_loss_connector_layers)

    def plan(self, n_audio, text_ids, resp_ids, device, audio_stride: Optional[int] = None,
             shared_prefix: bool = False) -> StepPlan:
        import numpy as np
        as_np = lambda xs: [x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x) for x in xs]
        d = build_plan_arrays(self.prefix, self.suffix, n_audio, as_np(text_ids), as_np(resp_ids),
                              with_teacher=(self.use_ld or self.use_fd), audio_stride=audio_stride,
                              shared_prefix=shared_prefix and len(self.prefix) > 0 and len(self.prefix) <= 64)
        # every index array of the plan travels in ONE pinned staging buffer and one H2D copy (each array starts on a
        # 16-byte boundary); the fp32 response lengths ride along as raw bits
        parts = {"row_src": d["row_src"], "cu_seqlens": d["cu_seqlens"], "positions": d["positions"],
                 "logit_rows": np.concatenate([d["student_rows"], d["teacher_rows"]]), "labels": d["labels"],
                 "row_offsets": d["row_offsets"], "audio_rows": d["audio_rows"], "seg": d["seg"]}
        offs, total = {}, 0
        for k, v in parts.items():
            offs[k] = total
            total += (len(v) + 3) // 4 * 4
        offs["resp_len_f"] = total
        nB = len(d["resp_lens"])
        total += (nB + 3) // 4 * 4
        host = torch.zeros(total, dtype=torch.int32).pin_memory()
        host_np = host.numpy()
        for k, v in parts.items():
            host_np[offs[k]:offs[k] + len(v)] = v
        host_np[offs["resp_len_f"]:offs["resp_len_f"] + nB] = np.asarray(d["resp_lens"], dtype=np.float32).view(np.int32)
        dev_buf = host.to(device, non_blocking=True)
        i32 = lambda k: dev_buf[offs[k]:offs[k] + len(parts[k])]
        return StepPlan(row_src=i32("row_src"), cu_seqlens=i32("cu_seqlens"), positions=i32("positions"),
                        logit_rows=i32("logit_rows"), labels=i32("labels"), row_offsets=i32("row_offsets"),
                        max_seqlen=d["max_seqlen"], rows=d["rows"], sum_r=d["sum_r"], resp_lens=d["resp_lens"],
                        L_audio=d["L_audio"], L_text=d["L_text"], seg=i32("seg").to(torch.int64),
                        resp_len_f=dev_buf[offs["resp_len_f"]:offs["resp_len_f"] + nB].view(torch.float32),
                        audio_rows=i32("audio_rows"), student_rows_total=d["student_rows_total"],
                        shared_prefix_len=d["shared_prefix_len"])

    @torch.no_grad()
    def forward_losses(self, waves: torch.Tensor, text_ids, resp_ids, plan: Optional[StepPlan] = None,
                       keep: bool = False, num_audio_embeds: Optional[int] = None,
                       lengths=None) -> Dict[str, torch.Tensor]:
        """waves: CUDA fp32 (B, T0), equal-length utterances (the reference's collate zero-pads a batch to one
        length, REF/trainer.py:146-149; batch-1 has no padding). Returns per-utterance device tensors
        ntp_loss / ld_loss / fd_loss / total_loss of shape (B,)."""
        if not waves.is_cuda:
            raise RuntimeError("AudioPromptStep needs CUDA inputs; there is no CPU path")
        dev = waves.device
        B = waves.shape[0]
        n_valid = None
        if lengths is not None and len(set(int(n) for n in lengths)) > 1:
            # ragged batch (zero-padded waveforms + samples per utterance): the ragged-aware kernels are the training
            # forward's; in eval mode it applies no regulariser, its saved activations are simply not used
            was_training = self.audio_encoder.training
            self.audio_encoder.eval()
            audio = self.audio_encoder.forward_train(waves, lengths=lengths)
            self.audio_encoder.train(was_training)
            n_valid = self.audio_encoder._train_ctx.get("n_valid")
        else:
            audio = self.audio_encoder.forward_fp32(waves)  # (B, A, C) fp32
        if num_audio_embeds is not None and num_audio_embeds < audio.shape[1]:
            # the trainer's un-padding (REF/trainer.py:280-291): Whisper's fixed 30 s window gives 374 pooled frames,
            # compute_num_audio_embeds keeps the ones covered by audio
            audio = audio[:, :num_audio_embeds].contiguous()
        A, Cdim = audio.shape[1], audio.shape[2]
        if plan is None:
            plan = self.plan(A if n_valid is None else n_valid, text_ids, resp_ids, dev, audio_stride=A,
                             shared_prefix=self.share_prefix)
        h = ops.embed_splice(self.llm.model.embed_tokens.weight, audio.view(B * A, Cdim), plan.row_src)
        with_teacher = self.use_ld or self.use_fd
        taps = [l for l in self.fd_layers if l > 0] if self.use_fd else []
        s_rows = plan.logit_rows[:plan.sum_r]
        t_rows = plan.logit_rows[plan.sum_r:] if with_teacher else None
        logits, fd_sq, _ = self.llm.prefill_packed(
            h, plan.cu_seqlens, plan.max_seqlen, plan.positions, plan.logit_rows, tap_layers=taps,
            tap_rows_a=s_rows if taps else None, tap_rows_b=t_rows if taps else None,
            shared_prefix_len=plan.shared_prefix_len)
        s_log = logits[:plan.sum_r]
        t_log = logits[plan.sum_r:] if with_teacher else s_log
        res = ops.kd_ce_loss(s_log, t_log, plan.labels, plan.row_offsets, scale_kd=self.w_ld, scale_ce=self.w_ntp)
        out = {"ntp_loss": res.loss_ntp}
        total = self.w_ntp * res.loss_ntp
        if self.use_ld:
            out["ld_loss"] = res.loss_ld
            total = total + self.w_ld * res.loss_ld
        if self.use_fd:
            # fd_u = sum over tapped layers of mean_{R_u x H}(diff^2); the layer-0 tap compares identical response
            # embeddings and is exactly 0, so it is skipped.
            fd = torch.zeros(B, device=dev, dtype=torch.float32)
            if fd_sq is not None:
                per_row = fd_sq.sum(dim=0)  # [sumR]
                fd.index_add_(0, plan.seg, per_row)
                fd = fd / (plan.resp_len_f * Cdim)
            out["fd_loss"] = fd
            total = total + self.w_fd * fd
        out["total_loss"] = total
        if keep:
            out["audio_embeds"] = audio
            out["student_logits"] = s_log
            out["teacher_logits"] = t_log
            out["kd_stats"] = res
            out["plan"] = plan
        return out

    @torch.no_grad()
    def prefill_prompts(self, waves: torch.Tensor, extra_text_ids=None, plan=None):
        """Batched form of the first LLM forward of `generate_audio_response` (REF/inference.py:95-135 -> :55-74):
        encoder -> prompt = prefix | [additional text] | audio embeddings | suffix[1:] for every utterance, all
        prompts packed into ONE prefill, logits of the last row of each (what greedy decoding consumes first).
        waves: CUDA fp32 (B, T0); extra_text_ids: optional list of B id tensors (BOS already stripped,
        REF/inference.py:116-118). Returns (last-row logits bf16 [B, V], plan) -- pass `plan` back to skip the host-side
        index building when the prompt geometry repeats."""
        if not waves.is_cuda:
            raise RuntimeError("AudioPromptStep needs CUDA inputs; there is no CPU path")
        dev = waves.device
        audio = self.audio_encoder.forward_fp32(waves)
        B, A, Cdim = audio.shape
        if plan is None:
            import numpy as np
            pre = np.asarray(self.prefix, dtype=np.int32)
            suf = np.asarray(self.suffix, dtype=np.int32)[1:]
            share = self.share_prefix and 0 < len(pre) <= 64
            head = np.zeros(0, dtype=np.int32) if share else pre
            seqs = []
            for i in range(B):
                extra = (np.asarray(extra_text_ids[i].detach().cpu() if torch.is_tensor(extra_text_ids[i])
                                    else extra_text_ids[i], dtype=np.int32).reshape(-1)
                         if extra_text_ids is not None else np.zeros(0, dtype=np.int32))
                seqs.append(np.concatenate([head, extra, -(i * A + np.arange(A, dtype=np.int32)) - 1, suf]))
            if share:  # the prefix rows once, as sequence 0 (build_plan_arrays' shared-prefix layout)
                seqs = [pre] + seqs
            lens = np.asarray([len(q) for q in seqs], dtype=np.int64)
            cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
            first = np.full(len(seqs), len(pre) if share else 0, dtype=np.int64)
            if share:
                first[0] = 0
            pos = (np.arange(int(cu[-1]), dtype=np.int64) - np.repeat(cu[:-1].astype(np.int64), lens) +
                   np.repeat(first, lens)).astype(np.int32)
            last = cu[2:] - 1 if share else cu[1:] - 1
            nseq = len(seqs)
            host = torch.from_numpy(np.concatenate([np.concatenate(seqs), cu, pos, last]).astype(np.int32)).pin_memory()
            d = host.to(dev, non_blocking=True)
            n = int(cu[-1])
            plan = dict(row_src=d[:n], cu=d[n:n + nseq + 1], pos=d[n + nseq + 1:2 * n + nseq + 1],
                        last=d[2 * n + nseq + 1:], max_len=int(lens.max()) + (len(pre) if share else 0), rows=n,
                        shared_prefix_len=len(pre) if share else 0)
        h = ops.embed_splice(self.llm.model.embed_tokens.weight, audio.view(B * A, Cdim), plan["row_src"])
        logits, _, _ = self.llm.prefill_packed(h, plan["cu"], plan["max_len"], plan["pos"], plan["last"],
                                               shared_prefix_len=plan.get("shared_prefix_len", 0))
        return logits, plan

    @torch.no_grad()
    def validation_losses(self, waves: torch.Tensor, text_ids, resp_ids,
                          num_audio_embeds: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Per-utterance next-token losses of the audio-prompt AND the text-prompt sequence (REF/trainer.py:438-451:
        `llm_audio_output.loss`, `llm_text_output.loss`), eval mode, from ONE packed pass. The fused loss kernel gives
        the student CE; called with the roles swapped it gives the text-prompt CE."""
        saved = self.use_ld, self.use_fd
        self.use_ld, self.use_fd = True, False  # teacher sequence needed, no taps
        try:
            out = self.forward_losses(waves, text_ids, resp_ids, keep=True, num_audio_embeds=num_audio_embeds)
        finally:
            self.use_ld, self.use_fd = saved
        plan = out["plan"]
        swapped = ops.kd_ce_loss(out["teacher_logits"], out["student_logits"], plan.labels, plan.row_offsets,
                                 scale_kd=0.0, scale_ce=1.0)
        return {"audio_ntp_loss": out["ntp_loss"], "text_ntp_loss": swapped.loss_ntp,
                "audio_embeds": out["audio_embeds"]}

    @torch.no_grad()
    def llm_forward_backward(self, audio: torch.Tensor, text_ids, resp_ids, loss_scale: float = 1.0,
                             plan: Optional[StepPlan] = None, n_audio=None, scaler=None) -> Dict[str, torch.Tensor]:
        """The LLM half of the TRAINING step for projected audio embeddings `audio` (B, A, C) fp32:
        splice -> training forward (activations kept) -> losses -> backward through the frozen LLM
        (REF/trainer.py:299-374). Returns the per-utterance losses and `d_audio_embeds` (B, A, C) fp32 =
        d( sum_u loss_scale * total_u ) / d audio  -- loss_scale = 1 / grad_accum_interval in the reference --
        times the dynamic scale of `scaler` (training.GradScaler, the reference's GradScaler, REF/trainer.py:374) when
        one is given. An fp16 LLM carries fp16 gradients, which underflow without that scale."""
        if scaler is None and torch.float16 in (self.llm.dtype, getattr(self.audio_encoder, "operand_dtype", None)):
            # fp16 gradients underflow without a loss scale: a caller that brings none (tests, one-off gradient checks)
            # gets a private scaler at GradScaler's initial scale; the gradients it receives carry out["grad_scale"]
            if getattr(self, "_own_scaler", None) is None:
                from .training import GradScaler
                self._own_scaler = GradScaler(audio.device)
            scaler = self._own_scaler
        scale_t = None if scaler is None else scaler.scale_tensor
        dev = audio.device
        B, A, Cdim = audio.shape
        if plan is None:  # n_audio: per-utterance counts of a ragged batch (the first n_audio[i] rows of audio[i])
            plan = self.plan(A if n_audio is None else n_audio, text_ids, resp_ids, dev, audio_stride=A)
        if plan.shared_prefix_len:
            raise ValueError("the training step needs the reference's layout: build its plan with shared_prefix=False")
        with_teacher = self.use_ld or self.use_fd
        saved, st = self.llm.alloc_saved(plan.rows, dev)
        ops.embed_splice(self.llm.model.embed_tokens.weight, audio.reshape(B * A, Cdim).contiguous(), plan.row_src,
                         out=st["h"][0])
        s_rows = plan.logit_rows[:plan.sum_r]
        t_rows = plan.logit_rows[plan.sum_r:] if with_teacher else None
        fd_layers = [l for l in self.fd_layers if l > 0] if self.use_fd else []
        logits, fd_sq, taps = self.llm.forward_train_packed(
            saved, st, plan.cu_seqlens, plan.max_seqlen, plan.positions, plan.logit_rows, tap_layers=fd_layers,
            tap_rows_a=s_rows if fd_layers else None, tap_rows_b=t_rows if fd_layers else None)
        s_log = logits[:plan.sum_r]
        t_log = logits[plan.sum_r:] if with_teacher else s_log
        res = ops.kd_ce_loss(s_log, t_log, plan.labels, plan.row_offsets,
                             scale_kd=(self.w_ld if self.use_ld else 0.0) * loss_scale, scale_ce=self.w_ntp * loss_scale)
        out = {"ntp_loss": res.loss_ntp}
        total = self.w_ntp * res.loss_ntp
        if self.use_ld:
            out["ld_loss"] = res.loss_ld
            total = total + self.w_ld * res.loss_ld
        tap_coef = None
        if self.use_fd:
            fd = torch.zeros(B, device=dev, dtype=torch.float32)
            if fd_sq is not None:
                fd.index_add_(0, plan.seg, fd_sq.sum(dim=0))
                fd = fd / (plan.resp_len_f * Cdim)
                # d fd_u / d h_s = 2 (h_s - h_t) / (R_u * C) per tapped layer
                tap_coef = (2.0 * self.w_fd * loss_scale / (plan.resp_len_f * Cdim))[plan.seg].contiguous()
            out["fd_loss"] = fd
            total = total + self.w_fd * fd
        out["total_loss"] = total
        d_logits = ops.kd_ce_loss_bwd(s_log, t_log, plan.labels, res, loss_scale=scale_t, out_dtype=self.llm.dtype)
        ns = len(plan.L_audio)
        dh0 = self.llm.backward_packed(saved, st, plan.student_rows_total, plan.cu_seqlens, ns, max(plan.L_audio),
                                       d_logits, s_rows, taps if tap_coef is not None else [],
                                       s_rows if tap_coef is not None else None,
                                       t_rows if tap_coef is not None else None, tap_coef, loss_scale=scale_t)
        out["d_audio_embeds"] = ops.gather_rows(dh0, plan.audio_rows).view(B, A, Cdim)
        out["grad_scale"] = scale_t if scale_t is not None else torch.ones(1, device=dev)
        out["plan"] = plan
        return out

    @torch.no_grad()
    def forward_backward(self, waves: torch.Tensor, text_ids, resp_ids, loss_scale: float = 1.0,
                         plan: Optional[StepPlan] = None, generator=None, draw=None,
                         num_audio_embeds: Optional[int] = None, lengths=None, scaler=None,
                         layer_events=None, comm_sms: int = 0) -> Dict[str, torch.Tensor]:
        """One training micro-batch (REF/trainer.py:270-374): encoder forward with kept activations -> LLM
        forward/backward -> encoder backward. Parameter gradients (x loss_scale x the scaler's scale) accumulate inside
        the encoder until `audio_encoder.flush_grads()`. `generator` / `draw` feed the encoder's train-mode
        regularisers (AudioEncoder.forward_train); `layer_events`: see AudioEncoder.backward."""
        if not waves.is_cuda:
            raise RuntimeError("AudioPromptStep needs CUDA inputs; there is no CPU path")
        kw = {}
        if generator is not None or draw is not None:
            kw.update(generator=generator, draw=draw)
        if lengths is not None:  # ragged batch: waves zero-padded on the right, samples per utterance in `lengths`
            kw.update(lengths=lengths)
        audio = self.audio_encoder.forward_train(waves, **kw)
        n_valid = (self.audio_encoder._train_ctx or {}).get("n_valid") if lengths is not None else None
        A_full = audio.shape[1]
        if num_audio_embeds is not None and num_audio_embeds < A_full:  # REF/trainer.py:280-291 (see forward_losses)
            audio = audio[:, :num_audio_embeds].contiguous()
        out = self.llm_forward_backward(audio, text_ids, resp_ids, loss_scale=loss_scale, plan=plan, n_audio=n_valid,
                                        scaler=scaler)
        d = out["d_audio_embeds"]
        if d.shape[1] < A_full:  # the cropped embeddings get no gradient
            d = torch.nn.functional.pad(d, (0, 0, 0, A_full - d.shape[1]))
        if layer_events is not None and comm_sms > 0:
            # gradient buckets are all-reduced on a communication stream WHILE this backward runs: the persistent GEMMs
            # must not count on the SMs the NCCL kernel occupies (b2s_set_sm_budget)
            from . import _lib
            lib = _lib.load()
            prop = torch.cuda.get_device_properties(waves.device)
            lib.b2s_set_sm_budget(max(2, prop.multi_processor_count - int(comm_sms)))
            try:
                self.audio_encoder.backward(d, layer_events=layer_events)
            finally:
                lib.b2s_set_sm_budget(0)
        else:
            self.audio_encoder.backward(d, layer_events=layer_events)
        return out

    def __call__(self, waves_host: torch.Tensor, text_ids, resp_ids, device) -> Dict[str, float]:
        """End-to-end call from HOST buffers: pinned H2D copy of the waveforms and ids, the fused step, and a D2H
        read of the per-utterance losses (what bench.py's `e2e` times)."""
        return self.submit(waves_host, text_ids, resp_ids, device).result()

    def submit(self, waves_host: torch.Tensor, text_ids, resp_ids, device) -> PendingStep:
        """Streaming form of `__call__`: enqueue H2D (side stream) + the fused step + an asynchronous D2H of the
        per-utterance losses, and return at once. Submitting batch i+1 before calling `result()` of batch i hides the
        host-side plan building and the PCIe copies behind the GPU work of batch i."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device)
        waves = _stage_to_device(waves_host, device, self._copy_stream)
        return _losses_to_host(self.forward_losses(waves, text_ids, resp_ids))
