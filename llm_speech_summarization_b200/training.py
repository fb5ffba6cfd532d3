"""Training step of the reference's Trainer (REF/trainer.py:98-110, 250-395, 518-526) on the B200 path.

    FlatAdamW     torch.optim.AdamW(audio_encoder.parameters(), lr, betas) (REF/trainer.py:98-105) with every
                  trainable tensor re-homed as a view of ONE flat fp32 buffer (params, grads, exp_avg, exp_avg_sq):
                  the update is one b2s_adamw_step launch and the data-parallel gradient exchange one SUM
                  all-reduce. state_dict()/load_state_dict() keep torch.optim.AdamW's layout, so the reference's
                  checkpoints (`optimizer` key, REF/trainer.py:518-526) load unchanged.
    PolynomialLR  torch.optim.lr_scheduler.PolynomialLR(power=1.0) (REF/trainer.py:106-110), closed form.
    EncoderTrainer.train_step
                  one micro-batch: encoder forward (activations kept) -> splice -> frozen-LLM student+teacher pass ->
                  CE/KD/FD losses -> backward to the encoder's parameters, scaled by 1/grad_accum_interval
                  (REF/trainer.py:372-374); optimizer.step / scheduler.step / zero_grad every grad_accum_interval
                  micro-batches (REF/trainer.py:376-384). Under torch.distributed every rank runs its shard of the
                  accumulation window and the flat gradient is SUM-all-reduced before the update.

The LLM is frozen (REF/trainer.py:62-64); its parameters sit in the reference optimizer's second param group without
ever receiving a gradient, so they carry no optimizer state -- FlatAdamW reproduces that group as an empty-state one.
`regularize=True` switches on HF's train-mode dropout / LayerDrop / SpecAugment for the HuBERT encoder (what the
reference's `audio_encoder.train()` does, REF/trainer.py:258; regularizers.py, csrc/rng.cuh); the default is the
deterministic step.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib


class GradScaler:
    """torch.cuda.amp.GradScaler (REF/trainer.py:252,374,381-382) with its state -- scale, growth tracker, the
    found-inf flag of the step in flight and the count of optimizer steps actually taken -- in DEVICE memory
    (include/b2s.h: b2s_grad_scaler_state), so neither the skip-on-overflow decision nor the scale update
    synchronises the host. fp16 operands mean fp16 gradients (a mixed bf16 x fp16 tcgen05.mma traps), and fp16
    gradients need the scale exactly as they do in the reference.

        loss gradient  *= scale        where it enters the backward pass (b2s_kd_ce_loss_bwd, the FD taps)
        optimizer step : check(flat gradient) -> AdamW(..., scaler) [no-op on overflow] -> update()

    enabled=False (bf16 operands): the scale is pinned to 1 and only the step counter is used."""

    def __init__(self, device, init_scale: float = 65536.0, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000, enabled: bool = True):
        if torch.device(device).type != "cuda":
            raise RuntimeError("GradScaler (B200 path) keeps its state on a CUDA device; there is no CPU path")
        self.enabled = bool(enabled)
        self.growth_factor = float(growth_factor) if enabled else 1.0
        self.backoff_factor = float(backoff_factor) if enabled else 1.0
        self.growth_interval = int(growth_interval)
        self.state = torch.zeros(8, device=device, dtype=torch.int32)
        self.state[:1].view(torch.float32).fill_(float(init_scale) if enabled else 1.0)

    @property
    def scale_tensor(self) -> torch.Tensor:
        """fp32 [1] view of the current scale (what the backward kernels read)."""
        return self.state[:1].view(torch.float32)

    def data_ptr(self) -> int:
        return self.state.data_ptr()

    def check(self, flat_grad: torch.Tensor) -> None:
        """found_inf |= any(non-finite) over the (already all-reduced) flat gradient."""
        if self.enabled:
            _lib.check(_lib.load().b2s_nonfinite_check(flat_grad.data_ptr(), flat_grad.numel(), self.state.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream), "nonfinite_check")

    def update(self) -> None:
        _lib.check(_lib.load().b2s_grad_scaler_update(self.state.data_ptr(), self.growth_factor, self.backoff_factor,
                                                      self.growth_interval, torch.cuda.current_stream().cuda_stream),
                   "grad_scaler_update")

    def read(self) -> Dict[str, float]:
        """Host copy of the state (synchronises; logging / checkpoints / tests only)."""
        host = self.state.cpu()
        return {"scale": float(host[:1].view(torch.float32)[0]), "growth_tracker": int(host[1]),
                "found_inf": int(host[2]), "opt_steps": int(host[3]), "skipped_steps": int(host[4])}

    def get_scale(self) -> float:
        return self.read()["scale"]

    def set_opt_steps(self, n: int) -> None:
        self.state[3] = int(n)

    def state_dict(self) -> Dict:
        r = self.read()
        return {"scale": r["scale"], "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor,
                "growth_interval": self.growth_interval, "_growth_tracker": r["growth_tracker"]}

    def load_state_dict(self, sd: Dict) -> None:
        if self.enabled:
            self.state[:1].view(torch.float32).fill_(float(sd["scale"]))
            self.state[1] = int(sd.get("_growth_tracker", 0))


class FlatAdamW:
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 frozen_params=(), exclude=(), order=None):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        skip = {id(p) for p in exclude}  # parameters that never receive a gradient: torch keeps no state for them
        self.trainable = [p for p in self.params if p.requires_grad and id(p) not in skip]
        if not self.trainable:
            raise ValueError("FlatAdamW: no trainable parameters")
        dev = self.trainable[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdamW (B200 path) needs CUDA parameters; there is no CPU path")
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.lr = lr
        self.n_frozen = len(list(frozen_params))
        n = sum(p.numel() for p in self.trainable)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        if order is not None:  # layout of the flat buffers (state_dict indices still follow `params`)
            keep = {id(p) for p in self.trainable}
            self.trainable = [p for p in order if id(p) in keep]
            assert len(self.trainable) == len(keep), "FlatAdamW: `order` must cover every trainable parameter"
        self.offsets = []
        off = 0
        for p in self.trainable:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1).float())
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            self.offsets.append(off)
            off += k
        self.step_count = 0
        self.scaler: Optional[GradScaler] = None  # set by the trainer: device-side step count + loss scale
        self.last_allreduce = None  # (events of the last exchange) for `allreduce_ms`

    def zero_grad(self, set_to_none: bool = False):
        self.grad.zero_()

    def span_of(self, params) -> Optional[tuple]:
        """(start, end) element range of the flat buffers covered by `params` if they are contiguous there."""
        off_of = {id(p): o for p, o in zip(self.trainable, self.offsets)}
        spans = sorted((off_of[id(p)], off_of[id(p)] + p.numel()) for p in params if id(p) in off_of)
        if not spans:
            return None
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            if a1 != b0:
                return None
        return spans[0][0], spans[-1][1]

    @staticmethod
    def _dist_on() -> bool:
        """True when the gradient has to be exchanged; makes sure the library's own communicator exists."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return False
        from . import dp
        return dp.ensure_native_comm() > 1

    @staticmethod
    def _raw_event(ev: torch.cuda.Event) -> int:
        if not getattr(ev, "_b2s_primed", False):  # a torch event owns a CUDA event only once it has been recorded
            ev.record()
            ev._b2s_primed = True
        return ev.cuda_event

    def all_reduce_grads(self):
        """One SUM all-reduce of the whole flat gradient, joined at once (no overlap)."""
        if self._dist_on():
            ev = torch.cuda.Event()
            ev.record()
            ev._b2s_primed = True
            self.all_reduce_buckets([(0, self.grad.numel())], [ev])
            self.finish_all_reduce()

    def all_reduce_buckets(self, buckets, ready_events) -> None:
        """SUM all-reduce of gradient ranges `buckets` = [(start, end), ...] through the C ABI (b2s_allreduce_grads) on
        the library's communication stream, each bucket as soon as `ready_events[i]` (recorded on the compute stream
        when the range's last gradient kernel was enqueued) has fired: the exchange of layer l's gradients runs under
        the backward of layers l-1 ... 0 (buckets per encoder layer; NVSwitch bandwidth is uniform, so
        bucket = layer). Several calls form one exchange; `finish_all_reduce()` makes the compute stream join."""
        if not self._dist_on():
            return
        import ctypes as C
        keep = [(a, b, ev) for (a, b), ev in zip(buckets, ready_events) if b > a]
        if not keep:
            return
        n = len(keep)
        starts = (C.c_int64 * n)(*[a for a, _, _ in keep])
        ends = (C.c_int64 * n)(*[b for _, b, _ in keep])
        evs = (C.c_void_p * n)(*[self._raw_event(ev) for _, _, ev in keep])
        pend = getattr(self, "_pending_comm", None)
        t0 = None
        if pend is None:
            t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        _lib.check(_lib.load().b2s_allreduce_grads(self.grad.data_ptr(), starts, ends, evs, n,
                                                   None if t0 is None else self._raw_event(t0), self._raw_event(t1)),
                   "allreduce_grads")
        self._pending_comm = (pend[0] if pend is not None else t0, t1)

    def finish_all_reduce(self) -> None:
        """Compute stream waits for every bucket launched by `all_reduce_buckets` (b2s_allreduce_join); records what
        bench.py reports: the communication-stream span of the exchange and how long the compute stream stood still."""
        pend = getattr(self, "_pending_comm", None)
        if pend is None:
            return
        t0, t1 = pend
        main = torch.cuda.current_stream()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(main)
        _lib.check(_lib.load().b2s_allreduce_join(main.cuda_stream), "allreduce_join")
        w1.record(main)
        self.last_allreduce = (t0, t1, w0, w1)
        self._pending_comm = None

    def allreduce_ms(self) -> Optional[Dict[str, float]]:
        """{'span': first bucket start -> last bucket end on the communication stream, 'exposed': time the compute
        stream waited} of the last exchange (synchronises on its events)."""
        if self.last_allreduce is None:
            return None
        t0, t1, w0, w1 = self.last_allreduce
        w1.synchronize()
        t1.synchronize()
        return {"span": float(t0.elapsed_time(t1)), "exposed": float(w0.elapsed_time(w1))}

    @torch.no_grad()
    def step(self):
        """AdamW on the flat buffers. With a GradScaler attached: inf/nan check -> update (skipped on overflow, gradient
        divided by the loss scale, bias correction from the device-side count of steps taken) -> scale update."""
        self.step_count += 1
        b1, b2 = self.defaults["betas"]
        sc = self.scaler
        if sc is not None:
            sc.check(self.grad)
        _lib.check(_lib.load().b2s_adamw_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), self.flat.numel(), float(self.lr), b1, b2,
                                              self.defaults["eps"], self.defaults["weight_decay"], self.step_count,
                                              1.0, None if sc is None else sc.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "adamw_step")
        if sc is not None:
            sc.update()

    def steps_taken(self) -> int:
        """Optimizer steps actually applied (overflow-skipped steps do not count; synchronises with a scaler)."""
        return self.step_count if self.scaler is None else self.scaler.read()["opt_steps"]

    # ---- torch.optim.AdamW-compatible state dict ------------------------------------------------
    def state_dict(self) -> Dict:
        state = {}
        idx_of = {id(p): i for i, p in enumerate(self.params)}
        taken = self.steps_taken()
        if taken > 0:
            for p, off in zip(self.trainable, self.offsets):
                k = p.numel()
                state[idx_of[id(p)]] = {"step": torch.tensor(float(taken)),
                                        "exp_avg": self.exp_avg[off:off + k].view(p.shape).clone(),
                                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view(p.shape).clone()}
        group = dict(lr=self.lr, betas=self.defaults["betas"], eps=self.defaults["eps"],
                     weight_decay=self.defaults["weight_decay"], amsgrad=False, maximize=False, foreach=None,
                     capturable=False, differentiable=False, fused=None, initial_lr=self.defaults["lr"])
        n = len(self.params)
        groups = [dict(group, params=list(range(n))), dict(group, params=list(range(n, n + self.n_frozen)))]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: Dict):
        idx_of = {id(p): i for i, p in enumerate(self.params)}
        steps = set()
        for p, off in zip(self.trainable, self.offsets):
            st = sd["state"].get(idx_of[id(p)])
            if st is None:
                continue
            k = p.numel()
            self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1).to(self.exp_avg.device, torch.float32))
            self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1).to(self.exp_avg.device, torch.float32))
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("FlatAdamW.load_state_dict: parameters with different step counts are not supported")
        self.step_count = steps.pop() if steps else 0
        if self.scaler is not None:
            self.scaler.set_opt_steps(self.step_count)
        if sd.get("param_groups"):
            self.lr = sd["param_groups"][0].get("lr", self.lr)


class PolynomialLR:
    """lr_t = base_lr * (1 - min(t, total_iters) / total_iters) ** power (torch's PolynomialLR, closed form)."""

    def __init__(self, optimizer: FlatAdamW, total_iters: int, power: float = 1.0):
        self.optimizer = optimizer
        self.total_iters = max(1, int(total_iters))
        self.power = power
        self.base_lr = optimizer.defaults["lr"]
        self.last_epoch = 0
        self._apply()

    def _apply(self):
        t = min(self.last_epoch, self.total_iters)
        self.optimizer.lr = self.base_lr * (1.0 - t / self.total_iters) ** self.power

    def step(self):
        self.last_epoch += 1
        self._apply()

    def get_last_lr(self):
        return [self.optimizer.lr]

    def state_dict(self):
        return {"total_iters": self.total_iters, "power": self.power, "base_lrs": [self.base_lr],
                "last_epoch": self.last_epoch, "_last_lr": [self.optimizer.lr]}

    def load_state_dict(self, sd):
        self.total_iters = sd.get("total_iters", self.total_iters)
        self.power = sd.get("power", self.power)
        self.base_lr = sd.get("base_lrs", [self.base_lr])[0]
        self.last_epoch = sd.get("last_epoch", 0)
        self._apply()


class EncoderTrainer:
    """The optimisation loop body of REF/trainer.py:250-395 (data loading, logging and validation stay with the
    caller). `step` counts micro-batches like the reference's `self.step`."""

    def __init__(self, step_fn, audio_encoder, llm=None, *, lr: float = 5e-5, betas=(0.9, 0.999),
                 grad_accum_interval: int = 16, total_optimizer_steps: int = 1000, weight_decay: float = 1e-2,
                 regularize: bool = False, generator: Optional[torch.Generator] = None):
        self.step_fn = step_fn
        self.audio_encoder = audio_encoder
        self.generator = generator
        frozen = list(llm.parameters()) if llm is not None else []
        spec_augment = False
        if regularize and getattr(audio_encoder, "encoder_base", None) == "hubert":
            from .regularizers import RegularizerConfig
            audio_encoder.regularizers = RegularizerConfig.from_arch(audio_encoder.encoder.arch)
            audio_encoder.train()
            cfg = audio_encoder.regularizers
            spec_augment = cfg.apply_spec_augment and cfg.mask_time_prob > 0
        # masked_spec_embed is only read by SpecAugment: without it the parameter never gets a gradient
        unused = [] if spec_augment else [p for n, p in audio_encoder.named_parameters()
                                          if n.endswith("masked_spec_embed")]
        lr, betas = float(lr), (float(betas[0]), float(betas[1]))  # yaml scalars may arrive as strings ('5e-5')
        self.optimizer = FlatAdamW(audio_encoder.parameters(), lr=lr, betas=betas, weight_decay=weight_decay,
                                   frozen_params=frozen, exclude=unused,
                                   order=getattr(audio_encoder, "flat_param_order", lambda: None)())
        self.lr_scheduler = PolynomialLR(self.optimizer, total_iters=total_optimizer_steps, power=1.0)
        self.grad_accum_interval = int(grad_accum_interval)
        self.step = 0
        self.start_epoch = 0
        self._micro = 0
        # fp16 operands anywhere on the path -> dynamic loss scaling like the reference's GradScaler (REF/trainer.py:252);
        # an all-bf16 path keeps the scale at 1 and uses only the device-side step counter
        llm_dt = getattr(llm, "dtype", None) if llm is not None else getattr(getattr(step_fn, "llm", None), "dtype", None)
        fp16 = torch.float16 in (llm_dt, getattr(audio_encoder, "operand_dtype", None))
        dev = self.optimizer.flat.device
        self.scaler = GradScaler(dev, enabled=fp16)
        self.optimizer.scaler = self.scaler
        # gradient exchange: one bucket per transformer layer, launched on a communication stream as soon as the
        # layer's backward has been enqueued (overlap with the layers below), plus one bucket for everything else
        self.overlap_allreduce = True
        # SMs left to the NCCL kernels while buckets are exchanged under the encoder backward (= NCCL_MAX_CTAS set by
        # dp.init_process_group); the persistent GEMMs of that backward size their grids for the rest
        import os
        self.comm_sms = int(os.environ.get("NCCL_MAX_CTAS", "0") or 0)
        self._layer_spans = None
        groups = getattr(audio_encoder, "layer_param_groups", lambda: [])()
        spans = [self.optimizer.span_of(g) for g in groups]
        if groups and all(sp is not None for sp in spans):
            self._layer_spans = spans
            self._layer_events = [torch.cuda.Event() for _ in spans]
            covered = sorted(spans)
            n = self.optimizer.grad.numel()
            rest, cur = [], 0
            for a, b in covered:
                if a > cur:
                    rest.append((cur, a))
                cur = max(cur, b)
            if cur < n:
                rest.append((cur, n))
            self._rest_spans = rest

    @classmethod
    def from_config(cls, config, step_fn, audio_encoder, llm=None, batches_per_epoch: int = 1):
        t = config.train
        total = (t.epochs * batches_per_epoch) // t.grad_accum_interval
        return cls(step_fn, audio_encoder, llm, lr=t.optimizer.lr, betas=(t.optimizer.beta1, t.optimizer.beta2),
                   grad_accum_interval=t.grad_accum_interval, total_optimizer_steps=total)

    def world(self) -> int:
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    @torch.no_grad()
    def train_step(self, waves: torch.Tensor, text_ids, resp_ids, last_batch: bool = False,
                   plan=None, num_audio_embeds: Optional[int] = None, lengths=None) -> Dict[str, torch.Tensor]:
        """One micro-batch of B utterances on this rank. The accumulation window counts utterances GLOBALLY:
        an optimizer step happens once world * (micro-batches * B) reaches grad_accum_interval (or at loader end)."""
        B = waves.shape[0]
        closes_window = self._micro + B * self.world() >= self.grad_accum_interval or last_batch
        overlap = (closes_window and self.overlap_allreduce and self._layer_spans is not None and self.world() > 1)
        out = self.step_fn.forward_backward(waves, text_ids, resp_ids, loss_scale=1.0 / self.grad_accum_interval,
                                            plan=plan, generator=self.generator, scaler=self.scaler,
                                            layer_events=self._layer_events if overlap else None,
                                            comm_sms=self.comm_sms if overlap else 0,
                                            **({} if num_audio_embeds is None else {"num_audio_embeds": num_audio_embeds}),
                                            **({} if lengths is None else {"lengths": lengths}))
        self._micro += B * self.world()
        self.step += 1
        out["optimizer_step"] = False
        if closes_window:
            if overlap:
                # the window's last backward is enqueued: layer L-1's gradients are final first, layer 0's last
                L = len(self._layer_spans)
                order = list(range(L - 1, -1, -1))
                self.optimizer.all_reduce_buckets([self._layer_spans[l] for l in order],
                                                  [self._layer_events[l] for l in order])
            self.audio_encoder.flush_grads()  # conv / positional-conv scratch -> .grad (the "rest" ranges)
            if overlap:
                ev = torch.cuda.Event()
                ev.record()
                self.optimizer.all_reduce_buckets(self._rest_spans, [ev] * len(self._rest_spans))
                self.optimizer.finish_all_reduce()
            else:
                self.optimizer.all_reduce_grads()
            if getattr(self, "capture_grad", False):  # measurement hook (bench.py check.grad_parity, tests)
                self.last_flat_grad = self.optimizer.grad.clone()
                self.last_flat_grad_scale = self.scaler.scale_tensor.clone()
            self.optimizer.step()
            self.audio_encoder.mark_weights_changed()
            self.lr_scheduler.step()
            self.optimizer.zero_grad()
            self._micro = 0
            out["optimizer_step"] = True
        return out

    def __call__(self, waves_host: torch.Tensor, text_ids, resp_ids, device) -> Dict[str, torch.Tensor]:
        """End-to-end micro-batch from HOST buffers: pinned H2D copy of the waveforms and ids, the training step, and
        a D2H read of the per-utterance losses (what bench.py's `e2e` times for the training workload)."""
        return self.submit(waves_host, text_ids, resp_ids, device).result()

    def submit(self, waves_host: torch.Tensor, text_ids, resp_ids, device, lengths=None):
        """Streaming form of `__call__` (see AudioPromptStep.submit): returns a PendingStep whose `result()` holds the
        losses and `optimizer_step`."""
        from .step import _losses_to_host, _stage_to_device
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device)
        out = self.train_step(_stage_to_device(waves_host, device, self._copy_stream), text_ids, resp_ids,
                              lengths=lengths)
        pending = _losses_to_host(out)
        pending.extra["optimizer_step"] = out["optimizer_step"]
        return pending

    @torch.no_grad()
    def validate(self, batches, epoch: int = 0, *, num_generate_samples: int = 0, tokenizer=None, writer=None,
                 save_path: Optional[str] = None) -> Dict:
        """REF/trainer.py:400-528. `batches` yields (waves CUDA (B, T0), text_ids, resp_ids[, num_audio_embeds]).
        Encoder in eval mode; per-utterance next-token NLL of the audio-prompt and the text-prompt sequence; perplexity
        = exp(mean NLL) over the whole set (:502-505); for the first `num_generate_samples` utterances both prompts
        are decoded greedily with max_new_tokens = 2 * (number of audio embeddings) (:460-476, :530-537); the
        checkpoint of :516-526 is written to `save_path` when given. `writer` (optional) receives the reference's
        LogWriter calls. The encoder's train/eval mode is restored on exit."""
        from .utils import merge_prompt_tokens
        enc, llm = self.audio_encoder, self.step_fn.llm
        was_training = enc.training
        enc.eval()
        audio_nlls, text_nlls, audio_resp, text_resp = [], [], [], []
        seen = 0
        try:
            for batch in batches:
                waves, text_ids, resp_ids = batch[:3]
                out = self.step_fn.validation_losses(waves, text_ids, resp_ids,
                                                     num_audio_embeds=batch[3] if len(batch) > 3 else None)
                audio_nlls.append(out["audio_ntp_loss"])
                text_nlls.append(out["text_ntp_loss"])
                if writer is not None:
                    for v in out["audio_ntp_loss"].tolist():
                        writer.log_validation({"ntp_loss": v}, self.step)
                for b in range(waves.shape[0]):
                    if seen >= num_generate_samples:
                        break
                    seen += 1
                    emb = out["audio_embeds"][b:b + 1].to(llm.dtype)
                    n_new = 2 * emb.shape[1]
                    llm_type = self.step_fn.llm_type
                    prompts = [merge_prompt_tokens(emb, tokenizer, llm.model.embed_tokens, llm_type, emb.device),
                               merge_prompt_tokens(llm.model.embed_tokens(torch.as_tensor(text_ids[b])[None].to(
                                   emb.device)), tokenizer, llm.model.embed_tokens, llm_type, emb.device)]
                    for dst, pe in zip((audio_resp, text_resp), prompts):
                        ids = llm.generate(input_ids=None, inputs_embeds=pe, max_new_tokens=n_new)
                        dst.append(tokenizer.batch_decode(ids, skip_special_tokens=True,
                                                          clean_up_tokenization_spaces=True)[0]
                                   if hasattr(tokenizer, "batch_decode") else ids[0].tolist())
        finally:
            enc.train(was_training)
        res = {"audio_perplexity": float(torch.exp(torch.cat(audio_nlls).mean())),
               "text_perplexity": float(torch.exp(torch.cat(text_nlls).mean())),
               "audio_nlls": torch.cat(audio_nlls), "text_nlls": torch.cat(text_nlls),
               "audio_responses": audio_resp, "text_responses": text_resp}
        if writer is not None:
            writer.log_validation_perplexity(res["audio_perplexity"], "audio", self.step)
            writer.log_validation_perplexity(res["text_perplexity"], "text", self.step)
        if save_path is not None:
            torch.save(self.checkpoint(epoch), save_path)
        return res

    def checkpoint(self, epoch: int, legacy_weight_norm_keys: bool = False) -> Dict:
        """Same keys as REF/trainer.py:518-526. `legacy_weight_norm_keys=True` spells the positional conv's weight-norm
        parameters `weight_g` / `weight_v` (what the reference's pinned torch 2.0 writes and expects); the default is
        the `parametrizations.weight.original0/1` spelling of current torch. Loading accepts both."""
        enc_sd = {k: v.detach().clone() for k, v in self.audio_encoder.state_dict().items()}
        if legacy_weight_norm_keys:
            for new, old in (("parametrizations.weight.original0", "weight_g"),
                             ("parametrizations.weight.original1", "weight_v")):
                for k in [k for k in enc_sd if k.endswith(new)]:
                    enc_sd[k[:-len(new)] + old] = enc_sd.pop(k)
        return {"audio_encoder": enc_sd,
                "optimizer": self.optimizer.state_dict(), "lr_scheduler": self.lr_scheduler.state_dict(),
                "epoch": epoch, "step": self.step}

    def load_checkpoint(self, ckpt: Dict):
        """REF/trainer.py:116-122. load_state_dict copies INTO the flat views, so the optimizer keeps its buffers."""
        self.audio_encoder.load_state_dict(ckpt["audio_encoder"])
        self.audio_encoder.mark_weights_changed()
        self.optimizer.load_state_dict(ckpt["optimizer"])
        self.lr_scheduler.load_state_dict(ckpt["lr_scheduler"])
        self.start_epoch = ckpt["epoch"]
        self.step = ckpt["step"]
