"""Drop-in for the reference's `Trainer(args, config, device).train()` (REF/trainer.py:24-132 constructor, :134-248
collate + dataloaders, :250-398 train loop, :400-528 validate, :530-545 generate) on the B200 path.

What stays the reference's: the constructor signature and attribute names, the yaml keys it reads, the collate
functions' output tuple (including the BOS-stripping quirks of :154-156), the loop structure (loss / accumulate /
optimizer + scheduler step every `grad_accum_interval` batches or at loader end, `self.step` counts batches, logging
and validation intervals), the checkpoint dictionary and the LogWriter calls.

What is different underneath: the batch goes through `EncoderTrainer.train_step` (one fused forward + backward of the
hand-written CUDA path, no autograd; fp16 operands and gradients like the reference's autocast, with its GradScaler
semantics kept on the device by `training.GradScaler`), the optimizer is the flat AdamW with
`torch.optim.AdamW`'s state layout, the Whisper log-mel features are computed on the GPU instead of inside the collate
function, and under `torch.distributed` every rank takes its shard of the utterances (DistributedSampler) with one SUM
all-reduce per optimizer step. A loader batch of utterances of DIFFERENT lengths is one ragged micro-batch (HuBERT:
the zero-padded waveforms plus their lengths go to `b2s_hubert_forward_train`, which keeps every utterance's batch-1
numerics; Whisper inputs are fixed 30 s windows anyway) -- the reference's own batch_size > 1 path attends over the
zero padding without a mask and is not reproduced.

Offline use (tests, benchmarks): `tokenizer`, `llm`, `train_dataset`, `val_dataset` and `writer` can be injected; left
as None they are loaded exactly like the reference does (AutoTokenizer / from_pretrained / datasets.load_from_disk /
TensorBoard).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .model.audio_encoder import AudioEncoder
from .model.audio_llama import AudioLlamaForCausalLM
from .step import AudioPromptStep
from .training import EncoderTrainer
from .utils import compute_num_audio_embeds

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")


class NullWriter:
    """LogWriter with the reference's method names (REF/writer.py) that keeps the scalars in memory."""

    def __init__(self):
        self.scalars: Dict[str, List] = {}
        self.texts: Dict[str, List] = {}

    def _add(self, tag, value, step):
        self.scalars.setdefault(tag, []).append((int(step), float(value)))

    def log_lr(self, lr, step):
        self._add("learning_rate", lr, step)

    def log_training(self, losses, step):
        for k, v in losses.items():
            self._add(f"train/{k}", v, step)

    def log_validation(self, losses, step):
        for k, v in losses.items():
            self._add(f"validation/{k}", v, step)

    def log_validation_perplexity(self, perplexity, input_type, step):
        self._add(f"validation/{input_type}_perplexity", perplexity, step)

    def log_audio_text_responses(self, prompt_audios, prompt_texts, audio_responses, text_responses, step):
        for i, (t, a, b) in enumerate(zip(prompt_texts, audio_responses, text_responses)):
            self.texts.setdefault(f"prompt_texts/prompt_{i}", []).append((int(step), t))
            self.texts.setdefault(f"llm_audio_responses/response_{i}", []).append((int(step), a))
            self.texts.setdefault(f"llm_text_responses/response_{i}", []).append((int(step), b))


def _make_writer(config, log_dir):
    try:
        from torch.utils.tensorboard import SummaryWriter
    except Exception:  # tensorboard not installed: keep the numbers in memory
        return NullWriter()

    class MyWriter(SummaryWriter):  # REF/writer.py
        def __init__(self, config, logdir):
            super().__init__(logdir)
            self.sample_rate = config.audio.sampling_rate

        def log_lr(self, lr, step):
            self.add_scalar("learning_rate", lr, step)

        def log_training(self, losses, step):
            for k, v in losses.items():
                self.add_scalar(f"train/{k}", v, step)

        def log_validation(self, losses, step):
            for k, v in losses.items():
                self.add_scalar(f"validation/{k}", v, step)

        def log_validation_perplexity(self, perplexity, input_type, step):
            self.add_scalar(f"validation/{input_type}_perplexity", perplexity, step)

        def log_audio_text_responses(self, prompt_audios, prompt_texts, audio_responses, text_responses, step):
            for i, (audio, text, ar, tr) in enumerate(zip(prompt_audios, prompt_texts, audio_responses, text_responses)):
                self.add_audio(f"prompt_audios/audio_{i}", audio, step, self.sample_rate)
                self.add_text(f"prompt_texts/prompt_{i}", text, step)
                self.add_text(f"llm_audio_responses/response_{i}", ar, step)
                self.add_text(f"llm_text_responses/response_{i}", tr, step)

    return MyWriter(config, log_dir)


WHISPER_WINDOW_SAMPLES = 480000  # WhisperFeatureExtractor pads / truncates every clip to 30 s


class Trainer():
    def __init__(self, args, config, device, *, tokenizer=None, llm=None, train_dataset=None, val_dataset=None,
                 writer=None, regularize: bool = True, generator: Optional[torch.Generator] = None) -> None:
        self.args = args
        self.config = config
        self.run_name = args.run_name
        self.device = device
        if torch.device(device).type != "cuda":
            raise RuntimeError("Trainer (B200 path) needs a CUDA device; there is no CPU path")

        torch.cuda.manual_seed(self.config.seed_everything)  # REF/trainer.py:32

        self.checkpoint_save_dir = os.path.join(self.config.log.checkpoint_dir, self.run_name)
        self.log_dir = os.path.join(self.config.log.log_dir, self.run_name)
        os.makedirs(self.checkpoint_save_dir, exist_ok=True)
        os.makedirs(self.log_dir, exist_ok=True)
        self.writer = writer if writer is not None else _make_writer(self.config, self.log_dir)

        self.encoder_base = self.config.model.audio_encoder.base
        self.audio_encoder = AudioEncoder(self.config, self.device)

        self.llm_type = self.config.model.llm_type
        if tokenizer is None:
            from transformers import AutoTokenizer
            tokenizer = AutoTokenizer.from_pretrained(self.llm_type, use_fast=False, padding_side="left")
            tokenizer.pad_token = tokenizer.eos_token
        self.tokenizer = tokenizer
        if llm is None:
            llm = AudioLlamaForCausalLM.from_pretrained(self.llm_type, use_cache=True, torch_dtype=torch.float16)
        self.llm = llm.eval()
        for param in self.llm.parameters():
            param.requires_grad = False

        self.train_dataset, self.val_dataset = train_dataset, val_dataset
        self.get_dataloaders()

        t = self.config.train
        self.use_ld_loss, self.use_fd_loss = t.use_ld_loss, t.use_fd_loss
        self.ntp_loss_weight, self.ld_loss_weight, self.fd_loss_weight = t.ntp_loss_weight, t.ld_loss_weight, t.fd_loss_weight
        self.fd_loss_connector_layers = t.fd_loss_connector_layers

        self.audio_encoder.to(self.device)
        self.llm.to(self.device)

        self.grad_accum_interval = t.grad_accum_interval
        self.num_epochs = t.epochs

        self.step_fn = AudioPromptStep.from_config(self.config, self.audio_encoder, self.llm, self.tokenizer)
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # the accumulation window in UTTERANCES (the reference counts batches of batch_size utterances, one rank)
        self.window = self.grad_accum_interval * t.batch_size * world
        total_iters = self.num_epochs * len(self.train_dataloader) // self.grad_accum_interval
        self.core = EncoderTrainer(self.step_fn, self.audio_encoder, self.llm, lr=t.optimizer.lr,
                                   betas=(t.optimizer.beta1, t.optimizer.beta2), grad_accum_interval=self.window,
                                   total_optimizer_steps=total_iters, regularize=regularize, generator=generator)
        self.optimizer = self.core.optimizer
        self.lr_scheduler = self.core.lr_scheduler
        self.step = 0
        self.start_epoch = 0

        if getattr(self.args, "checkpoint_path", None):
            self.load_checkpoint(self.args.checkpoint_path)

    # ------------------------------------------------------------------------------------------ checkpoints
    def load_checkpoint(self, checkpoint_path):
        # tensors, numbers, dicts and lists only: a checkpoint is data, never code (torch.load's pickle is restricted)
        checkpoint = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
        self.core.load_checkpoint(checkpoint)
        self.start_epoch = checkpoint["epoch"]
        self.step = checkpoint["step"]
        print(f"Loaded checkpoint from {checkpoint_path}.\n")

    # ------------------------------------------------------------------------------------------ data
    @staticmethod
    def _ids(x):
        return x if torch.is_tensor(x) else torch.as_tensor(x)

    def collate_audio_batch_hubert(self, data):
        """REF/trainer.py:134-166: zero-pad the waveforms on the right to the longest clip; the token ids lose the
        BOS added in preprocessing (`text_input_ids[1:]`, `response_input_ids[0][1:]`) and stay un-padded."""
        raw_audios = [self._ids(x["audio"]["array"]) for x in data]
        audio_len_samples = [len(a) for a in raw_audios]
        max_len = max(audio_len_samples)
        prompt_texts = [x["text"] for x in data]
        ctc_pool_ranges = [x.get("pool_ranges_4") if hasattr(x, "get") else x["pool_ranges_4"] for x in data]
        padded_audios = torch.stack([F.pad(a, (0, max_len - len(a)), mode="constant") for a in raw_audios], dim=0).float()
        text_input_ids = [self._ids(x["text_input_ids"])[1:] for x in data]
        response_input_ids = [self._ids(x["response_input_ids"])[0][1:] for x in data]
        return (raw_audios, padded_audios, audio_len_samples, prompt_texts, text_input_ids, response_input_ids,
                ctc_pool_ranges)

    def collate_audio_batch_whisper(self, data):
        """REF/trainer.py:168-199 with the feature extractor moved to the GPU: the waveforms are padded / truncated to
        the extractor's 30 s window here and turned into log-mel features by `AudioEncoder.extract_features` inside the
        step (b2s_whisper_log_mel), so the DataLoader workers do no STFT."""
        raw_audios = [self._ids(x["audio"]["array"]) for x in data]
        audio_len_samples = [len(a) for a in raw_audios]
        prompt_texts = [x["text"] for x in data]
        ctc_pool_ranges = [x.get("pool_ranges_4") if hasattr(x, "get") else x["pool_ranges_4"] for x in data]
        W = WHISPER_WINDOW_SAMPLES
        padded = torch.stack([F.pad(a[:W].float(), (0, max(0, W - len(a)))) for a in raw_audios], dim=0)
        text_input_ids = [self._ids(x["text_input_ids"])[1:] for x in data]
        response_input_ids = [self._ids(x["response_input_ids"])[0][1:] for x in data]
        return (raw_audios, padded, audio_len_samples, prompt_texts, text_input_ids, response_input_ids,
                ctc_pool_ranges)

    def _load_sets(self, names):
        from datasets import concatenate_datasets, load_from_disk
        sets = []
        for name in names:
            ds = load_from_disk(os.path.join(self.config.data.base_path, name))
            ds.set_format(type="torch")
            sets.append(ds)
        return concatenate_datasets(sets)

    def get_dataloaders(self):
        """REF/trainer.py:201-248 (datasets from `config.data`, shuffled train loader, batch-1 validation loader).
        Under torch.distributed each rank reads its own shard of the shuffled index list."""
        if self.train_dataset is None:
            self.train_dataset = self._load_sets(self.config.data.train_set)
        if self.val_dataset is None:
            self.val_dataset = self._load_sets(self.config.data.val_set)
        collate = self.collate_audio_batch_hubert if self.encoder_base == "hubert" else self.collate_audio_batch_whisper
        t = self.config.train
        sampler = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            sampler = torch.utils.data.distributed.DistributedSampler(self.train_dataset, shuffle=True,
                                                                      seed=self.config.seed_everything)
        self.train_sampler = sampler
        self.train_dataloader = torch.utils.data.DataLoader(
            dataset=self.train_dataset, batch_size=t.batch_size, shuffle=sampler is None, sampler=sampler,
            num_workers=t.num_workers, pin_memory=True, collate_fn=collate)
        self.val_dataloader = torch.utils.data.DataLoader(
            dataset=self.val_dataset, batch_size=1, shuffle=False, num_workers=t.num_workers, pin_memory=True,
            collate_fn=collate)

    # ------------------------------------------------------------------------------------------ steps
    def _micro_batches(self, padded_inputs, audio_len_samples, text_input_ids, response_input_ids):
        """Utterances of one loader batch grouped by length: each group is one packed micro-batch
        (waves (b, T) un-padded, ids, ids, embeddings to keep per utterance or None)."""
        groups: Dict[int, List[int]] = {}
        for i, n in enumerate(audio_len_samples):
            groups.setdefault(int(n), []).append(i)
        out = []
        for n, idx in groups.items():
            if self.encoder_base == "whisper":
                waves = padded_inputs[idx]
                keep = compute_num_audio_embeds(min(n, WHISPER_WINDOW_SAMPLES), sr=self.config.audio.sampling_rate)
            else:
                waves = padded_inputs[idx, :n]
                keep = None  # un-padded: every pooled frame belongs to the utterance (REF/trainer.py:292-294)
            out.append((waves, [text_input_ids[i] for i in idx], [response_input_ids[i] for i in idx], keep))
        return out

    def _encoder_input(self, waves):
        waves = waves.to(self.device, non_blocking=True)
        if self.encoder_base == "whisper":
            return self.audio_encoder.extract_features(waves)
        return waves

    def train(self):
        log = self.config.log
        n_batches = len(self.train_dataloader)
        for epoch in range(self.start_epoch, self.start_epoch + self.num_epochs):
            print(f"Epoch {epoch}")
            self.audio_encoder.train()
            self.optimizer.zero_grad()
            if self.train_sampler is not None:
                self.train_sampler.set_epoch(epoch)
            for batch_idx, (_, padded_inputs, audio_len_samples, _, text_input_ids, response_input_ids,
                            _) in enumerate(self.train_dataloader):
                if self.encoder_base == "hubert":  # one ragged micro-batch (lengths=None when they are all equal)
                    ragged = len(set(int(n) for n in audio_len_samples)) > 1
                    micro = [(padded_inputs, text_input_ids, response_input_ids, None,
                              [int(n) for n in audio_len_samples] if ragged else None)]
                else:
                    micro = [m + (None,) for m in self._micro_batches(padded_inputs, audio_len_samples, text_input_ids,
                                                                      response_input_ids)]
                last_of_loader = batch_idx + 1 == n_batches
                sums: Dict[str, float] = {}
                count = 0
                for j, (waves, t_ids, r_ids, keep, lengths) in enumerate(micro):
                    out = self.core.train_step(self._encoder_input(waves), t_ids, r_ids,
                                               last_batch=last_of_loader and j + 1 == len(micro),
                                               num_audio_embeds=keep, lengths=lengths)
                    if (self.step + 1) % log.log_interval == 0:  # .item()-style syncs only when logging
                        for k in ("ntp_loss", "ld_loss", "fd_loss"):
                            if k in out:
                                sums[k] = sums.get(k, 0.0) + float(out[k].sum())
                        count += len(t_ids)
                self.step += 1
                self.core.step = self.step
                if self.step % log.log_interval == 0 and count:
                    self.writer.log_training({k: v / count for k, v in sums.items()}, self.step)
                    self.writer.log_lr(self.lr_scheduler.get_last_lr()[0], self.step)
                if self.step % log.validation_interval == 0:
                    self.validate(epoch)
            self.validate(epoch)

    def validate(self, epoch):
        """REF/trainer.py:400-528 through EncoderTrainer.validate. The reference is single-process; under
        torch.distributed rank 0 alone validates, logs and writes `epoch_{e}_step_{s}.pt` (every rank holds identical
        parameters after the all-reduced step), the others wait at the barrier -- no duplicate log entries, no
        concurrent torch.save to one path."""
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if distributed and dist.get_rank() != 0:
            dist.barrier()
            return None

        def batches():
            for (_, padded, lens, _, t_ids, r_ids, _) in self.val_dataloader:
                for waves, t, r, _keep in self._micro_batches(padded, lens, t_ids, r_ids):
                    # validation feeds the UN-cropped embeddings (REF/trainer.py:418-447 has no un-padding step, unlike
                    # the training loop's :280-291), so Whisper perplexities are comparable with the reference's
                    yield self._encoder_input(waves), t, r, None

        texts, audios = [], []
        for i, item in enumerate(self.val_dataloader):
            if i >= self.config.log.num_generate_samples:
                break
            texts.append(item[3][0])
            audios.append(item[0][0].squeeze().cpu().numpy())
        save_path = os.path.join(self.checkpoint_save_dir, f"epoch_{epoch}_step_{self.step}.pt")
        self.core.step = self.step
        res = self.core.validate(batches(), epoch, num_generate_samples=self.config.log.num_generate_samples,
                                 tokenizer=self.tokenizer, writer=self.writer, save_path=save_path)
        self.writer.log_audio_text_responses(prompt_audios=audios, prompt_texts=texts,
                                             audio_responses=[str(r) for r in res["audio_responses"]],
                                             text_responses=[str(r) for r in res["text_responses"]], step=self.step)
        print(f"Saved checkpoint for epoch {epoch} to {save_path}.\n")
        if distributed:
            dist.barrier()
        return res

    def generate_llm_response(self, inputs_embeds, len_inputs=60):
        """REF/trainer.py:530-545."""
        with torch.no_grad():
            generate_ids = self.llm.generate(input_ids=None, inputs_embeds=inputs_embeds, max_new_tokens=2 * len_inputs)
        return self.tokenizer.batch_decode(generate_ids, skip_special_tokens=True, clean_up_tokenization_spaces=True)
