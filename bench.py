#!/usr/bin/env python
"""bench.py -- utterances/s of the audio-prompt forward step on N B200 GPUs (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1], the one the metric is quoted on): Llama-3.2-3B + HuBERT-large, synthetic 10 s
16 kHz utterances, random-init weights. One step = one micro-batch of `--batch` utterances per GPU through
    encoder -> splice -> packed student (audio prompt) + teacher (text prompt) prefill -> fused CE + KD (+FD) loss,
i.e. "encoder + prefill + KD loss" (forward). `--workload train` times BASELINE.json configs[2] instead: the same
forward with activations kept, the backward through the frozen LLM into every encoder/projector parameter, the SUM
all-reduce of the flat gradient over the ranks and one AdamW update per step (grad_accum window = the global batch).

`value`  : utterances/s, inputs already resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same through the public streaming call AudioPromptStep.submit(...) / .result() with HOST (pinned)
           inputs: every step's H2D of the waveforms + ids and the D2H read of its per-utterance losses happen inside
           the timed region; batch i+1 is submitted before batch i's losses are read, so the copies and the host-side
           plan building overlap the GPU work (the blocking form is `__call__` = submit(...).result()).
`roofline`: the dominant kernel family (the tcgen05 GEMM): algorithmic FLOPs of every GEMM launch of one step
           divided by the summed per-launch CUDA-event durations (b2s_gemm_timing_*), vs the measured sustained
           bf16 peak in MEASURED_PEAKS.json.
`cpu_baseline`: the oracle (CPU port of the reference path) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line under torchrun

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "utterances/sec (10 s audio, encoder+prefill+KD loss)"
UNIT = "utterances/s"
SAMPLES = 160000
T_TEXT, R_RESP = 40, 64


def ncu_traffic_per_launch(train: bool):
    """DRAM bytes per GEMM launch from the committed ncu launch list of the same command (dram__bytes_read.sum +
    dram__bytes_write.sum, averaged over the GEMM launches of one step); None when the profile is not there."""
    name = "r01_launch_summary_train.txt" if train else "r01_launch_summary.txt"
    path = os.path.join(ROOT, "profiles", name)
    try:
        tot_gb, n = 0.0, 0
        for line in open(path):
            if "gemm_bf16_tcgen05_kernel" in line and "dram=" in line:
                n += int(line.split("n=")[1].split()[0])
                tot_gb += float(line.split("dram=")[1].split()[0])
        if n:
            return tot_gb * 1e9 / n, f"profiles/{name} (ncu, {n} GEMM launches of one step, mean per launch)"
    except Exception:
        pass
    return None, None


def write_gemm_shapes(lib, n, path, steps):
    """Per-shape aggregation of the per-launch CUDA-event timings (hot, inside the real step)."""
    import ctypes as C
    agg = {}
    ms, shape = C.c_double(), (C.c_int32 * 10)()
    for i in range(n):
        lib.b2s_gemm_timing_get(i, C.byref(ms), shape)
        key = tuple(shape)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ms.value
    epi = {0: "bf16", 1: "resid_f32", 2: "swiglu", 3: "rope", 4: "f32", 5: "accum_f32"}
    mode = {0: "fwd", 1: "fwd+", 2: "dgrad", 3: "wgrad"}
    rows = []
    for (M, N, K, b, g, e, act, md, bn, cg), (cnt, t) in agg.items():
        fl = 2.0 * M * N * K * b * g * cnt
        rows.append((t / steps, cnt // steps, M, N, K, b, g, epi.get(e, e), "gelu" if act else "-", mode.get(md, md),
                     bn, cg, fl / (t / 1e3) / 1e12))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    with open(path, "w") as f:
        f.write(f"# per-shape GEMM timings inside the step (CUDA events per launch, hot); {tot:.2f} ms of GEMM per step\n")
        f.write("# ms/step  launches  M N K batches groups  epilogue act mode  bn cg  TFLOP/s\n")
        for r in rows:
            f.write(f"{r[0]:8.3f} {r[1]:4d}  {r[2]:6d} {r[3]:6d} {r[4]:7d} {r[5]:3d} {r[6]:3d}  {r[7]:>9s} {r[8]:>4s} "
                    f"{r[9]:>5s}  {r[10]:3d} {r[11]}  {r[12]:7.1f}\n")


def loss_roofline(step, lib, _lib, resident0, plan0, dev, peaks):
    """The HBM-bound kernel the north star names: the fused CE + KD forward (kd_ce_partial + kd_ce_finalize) timed alone
    on the step's own logits; 256 MB rewritten before every launch (> 126 MB L2); straight through the C ABI with
    pre-allocated outputs so no allocator work sits between the event pair."""
    import ctypes as C
    kept = step.forward_losses(*resident0, plan=plan0, keep=True)
    s_log, t_log, pl = kept["student_logits"], kept["teacher_logits"], kept["plan"]
    flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
    rows_l, V_l = s_log.shape
    n_utt = pl.row_offsets.numel() - 1
    ws_l = torch.empty(lib.b2s_kd_ce_workspace_bytes(rows_l, V_l), device=dev, dtype=torch.uint8)
    f32 = lambda n: torch.empty(n, device=dev, dtype=torch.float32)
    lse_s, lse_t, ck, cc, ld_o, ntp_o = f32(rows_l), f32(rows_l), f32(rows_l), f32(rows_l), f32(n_utt), f32(n_utt)
    st_l = torch.cuda.current_stream().cuda_stream
    evs, loss_launches = [], 0
    for _ in range(3 + 10):
        flush.zero_()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches_l0 = lib.b2s_launch_count()
        a.record()
        _lib.check(lib.b2s_kd_ce_loss_fwd(s_log.data_ptr(), t_log.data_ptr(), s_log.stride(0), t_log.stride(0), rows_l,
                                          V_l, pl.labels.data_ptr(), pl.row_offsets.data_ptr(), n_utt, C.c_float(0.5),
                                          C.c_float(0.5), ws_l.data_ptr(), lse_s.data_ptr(), lse_t.data_ptr(),
                                          ck.data_ptr(), cc.data_ptr(), ld_o.data_ptr(), ntp_o.data_ptr(), st_l),
                   "kd_ce_loss_fwd")
        b_.record()
        evs.append((a, b_))
        loss_launches = lib.b2s_launch_count() - launches_l0
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b_) for a, b_ in evs[3:])
    loss_ms = ts[len(ts) // 2]
    hbm_peak = float(peaks.get("hbm_gbs", 6400.0))
    alg_bytes = 4.0 * rows_l * V_l  # student + teacher bf16 logits read once
    return {"achieved": alg_bytes / (loss_ms / 1e3) / 1e9, "frac": alg_bytes / (loss_ms / 1e3) / 1e9 / hbm_peak,
            "ms": loss_ms, "launches": int(loss_launches), "rows": int(rows_l), "vocab": int(V_l),
            "algorithmic_bytes": alg_bytes, "peak": hbm_peak}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="prefill", choices=["prefill", "train"],
                    help="prefill = configs[1] (default, the metric's configuration); train = configs[2]")
    ap.add_argument("--regularize", default="none", choices=["none", "dropout", "all"],
                    help="train workload: HF train-mode regularisers of the encoder (REF/trainer.py:258). 'dropout' = "
                         "every dropout site + SpecAugment (same work as the deterministic step plus the mask "
                         "generation), 'all' = also LayerDrop (skips ~10%% of the encoder layers, like the reference)")
    ap.add_argument("--ragged", action="store_true",
                    help="train workload: utterance lengths drawn uniformly from [5 s, 10 s] (one ragged micro-batch "
                         "per step, zero-padded waveforms + lengths) instead of 10 s each")
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit operand format of the encoder and the LLM (fp16 = the reference's autocast dtype and the "
                         "default; bf16 = the round-1 path, for A/B)")
    ap.add_argument("--gemm-shapes", default="", help="write the per-shape GEMM table of the instrumented steps here")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-utts", type=int, default=3, help="utterances timed for the CPU baseline sample")
    ap.add_argument("--profile-mode", action="store_true",
                    help="only warm-up + timed steps (no e2e / roofline / CPU legs): for runs under ncu")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ synthetic model
def synth_weights(device):
    """Random-init weights of the named architectures, generated on the GPU (HF-default distributions:
    Linear N(0, 0.02), conv kaiming-normal, norms ~1/0) in the reference's checkpoint layouts."""
    import math
    from llm_speech_summarization_b200.config import EncoderArch, KNOWN_LLMS
    ea, la = EncoderArch(), KNOWN_LLMS["meta-llama/Llama-3.2-3B-Instruct"]
    g = torch.Generator(device=device).manual_seed(1234)
    rn = lambda *s, std=1.0, mean=0.0, dt=torch.float32: (torch.randn(*s, generator=g, device=device) * std + mean).to(dt)
    enc = {"encoder.masked_spec_embed": rn(ea.hidden)}
    cin = 1
    for i, (co, k) in enumerate(zip(ea.conv_dim, ea.conv_kernel)):
        p = f"encoder.feature_extractor.conv_layers.{i}."
        enc[p + "conv.weight"] = rn(co, cin, k, std=math.sqrt(2.0 / (cin * k)))
        enc[p + "conv.bias"] = rn(co, std=0.05)
        enc[p + "layer_norm.weight"] = rn(co, std=0.1, mean=1.0)
        enc[p + "layer_norm.bias"] = rn(co, std=0.1)
        cin = co
    enc["encoder.feature_projection.layer_norm.weight"] = rn(cin, std=0.1, mean=1.0)
    enc["encoder.feature_projection.layer_norm.bias"] = rn(cin, std=0.1)
    enc["encoder.feature_projection.projection.weight"] = rn(ea.hidden, cin, std=0.02)
    enc["encoder.feature_projection.projection.bias"] = rn(ea.hidden, std=0.02)
    pc = "encoder.encoder.pos_conv_embed.conv."
    v = rn(ea.hidden, ea.hidden // ea.pos_groups, ea.pos_k, std=2 * math.sqrt(1.0 / (ea.pos_k * ea.hidden)))
    enc[pc + "bias"] = rn(ea.hidden, std=0.02)
    enc[pc + "parametrizations.weight.original0"] = v.norm(dim=(0, 1), keepdim=True)
    enc[pc + "parametrizations.weight.original1"] = v
    enc["encoder.encoder.layer_norm.weight"] = rn(ea.hidden, std=0.1, mean=1.0)
    enc["encoder.encoder.layer_norm.bias"] = rn(ea.hidden, std=0.1)
    for l in range(ea.layers):
        p = f"encoder.encoder.layers.{l}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            enc[p + f"attention.{n}.weight"] = rn(ea.hidden, ea.hidden, std=0.02)
            enc[p + f"attention.{n}.bias"] = rn(ea.hidden, std=0.02)
        enc[p + "layer_norm.weight"] = rn(ea.hidden, std=0.1, mean=1.0)
        enc[p + "layer_norm.bias"] = rn(ea.hidden, std=0.1)
        enc[p + "feed_forward.intermediate_dense.weight"] = rn(ea.ffn, ea.hidden, std=0.02)
        enc[p + "feed_forward.intermediate_dense.bias"] = rn(ea.ffn, std=0.02)
        enc[p + "feed_forward.output_dense.weight"] = rn(ea.hidden, ea.ffn, std=0.02)
        enc[p + "feed_forward.output_dense.bias"] = rn(ea.hidden, std=0.02)
        enc[p + "final_layer_norm.weight"] = rn(ea.hidden, std=0.1, mean=1.0)
        enc[p + "final_layer_norm.bias"] = rn(ea.hidden, std=0.1)
    enc["embed_projection.weight"] = rn(la.hidden, ea.hidden, std=0.02)
    enc["embed_projection.bias"] = rn(la.hidden, std=0.02)
    bf = torch.bfloat16
    llm = {"model.embed_tokens.weight": rn(la.vocab, la.hidden, std=0.02, dt=bf)}
    D = la.head_dim
    for l in range(la.layers):
        p = f"model.layers.{l}."
        llm[p + "input_layernorm.weight"] = rn(la.hidden, std=0.1, mean=1.0, dt=bf)
        llm[p + "self_attn.q_proj.weight"] = rn(la.heads * D, la.hidden, std=0.02, dt=bf)
        llm[p + "self_attn.k_proj.weight"] = rn(la.kv_heads * D, la.hidden, std=0.02, dt=bf)
        llm[p + "self_attn.v_proj.weight"] = rn(la.kv_heads * D, la.hidden, std=0.02, dt=bf)
        llm[p + "self_attn.o_proj.weight"] = rn(la.hidden, la.heads * D, std=0.02, dt=bf)
        llm[p + "post_attention_layernorm.weight"] = rn(la.hidden, std=0.1, mean=1.0, dt=bf)
        llm[p + "mlp.gate_proj.weight"] = rn(la.ffn, la.hidden, std=0.02, dt=bf)
        llm[p + "mlp.up_proj.weight"] = rn(la.ffn, la.hidden, std=0.02, dt=bf)
        llm[p + "mlp.down_proj.weight"] = rn(la.hidden, la.ffn, std=0.02, dt=bf)
    llm["model.norm.weight"] = rn(la.hidden, std=0.1, mean=1.0, dt=bf)
    llm["lm_head.weight"] = llm["model.embed_tokens.weight"]
    return enc, llm


class FixedTokenizer:
    """P = 9 prefix ids / S = 6 suffix ids of the Llama-3 template (first = BOS); tokenizer files are unreachable
    offline, so the counts are benchmark constants."""

    def __init__(self, vocab, bos):
        g = torch.Generator().manual_seed(99)
        self.ids = {}
        from llm_speech_summarization_b200.utils import LLAMA_PROMPT_PREFIX, LLAMA_PROMPT_SUFFIX
        self.ids[LLAMA_PROMPT_PREFIX] = [bos] + torch.randint(0, vocab - 256, (8,), generator=g).tolist()
        self.ids[LLAMA_PROMPT_SUFFIX] = [bos] + torch.randint(0, vocab - 256, (5,), generator=g).tolist()

    def __call__(self, text, return_tensors="pt"):
        class O:
            pass
        o = O()
        o.input_ids = torch.tensor([self.ids[text]], dtype=torch.long)
        return o


def synth_batch(batch, vocab, seed):
    g = torch.Generator().manual_seed(seed)
    waves = torch.randn(batch, SAMPLES, generator=g) * 0.1
    text = [torch.randint(0, vocab - 256, (T_TEXT,), generator=g) for _ in range(batch)]
    resp = [torch.randint(0, vocab - 256, (R_RESP,), generator=g) for _ in range(batch)]
    return waves, text, resp


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(n_utts, enc_sd=None, llm_sd=None, steps=1, warmup=0):
    """The oracle (CPU restatement of the reference path, oracle/reference_math.py) on the host cores: per
    utterance encoder forward + student and teacher prefill + CE/KD/FD losses, fp32, batch 1 like the reference."""
    from oracle import configs, reference_math as rm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    enc_cfg, llm_cfg = configs.HUBERT_LARGE, configs.LLAMA32_3B
    if enc_sd is None:
        enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=1234)
        llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4321, dtype=torch.bfloat16)
        llm_sd = {k: v.float() for k, v in llm_sd.items()}
        llm_sd["lm_head.weight"] = llm_sd["model.embed_tokens.weight"]
    tok = configs.stub_tokenizer(llm_cfg)
    utts = [configs.synthetic_utterance(llm_cfg, i, SAMPLES, T=T_TEXT, R=R_RESP) for i in range(n_utts)]
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            for a, t, r in utts:
                rm.train_step_losses(enc_sd, llm_sd, enc_cfg, llm_cfg, tok, a, t, r)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {"value": n_utts / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_utts} utterances x {steps} timed pass(es) (10 s audio, L_audio=200, L_text=117, R=64), "
                      f"fp32 oracle, batch 1, {sec:.2f} s per pass"}, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, sec = cpu_reference(args.cpu_utts, steps=max(1, args.steps), warmup=min(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "utterances_per_step": args.cpu_utts},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


METRIC_TRAIN = "utterances/sec (10 s audio, training step: encoder+prefill+KD loss forward, backward, AdamW)"
WORKLOAD_TRAIN = ("configs[2] Llama-3.2-3B + HuBERT-large training step (encoder/projector trainable, LLM frozen), "
                  "CE + logit-KD + FD loss, synthetic 10 s utterances (L_audio=200, L_text=117, R=64); train-mode "
                  "regularisers per config.regularize (none = deterministic step)")
WORKLOAD = ("configs[1] Llama-3.2-3B + HuBERT-large audio-prompt forward: encoder + packed student&teacher prefill "
            "+ fused CE/KD/FD loss, synthetic 10 s utterances (L_audio=200, L_text=117, R=64)")


# ------------------------------------------------------------------------------------------ main arm
def gemm_flops_per_utt(train=False):
    """Algorithmic FLOPs per utterance credited to the GEMM kernel (attention and conv0
    excluded, LM head on the 2*R consumed rows only). Training adds dgrad + wgrad for the encoder (no dgrad into the
    waveform) and dgrad only for the student sequence of the frozen LLM."""
    conv1 = 2 * 512 * 512 * 3 * 15999
    conv = conv1 + 2 * 512 * 512 * (3 * (7999 + 3999 + 1999) + 2 * (999 + 499))
    N = 499
    enc = conv + 2 * N * 512 * 1024 + 2 * N * 1024 * 64 * 128 + 24 * (2 * N * 1024 * (4 * 1024 + 2 * 4096)) \
        + 2 * 123 * 1024 * 3072
    per_tok = 2 * 3072 * (5120 + 3072 + 2 * 8192 + 8192)
    llm = 28 * per_tok * (200 + 117) + 2 * 2 * R_RESP * 3072 * 128256
    if not train:
        # the forward runs the last layer's out-projection / MLP on the 2 * R consumed rows only: credit what is done
        skipped = (200 + 117 - 2 * R_RESP) * 2 * 3072 * (3072 + 2 * 8192 + 8192)
        return enc + llm - skipped
    return 3 * enc + llm + 28 * per_tok * 200 + 2 * R_RESP * 3072 * 128256


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback exists); use --impl reference for the CPU arm")
    from llm_speech_summarization_b200 import _lib, dp, ops
    from llm_speech_summarization_b200.config import KNOWN_LLMS, to_namespace
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    from llm_speech_summarization_b200.model.audio_llama import AudioLlamaForCausalLM
    from llm_speech_summarization_b200.step import AudioPromptStep

    rank, local_rank, world = dp.init_process_group()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    cfg = to_namespace({"model": {"audio_encoder": {"base": "hubert", "type": "facebook/hubert-large-ls960-ft",
                                                    "random_init": True,  # no hub offline: synthetic weights below
                                                    "downsample_method": "pool", "downsample_factor": 4,
                                                    "pooling": {"kernel_size": 8, "stride": 4}},
                                  "llm_type": "meta-llama/Llama-3.2-3B-Instruct", "llm_embedding_channels": 3072}})
    la = KNOWN_LLMS[cfg.model.llm_type]
    enc_sd, llm_sd = synth_weights(dev)
    op_dtype = torch.float16 if args.dtype == "fp16" else torch.bfloat16
    enc = AudioEncoder(cfg, dev)
    enc.operand_dtype = op_dtype
    enc.load_state_dict(enc_sd, strict=True)
    enc.eval().to(dev)
    llm = AudioLlamaForCausalLM(la, dtype=op_dtype)
    llm.load_state_dict(llm_sd, strict=True)
    llm.eval().to(dev)
    tok = FixedTokenizer(la.vocab, la.bos)
    step = AudioPromptStep(enc, llm, tok, cfg.model.llm_type)
    train = args.workload == "train"
    B = args.batch
    trainer = None
    if train:
        from llm_speech_summarization_b200.training import EncoderTrainer
        trainer = EncoderTrainer(step, enc, llm, lr=5e-5, betas=(0.9, 0.999), grad_accum_interval=B * world,
                                 total_optimizer_steps=10 ** 6, regularize=args.regularize != "none",
                                 generator=torch.Generator().manual_seed(1234 + rank))
        if args.regularize == "dropout":
            enc.regularizers.layerdrop = 0.0

    n_pool = 2  # rotate through distinct micro-batches
    host = [synth_batch(B, la.vocab, 1000 * (rank + 1) + i) for i in range(n_pool)]
    host = [(w.pin_memory(), t, r) for (w, t, r) in host]
    resident = [(w.to(dev), t, r) for (w, t, r) in host]
    plans = [step.plan(123, t, r, dev) for (_, t, r) in resident]
    torch.cuda.synchronize()

    ragged_lens = None
    if args.ragged:
        assert train, "--ragged applies to --workload train"
        gl = torch.Generator().manual_seed(4242 + rank)
        ragged_lens = [[int(x) for x in torch.randint(SAMPLES // 2, SAMPLES + 1, (B,), generator=gl)] for _ in range(n_pool)]
        for k in range(n_pool):
            ragged_lens[k][0] = SAMPLES  # the padded length is the longest utterance
            for b_, n_ in enumerate(ragged_lens[k]):
                resident[k][0][b_, n_:] = 0
                host[k][0][b_, n_:] = 0

    def run_resident(i):
        w, t, r = resident[i % n_pool]
        if train and ragged_lens is not None:
            return trainer.train_step(w, t, r, lengths=ragged_lens[i % n_pool])
        if train:
            return trainer.train_step(w, t, r, plan=plans[i % n_pool])
        return step.forward_losses(w, t, r, plan=plans[i % n_pool])

    public = trainer if train else step

    lib = _lib.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    loss_alone = None
    if not train and not args.profile_mode:
        loss_alone = loss_roofline(step, lib, _lib, resident[0], plans[0], dev, peaks)
    # ---- value: inputs resident in HBM
    for i in range(args.warmup):
        run_resident(i)
    torch.cuda.synchronize()
    dp.barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = lib.b2s_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        out = run_resident(i)
    e1.record()
    torch.cuda.synchronize()
    dp.barrier()
    launches = lib.b2s_launch_count() - launches0
    clk = clocks.stop()
    ms = dp.max_over_ranks(e0.elapsed_time(e1), dev)
    total_utts = dp.sum_over_ranks(float(B * args.steps), dev)
    value = total_utts / (ms / 1e3)
    loss_check = float(out["total_loss"].mean())

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"profile_mode": True, "ms_per_step": ms / args.steps, "gpu_launches": int(launches)}))
        return
    # ---- e2e: host buffers -> H2D -> step -> D2H, through the public call
    rag = (lambda i: {"lengths": ragged_lens[i % n_pool]}) if ragged_lens is not None else (lambda i: {})
    # warm-up in the same pipelined pattern as the timed loop (two batches in flight), so the second set of pinned /
    # device staging buffers exists before the clock starts (cudaHostAlloc under 8 processes is slow)
    warm = None
    for i in range(max(3, args.warmup)):
        nxt = public.submit(host[i % n_pool][0], host[i % n_pool][1], host[i % n_pool][2], dev, **rag(i))
        if warm is not None:
            warm.result()
        warm = nxt
    warm.result()
    torch.cuda.synchronize()
    dp.barrier()
    t0 = time.perf_counter()
    pending = None  # the streaming form of the public call: batch i+1 is submitted before batch i's losses are read
    for i in range(args.steps):
        w, t, r = host[i % n_pool]
        nxt = public.submit(w, t, r, dev, **rag(i))
        if pending is not None:
            res = pending.result()
        pending = nxt
    res = pending.result()
    torch.cuda.synchronize()
    e2e_s = dp.max_over_ranks(time.perf_counter() - t0, dev)
    dp.barrier()
    e2e_value = total_utts / e2e_s
    plan0 = plans[0]
    h2d = B * SAMPLES * 4 + 4 * (plan0.row_src.numel() + plan0.cu_seqlens.numel() + plan0.positions.numel() +
                                 plan0.logit_rows.numel() + plan0.labels.numel() + plan0.row_offsets.numel()) \
        + 4 * plan0.audio_rows.numel() + 4 * plan0.seg.numel() + 4 * plan0.resp_len_f.numel()
    d2h = 4 * 4 * B

    # ---- roofline of the dominant kernel (tcgen05 GEMM): per-launch CUDA events over instrumented steps
    lib.b2s_gemm_timing_enable(1)
    for i in range(2):
        run_resident(i)
    torch.cuda.synchronize()
    import ctypes as C
    g_ms, g_n = C.c_double(), C.c_longlong()
    lib.b2s_gemm_timing_read(C.byref(g_ms), C.byref(g_n))
    if args.gemm_shapes and rank == 0:
        write_gemm_shapes(lib, g_n.value, args.gemm_shapes, steps=2)
    lib.b2s_gemm_timing_enable(0)
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md sustained)"
    flops_step = gemm_flops_per_utt(train) * B
    gemm_ms_step = g_ms.value / 2
    achieved = flops_step / (gemm_ms_step / 1e3) / 1e12 if gemm_ms_step > 0 else None
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all launches of one step)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved / peak_tf) if achieved else None, "traffic": None, "peak_source": peak_src,
                "launches_per_step": g_n.value // 2, "gemm_ms_per_step": gemm_ms_step,
                "gemm_share_of_step": gemm_ms_step / (ms / args.steps) if ms > 0 else None,
                "algorithmic_gflop_per_utt": gemm_flops_per_utt(train) / 1e9}
    roofline["traffic"], roofline["traffic_source"] = ncu_traffic_per_launch(train)
    # ---- the HBM-bound kernel the north star names (fused CE + KD loss): again, hot, right after the timed steps
    roofline_loss = None
    if not train and loss_alone is not None:
        hot = loss_roofline(step, lib, _lib, resident[0], plans[0], dev, peaks)
        roofline_loss = {"bound": "hbm", "kernel": "kd_ce_partial_kernel + kd_ce_finalize_kernel (fused CE + KD forward)",
                         "unit": "GB/s", **loss_alone,
                         "timed": "alone, before the step heats the part (SM clock at its maximum): the burst HBM peak "
                                  "applies; `hot` is the same measurement right after the timed steps, under the "
                                  "power cap (the kernel issues 2 MUFU.EX2 per logit pair and is SM-clock sensitive)",
                         "hot": {k: hot[k] for k in ("achieved", "frac", "ms")},
                         "l2": "256 MB buffer rewritten before every timed launch",
                         "traffic": 1.0507e9 * loss_alone["rows"] / 2048.0,
                         "traffic_source": "profiles/r01_ncu_full_captures.txt (dram read 1.0507 GB at 2048 rows)",
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not train:
            enc_cpu = {k: v.float().cpu() for k, v in enc_sd.items()}
            llm_cpu = {k: v.float().cpu() for k, v in llm_sd.items() if k != "lm_head.weight"}
            llm_cpu["lm_head.weight"] = llm_cpu["model.embed_tokens.weight"]
            cpu, _ = cpu_reference(args.cpu_utts, enc_cpu, llm_cpu, steps=1, warmup=0)
        workload = WORKLOAD_TRAIN if train else WORKLOAD
        timed = ("forward with kept activations + backward to all encoder/projector parameters + gradient all-reduce "
                 "+ AdamW, every step" if train else "forward only (encoder + student/teacher prefill + CE/KD/FD)")
        line = {"metric": METRIC_TRAIN if train else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": workload, "utterances_per_step_per_gpu": B, "parallelism": f"dp{world}",
                           "l2": "no flush needed: every step streams ~7 GB of weights and >2 GB of activations "
                                 "(>> 126 MB L2); two distinct micro-batches alternate",
                           "timed": timed, **({"regularize": args.regularize} if train else {}),
                           **({"ragged": "utterance lengths uniform in [5 s, 10 s], mean %.2f s" % (
                               sum(map(sum, ragged_lens)) / (len(ragged_lens) * B) / 16000.0)} if ragged_lens else {})},
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "api": "submit()/result(), one batch in flight ahead of the one being read"},
                "gpu_launches": int(launches), "roofline": roofline, "roofline_loss": roofline_loss,
                "cpu_baseline": cpu,
                "check": {"mean_total_loss": loss_check}}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
