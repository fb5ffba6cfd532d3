"""How long does the HOST take to enqueue one training step (no sync)? If it approaches the GPU time the step is
launch-bound somewhere. Prints enqueue ms and device ms per step, plus per-phase enqueue times."""
import sys, time, json
import torch
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
import bench
from llm_speech_summarization_b200.config import KNOWN_LLMS, to_namespace
from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
from llm_speech_summarization_b200.model.audio_llama import AudioLlamaForCausalLM
from llm_speech_summarization_b200.step import AudioPromptStep
from llm_speech_summarization_b200.training import EncoderTrainer

dev = torch.device("cuda")
cfg = to_namespace({"model": {"audio_encoder": {"base": "hubert", "type": "facebook/hubert-large-ls960-ft",
                                                "downsample_method": "pool", "downsample_factor": 4,
                                                "pooling": {"kernel_size": 8, "stride": 4}},
                              "llm_type": "meta-llama/Llama-3.2-3B-Instruct", "llm_embedding_channels": 3072}})
la = KNOWN_LLMS[cfg.model.llm_type]
enc_sd, llm_sd = bench.synth_weights(dev)
enc = AudioEncoder(cfg, dev); enc.load_state_dict(enc_sd, strict=True); enc.eval().to(dev)
llm = AudioLlamaForCausalLM(la); llm.load_state_dict(llm_sd, strict=True); llm.eval().to(dev)
tok = bench.FixedTokenizer(la.vocab, la.bos)
step = AudioPromptStep(enc, llm, tok, cfg.model.llm_type)
B = 32
tr = EncoderTrainer(step, enc, llm, lr=5e-5, grad_accum_interval=B, total_optimizer_steps=10 ** 6)
w, t, r = bench.synth_batch(B, la.vocab, 1)
w = w.to(dev)
plan = step.plan(123, t, r, dev)
for _ in range(3):
    tr.train_step(w, t, r, plan=plan)
torch.cuda.synchronize()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    ta = time.perf_counter(); audio = enc.forward_train(w); tb = time.perf_counter()
    out = step.llm_forward_backward(audio, t, r, loss_scale=1.0 / B, plan=plan); tc = time.perf_counter()
    enc.backward(out["d_audio_embeds"]); td = time.perf_counter()
    enc.flush_grads(); tr.optimizer.all_reduce_grads(); tr.optimizer.step(); enc.mark_weights_changed(); tr.optimizer.zero_grad()
    te = time.perf_counter(); e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(json.dumps({"enqueue_ms": round((t1 - t0) * 1e3, 2), "device_ms": round(e0.elapsed_time(e1), 2),
                      "wall_ms": round((t2 - t0) * 1e3, 2), "enc_fwd": round((tb - ta) * 1e3, 2),
                      "llm_fwd_bwd": round((tc - tb) * 1e3, 2), "enc_bwd": round((td - tc) * 1e3, 2),
                      "opt": round((te - td) * 1e3, 2)}))
