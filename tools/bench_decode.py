"""Decode-step micro-benchmark (GPU box only): Llama-3.2-3B random-init, prompt L=137 (configs[1] prefill length),
greedy decode with the KV cache. Prints ms/token (CUDA-event timed over 64 steps after 8 warm-up steps) and the HBM
fraction: a decode step must stream every weight once (algorithmic bytes = bf16 weights of 28 layers + LM head)."""
import json
import os
import sys

import torch

ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from llm_speech_summarization_b200 import _lib  # noqa: E402
from llm_speech_summarization_b200.config import KNOWN_LLMS  # noqa: E402
from llm_speech_summarization_b200.model.audio_llama import AudioLlamaForCausalLM  # noqa: E402


def main():
    dev = torch.device("cuda")
    la = KNOWN_LLMS["meta-llama/Llama-3.2-3B-Instruct"]
    _, llm_sd = bench.synth_weights(dev)
    llm = AudioLlamaForCausalLM(la)
    llm.load_state_dict(llm_sd, strict=True)
    llm.eval().to(dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6461.5))
    D = la.head_dim
    per_layer = la.hidden * (la.heads + 2 * la.kv_heads) * D + la.heads * D * la.hidden + 3 * la.hidden * la.ffn
    wbytes = 2 * (la.layers * per_layer + la.vocab * la.hidden)
    for B in [int(x) for x in os.environ.get("DECODE_BATCHES", "1,2,4,8,32").split(",")]:
        g = torch.Generator(device=dev).manual_seed(1)
        prompts = [torch.randn(137, la.hidden, device=dev, generator=g) * 0.02 for _ in range(B)]
        steps, warm = 64, 8
        logits, state = llm.prefill_with_cache(prompts, steps + warm + 1)
        tok = logits.float().argmax(dim=-1).to(torch.int32)
        lib = _lib.load()
        n0 = None
        for i in range(warm + steps):
            if i == warm:
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n0 = lib.b2s_launch_count()
                e0.record()
            logits = llm.decode_step(tok, state)
            tok = logits.float().argmax(dim=-1).to(torch.int32)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"kernel": "decode_step", "batch": B, "ms_per_step": round(ms, 4),
                          "tokens_per_s": round(B * 1e3 / ms, 1), "weight_GB": round(wbytes / 1e9, 3),
                          "achieved_GBps": round(wbytes / ms / 1e6, 1), "hbm_frac": round(wbytes / ms / 1e6 / hbm, 3),
                          "launches_per_step": (lib.b2s_launch_count() - n0) // steps,
                          "path": "megakernel" if (B <= 4 and os.environ.get("B2S_DECODE_MEGA", "1") != "0") else "multi-kernel"}),
              flush=True)


if __name__ == "__main__":
    main()
