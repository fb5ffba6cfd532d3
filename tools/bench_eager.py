"""Library-kernel bar: the reference path restated in plain PyTorch (oracle/reference_math.py),
run ON THE SAME B200 under bf16 autocast with cuBLAS / cuDNN / SDPA kernels, batch 1 like the reference
(REF/trainer.py:270-374: encoder forward, student + teacher LLM forward with all-row logits, CE / KD / FD losses).
Forward only, and forward + backward + AdamW (torch autograd, torch.optim.AdamW) for the training step.
CUDA-event timed, prints one JSON line per workload. Measurement aid only -- not a product path."""
import json
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from oracle import configs, reference_math as rm  # noqa: E402

dev = torch.device("cuda")
enc_cfg, llm_cfg = configs.HUBERT_LARGE, configs.LLAMA32_3B
enc_sd = {k: v.to(dev) for k, v in configs.make_encoder_state_dict(enc_cfg, seed=1234).items()}
llm_sd = {k: v.to(dev) for k, v in configs.make_llm_state_dict(llm_cfg, seed=4321, dtype=torch.bfloat16).items()}
llm_sd["lm_head.weight"] = llm_sd["model.embed_tokens.weight"]
tok = configs.stub_tokenizer(llm_cfg)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
utts = [configs.synthetic_utterance(llm_cfg, i, 160000, T=40, R=64) for i in range(n)]
utts = [(a.to(dev), torch.as_tensor(t).to(dev), torch.as_tensor(r).to(dev)) for a, t, r in utts]
torch.set_default_device(dev)  # prompt ids, masks and rope tables the oracle creates on the fly land on the GPU


def fwd(sd):
    tot = 0.0
    for a, t, r in utts:
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            out = rm.train_step_losses(sd, llm_sd, enc_cfg, llm_cfg, tok, a, t, r)
        tot = tot + out["total_loss"] / 16
    return tot


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


with torch.no_grad():
    s = timed(lambda: fwd(enc_sd))
print(json.dumps({"impl": "pytorch-eager bf16 autocast (oracle on the GPU), batch 1", "workload": "forward",
                  "utterances_per_s": n / s, "s_per_pass": s, "utterances": n}), flush=True)

params = {k: v.clone().requires_grad_(True) for k, v in enc_sd.items() if v.is_floating_point()}
opt = torch.optim.AdamW(list(params.values()), lr=5e-5, fused=True)


def train():
    opt.zero_grad(set_to_none=True)
    for a, t, r in utts:
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            out = rm.train_step_losses(params, llm_sd, enc_cfg, llm_cfg, tok, a, t, r)
        (out["total_loss"] / 16).backward()
    opt.step()


s = timed(train, reps=2)
print(json.dumps({"impl": "pytorch-eager bf16 autocast + autograd + fused AdamW, batch 1", "workload": "train",
                  "utterances_per_s": n / s, "s_per_pass": s, "utterances": n}), flush=True)
