import sys, json, math, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_speech_summarization_b200 import ops
from tools.bench_kernels import time_fn
dev = torch.device("cuda")
# ramp the clocks: ~0.5 s of dense work before any timing
_a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(60):
    _a @ _a
torch.cuda.synchronize()
def run(label, lens, Hq, Hkv, D, causal):
    rows = sum(lens)
    qkv = torch.randn(rows, (Hq + 2 * Hkv) * D, device=dev).to(torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    for impl in (0, 1, 0, 1):
        ops.attention_set_impl(impl)
        ms = time_fn(lambda: ops.attention(qkv, cu, max(lens), Hq, Hkv, D, 1.0 / math.sqrt(D), causal), iters=10)
        print(label, "impl", impl, "ms", round(ms, 4), flush=True)
run("A 512x128 D64", [128] * 512, 16, 16, 64, False)
run("B 128x512 D64", [512] * 128, 16, 16, 64, False)
run("C 32x2048 D64", [2048] * 32, 16, 16, 64, False)
run("D 128x512 D128", [512] * 128, 8, 8, 128, False)
run("E 32x2048 D128", [2048] * 32, 8, 8, 128, False)
