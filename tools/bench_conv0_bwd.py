"""Times b2s_conv0_bwd alone (CUDA events, L2 flushed between launches) for both 16-bit formats and for finite / inf dy.
Usage: python tools/bench_conv0_bwd.py   (B = 32 utterances of 10 s, HuBERT-large conv0: 512 channels, k = 10, s = 5)"""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_speech_summarization_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
B, S, C, K, ST = 32, 160000, 512, 10, 5
frames = (S - K) // ST + 1
g = torch.Generator(device=dev).manual_seed(0)
wave = torch.randn(B, S, device=dev, generator=g)
w = torch.randn(C, K, device=dev, generator=g) * 0.3
bias = torch.randn(C, device=dev, generator=g) * 0.1
gamma = 1 + 0.1 * torch.randn(C, device=dev, generator=g)
beta = 0.1 * torch.randn(C, device=dev, generator=g)
dW, db, dg, dbt = (torch.zeros(n, device=dev) for n in (C * K, C, C, C))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(dy, fmt, n=5):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.b2s_conv0_bwd(p(wave), S, B, S, p(w), p(bias), p(gamma), p(beta), 1e-5, p(dy), frames, p(dW), p(db),
                               p(dg), p(dbt), fmt, st)
        assert rc == 0
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2]


for name, dt, fmt in (("bf16", torch.bfloat16, 0), ("fp16", torch.float16, 1)):
    base = torch.randn(B * frames, C, device=dev, generator=g)
    for label, scale in (("1e-3", 1e-3), ("1", 1.0), ("6e4 (some inf in fp16)", 3e4)):
        dy = (base * scale).to(dt)
        print(f"conv0_bwd {name} dy x {label}: {run(dy, fmt):.3f} ms   nonfinite {int((~torch.isfinite(dy)).sum())}")
    dy = torch.full((B * frames, C), float("nan"), device=dev, dtype=dt)
    print(f"conv0_bwd {name} dy all-NaN: {run(dy, fmt):.3f} ms")
    dy = torch.zeros(B * frames, C, device=dev, dtype=dt)
    print(f"conv0_bwd {name} dy zeros: {run(dy, fmt):.3f} ms")
    dy = (base * 1e-7).to(dt)
    print(f"conv0_bwd {name} dy x 1e-7 (fp16 subnormal): {run(dy, fmt):.3f} ms")
