"""Backward-pass kernel micro-benchmarks (GPU box only): dgrad / wgrad GEMM TFLOP/s per shape, attention backward,
conv0 backward, LayerNorm backward. CUDA-event timed, median of 10 after 3 warm-ups. One JSON line each."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from llm_speech_summarization_b200 import _lib, ops  # noqa: E402
from tools.bench_kernels import time_fn  # noqa: E402

dev = torch.device("cuda")
bf = torch.bfloat16


def rnd(*s, std=0.5):
    return (torch.randn(*s, device=dev) * std).to(bf)


def bench_dgrad(M, N, K, label, **kw):
    dy, w = rnd(M, N), rnd(N, K, std=0.05)
    ms = time_fn(lambda: ops.gemm_dgrad(dy, w, **kw))
    print(json.dumps({"kernel": "dgrad", "label": label, "M": M, "N": N, "K": K, **kw, "ms": round(ms, 4),
                      "tflops": round(2 * M * N * K / ms / 1e9, 1)}), flush=True)


def bench_wgrad(rows, N, K, label, **kw):
    dy, x = rnd(rows, N), rnd(rows, K)
    out = torch.zeros(N, K, device=dev)
    ms = time_fn(lambda: ops.gemm_wgrad(dy, x, out, **kw))
    print(json.dumps({"kernel": "wgrad", "label": label, "rows": rows, "N": N, "K": K, **kw, "ms": round(ms, 4),
                      "tflops": round(2 * rows * N * K / ms / 1e9, 1)}), flush=True)


def bench_attn_bwd(label, lens, Hq, Hkv, D, causal):
    rows = sum(lens)
    qkv = rnd(rows, (Hq + 2 * Hkv) * D)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    o, lse = ops.attention(qkv, cu, max(lens), Hq, Hkv, D, D ** -0.5, causal, return_lse=True)
    do = rnd(rows, Hq * D)
    ms = time_fn(lambda: ops.attention_bwd(qkv, o, do, lse, cu, max(lens), Hq, Hkv, D, D ** -0.5, causal))
    fl = sum(4 * L * L * D * Hq * (0.5 if causal else 1.0) for L in lens) * 2.5
    print(json.dumps({"kernel": "attn_bwd", "label": label, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}),
          flush=True)


def bench_conv0_bwd(B, samples):
    lib = _lib.load()
    T1 = (samples - 10) // 5 + 1
    wave = torch.randn(B, samples, device=dev) * 0.1
    w, b = torch.randn(512, 10, device=dev) * 0.3, torch.randn(512, device=dev) * 0.05
    g, bt = torch.ones(512, device=dev), torch.zeros(512, device=dev)
    dy = rnd(B, T1, 512)
    dW, db, dg, dbt = (torch.zeros(512, 10, device=dev), torch.zeros(512, device=dev), torch.zeros(512, device=dev),
                       torch.zeros(512, device=dev))
    st = torch.cuda.current_stream().cuda_stream
    ms = time_fn(lambda: _lib.check(lib.b2s_conv0_bwd(wave.data_ptr(), samples, B, samples, w.data_ptr(), b.data_ptr(),
                                                      g.data_ptr(), bt.data_ptr(), 1e-5, dy.data_ptr(), T1,
                                                      dW.data_ptr(), db.data_ptr(), dg.data_ptr(), dbt.data_ptr(), st),
                                    "conv0_bwd"))
    print(json.dumps({"kernel": "conv0_bwd", "B": B, "ms": round(ms, 4),
                      "GBps": round(dy.numel() * 2 / ms / 1e6, 1)}), flush=True)


if __name__ == "__main__":  # noqa
    R = 32 * 499
    if len(sys.argv) > 1 and sys.argv[1] == "splits":
        for (N, K, lab) in [(1024, 4096, "enc w2 wgrad"), (4096, 1024, "enc w1 wgrad"), (1024, 1024, "enc wo wgrad"),
                            (3072, 1024, "enc wqkv wgrad"), (3072, 1024, "proj wgrad rows=3936"),
                            (512, 1536, "conv1 wgrad rows=511968"), (512, 1536, "conv4 wgrad rows=63968"),
                            (1024, 512, "featproj wgrad")]:
            rows = 3936 if "proj wgrad" in lab else (int(lab.split("=")[1]) if "=" in lab else R)
            for ks in (1, 2, 3, 4, 6, 8, 9, 12, 16):
                bench_wgrad(rows, N, K, lab, k_splits=ks)
        sys.exit(0)
    bench_conv0_bwd(32, 160000)
    for (N, K, lab) in [(1024, 4096, "enc w2 dgrad"), (4096, 1024, "enc w1 dgrad"), (1024, 1024, "enc wo dgrad"),
                        (3072, 1024, "enc wqkv dgrad")]:
        bench_dgrad(R, N, K, lab)
    bench_dgrad(32 * 15999, 512, 1536, "conv1 dgrad")
    for (N, K, lab) in [(1024, 4096, "enc w2 wgrad"), (4096, 1024, "enc w1 wgrad"), (1024, 1024, "enc wo wgrad"),
                        (3072, 1024, "enc wqkv wgrad"), (3072, 1024, "proj wgrad rows=3936")]:
        rows = 3936 if "proj" in lab else R
        for kw in ({}, {"k_splits": 1}, {"block_n": 128, "cta_group": 1}, {"block_n": 256, "cta_group": 1}):
            bench_wgrad(rows, N, K, lab, **kw)
    bench_attn_bwd("hubert 32x499 H16 D64", [499] * 32, 16, 16, 64, False)
    bench_attn_bwd("llama 32x200 H24/8 D128 causal", [200] * 32, 24, 8, 128, True)
