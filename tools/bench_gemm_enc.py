"""GEMM micro-benchmark of the shapes of one step (GPU box): default tile configuration only, plus GELU variants."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from llm_speech_summarization_b200 import ops  # noqa: E402
from tools.bench_kernels import time_fn  # noqa: E402

dev = torch.device("cuda")


def run(M, N, K, epi, label, act=0, bias=False):
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=dev) if bias else None
    if epi == ops.EPI_RESID_F32:
        h = torch.zeros(M, N, device=dev)
        kw = dict(resid=h, out=h)
    elif epi == ops.EPI_SWIGLU:
        kw = dict(out=torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16))
    else:
        kw = dict(out=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
    ms = time_fn(lambda: ops.gemm(a, w, epi=epi, act=act, bias=b, **kw), iters=20)
    print(json.dumps({"label": label, "M": M, "N": N, "K": K, "ms": round(ms, 4),
                      "tflops": round(2 * M * N * K / ms / 1e9, 1)}), flush=True)


R = 32 * 499
run(R, 3072, 1024, ops.EPI_BF16, "enc_qkv", bias=True)
run(R, 1024, 1024, ops.EPI_RESID_F32, "enc_o", bias=True)
run(R, 4096, 1024, ops.EPI_BF16, "enc_ffn1_gelu", act=ops.ACT_GELU, bias=True)
run(R, 4096, 1024, ops.EPI_BF16, "enc_ffn1_nogelu", bias=True)
run(R, 1024, 4096, ops.EPI_RESID_F32, "enc_ffn2", bias=True)
M = 32 * 317
run(M, 5120, 3072, ops.EPI_BF16, "llm_qkv")
run(M, 3072, 3072, ops.EPI_RESID_F32, "llm_o")
run(M, 16384, 3072, ops.EPI_SWIGLU, "llm_gu")
run(M, 3072, 8192, ops.EPI_RESID_F32, "llm_down")
run(4096, 128256, 3072, ops.EPI_BF16, "lm_head")
