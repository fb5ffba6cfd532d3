"""Tile-configuration sweep for the residual-epilogue GEMM shapes (GPU box)."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from llm_speech_summarization_b200 import ops  # noqa: E402
from tools.bench_kernels import time_fn  # noqa: E402

dev = torch.device("cuda")
for (M, N, K, label) in [(15968, 1024, 1024, "enc_o"), (15968, 1024, 4096, "enc_ffn2"), (10144, 3072, 3072, "llm_o"),
                         (10144, 3072, 8192, "llm_down")]:
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    h = torch.zeros(M, N, device=dev)
    for bn, cg in ((256, 2), (128, 2), (256, 1), (128, 1)):
        ms = time_fn(lambda: ops.gemm(a, w, epi=ops.EPI_RESID_F32, bias=b, resid=h, out=h, block_n=bn, cta_group=cg),
                     iters=20)
        print(json.dumps({"label": label, "bn": bn, "cg": cg, "ms": round(ms, 4),
                          "tflops": round(2 * M * N * K / ms / 1e9, 1)}), flush=True)
