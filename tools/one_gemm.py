"""Launch one GEMM shape a few times (for an ncu capture): python tools/one_gemm.py enc_o|enc_ffn1"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from llm_speech_summarization_b200 import ops  # noqa: E402

dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "enc_o"
M = 15968
if which == "enc_o":
    N, K = 1024, 1024
else:
    N, K = 4096, 1024
a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
b = torch.randn(N, device=dev)
for _ in range(4):
    if which == "enc_o":
        h = torch.zeros(M, N, device=dev)
        ops.gemm(a, w, epi=ops.EPI_RESID_F32, bias=b, resid=h, out=h)
    else:
        ops.gemm(a, w, epi=ops.EPI_BF16, act=ops.ACT_GELU, bias=b)
torch.cuda.synchronize()
