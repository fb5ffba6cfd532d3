"""ncu target: one attention backward per configuration (tcgen05 and mma.sync), for per-kernel durations."""
import sys
import torch
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from llm_speech_summarization_b200 import ops

dev = torch.device("cuda")
for (lens, Hq, Hkv, D, causal) in (([499] * 32, 16, 16, 64, False), ([200] * 32, 24, 8, 128, True)):
    rows = sum(lens)
    qkv = (torch.randn(rows, (Hq + 2 * Hkv) * D, device=dev) * 0.5).to(torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    o, lse = ops.attention(qkv, cu, max(lens), Hq, Hkv, D, D ** -0.5, causal, return_lse=True)
    do = (torch.randn(rows, Hq * D, device=dev) * 0.5).to(torch.bfloat16)
    for impl in (1, 0, 1, 0):
        ops.attention_set_impl(impl)
        ops.attention_bwd(qkv, o, do, lse, cu, max(lens), Hq, Hkv, D, D ** -0.5, causal)
    ops.attention_set_impl(1)
torch.cuda.synchronize()
