import os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llm_speech_summarization_b200 import ops
dev = torch.device("cuda")
lens = [512] * 128
Hq = Hkv = 16; D = 64
qkv = torch.randn(sum(lens), (Hq + 2 * Hkv) * D, device=dev).to(torch.bfloat16)
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
ops.attention_set_impl(1)
for _ in range(3):
    ops.attention(qkv, cu, max(lens), Hq, Hkv, D, 1.0 / math.sqrt(D), False)
torch.cuda.synchronize()
