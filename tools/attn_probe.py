"""One attention forward launch per configuration on the HuBERT bench shape (for ncu captures): ATTN_CFG="64,3" etc."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_speech_summarization_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")
shape = os.environ.get("ATTN_SHAPE", "hubert")
if shape == "hubert":
    lens, Hq, Hkv, D, causal = [499] * 32, 16, 16, 64, False
else:
    lens, Hq, Hkv, D, causal = [200] * 32 + [117] * 32, 24, 8, 128, True
bn, kvs = (int(v) for v in os.environ.get("ATTN_CFG", "0,0").split(","))
lib.b2s_set_option(_lib.OPT_ATTN_KEYS_PER_STEP, bn)
lib.b2s_set_option(_lib.OPT_ATTN_KV_STAGES, kvs)
qkv = torch.randn(sum(lens), (Hq + 2 * Hkv) * D, device=dev).half()
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
for _ in range(3):
    o = ops.attention(qkv, cu, max(lens), Hq, Hkv, D, 1.0 / math.sqrt(D), causal)
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
