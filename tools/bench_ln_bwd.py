"""Times b2s_layernorm_bwd_ex alone on the two shapes of the training step (L2 flushed between launches, CUDA events):
transformer LN (15968 x 1024, x fp32, dy fp16, dh accumulated, fp16 copy, fused column sum) and the conv front end's
LN+GELU (conv layer 1: 32 x 15999 rows x 512, x / dy fp16, fp16 dx only). Prints us and the GB/s of the algorithmic bytes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_speech_summarization_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, n=7):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[n // 2]


rows, C = 15968, 1024
x = torch.randn(rows, C, device=dev, generator=g)
dy = torch.randn(rows, C, device=dev, generator=g).half()
gm, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
dh = torch.zeros(rows, C, device=dev)
dg, db, cs = (torch.zeros(C, device=dev) for _ in range(3))
for fused in (False, True):
    t = timed(lambda: ops.layernorm_bwd_ex(x, gm, bt, dy, 1e-5, dh=dh, accumulate=True, dx_dtype=torch.float16, dgamma=dg,
                                           dbeta=db, dh_colsum=cs if fused else None))
    by = rows * C * (4 + 2 + 4 + 4 + 2)
    print(f"transformer LN bwd {rows} x {C} (colsum fused: {fused}): {t * 1e3:7.1f} us  {by / t / 1e6:7.1f} GB/s")
rows, C = 32 * 15999, 512
x = torch.randn(rows, C, device=dev, generator=g).half()
dy = torch.randn(rows, C, device=dev, generator=g).half()
gm, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
dg, db = (torch.zeros(C, device=dev) for _ in range(2))
t = timed(lambda: ops.layernorm_bwd_ex(x, gm, bt, dy, 1e-5, gelu=True, dx_dtype=torch.float16, dgamma=dg, dbeta=db))
print(f"conv LN+GELU bwd {rows} x {C}: {t * 1e3:7.1f} us  {rows * C * 6 / t / 1e6:7.1f} GB/s")
