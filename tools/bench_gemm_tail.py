"""A/B of the GEMM tail split (B2S_OPT_GEMM_TAIL_SPLIT) on the shapes of the path whose tile count leaves a partly filled
last round on 148 SMs. CUDA events over 20 back-to-back launches per shape (hot, like the step), fp16 operands."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_speech_summarization_b200 import _lib, ops

lib = _lib.load()
dev = torch.device("cuda:0")
SHAPES = [  # M, N, K, epilogue
    (6400, 3072, 16384, "f32"), (6400, 3072, 5120, "f32"), (10144, 3072, 8192, "resid"), (10144, 3072, 3072, "resid"),
    (15968, 1024, 4096, "resid"), (15968, 1024, 1024, "resid"), (4096, 3072, 8192, "resid"), (4384, 3072, 8192, "resid"),
]
g = torch.Generator(device=dev).manual_seed(0)
for M, N, K, epi in SHAPES:
    a = (torch.randn(M, K, device=dev, generator=g) * 0.5).half()
    w = (torch.randn(N, K, device=dev, generator=g) * 0.05).half()
    out = torch.zeros(M, N, device=dev)
    res = {}
    for split in (0, 1, 0, 1):
        _lib.check(lib.b2s_set_option(_lib.OPT_GEMM_TAIL_SPLIT, split), "set_option")
        kw = dict(epi=ops.EPI_F32, out=out) if epi == "f32" else dict(epi=ops.EPI_RESID_F32, resid=out, out=out)
        for _ in range(5):
            ops.gemm(a, w, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm(a, w, **kw)
        e1.record()
        torch.cuda.synchronize()
        res.setdefault(split, []).append(e0.elapsed_time(e1) / 20)
    t0, t1 = min(res[0]), min(res[1])
    tf = lambda t: 2.0 * M * N * K / t / 1e9
    print(f"{M:6d} {N:5d} {K:6d} {epi:6s}  unsplit {t0 * 1e3:8.1f} us {tf(t0):7.1f} TF/s   tail split {t1 * 1e3:8.1f} us "
          f"{tf(t1):7.1f} TF/s   {100 * (t0 / t1 - 1):+5.1f} %")
