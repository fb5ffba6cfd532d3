"""Kernel micro-benchmarks (GPU box only): tcgen05 GEMM TFLOP/s per shape/config and fused-loss GB/s.

CUDA-event timed on the launching stream, >=3 warm-ups, L2 flushed between timed iterations by rotating through
operand sets larger than L2 where sizes allow. Prints one JSON line per measurement.
"""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from llm_speech_summarization_b200 import ops  # noqa: E402


def time_fn(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def bench_gemm(M, N, K, epi, bn, cg, label):
    dev = torch.device("cuda")
    a = (torch.randn(M, K, device=dev) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.05).to(torch.bfloat16)
    kw = {}
    if epi == ops.EPI_RESID_F32:
        h = torch.zeros(M, N, device=dev)
        kw = dict(resid=h, out=h)
    elif epi == ops.EPI_SWIGLU:
        kw = dict(out=torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16))
    else:
        kw = dict(out=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
    ms = time_fn(lambda: ops.gemm(a, w, epi=epi, block_n=bn, cta_group=cg, **kw))
    ref = time_fn(lambda: torch.matmul(a, w.t()))
    print(json.dumps({"kernel": "gemm", "label": label, "M": M, "N": N, "K": K, "epi": epi, "bn": bn, "cg": cg,
                      "ms": round(ms, 4), "tflops": round(2 * M * N * K / ms / 1e9, 1),
                      "cublas_ms": round(ref, 4), "cublas_tflops": round(2 * M * N * K / ref / 1e9, 1)}), flush=True)


def bench_loss(rows, V):
    dev = torch.device("cuda")
    s = torch.randn(rows, V, device=dev).to(torch.bfloat16)
    t = torch.randn(rows, V, device=dev).to(torch.bfloat16)
    labels = torch.randint(0, V, (rows,), device=dev, dtype=torch.int32)
    offs = torch.arange(0, rows + 1, 64, device=dev, dtype=torch.int32)
    res = ops.kd_ce_loss(s, t, labels, offs)
    # time the two launches alone (pre-allocated outputs, direct C-ABI call: no allocator / wrapper overhead)
    from llm_speech_summarization_b200 import _lib
    lib = _lib.load()
    ws = torch.empty(lib.b2s_kd_ce_workspace_bytes(rows, V), device=dev, dtype=torch.uint8)
    U = offs.numel() - 1
    ld_o, ntp_o = torch.empty(U, device=dev), torch.empty(U, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def fwd():
        _lib.check(lib.b2s_kd_ce_loss_fwd(s.data_ptr(), t.data_ptr(), V, V, rows, V, labels.data_ptr(), offs.data_ptr(),
                                          U, 1.0, 1.0, ws.data_ptr(), res.lse_s.data_ptr(), res.lse_t.data_ptr(),
                                          res.coef_kd.data_ptr(), res.coef_ce.data_ptr(), ld_o.data_ptr(),
                                          ntp_o.data_ptr(), st))
    ms = time_fn(fwd, iters=20)
    ds = torch.empty_like(s)
    msb = time_fn(lambda: ops.kd_ce_loss_bwd(s, t, labels, res, out=ds))
    print(json.dumps({"kernel": "kd_ce_loss", "rows": rows, "V": V, "fwd_ms": round(ms, 4),
                      "fwd_gbs": round(4 * rows * V / ms / 1e6, 1), "bwd_ms": round(msb, 4),
                      "bwd_gbs": round(6 * rows * V / msb / 1e6, 1)}), flush=True)


def bench_attn(label, lens, Hq, Hkv, D, causal):
    import math
    dev = torch.device("cuda")
    rows = sum(lens)
    qkv = torch.randn(rows, (Hq + 2 * Hkv) * D, device=dev).to(torch.bfloat16)
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    flops = sum(4 * L * L * Hq * D * (0.5 if causal else 1.0) for L in lens)
    from llm_speech_summarization_b200 import _lib
    lib = _lib.load()
    for cfg, cname in (((0, 0), "default"), ((64, 3), "pipelined")):
        if cfg[1] == 3 and max(lens) > 1024 and D == 128:
            continue
        lib.b2s_set_option(_lib.OPT_ATTN_KEYS_PER_STEP, cfg[0])
        lib.b2s_set_option(_lib.OPT_ATTN_KV_STAGES, cfg[1])
        for dt, name in ((torch.float16, "fp16"), (torch.bfloat16, "bf16")):
            x = qkv.to(dt)
            ms = time_fn(lambda: ops.attention(x, cu, max(lens), Hq, Hkv, D, 1.0 / math.sqrt(D), causal), iters=20)
            print(json.dumps({"kernel": "attention", "label": label, "config": cname, "operands": name,
                              "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}), flush=True)
    lib.b2s_set_option(_lib.OPT_ATTN_KEYS_PER_STEP, 0)
    lib.b2s_set_option(_lib.OPT_ATTN_KV_STAGES, 0)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "attn"):
        bench_attn("hubert_b32", [499] * 32, 16, 16, 64, False)
        bench_attn("whisper_b8", [1500] * 8, 16, 16, 64, False)
        bench_attn("llama_b32_student+teacher", [200] * 32 + [117] * 32, 24, 8, 128, True)
        bench_attn("minichat_b8_L400", [400] * 8, 24, 24, 128, True)
        bench_attn("llama_b1_prompt137 (configs[1] inference prefill)", [137], 24, 8, 128, True)
        bench_attn("hubert_b1", [499], 16, 16, 64, False)
        bench_attn("llama_b4_prompt137", [137] * 4, 24, 8, 128, True)
    if which in ("all", "loss"):  # before the GEMMs heat the part up (the loss kernel is SM-clock sensitive)
        for rows in (64, 2048):
            bench_loss(rows, 128256)
    if which in ("all", "gemm"):
        shapes = [
            (15968, 3072, 1024, ops.EPI_BF16, "hubert_qkv_b32"),
            (15968, 1024, 1024, ops.EPI_RESID_F32, "hubert_out_b32"),
            (15968, 4096, 1024, ops.EPI_BF16, "hubert_ffn1_b32"),
            (15968, 1024, 4096, ops.EPI_RESID_F32, "hubert_ffn2_b32"),
            (10144, 5120, 3072, ops.EPI_BF16, "llama_qkv_b32"),
            (10144, 3072, 3072, ops.EPI_RESID_F32, "llama_o_b32"),
            (10144, 16384, 3072, ops.EPI_SWIGLU, "llama_gateup_b32"),
            (10144, 3072, 8192, ops.EPI_RESID_F32, "llama_down_b32"),
            (4096, 128256, 3072, ops.EPI_BF16, "lm_head_b32"),
        ]
        for (M, N, K, epi, label) in shapes:
            for bn, cg in ((256, 1), (128, 1), (256, 2), (128, 2)):
                try:
                    bench_gemm(M, N, K, epi, bn, cg, label)
                except Exception as e:  # keep going: one bad config must not hide the others
                    print(json.dumps({"kernel": "gemm", "label": label, "bn": bn, "cg": cg, "error": str(e)[:300]}),
                          flush=True)
    if which in ("all",):  # and again hot, right after sustained tensor-core load (power-capped clocks)
        for rows in (2048,):
            bench_loss(rows, 128256)
