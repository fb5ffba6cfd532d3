"""Tensor-level wrappers over the C ABI: torch supplies device memory and the current stream, nothing else.

Every function requires CUDA tensors and raises otherwise (there is no CPU path in this package; the CPU
restatement lives in oracle/ and is test infrastructure only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs

EPI_BF16, EPI_RESID_F32, EPI_SWIGLU, EPI_ROPE, EPI_F32, EPI_ACCUM_F32 = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_GELU = 0, 1
FMT_BF16, FMT_F16 = 0, 1
_H16 = (torch.bfloat16, torch.float16)


def fmt_of(dtype: torch.dtype) -> int:
    """B2S_FMT_* id of a 16-bit torch dtype (include/b2s.h)."""
    if dtype == torch.float16:
        return FMT_F16
    if dtype == torch.bfloat16:
        return FMT_BF16
    raise TypeError(f"16-bit operand format must be torch.bfloat16 or torch.float16, got {dtype}")


def dtype_of(fmt: int) -> torch.dtype:
    return torch.float16 if fmt == FMT_F16 else torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("llm_speech_summarization_b200 ops need CUDA tensors (no CPU fallback exists)")


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, epi: int = EPI_BF16,
         act: int = ACT_NONE, resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         rope_cs: Optional[torch.Tensor] = None, positions: Optional[torch.Tensor] = None, rope_cols: int = 0,
         block_n: int = 0, cta_group: int = 0, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """out = epi(a @ w.T) for row-major a [M, K] and w [N, K], each bf16 or fp16 (they may differ); 16-bit outputs
    take `out_dtype` (default: a's dtype)."""
    _need_cuda(a, w, bias, resid, out)
    assert a.dtype in _H16 and w.dtype in _H16 and a.is_contiguous() and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if epi == EPI_SWIGLU else N
    if out is None:
        dt = torch.float32 if epi in (EPI_RESID_F32, EPI_F32) else (out_dtype or a.dtype)
        out = torch.empty(M, n_out, device=a.device, dtype=dt)
    g = GemmArgs()
    g.a_fmt, g.w_fmt = fmt_of(a.dtype), fmt_of(w.dtype)
    g.out_fmt = fmt_of(out.dtype) if out.dtype in _H16 else FMT_BF16
    g.A, g.a_dim0, g.a_row_stride, g.a_batch_stride, g.a_rows = a.data_ptr(), K, K, 0, M
    g.W, g.w_rows, g.w_cols = w.data_ptr(), N, K
    g.M, g.N, g.batches, g.groups, g.taps, g.k_per_tap = M, N, 1, 1, 1, K
    g.epi, g.act = epi, act
    g.bias = _ptr(bias)
    g.out, g.ldo, g.out_batch_rows = out.data_ptr(), out.stride(0), 0
    g.resid = _ptr(resid)
    g.rope_cs, g.positions, g.rope_cols = _ptr(rope_cs), _ptr(positions), rope_cols
    g.block_n, g.cta_group = block_n, cta_group
    _lib.check(_lib.load().b2s_gemm_bf16(C.byref(g), _stream()), "gemm")
    return out


def gemm_dgrad(dy: torch.Tensor, w: torch.Tensor, *, out_f32: bool = False, block_n: int = 0,
               cta_group: int = 0) -> torch.Tensor:
    """dx[M, K] = dy[M, N] @ w[N, K] with w in its nn.Linear [out, in] layout (MN-major B operand, no transpose).
    dy / w: bf16 or fp16 independently (bf16 gradient against an fp16 weight); a 16-bit dx takes dy's dtype."""
    _need_cuda(dy, w)
    assert dy.dtype in _H16 and w.dtype in _H16 and dy.is_contiguous() and w.is_contiguous()
    M, N = dy.shape
    K = w.shape[1]
    assert w.shape[0] == N
    out = torch.empty(M, K, device=dy.device, dtype=torch.float32 if out_f32 else dy.dtype)
    g = GemmArgs()
    g.a_fmt, g.w_fmt, g.out_fmt = fmt_of(dy.dtype), fmt_of(w.dtype), fmt_of(dy.dtype)
    g.A, g.a_dim0, g.a_row_stride, g.a_batch_stride, g.a_rows = dy.data_ptr(), N, N, 0, M
    g.W, g.w_rows, g.w_cols, g.b_mn = w.data_ptr(), N, K, 1
    g.M, g.N, g.batches, g.groups, g.taps, g.k_per_tap = M, K, 1, 1, 1, N
    g.epi = EPI_F32 if out_f32 else EPI_BF16
    g.out, g.ldo = out.data_ptr(), K
    g.block_n, g.cta_group = block_n, cta_group
    _lib.check(_lib.load().b2s_gemm_bf16(C.byref(g), _stream()), "gemm_dgrad")
    return out


def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor, out: torch.Tensor, *, k_splits: int = 0, block_n: int = 0,
               cta_group: int = 0) -> torch.Tensor:
    """out[N, K] += sum over rows of dy[.., N]^T x[.., K] (both operands MN-major, atomic fp32 accumulation).
    dy / x: [rows, N] / [rows, K] or batched [B, T, N] / [B, T, K], bf16 or fp16 independently."""
    _need_cuda(dy, x, out)
    assert dy.dtype in _H16 and x.dtype in _H16 and dy.is_contiguous() and x.is_contiguous()
    assert out.dtype == torch.float32 and out.is_contiguous()
    if dy.dim() == 2:
        dy, x = dy[None], x[None]
    B, T, N = dy.shape
    K = x.shape[2]
    assert out.shape == (N, K) and x.shape[:2] == (B, T)
    g = GemmArgs()
    g.a_fmt, g.w_fmt = fmt_of(dy.dtype), fmt_of(x.dtype)
    g.A, g.a_dim0, g.a_row_stride, g.a_batch_stride, g.a_rows, g.a_mn = dy.data_ptr(), N, N, T * N, T, 1
    g.W, g.w_rows, g.w_cols, g.b_mn, g.w_row_stride, g.w_batch_stride = x.data_ptr(), T, K, 1, K, T * K
    g.M, g.N, g.batches, g.groups, g.taps, g.k_per_tap, g.k_batches = N, K, 1, 1, 1, T, B
    g.epi = EPI_ACCUM_F32
    g.out, g.ldo = out.data_ptr(), K
    g.k_splits, g.block_n, g.cta_group = k_splits, block_n, cta_group
    _lib.check(_lib.load().b2s_gemm_bf16(C.byref(g), _stream()), "gemm_wgrad")
    return out


def gemm_raw(args: GemmArgs) -> None:
    _lib.check(_lib.load().b2s_gemm_bf16(C.byref(args), _stream()), "gemm")


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, gelu: bool = False,
              out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """x fp32 or 16-bit; the output (and a 16-bit input) is in ONE 16-bit format: out's / x's / out_dtype."""
    _need_cuda(x, gamma, beta)
    assert x.is_contiguous() and x.dtype in (torch.float32,) + _H16
    C_ = x.shape[-1]
    rows = x.numel() // C_
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=x.dtype if x.dtype in _H16 else out_dtype)
    assert x.dtype == torch.float32 or x.dtype == out.dtype
    _lib.check(_lib.load().b2s_layernorm_fwd(x.data_ptr(), int(x.dtype in _H16), gamma.data_ptr(),
                                             beta.data_ptr(), eps, int(gelu), out.data_ptr(), rows, C_,
                                             fmt_of(out.dtype), _stream()), "layernorm")
    return out


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, rows_index: Optional[torch.Tensor] = None,
            out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    _need_cuda(x, w, rows_index)
    assert x.is_contiguous() and x.dtype == torch.float32
    C_ = x.shape[-1]
    fmt = fmt_of(out_dtype)
    if rows_index is None:
        rows = x.numel() // C_
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
        _lib.check(_lib.load().b2s_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), eps, out.data_ptr(), rows, C_, fmt,
                                               _stream()), "rmsnorm")
    else:
        assert rows_index.dtype == torch.int32
        rows = rows_index.numel()
        out = torch.empty(rows, C_, device=x.device, dtype=out_dtype)
        _lib.check(_lib.load().b2s_rmsnorm_gather_fwd(x.data_ptr(), rows_index.data_ptr(), w.data_ptr(), eps,
                                                      out.data_ptr(), rows, C_, fmt, _stream()), "rmsnorm_gather")
    return out


def layernorm_avgpool(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, kernel: int,
                      stride: int, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """x fp32 [B, T, C] -> bf16 [B, (T-kernel)//stride+1, C] = AvgPool1d(LN(x)) over time."""
    _need_cuda(x, gamma, beta)
    B, T, C_ = x.shape
    To = (T - kernel) // stride + 1 if T >= kernel else 0
    out = torch.empty(B, To, C_, device=x.device, dtype=out_dtype)
    _lib.check(_lib.load().b2s_layernorm_avgpool_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps,
                                                     out.data_ptr(), B, T, C_, kernel, stride, To,
                                                     fmt_of(out_dtype), _stream()),
               "layernorm_avgpool")
    return out


def conv0_ln_gelu(wave: torch.Tensor, w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    _need_cuda(wave, w, b, gamma, beta)
    assert wave.dtype == torch.float32 and wave.dim() == 2 and wave.stride(1) == 1
    B, T = wave.shape
    To = (T - 10) // 5 + 1
    out = torch.empty(B, To, 512, device=wave.device, dtype=out_dtype)
    _lib.check(_lib.load().b2s_conv0_ln_gelu_fwd(wave.data_ptr(), wave.stride(0), B, T, w.data_ptr(), b.data_ptr(),
                                                 gamma.data_ptr(), beta.data_ptr(), eps, out.data_ptr(), To,
                                                 fmt_of(out_dtype), _stream()),
               "conv0_ln_gelu")
    return out


def embed_splice(table: torch.Tensor, audio: Optional[torch.Tensor], row_src: torch.Tensor,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(table, audio, row_src)
    assert table.dtype in _H16 and row_src.dtype == torch.int32
    C_ = table.shape[1]
    rows = row_src.numel()
    if out is None:
        out = torch.empty(rows, C_, device=table.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (rows, C_)
    _lib.check(_lib.load().b2s_embed_splice_fwd(table.data_ptr(), _ptr(audio), row_src.data_ptr(), out.data_ptr(), rows,
                                                C_, fmt_of(table.dtype), _stream()), "embed_splice")
    return out


def rowpair_sqdiff(h: torch.Tensor, rows_a: torch.Tensor, rows_b: torch.Tensor) -> torch.Tensor:
    _need_cuda(h, rows_a, rows_b)
    out = torch.empty(rows_a.numel(), device=h.device, dtype=torch.float32)
    _lib.check(_lib.load().b2s_rowpair_sqdiff_fwd(h.data_ptr(), rows_a.data_ptr(), rows_b.data_ptr(), out.data_ptr(),
                                                  rows_a.numel(), h.shape[-1], _stream()), "rowpair_sqdiff")
    return out


def posconv_weight_pack(g: torch.Tensor, v: torch.Tensor, out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """weight-norm(dim=2) then repack [cout, cin_g, k] -> 16-bit [cout, k*cin_g] (column = tap*cin_g + c_in)."""
    _need_cuda(g, v)
    cout, cin_g, k = v.shape
    out = torch.empty(cout, k * cin_g, device=v.device, dtype=out_dtype)
    _lib.check(_lib.load().b2s_posconv_weight_pack(g.contiguous().data_ptr(), v.contiguous().data_ptr(),
                                                   out.data_ptr(), cout, cin_g, k, fmt_of(out_dtype), _stream()),
               "posconv_weight_pack")
    return out


def attention(qkv: torch.Tensor, cu_seqlens: torch.Tensor, max_seqlen: int, Hq: int, Hkv: int, D: int, scale: float,
              causal: bool, return_lse: bool = False, shared_prefix_len: int = 0):
    """qkv 16-bit [rows, (Hq+2Hkv)*D] (q | k | v) -> o [rows, Hq*D] in the same dtype (and the fp32 [rows, Hq]
    log2-domain lse). shared_prefix_len > 0 (causal only): sequence 0 is a prefix of that many rows which every other
    sequence attends to in front of its own rows."""
    _need_cuda(qkv, cu_seqlens)
    assert qkv.dtype in _H16 and qkv.is_contiguous() and cu_seqlens.dtype == torch.int32
    rows, ld = qkv.shape
    o = torch.empty(rows, Hq * D, device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty(rows, Hq, device=qkv.device, dtype=torch.float32) if return_lse else None
    base = qkv.data_ptr()
    if shared_prefix_len > 0:
        assert causal
        _lib.check(_lib.load().b2s_attention_fwd_prefix(base, base + 2 * Hq * D, base + 2 * (Hq + Hkv) * D, ld,
                                                        o.data_ptr(), Hq * D, cu_seqlens.data_ptr(),
                                                        cu_seqlens.numel() - 1, max_seqlen, rows, Hq, Hkv, D, scale,
                                                        _ptr(lse), fmt_of(qkv.dtype), int(shared_prefix_len), _stream()),
                   "attention (shared prefix)")
        return (o, lse) if return_lse else o
    _lib.check(_lib.load().b2s_attention_fwd(base, base + 2 * Hq * D, base + 2 * (Hq + Hkv) * D, ld, o.data_ptr(),
                                             Hq * D, cu_seqlens.data_ptr(), cu_seqlens.numel() - 1, max_seqlen, rows,
                                             Hq, Hkv, D, scale, int(causal), _ptr(lse), fmt_of(qkv.dtype), _stream()),
               "attention")
    return (o, lse) if return_lse else o


def attention_bwd(qkv: torch.Tensor, o: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor, cu_seqlens: torch.Tensor,
                  max_seqlen: int, Hq: int, Hkv: int, D: int, scale: float, causal: bool,
                  rope_cs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Backward of attention(): returns dqkv [rows, (Hq+2Hkv)*D] in the same q | k | v column layout. qkv, o, dout
    (and dqkv) share one 16-bit dtype."""
    _need_cuda(qkv, o, dout, lse, cu_seqlens, rope_cs)
    assert qkv.dtype in _H16 and o.dtype == qkv.dtype and dout.dtype == qkv.dtype
    rows, ld = qkv.shape
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(rows, Hq, device=qkv.device, dtype=torch.float32)
    b, db = qkv.data_ptr(), dqkv.data_ptr()
    _lib.check(_lib.load().b2s_attention_bwd(b, b + 2 * Hq * D, b + 2 * (Hq + Hkv) * D, ld, o.data_ptr(), o.stride(0),
                                             dout.data_ptr(), dout.stride(0), lse.data_ptr(), delta.data_ptr(), db,
                                             db + 2 * Hq * D, db + 2 * (Hq + Hkv) * D, ld, cu_seqlens.data_ptr(),
                                             cu_seqlens.numel() - 1, max_seqlen, rows, Hq, Hkv, D, scale, int(causal),
                                             _ptr(rope_cs), fmt_of(qkv.dtype), _stream()), "attention_bwd")
    return dqkv


class KdCeResult:
    __slots__ = ("loss_ld", "loss_ntp", "lse_s", "lse_t", "coef_kd", "coef_ce")


def kd_ce_loss(student: torch.Tensor, teacher: torch.Tensor, labels: torch.Tensor, row_offsets: torch.Tensor,
               scale_kd: float = 1.0, scale_ce: float = 1.0) -> KdCeResult:
    """Fused CE + KD over packed rows; returns per-utterance means and the statistics the backward needs."""
    _need_cuda(student, teacher, labels, row_offsets)
    assert student.dtype == torch.bfloat16 and teacher.dtype == torch.bfloat16
    assert student.stride(1) == 1 and teacher.stride(1) == 1
    assert labels.dtype == torch.int32 and row_offsets.dtype == torch.int32
    rows, V = student.shape
    U = row_offsets.numel() - 1
    lib = _lib.load()
    dev = student.device
    ws = torch.empty(max(1, lib.b2s_kd_ce_workspace_bytes(rows, V)), device=dev, dtype=torch.uint8)
    r = KdCeResult()
    r.lse_s = torch.empty(rows, device=dev, dtype=torch.float32)
    r.lse_t = torch.empty(rows, device=dev, dtype=torch.float32)
    r.coef_kd = torch.empty(rows, device=dev, dtype=torch.float32)
    r.coef_ce = torch.empty(rows, device=dev, dtype=torch.float32)
    r.loss_ld = torch.empty(U, device=dev, dtype=torch.float32)
    r.loss_ntp = torch.empty(U, device=dev, dtype=torch.float32)
    _lib.check(lib.b2s_kd_ce_loss_fwd(student.data_ptr(), teacher.data_ptr(), student.stride(0), teacher.stride(0), rows,
                                      V, labels.data_ptr(), row_offsets.data_ptr(), U, scale_kd, scale_ce,
                                      ws.data_ptr(), r.lse_s.data_ptr(), r.lse_t.data_ptr(), r.coef_kd.data_ptr(),
                                      r.coef_ce.data_ptr(), r.loss_ld.data_ptr(), r.loss_ntp.data_ptr(), _stream()),
               "kd_ce_loss_fwd")
    return r


def kd_ce_loss_bwd(student: torch.Tensor, teacher: torch.Tensor, labels: torch.Tensor, res: KdCeResult,
                   out: Optional[torch.Tensor] = None, loss_scale: Optional[torch.Tensor] = None,
                   out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """d(loss)/d(student logits) in `out_dtype` (the gradient format of the model), times the device scalar
    `loss_scale` when given (GradScaler)."""
    _need_cuda(student, teacher, labels, loss_scale)
    rows, V = student.shape
    if out is None:
        out = torch.empty(rows, V, device=student.device, dtype=out_dtype)
    assert loss_scale is None or loss_scale.dtype == torch.float32
    _lib.check(_lib.load().b2s_kd_ce_loss_bwd(student.data_ptr(), teacher.data_ptr(), student.stride(0),
                                              teacher.stride(0), rows, V, labels.data_ptr(), res.lse_s.data_ptr(),
                                              res.lse_t.data_ptr(), res.coef_kd.data_ptr(), res.coef_ce.data_ptr(),
                                              _ptr(loss_scale), out.data_ptr(), out.stride(0), fmt_of(out.dtype),
                                              _stream()), "kd_ce_loss_bwd")
    return out


def gather_rows(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """out[i] = src[index[i]] for fp32 rows (index < 0 -> zeros)."""
    _need_cuda(src, index)
    assert src.dtype == torch.float32 and src.is_contiguous() and index.dtype == torch.int32
    out = torch.empty(index.numel(), src.shape[1], device=src.device, dtype=torch.float32)
    _lib.check(_lib.load().b2s_gather_rows_f32(src.data_ptr(), index.data_ptr(), out.data_ptr(), index.numel(),
                                               src.shape[1], _stream()), "gather_rows")
    return out


def whisper_log_mel(waves: torch.Tensor, mel_filters: torch.Tensor) -> torch.Tensor:
    """(B, samples) fp32 waveforms -> (B, 80, samples // 160) fp32 log-mel features, WhisperFeatureExtractor's numbers
    (TF/models/whisper/feature_extraction_whisper.py:105-133). mel_filters: fp32 [201, 80]."""
    _need_cuda(waves, mel_filters)
    assert waves.dtype == torch.float32 and waves.dim() == 2 and waves.stride(1) == 1
    assert mel_filters.dtype == torch.float32 and mel_filters.shape == (201, 80) and mel_filters.is_contiguous()
    B, n = waves.shape
    frames = n // 160
    out = torch.empty(B, 80, frames, device=waves.device, dtype=torch.float32)
    mx = torch.empty(B, device=waves.device, dtype=torch.int32)
    _lib.check(_lib.load().b2s_whisper_log_mel(waves.data_ptr(), waves.stride(0), B, n, mel_filters.data_ptr(),
                                               out.data_ptr(), frames, mx.data_ptr(), _stream()), "whisper_log_mel")
    return out


def launch_count() -> int:
    return int(_lib.load().b2s_launch_count())


def layernorm_bwd_ex(x: torch.Tensor, gamma: torch.Tensor, beta: Optional[torch.Tensor], dy: torch.Tensor, eps: float,
                     *, gelu: bool = False, dh: Optional[torch.Tensor] = None, accumulate: bool = False,
                     dx_dtype: Optional[torch.dtype] = None, dgamma: Optional[torch.Tensor] = None,
                     dbeta: Optional[torch.Tensor] = None, dh_colsum: Optional[torch.Tensor] = None):
    """Backward of y = [gelu](LayerNorm(x)): returns (dh fp32 or None, dx 16-bit or None, dgamma, dbeta); dgamma / dbeta /
    dh_colsum are ACCUMULATED into when given. x, dy: fp32 or one 16-bit format; C in {256, 512, 1024}."""
    _need_cuda(x, gamma, beta, dy, dh, dgamma, dbeta, dh_colsum)
    C_ = x.shape[-1]
    rows = x.numel() // C_
    h16 = [t.dtype for t in (x, dy) if t.dtype in _H16] + ([dx_dtype] if dx_dtype is not None else [])
    assert len(set(h16)) <= 1, "one 16-bit format per call"
    fmt = fmt_of(h16[0]) if h16 else FMT_BF16
    if dh is None and (dx_dtype is None or accumulate):
        dh = torch.zeros(x.shape, device=x.device, dtype=torch.float32)
    dx = torch.empty(x.shape, device=x.device, dtype=dx_dtype) if dx_dtype is not None else None
    dgamma = torch.zeros(C_, device=x.device) if dgamma is None else dgamma
    dbeta = torch.zeros(C_, device=x.device) if dbeta is None else dbeta
    _lib.check(_lib.load().b2s_layernorm_bwd_ex(x.data_ptr(), int(x.dtype in _H16), gamma.data_ptr(), _ptr(beta),
                                                int(gelu), eps, dy.data_ptr(), int(dy.dtype in _H16), _ptr(dh),
                                                int(accumulate), _ptr(dx), dgamma.data_ptr(), dbeta.data_ptr(), rows, C_,
                                                fmt, _ptr(dh_colsum), _stream()), "layernorm_bwd_ex")
    return dh, dx, dgamma, dbeta


def conv0_bwd(wave: torch.Tensor, w: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
              eps: float, dy: torch.Tensor):
    """Parameter gradients of y = gelu(LayerNorm(conv1d(wave, w, b, stride 5))) (HuBERT-large layer 0: 512 x 1 x 10):
    returns (dW [512, 10], db, dgamma, dbeta), fp32. dy: [B, frames, 512] in one 16-bit format."""
    _need_cuda(wave, w, b, gamma, beta, dy)
    assert wave.dtype == torch.float32 and wave.is_contiguous() and dy.dtype in _H16 and dy.is_contiguous()
    B, S = wave.shape
    frames = dy.shape[1]
    outs = [torch.zeros(n, device=wave.device) for n in (w.numel(), 512, 512, 512)]
    _lib.check(_lib.load().b2s_conv0_bwd(wave.data_ptr(), S, B, S, w.data_ptr(), b.data_ptr(), gamma.data_ptr(),
                                         beta.data_ptr(), eps, dy.data_ptr(), frames, *[o.data_ptr() for o in outs],
                                         fmt_of(dy.dtype), _stream()), "conv0_bwd")
    return outs[0].view_as(w), outs[1], outs[2], outs[3]


def gelu_bwd(pre: torch.Tensor, dy: torch.Tensor, colsum: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dpre = dy * gelu'(pre) (erf form) for 16-bit [rows, F] tensors of one format; `colsum` ([F] fp32, F % 2048 == 0)
    additionally receives += the column sums of dpre (the bias gradient of the Linear that produced `pre`)."""
    _need_cuda(pre, dy, colsum)
    assert pre.dtype in _H16 and dy.dtype == pre.dtype and pre.is_contiguous() and dy.is_contiguous()
    out = torch.empty_like(pre)
    F_ = pre.shape[-1]
    _lib.check(_lib.load().b2s_gelu_bwd(pre.data_ptr(), dy.data_ptr(), out.data_ptr(), pre.numel(), fmt_of(pre.dtype),
                                        _ptr(colsum), F_ if colsum is not None else 0, _stream()), "gelu_bwd")
    return out
