"""GPU parity of the tcgen05 GEMM (through the C ABI) against a torch fp32 reference of the same op.

Tolerances: operands are bf16-exact in both paths and accumulation is fp32, so fp32-output epilogues must agree
to ~1e-5 relative L2; bf16-output epilogues to bf16 rounding (2^-9 ~ 2e-3 relative per element -> 3e-3 L2).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5
TOL_BF16 = 3e-3


def _mk(M, N, K, dev, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    return a, w, b


@pytest.mark.parametrize("bn,cg", [(256, 1), (128, 1), (64, 1), (256, 2), (128, 2)],
                         ids=["bn256cg1", "bn128cg1", "bn64cg1", "bn256cg2", "bn128cg2"])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (499, 1024, 1024), (1000, 3072, 1024),
                                   (130, 264, 200), (4096, 512, 1536), (317, 128256 // 4, 512)])
def test_gemm_plain_f32(cuda, M, N, K, bn, cg):
    from llm_speech_summarization_b200 import ops
    a, w, b = _mk(M, N, K, cuda)
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_F32, block_n=bn, cta_group=cg)
    ref = a.float() @ w.float().t() + b
    torch.cuda.synchronize()
    assert rel_l2(out, ref) < TOL_F32


@pytest.mark.parametrize("cg", [1, 2], ids=["cg1", "cg2"])
def test_gemm_bf16_gelu(cuda, cg):
    from llm_speech_summarization_b200 import ops
    a, w, b = _mk(1497, 4096, 1024, cuda, seed=1)
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_BF16, act=ops.ACT_GELU, cta_group=cg)
    ref = F.gelu(a.float() @ w.float().t() + b)
    assert rel_l2(out.float(), ref) < TOL_BF16


@pytest.mark.parametrize("cg", [1, 2], ids=["cg1", "cg2"])
def test_gemm_resid_inplace(cuda, cg):
    from llm_speech_summarization_b200 import ops
    a, w, b = _mk(777, 1024, 4096, cuda, seed=2)
    h = torch.randn(777, 1024, device=cuda)
    ref = h + (a.float() @ w.float().t() + b)
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_RESID_F32, resid=h, out=h, cta_group=cg)
    assert out.data_ptr() == h.data_ptr()
    assert rel_l2(out, ref) < TOL_F32


@pytest.mark.parametrize("cg", [1, 2], ids=["cg1", "cg2"])
def test_gemm_swiglu(cuda, cg):
    from llm_speech_summarization_b200 import ops
    from llm_speech_summarization_b200.packing import pack_gate_up
    M, Hd, Fd = 333, 512, 1024
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(M, Hd, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    wg = (torch.randn(Fd, Hd, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    wu = (torch.randn(Fd, Hd, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    out = ops.gemm(a, pack_gate_up(wg, wu), epi=ops.EPI_SWIGLU, cta_group=cg)
    ref = F.silu(a.float() @ wg.float().t()) * (a.float() @ wu.float().t())
    assert out.shape == (M, Fd)
    assert rel_l2(out.float(), ref) < TOL_BF16


@pytest.mark.parametrize("cg", [1, 2], ids=["cg1", "cg2"])
def test_gemm_rope(cuda, cg):
    from llm_speech_summarization_b200 import ops
    M, Hd, Hq, Hkv, D = 300, 512, 4, 2, 128
    g = torch.Generator().manual_seed(4)
    a = (torch.randn(M, Hd, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    w = (torch.randn((Hq + 2 * Hkv) * D, Hd, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    pos = torch.randint(0, 400, (M,), generator=g).to(torch.int32).to(cuda)
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
    ang = torch.arange(512).float()[:, None] * inv[None, :]
    cs = torch.cat([ang.cos(), ang.sin()], dim=1).contiguous().to(cuda)
    out = ops.gemm(a, w, epi=ops.EPI_ROPE, rope_cs=cs, positions=pos, rope_cols=(Hq + Hkv) * D, cta_group=cg)
    y = (a.float() @ w.float().t()).view(M, Hq + 2 * Hkv, D)
    cos = torch.cat([ang.cos(), ang.cos()], 1).to(cuda)[pos.long()][:, None, :]
    sin = torch.cat([ang.sin(), ang.sin()], 1).to(cuda)[pos.long()][:, None, :]
    rot = torch.cat([-y[..., D // 2:], y[..., :D // 2]], dim=-1)
    yr = y * cos + rot * sin
    ref = torch.cat([yr[:, :Hq + Hkv], y[:, Hq + Hkv:]], dim=1).reshape(M, -1)
    assert rel_l2(out.float(), ref) < TOL_BF16


@pytest.mark.parametrize("k,s,T", [(3, 2, 1001), (2, 2, 640), (3, 2, 15999)])
def test_gemm_strided_conv_view(cuda, k, s, T):
    """Conv1d(512->512, k, stride s) as a GEMM over the overlapping window view."""
    from llm_speech_summarization_b200 import ops
    from llm_speech_summarization_b200._lib import GemmArgs
    B, Cc = 3, 512
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(B, T, Cc, generator=g) * 0.5).to(torch.bfloat16).to(cuda)  # channels-last
    w = (torch.randn(Cc, Cc, k, generator=g) * 0.03).to(torch.bfloat16).to(cuda)
    b = torch.randn(Cc, generator=g).to(cuda)
    To = (T - k) // s + 1
    wp = w.permute(0, 2, 1).reshape(Cc, k * Cc).contiguous()
    out = torch.empty(B, To, Cc, device=cuda, dtype=torch.float32)
    a = GemmArgs()
    a.A, a.a_dim0, a.a_row_stride, a.a_batch_stride, a.a_rows = x.data_ptr(), k * Cc, s * Cc, T * Cc, To
    a.W, a.w_rows, a.w_cols = wp.data_ptr(), Cc, k * Cc
    a.M, a.N, a.batches, a.groups, a.taps, a.k_per_tap = To, Cc, B, 1, 1, k * Cc
    a.epi, a.bias, a.out, a.ldo, a.out_batch_rows = ops.EPI_F32, b.data_ptr(), out.data_ptr(), Cc, To
    ops.gemm_raw(a)
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), b, stride=s).transpose(1, 2)
    assert rel_l2(out, ref) < TOL_F32


def test_gemm_grouped_posconv(cuda):
    """Grouped Conv1d(1024,1024,k=128,pad=64,groups=16), last frame dropped, GELU, residual add."""
    from llm_speech_summarization_b200 import ops
    from llm_speech_summarization_b200._lib import GemmArgs
    B, T, H, K, G = 2, 499, 1024, 128, 16
    gen = torch.Generator().manual_seed(6)
    h = torch.randn(B, T, H, generator=gen).to(cuda)
    v = (torch.randn(H, H // G, K, generator=gen) * 0.02).to(cuda)
    gg = (torch.rand(1, 1, K, generator=gen) + 0.5).to(cuda)
    bias = torch.randn(H, generator=gen).to(cuda)
    wp = ops.posconv_weight_pack(gg, v)
    wn = gg * v / v.norm(dim=(0, 1), keepdim=True)
    wp_ref = wn.permute(0, 2, 1).reshape(H, K * (H // G))
    assert rel_l2(wp.float(), wp_ref) < 3e-3
    x = h.to(torch.bfloat16)
    out = h.clone()
    a = GemmArgs()
    a.A, a.a_dim0, a.a_row_stride, a.a_batch_stride, a.a_rows = x.data_ptr(), H, H, T * H, T
    a.W, a.w_rows, a.w_cols = wp.data_ptr(), H, K * 64
    a.M, a.N, a.batches, a.groups, a.taps, a.k_per_tap = T, 64, B, G, K, 64
    a.a_pad, a.a_group_off, a.w_group_off = K // 2, 64, 64
    a.epi, a.act, a.bias = ops.EPI_RESID_F32, ops.ACT_GELU, bias.data_ptr()
    a.out, a.resid, a.ldo, a.out_batch_rows = out.data_ptr(), out.data_ptr(), H, T
    ops.gemm_raw(a)
    wq = wp.float().view(H, K, H // G).permute(0, 2, 1).contiguous()  # bf16-rounded normalised weight
    y = F.conv1d(x.float().transpose(1, 2), wq, bias, padding=K // 2, groups=G)[:, :, :-1]
    ref = h + F.gelu(y).transpose(1, 2)
    assert rel_l2(out, ref) < TOL_F32


def test_gemm_rejects_bad_args(cuda):
    from llm_speech_summarization_b200 import ops
    a = torch.zeros(8, 60, device=cuda, dtype=torch.bfloat16)  # K stride not a multiple of 8
    w = torch.zeros(16, 60, device=cuda, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, w)


# ------------------------------------------------------------------------------------------------ backward GEMMs
@pytest.mark.parametrize("bn,cg", [(256, 1), (128, 1), (64, 1), (256, 2), (128, 2)],
                         ids=["bn256cg1", "bn128cg1", "bn64cg1", "bn256cg2", "bn128cg2"])
@pytest.mark.parametrize("M,N,K", [(128, 64, 128), (499, 1024, 1024), (1000, 4096, 1024), (130, 200, 264),
                                   (998, 512, 1536)])
def test_gemm_dgrad_mn_major_w(cuda, M, N, K, bn, cg):
    """dX = dY @ W with W in nn.Linear layout: the B operand is read MN-major (no transposed copy)."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(3)
    dy = (torch.randn(M, N, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(cuda)
    out = ops.gemm_dgrad(dy, w, out_f32=True, block_n=bn, cta_group=cg)
    ref = dy.float() @ w.float()
    assert rel_l2(out, ref) < TOL_F32
    out_bf = ops.gemm_dgrad(dy, w, block_n=bn, cta_group=cg)
    assert rel_l2(out_bf.float(), ref) < TOL_BF16


@pytest.mark.parametrize("bn,cg", [(256, 1), (128, 1), (64, 1), (256, 2), (128, 2)],
                         ids=["bn256cg1", "bn128cg1", "bn64cg1", "bn256cg2", "bn128cg2"])
@pytest.mark.parametrize("B,T,N,K,splits", [(1, 128, 128, 64, 1), (1, 499, 1024, 1024, 0), (3, 333, 3072, 1024, 0),
                                            (2, 1000, 264, 200, 3), (4, 499, 1024, 4096, 1)])
def test_gemm_wgrad_mn_major_both(cuda, B, T, N, K, splits, bn, cg):
    """dW += sum_{b,t} dY[b,t,:]^T X[b,t,:]: both operands MN-major, reduction over batches, split-K + atomic adds."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(4)
    dy = (torch.randn(B, T, N, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    x = (torch.randn(B, T, K, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    init = torch.randn(N, K, generator=g).to(cuda)
    out = init.clone()
    ops.gemm_wgrad(dy, x, out, k_splits=splits, block_n=bn, cta_group=cg)
    ref = init + torch.einsum("btn,btk->nk", dy.float(), x.float())
    assert rel_l2(out, ref) < TOL_F32
    ops.gemm_wgrad(dy, x, out, k_splits=splits, block_n=bn, cta_group=cg)  # accumulates
    assert rel_l2(out, 2 * ref - init) < TOL_F32


@pytest.mark.parametrize("k,s,T", [(3, 2, 1001), (2, 2, 640)])
def test_gemm_wgrad_strided_conv_view(cuda, k, s, T):
    """Conv1d weight gradient: X operand = the overlapping window view [B, Tout, k*C] (row stride s*C) read MN-major."""
    from llm_speech_summarization_b200 import ops
    from llm_speech_summarization_b200._lib import GemmArgs
    Cin, Cout, B = 512, 512, 2
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(B, T, Cin, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    To = (T - k) // s + 1
    dy = (torch.randn(B, To, Cout, generator=g) * 0.5).to(torch.bfloat16).to(cuda)
    out = torch.zeros(Cout, k * Cin, device=cuda)
    a = GemmArgs()
    a.A, a.a_dim0, a.a_row_stride, a.a_batch_stride, a.a_rows, a.a_mn = dy.data_ptr(), Cout, Cout, To * Cout, To, 1
    a.W, a.w_rows, a.w_cols, a.b_mn, a.w_row_stride, a.w_batch_stride = x.data_ptr(), To, k * Cin, 1, s * Cin, T * Cin
    a.M, a.N, a.batches, a.groups, a.taps, a.k_per_tap, a.k_batches = Cout, k * Cin, 1, 1, 1, To, B
    a.epi, a.out, a.ldo = ops.EPI_ACCUM_F32, out.data_ptr(), k * Cin
    ops.gemm_raw(a)
    # reference: conv1d weight gradient in torch's [Cout, Cin, k] layout -> packed [Cout, k*Cin]
    xw = x.float().permute(0, 2, 1).requires_grad_(False)
    w = torch.zeros(Cout, Cin, k, device=cuda, requires_grad=True)
    y = F.conv1d(xw, w, stride=s)
    (gw,) = torch.autograd.grad(y, w, dy.float().permute(0, 2, 1))
    ref = gw.permute(0, 2, 1).reshape(Cout, k * Cin)
    assert rel_l2(out, ref) < TOL_F32


# ---- 16-bit operand formats (b2s.h B2S_FMT_*): fp16 and bf16, never mixed
_DT = {"bf16": torch.bfloat16, "f16": torch.float16}


@pytest.mark.parametrize("cg", [1, 2], ids=["cg1", "cg2"])
def test_gemm_f16_operands_f32(cuda, cg):
    """fp16 operands: values exactly representable in both 16-bit formats, so the fp32 result matches to fp32 rounding."""
    from llm_speech_summarization_b200 import ops
    a, w, b = _mk(777, 1024, 512, cuda, seed=11)     # bf16-exact values of moderate magnitude: exact in fp16 as well
    a, w = a.to(torch.float16), w.to(torch.float16)
    out = ops.gemm(a, w, bias=b, epi=ops.EPI_F32, cta_group=cg)
    ref = a.float() @ w.float().t() + b
    assert rel_l2(out, ref) < TOL_F32


def test_gemm_rejects_mixed_operand_formats(cuda):
    """tcgen05 kind::f16 encodes a_format / b_format separately, but a bf16 x fp16 pair traps with an illegal instruction
    on B200 (measured in round 2): the library refuses it up front -- which is why fp16 activations imply fp16
    gradients (and a loss scale) on the training path."""
    from llm_speech_summarization_b200 import ops
    a, w, _ = _mk(128, 256, 64, cuda, seed=14)
    with pytest.raises(RuntimeError, match="share one 16-bit format"):
        ops.gemm(a.to(torch.float16), w, epi=ops.EPI_F32)


def test_gemm_f16_needs_the_format_bit(cuda):
    """fp16 data fed with fp16 mantissa bits that bf16 cannot hold: the result only matches if the MMA really reads
    the operands as fp16 (guards against a silently ignored format field)."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(12)
    a = (torch.randn(256, 256, generator=g) * 0.5).to(torch.float16).to(cuda)
    w = (torch.randn(512, 256, generator=g) * 0.05).to(torch.float16).to(cuda)
    out = ops.gemm(a, w, epi=ops.EPI_F32)
    assert rel_l2(out, a.float() @ w.float().t()) < TOL_F32
    out16 = ops.gemm(a, w, epi=ops.EPI_BF16)           # 16-bit output in a's format (fp16): 2^-12 per element
    assert out16.dtype == torch.float16
    assert rel_l2(out16.float(), a.float() @ w.float().t()) < 5e-4


@pytest.mark.parametrize("dy_dt,w_dt", [("f16", "f16")])
def test_gemm_dgrad_wgrad_f16(cuda, dy_dt, w_dt):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(13)
    dy = (torch.randn(640, 768, generator=g) * 0.25).to(torch.bfloat16).to(_DT[dy_dt]).to(cuda)
    w = (torch.randn(768, 512, generator=g) * 0.05).to(torch.bfloat16).to(_DT[w_dt]).to(cuda)
    x = (torch.randn(640, 512, generator=g) * 0.5).to(torch.bfloat16).to(_DT[w_dt]).to(cuda)
    dx = ops.gemm_dgrad(dy, w, out_f32=True)
    assert rel_l2(dx, dy.float() @ w.float()) < TOL_F32
    dw = torch.zeros(768, 512, device=cuda)
    ops.gemm_wgrad(dy, x, dw)
    assert rel_l2(dw, dy.float().t() @ x.float()) < TOL_F32


def _set_tail_split(v):
    from llm_speech_summarization_b200 import _lib
    lib = _lib.load()
    old = C.c_int32(0)
    _lib.check(lib.b2s_get_option(_lib.OPT_GEMM_TAIL_SPLIT, C.byref(old)), "get_option")
    _lib.check(lib.b2s_set_option(_lib.OPT_GEMM_TAIL_SPLIT, v), "set_option")
    return old.value


@pytest.mark.timeout(180)
@pytest.mark.parametrize("M,N,K", [(6400, 3072, 2048), (6400, 3072, 1024), (10144, 3072, 1024), (15968, 1024, 1024),
                                   (9601, 3072 - 8, 1536)],
                         ids=["dgrad6400", "dgrad6400_k1024", "llm_resid", "enc_out_proj", "ragged_edges"])
@pytest.mark.parametrize("epi", ["f32", "resid"])
def test_gemm_tail_split_matches_unsplit(cuda, M, N, K, epi):
    """The last, partly filled round of the persistent grid is K-sliced (KParams.tail_*): slice 0 stores / every slice
    reduce-adds. Same numbers as the unsplit kernel up to fp32 summation order, twice in a row (the flags re-arm), and
    the shapes here all HAVE a tail on 148 SMs (300, 300, 480, 252, 456 tiles over 74 CTA pairs)."""
    from llm_speech_summarization_b200 import ops
    a, w, b = _mk(M, N, K, cuda, seed=11)
    ref = a.float() @ w.float().t() + b
    h0 = torch.randn(M, N, device=cuda)
    old = _set_tail_split(1)
    try:
        for rep in range(3):
            if epi == "f32":
                out = torch.full((M, N), float("nan"), device=cuda)
                ops.gemm(a, w, bias=b, epi=ops.EPI_F32, out=out)
                want = ref
            else:
                out = h0.clone()
                ops.gemm(a, w, bias=b, epi=ops.EPI_RESID_F32, resid=out, out=out)
                want = h0 + ref
            torch.cuda.synchronize()
            assert bool(torch.isfinite(out).all()), rep
            assert rel_l2(out, want) < TOL_F32, rep
        _set_tail_split(0)
        plain = torch.empty(M, N, device=cuda)
        ops.gemm(a, w, bias=b, epi=ops.EPI_F32, out=plain)
        assert rel_l2(plain, ref) < TOL_F32
    finally:
        _set_tail_split(old)


@pytest.mark.parametrize("epi", ["f32", "h16"])
def test_gemm_store_epilogue_is_repeatable_with_rows_beyond_m(cuda, epi):
    """M = 10144 (the packed LLM row count): in the last 256-row tile some epilogue warps own no valid row. They issue no
    bulk store, but their staging tiles are still rewritten every chunk -- the wait_group accounting has to advance with
    every chunk or a tile still being read by the previous store gets overwritten (a race found in round 2 as a flaky
    5e-3 error). 40 launches must be bit-identical and match the reference."""
    from llm_speech_summarization_b200 import ops
    M, N, K = 10144, 3072, 1024
    a, w, b = _mk(M, N, K, cuda, seed=21)
    ref = a.float() @ w.float().t() + b
    old = _set_tail_split(0)  # K-sliced tails add in arrival order: bit-exactness is only promised without them
    try:
        kw = dict(bias=b, epi=ops.EPI_F32) if epi == "f32" else dict(bias=b, epi=ops.EPI_BF16)
        first = ops.gemm(a, w, **kw).clone()
        assert rel_l2(first.float(), ref) < (TOL_F32 if epi == "f32" else TOL_BF16)
        for _ in range(40):
            out = ops.gemm(a, w, **kw)
            assert torch.equal(out, first)
    finally:
        _set_tail_split(old)


@pytest.mark.parametrize("epi,act", [("h16", "gelu"), ("h16", "none"), ("f32", "none"), ("resid", "none")])
def test_gemm_eight_epilogue_warps_match_four(cuda, epi, act):
    """MODE 5 (two epilogue warps per TMEM lane quarter, one 32-column chunk staged at a time) against the four-warp
    kernel and the torch reference, on an encoder-shaped problem with ragged M / N edges."""
    from llm_speech_summarization_b200 import _lib, ops
    lib = _lib.load()
    M, N, K = 1497, 4096 - 24, 1024
    a, w, b = _mk(M, N, K, cuda, seed=5)
    ref = a.float() @ w.float().t() + b
    if act == "gelu":
        ref = F.gelu(ref)
    h0 = torch.randn(M, N, device=cuda)
    outs = []
    for force in (2, 0):
        _lib.check(lib.b2s_set_option(_lib.OPT_GEMM_EPI8, force), "set_option")
        try:
            kw = dict(bias=b, act=ops.ACT_GELU if act == "gelu" else ops.ACT_NONE)
            if epi == "h16":
                out = ops.gemm(a, w, epi=ops.EPI_BF16, **kw).float()
            elif epi == "f32":
                out = ops.gemm(a, w, epi=ops.EPI_F32, **kw)
            else:
                out = h0.clone()
                ops.gemm(a, w, epi=ops.EPI_RESID_F32, resid=out, out=out, **kw)
                out = out - h0
            outs.append(out)
        finally:
            _lib.check(lib.b2s_set_option(_lib.OPT_GEMM_EPI8, 1), "set_option")
    tol = TOL_BF16 if epi == "h16" else (TOL_F32 if epi == "f32" else 2e-4)
    assert rel_l2(outs[0], ref) < tol and rel_l2(outs[1], ref) < tol
    if epi != "resid":
        assert torch.equal(outs[0], outs[1])  # same arithmetic, only the warp that stores a column differs
