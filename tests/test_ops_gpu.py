"""GPU parity of the memory-bound kernels and attention (through the C ABI) against torch fp32 references."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _loss_ref(s, t, labels, offs):
    """REF/utils.py:167-178 (soft_cross_entropy) and REF/model/audio_llama.py:84-98 restated per utterance."""
    s, t = s.float(), t.float()
    lds, ntps = [], []
    for u in range(len(offs) - 1):
        a, b = offs[u], offs[u + 1]
        ls = F.log_softmax(s[a:b], -1)
        pt = F.softmax(t[a:b], -1)
        lds.append((-(pt * ls).sum(-1)).mean())
        lab = labels[a:b]
        m = lab >= 0
        ntps.append(F.cross_entropy(s[a:b][m], lab[m].long()) if m.any() else s.new_zeros(()))
    return torch.stack(lds), torch.stack(ntps)


@pytest.mark.parametrize("V,R,U", [(128256, 64, 2), (49216, 33, 3), (512, 5, 1), (16384 + 8, 7, 2)])
def test_kd_ce_loss_fwd_bwd(cuda, V, R, U):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(V + R)
    rows = R * U
    s = (torch.randn(rows, V, generator=g) * 2.0).to(torch.bfloat16).to(cuda)
    t = (torch.randn(rows, V, generator=g) * 2.0).to(torch.bfloat16).to(cuda)
    labels = torch.randint(0, V, (rows,), generator=g).to(torch.int32)
    offs = [u * R for u in range(U + 1)]
    for u in range(U):
        labels[offs[u + 1] - 1] = -1  # the last row of each utterance has no next-token target
    labels = labels.to(cuda)
    offs_t = torch.tensor(offs, dtype=torch.int32, device=cuda)
    w_kd, w_ce = 0.5, 0.25
    res = ops.kd_ce_loss(s, t, labels, offs_t, scale_kd=w_kd, scale_ce=w_ce)
    s32 = s.float().requires_grad_(True)
    ld_ref, ntp_ref = _loss_ref(s32, t, labels, offs)
    # KD loss within 1e-3 relative is the north-star tolerance; same-input fp32 math should be far tighter
    assert torch.allclose(res.loss_ld, ld_ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(res.loss_ntp, ntp_ref, rtol=1e-5, atol=1e-5)
    (w_kd * ld_ref.sum() + w_ce * ntp_ref.sum()).backward()
    ds = ops.kd_ce_loss_bwd(s, t, labels, res)
    assert rel_l2(ds.float(), s32.grad) < 4e-3  # bf16 rounding of the stored gradient


def test_kd_ce_loss_empty_rows(cuda):
    from llm_speech_summarization_b200 import ops
    s = torch.zeros(0, 512, device=cuda, dtype=torch.bfloat16)
    labels = torch.zeros(0, dtype=torch.int32, device=cuda)
    offs = torch.zeros(2, dtype=torch.int32, device=cuda)
    res = ops.kd_ce_loss(s, s, labels, offs)
    assert float(res.loss_ld[0]) == 0.0 and float(res.loss_ntp[0]) == 0.0


@pytest.mark.parametrize("C_,dtype,gelu", [(512, torch.bfloat16, True), (512, torch.float32, True),
                                           (1024, torch.float32, False), (3072, torch.float32, False)])
def test_layernorm(cuda, C_, dtype, gelu):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(C_)
    x = (torch.randn(1237, C_, generator=g) * 3 + 0.7).to(dtype).to(cuda)
    gm = torch.randn(C_, generator=g).to(cuda)
    bt = torch.randn(C_, generator=g).to(cuda)
    y = ops.layernorm(x, gm, bt, 1e-5, gelu=gelu)
    ref = F.layer_norm(x.float(), (C_,), gm, bt, 1e-5)
    if gelu:
        ref = F.gelu(ref)
    assert rel_l2(y.float(), ref) < 3e-3


def test_rmsnorm_and_gather(cuda):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(517, 3072, generator=g) * 2).to(cuda)
    w = (torch.rand(3072, generator=g) + 0.5).to(cuda)
    ref = w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5))
    assert rel_l2(ops.rmsnorm(x, w, 1e-5).float(), ref) < 3e-3
    idx = torch.tensor([5, 516, 0, 77], dtype=torch.int32, device=cuda)
    assert rel_l2(ops.rmsnorm(x, w, 1e-5, rows_index=idx).float(), ref[idx.long()]) < 3e-3


def test_layernorm_avgpool(cuda):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.randn(3, 499, 1024, generator=g).to(cuda)
    gm = torch.randn(1024, generator=g).to(cuda)
    bt = torch.randn(1024, generator=g).to(cuda)
    y = ops.layernorm_avgpool(x, gm, bt, 1e-5, 8, 4)
    ref = F.avg_pool1d(F.layer_norm(x, (1024,), gm, bt, 1e-5).transpose(1, 2), 8, 4).transpose(1, 2)
    assert y.shape == (3, 123, 1024)
    assert rel_l2(y.float(), ref) < 3e-3


@pytest.mark.parametrize("T", [400, 16000, 12345])
def test_conv0_ln_gelu(cuda, T):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(T)
    wave = (torch.randn(2, T, generator=g) * 0.1).to(cuda)
    w = (torch.randn(512, 1, 10, generator=g) * 0.3).to(cuda)
    b = torch.randn(512, generator=g).to(cuda)
    gm = torch.randn(512, generator=g).to(cuda)
    bt = torch.randn(512, generator=g).to(cuda)
    y = ops.conv0_ln_gelu(wave, w.reshape(512, 10).contiguous(), b, gm, bt, 1e-5)
    c = F.conv1d(wave[:, None, :], w, b, stride=5).transpose(1, 2)
    ref = F.gelu(F.layer_norm(c, (512,), gm, bt, 1e-5))
    assert y.shape == ref.shape
    assert rel_l2(y.float(), ref) < 3e-3


def test_embed_splice_and_sqdiff(cuda):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(9)
    table = torch.randn(1000, 3072, generator=g).to(torch.bfloat16).to(cuda)
    audio = torch.randn(50, 3072, generator=g).to(cuda)
    src = torch.tensor([3, 999, -1, -50, 0, -7], dtype=torch.int32, device=cuda)
    h = ops.embed_splice(table, audio, src)
    ref = torch.stack([table[3].float(), table[999].float(), audio[0], audio[49], table[0].float(), audio[6]])
    assert torch.equal(h, ref)
    ra = torch.tensor([0, 2], dtype=torch.int32, device=cuda)
    rb = torch.tensor([1, 3], dtype=torch.int32, device=cuda)
    sq = ops.rowpair_sqdiff(h, ra, rb)
    ref_sq = torch.stack([(ref[0] - ref[1]).pow(2).sum(), (ref[2] - ref[3]).pow(2).sum()])
    assert torch.allclose(sq, ref_sq, rtol=1e-5)


def _attn_ref(qkv, cu, Hq, Hkv, D, scale, causal):
    rows = qkv.shape[0]
    q, k, v = qkv.float().split([Hq * D, Hkv * D, Hkv * D], dim=1)
    out = torch.zeros(rows, Hq * D, device=qkv.device)
    for s in range(len(cu) - 1):
        a, b = cu[s], cu[s + 1]
        qs = q[a:b].view(b - a, Hq, D).transpose(0, 1)
        ks = k[a:b].view(b - a, Hkv, D).transpose(0, 1).repeat_interleave(Hq // Hkv, 0)
        vs = v[a:b].view(b - a, Hkv, D).transpose(0, 1).repeat_interleave(Hq // Hkv, 0)
        o = F.scaled_dot_product_attention(qs, ks, vs, is_causal=causal, scale=scale)
        out[a:b] = o.transpose(0, 1).reshape(b - a, Hq * D)
    return out


@pytest.mark.parametrize("Hq,Hkv,D,causal,lens", [
    (16, 16, 64, False, [499, 499]),
    (16, 16, 64, False, [1, 63, 64, 65, 130]),
    (24, 8, 128, True, [200, 117, 1, 64, 129]),
    (24, 24, 128, True, [136, 400]),
    (4, 2, 128, False, [77]),
    (16, 16, 64, False, [1500]),
    (2, 1, 128, True, [128, 256, 257, 383]),
    (2, 2, 64, True, [300, 5]),
])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_attention(cuda, Hq, Hkv, D, causal, lens, dtype):
    """tcgen05 flash attention in both operand formats (Q, K, V, P and O all in `dtype`)."""
    _run_attention_case(cuda, Hq, Hkv, D, causal, lens, dtype)


@pytest.fixture
def attn_pipelined():
    """Selects the pipelined attention forward (S double-buffered in TMEM: 64 keys per step, 'K/V stages' = 3) for the
    duration of a test, through the handle options a caller would use."""
    from llm_speech_summarization_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    old = [C.c_int32(0), C.c_int32(0)]
    for o, opt in zip(old, (_lib.OPT_ATTN_KEYS_PER_STEP, _lib.OPT_ATTN_KV_STAGES)):
        _lib.check(lib.b2s_get_option(opt, C.byref(o)), "get_option")
    _lib.check(lib.b2s_set_option(_lib.OPT_ATTN_KEYS_PER_STEP, 64), "set_option")
    _lib.check(lib.b2s_set_option(_lib.OPT_ATTN_KV_STAGES, 3), "set_option")
    yield
    _lib.check(lib.b2s_set_option(_lib.OPT_ATTN_KEYS_PER_STEP, old[0].value), "set_option")
    _lib.check(lib.b2s_set_option(_lib.OPT_ATTN_KV_STAGES, old[1].value), "set_option")


@pytest.mark.timeout(120)
@pytest.mark.parametrize("Hq,Hkv,D,causal,lens", [
    (16, 16, 64, False, [499] * 4),
    (16, 16, 64, False, [1, 63, 64, 65, 130]),
    (24, 8, 128, True, [200, 117, 1, 64, 129] * 3),
    (24, 24, 128, True, [136, 400]),
    (16, 16, 64, False, [1500]),
    (2, 1, 128, True, [128, 256, 257, 383]),
    (2, 2, 64, True, [300, 5]),
    (16, 16, 64, False, [499] * 32),
    (24, 8, 128, True, [200] * 32 + [117] * 32),
])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_attention_pipelined(cuda, attn_pipelined, Hq, Hkv, D, causal, lens, dtype):
    """The same cases (plus the two full bench shapes, where every persistent CTA crosses many item boundaries) through
    the pipelined forward: the issuer runs two key blocks ahead across (sequence, head, query-block) items."""
    _run_attention_case(cuda, Hq, Hkv, D, causal, lens, dtype)


def test_attention_pipelined_lse_feeds_backward(cuda, attn_pipelined):
    from llm_speech_summarization_b200 import ops
    Hq, Hkv, D, lens = 4, 2, 128, [200, 117, 77]
    g = torch.Generator().manual_seed(12)
    rows = sum(lens)
    qkv = (torch.randn(rows, (Hq + 2 * Hkv) * D, generator=g) * 0.7).to(torch.float16).to(cuda)
    dout = torch.randn(rows, Hq * D, generator=g).to(torch.float16).to(cuda)
    cu = [0, 200, 317, 394]
    cu_t = torch.tensor(cu, dtype=torch.int32, device=cuda)
    o, lse = ops.attention(qkv, cu_t, max(lens), Hq, Hkv, D, 1.0 / math.sqrt(D), True, return_lse=True)
    dqkv = ops.attention_bwd(qkv, o, dout, lse, cu_t, max(lens), Hq, Hkv, D, 1.0 / math.sqrt(D), True)
    x = qkv.float().requires_grad_(True)
    _attn_ref(x, cu, Hq, Hkv, D, 1.0 / math.sqrt(D), True).backward(dout.float())
    assert rel_l2(dqkv.float(), x.grad) < 3e-3


def test_attention_f16_sharp_scores_stay_finite(cuda):
    """fp16 P tops out at 65504: a later key block whose scores outgrow the running reference by far more than 2^14
    must be handled by the immediate re-reference, not overflow to inf."""
    from llm_speech_summarization_b200 import ops
    Hq = Hkv = 2
    D, L = 64, 300
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(L, 3 * Hq * D, generator=g)
    qkv[:, :Hq * D] *= 3.0
    qkv[200:, Hq * D:2 * Hq * D] *= 9.0  # keys of the later blocks score much higher
    qkv = qkv.to(torch.float16).to(cuda)
    cu = torch.tensor([0, L], dtype=torch.int32, device=cuda)
    o = ops.attention(qkv, cu, L, Hq, Hkv, D, 0.125, False)
    ref = _attn_ref(qkv, [0, L], Hq, Hkv, D, 0.125, False)
    assert bool(torch.isfinite(o.float()).all())
    assert rel_l2(o.float(), ref) < 6e-3


def _run_attention_case(cuda, Hq, Hkv, D, causal, lens, dtype=torch.bfloat16):
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(sum(lens) + D)
    rows = sum(lens)
    qkv = torch.randn(rows, (Hq + 2 * Hkv) * D, generator=g).to(dtype).to(cuda)
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=cuda)
    scale = 1.0 / math.sqrt(D)
    o = ops.attention(qkv, cu_t, max(lens), Hq, Hkv, D, scale, causal)
    ref = _attn_ref(qkv, cu, Hq, Hkv, D, scale, causal)
    assert o.dtype == dtype
    # P is rounded to the operand format before P.V and the output is rounded to it: 2^-9 (bf16) / 2^-12 (fp16)
    assert rel_l2(o.float(), ref) < (6e-3 if dtype == torch.bfloat16 else 1e-3)


@pytest.mark.parametrize("Hq,Hkv,D,causal,lens", [
    (16, 16, 64, False, [499, 130]),
    (4, 4, 64, False, [1, 63, 64, 65]),
    (24, 8, 128, True, [200, 117, 1, 64, 129]),
    (6, 6, 128, True, [136, 300]),
    (4, 2, 64, True, [77, 200]),
])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_attention_backward(cuda, Hq, Hkv, D, causal, lens, dtype):
    """dQ / dK / dV of the packed attention vs autograd through fp32 SDPA on the same 16-bit inputs (every 16-bit
    tensor of the call -- q, k, v, o, dout, dq, dk, dv -- shares `dtype`)."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(sum(lens) * D + Hq)
    rows = sum(lens)
    qkv = (torch.randn(rows, (Hq + 2 * Hkv) * D, generator=g) * 0.7).to(dtype).to(cuda)
    dout = torch.randn(rows, Hq * D, generator=g).to(dtype).to(cuda)
    cu = [0]
    for L in lens:
        cu.append(cu[-1] + L)
    cu_t = torch.tensor(cu, dtype=torch.int32, device=cuda)
    scale = 1.0 / math.sqrt(D)
    o, lse = ops.attention(qkv, cu_t, max(lens), Hq, Hkv, D, scale, causal, return_lse=True)
    dqkv = ops.attention_bwd(qkv, o, dout, lse, cu_t, max(lens), Hq, Hkv, D, scale, causal)
    assert dqkv.dtype == dtype
    x = qkv.float().requires_grad_(True)
    ref = _attn_ref(x, cu, Hq, Hkv, D, scale, causal)
    ref.backward(dout.float())
    # lse sanity: natural-log lse of the scaled scores = lse2 * ln 2
    q, k, _ = qkv.float().split([Hq * D, Hkv * D, Hkv * D], dim=1)
    a, b = cu[0], cu[1]
    s = (q[a:b].view(b - a, Hq, D).transpose(0, 1) @ k[a:b].view(b - a, Hkv, D).transpose(0, 1).repeat_interleave(
        Hq // Hkv, 0).transpose(1, 2)) * scale
    if causal:
        s = s.masked_fill(~torch.ones(b - a, b - a, dtype=torch.bool, device=cuda).tril(), float("-inf"))
    assert torch.allclose(lse[a:b].t() * math.log(2.0), torch.logsumexp(s, -1), atol=2e-3, rtol=2e-3)
    for name, sl in (("dq", slice(0, Hq * D)), ("dk", slice(Hq * D, (Hq + Hkv) * D)), ("dv", slice((Hq + Hkv) * D, None))):
        assert rel_l2(dqkv[:, sl].float(), x.grad[:, sl]) < (1.5e-2 if dtype == torch.bfloat16 else 3e-3), name


# ------------------------------------------------------------------------------------------------ Whisper log-mel (f2)
@pytest.mark.parametrize("B,seconds_of_signal", [(1, 30.0), (3, 7.3)])
def test_whisper_log_mel_matches_oracle(cuda, B, seconds_of_signal):
    """GPU log-mel == the oracle's restatement of WhisperFeatureExtractor (itself pinned against transformers in the
    CPU suite). Tolerance 2e-4 absolute on features that live in (-1, 1.6): fp32 DFT vs float64 FFT."""
    transformers = pytest.importorskip("transformers")
    import numpy as np
    from oracle import reference_math as rm
    from llm_speech_summarization_b200 import ops
    mel = np.asarray(transformers.WhisperFeatureExtractor().mel_filters, dtype=np.float32)
    g = torch.Generator().manual_seed(3 + B)
    n = int(seconds_of_signal * 16000)
    waves = torch.zeros(B, 480000)
    waves[:, :n] = torch.randn(B, n, generator=g) * 0.1 * torch.linspace(0.2, 1.0, n)  # zero-padded to the 30 s window
    out = ops.whisper_log_mel(waves.to(cuda), torch.from_numpy(mel).to(cuda))
    assert out.shape == (B, 80, 3000)
    for b in range(B):
        ref = rm.whisper_log_mel(waves[b].numpy(), mel)
        assert float((out[b].cpu().double() - torch.from_numpy(ref)).abs().max()) < 2e-4


def test_whisper_extract_features_feeds_the_encoder(cuda):
    """AudioEncoder.extract_features (GPU) == the extractor's input_features, and the encoder consumes it directly."""
    transformers = pytest.importorskip("transformers")
    from oracle import configs
    from helpers import ns_config_whisper
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    cfg = configs.WhisperCfg(hidden=256, layers=1, heads=4, ffn=512, max_positions=1500, llm_dim=256)
    enc = AudioEncoder(ns_config_whisper(cfg), cuda)
    enc.load_state_dict(configs.make_whisper_state_dict(cfg, seed=5), strict=True)
    enc.eval().to(cuda)
    g = torch.Generator().manual_seed(8)
    wave = torch.randn(52000, generator=g) * 0.05
    want = enc.feature_extractor(wave.numpy(), sampling_rate=16000, return_tensors="pt").input_features  # CPU STFT
    padded = torch.zeros(1, 480000)
    padded[0, :52000] = wave
    got = enc.extract_features(padded.to(cuda))
    assert got.shape == want.shape and float((got.cpu() - want).abs().max()) < 2e-4
    with torch.no_grad():
        a = enc.forward_fp32(got)
        b = enc.forward_fp32(want.to(cuda))
    assert rel_l2(a, b) < 5e-3


# ------------------------------------------------------------------------------------------------ fp16 operand format
def test_memory_bound_ops_in_f16(cuda):
    """The norm / splice / cast family writes (and reads) fp16 when asked: same math, 2^-12 rounding instead of 2^-9."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(31)
    x = (torch.randn(333, 1024, generator=g) * 3 + 0.7).to(cuda)
    gm, bt = torch.randn(1024, generator=g).to(cuda), torch.randn(1024, generator=g).to(cuda)
    y = ops.layernorm(x, gm, bt, 1e-5, out_dtype=torch.float16)
    assert y.dtype == torch.float16
    assert rel_l2(y.float(), F.layer_norm(x, (1024,), gm, bt, 1e-5)) < 4e-4
    y16 = ops.layernorm(y, gm, bt, 1e-5, gelu=True)  # fp16 in -> fp16 out
    assert y16.dtype == torch.float16
    assert rel_l2(y16.float(), F.gelu(F.layer_norm(y.float(), (1024,), gm, bt, 1e-5))) < 4e-4
    w = (torch.rand(1024, generator=g) + 0.5).to(cuda)
    ref = w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5))
    assert rel_l2(ops.rmsnorm(x, w, 1e-5, out_dtype=torch.float16).float(), ref) < 4e-4
    xp = torch.randn(2, 499, 1024, generator=g).to(cuda)
    yp = ops.layernorm_avgpool(xp, gm, bt, 1e-5, 8, 4, out_dtype=torch.float16)
    refp = F.avg_pool1d(F.layer_norm(xp, (1024,), gm, bt, 1e-5).transpose(1, 2), 8, 4).transpose(1, 2)
    assert yp.dtype == torch.float16 and rel_l2(yp.float(), refp) < 4e-4
    table = torch.randn(100, 512, generator=g).to(torch.float16).to(cuda)
    src = torch.tensor([3, 99, 0], dtype=torch.int32, device=cuda)
    assert torch.equal(ops.embed_splice(table, None, src), table[src.long()].float())


def test_kd_ce_loss_bwd_f16_with_loss_scale(cuda):
    """The gradient enters the backward pass in the model's gradient format, multiplied by the device-resident loss
    scale (GradScaler, REF/trainer.py:374): fp16 keeps 2^-12 relative precision where an unscaled fp16 store would
    flush most of these ~1e-6 values to zero."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(77)
    V, R = 4096, 12
    s = (torch.randn(R, V, generator=g) * 2.0).to(torch.bfloat16).to(cuda)
    t = (torch.randn(R, V, generator=g) * 2.0).to(torch.bfloat16).to(cuda)
    labels = torch.randint(0, V, (R,), generator=g).to(torch.int32)
    labels[-1] = -1
    labels = labels.to(cuda)
    offs = torch.tensor([0, R], dtype=torch.int32, device=cuda)
    res = ops.kd_ce_loss(s, t, labels, offs, scale_kd=0.5 / 16, scale_ce=0.5 / 16)
    s32 = s.float().requires_grad_(True)
    ld_ref, ntp_ref = _loss_ref(s32, t, labels, [0, R])
    ((0.5 * ld_ref.sum() + 0.5 * ntp_ref.sum()) / 16).backward()
    scale = torch.tensor([65536.0], device=cuda)
    ds = ops.kd_ce_loss_bwd(s, t, labels, res, loss_scale=scale, out_dtype=torch.float16)
    assert ds.dtype == torch.float16
    assert rel_l2(ds.float() / 65536.0, s32.grad) < 6e-4
    # why the scale exists: the bulk of this gradient (|g| ~ 1e-7) lands on fp16's subnormal grid (step 6e-8) unscaled
    small = s32.grad.abs() < 1e-6
    unscaled = ops.kd_ce_loss_bwd(s, t, labels, res, out_dtype=torch.float16)
    assert rel_l2(unscaled.float()[small], s32.grad[small]) > 2e-2
    assert rel_l2((ds.float() / 65536.0)[small], s32.grad[small]) < 1e-3


def test_grad_scaler_device_state(cuda):
    """training.GradScaler == torch.cuda.amp.GradScaler's state machine, without host synchronisation: an overflowing
    step is skipped (parameters and moments untouched) and halves the scale; clean steps count towards growth; AdamW
    divides the scale out and takes its bias correction from the number of steps actually applied."""
    from llm_speech_summarization_b200.training import FlatAdamW, GradScaler
    p = torch.nn.Parameter(torch.linspace(-1, 1, 1024, device=cuda))
    ref = torch.nn.Parameter(p.detach().clone())
    opt = FlatAdamW([p], lr=1e-2)
    opt.scaler = GradScaler(cuda, init_scale=1024.0, growth_interval=2)
    topt = torch.optim.AdamW([ref], lr=1e-2)
    g = torch.Generator().manual_seed(0)
    for i in range(5):
        grad = torch.randn(1024, generator=g).to(cuda)
        scale = opt.scaler.get_scale()
        opt.grad.copy_(grad * scale)
        if i == 1:
            opt.grad[7] = float("inf")  # overflow: this step must be a no-op
        before = p.detach().clone()
        opt.step()
        if i == 1:
            assert torch.equal(p.detach(), before)
            assert opt.scaler.get_scale() == scale * 0.5
        else:
            ref.grad = grad.clone()
            topt.step()
            assert torch.allclose(p.detach(), ref.detach(), rtol=1e-5, atol=1e-6), i
        opt.zero_grad()
    st = opt.scaler.read()
    assert st["opt_steps"] == 4 and st["skipped_steps"] == 1 and st["found_inf"] == 0
    # 1024 -> (overflow) 512 -> two clean steps -> 1024 -> one more clean step (tracker 1)
    assert st["scale"] == 1024.0 and st["growth_tracker"] == 1
    assert opt.steps_taken() == 4 and opt.state_dict()["state"][0]["step"].item() == 4.0


# ------------------------------------------------------------------- encoder backward: the memory-bound kernels, directly
@pytest.mark.parametrize("C", [256, 512, 1024])
@pytest.mark.parametrize("gelu", [False, True], ids=["ln", "ln_gelu"])
@pytest.mark.parametrize("dt", [torch.float32, torch.float16, torch.bfloat16], ids=["f32", "f16", "bf16"])
def test_layernorm_bwd_ex_vs_autograd(cuda, C, gelu, dt):
    """dh (+)= d/dx [gelu](LN(x)) . dy, dgamma / dbeta, the 16-bit copy of dh, and the fused column sum of dh (the bias
    gradient of the linear layer underneath) against torch autograd on the same (rounded) inputs."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(5 + C)
    rows = 1237  # not a multiple of anything the kernel strides by
    x = (torch.randn(rows, C, generator=g) * 2 + 0.3).to(dt).to(cuda)
    dy = torch.randn(rows, C, generator=g).to(dt).to(cuda)
    gm = (1 + 0.2 * torch.randn(C, generator=g)).to(cuda)
    bt = (0.3 * torch.randn(C, generator=g)).to(cuda)
    prev = torch.randn(rows, C, generator=g).to(cuda)
    xr = x.float().requires_grad_(True)
    gr, br = gm.clone().requires_grad_(True), bt.clone().requires_grad_(True)
    y = F.layer_norm(xr, (C,), gr, br, 1e-5)
    if gelu:
        y = F.gelu(y)
    y.backward(dy.float())
    dh = prev.clone()
    cs = torch.full((C,), 2.0, device=cuda)
    h16 = dt if dt != torch.float32 else torch.float16
    dh, dx, dg, db = ops.layernorm_bwd_ex(x, gm, bt, dy, 1e-5, gelu=gelu, dh=dh, accumulate=True, dx_dtype=h16,
                                          dh_colsum=cs)
    want = prev + xr.grad
    assert rel_l2(dh, want) < 2e-5
    assert rel_l2(dx.float(), want) < (6e-4 if h16 == torch.float16 else 4e-3)
    assert rel_l2(dg, gr.grad) < 2e-5 and rel_l2(db, br.grad) < 2e-5
    assert rel_l2(cs - 2.0, want.sum(0)) < 2e-5
    # 16-bit output only, no accumulation (the conv front end's use)
    _, dx2, _, _ = ops.layernorm_bwd_ex(x, gm, bt, dy, 1e-5, gelu=gelu, dx_dtype=h16)
    assert rel_l2(dx2.float(), xr.grad) < (6e-4 if h16 == torch.float16 else 4e-3)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_conv0_bwd_vs_autograd(cuda, dt):
    """Layer 0 of the conv front end (REF: HubertLayerNormConvLayer via transformers): parameter gradients of
    gelu(LN(conv1d(wave))) recomputed from the waveform, against autograd."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, S = 3, 4000
    frames = (S - 10) // 5 + 1
    wave = torch.randn(B, S, generator=g).to(cuda)
    w = (0.3 * torch.randn(512, 1, 10, generator=g)).to(cuda).requires_grad_(True)
    b = (0.1 * torch.randn(512, generator=g)).to(cuda).requires_grad_(True)
    gm = (1 + 0.1 * torch.randn(512, generator=g)).to(cuda).requires_grad_(True)
    bt = (0.1 * torch.randn(512, generator=g)).to(cuda).requires_grad_(True)
    dy = torch.randn(B, frames, 512, generator=g).to(dt).to(cuda)
    y = F.gelu(F.layer_norm(F.conv1d(wave[:, None], w, b, stride=5).transpose(1, 2), (512,), gm, bt, 1e-5))
    y.backward(dy.float())
    dW, db, dg, dbt = ops.conv0_bwd(wave, w.detach().view(512, 10).contiguous(), b.detach(), gm.detach(), bt.detach(),
                                    1e-5, dy)
    for got, ref in ((dW, w.grad.view(512, 10)), (db, b.grad), (dg, gm.grad), (dbt, bt.grad)):
        assert rel_l2(got, ref) < 5e-5


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_gelu_bwd_with_fused_bias_gradient(cuda, dt):
    """dpre = dy * gelu'(pre) against autograd, plain and with the column sums (the W1 bias gradient) accumulated by the
    same pass; 1237 rows so that the row-striding blocks of the fused form see uneven trip counts."""
    from llm_speech_summarization_b200 import ops
    g = torch.Generator().manual_seed(8)
    rows, Fd = 1237, 4096
    pre = (torch.randn(rows, Fd, generator=g) * 1.5).to(dt).to(cuda)
    dy = torch.randn(rows, Fd, generator=g).to(dt).to(cuda)
    x = pre.float().requires_grad_(True)
    F.gelu(x).backward(dy.float())
    tol = 6e-4 if dt == torch.float16 else 4e-3
    plain = ops.gelu_bwd(pre, dy)
    assert rel_l2(plain.float(), x.grad) < tol
    cs = torch.full((Fd,), 3.0, device=cuda)
    fused = ops.gelu_bwd(pre, dy, colsum=cs)
    assert torch.equal(fused, plain)
    assert rel_l2(cs - 3.0, x.grad.sum(0)) < 1e-4


@pytest.mark.timeout(120)
@pytest.mark.parametrize("P,Hq,Hkv,own", [(9, 24, 8, [191, 108, 191, 108]), (9, 4, 2, [1, 55, 56, 64, 65, 130]),
                                         (64, 2, 2, [100, 7]), (1, 2, 1, [300])])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_attention_shared_prefix_equals_per_sequence_copies(cuda, P, Hq, Hkv, own, dtype):
    """Sequence 0 = the P prefix rows, stored once; sequences 1.. hold only their own rows and attend to the prefix keys in
    front of them. Must equal plain causal attention over sequences that each carry their own copy of the prefix."""
    from llm_speech_summarization_b200 import ops
    D = 128
    g = torch.Generator().manual_seed(P * 1000 + sum(own))
    width = (Hq + 2 * Hkv) * D
    pre = torch.randn(P, width, generator=g)
    owns = [torch.randn(n, width, generator=g) for n in own]
    packed = torch.cat([pre] + owns).to(dtype).to(cuda)
    cu_p = [0, P]
    for n in own:
        cu_p.append(cu_p[-1] + n)
    o = ops.attention(packed, torch.tensor(cu_p, dtype=torch.int32, device=cuda), max(own + [P]), Hq, Hkv, D,
                      1.0 / math.sqrt(D), True, shared_prefix_len=P)
    # reference: every sequence with its own copy of the prefix
    full = torch.cat([torch.cat([pre, x]) for x in owns]).to(dtype).to(cuda)
    cu_f = [0]
    for n in own:
        cu_f.append(cu_f[-1] + P + n)
    ref = _attn_ref(full, cu_f, Hq, Hkv, D, 1.0 / math.sqrt(D), True)
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    assert rel_l2(o[:P].float(), ref[:P]) < tol  # the prefix rows themselves (= the first copy's)
    for i, n in enumerate(own):
        got = o[cu_p[i + 1]:cu_p[i + 2]].float()
        want = ref[cu_f[i] + P:cu_f[i + 1]]
        assert rel_l2(got, want) < tol, (i, n)


@pytest.mark.timeout(120)
def test_attention_shared_prefix_pipelined(cuda, attn_pipelined):
    """The shared-prefix key block through the pipelined issuer (two-block look-ahead cursor across items)."""
    test_attention_shared_prefix_equals_per_sequence_copies(cuda, 9, 24, 8, [191, 108, 191, 108], torch.float16)
    test_attention_shared_prefix_equals_per_sequence_copies(cuda, 9, 4, 2, [1, 55, 56, 64, 65, 130], torch.bfloat16)
