"""Training-step parity (REF/trainer.py:270-384): gradients of the CUDA path against
autograd through the CPU oracle on the same seeded weights and inputs.

Tolerance: gradients flow through 16-bit GEMMs (dgrad operands are rounded to the model's operand format exactly like
the forward's: fp16 by default, so the gradients carry GradScaler's loss scale until it is divided out), so they are
compared in relative L2 against the fp32 oracle with the bound written at each assert.
"""
import pytest
import torch

from conftest import rel_l2
from helpers import bf16_round_sd, build_product

pytestmark = pytest.mark.gpu

TOL_GRAD = 3e-2


def _tiny(llm_name="llama"):
    from oracle import configs
    enc_cfg = configs.TINY_ENCODER
    llm_cfg = configs.TINY_LLAMA if llm_name == "llama" else configs.TINY_MINICHAT
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=3)
    llm_sd = bf16_round_sd(configs.make_llm_state_dict(llm_cfg, seed=4))
    return configs, enc_cfg, llm_cfg, enc_sd, llm_sd


@pytest.mark.parametrize("llm_name", ["llama", "minichat"])
@pytest.mark.parametrize("use_ld,use_fd", [(True, True), (True, False), (False, False)])
def test_llm_backward_matches_oracle_autograd(cuda, llm_name, use_ld, use_fd):
    """d total_loss / d audio_embeds through the frozen LLM (CE + KD + FD terms)."""
    from oracle import reference_math as rm
    from llm_speech_summarization_b200.step import AudioPromptStep
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny(llm_name)
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    fd_layers = sorted({0, 1, llm_cfg.layers - 1})
    g = torch.Generator().manual_seed(11)
    A = 9
    audio_embeds = (torch.randn(1, A, llm_cfg.hidden, generator=g) * 0.05)
    _, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, 4000, T=7, R=6)

    ae = audio_embeds.clone().requires_grad_(True)
    ref = rm.losses_from_audio_embeds(ae, llm_sd, llm_cfg, tok, text_ids, resp_ids, use_ld=use_ld, use_fd=use_fd,
                                      fd_layers=fd_layers)
    (g_ref,) = torch.autograd.grad(ref["total_loss"], ae)

    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, use_ld_loss=use_ld, use_fd_loss=use_fd,
                           fd_loss_connector_layers=fd_layers)
    out = step.llm_forward_backward(audio_embeds.to(cuda), [text_ids], [resp_ids])
    assert abs(float(out["total_loss"][0]) - float(ref["total_loss"])) / abs(float(ref["total_loss"])) < 1e-2
    err = rel_l2((out["d_audio_embeds"] / out["grad_scale"]).cpu(), g_ref)
    assert err < TOL_GRAD, err


def test_llm_backward_batched_and_scaled(cuda):
    """Two utterances of different lengths in one packed pass == each alone; loss_scale multiplies the gradient."""
    from llm_speech_summarization_b200.step import AudioPromptStep
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=[0, 1, 2])
    g = torch.Generator().manual_seed(5)
    A = 8
    audio = (torch.randn(2, A, llm_cfg.hidden, generator=g) * 0.05).to(cuda)
    _, t0, r0 = configs.synthetic_utterance(llm_cfg, 0, 4000, T=5, R=4)
    _, t1, r1 = configs.synthetic_utterance(llm_cfg, 1, 4000, T=9, R=7)
    both = step.llm_forward_backward(audio, [t0, t1], [r0, r1], loss_scale=0.25)
    one0 = step.llm_forward_backward(audio[0:1], [t0], [r0])
    one1 = step.llm_forward_backward(audio[1:2], [t1], [r1])
    assert rel_l2(both["d_audio_embeds"][0] * 4, one0["d_audio_embeds"][0]) < 5e-3
    assert rel_l2(both["d_audio_embeds"][1] * 4, one1["d_audio_embeds"][0]) < 5e-3
    assert torch.allclose(both["total_loss"], torch.cat([one0["total_loss"], one1["total_loss"]]), rtol=1e-3)


def _encoder_grads_vs_oracle(cuda, B, samples, seed=21):
    """Per-parameter gradients of sum(audio_embeds * R) for a fixed random R: CUDA path vs oracle autograd."""
    from oracle import reference_math as rm
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    g = torch.Generator().manual_seed(seed)
    wave = torch.randn(B, samples, generator=g) * 0.1
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in enc_sd.items()}
    out_ref = rm.audio_encoder_forward(sd, wave, enc_cfg)
    R = torch.randn(out_ref.shape, generator=g)
    names = [k for k, v in sd.items() if torch.is_tensor(v) and v.requires_grad]
    grads = torch.autograd.grad((out_ref * R).sum(), [sd[k] for k in names], allow_unused=True)
    ref = {k: gr for k, gr in zip(names, grads) if gr is not None}

    out = enc.forward_train(wave.to(cuda))
    assert rel_l2(out.cpu(), out_ref.detach()) < 2e-2
    enc.backward(R.to(cuda))
    enc.flush_grads()
    got = {k: p.grad for k, p in enc.named_parameters() if p.grad is not None}
    return ref, got


@pytest.mark.parametrize("B,samples", [(1, 4000), (3, 6000)])
def test_encoder_backward_matches_oracle_autograd(cuda, B, samples):
    ref, got = _encoder_grads_vs_oracle(cuda, B, samples)
    missing = [k for k in ref if k not in got and float(ref[k].norm()) > 0]
    assert not missing, missing
    worst = {}
    for k, gr in ref.items():
        n = float(gr.norm())
        if n < 1e-6:  # e.g. k_proj.bias: softmax is shift-invariant, the true gradient is zero
            assert float(got[k].norm()) < 1e-2
            continue
        worst[k] = rel_l2(got[k].cpu().float(), gr)
    bad = {k: v for k, v in worst.items() if v > 4e-2}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:12]


def test_full_step_gradients_match_oracle_autograd(cuda):
    """Whole training micro-batch: wave -> encoder -> LLM -> CE/KD/FD -> every encoder parameter's gradient."""
    from oracle import reference_math as rm
    from llm_speech_summarization_b200.step import AudioPromptStep
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    fd_layers = [0, 1, 2]
    wave, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, 6000, T=7, R=6)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in enc_sd.items()}
    ae = rm.audio_encoder_forward(sd, wave[None, :], enc_cfg)
    ref = rm.losses_from_audio_embeds(ae, llm_sd, llm_cfg, tok, text_ids, resp_ids, fd_layers=fd_layers)
    names = [k for k, v in sd.items() if torch.is_tensor(v) and v.requires_grad]
    grads = torch.autograd.grad(ref["total_loss"] / 16, [sd[k] for k in names], allow_unused=True)
    gref = {k: g for k, g in zip(names, grads) if g is not None}

    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=fd_layers)
    out = step.forward_backward(wave[None, :].to(cuda), [text_ids], [resp_ids], loss_scale=1.0 / 16)
    enc.flush_grads()
    assert abs(float(out["total_loss"][0]) - float(ref["total_loss"])) / abs(float(ref["total_loss"])) < 1e-2
    S = float(out["grad_scale"])
    got = {k: p.grad.cpu().float() / S for k, p in enc.named_parameters() if p.grad is not None}
    # the loss gradient crosses ~10 16-bit GEMM layers: compare the dominant tensors tightly, all of them loosely
    total_ref = torch.cat([g.reshape(-1) for g in gref.values()])
    total_got = torch.cat([got[k].reshape(-1) for k in gref])
    assert rel_l2(total_got, total_ref) < 3e-2
    for k, g in gref.items():
        if float(g.norm()) > 1e-3 * float(total_ref.norm()):
            assert rel_l2(got[k], g) < 6e-2, k


def test_trainer_accumulates_and_matches_torch_adamw(cuda):
    """4 micro-batches with grad_accum_interval=4, one optimizer step: parameters match torch.optim.AdamW applied to
    the oracle's accumulated gradient; the second window starts from zeroed gradients; checkpoints round-trip."""
    from oracle import reference_math as rm
    from llm_speech_summarization_b200.step import AudioPromptStep
    from llm_speech_summarization_b200.training import EncoderTrainer
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    fd_layers = [0, 1, 2]
    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=fd_layers)
    tr = EncoderTrainer(step, enc, llm, lr=1e-3, grad_accum_interval=4, total_optimizer_steps=10)
    utts = [configs.synthetic_utterance(llm_cfg, i, 5000, T=5 + i, R=4 + i) for i in range(4)]

    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in enc_sd.items()}
    names = [k for k in sd if torch.is_tensor(sd[k]) and sd[k].requires_grad and not k.endswith("masked_spec_embed")]
    opt = torch.optim.AdamW([sd[k] for k in names], lr=1e-3, betas=(0.9, 0.999))
    for wave, t, r in utts:
        ae = rm.audio_encoder_forward(sd, wave[None, :], enc_cfg)
        loss = rm.losses_from_audio_embeds(ae, llm_sd, llm_cfg, tok, t, r, fd_layers=fd_layers)["total_loss"] / 4
        loss.backward()
    opt.step()

    for i, (wave, t, r) in enumerate(utts):
        out = tr.train_step(wave[None, :].to(cuda), [t], [r])
        assert out["optimizer_step"] == (i == 3)
    assert tr.optimizer.step_count == 1 and float(tr.optimizer.grad.abs().max()) == 0.0
    assert abs(tr.lr_scheduler.get_last_lr()[0] - 1e-3 * 0.9) < 1e-12
    got = dict(enc.named_parameters())
    for k in names:
        # first Adam step moves every element by ~lr * sign(g): compare the UPDATE direction where |g| is not tiny
        upd_ref = (sd[k].detach() - enc_sd[k]).reshape(-1)
        upd_got = (got[k].detach().cpu() - enc_sd[k]).reshape(-1)
        if k.endswith("k_proj.bias"):  # true gradient is zero (softmax shift invariance): Adam normalises noise
            continue
        big = sd[k].grad.reshape(-1).abs() > 0.05 * sd[k].grad.abs().max()
        if int(big.sum()) == 0:
            continue
        agree = (torch.sign(upd_ref[big]) == torch.sign(upd_got[big])).float().mean()
        assert float(agree) > 0.98, (k, float(agree))
    # the forward sees the updated weights (packed copies rebuilt)
    o2 = enc.forward_fp32(utts[0][0][None, :].to(cuda))
    with torch.no_grad():
        o2_ref = rm.audio_encoder_forward({k: v.detach() if torch.is_tensor(v) else v for k, v in sd.items()},
                                          utts[0][0][None, :], enc_cfg)
    assert rel_l2(o2.cpu(), o2_ref) < 2e-2
    ck = tr.checkpoint(epoch=0)
    assert set(ck) == {"audio_encoder", "optimizer", "lr_scheduler", "epoch", "step"} and ck["step"] == 4
    torch_opt = torch.optim.AdamW([{"params": list(enc.parameters())}, {"params": list(llm.parameters())}], lr=1e-3)
    torch_opt.load_state_dict(ck["optimizer"])  # layout-compatible with the reference's optimizer
    tr.load_checkpoint(ck)
    assert tr.step == 4 and tr.optimizer.step_count == 1
    # torch-2.0 spelling of the weight-norm parameters (what the reference's pinned torch writes) round-trips too
    legacy = tr.checkpoint(epoch=0, legacy_weight_norm_keys=True)
    assert any(k.endswith("pos_conv_embed.conv.weight_g") for k in legacy["audio_encoder"])
    assert not any("parametrizations" in k for k in legacy["audio_encoder"])
    before = {k: v.clone() for k, v in enc.state_dict().items()}
    tr.load_checkpoint(legacy)
    after = enc.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before)


@pytest.mark.parametrize("B", [1, 2])
def test_whisper_encoder_backward_matches_oracle_autograd(cuda, B):
    """AudioEncoder(base="whisper") training path: per-parameter gradients of sum(audio_embeds * R) vs oracle autograd
    (conv1 / conv2 with their GELUs, the shared pre-LN stack, final LN + pool + projector; frozen sinusoid table)."""
    from oracle import configs, reference_math as rm
    from helpers import ns_config_whisper
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    cfg = configs.TINY_WHISPER
    sd0 = configs.make_whisper_state_dict(cfg, seed=6)
    enc = AudioEncoder(ns_config_whisper(cfg), cuda)
    enc.load_state_dict(sd0, strict=True)
    enc.eval().to(cuda)
    mel = configs.synthetic_log_mel(cfg, 3, batch=B)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "embed_positions" not in k else v)
          for k, v in sd0.items()}
    out_ref = rm.audio_encoder_forward_whisper(sd, mel, cfg)
    R = torch.randn(out_ref.shape, generator=torch.Generator().manual_seed(1))
    names = [k for k, v in sd.items() if torch.is_tensor(v) and v.requires_grad]
    grads = torch.autograd.grad((out_ref * R).sum(), [sd[k] for k in names], allow_unused=True)
    ref = {k: g for k, g in zip(names, grads) if g is not None}
    out = enc.forward_train(mel.to(cuda))
    assert rel_l2(out.cpu(), out_ref.detach()) < 2e-2
    enc.backward(R.to(cuda))
    enc.flush_grads()
    got = {k: p.grad for k, p in enc.named_parameters() if p.grad is not None}
    assert "encoder.embed_positions.weight" not in got
    bad = {}
    for k, gr in ref.items():
        assert k in got, k
        e = rel_l2(got[k].cpu().float(), gr)
        if e > 4e-2:
            bad[k] = e
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:10]


# ----------------------------------------------------------------------------------------------------------
# train-mode regularisers of the HuBERT encoder: dropout sites, LayerDrop,
# SpecAugment. Exact-mask protocol: the kernels regenerate their keep/drop decisions from (seed, site, element), the
# oracle is run under the SAME decisions (oracle/regularizers.py), so outputs and gradients are compared with the
# tolerances of the deterministic step; the golden fixture holds the REFERENCE's own train-mode module outputs.
def test_drop_mask_stream_is_bit_exact_with_the_oracle_generator(cuda):
    import ctypes as C
    import numpy as np
    from oracle import regularizers as rg
    from llm_speech_summarization_b200 import _lib
    lib = _lib.load()
    for seed, site, a, b, p, first in [(1, rg.SITE_FEAT_PROJ, 0, 0, 0.1, 0), (2 ** 61 + 12345, rg.site_ff_act(23), 0, 0, 0.1, 7),
                                       (987654321987, rg.site_attn_prob(5), 31, 15, 0.1, (498 << 16) | 480),
                                       (5, rg.SITE_POS_ADD, 0, 0, 0.5, 2 ** 32 - 5000)]:
        n = 4096
        out = torch.empty(n, dtype=torch.uint8, device=cuda)
        _lib.check(lib.b2s_drop_mask_dump(out.data_ptr(), n, seed, site, a, b, p, first, None), "drop_mask_dump")
        torch.cuda.synchronize()
        e = (np.arange(n, dtype=np.uint64) + np.uint64(first)) & np.uint64(0xFFFFFFFF)
        ref = rg.keep(e, seed, site, a, b, p)
        assert np.array_equal(out.cpu().numpy().astype(bool), ref), (seed, site)


def _train_mode_setup(cuda, gold):
    import dataclasses
    from types import SimpleNamespace as NS
    from oracle import configs
    from helpers import ns_config
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    enc_cfg = dataclasses.replace(configs.TINY_ENCODER, layers=gold["enc_cfg"]["layers"])
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=gold["enc_seed"])
    enc_sd["encoder.masked_spec_embed"] = gold["masked_spec_embed"]
    enc = AudioEncoder(ns_config(enc_cfg, configs.TINY_LLAMA), cuda)
    enc.load_state_dict(enc_sd, strict=True)
    enc.train().to(cuda)
    return enc_cfg, enc_sd, enc


def _draw_from(gold, cuda, cfg=None):
    from llm_speech_summarization_b200.regularizers import RegularizerConfig, RegularizerDraw
    return RegularizerDraw(seed=int(gold["seed"]), layer_skip=gold["layer_skip"].numpy().astype("uint8").copy(),
                           time_mask=gold["time_mask"].reshape(-1).to(torch.uint8).to(cuda),
                           cfg=cfg or RegularizerConfig())


def test_encoder_train_mode_matches_reference_golden_and_oracle(cuda):
    """forward_train / backward under dropout + LayerDrop + SpecAugment vs (a) the reference's own train-mode module
    (golden fixture) and (b) autograd through the oracle under the same masks, for EVERY parameter."""
    import os
    from helpers import GOLDEN
    from oracle import reference_math as rm, regularizers as rg
    gold = torch.load(os.path.join(GOLDEN, "tiny_hubert_train_mode.pt"), weights_only=False)
    enc_cfg, enc_sd, enc = _train_mode_setup(cuda, gold)
    out = enc.forward_train(gold["wave"].to(cuda), draw=_draw_from(gold, cuda))
    err = rel_l2(out.cpu(), gold["audio_embeds"])
    assert err < 2e-2, err                                     # north-star tolerance for the projected embeddings
    enc.backward(gold["R"].to(cuda))
    enc.flush_grads()
    got = {k: p.grad.cpu().float() for k, p in enc.named_parameters() if p.grad is not None}
    for k, g in gold["grads"].items():
        if g is None:
            continue
        if float(g.norm()) == 0.0:                             # parameters of the LayerDrop-skipped layer
            assert k not in got or float(got[k].norm()) == 0.0, k
        else:
            assert rel_l2(got[k], g) < 4e-2, (k, rel_l2(got[k], g))
    # every parameter against the oracle's autograd under the same masks
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in enc_sd.items()}
    reg = rg.OracleRegularizers(seed=int(gold["seed"]), layer_skip=gold["layer_skip"].numpy(),
                                time_mask=gold["time_mask"].numpy())
    o = rm.audio_encoder_forward(sd, gold["wave"], enc_cfg, reg=reg)
    names = [k for k, v in sd.items() if torch.is_tensor(v) and v.requires_grad]
    grads = torch.autograd.grad((o * gold["R"]).sum(), [sd[k] for k in names], allow_unused=True)
    bad = {}
    for k, g in zip(names, grads):
        if g is None or float(g.norm()) < 1e-6:
            assert k not in got or float(got[k].norm()) < 1e-2, k
            continue
        e = rel_l2(got[k], g)
        if e > 4e-2:
            bad[k] = e
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:12]
    assert float(got["encoder.masked_spec_embed"].norm()) > 0


def test_each_regulariser_alone_and_none_equals_eval(cuda):
    """One regulariser at a time (isolates every site's forward AND backward), and the all-off block is bit-identical
    to the deterministic path."""
    import os
    import numpy as np
    from helpers import GOLDEN
    from oracle import reference_math as rm, regularizers as rg
    from llm_speech_summarization_b200.regularizers import RegularizerConfig
    gold = torch.load(os.path.join(GOLDEN, "tiny_hubert_train_mode.pt"), weights_only=False)
    enc_cfg, enc_sd, enc = _train_mode_setup(cuda, gold)
    wave, R = gold["wave"], gold["R"]
    L = enc_cfg.layers
    base = enc.forward_train(wave.to(cuda)).clone()            # no draw, enc.regularizers is None -> deterministic
    off = RegularizerConfig(0.0, 0.0, 0.0, 0.0, 0.0, False)
    d = _draw_from(gold, cuda, off)
    d.layer_skip[:] = 0
    d.time_mask = None
    assert torch.equal(enc.forward_train(wave.to(cuda), draw=d), base)
    cases = {"feat_proj": dict(p_feat_proj=0.3), "hidden": dict(p_hidden=0.3), "attention": dict(p_attention=0.3),
             "activation": dict(p_activation=0.3), "layerdrop": dict(layer_skip=np.array([1] + [0] * (L - 1))),
             "specaug": dict(time_mask=gold["time_mask"].numpy())}
    names = [k for k, v in enc_sd.items() if v.is_floating_point()]
    for label, kw in cases.items():
        ps = dict(p_feat_proj=0.0, p_hidden=0.0, p_attention=0.0, p_activation=0.0)
        ps.update({k: v for k, v in kw.items() if k.startswith("p_")})
        reg = rg.OracleRegularizers(seed=99 + len(label), layer_skip=kw.get("layer_skip"), time_mask=kw.get("time_mask"),
                                    **ps)
        cfg = RegularizerConfig(ps["p_feat_proj"], ps["p_hidden"], ps["p_attention"], ps["p_activation"], 0.0, False)
        d = _draw_from(gold, cuda, cfg)
        d.seed = reg.seed
        d.layer_skip = (np.zeros(L, dtype=np.uint8) if reg.layer_skip is None else reg.layer_skip.astype(np.uint8))
        d.time_mask = None if reg.time_mask is None else torch.from_numpy(reg.time_mask.reshape(-1).astype(np.uint8)).to(cuda)
        for p in enc.parameters():
            p.grad = None
        enc._grads = None
        out = enc.forward_train(wave.to(cuda), draw=d)
        enc.backward(R.to(cuda))
        enc.flush_grads()
        sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in enc_sd.items()}
        o = rm.audio_encoder_forward(sd, wave, enc_cfg, reg=reg)
        assert rel_l2(out.cpu(), o.detach()) < 2e-2, (label, rel_l2(out.cpu(), o.detach()))
        assert rel_l2(o.detach(), base.cpu()) > 0.02, label    # the regulariser did something
        grads = torch.autograd.grad((o * R).sum(), [sd[k] for k in names], allow_unused=True)
        got = dict(enc.named_parameters())
        tot_ref = torch.cat([g.reshape(-1) for g in grads if g is not None])
        tot_got = torch.cat([got[k].grad.cpu().float().reshape(-1) for k, g in zip(names, grads) if g is not None])
        assert rel_l2(tot_got, tot_ref) < 3e-2, (label, rel_l2(tot_got, tot_ref))


def test_trainer_with_regularisers_runs_and_is_reproducible(cuda):
    """EncoderTrainer(regularize=True): train-mode step end to end; same generator seed -> same losses and the same
    parameter update; masked_spec_embed is optimised; a different seed changes the loss."""
    from llm_speech_summarization_b200.step import AudioPromptStep
    from llm_speech_summarization_b200.training import EncoderTrainer
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    utts = [configs.synthetic_utterance(llm_cfg, i, 8000, T=5 + i, R=4 + i) for i in range(2)]

    def run(seed):
        cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
        tok = configs.stub_tokenizer(llm_cfg)
        step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=[0, 1, 2])
        tr = EncoderTrainer(step, enc, llm, lr=1e-3, grad_accum_interval=2, total_optimizer_steps=10, regularize=True,
                            generator=torch.Generator().manual_seed(seed))
        assert enc.training and enc.regularizers is not None
        losses = []
        for i, (wave, t, r) in enumerate(utts):
            out = tr.train_step(wave[None, :].to(cuda), [t], [r])
            losses.append(float(out["total_loss"][0]))
            assert out["optimizer_step"] == (i == 1)
        return losses, {k: v.detach().clone() for k, v in enc.state_dict().items()}

    l1, p1 = run(5)
    l2, p2 = run(5)
    l3, _ = run(6)
    assert l1 == l2   # both micro-batches run before the first update: forward determinism under a fixed seed
    # (the parameter update itself goes through fp32 atomics in the wgrad GEMMs: equal up to summation order)
    big = "encoder.encoder.layers.0.feed_forward.intermediate_dense.weight"
    assert rel_l2(p1[big].float().cpu() - enc_sd[big], p2[big].float().cpu() - enc_sd[big]) < 0.1
    assert l1 != l3 and all(x == x and abs(x) < 1e4 for x in l1 + l3)
    assert not torch.equal(p1["encoder.masked_spec_embed"].cpu(), enc_sd["encoder.masked_spec_embed"])


def test_validate_matches_oracle_perplexities(cuda):
    """EncoderTrainer.validate (REF/trainer.py:400-528): audio / text prompt perplexities vs the oracle's CE on the
    same utterances, greedy generations of both prompts, checkpoint written, train mode restored."""
    import math
    import os
    import tempfile
    from oracle import reference_math as rm
    from llm_speech_summarization_b200.step import AudioPromptStep
    from llm_speech_summarization_b200.training import EncoderTrainer
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=[0, 1, 2])
    tr = EncoderTrainer(step, enc, llm, regularize=True)
    utts = [configs.synthetic_utterance(llm_cfg, i, 6000, T=5 + i, R=4 + i) for i in range(3)]
    a_ref, t_ref = [], []
    with torch.no_grad():
        for wave, t, r in utts:
            o = rm.train_step_losses(enc_sd, llm_sd, enc_cfg, llm_cfg, tok, wave, t, r, fd_layers=[0, 1, 2], keep=True)
            a_ref.append(float(o["ntp_loss"]))
            lt = o["teacher_logits"].reshape(-1, llm_cfg.vocab)          # last R rows of the text-prompt sequence
            t_ref.append(float(torch.nn.functional.cross_entropy(lt[:-1], torch.as_tensor(r)[1:].long())))
    batches = [(torch.stack([utts[0][0], utts[1][0]]).to(cuda), [utts[0][1], utts[1][1]], [utts[0][2], utts[1][2]]),
               (utts[2][0][None].to(cuda), [utts[2][1]], [utts[2][2]])]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "epoch_0_step_0.pt")
        res = tr.validate(batches, epoch=0, num_generate_samples=1, tokenizer=tok, save_path=path)
        ck = torch.load(path, weights_only=False)
    assert set(ck) == {"audio_encoder", "optimizer", "lr_scheduler", "epoch", "step"}
    assert enc.training                                                   # restored (regularize=True put it in train)
    assert abs(res["audio_perplexity"] / math.exp(sum(a_ref) / 3) - 1) < 2e-2
    assert abs(res["text_perplexity"] / math.exp(sum(t_ref) / 3) - 1) < 2e-2
    assert torch.allclose(res["audio_nlls"].cpu(), torch.tensor(a_ref), rtol=1e-2)
    assert torch.allclose(res["text_nlls"].cpu(), torch.tensor(t_ref), rtol=1e-2)
    assert len(res["audio_responses"]) == len(res["text_responses"]) == 1
    n_audio = rm.compute_num_audio_embeds(6000)
    assert 1 <= len(res["audio_responses"][0]) <= 2 * (n_audio + 1)


def test_trainer_dropin_runs_the_reference_loop(cuda, tmp_path):
    """`Trainer(args, config, device).train()` (REF/train.py:25-27, REF/trainer.py:250-398): loop bookkeeping of the
    reference -- `step` counts batches, an optimizer + scheduler step every grad_accum_interval batches and at loader
    end, logging / validation intervals, checkpoint files in <checkpoint_dir>/<run_name>/, resume from a checkpoint."""
    import os
    from types import SimpleNamespace as NS
    from helpers import ns_config
    from llm_speech_summarization_b200.config import llm_arch_from_config
    from llm_speech_summarization_b200.model.audio_llama import AudioLlamaForCausalLM
    from llm_speech_summarization_b200.trainer import NullWriter, Trainer
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    cfg = ns_config(enc_cfg, llm_cfg)
    cfg.train = NS(num_gpus=1, num_workers=0, optimizer=NS(lr=1e-3, beta1=0.9, beta2=0.999), batch_size=1,
                   grad_accum_interval=2, epochs=1, use_ld_loss=True, use_fd_loss=True, ntp_loss_weight=0.5,
                   ld_loss_weight=0.5, fd_loss_weight=1.0, fd_loss_connector_layers=[0, 1, 2])
    cfg.log = NS(checkpoint_dir=str(tmp_path / "ck"), log_dir=str(tmp_path / "logs"), log_interval=1,
                 validation_interval=4, num_generate_samples=1)
    llm = AudioLlamaForCausalLM(llm_arch_from_config(cfg))
    llm.load_state_dict({k: v.to(torch.bfloat16) for k, v in llm_sd.items()}, strict=True)
    tok = configs.stub_tokenizer(llm_cfg)

    def item(i, samples):
        wave, t, r = configs.synthetic_utterance(llm_cfg, i, samples, T=5 + i, R=4 + i)
        bos = torch.tensor([llm_cfg.bos])
        return {"audio": {"array": wave}, "text": f"utt {i}", "text_input_ids": torch.cat([bos, torch.as_tensor(t)]),
                "response_input_ids": torch.cat([bos, torch.as_tensor(r)])[None], "pool_ranges_4": []}

    train_set = [item(i, 5000 + 400 * (i % 2)) for i in range(5)]
    val_set = [item(10 + i, 5000) for i in range(2)]
    args = NS(run_name="run0", checkpoint_path=None, gpu_idx=0)
    w = NullWriter()
    tr = Trainer(args, cfg, cuda, tokenizer=tok, llm=llm, train_dataset=train_set, val_dataset=val_set, writer=w,
                 regularize=True, generator=torch.Generator().manual_seed(0))
    tr.audio_encoder.load_state_dict(enc_sd)
    tr.audio_encoder.mark_weights_changed()
    before = {k: v.detach().clone() for k, v in tr.audio_encoder.state_dict().items()}
    tr.train()
    assert tr.step == 5                                            # one per loader batch
    assert tr.optimizer.step_count == 3                            # batches 2, 4 and the loader end (5)
    assert tr.lr_scheduler.last_epoch == 3
    assert len(w.scalars["train/ntp_loss"]) == 5 and len(w.scalars["learning_rate"]) == 5
    assert [s for s, _ in w.scalars["validation/audio_perplexity"]] == [4, 5]   # interval + end of epoch
    assert all(v == v and v > 1.0 for _, v in w.scalars["validation/text_perplexity"])
    files = sorted(os.listdir(tmp_path / "ck" / "run0"))
    assert files == ["epoch_0_step_4.pt", "epoch_0_step_5.pt"]
    after = tr.audio_encoder.state_dict()
    assert any(not torch.equal(after[k], before[k]) for k in before)
    assert "llm_audio_responses/response_0" in w.texts
    # resume (REF/trainer.py:112-132)
    args2 = NS(run_name="run1", checkpoint_path=str(tmp_path / "ck" / "run0" / "epoch_0_step_5.pt"), gpu_idx=0)
    tr2 = Trainer(args2, cfg, cuda, tokenizer=tok, llm=llm, train_dataset=train_set, val_dataset=val_set,
                  writer=NullWriter(), regularize=False)
    assert tr2.step == 5 and tr2.start_epoch == 0 and tr2.optimizer.step_count == 3
    for k, v in tr2.audio_encoder.state_dict().items():
        assert torch.equal(v.cpu(), after[k].cpu()), k
    # batch_size 2: every loader batch holds two utterances of different lengths -> one ragged micro-batch each
    cfg.train.batch_size = 2
    cfg.train.grad_accum_interval = 1
    tr3 = Trainer(NS(run_name="run2", checkpoint_path=None, gpu_idx=0), cfg, cuda, tokenizer=tok, llm=llm,
                  train_dataset=train_set[:4], val_dataset=val_set, writer=NullWriter(), regularize=False)
    tr3.audio_encoder.load_state_dict(enc_sd)
    tr3.audio_encoder.mark_weights_changed()
    tr3.train()
    assert tr3.step == 2 and tr3.optimizer.step_count == 2


def test_ragged_batch_equals_each_utterance_alone(cuda):
    """Utterances of different lengths in ONE micro-batch (zero-padded waveforms + `lengths`, the reference's collate
    layout): projected embeddings, losses and the accumulated parameter gradients equal running every utterance on
    its own -- the reference's own batch_size > 1 path attends over the padding and does not have this property."""
    from llm_speech_summarization_b200.step import AudioPromptStep
    configs, enc_cfg, llm_cfg, enc_sd, llm_sd = _tiny()
    lens = [6000, 4800, 5610]
    utts = [configs.synthetic_utterance(llm_cfg, i, n, T=5 + i, R=4 + i) for i, n in enumerate(lens)]
    T0 = max(lens)
    padded = torch.zeros(len(lens), T0)
    for i, (w, _, _) in enumerate(utts):
        padded[i, :lens[i]] = w
    t_ids, r_ids = [u[1] for u in utts], [u[2] for u in utts]

    def fresh():
        cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
        tok = configs.stub_tokenizer(llm_cfg)
        return enc, AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=[0, 1, 2])

    enc, step = fresh()
    emb = enc.forward_train(padded.to(cuda), lengths=lens).clone()
    n_valid = [enc.num_audio_embeds(n) for n in lens]
    assert len(set(n_valid)) > 1 and emb.shape[1] == max(n_valid)
    out = step.forward_backward(padded.to(cuda), t_ids, r_ids, loss_scale=0.5, lengths=lens)
    enc.flush_grads()
    g_ragged = {k: p.grad.detach().clone() for k, p in enc.named_parameters() if p.grad is not None}

    enc1, step1 = fresh()
    losses = []
    for i, (w, t, r) in enumerate(utts):
        e1 = enc1.forward_train(w[None].to(cuda))
        assert e1.shape[1] == n_valid[i]
        assert rel_l2(emb[i, :n_valid[i]].cpu(), e1[0].cpu()) < 1e-3, i
        o1 = step1.forward_backward(w[None].to(cuda), [t], [r], loss_scale=0.5)
        losses.append(float(o1["total_loss"][0]))
    enc1.flush_grads()
    assert torch.allclose(out["total_loss"].cpu(), torch.tensor(losses), rtol=1e-3)
    g_single = {k: p.grad for k, p in enc1.named_parameters() if p.grad is not None}
    tot_a = torch.cat([g_ragged[k].reshape(-1) for k in g_single])
    tot_b = torch.cat([g_single[k].reshape(-1) for k in g_single])
    assert rel_l2(tot_a.cpu(), tot_b.cpu()) < 1e-2, rel_l2(tot_a.cpu(), tot_b.cpu())
    for k in g_single:
        if float(g_single[k].norm()) > 1e-2 * float(tot_b.norm()):
            assert rel_l2(g_ragged[k].cpu(), g_single[k].cpu()) < 2e-2, k
    # eval-mode losses of the ragged batch (validation / inference batches) = the per-utterance losses too
    ev = step.forward_losses(padded.to(cuda), t_ids, r_ids, lengths=lens)
    assert torch.allclose(ev["total_loss"].cpu(), torch.tensor(losses), rtol=1e-3)
    # with regularisers on the ragged path still runs and stays finite
    from llm_speech_summarization_b200.regularizers import RegularizerConfig
    enc.regularizers = RegularizerConfig()
    enc.train()
    o2 = step.forward_backward(padded.to(cuda), t_ids, r_ids, loss_scale=0.5, lengths=lens,
                               generator=torch.Generator().manual_seed(0))
    assert bool(torch.isfinite(o2["total_loss"]).all())
