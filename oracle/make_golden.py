"""Pin the oracle: run the REFERENCE's own modules (imported from /root/reference, with the installed
transformers as the backbone) on seeded tiny configurations, check oracle/reference_math.py against them and
write the golden vectors to tests/golden/.

TEST INFRASTRUCTURE; runs only in the build container (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/*.pt and prints the oracle-vs-reference errors
    python -m oracle.make_golden --full     # additionally checks the full-size architectures (slow, ~13 GB RAM)

Shims needed to import the reference offline:
  * transformers.models.llama.modeling_llama.KwargsForCausalLM no longer exists after 4.47 -> alias of
    TransformersKwargs before importing REF/model/audio_llama.py:8;
  * model.audio_encoder.load_hubert_encoder needs the hub -> replaced by HubertModel(HubertConfig(...)) built from
    the explicit architecture; weights are then overwritten from oracle.configs.make_encoder_state_dict;
  * tokenizer -> oracle.configs.StubTokenizer (fixed ids).
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    import transformers
    from transformers.models.llama import modeling_llama
    from transformers.utils import TransformersKwargs
    if not hasattr(modeling_llama, "KwargsForCausalLM"):
        modeling_llama.KwargsForCausalLM = TransformersKwargs
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import model.audio_encoder as ref_audio_encoder  # noqa
    import model.audio_llama as ref_audio_llama  # noqa
    import utils as ref_utils  # noqa
    return ref_audio_encoder, ref_audio_llama, ref_utils


def hubert_config(cfg):
    from transformers import HubertConfig
    return HubertConfig(
        hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
        intermediate_size=cfg.ffn, feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True,
        conv_dim=tuple(cfg.conv_dim), conv_kernel=tuple(cfg.conv_kernel), conv_stride=tuple(cfg.conv_stride),
        num_conv_pos_embeddings=cfg.pos_k, num_conv_pos_embedding_groups=cfg.pos_groups, hidden_act="gelu",
        feat_extract_activation="gelu", layer_norm_eps=cfg.ln_eps, feat_proj_layer_norm=True, vocab_size=32)


def llama_config(cfg):
    from transformers import LlamaConfig
    kw = dict(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.ffn,
              num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads,
              head_dim=cfg.head_dim, max_position_embeddings=cfg.max_pos, rms_norm_eps=cfg.rms_eps,
              tie_word_embeddings=cfg.tie_embeddings, bos_token_id=cfg.bos, attention_bias=False, mlp_bias=False)
    rp = dict(rope_type="default", rope_theta=cfg.rope_theta)
    if cfg.rope_scaling is not None:
        rp = dict(cfg.rope_scaling, rope_theta=cfg.rope_theta)
    kw["rope_parameters"] = rp
    return LlamaConfig(**kw)


def ns(**kw):
    return types.SimpleNamespace(**kw)


def build_reference_models(ref_audio_encoder, ref_audio_llama, enc_cfg, llm_cfg, enc_sd, llm_sd):
    from transformers import HubertModel
    ref_audio_encoder.load_hubert_encoder = lambda config: HubertModel(hubert_config(enc_cfg))
    config = ns(model=ns(audio_encoder=ns(base="hubert", type="facebook/hubert-large-ls960-ft",
                                          downsample_method="pool", downsample_factor=4,
                                          pooling=ns(kernel_size=enc_cfg.pool_kernel, stride=enc_cfg.pool_stride)),
                         llm_type=llm_cfg.llm_type, llm_embedding_channels=enc_cfg.llm_dim))
    enc = ref_audio_encoder.AudioEncoder(config, torch.device("cpu"))
    missing, unexpected = enc.load_state_dict(enc_sd, strict=True), None
    enc.eval()
    llm = ref_audio_llama.AudioLlamaForCausalLM(llama_config(llm_cfg))
    sd = {k: v for k, v in llm_sd.items() if not (llm_cfg.tie_embeddings and k == "lm_head.weight")}
    res = llm.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k == "lm_head.weight" for k in res.missing_keys), res.missing_keys
    if llm_cfg.tie_embeddings:
        llm.tie_weights()
    llm.eval()
    return enc, llm


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def run_case(name, enc_cfg, llm_cfg, samples, T, R, fd_layers, extra_text=0, write=True):
    from oracle import configs, reference_math as rm
    ref_audio_encoder, ref_audio_llama, ref_utils = import_reference()
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=1234)
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4321)
    enc, llm = build_reference_models(ref_audio_encoder, ref_audio_llama, enc_cfg, llm_cfg, enc_sd, llm_sd)
    tok = configs.stub_tokenizer(llm_cfg)
    audio, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, samples, T=T, R=R)
    dev = torch.device("cpu")

    with torch.no_grad():
        # ---- the reference path, restating only the glue of REF/trainer.py:270-374 around its own modules
        ref_embeds = enc(audio[None, :])                                           # REF/trainer.py:278
        a_seq, a_mask, t_seq, t_mask = ref_utils.batch_full_embed_sequence(        # REF/trainer.py:299-313
            all_audio_embeds=ref_embeds, all_text_input_ids=[text_ids], all_response_input_ids=[resp_ids],
            tokenizer=tok, embed_tokens=llm.model.embed_tokens, llm_type=llm_cfg.llm_type, device=dev,
            process_text=True)
        s_out = llm(inputs_embeds=a_seq, labels=[resp_ids], output_hidden_states=True, attention_mask=a_mask)
        t_out = llm(inputs_embeds=t_seq, labels=[resp_ids], output_hidden_states=True, attention_mask=t_mask)
        ntp = s_out.loss
        ld = ref_utils.soft_cross_entropy(s_out.logits[:, -R:, :], t_out.logits[:, -R:, :])
        fd = 0.0
        for l in fd_layers:
            fd = fd + torch.nn.functional.mse_loss(s_out.hidden_states[l][:, -R:, :],
                                                   t_out.hidden_states[l][:, -R:, :])
        n_embeds = ref_utils.compute_num_audio_embeds(samples, sr=16000)
        # ---- inference prompt assembly + prefill (REF/inference.py:113-135)
        extra_ids = None
        combined = ref_embeds
        if extra_text > 0:
            g = torch.Generator().manual_seed(7)
            extra_ids = torch.randint(0, llm_cfg.vocab - 256, (extra_text,), generator=g)
            combined = torch.cat([llm.model.embed_tokens(extra_ids[None, :]), ref_embeds], dim=1)
        prompt = ref_utils.merge_prompt_tokens(inputs_embeds=combined, tokenizer=tok,
                                               embed_tokens=llm.model.embed_tokens, llm_type=llm_cfg.llm_type,
                                               device=dev)
        pre_logits = llm(inputs_embeds=prompt).logits[:, -1, :]

        # ---- the oracle restatement on the same inputs
        o = rm.train_step_losses(enc_sd, llm_sd, enc_cfg, llm_cfg, tok, audio, text_ids, resp_ids,
                                 fd_layers=fd_layers, keep=True)
        o_embeds, o_prompt, o_pre = rm.audio_prompt_prefill(enc_sd, llm_sd, enc_cfg, llm_cfg, tok, audio, extra_ids)

    errs = {
        "audio_embeds": rel(o["audio_embeds"], ref_embeds),
        "student_logits": rel(o["student_logits"], s_out.logits[:, -R:, :]),
        "teacher_logits": rel(o["teacher_logits"], t_out.logits[:, -R:, :]),
        "ntp": abs(float(o["ntp_loss"]) - float(ntp)) / abs(float(ntp)),
        "ld": abs(float(o["ld_loss"]) - float(ld)) / abs(float(ld)),
        "fd": abs(float(o["fd_loss"]) - float(fd)) / max(abs(float(fd)), 1e-30),
        "prompt": rel(o_prompt, prompt),
        "prefill_logits": rel(o_pre, pre_logits),
        "num_audio_embeds": int(rm.compute_num_audio_embeds(samples) != n_embeds),
    }
    print(f"[{name}] oracle vs reference:", {k: f"{v:.2e}" for k, v in errs.items()},
          f"L_audio={a_seq.shape[1]} L_text={t_seq.shape[1]} A={ref_embeds.shape[1]}")
    bad = {k: v for k, v in errs.items() if v > 2e-4}
    assert not bad, f"oracle disagrees with the reference: {bad}"
    assert n_embeds <= ref_embeds.shape[1]

    if write:
        os.makedirs(GOLD, exist_ok=True)
        torch.save({
            "case": name, "samples": samples, "T": T, "R": R, "fd_layers": list(fd_layers), "extra_text": extra_text,
            "enc_cfg": configs.cfg_dict(enc_cfg), "llm_cfg": configs.cfg_dict(llm_cfg),
            "enc_seed": 1234, "llm_seed": 4321,
            # outputs of the REFERENCE modules (fp32)
            "audio_embeds": ref_embeds[0].clone(),
            "student_logits": s_out.logits[0, -R:, :].clone(),
            "teacher_logits": t_out.logits[0, -R:, :].clone(),
            "ntp_loss": float(ntp), "ld_loss": float(ld), "fd_loss": float(fd),
            "L_audio": a_seq.shape[1], "L_text": t_seq.shape[1],
            "num_audio_embeds": n_embeds,
            "prefill_logits": pre_logits[0].clone(),
            "extra_ids": extra_ids,
            "torch": torch.__version__,
        }, os.path.join(GOLD, f"{name}.pt"))
    return errs


def run_whisper_case(name, cfg, write=True):
    """AudioEncoder(base="whisper") of the reference (REF/model/audio_encoder.py:10-13,25-27,56-88) on synthetic
    log-mel features, plus the trainer's crop to compute_num_audio_embeds (REF/trainer.py:280-291)."""
    from transformers import WhisperConfig, WhisperModel
    from oracle import configs, reference_math as rm
    ref_audio_encoder, _, ref_utils = import_reference()
    wc = WhisperConfig(d_model=cfg.hidden, encoder_layers=cfg.layers, encoder_attention_heads=cfg.heads,
                       encoder_ffn_dim=cfg.ffn, num_mel_bins=cfg.mel_bins, max_source_positions=cfg.max_positions,
                       decoder_layers=1, decoder_attention_heads=cfg.heads, decoder_ffn_dim=cfg.ffn,
                       activation_function="gelu", dropout=0.0, encoder_layerdrop=0.0, vocab_size=64,
                       pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    ref_audio_encoder.load_whisper_encoder = lambda config: (WhisperModel(wc).encoder, None)
    config = ns(model=ns(audio_encoder=ns(base="whisper", type="openai/whisper-medium", downsample_method="pool",
                                          downsample_factor=4,
                                          pooling=ns(kernel_size=cfg.pool_kernel, stride=cfg.pool_stride)),
                         llm_type="meta-llama/Llama-3.2-3B-Instruct", llm_embedding_channels=cfg.llm_dim))
    enc = ref_audio_encoder.AudioEncoder(config, torch.device("cpu"))
    sd = configs.make_whisper_state_dict(cfg)
    enc.load_state_dict(sd, strict=True)
    enc.eval()
    mel = configs.synthetic_log_mel(cfg, 0, batch=2)
    with torch.no_grad():
        ref = enc(mel)
        ora = rm.audio_encoder_forward_whisper(sd, mel, cfg)
    # the trainer's un-padding for whisper: 20 ms frames -> samples covered by the mel input
    samples = 2 * cfg.max_positions * 160
    n = ref_utils.compute_num_audio_embeds(samples, sr=16000)
    err = rel(ora, ref)
    print(f"[{name}] oracle vs reference: audio_embeds {err:.2e}; shape {tuple(ref.shape)}, crop to {n}")
    assert err < 2e-4 and n <= ref.shape[1]
    if write:
        os.makedirs(GOLD, exist_ok=True)
        torch.save({"case": name, "cfg": configs.cfg_dict(cfg), "seed": 2468, "audio_embeds": ref.clone(),
                    "num_audio_embeds": n, "torch": torch.__version__}, os.path.join(GOLD, f"{name}.pt"))


def run_train_mode_case(name, enc_cfg, samples, batch, seed, layer_skip, write=True):
    """The reference's AudioEncoder in .train() mode (REF/trainer.py:258) with HF's host / device randomness replaced
    by the masks of oracle/regularizers.py: torch.nn.functional.dropout (every nn.Dropout and the eager attention's
    probability dropout route through it) consumes the sites in HF's call order, torch.rand([]) (LayerDrop) returns
    preset values, _compute_mask_indices (SpecAugment) returns the preset frame mask. Pins the oracle's train-mode
    restatement (forward and autograd gradients) and writes the fixture the CUDA tests compare against."""
    import numpy as np
    from transformers import HubertModel
    from transformers.models.hubert import modeling_hubert
    from oracle import configs, reference_math as rm, regularizers as rg
    from llm_speech_summarization_b200.regularizers import compute_time_mask
    ref_audio_encoder, _, _ = import_reference()
    hc = hubert_config(enc_cfg)
    hc.feat_proj_dropout = hc.hidden_dropout = hc.attention_dropout = hc.activation_dropout = 0.1
    hc.layerdrop = 0.1
    hc.apply_spec_augment, hc.mask_time_prob, hc.mask_time_length, hc.mask_time_min_masks = True, 0.05, 10, 2
    hc._attn_implementation = "eager"
    ref_audio_encoder.load_hubert_encoder = lambda config: HubertModel(hc)
    config = ns(model=ns(audio_encoder=ns(base="hubert", type="facebook/hubert-large-ls960-ft",
                                          downsample_method="pool", downsample_factor=4,
                                          pooling=ns(kernel_size=enc_cfg.pool_kernel, stride=enc_cfg.pool_stride)),
                         llm_type="meta-llama/Llama-3.2-3B-Instruct", llm_embedding_channels=enc_cfg.llm_dim))
    enc = ref_audio_encoder.AudioEncoder(config, torch.device("cpu"))
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=1234)
    g = torch.Generator().manual_seed(77)
    enc_sd["encoder.masked_spec_embed"] = torch.randn(enc_cfg.hidden, generator=g) * 0.5
    enc.load_state_dict(enc_sd, strict=True)
    enc.train()
    wave = torch.randn(batch, samples, generator=g) * 0.1
    frames = samples
    for k, st in zip(enc_cfg.conv_kernel, enc_cfg.conv_stride):
        frames = (frames - k) // st + 1
    time_mask = compute_time_mask(batch, frames, 0.05, 10, 2, np.random.default_rng(5))
    skip = np.asarray(layer_skip, dtype=np.uint8)
    reg = rg.OracleRegularizers(seed=seed, layer_skip=skip, time_mask=time_mask)
    H, nh, L = enc_cfg.hidden, enc_cfg.heads, enc_cfg.layers

    sites = [("elt", rg.SITE_FEAT_PROJ, reg.p_feat_proj), ("elt", rg.SITE_POS_ADD, reg.p_hidden)]
    for l in range(L):
        if not skip[l]:
            sites += [("att", l, reg.p_attention), ("elt", rg.site_attn_out(l), reg.p_hidden),
                      ("elt", rg.site_ff_act(l), reg.p_activation), ("elt", rg.site_ff_out(l), reg.p_hidden)]
    it = iter(sites)

    def fake_dropout(x, p=0.5, training=True, inplace=False):
        kind, site, ps = next(it)
        assert training and abs(p - ps) < 1e-9, (kind, site, p, ps)
        if kind == "att":
            assert x.shape == (batch, nh, frames, frames), x.shape
            return x * rg.attention_multiplier(seed, site, ps, batch, nh, frames)
        assert x.shape[:2] == (batch, frames), x.shape
        return x * rg.elementwise_multiplier(seed, site, ps, batch * frames, x.shape[-1]).view(x.shape)

    rand_vals = iter([0.0 if s else 1.0 for s in skip])  # < layerdrop  <=>  skip
    real_rand = torch.rand

    def fake_rand(*a, **kw):
        if len(a) == 1 and isinstance(a[0], (list, tuple)) and len(a[0]) == 0:
            return torch.tensor(next(rand_vals))
        return real_rand(*a, **kw)

    real = (torch.nn.functional.dropout, modeling_hubert._compute_mask_indices)
    torch.nn.functional.dropout = fake_dropout
    torch.rand = fake_rand
    modeling_hubert._compute_mask_indices = lambda *a, **kw: time_mask
    try:
        ref_out = enc(wave)
    finally:
        torch.nn.functional.dropout, modeling_hubert._compute_mask_indices = real
        torch.rand = real_rand
    assert next(it, None) is None, "the reference consumed fewer dropout sites than expected"
    R = torch.randn(ref_out.shape, generator=g)
    names = ["encoder.masked_spec_embed", "encoder.feature_projection.projection.bias",
             "encoder.encoder.layers.0.attention.q_proj.weight", "encoder.encoder.layers.0.attention.v_proj.bias",
             f"encoder.encoder.layers.{L - 1}.feed_forward.intermediate_dense.weight",
             f"encoder.encoder.layers.{L - 1}.feed_forward.output_dense.bias",
             "encoder.encoder.layer_norm.weight", "encoder.feature_extractor.conv_layers.6.conv.bias",
             "encoder.encoder.pos_conv_embed.conv.bias", "embed_projection.weight"]
    skipped = [l for l in range(L) if skip[l]]
    if skipped:
        names.append(f"encoder.encoder.layers.{skipped[0]}.attention.out_proj.weight")
    ref_params = dict(enc.named_parameters())
    ref_grads = torch.autograd.grad((ref_out * R).sum(), [ref_params[k] for k in names], allow_unused=True)

    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in enc_sd.items()}
    ora_out = rm.audio_encoder_forward(sd, wave, enc_cfg, reg=reg)
    ora_grads = torch.autograd.grad((ora_out * R).sum(), [sd[k] for k in names], allow_unused=True)
    errs = {"audio_embeds": rel(ora_out, ref_out)}
    for k, a, b in zip(names, ora_grads, ref_grads):
        if b is None or float(b.norm()) == 0.0:
            assert a is None or float(a.norm()) == 0.0, k
            errs["grad:" + k] = 0.0
        else:
            errs["grad:" + k] = rel(a, b)
    with torch.no_grad():
        eval_out = rm.audio_encoder_forward(enc_sd, wave, enc_cfg)
    print(f"[{name}] train-mode oracle vs reference:", {k: f"{v:.2e}" for k, v in errs.items()},
          f"frames={frames} masked={int(time_mask.sum())} skip={skip.tolist()} "
          f"train-vs-eval={rel(ref_out.detach(), eval_out):.3f}")
    bad = {k: v for k, v in errs.items() if v > 2e-4}
    assert not bad, f"train-mode oracle disagrees with the reference: {bad}"
    if write:
        torch.save({"case": name, "enc_cfg": configs.cfg_dict(enc_cfg), "enc_seed": 1234, "samples": samples,
                    "batch": batch, "seed": seed, "layer_skip": torch.from_numpy(skip.copy()),
                    "time_mask": torch.from_numpy(time_mask.copy()), "wave": wave.clone(), "R": R.clone(),
                    "masked_spec_embed": enc_sd["encoder.masked_spec_embed"].clone(),
                    "audio_embeds": ref_out.detach().clone(),
                    "grads": {k: (None if gr is None else gr.detach().clone()) for k, gr in zip(names, ref_grads)},
                    "torch": torch.__version__}, os.path.join(GOLD, f"{name}.pt"))
    return errs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also check the full-size architectures (no fixture written)")
    args = ap.parse_args()
    from oracle import configs
    torch.manual_seed(0)
    run_case("tiny_llama_hubert", configs.TINY_ENCODER, configs.TINY_LLAMA, samples=16000, T=12, R=9,
             fd_layers=(0, 1, 2))
    run_case("tiny_minichat_hubert", configs.TINY_ENCODER, configs.TINY_MINICHAT, samples=8000, T=5, R=4,
             fd_layers=(0, 1), extra_text=3)
    run_whisper_case("tiny_whisper", configs.TINY_WHISPER)
    import dataclasses
    run_train_mode_case("tiny_hubert_train_mode", dataclasses.replace(configs.TINY_ENCODER, layers=3), samples=8000,
                        batch=2, seed=0x1234_5678_9ABC_DEF0 >> 2, layer_skip=(0, 1, 0))
    if args.full:
        run_whisper_case("full_whisper_medium", configs.WHISPER_MEDIUM, write=False)
        run_case("full_llama32_hubert", configs.HUBERT_LARGE, configs.LLAMA32_3B, samples=160000, T=40, R=64,
                 fd_layers=(0, 5, 11, 17, 23), write=False)


if __name__ == "__main__":
    main()
