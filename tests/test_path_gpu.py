"""End-to-end parity of the CUDA path (through the drop-in modules and the C ABI) against
  (1) the golden vectors produced by the REFERENCE's own modules (tests/golden/*.pt, oracle/make_golden.py) and
  (2) the CPU oracle (oracle/reference_math.py) on the same seeded inputs.

Tolerances are the north star's: projected audio embeddings and logits within 2e-2 relative (bf16 vs the
reference's fp32), KD loss within 1e-3 relative.
"""
import os

import pytest
import torch

from conftest import rel_l2
from helpers import GOLDEN, bf16_round_sd, build_product

pytestmark = pytest.mark.gpu

TOL_EMBED = 2e-2
TOL_LOGITS = 2e-2
TOL_KD = 1e-3


def _load_case(name):
    from oracle import configs
    g = torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)
    enc_cfg = configs.EncoderCfg(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in g["enc_cfg"].items()})
    llm_cfg = configs.LlmCfg(**g["llm_cfg"])
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=g["enc_seed"])
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=g["llm_seed"])
    return g, enc_cfg, llm_cfg, enc_sd, llm_sd


@pytest.mark.parametrize("name", ["tiny_llama_hubert", "tiny_minichat_hubert"])
def test_golden_step(cuda, name):
    """Fused step (encoder -> splice -> packed student+teacher prefill -> CE/KD/FD) vs the reference's outputs."""
    from oracle import configs
    from llm_speech_summarization_b200.step import AudioPromptStep
    g, enc_cfg, llm_cfg, enc_sd, llm_sd = _load_case(name)
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    audio, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, g["samples"], T=g["T"], R=g["R"])
    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=g["fd_layers"])
    out = step.forward_losses(audio[None, :].to(cuda), [text_ids], [resp_ids], keep=True)
    plan = out["plan"]
    assert plan.L_audio[0] == g["L_audio"] and plan.L_text[0] == g["L_text"]
    assert out["audio_embeds"].shape[1] >= g["num_audio_embeds"]
    assert rel_l2(out["audio_embeds"][0].cpu(), g["audio_embeds"]) < TOL_EMBED
    assert rel_l2(out["student_logits"].float().cpu(), g["student_logits"]) < TOL_LOGITS
    assert rel_l2(out["teacher_logits"].float().cpu(), g["teacher_logits"]) < TOL_LOGITS
    assert abs(float(out["ld_loss"][0]) - g["ld_loss"]) / abs(g["ld_loss"]) < TOL_KD
    assert abs(float(out["ntp_loss"][0]) - g["ntp_loss"]) / abs(g["ntp_loss"]) < 5e-3
    assert abs(float(out["fd_loss"][0]) - g["fd_loss"]) / abs(g["fd_loss"]) < 3e-2
    total = 0.5 * g["ntp_loss"] + 0.5 * g["ld_loss"] + 1.0 * g["fd_loss"]
    assert abs(float(out["total_loss"][0]) - total) / abs(total) < 1e-2


def test_dropin_modules_match_fused_and_oracle(cuda):
    """The reference-signature calls (AudioEncoder.forward, utils.batch_full_embed_sequence,
    AudioLlamaForCausalLM.forward with labels / hidden states, utils.soft_cross_entropy) reproduce the trainer's
    step (REF/trainer.py:278-370) and agree with the oracle."""
    from oracle import configs, reference_math as rm
    from llm_speech_summarization_b200 import utils as U
    g, enc_cfg, llm_cfg, enc_sd, llm_sd = _load_case("tiny_llama_hubert")
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    audio, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, g["samples"], T=g["T"], R=g["R"])
    R = g["R"]
    with torch.no_grad():
        embeds = enc(audio[None, :].to(cuda))  # (1, A, C) fp16 (the encoder's operand dtype)
        a_seq, a_mask, t_seq, t_mask = U.batch_full_embed_sequence(
            all_audio_embeds=embeds, all_text_input_ids=[text_ids], all_response_input_ids=[resp_ids], tokenizer=tok,
            embed_tokens=llm.model.embed_tokens, llm_type=llm_cfg.llm_type, device=cuda, process_text=True)
        assert a_seq.shape[1] == g["L_audio"] and t_seq.shape[1] == g["L_text"]
        s_out = llm(inputs_embeds=a_seq, labels=[resp_ids.to(cuda)], output_hidden_states=True,
                    attention_mask=a_mask.to(cuda))
        t_out = llm(inputs_embeds=t_seq, labels=[resp_ids.to(cuda)], output_hidden_states=True,
                    attention_mask=t_mask.to(cuda))
        ld = U.soft_cross_entropy(s_out.logits[:, -R:, :], t_out.logits[:, -R:, :])
        fd = sum(torch.nn.functional.mse_loss(s_out.hidden_states[l][:, -R:, :], t_out.hidden_states[l][:, -R:, :])
                 for l in g["fd_layers"])
    assert len(s_out.hidden_states) == llm_cfg.layers + 1
    assert rel_l2(s_out.logits[0, -R:].float().cpu(), g["student_logits"]) < TOL_LOGITS
    assert abs(float(s_out.loss) - g["ntp_loss"]) / abs(g["ntp_loss"]) < 5e-3
    assert abs(float(ld) - g["ld_loss"]) / abs(g["ld_loss"]) < TOL_KD
    assert abs(float(fd) - g["fd_loss"]) / abs(g["fd_loss"]) < 3e-2
    # oracle with the same bf16-rounded LLM weights: tighter agreement on the hidden states
    _, hs = rm.llama_model_forward(bf16_round_sd(llm_sd), a_seq.float().cpu(), a_mask, llm_cfg, output_hidden_states=True)
    for l in (1, llm_cfg.layers):
        assert rel_l2(s_out.hidden_states[l].float().cpu(), hs[l]) < 1e-2


def test_prefill_inference_prompt(cuda):
    """generate_audio_response's prompt assembly + prefill (REF/inference.py:95-135): additional text prompt before
    the audio, merge_prompt_tokens, last-row logits; plus two greedy steps of generate()."""
    from oracle import configs
    from llm_speech_summarization_b200 import utils as U
    g, enc_cfg, llm_cfg, enc_sd, llm_sd = _load_case("tiny_minichat_hubert")
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    audio, _, _ = configs.synthetic_utterance(llm_cfg, 0, g["samples"], T=g["T"], R=g["R"])
    with torch.no_grad():
        embeds = enc(audio[None, :].to(cuda), ctc_pool_ranges=None)
        text = llm.model.embed_tokens(g["extra_ids"][None, :].to(cuda))
        combined = torch.cat([text, embeds], dim=1)
        prompt = U.merge_prompt_tokens(inputs_embeds=combined, tokenizer=tok, embed_tokens=llm.model.embed_tokens,
                                       llm_type=llm_cfg.llm_type, device=cuda)
        logits = llm(inputs_embeds=prompt, num_logits_to_keep=1).logits[0, -1]
        ids = llm.generate(input_ids=None, inputs_embeds=prompt, max_new_tokens=2)
    assert rel_l2(logits.float().cpu(), g["prefill_logits"]) < TOL_LOGITS
    assert ids.shape[0] == 1 and 1 <= ids.shape[1] <= 2
    assert int(ids[0, 0]) == int(g["prefill_logits"].argmax()) or \
        float(g["prefill_logits"].topk(2).values.diff().abs()) < 1e-2  # greedy token unless a near-tie


def test_batched_step_equals_per_utterance(cuda):
    """Packing B utterances into one launch must give each utterance the numbers it gets alone (batch-1 semantics)."""
    from oracle import configs
    from llm_speech_summarization_b200.step import AudioPromptStep
    g, enc_cfg, llm_cfg, enc_sd, llm_sd = _load_case("tiny_llama_hubert")
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=g["fd_layers"])
    utts = [configs.synthetic_utterance(llm_cfg, i, 8000, T=5 + 3 * i, R=4 + 2 * i) for i in range(3)]
    waves = torch.stack([u[0] for u in utts]).to(cuda)
    both = step.forward_losses(waves, [u[1] for u in utts], [u[2] for u in utts])
    for i, (a, t, r) in enumerate(utts):
        one = step.forward_losses(a[None].to(cuda), [t], [r])
        for k in ("ntp_loss", "ld_loss", "fd_loss", "total_loss"):
            assert abs(float(both[k][i]) - float(one[k][0])) <= 2e-3 * abs(float(one[k][0])) + 1e-6, (k, i)


def test_no_cpu_path():
    """The product must fail loudly without CUDA tensors instead of falling back."""
    from oracle import configs
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    from helpers import ns_config
    enc = AudioEncoder(ns_config(configs.TINY_ENCODER, configs.TINY_LLAMA), torch.device("cpu")).eval()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 4000))


@pytest.mark.slow
def test_full_size_vs_oracle(cuda):
    """Full architectures (HuBERT-large + Llama-3.2-3B shapes, random init, 10 s audio): CUDA path vs the CPU
    oracle run on this box's host cores. ~2-4 min of CPU time."""
    from oracle import configs, reference_math as rm
    from llm_speech_summarization_b200.step import AudioPromptStep
    enc_cfg, llm_cfg = configs.HUBERT_LARGE, configs.LLAMA32_3B
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=1234)
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4321, dtype=torch.bfloat16)
    tok = configs.stub_tokenizer(llm_cfg)
    audio, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, 160000, T=40, R=64)
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type)
    out = step.forward_losses(audio[None].to(cuda), [text_ids], [resp_ids], keep=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = rm.train_step_losses(enc_sd, {k: v.float() for k, v in llm_sd.items()}, enc_cfg, llm_cfg, tok, audio,
                                   text_ids, resp_ids, keep=True)
    assert out["plan"].L_audio[0] == 200 and out["plan"].L_text[0] == 117
    assert rel_l2(out["audio_embeds"].cpu(), ref["audio_embeds"]) < TOL_EMBED
    for k in ("ld_loss", "ntp_loss", "fd_loss", "total_loss"):
        assert abs(float(out[k][0]) - float(ref[k])) / abs(float(ref[k])) < TOL_KD, k
    # The north star's bound: logits within 2e-2 of the fp32 reference. Met with fp16 operands -- what the reference
    # itself computes in (REF/trainer.py:57-61,270) -- on BOTH the encoder and the LLM: bf16 operands land at
    # 2.5e-2 / 2.9e-2 on these random-init 28-layer weights (profiles/r01_precision.md, profiles/r02_precision.md),
    # and the reference's own Llama math run in bf16 by PyTorch eager on this GPU is ~4.8e-2 away (printed below).
    err_s = rel_l2(out["student_logits"].float().cpu(), ref["student_logits"][0])
    err_t = rel_l2(out["teacher_logits"].float().cpu(), ref["teacher_logits"][0])
    eager = _eager_bf16_teacher_logits(llm_sd, llm_cfg, tok, text_ids, resp_ids, cuda)
    err_eager = rel_l2(eager.float().cpu(), ref["teacher_logits"][0])
    print(f"full-size logits rel err: student {err_s:.3e} teacher {err_t:.3e} torch-eager-bf16 teacher {err_eager:.3e}")
    assert err_s < TOL_LOGITS and err_t < TOL_LOGITS


def _eager_bf16_teacher_logits(llm_sd, llm_cfg, tok, text_ids, resp_ids, dev):
    """The oracle's Llama math executed in bf16 by PyTorch eager (cuBLAS/SDPA) on the GPU -- the 'library kernel'
    execution of the reference modules, used only as a precision yardstick."""
    import torch.nn.functional as F
    from oracle import reference_math as R
    sd = {k: v.to(dev).to(torch.bfloat16) for k, v in llm_sd.items()}
    embed = lambda ids: F.embedding(ids.to(dev), sd["model.embed_tokens.weight"])
    pre = tok(R.prompt_strings(llm_cfg.llm_type)[0]).input_ids
    suf = tok(R.prompt_strings(llm_cfg.llm_type)[1]).input_ids
    x = R.merge_prompt_response_tokens(pre, suf, embed(text_ids[None]), resp_ids[None], embed)
    B, L, H = x.shape
    nh, nkv, D = llm_cfg.heads, llm_cfg.kv_heads, llm_cfg.head_dim
    ang = torch.arange(L, dtype=torch.float32)[:, None] * R.rope_inv_freq(llm_cfg)[None, :]
    emb = torch.cat((ang, ang), -1).to(dev)
    cos, sin = emb.cos().to(x.dtype)[None, None], emb.sin().to(x.dtype)[None, None]

    def rms(v, w):
        return w * (v.float() * torch.rsqrt(v.float().pow(2).mean(-1, keepdim=True) + llm_cfg.rms_eps)).to(x.dtype)

    for l in range(llm_cfg.layers):
        p = f"model.layers.{l}."
        y = rms(x, sd[p + "input_layernorm.weight"])
        q = F.linear(y, sd[p + "self_attn.q_proj.weight"]).view(B, L, nh, D).transpose(1, 2)
        k = F.linear(y, sd[p + "self_attn.k_proj.weight"]).view(B, L, nkv, D).transpose(1, 2)
        v = F.linear(y, sd[p + "self_attn.v_proj.weight"]).view(B, L, nkv, D).transpose(1, 2)
        q = q * cos + R._rotate_half(q) * sin
        k = k * cos + R._rotate_half(k) * sin
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
        x = x + F.linear(a.transpose(1, 2).reshape(B, L, nh * D), sd[p + "self_attn.o_proj.weight"])
        y = rms(x, sd[p + "post_attention_layernorm.weight"])
        y = F.silu(F.linear(y, sd[p + "mlp.gate_proj.weight"])) * F.linear(y, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(y, sd[p + "mlp.down_proj.weight"])
    x = rms(x, sd["model.norm.weight"])
    R_ = resp_ids.shape[0]
    return F.linear(x[0, -R_:], sd["lm_head.weight"])


def test_whisper_encoder_golden(cuda):
    """AudioEncoder(base="whisper") on the CUDA path vs the reference module's output (tests/golden/tiny_whisper.pt),
    plus the trainer's crop to compute_num_audio_embeds (REF/trainer.py:280-291) and the input-length check."""
    from oracle import configs
    from helpers import ns_config_whisper
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    from llm_speech_summarization_b200.utils import compute_num_audio_embeds
    g = torch.load(os.path.join(GOLDEN, "tiny_whisper.pt"), weights_only=False)
    cfg = configs.WhisperCfg(**g["cfg"])
    enc = AudioEncoder(ns_config_whisper(cfg), cuda)
    enc.load_state_dict(configs.make_whisper_state_dict(cfg, seed=g["seed"]), strict=True)
    enc.eval().to(cuda)
    mel = configs.synthetic_log_mel(cfg, 0, batch=2)
    with torch.no_grad():
        out32 = enc.forward_fp32(mel.to(cuda))
        out = enc(mel.to(cuda), None)
    assert out.shape == g["audio_embeds"].shape and out.dtype == enc.operand_dtype == torch.float16
    assert rel_l2(out32.cpu(), g["audio_embeds"]) < TOL_EMBED
    n = compute_num_audio_embeds(2 * cfg.max_positions * 160, sr=16000)
    assert n == g["num_audio_embeds"] and out[0, :n].shape[0] == n
    with pytest.raises(ValueError, match="Whisper expects the mel input features"):
        enc(mel[:, :, :-2].to(cuda))


@pytest.mark.slow
def test_whisper_medium_full_size_vs_oracle(cuda):
    """Whisper-medium encoder shapes (24 x 1024, 3000 mel frames -> 1500 -> 374 pooled), random init, vs the CPU oracle."""
    from oracle import configs, reference_math as rm
    from helpers import ns_config_whisper
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    cfg = configs.WHISPER_MEDIUM
    sd = configs.make_whisper_state_dict(cfg)
    enc = AudioEncoder(ns_config_whisper(cfg), cuda)
    enc.load_state_dict(sd, strict=True)
    enc.eval().to(cuda)
    mel = configs.synthetic_log_mel(cfg, 0, batch=2)
    with torch.no_grad():
        out = enc.forward_fp32(mel.to(cuda))
        ref = rm.audio_encoder_forward_whisper(sd, mel, cfg)
    assert out.shape == (2, 374, 3072)
    assert rel_l2(out.cpu(), ref) < TOL_EMBED


def test_streaming_submit_equals_blocking_call(cuda):
    """AudioPromptStep.submit(...).result() (side-stream H2D, async D2H into pinned memory, several batches in flight)
    returns exactly what the blocking __call__ returns for each batch."""
    from oracle import configs
    from llm_speech_summarization_b200.step import AudioPromptStep
    enc_cfg, llm_cfg = configs.TINY_ENCODER, configs.TINY_LLAMA
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=1234)
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4321)
    _, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, fd_loss_connector_layers=(0, 1, 2))
    batches = []
    for b in range(4):
        utts = [configs.synthetic_utterance(llm_cfg, 10 * b + i, 8000, T=5 + i, R=4 + b) for i in range(2)]
        batches.append((torch.stack([u[0] for u in utts]).pin_memory(), [u[1] for u in utts], [u[2] for u in utts]))
    blocking = [step(w, t, r, cuda) for (w, t, r) in batches]
    pending = [step.submit(w, t, r, cuda) for (w, t, r) in batches]          # all four in flight
    for ref, p in zip(blocking, reversed(list(reversed(pending)))):
        got = p.result()
        assert set(got) == set(ref)
        for k in ref:
            assert torch.equal(got[k], ref[k]), k


def test_shared_prefix_forward_equals_reference_layout(cuda):
    """forward_losses with the prompt prefix computed once per step (share_prefix=True, the default) against the same
    step in the reference's layout (every sequence carries its own prefix rows): same losses, same logits on the consumed
    rows, up to the rounding of a different attention tiling."""
    from oracle import configs
    from llm_speech_summarization_b200.step import AudioPromptStep
    enc_cfg, llm_cfg = configs.TINY_ENCODER, configs.TINY_LLAMA
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=3)
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4, dtype=torch.bfloat16)
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    tok = configs.stub_tokenizer(llm_cfg)
    utts = [configs.synthetic_utterance(llm_cfg, 40 + i, 16000, T=7 + i, R=5 + i) for i in range(3)]
    waves = torch.stack([u[0] for u in utts]).to(cuda)
    outs = {}
    for share in (True, False):
        step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, share_prefix=share, fd_loss_connector_layers=(0, 1, 2))
        outs[share] = step.forward_losses(waves, [u[1] for u in utts], [u[2] for u in utts], keep=True)
    a, b = outs[True], outs[False]
    assert a["plan"].shared_prefix_len == len(step.prefix) > 0 and b["plan"].shared_prefix_len == 0
    assert a["plan"].rows == b["plan"].rows - (2 * 3 - 1) * len(step.prefix)
    for k in ("ntp_loss", "ld_loss", "fd_loss", "total_loss"):
        assert torch.allclose(a[k], b[k], rtol=2e-3, atol=1e-5), k
    assert rel_l2(a["student_logits"].float(), b["student_logits"].float()) < 4e-3
    assert rel_l2(a["teacher_logits"].float(), b["teacher_logits"].float()) < 4e-3
    # the batched inference prefill shares the prefix the same way
    la, _ = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, share_prefix=True,
                            fd_loss_connector_layers=(0, 1)).prefill_prompts(waves)
    lb, _ = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type, share_prefix=False,
                            fd_loss_connector_layers=(0, 1)).prefill_prompts(waves)
    assert rel_l2(la.float(), lb.float()) < 4e-3
