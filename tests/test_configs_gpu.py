"""Full-size parity cases for the remaining BASELINE.json configs (the bench workload, configs[1]/[2] forward, is
covered by tests/test_path_gpu.py::test_full_size_vs_oracle):

  configs[0] / configs[3]: MiniChat-3B + HuBERT-large, interleaved text + speech prompt (additional_text_prompt),
                           30 s utterance, inference prefill (REF/inference.py:95-135);
  configs[4]:              Llama-3.2-3B + Whisper-medium, 30 s log-mel input, pooled embeddings cropped to
                           compute_num_audio_embeds like the trainer does (REF/trainer.py:280-291), prefill.

Random-init weights of the named architectures; the checker is the CPU oracle on this box's host cores.
Tolerances are the north star's, as in test_path_gpu.py: embeddings and last-row logits within 2e-2 of the fp32 oracle
(fp16 operands on the encoder and the LLM, like the reference's autocast; bf16 operands measured 2.7e-2 in round 1).
"""
import pytest
import torch

from conftest import rel_l2
from helpers import build_product, ns_config_whisper

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

TOL_EMBED = 2e-2
TOL_LOGITS = 2e-2


def test_config3_minichat_hubert_30s_text_plus_speech(cuda):
    from oracle import configs, reference_math as rm
    from llm_speech_summarization_b200 import utils as U
    enc_cfg, llm_cfg = configs.HUBERT_LARGE, configs.MINICHAT_3B
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=1234)
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4321, dtype=torch.bfloat16)
    tok = configs.stub_tokenizer(llm_cfg)
    audio, _, _ = configs.synthetic_utterance(llm_cfg, 3, 480000)
    extra = torch.randint(0, llm_cfg.vocab - 256, (14,), generator=torch.Generator().manual_seed(7))
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    with torch.no_grad():
        embeds = enc(audio[None].to(cuda))
        assert embeds.shape == (1, 373, 3072)  # 30 s -> N = 1499 -> A = 373
        text = llm.model.embed_tokens(extra[None].to(cuda))
        prompt = U.merge_prompt_tokens(inputs_embeds=torch.cat([text, embeds], dim=1), tokenizer=tok,
                                       embed_tokens=llm.model.embed_tokens, llm_type=llm_cfg.llm_type, device=cuda)
        assert prompt.shape[1] == 6 + 14 + 373 + 7  # P + X + A + (S - 1) = 400
        logits = llm(inputs_embeds=prompt, num_logits_to_keep=1).logits[0, -1].float().cpu()
        emb32 = enc.forward_fp32(audio[None].to(cuda)).cpu()
        ref_emb, ref_prompt, ref_logits = rm.audio_prompt_prefill(
            enc_sd, {k: v.float() for k, v in llm_sd.items()}, enc_cfg, llm_cfg, tok, audio, extra)
    assert rel_l2(emb32, ref_emb) < TOL_EMBED
    assert rel_l2(prompt.float().cpu(), ref_prompt) < TOL_EMBED
    err = rel_l2(logits, ref_logits[0])
    print(f"config[3] last-row logits rel err {err:.3e}")
    assert err < TOL_LOGITS


def test_config4_llama_whisper_30s(cuda):
    from oracle import configs, reference_math as rm
    from llm_speech_summarization_b200 import utils as U
    from llm_speech_summarization_b200.config import llm_arch_from_config
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    from llm_speech_summarization_b200.model.audio_llama import AudioLlamaForCausalLM
    wcfg, llm_cfg = configs.WHISPER_MEDIUM, configs.LLAMA32_3B
    wsd = configs.make_whisper_state_dict(wcfg)
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4321, dtype=torch.bfloat16)
    tok = configs.stub_tokenizer(llm_cfg)
    cfg = ns_config_whisper(wcfg, llm_type=llm_cfg.llm_type)
    enc = AudioEncoder(cfg, cuda)
    enc.load_state_dict(wsd, strict=True)
    enc.eval().to(cuda)
    llm = AudioLlamaForCausalLM(llm_arch_from_config(cfg))
    llm.load_state_dict(llm_sd, strict=True)
    llm.eval().to(cuda)
    mel = configs.synthetic_log_mel(wcfg, 1, batch=1)
    n = U.compute_num_audio_embeds(480000, sr=16000)
    assert n == 373
    llm32 = {k: v.float() for k, v in llm_sd.items()}
    embed = lambda ids: torch.nn.functional.embedding(ids, llm32["model.embed_tokens.weight"])
    with torch.no_grad():
        emb32 = enc.forward_fp32(mel.to(cuda))
        assert emb32.shape == (1, 374, 3072)
        cropped = enc(mel.to(cuda))[:, :n, :]  # the trainer's un-padding for whisper
        prompt = U.merge_prompt_tokens(inputs_embeds=cropped, tokenizer=tok, embed_tokens=llm.model.embed_tokens,
                                       llm_type=llm_cfg.llm_type, device=cuda)
        assert prompt.shape[1] == 9 + 373 + 5  # 387
        logits = llm(inputs_embeds=prompt, num_logits_to_keep=1).logits[0, -1].float().cpu()
        ref_emb = rm.audio_encoder_forward_whisper(wsd, mel, wcfg)
        ref_prompt = rm.merge_prompt_tokens(ref_emb[:, :n, :], tok, embed, llm_cfg.llm_type)
        _, ref_logits, _ = rm.audio_llama_forward(llm32, ref_prompt, None, None, llm_cfg, num_logits_to_keep=1)
    assert rel_l2(emb32.cpu(), ref_emb) < TOL_EMBED
    err = rel_l2(logits, ref_logits[0, -1])
    print(f"config[4] last-row logits rel err {err:.3e}")
    assert err < TOL_LOGITS
