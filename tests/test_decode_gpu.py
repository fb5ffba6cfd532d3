"""KV-cache greedy decode (REF/inference.py:55-74 -> HF generate with past_key_values):
every decode step must reproduce the logits a full prefill over the grown sequence gives, and the oracle's fp32 logits.
Tolerance: 2e-2 relative L2 on logits (the north star's bf16-vs-fp32 bound), 1e-2 between the two CUDA paths."""
import pytest
import torch

from conftest import rel_l2
from helpers import bf16_round_sd, build_product

pytestmark = pytest.mark.gpu


def _setup(cuda, llm_name="llama"):
    from oracle import configs
    enc_cfg = configs.TINY_ENCODER
    llm_cfg = configs.TINY_LLAMA if llm_name == "llama" else configs.TINY_MINICHAT
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=3)
    llm_sd = bf16_round_sd(configs.make_llm_state_dict(llm_cfg, seed=4))
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, cuda)
    return configs, llm_cfg, llm_sd, llm


@pytest.mark.parametrize("lens", [[11, 5], [7], [7, 9, 4, 6, 5]], ids=["b2_gemv", "b1_gemv", "b5_gemm"])
@pytest.mark.parametrize("llm_name", ["llama", "minichat"])
def test_decode_steps_match_full_prefill_and_oracle(cuda, llm_name, lens):
    from oracle import reference_math as rm
    configs, llm_cfg, llm_sd, llm = _setup(cuda, llm_name)
    g = torch.Generator().manual_seed(7)
    nb = len(lens)
    prompts = [(torch.randn(L, llm_cfg.hidden, generator=g) * 0.05).to(torch.bfloat16).float() for L in lens]
    new_tokens = torch.randint(0, llm_cfg.vocab - 256, (nb, 4), generator=g)
    logits, state = llm.prefill_with_cache([p.to(cuda) for p in prompts], max_new_tokens=8)
    table = llm_sd["model.embed_tokens.weight"]
    seqs = [p.clone() for p in prompts]
    for step in range(new_tokens.shape[1] + 1):
        for b in range(nb):
            full = llm(inputs_embeds=seqs[b][None].to(cuda).to(torch.bfloat16), num_logits_to_keep=1).logits[0, -1]
            assert rel_l2(logits[b].float(), full.float()) < 1e-2, (step, b)
            _, ref, _ = rm.audio_llama_forward(llm_sd, seqs[b][None].to(torch.bfloat16).float(), None, None, llm_cfg,
                                               num_logits_to_keep=1)
            assert rel_l2(logits[b].float().cpu(), ref[0, -1]) < 2e-2, (step, b)
        if step == new_tokens.shape[1]:
            break
        tok = new_tokens[:, step]
        logits = llm.decode_step(tok.to(cuda), state)  # step 0 eager, step 1 captures a CUDA graph, then replays
        for b in range(nb):
            seqs[b] = torch.cat([seqs[b], table[tok[b]][None].float()], dim=0)


def test_generate_with_cache_matches_reprefill(cuda):
    configs, llm_cfg, llm_sd, llm = _setup(cuda)
    g = torch.Generator().manual_seed(9)
    prompt = (torch.randn(1, 13, llm_cfg.hidden, generator=g) * 0.05).to(cuda)
    a = llm.generate(inputs_embeds=prompt, max_new_tokens=12)
    b = llm.generate(inputs_embeds=prompt, max_new_tokens=12, use_kv_cache=False)
    assert a.shape == b.shape and a.dtype == torch.long
    # greedy argmax over near-flat random-init logits can flip on a bf16 tie: demand the common prefix be long
    same = int((a[0] == b[0]).int().cumprod(0).sum())
    assert same >= 8, (a.tolist(), b.tolist())
    assert llm.generate(inputs_embeds=prompt, max_new_tokens=0).shape == (1, 0)
