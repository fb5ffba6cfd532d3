"""Per-stage error of the CUDA path vs the CPU oracle at full size (GPU box). Diagnostic only."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import build_product  # noqa: E402
from oracle import configs, reference_math as rm  # noqa: E402
from llm_speech_summarization_b200.step import AudioPromptStep  # noqa: E402
from llm_speech_summarization_b200 import utils as U  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm())


def main():
    dev = torch.device("cuda")
    enc_cfg, llm_cfg = configs.HUBERT_LARGE, configs.LLAMA32_3B
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=1234)
    rs = float(os.environ.get("RESID_SCALE", "1.0"))
    print("resid_scale", rs)
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=4321, dtype=torch.bfloat16, resid_scale=rs)
    tok = configs.stub_tokenizer(llm_cfg)
    audio, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, 160000, T=40, R=64)
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("DTYPE", "fp16")]
    enc_dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[os.environ.get("ENC_DTYPE", os.environ.get("DTYPE", "fp16"))]
    print("operand dtype: llm", dt, "encoder", enc_dt)
    cfg, enc, llm = build_product(enc_cfg, llm_cfg, enc_sd, llm_sd, dev, dtype=dt)
    enc.operand_dtype = enc_dt
    llm32 = {k: v.float() for k, v in llm_sd.items()}
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        t0 = time.time()
        feats = rm.hubert_feature_extractor(enc_sd, audio[None], enc_cfg).transpose(1, 2)
        pre = rm.hubert_encoder(enc_sd, feats, enc_cfg, return_pre_norm=True)
        emb_ref = rm.audio_encoder_forward(enc_sd, audio[None], enc_cfg)
        print(f"oracle encoder {time.time() - t0:.1f}s")
        emb, last = enc.forward_fp32(audio[None].to(dev), return_last_hidden=True)
        print("encoder pre-norm hidden rel err", rel(last, pre), "|h| rms", float(pre.pow(2).mean().sqrt()))
        print("audio embeds rel err", rel(emb, emb_ref), "rms", float(emb_ref.pow(2).mean().sqrt()))

        embed = lambda ids: torch.nn.functional.embedding(ids, llm32["model.embed_tokens.weight"])
        for name, src in (("oracle-embeds", emb_ref), ("cuda-embeds", emb.cpu())):
            a_seq, a_mask, t_seq, t_mask = rm.batch_full_embed_sequence(src, [text_ids], [resp_ids], tok, embed,
                                                                        llm_cfg.llm_type, process_text=True)
            if name == "oracle-embeds":
                t0 = time.time()
                _, s_logits_ref, s_hs = rm.audio_llama_forward(llm32, a_seq, a_mask, None, llm_cfg, True)
                _, t_logits_ref, t_hs = rm.audio_llama_forward(llm32, t_seq, t_mask, None, llm_cfg, True)
                print(f"oracle llm {time.time() - t0:.1f}s")
            out_s = llm(inputs_embeds=a_seq.to(dev), output_hidden_states=True)
            print(f"[student, LLM input = {name}] logits(last 64) rel err",
                  rel(out_s.logits[0, -64:].float(), s_logits_ref[0, -64:]))
            for l in (1, 2, 4, 8, 16, 27, 28):
                print(f"   hidden[{l}] rel err {rel(out_s.hidden_states[l], s_hs[l]):.4e}  rms {float(s_hs[l].pow(2).mean().sqrt()):.3f}")
        out_t = llm(inputs_embeds=t_seq.to(dev), output_hidden_states=True)
        print("[teacher] logits(last 64) rel err", rel(out_t.logits[0, -64:].float(), t_logits_ref[0, -64:]))
        for l in (1, 2, 4, 8, 16, 27, 28):
            print(f"   hidden[{l}] rel err {rel(out_t.hidden_states[l], t_hs[l]):.4e}")
        step = AudioPromptStep(enc, llm, tok, llm_cfg.llm_type)
        o = step.forward_losses(audio[None].to(dev), [text_ids], [resp_ids])
        r = rm.train_step_losses(enc_sd, llm32, enc_cfg, llm_cfg, tok, audio, text_ids, resp_ids)
        for k in ("ntp_loss", "ld_loss", "fd_loss", "total_loss"):
            print(k, "cuda", float(o[k][0]), "oracle", float(r[k]), "rel", abs(float(o[k][0]) - float(r[k])) / abs(float(r[k])))
        # reference modules in plain torch bf16 on the GPU for context (what an eager bf16 run of the oracle gives)
        sd_bf = {k: v.to(dev).to(torch.bfloat16) for k, v in llm_sd.items()}
        _, lg, _ = rm_gpu_forward(sd_bf, t_seq.to(dev).to(torch.bfloat16), llm_cfg)
        print("[teacher] torch-eager bf16 logits rel err", rel(lg[0, -64:].float(), t_logits_ref[0, -64:]))


def rm_gpu_forward(sd, x, cfg):
    import oracle.reference_math as R
    # reference_math builds its masks / rope on CPU; run a device-agnostic copy of the same math
    B, L, H = x.shape
    dev = x.device
    nh, nkv, D = cfg.heads, cfg.kv_heads, cfg.head_dim
    ang = torch.arange(L, dtype=torch.float32)[:, None] * R.rope_inv_freq(cfg)[None, :]
    emb = torch.cat((ang, ang), -1).to(dev)
    cos, sin = emb.cos().to(x.dtype)[None, None], emb.sin().to(x.dtype)[None, None]
    F = torch.nn.functional

    def rms(v, w):
        return w * (v.float() * torch.rsqrt(v.float().pow(2).mean(-1, keepdim=True) + cfg.rms_eps)).to(x.dtype)

    for l in range(cfg.layers):
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
    return None, F.linear(x, sd["lm_head.weight"]), None


if __name__ == "__main__":
    main()
