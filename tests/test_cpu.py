"""CPU suite (`-m "not gpu"`): the oracle against the golden vectors produced by the reference, the host logic
(index plans, config, checkpoint layout), the C-ABI library's symbol table, and the world-size-2 gloo path."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import rel_l2
from helpers import GOLDEN, ROOT, ns_config, ns_config_whisper


# --------------------------------------------------------------------------------------- oracle vs golden
def _load_case(name):
    from oracle import configs
    g = torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)
    enc_cfg = configs.EncoderCfg(**g["enc_cfg"])
    llm_cfg = configs.LlmCfg(**g["llm_cfg"])
    return g, enc_cfg, llm_cfg


@pytest.mark.parametrize("name", ["tiny_llama_hubert", "tiny_minichat_hubert"])
def test_oracle_matches_reference_golden(name):
    """The restatement must reproduce what the reference's own modules produced (oracle/make_golden.py)."""
    from oracle import configs, reference_math as rm
    g, enc_cfg, llm_cfg = _load_case(name)
    enc_sd = configs.make_encoder_state_dict(enc_cfg, seed=g["enc_seed"])
    llm_sd = configs.make_llm_state_dict(llm_cfg, seed=g["llm_seed"])
    tok = configs.stub_tokenizer(llm_cfg)
    audio, text_ids, resp_ids = configs.synthetic_utterance(llm_cfg, 0, g["samples"], T=g["T"], R=g["R"])
    with torch.no_grad():
        o = rm.train_step_losses(enc_sd, llm_sd, enc_cfg, llm_cfg, tok, audio, text_ids, resp_ids,
                                 fd_layers=g["fd_layers"], keep=True)
        _, _, pre = rm.audio_prompt_prefill(enc_sd, llm_sd, enc_cfg, llm_cfg, tok, audio, g["extra_ids"])
    assert o["L_audio"] == g["L_audio"] and o["L_text"] == g["L_text"]
    assert rel_l2(o["audio_embeds"][0], g["audio_embeds"]) < 1e-4
    assert rel_l2(o["student_logits"][0], g["student_logits"]) < 1e-4
    assert rel_l2(o["teacher_logits"][0], g["teacher_logits"]) < 1e-4
    assert rel_l2(pre[0], g["prefill_logits"]) < 1e-4
    for k in ("ntp_loss", "ld_loss", "fd_loss"):
        assert abs(float(o[k]) - g[k]) <= 1e-5 * abs(g[k]) + 1e-7
    assert rm.compute_num_audio_embeds(g["samples"]) == g["num_audio_embeds"]


def test_oracle_identities():
    """D1/D2 (single-pass KD and CE from running statistics) and D5 (pool <-> projector)."""
    from oracle import reference_math as rm
    g = torch.Generator().manual_seed(0)
    s = torch.randn(7, 1000, generator=g, dtype=torch.float64) * 3
    t = torch.randn(7, 1000, generator=g, dtype=torch.float64) * 3
    ref = rm.soft_cross_entropy(s, t)
    lse = torch.logsumexp(s, -1)
    dot = (torch.softmax(t, -1) * s).sum(-1)
    assert abs(float((lse - dot).mean()) - float(ref)) < 1e-12
    h = torch.randn(1, 499, 64, generator=g, dtype=torch.float64)
    w = torch.randn(32, 64, generator=g, dtype=torch.float64)
    pool = lambda x: torch.nn.functional.avg_pool1d(x.transpose(1, 2), 8, 4).transpose(1, 2)
    assert rel_l2(pool(h) @ w.t(), pool(h @ w.t())) < 1e-12
    assert pool(h).shape[1] == 123


def test_compute_num_audio_embeds_matches_reference_values():
    """REF/utils.py:13-24 on the benchmark lengths (160 000 -> 123, 480 000 -> 373)."""
    from llm_speech_summarization_b200.utils import compute_num_audio_embeds
    from oracle.reference_math import compute_num_audio_embeds as ref
    for n, want in ((160000, 123), (480000, 373)):
        assert compute_num_audio_embeds(n) == want == ref(n)
    for n in (400, 8000, 16000, 123457):
        assert compute_num_audio_embeds(n, sr=16000) == ref(n, sr=16000)


# --------------------------------------------------------------------------------------- host logic
def test_build_plan_matches_reference_layout():
    """Packed index plan == the reference's sequence construction (double BOS strip, label alignment, lengths)."""
    from oracle import configs, reference_math as rm
    from llm_speech_summarization_b200.step import build_plan
    llm_cfg = configs.TINY_LLAMA
    tok = configs.stub_tokenizer(llm_cfg)
    prefix, suffix = tok.prefix_ids, tok.suffix_ids
    utts = [configs.synthetic_utterance(llm_cfg, i, 16, T=4 + i, R=3 + 2 * i) for i in range(3)]
    A = 5
    plan = build_plan(prefix, suffix, A, [u[1].tolist() for u in utts], [u[2].tolist() for u in utts])
    table = torch.arange(llm_cfg.vocab, dtype=torch.float32)[:, None].repeat(1, 2)  # embedding = token id
    embed = lambda ids: torch.nn.functional.embedding(ids, table)
    audio = [-(torch.arange(i * A, (i + 1) * A, dtype=torch.float32) + 1)[:, None].repeat(1, 2) for i in range(3)]
    a_seq, a_mask, t_seq, t_mask = rm.batch_full_embed_sequence(audio, [u[1] for u in utts], [u[2] for u in utts],
                                                                tok, embed, llm_cfg.llm_type, process_text=True)
    cu = plan["cu_seqlens"]
    src = torch.tensor(plan["row_src"], dtype=torch.float32)
    for i in range(3):
        La, Lt = int(a_mask[i].sum()), int(t_mask[i].sum())
        assert plan["L_audio"][i] == La and plan["L_text"][i] == Lt
        assert torch.equal(src[cu[i]:cu[i + 1]], a_seq[i, a_seq.shape[1] - La:, 0])
        assert torch.equal(src[cu[3 + i]:cu[3 + i + 1]], t_seq[i, t_seq.shape[1] - Lt:, 0])
        assert plan["positions"][cu[i]:cu[i + 1]] == list(range(La))
        R = len(utts[i][2])
        rows = plan["student_rows"][plan["row_offsets"][i]:plan["row_offsets"][i + 1]]
        assert rows == list(range(cu[i + 1] - R, cu[i + 1]))
        labels = plan["labels"][plan["row_offsets"][i]:plan["row_offsets"][i + 1]]
        assert labels == utts[i][2].tolist()[1:] + [-1]  # logits[-R:-1] vs labels[1:], REF/model/audio_llama.py:84-89
    assert plan["rows"] == cu[-1] and plan["sum_r"] == sum(len(u[2]) for u in utts)
    # P=9, A=123, S=6, R=64, T=40 -> 200 and 117
    p = build_plan(list(range(9)), list(range(6)), 123, [list(range(40))], [list(range(64))])
    assert p["L_audio"] == [200] and p["L_text"] == [117]


def test_build_plan_edge_cases():
    from llm_speech_summarization_b200.step import build_plan
    p = build_plan([1, 2], [1, 3], 2, [[]], [[9]])  # empty transcript, 1-token response (no CE rows)
    assert p["L_audio"] == [2 + 2 + 1 + 0] and p["L_text"] == [2 + 0 + 1 + 0]
    assert p["labels"] == [-1] and p["sum_r"] == 1
    p = build_plan([1], [1], 0, [[5, 6]], [[7, 8, 9]], with_teacher=False)  # no audio rows, no teacher
    assert p["teacher_rows"] == [] and p["L_text"] == [] and p["L_audio"] == [1 + 0 + 0 + 2]
    with pytest.raises(ValueError):
        build_plan([], [1], 0, [[]], [[1, 2, 3, 4]])  # response longer than the sequence it must end


def test_state_dict_layout_matches_reference_checkpoint():
    """AudioEncoder.state_dict() == the reference's checkpoint layout (424 tensors for
    HuBERT-large + pool), both weight-norm spellings load, and the trainer / inference checkpoint dicts round-trip."""
    from oracle import configs
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    cfg = ns_config(configs.TINY_ENCODER, configs.TINY_LLAMA)
    enc = AudioEncoder(cfg, torch.device("cpu"))
    sd = configs.make_encoder_state_dict(configs.TINY_ENCODER)
    assert set(enc.state_dict().keys()) == set(sd.keys())
    for k, v in enc.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    enc.load_state_dict(sd, strict=True)
    legacy = dict(sd)
    pc = "encoder.encoder.pos_conv_embed.conv."
    legacy[pc + "weight_g"] = legacy.pop(pc + "parametrizations.weight.original0")
    legacy[pc + "weight_v"] = legacy.pop(pc + "parametrizations.weight.original1")
    enc2 = AudioEncoder(cfg, torch.device("cpu"))
    enc2.load_state_dict(legacy, strict=True)
    assert torch.equal(enc2.state_dict()[pc + "parametrizations.weight.original1"], sd[pc + "parametrizations.weight.original1"])
    # full-size key count
    from llm_speech_summarization_b200.config import EncoderArch
    from llm_speech_summarization_b200.model.audio_encoder import HubertBackbone
    n = len(HubertBackbone(EncoderArch(layers=24)).state_dict()) + 2  # + embed_projection.{weight,bias}
    assert n == 424


def test_config_loader_reads_reference_yaml_keys(tmp_path):
    from llm_speech_summarization_b200.config import load_config, llm_arch_from_config, encoder_arch_from_config
    y = tmp_path / "c.yaml"
    y.write_text("""
seed_everything: 1234
model:
  audio_encoder:
    base: hubert
    type: facebook/hubert-large-ls960-ft
    downsample_method: pool
    downsample_factor: 4
    pooling:
      kernel_size: 8
      stride: 4
  llm_type: "meta-llama/Llama-3.2-3B-Instruct"
  llm_embedding_channels: 3072
audio:
  sampling_rate: 16000
train:
  grad_accum_interval: 16
  use_ld_loss: True
  fd_loss_connector_layers: [0, 5, 11, 17, 23]
""")
    c = load_config(str(y))
    assert c.model.audio_encoder.pooling.kernel_size == 8 and c.train.fd_loss_connector_layers == [0, 5, 11, 17, 23]
    a = llm_arch_from_config(c)
    assert (a.layers, a.heads, a.kv_heads, a.vocab) == (28, 24, 8, 128256)
    assert encoder_arch_from_config(c).hidden == 1024
    c.model.llm_type = "nope"
    with pytest.raises(Exception, match="Unknown LLM type."):
        llm_arch_from_config(c)


def test_reference_error_texts():
    from oracle import configs
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    from llm_speech_summarization_b200 import utils as U
    cfg = ns_config(configs.TINY_ENCODER, configs.TINY_LLAMA)
    cfg.model.audio_encoder.base = "wav2vec"
    with pytest.raises(Exception, match="Unexpected encoder type in config."):
        AudioEncoder(cfg, torch.device("cpu"))
    with pytest.raises(Exception, match="Unknown LLM type."):
        U.merge_prompt_tokens(None, None, None, "gpt2", "cpu")


# --------------------------------------------------------------------------------------- C ABI
def test_abi_exports_every_declared_symbol():
    """libb2s.so loads without a GPU and exports exactly the functions include/b2s.h declares."""
    from llm_speech_summarization_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "b2s.h")).read()
    declared = set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b2s.h but not exported"
    assert declared == set(_lib.PROTOTYPES.keys())
    assert lib.b2s_version() == 1
    assert lib.b2s_kd_ce_workspace_bytes(64, 128256) == 64 * 16 * 24  # rows x max slices x sizeof(Partial)


def test_abi_fails_loudly_without_gpu():
    """No CPU fallback: a compute call on a box without a CUDA device returns an error status + message."""
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from llm_speech_summarization_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_float * 1024)()
    rc = lib.b2s_cast_f32_to_h16(ctypes.addressof(buf), ctypes.addressof(buf), 1024, 1, None)
    assert rc != 0 and len(lib.b2s_last_error()) > 0


# --------------------------------------------------------------------------------------- world-size-2 (gloo)
_DIST_SCRIPT = r"""
import os, sys, torch
sys.path.insert(0, os.environ["B2S_ROOT"])
from llm_speech_summarization_b200 import dp
rank, local_rank, world = dp.init_process_group("gloo")
assert world == 2
# sharding: disjoint, complete, balanced
mine = dp.shard_indices(17, rank, world)
cnt = torch.zeros(17); cnt[mine] = 1
torch.distributed.all_reduce(cnt)
assert torch.equal(cnt, torch.ones(17)) and abs(len(mine) - 17 / 2) <= 0.5
# summed-gradient invariant: 16 utterances split over 2 ranks == 1 rank doing all 16 (REF/trainer.py:372-384)
torch.manual_seed(0)
w = torch.randn(5, 3, dtype=torch.float64)
xs = torch.randn(16, 3, dtype=torch.float64)
def grads(idx):
    g1, g2 = torch.zeros(5, 3, dtype=torch.float64), torch.zeros(7, dtype=torch.float32)
    for i in idx:
        wi = w.clone().requires_grad_(True)
        (torch.tanh(wi @ xs[i]).sum() * dp.local_accum_scale(16)).backward()
        g1 += wi.grad; g2 += float(i)
    return [g1, g2]
local = grads(dp.shard_indices(16, rank, world))
dp.allreduce_sum_(local, bucket_bytes=64)
full = grads(range(16))
assert torch.allclose(local[0], full[0], atol=1e-12) and torch.allclose(local[1], full[1])
assert dp.max_over_ranks(float(rank), "cpu") == 1.0 and dp.sum_over_ranks(1.0, "cpu") == 2.0
dp.barrier()
print("OK", rank)
"""


def test_data_parallel_world_size_2_gloo(tmp_path):
    script = tmp_path / "dist_check.py"
    script.write_text(_DIST_SCRIPT)
    env = dict(os.environ, B2S_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2


# --------------------------------------------------------------------------------------- whisper variant
def test_whisper_oracle_matches_reference_golden():
    from oracle import configs, reference_math as rm
    g = torch.load(os.path.join(GOLDEN, "tiny_whisper.pt"), weights_only=False)
    cfg = configs.WhisperCfg(**g["cfg"])
    sd = configs.make_whisper_state_dict(cfg, seed=g["seed"])
    with torch.no_grad():
        out = rm.audio_encoder_forward_whisper(sd, configs.synthetic_log_mel(cfg, 0, batch=2), cfg)
    assert out.shape == g["audio_embeds"].shape
    assert rel_l2(out, g["audio_embeds"]) < 1e-4
    assert rm.compute_num_audio_embeds(2 * cfg.max_positions * 160) == g["num_audio_embeds"]
    assert rm.compute_num_audio_embeds(480000) == 373  # 30 s -> pooled 374 cropped to 373


def test_whisper_state_dict_layout():
    from oracle import configs
    from helpers import ns_config_whisper
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    cfg = configs.TINY_WHISPER
    enc = AudioEncoder(ns_config_whisper(cfg), torch.device("cpu"))
    sd = configs.make_whisper_state_dict(cfg)
    assert set(enc.state_dict().keys()) == set(sd.keys())
    enc.load_state_dict(sd, strict=True)
    assert "encoder.layers.0.self_attn.k_proj.bias" not in enc.state_dict()


# --------------------------------------------------------------------------------------- training host logic
def test_polynomial_lr_matches_torch():
    """PolynomialLR(power=1.0) of REF/trainer.py:106-110 in closed form == torch's scheduler, step by step."""
    from llm_speech_summarization_b200.training import PolynomialLR

    class _Opt:
        defaults = {"lr": 5e-5}
        lr = 5e-5

    ours = PolynomialLR(_Opt(), total_iters=7, power=1.0)
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.AdamW([p], lr=5e-5)
    theirs = torch.optim.lr_scheduler.PolynomialLR(topt, total_iters=7, power=1.0)
    for _ in range(10):
        assert abs(ours.get_last_lr()[0] - theirs.get_last_lr()[0]) < 1e-12
        topt.step()
        theirs.step()
        ours.step()
    sd = ours.state_dict()
    again = PolynomialLR(_Opt(), total_iters=3)
    again.load_state_dict(sd)
    assert again.get_last_lr() == ours.get_last_lr() and again.last_epoch == ours.last_epoch


def test_flat_param_order_and_grad_spec_cover_every_parameter():
    """The flat optimizer layout lists every parameter once with q|k|v weights (and biases) of a layer adjacent, and
    the packed gradient buffers tile exactly the parameters the reference optimises (REF/trainer.py:98-105)."""
    from oracle import configs
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    enc = AudioEncoder(ns_config(configs.TINY_ENCODER, configs.TINY_LLAMA), "cpu")
    params = list(enc.parameters())
    order = enc.flat_param_order()
    assert len(order) == len(params) and {id(p) for p in order} == {id(p) for p in params}
    pos = {id(p): i for i, p in enumerate(order)}
    for lay in enc.encoder.encoder.layers:
        a = lay.attention
        assert pos[id(a.k_proj.weight)] == pos[id(a.q_proj.weight)] + 1 == pos[id(a.v_proj.weight)] - 1
        assert pos[id(a.k_proj.bias)] == pos[id(a.q_proj.bias)] + 1 == pos[id(a.v_proj.bias)] - 1
    covered = {}
    for name, shape, ps in enc._grad_spec():
        assert sum(p.numel() for p in ps) == int(torch.tensor(shape).prod()), name
        for p in ps:
            assert id(p) not in covered, name
            covered[id(p)] = name
    named = dict(enc.named_parameters())
    missing = [n for n, p in named.items() if id(p) not in covered]
    # conv weights (permuted layout), the weight-normed positional conv and the unused SpecAugment embedding go
    # through scratch / get no gradient; everything else accumulates in place
    assert all(("conv.weight" in n and "conv_layers.0" not in n) or "parametrizations" in n or n.endswith(
        "masked_spec_embed") for n in missing), missing


def test_training_path_fails_loudly_without_gpu():
    from oracle import configs
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    from llm_speech_summarization_b200.training import FlatAdamW
    enc = AudioEncoder(ns_config(configs.TINY_ENCODER, configs.TINY_LLAMA), "cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        FlatAdamW(enc.parameters())
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc.forward_train(torch.zeros(1, 4000))


def test_whisper_log_mel_oracle_matches_transformers_extractor():
    """oracle.whisper_log_mel restates WhisperFeatureExtractor (the reference's collate-side dependency,
    REF/trainer.py:178-182); pinned here against the extractor itself on seeded audio."""
    transformers = pytest.importorskip("transformers")
    import numpy as np
    from oracle import reference_math as rm
    fe = transformers.WhisperFeatureExtractor()
    g = torch.Generator().manual_seed(77)
    wave = (torch.randn(48000, generator=g) * 0.1).numpy()
    feats = fe(wave, sampling_rate=16000, return_tensors="np").input_features[0]   # pads to 30 s
    padded = np.zeros(480000, dtype=np.float32)
    padded[:48000] = wave
    ours = rm.whisper_log_mel(padded, fe.mel_filters)
    assert ours.shape == feats.shape == (80, 3000)
    assert float(np.abs(ours - feats).max()) < 2e-4


# ----------------------------------------------------------------------------------------------------------
# train-mode regularisers
def test_train_mode_oracle_matches_reference_golden():
    """The oracle's train-mode HuBERT (dropout sites, LayerDrop, SpecAugment under the explicit masks of
    oracle/regularizers.py) against the fixture produced by the REFERENCE's AudioEncoder in .train() mode with HF's
    randomness replaced by the same masks (oracle/make_golden.py:run_train_mode_case): outputs and gradients."""
    import dataclasses
    import numpy as np
    from oracle import configs, reference_math as rm, regularizers as rg
    gold = torch.load(os.path.join(GOLDEN, "tiny_hubert_train_mode.pt"), weights_only=False)
    cfg = dataclasses.replace(configs.TINY_ENCODER, layers=gold["enc_cfg"]["layers"])
    sd = configs.make_encoder_state_dict(cfg, seed=gold["enc_seed"])
    sd["encoder.masked_spec_embed"] = gold["masked_spec_embed"]
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    reg = rg.OracleRegularizers(seed=gold["seed"], layer_skip=gold["layer_skip"].numpy(),
                                time_mask=gold["time_mask"].numpy())
    out = rm.audio_encoder_forward(sd, gold["wave"], cfg, reg=reg)
    assert rel_l2(out.detach(), gold["audio_embeds"]) < 1e-5
    names = [k for k, g in gold["grads"].items() if g is not None]
    grads = torch.autograd.grad((out * gold["R"]).sum(), [sd[k] for k in names], allow_unused=True)
    for k, g in zip(names, grads):
        ref = gold["grads"][k]
        if float(ref.norm()) == 0.0:  # the LayerDrop-skipped layer
            assert g is None or float(g.norm()) == 0.0, k
        else:
            assert rel_l2(g, ref) < 1e-4, k
    # and the masks really bite: the eval-mode output is far away
    with torch.no_grad():
        ev = rm.audio_encoder_forward({k: (v.detach() if torch.is_tensor(v) else v) for k, v in sd.items()},
                                      gold["wave"], cfg)
    assert rel_l2(ev, gold["audio_embeds"]) > 0.2


def test_mask_generator_statistics():
    """The counter-based generator (oracle restatement of csrc/rng.cuh): keep rate = 1 - p, streams independent."""
    import numpy as np
    from oracle import regularizers as rg
    e = np.arange(1 << 18, dtype=np.uint64)
    for p in (0.05, 0.1, 0.5):
        k = rg.keep(e, 1234567, rg.site_ff_act(3), 0, 0, p)
        assert abs(float(k.mean()) - (1 - p)) < 4e-3
    a = rg.keep(e, 42, rg.site_attn_out(0), 0, 0, 0.5)
    for other in (rg.keep(e, 43, rg.site_attn_out(0), 0, 0, 0.5), rg.keep(e, 42, rg.site_attn_out(1), 0, 0, 0.5),
                  rg.keep(e, 42, rg.site_attn_prob(0), 1, 0, 0.5), rg.keep(e, 42, rg.site_attn_prob(0), 0, 1, 0.5)):
        assert abs(float((a == other).mean()) - 0.5) < 5e-3  # uncorrelated with any other stream
    assert abs(float((a[1:] == a[:-1]).mean()) - 0.5) < 5e-3  # and along the element index
    assert rg.keep(e, 42, 1, 0, 0, 0.0).all() and rg.threshold(0.0) == 0
    m = rg.attention_multiplier(7, 2, 0.1, 1, 2, 64)
    assert set(torch.unique(m).tolist()) == {0.0, float(np.float32(1) / (np.float32(1) - np.float32(0.1)))}


def test_spec_augment_time_mask_follows_hf_sampling():
    """compute_time_mask (host side of SpecAugment) against HF's _compute_mask_indices
    (TF/models/hubert/modeling_hubert.py, the reference's dependency): same span structure and the same mean number
    of masked frames."""
    import numpy as np
    from llm_speech_summarization_b200.regularizers import compute_time_mask
    rng = np.random.default_rng(0)
    ours = np.stack([compute_time_mask(4, 499, 0.05, 10, 2, rng) for _ in range(200)])
    assert ours.shape == (200, 4, 499) and ours.dtype == bool
    per_row = ours.sum(-1)
    assert per_row.min() >= 10 and per_row.max() <= 30      # 2..3 spans of 10 frames, overlaps allowed
    # every masked run is a union of length-10 spans: no run shorter than 10
    for row in ours.reshape(-1, 499)[:200]:
        d = np.diff(np.concatenate([[0], row.astype(np.int8), [0]]))
        runs = np.flatnonzero(d == -1) - np.flatnonzero(d == 1)
        assert runs.min() >= 10
    assert compute_time_mask(2, 24, 0.05, 10, 2, rng).sum(-1).max() <= 20   # capped at frames // mask_length spans
    with pytest.raises(ValueError):
        compute_time_mask(1, 5, 0.05, 10, 2, rng)
    hub = pytest.importorskip("transformers.models.hubert.modeling_hubert")
    np.random.seed(0)
    theirs = np.stack([hub._compute_mask_indices((4, 499), 0.05, 10, min_masks=2) for _ in range(200)])
    assert abs(float(ours.sum(-1).mean()) - float(theirs.sum(-1).mean())) < 0.6


def test_regularizer_draw_is_reproducible_and_layerdrop_rate():
    from llm_speech_summarization_b200.regularizers import RegularizerConfig, draw
    from llm_speech_summarization_b200.config import EncoderArch
    cfg = RegularizerConfig.from_arch(EncoderArch())
    assert (cfg.hidden_dropout, cfg.layerdrop, cfg.mask_time_prob, cfg.mask_time_length) == (0.1, 0.1, 0.05, 10)
    g1, g2 = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
    a, b = draw(cfg, 2, 499, 24, "cpu", g1), draw(cfg, 2, 499, 24, "cpu", g2)
    assert a.seed == b.seed and (a.layer_skip == b.layer_skip).all() and torch.equal(a.time_mask, b.time_mask)
    c = draw(cfg, 2, 499, 24, "cpu", g1)
    assert c.seed != a.seed
    g = torch.Generator().manual_seed(1)
    skips = sum(int(draw(cfg, 1, 499, 24, "cpu", g).layer_skip.sum()) for _ in range(200))
    assert abs(skips / (200 * 24) - 0.1) < 0.02
    off = RegularizerConfig(apply_spec_augment=False)
    assert draw(off, 1, 499, 24, "cpu", g).time_mask is None


def test_trainer_collate_matches_reference_semantics():
    """Trainer.collate_audio_batch_hubert / _whisper (REF/trainer.py:134-199): right zero-padding to the longest clip,
    BOS stripped from the text ids and from row 0 of the nested response ids, everything else passed through."""
    from llm_speech_summarization_b200.trainer import Trainer, WHISPER_WINDOW_SAMPLES
    tr = object.__new__(Trainer)
    data = [{"audio": {"array": torch.arange(5.0)}, "text": "a", "text_input_ids": torch.tensor([1, 7, 8]),
             "response_input_ids": torch.tensor([[1, 4, 5, 6]]), "pool_ranges_4": [0]},
            {"audio": {"array": torch.arange(3.0)}, "text": "b", "text_input_ids": torch.tensor([1, 9]),
             "response_input_ids": torch.tensor([[1, 2]]), "pool_ranges_4": [1]}]
    raw, padded, lens, texts, t_ids, r_ids, ranges = tr.collate_audio_batch_hubert(data)
    assert lens == [5, 3] and texts == ["a", "b"] and ranges == [[0], [1]]
    assert padded.shape == (2, 5) and padded.dtype == torch.float32
    assert padded[1].tolist() == [0.0, 1.0, 2.0, 0.0, 0.0]
    assert [t.tolist() for t in t_ids] == [[7, 8], [9]] and [r.tolist() for r in r_ids] == [[4, 5, 6], [2]]
    _, pw, lens_w, _, t_w, r_w, _ = tr.collate_audio_batch_whisper(data)
    assert pw.shape == (2, WHISPER_WINDOW_SAMPLES) and lens_w == [5, 3] and float(pw[0, 5:].abs().sum()) == 0.0
    assert [t.tolist() for t in t_w] == [[7, 8], [9]] and [r.tolist() for r in r_w] == [[4, 5, 6], [2]]
    # micro-batches: equal-length utterances are packed together, others run on their own, nothing is padded
    tr.encoder_base = "hubert"
    data3 = data + [dict(data[0], text="c")]
    _, padded, lens, _, t_ids, r_ids, _ = tr.collate_audio_batch_hubert(data3)
    micro = tr._micro_batches(padded, lens, t_ids, r_ids)
    assert sorted((m[0].shape for m in micro)) == [(1, 3), (2, 5)] and all(m[3] is None for m in micro)
    with pytest.raises(RuntimeError, match="no CPU path"):
        from types import SimpleNamespace as NS
        Trainer(NS(run_name="x", checkpoint_path=None), NS(), "cpu")


def test_build_plan_ragged_audio_counts():
    """Per-utterance audio-embedding counts (ragged batch): sources index the [B, stride] embedding layout, padding
    rows never enter a sequence and map to -1 in the gradient gather; an int count is the uniform special case."""
    from llm_speech_summarization_b200.step import build_plan
    prefix, suffix = [1, 2], [9, 8, 7]
    d = build_plan(prefix, suffix, [3, 1], [[5, 6], [5]], [[0, 4, 4], [0, 4]], audio_stride=3)
    assert d["row_src"][:9] == [1, 2, -1, -2, -3, 8, 7, 4, 4]            # utterance 0: 3 audio rows (sources 0..2)
    assert d["row_src"][9:15] == [1, 2, -4, 8, 7, 4]                     # utterance 1: 1 audio row (source 3 = 1*3+0)
    assert d["L_audio"] == [9, 6] and d["cu_seqlens"][:3] == [0, 9, 15]
    assert d["audio_rows"] == [2, 3, 4, 11, -1, -1]
    u = build_plan(prefix, suffix, 2, [[5, 6], [5]], [[0, 4, 4], [0, 4]])
    assert u == build_plan(prefix, suffix, [2, 2], [[5, 6], [5]], [[0, 4, 4], [0, 4]], audio_stride=2)
    with pytest.raises(AssertionError):
        build_plan(prefix, suffix, [3, 1], [[5, 6], [5]], [[0, 4, 4], [0, 4]], audio_stride=2)


def test_rng_header_host_functions_match_the_oracle_generator(tmp_path):
    """csrc/rng.cuh is host + device code: compile its HOST side with nvcc (no GPU needed) and compare the mixer, the
    stream keys, the threshold and keep/drop decisions with oracle/regularizers.py, the numpy restatement the
    train-mode parity tests rely on."""
    import shutil
    import subprocess
    import numpy as np
    from oracle import regularizers as rg
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "rng_host.cu"
    src.write_text('''
#include <cstdio>
#include "%s/llm_speech_summarization_b200/csrc/rng.cuh"
int main() {
  using namespace b2s;
  const unsigned long long seeds[3] = {1ull, 0x123456789ABCDEFull, (1ull << 61) + 12345ull};
  for (auto seed : seeds) {
    for (unsigned site : {1u, 2u, 16u, 19u, 111u}) {
      unsigned k1, k2;
      rng_stream_key(seed, site, 31u, 15u, &k1, &k2);
      std::printf("K %%llu %%u %%u %%u\\n", seed, site, k1, k2);
      const unsigned th = drop_threshold(0.1f);
      for (unsigned e : {0u, 1u, 12345u, 0xFFFFFFFFu, (498u << 16) | 480u})
        std::printf("E %%u %%u %%d\\n", e, rng_mix(e ^ k1, k2), rng_mix(e ^ k1, k2) >= th ? 1 : 0);
    }
  }
  std::printf("T %%u %%u %%u\\n", drop_threshold(0.0f), drop_threshold(0.1f), drop_threshold(0.5f));
  return 0;
}
''' % ROOT)
    exe = tmp_path / "rng_host"
    subprocess.run([nvcc, "-std=c++17", "-O1", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    k1 = k2 = seed = site = None
    n_checked = 0
    for line in out:
        f = line.split()
        if f[0] == "K":
            seed, site, k1, k2 = int(f[1]), int(f[2]), int(f[3]), int(f[4])
            assert rg.stream_key(seed, site, 31, 15) == (k1, k2), line
        elif f[0] == "E":
            e, h, keep = int(f[1]), int(f[2]), int(f[3])
            assert int(rg.mix(np.uint64(e ^ k1), k2)) == h, line
            assert bool(rg.keep(np.array([e], dtype=np.uint64), seed, site, 31, 15, 0.1)[0]) == bool(keep), line
            n_checked += 1
        elif f[0] == "T":
            assert [int(x) for x in f[1:]] == [rg.threshold(0.0), rg.threshold(0.1), rg.threshold(0.5)]
    assert n_checked == 75


def test_build_plan_arrays_equals_build_plan():
    """The numpy plan builder the step uses per micro-batch == the pure-python reference construction, on random
    ragged batches (with and without the teacher sequences, uniform and per-utterance audio counts)."""
    import random
    import numpy as np
    from llm_speech_summarization_b200.step import build_plan, build_plan_arrays
    rnd = random.Random(7)
    for trial in range(40):
        B = rnd.randint(1, 6)
        prefix = [rnd.randrange(100) for _ in range(rnd.randint(1, 9))]
        suffix = [rnd.randrange(100) for _ in range(rnd.randint(2, 6))]
        text = [[rnd.randrange(1000) for _ in range(rnd.randint(1, 12))] for _ in range(B)]
        resp = [[rnd.randrange(1000) for _ in range(rnd.randint(2, 9))] for _ in range(B)]
        if trial % 2:
            n_audio, stride = [rnd.randint(1, 7) for _ in range(B)], None
            if trial % 4 == 1:
                stride = max(n_audio) + 2
        else:
            n_audio, stride = rnd.randint(1, 7), None
        for with_teacher in (True, False):
            a = build_plan(prefix, suffix, n_audio, text, resp, with_teacher=with_teacher, audio_stride=stride)
            b = build_plan_arrays(prefix, suffix, n_audio, [np.asarray(t) for t in text], [np.asarray(r) for r in resp],
                                  with_teacher=with_teacher, audio_stride=stride)
            for k, v in a.items():
                got = b[k].tolist() if hasattr(b[k], "tolist") else b[k]
                assert got == v, (trial, with_teacher, k)
            assert b["seg"].tolist() == [i for i, r in enumerate(resp) for _ in r]
    with pytest.raises(ValueError):
        build_plan_arrays([], [2], 0, [np.asarray([4])], [np.asarray([5, 6, 7, 8, 9])], with_teacher=False)


def test_bench_algorithmic_flops_match_the_survey_figures():
    """bench.py's roofline numerator: SURVEY.md section 8d gives 384.6 GFLOP for the HuBERT-large forward (10 s) and
    1184.7 + 712.3 GFLOP for the student + teacher prefill with the LM head on the consumed rows; the forward credits
    what is computed (the last layer's out-projection / MLP run on the 2 * R consumed rows only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fwd, train = bench.gemm_flops_per_utt(False), bench.gemm_flops_per_utt(True)
    skipped = (200 + 117 - 128) * 2 * 3072 * (3072 + 2 * 8192 + 8192)
    # SURVEY counts attention (LLM: causal-halved; encoder: 4 * N^2 * H per layer) and conv layer 0 too, which are not
    # launches of the GEMM kernel here
    attn = 28 * 2 * 3072 * (200 ** 2 + 117 ** 2) + 24 * 4 * 499 ** 2 * 1024
    conv0 = 2 * 512 * 10 * 31999
    survey_total = (384.6 + 1184.7 + 712.3) * 1e9
    assert abs((fwd + skipped + attn + conv0) - survey_total) / survey_total < 5e-3
    assert abs(fwd / 1e9 - 2215.5) < 1.0
    assert train > 2.5 * fwd * 0.6 and train < 3.0 * fwd


# --------------------------------------------------------------------------------------- round-2 advisor findings
def test_yaml_scientific_floats_reach_the_optimizer(tmp_path):
    """`lr: 5e-5` -- the spelling of every shipped reference yaml (REF/config/*.yaml) -- must load as a float (PyYAML's
    YAML 1.1 resolver reads it as a string) so that PolynomialLR / AdamW can multiply it."""
    from llm_speech_summarization_b200.config import load_config
    from llm_speech_summarization_b200.training import PolynomialLR
    y = tmp_path / "c.yaml"
    y.write_text("train:\n  epochs: 2\n  optimizer:\n    lr: 5e-5\n    beta1: 0.9\n    beta2: 0.999\n"
                 "  tiny: 1E-8\n  neg: -2.5e+3\n  name: e5\n  ver: 1.0\n  n: 16\n")
    c = load_config(str(y))
    o = c.train.optimizer
    assert isinstance(o.lr, float) and o.lr == 5e-5 and isinstance(o.beta1, float)
    assert c.train.tiny == 1e-8 and c.train.neg == -2500.0 and c.train.name == "e5" and c.train.n == 16

    class Opt:  # the scheduler only needs .defaults / .lr
        defaults = {"lr": o.lr}
        lr = o.lr
    sch = PolynomialLR(Opt, total_iters=10)
    sch.step()
    assert abs(sch.get_last_lr()[0] - 4.5e-5) < 1e-12
    ref_dir = "/root/reference/config"  # only present in the build container; the GPU box skips this part
    if os.path.isdir(ref_dir):
        for name in sorted(os.listdir(ref_dir)):
            if name.endswith(".yaml"):
                t = load_config(os.path.join(ref_dir, name)).train
                assert isinstance(t.optimizer.lr, float) and isinstance(t.optimizer.beta2, float), name


def test_encoder_is_never_left_at_zero_weights(monkeypatch):
    """AudioEncoder(config) loads the pretrained backbone like the reference (REF/model/audio_encoder.py:6-13); when the
    hub is unreachable it RAISES unless the config asks for random_init -- and random_init gives HF's default init
    (LayerNorm gamma 1, N(0, 0.02) linears), never an all-zero module."""
    from oracle import configs
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    cfg = ns_config(configs.TINY_ENCODER, configs.TINY_LLAMA)
    cfg.model.audio_encoder.random_init = False
    cfg.model.audio_encoder.type = "definitely/not-a-local-checkpoint"
    with pytest.raises(RuntimeError, match="could not load the pretrained audio encoder"):
        AudioEncoder(cfg, torch.device("cpu"))
    cfg.model.audio_encoder.random_init = True
    enc = AudioEncoder(cfg, torch.device("cpu"))
    sd = enc.state_dict()
    assert float(sd["encoder.encoder.layer_norm.weight"].min()) == 1.0
    assert float(sd["encoder.encoder.layers.0.attention.q_proj.weight"].std()) > 0.01
    assert float(sd["encoder.feature_extractor.conv_layers.1.conv.weight"].abs().max()) > 0
    v = sd["encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    g0 = sd["encoder.encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    assert torch.allclose(g0, v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt())
    wcfg = ns_config_whisper(configs.TINY_WHISPER)
    wenc = AudioEncoder(wcfg, torch.device("cpu"))
    pos = wenc.state_dict()["encoder.embed_positions.weight"]
    assert float(pos[0, : pos.shape[1] // 2].abs().max()) == 0.0 and float(pos[0, pos.shape[1] // 2:].min()) == 1.0


def test_allreduce_buckets_tile_the_flat_gradient():
    """The training step exchanges gradients in one bucket per transformer layer (+ one for everything else): in
    AudioEncoder.flat_param_order every layer's parameters form ONE contiguous range of the flat optimizer buffer, the
    ranges are disjoint, q|k|v stay adjacent (the fused-QKV gradient aliases them), and layers + rest cover every
    element exactly once."""
    from oracle import configs
    from llm_speech_summarization_b200.model.audio_encoder import AudioEncoder
    for cfg in (ns_config(configs.TINY_ENCODER, configs.TINY_LLAMA), ns_config_whisper(configs.TINY_WHISPER)):
        enc = AudioEncoder(cfg, torch.device("cpu"))
        order = [p for p in enc.flat_param_order() if p.requires_grad]
        assert len({id(p) for p in order}) == len(order) == sum(1 for p in enc.parameters() if p.requires_grad)
        off, at = 0, {}
        for p in order:
            at[id(p)] = (off, off + p.numel())
            off += p.numel()
        covered = []
        for group in enc.layer_param_groups():
            spans = sorted(at[id(p)] for p in group)
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])), "layer parameters are not contiguous"
            covered.append((spans[0][0], spans[-1][1]))
        covered.sort()
        assert all(a[1] <= b[0] for a, b in zip(covered, covered[1:])), "layer buckets overlap"
        assert covered[0][0] == 0  # the transformer layers come first, everything else follows as the last bucket
        layers = enc.transformer_layers()
        a = layers[0].attention if enc.encoder_base == "hubert" else layers[0].self_attn
        q, k, v = (at[id(m.weight)] for m in (a.q_proj, a.k_proj, a.v_proj))
        assert q[1] == k[0] and k[1] == v[0]
        assert sum(b - a_ for a_, b in covered) + (off - covered[-1][1]) == off


def test_handles_own_their_state():
    """include/b2s.h: options, SM budget and the launch counter belong to a b2s_handle, not to the process: two handles
    do not see each other's settings, the process default is untouched, destroy() falls back to the default."""
    import ctypes as C
    from llm_speech_summarization_b200 import _lib
    lib = _lib.load()
    get = lambda opt: (lambda v: (lib.b2s_get_option(opt, C.byref(v)), v.value)[1])(C.c_int32())
    base_pdl, base_budget = get(_lib.OPT_PDL), get(_lib.OPT_SM_BUDGET)
    h1, h2 = C.c_void_p(), C.c_void_p()
    assert lib.b2s_create(C.byref(h1)) == 0 and lib.b2s_create(C.byref(h2)) == 0
    try:
        assert lib.b2s_make_current(h1) == 0
        assert lib.b2s_set_option(_lib.OPT_PDL, 0) == 0 and lib.b2s_set_option(_lib.OPT_TMA_EPILOGUE, 0) == 0
        lib.b2s_set_sm_budget(140)
        assert (get(_lib.OPT_PDL), get(_lib.OPT_TMA_EPILOGUE), lib.b2s_get_sm_budget()) == (0, 0, 140)
        assert lib.b2s_make_current(h2) == 0
        assert (get(_lib.OPT_PDL), get(_lib.OPT_TMA_EPILOGUE), lib.b2s_get_sm_budget()) == (base_pdl, 1, 0)
        assert lib.b2s_set_option(99, 1) != 0 and b"unknown option" in lib.b2s_last_error()
        assert lib.b2s_make_current(None) == 0
        assert (get(_lib.OPT_PDL), get(_lib.OPT_SM_BUDGET)) == (base_pdl, base_budget)
        assert lib.b2s_make_current(h1) == 0 and get(_lib.OPT_PDL) == 0
    finally:
        assert lib.b2s_destroy(h1) == 0 and lib.b2s_destroy(h2) == 0
    assert get(_lib.OPT_PDL) == base_pdl  # destroying the current handle falls back to the process default
    rank, world = C.c_int32(-1), C.c_int32(-1)
    assert lib.b2s_comm_world(C.byref(rank), C.byref(world)) == 0 and (rank.value, world.value) == (0, 1)
    flat = (C.c_float * 4)()
    assert lib.b2s_allreduce_join(None) != 0 and b"no communicator" in lib.b2s_last_error()


def test_option_ids_match_the_header():
    """_lib.OPT_* are positional mirrors of the B2S_OPT_* enum in include/b2s.h: a new option added on one side only would
    silently set a different knob."""
    import re
    from llm_speech_summarization_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "b2s.h")).read()
    enum = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"B2S_OPT_([A-Z0-9_]+)\s*=\s*(\d+)", hdr))
    mirror = {"PDL": _lib.OPT_PDL, "RESID_RED": _lib.OPT_RESID_RED, "TMA_EPILOGUE": _lib.OPT_TMA_EPILOGUE,
              "ATTN_KEYS_PER_STEP": _lib.OPT_ATTN_KEYS_PER_STEP, "ATTN_KV_STAGES": _lib.OPT_ATTN_KV_STAGES,
              "SM_BUDGET": _lib.OPT_SM_BUDGET, "GEMM_GROUP_M": _lib.OPT_GEMM_GROUP_M,
              "GEMM_TAIL_SPLIT": _lib.OPT_GEMM_TAIL_SPLIT, "GEMM_EPI8": _lib.OPT_GEMM_EPI8}
    assert enum == mirror


def test_shared_prefix_plan_expands_to_the_reference_layout():
    """build_plan_arrays(shared_prefix=True): the prefix is sequence 0, every other sequence is the reference's sequence
    minus its prefix, positions continue after the prefix, and every index the losses / the splice use points at the
    same token with the same position as in the reference layout."""
    import numpy as np
    from llm_speech_summarization_b200.step import build_plan_arrays
    rng = np.random.default_rng(0)
    pre, suf = [1, 2, 3, 4], [9, 8, 7]
    B, A = 5, 6
    n_audio = [6, 3, 6, 1, 4]
    text = [rng.integers(10, 99, size=rng.integers(1, 9)).astype(np.int32) for _ in range(B)]
    resp = [np.concatenate([[5], rng.integers(10, 99, size=rng.integers(1, 5))]).astype(np.int32) for _ in range(B)]
    a = build_plan_arrays(pre, suf, n_audio, text, resp, audio_stride=A)
    b = build_plan_arrays(pre, suf, n_audio, text, resp, audio_stride=A, shared_prefix=True)
    seqs = lambda d: [d["row_src"][d["cu_seqlens"][i]:d["cu_seqlens"][i + 1]] for i in range(len(d["cu_seqlens"]) - 1)]
    poss = lambda d: [d["positions"][d["cu_seqlens"][i]:d["cu_seqlens"][i + 1]] for i in range(len(d["cu_seqlens"]) - 1)]
    sa, sb, pa, pb = seqs(a), seqs(b), poss(a), poss(b)
    assert b["shared_prefix_len"] == 4 and a.get("shared_prefix_len", 0) == 0
    assert len(sb) == 2 * B + 1 and np.array_equal(sb[0], pre) and np.array_equal(pb[0], np.arange(4))
    for i in range(2 * B):
        assert np.array_equal(np.concatenate([sb[0], sb[i + 1]]), sa[i])
        assert np.array_equal(np.concatenate([pb[0], pb[i + 1]]), pa[i])
    for key in ("student_rows", "teacher_rows"):
        assert np.array_equal(a["row_src"][a[key]], b["row_src"][b[key]])
        assert np.array_equal(a["positions"][a[key]], b["positions"][b[key]])
    for x, y in zip(a["audio_rows"], b["audio_rows"]):
        assert (x < 0) == (y < 0) and (x < 0 or a["row_src"][x] == b["row_src"][y])
    for key in ("labels", "row_offsets", "seg", "resp_lens", "L_audio", "L_text", "sum_r"):
        assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), key
    assert b["rows"] == a["rows"] - (2 * B - 1) * 4 and b["max_seqlen"] == a["max_seqlen"]
