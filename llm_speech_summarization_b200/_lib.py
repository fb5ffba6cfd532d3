"""ctypes binding of libb2s.so (the C ABI in include/b2s.h).

There is deliberately no fallback: if the library is missing it is built in-tree with nvcc; if that fails,
or a compute entry point is called without a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2s.so")

c_void_p, c_int, c_int64, c_float, c_size_t = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
P_int = C.c_void_p  # device int32*
P_f32 = C.c_void_p  # device float*


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("a_dim0", c_int), ("a_row_stride", c_int64), ("a_batch_stride", c_int64),
        ("a_rows", c_int), ("W", c_void_p), ("w_rows", c_int), ("w_cols", c_int),
        ("M", c_int), ("N", c_int), ("batches", c_int), ("groups", c_int), ("taps", c_int),
        ("k_per_tap", c_int), ("a_pad", c_int), ("a_group_off", c_int), ("w_group_off", c_int),
        ("epi", c_int), ("act", c_int), ("bias", c_void_p), ("out", c_void_p), ("ldo", c_int64),
        ("out_batch_rows", c_int64), ("resid", c_void_p), ("rope_cs", c_void_p), ("positions", c_void_p),
        ("rope_cols", c_int), ("resid_bcast", c_int), ("out2", c_void_p), ("ld2", c_int64), ("block_n", c_int),
        ("cta_group", c_int), ("a_mn", c_int), ("b_mn", c_int), ("k_batches", c_int), ("w_row_stride", c_int64),
        ("w_batch_stride", c_int64), ("k_splits", c_int), ("b_tap_atoms", c_int), ("out_group_rows", c_int),
        ("out_group_cols", c_int), ("a_fmt", c_int), ("w_fmt", c_int), ("out_fmt", c_int),
    ]


class EncoderLayer(C.Structure):
    _fields_ = [(n, c_void_p) for n in
                ("ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2")]


class HubertWeights(C.Structure):
    _fields_ = [
        ("conv0_w", c_void_p), ("conv0_b", c_void_p), ("conv0_ln_g", c_void_p), ("conv0_ln_b", c_void_p),
        ("conv_w", c_void_p * 6), ("conv_b", c_void_p * 6), ("conv_ln_g", c_void_p * 6), ("conv_ln_b", c_void_p * 6),
        ("conv_k", c_int * 6), ("conv_stride", c_int * 6),
        ("fp_ln_g", c_void_p), ("fp_ln_b", c_void_p), ("fp_w", c_void_p), ("fp_b", c_void_p),
        ("pos_w", c_void_p), ("pos_b", c_void_p), ("pos_k", c_int), ("pos_groups", c_int),
        ("layers", C.POINTER(EncoderLayer)), ("num_layers", c_int), ("hidden", c_int), ("heads", c_int),
        ("ffn", c_int), ("final_ln_g", c_void_p), ("final_ln_b", c_void_p), ("ln_eps", c_float),
        ("pool_kernel", c_int), ("pool_stride", c_int), ("proj_w", c_void_p), ("proj_b", c_void_p),
        ("llm_dim", c_int), ("fmt", c_int),
    ]


class EncoderLayerGrads(C.Structure):
    _fields_ = [(n, c_void_p) for n in
                ("ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo", "ln2_g", "ln2_b", "w1", "b1", "w2", "b2")]


class HubertGrads(C.Structure):
    _fields_ = [
        ("conv0_w", c_void_p), ("conv0_b", c_void_p), ("conv0_ln_g", c_void_p), ("conv0_ln_b", c_void_p),
        ("conv_w", c_void_p * 6), ("conv_b", c_void_p * 6), ("conv_ln_g", c_void_p * 6), ("conv_ln_b", c_void_p * 6),
        ("fp_ln_g", c_void_p), ("fp_ln_b", c_void_p), ("fp_w", c_void_p), ("fp_b", c_void_p),
        ("pos_w", c_void_p), ("pos_b", c_void_p), ("layers", C.POINTER(EncoderLayerGrads)),
        ("final_ln_g", c_void_p), ("final_ln_b", c_void_p), ("proj_w", c_void_p), ("proj_b", c_void_p),
    ]


class EncoderRegularizers(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("p_feat_proj", C.c_float), ("p_hidden", C.c_float), ("p_attention", C.c_float),
                ("p_activation", C.c_float), ("layer_skip", c_void_p), ("time_mask", c_void_p),
                ("masked_spec_embed", c_void_p), ("g_masked_spec_embed", c_void_p)]


class WhisperGrads(C.Structure):
    _fields_ = [("conv1_w", c_void_p), ("conv1_b", c_void_p), ("conv2_w", c_void_p), ("conv2_b", c_void_p),
                ("layers", C.POINTER(EncoderLayerGrads)), ("final_ln_g", c_void_p), ("final_ln_b", c_void_p),
                ("proj_w", c_void_p), ("proj_b", c_void_p)]


class WhisperWeights(C.Structure):
    _fields_ = [
        ("conv1_w", c_void_p), ("conv1_b", c_void_p), ("conv2_w", c_void_p), ("conv2_b", c_void_p),
        ("pos_emb", c_void_p), ("layers", C.POINTER(EncoderLayer)), ("num_layers", c_int), ("hidden", c_int),
        ("heads", c_int), ("ffn", c_int), ("mel_bins", c_int), ("max_positions", c_int),
        ("final_ln_g", c_void_p), ("final_ln_b", c_void_p), ("ln_eps", c_float),
        ("pool_kernel", c_int), ("pool_stride", c_int), ("proj_w", c_void_p), ("proj_b", c_void_p),
        ("llm_dim", c_int), ("fmt", c_int),
    ]


class LlamaLayer(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("ln1_w", "wqkv", "wo", "ln2_w", "wgu", "wd")]


class LlamaWeights(C.Structure):
    _fields_ = [
        ("layers", C.POINTER(LlamaLayer)), ("num_layers", c_int), ("hidden", c_int), ("heads", c_int),
        ("kv_heads", c_int), ("head_dim", c_int), ("ffn", c_int), ("vocab", c_int), ("rms_eps", c_float),
        ("final_norm_w", c_void_p), ("lm_head", c_void_p), ("rope_cs", c_void_p), ("max_pos", c_int),
        ("fmt", c_int),
    ]


class LlamaLayerT(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("wqkv_t", "wo_t", "wgu_t", "wd_t")]


class LlamaWeightsT(C.Structure):
    _fields_ = [("layers", C.POINTER(LlamaLayerT)), ("lm_head_t", c_void_p)]


class GradScalerState(C.Structure):
    """b2s_grad_scaler_state (device memory; this mirror is for sizing and for reading it back)."""
    _fields_ = [("scale", c_float), ("growth_tracker", c_int), ("found_inf", c_int), ("opt_steps", c_int),
                ("skipped_steps", c_int), ("reserved", c_int * 3)]


class LlamaSaved(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("h", "h_mid", "qkv", "ao", "lse", "gu")]


# name -> (restype, argtypes); mirrors include/b2s.h one to one (tests/test_abi.py checks the symbol list)
PROTOTYPES = {
    "b2s_last_error": (C.c_char_p, []),
    "b2s_version": (c_int, []),
    "b2s_launch_count": (C.c_longlong, []),
    "b2s_set_sm_budget": (None, [c_int]),
    "b2s_get_sm_budget": (c_int, []),
    "b2s_create": (c_int, [C.POINTER(c_void_p)]),
    "b2s_destroy": (c_int, [c_void_p]),
    "b2s_make_current": (c_int, [c_void_p]),
    "b2s_set_option": (c_int, [c_int, c_int]),
    "b2s_get_option": (c_int, [c_int, C.POINTER(c_int)]),
    "b2s_comm_unique_id": (c_int, [c_void_p]),
    "b2s_comm_init": (c_int, [c_void_p, c_int, c_int]),
    "b2s_comm_world": (c_int, [C.POINTER(c_int), C.POINTER(c_int)]),
    "b2s_comm_destroy": (c_int, []),
    "b2s_allreduce_grads": (c_int, [P_f32, C.POINTER(c_int64), C.POINTER(c_int64), C.POINTER(c_void_p), c_int,
                                    c_void_p, c_void_p]),
    "b2s_allreduce_join": (c_int, [c_void_p]),
    "b2s_gemm_bf16": (c_int, [C.POINTER(GemmArgs), c_void_p]),
    "b2s_gemm_timing_enable": (None, [c_int]),
    "b2s_gemm_timing_read": (c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "b2s_gemm_timing_get": (c_int, [c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "b2s_kd_ce_workspace_bytes": (c_size_t, [c_int, c_int]),
    "b2s_kd_ce_loss_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, P_int, P_int, c_int, c_float,
                                   c_float, c_void_p, P_f32, P_f32, P_f32, P_f32, P_f32, P_f32, c_void_p]),
    "b2s_kd_ce_loss_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, P_int, P_f32, P_f32, P_f32,
                                   P_f32, P_f32, c_void_p, c_int64, c_int, c_void_p]),
    "b2s_layernorm_fwd": (c_int, [c_void_p, c_int, P_f32, P_f32, c_float, c_int, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "b2s_rmsnorm_fwd": (c_int, [P_f32, P_f32, c_float, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "b2s_rmsnorm_gather_fwd": (c_int, [P_f32, P_int, P_f32, c_float, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "b2s_layernorm_avgpool_fwd": (c_int, [P_f32, P_f32, P_f32, c_float, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                          c_int, c_int, c_void_p]),
    "b2s_conv0_ln_gelu_fwd": (c_int, [P_f32, c_int64, c_int, c_int, P_f32, P_f32, P_f32, P_f32, c_float, c_void_p,
                                      c_int, c_int, c_void_p]),
    "b2s_embed_splice_fwd": (c_int, [c_void_p, P_f32, P_int, P_f32, c_int64, c_int, c_int, c_void_p]),
    "b2s_rowpair_sqdiff_fwd": (c_int, [P_f32, P_int, P_int, P_f32, c_int, c_int, c_void_p]),
    "b2s_posconv_weight_pack": (c_int, [P_f32, P_f32, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b2s_cast_f32_to_h16": (c_int, [P_f32, c_void_p, c_int64, c_int, c_void_p]),
    "b2s_cast_h16_to_f32": (c_int, [c_void_p, P_f32, c_int64, c_int, c_void_p]),
    "b2s_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, P_int, c_int, c_int,
                                  c_int64, c_int, c_int, c_int, c_float, c_int, P_f32, c_int, c_void_p]),
    "b2s_attention_fwd_prefix": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, P_int, c_int, c_int,
                                         c_int64, c_int, c_int, c_int, c_float, P_f32, c_int, c_int, c_void_p]),
    "b2s_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, P_f32,
                                  P_f32, c_void_p, c_void_p, c_void_p, c_int64, P_int, c_int, c_int, c_int64, c_int,
                                  c_int, c_int, c_float, c_int, P_f32, c_int, c_void_p]),
    "b2s_hubert_num_frames": (c_int, [C.POINTER(HubertWeights), c_int, C.POINTER(c_int), C.POINTER(c_int)]),
    "b2s_hubert_workspace_bytes": (c_size_t, [C.POINTER(HubertWeights), c_int, c_int]),
    "b2s_hubert_forward": (c_int, [C.POINTER(HubertWeights), P_f32, c_int64, c_int, c_int, c_void_p, c_size_t, P_f32,
                                   P_f32, c_void_p]),
    "b2s_whisper_workspace_bytes": (c_size_t, [C.POINTER(WhisperWeights), c_int]),
    "b2s_whisper_forward": (c_int, [C.POINTER(WhisperWeights), P_f32, c_int, c_int, c_void_p, c_size_t, P_f32, P_f32,
                                    c_void_p]),
    "b2s_llama_workspace_bytes": (c_size_t, [C.POINTER(LlamaWeights), c_int, c_int]),
    "b2s_llama_prefill": (c_int, [C.POINTER(LlamaWeights), P_f32, c_int, P_int, c_int, c_int, P_int, P_int, c_int,
                                  c_void_p, C.POINTER(c_int), c_int, P_int, P_int, c_int, P_f32, P_f32, c_void_p,
                                  c_size_t, c_void_p]),
    "b2s_llama_prefill_prefix": (c_int, [C.POINTER(LlamaWeights), P_f32, c_int, P_int, c_int, c_int, P_int, P_int, c_int,
                                         c_void_p, C.POINTER(c_int), c_int, P_int, P_int, c_int, P_f32, P_f32, c_void_p,
                                         c_size_t, c_int, c_void_p]),
    "b2s_llama_train_workspace_bytes": (c_size_t, [C.POINTER(LlamaWeights), c_int, c_int]),
    "b2s_llama_backward_workspace_bytes": (c_size_t, [C.POINTER(LlamaWeights), c_int, c_int]),
    "b2s_llama_forward_train": (c_int, [C.POINTER(LlamaWeights), C.POINTER(LlamaSaved), c_int, P_int, c_int, c_int,
                                        P_int, P_int, c_int, c_void_p, C.POINTER(c_int), c_int, P_int, P_int, c_int,
                                        P_f32, c_void_p, c_size_t, c_void_p]),
    "b2s_llama_backward": (c_int, [C.POINTER(LlamaWeights), C.POINTER(LlamaWeightsT), C.POINTER(LlamaSaved), c_int,
                                   c_int, P_int, c_int, c_int, c_void_p, P_int, c_int, C.POINTER(c_int), c_int, P_int,
                                   P_int, P_f32, P_f32, c_int, P_f32, c_void_p, c_size_t, c_void_p]),
    "b2s_rmsnorm_bwd": (c_int, [P_f32, P_int, P_f32, c_float, P_f32, P_f32, P_int, c_void_p, c_int64, c_int,
                                c_int, c_void_p]),
    "b2s_layernorm_bwd": (c_int, [P_f32, P_f32, c_float, c_void_p, c_int, P_f32, c_int, c_void_p, P_f32, P_f32,
                                  c_int64, c_int, c_int, c_void_p]),
    "b2s_swiglu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "b2s_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "b2s_gather_rows_f32": (c_int, [P_f32, P_int, P_f32, c_int64, c_int, c_void_p]),
    "b2s_whisper_log_mel": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b2s_llama_kv_cache_bytes": (C.c_size_t, [C.POINTER(LlamaWeights), c_int]),
    "b2s_llama_prefill_kv": (c_int, [C.POINTER(LlamaWeights), c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                     c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, C.c_size_t,
                                     c_void_p]),
    "b2s_llama_decode_workspace_bytes": (C.c_size_t, [C.POINTER(LlamaWeights), c_int]),
    "b2s_llama_decode_step": (c_int, [C.POINTER(LlamaWeights), c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "b2s_hubert_saved_bytes": (C.c_size_t, [C.POINTER(HubertWeights), c_int, c_int]),
    "b2s_hubert_backward_workspace_bytes": (C.c_size_t, [C.POINTER(HubertWeights), c_int, c_int]),
    "b2s_hubert_forward_train": (c_int, [C.POINTER(HubertWeights), c_void_p, c_int64, c_int, c_int,
                                         C.POINTER(C.c_int32), c_void_p, C.c_size_t, c_void_p,
                                         C.POINTER(EncoderRegularizers), c_void_p]),
    "b2s_hubert_backward": (c_int, [C.POINTER(HubertWeights), c_void_p, C.POINTER(HubertGrads), c_void_p, c_int64,
                                    c_int, c_int, C.POINTER(C.c_int32), c_void_p, C.c_size_t, c_void_p, c_void_p,
                                    C.c_size_t, C.POINTER(EncoderRegularizers), C.POINTER(c_void_p), c_void_p]),
    "b2s_drop_mask_dump": (c_int, [c_void_p, c_int64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, c_float,
                                   C.c_uint32, c_void_p]),
    "b2s_whisper_saved_bytes": (C.c_size_t, [C.POINTER(WhisperWeights), c_int]),
    "b2s_whisper_backward_workspace_bytes": (C.c_size_t, [C.POINTER(WhisperWeights), c_int]),
    "b2s_whisper_forward_train": (c_int, [C.POINTER(WhisperWeights), c_void_p, c_int, c_int, c_void_p, C.c_size_t,
                                          c_void_p, c_void_p]),
    "b2s_whisper_backward": (c_int, [C.POINTER(WhisperWeights), C.POINTER(WhisperGrads), c_int, c_void_p, C.c_size_t,
                                     c_void_p, c_void_p, C.c_size_t, C.POINTER(c_void_p), c_void_p]),
    "b2s_layernorm_bwd_ex": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int, c_void_p,
                                     c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p,
                                     c_void_p]),
    "b2s_colsum_accum": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "b2s_avgpool_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b2s_col2im_add": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b2s_conv0_bwd": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                              c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "b2s_adamw_step": (c_int, [P_f32, P_f32, P_f32, P_f32, c_int64, c_float, c_float, c_float, c_float, c_float,
                               c_int, c_float, c_void_p, c_void_p]),
    "b2s_nonfinite_check": (c_int, [P_f32, c_int64, c_void_p, c_void_p]),
    "b2s_grad_scaler_update": (c_int, [c_void_p, c_float, c_float, c_int, c_void_p]),
}

OPT_PDL, OPT_RESID_RED, OPT_TMA_EPILOGUE, OPT_ATTN_KEYS_PER_STEP, OPT_ATTN_KV_STAGES, OPT_SM_BUDGET, OPT_GEMM_GROUP_M, \
    OPT_GEMM_TAIL_SPLIT, OPT_GEMM_EPI8 = range(9)
COMM_ID_BYTES = 128

_lib = None


def load() -> C.CDLL:
    """Load (building first if necessary) libb2s.so and attach prototypes. Raises on any failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build(verbose=False)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().b2s_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libb2s {what} failed with status {status}: {msg}")
