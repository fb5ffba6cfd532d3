/* b2s.h -- C ABI of libb2s.so: the B200-native (sm_100a) audio-prompt forward / loss path.
 *
 * Drop-in boundary for wonjune-kang/llm-speech-summarization's one hot path. The reference
 * is pure Python, so "the FFI a maintainer would bind" is ctypes; every entry point
 * below names the reference code it replaces (REF/ = /root/reference, TF/ = transformers).
 *
 * Conventions
 *   - every function returns 0 on success or a negative b2s_status; b2s_last_error() gives the message
 *     (thread-local); the Python shim turns a nonzero status into RuntimeError;
 *   - all pointers are DEVICE pointers unless marked host; sizes are explicit; `stream` is a cudaStream_t
 *     passed as void* (the caller's current stream); no function synchronises or allocates caller-visible
 *     memory: outputs and workspaces are allocated by the caller (torch.empty) and passed in;
 *   - bf16 tensors are `void*` to keep the header free of CUDA types; fp32 tensors are `float*`;
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails with B2S_ERR_CUDA.
 */
#ifndef B2S_H_
#define B2S_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B2S_STATUS_OK = 0,
  B2S_STATUS_INVALID = -1,
  B2S_STATUS_CUDA = -2,
  B2S_STATUS_UNSUPPORTED = -3
} b2s_status;

/* epilogue ids of b2s_gemm_bf16 */
enum { B2S_EPI_BF16 = 0, B2S_EPI_RESID_F32 = 1, B2S_EPI_SWIGLU = 2, B2S_EPI_ROPE = 3, B2S_EPI_F32 = 4 };
enum { B2S_ACT_NONE = 0, B2S_ACT_GELU = 1 };
/* 16-bit storage format of a tensor. Arithmetic is fp32 everywhere; the format only says how GEMM / attention operands
 * are rounded when they are stored. The reference runs this path under fp16 autocast with an fp16 LLM
 * (REF/trainer.py:57-61,270; REF/inference.py:97-99), so B2S_FMT_F16 is the default of the Python shim for weights and
 * activations; gradients are always B2S_FMT_BF16 (fp16's range would need the reference's GradScaler). */
enum { B2S_FMT_BF16 = 0, B2S_FMT_F16 = 1 };

const char* b2s_last_error(void);
int b2s_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long b2s_launch_count(void);
/* ---------------------------------------------------------------------------------------------
 * Instance state ("persistent state is owned by an opaque handle created / destroyed explicitly").
 * A b2s_handle owns everything that is not an argument of an entry point: the A/B options below, the SM budget, the
 * launch counter, the GEMM timing window and the NCCL communicator of the gradient exchange. Entry points act on the
 * calling host thread's CURRENT handle (b2s_make_current, like a CUDA context); a process-wide default handle exists, so
 * callers that never create one keep working. Calls are reentrant across handles; one handle is not thread-safe. */
typedef struct b2s_handle b2s_handle;
int b2s_create(b2s_handle** out);
int b2s_destroy(b2s_handle* h);        /* also tears down its communicator */
int b2s_make_current(b2s_handle* h);   /* NULL = back to the process default */
enum {
  B2S_OPT_PDL = 0,                /* programmatic dependent launch between kernels (default 1; env seed B2S_PDL) */
  B2S_OPT_RESID_RED = 1,          /* in-place residual epilogues as L2 reductions (default 1; B2S_RESID_RED) */
  B2S_OPT_TMA_EPILOGUE = 2,       /* forward GEMM outputs through TMA stores / reductions (default 1; B2S_TMA_EPI) */
  B2S_OPT_ATTN_KEYS_PER_STEP = 3, /* attention forward tile override, 0 = auto (B2S_ATTN_CFG="keys,stages") */
  B2S_OPT_ATTN_KV_STAGES = 4,
  B2S_OPT_SM_BUDGET = 5,          /* = b2s_set_sm_budget */
  B2S_OPT_GEMM_GROUP_M = 6,       /* GEMM tile order: M tiles per group, 0 = default 8 (B2S_GEMM_GROUP_M) */
  B2S_OPT_GEMM_TAIL_SPLIT = 7,    /* K-slice the tiles of a partly filled last round of the persistent GEMM grid (fp32
                                     outputs; B2S_GEMM_TAIL_SPLIT). GEMMs of ONE handle must be stream-ordered: the
                                     slices of a tile hand over through flags the handle owns. Default 1 */
  B2S_OPT_GEMM_EPI8 = 8           /* eight epilogue warps in the forward GEMM kernel: 0 never, 1 for epilogues with an
                                     activation and a reduction of at most 2048 (default), 2 always (B2S_GEMM_EPI8) */
};
int b2s_set_option(int32_t option, int32_t value); /* on the current handle; measurement A/B only */
int b2s_get_option(int32_t option, int32_t* value);

/* Gradient exchange of the training step (REF/trainer.py:372-384 accumulation semantics, summed
 * over ranks): the current handle owns an NCCL communicator and a communication stream.
 *   b2s_comm_unique_id  rank 0 mints the id (host buffer of B2S_COMM_ID_BYTES), the caller distributes it
 *                       (torch.distributed broadcast, a file, MPI ...);
 *   b2s_comm_init       every rank joins (collective: blocks until all `world` ranks have called it);
 *   b2s_allreduce_grads SUM all-reduce, in place, of the element ranges [starts[i], ends[i]) of the flat fp32 gradient
 *                       on the communication stream; bucket i is held back until ready_events[i] (a cudaEvent_t the
 *                       compute stream recorded, e.g. b2s_hubert_backward's layer_done_events; NULL = ready) has fired.
 *                       span_begin_event / span_end_event (optional timing events) are recorded on the communication
 *                       stream before the first / after the last bucket;
 *   b2s_allreduce_join  `compute_stream` waits for every bucket enqueued so far.
 * NCCL is resolved with dlopen at the first call: the library itself loads without it. */
#define B2S_COMM_ID_BYTES 128
int b2s_comm_unique_id(uint8_t* id /* host, B2S_COMM_ID_BYTES */);
int b2s_comm_init(const uint8_t* id, int32_t rank, int32_t world);
int b2s_comm_world(int32_t* rank, int32_t* world);
int b2s_comm_destroy(void);
int b2s_allreduce_grads(float* flat, const int64_t* starts /* host */, const int64_t* ends /* host */,
                        void* const* ready_events /* host array of cudaEvent_t, may be NULL */, int32_t n_buckets,
                        void* span_begin_event, void* span_end_event);
int b2s_allreduce_join(void* compute_stream);

/* SM budget of the current handle: the persistent kernels (GEMM, attention
 * forward) size their grids for min(device SMs, budget); 0 = all. Their tile assignment is static, so a grid that counts
 * on SMs a concurrent NCCL kernel occupies would serialise behind it: the training step lowers the budget by the
 * communication kernel's CTA count while gradient buckets are being all-reduced under the encoder backward. */
void b2s_set_sm_budget(int32_t sms);
int32_t b2s_get_sm_budget(void);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction: out = epi(A . W^T).  Replaces every nn.Linear / Conv1d(C_in >= 512) on the path
 * (TF/models/hubert/modeling_hubert.py:45-92,127-151,216-231,262-369; REF/model/audio_encoder.py:87;
 *  TF/models/llama/modeling_llama.py:171-289; REF/model/audio_llama.py:67).
 * A is a bf16 3-D view (k contiguous | row stride | batch stride); see csrc/gemm_sm100.cuh for the meaning
 * of taps / a_pad / groups (strided and grouped convolutions as implicit GEMM, zero padding by TMA). */
typedef struct {
  const void* A;
  int32_t a_dim0;
  int64_t a_row_stride;
  int64_t a_batch_stride;
  int32_t a_rows;
  const void* W;
  int32_t w_rows;
  int32_t w_cols;
  int32_t M, N, batches, groups, taps, k_per_tap, a_pad, a_group_off, w_group_off;
  int32_t epi, act;
  const float* bias;
  void* out;
  int64_t ldo;
  int64_t out_batch_rows;
  const float* resid;
  const float* rope_cs;
  const int32_t* positions;
  int32_t rope_cols;
  int32_t resid_bcast; /* 1: resid is [M, ldo] shared by every batch (e.g. a positional table) */
  void* out2;          /* optional bf16 [rows, ld2]: pre-activation copy (EPI_BF16+act) or raw gate|up (EPI_SWIGLU) */
  int64_t ld2;
  int32_t block_n;   /* 0 = auto, else 64/128/256 */
  int32_t cta_group; /* 0 = auto, else 1/2 */
  /* backward GEMMs (no transposed copies): b_mn = W stored [K rows, N cols] (dgrad dX = dY . W with W in nn.Linear
   * layout); a_mn = A stored [K rows, M cols], reduction also runs over k_batches batches of both operands
   * (wgrad dW[m,n] = sum_{b,t} dY[b,t,m] X[b,t,n], k_per_tap = rows per batch); epi 5 = out_f32 += acc (atomic). */
  int32_t a_mn, b_mn, k_batches;
  int64_t w_row_stride;   /* MN-major W: elements between rows (0 = w_cols) */
  int64_t w_batch_stride; /* MN-major W: elements between k-batches */
  int32_t k_splits;       /* 0 = auto (epi 5 only), 1 = none */
  int32_t b_tap_atoms;    /* grouped-conv wgrad: N atom j = W columns [g*w_group_off, +64) at rows k + j - a_pad */
  int32_t out_group_rows; /* output row offset per group */
  int32_t out_group_cols; /* output column offset per group (default N when out_group_rows == 0) */
  /* 16-bit storage formats (B2S_FMT_*) of A, of W and of the 16-bit outputs (out / out2). A and W must agree (a mixed
   * bf16 x fp16 tcgen05.mma traps on B200); the output format is free. */
  int32_t a_fmt, w_fmt, out_fmt;
} b2s_gemm_args;
int b2s_gemm_bf16(const b2s_gemm_args* args, void* stream);
/* measurement hook (bench.py roofline leg): while enabled, every GEMM launch is bracketed by CUDA events on its
 * stream; read() synchronises them and returns the summed device time (host pointers) and the launch count. */
void b2s_gemm_timing_enable(int32_t on);
int b2s_gemm_timing_read(double* total_ms, long long* launches);
/* launch `index` of the instrumented window: duration and shape[10] = M, N, K (whole reduction), batches, groups,
 * epilogue id, activation id, kernel mode (0 fwd / 1 train-fwd / 2 dgrad / 3 wgrad), block_n, cta_group */
int b2s_gemm_timing_get(int64_t index, double* ms, int32_t* shape);

/* ---------------------------------------------------------------------------------------------
 * Fused CE + logit-KD loss over packed response rows.  Replaces utils.soft_cross_entropy
 * (REF/utils.py:167-178, call site REF/trainer.py:349-352) and the per-sample CrossEntropyLoss of
 * AudioLlamaForCausalLM.forward (REF/model/audio_llama.py:72-101).
 *   student/teacher: bf16 [rows, V] (leading dims lds/ldt); labels[row] = target id for the CE term or -1;
 *   row_offsets: int32 [utterances+1] segment boundaries; outputs per utterance: loss_ld, loss_ntp (means);
 *   lse_s/lse_t/coef_kd/coef_ce: fp32 [rows] saved for the backward (coef = scale / count). */
size_t b2s_kd_ce_workspace_bytes(int32_t rows, int32_t V);
int b2s_kd_ce_loss_fwd(const void* student, const void* teacher, int64_t lds, int64_t ldt, int32_t rows, int32_t V,
                       const int32_t* labels, const int32_t* row_offsets, int32_t utterances, float scale_kd,
                       float scale_ce, void* workspace, float* lse_s, float* lse_t, float* coef_kd, float* coef_ce,
                       float* loss_ld, float* loss_ntp, void* stream);
/* d(sum_u scale_kd*ld_u + scale_ce*ntp_u)/d student, 16-bit [rows, V] in format fmt, multiplied by *loss_scale (device
 * scalar, may be NULL = 1: the GradScaler scale, REF/trainer.py:374); the teacher gets no gradient (.detach(),
 * REF/trainer.py:351). The logits themselves are bf16. */
int b2s_kd_ce_loss_bwd(const void* student, const void* teacher, int64_t lds, int64_t ldt, int32_t rows, int32_t V,
                       const int32_t* labels, const float* lse_s, const float* lse_t, const float* coef_kd,
                       const float* coef_ce, const float* loss_scale, void* d_student, int64_t ldd, int32_t fmt,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Normalisation family (TF/models/hubert/modeling_hubert.py:127-151,216-231,505-548,613;
 * TF/models/llama/modeling_llama.py:53-67; REF/model/audio_encoder.py:59-63). */
int b2s_layernorm_fwd(const void* x, int32_t in_bf16, const float* gamma, const float* beta, float eps,
                      int32_t act_gelu, void* y_bf16, int64_t rows, int32_t C, int32_t fmt, void* stream);
int b2s_rmsnorm_fwd(const float* x, const float* w, float eps, void* y_bf16, int64_t rows, int32_t C, int32_t fmt, void* stream);
int b2s_rmsnorm_gather_fwd(const float* x, const int32_t* row_index, const float* w, float eps, void* y_bf16,
                           int64_t rows, int32_t C, int32_t fmt, void* stream);
int b2s_layernorm_avgpool_fwd(const float* x, const float* gamma, const float* beta, float eps, void* y_bf16,
                              int32_t batches, int32_t frames, int32_t C, int32_t kernel, int32_t stride,
                              int32_t out_frames, int32_t fmt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Remaining memory-bound ops. */
/* HuBERT feature-extractor layer 0 (TF/models/hubert/modeling_hubert.py:127-151, layer_id 0) */
int b2s_conv0_ln_gelu_fwd(const float* wave, int64_t wave_stride, int32_t batches, int32_t samples, const float* w,
                          const float* bias, const float* gamma, const float* beta, float eps, void* y_bf16,
                          int32_t out_frames, int32_t fmt, void* stream);
/* embedding gather + audio splice (REF/utils.py:27-46,49-73,85-164; REF/inference.py:113-134):
 * h0[row] = row_src[row] >= 0 ? embed_tokens[row_src[row]] : audio_embeds[-(row_src[row]+1)];
 * row_src[row] == INT32_MIN writes a zero row (the reference's left padding). */
int b2s_embed_splice_fwd(const void* embed_table_bf16, const float* audio_embeds, const int32_t* row_src, float* h0,
                         int64_t rows, int32_t C, int32_t fmt, void* stream);
/* per-row-pair sum of squared differences, the core of the FD MSE (REF/trainer.py:358-370) */
int b2s_rowpair_sqdiff_fwd(const float* h, const int32_t* rows_a, const int32_t* rows_b, float* out, int32_t pairs,
                           int32_t C, void* stream);
/* weight-norm(dim=2) + K-major repack of the positional conv weight (TF/models/hubert/modeling_hubert.py:45-92) */
int b2s_posconv_weight_pack(const float* g, const float* v, void* w_packed_bf16, int32_t cout, int32_t cin_g,
                            int32_t k, int32_t fmt, void* stream);
int b2s_cast_f32_to_h16(const float* x, void* y, int64_t n, int32_t fmt, void* stream);
int b2s_cast_h16_to_f32(const void* x, float* y, int64_t n, int32_t fmt, void* stream);

/* Packed varlen attention forward (TF/models/hubert/modeling_hubert.py:262-345;
 * TF/models/llama/modeling_llama.py:225-289). */
int b2s_attention_fwd(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                      const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, int64_t total_rows, int32_t Hq,
                      int32_t Hkv, int32_t D, float scale, int32_t causal, float* lse /* optional [rows, Hq] */,
                      int32_t fmt, void* stream);
/* The same with a SHARED PREFIX (causal only): sequence 0 = rows [cu[0], cu[1]) holds the shared_prefix_len prompt-prefix
 * rows once; every other sequence holds only its own rows and additionally attends to the keys / values of sequence 0
 * (all of them: own positions start at shared_prefix_len). Equivalent to b2s_attention_fwd on sequences that each carry
 * their own copy of the prefix (REF/utils.py:27-46 builds them that way); the copies' rows are simply not computed. */
int b2s_attention_fwd_prefix(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                             const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, int64_t total_rows,
                             int32_t Hq, int32_t Hkv, int32_t D, float scale, float* lse, int32_t fmt,
                             int32_t shared_prefix_len, void* stream);
/* Backward of b2s_attention_fwd (autograd through HubertAttention / LlamaAttention, REF/trainer.py:373-374).
 * lse: the forward's saved log-sum-exp; delta_ws: fp32 [rows, Hq] scratch; dq/dk/dv: bf16, row stride ld_dqkv;
 * rope_cs (optional): fuses the inverse rotary rotation into the dq / dk stores. */
int b2s_attention_bwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const void* o, int64_t ld_o,
                      const void* dout, int64_t ld_do, const float* lse, float* delta_ws, void* dq, void* dk, void* dv,
                      int64_t ld_dqkv, const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen,
                      int64_t total_rows, int32_t Hq, int32_t Hkv, int32_t D, float scale, int32_t causal,
                      const float* rope_cs, int32_t fmt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-model entry points (the layer loops run in C++, one call per forward). */
typedef struct {
  const float *ln1_g, *ln1_b;
  const void* wqkv;  /* bf16 [3H, H]: q | k | v rows */
  const float* bqkv; /* fp32 [3H] (k bias zero for Whisper) */
  const void* wo;
  const float* bo;
  const float *ln2_g, *ln2_b;
  const void* w1; /* bf16 [F, H] */
  const float* b1;
  const void* w2; /* bf16 [H, F] */
  const float* b2;
} b2s_encoder_layer;

typedef struct {
  /* conv feature extractor: layer 0 on CUDA cores, layers 1..6 as implicit GEMM */
  const float *conv0_w, *conv0_b, *conv0_ln_g, *conv0_ln_b;
  const void* conv_w[6]; /* bf16 [512, k*512], column = tap*512 + c_in */
  const float* conv_b[6];
  const float* conv_ln_g[6];
  const float* conv_ln_b[6];
  int32_t conv_k[6];
  int32_t conv_stride[6];
  /* feature projection */
  const float *fp_ln_g, *fp_ln_b;
  const void* fp_w; /* bf16 [H, 512] */
  const float* fp_b;
  /* positional conv (weight-normed, packed by b2s_posconv_weight_pack) */
  const void* pos_w; /* bf16 [H][K=128][H/groups] */
  const float* pos_b;
  int32_t pos_k, pos_groups;
  /* transformer */
  const b2s_encoder_layer* layers; /* host array */
  int32_t num_layers, hidden, heads, ffn;
  const float *final_ln_g, *final_ln_b;
  float ln_eps;
  /* AudioEncoder pooling + projector (REF/model/audio_encoder.py:34-42,59-63,87) */
  int32_t pool_kernel, pool_stride;
  const void* proj_w; /* bf16 [llm_dim, H] */
  const float* proj_b;
  int32_t llm_dim;
  int32_t fmt; /* B2S_FMT_*: format of every 16-bit weight above ("bf16" in the comments = "16-bit in fmt") and of every
                  16-bit activation / gradient buffer the entry points taking this struct read or write */
} b2s_hubert_weights;

/* frames produced by the conv stack for `samples` input samples; pooled = AvgPool1d output length */
int b2s_hubert_num_frames(const b2s_hubert_weights* w, int32_t samples, int32_t* frames, int32_t* pooled);
size_t b2s_hubert_workspace_bytes(const b2s_hubert_weights* w, int32_t batches, int32_t samples);
/* AudioEncoder.forward for the HuBERT + "pool" configuration (REF/model/audio_encoder.py:56-88 over
 * TF/models/hubert/modeling_hubert.py:889-958), eval mode. wave: fp32 [batches, samples] (row stride
 * wave_stride). audio_embeds: fp32 [batches*pooled, llm_dim]. last_hidden (optional, may be NULL):
 * fp32 [batches*frames, hidden] pre-final-LN residual stream, for parity tests. */
int b2s_hubert_forward(const b2s_hubert_weights* w, const float* wave, int64_t wave_stride, int32_t batches,
                       int32_t samples, void* workspace, size_t workspace_bytes, float* audio_embeds,
                       float* last_hidden, void* stream);

/* Whisper encoder variant of AudioEncoder.forward (REF/model/audio_encoder.py:10-13,56-88 over WhisperEncoder.forward,
 * TF/models/whisper/modeling_whisper.py:593-647), eval mode. */
typedef struct {
  const void* conv1_w; /* bf16 [H, 3*mel], column = tap*mel + c_in */
  const float* conv1_b;
  const void* conv2_w; /* bf16 [H, 3*H] */
  const float* conv2_b;
  const float* pos_emb; /* fp32 [max_positions, H] (embed_positions.weight) */
  const b2s_encoder_layer* layers; /* host array; k_proj has no bias -> zeros in bqkv */
  int32_t num_layers, hidden, heads, ffn, mel_bins, max_positions;
  const float *final_ln_g, *final_ln_b;
  float ln_eps;
  int32_t pool_kernel, pool_stride;
  const void* proj_w;
  const float* proj_b;
  int32_t llm_dim;
  int32_t fmt; /* B2S_FMT_*, as in b2s_hubert_weights */
} b2s_whisper_weights;

size_t b2s_whisper_workspace_bytes(const b2s_whisper_weights* w, int32_t batches);
/* mel: fp32 [batches, mel_bins, frames_in] (frames_in must be 2*max_positions, like the reference);
 * audio_embeds: fp32 [batches * pooled, llm_dim], pooled = (max_positions - pool_kernel)/pool_stride + 1;
 * last_hidden (optional): fp32 [batches * max_positions, hidden] pre-final-LN residual stream. */
int b2s_whisper_forward(const b2s_whisper_weights* w, const float* mel, int32_t batches, int32_t frames_in,
                        void* workspace, size_t workspace_bytes, float* audio_embeds, float* last_hidden,
                        void* stream);

typedef struct {
  const float* ln1_w;
  const void* wqkv; /* bf16 [(Hq+2Hkv)*D, H] */
  const void* wo;   /* bf16 [H, Hq*D] */
  const float* ln2_w;
  const void* wgu;  /* bf16 [2F, H], 64 gate rows | 64 up rows interleaved */
  const void* wd;   /* bf16 [H, F] */
} b2s_llama_layer;

typedef struct {
  const b2s_llama_layer* layers; /* host array */
  int32_t num_layers, hidden, heads, kv_heads, head_dim, ffn, vocab;
  float rms_eps;
  const float* final_norm_w;
  const void* lm_head;   /* bf16 [vocab, H] */
  const float* rope_cs;  /* fp32 [max_pos, head_dim]: cos[0:D/2] | sin[0:D/2] */
  int32_t max_pos;
  int32_t fmt; /* B2S_FMT_*: format of the weights, of the embedding table handed to decode, of the KV cache and of
                  every 16-bit activation / gradient buffer (logits stay bf16: the fused loss reads them once, their
                  2^-9 rounding is 1e-3 of the 2e-2 budget) */
} b2s_llama_weights;

size_t b2s_llama_workspace_bytes(const b2s_llama_weights* w, int32_t rows, int32_t logit_rows);
/* LlamaModel.forward + lm_head over a packed batch of sequences (REF/model/audio_llama.py:49-67 over
 * TF/models/llama/modeling_llama.py:375-425), eval mode, causal per sequence.
 *   h: fp32 [rows, hidden] in: inputs_embeds (spliced); out: the residual stream after the last layer (pre final
 *   norm) when all_hidden is requested or no / all logits are; otherwise the INPUT of the last layer (the rest of the
 *   last layer then runs on the logit rows only, in the workspace);
 *   cu_seqlens int32 [num_seqs+1]; positions int32 [rows];
 *   logit_rows_index int32 [logit_rows]: rows whose logits are produced -> logits bf16 [logit_rows, vocab];
 *   fd taps (optional): for each t < num_taps, before layer tap_layers[t] runs, out
 *   fd_sq[t*pairs + i] = sum_c (h[tap_rows_a[i], c] - h[tap_rows_b[i], c])^2   (REF/trainer.py:358-370);
 *   all_hidden (optional, may be NULL): fp32 [num_layers+1, rows, hidden], output_hidden_states=True layout
 *   ([l] = input of layer l, [num_layers] = output of the final norm). */
int b2s_llama_prefill(const b2s_llama_weights* w, float* h, int32_t rows, const int32_t* cu_seqlens,
                      int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                      const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16,
                      const int32_t* tap_layers /*host*/, int32_t num_taps, const int32_t* tap_rows_a,
                      const int32_t* tap_rows_b, int32_t pairs, float* fd_sq, float* all_hidden, void* workspace,
                      size_t workspace_bytes, void* stream);
/* b2s_llama_prefill with the prompt prefix SHARED between the sequences (see b2s_attention_fwd_prefix): sequence 0 of
 * cu_seqlens is the shared_prefix_len prefix rows (positions 0 ..), every other sequence holds its own rows only
 * (positions from shared_prefix_len on). Under the causal mask the prefix rows are identical in every sequence the
 * reference builds (REF/utils.py:27-46), so they are computed once: 5.6 % fewer LLM rows at 32 utterances. */
int b2s_llama_prefill_prefix(const b2s_llama_weights* w, float* h, int32_t rows, const int32_t* cu_seqlens,
                             int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                             const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16,
                             const int32_t* tap_layers /*host*/, int32_t num_taps, const int32_t* tap_rows_a,
                             const int32_t* tap_rows_b, int32_t pairs, float* fd_sq, float* all_hidden, void* workspace,
                             size_t workspace_bytes, int32_t shared_prefix_len, void* stream);

/* ---- Whisper log-mel features on the GPU (row f2): replaces transformers' WhisperFeatureExtractor call in the
 * reference's collate function (REF/trainer.py:178-182; TF/models/whisper/feature_extraction_whisper.py:105-133).
 * wave fp32 [batches, samples] (row stride wave_stride), mel_filters fp32 [201, 80] (slaney, as the extractor holds
 * them), out fp32 [batches, 80, frames] with frames = samples / 160, max_ws int32 [batches] scratch. */
int b2s_whisper_log_mel(const float* wave, int64_t wave_stride, int32_t batches, int32_t samples,
                        const float* mel_filters, float* out, int32_t frames, int32_t* max_ws, void* stream);

/* ---- greedy decode with a KV cache (REF/inference.py:55-74, REF/trainer.py:530-545: HF generate after the prefill).
 * Cache: bf16 [layers][slots][2*kv_heads*head_dim], one row = k (post-RoPE) | v of one token.
 * b2s_llama_prefill_kv = b2s_llama_prefill that also stores packed row r of every layer into slot kv_slot_of_row[r]
 * (< 0 = skip). b2s_llama_decode_step consumes ONE new token per sequence: its k|v are appended at slot
 * seq_start[b] + seq_len[b] (seq_len = tokens already cached = the new token's position) and logits bf16 [batch, vocab]
 * come back; the caller advances seq_len. */
size_t b2s_llama_kv_cache_bytes(const b2s_llama_weights* w, int32_t slots);
int b2s_llama_prefill_kv(const b2s_llama_weights* w, float* h, int32_t rows, const int32_t* cu_seqlens,
                         int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                         const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16, void* kv_cache,
                         int32_t kv_slots, const int32_t* kv_slot_of_row, void* workspace, size_t workspace_bytes,
                         void* stream);
size_t b2s_llama_decode_workspace_bytes(const b2s_llama_weights* w, int32_t batch);
int b2s_llama_decode_step(const b2s_llama_weights* w, const void* embed_table_bf16, const int32_t* token_ids,
                          int32_t batch, void* kv_cache, int32_t kv_slots, const int32_t* seq_start,
                          const int32_t* seq_len, void* logits_bf16, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training step (REF/trainer.py:270-384): forward with saved activations + backward of the frozen LLM
 * (data gradients only: REF/trainer.py:62-64), memory-bound backward kernels, AdamW. */
typedef struct {
  const void* wqkv_t; /* bf16 [H, (Hq+2Hkv)*D]  = wqkv^T */
  const void* wo_t;   /* bf16 [Hq*D, H]         = wo^T   */
  const void* wgu_t;  /* bf16 [H, 2F]           = wgu^T (packed gate|up order) */
  const void* wd_t;   /* bf16 [F, H]            = wd^T   */
} b2s_llama_layer_t;
typedef struct {
  const b2s_llama_layer_t* layers; /* host array */
  const void* lm_head_t;           /* bf16 [H, vocab] */
} b2s_llama_weights_t;
/* per-layer activations kept by the training forward (caller-allocated, rows = all packed rows):
 *   h [L+1][rows][H] fp32 (h[0] = spliced input on entry, h[l] = input of layer l, h[L] = final stream),
 *   h_mid [L][rows][H] fp32, qkv [L][rows][(Hq+2Hkv)D] bf16 (post-RoPE), ao [L][rows][Hq*D] bf16,
 *   lse [L][rows][Hq] fp32, gu [L][rows][2F] bf16 (raw gate|up). */
typedef struct {
  float* h;
  float* h_mid;
  void* qkv;
  void* ao;
  float* lse;
  void* gu;
} b2s_llama_saved;
size_t b2s_llama_train_workspace_bytes(const b2s_llama_weights* w, int32_t rows, int32_t logit_rows);
size_t b2s_llama_backward_workspace_bytes(const b2s_llama_weights* w, int32_t rows_bwd, int32_t n_dl);
int b2s_llama_forward_train(const b2s_llama_weights* w, const b2s_llama_saved* saved, int32_t rows,
                            const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                            const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16,
                            const int32_t* tap_layers /*host*/, int32_t num_taps, const int32_t* tap_rows_a,
                            const int32_t* tap_rows_b, int32_t pairs, float* fd_sq, void* workspace,
                            size_t workspace_bytes, void* stream);
/* Gradient w.r.t. the LLM input rows [0, rows_bwd) (the student sequences, packed first):
 *   d_logits 16-bit (w->fmt) [n_dl, vocab] for rows dl_rows_index; FD term: before layer l runs backward, rows
 *   tap_rows_a get tap_coef[i] * (*loss_scale) * (h[l+1][a_i] - h[l+1][b_i]) for every tap layer l+1 (loss_scale: device
 *   scalar or NULL = 1, the same one b2s_kd_ce_loss_bwd was given); dh fp32 [rows_bwd, H] = dL/d h[0] (scaled). */
int b2s_llama_backward(const b2s_llama_weights* w, const b2s_llama_weights_t* wt, const b2s_llama_saved* saved,
                       int32_t rows, int32_t rows_bwd, const int32_t* cu_seqlens, int32_t num_seqs_bwd,
                       int32_t max_seqlen, const void* d_logits, const int32_t* dl_rows_index, int32_t n_dl,
                       const int32_t* tap_layers /*host*/, int32_t num_taps, const int32_t* tap_rows_a,
                       const int32_t* tap_rows_b, const float* tap_coef, const float* loss_scale, int32_t pairs,
                       float* dh, void* workspace, size_t workspace_bytes, void* stream);
/* ---- trainable HuBERT audio encoder (REF/trainer.py:98-105: every AudioEncoder parameter is optimised).
 * Gradient accumulators: fp32, zeroed by the caller, shaped like the packed tensors of b2s_hubert_weights
 * (conv_w[i] [512, k*512], wqkv [3H, H], pos_w [H][k][H/groups] = gradient w.r.t. the weight-normed effective
 * weight; the caller applies the weight-norm chain rule once per optimizer step). Accumulated with += so
 * gradient accumulation over micro-batches (REF/trainer.py:372-380) needs no extra pass. */
typedef struct {
  float *ln1_g, *ln1_b, *wqkv, *bqkv, *wo, *bo, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
} b2s_encoder_layer_grads;
typedef struct {
  float *conv0_w, *conv0_b, *conv0_ln_g, *conv0_ln_b;
  float* conv_w[6];
  float* conv_b[6];
  float* conv_ln_g[6];
  float* conv_ln_b[6];
  float *fp_ln_g, *fp_ln_b, *fp_w, *fp_b;
  float *pos_w, *pos_b;
  const b2s_encoder_layer_grads* layers; /* host array */
  float *final_ln_g, *final_ln_b, *proj_w, *proj_b;
} b2s_hubert_grads;
size_t b2s_hubert_saved_bytes(const b2s_hubert_weights* w, int32_t batches, int32_t samples);
size_t b2s_hubert_backward_workspace_bytes(const b2s_hubert_weights* w, int32_t batches, int32_t samples);
/* Train-mode regularisers of HubertModel (REF/trainer.py:258 puts the encoder in .train()): the nn.Dropout sites of
 * TF/models/hubert/modeling_hubert.py (:223-230 feature projection, :585-587 after hidden + positional conv, :254
 * attention probabilities, :383-393 / :536 after the attention block, :351-368 FFN activation and output), LayerDrop
 * (:596-599) and SpecAugment time masking (:842-886). Keep / drop decisions are a pure function of
 * (seed, site, element index) -- csrc/rng.cuh, restated in oracle/regularizers.py -- so forward and backward take the
 * SAME block and no mask is stored. LayerDrop decisions and the SpecAugment frame mask are drawn by the caller (the
 * reference draws both on the host: torch.rand([]) per layer, numpy in _compute_mask_indices). NULL = eval behaviour. */
typedef struct {
  uint64_t seed;                  /* one value per micro-batch */
  float p_feat_proj;              /* config.feat_proj_dropout */
  float p_hidden;                 /* config.hidden_dropout */
  float p_attention;              /* config.attention_dropout */
  float p_activation;             /* config.activation_dropout */
  const uint8_t* layer_skip;      /* HOST [num_layers], 1 = LayerDrop skips the layer; NULL = none */
  const uint8_t* time_mask;       /* DEVICE [batches*frames], 1 = frame replaced by masked_spec_embed; NULL = none */
  const float* masked_spec_embed; /* DEVICE fp32 [hidden] (required with time_mask) */
  float* g_masked_spec_embed;     /* DEVICE fp32 [hidden] gradient accumulator (backward only; may be NULL) */
} b2s_encoder_regularizers;
/* b2s_hubert_forward keeping every activation the backward needs in `saved`; reg = NULL: deterministic (eval) math.
 * samples_per_utt (HOST int32 [batches], NULL = every utterance has `samples` samples): a ragged micro-batch. The
 * waveforms are zero-padded on the right to `samples` (the reference's collate, REF/trainer.py:146-149), every
 * utterance still gets exactly the numbers it would get alone: the front end runs on the padded layout, the
 * transformer stack on the packed valid frames with per-utterance attention, the pooling windows per utterance.
 * audio_embeds is [batches, pooled(samples), llm_dim]; rows >= pooled(samples_per_utt[b]) of utterance b are padding. */
int b2s_hubert_forward_train(const b2s_hubert_weights* w, const float* wave, int64_t wave_stride, int32_t batches,
                             int32_t samples, const int32_t* samples_per_utt, void* saved, size_t saved_bytes,
                             float* audio_embeds, const b2s_encoder_regularizers* reg, void* stream);
/* test hook: out[i] = 1 if element e_first + i of stream (seed, site, a, b) is KEPT at drop probability p */
int b2s_drop_mask_dump(uint8_t* out, int64_t n, uint64_t seed, uint32_t site, uint32_t a, uint32_t b, float p,
                       uint32_t e_first, void* stream);
/* d_audio_embeds fp32 [batches*pooled, llm_dim] -> grads (+=). pos_w_dgrad: bf16 [H][k][H/groups], the packed
 * positional-conv weight with taps reversed and each (out, in) block transposed (the conv's transpose).
 * samples_per_utt: as given to the forward; the padding rows of d_audio_embeds must be zero.
 * layer_done_events (HOST array of num_layers cudaEvent_t, or NULL): event l is recorded on `stream` once every kernel
 * that writes a gradient of transformer layer l has been enqueued (layers run last to first), so a communication
 * stream can all-reduce layer l's gradients under the backward of layers l-1 ... 0. */
int b2s_hubert_backward(const b2s_hubert_weights* w, const void* pos_w_dgrad, const b2s_hubert_grads* grads,
                        const float* wave, int64_t wave_stride, int32_t batches, int32_t samples,
                        const int32_t* samples_per_utt, void* saved,
                        size_t saved_bytes, const float* d_audio_embeds, void* workspace, size_t workspace_bytes,
                        const b2s_encoder_regularizers* reg /* the block the forward ran with */,
                        void* const* layer_done_events, void* stream);
/* Whisper encoder (REF/config/llama3_whisper.yaml trains it like the HuBERT one): same contract; mel fp32
 * [batches, mel_bins, 2*max_positions]; conv weights' gradients in the packed [H, 3*C_in] layout; the sinusoid table is
 * frozen; the k_proj slot of bqkv's gradient has no parameter behind it. */
typedef struct {
  float *conv1_w, *conv1_b, *conv2_w, *conv2_b;
  const b2s_encoder_layer_grads* layers; /* host array */
  float *final_ln_g, *final_ln_b, *proj_w, *proj_b;
} b2s_whisper_grads;
size_t b2s_whisper_saved_bytes(const b2s_whisper_weights* w, int32_t batches);
size_t b2s_whisper_backward_workspace_bytes(const b2s_whisper_weights* w, int32_t batches);
int b2s_whisper_forward_train(const b2s_whisper_weights* w, const float* mel, int32_t batches, int32_t frames_in,
                              void* saved, size_t saved_bytes, float* audio_embeds, void* stream);
int b2s_whisper_backward(const b2s_whisper_weights* w, const b2s_whisper_grads* grads, int32_t batches, void* saved,
                         size_t saved_bytes, const float* d_audio_embeds, void* workspace, size_t workspace_bytes,
                         void* const* layer_done_events, void* stream);
/* memory-bound backward kernels of the encoder (see csrc/backward_enc.cu).
 * layernorm_bwd_ex: dh (+)= LN'(x) dy [optionally through an erf-GELU], dgamma / dbeta += ...; dh_colsum (optional, [C])
 * += the column sums of the dh it writes -- the bias gradient of the linear layer underneath, so that the transformer
 * backward does not re-read dh for it (REF: torch autograd of nn.LayerNorm / nn.Linear.bias in HubertEncoderLayer). */
int b2s_layernorm_bwd_ex(const void* x, int32_t x_bf16, const float* gamma, const float* beta, int32_t act_gelu,
                         float eps, const void* dy, int32_t dy_bf16, float* dh, int32_t accumulate, void* dx_bf16,
                         float* dgamma, float* dbeta, int64_t rows, int32_t C, int32_t fmt, float* dh_colsum,
                         void* stream);
int b2s_colsum_accum(const void* x, int32_t x_bf16, float* out, int64_t rows, int32_t C, int32_t fmt, void* stream);
int b2s_avgpool_bwd(const float* dpooled, float* dx, int32_t batches, int32_t frames, int32_t C, int32_t kernel,
                    int32_t stride, int32_t pooled, void* stream);
int b2s_col2im_add(const void* dcol_bf16, void* dx_bf16, int32_t batches, int32_t tin, int32_t tout, int32_t k,
                   int32_t s, int32_t C, int32_t fmt, void* stream);
int b2s_conv0_bwd(const float* wave, int64_t wave_stride, int32_t batches, int32_t samples, const float* w,
                  const float* bias, const float* gamma, const float* beta, float eps, const void* dy_bf16,
                  int32_t frames, float* dW, float* db, float* dgamma, float* dbeta, int32_t fmt, void* stream);
int b2s_rmsnorm_bwd(const float* x, const int32_t* x_index, const float* w, float eps, const float* dy, float* dh,
                    const int32_t* dh_index, void* dh_bf16, int64_t rows, int32_t C, int32_t fmt, void* stream);
int b2s_layernorm_bwd(const float* x, const float* gamma, float eps, const void* dy, int32_t dy_bf16, float* dh,
                      int32_t accumulate, void* dh_bf16, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                      int32_t fmt, void* stream);
int b2s_swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t rows, int32_t F, int32_t fmt, void* stream);
/* dpre = dy * gelu'(pre) (erf form). colsum (optional, with the row width F, F % 2048 == 0): += column sums of dpre, the
 * bias gradient of the Linear that produced `pre`, accumulated by the same pass. */
int b2s_gelu_bwd(const void* pre, const void* dy, void* dpre, int64_t n, int32_t fmt, float* colsum, int32_t F,
                 void* stream);
int b2s_gather_rows_f32(const float* src, const int32_t* index, float* out, int64_t rows, int32_t C, void* stream);
/* Dynamic loss scaling, torch.cuda.amp.GradScaler semantics (REF/trainer.py:252,374,381-382), with its state in DEVICE
 * memory so that neither the skip-on-overflow decision nor the scale update synchronises the host:
 *   forward/backward: the loss gradient is multiplied by state->scale where it enters the backward pass
 *                     (b2s_kd_ce_loss_bwd / b2s_llama_backward take &state->scale);
 *   optimizer step  : b2s_nonfinite_check(flat gradient) -> b2s_adamw_step(..., state) -> b2s_grad_scaler_update. */
typedef struct {
  float scale;            /* current loss scale (GradScaler init_scale = 65536) */
  int32_t growth_tracker; /* clean optimizer steps since the last scale change */
  int32_t found_inf;      /* set by b2s_nonfinite_check, cleared by b2s_grad_scaler_update */
  int32_t opt_steps;      /* optimizer steps actually taken (AdamW bias-correction step) */
  int32_t skipped_steps;  /* steps skipped on overflow */
  int32_t reserved[3];
} b2s_grad_scaler_state;
/* state->found_inf |= (any element of g is inf / nan); g: 16-byte aligned, n % 4 == 0 */
int b2s_nonfinite_check(const float* g, int64_t n, b2s_grad_scaler_state* state, void* stream);
/* GradScaler.update(): on overflow scale *= backoff_factor (and the step counts as skipped), else after growth_interval
 * consecutive clean steps scale *= growth_factor; clears found_inf */
int b2s_grad_scaler_update(b2s_grad_scaler_state* state, float growth_factor, float backoff_factor,
                           int32_t growth_interval, void* stream);
/* torch.optim.AdamW step on one flat fp32 tensor (REF/trainer.py:98-105,381); grad_scale multiplies the gradient.
 * scaler (device, may be NULL): the gradient is additionally divided by scaler->scale, the whole step is a no-op when
 * scaler->found_inf is set (GradScaler.step), and the bias corrections use scaler->opt_steps + 1 instead of `step`. */
int b2s_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int32_t step, float grad_scale, const b2s_grad_scaler_state* scaler,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H_ */
