// api.cu -- the extern "C" surface declared in include/b2s.h, plus status/error plumbing.
#include <atomic>
#include <cstdarg>
#include <new>
#include <cstdio>
#include <cstdlib>

#include "../../include/b2s.h"
#include "b2s_common.cuh"
#include "gemm_sm100.cuh"
#include "ops.cuh"

namespace b2s {

namespace {
thread_local char g_err[1024] = "";
thread_local Context* g_current = nullptr;
Context& default_context() {
  static Context c;
  return c;
}
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
}  // namespace

Context::Context() {
  pdl = env_int("B2S_PDL", 1) != 0;
  resid_red = env_int("B2S_RESID_RED", 1) != 0;
  tma_epi = env_int("B2S_TMA_EPI", 1) != 0;
  if (const char* e = getenv("B2S_ATTN_CFG")) sscanf(e, "%d,%d", &attn_bn, &attn_kvs);
  gemm_group_m = env_int("B2S_GEMM_GROUP_M", 0);
  gemm_tail_split = env_int("B2S_GEMM_TAIL_SPLIT", 1) != 0;
  gemm_epi8 = env_int("B2S_GEMM_EPI8", 1);
}

Context& ctx() { return g_current != nullptr ? *g_current : default_context(); }

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_last_error("CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e), file, line, what);
  return B2S_ERR_CUDA;
}

void count_launch() { ctx().launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return ctx().launches.load(std::memory_order_relaxed); }

bool pdl_enabled() { return ctx().pdl != 0; }

void set_sm_budget(int sms) { ctx().sm_budget = sms > 0 ? sms : 0; }
int sm_budget() { return ctx().sm_budget; }

// SMs the persistent kernels size their grids for: the device's count, or the current context's budget when one is set
// (a persistent grid with static tile assignment must not count on SMs a concurrent communication kernel occupies)
int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
      sms = 148;
    }
  }
  const int b = ctx().sm_budget;
  return (b > 0 && b < sms) ? b : sms;
}

int comm_destroy(Context& c);  // comm.cu

// models.cu
int hubert_num_frames(const b2s_hubert_weights* w, int samples, int* frames, int* pooled);
size_t hubert_workspace_bytes(const b2s_hubert_weights* w, int batches, int samples);
int hubert_forward(const b2s_hubert_weights* w, const float* wave, long long wave_stride, int batches, int samples,
                   void* workspace, size_t workspace_bytes, float* audio_embeds, float* last_hidden,
                   cudaStream_t stream);
size_t whisper_workspace_bytes(const b2s_whisper_weights* w, int batches);
int whisper_forward(const b2s_whisper_weights* w, const float* mel, int batches, int frames_in, void* workspace,
                    size_t workspace_bytes, float* audio_embeds, float* last_hidden, cudaStream_t stream);
size_t llama_workspace_bytes(const b2s_llama_weights* w, int rows, int logit_rows);
int llama_prefill(const b2s_llama_weights* w, float* h, int rows, const int* cu_seqlens, int num_seqs, int max_seqlen,
                  const int* positions, const int* logit_rows_index, int logit_rows, void* logits_bf16,
                  const int* tap_layers, int num_taps, const int* tap_rows_a, const int* tap_rows_b, int pairs,
                  float* fd_sq, float* all_hidden, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                  int shared_prefix_len = 0);

size_t llama_train_workspace_bytes(const b2s_llama_weights* w, int rows, int logit_rows);
size_t whisper_saved_bytes(const b2s_whisper_weights* w, int batches);
size_t whisper_backward_workspace_bytes(const b2s_whisper_weights* w, int batches);
int whisper_forward_train(const b2s_whisper_weights* w, const float* mel, int batches, int frames_in, void* saved,
                          size_t saved_bytes, float* audio_embeds, cudaStream_t stream);
int whisper_backward(const b2s_whisper_weights* w, const b2s_whisper_grads* gr, int batches, void* saved,
                     size_t saved_bytes, const float* d_audio_embeds, void* workspace, size_t workspace_bytes,
                     void* const* layer_done, cudaStream_t stream);
size_t llama_kv_cache_bytes(const b2s_llama_weights* w, int slots);
int llama_prefill_kv(const b2s_llama_weights* w, float* h, int rows, const int* cu_seqlens, int num_seqs,
                     int max_seqlen, const int* positions, const int* logit_rows_index, int logit_rows,
                     void* logits_bf16, const int* tap_layers, int num_taps, const int* tap_rows_a,
                     const int* tap_rows_b, int pairs, float* fd_sq, float* all_hidden, void* kv_cache, int kv_slots,
                     const int* kv_slot_of_row, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                     int shared_prefix_len = 0);
size_t llama_decode_workspace_bytes(const b2s_llama_weights* w, int batch);
int llama_decode_step(const b2s_llama_weights* w, const void* embed_table, const int* token_ids, int batch,
                      void* kv_cache, int kv_slots, const int* seq_start, const int* seq_len, void* logits_bf16,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t hubert_saved_bytes(const b2s_hubert_weights* w, int batches, int samples);
size_t hubert_backward_workspace_bytes(const b2s_hubert_weights* w, int batches, int samples);
int hubert_forward_train(const b2s_hubert_weights* w, const float* wave, long long wave_stride, int batches, int samples,
                         const int* samples_per_utt, void* saved, size_t saved_bytes, float* audio_embeds,
                         const b2s_encoder_regularizers* reg, cudaStream_t stream);
int hubert_backward(const b2s_hubert_weights* w, const void* pos_w_dgrad, const b2s_hubert_grads* gr, const float* wave,
                    long long wave_stride, int batches, int samples, const int* samples_per_utt, void* saved,
                    size_t saved_bytes, const float* d_audio_embeds, void* workspace, size_t workspace_bytes,
                    const b2s_encoder_regularizers* reg, void* const* layer_done, cudaStream_t stream);
size_t llama_backward_workspace_bytes(const b2s_llama_weights* w, int rows_bwd, int n_dl);
int llama_forward_train(const b2s_llama_weights* w, const b2s_llama_saved* sv, int rows, const int* cu_seqlens,
                        int num_seqs, int max_seqlen, const int* positions, const int* logit_rows_index,
                        int logit_rows, void* logits_bf16, const int* tap_layers, int num_taps, const int* tap_rows_a,
                        const int* tap_rows_b, int pairs, float* fd_sq, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream);
int llama_backward(const b2s_llama_weights* w, const b2s_llama_weights_t* wt, const b2s_llama_saved* sv, int rows,
                   int rows_bwd, const int* cu_seqlens, int num_seqs_bwd, int max_seqlen, const void* d_logits,
                   const int* dl_rows_index, int n_dl, const int* tap_layers, int num_taps, const int* tap_rows_a,
                   const int* tap_rows_b, const float* tap_coef, const float* loss_scale, int pairs, float* dh,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace b2s

using namespace b2s;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* b2s_last_error(void) { return g_err; }
int b2s_version(void) { return 1; }
long long b2s_launch_count(void) { return launch_count(); }
void b2s_set_sm_budget(int32_t sms) { set_sm_budget(sms); }
int32_t b2s_get_sm_budget(void) { return sm_budget(); }

struct b2s_handle {
  Context c;
};
int b2s_create(b2s_handle** out) {
  if (out == nullptr) {
    set_last_error("b2s_create: null output pointer");
    return B2S_ERR_INVALID;
  }
  *out = new (std::nothrow) b2s_handle();
  if (*out == nullptr) {
    set_last_error("b2s_create: out of memory");
    return B2S_ERR_INVALID;
  }
  return B2S_OK;
}
int b2s_destroy(b2s_handle* h) {
  if (h == nullptr) return B2S_OK;
  if (g_current == &h->c) g_current = nullptr;
  const int rc = comm_destroy(h->c);
  if (h->c.tail_flags != nullptr) cudaFree(h->c.tail_flags);
  for (auto& e : h->c.events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  delete h;
  return rc;
}
int b2s_make_current(b2s_handle* h) {
  g_current = h != nullptr ? &h->c : nullptr;
  return B2S_OK;
}
int b2s_set_option(int32_t option, int32_t value) {
  Context& c = ctx();
  switch (option) {
    case B2S_OPT_PDL: c.pdl = value != 0; break;
    case B2S_OPT_RESID_RED: c.resid_red = value != 0; break;
    case B2S_OPT_TMA_EPILOGUE: c.tma_epi = value != 0; break;
    case B2S_OPT_ATTN_KEYS_PER_STEP: c.attn_bn = value; break;
    case B2S_OPT_ATTN_KV_STAGES: c.attn_kvs = value; break;
    case B2S_OPT_SM_BUDGET: c.sm_budget = value > 0 ? value : 0; break;
    case B2S_OPT_GEMM_GROUP_M: c.gemm_group_m = value > 0 ? value : 0; break;
    case B2S_OPT_GEMM_TAIL_SPLIT: c.gemm_tail_split = value != 0; break;
    case B2S_OPT_GEMM_EPI8: c.gemm_epi8 = value < 0 ? 0 : (value > 2 ? 2 : value); break;
    default:
      set_last_error("b2s_set_option: unknown option %d", option);
      return B2S_ERR_INVALID;
  }
  return B2S_OK;
}
int b2s_get_option(int32_t option, int32_t* value) {
  Context& c = ctx();
  if (value == nullptr) {
    set_last_error("b2s_get_option: null output pointer");
    return B2S_ERR_INVALID;
  }
  switch (option) {
    case B2S_OPT_PDL: *value = c.pdl; break;
    case B2S_OPT_RESID_RED: *value = c.resid_red; break;
    case B2S_OPT_TMA_EPILOGUE: *value = c.tma_epi; break;
    case B2S_OPT_ATTN_KEYS_PER_STEP: *value = c.attn_bn; break;
    case B2S_OPT_ATTN_KV_STAGES: *value = c.attn_kvs; break;
    case B2S_OPT_SM_BUDGET: *value = c.sm_budget; break;
    case B2S_OPT_GEMM_GROUP_M: *value = c.gemm_group_m; break;
    case B2S_OPT_GEMM_TAIL_SPLIT: *value = c.gemm_tail_split; break;
    case B2S_OPT_GEMM_EPI8: *value = c.gemm_epi8; break;
    default:
      set_last_error("b2s_get_option: unknown option %d", option);
      return B2S_ERR_INVALID;
  }
  return B2S_OK;
}

int b2s_gemm_bf16(const b2s_gemm_args* a, void* stream) {
  if (a == nullptr) {
    set_last_error("b2s_gemm_bf16: null args");
    return B2S_ERR_INVALID;
  }
  GemmArgs g{};
  g.A = a->A;
  g.a_dim0 = a->a_dim0;
  g.a_row_stride = a->a_row_stride;
  g.a_batch_stride = a->a_batch_stride;
  g.a_rows = a->a_rows;
  g.W = a->W;
  g.w_rows = a->w_rows;
  g.w_cols = a->w_cols;
  g.M = a->M;
  g.N = a->N;
  g.batches = a->batches;
  g.groups = a->groups;
  g.taps = a->taps;
  g.k_per_tap = a->k_per_tap;
  g.a_pad = a->a_pad;
  g.a_group_off = a->a_group_off;
  g.w_group_off = a->w_group_off;
  g.epi = a->epi;
  g.act = a->act;
  g.bias = a->bias;
  g.out = a->out;
  g.ldo = a->ldo;
  g.out_batch_rows = a->out_batch_rows;
  g.resid = a->resid;
  g.resid_bcast = a->resid_bcast;
  g.out2 = a->out2;
  g.ld2 = a->ld2;
  g.rope_cs = a->rope_cs;
  g.positions = a->positions;
  g.rope_cols = a->rope_cols;
  g.block_n = a->block_n;
  g.cta_group = a->cta_group;
  g.a_mn = a->a_mn;
  g.b_mn = a->b_mn;
  g.k_batches = a->k_batches;
  g.w_row_stride = a->w_row_stride;
  g.w_batch_stride = a->w_batch_stride;
  g.k_splits = a->k_splits;
  g.b_tap_atoms = a->b_tap_atoms;
  g.out_group_rows = a->out_group_rows;
  g.out_group_cols = a->out_group_cols;
  g.a_fmt = a->a_fmt;
  g.w_fmt = a->w_fmt;
  g.out_fmt = a->out_fmt;
  return gemm_bf16_launch(g, S(stream));
}

void b2s_gemm_timing_enable(int32_t on) { gemm_timing_enable(on); }
int b2s_gemm_timing_read(double* total_ms, long long* launches) { return gemm_timing_read(total_ms, launches); }
int b2s_gemm_timing_get(int64_t index, double* ms, int32_t* shape) { return gemm_timing_get(index, ms, shape); }

size_t b2s_kd_ce_workspace_bytes(int32_t rows, int32_t V) { return kd_ce_workspace_bytes(rows, V); }

int b2s_kd_ce_loss_fwd(const void* student, const void* teacher, int64_t lds, int64_t ldt, int32_t rows, int32_t V,
                       const int32_t* labels, const int32_t* row_offsets, int32_t utterances, float scale_kd,
                       float scale_ce, void* workspace, float* lse_s, float* lse_t, float* coef_kd, float* coef_ce,
                       float* loss_ld, float* loss_ntp, void* stream) {
  return kd_ce_loss_fwd(student, teacher, lds, ldt, rows, V, labels, row_offsets, utterances, scale_kd, scale_ce,
                        workspace, lse_s, lse_t, coef_kd, coef_ce, loss_ld, loss_ntp, S(stream));
}

int b2s_kd_ce_loss_bwd(const void* student, const void* teacher, int64_t lds, int64_t ldt, int32_t rows, int32_t V,
                       const int32_t* labels, const float* lse_s, const float* lse_t, const float* coef_kd,
                       const float* coef_ce, const float* loss_scale, void* d_student, int64_t ldd, int32_t fmt,
                       void* stream) {
  return kd_ce_loss_bwd(student, teacher, lds, ldt, rows, V, labels, lse_s, lse_t, coef_kd, coef_ce, loss_scale,
                        d_student, ldd, fmt, S(stream));
}

int b2s_layernorm_fwd(const void* x, int32_t in_bf16, const float* gamma, const float* beta, float eps,
                      int32_t act_gelu, void* y_bf16, int64_t rows, int32_t C, int32_t fmt, void* stream) {
  return layernorm_fwd(x, in_bf16, gamma, beta, eps, act_gelu, y_bf16, rows, C, fmt, S(stream));
}
int b2s_rmsnorm_fwd(const float* x, const float* w, float eps, void* y_bf16, int64_t rows, int32_t C, int32_t fmt, void* stream) {
  return rmsnorm_fwd(x, w, eps, y_bf16, rows, C, fmt, S(stream));
}
int b2s_rmsnorm_gather_fwd(const float* x, const int32_t* row_index, const float* w, float eps, void* y_bf16,
                           int64_t rows, int32_t C, int32_t fmt, void* stream) {
  return rmsnorm_gather_fwd(x, row_index, w, eps, y_bf16, rows, C, fmt, S(stream));
}
int b2s_layernorm_avgpool_fwd(const float* x, const float* gamma, const float* beta, float eps, void* y_bf16,
                              int32_t batches, int32_t frames, int32_t C, int32_t kernel, int32_t stride,
                              int32_t out_frames, int32_t fmt, void* stream) {
  return layernorm_avgpool_fwd(x, gamma, beta, eps, y_bf16, batches, frames, C, kernel, stride, out_frames, fmt, S(stream));
}
int b2s_conv0_ln_gelu_fwd(const float* wave, int64_t wave_stride, int32_t batches, int32_t samples, const float* w,
                          const float* bias, const float* gamma, const float* beta, float eps, void* y_bf16,
                          int32_t out_frames, int32_t fmt, void* stream) {
  return conv0_ln_gelu_fwd(wave, wave_stride, batches, samples, w, bias, gamma, beta, eps, y_bf16, out_frames,
                           fmt, S(stream));
}
int b2s_embed_splice_fwd(const void* embed_table_bf16, const float* audio_embeds, const int32_t* row_src, float* h0,
                         int64_t rows, int32_t C, int32_t fmt, void* stream) {
  return embed_splice_fwd(embed_table_bf16, audio_embeds, row_src, h0, rows, C, fmt, S(stream));
}
int b2s_rowpair_sqdiff_fwd(const float* h, const int32_t* rows_a, const int32_t* rows_b, float* out, int32_t pairs,
                           int32_t C, void* stream) {
  return rowpair_sqdiff_fwd(h, rows_a, rows_b, out, pairs, C, S(stream));
}
int b2s_posconv_weight_pack(const float* g, const float* v, void* w_packed_bf16, int32_t cout, int32_t cin_g,
                            int32_t k, int32_t fmt, void* stream) {
  return posconv_weight_pack(g, v, w_packed_bf16, cout, cin_g, k, fmt, S(stream));
}
int b2s_cast_f32_to_h16(const float* x, void* y, int64_t n, int32_t fmt, void* stream) {
  return cast_f32_to_h16(x, y, n, fmt, S(stream));
}
int b2s_cast_h16_to_f32(const void* x, float* y, int64_t n, int32_t fmt, void* stream) {
  return cast_h16_to_f32(x, y, n, fmt, S(stream));
}
int b2s_attention_fwd(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                      const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, int64_t total_rows, int32_t Hq,
                      int32_t Hkv, int32_t D, float scale, int32_t causal, float* lse, int32_t fmt, void* stream) {
  return attention_fwd(q, k, v, ld_qkv, o, ld_o, cu_seqlens, num_seqs, max_seqlen, total_rows, Hq, Hkv, D, scale,
                       causal, lse, fmt, S(stream));
}
int b2s_attention_fwd_prefix(const void* q, const void* k, const void* v, int64_t ld_qkv, void* o, int64_t ld_o,
                             const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, int64_t total_rows,
                             int32_t Hq, int32_t Hkv, int32_t D, float scale, float* lse, int32_t fmt,
                             int32_t shared_prefix_len, void* stream) {
  return attention_fwd(q, k, v, ld_qkv, o, ld_o, cu_seqlens, num_seqs, max_seqlen, total_rows, Hq, Hkv, D, scale, 1, lse,
                       fmt, S(stream), nullptr, shared_prefix_len);
}
int b2s_attention_bwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const void* o, int64_t ld_o,
                      const void* dout, int64_t ld_do, const float* lse, float* delta_ws, void* dq, void* dk, void* dv,
                      int64_t ld_dqkv, const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen,
                      int64_t total_rows, int32_t Hq, int32_t Hkv, int32_t D, float scale, int32_t causal,
                      const float* rope_cs, int32_t fmt, void* stream) {
  return attention_bwd(q, k, v, ld_qkv, o, ld_o, dout, ld_do, lse, delta_ws, dq, dk, dv, ld_dqkv, cu_seqlens, num_seqs,
                       max_seqlen, total_rows, Hq, Hkv, D, scale, causal, rope_cs, fmt, S(stream));
}

int b2s_hubert_num_frames(const b2s_hubert_weights* w, int32_t samples, int32_t* frames, int32_t* pooled) {
  return hubert_num_frames(w, samples, frames, pooled);
}
size_t b2s_hubert_workspace_bytes(const b2s_hubert_weights* w, int32_t batches, int32_t samples) {
  return hubert_workspace_bytes(w, batches, samples);
}
int b2s_hubert_forward(const b2s_hubert_weights* w, const float* wave, int64_t wave_stride, int32_t batches,
                       int32_t samples, void* workspace, size_t workspace_bytes, float* audio_embeds,
                       float* last_hidden, void* stream) {
  return hubert_forward(w, wave, wave_stride, batches, samples, workspace, workspace_bytes, audio_embeds, last_hidden,
                        S(stream));
}
size_t b2s_whisper_workspace_bytes(const b2s_whisper_weights* w, int32_t batches) {
  return whisper_workspace_bytes(w, batches);
}
int b2s_whisper_forward(const b2s_whisper_weights* w, const float* mel, int32_t batches, int32_t frames_in,
                        void* workspace, size_t workspace_bytes, float* audio_embeds, float* last_hidden,
                        void* stream) {
  return whisper_forward(w, mel, batches, frames_in, workspace, workspace_bytes, audio_embeds, last_hidden, S(stream));
}
size_t b2s_llama_workspace_bytes(const b2s_llama_weights* w, int32_t rows, int32_t logit_rows) {
  return llama_workspace_bytes(w, rows, logit_rows);
}
int b2s_llama_prefill(const b2s_llama_weights* w, float* h, int32_t rows, const int32_t* cu_seqlens,
                      int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                      const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16,
                      const int32_t* tap_layers, int32_t num_taps, const int32_t* tap_rows_a,
                      const int32_t* tap_rows_b, int32_t pairs, float* fd_sq, float* all_hidden, void* workspace,
                      size_t workspace_bytes, void* stream) {
  return llama_prefill(w, h, rows, cu_seqlens, num_seqs, max_seqlen, positions, logit_rows_index, logit_rows,
                       logits_bf16, tap_layers, num_taps, tap_rows_a, tap_rows_b, pairs, fd_sq, all_hidden, workspace,
                       workspace_bytes, S(stream));
}

int b2s_llama_prefill_prefix(const b2s_llama_weights* w, float* h, int32_t rows, const int32_t* cu_seqlens,
                             int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                             const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16,
                             const int32_t* tap_layers, int32_t num_taps, const int32_t* tap_rows_a,
                             const int32_t* tap_rows_b, int32_t pairs, float* fd_sq, float* all_hidden, void* workspace,
                             size_t workspace_bytes, int32_t shared_prefix_len, void* stream) {
  return llama_prefill(w, h, rows, cu_seqlens, num_seqs, max_seqlen, positions, logit_rows_index, logit_rows,
                       logits_bf16, tap_layers, num_taps, tap_rows_a, tap_rows_b, pairs, fd_sq, all_hidden, workspace,
                       workspace_bytes, S(stream), shared_prefix_len);
}

size_t b2s_llama_train_workspace_bytes(const b2s_llama_weights* w, int32_t rows, int32_t logit_rows) {
  return llama_train_workspace_bytes(w, rows, logit_rows);
}
size_t b2s_llama_backward_workspace_bytes(const b2s_llama_weights* w, int32_t rows_bwd, int32_t n_dl) {
  return llama_backward_workspace_bytes(w, rows_bwd, n_dl);
}
int b2s_llama_forward_train(const b2s_llama_weights* w, const b2s_llama_saved* saved, int32_t rows,
                            const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                            const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16,
                            const int32_t* tap_layers, int32_t num_taps, const int32_t* tap_rows_a,
                            const int32_t* tap_rows_b, int32_t pairs, float* fd_sq, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return llama_forward_train(w, saved, rows, cu_seqlens, num_seqs, max_seqlen, positions, logit_rows_index, logit_rows,
                             logits_bf16, tap_layers, num_taps, tap_rows_a, tap_rows_b, pairs, fd_sq, workspace,
                             workspace_bytes, S(stream));
}
int b2s_llama_backward(const b2s_llama_weights* w, const b2s_llama_weights_t* wt, const b2s_llama_saved* saved,
                       int32_t rows, int32_t rows_bwd, const int32_t* cu_seqlens, int32_t num_seqs_bwd,
                       int32_t max_seqlen, const void* d_logits, const int32_t* dl_rows_index, int32_t n_dl,
                       const int32_t* tap_layers, int32_t num_taps, const int32_t* tap_rows_a,
                       const int32_t* tap_rows_b, const float* tap_coef, const float* loss_scale, int32_t pairs,
                       float* dh, void* workspace, size_t workspace_bytes, void* stream) {
  return llama_backward(w, wt, saved, rows, rows_bwd, cu_seqlens, num_seqs_bwd, max_seqlen, d_logits, dl_rows_index,
                        n_dl, tap_layers, num_taps, tap_rows_a, tap_rows_b, tap_coef, loss_scale, pairs, dh, workspace,
                        workspace_bytes, S(stream));
}
int b2s_rmsnorm_bwd(const float* x, const int32_t* x_index, const float* w, float eps, const float* dy, float* dh,
                    const int32_t* dh_index, void* dh_bf16, int64_t rows, int32_t C, int32_t fmt, void* stream) {
  return rmsnorm_bwd(x, x_index, w, eps, dy, dh, dh_index, dh_bf16, rows, C, fmt, S(stream));
}
int b2s_layernorm_bwd(const float* x, const float* gamma, float eps, const void* dy, int32_t dy_bf16, float* dh,
                      int32_t accumulate, void* dh_bf16, float* dgamma, float* dbeta, int64_t rows, int32_t C,
                      int32_t fmt, void* stream) {
  return layernorm_bwd(x, gamma, eps, dy, dy_bf16, dh, accumulate, dh_bf16, dgamma, dbeta, rows, C, fmt, S(stream));
}
int b2s_swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t rows, int32_t F, int32_t fmt, void* stream) {
  return swiglu_bwd(gu, dact, dgu, rows, F, fmt, S(stream));
}
int b2s_gelu_bwd(const void* pre, const void* dy, void* dpre, int64_t n, int32_t fmt, float* colsum, int32_t F,
                 void* stream) {
  return gelu_bwd(pre, dy, dpre, n, fmt, S(stream), nullptr, colsum, F);
}
int b2s_gather_rows_f32(const float* src, const int32_t* index, float* out, int64_t rows, int32_t C, void* stream) {
  return gather_rows_f32(src, index, out, rows, C, S(stream));
}
int b2s_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int32_t step, float grad_scale, const b2s_grad_scaler_state* scaler,
                   void* stream) {
  return adamw_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                    reinterpret_cast<const GradScalerState*>(scaler), S(stream));
}
int b2s_nonfinite_check(const float* g, int64_t n, b2s_grad_scaler_state* scaler, void* stream) {
  return nonfinite_check(g, n, reinterpret_cast<GradScalerState*>(scaler), S(stream));
}
int b2s_grad_scaler_update(b2s_grad_scaler_state* scaler, float growth_factor, float backoff_factor,
                           int32_t growth_interval, void* stream) {
  return grad_scaler_update(reinterpret_cast<GradScalerState*>(scaler), growth_factor, backoff_factor, growth_interval,
                            S(stream));
}
static_assert(sizeof(b2s_grad_scaler_state) == sizeof(GradScalerState), "scaler state layouts must match");

size_t b2s_hubert_saved_bytes(const b2s_hubert_weights* w, int32_t batches, int32_t samples) {
  return hubert_saved_bytes(w, batches, samples);
}
size_t b2s_hubert_backward_workspace_bytes(const b2s_hubert_weights* w, int32_t batches, int32_t samples) {
  return hubert_backward_workspace_bytes(w, batches, samples);
}
int b2s_hubert_forward_train(const b2s_hubert_weights* w, const float* wave, int64_t wave_stride, int32_t batches,
                             int32_t samples, const int32_t* samples_per_utt, void* saved, size_t saved_bytes,
                             float* audio_embeds, const b2s_encoder_regularizers* reg, void* stream) {
  return hubert_forward_train(w, wave, wave_stride, batches, samples, samples_per_utt, saved, saved_bytes, audio_embeds,
                              reg, S(stream));
}
int b2s_hubert_backward(const b2s_hubert_weights* w, const void* pos_w_dgrad, const b2s_hubert_grads* grads,
                        const float* wave, int64_t wave_stride, int32_t batches, int32_t samples,
                        const int32_t* samples_per_utt, void* saved, size_t saved_bytes, const float* d_audio_embeds,
                        void* workspace, size_t workspace_bytes, const b2s_encoder_regularizers* reg,
                        void* const* layer_done_events, void* stream) {
  return hubert_backward(w, pos_w_dgrad, grads, wave, wave_stride, batches, samples, samples_per_utt, saved, saved_bytes,
                         d_audio_embeds, workspace, workspace_bytes, reg, layer_done_events, S(stream));
}
int b2s_drop_mask_dump(uint8_t* out, int64_t n, uint64_t seed, uint32_t site, uint32_t a, uint32_t b, float p,
                       uint32_t e_first, void* stream) {
  return drop_mask_dump(out, n, seed, site, a, b, p, e_first, S(stream));
}
int b2s_layernorm_bwd_ex(const void* x, int32_t x_bf16, const float* gamma, const float* beta, int32_t act_gelu,
                         float eps, const void* dy, int32_t dy_bf16, float* dh, int32_t accumulate, void* dx_bf16,
                         float* dgamma, float* dbeta, int64_t rows, int32_t C, int32_t fmt, float* dh_colsum,
                         void* stream) {
  return layernorm_bwd_ex(x, x_bf16, gamma, beta, act_gelu, eps, dy, dy_bf16, dh, accumulate, dx_bf16, dgamma, dbeta,
                          rows, C, fmt, S(stream), dh_colsum);
}
int b2s_colsum_accum(const void* x, int32_t x_bf16, float* out, int64_t rows, int32_t C, int32_t fmt, void* stream) {
  return colsum_accum(x, x_bf16, out, rows, C, fmt, S(stream));
}
int b2s_avgpool_bwd(const float* dpooled, float* dx, int32_t batches, int32_t frames, int32_t C, int32_t kernel,
                    int32_t stride, int32_t pooled, void* stream) {
  return avgpool_bwd(dpooled, dx, batches, frames, C, kernel, stride, pooled, S(stream));
}
int b2s_col2im_add(const void* dcol_bf16, void* dx_bf16, int32_t batches, int32_t tin, int32_t tout, int32_t k,
                   int32_t s, int32_t C, int32_t fmt, void* stream) {
  return col2im_add(dcol_bf16, dx_bf16, batches, tin, tout, k, s, C, fmt, S(stream));
}
int b2s_conv0_bwd(const float* wave, int64_t wave_stride, int32_t batches, int32_t samples, const float* w,
                  const float* bias, const float* gamma, const float* beta, float eps, const void* dy_bf16,
                  int32_t frames, float* dW, float* db, float* dgamma, float* dbeta, int32_t fmt, void* stream) {
  return conv0_bwd(wave, wave_stride, batches, samples, w, bias, gamma, beta, eps, dy_bf16, frames, dW, db, dgamma,
                   dbeta, fmt, S(stream));
}

size_t b2s_llama_kv_cache_bytes(const b2s_llama_weights* w, int32_t slots) { return llama_kv_cache_bytes(w, slots); }
int b2s_llama_prefill_kv(const b2s_llama_weights* w, float* h, int32_t rows, const int32_t* cu_seqlens,
                         int32_t num_seqs, int32_t max_seqlen, const int32_t* positions,
                         const int32_t* logit_rows_index, int32_t logit_rows, void* logits_bf16, void* kv_cache,
                         int32_t kv_slots, const int32_t* kv_slot_of_row, void* workspace, size_t workspace_bytes,
                         void* stream) {
  return llama_prefill_kv(w, h, rows, cu_seqlens, num_seqs, max_seqlen, positions, logit_rows_index, logit_rows,
                          logits_bf16, nullptr, 0, nullptr, nullptr, 0, nullptr, nullptr, kv_cache, kv_slots,
                          kv_slot_of_row, workspace, workspace_bytes, S(stream));
}
size_t b2s_llama_decode_workspace_bytes(const b2s_llama_weights* w, int32_t batch) {
  return llama_decode_workspace_bytes(w, batch);
}
int b2s_llama_decode_step(const b2s_llama_weights* w, const void* embed_table_bf16, const int32_t* token_ids,
                          int32_t batch, void* kv_cache, int32_t kv_slots, const int32_t* seq_start,
                          const int32_t* seq_len, void* logits_bf16, void* workspace, size_t workspace_bytes,
                          void* stream) {
  return llama_decode_step(w, embed_table_bf16, token_ids, batch, kv_cache, kv_slots, seq_start, seq_len, logits_bf16,
                           workspace, workspace_bytes, S(stream));
}

int b2s_whisper_log_mel(const float* wave, int64_t wave_stride, int32_t batches, int32_t samples,
                        const float* mel_filters, float* out, int32_t frames, int32_t* max_ws, void* stream) {
  return whisper_log_mel(wave, wave_stride, batches, samples, mel_filters, out, frames, max_ws, S(stream));
}

size_t b2s_whisper_saved_bytes(const b2s_whisper_weights* w, int32_t batches) { return whisper_saved_bytes(w, batches); }
size_t b2s_whisper_backward_workspace_bytes(const b2s_whisper_weights* w, int32_t batches) {
  return whisper_backward_workspace_bytes(w, batches);
}
int b2s_whisper_forward_train(const b2s_whisper_weights* w, const float* mel, int32_t batches, int32_t frames_in,
                              void* saved, size_t saved_bytes, float* audio_embeds, void* stream) {
  return whisper_forward_train(w, mel, batches, frames_in, saved, saved_bytes, audio_embeds, S(stream));
}
int b2s_whisper_backward(const b2s_whisper_weights* w, const b2s_whisper_grads* grads, int32_t batches, void* saved,
                         size_t saved_bytes, const float* d_audio_embeds, void* workspace, size_t workspace_bytes,
                         void* const* layer_done_events, void* stream) {
  return whisper_backward(w, grads, batches, saved, saved_bytes, d_audio_embeds, workspace, workspace_bytes,
                          layer_done_events, S(stream));
}

}  // extern "C"
