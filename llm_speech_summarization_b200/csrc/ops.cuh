// ops.cuh -- internal C++ launchers for the memory-bound kernels and attention (see include/b2s.h for the
// C ABI that wraps them). All take raw device pointers + sizes + the stream to enqueue on; none synchronise.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "rng.cuh"

namespace b2s {

// ---- loss.cu
size_t kd_ce_workspace_bytes(int rows, int V);
int kd_ce_loss_fwd(const void* S, const void* T, long long lds, long long ldt, int rows, int V, const int* labels,
                   const int* row_offsets, int utterances, float scale_kd, float scale_ce, void* workspace,
                   float* lse_s, float* lse_t, float* coef_kd, float* coef_ce, float* loss_ld, float* loss_ntp,
                   cudaStream_t stream);
// dS is written in format `fmt` and multiplied by *loss_scale (device scalar, optional: the GradScaler state's scale)
int kd_ce_loss_bwd(const void* S, const void* T, long long lds, long long ldt, int rows, int V, const int* labels,
                   const float* lse_s, const float* lse_t, const float* coef_kd, const float* coef_ce,
                   const float* loss_scale, void* dS, long long ldd, int fmt, cudaStream_t stream);

// ---- norm.cu
// Every `fmt` below is the 16-bit storage format (B2S_FMT_BF16 = 0 / B2S_FMT_F16 = 1, include/b2s.h) shared by ALL the
// 16-bit tensors of the call; parameters named *_bf16 are "16-bit in fmt" (the names predate fp16 support).
// y[r, :] = act( (x[r,:] - mean) * rstd * gamma + beta ); x is fp32 (in_bf16 = 0) or 16-bit (in_bf16 = 1)
int layernorm_fwd(const void* x, int in_bf16, const float* gamma, const float* beta, float eps, int act_gelu,
                  void* y_bf16, long long rows, int C, int fmt, cudaStream_t stream);
// y[r, :] = x[r,:] * rsqrt(mean(x^2) + eps) * w ; x fp32
int rmsnorm_fwd(const float* x, const float* w, float eps, void* y_bf16, long long rows, int C, int fmt,
                cudaStream_t stream);
// rows gathered through an index list (final norm on the consumed rows only)
int rmsnorm_gather_fwd(const float* x, const int* row_index, const float* w, float eps, void* y_bf16, long long rows,
                       int C, int fmt, cudaStream_t stream);
// final LayerNorm of the encoder fused with AvgPool1d(kernel, stride) over time:
// y[b, j, :] = mean_{r<kernel} LN(x[b, j*stride + r, :])   (REF/model/audio_encoder.py:59-63)
int layernorm_avgpool_fwd(const float* x, const float* gamma, const float* beta, float eps, void* y_bf16, int batches,
                          int frames, int C, int kernel, int stride, int out_frames, int fmt, cudaStream_t stream);

// ---- misc.cu
// HuBERT conv layer 0: Conv1d(1->512,k=10,s=5)+bias -> LayerNorm(512) -> GELU, channels-last bf16 out
int conv0_ln_gelu_fwd(const float* wave, long long wave_stride, int batches, int samples, const float* w /*[512,10]*/,
                      const float* bias, const float* gamma, const float* beta, float eps, void* y_bf16, int out_frames,
                      int fmt, cudaStream_t stream);
// h0[row, :] = src >= 0 ? embed_table[src, :] : audio_embeds[-(src+1), :]
int embed_splice_fwd(const void* embed_table_bf16, const float* audio_embeds, const int* row_src, float* h0,
                     long long rows, int C, int fmt, cudaStream_t stream);
// sum over columns of (a[ra[i], :] - b[rb[i], :])^2 -> out[i]
int rowpair_sqdiff_fwd(const float* h, const int* rows_a, const int* rows_b, float* out, int pairs, int C,
                       cudaStream_t stream);
// w[co, ci, k] = g[k] * v[co, ci, k] / ||v[:, :, k]||, repacked bf16 as [co][k][ci] (K-major per tap)
int posconv_weight_pack(const float* g, const float* v, void* w_packed_bf16, int cout, int cin_g, int k, int fmt,
                        cudaStream_t stream);
// log-mel (B, C, T) fp32 -> channels-last bf16 (B, T+2, C), zero row before/after each utterance (conv padding)
int mel_to_padded_cl(const float* x, void* y_bf16, int batches, int channels, int frames, int fmt, cudaStream_t stream);
int gather_rows_bf16(const void* src, const int* index, void* out, long long rows, int C, cudaStream_t stream);
int cast_f32_to_h16(const float* x, void* y, long long n, int fmt, cudaStream_t stream);
int cast_h16_to_f32(const void* x, float* y, long long n, int fmt, cudaStream_t stream);
// y_bf16 = gelu(x_f32 + residual) etc. are fused in GEMM epilogues; nothing else elementwise is needed.

// ---- logmel.cu: WhisperFeatureExtractor on the GPU (reflect-padded STFT 400/160, 80 slaney mels, log10, clamp, scale)
// out fp32 [batches, 80, frames], frames = samples / 160; max_ws int32 [batches] scratch
int whisper_log_mel(const float* wave, long long wave_stride, int batches, int samples, const float* mel_filters,
                    float* out, int frames, int* max_ws, cudaStream_t stream);

// ---- backward.cu (training step: memory-bound backward kernels + optimizer)
// dh[dst] += RMSNorm^T(dy) ; x / dst rows optionally gathered through index lists; optional bf16 copy of dh rows
int rmsnorm_bwd(const float* x, const int* x_index, const float* w, float eps, const float* dy, float* dh,
                const int* dh_index, void* dh_bf16, long long rows, int C, int fmt, cudaStream_t stream);
// dh (+)= LayerNorm^T(dy); dgamma / dbeta accumulated with atomics; dy fp32 or bf16
int layernorm_bwd(const float* x, const float* gamma, float eps, const void* dy, int dy_bf16, float* dh, int accumulate,
                  void* dh_bf16, float* dgamma, float* dbeta, long long rows, int C, int fmt, cudaStream_t stream);
int swiglu_bwd(const void* gu, const void* dact, void* dgu, long long rows, int F, int fmt, cudaStream_t stream);
// drop (optional): dy is the gradient w.r.t. dropout(gelu(pre)) of that site (element index = linear index)
// colsum (optional, with the row width F, F % 2048 == 0): colsum[c] += sum over rows of dpre[., c] -- the bias gradient of
// the Linear that produced `pre`, accumulated on the way instead of by a second pass over dpre
int gelu_bwd(const void* pre, const void* dy, void* dpre, long long n, int fmt, cudaStream_t stream,
             const DropSpec* drop = nullptr, float* colsum = nullptr, int F = 0);
// dh[rows_a[i]] += coef[i] * (*loss_scale) * (h[rows_a[i]] - h[rows_b[i]]); loss_scale: optional device scalar
int add_rowdiff(const float* h, const int* rows_a, const int* rows_b, const float* coef, const float* loss_scale,
                float* dh, void* dh_bf16, int pairs, int C, int fmt, cudaStream_t stream);
int gather_rows_f32(const float* src, const int* index, float* out, long long rows, int C, cudaStream_t stream);
// Dynamic loss scaling on the device (torch.cuda.amp.GradScaler, REF/trainer.py:252,374,381-382): the state lives in
// device memory so that neither the skip-on-overflow decision nor the scale update synchronises the host.
struct GradScalerState {  // mirrors b2s_grad_scaler_state (include/b2s.h)
  float scale;         // current loss scale S: gradients carry the factor S until adamw_step divides it out
  int growth_tracker;  // consecutive clean optimizer steps since the last scale change
  int found_inf;       // set by nonfinite_check for the step in flight; cleared by grad_scaler_update
  int opt_steps;       // optimizer steps actually taken (the `step` of AdamW's bias correction)
  int skipped_steps;   // steps skipped because of an overflow
  int reserved[3];
};
// scaler != nullptr: gradient multiplier grad_scale / scaler->scale, nothing is updated when scaler->found_inf is set,
// bias correction from scaler->opt_steps + 1 (`step` ignored)
int adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int step, float grad_scale, const GradScalerState* scaler, cudaStream_t stream);
int nonfinite_check(const float* g, long long n, GradScalerState* scaler, cudaStream_t stream);
int grad_scaler_update(GradScalerState* scaler, float growth, float backoff, int interval, cudaStream_t stream);

// ---- backward_enc.cu (trainable audio encoder)
// LayerNorm (+ optional erf-GELU on its output) backward; x / dy fp32 or bf16; dh fp32 (+=) and/or bf16 dx outputs
int layernorm_bwd_ex(const void* x, int x_bf16, const float* gamma, const float* beta, int act_gelu, float eps,
                     const void* dy, int dy_bf16, float* dh, int accumulate, void* dx_bf16, float* dgamma, float* dbeta,
                     long long rows, int C, int fmt, cudaStream_t stream, float* dh_colsum = nullptr);
int colsum_accum(const void* x, int x_bf16, float* out, long long rows, int C, int fmt, cudaStream_t stream);
int avgpool_bwd(const float* dpooled, float* dx, int batches, int frames, int C, int kernel, int stride, int pooled,
                cudaStream_t stream);
int col2im_add(const void* dcol_bf16, void* dx_bf16, int batches, int tin, int tout, int k, int s, int C, int fmt,
               cudaStream_t stream);
int conv0_bwd(const float* wave, long long wave_stride, int batches, int samples, const float* w, const float* bias,
              const float* gamma, const float* beta, float eps, const void* dy_bf16, int frames, float* dW, float* db,
              float* dgamma, float* dbeta, int fmt, cudaStream_t stream);

// ---- regularize.cu (train-mode regularisers of the HuBERT encoder; decisions regenerated from counters, rng.cuh)
// x *= keep ? 1/(1-p) : 0 in place on an fp32 tensor and / or its bf16 copy (element index = linear index)
int dropout_apply(float* x_f32, void* x_bf16, long long n, const DropSpec& d, int fmt, cudaStream_t stream);
// SpecAugment: h[row, :] = embed where time_mask[row] != 0
int mask_rows_f32(float* h, const unsigned char* time_mask, const float* embed, long long rows, int C,
                  cudaStream_t stream);
// backward of [feature-projection dropout -> SpecAugment] on dh in place; g_embed += gradient of the replaced rows
int featproj_reg_bwd(float* dh, const unsigned char* time_mask, float* g_embed, long long rows, int C, const DropSpec& d,
                     cudaStream_t stream);
// test hook: out[i] = 1 if element e_first + i of stream (seed, site, a, b) is kept at drop probability p
int drop_mask_dump(unsigned char* out, long long n, unsigned long long seed, uint32_t site, uint32_t a, uint32_t b,
                   float p, uint32_t e_first, cudaStream_t stream);

// ---- attention_tc.cu / attention_bwd_tc.cu (tcgen05 / TMEM / TMA flash attention)
// Packed variable-length attention. q/k/v are 16-bit views (format fmt) into one [rows, ld] buffer (fused QKV output):
// head h of row r lives at base + r*ld + h*D. Sequences are rows [cu[s], cu[s+1]).
// GQA: query head h uses kv head h / (Hq / Hkv). Output o [rows, Hq*D] in the same format.
// total_rows = rows of the packed buffers (the TMA tensor maps zero-fill beyond it).
int attention_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* o, long long ld_o,
                  const int* cu_seqlens, int num_seqs, int max_seqlen, long long total_rows, int Hq, int Hkv, int D,
                  float scale, int causal, float* lse /* optional [rows, Hq] */, int fmt, cudaStream_t stream,
                  const AttnDrop* drop = nullptr /* attention-probability dropout */,
                  int shared_prefix_len = 0 /* > 0: sequence 0 is a prefix every other sequence also attends to */);
// Backward of attention_fwd. lse = the forward's saved log-sum-exp; delta_ws = fp32 [rows, Hq] scratch.
// dq / dk / dv are 16-bit views with row stride ld_dqkv (head h at column h*D). rope_cs (optional, [npos, D]) fuses
// the inverse rotary rotation into the dq / dk stores (positions = row index inside its sequence).
// q, k, v, o, dout, dq, dk, dv all share `fmt` (a mixed-format tcgen05.mma traps).
int attention_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* o, long long ld_o,
                  const void* dout, long long ld_do, const float* lse, float* delta_ws, void* dq, void* dk, void* dv,
                  long long ld_dqkv, const int* cu_seqlens, int num_seqs, int max_seqlen, long long total_rows, int Hq,
                  int Hkv, int D, float scale, int causal, const float* rope_cs, int fmt, cudaStream_t stream,
                  const AttnDrop* drop = nullptr);

}  // namespace b2s
