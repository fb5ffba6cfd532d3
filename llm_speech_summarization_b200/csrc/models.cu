// models.cu -- whole-model forward passes: the layer loops of the audio encoder and of the LLM prefill run
// here in C++ (one C-ABI call per forward; every step is an asynchronous launch on the caller's stream).
//
//   hubert_forward : AudioEncoder.forward, HuBERT + "pool" branch  (REF/model/audio_encoder.py:56-88 over
//                    HubertModel.forward TF/models/hubert/modeling_hubert.py:889-958, eval mode)
//   llama_prefill  : LlamaModel.forward + lm_head over packed sequences (REF/model/audio_llama.py:49-67 over
//                    TF/models/llama/modeling_llama.py:375-425), plus the FD-loss hidden-state taps
//                    (REF/trainer.py:358-370)
//
// Numerics: 16-bit GEMM operands in the weights struct's `fmt` (fp16 like the reference's autocast, or bf16), fp32
// accumulation, fp32 residual streams, fp32 norm statistics, fp32 pre-norm conv activations.
#include "../../include/b2s.h"
#include "b2s_common.cuh"
#include "gemm_sm100.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

struct Carver {
  uint8_t* base;
  size_t off = 0;
  size_t cap;
  Carver(void* b, size_t c) : base(reinterpret_cast<uint8_t*>(b)), cap(c) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  bool ok() const { return off <= cap; }
};

int conv_out_len(int in, int k, int s) { return in < k ? 0 : (in - k) / s + 1; }

struct HubertPlan {
  int t[8];  // t[0] = samples, t[i+1] = frames after conv layer i
  int frames, pooled;
  size_t bytes;
  // workspace pieces
  void *xa, *xb;   // bf16 ping-pong channels-last conv activations
  float* pre;      // fp32 pre-norm conv output
  float* h;        // fp32 residual stream [B*frames, H]
  void* xn;        // bf16 normalised activations [B*frames, H]
  void* qkv;       // bf16 [B*frames, 3H]
  void* ao;        // bf16 [B*frames, H]
  void* ff;        // bf16 [B*frames, F]
  void* pooled_x;  // bf16 [B*pooled, H]
  int* cu;         // int32 [B+1]
};

int plan_hubert(const b2s_hubert_weights* w, int batches, int samples, void* ws, size_t ws_bytes, HubertPlan* pl) {
  pl->t[0] = samples;
  pl->t[1] = conv_out_len(samples, 10, 5);
  for (int i = 0; i < 6; ++i) pl->t[i + 2] = conv_out_len(pl->t[i + 1], w->conv_k[i], w->conv_stride[i]);
  pl->frames = pl->t[7];
  pl->pooled = pl->frames >= w->pool_kernel ? (pl->frames - w->pool_kernel) / w->pool_stride + 1 : 0;
  const size_t B = batches;
  const size_t rows = B * pl->frames;
  Carver c(ws, ws_bytes);
  pl->xa = c.take(B * pl->t[1] * 512 * 2 + 4096);
  pl->xb = c.take(B * pl->t[2] * 512 * 2 + 4096);
  pl->pre = reinterpret_cast<float*>(c.take(B * pl->t[2] * 512 * 4));
  pl->h = reinterpret_cast<float*>(c.take(rows * w->hidden * 4));
  pl->xn = c.take(rows * w->hidden * 2);
  pl->qkv = c.take(rows * 3 * w->hidden * 2);
  pl->ao = c.take(rows * w->hidden * 2);
  pl->ff = c.take(rows * w->ffn * 2);
  pl->pooled_x = c.take(B * (pl->pooled > 0 ? pl->pooled : 1) * w->hidden * 2);
  pl->cu = reinterpret_cast<int*>(c.take((B + 1) * sizeof(int)));
  pl->bytes = c.off + 256;
  if (ws != nullptr && !c.ok()) {
    set_last_error("hubert workspace too small: need %zu bytes, got %zu", pl->bytes, ws_bytes);
    return B2S_ERR_INVALID;
  }
  return B2S_OK;
}

__global__ void iota_scaled_kernel(int* out, int n, int scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i * scale;
}

GemmArgs plain_gemm(const void* A, const void* W, long long M, int N, int K) {
  GemmArgs g{};
  g.A = A;
  g.a_dim0 = K;
  g.a_row_stride = K;
  g.a_batch_stride = 0;
  g.a_rows = static_cast<int>(M);
  g.W = W;
  g.w_rows = N;
  g.w_cols = K;
  g.M = static_cast<int>(M);
  g.N = N;
  g.batches = 1;
  g.groups = 1;
  g.taps = 1;
  g.k_per_tap = K;
  g.ldo = N;
  g.out_batch_rows = 0;
  return g;
}

}  // namespace

// The pre-LN transformer stack shared by HuBERT (stable-layer-norm variant) and the Whisper encoder:
//   h += out_proj(attention(qkv(LN1(h))));  h += W2 gelu(W1 LN2(h))      (fp32 residual stream h, bf16 operands)
int encoder_stack(const b2s_encoder_layer* layers, int num_layers, int H, int F, int heads, float eps, float* h,
                  void* xn, void* qkv_buf, void* ao, void* ff, const int* cu, int B, int frames, int fmt,
                  cudaStream_t stream) {
  const long long rows = static_cast<long long>(B) * frames;
  int rc;
  for (int l = 0; l < num_layers; ++l) {
    const b2s_encoder_layer& L = layers[l];
    rc = layernorm_fwd(h, 0, L.ln1_g, L.ln1_b, eps, 0, xn, rows, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    {
      GemmArgs g = plain_gemm(xn, L.wqkv, rows, 3 * H, H);
      g.epi = EPI_BF16;
      g.bias = L.bqkv;
      g.out = qkv_buf;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    {
      const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(qkv_buf);
      rc = attention_fwd(qkv, qkv + H, qkv + 2 * H, 3 * H, ao, H, cu, B, frames, rows, heads, heads, 64, 0.125f, 0,
                         nullptr, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    {
      GemmArgs g = plain_gemm(ao, L.wo, rows, H, H);
      g.epi = EPI_RESID_F32;
      g.bias = L.bo;
      g.out = h;
      g.resid = h;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    rc = layernorm_fwd(h, 0, L.ln2_g, L.ln2_b, eps, 0, xn, rows, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    {
      GemmArgs g = plain_gemm(xn, L.w1, rows, F, H);
      g.epi = EPI_BF16;
      g.act = ACT_GELU;
      g.bias = L.b1;
      g.out = ff;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    {
      GemmArgs g = plain_gemm(ff, L.w2, rows, H, F);
      g.epi = EPI_RESID_F32;
      g.bias = L.b2;
      g.out = h;
      g.resid = h;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
  }
  return B2S_OK;
}

int hubert_num_frames(const b2s_hubert_weights* w, int samples, int* frames, int* pooled) {
  B2S_REQUIRE(w != nullptr, "hubert: null weights");
  HubertPlan pl;
  int rc = plan_hubert(w, 1, samples, nullptr, 0, &pl);
  if (rc != B2S_OK) return rc;
  if (frames) *frames = pl.frames;
  if (pooled) *pooled = pl.pooled;
  return B2S_OK;
}

size_t hubert_workspace_bytes(const b2s_hubert_weights* w, int batches, int samples) {
  if (w == nullptr || batches <= 0 || samples <= 0) return 0;
  HubertPlan pl;
  plan_hubert(w, batches, samples, nullptr, 0, &pl);
  return pl.bytes;
}

int hubert_forward(const b2s_hubert_weights* w, const float* wave, long long wave_stride, int batches, int samples,
                   void* workspace, size_t workspace_bytes, float* audio_embeds, float* last_hidden,
                   cudaStream_t stream) {
  B2S_REQUIRE(w && wave && workspace && audio_embeds, "hubert_forward: null pointer");
  B2S_REQUIRE(batches > 0 && samples > 0, "hubert_forward: empty batch");
  B2S_REQUIRE(w->hidden % 256 == 0 && w->hidden % w->heads == 0 && w->hidden / w->heads == 64,
              "hubert_forward: hidden/heads must give head_dim 64");
  B2S_REQUIRE(w->pos_groups > 0 && w->hidden / w->pos_groups == 64, "hubert_forward: positional conv needs 64 ch/group");
  HubertPlan pl;
  int rc = plan_hubert(w, batches, samples, workspace, workspace_bytes, &pl);
  if (rc != B2S_OK) return rc;
  B2S_REQUIRE(pl.frames > 0 && pl.pooled > 0, "hubert_forward: audio too short (%d samples -> %d frames)", samples,
              pl.frames);
  const int B = batches, H = w->hidden, F = w->ffn, fmt = w->fmt;
  const long long rows = static_cast<long long>(B) * pl.frames;
  const float eps = w->ln_eps;

  // ---- conv feature extractor
  rc = conv0_ln_gelu_fwd(wave, wave_stride, B, samples, w->conv0_w, w->conv0_b, w->conv0_ln_g, w->conv0_ln_b, eps,
                         pl.xa, pl.t[1], fmt, stream);
  if (rc != B2S_OK) return rc;
  void* cur = pl.xa;
  void* nxt = pl.xb;
  for (int i = 0; i < 6; ++i) {
    const int tin = pl.t[i + 1], tout = pl.t[i + 2];
    const int k = w->conv_k[i], s = w->conv_stride[i];
    GemmArgs g{};
    g.A = cur;
    g.a_dim0 = k * 512;                                // overlapping window view: row t = x[s*t : s*t+k, :]
    g.a_row_stride = static_cast<long long>(s) * 512;
    g.a_batch_stride = static_cast<long long>(tin) * 512;
    g.a_rows = tout;
    g.W = w->conv_w[i];
    g.w_rows = 512;
    g.w_cols = k * 512;
    g.M = tout;
    g.N = 512;
    g.batches = B;
    g.groups = 1;
    g.taps = 1;
    g.k_per_tap = k * 512;
    g.epi = EPI_F32;
    g.act = ACT_NONE;
    g.bias = w->conv_b[i];
    g.out = pl.pre;
    g.ldo = 512;
    g.out_batch_rows = tout;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
    rc = layernorm_fwd(pl.pre, 0, w->conv_ln_g[i], w->conv_ln_b[i], eps, 1, nxt, static_cast<long long>(B) * tout, 512,
                       fmt,
                       stream);
    if (rc != B2S_OK) return rc;
    void* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  // cur: bf16 [B, frames, 512]

  // ---- feature projection: LN(512) -> Linear(512 -> H)
  rc = layernorm_fwd(cur, 1, w->fp_ln_g, w->fp_ln_b, eps, 0, nxt, rows, 512, fmt, stream);
  if (rc != B2S_OK) return rc;
  {
    GemmArgs g = plain_gemm(nxt, w->fp_w, rows, H, 512);
    g.epi = EPI_F32;
    g.bias = w->fp_b;
    g.out = pl.h;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  // ---- positional conv embedding: h += gelu(grouped_conv(h) + b)
  rc = cast_f32_to_h16(pl.h, pl.xn, rows * H, fmt, stream);
  if (rc != B2S_OK) return rc;
  {
    GemmArgs g{};
    g.A = pl.xn;
    g.a_dim0 = H;
    g.a_row_stride = H;
    g.a_batch_stride = static_cast<long long>(pl.frames) * H;
    g.a_rows = pl.frames;
    g.W = w->pos_w;
    g.w_rows = H;
    g.w_cols = w->pos_k * 64;
    g.M = pl.frames;
    g.N = 64;
    g.batches = B;
    g.groups = w->pos_groups;
    g.taps = w->pos_k;
    g.k_per_tap = 64;
    g.a_pad = w->pos_k / 2;
    g.a_group_off = 64;
    g.w_group_off = 64;
    g.epi = EPI_RESID_F32;
    g.act = ACT_GELU;
    g.bias = w->pos_b;
    g.out = pl.h;
    g.resid = pl.h;
    g.ldo = H;
    g.out_batch_rows = pl.frames;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  // ---- transformer layers (stable layer norm = pre-LN)
  iota_scaled_kernel<<<(B + 1 + 255) / 256, 256, 0, stream>>>(pl.cu, B + 1, pl.frames);
  B2S_LAUNCH_CHECK();
  rc = encoder_stack(w->layers, w->num_layers, H, F, w->heads, eps, pl.h, pl.xn, pl.qkv, pl.ao, pl.ff, pl.cu, B,
                     pl.frames, fmt, stream);
  if (rc != B2S_OK) return rc;
  if (last_hidden != nullptr) {
    B2S_CUDA_CHECK(cudaMemcpyAsync(last_hidden, pl.h, rows * H * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  // ---- final LN + AvgPool1d + projector
  rc = layernorm_avgpool_fwd(pl.h, w->final_ln_g, w->final_ln_b, eps, pl.pooled_x, B, pl.frames, H, w->pool_kernel,
                             w->pool_stride, pl.pooled, fmt, stream);
  if (rc != B2S_OK) return rc;
  {
    GemmArgs g = plain_gemm(pl.pooled_x, w->proj_w, static_cast<long long>(B) * pl.pooled, w->llm_dim, H);
    g.epi = EPI_F32;
    g.bias = w->proj_b;
    g.out = audio_embeds;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  return B2S_OK;
}

// --------------------------------------------------------------------------------------------- Whisper
namespace {
struct WhisperPlan {
  int frames_in, frames, pooled;
  void* x0;   // bf16 [B, T+2, mel]   zero-padded channels-last log-mel
  void* x1;   // bf16 [B, T+2, H]     zero-padded conv1 output
  float* h;
  void *xn, *qkv, *ao, *ff, *pooled_x;
  int* cu;
  size_t bytes;
};
int plan_whisper(const b2s_whisper_weights* w, int batches, void* ws, size_t ws_bytes, WhisperPlan* pl) {
  pl->frames_in = 2 * w->max_positions;
  pl->frames = w->max_positions;
  pl->pooled = pl->frames >= w->pool_kernel ? (pl->frames - w->pool_kernel) / w->pool_stride + 1 : 0;
  const size_t B = batches, H = w->hidden;
  const size_t rows = B * pl->frames;
  Carver c(ws, ws_bytes);
  pl->x0 = c.take(B * (pl->frames_in + 2) * w->mel_bins * 2 + 4096);
  pl->x1 = c.take(B * (pl->frames_in + 2) * H * 2 + 4096);
  pl->h = reinterpret_cast<float*>(c.take(rows * H * 4));
  pl->xn = c.take(rows * H * 2);
  pl->qkv = c.take(rows * 3 * H * 2);
  pl->ao = c.take(rows * H * 2);
  pl->ff = c.take(rows * w->ffn * 2);
  pl->pooled_x = c.take(B * (pl->pooled > 0 ? pl->pooled : 1) * H * 2);
  pl->cu = reinterpret_cast<int*>(c.take((B + 1) * sizeof(int)));
  pl->bytes = c.off + 256;
  if (ws != nullptr && !c.ok()) {
    set_last_error("whisper workspace too small: need %zu bytes, got %zu", pl->bytes, ws_bytes);
    return B2S_ERR_INVALID;
  }
  return B2S_OK;
}
}  // namespace

size_t whisper_workspace_bytes(const b2s_whisper_weights* w, int batches) {
  if (w == nullptr || batches <= 0) return 0;
  WhisperPlan pl;
  plan_whisper(w, batches, nullptr, 0, &pl);
  return pl.bytes;
}

int whisper_forward(const b2s_whisper_weights* w, const float* mel, int batches, int frames_in, void* workspace,
                    size_t workspace_bytes, float* audio_embeds, float* last_hidden, cudaStream_t stream) {
  B2S_REQUIRE(w && mel && workspace && audio_embeds, "whisper_forward: null pointer");
  B2S_REQUIRE(batches > 0, "whisper_forward: empty batch");
  // WhisperEncoder.forward raises unless the mel input has exactly max_source_positions * 2 frames
  // (TF/models/whisper/modeling_whisper.py:613-617)
  B2S_REQUIRE(frames_in == 2 * w->max_positions,
              "Whisper expects the mel input features to be of length %d, but found %d", 2 * w->max_positions,
              frames_in);
  B2S_REQUIRE(w->hidden % 256 == 0 && w->hidden / w->heads == 64 && w->mel_bins % 8 == 0,
              "whisper_forward: hidden %% 256, head_dim 64 and mel_bins %% 8 required");
  WhisperPlan pl;
  int rc = plan_whisper(w, batches, workspace, workspace_bytes, &pl);
  if (rc != B2S_OK) return rc;
  B2S_REQUIRE(pl.pooled > 0, "whisper_forward: too few frames to pool");
  const int B = batches, H = w->hidden, F = w->ffn, C = w->mel_bins, T = frames_in, fmt = w->fmt;
  const long long rows = static_cast<long long>(B) * pl.frames;

  rc = mel_to_padded_cl(mel, pl.x0, B, C, T, fmt, stream);
  if (rc != B2S_OK) return rc;
  // conv1: Conv1d(mel -> H, k=3, pad=1) + GELU; output row t lands on padded row t+1 of x1, whose first and last
  // rows stay zero (they are conv2's padding)
  B2S_CUDA_CHECK(cudaMemsetAsync(pl.x1, 0, static_cast<size_t>(B) * (T + 2) * H * 2, stream));
  {
    GemmArgs g{};
    g.A = pl.x0;
    g.a_dim0 = 3 * C;  // window view over the padded input: row t = x[t : t+3, :]
    g.a_row_stride = C;
    g.a_batch_stride = static_cast<long long>(T + 2) * C;
    g.a_rows = T;
    g.W = w->conv1_w;
    g.w_rows = H;
    g.w_cols = 3 * C;
    g.M = T;
    g.N = H;
    g.batches = B;
    g.groups = 1;
    g.taps = 1;
    g.k_per_tap = 3 * C;
    g.epi = EPI_BF16;
    g.act = ACT_GELU;
    g.bias = w->conv1_b;
    g.out = reinterpret_cast<__nv_bfloat16*>(pl.x1) + H;  // skip the leading zero row
    g.ldo = H;
    g.out_batch_rows = T + 2;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  // conv2: Conv1d(H -> H, k=3, stride=2, pad=1) + GELU, + positional table -> fp32 residual stream
  {
    GemmArgs g{};
    g.A = pl.x1;
    g.a_dim0 = 3 * H;  // row t = x1_padded[2t : 2t+3, :]
    g.a_row_stride = 2LL * H;
    g.a_batch_stride = static_cast<long long>(T + 2) * H;
    g.a_rows = pl.frames;
    g.W = w->conv2_w;
    g.w_rows = H;
    g.w_cols = 3 * H;
    g.M = pl.frames;
    g.N = H;
    g.batches = B;
    g.groups = 1;
    g.taps = 1;
    g.k_per_tap = 3 * H;
    g.epi = EPI_RESID_F32;
    g.act = ACT_GELU;
    g.bias = w->conv2_b;
    g.out = pl.h;
    g.resid = w->pos_emb;
    g.resid_bcast = 1;
    g.ldo = H;
    g.out_batch_rows = pl.frames;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  iota_scaled_kernel<<<(B + 1 + 255) / 256, 256, 0, stream>>>(pl.cu, B + 1, pl.frames);
  B2S_LAUNCH_CHECK();
  rc = encoder_stack(w->layers, w->num_layers, H, F, w->heads, w->ln_eps, pl.h, pl.xn, pl.qkv, pl.ao, pl.ff, pl.cu, B,
                     pl.frames, fmt, stream);
  if (rc != B2S_OK) return rc;
  if (last_hidden != nullptr) {
    B2S_CUDA_CHECK(cudaMemcpyAsync(last_hidden, pl.h, rows * H * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  rc = layernorm_avgpool_fwd(pl.h, w->final_ln_g, w->final_ln_b, w->ln_eps, pl.pooled_x, B, pl.frames, H,
                             w->pool_kernel, w->pool_stride, pl.pooled, fmt, stream);
  if (rc != B2S_OK) return rc;
  {
    GemmArgs g = plain_gemm(pl.pooled_x, w->proj_w, static_cast<long long>(B) * pl.pooled, w->llm_dim, H);
    g.epi = EPI_F32;
    g.bias = w->proj_b;
    g.out = audio_embeds;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  return B2S_OK;
}

// --------------------------------------------------------------------------------------------- Llama
namespace {
struct LlamaPlan {
  void* xn;   // bf16 [rows, H]
  void* qkv;  // bf16 [rows, (Hq+2Hkv)*D]
  void* ao;   // bf16 [rows, Hq*D]
  void* act;  // bf16 [rows, F]
  void* xf;   // bf16 [logit_rows, H]
  // last layer on the consumed rows only (see llama_prefill_kv)
  float* hsel;  // fp32 [logit_rows, H]
  void* aosel;  // bf16 [logit_rows, Hq*D]
  size_t bytes;
};
int plan_llama(const b2s_llama_weights* w, long long rows, long long logit_rows, void* ws, size_t ws_bytes,
               LlamaPlan* pl) {
  const size_t qkv_cols = static_cast<size_t>(w->heads + 2 * w->kv_heads) * w->head_dim;
  Carver c(ws, ws_bytes);
  pl->xn = c.take(rows * w->hidden * 2);
  pl->qkv = c.take(rows * qkv_cols * 2);
  pl->ao = c.take(static_cast<size_t>(rows) * w->heads * w->head_dim * 2);
  pl->act = c.take(static_cast<size_t>(rows) * w->ffn * 2);
  pl->xf = c.take(static_cast<size_t>(logit_rows > 0 ? logit_rows : 1) * w->hidden * 2);
  pl->hsel = reinterpret_cast<float*>(c.take(static_cast<size_t>(logit_rows > 0 ? logit_rows : 1) * w->hidden * 4));
  pl->aosel = c.take(static_cast<size_t>(logit_rows > 0 ? logit_rows : 1) * w->heads * w->head_dim * 2);
  pl->bytes = c.off + 256;
  if (ws != nullptr && !c.ok()) {
    set_last_error("llama workspace too small: need %zu bytes, got %zu", pl->bytes, ws_bytes);
    return B2S_ERR_INVALID;
  }
  return B2S_OK;
}
}  // namespace

size_t llama_workspace_bytes(const b2s_llama_weights* w, int rows, int logit_rows) {
  if (w == nullptr || rows <= 0) return 0;
  LlamaPlan pl;
  plan_llama(w, rows, logit_rows, nullptr, 0, &pl);
  return pl.bytes;
}

int kv_cache_store(const b2s_llama_weights* w, int layer, const void* qkv, long long rows, void* kv_cache, int kv_slots,
                   const int* slot_of_row, cudaStream_t stream);  // decode.cu

int llama_prefill_kv(const b2s_llama_weights* w, float* h, int rows, const int* cu_seqlens, int num_seqs,
                     int max_seqlen, const int* positions, const int* logit_rows_index, int logit_rows,
                     void* logits_bf16, const int* tap_layers, int num_taps, const int* tap_rows_a,
                     const int* tap_rows_b, int pairs, float* fd_sq, float* all_hidden, void* kv_cache, int kv_slots,
                     const int* kv_slot_of_row, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                     int shared_prefix_len = 0);

int llama_prefill(const b2s_llama_weights* w, float* h, int rows, const int* cu_seqlens, int num_seqs, int max_seqlen,
                  const int* positions, const int* logit_rows_index, int logit_rows, void* logits_bf16,
                  const int* tap_layers, int num_taps, const int* tap_rows_a, const int* tap_rows_b, int pairs,
                  float* fd_sq, float* all_hidden, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                  int shared_prefix_len) {
  return llama_prefill_kv(w, h, rows, cu_seqlens, num_seqs, max_seqlen, positions, logit_rows_index, logit_rows,
                          logits_bf16, tap_layers, num_taps, tap_rows_a, tap_rows_b, pairs, fd_sq, all_hidden, nullptr, 0,
                          nullptr, workspace, workspace_bytes, stream, shared_prefix_len);
}

// llama_prefill that additionally leaves every layer's post-RoPE k | v rows in a KV cache (decode.cu) so a greedy
// decode loop can continue from the prompt (REF/inference.py:55-74).
int llama_prefill_kv(const b2s_llama_weights* w, float* h, int rows, const int* cu_seqlens, int num_seqs,
                     int max_seqlen, const int* positions, const int* logit_rows_index, int logit_rows,
                     void* logits_bf16, const int* tap_layers, int num_taps, const int* tap_rows_a,
                     const int* tap_rows_b, int pairs, float* fd_sq, float* all_hidden, void* kv_cache, int kv_slots,
                     const int* kv_slot_of_row, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                     int shared_prefix_len) {
  B2S_REQUIRE(kv_cache == nullptr || (kv_slots > 0 && kv_slot_of_row != nullptr), "llama_prefill: bad KV cache arguments");
  B2S_REQUIRE(shared_prefix_len == 0 || kv_cache == nullptr, "llama_prefill: the shared prefix is not combined with a KV cache");
  B2S_REQUIRE(w && h && cu_seqlens && positions && workspace, "llama_prefill: null pointer");
  B2S_REQUIRE(rows > 0 && num_seqs > 0 && max_seqlen > 0, "llama_prefill: empty batch");
  B2S_REQUIRE(w->head_dim == 128, "llama_prefill: head_dim must be 128 (got %d)", w->head_dim);
  B2S_REQUIRE(w->hidden % 256 == 0 && w->ffn % 64 == 0, "llama_prefill: hidden %% 256, ffn %% 64 required");
  B2S_REQUIRE(logit_rows == 0 || (logit_rows_index && logits_bf16), "llama_prefill: logits requested without buffers");
  B2S_REQUIRE(num_taps == 0 || (tap_layers && tap_rows_a && tap_rows_b && fd_sq), "llama_prefill: taps without buffers");
  LlamaPlan pl;
  int rc = plan_llama(w, rows, logit_rows, workspace, workspace_bytes, &pl);
  if (rc != B2S_OK) return rc;
  const int H = w->hidden, D = w->head_dim, Hq = w->heads, Hkv = w->kv_heads, F = w->ffn, fmt = w->fmt;
  const int qkv_cols = (Hq + 2 * Hkv) * D;
  const float scale = 1.0f / sqrtf(static_cast<float>(D));

  const size_t h_bytes = static_cast<size_t>(rows) * H * sizeof(float);
  // Only the rows whose logits are requested leave the last layer (REF/model/audio_llama.py:67 computes all of them;
  // the trainer reads the last R, REF/trainer.py:334,350-351, generate() the last one): after the last layer's
  // attention -- which needs every row's K / V -- the out-projection, the MLP and the final norm run on those rows
  // only. Not when the caller wants every hidden state. `h` then keeps the INPUT of the last layer on exit.
  const bool last_on_selected = all_hidden == nullptr && logit_rows > 0 && logit_rows < rows;
  for (int l = 0; l < w->num_layers; ++l) {
    if (all_hidden != nullptr) {  // output_hidden_states: hidden_states[l] = input of layer l
      B2S_CUDA_CHECK(cudaMemcpyAsync(all_hidden + static_cast<size_t>(l) * rows * H, h, h_bytes,
                                     cudaMemcpyDeviceToDevice, stream));
    }
    for (int t = 0; t < num_taps; ++t) {
      if (tap_layers[t] == l && pairs > 0) {
        rc = rowpair_sqdiff_fwd(h, tap_rows_a, tap_rows_b, fd_sq + static_cast<long long>(t) * pairs, pairs, H, stream);
        if (rc != B2S_OK) return rc;
      }
    }
    const b2s_llama_layer& L = w->layers[l];
    rc = rmsnorm_fwd(h, L.ln1_w, w->rms_eps, pl.xn, rows, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    {
      GemmArgs g = plain_gemm(pl.xn, L.wqkv, rows, qkv_cols, H);
      g.epi = EPI_ROPE;
      g.out = pl.qkv;
      g.rope_cs = w->rope_cs;
      g.positions = positions;
      g.rope_cols = (Hq + Hkv) * D;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    if (kv_cache != nullptr) {
      rc = kv_cache_store(w, l, pl.qkv, rows, kv_cache, kv_slots, kv_slot_of_row, stream);
      if (rc != B2S_OK) return rc;
    }
    {
      const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(pl.qkv);
      rc = attention_fwd(qkv, qkv + Hq * D, qkv + (Hq + Hkv) * D, qkv_cols, pl.ao, Hq * D, cu_seqlens, num_seqs,
                         max_seqlen, rows, Hq, Hkv, D, scale, 1, nullptr, fmt, stream, nullptr, shared_prefix_len);
      if (rc != B2S_OK) return rc;
    }
    const bool sel = last_on_selected && l == w->num_layers - 1;
    const long long m = sel ? logit_rows : rows;  // rows that go through the rest of this layer
    float* hs = sel ? pl.hsel : h;
    const void* ao = pl.ao;
    if (sel) {
      rc = gather_rows_f32(h, logit_rows_index, pl.hsel, m, H, stream);
      if (rc != B2S_OK) return rc;
      rc = gather_rows_bf16(pl.ao, logit_rows_index, pl.aosel, m, Hq * D, stream);
      if (rc != B2S_OK) return rc;
      ao = pl.aosel;
    }
    {
      GemmArgs g = plain_gemm(ao, L.wo, m, H, Hq * D);
      g.epi = EPI_RESID_F32;
      g.out = hs;
      g.resid = hs;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    rc = rmsnorm_fwd(hs, L.ln2_w, w->rms_eps, pl.xn, m, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    {
      GemmArgs g = plain_gemm(pl.xn, L.wgu, m, 2 * F, H);
      g.epi = EPI_SWIGLU;
      g.out = pl.act;
      g.ldo = F;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    {
      GemmArgs g = plain_gemm(pl.act, L.wd, m, H, F);
      g.epi = EPI_RESID_F32;
      g.out = hs;
      g.resid = hs;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
  }
  for (int t = 0; t < num_taps; ++t) {
    if (tap_layers[t] == w->num_layers && pairs > 0) {
      set_last_error("llama_prefill: tap at the post-norm output (index num_layers) is not supported");
      return B2S_ERR_UNSUPPORTED;
    }
  }
  if (all_hidden != nullptr) {  // hidden_states[-1] = output of the final norm (rounded through bf16)
    rc = rmsnorm_fwd(h, w->final_norm_w, w->rms_eps, pl.xn, rows, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    rc = cast_h16_to_f32(pl.xn, all_hidden + static_cast<size_t>(w->num_layers) * rows * H,
                          static_cast<long long>(rows) * H, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  if (logit_rows > 0) {
    if (last_on_selected) rc = rmsnorm_fwd(pl.hsel, w->final_norm_w, w->rms_eps, pl.xf, logit_rows, H, fmt, stream);
    else rc = rmsnorm_gather_fwd(h, logit_rows_index, w->final_norm_w, w->rms_eps, pl.xf, logit_rows, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    GemmArgs g = plain_gemm(pl.xf, w->lm_head, logit_rows, w->vocab, H);
    g.epi = EPI_BF16;
    g.out = logits_bf16;
    g.a_fmt = g.w_fmt = fmt;
    g.out_fmt = 0;  // logits are bf16 whatever the operand format (what the fused loss kernel reads)
    rc = gemm_bf16_launch(g, stream);
    if (rc != B2S_OK) return rc;
  }
  return B2S_OK;
}

}  // namespace b2s
