// train_enc.cu -- training-mode forward (activations kept) and full backward of the trainable HuBERT audio encoder
// (REF/trainer.py:98-105: every AudioEncoder parameter is optimised; REF/trainer.py:278,373-374: the loss gradient
// comes back through the projected audio embeddings).
//
//   hubert_forward_train : hubert_forward (models.cu) with every activation the backward needs written into one
//                          caller-owned "saved" region (no recomputation except the cheap LayerNorm outputs).
//   hubert_backward      : d(audio_embeds) -> fp32 gradient accumulators for all 422 trainable tensors
//                          (TF/models/hubert/modeling_hubert.py:127-231,45-92,505-624; REF/model/audio_encoder.py:59-87).
//
// Every dense contraction is the tcgen05 GEMM: dgrad reads the forward's bf16 weights MN-major (no transposed
// copies, the weights change every optimizer step), wgrad reads dY and X MN-major and accumulates with fp32 atomics
// (split-K fills the 148 SMs; accumulation across micro-batches = REF/trainer.py:372-380 comes for free).
// Train-mode regularisers (TF/.../modeling_hubert.py:223-230 feature-projection dropout, :842-886 SpecAugment, :585-587
// hidden dropout, :596-599 LayerDrop, :254 attention dropout, :351-368 FFN dropouts, :383-393 layer dropout) are applied
// when the caller passes a b2s_encoder_regularizers block; with NULL the step is the
// deterministic one. Elementwise dropouts are GEMM-epilogue fusions, masks are regenerated from counters (rng.cuh).
#include "../../include/b2s.h"
#include "b2s_common.cuh"
#include "gemm_sm100.cuh"
#include "ops.cuh"

// Ragged micro-batches (utterances of different lengths, `samples_per_utt`): the convolutional front end, the
// feature projection, the positional convolution and the final LayerNorm + AvgPool run on the zero-padded [B, T_max]
// layout (valid frames never read padded ones; the positional conv sees zeroed tail rows = its own zero padding),
// the transformer stack runs on the PACKED valid rows with per-utterance cu_seqlens (attention is varlen already),
// and two row gathers convert between the layouts. Every utterance gets exactly the numbers it would get alone.
#include <cstring>
#include <vector>

namespace b2s {
namespace {

struct Carve {
  uint8_t* base;
  size_t off = 0, cap;
  Carve(void* b, size_t c) : base(reinterpret_cast<uint8_t*>(b)), cap(c) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

int conv_len(int in, int k, int s) { return in < k ? 0 : (in - k) / s + 1; }

struct Saved {
  int t[8], frames, pooled;
  void* conv_x[7];     // bf16 [B, t[i+1], 512] : output of conv layer i
  float* conv_pre[6];  // fp32 [B, t[i+2], 512] : pre-norm output of conv layer i+1
  void* hp_bf;         // bf16 [rows, H] feature-projection output (positional conv input)
  void* pos_pre;       // bf16 [rows, H] positional conv output before GELU
  float* h;            // fp32 [L+1][rows][H]
  float* h_mid;        // fp32 [L][rows][H]
  void* qkv;           // bf16 [L][rows][3H]
  void* ao;            // bf16 [L][rows][H]
  float* lse;          // fp32 [L][rows][heads]
  void* ff_pre;        // bf16 [L][rows][F]
  void* ff;            // bf16 [L][rows][F]
  void* pooled_x;      // bf16 [B*pooled][H]
  void* xn;            // bf16 [rows][max(H,512)] transient
  int* cu;
  // ragged batches only (rows = B * frames is the PADDED row count, packed rows <= rows)
  float* hpad;         // fp32 [rows][H] padded-layout stream: h0 before the gather, then h[L] scattered back
  float* zeros_h;      // fp32 [H] zeros (tail-row "embedding" for mask_rows_f32)
  int* pack_idx;       // int [rows] packed row -> padded row
  int* unpack_idx;     // int [rows] padded row -> packed row, -1 on tail rows
  unsigned char* tail; // [rows] 1 on tail rows
  size_t bytes;
};

void plan_saved(const b2s_hubert_weights* w, int batches, int samples, void* ws, size_t cap, Saved* s) {
  s->t[0] = samples;
  s->t[1] = conv_len(samples, 10, 5);
  for (int i = 0; i < 6; ++i) s->t[i + 2] = conv_len(s->t[i + 1], w->conv_k[i], w->conv_stride[i]);
  s->frames = s->t[7];
  s->pooled = s->frames >= w->pool_kernel ? (s->frames - w->pool_kernel) / w->pool_stride + 1 : 0;
  const size_t B = batches, H = w->hidden, F = w->ffn, L = w->num_layers;
  const size_t rows = B * (s->frames > 0 ? s->frames : 1);
  Carve c(ws, cap);
  for (int i = 0; i < 7; ++i) s->conv_x[i] = c.take(B * (s->t[i + 1] > 0 ? s->t[i + 1] : 1) * 512 * 2 + 4096);
  for (int i = 0; i < 6; ++i)
    s->conv_pre[i] = reinterpret_cast<float*>(c.take(B * (s->t[i + 2] > 0 ? s->t[i + 2] : 1) * 512 * 4));
  s->hp_bf = c.take(rows * H * 2 + 4096);
  s->pos_pre = c.take(rows * H * 2);
  s->h = reinterpret_cast<float*>(c.take((L + 1) * rows * H * 4));
  s->h_mid = reinterpret_cast<float*>(c.take(L * rows * H * 4));
  s->qkv = c.take(L * rows * 3 * H * 2);
  s->ao = c.take(L * rows * H * 2);
  s->lse = reinterpret_cast<float*>(c.take(L * rows * w->heads * 4));
  s->ff_pre = c.take(L * rows * F * 2);
  s->ff = c.take(L * rows * F * 2);
  s->pooled_x = c.take(B * (s->pooled > 0 ? s->pooled : 1) * H * 2);
  s->xn = c.take(rows * (H > 512 ? H : 512) * 2);
  s->cu = reinterpret_cast<int*>(c.take((B + 1) * sizeof(int)));
  s->hpad = reinterpret_cast<float*>(c.take(rows * H * 4));
  s->zeros_h = reinterpret_cast<float*>(c.take(H * 4));
  s->pack_idx = reinterpret_cast<int*>(c.take(rows * sizeof(int)));
  s->unpack_idx = reinterpret_cast<int*>(c.take(rows * sizeof(int)));
  s->tail = reinterpret_cast<unsigned char*>(c.take(rows));
  s->bytes = c.off + 256;
}

// host-side geometry of a ragged batch
struct Ragged {
  bool on = false;
  long long rows_packed = 0;
  int max_frames = 0;
  std::vector<int> frames, cu, pack, unpack;
  std::vector<unsigned char> tail;
};

int plan_ragged(const b2s_hubert_weights* w, int batches, int samples, const int* samples_per_utt, int frames_pad,
                Ragged* r) {
  r->on = samples_per_utt != nullptr;
  if (!r->on) return B2S_OK;
  r->frames.resize(batches);
  r->cu.assign(batches + 1, 0);
  r->pack.clear();
  r->unpack.assign(static_cast<size_t>(batches) * frames_pad, -1);
  r->tail.assign(static_cast<size_t>(batches) * frames_pad, 1);
  for (int b = 0; b < batches; ++b) {
    B2S_REQUIRE(samples_per_utt[b] > 0 && samples_per_utt[b] <= samples,
                "ragged batch: utterance %d has %d samples (padded length %d)", b, samples_per_utt[b], samples);
    int t = conv_len(samples_per_utt[b], 10, 5);
    for (int i = 0; i < 6; ++i) t = conv_len(t, w->conv_k[i], w->conv_stride[i]);
    B2S_REQUIRE(t >= w->pool_kernel, "ragged batch: utterance %d is too short (%d frames)", b, t);
    r->frames[b] = t;
    r->cu[b + 1] = r->cu[b] + t;
    if (t > r->max_frames) r->max_frames = t;
    for (int f = 0; f < t; ++f) {
      r->unpack[static_cast<size_t>(b) * frames_pad + f] = static_cast<int>(r->pack.size());
      r->tail[static_cast<size_t>(b) * frames_pad + f] = 0;
      r->pack.push_back(b * frames_pad + f);
    }
  }
  r->rows_packed = r->cu[batches];
  return B2S_OK;
}

// Pinned staging for the ragged index arrays: cudaMemcpyAsync from pageable memory synchronises the stream before the
// copy starts, which would stop the host from running ahead of the GPU. A small ring of pinned slots (one event each,
// waited on before a slot is reused) keeps the upload asynchronous.
class PinnedRing {
 public:
  // copies `bytes` from `src` into a pinned slot and enqueues the H2D copy to `dst`
  int upload(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    Slot& sl = slots_[next_];
    next_ = (next_ + 1) % kSlots;
    if (sl.event != nullptr) B2S_CUDA_CHECK(cudaEventSynchronize(sl.event));
    if (sl.cap < bytes) {
      if (sl.ptr != nullptr) B2S_CUDA_CHECK(cudaFreeHost(sl.ptr));
      sl.cap = (bytes + 65535) & ~static_cast<size_t>(65535);
      B2S_CUDA_CHECK(cudaHostAlloc(&sl.ptr, sl.cap, cudaHostAllocDefault));
    }
    if (sl.event == nullptr) B2S_CUDA_CHECK(cudaEventCreateWithFlags(&sl.event, cudaEventDisableTiming));
    memcpy(sl.ptr, src, bytes);
    B2S_CUDA_CHECK(cudaMemcpyAsync(dst, sl.ptr, bytes, cudaMemcpyHostToDevice, stream));
    B2S_CUDA_CHECK(cudaEventRecord(sl.event, stream));
    return B2S_OK;
  }

 private:
  static constexpr int kSlots = 16;
  struct Slot {
    void* ptr = nullptr;
    size_t cap = 0;
    cudaEvent_t event = nullptr;
  };
  Slot slots_[kSlots];
  int next_ = 0;
};
PinnedRing g_ring;  // one host thread per process drives the library (include/b2s.h: not thread-safe per handle)

struct BwdWs {
  float *dh, *dxn_f, *dpool, *delta;
  void *dyb, *dbig, *dsm, *xn, *da, *dpre, *dcol, *dxa, *dxb;
  float* dh_pad;  // ragged batches: the padded-layout twin of dh (fp32 [rows][H]) ...
  void* dyb_pad;  // ... and of dyb (bf16)
  size_t bytes;
};

void plan_bwd(const b2s_hubert_weights* w, int batches, const Saved& s, void* ws, size_t cap, BwdWs* p) {
  const size_t B = batches, H = w->hidden, F = w->ffn;
  const size_t rows = B * (s.frames > 0 ? s.frames : 1);
  const size_t big = F > 3 * H ? F : 3 * H;
  const size_t np = B * (s.pooled > 0 ? s.pooled : 1);
  Carve c(ws, cap);
  p->dh = reinterpret_cast<float*>(c.take(rows * H * 4));
  p->dxn_f = reinterpret_cast<float*>(c.take(rows * H * 4));
  p->dpool = reinterpret_cast<float*>(c.take(np * H * 4));
  p->delta = reinterpret_cast<float*>(c.take(rows * w->heads * 4));
  p->dyb = c.take(rows * H * 2 + 4096);
  p->dbig = c.take(rows * big * 2 + 4096);
  p->dsm = c.take(rows * H * 2 + 4096);
  p->xn = c.take(rows * (H > 512 ? H : 512) * 2 + 4096);
  p->da = c.take(np * w->llm_dim * 2 + 4096);
  const size_t t1 = s.t[1] > 0 ? s.t[1] : 1, t2 = s.t[2] > 0 ? s.t[2] : 1;
  p->dpre = c.take(B * t2 * 512 * 2 + 4096);
  p->dcol = c.take(B * t2 * 3 * 512 * 2 + 4096);
  p->dxa = c.take(B * t1 * 512 * 2 + 4096);
  p->dxb = c.take(B * t1 * 512 * 2 + 4096);
  p->dh_pad = reinterpret_cast<float*>(c.take(rows * H * 4));
  p->dyb_pad = c.take(rows * H * 2 + 4096);
  p->bytes = c.off + 256;
}

__global__ void iota_scaled(int* out, int n, int scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i * scale;
}

GemmArgs lin(const void* A, const void* W, long long M, int N, int K) {
  GemmArgs g{};
  g.A = A;
  g.a_dim0 = K;
  g.a_row_stride = K;
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
  return g;
}

// dX[M, K] = dY[M, N] . W[N, K]   (W in its forward layout, read MN-major)
int dgrad(const void* dy, const void* W, long long M, int N, int K, int epi, void* out, int fmt, cudaStream_t st) {
  GemmArgs g{};
  g.A = dy;
  g.a_dim0 = N;
  g.a_row_stride = N;
  g.a_rows = static_cast<int>(M);
  g.W = W;
  g.w_rows = N;
  g.w_cols = K;
  g.b_mn = 1;
  g.M = static_cast<int>(M);
  g.N = K;
  g.batches = 1;
  g.groups = 1;
  g.taps = 1;
  g.k_per_tap = N;
  g.epi = epi;
  g.out = out;
  g.ldo = K;
  return gemm_launch_fmt(g, fmt, st);
}

// dW[N, K] += dY[rows, N]^T . X[rows, K]
int wgrad(const void* dy, const void* x, long long rows, int N, int K, float* dW, int fmt, cudaStream_t st) {
  GemmArgs g{};
  g.A = dy;
  g.a_dim0 = N;
  g.a_row_stride = N;
  g.a_rows = static_cast<int>(rows);
  g.a_mn = 1;
  g.W = x;
  g.w_rows = static_cast<int>(rows);
  g.w_cols = K;
  g.w_row_stride = K;
  g.b_mn = 1;
  g.M = N;
  g.N = K;
  g.batches = 1;
  g.groups = 1;
  g.taps = 1;
  g.k_per_tap = static_cast<int>(rows);
  g.k_batches = 1;
  g.epi = EPI_ACCUM_F32;
  g.out = dW;
  g.ldo = K;
  return gemm_launch_fmt(g, fmt, st);
}

#define RC(expr)                 \
  do {                           \
    int _rc = (expr);            \
    if (_rc != B2S_OK) return _rc; \
  } while (0)


// ---- pieces shared by the HuBERT and Whisper encoders: the pre-LN transformer stack and the pool + projector head
struct StackBufs {   // per-layer saved activations (see Saved)
  float *h, *h_mid, *lse;
  void *qkv, *ao, *ff_pre, *ff;
};
struct StackScratch {  // backward scratch
  float *dh, *delta;
  void *dyb, *dbig, *dsm, *xn;
};

void set_drop(GemmArgs& g, const DropSpec& d) {
  g.drop_k1 = d.k1;
  g.drop_k2 = d.k2;
  g.drop_thresh = d.thresh;
  g.drop_inv_keep = d.inv_keep;
}

AttnDrop attn_drop(const b2s_encoder_regularizers* reg, int l) {
  AttnDrop a{};
  if (reg != nullptr) {
    a.seed = reg->seed;
    a.site = site_attn_prob(l);
    a.thresh = drop_threshold(reg->p_attention);
    a.inv_keep = reg->p_attention < 1.f ? 1.0f / (1.0f - reg->p_attention) : 0.f;
  }
  return a;
}

bool layer_skipped(const b2s_encoder_regularizers* reg, int l) {
  return reg != nullptr && reg->layer_skip != nullptr && reg->layer_skip[l] != 0;
}

int stack_forward_train(const b2s_encoder_layer* layers, int L, int H, int F, int heads, float eps, const StackBufs& s,
                        void* xn, const int* cu, int B, int frames, int fmt, cudaStream_t stream,
                        const b2s_encoder_regularizers* reg = nullptr, long long rows_packed = -1) {
  // `frames` is the longest sequence (attention grid); rows = all packed rows (B * frames unless the batch is ragged)
  const long long rows = rows_packed >= 0 ? rows_packed : static_cast<long long>(B) * frames;
  const size_t rH = static_cast<size_t>(rows) * H;
  for (int l = 0; l < L; ++l) {
    const b2s_encoder_layer& Ly = layers[l];
    float* h_in = s.h + l * rH;
    float* h_mid = s.h_mid + l * rH;
    float* h_out = s.h + (l + 1) * rH;
    if (layer_skipped(reg, l)) {  // LayerDrop: the layer is the identity for this micro-batch
      B2S_CUDA_CHECK(cudaMemcpyAsync(h_out, h_in, rH * sizeof(float), cudaMemcpyDeviceToDevice, stream));
      continue;
    }
    const AttnDrop adrop = attn_drop(reg, l);
    __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(s.qkv) + l * 3 * rH;
    __nv_bfloat16* ao = reinterpret_cast<__nv_bfloat16*>(s.ao) + l * rH;
    __nv_bfloat16* ffp = reinterpret_cast<__nv_bfloat16*>(s.ff_pre) + static_cast<size_t>(l) * rows * F;
    __nv_bfloat16* ff = reinterpret_cast<__nv_bfloat16*>(s.ff) + static_cast<size_t>(l) * rows * F;
    RC(layernorm_fwd(h_in, 0, Ly.ln1_g, Ly.ln1_b, eps, 0, xn, rows, H, fmt, stream));
    {
      GemmArgs g = lin(xn, Ly.wqkv, rows, 3 * H, H);
      g.epi = EPI_BF16;
      g.bias = Ly.bqkv;
      g.out = qkv;
      RC(gemm_launch_fmt(g, fmt, stream));
    }
    RC(attention_fwd(qkv, qkv + H, qkv + 2 * H, 3 * H, ao, H, cu, B, frames, rows, heads, heads, 64, 0.125f, 0,
                     s.lse + static_cast<size_t>(l) * rows * heads, fmt, stream, reg ? &adrop : nullptr));
    {
      GemmArgs g = lin(ao, Ly.wo, rows, H, H);
      g.epi = EPI_RESID_F32;
      g.bias = Ly.bo;
      g.out = h_mid;
      g.resid = h_in;
      if (reg) set_drop(g, make_drop_spec(reg->seed, site_attn_out(l), reg->p_hidden));
      RC(gemm_launch_fmt(g, fmt, stream));
    }
    RC(layernorm_fwd(h_mid, 0, Ly.ln2_g, Ly.ln2_b, eps, 0, xn, rows, H, fmt, stream));
    {
      GemmArgs g = lin(xn, Ly.w1, rows, F, H);
      g.epi = EPI_BF16;
      g.act = ACT_GELU;
      g.bias = Ly.b1;
      g.out = ff;
      g.out2 = ffp;
      g.ld2 = F;
      if (reg) set_drop(g, make_drop_spec(reg->seed, site_ff_act(l), reg->p_activation));
      RC(gemm_launch_fmt(g, fmt, stream));
    }
    {
      GemmArgs g = lin(ff, Ly.w2, rows, H, F);
      g.epi = EPI_RESID_F32;
      g.bias = Ly.b2;
      g.out = h_out;
      g.resid = h_mid;
      if (reg) set_drop(g, make_drop_spec(reg->seed, site_ff_out(l), reg->p_hidden));
      RC(gemm_launch_fmt(g, fmt, stream));
    }
  }
  return B2S_OK;
}

// on entry b.dh / b.dyb hold d(loss)/d(h[L]) (fp32 / bf16); on exit d(loss)/d(h[0])
int stack_backward(const b2s_encoder_layer* layers, const b2s_encoder_layer_grads* grads, int L, int H, int F, int heads,
                   float eps, const StackBufs& s, const StackScratch& b, const int* cu, int B, int frames, int fmt,
                   cudaStream_t stream, const b2s_encoder_regularizers* reg = nullptr, long long rows_packed = -1,
                   void* const* layer_done = nullptr) {
  const long long rows = rows_packed >= 0 ? rows_packed : static_cast<long long>(B) * frames;
  const size_t rH = static_cast<size_t>(rows) * H;
  const bool drop_h = reg != nullptr && reg->p_hidden > 0.f;
  // Without hidden dropout the bias gradients of W2 / Wo are column sums of dh itself; the LayerNorm backward that
  // writes dh accumulates them on the way (layernorm_bwd_ex's dh_colsum) instead of a colsum pass re-reading dh.
  bool b2_done = false;  // layer l's b2 gradient was produced by the layer above
  for (int l = L - 1; l >= 0; --l) {
    if (layer_skipped(reg, l)) {  // identity layer: the gradient passes through unchanged
      if (layer_done != nullptr && layer_done[l] != nullptr)
        B2S_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(layer_done[l]), stream));
      continue;
    }
    const b2s_encoder_layer& Ly = layers[l];
    const b2s_encoder_layer_grads& G = grads[l];
    const float* h_in = s.h + l * rH;
    const float* h_mid = s.h_mid + l * rH;
    const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(s.qkv) + l * 3 * rH;
    const __nv_bfloat16* ao = reinterpret_cast<const __nv_bfloat16*>(s.ao) + l * rH;
    const __nv_bfloat16* ffp = reinterpret_cast<const __nv_bfloat16*>(s.ff_pre) + static_cast<size_t>(l) * rows * F;
    const __nv_bfloat16* ff = reinterpret_cast<const __nv_bfloat16*>(s.ff) + static_cast<size_t>(l) * rows * F;
    // feed-forward: h_out = h_mid + drop(W2 drop(gelu(W1 LN2(h_mid) + b1)) + b2)
    if (drop_h) {  // the branch gradient is the masked copy; the residual path keeps b.dh
      RC(dropout_apply(nullptr, b.dyb, rows * H, make_drop_spec(reg->seed, site_ff_out(l), reg->p_hidden), fmt, stream));
      RC(colsum_accum(b.dyb, 1, G.b2, rows, H, fmt, stream));
    } else if (!b2_done) {
      RC(colsum_accum(b.dh, 0, G.b2, rows, H, fmt, stream));
    }
    b2_done = false;
    RC(wgrad(b.dyb, ff, rows, H, F, G.w2, fmt, stream));
    RC(dgrad(b.dyb, Ly.w2, rows, H, F, EPI_BF16, b.dbig, fmt, stream));
    {
      const DropSpec dact = reg ? make_drop_spec(reg->seed, site_ff_act(l), reg->p_activation) : DropSpec{};
      if (F % 2048 == 0) {  // the bias gradient of W1 is accumulated by the same pass
        RC(gelu_bwd(ffp, b.dbig, b.dbig, rows * F, fmt, stream, dact.thresh != 0u ? &dact : nullptr, G.b1, F));
      } else {
        RC(gelu_bwd(ffp, b.dbig, b.dbig, rows * F, fmt, stream, dact.thresh != 0u ? &dact : nullptr));
        RC(colsum_accum(b.dbig, 1, G.b1, rows, F, fmt, stream));
      }
    }
    RC(layernorm_fwd(h_mid, 0, Ly.ln2_g, Ly.ln2_b, eps, 0, b.xn, rows, H, fmt, stream));
    RC(wgrad(b.dbig, b.xn, rows, F, H, G.w1, fmt, stream));
    RC(dgrad(b.dbig, Ly.w1, rows, F, H, EPI_BF16, b.dsm, fmt, stream));
    RC(layernorm_bwd_ex(h_mid, 0, Ly.ln2_g, Ly.ln2_b, 0, eps, b.dsm, 1, b.dh, 1, b.dyb, G.ln2_g, G.ln2_b, rows, H,
                        fmt, stream, drop_h ? nullptr : G.bo));
    // attention: h_mid = h_in + drop(Wo attn(Wqkv LN1(h_in) + bqkv) + bo)
    if (drop_h) {
      RC(dropout_apply(nullptr, b.dyb, rows * H, make_drop_spec(reg->seed, site_attn_out(l), reg->p_hidden), fmt, stream));
      RC(colsum_accum(b.dyb, 1, G.bo, rows, H, fmt, stream));
    }
    RC(wgrad(b.dyb, ao, rows, H, H, G.wo, fmt, stream));
    RC(dgrad(b.dyb, Ly.wo, rows, H, H, EPI_BF16, b.dsm, fmt, stream));
    {
      __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(b.dbig);
      const AttnDrop adrop = attn_drop(reg, l);
      RC(attention_bwd(qkv, qkv + H, qkv + 2 * H, 3 * H, ao, H, b.dsm, H, s.lse + static_cast<size_t>(l) * rows * heads,
                       b.delta, dqkv, dqkv + H, dqkv + 2 * H, 3 * H, cu, B, frames, rows, heads, heads, 64, 0.125f, 0,
                       nullptr, fmt, stream, reg ? &adrop : nullptr));
    }
    RC(colsum_accum(b.dbig, 1, G.bqkv, rows, 3 * H, fmt, stream));
    RC(layernorm_fwd(h_in, 0, Ly.ln1_g, Ly.ln1_b, eps, 0, b.xn, rows, H, fmt, stream));
    RC(wgrad(b.dbig, b.xn, rows, 3 * H, H, G.wqkv, fmt, stream));
    RC(dgrad(b.dbig, Ly.wqkv, rows, 3 * H, H, EPI_BF16, b.dsm, fmt, stream));
    {
      int below = l - 1;  // the next layer down that runs: its W2 bias gradient is the column sum of the dh written here
      while (below >= 0 && layer_skipped(reg, below)) --below;
      float* b2_below = (!drop_h && below >= 0) ? grads[below].b2 : nullptr;
      RC(layernorm_bwd_ex(h_in, 0, Ly.ln1_g, Ly.ln1_b, 0, eps, b.dsm, 1, b.dh, 1, b.dyb, G.ln1_g, G.ln1_b, rows, H, fmt,
                          stream, b2_below));
      b2_done = b2_below != nullptr;
    }
    // every gradient of layer l has been enqueued: a communication stream may start exchanging them (training.py)
    if (layer_done != nullptr && layer_done[l] != nullptr)
      B2S_CUDA_CHECK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(layer_done[l]), stream));
  }
  return B2S_OK;
}

// projector + AvgPool + final LayerNorm backward: d_audio_embeds -> b.dh / b.dyb = d(loss)/d(h[L])
int head_backward(const void* proj_w, const float* final_ln_g, const float* final_ln_b, float* g_proj_w, float* g_proj_b,
                  float* g_ln_g, float* g_ln_b, const float* h_last, const void* pooled_x, const float* d_audio_embeds,
                  void* da, float* dpool, float* dxn_f, const StackScratch& b, int B, int frames, int pooled, int H,
                  int Cl, int pool_kernel, int pool_stride, float eps, int fmt, cudaStream_t stream) {
  const long long rows = static_cast<long long>(B) * frames;
  const long long np = static_cast<long long>(B) * pooled;
  RC(cast_f32_to_h16(d_audio_embeds, da, np * Cl, fmt, stream));
  RC(colsum_accum(d_audio_embeds, 0, g_proj_b, np, Cl, fmt, stream));
  RC(wgrad(da, pooled_x, np, Cl, H, g_proj_w, fmt, stream));
  RC(dgrad(da, proj_w, np, Cl, H, EPI_F32, dpool, fmt, stream));
  RC(avgpool_bwd(dpool, dxn_f, B, frames, H, pool_kernel, pool_stride, pooled, stream));
  RC(layernorm_bwd_ex(h_last, 0, final_ln_g, final_ln_b, 0, eps, dxn_f, 0, b.dh, 0, b.dyb, g_ln_g, g_ln_b, rows, H,
                      fmt,
                      stream));
  return B2S_OK;
}

}  // namespace

size_t hubert_saved_bytes(const b2s_hubert_weights* w, int batches, int samples) {
  if (w == nullptr || batches <= 0 || samples <= 0) return 0;
  Saved s;
  plan_saved(w, batches, samples, nullptr, 0, &s);
  return s.bytes;
}

size_t hubert_backward_workspace_bytes(const b2s_hubert_weights* w, int batches, int samples) {
  if (w == nullptr || batches <= 0 || samples <= 0) return 0;
  Saved s;
  plan_saved(w, batches, samples, nullptr, 0, &s);
  BwdWs b;
  plan_bwd(w, batches, s, nullptr, 0, &b);
  return b.bytes;
}

int check_regularizers(const b2s_encoder_regularizers* reg) {
  if (reg == nullptr) return B2S_OK;
  const float ps[4] = {reg->p_feat_proj, reg->p_hidden, reg->p_attention, reg->p_activation};
  for (float p : ps) B2S_REQUIRE(p >= 0.f && p < 1.f, "regularizers: dropout probabilities must lie in [0, 1)");
  B2S_REQUIRE(reg->time_mask == nullptr || reg->masked_spec_embed != nullptr,
              "regularizers: time_mask needs masked_spec_embed");
  return B2S_OK;
}

int hubert_forward_train(const b2s_hubert_weights* w, const float* wave, long long wave_stride, int batches, int samples,
                         const int* samples_per_utt, void* saved, size_t saved_bytes, float* audio_embeds,
                         const b2s_encoder_regularizers* reg, cudaStream_t stream) {
  B2S_REQUIRE(w && wave && saved && audio_embeds, "hubert_forward_train: null pointer");
  RC(check_regularizers(reg));
  B2S_REQUIRE(batches > 0 && samples > 0, "hubert_forward_train: empty batch");
  B2S_REQUIRE(w->hidden % 256 == 0 && w->hidden % w->heads == 0 && w->hidden / w->heads == 64,
              "hubert_forward_train: hidden/heads must give head_dim 64");
  B2S_REQUIRE(w->pos_groups > 0 && w->hidden / w->pos_groups == 64, "hubert_forward_train: positional conv needs 64 ch/group");
  Saved s;
  plan_saved(w, batches, samples, saved, saved_bytes, &s);
  B2S_REQUIRE(s.bytes <= saved_bytes, "hubert_forward_train: saved region too small: need %zu bytes, got %zu", s.bytes,
              saved_bytes);
  B2S_REQUIRE(s.frames > 0 && s.pooled > 0, "hubert_forward_train: audio too short (%d samples -> %d frames)", samples,
              s.frames);
  const int B = batches, H = w->hidden, F = w->ffn, L = w->num_layers, fmt = w->fmt;
  const long long rows = static_cast<long long>(B) * s.frames;
  const float eps = w->ln_eps;
  Ragged rg;
  RC(plan_ragged(w, B, samples, samples_per_utt, s.frames, &rg));
  float* const h0 = rg.on ? s.hpad : s.h;  // padded-layout stream the front end writes
  if (rg.on) {
    RC(g_ring.upload(s.cu, rg.cu.data(), (B + 1) * sizeof(int), stream));
    RC(g_ring.upload(s.pack_idx, rg.pack.data(), rg.pack.size() * sizeof(int), stream));
    RC(g_ring.upload(s.unpack_idx, rg.unpack.data(), rg.unpack.size() * sizeof(int), stream));
    RC(g_ring.upload(s.tail, rg.tail.data(), rg.tail.size(), stream));
    B2S_CUDA_CHECK(cudaMemsetAsync(s.zeros_h, 0, static_cast<size_t>(H) * 4, stream));
  }

  RC(conv0_ln_gelu_fwd(wave, wave_stride, B, samples, w->conv0_w, w->conv0_b, w->conv0_ln_g, w->conv0_ln_b, eps,
                       s.conv_x[0], s.t[1], fmt, stream));
  for (int i = 0; i < 6; ++i) {
    const int tin = s.t[i + 1], tout = s.t[i + 2];
    const int k = w->conv_k[i], sd = w->conv_stride[i];
    GemmArgs g{};
    g.A = s.conv_x[i];
    g.a_dim0 = k * 512;
    g.a_row_stride = static_cast<long long>(sd) * 512;
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
    g.bias = w->conv_b[i];
    g.out = s.conv_pre[i];
    g.ldo = 512;
    g.out_batch_rows = tout;
    RC(gemm_launch_fmt(g, fmt, stream));
    RC(layernorm_fwd(s.conv_pre[i], 0, w->conv_ln_g[i], w->conv_ln_b[i], eps, 1, s.conv_x[i + 1],
                     static_cast<long long>(B) * tout, 512, fmt, stream));
  }
  // feature projection -> h[0]
  RC(layernorm_fwd(s.conv_x[6], 1, w->fp_ln_g, w->fp_ln_b, eps, 0, s.xn, rows, 512, fmt, stream));
  {
    GemmArgs g = lin(s.xn, w->fp_w, rows, H, 512);
    g.epi = EPI_F32;
    g.bias = w->fp_b;
    g.out = h0;
    if (reg) set_drop(g, make_drop_spec(reg->seed, SITE_FEAT_PROJ, reg->p_feat_proj));
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  if (reg && reg->time_mask)  // SpecAugment: masked frames become masked_spec_embed (after the projection dropout)
    RC(mask_rows_f32(h0, reg->time_mask, reg->masked_spec_embed, rows, H, stream));
  if (rg.on)  // frames past an utterance's end become the zero padding its positional conv must see
    RC(mask_rows_f32(h0, s.tail, s.zeros_h, rows, H, stream));
  RC(cast_f32_to_h16(h0, s.hp_bf, rows * H, fmt, stream));
  {
    GemmArgs g{};
    g.A = s.hp_bf;
    g.a_dim0 = H;
    g.a_row_stride = H;
    g.a_batch_stride = static_cast<long long>(s.frames) * H;
    g.a_rows = s.frames;
    g.W = w->pos_w;
    g.w_rows = H;
    g.w_cols = w->pos_k * 64;
    g.M = s.frames;
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
    g.out = h0;
    g.resid = h0;
    g.ldo = H;
    g.out_batch_rows = s.frames;
    g.out2 = s.pos_pre;
    g.ld2 = H;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  if (reg)  // dropout(hidden + positional embedding): after the residual add, so not an epilogue of that GEMM
    RC(dropout_apply(h0, nullptr, rows * H, make_drop_spec(reg->seed, SITE_POS_ADD, reg->p_hidden), fmt, stream));
  const long long srows = rg.on ? rg.rows_packed : rows;  // rows of the transformer stack
  if (rg.on) {
    RC(gather_rows_f32(h0, s.pack_idx, s.h, srows, H, stream));  // padded -> packed valid rows
  } else {
    iota_scaled<<<(B + 1 + 255) / 256, 256, 0, stream>>>(s.cu, B + 1, s.frames);
    B2S_LAUNCH_CHECK();
  }
  const size_t rH = static_cast<size_t>(srows) * H;
  {
    StackBufs sb{s.h, s.h_mid, s.lse, s.qkv, s.ao, s.ff_pre, s.ff};
    RC(stack_forward_train(w->layers, L, H, F, w->heads, eps, sb, s.xn, s.cu, B, rg.on ? rg.max_frames : s.frames,
                           fmt, stream, reg, rg.on ? srows : -1));
  }
  const float* h_last = s.h + L * rH;
  if (rg.on) {  // packed -> padded (tail rows zero) for the per-utterance pooling windows
    RC(gather_rows_f32(s.h + L * rH, s.unpack_idx, s.hpad, rows, H, stream));
    h_last = s.hpad;
  }
  RC(layernorm_avgpool_fwd(h_last, w->final_ln_g, w->final_ln_b, eps, s.pooled_x, B, s.frames, H, w->pool_kernel,
                           w->pool_stride, s.pooled, fmt, stream));
  {
    GemmArgs g = lin(s.pooled_x, w->proj_w, static_cast<long long>(B) * s.pooled, w->llm_dim, H);
    g.epi = EPI_F32;
    g.bias = w->proj_b;
    g.out = audio_embeds;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  return B2S_OK;
}

int hubert_backward(const b2s_hubert_weights* w, const void* pos_w_dgrad, const b2s_hubert_grads* gr, const float* wave,
                    long long wave_stride, int batches, int samples, const int* samples_per_utt, void* saved,
                    size_t saved_bytes, const float* d_audio_embeds, void* workspace, size_t workspace_bytes,
                    const b2s_encoder_regularizers* reg, void* const* layer_done, cudaStream_t stream) {
  B2S_REQUIRE(w && pos_w_dgrad && gr && gr->layers && wave && saved && d_audio_embeds && workspace,
              "hubert_backward: null pointer");
  RC(check_regularizers(reg));
  Saved s;
  plan_saved(w, batches, samples, saved, saved_bytes, &s);
  B2S_REQUIRE(s.bytes <= saved_bytes && s.frames > 0 && s.pooled > 0, "hubert_backward: bad saved region");
  BwdWs b;
  plan_bwd(w, batches, s, workspace, workspace_bytes, &b);
  B2S_REQUIRE(b.bytes <= workspace_bytes, "hubert_backward: workspace too small: need %zu bytes, got %zu", b.bytes,
              workspace_bytes);
  const int B = batches, H = w->hidden, F = w->ffn, L = w->num_layers, Cl = w->llm_dim, fmt = w->fmt;
  const long long rows = static_cast<long long>(B) * s.frames;
  const float eps = w->ln_eps;
  Ragged rg;
  RC(plan_ragged(w, B, samples, samples_per_utt, s.frames, &rg));  // device index arrays are still in `saved`
  const long long srows = rg.on ? rg.rows_packed : rows;
  const size_t rH = static_cast<size_t>(srows) * H;

  // ---- projector + AvgPool + final LayerNorm, then the transformer layers last to first
  StackBufs sb{s.h, s.h_mid, s.lse, s.qkv, s.ao, s.ff_pre, s.ff};
  StackScratch sc{b.dh, b.delta, b.dyb, b.dbig, b.dsm, b.xn};
  if (rg.on) {
    // padded layout: h[L] scattered back by the forward lives in s.hpad; the head's gradient goes to the padded twins
    StackScratch sc_pad{b.dh_pad, b.delta, b.dyb_pad, b.dbig, b.dsm, b.xn};
    RC(head_backward(w->proj_w, w->final_ln_g, w->final_ln_b, gr->proj_w, gr->proj_b, gr->final_ln_g, gr->final_ln_b,
                     s.hpad, s.pooled_x, d_audio_embeds, b.da, b.dpool, b.dxn_f, sc_pad, B, s.frames, s.pooled, H, Cl,
                     w->pool_kernel, w->pool_stride, eps, fmt, stream));
    RC(gather_rows_f32(b.dh_pad, s.pack_idx, b.dh, srows, H, stream));
    RC(cast_f32_to_h16(b.dh, b.dyb, srows * H, fmt, stream));
  } else {
    RC(head_backward(w->proj_w, w->final_ln_g, w->final_ln_b, gr->proj_w, gr->proj_b, gr->final_ln_g, gr->final_ln_b,
                     s.h + L * rH, s.pooled_x, d_audio_embeds, b.da, b.dpool, b.dxn_f, sc, B, s.frames, s.pooled, H,
                     Cl, w->pool_kernel, w->pool_stride, eps, fmt, stream));
  }
  RC(stack_backward(w->layers, gr->layers, L, H, F, w->heads, eps, sb, sc, s.cu, B, rg.on ? rg.max_frames : s.frames,
                    fmt, stream, reg, rg.on ? srows : -1, layer_done));
  if (rg.on) {  // packed -> padded gradient of h0 (tail rows zero); the front end's backward runs on the padded layout
    RC(gather_rows_f32(b.dh, s.unpack_idx, b.dh_pad, rows, H, stream));
    RC(cast_f32_to_h16(b.dh_pad, b.dyb_pad, rows * H, fmt, stream));
    b.dh = b.dh_pad;
    b.dyb = b.dyb_pad;
  }

  // ---- positional conv embedding: h0 = drop(hp + gelu(conv(hp) + b)); dh / dyb = gradient w.r.t. h0
  if (reg) RC(dropout_apply(b.dh, b.dyb, rows * H, make_drop_spec(reg->seed, SITE_POS_ADD, reg->p_hidden), fmt, stream));
  RC(gelu_bwd(s.pos_pre, b.dyb, b.dsm, rows * H, fmt, stream));  // dsm = d(pre-GELU conv output), bf16
  RC(colsum_accum(b.dsm, 1, gr->pos_b, rows, H, fmt, stream));
  {
    // grouped wgrad: dW[g*64 + o][tap*64 + i] += sum_{b,t} dsm[b, t, g*64 + o] * hp[b, t + tap - pad, g*64 + i]
    GemmArgs g{};
    g.A = b.dsm;
    g.a_dim0 = H;
    g.a_row_stride = H;
    g.a_batch_stride = static_cast<long long>(s.frames) * H;
    g.a_rows = s.frames;
    g.a_mn = 1;
    g.a_group_off = 64;
    g.W = s.hp_bf;
    g.w_rows = s.frames;
    g.w_cols = H;
    g.w_row_stride = H;
    g.w_batch_stride = static_cast<long long>(s.frames) * H;
    g.w_group_off = 64;
    g.b_mn = 1;
    g.b_tap_atoms = 1;
    g.a_pad = w->pos_k / 2;
    g.M = 64;
    g.N = w->pos_k * 64;
    g.batches = 1;
    g.groups = w->pos_groups;
    g.taps = 1;
    g.k_per_tap = s.frames;
    g.k_batches = B;
    g.epi = EPI_ACCUM_F32;
    g.out = gr->pos_w;
    g.ldo = static_cast<long long>(w->pos_k) * 64;
    g.out_group_rows = 64;
    g.out_group_cols = 0;
    g.cta_group = 1;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  {
    // grouped dgrad: dh += conv^T(dsm) -- the forward's tap walk with flipped taps and transposed 64x64 blocks
    GemmArgs g{};
    g.A = b.dsm;
    g.a_dim0 = H;
    g.a_row_stride = H;
    g.a_batch_stride = static_cast<long long>(s.frames) * H;
    g.a_rows = s.frames;
    g.W = pos_w_dgrad;
    g.w_rows = H;
    g.w_cols = w->pos_k * 64;
    g.M = s.frames;
    g.N = 64;
    g.batches = B;
    g.groups = w->pos_groups;
    g.taps = w->pos_k;
    g.k_per_tap = 64;
    g.a_pad = w->pos_k - 1 - w->pos_k / 2;
    g.a_group_off = 64;
    g.w_group_off = 64;
    g.epi = EPI_RESID_F32;
    g.out = b.dh;
    g.resid = b.dh;
    g.ldo = H;
    g.out_batch_rows = s.frames;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  // ---- feature projection: hp = specaug(drop(fp_w LN(conv_x[6]) + fp_b)) ; dh = gradient w.r.t. hp
  if (rg.on)  // the forward zeroed the tail rows: they pass no gradient on (the conv transpose wrote into them)
    RC(mask_rows_f32(b.dh, s.tail, s.zeros_h, rows, H, stream));
  if (reg)
    RC(featproj_reg_bwd(b.dh, reg->time_mask, reg->g_masked_spec_embed, rows, H,
                        make_drop_spec(reg->seed, SITE_FEAT_PROJ, reg->p_feat_proj), stream));
  RC(colsum_accum(b.dh, 0, gr->fp_b, rows, H, fmt, stream));
  RC(cast_f32_to_h16(b.dh, b.dyb, rows * H, fmt, stream));
  RC(layernorm_fwd(s.conv_x[6], 1, w->fp_ln_g, w->fp_ln_b, eps, 0, b.xn, rows, 512, fmt, stream));
  RC(wgrad(b.dyb, b.xn, rows, H, 512, gr->fp_w, fmt, stream));
  RC(dgrad(b.dyb, w->fp_w, rows, H, 512, EPI_BF16, b.dsm, fmt, stream));
  void* dcur = b.dxa;  // gradient w.r.t. conv_x[6], bf16 [B, frames, 512]
  void* dnxt = b.dxb;
  RC(layernorm_bwd_ex(s.conv_x[6], 1, w->fp_ln_g, w->fp_ln_b, 0, eps, b.dsm, 1, nullptr, 0, dcur, gr->fp_ln_g,
                      gr->fp_ln_b, rows, 512, fmt, stream));
  // ---- conv feature extractor, layers 6..1 (implicit GEMM) then layer 0
  for (int i = 5; i >= 0; --i) {
    const int tin = s.t[i + 1], tout = s.t[i + 2];
    const int k = w->conv_k[i], sd = w->conv_stride[i];
    const long long orows = static_cast<long long>(B) * tout;
    RC(layernorm_bwd_ex(s.conv_pre[i], 0, w->conv_ln_g[i], w->conv_ln_b[i], 1, eps, dcur, 1, nullptr, 0, b.dpre,
                        gr->conv_ln_g[i], gr->conv_ln_b[i], orows, 512, fmt, stream));
    RC(colsum_accum(b.dpre, 1, gr->conv_b[i], orows, 512, fmt, stream));
    {
      GemmArgs g{};
      g.A = b.dpre;
      g.a_dim0 = 512;
      g.a_row_stride = 512;
      g.a_batch_stride = static_cast<long long>(tout) * 512;
      g.a_rows = tout;
      g.a_mn = 1;
      g.W = s.conv_x[i];
      g.w_rows = tout;
      g.w_cols = k * 512;
      g.w_row_stride = static_cast<long long>(sd) * 512;
      g.w_batch_stride = static_cast<long long>(tin) * 512;
      g.b_mn = 1;
      g.M = 512;
      g.N = k * 512;
      g.batches = 1;
      g.groups = 1;
      g.taps = 1;
      g.k_per_tap = tout;
      g.k_batches = B;
      g.epi = EPI_ACCUM_F32;
      g.out = gr->conv_w[i];
      g.ldo = static_cast<long long>(k) * 512;
      RC(gemm_launch_fmt(g, fmt, stream));
    }
    RC(dgrad(b.dpre, w->conv_w[i], orows, 512, k * 512, EPI_BF16, b.dcol, fmt, stream));
    RC(col2im_add(b.dcol, dnxt, B, tin, tout, k, sd, 512, fmt, stream));
    void* tmp = dcur;
    dcur = dnxt;
    dnxt = tmp;
  }
  RC(conv0_bwd(wave, wave_stride, B, samples, w->conv0_w, w->conv0_b, w->conv0_ln_g, w->conv0_ln_b, eps, dcur, s.t[1],
               gr->conv0_w, gr->conv0_b, gr->conv0_ln_g, gr->conv0_ln_b, fmt, stream));
  return B2S_OK;
}

// ================================================================================================ Whisper
// Same training path for AudioEncoder(base="whisper") (REF/config/llama3_whisper.yaml trains it the same way):
// conv1 (k3, pad 1) + GELU -> conv2 (k3, stride 2, pad 1) + GELU + frozen sinusoid table -> the shared pre-LN stack ->
// final LN + AvgPool + projector (TF/models/whisper/modeling_whisper.py:593-647). The positional table is frozen
// (requires_grad False in HF), k_proj has no bias (its slot of the fused-QKV bias gradient is discarded by the caller).
namespace {

struct WSaved {
  int frames_in, frames, pooled;
  void* x0;       // bf16 [B, T+2, mel]  zero-padded channels-last log-mel
  void* x1;       // bf16 [B, T+2, H]    zero-padded gelu(conv1)
  void* pre1;     // bf16 [B, T+2, H]    conv1 output before GELU (same geometry as x1)
  void* pre2;     // bf16 [B*frames, H]  conv2 output before GELU
  float *h, *h_mid, *lse;
  void *qkv, *ao, *ff_pre, *ff, *pooled_x, *xn;
  int* cu;
  size_t bytes;
};

void plan_wsaved(const b2s_whisper_weights* w, int batches, void* ws, size_t cap, WSaved* s) {
  s->frames_in = 2 * w->max_positions;
  s->frames = w->max_positions;
  s->pooled = s->frames >= w->pool_kernel ? (s->frames - w->pool_kernel) / w->pool_stride + 1 : 0;
  const size_t B = batches, H = w->hidden, F = w->ffn, L = w->num_layers, T = s->frames_in;
  const size_t rows = B * s->frames;
  Carve c(ws, cap);
  s->x0 = c.take(B * (T + 2) * w->mel_bins * 2 + 4096);
  s->x1 = c.take(B * (T + 2) * H * 2 + 4096);
  s->pre1 = c.take(B * (T + 2) * H * 2 + 4096);
  s->pre2 = c.take(rows * H * 2 + 4096);
  s->h = reinterpret_cast<float*>(c.take((L + 1) * rows * H * 4));
  s->h_mid = reinterpret_cast<float*>(c.take(L * rows * H * 4));
  s->qkv = c.take(L * rows * 3 * H * 2);
  s->ao = c.take(L * rows * H * 2);
  s->lse = reinterpret_cast<float*>(c.take(L * rows * w->heads * 4));
  s->ff_pre = c.take(L * rows * F * 2);
  s->ff = c.take(L * rows * F * 2);
  s->pooled_x = c.take(B * (s->pooled > 0 ? s->pooled : 1) * H * 2);
  s->xn = c.take(rows * H * 2 + 4096);
  s->cu = reinterpret_cast<int*>(c.take((B + 1) * sizeof(int)));
  s->bytes = c.off + 256;
}

struct WBwdWs {
  float *dh, *dxn_f, *dpool, *delta;
  void *dyb, *dbig, *dsm, *xn, *da, *dpre2, *dcol, *dx1;
  size_t bytes;
};

void plan_wbwd(const b2s_whisper_weights* w, int batches, const WSaved& s, void* ws, size_t cap, WBwdWs* p) {
  const size_t B = batches, H = w->hidden, F = w->ffn, T = s.frames_in;
  const size_t rows = B * s.frames;
  const size_t big = F > 3 * H ? F : 3 * H;
  const size_t np = B * (s.pooled > 0 ? s.pooled : 1);
  Carve c(ws, cap);
  p->dh = reinterpret_cast<float*>(c.take(rows * H * 4));
  p->dxn_f = reinterpret_cast<float*>(c.take(rows * H * 4));
  p->dpool = reinterpret_cast<float*>(c.take(np * H * 4));
  p->delta = reinterpret_cast<float*>(c.take(rows * w->heads * 4));
  p->dyb = c.take(rows * H * 2 + 4096);
  p->dbig = c.take(rows * big * 2 + 4096);
  p->dsm = c.take(rows * H * 2 + 4096);
  p->xn = c.take(rows * H * 2 + 4096);
  p->da = c.take(np * w->llm_dim * 2 + 4096);
  p->dpre2 = c.take(rows * H * 2 + 4096);
  p->dcol = c.take(rows * 3 * H * 2 + 4096);
  p->dx1 = c.take(B * (T + 2) * H * 2 + 4096);
  p->bytes = c.off + 256;
}

}  // namespace

size_t whisper_saved_bytes(const b2s_whisper_weights* w, int batches) {
  if (w == nullptr || batches <= 0) return 0;
  WSaved s;
  plan_wsaved(w, batches, nullptr, 0, &s);
  return s.bytes;
}

size_t whisper_backward_workspace_bytes(const b2s_whisper_weights* w, int batches) {
  if (w == nullptr || batches <= 0) return 0;
  WSaved s;
  plan_wsaved(w, batches, nullptr, 0, &s);
  WBwdWs b;
  plan_wbwd(w, batches, s, nullptr, 0, &b);
  return b.bytes;
}

int whisper_forward_train(const b2s_whisper_weights* w, const float* mel, int batches, int frames_in, void* saved,
                          size_t saved_bytes, float* audio_embeds, cudaStream_t stream) {
  B2S_REQUIRE(w && mel && saved && audio_embeds, "whisper_forward_train: null pointer");
  B2S_REQUIRE(batches > 0, "whisper_forward_train: empty batch");
  B2S_REQUIRE(frames_in == 2 * w->max_positions,
              "Whisper expects the mel input features to be of length %d, but found %d", 2 * w->max_positions,
              frames_in);
  B2S_REQUIRE(w->hidden % 256 == 0 && w->hidden / w->heads == 64 && w->mel_bins % 8 == 0,
              "whisper_forward_train: hidden %% 256, head_dim 64 and mel_bins %% 8 required");
  WSaved s;
  plan_wsaved(w, batches, saved, saved_bytes, &s);
  B2S_REQUIRE(s.bytes <= saved_bytes && s.pooled > 0, "whisper_forward_train: saved region too small (%zu < %zu)",
              saved_bytes, s.bytes);
  const int B = batches, H = w->hidden, F = w->ffn, C = w->mel_bins, T = frames_in, L = w->num_layers, fmt = w->fmt;
  const long long rows = static_cast<long long>(B) * s.frames;

  RC(mel_to_padded_cl(mel, s.x0, B, C, T, fmt, stream));
  B2S_CUDA_CHECK(cudaMemsetAsync(s.x1, 0, static_cast<size_t>(B) * (T + 2) * H * 2, stream));
  B2S_CUDA_CHECK(cudaMemsetAsync(s.pre1, 0, static_cast<size_t>(B) * (T + 2) * H * 2, stream));
  {
    GemmArgs g{};
    g.A = s.x0;
    g.a_dim0 = 3 * C;
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
    g.out = reinterpret_cast<__nv_bfloat16*>(s.x1) + H;
    g.ldo = H;
    g.out_batch_rows = T + 2;
    g.out2 = reinterpret_cast<__nv_bfloat16*>(s.pre1) + H;
    g.ld2 = H;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  {
    GemmArgs g{};
    g.A = s.x1;
    g.a_dim0 = 3 * H;
    g.a_row_stride = 2LL * H;
    g.a_batch_stride = static_cast<long long>(T + 2) * H;
    g.a_rows = s.frames;
    g.W = w->conv2_w;
    g.w_rows = H;
    g.w_cols = 3 * H;
    g.M = s.frames;
    g.N = H;
    g.batches = B;
    g.groups = 1;
    g.taps = 1;
    g.k_per_tap = 3 * H;
    g.epi = EPI_RESID_F32;
    g.act = ACT_GELU;
    g.bias = w->conv2_b;
    g.out = s.h;
    g.resid = w->pos_emb;
    g.resid_bcast = 1;
    g.ldo = H;
    g.out_batch_rows = s.frames;
    g.out2 = s.pre2;
    g.ld2 = H;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  iota_scaled<<<(B + 1 + 255) / 256, 256, 0, stream>>>(s.cu, B + 1, s.frames);
  B2S_LAUNCH_CHECK();
  StackBufs sb{s.h, s.h_mid, s.lse, s.qkv, s.ao, s.ff_pre, s.ff};
  RC(stack_forward_train(w->layers, L, H, F, w->heads, w->ln_eps, sb, s.xn, s.cu, B, s.frames, fmt, stream));
  RC(layernorm_avgpool_fwd(s.h + static_cast<size_t>(L) * rows * H, w->final_ln_g, w->final_ln_b, w->ln_eps, s.pooled_x,
                           B, s.frames, H, w->pool_kernel, w->pool_stride, s.pooled, fmt, stream));
  {
    GemmArgs g = lin(s.pooled_x, w->proj_w, static_cast<long long>(B) * s.pooled, w->llm_dim, H);
    g.epi = EPI_F32;
    g.bias = w->proj_b;
    g.out = audio_embeds;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  return B2S_OK;
}

int whisper_backward(const b2s_whisper_weights* w, const b2s_whisper_grads* gr, int batches, void* saved,
                     size_t saved_bytes, const float* d_audio_embeds, void* workspace, size_t workspace_bytes,
                     void* const* layer_done, cudaStream_t stream) {
  B2S_REQUIRE(w && gr && gr->layers && saved && d_audio_embeds && workspace, "whisper_backward: null pointer");
  WSaved s;
  plan_wsaved(w, batches, saved, saved_bytes, &s);
  B2S_REQUIRE(s.bytes <= saved_bytes && s.pooled > 0, "whisper_backward: bad saved region");
  WBwdWs b;
  plan_wbwd(w, batches, s, workspace, workspace_bytes, &b);
  B2S_REQUIRE(b.bytes <= workspace_bytes, "whisper_backward: workspace too small: need %zu bytes, got %zu", b.bytes,
              workspace_bytes);
  const int B = batches, H = w->hidden, F = w->ffn, C = w->mel_bins, T = s.frames_in, L = w->num_layers, fmt = w->fmt;
  const long long rows = static_cast<long long>(B) * s.frames;
  const size_t rH = static_cast<size_t>(rows) * H;
  StackBufs sb{s.h, s.h_mid, s.lse, s.qkv, s.ao, s.ff_pre, s.ff};
  StackScratch sc{b.dh, b.delta, b.dyb, b.dbig, b.dsm, b.xn};
  RC(head_backward(w->proj_w, w->final_ln_g, w->final_ln_b, gr->proj_w, gr->proj_b, gr->final_ln_g, gr->final_ln_b,
                   s.h + L * rH, s.pooled_x, d_audio_embeds, b.da, b.dpool, b.dxn_f, sc, B, s.frames, s.pooled, H,
                   w->llm_dim, w->pool_kernel, w->pool_stride, w->ln_eps, fmt, stream));
  RC(stack_backward(w->layers, gr->layers, L, H, F, w->heads, w->ln_eps, sb, sc, s.cu, B, s.frames, fmt, stream, nullptr,
                    -1, layer_done));
  // ---- conv2: h0 = gelu(conv2(x1) + b2) + pos (pos frozen); dyb = bf16 d(loss)/d(h0)
  RC(gelu_bwd(s.pre2, b.dyb, b.dpre2, rows * H, fmt, stream));
  RC(colsum_accum(b.dpre2, 1, gr->conv2_b, rows, H, fmt, stream));
  {
    GemmArgs g{};  // dW2[o, (j, c)] += sum_{b,t} dpre2[b,t,o] * x1p[b, 2t + j, c]
    g.A = b.dpre2;
    g.a_dim0 = H;
    g.a_row_stride = H;
    g.a_batch_stride = static_cast<long long>(s.frames) * H;
    g.a_rows = s.frames;
    g.a_mn = 1;
    g.W = s.x1;
    g.w_rows = s.frames;
    g.w_cols = 3 * H;
    g.w_row_stride = 2LL * H;
    g.w_batch_stride = static_cast<long long>(T + 2) * H;
    g.b_mn = 1;
    g.M = H;
    g.N = 3 * H;
    g.batches = 1;
    g.groups = 1;
    g.taps = 1;
    g.k_per_tap = s.frames;
    g.k_batches = B;
    g.epi = EPI_ACCUM_F32;
    g.out = gr->conv2_w;
    g.ldo = 3LL * H;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  RC(dgrad(b.dpre2, w->conv2_w, rows, H, 3 * H, EPI_BF16, b.dcol, fmt, stream));
  RC(col2im_add(b.dcol, b.dx1, B, T + 2, s.frames, 3, 2, H, fmt, stream));  // gradient on the PADDED x1 rows
  // ---- conv1: x1 = gelu(conv1(x0) + b1); the two padding rows of every utterance carry no gradient
  RC(gelu_bwd(s.pre1, b.dx1, b.dx1, static_cast<long long>(B) * (T + 2) * H, fmt, stream));
  B2S_CUDA_CHECK(cudaMemset2DAsync(b.dx1, static_cast<size_t>(T + 2) * H * 2, 0, static_cast<size_t>(H) * 2, B, stream));
  B2S_CUDA_CHECK(cudaMemset2DAsync(reinterpret_cast<__nv_bfloat16*>(b.dx1) + static_cast<size_t>(T + 1) * H,
                                   static_cast<size_t>(T + 2) * H * 2, 0, static_cast<size_t>(H) * 2, B, stream));
  RC(colsum_accum(b.dx1, 1, gr->conv1_b, static_cast<long long>(B) * (T + 2), H, fmt, stream));
  {
    GemmArgs g{};  // dW1[o, (j, c)] += sum_{b,t} dpre1[b,t,o] * x0p[b, t + j, c]
    g.A = reinterpret_cast<__nv_bfloat16*>(b.dx1) + H;
    g.a_dim0 = H;
    g.a_row_stride = H;
    g.a_batch_stride = static_cast<long long>(T + 2) * H;
    g.a_rows = T;
    g.a_mn = 1;
    g.W = s.x0;
    g.w_rows = T;
    g.w_cols = 3 * C;
    g.w_row_stride = C;
    g.w_batch_stride = static_cast<long long>(T + 2) * C;
    g.b_mn = 1;
    g.M = H;
    g.N = 3 * C;
    g.batches = 1;
    g.groups = 1;
    g.taps = 1;
    g.k_per_tap = T;
    g.k_batches = B;
    g.epi = EPI_ACCUM_F32;
    g.out = gr->conv1_w;
    g.ldo = 3LL * C;
    RC(gemm_launch_fmt(g, fmt, stream));
  }
  return B2S_OK;
}

}  // namespace b2s
