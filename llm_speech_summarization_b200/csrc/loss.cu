// loss.cu -- fused streaming cross-entropy + logit-distillation (soft cross-entropy) over the vocabulary.
//
// Reference semantics (per utterance u with R_u response rows, batch-1 in the reference):
//   ld_u  = mean_{r<R_u}   ( lse(s_r) - sum_v softmax(t_r)_v * s_{r,v} )        REF/utils.py:167-178
//   ntp_u = mean_{r<R_u-1} ( lse(s_r) - s_r[label_{r+1}] )                       REF/model/audio_llama.py:84-98
// Rows are the last R_u positions of the student (audio prompt) / teacher (text prompt) logits
// (REF/trainer.py:334,349-352). The student and teacher rows are read exactly once (4*N*V bytes), in
// one pass: per row an online softmax over s and t simultaneously keeps (m_s, Z_s, m_t, Z_t, B_t) with
// B_t = sum e^{t-m_t} s, so sum softmax(t) s = B_t / Z_t; probabilities are never materialised.
// Backward writes only ds (6*N*V bytes total): ds = ck (softmax(s) - softmax(t)) + cc (softmax(s) - onehot).
//
// Forward kernel: HBM-bound on paper, but 2 MUFU.EX2 per logit pair (16 lanes/clk/SM) keep the SFU pipe >55 % busy
// at full bandwidth, so memory latency must be hidden completely and nothing may serialise the warps. Three
// layouts were measured on B200 (profiles/r01_loss_kernel.md): register staging, whole-slice bulk-copy staging in
// shared memory, and a persistent producer/consumer ring; the shared-memory variants lost to plain registers
// because their CTA-wide mbarrier phases make the MUFU demand bursty. Kept: registers + 2-deep software pipeline.
#include "b2s_common.cuh"
#include "b2s_ptx.cuh"
#include "ops.cuh"

namespace b2s {

namespace {

constexpr int kLossThreads = 256;
constexpr int kChunkCols = 16384;     // backward: vocabulary columns per CTA
constexpr int kMaxFwdChunkCols = 65536;  // forward: largest slice per CTA (shrunk when there are few rows)
constexpr int kMinFwdChunkCols = 8192;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {  // single MUFU.EX2; ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Partial {
  float m_s, z_s, m_t, z_t, b_t, pad;
};

__device__ __forceinline__ void merge_sz(float& m, float& z, float m2, float z2) {
  const float mn = fmaxf(m, m2);
  z = z * exp2f((m - mn) * kLog2e) + z2 * exp2f((m2 - mn) * kLog2e);
  m = mn;
}
__device__ __forceinline__ void merge_tzb(float& m, float& z, float& b, float m2, float z2, float b2) {
  const float mn = fmaxf(m, m2);
  const float f1 = exp2f((m - mn) * kLog2e), f2 = exp2f((m2 - mn) * kLog2e);
  z = z * f1 + z2 * f2;
  b = b * f1 + b2 * f2;
  m = mn;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x);
  f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z);
  f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// grid = (chunks, rows); each CTA streams `chunk_cols` columns of one (student, teacher) row pair straight from
// HBM into registers with a 2-deep software pipeline (the loads of iteration i+1 are in flight while iteration i
// is being reduced), 16-byte L1-bypassing loads, 4 CTAs (32 warps) per SM; no block-wide barrier until the end.
__global__ void __launch_bounds__(kLossThreads, 4)
kd_ce_partial_kernel(const __nv_bfloat16* __restrict__ S, const __nv_bfloat16* __restrict__ T, long long lds,
                     long long ldt, int V, Partial* __restrict__ part, int chunks, int chunk_cols) {
  const int row = blockIdx.y;
  const int chunk = blockIdx.x;
  const int c0 = chunk * chunk_cols;
  const int c1 = min(V, c0 + chunk_cols);
  const int nvec = (c1 - c0) >> 3;  // V % 8 == 0 is required by the launcher
  const uint4* sv = reinterpret_cast<const uint4*>(S + static_cast<long long>(row) * lds + c0);
  const uint4* tv = reinterpret_cast<const uint4*>(T + static_cast<long long>(row) * ldt + c0);

  float m_s = -INFINITY, z_s = 0.f, m_t = -INFINITY, z_t = 0.f, b_t = 0.f;
  float zs2 = 0.f, zt2 = 0.f, bt2 = 0.f;  // second accumulator set: halves the dependent-add chains

  constexpr int kPer = 2;  // vector pairs per thread per pipeline stage
  uint4 cs[kPer], ct[kPer], ns[kPer], nt[kPer];
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    const int i = threadIdx.x + u * kLossThreads;
    if (i < nvec) {
      cs[u] = ld_stream_u4(sv + i);
      ct[u] = ld_stream_u4(tv + i);
    }
  }
  for (int v = threadIdx.x; v < nvec; v += kPer * kLossThreads) {
#pragma unroll
    for (int u = 0; u < kPer; ++u) {  // prefetch the next stage
      const int i = v + (kPer + u) * kLossThreads;
      if (i < nvec) {
        ns[u] = ld_stream_u4(sv + i);
        nt[u] = ld_stream_u4(tv + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      if (v + u * kLossThreads < nvec) {
        float fs[8], ft[8];
        unpack8(cs[u], fs);
        unpack8(ct[u], ft);
        float vs = fs[0], vt = ft[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) {
          vs = fmaxf(vs, fs[j]);
          vt = fmaxf(vt, ft[j]);
        }
        if (vs > m_s) {
          const float f = ex2((m_s - vs) * kLog2e);
          z_s *= f;
          zs2 *= f;
          m_s = vs;
        }
        if (vt > m_t) {
          const float f = ex2((m_t - vt) * kLog2e);
          z_t *= f;
          zt2 *= f;
          b_t *= f;
          bt2 *= f;
          m_t = vt;
        }
        const float ms2 = m_s * kLog2e, mt2 = m_t * kLog2e;
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          z_s += ex2(fmaf(fs[j], kLog2e, -ms2));
          zs2 += ex2(fmaf(fs[j + 1], kLog2e, -ms2));
          const float e0 = ex2(fmaf(ft[j], kLog2e, -mt2));
          const float e1 = ex2(fmaf(ft[j + 1], kLog2e, -mt2));
          z_t += e0;
          zt2 += e1;
          b_t = fmaf(e0, fs[j], b_t);
          bt2 = fmaf(e1, fs[j + 1], bt2);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      cs[u] = ns[u];
      ct[u] = nt[u];
    }
  }
  z_s += zs2;
  z_t += zt2;
  b_t += bt2;

  // warp then block merge
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m_s, o), z2 = __shfl_xor_sync(0xffffffffu, z_s, o);
    const float n2 = __shfl_xor_sync(0xffffffffu, m_t, o), y2 = __shfl_xor_sync(0xffffffffu, z_t, o);
    const float b2 = __shfl_xor_sync(0xffffffffu, b_t, o);
    if (m2 > -INFINITY) merge_sz(m_s, z_s, m2, z2);
    if (n2 > -INFINITY) merge_tzb(m_t, z_t, b_t, n2, y2, b2);
  }
  __shared__ float sh[kLossThreads / 32][5];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[warp][0] = m_s; sh[warp][1] = z_s; sh[warp][2] = m_t; sh[warp][3] = z_t; sh[warp][4] = b_t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kLossThreads / 32; ++w) {
      if (sh[w][0] > -INFINITY) merge_sz(m_s, z_s, sh[w][0], sh[w][1]);
      if (sh[w][2] > -INFINITY) merge_tzb(m_t, z_t, b_t, sh[w][2], sh[w][3], sh[w][4]);
    }
    Partial pr;
    pr.m_s = m_s; pr.z_s = z_s; pr.m_t = m_t; pr.z_t = z_t; pr.b_t = b_t; pr.pad = 0.f;
    part[static_cast<long long>(row) * chunks + chunk] = pr;
  }
}

// grid = utterances; merges chunk partials per row, then reduces rows -> per-utterance ld / ntp.
__global__ void __launch_bounds__(256)
kd_ce_finalize_kernel(const Partial* __restrict__ part, int chunks, const __nv_bfloat16* __restrict__ S,
                      long long lds, const int* __restrict__ labels, const int* __restrict__ row_offsets,
                      float scale_kd, float scale_ce, float* __restrict__ lse_s, float* __restrict__ lse_t,
                      float* __restrict__ coef_kd, float* __restrict__ coef_ce, float* __restrict__ loss_ld,
                      float* __restrict__ loss_ntp) {
  const int u = blockIdx.x;
  const int r0 = row_offsets[u], r1 = row_offsets[u + 1];
  float kd_sum = 0.f, ce_sum = 0.f, ce_cnt = 0.f;
  for (int row = r0 + threadIdx.x; row < r1; row += blockDim.x) {
    const Partial* pr = part + static_cast<long long>(row) * chunks;
    float m_s = pr[0].m_s, z_s = pr[0].z_s, m_t = pr[0].m_t, z_t = pr[0].z_t, b_t = pr[0].b_t;
    for (int c = 1; c < chunks; ++c) {
      merge_sz(m_s, z_s, pr[c].m_s, pr[c].z_s);
      merge_tzb(m_t, z_t, b_t, pr[c].m_t, pr[c].z_t, pr[c].b_t);
    }
    const float ls = m_s + logf(z_s);
    const float lt = m_t + logf(z_t);
    lse_s[row] = ls;
    lse_t[row] = lt;
    kd_sum += ls - b_t / z_t;
    const int lab = labels[row];
    if (lab >= 0) {
      ce_sum += ls - __bfloat162float(S[static_cast<long long>(row) * lds + lab]);
      ce_cnt += 1.f;
    }
  }
  __shared__ float sh[3][8];
  kd_sum = warp_sum(kd_sum);
  ce_sum = warp_sum(ce_sum);
  ce_cnt = warp_sum(ce_cnt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[0][warp] = kd_sum; sh[1][warp] = ce_sum; sh[2][warp] = ce_cnt;
  }
  __syncthreads();
  __shared__ float tot[3];
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) {
      a += sh[0][w]; b += sh[1][w]; c += sh[2][w];
    }
    tot[0] = a; tot[1] = b; tot[2] = c;
    const int R = r1 - r0;
    loss_ld[u] = R > 0 ? a / R : 0.f;
    loss_ntp[u] = c > 0.f ? b / c : 0.f;
  }
  __syncthreads();
  // per-row gradient coefficients for the backward pass (D3 in SURVEY.md appendix D)
  const int R = r1 - r0;
  const float ck = R > 0 ? scale_kd / R : 0.f;
  const float cc = tot[2] > 0.f ? scale_ce / tot[2] : 0.f;
  for (int row = r0 + threadIdx.x; row < r1; row += blockDim.x) {
    coef_kd[row] = ck;
    coef_ce[row] = labels[row] >= 0 ? cc : 0.f;
  }
}

// grid = (chunks, rows): ds = ck*(p_s - p_t) + cc*(p_s - onehot(label))
__global__ void __launch_bounds__(kLossThreads)
kd_ce_bwd_kernel(const __nv_bfloat16* __restrict__ S, const __nv_bfloat16* __restrict__ T, long long lds,
                 long long ldt, int V, const int* __restrict__ labels, const float* __restrict__ lse_s,
                 const float* __restrict__ lse_t, const float* __restrict__ coef_kd,
                 const float* __restrict__ coef_ce, const float* __restrict__ loss_scale,
                 __nv_bfloat16* __restrict__ dS, long long ldd, int f16) {
  const int row = blockIdx.y;
  const int c0 = blockIdx.x * kChunkCols;
  const int c1 = min(V, c0 + kChunkCols);
  // the gradient enters the backward pass here: the dynamic loss scale (GradScaler, REF/trainer.py:374) multiplies it once
  const float ls = loss_scale != nullptr ? *loss_scale : 1.0f;
  const float ck = coef_kd[row] * ls, cc = coef_ce[row] * ls;
  const float ls2 = lse_s[row] * kLog2e, lt2 = lse_t[row] * kLog2e;
  const int lab = labels[row];
  const uint4* sv = reinterpret_cast<const uint4*>(S + static_cast<long long>(row) * lds + c0);
  const uint4* tv = reinterpret_cast<const uint4*>(T + static_cast<long long>(row) * ldt + c0);
  uint4* dv = reinterpret_cast<uint4*>(dS + static_cast<long long>(row) * ldd + c0);
  const int nvec = (c1 - c0) >> 3;
  const float cs = ck + cc;
  constexpr int kUnroll = 4;
  for (int base = threadIdx.x; base < nvec; base += kLossThreads * kUnroll) {
    uint4 su[kUnroll], tu[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kLossThreads;
      if (i < nvec) {
        su[u] = ld_stream_u4(sv + i);
        tu[u] = ld_stream_u4(tv + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int i = base + u * kLossThreads;
      if (i < nvec) {
        float fs[8], ft[8], g[8];
        unpack8(su[u], fs);
        unpack8(tu[u], ft);
        const int col = c0 + i * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float ps = ex2(fmaf(fs[j], kLog2e, -ls2));
          const float pt = ex2(fmaf(ft[j], kLog2e, -lt2));
          g[j] = cs * ps - ck * pt;
          if (col + j == lab) g[j] -= cc;
        }
        st_stream_u4(dv + i, pack8_h16(g, f16));
      }
    }
  }
}

}  // namespace

namespace {
int fwd_chunk_cols(int rows, int V) {
  int cols = kMaxFwdChunkCols;
  const long long want = 4LL * num_sms();
  while (cols > kMinFwdChunkCols && static_cast<long long>(rows) * ((V + cols - 1) / cols) < want) cols /= 2;
  return cols;
}
}  // namespace

size_t kd_ce_workspace_bytes(int rows, int V) {
  const int chunks = (V + kMinFwdChunkCols - 1) / kMinFwdChunkCols;  // upper bound over every slice size
  return static_cast<size_t>(rows) * chunks * sizeof(Partial);
}

int kd_ce_loss_fwd(const void* S, const void* T, long long lds, long long ldt, int rows, int V, const int* labels,
                   const int* row_offsets, int utterances, float scale_kd, float scale_ce, void* workspace,
                   float* lse_s, float* lse_t, float* coef_kd, float* coef_ce, float* loss_ld, float* loss_ntp,
                   cudaStream_t stream) {
  B2S_REQUIRE(utterances > 0 && rows >= 0 && loss_ld && loss_ntp, "kd_ce_loss_fwd: bad sizes");
  if (rows == 0) {  // no response rows at all: every utterance's mean over an empty set is reported as 0
    B2S_CUDA_CHECK(cudaMemsetAsync(loss_ld, 0, sizeof(float) * utterances, stream));
    B2S_CUDA_CHECK(cudaMemsetAsync(loss_ntp, 0, sizeof(float) * utterances, stream));
    return B2S_OK;
  }
  B2S_REQUIRE(S && T && labels && row_offsets && workspace && lse_s && lse_t && coef_kd && coef_ce,
              "kd_ce_loss_fwd: null pointer");
  B2S_REQUIRE(V > 0 && V % 8 == 0 && lds % 8 == 0 && ldt % 8 == 0, "kd_ce_loss_fwd: V/ld must be multiples of 8");
  B2S_REQUIRE((reinterpret_cast<uintptr_t>(S) & 15) == 0 && (reinterpret_cast<uintptr_t>(T) & 15) == 0,
              "kd_ce_loss_fwd: logits must be 16-byte aligned");
  const int chunk_cols = fwd_chunk_cols(rows, V);
  const int chunks = (V + chunk_cols - 1) / chunk_cols;
  dim3 grid(chunks, rows);
  kd_ce_partial_kernel<<<grid, kLossThreads, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(S), reinterpret_cast<const __nv_bfloat16*>(T), lds, ldt, V,
      reinterpret_cast<Partial*>(workspace), chunks, chunk_cols);
  B2S_LAUNCH_CHECK();
  kd_ce_finalize_kernel<<<utterances, 256, 0, stream>>>(reinterpret_cast<const Partial*>(workspace), chunks,
                                                        reinterpret_cast<const __nv_bfloat16*>(S), lds, labels,
                                                        row_offsets, scale_kd, scale_ce, lse_s, lse_t, coef_kd,
                                                        coef_ce, loss_ld, loss_ntp);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int kd_ce_loss_bwd(const void* S, const void* T, long long lds, long long ldt, int rows, int V, const int* labels,
                   const float* lse_s, const float* lse_t, const float* coef_kd, const float* coef_ce,
                   const float* loss_scale, void* dS, long long ldd, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(S && T && labels && lse_s && lse_t && coef_kd && coef_ce && dS, "kd_ce_loss_bwd: null pointer");
  B2S_REQUIRE(V > 0 && V % 8 == 0 && lds % 8 == 0 && ldt % 8 == 0 && ldd % 8 == 0,
              "kd_ce_loss_bwd: V/ld must be multiples of 8");
  if (rows == 0) return B2S_OK;
  const int chunks = (V + kChunkCols - 1) / kChunkCols;
  dim3 grid(chunks, rows);
  kd_ce_bwd_kernel<<<grid, kLossThreads, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(S), reinterpret_cast<const __nv_bfloat16*>(T), lds, ldt, V, labels,
      lse_s, lse_t, coef_kd, coef_ce, loss_scale, reinterpret_cast<__nv_bfloat16*>(dS), ldd, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
