// attention_bwd.cu -- packed variable-length flash-attention BACKWARD (bf16 in, fp32 accumulate, bf16 out).
//
// Needed by the training step (REF/trainer.py:373-374 back-propagates through HubertAttention and LlamaAttention,
// TF/models/hubert/modeling_hubert.py:262-345, TF/models/llama/modeling_llama.py:225-289). Two kernels, FA2-style
// recomputation, no atomics:
//   attn_bwd_dq_kernel : one CTA per (sequence, query head, 64-query block), loops over key blocks:
//                        S = Q K^T, P = exp2(S*c - lse), dP = dO V^T, dS = P (dP - delta) * scale, dQ += dS K
//   attn_bwd_dkv_kernel: one CTA per (sequence, kv head, 64-key block), loops over the query heads of the GQA group
//                        and over query blocks, working on TRANSPOSED tiles so no register transposes are needed:
//                        S^T = K Q^T, P^T, dP^T = V dO^T, dS^T, dV += P^T dO, dK += dS^T Q
// `lse` is the forward's log2-domain log-sum-exp of the SCALED scores (attention_tc.cu), `delta` = rowsum(dO * O).
// Rotary embedding backward (the inverse rotation) is fused into the dQ / dK stores when `rope_cs` is given.
// Round-1 implementation on mma.sync.m16n8k16 (the backward is ~2.5x the forward's FLOPs but still < 4 % of the
// training step); a tcgen05 version follows the forward kernel's pattern.
#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kT = 64;  // tile edge (queries and keys)
constexpr int kBwdThreads = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return static_cast<uint32_t>(row * (D * 2) + ((chunk ^ (row & 7)) << 4));
}

// rows_valid rows of D bf16 (row stride ld) -> swizzled [64][D] tile, zero-filled beyond
template <int D>
__device__ __forceinline__ void load_tile(uint32_t smem, const __nv_bfloat16* g, long long ld, int rows_valid) {
  constexpr int CH = D / 8;
  for (int i = threadIdx.x; i < kT * CH; i += kBwdThreads) {
    const int r = i / CH, c = i - r * CH;
    const bool ok = r < rows_valid;
    cp_async16(smem + swz<D>(r, c), ok ? static_cast<const void*>(g + r * ld + c * 8) : static_cast<const void*>(g), ok);
  }
}

// A-operand fragments (16 rows x D) of a row-major tile for warp-row block `rb`
template <int D>
__device__ __forceinline__ void load_a_frags(uint32_t tile, int rb, int lane, uint32_t (&f)[D / 16][4]) {
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    const int row = rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int chunk = ks * 2 + (lane >> 4);
    ldsm4(tile + swz<D>(row, chunk), f[ks][0], f[ks][1], f[ks][2], f[ks][3]);
  }
}

// acc[16 x 64] += A(16 x D, fragments) * T^T where T is a row-major [64][D] tile (B[n][k] = T[n][k])
template <int D>
__device__ __forceinline__ void mm_a_tt(float (&acc)[kT / 8][4], const uint32_t (&a)[D / 16][4], uint32_t tile, int lane) {
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
#pragma unroll
    for (int nb = 0; nb < kT / 16; ++nb) {
      uint32_t b0, b1, b2, b3;
      const int row = nb * 16 + (lane & 7) + (lane >> 4) * 8;
      const int chunk = ks * 2 + ((lane >> 3) & 1);
      ldsm4(tile + swz<D>(row, chunk), b0, b1, b2, b3);
      mma16816(acc[2 * nb], a[ks], b0, b1);
      mma16816(acc[2 * nb + 1], a[ks], b2, b3);
    }
  }
}

// acc[16 x D] += P(16 x 64, bf16 A-fragments) * T where T is a row-major [64][D] tile (B[k][n] = T[k][n])
template <int D>
__device__ __forceinline__ void mm_p_t(float (&acc)[D / 8][4], const uint32_t (&p)[kT / 16][4], uint32_t tile, int lane) {
#pragma unroll
  for (int kk = 0; kk < kT / 16; ++kk) {
#pragma unroll
    for (int db = 0; db < D / 16; ++db) {
      uint32_t b0, b1, b2, b3;
      const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      const int chunk = db * 2 + (lane >> 4);
      ldsm4t(tile + swz<D>(row, chunk), b0, b1, b2, b3);
      mma16816(acc[2 * db], p[kk], b0, b1);
      mma16816(acc[2 * db + 1], p[kk], b2, b3);
    }
  }
}

struct BwdParams {
  const __nv_bfloat16 *q, *k, *v, *dout;
  long long ld_qkv, ld_do;
  const float* lse;    // [rows, Hq] log2-domain lse of the scaled scores
  const float* delta;  // [rows, Hq]
  __nv_bfloat16 *dq, *dk, *dv;
  long long ld_dqkv;
  const int* cu;
  int Hq, Hkv;
  float scale, scale_log2;
  int causal;
  const float* rope_cs;  // [npos, D]: cos[0:D/2] | sin[0:D/2], or null
};

// inverse rotary rotation on a C-fragment row pair: y = x*cos + rot(x)*sin  =>  dx = dy*cos - rot(dy)*sin,
// with rot(v) = (-v_hi, v_lo): dx_lo = dy_lo*cos + dy_hi*sin ; dx_hi = dy_hi*cos - dy_lo*sin
template <int D>
__device__ __forceinline__ void rope_bwd_frag(float (&acc)[D / 8][4], const float* rope_cs, int pos0, int pos1, int t) {
  constexpr int HB = D / 16;  // 8-column blocks per half
#pragma unroll
  for (int db = 0; db < HB; ++db) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int col = db * 8 + 2 * t + (e & 1);
      const int pos = (e >> 1) ? pos1 : pos0;
      const float c = __ldg(rope_cs + static_cast<long long>(pos) * D + col);
      const float s = __ldg(rope_cs + static_cast<long long>(pos) * D + D / 2 + col);
      const float lo = acc[db][e], hi = acc[db + HB][e];
      acc[db][e] = lo * c + hi * s;
      acc[db + HB][e] = hi * c - lo * s;
    }
  }
}

template <int D>
__device__ __forceinline__ void store_frag_rows(const float (&acc)[D / 8][4], __nv_bfloat16* base, long long ld, int row0,
                                                int rows_valid, int g, int t) {
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + g + r * 8;
    if (row < rows_valid) {
      __nv_bfloat16* o = base + static_cast<long long>(row) * ld;
#pragma unroll
      for (int db = 0; db < D / 8; ++db)
        *reinterpret_cast<uint32_t*>(o + db * 8 + 2 * t) = pack_bf16(acc[db][2 * r], acc[db][2 * r + 1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ dQ
template <int D>
__global__ void __launch_bounds__(kBwdThreads)
attn_bwd_dq_kernel(const BwdParams p) {
  constexpr int ROWB = D * 2;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const uint32_t sDO = sQ + kT * ROWB;
  const uint32_t sK = sDO + kT * ROWB;       // 2 stages
  const uint32_t sV = sK + 2 * kT * ROWB;    // 2 stages

  const int seq = blockIdx.z, h = blockIdx.y;
  const int s0 = p.cu[seq];
  const int L = p.cu[seq + 1] - s0;
  const int q0 = blockIdx.x * kT;
  if (q0 >= L) return;
  const int hk = h / (p.Hq / p.Hkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int qrows = min(kT, L - q0);

  const __nv_bfloat16* qg = p.q + static_cast<long long>(s0 + q0) * p.ld_qkv + h * D;
  const __nv_bfloat16* dog = p.dout + static_cast<long long>(s0 + q0) * p.ld_do + h * D;
  const __nv_bfloat16* kg = p.k + static_cast<long long>(s0) * p.ld_qkv + hk * D;
  const __nv_bfloat16* vg = p.v + static_cast<long long>(s0) * p.ld_qkv + hk * D;
  const int kv_len = p.causal ? min(L, q0 + kT) : L;
  const int nblk = (kv_len + kT - 1) / kT;

  load_tile<D>(sQ, qg, p.ld_qkv, qrows);
  load_tile<D>(sDO, dog, p.ld_do, qrows);
  load_tile<D>(sK, kg, p.ld_qkv, min(kT, L));
  load_tile<D>(sV, vg, p.ld_qkv, min(kT, L));
  cp_async_commit();

  const int r0 = q0 + warp * 16 + g;  // this thread's two query rows (sequence-local): r0, r0 + 8
  float lse_r[2], dl_r[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = r0 + r * 8;
    const bool ok = row < L;
    lse_r[r] = ok ? p.lse[static_cast<long long>(s0 + row) * p.Hq + h] : 0.f;
    dl_r[r] = ok ? p.delta[static_cast<long long>(s0 + row) * p.Hq + h] : 0.f;
  }

  uint32_t qf[D / 16][4], dof[D / 16][4];
  float dq[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dq[i][e] = 0.f;

  for (int j = 0; j < nblk; ++j) {
    const int st = j & 1;
    if (j + 1 < nblk) {
      const int kn = (j + 1) * kT;
      load_tile<D>(sK + (st ^ 1) * kT * ROWB, kg + static_cast<long long>(kn) * p.ld_qkv, p.ld_qkv, min(kT, L - kn));
      load_tile<D>(sV + (st ^ 1) * kT * ROWB, vg + static_cast<long long>(kn) * p.ld_qkv, p.ld_qkv, min(kT, L - kn));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      load_a_frags<D>(sQ, warp, lane, qf);
      load_a_frags<D>(sDO, warp, lane, dof);
    }
    float s[kT / 8][4], dp[kT / 8][4];
#pragma unroll
    for (int i = 0; i < kT / 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[i][e] = dp[i][e] = 0.f;
    mm_a_tt<D>(s, qf, sK + st * kT * ROWB, lane);
    mm_a_tt<D>(dp, dof, sV + st * kT * ROWB, lane);
    uint32_t dsf[kT / 16][4];
    const int kbase = j * kT;
#pragma unroll
    for (int nb = 0; nb < kT / 8; ++nb) {
      float ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + nb * 8 + 2 * t + (e & 1);
        const int qr = r0 + (e >> 1) * 8;
        const bool ok = key < L && (!p.causal || key <= qr);
        const float pv = ok ? ex2a(fmaf(s[nb][e], p.scale_log2, -lse_r[e >> 1])) : 0.f;
        ds[e] = pv * (dp[nb][e] - dl_r[e >> 1]) * p.scale;
      }
      dsf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
      dsf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
    mm_p_t<D>(dq, dsf, sK + st * kT * ROWB, lane);
    __syncthreads();
  }
  if (p.rope_cs != nullptr) rope_bwd_frag<D>(dq, p.rope_cs, r0, r0 + 8, t);
  store_frag_rows<D>(dq, p.dq + static_cast<long long>(s0 + q0) * p.ld_dqkv + h * D, p.ld_dqkv, warp * 16, qrows, g, t);
}

// ------------------------------------------------------------------------------------------------ dK, dV
template <int D>
__global__ void __launch_bounds__(kBwdThreads)
attn_bwd_dkv_kernel(const BwdParams p) {
  constexpr int ROWB = D * 2;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sK = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const uint32_t sV = sK + kT * ROWB;
  const uint32_t sQ = sV + kT * ROWB;        // 2 stages
  const uint32_t sDO = sQ + 2 * kT * ROWB;   // 2 stages
  float* s_lse = reinterpret_cast<float*>(smem_raw + 6 * kT * ROWB);  // [2][64]
  float* s_dl = s_lse + 2 * kT;                                       // [2][64]

  const int seq = blockIdx.z, hk = blockIdx.y;
  const int s0 = p.cu[seq];
  const int L = p.cu[seq + 1] - s0;
  const int k0 = blockIdx.x * kT;
  if (k0 >= L) return;
  const int G = p.Hq / p.Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int krows = min(kT, L - k0);
  const int nq = (L + kT - 1) / kT;
  const int i0 = p.causal ? (k0 / kT) : 0;  // first query block that can see this key block
  const int iters = G * (nq - i0);

  const __nv_bfloat16* kg = p.k + static_cast<long long>(s0 + k0) * p.ld_qkv + hk * D;
  const __nv_bfloat16* vg = p.v + static_cast<long long>(s0 + k0) * p.ld_qkv + hk * D;

  auto load_q = [&](int it, int st) {
    const int hq = hk * G + it / (nq - i0);
    const int qb = i0 + it % (nq - i0);
    const int q0 = qb * kT;
    const int rows = min(kT, L - q0);
    load_tile<D>(sQ + st * kT * ROWB, p.q + static_cast<long long>(s0 + q0) * p.ld_qkv + hq * D, p.ld_qkv, rows);
    load_tile<D>(sDO + st * kT * ROWB, p.dout + static_cast<long long>(s0 + q0) * p.ld_do + hq * D, p.ld_do, rows);
    if (threadIdx.x < kT) {
      const int row = q0 + threadIdx.x;
      const bool ok = row < L;
      s_lse[st * kT + threadIdx.x] = ok ? p.lse[static_cast<long long>(s0 + row) * p.Hq + hq] : 0.f;
      s_dl[st * kT + threadIdx.x] = ok ? p.delta[static_cast<long long>(s0 + row) * p.Hq + hq] : 0.f;
    }
  };

  load_tile<D>(sK, kg, p.ld_qkv, krows);
  load_tile<D>(sV, vg, p.ld_qkv, krows);
  load_q(0, 0);
  cp_async_commit();

  uint32_t kf[D / 16][4], vf[D / 16][4];
  float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[i][e] = dv[i][e] = 0.f;
  const int kr0 = k0 + warp * 16 + g;  // this thread's two key rows (sequence-local)

  for (int it = 0; it < iters; ++it) {
    const int st = it & 1;
    if (it + 1 < iters) {
      load_q(it + 1, st ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
      load_a_frags<D>(sK, warp, lane, kf);
      load_a_frags<D>(sV, warp, lane, vf);
    }
    const int qb = i0 + it % (nq - i0);
    const int q0 = qb * kT;
    // transposed tiles: rows = keys (16 per warp), columns = 64 queries
    float st_[kT / 8][4], dpt[kT / 8][4];
#pragma unroll
    for (int i = 0; i < kT / 8; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) st_[i][e] = dpt[i][e] = 0.f;
    mm_a_tt<D>(st_, kf, sQ + st * kT * ROWB, lane);
    mm_a_tt<D>(dpt, vf, sDO + st * kT * ROWB, lane);
    uint32_t ptf[kT / 16][4], dstf[kT / 16][4];
#pragma unroll
    for (int nb = 0; nb < kT / 8; ++nb) {
      float pv[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qc = nb * 8 + 2 * t + (e & 1);  // query column inside the tile
        const int qr = q0 + qc;
        const int key = kr0 + (e >> 1) * 8;
        const bool ok = key < L && qr < L && (!p.causal || key <= qr);
        pv[e] = ok ? ex2a(fmaf(st_[nb][e], p.scale_log2, -s_lse[st * kT + qc])) : 0.f;
        ds[e] = pv[e] * (dpt[nb][e] - s_dl[st * kT + qc]) * p.scale;
      }
      ptf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
      ptf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
      dstf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(ds[0], ds[1]);
      dstf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
    }
    mm_p_t<D>(dv, ptf, sDO + st * kT * ROWB, lane);  // dV += P^T dO
    mm_p_t<D>(dk, dstf, sQ + st * kT * ROWB, lane);  // dK += dS^T Q
    __syncthreads();
  }
  if (p.rope_cs != nullptr) rope_bwd_frag<D>(dk, p.rope_cs, kr0, kr0 + 8, t);
  store_frag_rows<D>(dk, p.dk + static_cast<long long>(s0 + k0) * p.ld_dqkv + hk * D, p.ld_dqkv, warp * 16, krows, g, t);
  store_frag_rows<D>(dv, p.dv + static_cast<long long>(s0 + k0) * p.ld_dqkv + hk * D, p.ld_dqkv, warp * 16, krows, g, t);
}

// delta[row, h] = sum_d dO[row, h, d] * O[row, h, d]; one warp per row, 16-byte loads, lanes of one head reduce together
template <int D>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, long long ld_o,
                  long long ld_do, float* __restrict__ delta, long long rows, int Hq) {
  constexpr int kLanesPerHead = D / 8;            // 8 (D = 64) or 16 (D = 128)
  constexpr int kHeadsPerIter = 32 / kLanesPerHead;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  for (int h0 = 0; h0 < Hq; h0 += kHeadsPerIter) {
    const int h = h0 + lane / kLanesPerHead;
    float s = 0.f;
    if (h < Hq) {
      float a[8], b[8];
      ld8bf(o + row * ld_o + h0 * D + lane * 8, a);
      ld8bf(dout + row * ld_do + h0 * D + lane * 8, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
    }
#pragma unroll
    for (int off = kLanesPerHead / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (h < Hq && lane % kLanesPerHead == 0) delta[row * Hq + h] = s;
  }
}

template <int D>
int launch_bwd(const BwdParams& p, int num_seqs, int max_seqlen, cudaStream_t stream) {
  constexpr int smem_dq = 6 * kT * D * 2;
  constexpr int smem_dkv = 6 * kT * D * 2 + 4 * kT * 4;
  static bool attr_set = false;
  if (!attr_set) {
    B2S_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq));
    B2S_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv));
    attr_set = true;
  }
  const int nb = (max_seqlen + kT - 1) / kT;
  attn_bwd_dq_kernel<D><<<dim3(nb, p.Hq, num_seqs), kBwdThreads, smem_dq, stream>>>(p);
  B2S_LAUNCH_CHECK();
  attn_bwd_dkv_kernel<D><<<dim3(nb, p.Hkv, num_seqs), kBwdThreads, smem_dkv, stream>>>(p);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace

int attention_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* o, long long ld_o,
                  const void* dout, long long ld_do, const float* lse, float* delta_ws, void* dq, void* dk, void* dv,
                  long long ld_dqkv, const int* cu_seqlens, int num_seqs, int max_seqlen, long long total_rows, int Hq,
                  int Hkv, int D, float scale, int causal, const float* rope_cs, cudaStream_t stream,
                  const AttnDrop* drop) {
  B2S_REQUIRE(q && k && v && o && dout && lse && delta_ws && dq && dk && dv && cu_seqlens, "attention_bwd: null pointer");
  if (drop != nullptr && drop->thresh == 0u) drop = nullptr;
  B2S_REQUIRE(num_seqs > 0 && max_seqlen > 0 && total_rows > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0,
              "attention_bwd: bad sizes");
  B2S_REQUIRE(ld_qkv % 8 == 0 && ld_do % 8 == 0 && ld_o % 2 == 0 && ld_dqkv % 2 == 0,
              "attention_bwd: strides must keep 16-byte row alignment");
  B2S_REQUIRE(D == 64 || D == 128, "attention_bwd: head_dim %d unsupported (64 or 128)", D);
  B2S_REQUIRE(ld_o % 8 == 0, "attention_bwd: the forward output needs 16-byte aligned rows");
  {
    const unsigned grid = static_cast<unsigned>((total_rows + 7) / 8);
    const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(o);
    const __nv_bfloat16* dob = reinterpret_cast<const __nv_bfloat16*>(dout);
    if (D == 64) attn_delta_kernel<64><<<grid, 256, 0, stream>>>(ob, dob, ld_o, ld_do, delta_ws, total_rows, Hq);
    else attn_delta_kernel<128><<<grid, 256, 0, stream>>>(ob, dob, ld_o, ld_do, delta_ws, total_rows, Hq);
    B2S_LAUNCH_CHECK();
  }
  if (attention_get_impl() == 1) {  // tcgen05 kernels (attention_bwd_tc.cu); 0 keeps the mma.sync kernels below
    return attention_bwd_tc(q, k, v, ld_qkv, dout, ld_do, lse, delta_ws, dq, dk, dv, ld_dqkv, cu_seqlens, num_seqs,
                            max_seqlen, total_rows, Hq, Hkv, D, scale, causal, rope_cs, stream, drop);
  }
  B2S_REQUIRE(drop == nullptr, "attention_bwd: attention dropout needs the tcgen05 kernels (impl 1)");
  BwdParams p{};
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.ld_qkv = ld_qkv;
  p.ld_do = ld_do;
  p.lse = lse;
  p.delta = delta_ws;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.ld_dqkv = ld_dqkv;
  p.cu = cu_seqlens;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.rope_cs = rope_cs;
  if (D == 64) return launch_bwd<64>(p, num_seqs, max_seqlen, stream);
  if (D == 128) return launch_bwd<128>(p, num_seqs, max_seqlen, stream);
  set_last_error("attention_bwd: head_dim %d unsupported (64 or 128)", D);
  return B2S_ERR_UNSUPPORTED;
}

}  // namespace b2s
