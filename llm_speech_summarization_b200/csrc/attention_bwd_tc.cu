// attention_bwd_tc.cu -- flash-attention backward on tcgen05 / TMEM / TMA (sm_100a), the training counterpart of
// attention_tc.cu: a delta kernel (rowsum(dO o O)) + a dQ kernel over query tiles + a dK/dV kernel over key tiles, scores
// recomputed from the saved log-sum-exp; every matmul is a tcgen05.mma with its accumulator in TMEM and its operands in
// 128B-swizzled shared memory. All 16-bit tensors share ONE format (template F16): tcgen05 kind::f16 rejects mixed
// bf16 / fp16 operands (measured: illegal instruction), so fp16 activations mean fp16 gradients (loss-scaled by the
// caller, as the reference's GradScaler does).
//
//   dQ kernel  (CTA = sequence x query head x 128 queries; keys in blocks of 64)
//       S  = Q K^T, dP = dO V^T            A = Q / dO (K-major), B = K / V (K-major)      TMEM [0,64) / [64,128)
//       dS = P o (dP - delta) * scale      128 threads, thread = query row, straight from TMEM -> bf16 smem tile
//       dQ += dS K                         A = dS (K-major), B = K read MN-major            TMEM [128,128+D)
//   dK/dV kernel (CTA = sequence x kv head x 128 keys; queries in blocks of 64, all query heads of the GQA group)
//       S^T = K Q^T, dP^T = V dO^T         A = K / V (K-major), B = Q / dO (K-major)      TMEM [0,64) / [64,128)
//       P^T, dS^T                          thread = key row                               -> two bf16 smem tiles
//       dV += P^T dO, dK += dS^T Q         A = P^T / dS^T, B = dO / Q read MN-major         TMEM [128,..) / [128+D,..)
//
// One thread issues TMA + MMA (warp 4), hand-offs are mbarriers with bounded waits, exactly like the forward kernel.
// The inverse rotary rotation (RoPE^T) is fused into the dQ / dK stores (LLM only).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "b2s_common.cuh"
#include "b2s_ptx.cuh"
#include "ops.cuh"

namespace b2s {

int encode_map_2d_bf16(CUtensorMap* map, const void* base, unsigned long long cols, unsigned long long rows,
                       unsigned long long row_stride_bytes, unsigned box_cols, unsigned box_rows);  // gemm_sm100.cu

namespace {

constexpr int kBig = 128;   // rows of the resident tile (queries in the dQ kernel, keys in the dK/dV kernel)
constexpr int kSmall = 64;  // rows of the streamed tile
constexpr int kThreadsB = 160;

struct BwdTcParams {
  const int* cu;
  const float* lse;    // [rows, Hq] log2-domain
  const float* delta;  // [rows, Hq]
  __nv_bfloat16 *dq, *dk, *dv;
  long long ld_d;
  int Hq, Hkv;
  float scale, scale_log2;
  int causal;
  const float* rope_cs;  // optional [npos, D]: cos[0:D/2] | sin[0:D/2]
  AttnDrop drop;         // attention-probability dropout of the forward (thresh 0 = off), regenerated here
};

template <bool F16>
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  return F16 ? pack_f16(lo, hi) : pack_bf16(lo, hi);
}

__device__ __forceinline__ float ex2b(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// write 32 consecutive bf16 columns [32c, 32c+32) of row r into a [rows x 64] K-major 128B-swizzled tile
__device__ __forceinline__ void st_row_chunk(uint32_t tile, int r, int c, const uint32_t (&pk)[16]) {
  const uint32_t row = tile + r * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int chunk = c * 4 + q;
    const uint32_t addr = row + ((chunk ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                 "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                 : "memory");
  }
}

// TMEM accumulator row (D fp32) -> optional RoPE^T -> 16-bit (bf16 / fp16) -> global row
template <int D, bool F16>
__device__ __forceinline__ void store_acc_row(uint32_t taddr, __nv_bfloat16* dst, bool valid, float mul,
                                              const float* rope_row /* [D] cos|sin or null */) {
  if (rope_row != nullptr && D == 128) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {  // column chunks c (low half) and c + 2 (high half) rotate together
      uint32_t lo[32], hi[32];
      ptx::tmem_ld_32x32(taddr + c * 32, lo);
      ptx::tmem_ld_32x32(taddr + (c + 2) * 32, hi);
      ptx::tmem_ld_wait();
      if (valid) {
        float a[32], b[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // 16-byte loads of the row's cos / sin (the table row is 512 B, 16-byte aligned)
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(rope_row + c * 32) + q);
          const float4 s4 = __ldg(reinterpret_cast<const float4*>(rope_row + 64 + c * 32) + q);
          const float cs[4] = {c4.x, c4.y, c4.z, c4.w}, sn[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = q * 4 + e;
            const float l = __uint_as_float(lo[i]) * mul, h = __uint_as_float(hi[i]) * mul;
            a[i] = l * cs[e] + h * sn[e];
            b[i] = h * cs[e] - l * sn[e];
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          st8h(dst + c * 32 + q * 8, *reinterpret_cast<const float(*)[8]>(&a[q * 8]), F16);
          st8h(dst + 64 + c * 32 + q * 8, *reinterpret_cast<const float(*)[8]>(&b[q * 8]), F16);
        }
      }
    }
    return;
  }
#pragma unroll 1
  for (int c = 0; c < D / 32; ++c) {
    uint32_t raw[32];
    ptx::tmem_ld_32x32(taddr + c * 32, raw);
    ptx::tmem_ld_wait();
    if (valid) {
      float a[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = __uint_as_float(raw[i]) * mul;
#pragma unroll
      for (int q = 0; q < 4; ++q) st8h(dst + c * 32 + q * 8, *reinterpret_cast<const float(*)[8]>(&a[q * 8]), F16);
    }
  }
}

template <int D>
struct DqCfg {
  static constexpr int kStages = (D == 64) ? 2 : 1;
  static constexpr int kBigBytes = kBig * D * 2;
  static constexpr int kSmallBytes = kSmall * D * 2;
  static constexpr int kDsBytes = kBig * kSmall * 2;
  // no alignment slack (the window is declared 1024-byte aligned): at D = 128 this is what lets two CTAs share an SM
  static constexpr int kSmemBytes = 2 * kBigBytes + kStages * 2 * kSmallBytes + kDsBytes + 256;
  static constexpr int kTmemCols = 256;  // S [0,64) dP [64,128) dQ [128,128+D)
};

// ------------------------------------------------------------------------------------------------ dQ
// With dropout O = (P o M) V, M = keep / (1 - p): dP = M o (dO V^T), dV = (P o M)^T dO, and delta = rowsum(dO o O) is
// unchanged (sum_j P_j M_j (dO . V_j) = dO . O), so only the two elementwise stages below see the mask.
template <int D, bool DROP, bool F16>
__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                      const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                      const BwdTcParams p) {
  using C = DqCfg<D>;
  constexpr int kAtoms = D / 64;
  constexpr int kStages = C::kStages;
  const int seq = blockIdx.z, h = blockIdx.y;
  const int s0 = p.cu[seq];
  const int L = p.cu[seq + 1] - s0;
  const int q0 = blockIdx.x * kBig;
  if (q0 >= L) return;
  const int hk = h / (p.Hq / p.Hkv);
  const int kv_len = p.causal ? min(L, q0 + kBig) : L;
  const int nblk = (kv_len + kSmall - 1) / kSmall;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();  // the swizzled tiles need 1024-byte alignment
  const uint32_t sQ = base;
  const uint32_t sDO = sQ + C::kBigBytes;
  const uint32_t sK = sDO + C::kBigBytes;
  const uint32_t sV = sK + kStages * C::kSmallBytes;
  const uint32_t sDS = sV + kStages * C::kSmallBytes;
  const uint32_t bars = sDS + C::kDsBytes;
  const uint32_t bar_q = bars, bar_kv0 = bars + 8, bar_s = bars + 24, bar_p = bars + 32, bar_o = bars + 40;
  const uint32_t tmem_slot = bars + 48;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_kv0, 1);
    ptx::mbar_init(bar_kv0 + 8, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, kBig);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) ptx::tmem_alloc<1>(tmem_slot, C::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_do);
      ptx::prefetch_tmap(&tmap_k);
      ptx::prefetch_tmap(&tmap_v);
      ptx::mbar_arrive_expect_tx(bar_q, 2 * C::kBigBytes);
#pragma unroll
      for (int a = 0; a < kAtoms; ++a) {
        ptx::tma_load_2d(&tmap_q, bar_q, sQ + a * (kBig * 128), h * D + a * 64, s0 + q0);
        ptx::tma_load_2d(&tmap_do, bar_q, sDO + a * (kBig * 128), h * D + a * 64, s0 + q0);
      }
      auto load_kv = [&](int j) {
        const int st = j % kStages;
        ptx::mbar_arrive_expect_tx(bar_kv0 + 8 * st, 2 * C::kSmallBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a) {
          ptx::tma_load_2d(&tmap_k, bar_kv0 + 8 * st, sK + st * C::kSmallBytes + a * (kSmall * 128), hk * D + a * 64,
                           s0 + j * kSmall);
          ptx::tma_load_2d(&tmap_v, bar_kv0 + 8 * st, sV + st * C::kSmallBytes + a * (kSmall * 128), hk * D + a * 64,
                           s0 + j * kSmall);
        }
      };
      load_kv(0);
      constexpr uint32_t idesc_s = ptx::make_idesc_f32acc(kBig, kSmall) | ptx::idesc_formats(F16, F16);
      constexpr uint32_t idesc_o =
          ptx::make_idesc_f32acc(kBig, D) | ptx::idesc_formats(F16, F16) | (1u << 16);  // B (= K) read MN-major
      ptx::mbar_wait(bar_q, 0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kStages;
        if (kStages == 2 && j + 1 < nblk) {
          if (j >= 1) ptx::mbar_wait(bar_o, (j - 1) & 1);  // dQ(j-1) was the last reader of stage st^1
          load_kv(j + 1);
        }
        ptx::mbar_wait(bar_kv0 + 8 * st, (j / kStages) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t big_off = (k >> 2) * (kBig * 128) + (k & 3) * 32;
          const uint32_t small_off = (k >> 2) * (kSmall * 128) + (k & 3) * 32;
          ptx::umma_bf16<1>(tmem_base, ptx::make_kmajor_sw128_desc(sQ + big_off),
                            ptx::make_kmajor_sw128_desc(sK + st * C::kSmallBytes + small_off), idesc_s, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t big_off = (k >> 2) * (kBig * 128) + (k & 3) * 32;
          const uint32_t small_off = (k >> 2) * (kSmall * 128) + (k & 3) * 32;
          ptx::umma_bf16<1>(tmem_base + kSmall, ptx::make_kmajor_sw128_desc(sDO + big_off),
                            ptx::make_kmajor_sw128_desc(sV + st * C::kSmallBytes + small_off), idesc_s, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit<1>(bar_s);
        ptx::mbar_wait(bar_p, j & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < kSmall / 16; ++k) {
          const uint64_t ds_desc = ptx::make_kmajor_sw128_desc(sDS + k * 32);
          const uint64_t k_desc = ptx::make_mnmajor_sw128_desc(sK + st * C::kSmallBytes + k * 2048, kSmall * 128);
          ptx::umma_bf16<1>(tmem_base + 2 * kSmall, ds_desc, k_desc, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit<1>(bar_o);
        if (kStages == 1 && j + 1 < nblk) {
          ptx::mbar_wait(bar_o, j & 1);
          load_kv(j + 1);
        }
      }
    }
  } else {
    const int r = warp * 32 + lane;
    const int qi = q0 + r;
    const bool valid = qi < L;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_base;
    const uint32_t tDP = tS + kSmall;
    const uint32_t tDQ = tS + 2 * kSmall;
    const long long grow = static_cast<long long>(s0 + qi);
    const float lse = valid ? p.lse[grow * p.Hq + h] : 0.f;
    const float dl = valid ? p.delta[grow * p.Hq + h] : 0.f;
    const int row_limit = valid ? (p.causal ? min(L, qi + 1) : L) : 0;
    const float sc2 = p.scale_log2, sc = p.scale, dls = dl * p.scale;
    uint32_t dk1 = 0u, dk2 = 0u;
    if (DROP) rng_stream_key(p.drop.seed, p.drop.site, static_cast<uint32_t>(seq), static_cast<uint32_t>(h), &dk1, &dk2);
    const uint32_t drow = static_cast<uint32_t>(qi) << 16;
    for (int j = 0; j < nblk; ++j) {
      ptx::mbar_wait(bar_s, j & 1);
      ptx::tc_fence_after();
      const int nvis = row_limit - j * kSmall;
      const bool full_blk = __all_sync(0xffffffffu, nvis >= kSmall);
#pragma unroll 1
      for (int c = 0; c < kSmall / 32; ++c) {
        uint32_t rs[32], rd[32];
        ptx::tmem_ld_32x32(tS + c * 32, rs);
        ptx::tmem_ld_32x32(tDP + c * 32, rd);
        ptx::tmem_ld_wait();
        const int nv = nvis - c * 32;
        uint32_t pk[16];
        if (DROP) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t e = drow | static_cast<uint32_t>(j * kSmall + c * 32 + i);
            rd[i] = rng_keep(e, dk1, dk2, p.drop.thresh) ? __float_as_uint(__uint_as_float(rd[i]) * p.drop.inv_keep) : 0u;
          }
        }
        // one code path: a partial block first turns its invisible scores into -inf (P = 0, dS = 0) behind a
        // warp-uniform branch (two loops get merged by the compiler into the masked form for EVERY block, see
        // profiles/r02_attention_pipelined.md)
        if (!full_blk) {
#pragma unroll
          for (int i = 0; i < 32; ++i) rs[i] = (i < nv) ? rs[i] : 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2b(fmaf(__uint_as_float(rs[i]), sc2, -lse));
          const float p1 = ex2b(fmaf(__uint_as_float(rs[i + 1]), sc2, -lse));
          pk[i >> 1] = pack_op<F16>(p0 * fmaf(__uint_as_float(rd[i]), sc, -dls),
                                 p1 * fmaf(__uint_as_float(rd[i + 1]), sc, -dls));
        }
        st_row_chunk(sDS, r, c, pk);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_p);
    }
    ptx::mbar_wait(bar_o, (nblk - 1) & 1);
    ptx::tc_fence_after();
    const float* rope_row = p.rope_cs != nullptr ? p.rope_cs + static_cast<long long>(valid ? qi : 0) * D : nullptr;
    store_acc_row<D, F16>(tDQ, p.dq + grow * p.ld_d + h * D, valid, 1.0f, rope_row);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ dK, dV
template <int D>
struct DkvCfg {
  static constexpr int kStages = 2;
  static constexpr int kBigBytes = kBig * D * 2;
  static constexpr int kSmallBytes = kSmall * D * 2;
  static constexpr int kPtBytes = kBig * kSmall * 2;
  static constexpr int kSmemBytes = 2 * kBigBytes + kStages * 2 * kSmallBytes + 2 * kPtBytes + 2 * 2 * kSmall * 4 + 256;
  static constexpr int kTmemCols = (D == 64) ? 256 : 512;  // S^T [0,64) dP^T [64,128) dV [128,128+D) dK [128+D,128+2D)
};

template <int D, bool DROP, bool F16>
__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                       const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                       const BwdTcParams p) {
  using C = DkvCfg<D>;
  constexpr int kAtoms = D / 64;
  constexpr int kStages = C::kStages;
  const int seq = blockIdx.z, hk = blockIdx.y;
  const int s0 = p.cu[seq];
  const int L = p.cu[seq + 1] - s0;
  const int k0 = blockIdx.x * kBig;
  if (k0 >= L) return;
  const int G = p.Hq / p.Hkv;
  const int nq = (L + kSmall - 1) / kSmall;
  const int i0 = p.causal ? (k0 / kSmall) : 0;
  const int per_head = nq - i0;
  const int iters = G * per_head;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();  // the swizzled tiles need 1024-byte alignment
  const uint32_t sK = base;
  const uint32_t sV = sK + C::kBigBytes;
  const uint32_t sQ = sV + C::kBigBytes;
  const uint32_t sDO = sQ + kStages * C::kSmallBytes;
  const uint32_t sPT = sDO + kStages * C::kSmallBytes;
  const uint32_t sDST = sPT + C::kPtBytes;
  const uint32_t sStat = sDST + C::kPtBytes;  // float [2 stages][lse 64 | delta 64]
  const uint32_t bars = sStat + 2 * 2 * kSmall * 4;
  const uint32_t bar_kv = bars, bar_q0 = bars + 8, bar_s = bars + 24, bar_p = bars + 32, bar_o = bars + 40;
  const uint32_t tmem_slot = bars + 48;
  float* stat = reinterpret_cast<float*>(smem_raw + (sStat - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_kv, 1);
    ptx::mbar_init(bar_q0, 1);
    ptx::mbar_init(bar_q0 + 8, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, kBig);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) ptx::tmem_alloc<1>(tmem_slot, C::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_do);
      ptx::prefetch_tmap(&tmap_k);
      ptx::prefetch_tmap(&tmap_v);
      ptx::mbar_arrive_expect_tx(bar_kv, 2 * C::kBigBytes);
#pragma unroll
      for (int a = 0; a < kAtoms; ++a) {
        ptx::tma_load_2d(&tmap_k, bar_kv, sK + a * (kBig * 128), hk * D + a * 64, s0 + k0);
        ptx::tma_load_2d(&tmap_v, bar_kv, sV + a * (kBig * 128), hk * D + a * 64, s0 + k0);
      }
      auto load_q = [&](int it) {
        const int st = it % kStages;
        const int hq = hk * G + it / per_head;
        const int qrow = s0 + (i0 + it % per_head) * kSmall;
        ptx::mbar_arrive_expect_tx(bar_q0 + 8 * st, 2 * C::kSmallBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a) {
          ptx::tma_load_2d(&tmap_q, bar_q0 + 8 * st, sQ + st * C::kSmallBytes + a * (kSmall * 128), hq * D + a * 64, qrow);
          ptx::tma_load_2d(&tmap_do, bar_q0 + 8 * st, sDO + st * C::kSmallBytes + a * (kSmall * 128), hq * D + a * 64,
                           qrow);
        }
      };
      load_q(0);
      constexpr uint32_t idesc_s = ptx::make_idesc_f32acc(kBig, kSmall) | ptx::idesc_formats(F16, F16);
      constexpr uint32_t idesc_o = ptx::make_idesc_f32acc(kBig, D) | ptx::idesc_formats(F16, F16) | (1u << 16);
      ptx::mbar_wait(bar_kv, 0);
      for (int it = 0; it < iters; ++it) {
        const int st = it % kStages;
        if (it + 1 < iters) {
          if (it >= 1) ptx::mbar_wait(bar_o, (it - 1) & 1);  // dV/dK(it-1) were the last readers of stage st^1
          load_q(it + 1);
        }
        ptx::mbar_wait(bar_q0 + 8 * st, (it / kStages) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t big_off = (k >> 2) * (kBig * 128) + (k & 3) * 32;
          const uint32_t small_off = (k >> 2) * (kSmall * 128) + (k & 3) * 32;
          ptx::umma_bf16<1>(tmem_base, ptx::make_kmajor_sw128_desc(sK + big_off),
                            ptx::make_kmajor_sw128_desc(sQ + st * C::kSmallBytes + small_off), idesc_s, k > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t big_off = (k >> 2) * (kBig * 128) + (k & 3) * 32;
          const uint32_t small_off = (k >> 2) * (kSmall * 128) + (k & 3) * 32;
          ptx::umma_bf16<1>(tmem_base + kSmall, ptx::make_kmajor_sw128_desc(sV + big_off),
                            ptx::make_kmajor_sw128_desc(sDO + st * C::kSmallBytes + small_off), idesc_s, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit<1>(bar_s);
        ptx::mbar_wait(bar_p, it & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < kSmall / 16; ++k) {
          ptx::umma_bf16<1>(tmem_base + 2 * kSmall, ptx::make_kmajor_sw128_desc(sPT + k * 32),
                            ptx::make_mnmajor_sw128_desc(sDO + st * C::kSmallBytes + k * 2048, kSmall * 128), idesc_o,
                            (it > 0 || k > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < kSmall / 16; ++k) {
          ptx::umma_bf16<1>(tmem_base + 2 * kSmall + D, ptx::make_kmajor_sw128_desc(sDST + k * 32),
                            ptx::make_mnmajor_sw128_desc(sQ + st * C::kSmallBytes + k * 2048, kSmall * 128), idesc_o,
                            (it > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit<1>(bar_o);
      }
    }
  } else {
    const int r = warp * 32 + lane;  // key row of this thread
    const int key = k0 + r;
    const bool kvalid = key < L;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_base;
    const uint32_t tDP = tS + kSmall;
    const float sc2 = p.scale_log2, sc = p.scale;
    for (int it = 0; it < iters; ++it) {
      const int hq = hk * G + it / per_head;
      const int q0 = (i0 + it % per_head) * kSmall;
      float* st_lse = stat + (it & 1) * 2 * kSmall;
      float* st_dl = st_lse + kSmall;
      if (r < kSmall) {
        const int qr = q0 + r;
        const bool ok = qr < L;
        const long long idx = static_cast<long long>(s0 + qr) * p.Hq + hq;
        st_lse[r] = ok ? p.lse[idx] : 0.f;
        st_dl[r] = ok ? p.delta[idx] * p.scale : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the 128 softmax threads only
      // whole tile visible to every key row of this CTA: all keys and queries in range (and, causal, queries >= keys)
      const bool full_blk = (k0 + kBig <= L) && (q0 + kSmall <= L) && (!p.causal || q0 >= k0 + kBig - 1);
      uint32_t dk1 = 0u, dk2 = 0u;
      if (DROP) rng_stream_key(p.drop.seed, p.drop.site, static_cast<uint32_t>(seq), static_cast<uint32_t>(hq), &dk1, &dk2);
      ptx::mbar_wait(bar_s, it & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kSmall / 32; ++c) {
        uint32_t rs[32], rd[32];
        ptx::tmem_ld_32x32(tS + c * 32, rs);
        ptx::tmem_ld_32x32(tDP + c * 32, rd);
        ptx::tmem_ld_wait();
        uint32_t pp[16], pd[16];
        float mk[32];  // dropout multiplier of (query q0 + 32c + i, this key); dead code without DROP
        if (DROP) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint32_t e = (static_cast<uint32_t>(q0 + c * 32 + i) << 16) | static_cast<uint32_t>(key);
            mk[i] = rng_keep(e, dk1, dk2, p.drop.thresh) ? p.drop.inv_keep : 0.f;
            rd[i] = __float_as_uint(__uint_as_float(rd[i]) * mk[i]);
          }
        }
        if (!full_blk) {  // one code path (see the dQ kernel): masked (query, key) pairs get a score of -inf
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int qa = q0 + c * 32 + i;
            const bool ok = kvalid && qa < L && (!p.causal || key <= qa);
            rs[i] = ok ? rs[i] : 0xff800000u;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 nl = *reinterpret_cast<const float2*>(st_lse + c * 32 + i);  // lse (log2 domain)
          const float2 ds = *reinterpret_cast<const float2*>(st_dl + c * 32 + i);   // delta * scale
          const float p0 = ex2b(fmaf(__uint_as_float(rs[i]), sc2, -nl.x));
          const float p1 = ex2b(fmaf(__uint_as_float(rs[i + 1]), sc2, -nl.y));
          pp[i >> 1] = DROP ? pack_op<F16>(p0 * mk[i], p1 * mk[i + 1]) : pack_op<F16>(p0, p1);
          pd[i >> 1] = pack_op<F16>(p0 * fmaf(__uint_as_float(rd[i]), sc, -ds.x),
                                 p1 * fmaf(__uint_as_float(rd[i + 1]), sc, -ds.y));
        }
        st_row_chunk(sPT, r, c, pp);
        st_row_chunk(sDST, r, c, pd);
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_p);
    }
    ptx::mbar_wait(bar_o, (iters - 1) & 1);
    ptx::tc_fence_after();
    const long long grow = static_cast<long long>(s0 + key);
    store_acc_row<D, F16>(tS + 2 * kSmall, p.dv + grow * p.ld_d + hk * D, kvalid, 1.0f, nullptr);
    const float* rope_row = p.rope_cs != nullptr ? p.rope_cs + static_cast<long long>(kvalid ? key : 0) * D : nullptr;
    store_acc_row<D, F16>(tS + 2 * kSmall + D, p.dk + grow * p.ld_d + hk * D, kvalid, 1.0f, rope_row);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, C::kTmemCols);
  }
}

template <int D, bool DROP, bool F16>
int launch_bwd_tc(const void* q, const void* k, const void* v, long long ld_qkv, const void* dout, long long ld_do,
                  const BwdTcParams& p, int num_seqs, int max_seqlen, long long total_rows, cudaStream_t stream) {
  auto kdq = attn_bwd_dq_tc_kernel<D, DROP, F16>;
  auto kdkv = attn_bwd_dkv_tc_kernel<D, DROP, F16>;
  static bool attr_set = false;
  if (!attr_set) {
    B2S_CUDA_CHECK(cudaFuncSetAttribute(kdq, cudaFuncAttributeMaxDynamicSharedMemorySize, DqCfg<D>::kSmemBytes));
    B2S_CUDA_CHECK(cudaFuncSetAttribute(kdkv, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvCfg<D>::kSmemBytes));
    attr_set = true;
  }
  const unsigned long long qcols = static_cast<unsigned long long>(p.Hq) * D, kcols = static_cast<unsigned long long>(p.Hkv) * D;
  CUtensorMap tq_big, tdo_big, tk_small, tv_small, tq_small, tdo_small, tk_big, tv_big;
  int rc;
  if ((rc = encode_map_2d_bf16(&tq_big, q, qcols, total_rows, ld_qkv * 2, 64, kBig)) != B2S_OK) return rc;
  if ((rc = encode_map_2d_bf16(&tdo_big, dout, qcols, total_rows, ld_do * 2, 64, kBig)) != B2S_OK) return rc;
  if ((rc = encode_map_2d_bf16(&tk_small, k, kcols, total_rows, ld_qkv * 2, 64, kSmall)) != B2S_OK) return rc;
  if ((rc = encode_map_2d_bf16(&tv_small, v, kcols, total_rows, ld_qkv * 2, 64, kSmall)) != B2S_OK) return rc;
  if ((rc = encode_map_2d_bf16(&tq_small, q, qcols, total_rows, ld_qkv * 2, 64, kSmall)) != B2S_OK) return rc;
  if ((rc = encode_map_2d_bf16(&tdo_small, dout, qcols, total_rows, ld_do * 2, 64, kSmall)) != B2S_OK) return rc;
  if ((rc = encode_map_2d_bf16(&tk_big, k, kcols, total_rows, ld_qkv * 2, 64, kBig)) != B2S_OK) return rc;
  if ((rc = encode_map_2d_bf16(&tv_big, v, kcols, total_rows, ld_qkv * 2, 64, kBig)) != B2S_OK) return rc;
  const int nb = (max_seqlen + kBig - 1) / kBig;
  kdq<<<dim3(nb, p.Hq, num_seqs), kThreadsB, DqCfg<D>::kSmemBytes, stream>>>(tq_big, tdo_big, tk_small, tv_small, p);
  B2S_LAUNCH_CHECK();
  kdkv<<<dim3(nb, p.Hkv, num_seqs), kThreadsB, DkvCfg<D>::kSmemBytes, stream>>>(tq_small, tdo_small, tk_big, tv_big, p);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}


// delta[row, h] = sum_d dO[row, h, d] * O[row, h, d]; one warp per row, 16-byte loads, lanes of one head reduce together
template <int D>
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, long long ld_o,
                  long long ld_do, float* __restrict__ delta, long long rows, int Hq, int f16) {
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
      ld8h(o + row * ld_o + h0 * D + lane * 8, a, f16);
      ld8h(dout + row * ld_do + h0 * D + lane * 8, b, f16);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(a[i], b[i], s);
    }
#pragma unroll
    for (int off = kLanesPerHead / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (h < Hq && lane % kLanesPerHead == 0) delta[row * Hq + h] = s;
  }
}

}  // namespace

// Backward of attention_fwd (ops.cuh). q / k / v / o / dout / dq / dk / dv share one 16-bit format (fmt: B2S_FMT_*);
// head h of a row sits at column h*D of EACH view.
int attention_bwd(const void* q, const void* k, const void* v, long long ld_qkv, const void* o, long long ld_o,
                  const void* dout, long long ld_do, const float* lse, float* delta_ws, void* dq, void* dk, void* dv,
                  long long ld_dqkv, const int* cu_seqlens, int num_seqs, int max_seqlen, long long total_rows, int Hq,
                  int Hkv, int D, float scale, int causal, const float* rope_cs, int fmt, cudaStream_t stream,
                  const AttnDrop* drop) {
  B2S_REQUIRE(q && k && v && o && dout && lse && delta_ws && dq && dk && dv && cu_seqlens, "attention_bwd: null pointer");
  if (drop != nullptr && drop->thresh == 0u) drop = nullptr;
  B2S_REQUIRE(num_seqs > 0 && max_seqlen > 0 && total_rows > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0,
              "attention_bwd: bad sizes");
  B2S_REQUIRE(ld_qkv % 8 == 0 && ld_do % 8 == 0 && ld_o % 8 == 0 && ld_dqkv % 8 == 0,
              "attention_bwd: strides must keep 16-byte row alignment");
  B2S_REQUIRE(D == 64 || D == 128, "attention_bwd: head_dim %d unsupported (64 or 128)", D);
  const bool f16 = fmt != 0;
  {
    const unsigned grid = static_cast<unsigned>((total_rows + 7) / 8);
    const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(o);
    const __nv_bfloat16* dob = reinterpret_cast<const __nv_bfloat16*>(dout);
    if (D == 64) attn_delta_kernel<64><<<grid, 256, 0, stream>>>(ob, dob, ld_o, ld_do, delta_ws, total_rows, Hq, fmt);
    else attn_delta_kernel<128><<<grid, 256, 0, stream>>>(ob, dob, ld_o, ld_do, delta_ws, total_rows, Hq, fmt);
    B2S_LAUNCH_CHECK();
  }
  BwdTcParams p{};
  p.cu = cu_seqlens;
  p.lse = lse;
  p.delta = delta_ws;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq);
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  p.ld_d = ld_dqkv;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.rope_cs = rope_cs;
#define B2S_BWD_GO(D_, DROP_)                                                                                          \
  return f16 ? launch_bwd_tc<D_, DROP_, true>(q, k, v, ld_qkv, dout, ld_do, p, num_seqs, max_seqlen, total_rows, stream) \
             : launch_bwd_tc<D_, DROP_, false>(q, k, v, ld_qkv, dout, ld_do, p, num_seqs, max_seqlen, total_rows, stream)
  if (drop != nullptr) {
    B2S_REQUIRE(D == 64 && max_seqlen < 65536, "attention_bwd: attention dropout supports head_dim 64, seqlen < 65536");
    p.drop = *drop;
    B2S_BWD_GO(64, true);
  }
  if (D == 64) { B2S_BWD_GO(64, false); }
  B2S_BWD_GO(128, false);
#undef B2S_BWD_GO
}

}  // namespace b2s
