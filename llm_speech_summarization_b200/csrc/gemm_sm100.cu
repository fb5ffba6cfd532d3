// gemm_sm100.cu -- the one dense-contraction kernel of the hot path: a persistent, warp-specialised
// tcgen05 GEMM for sm_100a.
//
//   * operands: bf16, both K-major (activations row-major, weights in nn.Linear [out, in] layout),
//     staged into shared memory by TMA with the 128-byte swizzle, BLOCK_K = 64 (one swizzle atom);
//   * math: tcgen05.mma kind::f16, fp32 accumulators in TMEM, UMMA 128 x BN x 16 (cta_group::1) or
//     256 x BN x 16 across a CTA pair (cta_group::2, A split by rows, W split by columns);
//   * pipeline: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//     warps 4-7 = epilogue (TMEM -> registers -> fused epilogue -> global); smem ring of kStages
//     full/empty mbarriers; TMEM accumulators double-buffered (2 x BN columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1; persistent CTAs walk a static, L2-friendly tile order.
//   * fused epilogues: bias, erf-GELU, fp32 residual add, SwiGLU (gate|up packed), rotate-half RoPE.
//
// Replaces (reference path): every nn.Linear / Conv1d(k>1, C_in=512|1024) the reference executes through
// transformers -- HubertAttention/FeedForward projections, the conv feature extractor layers 1-6 and the
// positional conv (TF/models/hubert/modeling_hubert.py:45-92,127-151,262-369), AudioEncoder.embed_projection
// (REF/model/audio_encoder.py:87), LlamaAttention/LlamaMLP projections and lm_head
// (TF/models/llama/modeling_llama.py:171-289, REF/model/audio_llama.py:67).
#include "gemm_sm100.cuh"

#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <utility>
#include <vector>

#include "b2s_common.cuh"
#include "b2s_ptx.cuh"
#include "rng.cuh"

namespace b2s {

namespace {

constexpr int kBlockK = 64;           // bf16 elements = 128 bytes = one swizzle atom
constexpr int kUmmaK = 16;
// warps 0-3: TMA producer, MMA issuer, TMEM allocator, spare; then the epilogue warps, one per TMEM lane quarter.
// Eight epilogue warps -- two per TMEM lane quarter, each taking half of the tile's columns -- are MODE 5 (= MODE 0
// otherwise), chosen by the host for short reductions where the epilogue, not the mainloop, is the critical path. Their
// first form (round 2, both 32-column chunks of a 16-bit staging tile in registers at once) spilled at the 168 registers
// 384 threads leave and was slower on every shape; MODE 5 stages one chunk at a time.
template <int MODE>
constexpr int kEpiWarps = (MODE == 5) ? 8 : 4;
template <int MODE>
constexpr int kThreads = 128 + 32 * kEpiWarps<MODE>;
constexpr int kATileBytes = 128 * kBlockK * 2;  // 16 KiB per CTA per stage
constexpr int kSmemBudget = 200 * 1024;

template <int BN, int CG>
struct Cfg {
  static constexpr int kBRows = BN / CG;                    // W rows loaded per CTA per stage
  static constexpr int kBTileBytes = kBRows * kBlockK * 2;
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = (2 * BN) < 32 ? 32 : 2 * BN;  // power of two for BN in {64,128,256}
  static constexpr int kBarBytes = 256;                  // (2*kStages + 4) mbarriers + the TMEM base slot
  // 32 KiB of epilogue staging, 1024-byte aligned (128B-swizzled TMA store sources): one 32-row x 128-byte tile for
  // each of MODE 0's eight epilogue warps (the bulk store of chunk c drains while chunk c+1 is read from TMEM and
  // activated), 8 KiB per warp for the four-warp modes
  static constexpr int kStagingBytes = 8 * 4096;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;  // +1024 align slack
  static_assert((2 * kStages + 4) * 8 + 16 <= kBarBytes, "barrier region too small");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

struct KParams {
  int M, N;
  int batches, groups;
  int m_tiles, n_tiles;
  int total_tiles;
  int num_kb, kb_per_tap, k_per_tap;
  int a_pad, a_group_off, w_group_off;
  int epi, act;
  const float* bias;
  void* out;
  long long ldo;
  long long out_batch_rows;
  const float* resid;
  int resid_bcast;
  void* out2;
  long long ld2;
  const float* rope_cs;
  const int* positions;
  int rope_cols;
  // MN-major operands / reduction over batches / split-K (backward GEMMs)
  int a_mn, b_mn;         // operand stored [k rows, m|n cols] row-major instead of K-major
  int kb_per_batch;       // k-blocks per reduction batch (num_kb = k_batches * kb_per_batch)
  int k_splits, kb_per_split;
  int tiles_per_split;
  int b_tap_atoms;        // MN-major B: 64-column atom j of the N axis = the same columns shifted by j - a_pad rows
  int og_rows, og_cols;   // output offset per group: rows += g * og_rows, cols += g * og_cols
  uint32_t drop_k1, drop_k2, drop_thresh;  // fused dropout (rng.cuh); thresh 0 = off
  float drop_inv_keep;
  int resid_red;  // in-place residual adds go through red.global.add (B2S_RESID_RED=0 keeps load + add + store, A/B)
  int tma_out;    // MODE 0 plain epilogues: output leaves through TMA stores / reductions (tmap_o is valid)
  int group_m;    // tile order: M tiles per group (a group's tiles walk N with M fastest, see decode_tile)
  // MODE 0 tail split: the persistent grid walks `units` work units. Units below tail_first are whole tiles; the
  // tiles of the last, partly filled round are cut into tail_split K-slices of tail_kb k-blocks each so that the round
  // costs 1 / tail_split of a tile instead of a whole one (M = 6400, N = 3072: 4.05 rounds -> 5 before, 4.25 now).
  // Slices of one tile meet in the fp32 output: slice 0 stores (or all slices reduce-add for the in-place residual),
  // the others reduce-add once slice 0's rows have landed (tail_flags, two ints per tile x CTA x epilogue warp).
  int units, tail_first, tail_split, tail_kb;
  int* tail_flags;
  uint32_t idesc_fmt;  // a_format / b_format bits of the instruction descriptor (bf16 = 1, fp16 = 0; may differ)
  int out_f16;         // 16-bit outputs (out for the *_BF16-class epilogues, out2) are written as fp16 instead of bf16
};

struct TileCoord {
  int b, g, m_t, n_t, split;
};

template <bool EXT>
__device__ __forceinline__ TileCoord decode_tile(const KParams& p, int t) {
  // Tiles are walked group by group: inside a group of `group_m` M-tiles, M is the fast index and N the slow one, so
  // the clusters running at the same time share few W tiles and a group's A tiles stay in L2 for its whole N sweep.
  // W is re-read once per group: DRAM traffic ~ A + W * ceil(m_tiles / group_m) as long as group_m A-tiles fit in L2.
  const int kGroupM = p.group_m;
  int split = 0;
  if constexpr (EXT) {
    split = t / p.tiles_per_split;
    t -= split * p.tiles_per_split;
  }
  const int per_bg = p.m_tiles * p.n_tiles;
  const int bg = t / per_bg;
  const int r = t - bg * per_bg;
  const int span = kGroupM * p.n_tiles;
  const int gid = r / span;
  const int first_m = gid * kGroupM;
  const int gsz = min(p.m_tiles - first_m, kGroupM);
  const int rr = r - gid * span;
  TileCoord c;
  c.m_t = first_m + rr % gsz;
  c.n_t = rr / gsz;
  c.b = bg / p.groups;
  c.g = bg - c.b * p.groups;
  c.split = split;
  return c;
}

struct Unit {
  int tile, split, kb_lo, kb_hi;
};

template <bool EXT>
__device__ __forceinline__ Unit decode_unit(const KParams& p, int u) {
  Unit r;
  if constexpr (EXT) {  // split-K over the whole problem: the split index is the slow part of the tile index
    r.tile = u;
    r.split = u / p.tiles_per_split;
    r.kb_lo = r.split * p.kb_per_split;
    r.kb_hi = min(p.num_kb, r.kb_lo + p.kb_per_split);
  } else if (u < p.tail_first) {
    r.tile = u;
    r.split = 0;
    r.kb_lo = 0;
    r.kb_hi = p.num_kb;
  } else {
    const int j = u - p.tail_first;
    const int q = j / p.tail_split;
    r.split = j - q * p.tail_split;
    r.tile = p.tail_first + q;
    r.kb_lo = r.split * p.tail_kb;
    r.kb_hi = min(p.num_kb, r.kb_lo + p.tail_kb);
  }
  return r;
}

// ---- epilogue helpers ------------------------------------------------------------------------
// nvalid: columns of this 32-wide chunk that exist (N is a multiple of 8, so whole float4 groups; the bias array ends at N)
__device__ __forceinline__ void add_bias_act(float (&v)[32], const float* bias, int act, int nvalid = 32) {
  if (bias != nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * q < nvalid) b4 = __ldg(reinterpret_cast<const float4*>(bias) + q);
      v[4 * q + 0] += b4.x;
      v[4 * q + 1] += b4.y;
      v[4 * q + 2] += b4.z;
      v[4 * q + 3] += b4.w;
    }
  }
  if (act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
  }
}

// ---- staged (transposing) stores ---------------------------------------------------------------
// Staging tile: 32 rows x 128 B (32 fp32), 16-byte chunk q of row r stored at chunk position q ^ (r & 7):
// the per-thread row writes and the per-row coalesced reads are both bank-conflict free.
__device__ __forceinline__ void stage_write(uint32_t stg, const float (&v)[32], int lane) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t addr = stg + lane * 128 + ((q ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * q]), "f"(v[4 * q + 1]),
                 "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                 : "memory");
  }
}
__device__ __forceinline__ float4 stage_read(uint32_t stg, int row, int chunk) {
  float4 x;
  const uint32_t addr = stg + row * 128 + ((chunk ^ (row & 7)) << 4);
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(addr));
  return x;
}

// v = this thread's row (lane) x 32 columns. out0 / resid0 point at element [warp's first row][first column].
__device__ __forceinline__ void emit_f32(uint32_t stg, const float (&v)[32], int lane, float* out0,
                                         const float* resid0, long long ld, int rows_valid, int cols_valid) {
  stage_write(stg, v, lane);
  __syncwarp();
  const int chunk = lane & 7;
  float4 x[8], r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    x[i] = stage_read(stg, row, chunk);
    r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (resid0 != nullptr && row < rows_valid && chunk * 4 < cols_valid) {
      r[i] = *reinterpret_cast<const float4*>(resid0 + row * ld + chunk * 4);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    if (row < rows_valid && chunk * 4 < cols_valid) {
      *reinterpret_cast<float4*>(out0 + row * ld + chunk * 4) =
          make_float4(x[i].x + r[i].x, x[i].y + r[i].y, x[i].z + r[i].z, x[i].w + r[i].w);
    }
  }
  __syncwarp();
}

// residual tile of one 32x32 chunk in the coalesced (row = i*4 + lane/8, 16-byte chunk = lane%8) register layout
__device__ __forceinline__ void load_resid(float4 (&r)[8], const float* resid0, long long ld, int lane, int rows_valid,
                                           int cols_valid) {
  const int chunk = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (resid0 != nullptr && row < rows_valid && chunk * 4 < cols_valid) {
      r[i] = *reinterpret_cast<const float4*>(resid0 + row * ld + chunk * 4);
    }
  }
}

// emit_f32 with the residual already in registers (loaded one chunk ahead: the global-load latency of the residual
// overlaps the TMEM read / transpose of the previous chunk instead of sitting between them)
__device__ __forceinline__ void emit_f32_pre(uint32_t stg, const float (&v)[32], int lane, float* out0,
                                             const float4 (&r)[8], long long ld, int rows_valid, int cols_valid) {
  stage_write(stg, v, lane);
  __syncwarp();
  const int chunk = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    const float4 x = stage_read(stg, row, chunk);
    if (row < rows_valid && chunk * 4 < cols_valid) {
      *reinterpret_cast<float4*>(out0 + row * ld + chunk * 4) =
          make_float4(x.x + r[i].x, x.y + r[i].y, x.z + r[i].z, x.w + r[i].w);
    }
  }
  __syncwarp();
}

// out += v (fp32 reduction in L2: split-K partial sums and gradient accumulation across micro-batches)
__device__ __forceinline__ void emit_accum_f32(uint32_t stg, const float (&v)[32], int lane, float* out0, long long ld,
                                               int rows_valid, int cols_valid) {
  stage_write(stg, v, lane);
  __syncwarp();
  const int chunk = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i * 4 + (lane >> 3);
    const float4 x = stage_read(stg, row, chunk);
    if (row < rows_valid && chunk * 4 < cols_valid) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out0 + row * ld + chunk * 4), "f"(x.x),
                   "f"(x.y), "f"(x.z), "f"(x.w)
                   : "memory");
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void emit_h16(uint32_t stg, const float (&v)[32], int lane, __nv_bfloat16* out0,
                                         long long ld, int rows_valid, int cols_valid, int f16) {
  // rows are packed to bf16 / fp16 BEFORE staging (64 B per row, half the shared-memory traffic of the fp32 tile);
  // 16-byte chunk q of row r sits at chunk position q ^ ((r >> 1) & 3): conflict-free both ways.
  // (f16 is CTA-uniform: one of the two conversion sequences runs, nothing is selected per element)
  if (f16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t addr = stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_f16(v[8 * q], v[8 * q + 1])),
                   "r"(pack_f16(v[8 * q + 2], v[8 * q + 3])), "r"(pack_f16(v[8 * q + 4], v[8 * q + 5])),
                   "r"(pack_f16(v[8 * q + 6], v[8 * q + 7]))
                   : "memory");
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t addr = stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(v[8 * q], v[8 * q + 1])),
                   "r"(pack_bf16(v[8 * q + 2], v[8 * q + 3])), "r"(pack_bf16(v[8 * q + 4], v[8 * q + 5])),
                   "r"(pack_bf16(v[8 * q + 6], v[8 * q + 7]))
                   : "memory");
    }
  }
  __syncwarp();
  const int j = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i * 8 + (lane >> 2);
    uint4 u;
    const uint32_t addr = stg + row * 64 + ((j ^ ((row >> 1) & 3)) << 4);
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
    if (row < rows_valid && j * 8 < cols_valid) {
      *reinterpret_cast<uint4*>(out0 + row * ld + j * 8) = u;
    }
  }
  __syncwarp();
}

// MODE 0: the forward instantiation (K-major operands, no split-K, no second output) -- the hot loops carry nothing
//         they do not need.
// MODE 1: K-major operands + split-K, the accumulate epilogue, per-group output offsets and the pre-activation copy
//         (training forward, decode).
// MODE 2: MODE 1 with an MN-major B operand (dgrad).   MODE 3: both operands MN-major, reduction over batches (wgrad).
// MODE 4: MODE 2 whose plain 16-bit / fp32 output leaves through TMA like MODE 0's (the encoder's dgrad GEMMs).
// Operand major-ness is a compile-time property so the single-thread producer / MMA-issue loops stay branch-free.
template <int BN, int CG, int MODE>
__global__ void __launch_bounds__(kThreads<MODE>, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_o, const KParams p) {
  using C = Cfg<BN, CG>;
  constexpr int kStages = C::kStages;
  constexpr bool EXT = MODE >= 1 && MODE <= 4;
  constexpr bool kAmn = MODE == 3, kBmn = MODE >= 2 && MODE <= 4;
  constexpr bool kTma = MODE == 0 || MODE == 4 || MODE == 5;  // plain outputs leave through TMA stores / reductions

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base + kStages * C::kStageBytes;  // epilogue staging tiles (1024-byte aligned)
  const uint32_t bar_base = stage_base + C::kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;
  // PDL (b2s_common.cuh): this grid may have been scheduled while its predecessor drains -- barrier init, TMEM
  // allocation and the cluster handshakes below touch no global memory and overlap the predecessor's tail; every role
  // passes pdl_wait() before its first global access. And a dependent GEMM may take the SMs this grid frees.
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(full_bar(s), CG);   // one arrive(+tx) per CTA producer, all on the leader's barrier
      ptx::mbar_init(empty_bar(s), 1);   // one tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(tfull_bar(s), 1);        // one tcgen05.commit
      ptx::mbar_init(tempty_bar(s), kEpiWarps<MODE> * CG);  // one arrive per epilogue warp (both CTAs -> leader)
    }
    ptx::fence_mbar_init();
  }
  if (CG == 2) ptx::cluster_sync_all();  // peer smem must be live before a 2-SM allocation
  if (warp == 2) {
    ptx::tmem_alloc<CG>(tmem_slot, C::kTmemCols);
  }
  ptx::tc_fence_before();
  if (CG == 2) {
    ptx::cluster_sync_all();
  } else {
    __syncthreads();
  }
  ptx::tc_fence_after();

  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // the predecessor's writes (A operand, residual, the buffer this grid overwrites) are complete from here

  if (warp == 0 && lane == 0) {
    // ============================== TMA producer ==============================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = cluster_id; t < p.units; t += num_clusters) {
      const Unit un = decode_unit<EXT>(p, t);
      const TileCoord tc = decode_tile<EXT>(p, un.tile);
      const int m0 = tc.m_t * (128 * CG) + static_cast<int>(cta_rank) * 128;
      const int n0 = tc.g * p.w_group_off + tc.n_t * BN + static_cast<int>(cta_rank) * C::kBRows;
      const int a_c0_base = tc.g * p.a_group_off;
      const int kb_lo = un.kb_lo, kb_hi = un.kb_hi;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * C::kStageBytes;
        const uint32_t sb = sa + kATileBytes;
        if (CG == 1) {
          ptx::mbar_arrive_expect_tx(full_bar(stage), C::kStageBytes);
        } else if (cta_rank == 0) {
          ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * C::kStageBytes);
        } else {
          ptx::mbar_arrive_cluster(full_bar(stage), 0);
        }
        if constexpr (!kBmn) {
          const int tap = kb / p.kb_per_tap;
          const int kk = (kb - tap * p.kb_per_tap) * kBlockK;
          if (CG == 1) {
            ptx::tma_load_3d(&tmap_a, full_bar(stage), sa, a_c0_base + kk, m0 + tap - p.a_pad, tc.b);
            ptx::tma_load_2d(&tmap_w, full_bar(stage), sb, tap * p.k_per_tap + kk, n0);
          } else {
            ptx::tma_load_3d_2sm(&tmap_a, full_bar(stage), sa, a_c0_base + kk, m0 + tap - p.a_pad, tc.b);
            ptx::tma_load_2d_2sm(&tmap_w, full_bar(stage), sb, tap * p.k_per_tap + kk, n0);
          }
        } else {
          // MN-major operands: the reduction index walks ROWS of the source (k-batch kbat, row kk); a tile is a
          // set of [64 k-rows x 64 columns] boxes, one per 64-wide M / N atom, 8 KiB apart.
          const int kbat = kb / p.kb_per_batch;
          const int kk = (kb - kbat * p.kb_per_batch) * kBlockK;
          if constexpr (kAmn) {
#pragma unroll
            for (int a = 0; a < 2; ++a) {
              if (CG == 1) ptx::tma_load_3d(&tmap_a, full_bar(stage), sa + a * 8192, a_c0_base + m0 + a * 64, kk, kbat);
              else ptx::tma_load_3d_2sm(&tmap_a, full_bar(stage), sa + a * 8192, a_c0_base + m0 + a * 64, kk, kbat);
            }
          } else {
            if (CG == 1) ptx::tma_load_3d(&tmap_a, full_bar(stage), sa, a_c0_base + kk, m0, tc.b);
            else ptx::tma_load_3d_2sm(&tmap_a, full_bar(stage), sa, a_c0_base + kk, m0, tc.b);
          }
#pragma unroll
          for (int a = 0; a < C::kBRows / 64; ++a) {
            int col, row;
            if (p.b_tap_atoms) {
              col = tc.g * p.w_group_off;
              row = kk + (tc.n_t * BN + static_cast<int>(cta_rank) * C::kBRows) / 64 + a - p.a_pad;
            } else {
              col = n0 + a * 64;
              row = kk;
            }
            if (CG == 1) ptx::tma_load_3d(&tmap_w, full_bar(stage), sb + a * 8192, col, row, kbat);
            else ptx::tma_load_3d_2sm(&tmap_w, full_bar(stage), sb + a * 8192, col, row, kbat);
          }
        }
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 && lane == 0 && cta_rank == 0) {
    // ============================== MMA issuer ==============================
    constexpr bool a_mn = kAmn, b_mn = kBmn;
    const uint32_t idesc =
        ptx::make_idesc_f32acc(128 * CG, BN) | p.idesc_fmt | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
    // K step of 16 inside a stage: K-major = 32 bytes along the swizzled row; MN-major = 16 rows of 128 bytes
    constexpr uint32_t a_kstep = a_mn ? (2048u >> 4) : 2u;
    constexpr uint32_t b_kstep = b_mn ? (2048u >> 4) : 2u;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = cluster_id; t < p.units; t += num_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * BN;
      const Unit un = decode_unit<EXT>(p, t);
      const int kb_lo = un.kb_lo, kb_hi = un.kb_hi;
      for (int kb = kb_lo; kb < kb_hi; ++kb) {
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * C::kStageBytes;
        const uint64_t adesc = a_mn ? ptx::make_mnmajor_sw128_desc(sa, 8192) : ptx::make_kmajor_sw128_desc(sa);
        const uint64_t bdesc = b_mn ? ptx::make_mnmajor_sw128_desc(sa + kATileBytes, 8192)
                                    : ptx::make_kmajor_sw128_desc(sa + kATileBytes);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          ptx::umma_bf16<CG>(tmem_d, adesc + a_kstep * k, bdesc + b_kstep * k, idesc,
                             ((kb - kb_lo) | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit<CG>(empty_bar(stage));  // frees this smem stage (both CTAs) when the MMAs retire
        if (kb == kb_hi - 1) ptx::umma_commit<CG>(tfull_bar(as));
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==============================
    // Each warp owns TMEM lanes [32*quarter, +32) = 32 output rows. A thread reads ONE row x 32 columns from TMEM;
    // writing that straight to global memory would touch 32 different cache lines per instruction, so every
    // 32x32 chunk is transposed through a private, XOR-swizzled 4 KiB shared-memory tile and leaves the SM as
    // fully coalesced 128-byte row segments (and the fp32 residual is read the same way).
    const int quarter = warp & 3;
    constexpr int kHalves = kEpiWarps<MODE> / 4;        // warps sharing one TMEM lane quarter split the tile's columns
    const int half = (warp - 4) >> 2;                   // 0 .. kHalves-1
    const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t stg = stage_base + static_cast<uint32_t>(warp - 4) * (kHalves == 2 ? 4096u : 8192u);
    uint32_t nst = 0;  // staging tiles this warp has handed to the TMA store engine (buffer = nst & 1)
    if (kTma && lane == 0) ptx::prefetch_tmap(&tmap_o);
    int it = 0;
    for (int t = cluster_id; t < p.units; t += num_clusters, ++it) {
      const Unit un = decode_unit<EXT>(p, t);
      const TileCoord tc = decode_tile<EXT>(p, un.tile);
      const int og_cols = EXT ? p.og_cols : p.N;
      void* const out2 = EXT ? p.out2 : nullptr;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int m0w = tc.m_t * (128 * CG) + static_cast<int>(cta_rank) * 128 + quarter * 32;  // warp's first row
      const int rows_valid = max(0, min(32, p.M - m0w));
      const bool row_ok = lane < rows_valid;
      const long long orow0 = static_cast<long long>(tc.b) * p.out_batch_rows +
                              (EXT ? static_cast<long long>(tc.g) * p.og_rows : 0LL) + m0w;
      const int ncol0 = tc.n_t * BN;  // column inside the group
      const float* bias = p.bias ? p.bias + static_cast<long long>(tc.g) * p.N : nullptr;
      (void)un;

      ptx::mbar_wait(tfull_bar(as), aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + as * BN;

      if (p.epi == EPI_BF16 || p.epi == EPI_F32 || p.epi == EPI_RESID_F32 || (EXT && p.epi == EPI_ACCUM_F32)) {
        if constexpr (kTma) {
        // (MODE 0: the host routes every plain epilogue it cannot express as a TMA store to the MODE 1 kernel; MODE 4 is
        // the dgrad kernel with this epilogue. Compiling both epilogues into one kernel cost the per-thread one 7-8 %
        // through register pressure, so a kernel has exactly one of them.)
        // ---- MODE 0 plain epilogues (bf16 / fp16 / fp32 output, fp32 in-place residual): TMEM -> registers ->
        // bias / activation -> 128B-swizzled staging tile -> ONE bulk tensor store (or fp32 add-reduction at the L2 for
        // h += proj(...)) per 32-row x 128-byte tile. No per-thread global stores, no read-back of the staging tile, the
        // write drains asynchronously while the next chunk is computed, and TMA clips the M / N edges.
        const bool h16 = p.epi == EPI_BF16;
        // K-slices of a tail tile: slice 0 carries the bias and (fp32 store) writes first; later slices add into it
        const bool tail_tile = t >= p.tail_first;
        const bool part = un.split > 0;
        const bool red = p.epi == EPI_RESID_F32 || p.epi == EPI_ACCUM_F32 || part;
        if (part) bias = nullptr;
        int* const tflag = p.tail_flags +
                           2 * ((((un.tile - p.tail_first) * CG + static_cast<int>(cta_rank)) * 4 + quarter) * 2 + half);
        if (part && p.epi == EPI_F32) {
          if (lane == 0) {
            // bounded like every other wait of this kernel: slice 0 runs on a lower-numbered, co-resident CTA pair and
            // publishes within microseconds; ~4 s without the flag means a broken launch, and a trap beats a hang
            unsigned spins = 0;
            while (ptx::ld_acquire_gpu(tflag) == 0) {
              __nanosleep(64);
              if (++spins > (1u << 26)) __trap();
            }
            ptx::fence_proxy_async_all();
          }
          __syncwarp();
        }
        const int cw = h16 ? 64 : 32;  // output columns per staging tile
        const int gcol0 = tc.g * p.N + ncol0;
        constexpr int kChunksPerWarp = (BN / 32) / kHalves;  // 32-column TMEM chunks this warp drains
        if constexpr (kHalves == 2) {
          // eight warps: one staging tile per warp, ONE 32-column chunk in registers at a time (a 16-bit tile is filled by
          // two consecutive chunks, left then right half of its 128-byte rows)
#pragma unroll 1
          for (int c = half * kChunksPerWarp; c < (half + 1) * kChunksPerWarp; c += (h16 ? 2 : 1)) {
            const int col = ncol0 + c * 32;
            if (col >= p.N) break;
            if (lane == 0) ptx::bulk_wait_read<0>();  // the store that last used this tile has read it
            __syncwarp();
#pragma unroll 1
            for (int hc = 0; hc < (h16 ? 2 : 1); ++hc) {
              uint32_t raw[32];
              ptx::tmem_ld_32x32(taddr + (c + hc) * 32, raw);
              ptx::tmem_ld_wait();
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
              const int cc = col + 32 * hc;
              add_bias_act(v, (bias && cc < p.N) ? bias + cc : nullptr, p.act, p.N - cc);
              if (h16) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float* s8 = &v[8 * q];
                  const uint32_t addr = stg + lane * 128 + (((4 * hc + q) ^ (lane & 7)) << 4);
                  const uint32_t w0 = p.out_f16 ? pack_f16(s8[0], s8[1]) : pack_bf16(s8[0], s8[1]);
                  const uint32_t w1 = p.out_f16 ? pack_f16(s8[2], s8[3]) : pack_bf16(s8[2], s8[3]);
                  const uint32_t w2 = p.out_f16 ? pack_f16(s8[4], s8[5]) : pack_bf16(s8[4], s8[5]);
                  const uint32_t w3 = p.out_f16 ? pack_f16(s8[6], s8[7]) : pack_bf16(s8[6], s8[7]);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                               : "memory");
                }
              } else {
                stage_write(stg, v, lane);
              }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (rows_valid > 0) {
                if (red) ptx::tma_reduce_add_3d(&tmap_o, stg, gcol0 + c * 32, m0w, tc.b);
                else ptx::tma_store_3d(&tmap_o, stg, gcol0 + c * 32, m0w, tc.b);
              }
              ptx::bulk_commit();
            }
          }
        } else {
#pragma unroll 1
        for (int c = half * kChunksPerWarp; c < (half + 1) * kChunksPerWarp; c += (h16 ? 2 : 1)) {
          const int col = ncol0 + c * 32;
          if (col >= p.N) break;
          uint32_t raw[32], raw2[32];
          ptx::tmem_ld_32x32(taddr + c * 32, raw);
          if (h16) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, raw2);
          ptx::tmem_ld_wait();
          float v[32], v2[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          add_bias_act(v, (bias && col < p.N) ? bias + col : nullptr, p.act, p.N - col);
          if (h16) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v2[i] = __uint_as_float(raw2[i]);
            add_bias_act(v2, (bias && col + 32 < p.N) ? bias + col + 32 : nullptr, p.act, p.N - col - 32);
          }
          // four warps: two staging tiles per warp, the store of chunk c drains while chunk c+1 is staged
          const uint32_t buf = stg + (kHalves == 1 ? (nst & 1u) * 4096u : 0u);
          if (lane == 0) ptx::bulk_wait_read<(kHalves == 1 ? 1 : 0)>();  // the store that last used this tile has read it
          __syncwarp();
          if (h16) {
            // row = lane: 64 values = 128 bytes = eight 16-byte chunks at chunk position q ^ (lane & 7)
            if (p.out_f16) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float* s = q < 4 ? &v[8 * q] : &v2[8 * (q - 4)];
                const uint32_t addr = buf + lane * 128 + ((q ^ (lane & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_f16(s[0], s[1])),
                             "r"(pack_f16(s[2], s[3])), "r"(pack_f16(s[4], s[5])), "r"(pack_f16(s[6], s[7]))
                             : "memory");
              }
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float* s = q < 4 ? &v[8 * q] : &v2[8 * (q - 4)];
                const uint32_t addr = buf + lane * 128 + ((q ^ (lane & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16(s[0], s[1])),
                             "r"(pack_bf16(s[2], s[3])), "r"(pack_bf16(s[4], s[5])), "r"(pack_bf16(s[6], s[7]))
                             : "memory");
              }
            }
          } else {
            stage_write(buf, v, lane);
          }
          ptx::fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy) read
          __syncwarp();
          if (lane == 0) {
            if (rows_valid > 0) {
              if (red) ptx::tma_reduce_add_3d(&tmap_o, buf, gcol0 + c * 32, m0w, tc.b);
              else ptx::tma_store_3d(&tmap_o, buf, gcol0 + c * 32, m0w, tc.b);
            }
            ptx::bulk_commit();  // (an empty group when the warp's rows lie beyond M: the two staging tiles' wait_group
                                 // accounting must advance with every chunk, stored or not)
          }
          ++nst;
          (void)cw;
        }
        }
        if (tail_tile && p.epi == EPI_F32 && lane == 0) {
          if (!part) {
            ptx::bulk_wait<0>();  // this warp's 32 rows of slice 0 are in memory
            ptx::fence_proxy_async_all();
            __threadfence();
            ptx::st_release_gpu(tflag, 1);
          } else if (atomicAdd(tflag + 1, 1) == p.tail_split - 2) {  // last adder: re-arm the pair for the next launch
            tflag[1] = 0;
            ptx::st_release_gpu(tflag, 0);
          }
        }
        } else {
        // In-place residual (out aliases resid: the inference forward's h += proj(...)): the add is done by the L2 as
        // an fp32 reduction (red.global.add.v4.f32, one per element, so still deterministic) and the residual never
        // travels to the SM -- the load -> add -> store chain with one 4 KiB chunk per warp in flight was what bound
        // the K = 1024 shapes (profiles/r01_gemm_shapes.md).
        const bool resid_inplace = p.resid_red && p.epi == EPI_RESID_F32 && static_cast<const void*>(p.resid) == p.out &&
                                   !p.resid_bcast;
        const bool has_resid = p.epi == EPI_RESID_F32 && !resid_inplace;
        const long long rbase = (p.resid_bcast ? static_cast<long long>(m0w) : orow0) * p.ldo +
                                static_cast<long long>(tc.g) * og_cols;
        float4 rnext[8];
        if (has_resid) load_resid(rnext, p.resid + rbase + ncol0, p.ldo, lane, rows_valid, min(32, p.N - ncol0));
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col = ncol0 + c * 32;
          if (col >= p.N) break;
          float4 rcur[8];
          if (has_resid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) rcur[i] = rnext[i];
            if (c + 1 < BN / 32 && col + 32 < p.N)
              load_resid(rnext, p.resid + rbase + col + 32, p.ldo, lane, rows_valid, min(32, p.N - col - 32));
          }
          uint32_t raw[32];
          ptx::tmem_ld_32x32(taddr + c * 32, raw);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
          const int valid = min(32, p.N - col);
          if (EXT && out2 != nullptr) {  // training: keep the pre-activation (acc + bias) for the backward
            add_bias_act(v, bias ? bias + col : nullptr, ACT_NONE);
            emit_h16(stg, v, lane,
                      reinterpret_cast<__nv_bfloat16*>(out2) + orow0 * p.ld2 + static_cast<long long>(tc.g) * og_cols + col,
                      p.ld2, rows_valid, valid, p.out_f16);
            add_bias_act(v, nullptr, p.act);
          } else {
            add_bias_act(v, bias ? bias + col : nullptr, p.act);
          }
          const long long off0 = orow0 * p.ldo + static_cast<long long>(tc.g) * og_cols + col;
          if (EXT && p.drop_thresh != 0u) {  // train-mode dropout: this thread's row, 32 consecutive columns
            const uint32_t e0 = static_cast<uint32_t>(off0 + static_cast<long long>(lane) * p.ldo);
#pragma unroll
            for (int i = 0; i < 32; ++i)
              v[i] = rng_keep(e0 + i, p.drop_k1, p.drop_k2, p.drop_thresh) ? v[i] * p.drop_inv_keep : 0.f;
          }
          if (p.epi == EPI_BF16) {
            emit_h16(stg, v, lane, reinterpret_cast<__nv_bfloat16*>(p.out) + off0, p.ldo, rows_valid, valid, p.out_f16);
          } else if ((EXT && p.epi == EPI_ACCUM_F32) || resid_inplace) {
            emit_accum_f32(stg, v, lane, reinterpret_cast<float*>(p.out) + off0, p.ldo, rows_valid, valid);
          } else if (has_resid) {
            // (resid_bcast: the residual is indexed by the row inside the batch only, e.g. a positional table)
            emit_f32_pre(stg, v, lane, reinterpret_cast<float*>(p.out) + off0, rcur, p.ldo, rows_valid, valid);
          } else {
            emit_f32(stg, v, lane, reinterpret_cast<float*>(p.out) + off0, nullptr, p.ldo, rows_valid, valid);
          }
        }
        }
      } else {
        // paired-chunk epilogues over 128-column blocks: chunk c pairs with chunk c+2
        // (SwiGLU: 64 gate | 64 up ; RoPE: head_dim 128 = first half | second half)
        int pos = 0;
        if (p.epi == EPI_ROPE && row_ok) pos = __ldg(p.positions + orow0 + lane);
        constexpr int kBlocks = BN / 128;
        constexpr int kBlkPerWarp = (kBlocks >= kHalves) ? kBlocks / kHalves : 1;
        const int blk_lo = (kBlocks >= kHalves) ? half * kBlkPerWarp : (half == 0 ? 0 : kBlocks);
#pragma unroll 1
        for (int blk = blk_lo; blk < min(kBlocks, blk_lo + kBlkPerWarp); ++blk) {
          const int bcol = ncol0 + blk * 128;
          if (bcol >= p.N) break;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t raw_lo[32], raw_hi[32];
            ptx::tmem_ld_32x32(taddr + blk * 128 + c * 32, raw_lo);
            ptx::tmem_ld_32x32(taddr + blk * 128 + (c + 2) * 32, raw_hi);
            ptx::tmem_ld_wait();
            float lo[32], hi[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              lo[i] = __uint_as_float(raw_lo[i]);
              hi[i] = __uint_as_float(raw_hi[i]);
            }
            add_bias_act(lo, bias ? bias + bcol + c * 32 : nullptr, ACT_NONE);
            add_bias_act(hi, bias ? bias + bcol + (c + 2) * 32 : nullptr, ACT_NONE);
            if (p.epi == EPI_SWIGLU) {
              if (EXT && out2 != nullptr) {  // training: raw gate | up (the GEMM's natural [M, N] layout)
                __nv_bfloat16* raw = reinterpret_cast<__nv_bfloat16*>(out2) + orow0 * p.ld2 +
                                     static_cast<long long>(tc.g) * p.N + bcol + c * 32;
                emit_h16(stg, lo, lane, raw, p.ld2, rows_valid, 32, p.out_f16);
                emit_h16(stg, hi, lane, raw + 64, p.ld2, rows_valid, 32, p.out_f16);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) lo[i] = silu(lo[i]) * hi[i];
              const long long off0 = orow0 * p.ldo + (static_cast<long long>(tc.g) * p.N + bcol) / 2 + c * 32;
              emit_h16(stg, lo, lane, reinterpret_cast<__nv_bfloat16*>(p.out) + off0, p.ldo, rows_valid, 32, p.out_f16);
            } else {  // EPI_ROPE
              if (bcol < p.rope_cols && row_ok) {
                const float* cs = p.rope_cs + static_cast<long long>(pos) * 128 + c * 32;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs) + q);
                  const float4 s4 = __ldg(reinterpret_cast<const float4*>(cs + 64) + q);
                  const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                  const float ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float a = lo[4 * q + j], b = hi[4 * q + j];
                    lo[4 * q + j] = a * cc[j] - b * ss[j];
                    hi[4 * q + j] = b * cc[j] + a * ss[j];
                  }
                }
              }
              const long long off0 = orow0 * p.ldo + static_cast<long long>(tc.g) * p.N + bcol + c * 32;
              emit_h16(stg, lo, lane, reinterpret_cast<__nv_bfloat16*>(p.out) + off0, p.ldo, rows_valid, 32, p.out_f16);
              emit_h16(stg, hi, lane, reinterpret_cast<__nv_bfloat16*>(p.out) + off0 + 64, p.ldo, rows_valid, 32, p.out_f16);
            }
          }
        }
      }
      // release this accumulator buffer back to the MMA issuer (leader CTA's barrier)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) {
          ptx::mbar_arrive(tempty_bar(as));
        } else {
          ptx::mbar_arrive_cluster(tempty_bar(as), 0);
        }
      }
    }
    if (kTma && lane == 0) ptx::bulk_wait<0>();  // every bulk store of this warp has landed before the CTA retires
  }

  ptx::tc_fence_before();
  if (CG == 2) {
    ptx::cluster_sync_all();
  } else {
    __syncthreads();
  }
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, C::kTmemCols);
  }
}

// ---------------------------------------------------------------- host side
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeFn>(sym);
    }
  });
  return fn;
}

int encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  EncodeFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return B2S_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dtype, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u)",
                   static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides_bytes[0],
                   (unsigned long long)(rank > 2 ? strides_bytes[1] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return B2S_ERR_CUDA;
  }
  return B2S_OK;
}

constexpr size_t kTailFlagBytes = 148 * 2 * 8 * 2 * sizeof(int);  // tiles of one round x CTAs x epilogue warps (<= 8) x {ready, count}

// (zeroed on the launching stream, not with a synchronous memset: the first tail-split GEMM of a process may be issued
// inside a stream capture, where a legacy-stream memset is an error)
int ensure_tail_flags(cudaStream_t stream) {
  Context& c = ctx();
  int dev = 0;
  B2S_CUDA_CHECK(cudaGetDevice(&dev));
  if (c.tail_flags != nullptr && c.tail_flags_dev == dev) return B2S_OK;
  if (c.tail_flags != nullptr) cudaFree(c.tail_flags);
  c.tail_flags = nullptr;
  B2S_CUDA_CHECK(cudaMalloc(&c.tail_flags, kTailFlagBytes));
  B2S_CUDA_CHECK(cudaMemsetAsync(c.tail_flags, 0, kTailFlagBytes, stream));
  c.tail_flags_dev = dev;
  return B2S_OK;
}

template <int BN, int CG, int MODE>
int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const KParams& p,
               cudaStream_t stream) {
  using C = Cfg<BN, CG>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, CG, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    B2S_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  const int sms = num_sms();
  int clusters = sms / CG;
  if (clusters > p.units) clusters = p.units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(kThreads<MODE>);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;  // B2S_PDL=0: plain stream serialisation (A/B runs)
  count_launch();
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  Context& cx = ctx();
  if (cx.timing) {
    B2S_CUDA_CHECK(cudaEventCreate(&ev0));
    B2S_CUDA_CHECK(cudaEventCreate(&ev1));
    B2S_CUDA_CHECK(cudaEventRecord(ev0, stream));
  }
  B2S_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tw, to, p));
  if (cx.timing) {
    B2S_CUDA_CHECK(cudaEventRecord(ev1, stream));
    cx.events.emplace_back(ev0, ev1);
    cx.shapes.push_back(cx.pending);
  }
  return B2S_OK;
}

}  // namespace

// 2-D bf16 tensor map with the 128-byte swizzle (shared with attention_tc.cu)
int encode_map_2d_bf16(CUtensorMap* map, const void* base, unsigned long long cols, unsigned long long rows,
                       unsigned long long row_stride_bytes, unsigned box_cols, unsigned box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  return encode_map(map, base, 2, dims, strides, box);
}

// Optional per-launch timing of this kernel (bench.py's roofline leg): CUDA events on the launching stream
// around every GEMM launch while enabled.
void gemm_timing_enable(int on) {
  Context& cx = ctx();
  for (auto& e : cx.events) {
    cudaEventDestroy(e.first);
    cudaEventDestroy(e.second);
  }
  cx.events.clear();
  cx.shapes.clear();
  cx.timing = on != 0;
}

// per-launch record: shape[10] = M, N, K (whole reduction), batches, groups, epilogue, activation, mode, block_n, cta_group
int gemm_timing_get(long long index, double* ms, int* shape) {
  Context& cx = ctx();
  B2S_REQUIRE(index >= 0 && index < static_cast<long long>(cx.events.size()) && ms && shape, "gemm_timing_get: bad index");
  auto& e = cx.events[static_cast<size_t>(index)];
  B2S_CUDA_CHECK(cudaEventSynchronize(e.second));
  float t = 0.f;
  B2S_CUDA_CHECK(cudaEventElapsedTime(&t, e.first, e.second));
  *ms = t;
  const TimedShape& d = cx.shapes[static_cast<size_t>(index)];
  const int v[10] = {d.M, d.N, d.K, d.batches, d.groups, d.epi, d.act, d.mode, d.bn, d.cg};
  for (int i = 0; i < 10; ++i) shape[i] = v[i];
  return B2S_OK;
}

int gemm_timing_read(double* total_ms, long long* launches) {
  Context& cx = ctx();
  double ms = 0.0;
  for (auto& e : cx.events) {
    B2S_CUDA_CHECK(cudaEventSynchronize(e.second));
    float t = 0.f;
    B2S_CUDA_CHECK(cudaEventElapsedTime(&t, e.first, e.second));
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (launches) *launches = static_cast<long long>(cx.events.size());
  return B2S_OK;
}

int gemm_bf16_launch(const GemmArgs& a, cudaStream_t stream) {
  B2S_REQUIRE(a.A && a.W && a.out, "gemm: null pointer");
  B2S_REQUIRE(a.M > 0 && a.N > 0 && a.k_per_tap > 0 && a.taps > 0 && a.batches > 0 && a.groups > 0,
              "gemm: non-positive dimension");
  B2S_REQUIRE(a.N % 8 == 0 && a.ldo % 8 == 0, "gemm: N and ldo must be multiples of 8");
  B2S_REQUIRE(a.a_row_stride % 8 == 0 && a.a_batch_stride % 8 == 0 && a.w_cols % 8 == 0,
              "gemm: strides must be multiples of 8 elements (16 bytes)");
  B2S_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
              "gemm: pointers must be 16-byte aligned");
  B2S_REQUIRE(a.epi >= EPI_BF16 && a.epi <= EPI_ACCUM_F32, "gemm: bad epilogue id %d", a.epi);
  // tcgen05 kind::f16 does encode a_format / b_format separately, but a bf16 x fp16 pair traps with an illegal
  // instruction on B200 (measured, round 2): both operands must share one 16-bit format
  B2S_REQUIRE((a.a_fmt != 0) == (a.w_fmt != 0), "gemm: A and W must share one 16-bit format (bf16 or fp16)");
  const bool mn = a.a_mn || a.b_mn;
  if (mn) {
    B2S_REQUIRE(a.taps == 1 && (a.epi == EPI_BF16 || a.epi == EPI_F32 || a.epi == EPI_ACCUM_F32 || a.epi == EPI_RESID_F32),
                "gemm: MN-major operands support plain epilogues and taps == 1 only");
    B2S_REQUIRE(!a.a_mn || a.batches == 1, "gemm: MN-major A reduces over k_batches; tile batches must be 1");
    B2S_REQUIRE(!a.a_mn || a.b_mn, "gemm: MN-major A needs MN-major W");
    B2S_REQUIRE(a.a_mn || a.k_batches <= 1, "gemm: k_batches needs MN-major A");
  } else {
    B2S_REQUIRE(a.k_batches <= 1 && !a.b_tap_atoms, "gemm: k_batches / b_tap_atoms need MN-major operands");
  }
  if (a.k_splits > 1) B2S_REQUIRE(a.epi == EPI_ACCUM_F32, "gemm: split-K needs the accumulate epilogue");
  if (a.epi == EPI_ACCUM_F32) B2S_REQUIRE(a.bias == nullptr && a.act == ACT_NONE, "gemm: accumulate epilogue is plain");
  if (a.epi == EPI_RESID_F32) B2S_REQUIRE(a.resid != nullptr, "gemm: residual epilogue needs resid");
  if (a.epi == EPI_ROPE) {
    B2S_REQUIRE(a.rope_cs && a.positions && a.N % 128 == 0 && a.rope_cols % 128 == 0,
                "gemm: rope epilogue needs tables and 128-aligned heads");
  }
  if (a.epi == EPI_SWIGLU) B2S_REQUIRE(a.N % 128 == 0, "gemm: swiglu epilogue needs N %% 128 == 0");

  int bn = a.block_n;
  int cg = a.cta_group;
  if (bn == 0) {
    if (a.N <= 64) bn = 64;
    else if (a.N <= 128) bn = 128;
    else bn = 256;
  }
  // measured on B200 (profiles/r01_trip1_kernel_microbench.jsonl): 256x256 tiles on a CTA pair win on every
  // shape of the path (halved W smem traffic per SM); small-N grouped tiles stay single-CTA.
  if (cg == 0) cg = (bn == 256 && a.M > 128) ? 2 : 1;
  B2S_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: block_n must be 64/128/256");
  B2S_REQUIRE(cg == 1 || cg == 2, "gemm: cta_group must be 1 or 2");
  if (a.epi == EPI_ROPE || a.epi == EPI_SWIGLU) B2S_REQUIRE(bn >= 128, "gemm: paired epilogues need block_n >= 128");
  if (a.groups > 1) B2S_REQUIRE(a.N % bn == 0 || a.N < bn, "gemm: grouped N must tile evenly");

  KParams p{};
  p.M = a.M;
  p.N = a.N;
  p.batches = a.batches;
  p.groups = a.groups;
  p.m_tiles = (a.M + 128 * cg - 1) / (128 * cg);
  p.n_tiles = (a.N + bn - 1) / bn;
  const long long total = 1LL * p.m_tiles * p.n_tiles * a.batches * a.groups;
  B2S_REQUIRE(total < (1LL << 31), "gemm: too many tiles");
  p.total_tiles = static_cast<int>(total);
  p.kb_per_tap = (a.k_per_tap + kBlockK - 1) / kBlockK;
  const int k_batches = a.k_batches > 0 ? a.k_batches : 1;
  p.kb_per_batch = p.kb_per_tap;
  p.num_kb = p.kb_per_tap * a.taps * k_batches;
  p.a_mn = a.a_mn ? 1 : 0;
  p.b_mn = a.b_mn ? 1 : 0;
  p.b_tap_atoms = a.b_tap_atoms ? 1 : 0;
  p.og_rows = a.out_group_rows;
  p.og_cols = a.out_group_cols > 0 || a.out_group_rows > 0 ? a.out_group_cols : a.N;
  {
    const int sms = num_sms();
    int splits = a.k_splits;
    if (splits == 0) {  // auto: only the accumulate epilogue can split
      splits = 1;
      if (a.epi == EPI_ACCUM_F32 && total > 0) {
        // fill whole waves of the persistent grid: maximise tiles*s / (waves * clusters), small cost per extra split
        // for the added atomic traffic (measured: profiles/r01_bench_splits.jsonl)
        const long long clusters = sms / cg;
        double best = -1.0;
        for (int sp = 1; sp <= 16 && sp * 4 <= p.num_kb; ++sp) {
          const long long work = total * sp;
          const long long waves = (work + clusters - 1) / clusters;
          const double eff = static_cast<double>(work) / static_cast<double>(waves * clusters) - 0.004 * sp;
          if (eff > best + 1e-9) {
            best = eff;
            splits = sp;
          }
        }
      }
    }
    if (splits > p.num_kb / 4) splits = p.num_kb / 4;
    if (splits < 1) splits = 1;
    p.kb_per_split = (p.num_kb + splits - 1) / splits;
    p.k_splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.tiles_per_split = p.total_tiles;
    B2S_REQUIRE(total * p.k_splits < (1LL << 31), "gemm: too many tiles");
    p.total_tiles = p.total_tiles * p.k_splits;
  }
  p.k_per_tap = a.k_per_tap;
  p.a_pad = a.a_pad;
  p.a_group_off = a.a_group_off;
  p.w_group_off = a.w_group_off;
  p.epi = a.epi;
  p.act = a.act;
  p.bias = a.bias;
  p.out = a.out;
  p.ldo = a.ldo;
  p.out_batch_rows = a.out_batch_rows;
  p.resid = a.resid;
  p.resid_bcast = a.resid_bcast;
  p.out2 = a.out2;
  p.ld2 = a.ld2;
  p.rope_cs = a.rope_cs;
  p.positions = a.positions;
  p.rope_cols = a.rope_cols;
  const int resid_red = ctx().resid_red;
  p.resid_red = resid_red;
  // group_m = 8 M-tiles per group. Swept on the forward step in round 2 (profiles/r02_gemm_tile_order.md): 8 / 16 / 32 and
  // an L2-sized automatic choice are within run-to-run noise of each other (32 slightly slower); larger groups do cut the
  // W re-reads (DRAM traffic ~ A + W * ceil(m_tiles / group_m)) but DRAM sits at 12-28 % of peak on these shapes.
  p.group_m = ctx().gemm_group_m > 0 ? ctx().gemm_group_m : 8;
  p.idesc_fmt = ptx::idesc_formats(a.a_fmt != 0, a.w_fmt != 0);
  p.out_f16 = a.out_fmt != 0 ? 1 : 0;
  p.drop_k1 = a.drop_k1;
  p.drop_k2 = a.drop_k2;
  p.drop_thresh = a.drop_thresh;
  p.drop_inv_keep = a.drop_inv_keep;
  if (a.drop_thresh != 0u) {
    B2S_REQUIRE(a.epi == EPI_BF16 || a.epi == EPI_F32 || a.epi == EPI_RESID_F32, "gemm: dropout needs a plain epilogue");
    B2S_REQUIRE(static_cast<long long>(a.batches) * (a.out_batch_rows > 0 ? a.out_batch_rows : a.M) * a.ldo < (1LL << 32),
                "gemm: dropout element index exceeds 32 bits");
  }

  CUtensorMap ta, tw;
  {
    const int nb = a.a_mn ? k_batches : a.batches;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(a.a_dim0), static_cast<cuuint64_t>(a.a_rows),
                          static_cast<cuuint64_t>(nb)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(a.a_row_stride) * 2,
                             static_cast<cuuint64_t>(nb > 1 ? a.a_batch_stride : a.a_row_stride * a.a_rows) * 2};
    if (strides[1] == 0) strides[1] = strides[0];
    cuuint32_t box[3] = {kBlockK, a.a_mn ? 64u : 128u, 1};
    int rc = encode_map(&ta, a.A, 3, dims, strides, box);
    if (rc != B2S_OK) return rc;
  }
  if (a.b_mn) {
    // MN-major W: [w_rows = reduction rows per k-batch, w_cols] with explicit row / batch strides
    const long long rs = a.w_row_stride > 0 ? a.w_row_stride : a.w_cols;
    B2S_REQUIRE(rs % 8 == 0 && a.w_batch_stride % 8 == 0, "gemm: W strides must be multiples of 8 elements");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(a.w_cols), static_cast<cuuint64_t>(a.w_rows),
                          static_cast<cuuint64_t>(k_batches)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(rs) * 2,
                             static_cast<cuuint64_t>(k_batches > 1 ? a.w_batch_stride : rs * a.w_rows) * 2};
    if (strides[1] == 0) strides[1] = strides[0];
    cuuint32_t box[3] = {kBlockK, 64, 1};
    int rc = encode_map(&tw, a.W, 3, dims, strides, box);
    if (rc != B2S_OK) return rc;
  } else {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(a.w_cols), static_cast<cuuint64_t>(a.w_rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(a.w_cols) * 2};
    cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(bn / cg)};
    int rc = encode_map(&tw, a.W, 2, dims, strides, box);
    if (rc != B2S_OK) return rc;
  }

  bool ext = mn || a.epi == EPI_ACCUM_F32 || p.k_splits > 1 || a.out2 != nullptr || a.out_group_rows != 0 ||
             (a.out_group_cols > 0 && a.out_group_cols != a.N) || a.drop_thresh != 0u;
  // MODE 0 writes its plain outputs through TMA (stores, or fp32 add-reductions for the in-place residual); whatever
  // that cannot express -- a residual read from another tensor, a broadcast residual -- takes the MODE 1 kernel, which
  // keeps the per-thread epilogue. B2S_OPT_TMA_EPILOGUE = 0 (env seed B2S_TMA_EPI) forces the per-thread path everywhere (A/B).
  const int tma_epi = ctx().tma_epi;
  CUtensorMap to = ta;
  p.tma_out = 0;
  {
    // what a tensor-map store / reduction can express: one [batches][M][groups * N] view of the output, no second
    // output, no per-group row offsets, no dropout; the residual only in place (a reduction cannot read another tensor)
    const bool plain = a.epi == EPI_BF16 || a.epi == EPI_F32 || a.epi == EPI_RESID_F32 || a.epi == EPI_ACCUM_F32;
    const bool inplace = a.epi == EPI_RESID_F32 && static_cast<const void*>(a.resid) == a.out && !a.resid_bcast &&
                         resid_red != 0;
    // MODE 0 (K-major, unsplit) and the dgrad form (MN-major B, plain store). The fp32 accumulation of wgrad / split-K
    // stays on red.global: as a TMA reduction it measured 4 % slower on the 15968-row wgrads (round 2).
    const bool kernel_ok = !ext || (a.b_mn && !a.a_mn && p.k_splits == 1 && (a.epi == EPI_BF16 || a.epi == EPI_F32));
    const bool view_ok = a.out2 == nullptr && a.out_group_rows == 0 &&
                         (a.out_group_cols == 0 || a.out_group_cols == a.N) && a.drop_thresh == 0u &&
                         a.epi != EPI_ACCUM_F32;
    if (plain && kernel_ok && view_ok && tma_epi != 0 && (a.epi != EPI_RESID_F32 || inplace)) {
      const bool h16 = a.epi == EPI_BF16;
      const cuuint64_t elt = h16 ? 2 : 4;
      const cuuint64_t cols = a.groups > 1 ? static_cast<cuuint64_t>(a.groups) * a.N : static_cast<cuuint64_t>(a.N);
      const cuuint64_t row_bytes = static_cast<cuuint64_t>(a.ldo) * elt;
      const cuuint64_t batch_rows = a.out_batch_rows > 0 ? a.out_batch_rows : a.M;
      cuuint64_t dims[3] = {cols, static_cast<cuuint64_t>(a.M), static_cast<cuuint64_t>(a.batches)};
      cuuint64_t strides[2] = {row_bytes, row_bytes * batch_rows};
      cuuint32_t box[3] = {h16 ? 64u : 32u, 32u, 1u};
      const CUtensorMapDataType dt = h16 ? (a.out_fmt != 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)
                                         : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
      if (row_bytes % 16 == 0 && encode_map(&to, a.out, 3, dims, strides, box, dt) == B2S_OK) p.tma_out = 1;
    }
    // MODE 0 has no per-thread plain epilogue: a K-major, unsplit problem that cannot use TMA runs as MODE 1
    if (!ext && !p.tma_out && plain) ext = true;
  }
  int mode = a.a_mn ? 3 : (a.b_mn ? (p.tma_out ? 4 : 2) : (ext ? 1 : 0));
  // short reductions with an ACTIVATION in a plain epilogue are bound by the epilogue's arithmetic (16 k-blocks of MMA
  // against 256 erf-GELUs per epilogue thread): eight epilogue warps (MODE 5) hide its latencies better -- FFN1 + GELU
  // 756 -> 805 TFLOP/s inside the step. Without an activation the same eight warps LOSE 2-9 % (QKV 928 -> 909, out-proj
  // 599 -> 574, conv 1150 -> 1046: more barrier traffic, one staging tile per warp), so they stay on four.
  // B2S_OPT_GEMM_EPI8: 0 = never, 1 = activation and K <= 2048 (default), 2 = every plain MODE 0 launch (tests)
  if (mode == 0 && (a.epi == EPI_BF16 || a.epi == EPI_F32 || a.epi == EPI_RESID_F32)) {
    const int e8 = ctx().gemm_epi8;
    if (e8 == 2 || (e8 == 1 && p.num_kb <= 32 && a.act != ACT_NONE)) mode = 5;
  }
  // Tail split (MODE 0, fp32 outputs that leave through TMA): see KParams. Only the last, partly filled round is cut, into
  // at most 4 K-slices of at least 8 k-blocks, one slice per cluster, so the round costs 1/S of a tile (+ one epilogue).
  p.units = p.total_tiles;
  p.tail_first = p.total_tiles;
  p.tail_split = 1;
  p.tail_kb = p.num_kb;
  p.tail_flags = nullptr;
  if ((mode == 0 || mode == 5) && p.tma_out && ctx().gemm_tail_split && a.act == ACT_NONE && (a.epi == EPI_F32 || a.epi == EPI_RESID_F32)) {
    const int G = num_sms() / cg;
    const int tail = G > 0 ? p.total_tiles % G : 0;
    const int full = p.total_tiles - tail;
    if (tail > 0 && full > 0) {
      int S = G / tail;
      if (S > 4) S = 4;
      if (S > p.num_kb / 8) S = p.num_kb / 8;
      if (S >= 2) {
        const int rc = ensure_tail_flags(stream);
        if (rc != B2S_OK) return rc;
        p.tail_kb = (p.num_kb + S - 1) / S;
        p.tail_split = (p.num_kb + p.tail_kb - 1) / p.tail_kb;
        p.tail_first = full;
        p.units = full + tail * p.tail_split;
        p.tail_flags = ctx().tail_flags;
      }
    }
  }
  if (ctx().timing)
    ctx().pending = TimedShape{a.M, a.N, a.k_per_tap * a.taps * k_batches, a.batches, a.groups, a.epi, a.act, mode, bn, cg};
#define B2S_GEMM_CASE(BN_, CG_)                                          \
  if (bn == BN_ && cg == CG_) {                                          \
    switch (mode) {                                                      \
      case 0: return launch_cfg<BN_, CG_, 0>(ta, tw, to, p, stream);     \
      case 1: return launch_cfg<BN_, CG_, 1>(ta, tw, to, p, stream);     \
      case 2: return launch_cfg<BN_, CG_, 2>(ta, tw, to, p, stream);     \
      case 4: return launch_cfg<BN_, CG_, 4>(ta, tw, to, p, stream);     \
      case 5: return launch_cfg<BN_, CG_, 5>(ta, tw, to, p, stream);     \
      default: return launch_cfg<BN_, CG_, 3>(ta, tw, to, p, stream);    \
    }                                                                    \
  }
  B2S_GEMM_CASE(256, 1);
  B2S_GEMM_CASE(128, 1);
  B2S_GEMM_CASE(64, 1);
  B2S_GEMM_CASE(256, 2);
  B2S_GEMM_CASE(128, 2);
#undef B2S_GEMM_CASE
  set_last_error("gemm: unsupported (block_n=%d, cta_group=%d)", bn, cg);
  return B2S_ERR_UNSUPPORTED;
}

}  // namespace b2s
