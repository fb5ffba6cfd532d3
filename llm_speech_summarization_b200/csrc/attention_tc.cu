// attention_tc.cu -- packed variable-length flash attention forward on tcgen05 / TMEM / TMA (sm_100a).
//
// Work item = (sequence, query head, 128-query block); persistent CTAs (as many as stay resident) walk the items.
// 5 warps per CTA:
//   warps 0-3 : softmax. Thread r owns query row r (TMEM lane r): it reads its row of S straight from TMEM
//               (no cross-thread reductions) ONCE per block, masks only in partial / diagonal blocks, keeps the
//               online-softmax state (reference max m, running sum l) in registers, writes P = exp2(S - m) as bf16
//               into a 128B-swizzled K-major shared-memory tile, and moves the reference (rescaling O in TMEM) only
//               when the row max outgrew it by 2^16 (lazy, deferred, warp-uniform).
//   warp 4    : lane 0 issues TMA loads (Q per item, K/V per step, 1 or 2 stages) and all tcgen05.mma:
//               S[128 x BN keys] = Q K^T   (A = Q smem K-major, B = K smem K-major, fp32 accumulators in TMEM cols [0,BN))
//               O[128 x D]      += P V      (A = P smem K-major, B = V smem MN-major, accumulators in TMEM cols [BN,BN+D))
//               and runs ahead over item boundaries (next item's Q / first K/V in flight during the output store).
// BN = 64 keys per step with single-buffered K/V is the default: 48 KiB + 128 TMEM columns per CTA at D = 64, so four
// CTAs (16 softmax warps) share an SM -- the kernel is bound by the per-step latency chain, not by the tensor pipe
// (profiles/r01_attention_configs.md).
// Hand-offs are mbarriers (tcgen05.commit -> softmax, softmax -> MMA issuer); every wait is bounded.
//
// Covers HuBERT / Whisper (non-causal, D=64) and Llama / MiniChat (causal GQA, D=128)
// (TF/models/hubert/modeling_hubert.py:262-345, TF/models/llama/modeling_llama.py:225-289). Operands (Q, K, V, P, O)
// are bf16 or fp16 (template F16): same tcgen05 kind::f16 rate, 3 more mantissa bits for fp16.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>

#include "b2s_common.cuh"
#include "b2s_ptx.cuh"
#include "ops.cuh"

namespace b2s {

int encode_map_2d_bf16(CUtensorMap* map, const void* base, unsigned long long cols, unsigned long long rows,
                       unsigned long long row_stride_bytes, unsigned box_cols, unsigned box_rows);  // gemm_sm100.cu

namespace {

constexpr int kBM = 128;   // queries per CTA
constexpr int kThreadsTc = 160;

// BN = keys per step (128, or 64 for short sequences at D = 128 so that two CTAs fit on an SM)
// KVS = 3 selects the PIPELINED form: S double-buffered in TMEM (the issuer runs two key blocks ahead of the softmax
// warps, so QK^T of block g+1 / g+2 overlaps the softmax of block g), K and V in two stages each with their own
// barriers, P single-buffered (rewritten once P.V of the previous block has retired).
template <int D, int BN, int KVS>
struct AttnCfg {
  static constexpr bool kPipe = KVS == 3;
  static constexpr int kQBytes = kBM * D * 2;
  static constexpr int kKBytes = BN * D * 2;
  static constexpr int kVBytes = BN * D * 2;
  static constexpr int kPBytes = kBM * BN * 2;
  // D = 64: single-buffered K/V keeps the CTA at ~81 KiB so two CTAs (2 x 256 TMEM columns) share an SM and overlap
  // each other's TMA / MMA / softmax phases; D = 128: one CTA per SM, K/V double-buffered inside it.
  static constexpr int kKvStages = kPipe ? 2 : KVS;  // K/V tiles in flight: 1 keeps the CTA small (more CTAs per SM), 2 prefetches
  // the dynamic shared-memory window is declared 1024-byte aligned (128B-swizzle atoms), so no alignment slack
  static constexpr int kSmemBytes = kQBytes + kKvStages * (kKBytes + kVBytes) + kPBytes + 256;
  static constexpr int kOCol = kPipe ? 2 * BN : BN;               // S: [0, BN) (+ [BN, 2 BN) pipelined), then O: D columns
  static constexpr int kTmemCols = (kOCol + D) <= 128 ? 128 : 256;  // power of two >= kOCol + D
  static_assert(kOCol + D <= 256, "TMEM budget: two CTAs per SM");
};

struct AttnParams {
  const int* cu;
  __nv_bfloat16* o;
  long long ldo;
  int Hq, Hkv;
  int q_col0, k_col0, v_col0;  // first column of head 0 inside each tensor map
  float scale_log2;
  int causal;
  float* lse;  // optional [rows, Hq]: log2-domain log-sum-exp of the scaled scores (saved for the backward)
  AttnDrop drop;  // attention-probability dropout (thresh 0 = off): softmax statistics stay those of the un-dropped row
  int n_qblk;       // 128-query blocks per (sequence, head) = ceil(max_seqlen / 128)
  int total_items;  // num_seqs * Hq * n_qblk
  // Shared prefix (the prompt prefix is the same token run in every sequence of a step, so under a causal mask its hidden
  // states are identical everywhere): sequence 0 holds the prefix rows ONCE; every other sequence holds only its own
  // rows and sees the prefix_len keys / values of sequence 0 as one extra, leading key block. 0 = off.
  int prefix_len;
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One work item = (sequence, query head, 128-query block). The kernel is PERSISTENT: a CTA walks items
// blockIdx.x, blockIdx.x + gridDim.x, ... with its TMEM columns, barriers and tensor-map prefetches set up once, and
// the issuing thread runs ahead across item boundaries (Q and the first K/V tile of the next item are in flight while
// the softmax warps store the current item's output), so the per-item fixed latencies that dominated the one-CTA-per-
// item version at prompt lengths of a few hundred tokens are amortised. All barrier phases are tracked with running
// counters (g = blocks processed by this CTA so far, items done so far).
struct AttnItem {
  int seq, h, hk, s0, L, q0, nblk;
  int pre;  // 1 = key block 0 of this item is the shared prefix (rows [cu[0], cu[0] + prefix_len)), own keys follow
};

template <int BN>
__device__ __forceinline__ bool attn_decode_item(const AttnParams& p, int item, AttnItem* it) {
  const int qb = item % p.n_qblk;
  const int r = item / p.n_qblk;
  it->h = r % p.Hq;
  it->seq = r / p.Hq;
  it->s0 = p.cu[it->seq];
  it->L = p.cu[it->seq + 1] - it->s0;
  it->q0 = qb * kBM;
  if (it->q0 >= it->L) return false;
  it->hk = it->h / (p.Hq / p.Hkv);
  const int kv_len = p.causal ? min(it->L, it->q0 + kBM) : it->L;
  it->pre = (p.prefix_len > 0 && it->seq > 0) ? 1 : 0;
  it->nblk = it->pre + (kv_len + BN - 1) / BN;
  return true;
}

// first valid item at or after `item` (stride gridDim.x); returns total_items when there is none
template <int BN>
__device__ __forceinline__ int attn_next_item(const AttnParams& p, int item, AttnItem* it) {
  for (; item < p.total_items; item += gridDim.x)
    if (attn_decode_item<BN>(p, item, it)) return item;
  return p.total_items;
}

template <int D, int BN, int KVS, bool DROP, bool F16>
__global__ void __launch_bounds__(kThreadsTc, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  using C = AttnCfg<D, BN, KVS>;
  constexpr int kBN = BN;
  constexpr int kAtoms = D / 64;  // 64-element (128 B) column atoms per row
  pdl_trigger();  // PDL: the out-projection GEMM that follows may take the SMs this grid frees (b2s_common.cuh)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();  // the swizzled tiles need 1024-byte alignment
  const uint32_t sQ = base;
  constexpr int kKvStages = C::kKvStages;
  const uint32_t sK = sQ + C::kQBytes;
  const uint32_t sV = sK + kKvStages * C::kKBytes;
  const uint32_t sP = sV + kKvStages * C::kVBytes;
  const uint32_t bars = sP + C::kPBytes;
  const uint32_t bar_q = bars, bar_kv0 = bars + 8, bar_s = bars + 24, bar_p = bars + 32, bar_o = bars + 40;
  const uint32_t bar_oe = bars + 48;  // softmax warps -> issuer: the item's O accumulator has been read out
  const uint32_t tmem_slot = bars + 56;
  // pipelined form: K / V stages have separate barriers (bar_kv0 + 8 s = K stage s, bar_v0 + 8 s = V stage s) and there
  // is one S barrier per TMEM buffer (bar_s, bar_s1)
  constexpr bool kPipe = C::kPipe;
  const uint32_t bar_v0 = bars + 64, bar_s1 = bars + 80;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_kv0, 1);
    ptx::mbar_init(bar_kv0 + 8, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, kBM);
    ptx::mbar_init(bar_o, 1);
    ptx::mbar_init(bar_oe, kBM);
    ptx::mbar_init(bar_v0, 1);
    ptx::mbar_init(bar_v0 + 8, 1);
    ptx::mbar_init(bar_s1, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) ptx::tmem_alloc<1>(tmem_slot, C::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // launched with PDL: barriers / TMEM are set up while the QKV GEMM drains; its output is complete from here

  if (warp == 4) {
    if (lane == 0) {
      // ---------------- TMA + MMA issuer ----------------
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_k);
      ptx::prefetch_tmap(&tmap_v);
      // first row of key block j of an item (block 0 of a prefixed item = the shared prefix rows of sequence 0)
      auto kv_row = [&](const AttnItem& it, int j) { return j < it.pre ? p.cu[0] : it.s0 + (j - it.pre) * kBN; };
      auto load_q = [&](const AttnItem& it) {
        ptx::mbar_arrive_expect_tx(bar_q, C::kQBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a)
          ptx::tma_load_2d(&tmap_q, bar_q, sQ + a * (kBM * 128), p.q_col0 + it.h * D + a * 64, it.s0 + it.q0);
      };
      auto load_kv = [&](const AttnItem& it, int j, uint32_t g) {  // key block j of item `it` = CTA-global block g
        const int st = g % kKvStages;
        ptx::mbar_arrive_expect_tx(bar_kv0 + 8 * st, C::kKBytes + C::kVBytes);
#pragma unroll
        for (int a = 0; a < kAtoms; ++a) {
          ptx::tma_load_2d(&tmap_k, bar_kv0 + 8 * st, sK + st * C::kKBytes + a * (kBN * 128),
                           p.k_col0 + it.hk * D + a * 64, kv_row(it, j));
          ptx::tma_load_2d(&tmap_v, bar_kv0 + 8 * st, sV + st * C::kVBytes + a * (kBN * 128),
                           p.v_col0 + it.hk * D + a * 64, kv_row(it, j));
        }
      };
      constexpr uint32_t idesc_s = ptx::make_idesc_f32acc(kBM, kBN) | ptx::idesc_formats(F16, F16);
      constexpr uint32_t idesc_o =
          ptx::make_idesc_f32acc(kBM, D) | ptx::idesc_formats(F16, F16) | (1u << 16);  // B (= V) is MN-major
      if constexpr (kPipe) {
        // ---- pipelined issue order. Blocks are numbered CTA-wide (g = 0, 1, ...) across item boundaries; block g uses
        // S buffer g & 1 and K / V stage g & 1. Iteration i: K(i+2) is requested as soon as S(i) has retired, P.V(i) is
        // issued when the softmax warps publish P(i), then S(i+2) (its buffer was read out for P(i)), then V(i+2) once
        // P.V(i) has retired. The softmax warps therefore always find S(i+1) waiting when they finish block i.
        struct Cur {
          AttnItem it;
          int item, j;
        };
        auto cur_valid = [&](const Cur& c) { return c.item < p.total_items; };
        auto cur_next = [&](Cur& c) {
          if (++c.j >= c.it.nblk) {
            c.item = attn_next_item<BN>(p, c.item + gridDim.x, &c.it);
            c.j = 0;
          }
        };
        auto load_k = [&](const Cur& c, uint32_t g) {
          const int st = g & 1;
          ptx::mbar_arrive_expect_tx(bar_kv0 + 8 * st, C::kKBytes);
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
            ptx::tma_load_2d(&tmap_k, bar_kv0 + 8 * st, sK + st * C::kKBytes + a * (kBN * 128),
                             p.k_col0 + c.it.hk * D + a * 64, kv_row(c.it, c.j));
        };
        auto load_v = [&](const Cur& c, uint32_t g) {
          const int st = g & 1;
          ptx::mbar_arrive_expect_tx(bar_v0 + 8 * st, C::kVBytes);
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
            ptx::tma_load_2d(&tmap_v, bar_v0 + 8 * st, sV + st * C::kVBytes + a * (kBN * 128),
                             p.v_col0 + c.it.hk * D + a * 64, kv_row(c.it, c.j));
        };
        uint32_t q_loads = 0;
        auto issue_s = [&](const Cur& c, uint32_t g) {
          const int st = g & 1;
          if (c.j == 0) {  // first block of an item: its Q rows may replace the previous item's once that item's last
                           // S = S(g-1) has retired
            if (g > 0) ptx::mbar_wait(((g - 1) & 1) ? bar_s1 : bar_s, ((g - 1) >> 1) & 1);
            load_q(c.it);
            ptx::mbar_wait(bar_q, q_loads & 1);
            ++q_loads;
          }
          ptx::mbar_wait(bar_kv0 + 8 * st, (g >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t q_off = (k >> 2) * (kBM * 128) + (k & 3) * 32;
            const uint32_t k_off = (k >> 2) * (kBN * 128) + (k & 3) * 32;
            ptx::umma_bf16<1>(tmem_base + st * kBN, ptx::make_kmajor_sw128_desc(sQ + q_off),
                              ptx::make_kmajor_sw128_desc(sK + st * C::kKBytes + k_off), idesc_s, k > 0 ? 1u : 0u);
          }
          ptx::umma_commit<1>(st ? bar_s1 : bar_s);
        };
        Cur c0, c2;
        c0.item = attn_next_item<BN>(p, blockIdx.x, &c0.it);
        c0.j = 0;
        if (cur_valid(c0)) {
          c2 = c0;
          load_k(c2, 0);
          load_v(c2, 0);
          issue_s(c2, 0);
          cur_next(c2);
          if (cur_valid(c2)) {
            load_k(c2, 1);
            load_v(c2, 1);
            issue_s(c2, 1);
            cur_next(c2);
          }
          uint32_t i = 0, items_done = 0;
          while (cur_valid(c0)) {
            const int st = i & 1;
            const bool have2 = cur_valid(c2);
            if (have2) {
              ptx::mbar_wait(st ? bar_s1 : bar_s, (i >> 1) & 1);  // S(i) has retired: its K stage is free
              load_k(c2, i + 2);
            }
            ptx::mbar_wait(bar_p, i & 1);  // P(i) published, S(i) read out
            if (c0.j == 0 && items_done > 0) ptx::mbar_wait(bar_oe, (items_done - 1) & 1);  // previous O read out
            ptx::mbar_wait(bar_v0 + 8 * st, (i >> 1) & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < kBN / 16; ++k) {
              const uint64_t pdesc = ptx::make_kmajor_sw128_desc(sP + (k >> 2) * (kBM * 128) + (k & 3) * 32);
              const uint64_t vdesc = ptx::make_mnmajor_sw128_desc(sV + st * C::kVBytes + k * 2048, kBN * 128);
              ptx::umma_bf16<1>(tmem_base + C::kOCol, pdesc, vdesc, idesc_o, (c0.j > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit<1>(bar_o);
            if (have2) {
              issue_s(c2, i + 2);
              ptx::mbar_wait(bar_o, i & 1);  // P.V(i) has retired: its V stage is free
              load_v(c2, i + 2);
              cur_next(c2);
            }
            if (c0.j + 1 == c0.it.nblk) ++items_done;
            cur_next(c0);
            ++i;
          }
        }
      } else {
        AttnItem cur, nxt;
        int item = attn_next_item<BN>(p, blockIdx.x, &cur);
        uint32_t g = 0, items_done = 0;
        if (item < p.total_items) {
          load_q(cur);
          load_kv(cur, 0, 0);
        }
        while (item < p.total_items) {
          const int item_nxt = attn_next_item<BN>(p, item + gridDim.x, &nxt);
          const bool have_nxt = item_nxt < p.total_items;
          ptx::mbar_wait(bar_q, items_done & 1);
          for (int j = 0; j < cur.nblk; ++j, ++g) {
            const int st = g % kKvStages;
            const bool last = j + 1 == cur.nblk;
            const bool has_next = !last || have_nxt;
            if (kKvStages == 2 && has_next) {
              // stage st^1 was last read by PV(g-1): refill it only after that MMA has retired
              if (g >= 1) ptx::mbar_wait(bar_o, (g - 1) & 1);
              if (!last) load_kv(cur, j + 1, g + 1);
              else load_kv(nxt, 0, g + 1);
            }
            ptx::mbar_wait(bar_kv0 + 8 * st, (g / kKvStages) & 1);
            ptx::tc_fence_after();
            // S = Q K^T
  #pragma unroll
            for (int k = 0; k < D / 16; ++k) {
              const uint32_t q_off = (k >> 2) * (kBM * 128) + (k & 3) * 32;  // 64-column atoms are kBM rows tall
              const uint32_t k_off = (k >> 2) * (kBN * 128) + (k & 3) * 32;  // ... and kBN rows tall for K
              ptx::umma_bf16<1>(tmem_base, ptx::make_kmajor_sw128_desc(sQ + q_off),
                                ptx::make_kmajor_sw128_desc(sK + st * C::kKBytes + k_off), idesc_s, k > 0 ? 1u : 0u);
            }
            ptx::umma_commit<1>(bar_s);
            // O += P V once the softmax warps have published P (and finished reading S / rescaling O)
            ptx::mbar_wait(bar_p, g & 1);
            if (j == 0 && items_done > 0) ptx::mbar_wait(bar_oe, (items_done - 1) & 1);  // previous O read out
            ptx::tc_fence_after();
  #pragma unroll
            for (int k = 0; k < kBN / 16; ++k) {
              const uint64_t pdesc = ptx::make_kmajor_sw128_desc(sP + (k >> 2) * (kBM * 128) + (k & 3) * 32);
              const uint64_t vdesc = ptx::make_mnmajor_sw128_desc(sV + st * C::kVBytes + k * 2048, kBN * 128);
              ptx::umma_bf16<1>(tmem_base + C::kOCol, pdesc, vdesc, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
            }
            ptx::umma_commit<1>(bar_o);
            // bar_p(g) has fired, so every S MMA of this item has retired: Q's tile can take the next item's rows
            if (last && have_nxt) load_q(nxt);
            if (kKvStages == 1 && has_next) {  // single buffer: refill once PV(g) has retired
              ptx::mbar_wait(bar_o, g & 1);
              if (!last) load_kv(cur, j + 1, g + 1);
              else load_kv(nxt, 0, g + 1);
            }
            // S(g+1) may now overwrite S (the softmax warps are done with it); P's smem and O are protected because
            // bar_s(g+1) -- a tcgen05.commit -- only fires after every earlier MMA of this thread, PV(g) included.
          }
          cur = nxt;
          item = item_nxt;
          ++items_done;
        }
      }
    }
  } else {
    // ---------------- softmax warps: thread = query row ----------------
    const int r = warp * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t tS = tmem_base + lane_base;
    const uint32_t tO = tmem_base + lane_base + C::kOCol;
    const float sc = p.scale_log2;  // > 0
    AttnItem cur;
    uint32_t g = 0;
    for (int item = attn_next_item<BN>(p, blockIdx.x, &cur); item < p.total_items;
         item = attn_next_item<BN>(p, item + gridDim.x, &cur)) {
    const int seq = cur.seq, h = cur.h, s0 = cur.s0, L = cur.L, q0 = cur.q0, nblk = cur.nblk;
    const int qi = q0 + r;  // sequence-local query index
    const int row_limit = p.causal ? min(L, qi + 1) : L;  // this row sees keys [0, row_limit)
    float m_ref = -INFINITY;  // log2-domain reference the accumulated O / l are expressed against
    float pend = -INFINITY;   // larger reference to adopt at the next safe point (no PV in flight)
    float l_run = 0.f;
    uint32_t dk1 = 0u, dk2 = 0u;  // dropout stream of this (sequence, head); element = (query << 16) | key
    if (DROP) rng_stream_key(p.drop.seed, p.drop.site, static_cast<uint32_t>(seq), static_cast<uint32_t>(h), &dk1, &dk2);
    const uint32_t drow = static_cast<uint32_t>(qi) << 16;

    // adopt `pend` as the new reference where it is larger: rescale l (registers) and O (TMEM); warp-uniform
    auto adopt = [&](bool o_valid) {
      const bool need = pend > m_ref;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ((m_ref == -INFINITY) ? 0.f : ex2f(m_ref - pend)) : 1.0f;
        if (o_valid) {
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tO + c * 32, raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
            tmem_st_32x32(tO + c * 32, raw);
          }
          tmem_st_wait();
        }
        if (need) {
          l_run *= alpha;
          m_ref = pend;
        }
      }
    };

    for (int j = 0; j < nblk; ++j, ++g) {
      // S(g) is in TMEM. Unpipelined: every earlier MMA (PV(g-1) included) has retired with it. Pipelined: S(g) was
      // issued two blocks ahead, so P.V(g-1) may still be running -- it is waited for (once per block, prev_pv) before
      // P's shared-memory tile is rewritten and before O is rescaled.
      if (kPipe) ptx::mbar_wait((g & 1) ? bar_s1 : bar_s, (g >> 1) & 1);
      else ptx::mbar_wait(bar_s, g & 1);
      ptx::tc_fence_after();
      const uint32_t tSg = tS + (kPipe ? (g & 1u) * kBN : 0u);
      bool pv_synced = !kPipe || g == 0;
      auto prev_pv = [&]() {
        if (!pv_synced) {
          ptx::mbar_wait(bar_o, (g - 1) & 1);
          ptx::tc_fence_after();
          pv_synced = true;
        }
      };
      // visible keys of this row inside this block (<= 0 .. >= 128); the shared-prefix block shows all its keys to every row
      const int nvis = (j < cur.pre) ? p.prefix_len : row_limit - (j - cur.pre) * kBN;
      const bool full_blk = __all_sync(0xffffffffu, nvis >= kBN);
      // D = 128 with 64-key steps (two CTAs per SM, registers to spare): the thread's whole row of S (64 fp32) is read
      // from TMEM ONCE, both 32-column loads in flight before the single wait, and stays in registers for the
      // first-block max pass, the main pass and a retry. Elsewhere the row is re-read chunk by chunk: at D = 64 the
      // four-CTAs-per-SM residency needs <= 102 registers per thread.
      constexpr bool kHold = (kBN == 64 && (D == 128 || kPipe));  // (pipelined: two CTAs per SM at either D)
      uint32_t held[kHold ? kBN : 1];
      if (kHold) {
        ptx::tmem_ld_32x32(tSg, *reinterpret_cast<uint32_t(*)[32]>(&held[0]));
        ptx::tmem_ld_32x32(tSg + 32, *reinterpret_cast<uint32_t(*)[32]>(&held[kHold ? 32 : 0]));
        ptx::tmem_ld_wait();
        if (!full_blk) {  // (register-held row) invisible scores become -inf once per block, in place
#pragma unroll
          for (int i = 0; i < kBN; ++i) held[kHold ? i : 0] = (i < nvis) ? held[kHold ? i : 0] : 0xff800000u;
        }
      }
      if (__builtin_expect(j == 0, 0)) {
        // first block: exact masked row max as the initial reference (nothing accumulated yet). (Kept behind a real
        // branch: if-converted, ptxas ran these 64 compare + select + max per row in EVERY block -- ncu, round 2.)
        float bm = -INFINITY;
        if (kHold) {
          asm volatile("" ::: "memory");
#pragma unroll
          for (int i = 0; i < kBN; ++i) bm = fmaxf(bm, __uint_as_float(held[kHold ? i : 0]));
        } else {
#pragma unroll 1
          for (int c = 0; c < kBN / 32; ++c) {
            uint32_t raw[32];
            ptx::tmem_ld_32x32(tSg + c * 32, raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) bm = fmaxf(bm, (c * 32 + i < nvis) ? __uint_as_float(raw[i]) : -INFINITY);
          }
        }
        pend = bm * sc;
      }
      // Single pass per block against the (possibly stale) reference: P = exp2(s*sc - m_ref) may exceed 1, which
      // fp32 / bf16 absorb; the reference is only moved when the block max outgrew it by 2^16 (deferred to the
      // next block, when no PV is in flight) or, to rule out overflow, immediately by re-running the block (> 2^60).
      for (int attempt = 0; attempt < 2; ++attempt) {
        if (kPipe && j > 0 && __any_sync(0xffffffffu, pend > m_ref)) prev_pv();  // O is about to be rescaled
        adopt(j > 0);
        const float mr = (m_ref == -INFINITY) ? 0.f : m_ref;
        float rs = 0.f, bmax = -INFINITY;
#pragma unroll(kHold ? 2 : 1)
        for (int c = 0; c < kBN / 32; ++c) {
          uint32_t raw_c[kHold ? 1 : 32];
          if (!kHold) {
            ptx::tmem_ld_32x32(tSg + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&raw_c[0]));
            ptx::tmem_ld_wait();
          }
          const uint32_t* raw = kHold ? &held[kHold ? c * 32 : 0] : &raw_c[0];
          uint32_t pk[16];
          // ONE code path for full and partial blocks: a partial block first overwrites its invisible scores with -inf
          // (exp2 -> 0, ignored by the max) behind a warp-uniform branch. (Two separate loops were merged by the compiler
          // into one that evaluated the 64 compare + select + max of the masked form in EVERY block -- ncu, round 2.)
          if (!kHold && !full_blk) {
            const int nv = nvis - c * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) raw_c[kHold ? 0 : i] = (i < nv) ? raw_c[kHold ? 0 : i] : 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float x0 = __uint_as_float(raw[i]), x1 = __uint_as_float(raw[i + 1]);
            bmax = fmaxf(bmax, fmaxf(x0, x1));
            float p0 = ex2f(fmaf(x0, sc, -mr)), p1 = ex2f(fmaf(x1, sc, -mr));
            rs += p0 + p1;
            if (DROP) {  // the row sum keeps every key; only the P fed to P.V is thinned (1/(1-p) folded into 1/l)
              const uint32_t e = drow | static_cast<uint32_t>(j * kBN + c * 32 + i);
              p0 = rng_keep(e, dk1, dk2, p.drop.thresh) ? p0 : 0.f;
              p1 = rng_keep(e + 1u, dk1, dk2, p.drop.thresh) ? p1 : 0.f;
            }
            pk[i >> 1] = F16 ? pack_f16(p0, p1) : pack_bf16(p0, p1);
          }
          // chunk c covers keys [32c, 32c+32) = 64 bytes = four 16-byte chunks of atom (c >> 1)
          prev_pv();  // (pipelined) P.V(g-1) still reads this tile until it retires
          const uint32_t atom = sP + (c >> 1) * (kBM * 128) + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = (c & 1) * 4 + q;
            const uint32_t addr = atom + ((chunk ^ (r & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                         "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                         : "memory");
          }
        }
        const float bm = bmax * sc;
        // P is stored in the operand format: bf16 shares fp32's exponent range (redo only past 2^60); fp16 tops out at
        // 65504, so a block whose max outgrew the reference by 2^14 is redone at once and 2^8 already moves it
        constexpr float kRedo = F16 ? 14.0f : 60.0f, kDefer = F16 ? 8.0f : 16.0f;
        const bool overflow_risk = bm > m_ref + kRedo;
        if (attempt == 0 && __any_sync(0xffffffffu, overflow_risk)) {
          if (overflow_risk) pend = bm;
          continue;  // adopt now and redo this block (P in smem is simply rewritten; nothing has consumed it)
        }
        l_run += rs;
        if (bm > m_ref + kDefer) pend = bm;
        break;
      }
      // publish P to the async proxy (UMMA reads smem) and hand S / O back to the MMA issuer
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_p);
    }
    // epilogue: O / l -> bf16 -> global (one row per thread)
    ptx::mbar_wait(bar_o, (g - 1) & 1);
    ptx::tc_fence_after();
    const float inv = (l_run > 0.f ? 1.0f / l_run : 0.f) * (DROP ? p.drop.inv_keep : 1.0f);
    if (p.lse != nullptr && qi < L) p.lse[static_cast<long long>(s0 + qi) * p.Hq + h] = m_ref + log2f(l_run);
    __nv_bfloat16* orow = p.o + static_cast<long long>(s0 + qi) * p.ldo + h * D;
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      uint32_t raw[32];
      ptx::tmem_ld_32x32(tO + c * 32, raw);
      ptx::tmem_ld_wait();
      if (c + 1 == D / 32) {  // the accumulator is in registers: the issuer may start the next item's P.V
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_oe);
      }
      if (qi < L) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_h16(__uint_as_float(raw[8 * q + 0]) * inv, __uint_as_float(raw[8 * q + 1]) * inv, F16);
          u.y = pack_h16(__uint_as_float(raw[8 * q + 2]) * inv, __uint_as_float(raw[8 * q + 3]) * inv, F16);
          u.z = pack_h16(__uint_as_float(raw[8 * q + 4]) * inv, __uint_as_float(raw[8 * q + 5]) * inv, F16);
          u.w = pack_h16(__uint_as_float(raw[8 * q + 6]) * inv, __uint_as_float(raw[8 * q + 7]) * inv, F16);
          *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = u;
        }
      }
    }
    }  // items
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<1>(tmem_base, C::kTmemCols);
  }
}

template <int D, int BN, int KVS, bool DROP, bool F16>
int launch_attn_tc(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, const int* cu,
                   int num_seqs, int max_seqlen, long long total_rows, int Hq, int Hkv, float scale, int causal,
                   float* lse, cudaStream_t stream, const AttnDrop* drop, int prefix_len) {
  using C = AttnCfg<D, BN, KVS>;
  constexpr int kBN = BN;
  auto kern = attn_fwd_tc_kernel<D, BN, KVS, DROP, F16>;
  static bool attr_set = false;
  if (!attr_set) {
    B2S_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  int rc = encode_map_2d_bf16(&tq, q, static_cast<unsigned long long>(Hq) * D, total_rows, ld * 2, 64, kBM);
  if (rc != B2S_OK) return rc;
  rc = encode_map_2d_bf16(&tk, k, static_cast<unsigned long long>(Hkv) * D, total_rows, ld * 2, 64, kBN);
  if (rc != B2S_OK) return rc;
  rc = encode_map_2d_bf16(&tv, v, static_cast<unsigned long long>(Hkv) * D, total_rows, ld * 2, 64, kBN);
  if (rc != B2S_OK) return rc;
  AttnParams p{};
  p.cu = cu;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.Hq = Hq;
  p.Hkv = Hkv;
  p.q_col0 = p.k_col0 = p.v_col0 = 0;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.lse = lse;
  if (DROP) p.drop = *drop;
  p.prefix_len = prefix_len;
  p.n_qblk = (max_seqlen + kBM - 1) / kBM;
  const long long total = static_cast<long long>(p.n_qblk) * Hq * num_seqs;
  B2S_REQUIRE(total < (1LL << 31), "attention_fwd_tc: too many work items");
  p.total_items = static_cast<int>(total);
  // resident CTAs per SM: limited by shared memory (227 KiB, 1 KiB reserved per CTA) and TMEM (512 columns)
  constexpr int by_smem = (227 * 1024) / (C::kSmemBytes + 1024);
  constexpr int by_tmem = 512 / C::kTmemCols;
  constexpr int ctas_per_sm = by_smem < by_tmem ? (by_smem < 1 ? 1 : by_smem) : by_tmem;
  const long long resident = static_cast<long long>(num_sms()) * ctas_per_sm;
  const unsigned grid = static_cast<unsigned>(total < resident ? total : resident);
  rc = launch_pdl_kernel(kern, dim3(grid), dim3(kThreadsTc), C::kSmemBytes, stream, tq, tk, tv, p);
  if (rc != B2S_OK) return rc;
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace

// Packed variable-length attention forward (ops.cuh). q / k / v / o share one 16-bit format (fmt: B2S_FMT_*).
int attention_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* o, long long ld_o,
                  const int* cu_seqlens, int num_seqs, int max_seqlen, long long total_rows, int Hq, int Hkv, int D,
                  float scale, int causal, float* lse, int fmt, cudaStream_t stream, const AttnDrop* drop,
                  int shared_prefix_len) {
  B2S_REQUIRE(q && k && v && o && cu_seqlens, "attention_fwd: null pointer");
  B2S_REQUIRE(shared_prefix_len >= 0 && shared_prefix_len <= 64 && (shared_prefix_len == 0 || (causal && drop == nullptr)),
              "attention_fwd: a shared prefix needs a causal mask, no dropout and at most 64 rows");
  if (drop != nullptr && drop->thresh == 0u) drop = nullptr;
  B2S_REQUIRE(total_rows > 0, "attention_fwd: total_rows must be the row count of the packed q/k/v buffers");
  B2S_REQUIRE(num_seqs > 0 && max_seqlen > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "attention_fwd: bad head counts");
  B2S_REQUIRE(ld_qkv % 8 == 0 && ld_o % 8 == 0, "attention_fwd: strides must keep 16-byte row alignment");
  B2S_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(o) & 15) == 0,
              "attention_fwd: q/k/v/o must be 16-byte aligned");
  B2S_REQUIRE(D == 64 || D == 128, "attention_fwd: head_dim %d unsupported (64 or 128)", D);
  const bool f16 = fmt != 0;
#define B2S_ATTN_ARGS q, k, v, ld_qkv, o, ld_o, cu_seqlens, num_seqs, max_seqlen, total_rows, Hq, Hkv, scale, causal, lse, stream
#define B2S_ATTN_GO(D_, BN_, KVS_, DROP_, DP_)                                     \
  return f16 ? launch_attn_tc<D_, BN_, KVS_, DROP_, true>(B2S_ATTN_ARGS, DP_, shared_prefix_len)      \
             : launch_attn_tc<D_, BN_, KVS_, DROP_, false>(B2S_ATTN_ARGS, DP_, shared_prefix_len)
  if (drop != nullptr) {  // train-mode HuBERT attention only (TF/.../modeling_hubert.py:254)
    B2S_REQUIRE(D == 64 && max_seqlen < 65536, "attention_fwd: attention dropout supports head_dim 64, seqlen < 65536");
    B2S_ATTN_GO(64, 64, 1, true, drop);
  }
  int bn = ctx().attn_bn, kvs = ctx().attn_kvs;  // experiment hook (B2S_OPT_ATTN_*, env seed B2S_ATTN_CFG); 0 = auto
  if (bn == 0) {
    // measured on B200 (profiles/r01_attention_configs.md): the kernel is bound by the per-step latency chain, so the
    // configuration that keeps the most CTAs resident wins at these sequence lengths
    bn = 64;
    kvs = 1;
    if (D == 128 && max_seqlen > 1024) bn = 128, kvs = 2;
  }
  if (D == 64 && bn == 64 && kvs == 1) { B2S_ATTN_GO(64, 64, 1, false, nullptr); }
  if (D == 64 && bn == 64 && kvs == 2) { B2S_ATTN_GO(64, 64, 2, false, nullptr); }
  if (D == 64 && bn == 128 && kvs == 1) { B2S_ATTN_GO(64, 128, 1, false, nullptr); }
  if (D == 64 && bn == 128 && kvs == 2) { B2S_ATTN_GO(64, 128, 2, false, nullptr); }
  if (D == 128 && bn == 64 && kvs == 1) { B2S_ATTN_GO(128, 64, 1, false, nullptr); }
  if (D == 128 && bn == 64 && kvs == 2) { B2S_ATTN_GO(128, 64, 2, false, nullptr); }
  if (D == 128 && bn == 128 && kvs == 2) { B2S_ATTN_GO(128, 128, 2, false, nullptr); }
  if (D == 64 && bn == 64 && kvs == 3) { B2S_ATTN_GO(64, 64, 3, false, nullptr); }    // pipelined (S double-buffered)
  if (D == 128 && bn == 64 && kvs == 3) { B2S_ATTN_GO(128, 64, 3, false, nullptr); }
#undef B2S_ATTN_GO
#undef B2S_ATTN_ARGS
  set_last_error("attention_fwd: tile configuration %d,%d unsupported for head_dim %d", bn, kvs, D);
  return B2S_ERR_UNSUPPORTED;
}

}  // namespace b2s
