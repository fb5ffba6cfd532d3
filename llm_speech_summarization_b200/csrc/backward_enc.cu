// backward_enc.cu -- memory-bound backward kernels of the trainable audio encoder (REF/trainer.py:98-105 puts every
// AudioEncoder parameter in the optimizer; autograd through TF/models/hubert/modeling_hubert.py:127-231,505-624 and
// REF/model/audio_encoder.py:56-88). The dense contractions (dgrad / wgrad) run on the tcgen05 GEMM with MN-major
// operands; what is left is here:
//   layernorm_bwd_ex : LayerNorm (+ optional erf-GELU on its output) backward with parameter gradients accumulated
//                      in registers across rows (one atomic per column per block)
//   colsum_accum     : bias gradients
//   avgpool_bwd      : AvgPool1d over time, transposed
//   col2im_add       : strided Conv1d data gradient = gather of the per-tap dgrad GEMM columns
//   conv0_bwd        : conv layer 0 (1 -> 512 channels, CUDA cores) fused with its LayerNorm + GELU backward
#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kWarps = 8;

// F16 is a template parameter, not a kernel argument: a run-time format test inside the row loop splits the unrolled
// loads into separate branch regions and serialises their latencies (+22 % on this kernel, measured round 2).
// One warp per row, 8 * GROUPS columns per lane, dgamma / dbeta sums in registers. Measured and rejected in round 2
// (tools/bench_ln_bwd.py): one block per row with 8 columns per thread (72-96 registers, 20-32 resident warps, three
// block reductions per row) ran at 110-140 us on the 15968 x 1024 case against 73 us for this form.
// dh_colsum (optional): column sums of the dh written here, accumulated in the warp's shared-memory slice (each cell is
// owned by one lane, so no synchronisation) -- the registers are all taken.
template <int GROUPS, bool GELU, bool F16>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_bwd_ex_kernel(const void* __restrict__ x, int x_bf16, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, const void* __restrict__ dy, int dy_bf16, float* dh,
                        int accumulate, __nv_bfloat16* dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        float* __restrict__ dh_colsum, long long rows) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  constexpr int C = GROUPS * 256;
  constexpr int f16 = F16 ? 1 : 0;
  __shared__ __align__(16) float s_red[kWarps][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[GROUPS][8], bt[GROUPS][8], adg[GROUPS][8], adb[GROUPS][8];
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    const int c = (g * 32 + lane) * 8;
    if (dh_colsum != nullptr) {
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      st8f(&s_red[warp][c], z);
    }
    ld8f(gamma + c, gm[g]);
    if constexpr (GELU) ld8f(beta + c, bt[g]);
#pragma unroll
    for (int j = 0; j < 8; ++j) adg[g][j] = adb[g][j] = 0.f;
  }
  for (long long row = static_cast<long long>(blockIdx.x) * kWarps + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * kWarps) {
    float xv[GROUPS][8], dv[GROUPS][8];
    float s = 0.f;
    if constexpr (GROUPS >= 4) {
      // the warp's NEXT row is requested into L2 now: with 8 resident warps per SM (the registers are all taken) nothing
      // else overlaps DRAM latency with this row's three reductions (15968 x 1024: 84 -> 74 us; at 512 columns, where
      // twice the warps are resident, the same prefetch cost 30 %)
      const long long nrow = row + static_cast<long long>(gridDim.x) * kWarps;
      if (nrow < rows) {
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
          const long long e = nrow * C + (g * 32 + lane) * 8;
          prefetch_l2(reinterpret_cast<const char*>(x) + e * (x_bf16 ? 2 : 4));
          prefetch_l2(reinterpret_cast<const char*>(dy) + e * (dy_bf16 ? 2 : 4));
          if (dh != nullptr && accumulate) prefetch_l2(dh + e);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int c = (g * 32 + lane) * 8;
      if (x_bf16) ld8h(reinterpret_cast<const __nv_bfloat16*>(x) + row * C + c, xv[g], f16);
      else ld8f(reinterpret_cast<const float*>(x) + row * C + c, xv[g]);
      if (dy_bf16) ld8h(reinterpret_cast<const __nv_bfloat16*>(dy) + row * C + c, dv[g], f16);
      else ld8f(reinterpret_cast<const float*>(dy) + row * C + c, dv[g]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += xv[g][j];
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xv[g][j] -= mean;
        q = fmaf(xv[g][j], xv[g][j], q);
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = xv[g][j] * rstd;
        float d = dv[g][j];
        if constexpr (GELU) d *= gelu_erf_grad(fmaf(gm[g][j], xh, bt[g][j]));
        adg[g][j] = fmaf(d, xh, adg[g][j]);
        adb[g][j] += d;
        d *= gm[g][j];
        xv[g][j] = xh;
        dv[g][j] = d;
        sg += d;
        sgx = fmaf(d, xh, sgx);
      }
    sg = warp_sum(sg) * (1.0f / C);
    sgx = warp_sum(sgx) * (1.0f / C);
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int c = (g * 32 + lane) * 8;
      float acc[8];
      if (dh != nullptr && accumulate) ld8f(dh + row * C + c, acc);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += rstd * (dv[g][j] - sg - xv[g][j] * sgx);
      if (dh != nullptr) st8f(dh + row * C + c, acc);
      if (dx_bf16 != nullptr) st8h(dx_bf16 + row * C + c, acc, f16);
      if (dh_colsum != nullptr) {
        float cs[8];
        ld8f(&s_red[warp][c], cs);
#pragma unroll
        for (int j = 0; j < 8; ++j) cs[j] += acc[j];
        st8f(&s_red[warp][c], cs);
      }
    }
  }
  if (dh_colsum != nullptr) {  // flush the column sums before the slices are reused for dgamma / dbeta
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += s_red[w][i];
      atomicAdd(dh_colsum + i, t);
    }
    __syncthreads();
  }
  // parameter gradients: registers -> shared (per warp) -> one atomic per column per block
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) st8f(&s_red[warp][(g * 32 + lane) * 8], pass == 0 ? adg[g] : adb[g]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += s_red[w][i];
      atomicAdd((pass == 0 ? dgamma : dbeta) + i, t);
    }
    __syncthreads();
  }
}

// out[c] += sum_r x[r, c]
__global__ void __launch_bounds__(kWarps * 32)
colsum_kernel(const void* __restrict__ x, int x_bf16, float* __restrict__ out, long long rows, int C, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  __shared__ float s_red[kWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long row = static_cast<long long>(blockIdx.y) * kWarps + warp; row < rows;
       row += static_cast<long long>(gridDim.y) * kWarps) {
    float v[8];
    if (x_bf16) ld8h(reinterpret_cast<const __nv_bfloat16*>(x) + row * C + c, v, f16);
    else ld8f(reinterpret_cast<const float*>(x) + row * C + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  st8f(&s_red[warp][lane * 8], acc);
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += s_red[w][threadIdx.x];
  atomicAdd(out + blockIdx.x * 256 + threadIdx.x, t);
}

// dx[b, t, :] = (1 / kernel) * sum_{j : j*stride <= t < j*stride + kernel} dpooled[b, j, :]
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dp, float* __restrict__ dx, int frames, int C, int kernel, int stride,
                   int pooled, long long total4) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const int c = static_cast<int>(i % c4) * 4;
  const long long bt = i / c4;
  const int t = static_cast<int>(bt % frames);
  const long long b = bt / frames;
  int j_hi = t / stride;
  if (j_hi > pooled - 1) j_hi = pooled - 1;
  int j_lo = t - kernel + 1 <= 0 ? 0 : (t - kernel + stride) / stride;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = j_lo; j <= j_hi; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(dp + (b * pooled + j) * C + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float inv = 1.0f / kernel;
  *reinterpret_cast<float4*>(dx + bt * C + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// dx[b, u, :] = sum_{j < k, (u - j) % s == 0, t = (u - j) / s < tout} dcol[b, t, j*C + :]
__global__ void __launch_bounds__(256)
col2im_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int tin, int tout, int k, int s,
              int C, long long total8, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int c = static_cast<int>(i % c8) * 8;
  const long long bu = i / c8;
  const int u = static_cast<int>(bu % tin);
  const long long b = bu / tin;
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  for (int j = 0; j < k; ++j) {
    const int d = u - j;
    if (d < 0 || d % s != 0) continue;
    const int t = d / s;
    if (t >= tout) continue;
    float v[8];
    ld8h(dcol + ((b * tout + t) * k + j) * C + c, v, f16);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += v[q];
  }
  st8h(dx + bu * C + c, acc, f16);
}

// ---- conv layer 0 backward ------------------------------------------------------------------------------------
// Tile = 32 output frames per block iteration. Phase 1 (warp per frame, lane = 16 channels, like the forward kernel):
// recompute conv + LayerNorm, back-propagate GELU and LayerNorm, leave d(pre-norm) in shared memory; dgamma / dbeta
// accumulate in registers. Phase 2 (thread = 2 channels): dW[c, j] += dpre[r, c] * x[r, j], db[c] += dpre[r, c] over
// the tile, accumulators in registers for the whole kernel. Two block barriers per 32 frames.
constexpr int kC0 = 512, kK0 = 10, kS0 = 5, kTile0 = 32, kTT0 = 2;
constexpr int kConv0BwdSmem = (kK0 * kC0 + kTile0 * kC0 + kTile0 * 12) * 4;

template <bool F16>
__global__ void __launch_bounds__(256)
conv0_bwd_kernel(const float* __restrict__ wave, long long wave_stride, int samples, int frames, long long total_rows,
                 const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dW,
                 float* __restrict__ db, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float sm0[];
  float* ws = sm0;                 // [kK0][kC0] transposed taps
  float* sd = ws + kK0 * kC0;      // [kTile0][kC0] d(pre-norm)
  float* sx = sd + kTile0 * kC0;   // [kTile0][12] input windows
  for (int i = threadIdx.x; i < kK0 * kC0; i += blockDim.x) {
    const int c = i / kK0, j = i - c * kK0;
    ws[j * kC0 + c] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float aw[2][kK0], ab[2] = {0.f, 0.f};
#pragma unroll
  for (int j = 0; j < kK0; ++j) aw[0][j] = aw[1][j] = 0.f;
  float adg[16], adb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) adg[i] = adb[i] = 0.f;

  const long long tiles = (total_rows + kTile0 - 1) / kTile0;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // ---------------- phase 1
#pragma unroll 1
    for (int it = 0; it < 4 / kTT0; ++it) {
      float xs[kTT0][kK0], acc[kTT0][16];
      long long rows[kTT0];
#pragma unroll
      for (int tt = 0; tt < kTT0; ++tt) {
        const int rt = warp * 4 + it * kTT0 + tt;
        rows[tt] = tile * kTile0 + rt;
        const bool ok = rows[tt] < total_rows;
        const long long b = ok ? rows[tt] / frames : 0;
        const long long t = ok ? rows[tt] - b * frames : 0;
#pragma unroll
        for (int j = 0; j < kK0; ++j) {
          const long long idx = t * kS0 + j;
          xs[tt][j] = (ok && idx < samples) ? __ldg(wave + b * wave_stride + idx) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kK0; ++j)
          if (lane == j) sx[rt * 12 + j] = xs[tt][j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 bb = *reinterpret_cast<const float2*>(bias + 64 * i + 2 * lane);
#pragma unroll
        for (int tt = 0; tt < kTT0; ++tt) {
          acc[tt][2 * i] = bb.x;
          acc[tt][2 * i + 1] = bb.y;
        }
      }
#pragma unroll
      for (int j = 0; j < kK0; ++j) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ww = *reinterpret_cast<const float2*>(&ws[j * kC0 + 64 * i + 2 * lane]);
#pragma unroll
          for (int tt = 0; tt < kTT0; ++tt) {
            acc[tt][2 * i] = fmaf(ww.x, xs[tt][j], acc[tt][2 * i]);
            acc[tt][2 * i + 1] = fmaf(ww.y, xs[tt][j], acc[tt][2 * i + 1]);
          }
        }
      }
#pragma unroll
      for (int tt = 0; tt < kTT0; ++tt) {
        const int rt = warp * 4 + it * kTT0 + tt;
        const bool ok = rows[tt] < total_rows;
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += acc[tt][i];
        const float mean = warp_sum(sum) * (1.0f / kC0);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          acc[tt][i] -= mean;
          q = fmaf(acc[tt][i], acc[tt][i], q);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / kC0) + eps);
        float g16[16];
        float sg = 0.f, sgx = 0.f;
        // all eight dy words of the frame are requested before the first is used (one load latency per frame, not 8)
        uint32_t du[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          du[i] = ok ? __ldg(reinterpret_cast<const uint32_t*>(dy + rows[tt] * kC0 + 64 * i + 2 * lane)) : 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 gm = *reinterpret_cast<const float2*>(gamma + 64 * i + 2 * lane);
          const float2 bt = *reinterpret_cast<const float2*>(beta + 64 * i + 2 * lane);
          const float2 d = F16 ? unpack_f16(du[i]) : unpack_bf16(du[i]);
          const float xh0 = acc[tt][2 * i] * rstd, xh1 = acc[tt][2 * i + 1] * rstd;
          const float dz0 = d.x * gelu_erf_grad(fmaf(gm.x, xh0, bt.x));
          const float dz1 = d.y * gelu_erf_grad(fmaf(gm.y, xh1, bt.y));
          adg[2 * i] = fmaf(dz0, xh0, adg[2 * i]);
          adg[2 * i + 1] = fmaf(dz1, xh1, adg[2 * i + 1]);
          adb[2 * i] += dz0;
          adb[2 * i + 1] += dz1;
          acc[tt][2 * i] = xh0;
          acc[tt][2 * i + 1] = xh1;
          g16[2 * i] = gm.x * dz0;
          g16[2 * i + 1] = gm.y * dz1;
          sg += g16[2 * i] + g16[2 * i + 1];
          sgx = fmaf(g16[2 * i], xh0, fmaf(g16[2 * i + 1], xh1, sgx));
        }
        sg = warp_sum(sg) * (1.0f / kC0);
        sgx = warp_sum(sgx) * (1.0f / kC0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float o0 = rstd * (g16[2 * i] - sg - acc[tt][2 * i] * sgx);
          const float o1 = rstd * (g16[2 * i + 1] - sg - acc[tt][2 * i + 1] * sgx);
          *reinterpret_cast<float2*>(&sd[rt * kC0 + 64 * i + 2 * lane]) = make_float2(o0, o1);
        }
      }
    }
    __syncthreads();
    // ---------------- phase 2: thread owns channels 2*tid, 2*tid + 1
#pragma unroll 4
    for (int r = 0; r < kTile0; ++r) {
      const float2 d = *reinterpret_cast<const float2*>(&sd[r * kC0 + 2 * threadIdx.x]);
      ab[0] += d.x;
      ab[1] += d.y;
#pragma unroll
      for (int j = 0; j < kK0; ++j) {
        const float xv = sx[r * 12 + j];
        aw[0][j] = fmaf(d.x, xv, aw[0][j]);
        aw[1][j] = fmaf(d.y, xv, aw[1][j]);
      }
    }
    __syncthreads();
  }
  const int c0 = 2 * threadIdx.x;
#pragma unroll
  for (int j = 0; j < kK0; ++j) {
    atomicAdd(dW + c0 * kK0 + j, aw[0][j]);
    atomicAdd(dW + (c0 + 1) * kK0 + j, aw[1][j]);
  }
  atomicAdd(db + c0, ab[0]);
  atomicAdd(db + c0 + 1, ab[1]);
  // dgamma / dbeta: per-warp registers -> shared [8][512] -> one atomic per channel per block
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* a = pass == 0 ? adg : adb;
      *reinterpret_cast<float2*>(&sd[warp * kC0 + 64 * i + 2 * lane]) = make_float2(a[2 * i], a[2 * i + 1]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < kC0; c += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) t += sd[wq * kC0 + c];
      atomicAdd((pass == 0 ? dgamma : dbeta) + c, t);
    }
    __syncthreads();
  }
}

}  // namespace

int layernorm_bwd_ex(const void* x, int x_bf16, const float* gamma, const float* beta, int act_gelu, float eps,
                     const void* dy, int dy_bf16, float* dh, int accumulate, void* dx_bf16, float* dgamma, float* dbeta,
                     long long rows, int C, int fmt, cudaStream_t stream, float* dh_colsum) {
  B2S_REQUIRE(x && gamma && dy && dgamma && dbeta && (dh || dx_bf16), "layernorm_bwd_ex: null pointer");
  B2S_REQUIRE(!act_gelu || beta != nullptr, "layernorm_bwd_ex: the GELU variant needs beta");
  if (rows <= 0) return B2S_OK;
  long long blocks = (rows + kWarps - 1) / kWarps;
  const long long cap = 4LL * num_sms();
  if (blocks > cap) blocks = cap;
  const unsigned grid = static_cast<unsigned>(blocks);
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define B2S_LNBWD_(G, A, F)                                                                                            \
  layernorm_bwd_ex_kernel<G, A, F><<<grid, kWarps * 32, 0, stream>>>(x, x_bf16, gamma, beta, eps, dy, dy_bf16, dh,      \
                                                                     accumulate, dxb, dgamma, dbeta, dh_colsum, rows)
#define B2S_LNBWD(G)                                  \
  if (act_gelu) {                                     \
    if (fmt) B2S_LNBWD_(G, true, true);               \
    else B2S_LNBWD_(G, true, false);                  \
  } else {                                            \
    if (fmt) B2S_LNBWD_(G, false, true);              \
    else B2S_LNBWD_(G, false, false);                 \
  }
  switch (C) {
    case 256: B2S_LNBWD(1); break;
    case 512: B2S_LNBWD(2); break;
    case 1024: B2S_LNBWD(4); break;
    default:
      set_last_error("layernorm_bwd_ex: unsupported width %d (256, 512, 1024)", C);
      return B2S_ERR_UNSUPPORTED;
  }
#undef B2S_LNBWD
#undef B2S_LNBWD_
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int colsum_accum(const void* x, int x_bf16, float* out, long long rows, int C, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(x && out, "colsum_accum: null pointer");
  B2S_REQUIRE(C % 256 == 0, "colsum_accum: C must be a multiple of 256");
  if (rows <= 0) return B2S_OK;
  long long chunks = (rows + kWarps * 16 - 1) / (kWarps * 16);
  const long long cap = (4LL * num_sms() + C / 256 - 1) / (C / 256);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  colsum_kernel<<<dim3(C / 256, static_cast<unsigned>(chunks)), kWarps * 32, 0, stream>>>(x, x_bf16, out, rows, C, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int avgpool_bwd(const float* dpooled, float* dx, int batches, int frames, int C, int kernel, int stride, int pooled,
                cudaStream_t stream) {
  B2S_REQUIRE(dpooled && dx && C % 4 == 0 && kernel > 0 && stride > 0 && pooled > 0, "avgpool_bwd: bad arguments");
  const long long total4 = static_cast<long long>(batches) * frames * (C / 4);
  if (total4 <= 0) return B2S_OK;
  avgpool_bwd_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(dpooled, dx, frames, C, kernel,
                                                                                      stride, pooled, total4);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int col2im_add(const void* dcol_bf16, void* dx_bf16, int batches, int tin, int tout, int k, int s, int C, int fmt,
               cudaStream_t stream) {
  B2S_REQUIRE(dcol_bf16 && dx_bf16 && C % 8 == 0 && k > 0 && s > 0, "col2im_add: bad arguments");
  const long long total8 = static_cast<long long>(batches) * tin * (C / 8);
  if (total8 <= 0) return B2S_OK;
  col2im_kernel<<<static_cast<unsigned>((total8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dcol_bf16), reinterpret_cast<__nv_bfloat16*>(dx_bf16), tin, tout, k, s, C,
      total8, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int conv0_bwd(const float* wave, long long wave_stride, int batches, int samples, const float* w, const float* bias,
              const float* gamma, const float* beta, float eps, const void* dy_bf16, int frames, float* dW, float* db,
              float* dgamma, float* dbeta, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(wave && w && bias && gamma && beta && dy_bf16 && dW && db && dgamma && dbeta, "conv0_bwd: null pointer");
  B2S_REQUIRE(frames == (samples - kK0) / kS0 + 1, "conv0_bwd: frames mismatch");
  const long long total = static_cast<long long>(batches) * frames;
  static bool attr_set = false;
  if (!attr_set) {
    B2S_CUDA_CHECK(cudaFuncSetAttribute(conv0_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv0BwdSmem));
    B2S_CUDA_CHECK(cudaFuncSetAttribute(conv0_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv0BwdSmem));
    attr_set = true;
  }
  long long blocks = (total + kTile0 - 1) / kTile0;
  const long long cap = num_sms();  // 177 registers x 256 threads + 87 KB shared memory: one resident block per SM
  if (blocks > cap) blocks = cap;
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(dy_bf16);
  if (fmt)
    conv0_bwd_kernel<true><<<static_cast<unsigned>(blocks), 256, kConv0BwdSmem, stream>>>(
        wave, wave_stride, samples, frames, total, w, bias, gamma, beta, eps, dy, dW, db, dgamma, dbeta);
  else
    conv0_bwd_kernel<false><<<static_cast<unsigned>(blocks), 256, kConv0BwdSmem, stream>>>(
        wave, wave_stride, samples, frames, total, w, bias, gamma, beta, eps, dy, dW, db, dgamma, dbeta);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
