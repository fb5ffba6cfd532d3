// backward.cu -- the memory-bound pieces of the training step's backward pass and the optimizer
// (REF/trainer.py:372-384: total_loss / grad_accum -> backward -> AdamW.step; autograd through the modules of
// TF/models/llama/modeling_llama.py and TF/models/hubert/modeling_hubert.py). One warp per row for the norm
// backwards (row in registers, fp32), vectorised elementwise kernels for the activation backwards.
#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kWarps = 8;

// RMSNorm backward (no weight gradient: the LLM is frozen, REF/trainer.py:63-64):
//   y = w * x * rstd  =>  dx = rstd * (w dy) - x * rstd^3 / C * sum_k x_k w_k dy_k ;   dh[dst] += dx
template <int GROUPS>
__global__ void __launch_bounds__(kWarps * 32)
rmsnorm_bwd_kernel(const float* __restrict__ x, const int* __restrict__ x_index, const float* __restrict__ w, float eps,
                   const float* __restrict__ dy, float* dh, const int* __restrict__ dh_index,
                   __nv_bfloat16* dh_bf16, long long rows, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  constexpr int C = GROUPS * 256;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long xr = x_index ? x_index[row] : row;
  const long long dr = dh_index ? dh_index[row] : row;
  float xv[GROUPS][8], gv[GROUPS][8];
  float ss = 0.f, dot = 0.f;
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    const int c = (g * 32 + lane) * 8;
    float wv[8], dv[8];
    ld8f(x + xr * C + c, xv[g]);
    ld8f(w + c, wv);
    ld8f(dy + row * C + c, dv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gv[g][j] = wv[j] * dv[j];
      ss = fmaf(xv[g][j], xv[g][j], ss);
      dot = fmaf(xv[g][j], gv[g][j], dot);
    }
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float rstd = rsqrtf(ss * (1.0f / C) + eps);
  const float coef = dot * rstd * rstd * rstd * (1.0f / C);
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    const int c = (g * 32 + lane) * 8;
    float acc[8];
    ld8f(dh + dr * C + c, acc);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += rstd * gv[g][j] - coef * xv[g][j];
    st8f(dh + dr * C + c, acc);
    if (dh_bf16 != nullptr) st8h(dh_bf16 + dr * C + c, acc, f16);
  }
}

// LayerNorm backward: xhat = (x - mean) rstd, y = gamma xhat + beta
//   dx = rstd * (g - mean(g) - xhat * mean(g xhat)),  g = gamma dy ;  dgamma += dy xhat ; dbeta += dy
// dx is ADDED to dh (residual stream gradient) or written (accumulate = 0).
template <int GROUPS, bool DY_BF16>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, float eps, const void* __restrict__ dy,
                     float* dh, int accumulate, __nv_bfloat16* dh_bf16, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, long long rows, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  constexpr int C = GROUPS * 256;
  __shared__ float s_dg[C], s_db[C];
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_dg[i] = s_db[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarps + (threadIdx.x >> 5);
  if (row < rows) {
    float xv[GROUPS][8], dv[GROUPS][8];
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int c = (g * 32 + lane) * 8;
      ld8f(x + row * C + c, xv[g]);
      if constexpr (DY_BF16) ld8h(reinterpret_cast<const __nv_bfloat16*>(dy) + row * C + c, dv[g], f16);
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
    for (int g = 0; g < GROUPS; ++g) {
      const int c = (g * 32 + lane) * 8;
      float gm[8];
      ld8f(gamma + c, gm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = xv[g][j] * rstd;
        atomicAdd(&s_dg[c + j], dv[g][j] * xh);
        atomicAdd(&s_db[c + j], dv[g][j]);
        xv[g][j] = xh;             // xhat
        dv[g][j] *= gm[j];         // g = gamma dy
        sg += dv[g][j];
        sgx = fmaf(dv[g][j], xh, sgx);
      }
    }
    sg = warp_sum(sg) * (1.0f / C);
    sgx = warp_sum(sgx) * (1.0f / C);
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int c = (g * 32 + lane) * 8;
      float acc[8];
      if (accumulate) ld8f(dh + row * C + c, acc);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += rstd * (dv[g][j] - sg - xv[g][j] * sgx);
      st8f(dh + row * C + c, acc);
      if (dh_bf16 != nullptr) st8h(dh_bf16 + row * C + c, acc, f16);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, s_dg[i]);
    atomicAdd(dbeta + i, s_db[i]);
  }
}

// SwiGLU backward on the packed layout (64 gate | 64 up per 128 columns):
//   a = silu(g) u ; dg = da u sig(g) (1 + g (1 - sig(g))) ; du = da silu(g)
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, const __nv_bfloat16* __restrict__ dact,
                  __nv_bfloat16* __restrict__ dgu, long long rows, int F, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // one 8-element vector of dact
  const long long per_row = F / 8;
  if (i >= rows * per_row) return;
  const long long row = i / per_row;
  const int v = static_cast<int>(i - row * per_row);
  const int col = v * 8;                       // column in [0, F)
  const int blk = col / 64, within = col % 64;  // 64-wide block
  const long long gbase = row * 2 * F + blk * 128 + within;
  float g[8], u[8], da[8], dg[8], du[8];
  ld8h(gu + gbase, g, f16);
  ld8h(gu + gbase + 64, u, f16);
  ld8h(dact + row * F + col, da, f16);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float sg = __fdividef(1.0f, 1.0f + __expf(-g[j]));
    dg[j] = da[j] * u[j] * sg * (1.0f + g[j] * (1.0f - sg));
    du[j] = da[j] * g[j] * sg;
  }
  st8h(dgu + gbase, dg, f16);
  st8h(dgu + gbase + 64, du, f16);
}

// erf-GELU backward: dpre = dy * (Phi(x) + x phi(x))
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dy,
                __nv_bfloat16* __restrict__ dpre, long long n8, DropSpec drop, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float x[8], d[8];
  ld8h(pre + i * 8, x, f16);
  ld8h(dy + i * 8, d, f16);
  if (drop.thresh != 0u) {  // dy is the gradient w.r.t. dropout(gelu(pre)): mask first (activation dropout site)
    const uint32_t e0 = static_cast<uint32_t>(i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = rng_keep(e0 + j, drop.k1, drop.k2, drop.thresh) ? d[j] * drop.inv_keep : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    d[j] *= gelu_erf_grad(x[j]);
  }
  st8h(dpre + i * 8, d, f16);
}

// The same with the column sums of dpre (the bias gradient) accumulated on the way: a thread keeps the SAME 8 columns for
// every row it visits (256 threads = 2048 columns, F / 2048 blocks side by side, the rest of the grid strides over rows),
// so the sums are 8 registers and one atomic per column per block.
__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dy,
                       __nv_bfloat16* __restrict__ dpre, long long rows, int F, float* __restrict__ colsum,
                       DropSpec drop, int f16) {
  pdl_trigger();
  const int tpr = F / 8, bpr = tpr / 256;
  const int cg = (blockIdx.x % bpr) * 256 + threadIdx.x;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  // four rows per iteration: eight 16-byte loads in flight per thread (one row at a time ran at 2/3 of the plain kernel)
  const long long rstep = gridDim.x / bpr;
  for (long long row0 = blockIdx.x / bpr; row0 < rows; row0 += 4 * rstep) {
    uint4 xr[4], dr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long row = row0 + u * rstep;
      if (row < rows) {
        const long long i = row * tpr + cg;
        xr[u] = *reinterpret_cast<const uint4*>(pre + i * 8);
        dr[u] = *reinterpret_cast<const uint4*>(dy + i * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long row = row0 + u * rstep;
      if (row < rows) {
        const long long i = row * tpr + cg;
        float x[8], d[8];
        unpack8_h16(xr[u], x, f16);
        unpack8_h16(dr[u], d, f16);
        if (drop.thresh != 0u) {
          const uint32_t e0 = static_cast<uint32_t>(i * 8);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            d[j] = rng_keep(e0 + j, drop.k1, drop.k2, drop.thresh) ? d[j] * drop.inv_keep : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          d[j] *= gelu_erf_grad(x[j]);
          acc[j] += d[j];
        }
        st8h(dpre + i * 8, d, f16);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(colsum + cg * 8 + j, acc[j]);
}

// dh[rows_a[i]] += c * (h[rows_a[i]] - h[rows_b[i]])  (feature-distillation MSE backward, REF/trainer.py:358-370)
__global__ void __launch_bounds__(256)
add_rowdiff_kernel(const float* __restrict__ h, const int* __restrict__ rows_a, const int* __restrict__ rows_b,
                   const float* __restrict__ coef, const float* __restrict__ loss_scale, float* dh,
                   __nv_bfloat16* dh_bf16, int pairs, int C, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= pairs) return;
  const long long ra = rows_a[i], rb = rows_b[i];
  const float c = coef[i] * (loss_scale != nullptr ? *loss_scale : 1.0f);
  for (int k = lane * 4; k < C; k += 128) {
    const float4 a = *reinterpret_cast<const float4*>(h + ra * C + k);
    const float4 b = *reinterpret_cast<const float4*>(h + rb * C + k);
    float4 d = *reinterpret_cast<float4*>(dh + ra * C + k);
    d.x += c * (a.x - b.x); d.y += c * (a.y - b.y); d.z += c * (a.z - b.z); d.w += c * (a.w - b.w);
    *reinterpret_cast<float4*>(dh + ra * C + k) = d;
    if (dh_bf16 != nullptr) {
      uint2 u;
      u.x = pack_h16(d.x, d.y, f16);
      u.y = pack_h16(d.z, d.w, f16);
      *reinterpret_cast<uint2*>(dh_bf16 + ra * C + k) = u;
    }
  }
}

// out[i, :] = src[index[i], :]   (fp32 rows; index < 0 -> zeros)
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ index, float* __restrict__ out, long long rows,
                   int C) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long i = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (i >= rows) return;
  const long long s = index[i];
  for (int k = lane * 4; k < C; k += 128) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s >= 0) v = *reinterpret_cast<const float4*>(src + s * C + k);
    *reinterpret_cast<float4*>(out + i * C + k) = v;
  }
}

// AdamW (torch.optim.AdamW semantics, REF/trainer.py:98-105): decoupled weight decay, bias-corrected moments.
// scaler (optional, device): the dynamic loss-scaling state of grad_scaler_update below -- the gradients carry the
// factor scaler->scale, and a step whose gradients held an inf / nan is skipped entirely (torch.cuda.amp.GradScaler.step,
// REF/trainer.py:381); the bias corrections then use the number of steps actually taken.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2,
             float grad_scale, const GradScalerState* __restrict__ scaler, int vec) {
  if (scaler != nullptr) {  // block-uniform: one thread evaluates the state, everybody reads it from shared memory
    __shared__ float s_par[3];
    __shared__ int s_skip;
    if (threadIdx.x == 0) {
      s_skip = scaler->found_inf;
      const float t = static_cast<float>(scaler->opt_steps + 1);
      s_par[0] = grad_scale / scaler->scale;
      s_par[1] = 1.0f - powf(beta1, t);
      s_par[2] = 1.0f - powf(beta2, t);
    }
    __syncthreads();
    if (s_skip != 0) return;
    grad_scale = s_par[0];
    bc1 = s_par[1];
    bc2 = s_par[2];
  }
  // grid-stride over float4 groups (the scaler prologue above is paid once per block, not once per 256 elements), then
  // the scalar tail; `vec` = all four buffers 16-byte aligned
  auto upd = [&](float& pv, float gr, float& mv, float& vv) {
    gr *= grad_scale;
    pv *= (1.0f - lr * wd);
    mv = beta1 * mv + (1.0f - beta1) * gr;
    vv = beta2 * vv + (1.0f - beta2) * gr * gr;
    pv -= lr * (mv / bc1) / (sqrtf(vv / bc2) + eps);
  };
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = gid; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
    upd(P.x, G.x, M.x, V.x);
    upd(P.y, G.y, M.y, V.y);
    upd(P.z, G.z, M.z, V.z);
    upd(P.w, G.w, M.w, V.w);
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    reinterpret_cast<float4*>(p)[i] = P;
  }
  for (long long i = n4 * 4 + gid; i < n; i += stride) {
    float pv = p[i], mv = m[i], vv = v[i];
    upd(pv, g[i], mv, vv);
    m[i] = mv;
    v[i] = vv;
    p[i] = pv;
  }
}

// found_inf |= any non-finite element of g (the unscale_ + inf check of GradScaler, fused: nothing is rewritten)
__global__ void __launch_bounds__(256)
nonfinite_check_kernel(const float* __restrict__ g, long long n4, GradScalerState* __restrict__ scaler) {
  bool bad = false;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    // x - x is 0 for finite x and nan for inf / nan
    bad |= !((v.x - v.x) + (v.y - v.y) + (v.z - v.z) + (v.w - v.w) == 0.0f);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(&scaler->found_inf, 1);
}

// GradScaler.update(): scale *= backoff on an overflow step, else *= growth after `interval` clean steps
__global__ void grad_scaler_update_kernel(GradScalerState* s, float growth, float backoff, int interval) {
  if (s->found_inf != 0) {
    s->scale *= backoff;
    s->growth_tracker = 0;
    s->skipped_steps += 1;
  } else {
    s->opt_steps += 1;
    if (++s->growth_tracker >= interval) {
      s->scale *= growth;
      s->growth_tracker = 0;
    }
  }
  s->found_inf = 0;
}

}  // namespace

#define B2S_GROUPS_SWITCH(C, ...)                                  \
  switch ((C) / 256) {                                             \
    case 1: { constexpr int G = 1; __VA_ARGS__; break; }           \
    case 2: { constexpr int G = 2; __VA_ARGS__; break; }           \
    case 4: { constexpr int G = 4; __VA_ARGS__; break; }           \
    case 12: { constexpr int G = 12; __VA_ARGS__; break; }         \
    default:                                                       \
      set_last_error("norm backward: unsupported width %d (256, 512, 1024, 3072)", (C)); \
      return B2S_ERR_UNSUPPORTED;                                  \
  }

int rmsnorm_bwd(const float* x, const int* x_index, const float* w, float eps, const float* dy, float* dh,
                const int* dh_index, void* dh_bf16, long long rows, int C, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(x && w && dy && dh, "rmsnorm_bwd: null pointer");
  B2S_REQUIRE(C % 256 == 0, "rmsnorm_bwd: C must be a multiple of 256");
  if (rows <= 0) return B2S_OK;
  const unsigned grid = static_cast<unsigned>((rows + kWarps - 1) / kWarps);
  B2S_GROUPS_SWITCH(C, (rmsnorm_bwd_kernel<G><<<grid, kWarps * 32, 0, stream>>>(
                           x, x_index, w, eps, dy, dh, dh_index, reinterpret_cast<__nv_bfloat16*>(dh_bf16), rows, fmt)));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int layernorm_bwd(const float* x, const float* gamma, float eps, const void* dy, int dy_bf16, float* dh, int accumulate,
                  void* dh_bf16, float* dgamma, float* dbeta, long long rows, int C, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(x && gamma && dy && dh && dgamma && dbeta, "layernorm_bwd: null pointer");
  B2S_REQUIRE(C % 256 == 0, "layernorm_bwd: C must be a multiple of 256");
  if (rows <= 0) return B2S_OK;
  const unsigned grid = static_cast<unsigned>((rows + kWarps - 1) / kWarps);
  if (dy_bf16) {
    B2S_GROUPS_SWITCH(C, (layernorm_bwd_kernel<G, true><<<grid, kWarps * 32, 0, stream>>>(
                             x, gamma, eps, dy, dh, accumulate, reinterpret_cast<__nv_bfloat16*>(dh_bf16), dgamma, dbeta,
                             rows, fmt)));
  } else {
    B2S_GROUPS_SWITCH(C, (layernorm_bwd_kernel<G, false><<<grid, kWarps * 32, 0, stream>>>(
                             x, gamma, eps, dy, dh, accumulate, reinterpret_cast<__nv_bfloat16*>(dh_bf16), dgamma, dbeta,
                             rows, fmt)));
  }
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int swiglu_bwd(const void* gu, const void* dact, void* dgu, long long rows, int F, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(gu && dact && dgu, "swiglu_bwd: null pointer");
  B2S_REQUIRE(F % 64 == 0, "swiglu_bwd: F must be a multiple of 64");
  if (rows <= 0) return B2S_OK;
  const long long n = rows * (F / 8);
  swiglu_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(gu), reinterpret_cast<const __nv_bfloat16*>(dact),
      reinterpret_cast<__nv_bfloat16*>(dgu), rows, F, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int gelu_bwd(const void* pre, const void* dy, void* dpre, long long n, int fmt, cudaStream_t stream,
             const DropSpec* drop, float* colsum, int F) {
  B2S_REQUIRE(pre && dy && dpre, "gelu_bwd: null pointer");
  B2S_REQUIRE(n % 8 == 0, "gelu_bwd: element count must be a multiple of 8");
  B2S_REQUIRE(drop == nullptr || n < (1LL << 32), "gelu_bwd: dropout element index exceeds 32 bits");
  if (n <= 0) return B2S_OK;
  if (colsum != nullptr) {
    B2S_REQUIRE(F > 0 && F % 2048 == 0 && n % F == 0, "gelu_bwd: the fused column sum needs a row width that is a multiple of 2048");
    const int bpr = F / 2048;
    const long long rows = n / F;
    long long per_col = 4LL * num_sms() / bpr;  // four 256-thread blocks per SM
    if (per_col > rows) per_col = rows;
    if (per_col < 1) per_col = 1;
    gelu_bwd_colsum_kernel<<<static_cast<unsigned>(per_col * bpr), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(pre), reinterpret_cast<const __nv_bfloat16*>(dy),
        reinterpret_cast<__nv_bfloat16*>(dpre), rows, F, colsum, drop ? *drop : DropSpec{}, fmt);
    B2S_LAUNCH_CHECK();
    return B2S_OK;
  }
  gelu_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(pre), reinterpret_cast<const __nv_bfloat16*>(dy),
      reinterpret_cast<__nv_bfloat16*>(dpre), n / 8, drop ? *drop : DropSpec{}, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int add_rowdiff(const float* h, const int* rows_a, const int* rows_b, const float* coef, const float* loss_scale,
                float* dh, void* dh_bf16, int pairs, int C, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(h && rows_a && rows_b && coef && dh, "add_rowdiff: null pointer");
  B2S_REQUIRE(C % 4 == 0, "add_rowdiff: C must be a multiple of 4");
  if (pairs <= 0) return B2S_OK;
  add_rowdiff_kernel<<<(pairs + 7) / 8, 256, 0, stream>>>(h, rows_a, rows_b, coef, loss_scale, dh,
                                                          reinterpret_cast<__nv_bfloat16*>(dh_bf16), pairs, C, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int gather_rows_f32(const float* src, const int* index, float* out, long long rows, int C, cudaStream_t stream) {
  B2S_REQUIRE(src && index && out, "gather_rows_f32: null pointer");
  B2S_REQUIRE(C % 4 == 0, "gather_rows_f32: C must be a multiple of 4");
  if (rows <= 0) return B2S_OK;
  gather_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(src, index, out, rows, C);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
               float weight_decay, int step, float grad_scale, const GradScalerState* scaler, cudaStream_t stream) {
  B2S_REQUIRE(p && g && m && v && (step >= 1 || scaler != nullptr), "adamw_step: bad arguments");
  if (n <= 0) return B2S_OK;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  const int vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                    reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  long long blocks = ((vec ? n / 4 : n) + 255) / 256;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1,
                                                                  bc2, grad_scale, scaler, vec);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int nonfinite_check(const float* g, long long n, GradScalerState* scaler, cudaStream_t stream) {
  B2S_REQUIRE(g && scaler && n % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0,
              "nonfinite_check: needs a 16-byte aligned buffer of a multiple of 4 floats");
  if (n <= 0) return B2S_OK;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  nonfinite_check_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(g, n / 4, scaler);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int grad_scaler_update(GradScalerState* scaler, float growth, float backoff, int interval, cudaStream_t stream) {
  B2S_REQUIRE(scaler && growth >= 1.0f && backoff > 0.0f && backoff <= 1.0f && interval >= 1,
              "grad_scaler_update: bad arguments");
  grad_scaler_update_kernel<<<1, 1, 0, stream>>>(scaler, growth, backoff, interval);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
