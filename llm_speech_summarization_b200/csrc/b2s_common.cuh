// b2s_common.cuh -- shared device helpers (bf16 packing, warp/block reductions, vector IO) and the
// host-side status/error plumbing behind the C ABI in include/b2s.h.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace b2s {

// ---- host-side status -----------------------------------------------------------------------
enum Status : int {
  B2S_OK = 0,
  B2S_ERR_INVALID = -1,   // bad argument / unsupported shape
  B2S_ERR_CUDA = -2,      // CUDA runtime / driver error
  B2S_ERR_UNSUPPORTED = -3,
};

void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define B2S_CUDA_CHECK(expr)                                                  \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return ::b2s::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define B2S_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::b2s::set_last_error(__VA_ARGS__);      \
      return ::b2s::B2S_ERR_INVALID;           \
    }                                          \
  } while (0)

// ---- per-instance state (the opaque b2s_handle of include/b2s.h) -------------------------------------------------
// Everything that is not an argument of an entry point lives here: the A/B toggles, the SM budget of the persistent
// kernels, the launch counter, the GEMM timing window and the NCCL communicator of the gradient exchange. A host thread
// works against its CURRENT context (b2s_make_current; a process-wide default exists so callers that never create one
// keep working). Entry points are reentrant across contexts; one context must not be driven by two threads at once.
struct TimedShape {
  int M, N, K, batches, groups, epi, act, mode, bn, cg;
};
struct Context {
  int pdl = 1;          // programmatic dependent launch (B2S_PDL seeds the default)
  int resid_red = 1;    // in-place residual epilogues as L2 reductions (B2S_RESID_RED)
  int tma_epi = 1;      // MODE 0 GEMM outputs through TMA stores (B2S_TMA_EPI)
  int attn_bn = 0, attn_kvs = 0;  // attention forward tile override (B2S_ATTN_CFG="keys per step,K/V stages"; 0 = auto)
  int sm_budget = 0;    // SMs the persistent kernels size their grids for (0 = all)
  int gemm_group_m = 0; // GEMM tile-order override: M tiles per group (0 = the default of 8, gemm_sm100.cu)
  int gemm_epi8 = 1;        // eight epilogue warps for short reductions (B2S_GEMM_EPI8: 0 never, 1 activation + K <= 2048, 2 always)
  int gemm_tail_split = 1;  // K-slice the tiles of a partly filled last round (B2S_GEMM_TAIL_SPLIT; gemm_sm100.cu)
  int* tail_flags = nullptr;  // device flags of the tail split (slice 0 stored -> later slices may add); owned
  int tail_flags_dev = -1;
  std::atomic<long long> launches{0};
  bool timing = false;  // GEMM timing window (bench.py roofline leg)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
  std::vector<TimedShape> shapes;
  TimedShape pending{};
  void* comm = nullptr;  // ncclComm_t of the gradient exchange (comm.cu); owned by the context
  int comm_rank = 0, comm_world = 1;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_done = nullptr;
  Context();
};
Context& ctx();  // the calling thread's current context

void count_launch();
long long launch_count();
#define B2S_LAUNCH_CHECK()              \
  do {                                  \
    ::b2s::count_launch();              \
    B2S_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

int num_sms();  // SMs persistent grids are sized for (the device's, or the calling thread's budget)
void set_sm_budget(int sms);
int sm_budget();

// Launch `kern` with the programmatic-stream-serialization attribute (PDL, see pdl_trigger / pdl_wait below): the grid
// may be scheduled while its predecessor drains and must call pdl_wait() before its first global access.
// B2S_PDL=0 in the environment falls back to plain stream serialisation (A/B runs).
bool pdl_enabled();
template <typename Kern, typename... Args>
int launch_pdl_kernel(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
  if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx", __FILE__, __LINE__);
  return B2S_OK;
}

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 p = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(p);
}
// 16-bit storage format of a tensor: B2S_FMT_BF16 = 0, B2S_FMT_F16 = 1 (include/b2s.h). `f16` is warp-uniform
// everywhere (a kernel argument), so these are a predicated pair of conversions, not divergence.
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 p = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float2 unpack_f16(uint32_t u) {
  __half2 p = *reinterpret_cast<__half2*>(&u);
  return __half22float2(p);
}
__device__ __forceinline__ uint32_t pack_h16(float lo, float hi, int f16) {
  return f16 ? pack_f16(lo, hi) : pack_bf16(lo, hi);
}
__device__ __forceinline__ float2 unpack_h16(uint32_t u, int f16) { return f16 ? unpack_f16(u) : unpack_bf16(u); }
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// erf-GELU (TF/activations.py:318 `gelu` = x * Phi(x), exact erf form). h(|x|) = 0.5 * erfc(|x| / sqrt 2) is evaluated
// with Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 on erf, i.e. the accuracy class of erff): 5 FMAs, one MUFU.RCP and
// one MUFU.EX2 -- about half the instructions of 0.5 * x * (1 + erff(x / sqrt 2)), which is what bounded the FFN1
// GEMM epilogue and the conv front end (profiles/r01_gemm_shapes.md). x * Phi(x) = max(x, 0) - |x| * h(|x|).
__device__ __forceinline__ float half_erfc_abs(float ax /* |x| */, float* gauss /* exp(-x^2/2) */) {
  const float z = ax * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  *gauss = e;
  return p * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float ax = fabsf(x);
  return fmaf(-ax, half_erfc_abs(ax, &e), fmaxf(x, 0.0f));
}
// x * sigmoid(x); the division is rcp.approx + multiply (2 ulp) -- the IEEE division costs ~10 more instructions per
// element, a fifth of everything the gate|up GEMM executes, and the steps run at the board's power cap
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// 8-element vector IO (32 B fp32 / 16 B bf16 per call)
__device__ __forceinline__ void ld8f(const float* p, float (&f)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void st8f(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void st8bf(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void ld8bf(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// format-generic twins (f16 = 1: IEEE half, 0: bfloat16). Pointers to 16-bit tensors are typed __nv_bfloat16* throughout
// the library purely as "2-byte element" pointers; the format travels separately.
__device__ __forceinline__ void unpack8_h16(const uint4& u, float (&f)[8], int f16) {
  if (f16) {
    const float2 a = unpack_f16(u.x), b = unpack_f16(u.y), c = unpack_f16(u.z), d = unpack_f16(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  } else {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
  }
}
__device__ __forceinline__ uint4 pack8_h16(const float (&f)[8], int f16) {
  uint4 u;
  u.x = pack_h16(f[0], f[1], f16); u.y = pack_h16(f[2], f[3], f16);
  u.z = pack_h16(f[4], f[5], f16); u.w = pack_h16(f[6], f[7], f16);
  return u;
}
__device__ __forceinline__ void ld8h(const __nv_bfloat16* p, float (&f)[8], int f16) {
  unpack8_h16(*reinterpret_cast<const uint4*>(p), f, f16);
}
__device__ __forceinline__ void st8h(__nv_bfloat16* p, const float (&f)[8], int f16) {
  *reinterpret_cast<uint4*>(p) = pack8_h16(f, f16);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float h16_to_float(uint16_t bits, int f16) {
  return f16 ? __half2float(__ushort_as_half(bits)) : __uint_as_float(static_cast<uint32_t>(bits) << 16);
}
__device__ __forceinline__ uint16_t float_to_h16(float x, int f16) {
  return f16 ? __half_as_ushort(__float2half_rn(x)) : __bfloat16_as_ushort(__float2bfloat16(x));
}

// erf-GELU derivative: Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float h = half_erfc_abs(fabsf(x), &e);
  return (x >= 0.0f ? 1.0f - h : h) + x * 0.3989422804014327f * e;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while its predecessor drains. pdl_trigger() at the top of a kernel lets such a successor's CTAs take the SMs
// this grid frees in its tail (it is a no-op when the successor was launched normally); the successor calls pdl_wait()
// before its first access to global memory, which blocks until the predecessor has completed and flushed.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// streaming 16-byte load that does not pollute L1
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

}  // namespace b2s
