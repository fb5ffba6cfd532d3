// b2s_common.cuh -- shared device helpers (bf16 packing, warp/block reductions, vector IO) and the
// host-side status/error plumbing behind the C ABI in include/b2s.h.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
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

void count_launch();
long long launch_count();
#define B2S_LAUNCH_CHECK()              \
  do {                                  \
    ::b2s::count_launch();              \
    B2S_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

int num_sms();

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
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

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
