// Shared host/device helpers for the modelcompose_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/modelcompose_b200.h"

namespace mc {

// ---- thread-local error string -------------------------------------------------------------
std::string& last_error();
int fail(int code, const char* fmt, ...);

#define MC_CUDA_OK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::mc::fail(MC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define MC_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return ::mc::fail(MC_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// per-device SM count (cached)
int sm_count();

// ---- programmatic dependent launch (decode chain) ---------------------------------------------
// mc_set_launch_mode(1) makes the decode-chain entry points of the calling thread launch their kernels with
// cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's CTAs may then become resident while its predecessor in the
// stream is still running (every kernel of the chain calls griddepcontrol.launch_dependents first thing), run whatever does
// not depend on the predecessor — barrier setup, descriptor fetch, and in the skinny-linear kernel the first ring of WEIGHT
// loads — and block in griddepcontrol.wait before touching anything a predecessor writes.  Without the attribute both
// instructions are no-ops, so the same kernels serve the ordinary launches.
int& launch_mode();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (launch_mode() & 1) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline size_t dtype_size(int dt) { return dt == MC_F32 ? 4 : 2; }
inline bool dtype_valid(int dt) { return dt == MC_F32 || dt == MC_F16 || dt == MC_BF16; }

// ---- device: streaming 128/256-bit global accesses ------------------------------------------
// Read-once streams: bypass L1 allocation; the 256-bit form also marks the line evict-first in L2
// (ptxas only accepts the bare .L2::evict_first qualifier on .v8.b32/.v4.b64 loads) so 50+ GB of
// single-use data does not displace anything useful (B200 L2 = 126 MB).
template <int BYTES>
struct Vec;
template <>
struct alignas(16) Vec<16> {
  uint32_t w[4];
};
template <>
struct alignas(32) Vec<32> {
  uint32_t w[8];
};
template <>
struct alignas(8) Vec<8> {
  uint32_t w[2];
};
template <>
struct alignas(4) Vec<4> {
  uint32_t w[1];
};

__device__ __forceinline__ Vec<16> ld_stream(const Vec<16>* p) {
  Vec<16> v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
               : "l"(p));
  return v;
}
__device__ __forceinline__ Vec<32> ld_stream(const Vec<32>* p) {
  Vec<32> v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]),
                 "=r"(v.w[7])
               : "l"(p));
  return v;
}
__device__ __forceinline__ Vec<8> ld_stream(const Vec<8>* p) {
  Vec<8> v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(v.w[0]), "=r"(v.w[1]) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(Vec<8>* p, const Vec<8>& v) {
  asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]) : "memory");
}
__device__ __forceinline__ void st_stream(Vec<16>* p, const Vec<16>& v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]),
               "r"(v.w[3])
               : "memory");
}
__device__ __forceinline__ void st_stream(Vec<32>* p, const Vec<32>& v) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]),
               "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4]), "r"(v.w[5]), "r"(v.w[6]), "r"(v.w[7])
               : "memory");
}

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- device: scalar conversions with defined rounding ---------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);  // round-to-nearest-even
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// One row of LlamaRMSNorm by ONE warp (fp32 variance, x * rsqrt(var + eps) rounded to the storage dtype, then weight * that, rounded
// again): the body of rmsnorm_kernel (mc_ops.cu), shared with the skinny-linear kernel that computes the norm under its weight ramp.
template <typename T>
__device__ __forceinline__ void rmsnorm_row(const T* __restrict__ x_row, const T* __restrict__ w, T* __restrict__ out_row, int hidden, float eps,
                                            int lane) {
  const int n_vec = hidden >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x_row);
  float ss = 0.f;
  for (int i = lane; i < n_vec; i += 32) {
    const uint4 u = xr[i];
    const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float f = to_f32<T>(e[k]);
      ss += f * f;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, d);
  const float inv = rsqrtf(ss / (float)hidden + eps);
  uint4* orow = reinterpret_cast<uint4*>(out_row);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  for (int i = lane; i < n_vec; i += 32) {
    const uint4 u = xr[i];  // second read hits L1/L2
    const uint4 wu = wr[i];
    const T* e = reinterpret_cast<const T*>(&u);
    const T* we = reinterpret_cast<const T*>(&wu);
    uint4 o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const T h = from_f32<T>(to_f32<T>(e[k]) * inv);
      oe[k] = from_f32<T>(to_f32<T>(we[k]) * to_f32<T>(h));
    }
    orow[i] = o;
  }
}

}  // namespace mc
