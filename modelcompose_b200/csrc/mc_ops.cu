// Row-wise glue of the decoder layer around the routed linears: RMSNorm and rotary position embedding.
//
// Replaces (reference paths, via transformers==4.31.0 which the reference star-imports):
//   LlamaRMSNorm.forward  (used at modelcompose/model/language_model/multimodal_llama.py:405-406,:441,:455,:603):
//       variance in fp32, x * rsqrt(var + eps) cast back to the storage dtype, then weight * that (storage dtype)
//   apply_rotary_pos_emb  (multimodal_llama.py:281-282): q*cos + rotate_half(q)*sin with the cos/sin cache cast to
//       the storage dtype; each product and the sum are rounded in the storage dtype like the eager ops.
// Both are HBM-bound row streams (read once, write once, 128-bit accesses).
#include <algorithm>

#include "mc_common.cuh"

namespace mc {

template <typename T>
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ out, long long rows, int hidden,
               long long ldx, long long ldo, float eps) {
  griddep_launch_dependents();  // decode chain (mc_set_launch_mode): no-ops in an ordinary launch
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps)
    rmsnorm_row<T>(x + row * ldx, w, out + row * ldo, hidden, eps, lane);
}

// One thread rotates 8 element pairs (i, i + D/2) of one head of q or k.
template <typename T>
__global__ void __launch_bounds__(256)
rope_kernel(T* __restrict__ q, T* __restrict__ k, const T* __restrict__ cos_t, const T* __restrict__ sin_t,
            long long tokens, int seq_len, int pos_offset, int n_heads, int head_dim, long long ldq, long long ldk) {
  const int half = head_dim >> 1;
  const int vec_per_head = half >> 3;
  const long long per_token = (long long)2 * n_heads * vec_per_head;  // q heads then k heads
  const long long total = tokens * per_token;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / per_token;
    int r = (int)(i % per_token);
    const bool is_k = r >= n_heads * vec_per_head;
    if (is_k) r -= n_heads * vec_per_head;
    const int head = r / vec_per_head, v = r % vec_per_head;
    T* base = (is_k ? k + t * ldk : q + t * ldq) + head * head_dim + v * 8;
    const int pos = pos_offset + (int)(t % seq_len);
    const T* cr = cos_t + (long long)pos * head_dim + v * 8;
    const T* sr = sin_t + (long long)pos * head_dim + v * 8;
    const uint4 lo = *reinterpret_cast<const uint4*>(base);
    const uint4 hi = *reinterpret_cast<const uint4*>(base + half);
    const uint4 c_lo = *reinterpret_cast<const uint4*>(cr), c_hi = *reinterpret_cast<const uint4*>(cr + half);
    const uint4 s_lo = *reinterpret_cast<const uint4*>(sr), s_hi = *reinterpret_cast<const uint4*>(sr + half);
    const T* a = reinterpret_cast<const T*>(&lo);
    const T* b = reinterpret_cast<const T*>(&hi);
    const T* cl = reinterpret_cast<const T*>(&c_lo);
    const T* ch = reinterpret_cast<const T*>(&c_hi);
    const T* sl = reinterpret_cast<const T*>(&s_lo);
    const T* sh = reinterpret_cast<const T*>(&s_hi);
    uint4 o_lo, o_hi;
    T* ol = reinterpret_cast<T*>(&o_lo);
    T* oh = reinterpret_cast<T*>(&o_hi);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x1 = to_f32<T>(a[e]), x2 = to_f32<T>(b[e]);
      // rotate_half(x) = cat(-x2, x1): out_lo = x1*cos_lo + (-x2)*sin_lo ; out_hi = x2*cos_hi + x1*sin_hi
      const T p1 = from_f32<T>(x1 * to_f32<T>(cl[e])), p2 = from_f32<T>(-x2 * to_f32<T>(sl[e]));
      const T p3 = from_f32<T>(x2 * to_f32<T>(ch[e])), p4 = from_f32<T>(x1 * to_f32<T>(sh[e]));
      ol[e] = from_f32<T>(to_f32<T>(p1) + to_f32<T>(p2));
      oh[e] = from_f32<T>(to_f32<T>(p3) + to_f32<T>(p4));
    }
    *reinterpret_cast<uint4*>(base) = o_lo;
    *reinterpret_cast<uint4*>(base + half) = o_hi;
  }
}

// dst[i] = src[index[i]]: one warp per row, four 128-bit loads in flight per lane
__global__ void __launch_bounds__(256)
gather_rows_kernel(const char* __restrict__ src, long long ld_src, char* __restrict__ dst, long long ld_dst,
                   const int* __restrict__ index, long long rows, int n_vec) {
  griddep_launch_dependents();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const Vec<16>* s = reinterpret_cast<const Vec<16>*>(src + (long long)index[row] * ld_src);
    Vec<16>* d = reinterpret_cast<Vec<16>*>(dst + row * ld_dst);
    int i = lane;
    for (; i + 96 < n_vec; i += 128) {
      const Vec<16> a = ld_stream(s + i), b = ld_stream(s + i + 32), c = ld_stream(s + i + 64), e = ld_stream(s + i + 96);
      d[i] = a;
      d[i + 32] = b;
      d[i + 64] = c;
      d[i + 96] = e;
    }
    for (; i < n_vec; i += 32) d[i] = ld_stream(s + i);
  }
}

}  // namespace mc

using namespace mc;

extern "C" int mc_gather_rows(const void* src, int64_t ld_src_bytes, void* dst, int64_t ld_dst_bytes, const int32_t* index,
                              int64_t rows, int row_bytes, mc_stream_t stream) {
  MC_REQUIRE(src && dst && index, "gather_rows: NULL pointer");
  MC_REQUIRE(rows >= 0 && row_bytes >= 16 && row_bytes % 16 == 0 && ld_src_bytes % 16 == 0 && ld_dst_bytes % 16 == 0,
             "gather_rows: row_bytes and leading dimensions must be multiples of 16 bytes");
  MC_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "gather_rows: pointers must be 16-byte aligned");
  if (rows == 0) return MC_OK;
  const int sms = sm_count();
  MC_REQUIRE(sms > 0, "no CUDA device");
  const int grid = (int)std::min<long long>((rows + 7) / 8, (long long)sms * 32);
  MC_CUDA_OK(launch_kernel(gather_rows_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const char*)src, (long long)ld_src_bytes,
                           (char*)dst, (long long)ld_dst_bytes, (const int*)index, (long long)rows, row_bytes / 16));
  return MC_OK;
}

extern "C" int mc_rmsnorm(const void* x, const void* weight, void* out, int64_t rows, int hidden, int64_t ldx, int64_t ldo,
                          float eps, int dtype, mc_stream_t stream) {
  MC_REQUIRE(x && weight && out, "rmsnorm: NULL pointer");
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "rmsnorm: dtype must be bf16 or fp16");
  MC_REQUIRE(rows >= 0 && hidden >= 8 && hidden % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "rmsnorm: hidden / ld must be multiples of 8");
  MC_REQUIRE((((uintptr_t)x | (uintptr_t)weight | (uintptr_t)out) & 15) == 0, "rmsnorm: pointers must be 16-byte aligned");
  if (rows == 0) return MC_OK;
  const int sms = sm_count();
  MC_REQUIRE(sms > 0, "no CUDA device");
  const int grid = (int)std::min<long long>((rows + 7) / 8, (long long)sms * 32);
  if (dtype == MC_BF16)
    MC_CUDA_OK(launch_kernel(rmsnorm_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)x,
                             (const __nv_bfloat16*)weight, (__nv_bfloat16*)out, (long long)rows, hidden, (long long)ldx, (long long)ldo, eps));
  else
    MC_CUDA_OK(launch_kernel(rmsnorm_kernel<__half>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const __half*)x, (const __half*)weight,
                             (__half*)out, (long long)rows, hidden, (long long)ldx, (long long)ldo, eps));
  return MC_OK;
}

extern "C" int mc_rope(void* q, void* k, const void* cos_table, const void* sin_table, int64_t tokens, int seq_len, int pos_offset,
                       int n_heads, int head_dim, int64_t ldq, int64_t ldk, int dtype, mc_stream_t stream) {
  MC_REQUIRE(q && k && cos_table && sin_table, "rope: NULL pointer");
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "rope: dtype must be bf16 or fp16");
  MC_REQUIRE(tokens >= 0 && seq_len >= 1 && pos_offset >= 0 && n_heads >= 1 && head_dim >= 16 && head_dim % 16 == 0,
             "rope: head_dim must be a multiple of 16");
  MC_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0, "rope: leading dimensions must be multiples of 8");
  MC_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)cos_table | (uintptr_t)sin_table) & 15) == 0, "rope: pointers must be 16-byte aligned");
  if (tokens == 0) return MC_OK;
  const int sms = sm_count();
  MC_REQUIRE(sms > 0, "no CUDA device");
  const long long total = tokens * 2 * n_heads * (head_dim / 16);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
  if (dtype == MC_BF16)
    rope_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)q, (__nv_bfloat16*)k, (const __nv_bfloat16*)cos_table,
                                                                      (const __nv_bfloat16*)sin_table, tokens, seq_len, pos_offset, n_heads, head_dim, ldq, ldk);
  else
    rope_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((__half*)q, (__half*)k, (const __half*)cos_table, (const __half*)sin_table,
                                                                tokens, seq_len, pos_offset, n_heads, head_dim, ldq, ldk);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}
