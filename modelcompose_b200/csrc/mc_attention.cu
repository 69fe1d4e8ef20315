// Causal self-attention of the prefill on tcgen05 tensor cores (flash-style: scores never leave the SM).
//
// Replaces (reference path): modelcompose/model/language_model/multimodal_llama.py:295-312 — the eager attention of
// LocalLoraAttention.forward (QK^T / sqrt(d) + causal mask, fp32 softmax, PV), which materialises [B, heads, S, S] scores.
//
// One CTA = 128 query rows of one (sequence, head); key/value tiles of 128 keys up to the diagonal.
//   warp 0      TMA producer: Q once, then K_j / V_j into two-stage rings (separate barriers, so QK_j never waits for V_j)
//   warp 1      MMA issuer (one thread): S_j = Q K_j^T (K-major operands) into one of two TMEM score buffers, issued one tile
//               ahead of the softmax; PV_j = P_j V_j with V consumed MN-major straight from its [keys, d] tile
//               (warp 1 also owns the TMEM allocation)
//   warps 2-9   softmax, kAttSplit = 2 warps per TMEM lane quadrant: thread = (query row, 64-key half).  S_j is read from TMEM
//               once and kept in registers (row max -> exchange with the partner warp through shared memory -> exp2 / sum /
//               16-bit P_j into the 128B-swizzled K-major shared-memory tile the PV MMA reads); the running output half-row
//               lives in 64 fp32 registers: O = (O + PV_{j-1}) * alpha_j, so the tensor core never rescales an accumulator.
//               (kAttSplit = 4, sixteen softmax warps at 96 registers, measured slower: profiles/r01_attention.txt)
// Roofline: tensor pipe (bf16 / fp16 dense), bounded in practice by the softmax warps (exp2 on the MUFU pipe).
// head_dim is fixed at 128 (vicuna-7B); other head sizes keep the library call in model.py.
#include <algorithm>
#include <cmath>

#include "mc_tc.cuh"

namespace mc {

constexpr int kAttTile = 128;       // query rows per CTA, keys per step, head_dim
constexpr int kAttSplit = 2;          // softmax warps per TMEM lane quadrant: each thread owns 128 / kAttSplit keys of its row
constexpr int kAttCW = kAttTile / kAttSplit;
constexpr int kAttThreads = 64 + kAttSplit * 128;  // warp 0 TMA, warp 1 MMA + TMEM, then the softmax warps
constexpr int kAttHalfBytes = kAttTile * 64 * 2;  // one [128 x 64] 16-bit block = 16 KB
constexpr int kAttTileBytes = 2 * kAttHalfBytes;  // [128 x 128] = 32 KB

struct AttParams {
  CUtensorMap tmQ, tmK, tmV;  // 2-D [tokens, hidden] maps, box 128 rows x 64 columns, 128B swizzle
  void* out;                  // [tokens, ld_out]
  const int* out_rowmap;      // output row of token t (NULL = t): writes straight into the modality-major buffer order
  long long ld_out;
  int seq_len, n_heads, is_f16;
  float scale_log2;           // softmax scale * log2(e)
  unsigned int idesc_qk, idesc_pv;
};

struct AttSmem {
  static constexpr int Q = 0;
  static constexpr int K = Q + kAttTileBytes;      // 2 stages
  static constexpr int V = K + 2 * kAttTileBytes;  // 2 stages
  static constexpr int P = V + 2 * kAttTileBytes;
  static constexpr int BAR = P + kAttTileBytes;    // 192 KB of tiles
  static constexpr int N_BAR = 16;
  static constexpr int XCH = BAR + N_BAR * 8 + 16;  // per-row partials of the column parts: float [3][kAttSplit][128]
  static constexpr int TOTAL = XCH + 3 * kAttSplit * kAttTile * 4;  // (row max, double-buffered by tile parity; row sum)
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

// MN-major B operand (V tile: rows = keys = K dimension, 64 head-dim columns per 128-byte row, two column halves 16 KB apart):
// canonical SW128 MN-major layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> LBO = distance between the two
// 64-wide head-dim halves, SBO = 8 keys x 128 B.
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(kAttHalfBytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void pair_barrier(int q) {  // the softmax warps of TMEM lane quadrant q
  asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "n"(kAttSplit * 32) : "memory");
}
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (F16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <bool F16>
__global__ void __launch_bounds__(kAttThreads, 1) attention_kernel(const __grid_constant__ AttParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttSmem::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]  MMA -> softmax: scores of a tile are in TMEM
  uint64_t* s_empty = bars + 11;  // [2]  softmax -> MMA: the score buffer may be overwritten
  uint64_t* p_full = bars + 13;   //      softmax -> MMA: P_j is in shared memory
  uint64_t* pv_full = bars + 14;  //      MMA -> softmax: P_j V_j is in TMEM (and the P tile is free again)
  uint64_t* pv_empty = bars + 15; //      softmax -> MMA: the PV buffer may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + AttSmem::N_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heaviest query tiles (most key tiles below the diagonal) are scheduled first
  const int qt = (int)gridDim.x - 1 - (int)blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int q0 = qt * kAttTile;
  const int n_kv = qt + 1;  // causal: key tiles 0 .. qt
  const long long row0 = (long long)seq * P.seq_len;
  const int col0 = head * kAttTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmQ);
    tma_prefetch_desc(&P.tmK);
    tma_prefetch_desc(&P.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4 * kAttSplit);
    }
    mbar_init(p_full, 4 * kAttSplit);
    mbar_init(pv_full, 1);
    mbar_init(pv_empty, 4 * kAttSplit);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s[2] = {tmem_base, tmem_base + 128u};
  const uint32_t tmem_pv = tmem_base + 256u;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kAttTileBytes);
      tma_load_2d(&P.tmQ, q_full, smem + AttSmem::Q, col0, (int)(row0 + q0));
      tma_load_2d(&P.tmQ, q_full, smem + AttSmem::Q + kAttHalfBytes, col0 + 64, (int)(row0 + q0));
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        const int krow = (int)(row0 + (long long)j * kAttTile);
        mbar_wait(&k_empty[st], ph ^ 1u);
        mbar_expect_tx(&k_full[st], kAttTileBytes);
        uint8_t* kd = smem + AttSmem::K + st * kAttTileBytes;
        tma_load_2d(&P.tmK, &k_full[st], kd, col0, krow);
        tma_load_2d(&P.tmK, &k_full[st], kd + kAttHalfBytes, col0 + 64, krow);
        mbar_wait(&v_empty[st], ph ^ 1u);
        mbar_expect_tx(&v_full[st], kAttTileBytes);
        uint8_t* vd = smem + AttSmem::V + st * kAttTileBytes;
        tma_load_2d(&P.tmV, &v_full[st], vd, col0, krow);
        tma_load_2d(&P.tmV, &v_full[st], vd + kAttHalfBytes, col0 + 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t q_addr = smem_u32(smem + AttSmem::Q), p_addr = smem_u32(smem + AttSmem::P);
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1);
        mbar_wait(&k_full[st], ph);
        mbar_wait(&s_empty[st], ph ^ 1u);  // the softmax has finished with S_{j-2}
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + AttSmem::K + st * kAttTileBytes);
        uint32_t accumulate = 0;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a_desc = umma_smem_desc(q_addr + kb * kAttHalfBytes), b_desc = umma_smem_desc(k_addr + kb * kAttHalfBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(tmem_s[st], a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), P.idesc_qk, accumulate);
            accumulate = 1;
          }
        }
        umma_commit(&s_full[st]);
        umma_commit(&k_empty[st]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_qk(j + 1);
        const int st = j & 1;
        const uint32_t ph = (uint32_t)((j >> 1) & 1), jph = (uint32_t)(j & 1);
        mbar_wait(&v_full[st], ph);
        mbar_wait(p_full, jph);
        mbar_wait(pv_empty, jph ^ 1u);  // the softmax has folded PV_{j-1} into its registers
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + AttSmem::V + st * kAttTileBytes);
        uint32_t accumulate = 0;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 keys per MMA: P advances 32 B inside its 64-key block, V advances 16 rows
          const uint64_t a_desc = umma_smem_desc(p_addr + (kk >> 2) * kAttHalfBytes) + (uint64_t)(2 * (kk & 3));
          const uint64_t b_desc = umma_smem_desc_mn(v_addr + kk * 2048);
          umma_f16(tmem_pv, a_desc, b_desc, P.idesc_pv, accumulate);
          accumulate = 1;
        }
        umma_commit(pv_full);
        umma_commit(&v_empty[st]);
      }
    }
  } else {
    constexpr int CW = kAttCW;
    const int q = warp & 3;             // TMEM lane quadrant (hardware: warp id % 4)
    const int part = (warp - 2) >> 2;   // which CW keys of the score tile / which CW head-dim columns of the output
    const int r = q * 32 + lane;        // query row inside the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t col_off = (uint32_t)(part * CW);
    float* xch = reinterpret_cast<float*>(smem + AttSmem::XCH);
    float o[CW];
#pragma unroll
    for (int i = 0; i < CW; ++i) o[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;
    // this thread's CW keys of its row of the P tile: 64-key block, then 16-byte chunks (swizzled per row)
    uint8_t* p_blk = smem + AttSmem::P + (col_off >> 6) * kAttHalfBytes + r * 128;
    const int chunk0 = (int)(col_off & 63u) >> 3;
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      const bool diag = j == qt;
      mbar_wait(&s_full[st], ph);
      tc_fence_after();
      // scores of this thread's keys: TMEM -> registers, once; the buffer is free for QK_{j+2} right away
      uint32_t sv[CW];
#pragma unroll
      for (int c = 0; c < CW; c += 32)
        tmem_ld_32x32(tmem_s[st] + lane_off + col_off + (uint32_t)c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[st]);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent chains
      if (diag) {  // warp-uniform: only the diagonal tile pays for the causal comparison
#pragma unroll
        for (int i = 0; i < CW; ++i)
          if ((int)col_off + i <= r) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
      } else {
#pragma unroll
        for (int i = 0; i < CW; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      float* xm = xch + (j & 1) * kAttSplit * kAttTile;  // double-buffered: a partner may still be reading the previous tile's slots
      xm[part * kAttTile + r] = mx;
      pair_barrier(q);
#pragma unroll
      for (int pp = 0; pp < kAttSplit; ++pp) mx = fmaxf(mx, xm[pp * kAttTile + r]);
      const float m_new = fmaxf(m_run, mx * P.scale_log2);  // key 0 is always visible: finite from the first tile on
      const float alpha = ex2_approx(m_run - m_new);
      m_run = m_new;
      // p = exp2(s * scale - m): fp32 row sum, 16-bit packed pairs kept in registers until the P tile is free
      float l4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      uint32_t w[CW / 2];
      if (diag) {
#pragma unroll
        for (int i = 0; i < CW; i += 2) {
          float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), P.scale_log2, -m_new));
          float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), P.scale_log2, -m_new));
          if ((int)col_off + i > r) p0 = 0.0f;
          if ((int)col_off + i + 1 > r) p1 = 0.0f;
          l4[(i >> 1) & 3] += p0 + p1;
          w[i >> 1] = pack2<F16>(p0, p1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < CW; i += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[i]), P.scale_log2, -m_new));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), P.scale_log2, -m_new));
          l4[(i >> 1) & 3] += p0 + p1;
          w[i >> 1] = pack2<F16>(p0, p1);
        }
      }
      const float l_add = (l4[0] + l4[1]) + (l4[2] + l4[3]);
      // fold the previous tile's P V into the running part of the row, then rescale: O = (O + PV_{j-1}) * alpha
      if (j > 0) {
        const bool rescale = __any_sync(0xffffffffu, alpha != 1.0f);
        mbar_wait(pv_full, (uint32_t)((j - 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < CW; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_pv + lane_off + col_off + (uint32_t)c, v);
          tmem_ld_wait();
          if (rescale) {
#pragma unroll
            for (int i = 0; i < 32; ++i) o[c + i] = (o[c + i] + __uint_as_float(v[i])) * alpha;
          } else {  // no row of this warp moved its maximum: alpha is exactly 1
#pragma unroll
            for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pv_empty);
      }
      // the P tile is free (PV_{j-1} has completed): 16-bit P into the swizzled K-major tile, 16-byte chunks of 8 keys
#pragma unroll
      for (int c = 0; c < CW / 8; ++c)
        *reinterpret_cast<uint4*>(p_blk + (((chunk0 + c) ^ (r & 7)) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
      l_run = l_run * alpha + l_add;
      fence_proxy_async_smem();  // the P tile was written through the generic proxy, the MMA reads it through the async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // last tile's P V, normalise by the full row sum (all parts), store
    float* xl = xch + 2 * kAttSplit * kAttTile;
    xl[part * kAttTile + r] = l_run;
    pair_barrier(q);
    float l_tot = 0.0f;
#pragma unroll
    for (int pp = 0; pp < kAttSplit; ++pp) l_tot += xl[pp * kAttTile + r];
    const float inv = 1.0f / l_tot;
    mbar_wait(pv_full, (uint32_t)((n_kv - 1) & 1));
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < CW; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_pv + lane_off + col_off + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c + i] = (o[c + i] + __uint_as_float(v[i])) * inv;
    }
    const int tok = q0 + r;
    if (tok < P.seq_len) {
      const long long t = row0 + tok;
      const long long orow = P.out_rowmap ? (long long)P.out_rowmap[t] : t;
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(P.out) + (orow * P.ld_out + col0 + (int)col_off) * 2);
#pragma unroll
      for (int c = 0; c < CW; c += 8)
        dst[c >> 3] = make_uint4(pack2<F16>(o[c], o[c + 1]), pack2<F16>(o[c + 2], o[c + 3]), pack2<F16>(o[c + 4], o[c + 5]),
                                 pack2<F16>(o[c + 6], o[c + 7]));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace mc

using namespace mc;

extern "C" int mc_attention_causal(const void* q, const void* k, const void* v, void* out, int64_t ld_qkv, int64_t ld_out,
                                   const int32_t* out_rowmap, int batch, int seq_len, int n_heads, int head_dim, float softmax_scale,
                                   int dtype, mc_stream_t stream) {
  MC_REQUIRE(q && k && v && out, "attention: NULL pointer");
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "attention: dtype must be bf16 or fp16");
  MC_REQUIRE(head_dim == kAttTile, "attention: head_dim must be %d", kAttTile);
  MC_REQUIRE(batch >= 1 && seq_len >= 1 && n_heads >= 1, "attention: batch, seq_len and n_heads must be positive");
  MC_REQUIRE(ld_qkv >= (int64_t)n_heads * head_dim && ld_out >= (int64_t)n_heads * head_dim && ld_qkv % 8 == 0 && ld_out % 8 == 0,
             "attention: leading dimensions must cover n_heads * head_dim and be multiples of 8");
  MC_REQUIRE((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0, "attention: pointers must be 16-byte aligned");
  const long long tokens = (long long)batch * seq_len;
  MC_REQUIRE(tokens < (1LL << 31), "attention: too many tokens");
  AttParams P;
  memset(&P, 0, sizeof(P));
  int rc = encode_operand(&P.tmQ, q, tokens, (long long)n_heads * head_dim, ld_qkv, kAttTile, dtype);
  if (rc == MC_OK) rc = encode_operand(&P.tmK, k, tokens, (long long)n_heads * head_dim, ld_qkv, kAttTile, dtype);
  if (rc == MC_OK) rc = encode_operand(&P.tmV, v, tokens, (long long)n_heads * head_dim, ld_qkv, kAttTile, dtype);
  if (rc != MC_OK) return rc;
  P.out = out;
  P.out_rowmap = out_rowmap;
  P.ld_out = ld_out;
  P.seq_len = seq_len;
  P.n_heads = n_heads;
  P.is_f16 = dtype == MC_F16;
  P.scale_log2 = softmax_scale * 1.4426950408889634f;
  // instruction descriptors: D = F32, A/B = bf16 / fp16, N = 128 (>> 3 at bit 17), M = 128 (>> 4 at bit 24);
  // bit 16 = B operand MN-major (the V tile is [keys, head_dim] with head_dim contiguous)
  const unsigned int fmt = dtype == MC_F16 ? 0u : 1u;
  P.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(kAttTile >> 3) << 17) | ((unsigned)(kAttTile >> 4) << 24);
  P.idesc_pv = P.idesc_qk | (1u << 16);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    MC_CUDA_OK(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem::DYN_BYTES));
    MC_CUDA_OK(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttSmem::DYN_BYTES));
    configured[dev] = true;
  }
  dim3 grid((seq_len + kAttTile - 1) / kAttTile, n_heads, batch);
  if (dtype == MC_F16)
    attention_kernel<true><<<grid, kAttThreads, AttSmem::DYN_BYTES, (cudaStream_t)stream>>>(P);
  else
    attention_kernel<false><<<grid, kAttThreads, AttSmem::DYN_BYTES, (cudaStream_t)stream>>>(P);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}
