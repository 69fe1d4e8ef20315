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
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes = kAttHalfBytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
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


// =====================================================================================================================
// v2: two query tiles per CTA in ping-pong, P and O resident in TMEM.
//
// One CTA = 256 query rows (tiles A and B of 128 rows) of one (sequence, head).  TMEM (512 columns): S_A | S_B | O_A | O_B,
// 128 fp32 columns each.  The MMA thread issues, per key tile i:   S_A = Q_A K_i^T,  O_B += P_B(i-1) V_{i-1},  S_B = Q_B K_i^T,
// O_A += P_A(i) V_i — so while the softmax warps of one tile work, the tensor core runs the two MMAs of the other tile.
//   warp 0      TMA producer: Q_A, Q_B once; K_0, V_0, K_1, V_1, ... through ONE ring of five 32 KB stages
//   warp 1      MMA issuer (one thread) + TMEM allocation.  P is the A operand FROM TMEM (tcgen05.mma with [tmem] A: 16-bit
//               pairs, two keys per 32-bit column, overwriting the first 64 columns of the score tile it came from), V the
//               MN-major B operand from shared memory, O accumulates in TMEM over the whole key loop
//   warps 4-7   softmax of tile A, warps 8-11 of tile B (setmaxnreg moves warpgroup 0's registers to them): thread = one query row, all 128 keys of the tile in registers
//               (no cross-thread reduction).  The running maximum is LAZY: it only moves (and O / l are only rescaled, by
//               this same thread, in TMEM) when the tile's maximum exceeds it by more than 2^8, so P stays <= 256 and the
//               rescale is rare; exp2 runs partly on the MUFU pipe (ex2.approx) and partly as a Cody-Waite split + cubic
//               on the FMA pipe (packed f32x2 FFMA2 / FADD2), because at 16 ex2 per clock and SM the MUFU pipe alone takes
//               as long as the tile's MMAs.  Ordering of O: tcgen05.commit on s_full(i) covers every earlier MMA, so when a
//               softmax thread sees S(i) its O already holds P(i-1) V(i-1); the P V(i) MMA is only issued after p_full(i).
// =====================================================================================================================
constexpr int kA2Threads = 128 + 256;  // warpgroup 0: TMA warp, MMA warp, two idle warps; warpgroups 1, 2: softmax of tile A, B
constexpr int kA2Stages = 5;

struct Att2Smem {
  static constexpr int Q = 0;                                  // Q_A, Q_B
  static constexpr int KV = Q + 2 * kAttTileBytes;             // ring
  static constexpr int BAR = KV + kA2Stages * kAttTileBytes;   // 224 KB of tiles
  static constexpr int N_BAR = 1 + 2 * kA2Stages + 6;
  static constexpr int TOTAL = BAR + N_BAR * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// packed fp32 pairs (FFMA2 / FADD2 on sm_100)
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// 2^x for x <= ~8 on the FMA pipe: x = n + f (n = floor(x) through a round-down add of 1.5 * 2^23, exact), 2^f by the cubic
// 1 + f (c1 + f (c2 + f c3)) (p(0) = 1, p(1) = 2, max relative error 1.03e-4 — below the 16-bit rounding P gets anyway), and n
// added to the exponent field.  x is clamped at -126 so the exponent never wraps (2^-126 rounds to nothing in the row sum).
__device__ __forceinline__ void exp2_poly2(uint64_t x, float& p0, float& p1) {
  float x0, x1;
  upk2(x, x0, x1);
  const uint64_t xc = pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t t = add2_rm(xc, pk2(12582912.0f, 12582912.0f));
  const uint64_t fl = add2(t, pk2(-12582912.0f, -12582912.0f));
  const uint64_t f = fma2(fl, pk2(-1.0f, -1.0f), xc);
  uint64_t p = fma2(f, pk2(0.07826796919107437f, 0.07826796919107437f), pk2(0.226307675242424f, 0.226307675242424f));
  p = fma2(p, f, pk2(0.6954243183135986f, 0.6954243183135986f));
  p = fma2(p, f, pk2(1.0f, 1.0f));
  float t0, t1, q0, q1;
  upk2(t, t0, t1);
  upk2(p, q0, q1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}

// PP = pairs out of every 4 whose exp2 runs on the FMA pipe (0 = all on MUFU); LAZY = 0 rescales on every new maximum
template <bool F16, int PP, bool LAZY>
__global__ void __launch_bounds__(kA2Threads, 1) attention2_kernel(const __grid_constant__ AttParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Att2Smem::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                     // [5]
  uint64_t* kv_empty = bars + 1 + kA2Stages;        // [5]
  uint64_t* s_full = bars + 1 + 2 * kA2Stages;      // [2]  MMA -> softmax of tile t: S_t(i) is in TMEM (and O_t holds every earlier P V)
  uint64_t* p_full = s_full + 2;                    // [2]  softmax -> MMA: P_t(i) is in TMEM, O_t rescaled
  uint64_t* o_full = p_full + 2;                    // [2]  MMA -> softmax: the last P V of tile t has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Att2Smem::N_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heaviest query blocks (most key tiles below the diagonal) are scheduled first
  const int qb = (int)gridDim.x - 1 - (int)blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int q0 = qb * 2 * kAttTile;
  const bool valid_b = q0 + kAttTile < P.seq_len;
  const int n_a = 2 * qb + 1;                 // causal: tile A (query tile 2 qb) sees key tiles 0 .. 2 qb
  const int n_b = valid_b ? 2 * qb + 2 : 0;   //         tile B one more
  const int n_max = max(n_a, n_b);
  const long long row0 = (long long)seq * P.seq_len;
  const int col0 = head * kAttTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmQ);
    tma_prefetch_desc(&P.tmK);
    tma_prefetch_desc(&P.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kA2Stages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&o_full[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // whole warp converged, one elected lane issues the copies
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * kAttTileBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        uint8_t* qd = smem + Att2Smem::Q + t * kAttTileBytes;
        tma_load_2d(&P.tmQ, q_full, qd, col0, (int)(row0 + q0 + t * kAttTile));
        tma_load_2d(&P.tmQ, q_full, qd + kAttHalfBytes, col0 + 64, (int)(row0 + q0 + t * kAttTile));
      }
    }
    __syncwarp();
    for (int u = 0; u < 2 * n_max; ++u) {  // K_0, V_0, K_1, V_1, ...
      const int st = u % kA2Stages;
      const uint32_t ph = (uint32_t)((u / kA2Stages) & 1);
      const int krow = (int)(row0 + (long long)(u >> 1) * kAttTile);
      mbar_wait(&kv_empty[st], ph ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(&kv_full[st], kAttTileBytes);
        uint8_t* d = smem + Att2Smem::KV + st * kAttTileBytes;
        const CUtensorMap* tm = (u & 1) ? &P.tmV : &P.tmK;
        tma_load_2d(tm, &kv_full[st], d, col0, krow);
        tma_load_2d(tm, &kv_full[st], d + kAttHalfBytes, col0 + 64, krow);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // the whole warp walks the schedule (converged); one elected lane issues the tcgen05 instructions
    const uint32_t q_addr = smem_u32(smem + Att2Smem::Q), kv_addr = smem_u32(smem + Att2Smem::KV);
    auto wait_kv = [&](int u) {
      mbar_wait(&kv_full[u % kA2Stages], (uint32_t)((u / kA2Stages) & 1));
      tc_fence_after();
    };
    auto issue_qk = [&](int t, int i) {  // S_t = Q_t K_i^T
      const uint32_t k_addr = kv_addr + (uint32_t)(((2 * i) % kA2Stages) * kAttTileBytes);
      const uint32_t qa = q_addr + (uint32_t)(t * kAttTileBytes);
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t a_desc = umma_smem_desc(qa + kb * kAttHalfBytes), b_desc = umma_smem_desc(k_addr + kb * kAttHalfBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + (uint32_t)(t * kAttTile), a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), P.idesc_qk,
                     (kb | k) ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    // O_t (+)= P_t(i) V_i, P from TMEM (8 columns = 16 keys per MMA); then the listed barriers are committed
    auto issue_pv = [&](int t, int i, uint64_t* bar0, uint64_t* bar1) {
      mbar_wait(&p_full[t], (uint32_t)(i & 1));
      tc_fence_after();
      const uint32_t v_addr = kv_addr + (uint32_t)(((2 * i + 1) % kA2Stages) * kAttTileBytes);
      const uint32_t p_tmem = tmem_base + (uint32_t)(t * kAttTile), o_tmem = tmem_base + 256u + (uint32_t)(t * kAttTile);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ts(o_tmem, p_tmem + (uint32_t)(kk * 8), umma_smem_desc_mn(v_addr + kk * 2048), P.idesc_pv, (i > 0 || kk > 0) ? 1u : 0u);
        if (bar0) umma_commit(bar0);
        if (bar1) umma_commit(bar1);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    for (int i = 0; i < n_max; ++i) {
      wait_kv(2 * i);
      if (i < n_a) issue_qk(0, i);
      if (i >= 1 && n_b > 0) {  // i - 1 < n_b always holds inside the loop
        wait_kv(2 * i - 1);
        issue_pv(1, i - 1, &kv_empty[(2 * i - 1) % kA2Stages], nullptr);
      }
      if (i < n_b) issue_qk(1, i);
      if (elect_one()) umma_commit(&kv_empty[(2 * i) % kA2Stages]);  // K_i: both score MMAs have been issued
      __syncwarp();
      if (i < n_a) {
        wait_kv(2 * i + 1);
        issue_pv(0, i, n_b == 0 ? &kv_empty[(2 * i + 1) % kA2Stages] : nullptr, i == n_a - 1 ? &o_full[0] : nullptr);
      }
    }
    if (n_b > 0) {
      wait_kv(2 * n_b - 1);
      issue_pv(1, n_b - 1, &kv_empty[(2 * n_b - 1) % kA2Stages], &o_full[1]);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;      // 0: tile A, 1: tile B
    const int q = warp & 3;             // TMEM lane quadrant (hardware: warp id % 4)
    const int r = q * 32 + lane;        // query row inside the tile = TMEM lane
    const int n_t = t == 0 ? n_a : n_b;
    const int qt = 2 * qb + t;          // this tile's diagonal key tile
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + (uint32_t)(t * kAttTile);
    const uint32_t o_tmem = tmem_base + lane_off + 256u + (uint32_t)(t * kAttTile);
    const uint64_t scale2 = pk2(P.scale_log2, P.scale_log2);
    float m_used = -INFINITY, l_run = 0.0f;
    for (int i = 0; i < n_t; ++i) {
      mbar_wait(&s_full[t], (uint32_t)(i & 1));
      tc_fence_after();
      uint32_t sv[kAttTile];
#pragma unroll
      for (int c = 0; c < kAttTile; c += 32) tmem_ld_32x32(s_tmem + (uint32_t)c, *reinterpret_cast<uint32_t(*)[32]>(&sv[c]));
      tmem_ld_wait();
      if (i == qt) {  // warp-uniform: the diagonal tile masks keys after the query
#pragma unroll
        for (int c = 0; c < kAttTile; ++c)
          if (c > r) sv[c] = 0xff800000u;  // -inf
      }
      float mx8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = max3(__uint_as_float(sv[c]), __uint_as_float(sv[c + 8]), __uint_as_float(sv[c + 16]));
#pragma unroll
      for (int c = 24; c < kAttTile - 8; c += 16)
#pragma unroll
        for (int e = 0; e < 8; ++e) mx8[e] = max3(mx8[e], __uint_as_float(sv[c + e]), __uint_as_float(sv[c + 8 + e]));
#pragma unroll
      for (int e = 0; e < 8; ++e) mx8[e] = fmaxf(mx8[e], __uint_as_float(sv[kAttTile - 8 + e]));
      const float mx = max3(max3(mx8[0], mx8[1], mx8[2]), max3(mx8[3], mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]));
      // key 0 is visible to every row, so the maximum is finite from the first tile on
      const float m_tile = mx * P.scale_log2;
      const bool grow = LAZY ? (m_tile > m_used + 8.0f) : (m_tile > m_used);
      float alpha = 1.0f;
      if (grow) {
        alpha = ex2_approx(m_used - m_tile);  // first tile: exp2(-inf) = 0
        m_used = m_tile;
        l_run *= alpha;
      }
      if (i > 0 && __any_sync(0xffffffffu, grow)) {
        // O_t holds P(0..i-1) V: s_full(i) was committed after those MMAs.  Rescale this thread's row in place.
        const uint64_t a2 = pk2(alpha, alpha);
#pragma unroll
        for (int c = 0; c < kAttTile; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(o_tmem + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float a, b;
            upk2(mul2(pk2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), a2), a, b);
            v[e] = __float_as_uint(a);
            v[e + 1] = __float_as_uint(b);
          }
          tmem_st_32x32(o_tmem + (uint32_t)c, v);
        }
      }
      // p = exp2(s * scale - m): fp32 row sum, 16-bit pairs into the first 64 columns of the score tile
      const uint64_t negm2 = pk2(-m_used, -m_used);
      uint64_t l2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int c = 0; c < kAttTile; c += 32) {
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const uint64_t x = fma2(pk2(__uint_as_float(sv[c + 2 * k]), __uint_as_float(sv[c + 2 * k + 1])), scale2, negm2);
          float p0, p1;
          if ((k & 3) < PP) {
            exp2_poly2(x, p0, p1);
          } else {
            float x0, x1;
            upk2(x, x0, x1);
            p0 = ex2_approx(x0);
            p1 = ex2_approx(x1);
          }
          l2[k & 3] = add2(l2[k & 3], pk2(p0, p1));
          w[k] = pack2<F16>(p0, p1);
        }
        tmem_st_32x16(s_tmem + (uint32_t)(c >> 1), w);
      }
      float la, lb;
      upk2(add2(add2(l2[0], l2[1]), add2(l2[2], l2[3])), la, lb);
      l_run += la + lb;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    if (n_t > 0) {
      // all P V of this tile have completed: normalise by the row sum, store
      const float inv = 1.0f / l_run;
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const int tok = q0 + t * kAttTile + r;
      const bool ok = tok < P.seq_len;
      const long long tt = row0 + tok;
      const long long orow = (ok && P.out_rowmap) ? (long long)P.out_rowmap[tt] : tt;
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(P.out) + (orow * P.ld_out + col0) * 2);
#pragma unroll
      for (int c = 0; c < kAttTile; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(o_tmem + (uint32_t)c, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 8)
            dst[(c + e) >> 3] = make_uint4(pack2<F16>(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv),
                                           pack2<F16>(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv),
                                           pack2<F16>(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv),
                                           pack2<F16>(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}


// =====================================================================================================================
// v3: v2 with 64-key steps and DOUBLE-BUFFERED scores, so the score MMA of step j + 1 never waits for the softmax of step j.
//
// ncu on v2 (profiles/r02_attention.txt): tensor pipe 47 %, and each softmax warp spends half its time waiting for its next
// score tile — per query tile the chain  S = Q K^T -> softmax -> O += P V -> next S  is strictly serial (P overwrites the only
// score buffer), so every iteration pays softmax + two MMAs + ~1300 cycles of commit / wake-up / tcgen05.ld latencies.
// Here TMEM holds S_A[2] | S_B[2] (64 columns each) | O_A | O_B: the MMA thread runs one step ahead with the scores
// (Q_A K_{j+1}^T, Q_B K_{j+1}^T, then P_A(j) V_j, P_B(j) V_j), both softmax warpgroups always have a score tile waiting and
// run back to back (two active warps per scheduler instead of one), and the chain per tile shrinks to the softmax itself.
// P_t(j) (64 keys = 32 TMEM columns of 16-bit pairs) overwrites the first half of the score buffer it came from.
// A softmax thread signals p_full(j) only after pv_done(j - 1): it may not get two phases ahead of the MMA thread's parity
// wait, and that wait is also what makes an (occasional, lazy) rescale of O in TMEM safe.
// =====================================================================================================================
constexpr int kA3Keys = 64;                          // keys per step
constexpr int kA3KvBytes = kA3Keys * kAttTile * 2;   // one K or V step tile: 16 KB (two [64 x 64] halves)
constexpr int kA3KStages = 5, kA3VStages = 5;

struct Att3Smem {
  static constexpr int Q = 0;                                   // Q_A, Q_B
  static constexpr int K = Q + 2 * kAttTileBytes;
  static constexpr int V = K + kA3KStages * kA3KvBytes;
  static constexpr int BAR = V + kA3VStages * kA3KvBytes;       // 224 KB of tiles
  static constexpr int N_BAR = 1 + 2 * kA3KStages + 2 * kA3VStages + 8;
  static constexpr int TOTAL = BAR + N_BAR * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

template <bool F16, int PP, bool LAZY>
__global__ void __launch_bounds__(kA2Threads, 1) attention3_kernel(const __grid_constant__ AttParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Att3Smem::BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kA3KStages;
  uint64_t* v_full = k_empty + kA3KStages;
  uint64_t* v_empty = v_full + kA3VStages;
  uint64_t* s_full = v_empty + kA3VStages;   // [2][2]  MMA -> softmax of tile t: S_t(j) is in TMEM buffer j & 1 (one barrier per buffer)
  uint64_t* p_full = s_full + 4;             // [2]  softmax -> MMA: P_t(j) is in TMEM (and O_t rescaled if the maximum grew)
  uint64_t* pv_done = p_full + 2;            // [2]  MMA -> softmax: O_t += P_t(j) V_j has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Att3Smem::N_BAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qb = (int)gridDim.x - 1 - (int)blockIdx.x, head = blockIdx.y, seq = blockIdx.z;  // heaviest query blocks first
  const int q0 = qb * 2 * kAttTile;
  const bool valid_b = q0 + kAttTile < P.seq_len;
  const int n_keys = (P.seq_len + kA3Keys - 1) / kA3Keys;   // steps that hold any key of the sequence
  const int n_a = min(4 * qb + 2, n_keys);                 // causal: tile A (rows q0 .. q0 + 127) sees keys < q0 + 128 = steps 0 .. 4 qb + 1
  const int n_b = valid_b ? min(4 * qb + 4, n_keys) : 0;   //         tile B two steps more
  const int n_max = max(n_a, n_b);
  const long long row0 = (long long)seq * P.seq_len;
  const int col0 = head * kAttTile;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmQ);
    tma_prefetch_desc(&P.tmK);
    tma_prefetch_desc(&P.tmV);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int st = 0; st < kA3KStages; ++st) {
      mbar_init(&k_full[st], 1);
      mbar_init(&k_empty[st], 1);
    }
    for (int st = 0; st < kA3VStages; ++st) {
      mbar_init(&v_full[st], 1);
      mbar_init(&v_empty[st], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[2 * t], 1);
      mbar_init(&s_full[2 * t + 1], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // TMA producer: whole warp converged, one elected lane issues.  Q_A, Q_B, then K_0, K_1, V_0, K_2, V_1, ...
      if (elect_one()) {
        mbar_expect_tx(q_full, 2 * kAttTileBytes);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          uint8_t* qd = smem + Att3Smem::Q + t * kAttTileBytes;
          tma_load_2d(&P.tmQ, q_full, qd, col0, (int)(row0 + q0 + t * kAttTile));
          tma_load_2d(&P.tmQ, q_full, qd + kAttHalfBytes, col0 + 64, (int)(row0 + q0 + t * kAttTile));
        }
      }
      __syncwarp();
      for (int j = 0; j <= n_max; ++j) {
        if (j < n_max) {  // K_j
          const int st = j % kA3KStages;
          mbar_wait(&k_empty[st], (uint32_t)(((j / kA3KStages) & 1) ^ 1));
          if (elect_one()) {
            mbar_expect_tx(&k_full[st], kA3KvBytes);
            uint8_t* d = smem + Att3Smem::K + st * kA3KvBytes;
            const int krow = (int)(row0 + (long long)j * kA3Keys);
            tma_load_2d(&P.tmK, &k_full[st], d, col0, krow);
            tma_load_2d(&P.tmK, &k_full[st], d + kA3KvBytes / 2, col0 + 64, krow);
          }
          __syncwarp();
        }
        if (j >= 1) {  // V_{j-1}
          const int jv = j - 1, st = jv % kA3VStages;
          mbar_wait(&v_empty[st], (uint32_t)(((jv / kA3VStages) & 1) ^ 1));
          if (elect_one()) {
            mbar_expect_tx(&v_full[st], kA3KvBytes);
            uint8_t* d = smem + Att3Smem::V + st * kA3KvBytes;
            const int krow = (int)(row0 + (long long)jv * kA3Keys);
            tma_load_2d(&P.tmV, &v_full[st], d, col0, krow);
            tma_load_2d(&P.tmV, &v_full[st], d + kA3KvBytes / 2, col0 + 64, krow);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // MMA issuer: whole warp converged, one elected lane issues
      const uint32_t q_addr = smem_u32(smem + Att3Smem::Q), k_addr0 = smem_u32(smem + Att3Smem::K), v_addr0 = smem_u32(smem + Att3Smem::V);
      auto issue_qk = [&](int t, int j) {  // S_t[j & 1] = Q_t K_j^T   (M 128, N 64, K 16 x 8)
        const uint32_t k_addr = k_addr0 + (uint32_t)((j % kA3KStages) * kA3KvBytes);
        const uint32_t qa = q_addr + (uint32_t)(t * kAttTileBytes);
        const uint32_t d_tmem = tmem_base + (uint32_t)(t * kAttTile + (j & 1) * kA3Keys);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t a_desc = umma_smem_desc(qa + kb * kAttHalfBytes), b_desc = umma_smem_desc(k_addr + kb * (kA3KvBytes / 2));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), P.idesc_qk, (kb | k) ? 1u : 0u);
          }
          umma_commit(&s_full[2 * t + (j & 1)]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int t, int j) {  // O_t (+)= P_t(j) V_j, P from TMEM: 8 columns = 16 keys per MMA
        mbar_wait(&p_full[t], (uint32_t)(j & 1));
        tc_fence_after();
        const uint32_t v_addr = v_addr0 + (uint32_t)((j % kA3VStages) * kA3KvBytes);
        const uint32_t p_tmem = tmem_base + (uint32_t)(t * kAttTile + (j & 1) * kA3Keys), o_tmem = tmem_base + 256u + (uint32_t)(t * kAttTile);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kA3Keys / 16; ++kk)
            umma_f16_ts(o_tmem, p_tmem + (uint32_t)(kk * 8), umma_smem_desc_mn(v_addr + kk * 2048, kA3KvBytes / 2), P.idesc_pv,
                        (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&pv_done[t]);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      if (n_b > 0) issue_qk(1, 0);
      if (elect_one()) umma_commit(&k_empty[0]);
      __syncwarp();
      for (int j = 0; j < n_max; ++j) {
        if (j + 1 < n_max) {  // scores of the next step: their buffers were released by the P V MMAs of step j - 1
          const int st = (j + 1) % kA3KStages;
          mbar_wait(&k_full[st], (uint32_t)((((j + 1) / kA3KStages)) & 1));
          tc_fence_after();
          if (j + 1 < n_a) issue_qk(0, j + 1);
          if (j + 1 < n_b) issue_qk(1, j + 1);
          if (elect_one()) umma_commit(&k_empty[st]);
          __syncwarp();
        }
        mbar_wait(&v_full[j % kA3VStages], (uint32_t)((j / kA3VStages) & 1));
        tc_fence_after();
        if (j < n_a) issue_pv(0, j);
        if (j < n_b) issue_pv(1, j);
        if (elect_one()) umma_commit(&v_empty[j % kA3VStages]);
        __syncwarp();
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;      // 0: tile A, 1: tile B
    const int q = warp & 3;             // TMEM lane quadrant (hardware: warp id % 4)
    const int r = q * 32 + lane;        // query row inside the tile = TMEM lane
    const int n_t = t == 0 ? n_a : n_b;
    const int qrow = q0 + t * kAttTile + r;   // query position inside the sequence
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + (uint32_t)(t * kAttTile);
    const uint32_t o_tmem = tmem_base + lane_off + 256u + (uint32_t)(t * kAttTile);
    const uint64_t scale2 = pk2(P.scale_log2, P.scale_log2);
    float m_used = -INFINITY, l_run = 0.0f;
    for (int j = 0; j < n_t; ++j) {
      const uint32_t sj = s_tmem + (uint32_t)((j & 1) * kA3Keys);
      mbar_wait(&s_full[2 * t + (j & 1)], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t sv[kA3Keys];
      tmem_ld_32x32(sj, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
      tmem_ld_32x32(sj + 32u, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
      tmem_ld_wait();
      const int key0 = j * kA3Keys;
      if (key0 + kA3Keys - 1 > q0 + t * kAttTile + q * 32) {  // warp-uniform: some key of the step lies after some query of this warp
#pragma unroll
        for (int c = 0; c < kA3Keys; ++c)
          if (key0 + c > qrow) sv[c] = 0xff800000u;  // -inf
      }
      float mx8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = max3(__uint_as_float(sv[c]), __uint_as_float(sv[c + 8]), __uint_as_float(sv[c + 16]));
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = max3(mx8[c], __uint_as_float(sv[c + 24]), __uint_as_float(sv[c + 32]));
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = max3(mx8[c], __uint_as_float(sv[c + 40]), __uint_as_float(sv[c + 48]));
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(mx8[c], __uint_as_float(sv[c + 56]));
      const float mx = max3(max3(mx8[0], mx8[1], mx8[2]), max3(mx8[3], mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]));
      // key 0 is visible to every row, so the running maximum is finite from the first step on; a later step may be
      // masked entirely for a row (mx = -inf): it neither moves the maximum nor adds to the sum
      const float m_tile = mx * P.scale_log2;
      const bool grow = LAZY ? (m_tile > m_used + 8.0f) : (m_tile > m_used);
      float alpha = 1.0f;
      if (grow) {
        alpha = ex2_approx(m_used - m_tile);  // first step: exp2(-inf) = 0
        m_used = m_tile;
        l_run *= alpha;
      }
      bool pv_waited = false;
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // rescale this thread's row of O in place; every earlier P V must have landed first
        mbar_wait(&pv_done[t], (uint32_t)((j - 1) & 1));
        tc_fence_after();
        pv_waited = true;
        const uint64_t a2 = pk2(alpha, alpha);
#pragma unroll
        for (int c = 0; c < kAttTile; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(o_tmem + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            float a, b;
            upk2(mul2(pk2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), a2), a, b);
            v[e] = __float_as_uint(a);
            v[e + 1] = __float_as_uint(b);
          }
          tmem_st_32x32(o_tmem + (uint32_t)c, v);
        }
      }
      // p = exp2(s * scale - m): fp32 row sum, 16-bit pairs into the first 32 columns of the score buffer
      const uint64_t negm2 = pk2(-m_used, -m_used);
      uint64_t l2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t w[kA3Keys / 2];
#pragma unroll
      for (int k = 0; k < kA3Keys / 2; ++k) {
        const uint64_t x = fma2(pk2(__uint_as_float(sv[2 * k]), __uint_as_float(sv[2 * k + 1])), scale2, negm2);
        float p0, p1;
        if ((k & 3) < PP) {
          exp2_poly2(x, p0, p1);
        } else {
          float x0, x1;
          upk2(x, x0, x1);
          p0 = ex2_approx(x0);
          p1 = ex2_approx(x1);
        }
        l2[k & 3] = add2(l2[k & 3], pk2(p0, p1));
        w[k] = pack2<F16>(p0, p1);
      }
      tmem_st_32x32(sj, w);
      float la, lb;
      upk2(add2(add2(l2[0], l2[1]), add2(l2[2], l2[3])), la, lb);
      l_run += la + lb;
      tmem_st_wait();
      if (j > 0 && !pv_waited) mbar_wait(&pv_done[t], (uint32_t)((j - 1) & 1));  // never two phases ahead of the MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    if (n_t > 0) {
      // all P V of this tile have completed: normalise by the row sum, store
      const float inv = 1.0f / l_run;
      mbar_wait(&pv_done[t], (uint32_t)((n_t - 1) & 1));
      tc_fence_after();
      const bool ok = qrow < P.seq_len;
      const long long tt = row0 + qrow;
      const long long orow = (ok && P.out_rowmap) ? (long long)P.out_rowmap[tt] : tt;
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(P.out) + (orow * P.ld_out + col0) * 2);
#pragma unroll
      for (int c = 0; c < kAttTile; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(o_tmem + (uint32_t)c, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 8)
            dst[(c + e) >> 3] = make_uint4(pack2<F16>(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv),
                                           pack2<F16>(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv),
                                           pack2<F16>(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv),
                                           pack2<F16>(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

}  // namespace mc

using namespace mc;

template <bool F16, int PP, bool LAZY>
static cudaError_t launch_attention2(const AttParams& P, dim3 grid, cudaStream_t stream, bool v3) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention2_kernel<F16, PP, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att2Smem::DYN_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention3_kernel<F16, PP, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att3Smem::DYN_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  if (v3)
    attention3_kernel<F16, PP, LAZY><<<grid, kA2Threads, Att3Smem::DYN_BYTES, stream>>>(P);
  else
    attention2_kernel<F16, PP, LAZY><<<grid, kA2Threads, Att2Smem::DYN_BYTES, stream>>>(P);
  return cudaGetLastError();
}

extern "C" int mc_attention_causal_tuned(const void* q, const void* k, const void* v, void* out, int64_t ld_qkv, int64_t ld_out,
                                         const int32_t* out_rowmap, int batch, int seq_len, int n_heads, int head_dim,
                                         float softmax_scale, int dtype, int tuning, mc_stream_t stream) {
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
  const int variant = tuning & 0xf;
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == 1) {  // v1: one query tile per CTA, P through shared memory, O in registers
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
      attention_kernel<true><<<grid, kAttThreads, AttSmem::DYN_BYTES, st>>>(P);
    else
      attention_kernel<false><<<grid, kAttThreads, AttSmem::DYN_BYTES, st>>>(P);
    MC_CUDA_OK(cudaGetLastError());
    return MC_OK;
  }
  // v2 / v3: bits 4-7 = pairs out of 4 whose exp2 runs on the FMA pipe + 1 (0 = default), bit 8 = rescale on every new maximum
  dim3 grid((seq_len + 2 * kAttTile - 1) / (2 * kAttTile), n_heads, batch);
  if (variant != 2) {  // v3 (default): 64-key steps, K / V maps with 64-row boxes
    rc = encode_operand(&P.tmK, k, tokens, (long long)n_heads * head_dim, ld_qkv, kA3Keys, dtype);
    if (rc == MC_OK) rc = encode_operand(&P.tmV, v, tokens, (long long)n_heads * head_dim, ld_qkv, kA3Keys, dtype);
    if (rc != MC_OK) return rc;
    P.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(kA3Keys >> 3) << 17) | ((unsigned)(kAttTile >> 4) << 24);
  }
  const int pp = ((tuning >> 4) & 0xf) ? ((tuning >> 4) & 0xf) - 1 : 2;
  const bool eager = (tuning >> 8) & 1;
  const bool f16 = dtype == MC_F16, v3 = variant != 2;
  cudaError_t e = cudaErrorInvalidValue;
#define MC_ATT2(PPV)                                                                                          \
  case PPV:                                                                                                   \
    e = f16 ? (eager ? launch_attention2<true, PPV, false>(P, grid, st, v3) : launch_attention2<true, PPV, true>(P, grid, st, v3))   \
            : (eager ? launch_attention2<false, PPV, false>(P, grid, st, v3) : launch_attention2<false, PPV, true>(P, grid, st, v3)); \
    break;
  switch (pp) {
    MC_ATT2(0)
    MC_ATT2(1)
    MC_ATT2(2)
    MC_ATT2(3)
    default:
      return fail(MC_ERR_INVALID, "attention: tuning 0x%x selects no kernel", tuning);
  }
#undef MC_ATT2
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "attention launch failed: %s", cudaGetErrorString(e));
  return MC_OK;
}

extern "C" int mc_attention_causal(const void* q, const void* k, const void* v, void* out, int64_t ld_qkv, int64_t ld_out,
                                   const int32_t* out_rowmap, int batch, int seq_len, int n_heads, int head_dim, float softmax_scale,
                                   int dtype, mc_stream_t stream) {
  return mc_attention_causal_tuned(q, k, v, out, ld_qkv, ld_out, out_rowmap, batch, seq_len, n_heads, head_dim, softmax_scale, dtype, 0,
                                   stream);
}
