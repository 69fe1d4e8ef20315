// Causal self-attention of the prefill on tcgen05 tensor cores (flash-style: scores never leave the SM).
//
// Replaces (reference path): modelcompose/model/language_model/multimodal_llama.py:295-312 — the eager attention of
// LocalLoraAttention.forward (QK^T / sqrt(d) + causal mask, fp32 softmax, PV), which materialises [B, heads, S, S] scores.
//
// One CTA = 256 query rows (two 128-row tiles A and B) of one (sequence, head), walking 64-key steps up to the diagonal.
//   warp 0      TMA producer: Q_A, Q_B once; K and V steps through two rings of five 16 KB stages
//   warp 1      score MMAs of both tiles (one elected lane issues; also owns the TMEM allocation)
//   warp 2      P V MMAs of both tiles: P is the A operand FROM TMEM, V the MN-major B operand straight from its [keys, d] tile,
//               O accumulates in TMEM over the whole key loop
//   warps 4-7   softmax of tile A, warps 8-11 of tile B (setmaxnreg moves warpgroup 0's registers to them): thread = one
//               query row, all keys of the step in registers, no cross-thread reduction.  The running maximum is LAZY (it
//               only moves — and O / l are only rescaled, by the same thread, in TMEM — when a step's maximum exceeds it by
//               more than 2^8), and a share of the exp2 can run as a Cody-Waite split + cubic on the FMA pipe (FFMA2 / FADD2).
// Roofline: tensor pipe (bf16 / fp16 dense); what bounds it in practice, and every variant measured on the way, is in
// profiles/r02_attention.txt and DESIGN.md §4.6.  head_dim is fixed at 128 (vicuna-7B).
#include <algorithm>
#include <atomic>
#include <cmath>

#include "mc_tc.cuh"

namespace mc {

constexpr int kAttTile = 128;                     // query rows per tile, head_dim
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
  int n_items;                // (sequence, head, 256-row query block) work items
  unsigned int* counter;      // device: next item to hand out (zeroed before the launch)
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
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (F16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

constexpr int kAttThreads = 128 + 256;  // warpgroup 0: TMA warp, two MMA warps, one idle warp; warpgroups 1, 2: softmax of tile A, B

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

// =====================================================================================================================
// The kernel: persistent CTAs, 64-key steps; scores, probabilities and output each have their own TMEM columns.
//
// An earlier version kept P in the first half of the score tile it came from (128-key steps, one score tile per query
// tile); ncu showed tensor pipe 47 %, MUFU pipe 48 %, and a clock64 timeline of one CTA why — per query tile the chain
// S = Q K^T -> softmax -> O += P V -> next S  is strictly serial because P overwrites the score buffer: the next score MMA
// can only be issued after the softmax has finished AND the P V MMA has consumed P.  Double-buffering the scores alone does
// not break the chain (tried: same time).  Here a step is 64 keys and a tile owns
//     S (64 columns) | P[0] | P[1] (32 columns each: 64 keys as 16-bit pairs) | O (128 columns)
// — 256 columns per tile, 512 for the CTA's two tiles.  The softmax warp copies S(j) into registers and releases the score
// columns AT ONCE (s_empty), so Q K_{j+1}^T runs underneath the softmax of step j and S(j + 1) is waiting when that softmax ends;
// P(j) goes to buffer j & 1, which only needs P V(j - 2) to have completed.  One MMA warp issues Q_t K_{j+1}^T (on
// s_empty_t(j)) for both tiles, another O_t += P_t(j) V_j (on p_full_t(j)): a warp that issues both kinds serialises two
// barrier waits per step and was the bottleneck.  Barriers that a waiter may lag by two phases are split per buffer (a
// parity wait cannot tell phase k from k + 2).
//
// Persistent: one CTA per SM takes (sequence, head, 256-row query block) work items from a global counter, a head's blocks
// consecutively and heaviest first (the CTAs that run next to each other share that head's K / V in L2).  A one-item-per-CTA
// launch spent ~11,000 cycles per CTA outside the steps (850 set-up, 6,400 until the first score tile — Q is 64 KB at one
// SM's share of the bandwidth —, 5,000 drain), 20 % of the kernel at S = 3046 and 40 % at S = 980.  Here every role walks the
// item sequence at its own pace (the scheduler warp broadcasts item ids through a 2-slot ring): the producer loads the next
// item's Q as soon as the last score MMA of the current one has read it, the first scores of the next item are computed while
// the softmax warps store the current output, and every ring / barrier phase simply keeps counting across items.
// =====================================================================================================================
constexpr int kA3Keys = 64;                          // keys per step
constexpr int kA3KvBytes = kA3Keys * kAttTile * 2;   // one K or V step tile: 16 KB (two [64 x 64] halves)
constexpr int kA3KStages = 5, kA3VStages = 5;
constexpr uint32_t kA3TileCols = 256, kA3ColP = 64, kA3ColO = 128;   // TMEM columns of a tile: S at 0, P[b] at 64 + 32 b, O at 128
constexpr int kA3Consumers = 11;                     // warps that read the item ring: TMA, Q K, P V, 8 softmax

struct Att3Smem {
  static constexpr int Q = 0;                                   // Q_A, Q_B
  static constexpr int K = Q + 2 * kAttTileBytes;
  static constexpr int V = K + kA3KStages * kA3KvBytes;
  static constexpr int BAR = V + kA3VStages * kA3KvBytes;       // 224 KB of tiles
  static constexpr int N_BAR = 2 * kA3KStages + 2 * kA3VStages + 2 + 2 + 2 + 2 + 4 + 4 + 2 + 4;
  static constexpr int ITEM = BAR + N_BAR * 8;                  // int [2]: the item ring
  static constexpr int TOTAL = ITEM + 16 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

struct AttItem {
  int q0, n_a, n_b, n_max, col0;
  long long row0;
};

__device__ __forceinline__ AttItem att_item(const AttParams& P, int idx, int n_qb) {
  // item order: (sequence, head) major, query blocks of a head heaviest first
  const int qb = n_qb - 1 - idx % n_qb;
  const int sh = idx / n_qb;
  const int head = sh % P.n_heads, seq = sh / P.n_heads;
  AttItem it;
  it.q0 = qb * 2 * kAttTile;
  const int n_keys = (P.seq_len + kA3Keys - 1) / kA3Keys;   // steps that hold any key of the sequence
  it.n_a = min(4 * qb + 2, n_keys);                          // causal: tile A (rows q0 .. q0 + 127) sees keys < q0 + 128
  it.n_b = it.q0 + kAttTile < P.seq_len ? min(4 * qb + 4, n_keys) : 0;   // tile B two steps more (none if it lies past the end)
  it.n_max = max(it.n_a, it.n_b);
  it.row0 = (long long)seq * P.seq_len;
  it.col0 = head * kAttTile;
  return it;
}

template <bool F16, int PP, bool LAZY>
__global__ void __launch_bounds__(kAttThreads, 1) attention3_kernel(const __grid_constant__ AttParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Att3Smem::BAR);
  uint64_t* k_full = bars;
  uint64_t* k_empty = k_full + kA3KStages;
  uint64_t* v_full = k_empty + kA3KStages;
  uint64_t* v_empty = v_full + kA3VStages;
  uint64_t* q_full = v_empty + kA3VStages;   // [2]     TMA -> Q K warp: Q_t of the item has landed
  uint64_t* q_empty = q_full + 2;            // [2]     Q K warp -> TMA: the last score MMA of the item has read Q_t
  uint64_t* s_full = q_empty + 2;            // [2]     Q K warp -> softmax of tile t: S_t is in TMEM
  uint64_t* s_empty = s_full + 2;            // [2]     softmax -> Q K warp: S_t has been copied to registers
  uint64_t* p_full = s_empty + 2;            // [2][2]  softmax -> P V warp: P_t is in TMEM buffer b (O_t rescaled if the maximum grew)
  uint64_t* pv_done = p_full + 4;            // [2][2]  P V warp -> softmax: O_t += P_t V has completed (buffer b is free)
  uint64_t* o_empty = pv_done + 4;           // [2]     softmax -> P V warp: O_t of the finished item has been read out
  uint64_t* it_full = o_empty + 2;           // [2]     scheduler -> everyone: item ring slot filled
  uint64_t* it_empty = it_full + 2;          // [2]     everyone -> scheduler: slot read
  volatile int* s_item = reinterpret_cast<volatile int*>(smem + Att3Smem::ITEM);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Att3Smem::ITEM + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qb = (P.seq_len + 2 * kAttTile - 1) / (2 * kAttTile);
  const int n_items = P.n_items;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmQ);
    tma_prefetch_desc(&P.tmK);
    tma_prefetch_desc(&P.tmV);
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < kA3KStages; ++st) {
      mbar_init(&k_full[st], 1);
      mbar_init(&k_empty[st], 1);
    }
    for (int st = 0; st < kA3VStages; ++st) {
      mbar_init(&v_full[st], 1);
      mbar_init(&v_empty[st], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1);
      mbar_init(&q_empty[t], 1);
      mbar_init(&s_full[t], 1);
      mbar_init(&s_empty[t], 4);
      mbar_init(&o_empty[t], 4);
      mbar_init(&it_full[t], 1);
      mbar_init(&it_empty[t], kA3Consumers);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&p_full[2 * t + b], 4);
        mbar_init(&pv_done[2 * t + b], 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // every consumer warp walks the same item sequence: n-th item from ring slot n & 1
  auto next_item = [&](int n) -> int {
    mbar_wait(&it_full[n & 1], (uint32_t)((n >> 1) & 1));
    const int idx = s_item[n & 1];
    __syncwarp();
    if (lane == 0) mbar_arrive(&it_empty[n & 1]);
    return idx;
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 3) {
      // ===== scheduler: hands out item ids (a sentinel >= n_items ends every role's loop)
      if (lane == 0) {
        for (int n = 0;; ++n) {
          if (n >= 2) mbar_wait(&it_empty[n & 1], (uint32_t)(((n >> 1) - 1) & 1));
          const int idx = (int)atomicAdd(P.counter, 1u);
          s_item[n & 1] = idx;
          mbar_arrive(&it_full[n & 1]);   // release: the slot is written before the arrive is observed
          if (idx >= n_items) break;
        }
      }
    } else if (warp == 0) {
      // ===== TMA producer: whole warp converged, one elected lane issues.  Per item: Q_A, Q_B (once the score MMAs of the
      // previous item are through with them), K_0, then (K_{j+1}, V_j) per step — the order the MMA warps consume them in.
      int kc = 0, vc = 0;   // K / V tiles loaded so far (ring position and phase keep counting across items)
      auto load_kv = [&](bool is_v, const AttItem& it, int j) {
        int& c = is_v ? vc : kc;
        const int stages = is_v ? kA3VStages : kA3KStages;
        const int st = c % stages;
        uint64_t* full = is_v ? &v_full[st] : &k_full[st];
        mbar_wait(is_v ? &v_empty[st] : &k_empty[st], (uint32_t)(((c / stages) & 1) ^ 1));
        if (elect_one()) {
          mbar_expect_tx(full, kA3KvBytes);
          uint8_t* d = smem + (is_v ? Att3Smem::V : Att3Smem::K) + st * kA3KvBytes;
          const int krow = (int)(it.row0 + (long long)j * kA3Keys);
          const CUtensorMap* tm = is_v ? &P.tmV : &P.tmK;
          tma_load_2d(tm, full, d, it.col0, krow);
          tma_load_2d(tm, full, d + kA3KvBytes / 2, it.col0 + 64, krow);
        }
        __syncwarp();
        ++c;
      };
      for (int n = 0;; ++n) {
        const int idx = next_item(n);
        if (idx >= n_items) break;
        const AttItem it = att_item(P, idx, n_qb);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (n > 0) mbar_wait(&q_empty[t], (uint32_t)((n - 1) & 1));
          if (elect_one()) {
            mbar_expect_tx(&q_full[t], kAttTileBytes);
            uint8_t* qd = smem + Att3Smem::Q + t * kAttTileBytes;
            tma_load_2d(&P.tmQ, &q_full[t], qd, it.col0, (int)(it.row0 + it.q0 + t * kAttTile));
            tma_load_2d(&P.tmQ, &q_full[t], qd + kAttHalfBytes, it.col0 + 64, (int)(it.row0 + it.q0 + t * kAttTile));
          }
          __syncwarp();
        }
        load_kv(false, it, 0);
        for (int j = 0; j < it.n_max; ++j) {
          if (j + 1 < it.n_max) load_kv(false, it, j + 1);
          load_kv(true, it, j);
        }
      }
    } else if (warp == 1) {
      // ===== score MMAs of both tiles (whole warp converged, one elected lane issues):  S_t = Q_t K_j^T as soon as the softmax
      // of tile t has taken the previous S_t out of TMEM
      const uint64_t kd0 = umma_smem_desc(smem_u32(smem + Att3Smem::K));
      const uint32_t idesc_qk = P.idesc_qk;
      auto issue_qk = [&](int t, int kst) {  // M 128, N 64, K 16 x 8; elected lane only
        const uint64_t qd = umma_smem_desc(smem_u32(smem + Att3Smem::Q + t * kAttTileBytes));
        const uint64_t kd = kd0 + (uint64_t)(kst * (kA3KvBytes >> 4));
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base + (uint32_t)t * kA3TileCols, qd + (uint64_t)(kb * (kAttHalfBytes >> 4) + 2 * k),
                     kd + (uint64_t)(kb * (kA3KvBytes >> 5) + 2 * k), idesc_qk, (kb | k) ? 1u : 0u);
        umma_commit(&s_full[t]);
      };
      int kst = 0;
      uint32_t kph = 0;
      int sc[2] = {0, 0};   // score tiles issued so far per tile (s_empty phase of the previous one = sc - 1)
      for (int n = 0;; ++n) {
        const int idx = next_item(n);
        if (idx >= n_items) break;
        const AttItem it = att_item(P, idx, n_qb);
        mbar_wait(&q_full[0], (uint32_t)(n & 1));
        mbar_wait(&q_full[1], (uint32_t)(n & 1));
        for (int j = 0; j < it.n_max; ++j) {
          mbar_wait(&k_full[kst], kph);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int n_t = t == 0 ? it.n_a : it.n_b;
            if (j < n_t) {
              if (sc[t] > 0) mbar_wait(&s_empty[t], (uint32_t)((sc[t] - 1) & 1));
              tc_fence_after();
              if (elect_one()) {
                issue_qk(t, kst);
                if (j == n_t - 1) umma_commit(&q_empty[t]);   // Q_t may be overwritten once these MMAs have read it
              }
              __syncwarp();
              ++sc[t];
            }
          }
          if (elect_one()) umma_commit(&k_empty[kst]);
          __syncwarp();
          if (++kst == kA3KStages) {
            kst = 0;
            kph ^= 1u;
          }
        }
        if (it.n_b == 0) {  // tile B lies past the end of the sequence: nothing read its Q
          if (elect_one()) umma_commit(&q_empty[1]);
          __syncwarp();
        }
      }
    } else {
      // ===== P V MMAs of both tiles:  O_t (+)= P_t V_j, P from TMEM buffer b (8 columns = 16 keys per MMA)
      const uint64_t vd0 = umma_smem_desc_mn(smem_u32(smem + Att3Smem::V), kA3KvBytes / 2);
      const uint32_t idesc_pv = P.idesc_pv;
      int vst = 0;
      uint32_t vph = 0;
      int pc[2] = {0, 0};   // P tiles consumed so far per tile: buffer pc & 1, phase pc >> 1
      for (int n = 0;; ++n) {
        const int idx = next_item(n);
        if (idx >= n_items) break;
        const AttItem it = att_item(P, idx, n_qb);
        for (int j = 0; j < it.n_max; ++j) {
          mbar_wait(&v_full[vst], vph);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (j < (t == 0 ? it.n_a : it.n_b)) {
              if (j == 0 && n > 0) mbar_wait(&o_empty[t], (uint32_t)((n - 1) & 1));   // the previous item's O_t has been read out
              const int b = pc[t] & 1;
              mbar_wait(&p_full[2 * t + b], (uint32_t)((pc[t] >> 1) & 1));
              tc_fence_after();
              if (elect_one()) {
                const uint64_t vd = vd0 + (uint64_t)(vst * (kA3KvBytes >> 4));
                const uint32_t tile_tm = tmem_base + (uint32_t)t * kA3TileCols;
#pragma unroll
                for (int kk = 0; kk < kA3Keys / 16; ++kk)
                  umma_f16_ts(tile_tm + kA3ColO, tile_tm + kA3ColP + (uint32_t)(b * 32 + kk * 8), vd + (uint64_t)(kk * (2048 >> 4)),
                              idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(&pv_done[2 * t + b]);
              }
              __syncwarp();
              ++pc[t];
            }
          }
          if (elect_one()) umma_commit(&v_empty[vst]);
          __syncwarp();
          if (++vst == kA3VStages) {
            vst = 0;
            vph ^= 1u;
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int t = (warp - 4) >> 2;      // 0: tile A, 1: tile B
    const int q = warp & 3;             // TMEM lane quadrant (hardware: warp id % 4)
    const int r = q * 32 + lane;        // query row inside the tile = TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_off + (uint32_t)t * kA3TileCols;
    const uint32_t p_tmem = s_tmem + kA3ColP, o_tmem = s_tmem + kA3ColO;
    const uint64_t scale2 = pk2(P.scale_log2, P.scale_log2);
    int c = 0;   // steps of this tile so far, all items: S phase c & 1, P buffer c & 1 with phase c >> 1
    for (int n = 0;; ++n) {
      const int idx = next_item(n);
      if (idx >= n_items) break;
      const AttItem it = att_item(P, idx, n_qb);
      const int n_t = t == 0 ? it.n_a : it.n_b;
      const int qrow = it.q0 + t * kAttTile + r;   // query position inside the sequence
      float m_used = -INFINITY, l_run = 0.0f;
      for (int j = 0; j < n_t; ++j, ++c) {
        mbar_wait(&s_full[t], (uint32_t)(c & 1));
        tc_fence_after();
        uint32_t sv[kA3Keys];
        tmem_ld_32x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        tmem_ld_32x32(s_tmem + 32u, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[t]);  // the score columns are free for the next Q K^T
        const int key0 = j * kA3Keys;
        if (key0 + kA3Keys - 1 > it.q0 + t * kAttTile + q * 32) {  // warp-uniform: some key of the step lies after some query of this warp
#pragma unroll
          for (int cc = 0; cc < kA3Keys; ++cc)
            if (key0 + cc > qrow) sv[cc] = 0xff800000u;  // -inf
        }
        float mx8[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) mx8[cc] = max3(__uint_as_float(sv[cc]), __uint_as_float(sv[cc + 8]), __uint_as_float(sv[cc + 16]));
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) mx8[cc] = max3(mx8[cc], __uint_as_float(sv[cc + 24]), __uint_as_float(sv[cc + 32]));
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) mx8[cc] = max3(mx8[cc], __uint_as_float(sv[cc + 40]), __uint_as_float(sv[cc + 48]));
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) mx8[cc] = fmaxf(mx8[cc], __uint_as_float(sv[cc + 56]));
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
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          // rescale this thread's row of O in place (rare: the maximum is lazy); every earlier P V must have landed first
          mbar_wait(&pv_done[2 * t + ((c - 1) & 1)], (uint32_t)(((c - 1) >> 1) & 1));
          tc_fence_after();
          const uint64_t a2 = pk2(alpha, alpha);
#pragma unroll
          for (int cc = 0; cc < kAttTile; cc += 32) {
            uint32_t v[32];
            tmem_ld_32x32(o_tmem + (uint32_t)cc, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              float a, b;
              upk2(mul2(pk2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), a2), a, b);
              v[e] = __float_as_uint(a);
              v[e + 1] = __float_as_uint(b);
            }
            tmem_st_32x32(o_tmem + (uint32_t)cc, v);
          }
        }
        // p = exp2(s * scale - m): fp32 row sum, 16-bit pairs
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
        float la, lb;
        upk2(add2(add2(l2[0], l2[1]), add2(l2[2], l2[3])), la, lb);
        l_run += la + lb;
        // P buffer c & 1 was last read by the P V of step c - 2 (possibly the previous item's)
        if (c >= 2) mbar_wait(&pv_done[2 * t + (c & 1)], (uint32_t)(((c >> 1) - 1) & 1));
        tmem_st_32x32(p_tmem + (uint32_t)((c & 1) * 32), w);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * t + (c & 1)]);
      }
      if (n_t > 0) {
        // all P V of this tile have completed (the tensor pipe completes them in order): normalise by the row sum, store
        const float inv = 1.0f / l_run;
        mbar_wait(&pv_done[2 * t + ((c - 1) & 1)], (uint32_t)(((c - 1) >> 1) & 1));
        tc_fence_after();
        const bool ok = qrow < P.seq_len;
        const long long tt = it.row0 + qrow;
        const long long orow = (ok && P.out_rowmap) ? (long long)P.out_rowmap[tt] : tt;
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(P.out) + (orow * P.ld_out + it.col0) * 2);
#pragma unroll
        for (int cc = 0; cc < kAttTile; cc += 32) {
          uint32_t v[32];
          tmem_ld_32x32(o_tmem + (uint32_t)cc, v);
          tmem_ld_wait();
          if (cc + 32 == kAttTile) {  // O_t is in registers: the next item's first P V may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[t]);
          }
          if (ok) {
#pragma unroll
            for (int e = 0; e < 32; e += 8)
              dst[(cc + e) >> 3] = make_uint4(pack2<F16>(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv),
                                              pack2<F16>(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv),
                                              pack2<F16>(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv),
                                              pack2<F16>(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv));
          }
        }
      } else {
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[t]);   // keeps the per-item phase of o_empty in step for a tile without work
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
static cudaError_t launch_attention(const AttParams& P, int grid, cudaStream_t stream) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(attention3_kernel<F16, PP, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att3Smem::DYN_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  attention3_kernel<F16, PP, LAZY><<<grid, kAttThreads, Att3Smem::DYN_BYTES, stream>>>(P);
  return cudaGetLastError();
}

// work counters: 64 per device, handed out round-robin (launches in flight on different streams never share one); each is
// zeroed on the launch stream right before its launch
static unsigned int* attention_counter() {
  static unsigned int* counters[64] = {};
  static std::atomic<unsigned int> next{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!counters[dev] && cudaMalloc(&counters[dev], 64 * sizeof(unsigned int)) != cudaSuccess) counters[dev] = nullptr;
  return counters[dev] ? counters[dev] + (next.fetch_add(1) & 63u) : nullptr;
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
  const long long items = (long long)batch * n_heads * ((seq_len + 2 * kAttTile - 1) / (2 * kAttTile));
  MC_REQUIRE(items < (1LL << 30), "attention: too many work items");
  AttParams P;
  memset(&P, 0, sizeof(P));
  // Q: 128-row boxes; K / V: 64-row boxes (one step)
  int rc = encode_operand(&P.tmQ, q, tokens, (long long)n_heads * head_dim, ld_qkv, kAttTile, dtype);
  if (rc == MC_OK) rc = encode_operand(&P.tmK, k, tokens, (long long)n_heads * head_dim, ld_qkv, kA3Keys, dtype);
  if (rc == MC_OK) rc = encode_operand(&P.tmV, v, tokens, (long long)n_heads * head_dim, ld_qkv, kA3Keys, dtype);
  if (rc != MC_OK) return rc;
  P.out = out;
  P.out_rowmap = out_rowmap;
  P.ld_out = ld_out;
  P.seq_len = seq_len;
  P.n_heads = n_heads;
  P.is_f16 = dtype == MC_F16;
  P.n_items = (int)items;
  P.scale_log2 = softmax_scale * 1.4426950408889634f;
  // instruction descriptors: D = F32, A/B = bf16 / fp16, N >> 3 at bit 17 (64 keys for the scores, 128 head-dim columns for
  // P V), M = 128 (>> 4 at bit 24); bit 16 = B operand MN-major (the V tile is [keys, head_dim] with head_dim contiguous)
  const unsigned int fmt = dtype == MC_F16 ? 0u : 1u;
  const unsigned int common = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(kAttTile >> 4) << 24);
  P.idesc_qk = common | ((unsigned)(kA3Keys >> 3) << 17);
  P.idesc_pv = common | ((unsigned)(kAttTile >> 3) << 17) | (1u << 16);
  cudaStream_t st = (cudaStream_t)stream;
  P.counter = attention_counter();
  MC_REQUIRE(P.counter != nullptr, "attention: no device memory for the work counter");
  MC_CUDA_OK(cudaMemsetAsync(P.counter, 0, sizeof(unsigned int), st));
  const int sms = sm_count();
  MC_REQUIRE(sms > 0, "no CUDA device");
  const int grid = (int)std::min<long long>(items, sms);
  // bits 4-7 = 1 + pairs out of 4 whose exp2 runs on the FMA pipe (0 = default), bit 8 = rescale on every new maximum
  const int pp = ((tuning >> 4) & 0xf) ? ((tuning >> 4) & 0xf) - 1 : 0;
  const bool eager = (tuning >> 8) & 1;
  const bool f16 = dtype == MC_F16;
  cudaError_t e = cudaErrorInvalidValue;
#define MC_ATT(PPV)                                                                                          \
  case PPV:                                                                                                  \
    e = f16 ? (eager ? launch_attention<true, PPV, false>(P, grid, st) : launch_attention<true, PPV, true>(P, grid, st))   \
            : (eager ? launch_attention<false, PPV, false>(P, grid, st) : launch_attention<false, PPV, true>(P, grid, st)); \
    break;
  switch (pp) {
    MC_ATT(0)
    MC_ATT(1)
    MC_ATT(2)
    MC_ATT(3)
    default:
      return fail(MC_ERR_INVALID, "attention: tuning 0x%x selects no kernel", tuning);
  }
#undef MC_ATT
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "attention launch failed: %s", cudaGetErrorString(e));
  return MC_OK;
}

extern "C" int mc_attention_causal(const void* q, const void* k, const void* v, void* out, int64_t ld_qkv, int64_t ld_out,
                                   const int32_t* out_rowmap, int batch, int seq_len, int n_heads, int head_dim, float softmax_scale,
                                   int dtype, mc_stream_t stream) {
  return mc_attention_causal_tuned(q, k, v, out, ld_qkv, ld_out, out_rowmap, batch, seq_len, n_heads, head_dim, softmax_scale, dtype, 0,
                                   stream);
}
