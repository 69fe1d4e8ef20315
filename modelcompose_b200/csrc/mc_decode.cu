// Decode step: one new token per sequence against a key/value cache (M = batch <= 64 rows).
//
// Replaces (reference paths):
//   modelcompose/model/multimodal_arch.py:290-293                     cache present: no splice, mask rebuilt over past + 1
//   modelcompose/model/language_model/multimodal_llama.py:436-438     modality masks dropped -> every row takes the default adapter
//   :120-160 LocalLoraLinear.forward at q_len = 1, :274-312 attention with past_key_value (cat of cached and new keys),
//   :380-390 MLP, :747-767 prepare_inputs_for_generation; the greedy sampler HF generate runs behind
//   modelcompose/eval/model_multimodal_qa_loader.py:93-102.
//
// With at most 64 rows every linear is a single pass over its weight matrix and attention a single pass over the cache:
// all kernels here are HBM streams (roofline: bytes of weights / cache per step over the copy bandwidth), so none of them
// uses tcgen05 — a 128-row UMMA tile would spend >= 50 % of the tensor pipe on padding and, worse, leave most SMs without
// a tile to stream (a 4096 x 4096 weight is 32 tiles of 128 rows on 148 SMs).
//   skinny_linear_kernel   weight rows -> registers (128-bit streaming loads) -> mma.sync.m16n8k16 fragments directly
//   decode_rope_append     RoPE of the new q / k (rounding points of mc_rope), k / v appended to the cache
//   decode_attention       split-KV online softmax, 16 lanes per key, deterministic combine by the last CTA
//   argmax_rows            greedy sampler on the device
#include <algorithm>
#include <cmath>
#include <cstring>

#include "mc_tc.cuh"

namespace mc {

// ================================================================================================ skinny linear
constexpr int kSkWarps = 8;
constexpr int kSkThreads = kSkWarps * 32;
constexpr int kSkMaxProb = MC_SKINNY_MAX_PROBLEMS;

struct SkProblem {
  const char* A0;
  const char* B0;
  const char* A1;
  const char* B1;
  const char* B0u;
  const char* A1u;
  const char* B1u;
  char* C;
  const char* residual;
  const float* col_scale;
  long long lda0, ldb0, lda1, ldb1, ldc, ldr;  // BYTES
  int M, N, K0, K1, epilogue, cta_end;
};

struct SkParams {
  SkProblem prob[kSkMaxProb];
  int n_prob;
};

template <bool F16>
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  if constexpr (F16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

__device__ __forceinline__ uint4 ld_w16(const char* p) {  // weights: read once, keep them out of L1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// One K range of one operand pair, accumulated into acc.  A warp owns every kSkWarps-th 64-element K step of the CTA's rows.
// Lane (g = lane / 4, t = lane % 4) loads the 16-byte chunks [k + 8t, k + 8t + 8) and [k + 32 + 8t, ...) of weight rows g and
// g + 8 of every 16-row tile, and the same chunks of activation row 8 * nt + g.  The two operands of an MMA only have to agree
// on WHICH k each fragment slot holds, so a chunk's four 32-bit words feed two MMAs as they are: slots (2t, 2t+1 | 2t+8, 2t+9)
// take words (0 | 1) in the first MMA and (2 | 3) in the second — no shuffle, no shared-memory staging, 128-byte row segments.
// Rows past N / M are clamped to the last valid row by the caller (their results are never stored), so the full 64-element
// steps run without predicates; the weights of the step after next are requested before the MMAs of the current one.
// SEPX: the row tiles multiply different activations (rank-space inputs of gate_proj / up_proj in the dual problem).
template <int RT>
struct SkW {
  uint4 v[RT][2][2];
};

template <int RT>
__device__ __forceinline__ void sk_load_w(SkW<RT>& w, const char* (&wrow)[RT][2], int k) {
#pragma unroll
  for (int rt = 0; rt < RT; ++rt)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int h = 0; h < 2; ++h) w.v[rt][r][h] = ld_w16(wrow[rt][r] + 2ll * (k + 32 * h));
}

template <int RT, int NT, bool F16, bool SEPX>
__device__ __forceinline__ void sk_step(float (&acc)[RT][NT][4], const SkW<RT>& w, const char* (&xrow)[SEPX ? RT : 1][NT], int k) {
  constexpr int XT = SEPX ? RT : 1;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4 x[XT][NT];
#pragma unroll
    for (int xt = 0; xt < XT; ++xt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) x[xt][nt] = *reinterpret_cast<const uint4*>(xrow[xt][nt] + 2ll * (k + 32 * h));
#pragma unroll
    for (int rt = 0; rt < RT; ++rt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint4& xv = x[SEPX ? rt : 0][nt];
        mma16816<F16>(acc[rt][nt], w.v[rt][0][h].x, w.v[rt][1][h].x, w.v[rt][0][h].y, w.v[rt][1][h].y, xv.x, xv.y);
        mma16816<F16>(acc[rt][nt], w.v[rt][0][h].z, w.v[rt][1][h].z, w.v[rt][0][h].w, w.v[rt][1][h].w, xv.z, xv.w);
      }
  }
}

template <int RT, int NT, bool F16, bool SEPX>
__device__ __forceinline__ void sk_accumulate(float (&acc)[RT][NT][4], const char* (&wrow)[RT][2], const char* (&xbase)[RT],
                                              long long ldx, int M, int K, int warp, int g, int t) {
  constexpr int XT = SEPX ? RT : 1;
  const char* xrow[XT][NT];
#pragma unroll
  for (int xt = 0; xt < XT; ++xt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) xrow[xt][nt] = xbase[xt] + min(nt * 8 + g, M - 1) * ldx;
  const int full = K >> 6;
  SkW<RT> w0, w1;
  if (warp < full) sk_load_w<RT>(w0, wrow, (warp << 6) + (t << 3));
  for (int s = warp; s < full; s += 2 * kSkWarps) {
    const int s1 = s + kSkWarps, s2 = s + 2 * kSkWarps;
    if (s1 < full) sk_load_w<RT>(w1, wrow, (s1 << 6) + (t << 3));
    sk_step<RT, NT, F16, SEPX>(acc, w0, xrow, (s << 6) + (t << 3));
    if (s2 < full) sk_load_w<RT>(w0, wrow, (s2 << 6) + (t << 3));
    if (s1 < full) sk_step<RT, NT, F16, SEPX>(acc, w1, xrow, (s1 << 6) + (t << 3));
  }
  if ((K & 63) && warp == (full & (kSkWarps - 1))) {  // K % 64 != 0: one partial step; chunks past K read as zeros (K % 8 == 0)
    const int k = (full << 6) + (t << 3);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = k + 32 * h;
      const bool in = kk < K;  // per lane: only the LOADS are predicated, the MMAs are warp-wide
#pragma unroll
      for (int rt = 0; rt < RT; ++rt) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        const uint4 wa = in ? ld_w16(wrow[rt][0] + 2ll * kk) : z, wb = in ? ld_w16(wrow[rt][1] + 2ll * kk) : z;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint4 xv = in ? *reinterpret_cast<const uint4*>(xrow[SEPX ? rt : 0][nt] + 2ll * kk) : z;
          mma16816<F16>(acc[rt][nt], wa.x, wb.x, wa.y, wb.y, xv.x, xv.y);
          mma16816<F16>(acc[rt][nt], wa.z, wb.z, wa.w, wb.w, xv.z, xv.w);
        }
      }
    }
  }
}

template <bool F16>
__device__ __forceinline__ float sk_round(float v) {
  if constexpr (F16) return __half2float(__float2half_rn(v));
  else return __bfloat162float(__float2bfloat16_rn(v));
}
template <bool F16>
__device__ __forceinline__ float sk_load(const char* p) {
  if constexpr (F16) return __half2float(*reinterpret_cast<const __half*>(p));
  else return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p));
}
template <bool F16>
__device__ __forceinline__ void sk_store(char* p, float v) {
  if constexpr (F16) *reinterpret_cast<__half*>(p) = __float2half_rn(v);
  else *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(v);
}

// RT 16-row tiles of weight rows per CTA (the dual problem uses tile 0 for gate_proj and tile 1 for up_proj rows), NT 8-row tiles of
// activations (M <= 8 NT).  K is split over the 8 warps; partial sums meet in shared memory and are added in warp order.
template <int RT, int NT, bool F16>
__global__ void __launch_bounds__(kSkThreads) skinny_linear_kernel(const __grid_constant__ SkParams P) {
  extern __shared__ float sk_red[];  // [kSkWarps][8 NT][16 RT + 1]
  constexpr int FT = 16 * RT, MT = 8 * NT, LDR = FT + 1;
  griddep_launch_dependents();
  griddep_wait();
  int p = 0, first = 0;
  while (p < P.n_prob - 1 && (int)blockIdx.x >= P.prob[p].cta_end) {
    first = P.prob[p].cta_end;
    ++p;
  }
  const SkProblem& pr = P.prob[p];
  const bool dual = pr.epilogue == MC_SKINNY_EPI_SILU_MUL;
  const int F = dual ? 16 : FT;  // output features of this CTA
  const int n0 = ((int)blockIdx.x - first) * F;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

  float acc[RT][NT][4];
#pragma unroll
  for (int rt = 0; rt < RT; ++rt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[rt][nt][e] = 0.f;

#pragma unroll 1
  for (int phase = 0; phase < 2; ++phase) {
    const int K = phase ? pr.K1 : pr.K0;
    if (K == 0) continue;
    const long long ldw = phase ? pr.ldb1 : pr.ldb0, ldx = phase ? pr.lda1 : pr.lda0;
    const char* wrow[RT][2];
    const char* xbase[RT];
#pragma unroll
    for (int rt = 0; rt < RT; ++rt) {
      const bool second = dual && rt == 1;
      const char* wb = second ? (phase ? pr.B1u : pr.B0u) : (phase ? pr.B1 : pr.B0);
      const int row = n0 + (dual ? 0 : rt * 16) + g;  // rows past N: clamped, computed, never stored
      wrow[rt][0] = wb + min(row, pr.N - 1) * ldw;
      wrow[rt][1] = wb + min(row + 8, pr.N - 1) * ldw;
      xbase[rt] = (second && phase) ? pr.A1u : (phase ? pr.A1 : pr.A0);
    }
    if (RT > 1 && dual && phase) sk_accumulate<RT, NT, F16, (RT > 1)>(acc, wrow, xbase, ldx, pr.M, K, warp, g, t);
    else sk_accumulate<RT, NT, F16, false>(acc, wrow, xbase, ldx, pr.M, K, warp, g, t);
  }

  // accumulator fragment: c0, c1 = (feature g, rows 2t, 2t + 1 of the activation tile), c2, c3 = feature g + 8
  float* mine = sk_red + warp * (MT * LDR);
#pragma unroll
  for (int rt = 0; rt < RT; ++rt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int m = nt * 8 + 2 * t, f = rt * 16 + g;
      mine[m * LDR + f] = acc[rt][nt][0];
      mine[(m + 1) * LDR + f] = acc[rt][nt][1];
      mine[m * LDR + f + 8] = acc[rt][nt][2];
      mine[(m + 1) * LDR + f + 8] = acc[rt][nt][3];
    }
  __syncthreads();
  for (int idx = threadIdx.x; idx < pr.M * F; idx += kSkThreads) {
    const int m = idx / F, f = idx % F, n = n0 + f;
    if (n >= pr.N) continue;
    float v = 0.f, u = 0.f;
#pragma unroll
    for (int w = 0; w < kSkWarps; ++w) {
      v += sk_red[w * (MT * LDR) + m * LDR + f];
      if (RT > 1 && dual) u += sk_red[w * (MT * LDR) + m * LDR + 16 + f];
    }
    if (pr.epilogue == MC_SKINNY_EPI_RESIDUAL) {
      v += sk_load<F16>(pr.residual + m * pr.ldr + 2ll * n);
    } else if (pr.epilogue == MC_SKINNY_EPI_COLSCALE) {
      v *= pr.col_scale[n];
    } else if (pr.epilogue == MC_SKINNY_EPI_SILU_MUL) {
      const float gate = sk_round<F16>(v);
      const float sg = sk_round<F16>(__fdividef(gate, 1.0f + __expf(-gate)));
      v = sg * sk_round<F16>(u);
    }
    sk_store<F16>(pr.C + m * pr.ldc + 2ll * n, v);
  }
}

// ================================================================================================ skinny linear, stream-K
// The same product as a persistent stream-K kernel fed by TMA.  Work = (problem, block of R weight rows, 128-element K chunk)
// iterations in row-block-major order; every CTA (one per SM) owns an equal contiguous span of them, so all SMs stream for the
// same time whatever N and K are (a 4096-row weight is 64 row blocks on 148 SMs: without the K split most SMs would idle).
// A producer lane issues, per iteration, two [R x 64] boxes of the weight matrix and two [8 NT x 64] boxes of the activations
// (cp.async.bulk.tensor, 128-byte swizzle, completion on an mbarrier) into a ring that holds ~170 KB in flight per SM — the
// bytes in flight that saturate HBM come from shared memory, not from registers.  (A first version staged every row segment
// with its own 256-byte cp.async.bulk: ~100 copies per iteration ran into the copy engine's request rate, 0.65 TB/s.)
// Eight consumer warps read the fragments with ldmatrix (swizzled: conflict-free) and run mma.sync.m16n8k16; at the end of a
// row block their partial sums meet in shared memory.  A row block whose K range is shared by several CTAs goes through a
// scratch tile per CTA; the last CTA to arrive adds the partial tiles in CTA order (deterministic) and runs the epilogue.
constexpr int kS2Consumers = 8;
constexpr int kS2Threads = (kS2Consumers + 1) * 32;  // 8 consumer warps + the producer warp (one elected lane issues the TMA loads)
constexpr int kS2MaxStages = 12;
constexpr int kS2TileFloats = MC_SKINNY_MAX_M * 64;  // scratch tile: [M][R] fp32, R <= 64

struct alignas(64) S2Problem {
  CUtensorMap tmB0, tmB1, tmB0u, tmB1u;  // weights: box [R (dual: R / 2) rows x 64]
  CUtensorMap tmA0, tmA1, tmA1u;         // activations: box [8 NT rows x 64], rows past M read as zeros
  SkProblem pr;
  int nrb;         // row blocks
  int nkc0, nkc1;  // K chunks of the two products
  int iter_end;    // prefix sum of iterations
  int rb_base;     // global index of this problem's first row block (counters)
};

struct alignas(64) S2Params {
  S2Problem prob[kSkMaxProb];
  int n_prob, total_iters, span, stages, xslots, stage_bytes;
  int nh;          // 64-element TMA boxes per K chunk (2: 128-element chunks, 4: 256)
  int copy_only;   // profiling aid: consumers release the stages without reading them (pipeline ceiling)
  float* scratch;  // [grid][2][kS2TileFloats]
  int* counters;   // [total row blocks], zero between launches
  // optional: the launch first computes norm_dst = rmsnorm(norm_src) * norm_w ([norm_rows, norm_hidden], one warp per row, the
  // arithmetic of mc_rmsnorm) — the activations its problems read as A0 — while the first ring of weight boxes is in flight
  const char* norm_src;
  const char* norm_w;
  char* norm_dst;
  long long norm_lds, norm_ldd;  // bytes
  int norm_rows, norm_hidden;
  float norm_eps;
  int* norm_cnt;   // [2]: CTAs that finished the norm, CTAs that finished the launch; zero between launches
};

struct S2Iter {  // position of the walk through the iteration space
  int p, rb, kc;
};

__device__ __forceinline__ S2Iter s2_locate(const S2Params& P, int it) {
  S2Iter w;
  w.p = 0;
  int first = 0;
  while (w.p < P.n_prob - 1 && it >= P.prob[w.p].iter_end) {
    first = P.prob[w.p].iter_end;
    ++w.p;
  }
  const int ipr = P.prob[w.p].nkc0 + P.prob[w.p].nkc1;
  w.rb = (it - first) / ipr;
  w.kc = (it - first) % ipr;
  return w;
}
__device__ __forceinline__ void s2_advance(const S2Params& P, S2Iter& w) {
  const S2Problem& q = P.prob[w.p];
  if (++w.kc == q.nkc0 + q.nkc1) {
    w.kc = 0;
    if (++w.rb == q.nrb) {
      w.rb = 0;
      ++w.p;
    }
  }
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kS2Consumers * 32) : "memory"); }

// RT 16-row weight tiles per row block (R = 16 RT rows; dual problems: first half gate_proj rows, second half up_proj rows of the
// same features), NT 8-row activation tiles.  Consumer warp w owns the tile pair w % (RT / 2) and every KP-th 16-element K step of
// a chunk, KP = 8 / (RT / 2).  Stage layout: weights K half 0 | weights K half 1 | activations (slot, K half) ..., every box a
// multiple of 1024 bytes so the 128-byte swizzle pattern (16-byte unit ^= row & 7) starts at row 0 of each.
template <int RT, int NT, bool F16>
__global__ void __launch_bounds__(kS2Threads, 1) skinny_streamk_kernel(const __grid_constant__ S2Params P) {
  constexpr int R = 16 * RT, MT = 8 * NT, TP = RT / 2, KP = kS2Consumers / TP;
  constexpr int LDT = R + 1;   // tile buffer [MT][R + 1]
  constexpr int LDR = 33;      // reduction buffer [consumer warp][MT][32 + 1]
  constexpr uint32_t W_HALF = R * 128, X_HALF = MT * 128;
  const uint32_t X_BASE = (uint32_t)P.nh * W_HALF, X_SLOT = (uint32_t)P.nh * X_HALF;
  const int KC = 64 * P.nh;
  extern __shared__ __align__(128) unsigned char s2_smem[];
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s2_smem) + 1023) & ~(uintptr_t)1023);
  unsigned char* tail = ring + (size_t)P.stages * P.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + kS2MaxStages;
  int* flag = reinterpret_cast<int*>(empty + kS2MaxStages);
  float* red = reinterpret_cast<float*>(tail + 256);
  float* tile = red + kS2Consumers * MT * LDR;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int it0 = min((int)blockIdx.x * P.span, P.total_iters), it1 = min(it0 + P.span, P.total_iters);
  griddep_launch_dependents();  // decode chain: the next kernel may become resident as soon as every CTA of this one runs
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, kS2Consumers);
    }
    fence_barrier_init();
  }
  __syncthreads();
  // fused RMSNorm (optional): every CTA's consumer warps take rows blockIdx.x * 8 + warp, + 8 * gridDim.x, ...; each CTA then counts
  // that owns rows counts itself on norm_cnt[0] (one thread fences after the CTA barrier), and the producers hold the ACTIVATION boxes back
  // until all of those have arrived (the weight boxes of the first ring are already in flight by then).  Generic-proxy stores, async-proxy (TMA)
  // reads: the reader issues fence.proxy.async after the acquire.
  auto norm_phase = [&]() {
    if (P.norm_rows <= 0 || warp >= kS2Consumers) return;
    griddep_wait();
    for (int row = (int)blockIdx.x * kS2Consumers + warp; row < P.norm_rows; row += (int)gridDim.x * kS2Consumers) {
      if constexpr (F16)
        rmsnorm_row<__half>(reinterpret_cast<const __half*>(P.norm_src + row * P.norm_lds), reinterpret_cast<const __half*>(P.norm_w),
                            reinterpret_cast<__half*>(P.norm_dst + row * P.norm_ldd), P.norm_hidden, P.norm_eps, lane);
      else
        rmsnorm_row<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(P.norm_src + row * P.norm_lds), reinterpret_cast<const __nv_bfloat16*>(P.norm_w),
                                   reinterpret_cast<__nv_bfloat16*>(P.norm_dst + row * P.norm_ldd), P.norm_hidden, P.norm_eps, lane);
    }
    if ((int)blockIdx.x * kS2Consumers >= P.norm_rows) return;  // only the CTAs that own rows arrive (the producers wait for exactly those)
    consumer_bar();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(P.norm_cnt, 1);
    }
  };
  auto norm_exit = [&]() {  // the last CTA to finish the launch re-arms the two counters (every producer has passed its wait by then)
    if (P.norm_rows > 0 && threadIdx.x == 0 && atomicAdd(P.norm_cnt + 1, 1) == (int)gridDim.x - 1) {
      P.norm_cnt[0] = 0;
      P.norm_cnt[1] = 0;
    }
  };
  if (it0 >= it1) {
    norm_phase();
    norm_exit();
    return;
  }

  if (warp == kS2Consumers) {
    // ------------------------------------------------------------------------------------------------ producer (one lane issues)
    // fills stage ii % stages with iteration `w`; parts: 1 = weight boxes (+ the barrier's byte count), 2 = activation boxes, 3 = both.
    // (Two other ways to fill the ring were measured and dropped, profiles/r02_decode.txt: 16-byte cp.async from four producer warps,
    // 3-4x slower, and activations by cp.async with weights on TMA, 12-20 % slower.)
    auto fill = [&](int ii, const S2Iter& w, int parts) {
      const int s = ii % P.stages;
      const S2Problem& q = P.prob[w.p];
      const SkProblem& pr = q.pr;
      const bool dual = pr.epilogue == MC_SKINNY_EPI_SILU_MUL;
      const bool phase = w.kc >= q.nkc0;
      const int kc = phase ? w.kc - q.nkc0 : w.kc, K = phase ? pr.K1 : pr.K0;
      const int k0 = kc * KC;
      const int halves = min(P.nh, (K - k0 + 63) >> 6);
      const int F = dual ? R / 2 : R, n0 = w.rb * F;
      const int nx = (dual && phase) ? 2 : 1;
      unsigned char* st = ring + (size_t)s * P.stage_bytes;
      if (elect_one()) {
        if (parts & 1) mbar_expect_tx(full + s, (uint32_t)halves * (W_HALF + (uint32_t)nx * X_HALF));
        for (int h = 0; h < halves; ++h) {
          const int k = k0 + 64 * h;
          if (parts & 1) {
            if (dual) {
              tma_load_2d(phase ? &q.tmB1 : &q.tmB0, full + s, st + h * W_HALF, k, n0);
              tma_load_2d(phase ? &q.tmB1u : &q.tmB0u, full + s, st + h * W_HALF + W_HALF / 2, k, n0);
            } else {
              tma_load_2d(phase ? &q.tmB1 : &q.tmB0, full + s, st + h * W_HALF, k, n0);
            }
          }
          if (parts & 2) {
            tma_load_2d(phase ? &q.tmA1 : &q.tmA0, full + s, st + X_BASE + h * X_HALF, k, 0);
            if (nx == 2) tma_load_2d(&q.tmA1u, full + s, st + X_BASE + X_SLOT + h * X_HALF, k, 0);
          }
        }
      }
      __syncwarp();
    };
    // Weights do not depend on the predecessor kernel: the first ring of weight boxes is requested BEFORE griddepcontrol.wait (under
    // programmatic dependent launch this CTA may be running while the previous kernel of the decode chain still is), the
    // activation boxes of those stages right after it.
    const int pre = min(P.stages, it1 - it0);
    S2Iter w = s2_locate(P, it0);
    for (int i = 0; i < pre; ++i) {
      fill(i, w, 1);
      s2_advance(P, w);
    }
    griddep_wait();
    if (P.norm_rows > 0) {
      if (lane == 0) {
        const int owners = min((int)gridDim.x, (P.norm_rows + kS2Consumers - 1) / kS2Consumers);
        while (*reinterpret_cast<volatile int*>(P.norm_cnt) < owners) {
        }
        __threadfence();
      }
      __syncwarp();
      asm volatile("fence.proxy.async.global;" ::: "memory");
    }
    w = s2_locate(P, it0);
    for (int i = 0; i < pre; ++i) {
      fill(i, w, 2);
      s2_advance(P, w);
    }
    for (int i = pre; i < it1 - it0; ++i) {
      mbar_wait(empty + i % P.stages, ((uint32_t)(i / P.stages) & 1u) ^ 1u);
      fill(i, w, 3);
      s2_advance(P, w);
    }
    return;
  }

  // -------------------------------------------------------------------------------------------------- consumers
  griddep_wait();
  norm_phase();
  const int tp = warp % TP, kp = warp / TP, g = lane >> 2, t = lane & 3;
  const int tid = threadIdx.x;  // 0 .. 255
  float acc[2][NT][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][nt][e] = 0.f;
  // ldmatrix lane addresses inside a box: A rows (lane & 7) + 8 * ((lane >> 3) & 1) of the tile, 16-byte unit (lane >> 4);
  // B rows (lane & 7) + 8 * (lane >> 4) of the activation tile pair, unit ((lane >> 3) & 1); swizzle: unit ^= row & 7 = lane & 7
  const uint32_t a_row = (uint32_t)((tp * 32 + (lane & 7) + ((lane >> 3) & 1) * 8) * 128), a_unit = (uint32_t)(lane >> 4);
  const uint32_t b_row = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * 128), b_unit = (uint32_t)((lane >> 3) & 1);
  const uint32_t swz = (uint32_t)(lane & 7);
  const uint32_t ring_u32 = smem_u32(ring);

  S2Iter w = s2_locate(P, it0);
  for (int it = it0, i = 0; it < it1; ++it, ++i) {
    const int s = i % P.stages;
    const uint32_t ph = (uint32_t)(i / P.stages) & 1u;
    const S2Problem& q = P.prob[w.p];
    const SkProblem& pr = q.pr;
    const bool dual = pr.epilogue == MC_SKINNY_EPI_SILU_MUL;
    const bool phase = w.kc >= q.nkc0;
    const int kc = phase ? w.kc - q.nkc0 : w.kc, K = phase ? pr.K1 : pr.K0;
    const int nk16 = P.copy_only ? 0 : (min(KC, K - kc * KC) + 15) >> 4;  // a partial last step reads the zeros TMA filled in
    mbar_wait(full + s, ph);
    const uint32_t st = ring_u32 + (uint32_t)s * (uint32_t)P.stage_bytes;
    // activation slot of the pair's two tiles (dual K1 chunks: up_proj tiles read the second slot)
    const bool sep = dual && phase;
    const uint32_t xs0 = (sep && (tp * 2) >= RT / 2) ? X_SLOT : 0u;
    const uint32_t xs1 = (sep && (tp * 2 + 1) >= RT / 2) ? X_SLOT : 0u;
    for (int ks = kp; ks < nk16; ks += KP) {
      const uint32_t half = (uint32_t)(ks >> 2), u0 = (uint32_t)((ks & 3) * 2);
      const uint32_t a_addr = st + half * W_HALF + a_row + (((u0 + a_unit) ^ swz) << 4);
      uint32_t a[2][4];
      ldsm_x4(a_addr, a[0][0], a[0][1], a[0][2], a[0][3]);
      ldsm_x4(a_addr + 16 * 128, a[1][0], a[1][1], a[1][2], a[1][3]);
      const uint32_t b_addr = st + X_BASE + half * X_HALF + b_row + (((u0 + b_unit) ^ swz) << 4);
      uint32_t b[2][NT][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (j == 1 && xs1 == xs0) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            b[1][nt][0] = b[0][nt][0];
            b[1][nt][1] = b[0][nt][1];
          }
        } else {
          const uint32_t xb = b_addr + (j ? xs1 : xs0);
          if constexpr (NT == 1) {
            ldsm_x2(xb, b[j][0][0], b[j][0][1]);
          } else {
#pragma unroll
            for (int nt = 0; nt < NT; nt += 2) ldsm_x4(xb + nt * 8 * 128, b[j][nt][0], b[j][nt][1], b[j][nt + 1][0], b[j][nt + 1][1]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma16816<F16>(acc[j][nt], a[j][0], a[j][1], a[j][2], a[j][3], b[j][nt][0], b[j][nt][1]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);

    const int ipr = q.nkc0 + q.nkc1;
    const bool rb_done = w.kc == ipr - 1, span_done = it == it1 - 1;
    if (rb_done || span_done) {
      // ---- the warps' partial sums of this row block -> red; fragment: c0, c1 = (weight row g, activation rows 2t, 2t + 1)
      float* mine = red + warp * (MT * LDR);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int m = nt * 8 + 2 * t, c = j * 16 + g;
          mine[m * LDR + c] = acc[j][nt][0];
          mine[(m + 1) * LDR + c] = acc[j][nt][1];
          mine[m * LDR + c + 8] = acc[j][nt][2];
          mine[(m + 1) * LDR + c + 8] = acc[j][nt][3];
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[j][nt][e] = 0.f;
        }
      consumer_bar();
      // iteration range of this row block and who shares it
      int first = w.p ? P.prob[w.p - 1].iter_end : 0;
      const int a_j = first + w.rb * ipr, b_j = a_j + ipr;
      const bool complete = it0 <= a_j && b_j <= it1;
      const int rb_global = q.rb_base + w.rb;
      float* my_slot = P.scratch + ((size_t)blockIdx.x * 2 + (((int)blockIdx.x * P.span >= a_j) ? 0 : 1)) * kS2TileFloats;
      for (int idx = tid; idx < pr.M * R; idx += kS2Consumers * 32) {
        const int m = idx / R, c = idx % R;
        const int tpc = c >> 5, cc = c & 31;
        float v = 0.f;
#pragma unroll
        for (int k2 = 0; k2 < KP; ++k2) v += red[(k2 * TP + tpc) * (MT * LDR) + m * LDR + cc];
        if (complete) tile[m * LDT + c] = v;
        else my_slot[idx] = v;
      }
      bool run_epilogue = complete;
      if (!complete) {
        // Publish the partial tile and count the arrival.  ONE thread fences: the CTA barrier orders every consumer's stores before
        // thread 0's fence, and the fence is cumulative, so they are visible at GPU scope before the counter moves (a fence in all
        // 256 threads cost ~2 us per split row block: two per CTA and launch).  Same on the reading side: atomic, fence, barrier.
        consumer_bar();
        if (tid == 0) {
          const int c_first = a_j / P.span, c_last = (b_j - 1) / P.span;
          __threadfence();
          *flag = atomicAdd(P.counters + rb_global, 1) == c_last - c_first;
          __threadfence();
        }
        consumer_bar();
        run_epilogue = *flag != 0;
        if (run_epilogue) {
          const int c_first = a_j / P.span, c_last = (b_j - 1) / P.span;
          for (int idx = tid; idx < pr.M * R; idx += kS2Consumers * 32) {
            float v = 0.f;
            for (int c = c_first; c <= c_last; ++c)
              v += __ldcg(P.scratch + ((size_t)c * 2 + ((c * P.span >= a_j) ? 0 : 1)) * kS2TileFloats + idx);
            tile[(idx / R) * LDT + idx % R] = v;
          }
          if (tid == 0) P.counters[rb_global] = 0;
        }
      }
      consumer_bar();
      if (run_epilogue) {
        const int F = dual ? R / 2 : R, n0 = w.rb * F;
        for (int idx = tid; idx < pr.M * F; idx += kS2Consumers * 32) {
          const int m = idx / F, f = idx % F, n = n0 + f;
          if (n >= pr.N) continue;
          float v = tile[m * LDT + f];
          if (pr.epilogue == MC_SKINNY_EPI_RESIDUAL) {
            v += sk_load<F16>(pr.residual + m * pr.ldr + 2ll * n);
          } else if (pr.epilogue == MC_SKINNY_EPI_COLSCALE) {
            v *= pr.col_scale[n];
          } else if (pr.epilogue == MC_SKINNY_EPI_SILU_MUL) {
            const float gate = sk_round<F16>(v);
            const float sg = sk_round<F16>(__fdividef(gate, 1.0f + __expf(-gate)));
            v = sg * sk_round<F16>(tile[m * LDT + F + f]);
          }
          sk_store<F16>(pr.C + m * pr.ldc + 2ll * n, v);
        }
      }
      consumer_bar();  // red / tile / flag are free again
    }
    s2_advance(P, w);
  }
  norm_exit();
}

// ================================================================================================ RoPE + cache append
template <typename T>
__device__ __forceinline__ void rope8(const uint4& lo, const uint4& hi, const uint4& c_lo, const uint4& c_hi, const uint4& s_lo,
                                      const uint4& s_hi, uint4& o_lo, uint4& o_hi) {
  const T* a = reinterpret_cast<const T*>(&lo);
  const T* b = reinterpret_cast<const T*>(&hi);
  const T* cl = reinterpret_cast<const T*>(&c_lo);
  const T* ch = reinterpret_cast<const T*>(&c_hi);
  const T* sl = reinterpret_cast<const T*>(&s_lo);
  const T* sh = reinterpret_cast<const T*>(&s_hi);
  T* ol = reinterpret_cast<T*>(&o_lo);
  T* oh = reinterpret_cast<T*>(&o_hi);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float x1 = to_f32<T>(a[e]), x2 = to_f32<T>(b[e]);
    // rotate_half(x) = cat(-x2, x1); every product and the sum rounded to the storage dtype (the eager ops of the reference)
    const T p1 = from_f32<T>(x1 * to_f32<T>(cl[e])), p2 = from_f32<T>(-x2 * to_f32<T>(sl[e]));
    const T p3 = from_f32<T>(x2 * to_f32<T>(ch[e])), p4 = from_f32<T>(x1 * to_f32<T>(sh[e]));
    ol[e] = from_f32<T>(to_f32<T>(p1) + to_f32<T>(p2));
    oh[e] = from_f32<T>(to_f32<T>(p3) + to_f32<T>(p4));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_rope_append_kernel(T* __restrict__ q, const T* __restrict__ k_new, const T* __restrict__ v_new, long long ld,
                          T* __restrict__ k_cache, T* __restrict__ v_cache, long long capacity, const int* __restrict__ d_pos,
                          const T* __restrict__ cos_t, const T* __restrict__ sin_t, int batch, int n_heads, int head_dim) {
  griddep_launch_dependents();
  griddep_wait();
  const int half = head_dim >> 1, vec = half >> 3;
  const int total = batch * n_heads * vec;
  const int pos = *d_pos;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / (n_heads * vec), r = i % (n_heads * vec), h = r / vec, v = r % vec;
    const long long src = b * ld + h * head_dim + v * 8;
    const long long dst = (((long long)b * n_heads + h) * capacity + pos) * head_dim + v * 8;
    const T* cr = cos_t + (long long)pos * head_dim + v * 8;
    const T* sr = sin_t + (long long)pos * head_dim + v * 8;
    const uint4 c_lo = *reinterpret_cast<const uint4*>(cr), c_hi = *reinterpret_cast<const uint4*>(cr + half);
    const uint4 s_lo = *reinterpret_cast<const uint4*>(sr), s_hi = *reinterpret_cast<const uint4*>(sr + half);
    uint4 o_lo, o_hi;
    rope8<T>(*reinterpret_cast<const uint4*>(q + src), *reinterpret_cast<const uint4*>(q + src + half), c_lo, c_hi, s_lo, s_hi, o_lo, o_hi);
    *reinterpret_cast<uint4*>(q + src) = o_lo;
    *reinterpret_cast<uint4*>(q + src + half) = o_hi;
    rope8<T>(*reinterpret_cast<const uint4*>(k_new + src), *reinterpret_cast<const uint4*>(k_new + src + half), c_lo, c_hi, s_lo, s_hi,
             o_lo, o_hi);
    *reinterpret_cast<uint4*>(k_cache + dst) = o_lo;
    *reinterpret_cast<uint4*>(k_cache + dst + half) = o_hi;
    *reinterpret_cast<uint4*>(v_cache + dst) = *reinterpret_cast<const uint4*>(v_new + src);
    *reinterpret_cast<uint4*>(v_cache + dst + half) = *reinterpret_cast<const uint4*>(v_new + src + half);
  }
}

// ================================================================================================ decode attention
constexpr int kDaThreads = 128;  // 8 half-warps, one key each at a time
constexpr int kDaD = 128;
constexpr int kDaUnroll = 2;     // keys per half-warp and block; two blocks in flight: 8 x 16 B per lane (k and v)

struct DaParams {
  const char* q;
  const char* k_cache;
  const char* v_cache;
  const unsigned char* key_mask;
  char* out;
  float* scratch;
  int* counters;
  const int* d_pos;
  long long capacity, ld_mask, ld_q, ld_out;  // ld_q / ld_out in bytes
  int n_heads, n_splits;
  float scale_log2e;
  // fused form (mc_decode_attention_fused): q / k_new / v_new are the raw projection outputs; the kernel rotates q and k_new itself
  // (rounding points of mc_rope), the CTA whose key range holds the new position appends k / v to the cache, and that key is read
  // from shared memory rather than back from the cache
  const char* k_new;
  const char* v_new;
  const char* cos_t;
  const char* sin_t;
  int fused;
};

template <typename T>
__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
  const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = to_f32<T>(e[i]);
}

template <typename T>
__global__ void __launch_bounds__(kDaThreads) decode_attention_kernel(const __grid_constant__ DaParams P) {
  __shared__ float sm_acc[8][kDaD];
  __shared__ float sm_m[8], sm_l[8];
  __shared__ int sm_last;
  griddep_launch_dependents();
  griddep_wait();
  const int bh = blockIdx.x, b = bh / P.n_heads, h = bh % P.n_heads, split = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, l16 = lane & 15, hw = tid >> 4;
  const int L = *P.d_pos + 1;
  int chunk = (L + P.n_splits - 1) / P.n_splits;
  chunk = (chunk + 8 * kDaUnroll - 1) / (8 * kDaUnroll) * (8 * kDaUnroll);
  const int start = split * chunk, end = min(L, start + chunk);

  // caches are [batch, heads, capacity, D]: the keys of one (sequence, head) are one contiguous stream, 512 B per warp-level load
  const long long key_stride = kDaD * 2;
  char* kc_row = const_cast<char*>(P.k_cache) + (long long)bh * P.capacity * kDaD * 2;
  char* vc_row = const_cast<char*>(P.v_cache) + (long long)bh * P.capacity * kDaD * 2;
  const char* kb = kc_row + l16 * 16;
  const char* vb = vc_row + l16 * 16;
  const unsigned char* mrow = P.key_mask ? P.key_mask + b * P.ld_mask : nullptr;
  const int pos = L - 1;
  __shared__ __align__(16) T s_q[kDaD], s_k[kDaD], s_v[kDaD];
  float qf[8];
  if (P.fused) {
    // RoPE of the new token's q and k for this (sequence, head): thread i < 64 rotates the pair (i, i + 64) with every product and
    // the sum rounded to the storage dtype (the eager ops of the reference, as rope8 / mc_rope); threads 64 .. 127 copy v
    const long long off = b * P.ld_q + h * kDaD * 2ll;
    if (tid < kDaD / 2) {
      const T* qp = reinterpret_cast<const T*>(P.q + off);
      const T* kp = reinterpret_cast<const T*>(P.k_new + off);
      const T* cr = reinterpret_cast<const T*>(P.cos_t) + (long long)pos * kDaD;
      const T* sr = reinterpret_cast<const T*>(P.sin_t) + (long long)pos * kDaD;
      const float cl = to_f32<T>(cr[tid]), ch = to_f32<T>(cr[tid + kDaD / 2]), sl = to_f32<T>(sr[tid]), sh = to_f32<T>(sr[tid + kDaD / 2]);
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const T* xp = which ? kp : qp;
        const float x1 = to_f32<T>(xp[tid]), x2 = to_f32<T>(xp[tid + kDaD / 2]);
        const T p1 = from_f32<T>(x1 * cl), p2 = from_f32<T>(-x2 * sl), p3 = from_f32<T>(x2 * ch), p4 = from_f32<T>(x1 * sh);
        T* dst = which ? s_k : s_q;
        dst[tid] = from_f32<T>(to_f32<T>(p1) + to_f32<T>(p2));
        dst[tid + kDaD / 2] = from_f32<T>(to_f32<T>(p3) + to_f32<T>(p4));
      }
    } else {
      const T* vp = reinterpret_cast<const T*>(P.v_new + off);
      const int i = (tid - kDaD / 2) * 2;
      s_v[i] = vp[i];
      s_v[i + 1] = vp[i + 1];
    }
    __syncthreads();
    if (pos >= start && pos < end && tid < 32) {  // append: 16 lanes x 16 bytes each for k and v
      const int i = (tid & 15) * 8;
      if (tid < 16) *reinterpret_cast<uint4*>(kc_row + pos * key_stride + i * 2) = *reinterpret_cast<const uint4*>(s_k + i);
      else *reinterpret_cast<uint4*>(vc_row + pos * key_stride + i * 2) = *reinterpret_cast<const uint4*>(s_v + i);
    }
    unpack8f<T>(*reinterpret_cast<const uint4*>(s_q + l16 * 8), qf);
  } else {
    unpack8f<T>(*reinterpret_cast<const uint4*>(P.q + b * P.ld_q + (h * kDaD + l16 * 8) * 2ll), qf);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) qf[e] *= P.scale_log2e;

  float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  struct Block {
    uint4 kv[kDaUnroll], vv[kDaUnroll];
    bool ok[kDaUnroll];
  };
  auto load = [&](Block& blk, int base) {
#pragma unroll
    for (int u = 0; u < kDaUnroll; ++u) {
      const int j = base + u * 8 + hw;
      blk.ok[u] = j < end && (mrow == nullptr || mrow[j] != 0);
      blk.kv[u] = blk.vv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (j < end) {
        if (P.fused && j == pos) {  // the new token's own key: from shared memory, never back from the cache in the same launch
          blk.kv[u] = *reinterpret_cast<const uint4*>(s_k + l16 * 8);
          blk.vv[u] = *reinterpret_cast<const uint4*>(s_v + l16 * 8);
        } else {
          blk.kv[u] = ld_w16(kb + j * key_stride);
          blk.vv[u] = ld_w16(vb + j * key_stride);
        }
      }
    }
  };
  auto compute = [&](const Block& blk) {
    float s[kDaUnroll];
#pragma unroll
    for (int u = 0; u < kDaUnroll; ++u) {
      float kf[8];
      unpack8f<T>(blk.kv[u], kf);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) d = fmaf(qf[e], kf[e], d);
      s[u] = d;
    }
#pragma unroll
    for (int sh = 8; sh; sh >>= 1)
#pragma unroll
      for (int u = 0; u < kDaUnroll; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], sh);
    float gm = m;
#pragma unroll
    for (int u = 0; u < kDaUnroll; ++u) {
      s[u] = blk.ok[u] ? s[u] : -INFINITY;
      gm = fmaxf(gm, s[u]);
    }
    const float mm = gm == -INFINITY ? 0.f : gm;
    const float corr = exp2f(m - mm);
    l *= corr;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
    for (int u = 0; u < kDaUnroll; ++u) {
      const float pw = exp2f(s[u] - mm);
      float vf[8];
      unpack8f<T>(blk.vv[u], vf);
      l += pw;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pw, vf[e], acc[e]);
    }
    m = gm;
  };
  // two blocks of 4 keys per half-warp in flight: the loads of block i + 1 are issued before the arithmetic of block i
  constexpr int kStep = 8 * kDaUnroll;
  Block b0, b1;
  if (start < end) load(b0, start);
  for (int base = start; base < end; base += 2 * kStep) {
    if (base + kStep < end) load(b1, base + kStep);
    compute(b0);
    if (base + 2 * kStep < end) load(b0, base + 2 * kStep);
    if (base + kStep < end) compute(b1);
  }
  if (l16 == 0) {
    sm_m[hw] = m;
    sm_l[hw] = l;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sm_acc[hw][l16 * 8 + e] = acc[e];
  __syncthreads();
  // thread d combines the 8 half-warp states of output dimension d
  float gm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) gm = fmaxf(gm, sm_m[i]);
  const float mm = gm == -INFINITY ? 0.f : gm;
  float lsum = 0.f, o = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float w = exp2f(sm_m[i] - mm);
    lsum = fmaf(sm_l[i], w, lsum);
    o = fmaf(sm_acc[i][tid], w, o);
  }
  char* orow = P.out + b * P.ld_out + (h * kDaD + tid) * 2ll;
  if (P.n_splits == 1) {
    *reinterpret_cast<T*>(orow) = from_f32<T>(lsum > 0.f ? o / lsum : 0.f);
    return;
  }
  float* part = P.scratch + ((long long)bh * P.n_splits + split) * (kDaD + 2);
  part[tid] = o;
  if (tid == 0) {
    part[kDaD] = gm;
    part[kDaD + 1] = lsum;
  }
  __syncthreads();
  if (tid == 0) {  // one thread fences: the barrier orders the CTA's stores before it, and the fence is cumulative
    __threadfence();
    sm_last = atomicAdd(P.counters + bh, 1) == P.n_splits - 1;
    __threadfence();
  }
  __syncthreads();
  if (!sm_last) return;
  const float* all = P.scratch + (long long)bh * P.n_splits * (kDaD + 2);
  float tm = -INFINITY;
  for (int sp = 0; sp < P.n_splits; ++sp) tm = fmaxf(tm, __ldcg(all + sp * (kDaD + 2) + kDaD));
  const float tmm = tm == -INFINITY ? 0.f : tm;
  float tl = 0.f, to = 0.f;
  for (int sp = 0; sp < P.n_splits; ++sp) {
    const float w = exp2f(__ldcg(all + sp * (kDaD + 2) + kDaD) - tmm);
    tl = fmaf(__ldcg(all + sp * (kDaD + 2) + kDaD + 1), w, tl);
    to = fmaf(__ldcg(all + sp * (kDaD + 2) + tid), w, to);
  }
  *reinterpret_cast<T*>(orow) = from_f32<T>(tl > 0.f ? to / tl : 0.f);
  if (tid == 0) P.counters[bh] = 0;  // ready for the next step (graph replay)
}

// ================================================================================================ argmax
template <typename T>
__global__ void __launch_bounds__(256) argmax_rows_kernel(const T* __restrict__ logits, long long ld, int cols, int* __restrict__ out_i32,
                                                          long long* __restrict__ out_i64, int* __restrict__ counter) {
  __shared__ float sv[8];
  __shared__ int si[8];
  griddep_launch_dependents();
  griddep_wait();
  const T* row = logits + (long long)blockIdx.x * ld;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  const int n_vec = cols >> 3;
  for (int i = threadIdx.x; i < n_vec; i += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + i * 8);
    const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float f = to_f32<T>(e[k]);
      if (f > best) {  // ascending scan: the first maximum of this thread wins
        best = f;
        idx = i * 8 + k;
      }
    }
  }
  for (int i = (n_vec << 3) + threadIdx.x; i < cols; i += 256) {
    const float f = to_f32<T>(row[i]);
    if (f > best || (f == best && i < idx)) {
      best = f;
      idx = i;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, d);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
    if (ob > best || (ob == best && oi < idx)) {
      best = ob;
      idx = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = idx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < idx)) {
        best = sv[w];
        idx = si[w];
      }
    if (idx == 0x7fffffff) idx = 0;  // a row of NaNs / -inf
    if (out_i32) out_i32[blockIdx.x] = idx;
    if (out_i64) out_i64[blockIdx.x] = idx;
    if (counter && blockIdx.x == 0) *counter += 1;  // step counter of a captured decode loop (nothing else touches it meanwhile)
  }
}

}  // namespace mc

using namespace mc;

template <int RT, int NT, bool F16>
static cudaError_t launch_skinny(const SkParams& P, int grid, cudaStream_t stream) {
  constexpr size_t smem = (size_t)kSkWarps * (8 * NT) * (16 * RT + 1) * sizeof(float);
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(skinny_linear_kernel<RT, NT, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  return launch_kernel(skinny_linear_kernel<RT, NT, F16>, dim3(grid), dim3(kSkThreads), smem, stream, P);
}

template <int RT, bool F16>
static cudaError_t launch_skinny_nt(const SkParams& P, int grid, int nt, cudaStream_t stream) {
  switch (nt) {
    case 1: return launch_skinny<RT, 1, F16>(P, grid, stream);
    case 2: return launch_skinny<RT, 2, F16>(P, grid, stream);
    case 4: return launch_skinny<RT, 4, F16>(P, grid, stream);
    default: return launch_skinny<RT, 8, F16>(P, grid, stream);
  }
}

template <int RT, int NT, bool F16>
static cudaError_t launch_streamk(const S2Params& P, int grid, size_t smem, cudaStream_t stream) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(skinny_streamk_kernel<RT, NT, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  return launch_kernel(skinny_streamk_kernel<RT, NT, F16>, dim3(grid), dim3(kS2Threads), smem, stream, P);
}

template <int RT, bool F16>
static cudaError_t launch_streamk_nt(const S2Params& P, int nt, int grid, size_t smem, cudaStream_t stream) {
  switch (nt) {
    case 1: return launch_streamk<RT, 1, F16>(P, grid, smem, stream);
    case 2: return launch_streamk<RT, 2, F16>(P, grid, smem, stream);
    case 4: return launch_streamk<RT, 4, F16>(P, grid, smem, stream);
    default: return launch_streamk<RT, 8, F16>(P, grid, smem, stream);
  }
}

constexpr int kS2MaxRowBlocks = 8192;

struct mc_skinny_plan {
  SkParams reg;   // register kernel (fallback / A-B)
  S2Params sk;    // stream-K kernel
  int n_prob, reg_ctas, reg_rt, nt, rt2, grid;
  bool streamk, f16;
  size_t smem;
  long long bytes;
};

extern "C" size_t mc_skinny_workspace_bytes(void) {
  const int sms = sm_count();
  return (size_t)kS2MaxRowBlocks * sizeof(int) + (size_t)std::max(sms, 1) * 2 * kS2TileFloats * sizeof(float);
}

extern "C" int mc_skinny_plan_create(mc_skinny_plan_t** out, const mc_skinny_desc_t* desc, int n_problems, int dtype, int tuning) {
  MC_REQUIRE(out != nullptr, "skinny plan out-pointer is NULL");
  *out = nullptr;
  MC_REQUIRE(desc != nullptr && n_problems >= 1 && n_problems <= kSkMaxProb, "skinny linear: n_problems %d outside [1, %d]", n_problems, kSkMaxProb);
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "skinny linear: dtype must be bf16 or fp16");
  mc_skinny_plan* p = new (std::nothrow) mc_skinny_plan();
  if (!p) return fail(MC_ERR_NOMEM, "host allocation failed");
  memset(&p->reg, 0, sizeof(p->reg));
  memset(&p->sk, 0, sizeof(p->sk));
  SkParams& P = p->reg;
  P.n_prob = p->n_prob = n_problems;
  p->f16 = dtype == MC_F16;
  p->bytes = 0;
  bool any_dual = false, dual_k1 = false;
  int max_m = 0;
  for (int i = 0; i < n_problems; ++i) {
    const bool dual = desc[i].epilogue == MC_SKINNY_EPI_SILU_MUL;
    any_dual |= dual;
    dual_k1 |= dual && desc[i].K1 > 0;
    max_m = std::max(max_m, (int)desc[i].M);
  }
  // tuning bits 0-3: row tiles per CTA of the register kernel (0 = default: 2, i.e. 32 weight rows per CTA; 1 = 16 rows; the dual
  // problem needs 2); bit 4: run the register kernel instead of the stream-K kernel; bit 5: stream-K with 32-row blocks instead of 64
  int rt = (tuning & 0xf) == 1 ? 1 : 2;
  if (any_dual) rt = 2;
  p->reg_rt = rt;
  int ctas = 0;
  int rc = MC_OK;
#define SK_REQUIRE(cond, ...)               \
  if (!(cond)) {                            \
    rc = fail(MC_ERR_INVALID, __VA_ARGS__); \
    break;                                  \
  }
  for (int i = 0; i < n_problems; ++i) {
    const mc_skinny_desc_t& d = desc[i];
    SK_REQUIRE(d.M >= 1 && d.M <= MC_SKINNY_MAX_M, "skinny linear: problem %d: M %d outside [1, %d]", i, d.M, MC_SKINNY_MAX_M);
    SK_REQUIRE(d.N >= 8 && d.N % 8 == 0 && d.K0 >= 8 && d.K0 % 8 == 0 && d.K1 >= 0 && d.K1 % 8 == 0,
               "skinny linear: problem %d: N, K0, K1 must be multiples of 8", i);
    SK_REQUIRE(d.A0 && d.B0 && d.C, "skinny linear: problem %d: A0 / B0 / C is NULL", i);
    SK_REQUIRE(d.lda0 >= d.K0 && d.ldb0 >= d.K0 && d.ldc >= d.N && d.lda0 % 8 == 0 && d.ldb0 % 8 == 0,
               "skinny linear: problem %d: leading dimensions must cover the row and be multiples of 8 elements", i);
    SK_REQUIRE(d.K1 == 0 || (d.A1 && d.B1 && d.lda1 >= d.K1 && d.ldb1 >= d.K1 && d.lda1 % 8 == 0 && d.ldb1 % 8 == 0),
               "skinny linear: problem %d: K1 > 0 needs A1 / B1 with valid leading dimensions", i);
    SK_REQUIRE(d.epilogue >= MC_SKINNY_EPI_NONE && d.epilogue <= MC_SKINNY_EPI_SILU_MUL, "skinny linear: problem %d: bad epilogue %d", i, d.epilogue);
    SK_REQUIRE(d.epilogue != MC_SKINNY_EPI_RESIDUAL || (d.residual && d.ldr >= d.N), "skinny linear: problem %d: residual missing", i);
    SK_REQUIRE(d.epilogue != MC_SKINNY_EPI_COLSCALE || d.col_scale, "skinny linear: problem %d: col_scale missing", i);
    const bool dual = d.epilogue == MC_SKINNY_EPI_SILU_MUL;
    SK_REQUIRE(!dual || (d.B0u && (d.K1 == 0 || (d.A1u && d.B1u))), "skinny linear: problem %d: SILU_MUL needs B0u (and A1u / B1u when K1 > 0)", i);
    SK_REQUIRE((((uintptr_t)d.A0 | (uintptr_t)d.B0 | (uintptr_t)d.A1 | (uintptr_t)d.B1 | (uintptr_t)d.B0u | (uintptr_t)d.A1u |
                 (uintptr_t)d.B1u) & 15) == 0 && (((uintptr_t)d.C | (uintptr_t)d.residual) & 1) == 0,
               "skinny linear: problem %d: operand pointers must be 16-byte aligned", i);
    SkProblem& pr = P.prob[i];
    pr.A0 = (const char*)d.A0; pr.B0 = (const char*)d.B0; pr.A1 = (const char*)d.A1; pr.B1 = (const char*)d.B1;
    pr.B0u = (const char*)d.B0u; pr.A1u = (const char*)d.A1u; pr.B1u = (const char*)d.B1u;
    pr.C = (char*)d.C; pr.residual = (const char*)d.residual; pr.col_scale = d.col_scale;
    pr.lda0 = d.lda0 * 2; pr.ldb0 = d.ldb0 * 2; pr.lda1 = d.lda1 * 2; pr.ldb1 = d.ldb1 * 2; pr.ldc = d.ldc * 2; pr.ldr = d.ldr * 2;
    pr.M = d.M; pr.N = d.N; pr.K0 = d.K0; pr.K1 = d.K1; pr.epilogue = d.epilogue;
    const int F = dual ? 16 : 16 * rt;
    ctas += (d.N + F - 1) / F;
    pr.cta_end = ctas;
    p->bytes += 2ll * d.N * (d.K0 + d.K1) * (dual ? 2 : 1);
  }
#undef SK_REQUIRE
  if (rc != MC_OK) {
    delete p;
    return rc;
  }
  p->reg_ctas = ctas;
  p->nt = max_m <= 8 ? 1 : (max_m <= 16 ? 2 : (max_m <= 32 ? 4 : 8));
  p->streamk = !((tuning >> 4) & 1);
  if (p->streamk) {
    const int sms = sm_count();
    if (sms <= 0) {
      delete p;
      return fail(MC_ERR_CUDA, "no CUDA device");
    }
    // row block: 64 weight rows; tuning bit 5: 32.  (128-row blocks — half the activation bytes per weight byte, but a ring of
    // 2-4 stages — measured 10-25 % slower at M >= 16: profiles/r02_decode.txt)
    p->rt2 = ((tuning >> 5) & 1) ? 2 : 4;
    const int R = 16 * p->rt2, MT = 8 * p->nt;
    S2Params& Q = p->sk;
    Q.n_prob = n_problems;
    // K chunk: 256 elements (four TMA boxes per operand and stage: 512 contiguous bytes per weight row) when the ring still gets
    // three stages, else 128; tuning bit 7 forces 128.  Measured (profiles/r02_decode.txt): 256 is 10-20 % faster at M <= 32.
    Q.copy_only = (tuning >> 6) & 1;              // tuning bit 6: profiling aid, results are garbage
    const size_t budget = 227 * 1024;
    Q.xslots = dual_k1 ? 2 : 1;
    {
      const size_t fixed256 = 1024 + 256 + (size_t)kS2Consumers * MT * 33 * 4 + (size_t)MT * (R + 1) * 4;
      const size_t stage256 = 4 * (size_t)(R * 128 + Q.xslots * MT * 128);
      Q.nh = (!((tuning >> 7) & 1) && (budget - fixed256) / stage256 >= 3) ? 4 : 2;
    }
    const int kc_elems = 64 * Q.nh;
    int iters = 0, rbs = 0;
    for (int i = 0; i < n_problems && rc == MC_OK; ++i) {
      S2Problem& q = Q.prob[i];
      const mc_skinny_desc_t& d = desc[i];
      q.pr = P.prob[i];
      const bool dual = d.epilogue == MC_SKINNY_EPI_SILU_MUL;
      const int F = dual ? R / 2 : R;
      q.nrb = (d.N + F - 1) / F;
      q.nkc0 = (d.K0 + kc_elems - 1) / kc_elems;
      q.nkc1 = (d.K1 + kc_elems - 1) / kc_elems;
      q.rb_base = rbs;
      rbs += q.nrb;
      iters += q.nrb * (q.nkc0 + q.nkc1);
      q.iter_end = iters;
      rc = encode_operand(&q.tmB0, d.B0, d.N, d.K0, d.ldb0, F, dtype);
      if (rc == MC_OK) rc = encode_operand(&q.tmA0, d.A0, d.M, d.K0, d.lda0, MT, dtype);
      if (rc == MC_OK && dual) rc = encode_operand(&q.tmB0u, d.B0u, d.N, d.K0, d.ldb0, F, dtype);
      if (rc == MC_OK && d.K1 > 0) rc = encode_operand(&q.tmB1, d.B1, d.N, d.K1, d.ldb1, F, dtype);
      if (rc == MC_OK && d.K1 > 0) rc = encode_operand(&q.tmA1, d.A1, d.M, d.K1, d.lda1, MT, dtype);
      if (rc == MC_OK && d.K1 > 0 && dual) rc = encode_operand(&q.tmB1u, d.B1u, d.N, d.K1, d.ldb1, F, dtype);
      if (rc == MC_OK && d.K1 > 0 && dual) rc = encode_operand(&q.tmA1u, d.A1u, d.M, d.K1, d.lda1, MT, dtype);
    }
    if (rc == MC_OK && rbs > kS2MaxRowBlocks - 2) rc = fail(MC_ERR_INVALID, "skinny linear: %d row blocks exceed the workspace (%d)", rbs, kS2MaxRowBlocks - 2);
    if (rc != MC_OK) {
      delete p;
      return rc;
    }
    Q.total_iters = iters;
    Q.stage_bytes = Q.nh * (R * 128 + Q.xslots * MT * 128);
    const size_t fixed = 1024 + 256 + (size_t)kS2Consumers * MT * 33 * 4 + (size_t)MT * (R + 1) * 4;
    Q.stages = (int)std::min<size_t>(kS2MaxStages, (budget - fixed) / Q.stage_bytes);
    if ((tuning >> 8) & 0xf) Q.stages = std::min(Q.stages, (tuning >> 8) & 0xf);  // tuning bits 8-11: cap the ring depth
    if (Q.stages < 2) {
      delete p;
      return fail(MC_ERR_UNSUPPORTED, "skinny linear: shared memory too small for a 2-stage ring");
    }
    p->smem = fixed + (size_t)Q.stages * Q.stage_bytes;
    p->grid = std::min(sms, iters);
    Q.span = (iters + p->grid - 1) / p->grid;
  }
  *out = p;
  return MC_OK;
}

extern "C" int mc_skinny_plan_run(const mc_skinny_plan_t* p, void* workspace, size_t workspace_bytes, mc_stream_t stream) {
  MC_REQUIRE(p != nullptr, "skinny plan is NULL");
  cudaError_t e;
  if (p->streamk) {
    MC_REQUIRE(workspace != nullptr && workspace_bytes >= mc_skinny_workspace_bytes() && ((uintptr_t)workspace & 15) == 0,
               "skinny linear: workspace missing, smaller than mc_skinny_workspace_bytes() or misaligned");
    S2Params Q = p->sk;
    Q.counters = (int*)workspace;
    Q.norm_cnt = Q.counters + kS2MaxRowBlocks - 2;  // the last two counters (row blocks use at most kS2MaxRowBlocks - 2)
    Q.scratch = (float*)((char*)workspace + (size_t)kS2MaxRowBlocks * sizeof(int));
    if (p->rt2 == 2) e = p->f16 ? launch_streamk_nt<2, true>(Q, p->nt, p->grid, p->smem, (cudaStream_t)stream)
                                : launch_streamk_nt<2, false>(Q, p->nt, p->grid, p->smem, (cudaStream_t)stream);
    else e = p->f16 ? launch_streamk_nt<4, true>(Q, p->nt, p->grid, p->smem, (cudaStream_t)stream)
                    : launch_streamk_nt<4, false>(Q, p->nt, p->grid, p->smem, (cudaStream_t)stream);
  } else if (p->reg_rt == 1) {
    e = p->f16 ? launch_skinny_nt<1, true>(p->reg, p->reg_ctas, p->nt, (cudaStream_t)stream)
               : launch_skinny_nt<1, false>(p->reg, p->reg_ctas, p->nt, (cudaStream_t)stream);
  } else {
    e = p->f16 ? launch_skinny_nt<2, true>(p->reg, p->reg_ctas, p->nt, (cudaStream_t)stream)
               : launch_skinny_nt<2, false>(p->reg, p->reg_ctas, p->nt, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "skinny linear launch failed: %s", cudaGetErrorString(e));
  return MC_OK;
}

extern "C" int mc_skinny_plan_set_norm(mc_skinny_plan_t* p, const void* src, int64_t ld_src, const void* weight, void* dst, int64_t ld_dst,
                                       int rows, int hidden, float eps) {
  MC_REQUIRE(p != nullptr, "skinny plan is NULL");
  MC_REQUIRE(p->streamk, "skinny linear: the fused RMSNorm needs the stream-K kernel (not tuning bit 4)");
  MC_REQUIRE(src && weight && dst && rows >= 1 && rows <= MC_SKINNY_MAX_M && hidden >= 8 && hidden % 8 == 0 && ld_src >= hidden &&
                 ld_dst >= hidden && ld_src % 8 == 0 && ld_dst % 8 == 0, "skinny linear: bad RMSNorm shape");
  MC_REQUIRE((((uintptr_t)src | (uintptr_t)weight | (uintptr_t)dst) & 15) == 0, "skinny linear: RMSNorm pointers must be 16-byte aligned");
  S2Params& Q = p->sk;
  Q.norm_src = (const char*)src;
  Q.norm_w = (const char*)weight;
  Q.norm_dst = (char*)dst;
  Q.norm_lds = ld_src * 2;
  Q.norm_ldd = ld_dst * 2;
  Q.norm_rows = rows;
  Q.norm_hidden = hidden;
  Q.norm_eps = eps;
  return MC_OK;
}

extern "C" int64_t mc_skinny_plan_bytes(const mc_skinny_plan_t* p) { return p ? p->bytes : 0; }

extern "C" int mc_skinny_plan_destroy(mc_skinny_plan_t* p) {
  delete p;
  return MC_OK;
}

extern "C" int mc_decode_rope_append(void* q, const void* k_new, const void* v_new, int64_t ld_qkv, void* k_cache, void* v_cache,
                                     int64_t capacity, const int32_t* d_pos, const void* cos_table, const void* sin_table, int batch,
                                     int n_heads, int head_dim, int dtype, mc_stream_t stream) {
  MC_REQUIRE(q && k_new && v_new && k_cache && v_cache && d_pos && cos_table && sin_table, "decode rope/append: NULL pointer");
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "decode rope/append: dtype must be bf16 or fp16");
  MC_REQUIRE(batch >= 1 && n_heads >= 1 && head_dim >= 16 && head_dim % 16 == 0 && capacity >= 1 && ld_qkv % 8 == 0 &&
                 ld_qkv >= (int64_t)n_heads * head_dim, "decode rope/append: bad shape");
  MC_REQUIRE((((uintptr_t)q | (uintptr_t)k_new | (uintptr_t)v_new | (uintptr_t)k_cache | (uintptr_t)v_cache | (uintptr_t)cos_table |
               (uintptr_t)sin_table) & 15) == 0, "decode rope/append: pointers must be 16-byte aligned");
  const int total = batch * n_heads * (head_dim / 16);
  const int grid = (total + 255) / 256;
  if (dtype == MC_BF16)
    MC_CUDA_OK(launch_kernel(decode_rope_append_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (__nv_bfloat16*)q,
                             (const __nv_bfloat16*)k_new, (const __nv_bfloat16*)v_new, (long long)ld_qkv, (__nv_bfloat16*)k_cache,
                             (__nv_bfloat16*)v_cache, (long long)capacity, (const int*)d_pos, (const __nv_bfloat16*)cos_table,
                             (const __nv_bfloat16*)sin_table, batch, n_heads, head_dim));
  else
    MC_CUDA_OK(launch_kernel(decode_rope_append_kernel<__half>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (__half*)q, (const __half*)k_new,
                             (const __half*)v_new, (long long)ld_qkv, (__half*)k_cache, (__half*)v_cache, (long long)capacity,
                             (const int*)d_pos, (const __half*)cos_table, (const __half*)sin_table, batch, n_heads, head_dim));
  return MC_OK;
}

static int decode_attention_launch(const void* q, const void* k_new, const void* v_new, const void* cos_table, const void* sin_table, int fused,
                                   const void* k_cache, const void* v_cache, int64_t capacity, const int32_t* d_pos, const uint8_t* key_mask,
                                   int64_t ld_mask, void* out, int64_t ld_q, int64_t ld_out, int batch, int n_heads, int head_dim,
                                   float softmax_scale, int n_splits, float* scratch, int32_t* counters, int dtype, mc_stream_t stream) {
  MC_REQUIRE(q && k_cache && v_cache && d_pos && out, "decode attention: NULL pointer");
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "decode attention: dtype must be bf16 or fp16");
  MC_REQUIRE(head_dim == kDaD, "decode attention: head_dim must be %d", kDaD);
  MC_REQUIRE(batch >= 1 && n_heads >= 1 && capacity >= 1 && n_splits >= 1 && n_splits <= 65535, "decode attention: bad shape");
  MC_REQUIRE(n_splits == 1 || (scratch && counters), "decode attention: n_splits > 1 needs scratch and counters");
  MC_REQUIRE(key_mask == nullptr || ld_mask >= capacity, "decode attention: ld_mask must cover the capacity");
  MC_REQUIRE((((uintptr_t)q | (uintptr_t)k_cache | (uintptr_t)v_cache) & 15) == 0 && ld_q % 8 == 0, "decode attention: q / caches must be 16-byte aligned");
  MC_REQUIRE(!fused || (k_new && v_new && cos_table && sin_table && (((uintptr_t)k_new | (uintptr_t)v_new) & 15) == 0),
             "decode attention: the fused form needs k_new, v_new and the cos / sin tables");
  DaParams P;
  memset(&P, 0, sizeof(P));
  P.q = (const char*)q; P.k_cache = (const char*)k_cache; P.v_cache = (const char*)v_cache; P.key_mask = key_mask;
  P.out = (char*)out; P.scratch = scratch; P.counters = counters; P.d_pos = d_pos;
  P.capacity = capacity; P.ld_mask = ld_mask; P.ld_q = ld_q * 2; P.ld_out = ld_out * 2;
  P.n_heads = n_heads; P.n_splits = n_splits;
  P.scale_log2e = softmax_scale * 1.4426950408889634f;
  P.k_new = (const char*)k_new; P.v_new = (const char*)v_new; P.cos_t = (const char*)cos_table; P.sin_t = (const char*)sin_table;
  P.fused = fused;
  const dim3 grid((unsigned)(batch * n_heads), (unsigned)n_splits);
  if (dtype == MC_BF16) MC_CUDA_OK(launch_kernel(decode_attention_kernel<__nv_bfloat16>, grid, dim3(kDaThreads), 0, (cudaStream_t)stream, P));
  else MC_CUDA_OK(launch_kernel(decode_attention_kernel<__half>, grid, dim3(kDaThreads), 0, (cudaStream_t)stream, P));
  return MC_OK;
}

extern "C" int mc_decode_attention(const void* q, const void* k_cache, const void* v_cache, int64_t capacity, const int32_t* d_pos,
                                   const uint8_t* key_mask, int64_t ld_mask, void* out, int64_t ld_q, int64_t ld_out, int batch,
                                   int n_heads, int head_dim, float softmax_scale, int n_splits, float* scratch, int32_t* counters,
                                   int dtype, mc_stream_t stream) {
  return decode_attention_launch(q, nullptr, nullptr, nullptr, nullptr, 0, k_cache, v_cache, capacity, d_pos, key_mask, ld_mask, out, ld_q, ld_out,
                                 batch, n_heads, head_dim, softmax_scale, n_splits, scratch, counters, dtype, stream);
}

extern "C" int mc_decode_attention_fused(const void* q, const void* k_new, const void* v_new, int64_t ld_qkv, void* k_cache, void* v_cache,
                                         int64_t capacity, const int32_t* d_pos, const void* cos_table, const void* sin_table,
                                         const uint8_t* key_mask, int64_t ld_mask, void* out, int64_t ld_out, int batch, int n_heads,
                                         int head_dim, float softmax_scale, int n_splits, float* scratch, int32_t* counters, int dtype,
                                         mc_stream_t stream) {
  return decode_attention_launch(q, k_new, v_new, cos_table, sin_table, 1, k_cache, v_cache, capacity, d_pos, key_mask, ld_mask, out, ld_qkv,
                                 ld_out, batch, n_heads, head_dim, softmax_scale, n_splits, scratch, counters, dtype, stream);
}

extern "C" int mc_argmax_rows(const void* logits, int64_t ld, int rows, int cols, int32_t* out_i32, int64_t* out_i64,
                              int32_t* d_counter, int dtype, mc_stream_t stream) {
  MC_REQUIRE(logits && (out_i32 || out_i64), "argmax: NULL pointer");
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "argmax: dtype must be bf16 or fp16");
  MC_REQUIRE(rows >= 0 && cols >= 1 && ld >= cols && ld % 8 == 0 && ((uintptr_t)logits & 15) == 0, "argmax: bad shape or alignment");
  if (rows == 0) return MC_OK;
  if (dtype == MC_BF16)
    MC_CUDA_OK(launch_kernel(argmax_rows_kernel<__nv_bfloat16>, dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const __nv_bfloat16*)logits,
                             (long long)ld, cols, (int*)out_i32, (long long*)out_i64, (int*)d_counter));
  else
    MC_CUDA_OK(launch_kernel(argmax_rows_kernel<__half>, dim3(rows), dim3(256), 0, (cudaStream_t)stream, (const __half*)logits, (long long)ld,
                             cols, (int*)out_i32, (long long*)out_i64, (int*)d_counter));
  return MC_OK;
}
