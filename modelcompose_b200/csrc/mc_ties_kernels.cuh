// Kernel templates of the TIES merge (trim / elect sign / disjoint merge); contract in mc_ties.cu.
#pragma once
#include <type_traits>

#include "mc_merge_kernels.cuh"

namespace mc {

// Device-resident state of one TIES run (radix-select progress, thresholds, sign census).
struct TiesState {
  unsigned long long k_rem[MC_MERGE_MAX_SRC];  // rank still to find inside the current prefix bucket (1-based)
  unsigned int prefix[MC_MERGE_MAX_SRC];       // magnitude key bits fixed so far
  float thr[MC_MERGE_MAX_SRC];                 // k-th smallest |x| per source (final)
  unsigned long long n_pos, n_neg, n_zero, n_amb;  // elected-sign census of the speculative merge pass
  int majority;                                // sign(n_pos - n_neg)
  int need_fix;                                // 0: speculative outputs are final; 1: sparse fix-up list; 2: dense re-merge
  unsigned int fix_count;                      // entries appended to the fix-up list (may exceed its capacity)
  // sampled bracket of the k-th magnitude (16-bit dtypes): keys in [win_lo, win_hi] are histogrammed, keys below counted
  unsigned int win_lo[MC_MERGE_MAX_SRC], win_hi[MC_MERGE_MAX_SRC];
  unsigned long long below[MC_MERGE_MAX_SRC];
  int need_full;                               // 1: run the full-range histogram passes (fp32, small inputs, bracket miss)
  // CTAs of the current launch that have flushed their counts for a source: the last one runs the per-source step that follows
  // (bracket / window select / radix select) in the same launch, and puts the counter back to 0
  unsigned int done[MC_MERGE_MAX_SRC];
};

constexpr int kTiesChunkBytes = 16384;  // one chunk = 1024 16-byte vectors of every source
constexpr int kTiesMergeThreads = 256;    // x 4 vectors per source per thread = one chunk; every load issued up front
constexpr int kTiesMetricsThreads = 512;
constexpr int kTiesHistThreads = 1024;
constexpr int kTiesCountThreads = 256;
constexpr int kTiesWindowBins = 2048;   // widest bracket the counting pass histograms (8 KB of shared memory)
constexpr int kTiesSampleEvery = 32;    // the sampling pass reads one 512-byte granule (1/32) of every chunk

// Elements whose surviving entries cancel exactly depend on the global majority sign, which is only known after the whole
// census: the speculative pass appends them here (packed chunk index << 32 | offset inside the chunk) and the fix-up kernel
// recomputes just those.  A full list (or MAX with a negative majority, which turns every empty element into -0) falls back
// to a dense re-merge.
constexpr unsigned int kTiesFixCapacity = 1u << 20;

// One output element.  `cls` receives 0 / 1 (elected sign + / -), 2 (no source survives the trim) or 3 (survivors
// cancel exactly: the output depends on the global majority sign).  Rounding points follow the reference's torch ops:
//   ties_merging.py:98-101  m = x * (|x| >= thr)
//   :121-124, :111-118      s = sign(round_dt(sum_src m)), zeros take the majority sign
//   :133-137                keep m where its sign agrees with s (s > 0 ? m > 0 : m < 0)
//   :144-153                sum: round_dt(sum kept) | mean: fp32(round_dt(sum kept)) / max(#kept != 0, 1) | max: round_dt(max |kept|) * s
// The kept entries all share one sign, so "sum kept" is the left-to-right fp32 sum of the positive (or of the negative)
// survivors with +0 in the other slots — both candidates are accumulated in the same sweep as the sign election and the
// elected one is picked afterwards.  Zero results are +0 (torch's reductions start from +0) except MAX, whose `* s` keeps -0.
// element e of a 16-byte vector of S as float32 (bf16: one shift / mask on the packed word, no byte permute)
template <typename S>
__device__ __forceinline__ float vec_elem_f32(const Vec<16>& v, int e) {
  if constexpr (std::is_same<S, __nv_bfloat16>::value) {
    const uint32_t w = v.w[e >> 1];
    return __uint_as_float((e & 1) ? (w & 0xffff0000u) : (w << 16));
  } else {
    return to_f32<S>(reinterpret_cast<const S*>(&v)[e]);
  }
}
template <int NSRC, typename S, typename D, int FUNC>
__device__ __forceinline__ D ties_one_ref(const float (&in)[NSRC], const float (&thr)[NSRC], float majority, int& cls) {
  float acc = 0.0f, pos = 0.0f, neg = 0.0f;  // pos / neg double as the running max / min for MAX
  int n_pos = 0, n_neg = 0;
#pragma unroll
  for (int s = 0; s < NSRC; ++s) {
    const float x = in[s];
    const float m = fabsf(x) >= thr[s] ? x : 0.0f;
    acc = __fadd_rn(acc, m);
    if (FUNC == MC_TIES_MAX) {
      pos = fmaxf(pos, m);
      neg = fminf(neg, m);
    } else {
      pos = __fadd_rn(pos, fmaxf(m, 0.0f));
      neg = __fadd_rn(neg, fminf(m, 0.0f));
    }
    if (FUNC == MC_TIES_MEAN) {
      n_pos += m > 0.0f ? 1 : 0;
      n_neg += m < 0.0f ? 1 : 0;
    }
  }
  const float total = to_f32<S>(from_f32<S>(acc));
  const bool any_nz = pos > 0.0f || neg < 0.0f;
  float sg = total > 0.0f ? 1.0f : (total < 0.0f ? -1.0f : 0.0f);
  cls = total > 0.0f ? 0 : (total < 0.0f ? 1 : (any_nz ? 3 : 2));
  if (sg == 0.0f) sg = majority;
  const bool up = sg > 0.0f;
  if (FUNC == MC_TIES_SUM) return from_f32<D>(up ? pos : neg);
  if (FUNC == MC_TIES_MEAN) {
    const int cnt = up ? n_pos : n_neg;
    return from_f32<D>(__fdiv_rn(to_f32<S>(from_f32<S>(up ? pos : neg)), (float)(cnt > 1 ? cnt : 1)));
  }
  return from_f32<D>(__fmul_rn(to_f32<S>(from_f32<S>(up ? pos : fabsf(neg))), sg));  // |.| first: (+0) * -1 = -0 as torch
}

// a * b rounded to nearest, then clamped to [0, 1] with NaN -> +0 (PTX .sat): mul_sat(k, +inf) is the indicator of k > 0
// for k >= +0 on the FMA pipe (0 * inf = NaN -> 0; any positive k, subnormals included, -> inf -> 1).
__device__ __forceinline__ float mul_sat(float a, float b) {
  float r;
  asm("mul.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float rcp_approx(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

// Packed trim of one 32-bit word (two 16-bit values): entries with |x| < thr become +0, the others keep their bits.
// thr2 holds the threshold (a value of the dtype: it is the k-th |x| of the data) in both halves.  Two instructions per
// word: HSET2.LT with the |.| operand modifier, then w & ~mask.  ptxas only folds the modifier into the fp16 form of HSET2
// (for bf16 it materialises |w| with an extra instruction), so bf16 pairs are compared AS fp16 bit patterns: both formats
// are sign-magnitude with the exponent above the mantissa, hence the magnitudes order like their bit patterns in either.
// The exception is a pattern whose magnitude bits are >= 0x7c01, a NaN to the fp16 compare, which answers false: the
// value is kept — right whenever the threshold itself is below 0x7c00 (bf16 2^121), which ties_fast_threshold checks
// (the kernel takes the scalar path of the unaligned tails otherwise).
template <typename S>
__device__ __forceinline__ uint32_t trim_word(uint32_t w, uint32_t thr2) {
  __half2_raw ra, rt;
  ra.x = (unsigned short)w; ra.y = (unsigned short)(w >> 16);
  rt.x = (unsigned short)thr2; rt.y = (unsigned short)(thr2 >> 16);
  return w & ~__hlt2_mask(__habs2(__half2(ra)), __half2(rt));
}
template <typename S>
__device__ __forceinline__ bool ties_fast_threshold(uint32_t thr2) {
  return !std::is_same<S, __nv_bfloat16>::value || (thr2 & 0xffffu) < 0x7c00u;
}
template <typename S>
__device__ __forceinline__ uint32_t pack_thr(float thr) {
  uint32_t h;  // the conversion is exact
  if constexpr (std::is_same<S, __nv_bfloat16>::value) h = __bfloat16_as_ushort(__float2bfloat16_rn(thr));
  else h = __half_as_ushort(__float2half_rn(thr));
  return h | (h << 16);
}

// 16-bit sources: the same results bit for bit as ties_one_ref with the arithmetic moved off the
// half-rate ALU pipe that bounded the first version of this pass (profiles/r01_ties.txt: 57 instructions per element,
// ALU 65 %, DRAM 39 %) — no predicate, select or min/max per source, everything but the trim compare is FMUL/FADD/FFMA:
//   * trim by multiplication, m = x * [|x| >= thr] (the reference's own form, ties_merging.py:98-101) in the scalar form; the vector
//     path trims the packed words before they are widened (trim_word);
//   * the sign is elected from the fp32 sum itself: rounding it to the 16-bit dtype first (:121) never changes the sign
//     and never turns a non-zero sum into zero (the sum is a multiple of the dtype's smallest subnormal, overflow keeps
//     the sign).  p = [acc > 0] and n = [acc < 0] are mul_sat(+-acc, inf), and half the elected sign is
//     hs = mh + p (0.5 - mh) - n (0.5 + mh) with mh = +-0.5 the majority default: exact, values in {-0.5, +0.5};
//   * only the elected candidate is summed, in a second sweep over the registers: k = max(sigma m, 0) computed as
//     0.5 |m| + hs m, exact for 16-bit values (halving never drops a bit in fp32), so sum k is the reference's
//     left-to-right sum of the kept entries with +0 in the other slots, negated for sigma = -1 (round-to-nearest is
//     symmetric); fma(sum k, sigma, +0) restores the reference's +0 when nothing is kept;
//   * MEAN: #kept != 0 is sum of mul_sat(k, inf); the float32 division by that small integer c is
//     q0 = x r, q1 = fma(fma(-q0, c, x), r, q0) with r ~ 1 / c (MUFU), the correctly rounded quotient for every 16-bit x,
//     c <= 8 and any r within 2 ulp of 1 / c (checked exhaustively: tests/test_ties_fastmath.py); min(q1, x) returns
//     x = inf when the 16-bit rounding of the sum overflowed (q1 = NaN there);
//   * survivors that cancel exactly (class 3) are acc == 0 with a non-empty candidate, for either default sign:
//     amb = [sum k > 0] (1 - p - n);
//   * MAX keeps the running maximum of k instead of the sum; (max k) * sigma is exact and keeps torch's -0.
// p, n, some (and amb in the scalar form) come back as 0.0 / 1.0, so the census of the vector path is three FFMAs per element
// (ties_chunk_fast).

// Election half of the element function on already trimmed entries m: the kept sum (or running max) ksum >= +0, the number of
// kept non-zero entries cnt (MEAN only), the elected sign sg = +-1, the census indicators p = [acc > 0], n = [acc < 0] and
// some = [a kept entry is non-zero].  DEFPOS: the default sign is known to be + (the speculative pass), so hs = 0.5 - n.
template <int NSRC, int FUNC, bool DEFPOS>
__device__ __forceinline__ void ties_elect_fast(const float (&m)[NSRC], float mh, float& ksum, float& cnt, float& sg, float& p, float& n,
                                                float& some) {
  const float inf = __int_as_float(0x7f800000);
  float acc = m[0];
#pragma unroll
  for (int s = 1; s < NSRC; ++s) acc = __fadd_rn(acc, m[s]);
  p = mul_sat(acc, inf);
  n = mul_sat(acc, -inf);
  const float hs = DEFPOS ? __fsub_rn(0.5f, n) : __fmaf_rn(-n, __fadd_rn(0.5f, mh), __fmaf_rn(p, __fsub_rn(0.5f, mh), mh));
  sg = __fadd_rn(hs, hs);
  ksum = 0.0f;
  cnt = 0.0f;
#pragma unroll
  for (int s = 0; s < NSRC; ++s) {
    const float k = __fmaf_rn(0.5f, fabsf(m[s]), __fmul_rn(m[s], hs));
    ksum = s == 0 ? k : (FUNC == MC_TIES_MAX ? fmaxf(ksum, k) : __fadd_rn(ksum, k));  // MAX: running max |kept|
    if (FUNC == MC_TIES_MEAN) {
      const float one = mul_sat(k, inf);
      cnt = s == 0 ? one : __fadd_rn(cnt, one);
    }
  }
  some = FUNC == MC_TIES_MEAN ? mul_sat(cnt, 1.0f) : mul_sat(ksum, inf);
}

// MEAN: x = the kept sum after its 16-bit rounding (>= +0), c = #kept.  c = 0 comes with x = 0: r = inf, q0 = q1 = NaN and
// min(NaN, 0) = 0, so no max(c, 1) is needed; fma(., sg, +0) puts the sign on and turns -0 into the reference's +0.
template <typename D>
__device__ __forceinline__ D ties_mean_quotient(float x, float c, float sg) {
  const float r = rcp_approx(c);
  const float q0 = __fmul_rn(x, r);
  const float q1 = __fmaf_rn(__fmaf_rn(-q0, c, x), r, q0);
  return from_f32<D>(__fmaf_rn(fminf(q1, x), sg, 0.0f));
}

// float32 values of a and b after rounding to the 16-bit dtype S: one packed conversion for the pair
template <typename S>
__device__ __forceinline__ void round_pair(float a, float b, float& xa, float& xb) {
  if constexpr (std::is_same<S, __nv_bfloat16>::value) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&h);
    xa = __uint_as_float(u << 16);
    xb = __uint_as_float(u & 0xffff0000u);
  } else {
    const float2 f = __half22float2(__floats2half2_rn(a, b));
    xa = f.x;
    xb = f.y;
  }
}

template <int NSRC, typename S, typename D, int FUNC>
__device__ __forceinline__ D ties_one_fast(const float (&in)[NSRC], const float (&thr)[NSRC], float mh, float& p, float& n, float& amb) {
  static_assert(sizeof(S) == 2, "16-bit sources");
  float m[NSRC];
#pragma unroll
  for (int s = 0; s < NSRC; ++s) m[s] = __fmul_rn(in[s], fabsf(in[s]) >= thr[s] ? 1.0f : 0.0f);
  float ksum, cnt, sg, some;
  ties_elect_fast<NSRC, FUNC, false>(m, mh, ksum, cnt, sg, p, n, some);
  amb = __fmaf_rn(-some, __fadd_rn(p, n), some);
  if (FUNC == MC_TIES_SUM) return from_f32<D>(__fmaf_rn(ksum, sg, 0.0f));
  if (FUNC == MC_TIES_MAX) return from_f32<D>(__fmul_rn(ksum, sg));  // a 16-bit value already; (+0) * -1 = -0 as torch
  return ties_mean_quotient<D>(to_f32<S>(from_f32<S>(ksum)), cnt, sg);
}

template <typename S, int FUNC>
struct TiesFast {
  static constexpr bool value = sizeof(S) == 2;
};

template <int NSRC, typename S, typename D, int FUNC>
__device__ __forceinline__ D ties_one(const float (&in)[NSRC], const float (&thr)[NSRC], float majority, int& cls) {
  if constexpr (TiesFast<S, FUNC>::value) {
    float p, n, amb;
    const D out = ties_one_fast<NSRC, S, D, FUNC>(in, thr, majority > 0.0f ? 0.5f : -0.5f, p, n, amb);
    cls = p != 0.0f ? 0 : (n != 0.0f ? 1 : (amb != 0.0f ? 3 : 2));
    // majority sign exactly 0 (n_pos == n_neg): the reference multiplies max |kept| by sign 0 (ties_merging.py:152-153), so
    // every element without an elected sign is +0 for MAX; SUM / MEAN never multiply by the sign and keep the negative side
    if (FUNC == MC_TIES_MAX && majority == 0.0f && cls >= 2) return from_f32<D>(0.0f);
    return out;
  } else {
    return ties_one_ref<NSRC, S, D, FUNC>(in, thr, majority, cls);
  }
}

// Ask L2 for `bytes` (multiple of 16) starting at the 16-byte aligned global address p: one instruction, no destination.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned int bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One whole, aligned chunk of 16-bit sources on the fast element function: 4 vectors per source per thread, trimmed two
// values per instruction on the packed words before they are widened, elements taken in pairs so that MEAN's 16-bit
// rounding of the kept sum is one packed conversion.  ONE basic block (ptxas sinks loads whose first use sits in a later
// block).  The census (DEFPOS, the speculative pass, only) is kept as one BIT per element in float accumulators — three FFMAs
// per element, mask += indicator * 2^bit, 16 elements per mask so the sums stay exact — for p = [sum > 0], n = [sum < 0] and
// some = [a kept entry is non-zero]; per 16 elements the counts are two popcounts and the class-3 elements (survivors cancel
// exactly: a kept entry and no elected sign) are the bits of some & ~(p | n), returned in amb_bits for the caller's rare path.
template <int NSRC, typename S, typename D, int FUNC, bool DEFPOS>
__device__ __forceinline__ void ties_chunk_fast(const void* const* __restrict__ row, const uint32_t (&thr2)[NSRC], float mh, float majority,
                                                unsigned int& c_pos, unsigned int& c_neg, unsigned int (&amb_bits)[2]) {
  constexpr int E = 16 / sizeof(S);
  constexpr int VPT = 4;
  static_assert(E == 8 && VPT == 4, "two masks of 16 bits");
  using VS = Vec<16>;
  using VD = Vec<E * sizeof(D)>;
  VS v[NSRC][VPT];
#pragma unroll
  for (int s = 0; s < NSRC; ++s) {
    const VS* p = reinterpret_cast<const VS*>(row[s]) + threadIdx.x;
#pragma unroll
    for (int j = 0; j < VPT; ++j) v[s][j] = ld_stream(p + j * kTiesMergeThreads);
  }
  VD* q = reinterpret_cast<VD*>(const_cast<void*>(row[NSRC])) + threadIdx.x;
  float m_pos = 0.0f, m_neg = 0.0f, m_some = 0.0f;
#pragma unroll
  for (int j = 0; j < VPT; ++j) {
    VD o;
    D* oe = reinterpret_cast<D*>(&o);
#pragma unroll
    for (int s = 0; s < NSRC; ++s) {
#pragma unroll
      for (int w = 0; w < 4; ++w) v[s][j].w[w] = trim_word<S>(v[s][j].w[w], thr2[s]);
    }
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      float ksum[2], cnt[2], sg[2], pn[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float in[NSRC];
#pragma unroll
        for (int s = 0; s < NSRC; ++s) in[s] = vec_elem_f32<S>(v[s][j], e + h);
        float p, n, some;
        ties_elect_fast<NSRC, FUNC, DEFPOS>(in, mh, ksum[h], cnt[h], sg[h], p, n, some);
        if (DEFPOS) {
          const float bit = (float)(1 << ((j & 1) * E + e + h));
          m_pos = __fmaf_rn(p, bit, m_pos);
          m_neg = __fmaf_rn(n, bit, m_neg);
          m_some = __fmaf_rn(some, bit, m_some);
        }
        pn[h] = FUNC == MC_TIES_MAX && !DEFPOS ? p + n : 1.0f;
      }
      if (FUNC == MC_TIES_MEAN) {
        float x[2];
        round_pair<S>(ksum[0], ksum[1], x[0], x[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) oe[e + h] = ties_mean_quotient<D>(x[h], cnt[h], sg[h]);
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          oe[e + h] = FUNC == MC_TIES_SUM ? from_f32<D>(__fmaf_rn(ksum[h], sg[h], 0.0f)) : from_f32<D>(__fmul_rn(ksum[h], sg[h]));
          // majority sign exactly 0 (see ties_one): MAX multiplies by sign 0 wherever no sign is elected
          if (FUNC == MC_TIES_MAX && !DEFPOS && majority == 0.0f && pn[h] == 0.0f) oe[e + h] = from_f32<D>(0.0f);
        }
      }
    }
    st_stream(q + j * kTiesMergeThreads, o);
    if (DEFPOS && (j & 1)) {
      const unsigned int bp = (unsigned int)m_pos, bn = (unsigned int)m_neg;
      c_pos += __popc(bp);
      c_neg += __popc(bn);
      amb_bits[j >> 1] = (unsigned int)m_some & ~(bp | bn);
      m_pos = m_neg = m_some = 0.0f;
    }
  }
}

// mode 0: speculative merge with majority = +1, census of the elected signs, list of majority-dependent elements.
// mode 1: dense re-merge with the real majority; exits at once unless the fix-up kernel asked for it (need_fix == 2).
template <int NSRC, typename S, typename D, int FUNC>
__global__ void __launch_bounds__(kTiesMergeThreads)
ties_merge_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, const void* const* __restrict__ vec,
                  int nchunks, TiesState* st, unsigned long long* __restrict__ fix_list, int mode, int pf_dist) {
  constexpr int E = 16 / sizeof(S);
  constexpr int VPT = 4;  // vectors per source per thread
  constexpr int CHUNK = kTiesChunkBytes / sizeof(S);
  static_assert(CHUNK == kTiesMergeThreads * VPT * E, "chunk geometry");
  using VS = Vec<16>;
  using VD = Vec<E * sizeof(D)>;
  float majority = 1.0f;
  if (mode == 1) {
    if (st->need_fix != 2) return;
    majority = (float)st->majority;
  }
  float thr[NSRC];
#pragma unroll
  for (int s = 0; s < NSRC; ++s) thr[s] = st->thr[s];
  constexpr bool kFast = TiesFast<S, FUNC>::value;
  const float mh = majority > 0.0f ? 0.5f : -0.5f;
  uint32_t thr2[NSRC];
  bool fast_ok = kFast;
  if constexpr (kFast) {
#pragma unroll
    for (int s = 0; s < NSRC; ++s) {
      thr2[s] = pack_thr<S>(thr[s]);
      fast_ok = fast_ok && ties_fast_threshold<S>(thr2[s]);
    }
  }
  unsigned int c_pos = 0u, c_neg = 0u, c_amb = 0u;  // elements without survivors are derived: total - pos - neg - amb
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    // ptxas sinks this long block's loads next to their first use (three 16-byte loads in flight per thread), so the
    // memory-level parallelism comes from L2 prefetches instead: one thread asks L2 for the whole chunk that the CTA
    // scheduled pf_dist blocks later will stream (one generation of resident CTAs ahead; the first generation asks
    // for its own), and the demand loads below mostly hit L2.
    // addresses of a whole, aligned chunk come from the plan's flat table (row = sources..., destination; NULL = tail)
    const void* const* row = vec + (long long)c * (NSRC + 1);
    if (pf_dist > 0 && threadIdx.x == 0) {
#pragma unroll 1
      for (int rep = 0; rep < 2; ++rep) {
        const long long cq = rep == 0 ? (long long)c + pf_dist : (long long)c;
        if (cq >= nchunks || (rep == 1 && c >= pf_dist)) continue;
        const void* const* prow = vec + cq * (NSRC + 1);
        if (prow[0] != nullptr) {
#pragma unroll
          for (int s = 0; s < NSRC; ++s) l2_prefetch_bulk(prow[s], (unsigned int)(CHUNK * sizeof(S)));
        }
      }
    }
    bool vector_path = row[0] != nullptr;
    if constexpr (kFast) {
      vector_path = vector_path && fast_ok;
      if (vector_path) {
        unsigned int amb_bits[2] = {0u, 0u};
        if (mode == 0) {
          ties_chunk_fast<NSRC, S, D, FUNC, true>(row, thr2, mh, majority, c_pos, c_neg, amb_bits);
          if ((amb_bits[0] | amb_bits[1]) != 0u) {  // rare: append this thread's class-3 elements to the fix-up list
#pragma unroll 1
            for (int jp = 0; jp < 2; ++jp) {
              unsigned int bits = jp == 0 ? amb_bits[0] : amb_bits[1];
              while (bits) {
                const int b = __ffs((int)bits) - 1;
                bits &= bits - 1u;
                c_amb += 1u;
                const unsigned int slot = atomicAdd(&st->fix_count, 1u);
                if (slot < kTiesFixCapacity)
                  fix_list[slot] = ((unsigned long long)c << 32) | (unsigned int)(((jp * 2 + (b >> 3)) * kTiesMergeThreads + threadIdx.x) * E + (b & 7));
              }
            }
          }
        } else {
          ties_chunk_fast<NSRC, S, D, FUNC, false>(row, thr2, mh, majority, c_pos, c_neg, amb_bits);
        }
      }
    } else if (vector_path) {
      VS v[NSRC][VPT];
#pragma unroll
      for (int s = 0; s < NSRC; ++s) {
        const VS* p = reinterpret_cast<const VS*>(row[s]) + threadIdx.x;
#pragma unroll
        for (int j = 0; j < VPT; ++j) v[s][j] = ld_stream(p + j * kTiesMergeThreads);
      }
      VD* q = reinterpret_cast<VD*>(const_cast<void*>(row[NSRC])) + threadIdx.x;
      // Class-3 elements (rare) are collected as one bit per element and appended to the fix-up list after the sweep
      unsigned int amb_mask[VPT];
      unsigned int amb_any = 0u;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        VD o;
        D* oe = reinterpret_cast<D*>(&o);
        amb_mask[j] = 0u;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          float in[NSRC];
#pragma unroll
          for (int s = 0; s < NSRC; ++s) in[s] = vec_elem_f32<S>(v[s][j], e);
          int cls;
          oe[e] = ties_one<NSRC, S, D, FUNC>(in, thr, majority, cls);
          c_pos += cls == 0;
          c_neg += cls == 1;
          amb_mask[j] |= cls == 3 ? 1u << e : 0u;
        }
        amb_any |= amb_mask[j];
        st_stream(q + j * kTiesMergeThreads, o);
      }
      if (amb_any != 0u) {
#pragma unroll 1
        for (int j = 0; j < VPT; ++j) {
          unsigned int bits = amb_mask[j];
          while (bits) {
            const int e = __ffs((int)bits) - 1;
            bits &= bits - 1u;
            c_amb += 1u;
            if (mode == 0) {
              const unsigned int slot = atomicAdd(&st->fix_count, 1u);
              if (slot < kTiesFixCapacity)
                fix_list[slot] = ((unsigned long long)c << 32) | (unsigned int)((j * kTiesMergeThreads + threadIdx.x) * E + e);
            }
          }
        }
      }
    }
    if (!vector_path) {
      const MergeChunk ch = chunks[c];
      const MergeSeg* sg = segs + ch.seg;
      const long long base = (long long)ch.idx * CHUNK;
      const long long rem = sg->numel - base;
      const long long n = rem < CHUNK ? rem : CHUNK;
      for (long long i = threadIdx.x; i < n; i += kTiesMergeThreads) {
        float in[NSRC];
#pragma unroll
        for (int s = 0; s < NSRC; ++s) in[s] = to_f32<S>(reinterpret_cast<const S*>(sg->src[s])[base + i]);
        int cls;
        reinterpret_cast<D*>(sg->dst)[base + i] = ties_one<NSRC, S, D, FUNC>(in, thr, majority, cls);
        c_pos += cls == 0;
        c_neg += cls == 1;
        if (cls == 3) {
          c_amb += 1u;
          if (mode == 0) {
            const unsigned int slot = atomicAdd(&st->fix_count, 1u);
            if (slot < kTiesFixCapacity) fix_list[slot] = ((unsigned long long)c << 32) | (unsigned int)i;
          }
        }
      }
    }
  }
  if (mode == 1) return;
  __shared__ unsigned int s_census[3];
  if (threadIdx.x < 3) s_census[threadIdx.x] = 0u;
  __syncthreads();
  unsigned int census[3] = {c_pos, c_neg, c_amb};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    unsigned int x = census[k];
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(&s_census[k], x);
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_census[threadIdx.x]) {
    unsigned long long* dst = threadIdx.x == 0 ? &st->n_pos : threadIdx.x == 1 ? &st->n_neg : &st->n_amb;
    atomicAdd(dst, (unsigned long long)s_census[threadIdx.x]);
  }
}

// After the speculative pass: the majority sign of the census, the decision between nothing / sparse fix-up / dense re-merge
// (every CTA derives it from the same counters; CTA 0 records it for the re-merge launch and the statistics), and the sparse
// fix-up itself: the listed majority-dependent elements recomputed with the real majority (need_fix == 1).
// The speculative pass used +1: wrong wherever survivors cancel exactly (listed), and MAX writes -0 (0 * -1) into every
// element without survivors when the majority is negative (dense re-merge).
template <int NSRC, typename S, typename D, int FUNC>
__global__ void __launch_bounds__(256)
ties_fix_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, TiesState* st,
                const unsigned long long* __restrict__ fix_list, unsigned long long total) {
  const unsigned long long n_p = st->n_pos, n_n = st->n_neg, n_a = st->n_amb;
  const unsigned int n = st->fix_count;
  const unsigned long long n_zero = total - n_p - n_n - n_a;
  const int maj = n_p > n_n ? 1 : (n_p < n_n ? -1 : 0);
  int fix = 0;
  if (maj != 1) {
    if (FUNC == MC_TIES_MAX && maj == -1 && n_zero > 0ull) fix = 2;
    else if (n_a > 0ull) fix = n <= kTiesFixCapacity ? 1 : 2;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->n_zero = n_zero;
    st->majority = maj;
    st->need_fix = fix;
  }
  if (fix != 1) return;
  constexpr int CHUNK = kTiesChunkBytes / sizeof(S);
  const float majority = (float)maj;
  float thr[NSRC];
#pragma unroll
  for (int s = 0; s < NSRC; ++s) thr[s] = st->thr[s];
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long packed = fix_list[i];
    const MergeChunk ch = chunks[(int)(packed >> 32)];
    const MergeSeg* sg = segs + ch.seg;
    const long long idx = (long long)ch.idx * CHUNK + (long long)(unsigned int)packed;
    float in[NSRC];
#pragma unroll
    for (int s = 0; s < NSRC; ++s) in[s] = to_f32<S>(reinterpret_cast<const S*>(sg->src[s])[idx]);
    int cls;
    reinterpret_cast<D*>(sg->dst)[idx] = ties_one<NSRC, S, D, FUNC>(in, thr, majority, cls);
  }
}

typedef void (*ties_fn_t)(const MergeSeg*, const MergeChunk*, const void* const*, int, TiesState*, unsigned long long*, int, int);
typedef void (*ties_fix_fn_t)(const MergeSeg*, const MergeChunk*, TiesState*, const unsigned long long*, unsigned long long);
struct TiesKernels {
  ties_fn_t merge;
  ties_fix_fn_t fix;
};

template <typename S, int FUNC>
static TiesKernels ties_pick_nsrc(int n_src) {
  using D = typename std::conditional<FUNC == MC_TIES_MEAN, float, S>::type;  // mean promotes to float32 (see ties_one)
  switch (n_src) {
#define MC_TIES_CASE(N) \
  case N: return TiesKernels{ties_merge_kernel<N, S, D, FUNC>, ties_fix_kernel<N, S, D, FUNC>};
    MC_TIES_CASE(1)
    MC_TIES_CASE(2)
    MC_TIES_CASE(3)
    MC_TIES_CASE(4)
    MC_TIES_CASE(5)
    MC_TIES_CASE(6)
    MC_TIES_CASE(7)
    MC_TIES_CASE(8)
#undef MC_TIES_CASE
  }
  return TiesKernels{nullptr, nullptr};
}

template <typename S>
static TiesKernels ties_pick_func(int n_src, int func) {
  switch (func) {
    case MC_TIES_SUM: return ties_pick_nsrc<S, MC_TIES_SUM>(n_src);
    case MC_TIES_MEAN: return ties_pick_nsrc<S, MC_TIES_MEAN>(n_src);
    case MC_TIES_MAX: return ties_pick_nsrc<S, MC_TIES_MAX>(n_src);
  }
  return TiesKernels{nullptr, nullptr};
}

// one translation unit per source dtype (mc_ties_inst.cu, -DMC_TIES_DT=k)
TiesKernels pick_ties_bf16(int n_src, int func);
TiesKernels pick_ties_f16(int n_src, int func);
TiesKernels pick_ties_f32(int n_src, int func);

// ---- interference metrics (reference calculate_metrics.py:26-37,53-64) ---------------------------------------------
// L2 and cosine distance between the first two sources, soft sign dissimilarity over all sources before and after the
// top-k trim.  Per-element arithmetic is fp32 exactly as the reference's torch ops (after its .float() on load); the
// reductions over elements accumulate in fp64 (the reference's fp32 tree sums agree to ~1e-6 relative).
struct TiesMetricSums {
  double d2, xy, xx, yy, ssd, tssd;
  unsigned long long ssd_n, tssd_n;
};

template <int NSRC, typename S>
__device__ __forceinline__ void metrics_one(const S (&in)[NSRC], const float (&thr)[NSRC], float (&acc)[6], unsigned int (&cnt)[2]) {
  float sum = 0.0f, asum = 0.0f, tsum = 0.0f, tasum = 0.0f, x0 = 0.0f, x1 = 0.0f;
#pragma unroll
  for (int s = 0; s < NSRC; ++s) {
    const float x = to_f32<S>(in[s]);
    const float m = fabsf(x) >= thr[s] ? x : 0.0f;
    sum = __fadd_rn(sum, x);
    asum = __fadd_rn(asum, fabsf(x));
    tsum = __fadd_rn(tsum, m);
    tasum = __fadd_rn(tasum, fabsf(m));
    if (s == 0) x0 = x;
    if (s == 1) x1 = x;
  }
  if (NSRC >= 2) {
    const float d = __fsub_rn(x0, x1);
    acc[0] += __fmul_rn(d, d);
    acc[1] += __fmul_rn(x0, x1);
    acc[2] += __fmul_rn(x0, x0);
    acc[3] += __fmul_rn(x1, x1);
  }
  if (asum != 0.0f) {
    acc[4] += fabsf(__fdiv_rn(sum, asum));
    cnt[0] += 1u;
  }
  if (tasum != 0.0f) {
    acc[5] += fabsf(__fdiv_rn(tsum, tasum));
    cnt[1] += 1u;
  }
}

template <int NSRC, typename S>
__global__ void __launch_bounds__(kTiesMetricsThreads)
ties_metrics_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, int nchunks, const TiesState* __restrict__ st,
                    TiesMetricSums* __restrict__ out) {
  constexpr int E = 16 / sizeof(S);
  constexpr int VPT = 2;
  constexpr int CHUNK = kTiesChunkBytes / sizeof(S);
  float thr[NSRC];
#pragma unroll
  for (int s = 0; s < NSRC; ++s) thr[s] = st->thr[s];
  double tot[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  unsigned long long n[2] = {0ull, 0ull};
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const MergeChunk ch = chunks[c];
    const MergeSeg* sg = segs + ch.seg;
    const long long base = (long long)ch.idx * CHUNK;
    const long long rem = sg->numel - base;
    float acc[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};  // fp32 partials over this thread's <= 16 elements of the chunk
    unsigned int cnt[2] = {0u, 0u};
    if (sg->aligned && rem >= CHUNK) {
      Vec<16> v[NSRC][VPT];
#pragma unroll
      for (int s = 0; s < NSRC; ++s) {
        const Vec<16>* p = reinterpret_cast<const Vec<16>*>(reinterpret_cast<const S*>(sg->src[s]) + base) + threadIdx.x;
#pragma unroll
        for (int j = 0; j < VPT; ++j) v[s][j] = ld_stream(p + j * kTiesMetricsThreads);
      }
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          S in[NSRC];
#pragma unroll
          for (int s = 0; s < NSRC; ++s) in[s] = reinterpret_cast<const S*>(&v[s][j])[e];
          metrics_one<NSRC, S>(in, thr, acc, cnt);
        }
      }
    } else {
      const long long m = rem < CHUNK ? rem : CHUNK;
      for (long long i = threadIdx.x; i < m; i += kTiesMetricsThreads) {
        S in[NSRC];
#pragma unroll
        for (int s = 0; s < NSRC; ++s) in[s] = reinterpret_cast<const S*>(sg->src[s])[base + i];
        metrics_one<NSRC, S>(in, thr, acc, cnt);
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) tot[k] += (double)acc[k];
    n[0] += cnt[0];
    n[1] += cnt[1];
  }
  __shared__ double s_tot[6];
  __shared__ unsigned long long s_n[2];
  if (threadIdx.x < 6) s_tot[threadIdx.x] = 0.0;
  if (threadIdx.x < 2) s_n[threadIdx.x] = 0ull;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double x = tot[k];
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_tot[k], x);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    unsigned long long x = n[k];
#pragma unroll
    for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_n[k], x);
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double* dst = threadIdx.x == 0 ? &out->d2 : threadIdx.x == 1 ? &out->xy : threadIdx.x == 2 ? &out->xx : threadIdx.x == 3 ? &out->yy
                  : threadIdx.x == 4 ? &out->ssd : &out->tssd;
    atomicAdd(dst, s_tot[threadIdx.x]);
  }
  if (threadIdx.x >= 32 && threadIdx.x < 34) atomicAdd(threadIdx.x == 32 ? &out->ssd_n : &out->tssd_n, s_n[threadIdx.x - 32]);
}

typedef void (*ties_metrics_fn_t)(const MergeSeg*, const MergeChunk*, int, const TiesState*, TiesMetricSums*);

template <typename S>
static ties_metrics_fn_t ties_pick_metrics(int n_src) {
  switch (n_src) {
    case 2: return ties_metrics_kernel<2, S>;
    case 3: return ties_metrics_kernel<3, S>;
    case 4: return ties_metrics_kernel<4, S>;
    case 5: return ties_metrics_kernel<5, S>;
    case 6: return ties_metrics_kernel<6, S>;
    case 7: return ties_metrics_kernel<7, S>;
    case 8: return ties_metrics_kernel<8, S>;
  }
  return nullptr;
}
ties_metrics_fn_t pick_ties_metrics_bf16(int n_src);
ties_metrics_fn_t pick_ties_metrics_f16(int n_src);
ties_metrics_fn_t pick_ties_metrics_f32(int n_src);

}  // namespace mc
