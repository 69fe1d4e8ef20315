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
};

constexpr int kTiesChunkBytes = 16384;  // one chunk = 1024 16-byte vectors of every source
constexpr int kTiesMergeThreads = 512;
constexpr int kTiesHistThreads = 1024;
constexpr int kTiesCountThreads = 512;
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
template <int NSRC, typename S, typename D, int FUNC>
__device__ __forceinline__ D ties_one(const S (&in)[NSRC], const float (&thr)[NSRC], float majority, int& cls) {
  float acc = 0.0f, pos = 0.0f, neg = 0.0f;  // pos / neg double as the running max / min for MAX
  int n_pos = 0, n_neg = 0;
#pragma unroll
  for (int s = 0; s < NSRC; ++s) {
    const float x = to_f32<S>(in[s]);
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

// mode 0: speculative merge with majority = +1, census of the elected signs, list of majority-dependent elements.
// mode 1: dense re-merge with the real majority; exits at once unless ties_finalize_kernel asked for it (need_fix == 2).
template <int NSRC, typename S, typename D, int FUNC>
__global__ void __launch_bounds__(kTiesMergeThreads)
ties_merge_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, int nchunks, TiesState* st,
                  unsigned long long* __restrict__ fix_list, int mode) {
  constexpr int E = 16 / sizeof(S);
  constexpr int VPT = 2;  // vectors per source per thread
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
  unsigned int c_pos = 0u, c_neg = 0u, c_amb = 0u;  // elements without survivors are derived: total - pos - neg - amb
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const MergeChunk ch = chunks[c];
    const MergeSeg* sg = segs + ch.seg;
    const long long base = (long long)ch.idx * CHUNK;
    const long long rem = sg->numel - base;
    if (sg->aligned && rem >= CHUNK) {
      VS v[NSRC][VPT];
#pragma unroll
      for (int s = 0; s < NSRC; ++s) {
        const VS* p = reinterpret_cast<const VS*>(reinterpret_cast<const S*>(sg->src[s]) + base) + threadIdx.x;
#pragma unroll
        for (int j = 0; j < VPT; ++j) v[s][j] = ld_stream(p + j * kTiesMergeThreads);
      }
      VD* q = reinterpret_cast<VD*>(reinterpret_cast<D*>(sg->dst) + base) + threadIdx.x;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        VD o;
        D* oe = reinterpret_cast<D*>(&o);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          S in[NSRC];
#pragma unroll
          for (int s = 0; s < NSRC; ++s) in[s] = reinterpret_cast<const S*>(&v[s][j])[e];
          int cls;
          oe[e] = ties_one<NSRC, S, D, FUNC>(in, thr, majority, cls);
          c_pos += cls == 0;
          c_neg += cls == 1;
          if (cls == 3) {
            c_amb += 1u;
            if (mode == 0) {
              const unsigned int slot = atomicAdd(&st->fix_count, 1u);
              if (slot < kTiesFixCapacity)
                fix_list[slot] = ((unsigned long long)c << 32) | (unsigned int)((j * kTiesMergeThreads + threadIdx.x) * E + e);
            }
          }
        }
        st_stream(q + j * kTiesMergeThreads, o);
      }
    } else {
      const long long n = rem < CHUNK ? rem : CHUNK;
      for (long long i = threadIdx.x; i < n; i += kTiesMergeThreads) {
        S in[NSRC];
#pragma unroll
        for (int s = 0; s < NSRC; ++s) in[s] = reinterpret_cast<const S*>(sg->src[s])[base + i];
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

// Sparse fix-up: recompute the listed majority-dependent elements with the real majority (need_fix == 1).
template <int NSRC, typename S, typename D, int FUNC>
__global__ void __launch_bounds__(256)
ties_fix_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, const TiesState* __restrict__ st,
                const unsigned long long* __restrict__ fix_list) {
  if (st->need_fix != 1) return;
  constexpr int CHUNK = kTiesChunkBytes / sizeof(S);
  const float majority = (float)st->majority;
  float thr[NSRC];
#pragma unroll
  for (int s = 0; s < NSRC; ++s) thr[s] = st->thr[s];
  const unsigned int n = st->fix_count;
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned long long packed = fix_list[i];
    const MergeChunk ch = chunks[(int)(packed >> 32)];
    const MergeSeg* sg = segs + ch.seg;
    const long long idx = (long long)ch.idx * CHUNK + (long long)(unsigned int)packed;
    S in[NSRC];
#pragma unroll
    for (int s = 0; s < NSRC; ++s) in[s] = reinterpret_cast<const S*>(sg->src[s])[idx];
    int cls;
    reinterpret_cast<D*>(sg->dst)[idx] = ties_one<NSRC, S, D, FUNC>(in, thr, majority, cls);
  }
}

typedef void (*ties_fn_t)(const MergeSeg*, const MergeChunk*, int, TiesState*, unsigned long long*, int);
typedef void (*ties_fix_fn_t)(const MergeSeg*, const MergeChunk*, const TiesState*, const unsigned long long*);
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
__global__ void __launch_bounds__(kTiesMergeThreads)
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
        for (int j = 0; j < VPT; ++j) v[s][j] = ld_stream(p + j * kTiesMergeThreads);
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
      for (long long i = threadIdx.x; i < m; i += kTiesMergeThreads) {
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
