// TIES merge: exact magnitude top-k trim, sign election, disjoint merge — multi-tensor, all on one stream.
//
// Replaces (reference paths): scripts/model_composition/ties_merging.py:88-179 as driven by do_merging (:182-222) and
// scripts/model_composition/merge_unimodal_modelcompose.py:59-64,75-85 (`ties-*`, `convert-drop-*`).
//
// Roofline: HBM.  Every pass streams the sources once with 128-bit L1::no_allocate loads:
//   select   bf16 / fp16: ties_sample_kernel (1/32 .. 1/256 of the data, 2^15 shared-memory bins; its last CTA per source turns the
//            sample into the key bracket of the k-th magnitude) -> ties_count_kernel (one pass: SIMD ">= t" counters per bin boundary
//            of the bracket; its last CTA per source picks the bin); a bracket miss, small inputs and fp32 take the full-range radix
//            passes (ties_hist_kernel, whose last CTA per source selects the digit; 1 pass of 15 bits or 3 passes of 11+10+10 bits)
//   merge    ties_merge_kernel, speculative majority +1, census of elected signs (mc_ties_kernels.cuh)
//   fix      ties_fix_kernel decides from the census: nothing / recompute the listed majority-dependent elements / a dense
//            re-merge (list overflow, or MAX with a negative majority); the unused kernels exit at once
// Algorithmic bytes per element = (passes + 1) * n_src * sizeof(src) + sizeof(dst)  (bf16, 3 sources, sum: 14 B).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mc_ties_kernels.cuh"

namespace mc {

template <typename S>
__device__ __forceinline__ unsigned int magnitude_key(S v);
template <>
__device__ __forceinline__ unsigned int magnitude_key<float>(float v) { return __float_as_uint(v) & 0x7fffffffu; }
template <>
__device__ __forceinline__ unsigned int magnitude_key<__half>(__half v) { return (unsigned int)(__half_as_ushort(v) & 0x7fffu); }
template <>
__device__ __forceinline__ unsigned int magnitude_key<__nv_bfloat16>(__nv_bfloat16 v) {
  return (unsigned int)(__bfloat16_as_ushort(v) & 0x7fffu);
}

// True in exactly one CTA of the launch per source: the one that arrives last, after every CTA's global atomics on the
// histogram are visible.  Call with all threads; the counter is back at 0 when the launch ends.
__device__ __forceinline__ bool ties_last_cta_of_source(TiesState* st, int src, unsigned int n_ctas) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&st->done[src], 1u);
    s_last = prev == n_ctas - 1u;
    if (s_last) st->done[src] = 0u;
  }
  __syncthreads();
  const bool last = s_last != 0;
  if (last) __threadfence();
  return last;
}

// Inclusive scan of one value per thread over a 1024-thread CTA: shuffles inside the warps, one pass over the 32 warp totals.
__device__ __forceinline__ unsigned long long ties_scan_1024(unsigned long long local, unsigned long long* s_warp /* [33] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long x = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  __syncthreads();  // s_warp may still be read by an earlier call
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = s_warp[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, w, d);
      if (lane >= d) w += y;
    }
    s_warp[lane] = w;
    if (lane == 31) s_warp[32] = w;  // grand total
  }
  __syncthreads();
  return x + (warp ? s_warp[warp - 1] : 0ull);
}

// The last CTA of a source in ties_hist_kernel: find the bin holding rank k_rem, append it to the prefix, clear the histogram for the next pass.
// `key_kind` on the last pass turns the finished key into the threshold value: 0 fp32 bits, 1 fp16 bits, 2 bf16 bits.
__device__ __forceinline__ void ties_select_cta(unsigned long long* gh, TiesState* st, int src, int bits, int last, int key_kind) {
  __shared__ unsigned long long s_warp[33];
  const int tid = threadIdx.x;
  const int nbins = 1 << bits;
  const int per = nbins >= 1024 ? nbins / 1024 : 1;
  const int b0 = tid * per;
  unsigned long long local = 0ull;
  if (b0 < nbins)
    for (int i = 0; i < per; ++i) local += __ldcg(gh + b0 + i);
  const unsigned long long incl = ties_scan_1024(local, s_warp), excl = incl - local;
  const unsigned long long k = st->k_rem[src];
  if (b0 < nbins && excl < k && k <= incl) {
    unsigned long long cum = excl;
    int b = b0;
    for (int i = 0; i < per; ++i) {
      const unsigned long long c = __ldcg(gh + b0 + i);
      if (cum + c >= k) {
        b = b0 + i;
        break;
      }
      cum += c;
    }
    const unsigned int key = (st->prefix[src] << bits) | (unsigned int)b;
    st->prefix[src] = key;
    st->k_rem[src] = k - cum;
    if (last) {
      float thr;
      if (key_kind == 0) thr = __uint_as_float(key);
      else if (key_kind == 1) thr = __half2float(__ushort_as_half((unsigned short)key));
      else thr = __uint_as_float(key << 16);
      st->thr[src] = thr;
    }
  }
  __syncthreads();
  if (b0 < nbins)
    for (int i = 0; i < per; ++i) gh[b0 + i] = 0ull;
}

// Histogram of digit ((key >> shift) & (2^bits - 1)) over the elements of source blockIdx.y whose higher key bits equal
// the prefix found by the earlier passes.  One shared-memory histogram per CTA (<= 128 KB), flushed with 64-bit global
// atomics; CTAs take super-chunks of 4 chunks so every thread has four 128-bit loads in flight.
template <typename S>
__global__ void __launch_bounds__(kTiesHistThreads, 1)
ties_hist_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, int nchunks, TiesState* st,
                 unsigned long long* __restrict__ ghist, int shift, int bits, int hi_shift, int last, int key_kind) {
  extern __shared__ unsigned int s_hist[];
  constexpr int E = 16 / sizeof(S);
  constexpr int CHUNK = kTiesChunkBytes / sizeof(S);
  constexpr int SUPER = 4;
  static_assert(CHUNK == kTiesHistThreads * E, "chunk geometry");
  if (!st->need_full) return;  // the sampled bracket already produced the thresholds
  const int src = blockIdx.y;
  const int nbins = 1 << bits;
  const unsigned int dmask = (unsigned int)nbins - 1u;
  const unsigned int prefix = st->prefix[src];
  for (int i = threadIdx.x; i < nbins; i += kTiesHistThreads) s_hist[i] = 0u;
  __syncthreads();
  const int nsuper = (nchunks + SUPER - 1) / SUPER;
  for (int sc = blockIdx.x; sc < nsuper; sc += gridDim.x) {
    Vec<16> v[SUPER];
    bool fast[SUPER];
#pragma unroll
    for (int j = 0; j < SUPER; ++j) {
      const int c = sc * SUPER + j;
      fast[j] = false;
      if (c < nchunks) {
        const MergeChunk ch = chunks[c];
        const MergeSeg* sg = segs + ch.seg;
        const long long base = (long long)ch.idx * CHUNK;
        if (sg->aligned && sg->numel - base >= CHUNK) {
          fast[j] = true;
          v[j] = ld_stream(reinterpret_cast<const Vec<16>*>(reinterpret_cast<const S*>(sg->src[src]) + base) + threadIdx.x);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < SUPER; ++j) {
      const int c = sc * SUPER + j;
      if (c >= nchunks) break;
      if (fast[j]) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const unsigned int key = magnitude_key<S>(reinterpret_cast<const S*>(&v[j])[e]);
          if ((key >> hi_shift) == prefix) atomicAdd(&s_hist[(key >> shift) & dmask], 1u);
        }
      } else {
        const MergeChunk ch = chunks[c];
        const MergeSeg* sg = segs + ch.seg;
        const long long base = (long long)ch.idx * CHUNK;
        const long long n = min((long long)CHUNK, sg->numel - base);
        for (long long i = threadIdx.x; i < n; i += kTiesHistThreads) {
          const unsigned int key = magnitude_key<S>(reinterpret_cast<const S*>(sg->src[src])[base + i]);
          if ((key >> hi_shift) == prefix) atomicAdd(&s_hist[(key >> shift) & dmask], 1u);
        }
      }
    }
  }
  __syncthreads();
  unsigned long long* gh = ghist + (size_t)src * nbins;
  for (int i = threadIdx.x; i < nbins; i += kTiesHistThreads) {
    const unsigned int c = s_hist[i];
    if (c) atomicAdd(gh + i, (unsigned long long)c);
  }
  if (ties_last_cta_of_source(st, src, gridDim.x)) ties_select_cta(gh, st, src, bits, last, key_kind);
}

__device__ __forceinline__ void ties_init_source(TiesState* st, int i, unsigned long long kth) {
  st->k_rem[i] = kth;
  st->prefix[i] = 0u;
  st->thr[i] = 0.0f;
  st->win_lo[i] = 0u;
  st->win_hi[i] = 0u;
  st->below[i] = 0ull;
}
__device__ __forceinline__ void ties_init_globals(TiesState* st, int need_full) {
  st->need_full = need_full;
  st->n_pos = st->n_neg = st->n_zero = st->n_amb = 0ull;
  st->majority = 1;
  st->need_fix = 0;
  st->fix_count = 0u;
}

// Start of a run without the sampling pass (the sampled path resets the state in ties_sample_kernel).
__global__ void ties_init_kernel(TiesState* st, unsigned long long kth, int n_src, int need_full) {
  const int i = threadIdx.x;
  if (i < MC_MERGE_MAX_SRC) ties_init_source(st, i, i < n_src ? kth : 0ull);
  if (i == 0) ties_init_globals(st, need_full);
}

// ---- sampled bracket (bf16 / fp16) --------------------------------------------------------------------------------
// A full-range shared-memory histogram costs one atomic per element (2.9 TB/s measured).  Instead: (1) histogram a 1/32
// sample (one 512-byte granule of every chunk), (2) bracket the k-th magnitude between the sample quantiles k/d -/+ 6 sigma,
// (3) stream everything once counting the keys below the bracket and histogramming only the keys inside it (a handful of
// bins), (4) pick the exact bin.  If the rank falls outside the bracket (or the bracket is wider than the window) the
// full-range passes run instead, so the result is always exact.
template <typename S>
__device__ __forceinline__ float key_to_float(unsigned int key);
template <>
__device__ __forceinline__ float key_to_float<__half>(unsigned int key) { return __half2float(__ushort_as_half((unsigned short)key)); }
template <>
__device__ __forceinline__ float key_to_float<__nv_bfloat16>(unsigned int key) { return __uint_as_float(key << 16); }
template <>
__device__ __forceinline__ float key_to_float<float>(unsigned int key) { return __uint_as_float(key); }

// The last sampling CTA of a source, over its 2^15-bin sample histogram: bins holding the sample ranks t - margin and t + margin,
// where t = k * n_sample / d.  Clears the histogram for the counting pass.  s_cnt: 2^15 + 2^10 words of shared memory.
__device__ __forceinline__ void ties_bracket_cta(unsigned long long* gh, TiesState* st, int src, unsigned long long kth,
                                                 unsigned long long d_total, unsigned int* s_cnt) {
  // s_cnt: the sample counts, padded (+1 word per 32) so a thread's 32 bins are conflict-free
  __shared__ unsigned long long s_warp[33];
  __shared__ unsigned int s_lo, s_hi;
  constexpr int NB = 1 << 15, PER = NB / 1024;
  const int tid = threadIdx.x;
  {  // coalesced read in two batches of 16 independent loads per thread, cleared for the counting pass on the way
    constexpr int BATCH = 16;
#pragma unroll
    for (int b0 = 0; b0 < NB; b0 += 1024 * BATCH) {
      unsigned long long g[BATCH];
#pragma unroll
      for (int i = 0; i < BATCH; ++i) g[i] = __ldcg(gh + b0 + i * 1024 + tid);
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int b = b0 + i * 1024 + tid;
        s_cnt[b + (b >> 5)] = (unsigned int)g[i];
        gh[b] = 0ull;
      }
    }
  }
  if (tid == 0) {
    s_lo = 0u;
    s_hi = (unsigned int)NB - 1u;
  }
  __syncthreads();
  unsigned long long local = 0ull;
#pragma unroll
  for (int i = 0; i < PER; ++i) local += s_cnt[tid * (PER + 1) + i];
  const unsigned long long incl = ties_scan_1024(local, s_warp);
  const float n_s = (float)s_warp[32];
  const float q = (float)((double)kth / (double)d_total);
  const float t = q * n_s, margin = 6.0f * sqrtf(n_s * q * (1.0f - q)) + 64.0f;
  const float r_lo = t - margin, r_hi = t + margin;  // sample ranks (1-based) that bracket the k-th element
  unsigned long long cum = incl - local;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const unsigned int c = s_cnt[tid * (PER + 1) + i];
    const float before = (float)cum, after = (float)(cum + c);
    if (c) {
      if (r_lo >= 1.0f && before < r_lo && r_lo <= after) s_lo = (unsigned int)(tid * PER + i);
      if (r_hi <= n_s && before < r_hi && r_hi <= after) s_hi = (unsigned int)(tid * PER + i);
    }
    cum += c;
  }
  __syncthreads();
  if (tid == 0) {
    // this source's part of the run state starts here.  A bracket wider than the window is clipped (no samples at all: bins
    // 0 .. window - 1): the window select notices if the rank lies beyond it and asks for the full passes
    ties_init_source(st, src, kth);
    st->win_lo[src] = s_lo;
    st->win_hi[src] = min(s_hi, s_lo + (unsigned int)kTiesWindowBins - 1u);
  }
}

// First kernel of a sampled run: CTA (0, 0) resets the run-wide part of the device state (nothing in this launch reads it), the
// last CTA of each source that source's part.  Each warp takes
// four chunks per trip — descriptors, then the four 512-byte granules, then the shared-memory atomics — so four DRAM round
// trips overlap instead of following one another.
template <typename S>
__global__ void __launch_bounds__(kTiesHistThreads, 1)
ties_sample_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, int nchunks, TiesState* st,
                   unsigned long long* __restrict__ ghist, unsigned long long kth, unsigned long long d_total, int every) {
  extern __shared__ unsigned int s_hist[];
  constexpr int E = 16 / sizeof(S);
  constexpr int CHUNK = kTiesChunkBytes / sizeof(S);
  constexpr int NB = 1 << 15;
  constexpr int U = 4;
  const int src = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ties_init_globals(st, 0);
  for (int i = threadIdx.x; i < NB; i += kTiesHistThreads) s_hist[i] = 0u;
  __syncthreads();
  // `every`: only every `every`-th chunk contributes its granule (large inputs: 2^21 samples per source bracket the rank as
  // tightly as 2^25 would — the bracket is 1-3 bins of the dtype either way — and the shared-memory atomics bound this pass)
  const int nsel = (nchunks + every - 1) / every;
  for (int c0 = blockIdx.x * 32 * U; c0 < nsel; c0 += gridDim.x * 32 * U) {
    const S* ptr[U];
    long long first[U], n[U];
    bool vec_ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int c = (c0 + u * 32 + warp) * every;
      ptr[u] = nullptr;
      if (c >= nchunks) continue;
      const MergeChunk ch = chunks[c];
      const MergeSeg* sg = segs + ch.seg;
      const long long base = (long long)ch.idx * CHUNK;
      n[u] = min((long long)CHUNK, sg->numel - base);
      const int granule = (int)(((unsigned int)c * 7u) % (unsigned int)kTiesSampleEvery);  // which 512 B of the chunk
      first[u] = (long long)(granule * 32 + lane) * E;
      ptr[u] = reinterpret_cast<const S*>(sg->src[src]) + base;
      vec_ok[u] = sg->aligned && first[u] + E <= n[u];
    }
    Vec<16> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (ptr[u] != nullptr && vec_ok[u]) v[u] = ld_stream(reinterpret_cast<const Vec<16>*>(ptr[u] + first[u]));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ptr[u] == nullptr) continue;
      if (vec_ok[u]) {
#pragma unroll
        for (int e = 0; e < E; ++e) atomicAdd(&s_hist[magnitude_key<S>(reinterpret_cast<const S*>(&v[u])[e])], 1u);
      } else {
        for (long long i = first[u]; i < min(first[u] + E, n[u]); ++i) atomicAdd(&s_hist[magnitude_key<S>(ptr[u][i])], 1u);
      }
    }
  }
  __syncthreads();
  unsigned long long* gh = ghist + (size_t)src * NB;
  for (int i = threadIdx.x; i < NB; i += kTiesHistThreads) {
    const unsigned int c = s_hist[i];
    if (c) atomicAdd(gh + i, (unsigned long long)c);
  }
  if (ties_last_cta_of_source(st, src, gridDim.x)) ties_bracket_cta(gh, st, src, kth, d_total, s_hist);
}

// The last counting CTA of a source: the bin of the bracket that holds rank k - below; a rank outside the bracket requests the
// full passes.  Cleans the bracket's bins for whatever pass comes next.
template <typename S>
__device__ __forceinline__ void ties_window_select_cta(unsigned long long* gh /* this source's 2^15 bins */, TiesState* st, int src,
                                                       unsigned long long kth) {
  const unsigned int lo = st->win_lo[src], hi = st->win_hi[src];
  if (threadIdx.x == 0) {
    const unsigned long long below = *reinterpret_cast<volatile unsigned long long*>(&st->below[src]);
    bool found = false;
    if (kth > below) {
      unsigned long long cum = below;
      for (unsigned int b = lo; b <= hi; ++b) {
        cum += __ldcg(gh + b);
        if (cum >= kth) {
          st->thr[src] = key_to_float<S>(b);
          found = true;
          break;
        }
      }
    }
    if (!found) atomicExch(&st->need_full, 1);
  }
  __syncthreads();
  for (unsigned int b = lo + threadIdx.x; b <= hi; b += blockDim.x) gh[b] = 0ull;
}

// The streaming pass: keys below the bracket are counted, keys inside it histogrammed (a few bins, rarely hit).
template <typename S>
__global__ void __launch_bounds__(kTiesCountThreads, 4)
ties_count_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, const void* const* __restrict__ vec,
                  int nchunks, TiesState* st, unsigned long long* __restrict__ ghist, unsigned long long kth) {
  __shared__ unsigned int s_win[kTiesWindowBins];
  __shared__ unsigned int s_below;
  if (st->need_full) return;
  constexpr int E = 16 / sizeof(S);
  constexpr int CHUNK = kTiesChunkBytes / sizeof(S);
  constexpr int VPT = 4, PAIR = 2;  // one chunk = 256 threads x 4 vectors; two chunks (eight 128-bit loads per thread) per iteration
  static_assert(CHUNK == kTiesCountThreads * VPT * E, "chunk geometry");
  const int src = blockIdx.y;
  const unsigned int lo = st->win_lo[src], span = st->win_hi[src] - lo;
  for (int i = threadIdx.x; i <= (int)span; i += kTiesCountThreads) s_win[i] = 0u;
  if (threadIdx.x == 0) s_below = 0u;
  __syncthreads();
  static_assert(sizeof(S) == 2, "the counting pass packs two 16-bit keys per word");
  const unsigned int lo2 = lo | (lo << 16), hi2 = (lo + span) | ((lo + span) << 16) | 0x80008000u;
  unsigned int below = 0u, ge_acc = 0u, n_fast = 0u;  // ge_acc = 128 x (#keys >= lo) over the n_fast vector-path elements
  unsigned long long below_total = 0ull;
  // narrow bracket (<= 4 bins, the usual outcome of the sampling): ">= t" counters for t = lo + 1 .. lo + 4 (clamped to
  // 0x8000, above every key, so that no borrow crosses the halves) replace the shared-memory bins of the vector path
  constexpr int NT = 4;
  const bool narrow = span < (unsigned int)NT;
  unsigned int tn2[NT], gen_acc[NT];
  unsigned long long bin_total[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const unsigned int t = min(lo + 1u + (unsigned int)i, 0x8000u);
    tn2[i] = t | (t << 16);
    asm volatile("" : "+r"(tn2[i]));  // keep the packed form in a register (else each use re-derives t * 0x10001)
    gen_acc[i] = 0u;
    bin_total[i] = 0ull;
  }
  // The address of every whole chunk comes from the plan's flat table (one 8-byte load per chunk, walked with a running
  // pointer; the chunk -> segment walk and its 64-bit index arithmetic cost three instructions per element), read one
  // iteration ahead so the streaming loads never wait for it.
  const int row = (int)gridDim.y + 1;
  const long long stride = (long long)gridDim.x * PAIR;
  long long c0 = (long long)blockIdx.x * PAIR;
  const void* const* vp = vec + c0 * row + src;
  const long long vstep = stride * row;
  const void* nxt[PAIR];  // whole, aligned chunk c0 + j of this source, or NULL (tail / unaligned / past the end)
#pragma unroll
  for (int j = 0; j < PAIR; ++j) nxt[j] = c0 + j < nchunks ? vp[j * row] : nullptr;
  // three instructions per threshold and word (subtract, mask, dp4a), no branch, no atomic
#define MC_TIES_COUNT_WORDS(N_T)                                                                       \
  _Pragma("unroll") for (int j = 0; j < PAIR; ++j) {                                                   \
    if (cur[j] == nullptr) continue;                                                                   \
    _Pragma("unroll") for (int u = 0; u < VPT; ++u) {                                                  \
      _Pragma("unroll") for (int w = 0; w < 4; ++w) {                                                  \
        const unsigned int k = v[j][u].w[w] | 0x80008000u;                                             \
        ge_acc = __dp4a((k - lo2) & 0x80008000u, 0x01010101u, ge_acc);                                 \
        _Pragma("unroll") for (int i = 0; i < (N_T); ++i)                                              \
          gen_acc[i] = __dp4a((k - tn2[i]) & 0x80008000u, 0x01010101u, gen_acc[i]);                    \
      }                                                                                                \
    }                                                                                                  \
  }
  for (; c0 < nchunks; c0 += stride) {
    const void* cur[PAIR];
    Vec<16> v[PAIR][VPT];
#pragma unroll
    for (int j = 0; j < PAIR; ++j) {
      cur[j] = nxt[j];
      if (cur[j] != nullptr) {
#pragma unroll
        for (int u = 0; u < VPT; ++u) v[j][u] = ld_stream(reinterpret_cast<const Vec<16>*>(cur[j]) + threadIdx.x + u * kTiesCountThreads);
      }
    }
    vp += vstep;
#pragma unroll
    for (int j = 0; j < PAIR; ++j) nxt[j] = c0 + stride + j < nchunks ? vp[j * row] : nullptr;
    // two 15-bit keys per 32-bit word: (0x8000 | key) - t keeps bit 15 iff key >= t (no borrow crosses the halves); the
    // flag bytes (0x80) are summed with one dp4a per word
    if (narrow) {
      // bracket of at most four bins: the divergent shared-memory atomic path below, taken by ~1 % of the elements,
      // doubled the instruction count of this pass.  Thresholds above the bracket are skipped (span is CTA-uniform).
      if (span == 0u) {
        MC_TIES_COUNT_WORDS(1)
      } else if (span == 1u) {
        MC_TIES_COUNT_WORDS(2)
      } else if (span == 2u) {
        MC_TIES_COUNT_WORDS(3)
      } else {
        MC_TIES_COUNT_WORDS(NT)
      }
    } else {
      // wide bracket: (0x8000 | hi) - key keeps bit 15 iff key <= hi; keys inside go to the shared-memory bins
#pragma unroll
      for (int j = 0; j < PAIR; ++j) {
        if (cur[j] == nullptr) continue;
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const unsigned int keys = v[j][u].w[w] & 0x7fff7fffu;
            const unsigned int ge = ((keys | 0x80008000u) - lo2) & 0x80008000u;
            ge_acc = __dp4a(ge, 0x01010101u, ge_acc);
            const unsigned int in = ge & (hi2 - keys);
            if (in) {
              if (in & 0x00008000u) atomicAdd(&s_win[(keys & 0xffffu) - lo], 1u);
              if (in & 0x80000000u) atomicAdd(&s_win[(keys >> 16) - lo], 1u);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PAIR; ++j) {
      if (cur[j] != nullptr) {
        n_fast += (unsigned int)(VPT * E);
      } else if (c0 + j < nchunks) {
        const MergeChunk ch = chunks[c0 + j];
        const MergeSeg* sg = segs + ch.seg;
        const long long base = (long long)ch.idx * CHUNK;
        const long long n = min((long long)CHUNK, sg->numel - base);
        for (long long i = threadIdx.x; i < n; i += kTiesCountThreads) {
          const unsigned int key = magnitude_key<S>(reinterpret_cast<const S*>(sg->src[src])[base + i]);
          below += key < lo ? 1u : 0u;
          if (key - lo <= span) atomicAdd(&s_win[key - lo], 1u);
        }
      }
    }
    if (n_fast > (1u << 24)) {  // keep the 32-bit per-thread counters (ge_acc counts in units of 128) from wrapping
      below_total += below + (n_fast - (ge_acc >> 7));
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        bin_total[i] += ((i == 0 ? ge_acc : gen_acc[i - 1]) >> 7) - (gen_acc[i] >> 7);
        if (i) gen_acc[i - 1] = 0u;
      }
      gen_acc[NT - 1] = 0u;
      below = ge_acc = n_fast = 0u;
    }
  }
#undef MC_TIES_COUNT_WORDS
  below_total += below + (n_fast - (ge_acc >> 7));
  if (narrow) {  // warp-reduced bin counts of the vector path join the scalar path's shared-memory bins
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      // counters of thresholds above the bracket were never advanced: bins beyond the span are not flushed
      unsigned long long b = bin_total[i] + (((i == 0 ? ge_acc : gen_acc[i - 1]) >> 7) - (gen_acc[i] >> 7));
#pragma unroll
      for (int d = 16; d; d >>= 1) b += __shfl_xor_sync(0xffffffffu, b, d);
      if ((threadIdx.x & 31) == 0 && b && (unsigned int)i <= span) atomicAdd(&s_win[i], (unsigned int)b);
    }
  }
  // block reduction of the below-counts (64-bit), then one global atomic per CTA
  unsigned long long x = below_total;
#pragma unroll
  for (int d = 16; d; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
  __shared__ unsigned long long s_red[kTiesCountThreads / 32];
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long tot = 0ull;
    for (int w = 0; w < kTiesCountThreads / 32; ++w) tot += s_red[w];
    if (tot) atomicAdd(&st->below[src], tot);
  }
  unsigned long long* gh = ghist + (size_t)src * (1 << 15) + lo;
  for (int i = threadIdx.x; i <= (int)span; i += kTiesCountThreads) {
    const unsigned int cnt = s_win[i];
    if (cnt) atomicAdd(gh + i, (unsigned long long)cnt);
  }
  if (ties_last_cta_of_source(st, src, gridDim.x)) ties_window_select_cta<S>(ghist + (size_t)src * (1 << 15), st, src, kth);
}

static TiesKernels pick_ties(int sdt, int n_src, int func) {
  if (sdt == MC_BF16) return pick_ties_bf16(n_src, func);
  if (sdt == MC_F16) return pick_ties_f16(n_src, func);
  if (sdt == MC_F32) return pick_ties_f32(n_src, func);
  return TiesKernels{nullptr, nullptr};
}

}  // namespace mc

using namespace mc;

struct mc_ties_plan {
  int n_src, src_dtype, dst_dtype, device, sms;
  int nsegs, nchunks;
  long long total_elems;
  MergeSeg* d_segs;
  MergeChunk* d_chunks;
  const void** d_vec;          // [nchunks][n_src + 1]: source / destination address of every whole, aligned chunk (NULL row = tail / unaligned)
  TiesState* d_state;
  unsigned long long* d_hist;  // n_src x 2^15 bins
  unsigned long long* d_fix;   // kTiesFixCapacity entries
  TiesMetricSums* d_metrics;
  bool has_dst;                // false: statistics-only plan (interference metrics), mc_ties_plan_run refuses
};

static const int kHistBinsMax = 1 << 15;

extern "C" int mc_ties_plan_create(mc_ties_plan_t** out, int n_tensors, int n_src, const void* const* src, void* const* dst,
                                   const int64_t* numel, int src_dtype, int dst_dtype) {
  MC_REQUIRE(out != nullptr, "plan out-pointer is NULL");
  *out = nullptr;
  MC_REQUIRE(n_tensors >= 0, "n_tensors < 0");
  MC_REQUIRE(n_src >= 1 && n_src <= MC_MERGE_MAX_SRC, "n_src %d outside [1, %d]", n_src, MC_MERGE_MAX_SRC);
  MC_REQUIRE(dtype_valid(src_dtype) && dtype_valid(dst_dtype), "bad dtype");
  MC_REQUIRE(dst_dtype == src_dtype || dst_dtype == MC_F32, "dst dtype must be the source dtype (sum / max) or float32 (mean)");
  MC_REQUIRE(n_tensors == 0 || (src && numel), "NULL table");
  const size_t ss = dtype_size(src_dtype), ds = dtype_size(dst_dtype);
  const long long CHUNK = kTiesChunkBytes / (long long)ss;
  const bool has_dst = dst != nullptr;  // dst == NULL: a statistics-only plan for mc_ties_plan_metrics
  std::vector<MergeSeg> segs;
  for (int t = 0; t < n_tensors; ++t) {
    MC_REQUIRE(numel[t] >= 0, "numel[%d] < 0", t);
    if (numel[t] == 0) continue;
    MergeSeg s{};
    for (int k = 0; k < n_src; ++k) {
      s.src[k] = src[(size_t)k * n_tensors + t];
      MC_REQUIRE(s.src[k] != nullptr, "src[%d][%d] is NULL", k, t);
    }
    s.dst = has_dst ? dst[t] : nullptr;
    MC_REQUIRE(!has_dst || s.dst != nullptr, "dst[%d] is NULL", t);
    s.numel = numel[t];
    bool fused = false;
    if (!segs.empty()) {
      MergeSeg& p = segs.back();
      bool contig = (!has_dst || (const char*)p.dst + p.numel * ds == (const char*)s.dst) && p.numel + s.numel < (1LL << 40);
      for (int k = 0; k < n_src && contig; ++k) contig = (const char*)p.src[k] + p.numel * ss == (const char*)s.src[k];
      if (contig) {
        p.numel += s.numel;
        fused = true;
      }
    }
    if (!fused) segs.push_back(s);
  }
  std::vector<MergeChunk> chunks;
  long long total = 0;
  for (size_t i = 0; i < segs.size(); ++i) {
    MergeSeg& s = segs[i];
    uintptr_t bits = (uintptr_t)s.dst;
    for (int k = 0; k < n_src; ++k) bits |= (uintptr_t)s.src[k];
    s.aligned = (bits & 31) == 0;
    const long long n = (s.numel + CHUNK - 1) / CHUNK;
    MC_REQUIRE(n < (1LL << 31) && (long long)chunks.size() + n < (1LL << 31), "too many chunks");
    for (long long j = 0; j < n; ++j) chunks.push_back(MergeChunk{(int)i, (int)j});
    total += s.numel;
  }
  mc_ties_plan* p = new (std::nothrow) mc_ties_plan();
  if (!p) return fail(MC_ERR_NOMEM, "host allocation failed");
  p->n_src = n_src;
  p->src_dtype = src_dtype;
  p->dst_dtype = dst_dtype;
  p->nsegs = (int)segs.size();
  p->nchunks = (int)chunks.size();
  p->total_elems = total;
  p->d_segs = nullptr;
  p->d_chunks = nullptr;
  p->d_vec = nullptr;
  p->d_state = nullptr;
  p->d_hist = nullptr;
  p->d_fix = nullptr;
  p->d_metrics = nullptr;
  p->has_dst = has_dst;
  p->sms = sm_count();
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess && p->sms <= 0) e = cudaErrorNoDevice;
  if (e == cudaSuccess) e = cudaMalloc(&p->d_state, sizeof(TiesState));
  if (e == cudaSuccess) e = cudaMemset(p->d_state, 0, sizeof(TiesState));
  if (e == cudaSuccess) e = cudaMalloc(&p->d_hist, (size_t)n_src * kHistBinsMax * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&p->d_fix, (size_t)kTiesFixCapacity * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc(&p->d_metrics, sizeof(TiesMetricSums));
  if (e == cudaSuccess) e = cudaMemset(p->d_hist, 0, (size_t)n_src * kHistBinsMax * sizeof(unsigned long long));
  if (e == cudaSuccess && p->nchunks > 0) {
    e = cudaMalloc(&p->d_segs, segs.size() * sizeof(MergeSeg));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_chunks, chunks.size() * sizeof(MergeChunk));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_segs, segs.data(), segs.size() * sizeof(MergeSeg), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_chunks, chunks.data(), chunks.size() * sizeof(MergeChunk), cudaMemcpyHostToDevice);
    // flat address table of the vector path: the streaming kernels read one row instead of walking chunk -> segment
    std::vector<const void*> vec(chunks.size() * (size_t)(n_src + 1), nullptr);
    for (size_t c = 0; c < chunks.size(); ++c) {
      const MergeSeg& sg = segs[chunks[c].seg];
      const long long base = (long long)chunks[c].idx * CHUNK;
      if (!sg.aligned || sg.numel - base < CHUNK) continue;
      for (int k = 0; k < n_src; ++k) vec[c * (n_src + 1) + k] = (const char*)sg.src[k] + base * (long long)ss;
      vec[c * (n_src + 1) + n_src] = has_dst ? (const char*)sg.dst + base * (long long)ds : nullptr;
    }
    if (e == cudaSuccess) e = cudaMalloc(&p->d_vec, vec.size() * sizeof(void*));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_vec, vec.data(), vec.size() * sizeof(void*), cudaMemcpyHostToDevice);
  }
  if (e != cudaSuccess) {
    cudaFree(p->d_segs);
    cudaFree(p->d_chunks);
    cudaFree(p->d_vec);
    cudaFree(p->d_state);
    cudaFree(p->d_hist);
    cudaFree(p->d_fix);
    cudaFree(p->d_metrics);
    delete p;
    return fail(MC_ERR_CUDA, "ties plan setup failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return MC_OK;
}

static int select_passes(int src_dtype) { return src_dtype == MC_F32 ? 3 : 1; }

// Enqueues the exact k-th-magnitude select of every source (thresholds land in the device state).
static int enqueue_select(const mc_ties_plan_t* p, int64_t kth, cudaStream_t s) {
  // 16-bit dtypes with enough data: sampled bracket + one counting pass; otherwise (and on a bracket miss, decided on
  // the device) the full-range radix select, most significant digit first.  Unneeded kernels exit at once.
  const bool sampled = p->src_dtype != MC_F32 && p->nchunks >= 1024;
  if (!sampled) ties_init_kernel<<<1, 32, 0, s>>>(p->d_state, (unsigned long long)kth, p->n_src, 1);
  if (sampled) {
    // 2^15 sample bins; the bracket step of the last CTA re-reads them in a padded layout (+1 word per 32)
    const size_t bsmem = (((size_t)1 << 15) + 1024) * sizeof(unsigned int);
    // one 512-byte granule of every `every`-th chunk: >= 2^13 chunks (2^21 samples) per source once the input is large enough
    const int every = std::max(1, std::min(8, p->nchunks >> 13));
    const int nsel = (p->nchunks + every - 1) / every;
    dim3 sgrid(std::max(1, std::min((nsel + 127) / 128, p->sms / p->n_src)), p->n_src);
    dim3 cgrid(std::max(1, std::min((p->nchunks + 1) / 2, p->sms * 4 / p->n_src)), p->n_src);
    if (p->src_dtype == MC_F16) {
      MC_CUDA_OK(cudaFuncSetAttribute(ties_sample_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
      ties_sample_kernel<__half><<<sgrid, kTiesHistThreads, bsmem, s>>>(p->d_segs, p->d_chunks, p->nchunks, p->d_state, p->d_hist,
                                                                          (unsigned long long)kth, (unsigned long long)p->total_elems, every);
      ties_count_kernel<__half><<<cgrid, kTiesCountThreads, 0, s>>>(p->d_segs, p->d_chunks, p->d_vec, p->nchunks, p->d_state, p->d_hist,
                                                                   (unsigned long long)kth);
    } else {
      MC_CUDA_OK(cudaFuncSetAttribute(ties_sample_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
      ties_sample_kernel<__nv_bfloat16><<<sgrid, kTiesHistThreads, bsmem, s>>>(p->d_segs, p->d_chunks, p->nchunks, p->d_state, p->d_hist,
                                                                          (unsigned long long)kth, (unsigned long long)p->total_elems, every);
      ties_count_kernel<__nv_bfloat16><<<cgrid, kTiesCountThreads, 0, s>>>(p->d_segs, p->d_chunks, p->d_vec, p->nchunks, p->d_state, p->d_hist,
                                                                   (unsigned long long)kth);
    }
  }
  struct Pass { int shift, bits, hi_shift; };
  static const Pass k16[] = {{0, 15, 15}};
  static const Pass k32[] = {{20, 11, 31}, {10, 10, 20}, {0, 10, 10}};
  const Pass* passes = p->src_dtype == MC_F32 ? k32 : k16;
  const int n_pass = select_passes(p->src_dtype);
  const int key_kind = p->src_dtype == MC_F32 ? 0 : (p->src_dtype == MC_F16 ? 1 : 2);
  const int nsuper = (p->nchunks + 3) / 4;
  const int gx = std::max(1, std::min(nsuper, p->sms / p->n_src));  // one resident CTA per SM (128 KB of bins), no second wave
  for (int i = 0; i < n_pass; ++i) {
    const size_t smem = ((size_t)1 << passes[i].bits) * sizeof(unsigned int);
    dim3 grid(gx, p->n_src);
    if (p->src_dtype == MC_F32) {
      MC_CUDA_OK(cudaFuncSetAttribute(ties_hist_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ties_hist_kernel<float><<<grid, kTiesHistThreads, smem, s>>>(p->d_segs, p->d_chunks, p->nchunks, p->d_state, p->d_hist,
                                                                    passes[i].shift, passes[i].bits, passes[i].hi_shift, i == n_pass - 1, key_kind);
    } else if (p->src_dtype == MC_F16) {
      MC_CUDA_OK(cudaFuncSetAttribute(ties_hist_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ties_hist_kernel<__half><<<grid, kTiesHistThreads, smem, s>>>(p->d_segs, p->d_chunks, p->nchunks, p->d_state, p->d_hist,
                                                                     passes[i].shift, passes[i].bits, passes[i].hi_shift, i == n_pass - 1, key_kind);
    } else {
      MC_CUDA_OK(cudaFuncSetAttribute(ties_hist_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ties_hist_kernel<__nv_bfloat16><<<grid, kTiesHistThreads, smem, s>>>(p->d_segs, p->d_chunks, p->nchunks, p->d_state, p->d_hist,
                                                                            passes[i].shift, passes[i].bits, passes[i].hi_shift, i == n_pass - 1, key_kind);
    }
  }
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}

extern "C" int mc_ties_plan_run(const mc_ties_plan_t* p, int64_t kth, int func, mc_stream_t stream) {
  MC_REQUIRE(p != nullptr, "plan is NULL");
  MC_REQUIRE(p->has_dst, "the plan was created without outputs (statistics only)");
  MC_REQUIRE(func == MC_TIES_SUM || func == MC_TIES_MEAN || func == MC_TIES_MAX, "bad merge function %d", func);
  MC_REQUIRE(p->dst_dtype == (func == MC_TIES_MEAN ? MC_F32 : p->src_dtype),
             "dst dtype must be float32 for MEAN and the source dtype for SUM / MAX");
  MC_REQUIRE(p->total_elems > 0, "nothing to merge");
  MC_REQUIRE(kth >= 1 && kth <= p->total_elems, "kth %lld outside [1, %lld]", (long long)kth, p->total_elems);
  const TiesKernels fn = pick_ties(p->src_dtype, p->n_src, func);
  if (!fn.merge) return fail(MC_ERR_UNSUPPORTED, "ties kernel for dtype %d not built", p->src_dtype);
  cudaStream_t s = (cudaStream_t)stream;
  const int rc = enqueue_select(p, kth, s);
  if (rc != MC_OK) return rc;
  // L2 prefetch distance of the merge pass, in chunks (= CTAs) ahead: a quarter of a generation of resident CTAs — one SM
  // count; swept on B200 in profiles/r02_ties.txt (longer leads lose the lines again before the demand loads arrive).
  // MC_TIES_PREFETCH=n overrides it (0 turns the prefetch off), a development switch for the ncu comparison.
  static const int pf_env = [] { const char* e = getenv("MC_TIES_PREFETCH"); return e ? atoi(e) : -1; }();
  const int pf_dist = pf_env >= 0 ? pf_env : p->sms;
  fn.merge<<<p->nchunks, kTiesMergeThreads, 0, s>>>(p->d_segs, p->d_chunks, p->d_vec, p->nchunks, p->d_state, p->d_fix, 0, pf_dist);
  fn.fix<<<std::min(p->sms * 4, (int)(kTiesFixCapacity / 256)), 256, 0, s>>>(p->d_segs, p->d_chunks, p->d_state, p->d_fix,
                                                                                (unsigned long long)p->total_elems);
  // dense re-merge: a grid-stride launch that is small when it turns out to be a no-op (4.4 us); each CTA asks L2 for its own next
  // chunk while it works on the current one (16 CTAs per SM x that lead: 0.604 ms against 0.613 without the prefetch and 0.646 with
  // one generation of resident CTAs, 3 x 160 M max / negative majority)
  // (MC_TIES_REMERGE="<CTAs per SM>,<prefetch distance in grids>": development switch for the sweep in profiles/r02_ties.txt)
  static const int rm_mult = [] { const char* e = getenv("MC_TIES_REMERGE"); return e ? std::max(1, atoi(e)) : 16; }();
  static const int rm_pf = [] { const char* e = getenv("MC_TIES_REMERGE"); const char* c = e ? strchr(e, ',') : nullptr; return c ? atoi(c + 1) : 1; }();
  const int rgrid = std::min(p->nchunks, p->sms * rm_mult);
  fn.merge<<<rgrid, kTiesMergeThreads, 0, s>>>(p->d_segs, p->d_chunks, p->d_vec, p->nchunks, p->d_state, p->d_fix, 1,
                                              pf_env == 0 ? 0 : rgrid * rm_pf);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}

extern "C" int mc_ties_plan_stats(const mc_ties_plan_t* p, mc_ties_stats_t* out, mc_stream_t stream) {
  MC_REQUIRE(p != nullptr && out != nullptr, "NULL argument");
  TiesState h;
  MC_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  MC_CUDA_OK(cudaMemcpy(&h, p->d_state, sizeof(h), cudaMemcpyDeviceToHost));
  for (int i = 0; i < MC_MERGE_MAX_SRC; ++i) out->threshold[i] = h.thr[i];
  out->n_pos = (int64_t)h.n_pos;
  out->n_neg = (int64_t)h.n_neg;
  out->n_zero = (int64_t)h.n_zero;
  out->n_ambiguous = (int64_t)h.n_amb;
  out->majority = h.majority;
  out->full_select_ran = h.need_full;
  out->fix_pass_ran = h.need_fix;  /* 0 none, 1 sparse fix-up of the listed elements, 2 dense re-merge */
  return MC_OK;
}

extern "C" int mc_ties_plan_metrics(const mc_ties_plan_t* p, int64_t kth, mc_interference_metrics_t* out, mc_stream_t stream) {
  MC_REQUIRE(p != nullptr && out != nullptr, "NULL argument");
  MC_REQUIRE(p->n_src >= 2, "interference metrics compare the first two sources: need n_src >= 2");
  MC_REQUIRE(p->total_elems > 0, "nothing to measure");
  MC_REQUIRE(kth >= 1 && kth <= p->total_elems, "kth %lld outside [1, %lld]", (long long)kth, p->total_elems);
  ties_metrics_fn_t fn = p->src_dtype == MC_BF16 ? pick_ties_metrics_bf16(p->n_src)
                         : p->src_dtype == MC_F16 ? pick_ties_metrics_f16(p->n_src) : pick_ties_metrics_f32(p->n_src);
  if (!fn) return fail(MC_ERR_UNSUPPORTED, "metrics kernel for dtype %d not built", p->src_dtype);
  cudaStream_t s = (cudaStream_t)stream;
  const int rc = enqueue_select(p, kth, s);
  if (rc != MC_OK) return rc;
  MC_CUDA_OK(cudaMemsetAsync(p->d_metrics, 0, sizeof(TiesMetricSums), s));
  fn<<<std::min(p->nchunks, p->sms * 8), kTiesMetricsThreads, 0, s>>>(p->d_segs, p->d_chunks, p->nchunks, p->d_state, p->d_metrics);
  MC_CUDA_OK(cudaGetLastError());
  TiesMetricSums h;
  TiesState hs;
  MC_CUDA_OK(cudaStreamSynchronize(s));
  MC_CUDA_OK(cudaMemcpy(&h, p->d_metrics, sizeof(h), cudaMemcpyDeviceToHost));
  MC_CUDA_OK(cudaMemcpy(&hs, p->d_state, sizeof(hs), cudaMemcpyDeviceToHost));
  // reference calculate_metrics.py:26-37 — L2 = sqrt(sum (x0 - x1)^2); Cosine = 1 - x0.x1 / max(|x0| |x1|, 1e-8);
  // SSD = 1 - mean over elements with sum |x| != 0 of |sum x| / sum |x| (NaN when there is none, as torch's empty mean)
  out->l2 = sqrt(h.d2);
  out->cosine = 1.0 - h.xy / std::max(sqrt(h.xx) * sqrt(h.yy), 1e-8);
  out->ssd = h.ssd_n ? 1.0 - h.ssd / (double)h.ssd_n : nan("");
  out->tssd = h.tssd_n ? 1.0 - h.tssd / (double)h.tssd_n : nan("");
  out->ssd_elements = (int64_t)h.ssd_n;
  out->tssd_elements = (int64_t)h.tssd_n;
  for (int i = 0; i < MC_MERGE_MAX_SRC; ++i) out->threshold[i] = hs.thr[i];
  return MC_OK;
}

extern "C" int mc_interference_host(int n_tensors, int n_src, const void* const* h_src, const int64_t* numel, int64_t kth,
                                    int src_dtype, mc_interference_metrics_t* out) {
  MC_REQUIRE(n_tensors >= 1 && h_src && numel && out, "NULL / empty table");
  MC_REQUIRE(n_src >= 2 && n_src <= MC_MERGE_MAX_SRC, "n_src %d outside [2, %d]", n_src, MC_MERGE_MAX_SRC);
  MC_REQUIRE(dtype_valid(src_dtype), "bad dtype");
  const size_t ss = dtype_size(src_dtype);
  std::vector<long long> off(n_tensors);
  long long total = 0;
  for (int t = 0; t < n_tensors; ++t) {
    MC_REQUIRE(numel[t] >= 0, "numel[%d] < 0", t);
    off[t] = total;
    total = (total + numel[t] + 15) & ~15LL;
  }
  MC_REQUIRE(total > 0, "nothing to measure");
  std::vector<char*> d_src(n_src, nullptr);
  mc_ties_plan_t* plan = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  for (int k = 0; k < n_src && e == cudaSuccess; ++k) e = cudaMalloc(&d_src[k], (size_t)total * ss);
  for (int k = 0; k < n_src && e == cudaSuccess; ++k)
    for (int t = 0; t < n_tensors && e == cudaSuccess; ++t)
      if (numel[t])
        e = cudaMemcpyAsync(d_src[k] + off[t] * ss, h_src[(size_t)k * n_tensors + t], (size_t)numel[t] * ss, cudaMemcpyHostToDevice, s);
  int rc = MC_OK;
  if (e == cudaSuccess) {
    std::vector<const void*> src_tab((size_t)n_src * n_tensors);
    for (int t = 0; t < n_tensors; ++t)
      for (int k = 0; k < n_src; ++k) src_tab[(size_t)k * n_tensors + t] = d_src[k] + off[t] * ss;
    rc = mc_ties_plan_create(&plan, n_tensors, n_src, src_tab.data(), nullptr, numel, src_dtype, src_dtype);
    if (rc == MC_OK) rc = mc_ties_plan_metrics(plan, kth, out, s);
  }
  if (s) {
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
  }
  if (plan) mc_ties_plan_destroy(plan);
  for (int k = 0; k < n_src; ++k) cudaFree(d_src[k]);
  if (s) cudaStreamDestroy(s);
  if (rc != MC_OK) return rc;
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "mc_interference_host failed: %s", cudaGetErrorString(e));
  return MC_OK;
}

extern "C" int64_t mc_ties_plan_bytes(const mc_ties_plan_t* p) {
  if (!p) return 0;
  const int64_t reads = (int64_t)(select_passes(p->src_dtype) + 1) * p->n_src * (int64_t)dtype_size(p->src_dtype);
  return (int64_t)p->total_elems * (reads + (int64_t)dtype_size(p->dst_dtype));
}

extern "C" int64_t mc_ties_plan_elements(const mc_ties_plan_t* p) { return p ? (int64_t)p->total_elems : 0; }

extern "C" int mc_ties_plan_destroy(mc_ties_plan_t* p) {
  if (!p) return MC_OK;
  cudaFree(p->d_segs);
  cudaFree(p->d_chunks);
  cudaFree(p->d_vec);
  cudaFree(p->d_state);
  cudaFree(p->d_hist);
  cudaFree(p->d_fix);
  cudaFree(p->d_metrics);
  delete p;
  return MC_OK;
}

extern "C" int mc_ties_host(int n_tensors, int n_src, const void* const* h_src, void* const* h_dst, const int64_t* numel,
                            int64_t kth, int func, int src_dtype, mc_ties_stats_t* stats) {
  MC_REQUIRE(n_tensors >= 1 && h_src && h_dst && numel, "NULL / empty table");
  MC_REQUIRE(n_src >= 1 && n_src <= MC_MERGE_MAX_SRC, "n_src %d outside [1, %d]", n_src, MC_MERGE_MAX_SRC);
  MC_REQUIRE(dtype_valid(src_dtype), "bad dtype");
  MC_REQUIRE(func == MC_TIES_SUM || func == MC_TIES_MEAN || func == MC_TIES_MAX, "bad merge function %d", func);
  const int dst_dtype = func == MC_TIES_MEAN ? MC_F32 : src_dtype;
  const size_t ss = dtype_size(src_dtype), ds = dtype_size(dst_dtype);
  // one slab per source and one for the output; tensor starts padded to 32 B so every tensor takes the vector path
  std::vector<long long> off(n_tensors);
  long long total = 0;
  for (int t = 0; t < n_tensors; ++t) {
    MC_REQUIRE(numel[t] >= 0, "numel[%d] < 0", t);
    off[t] = total;
    total = (total + numel[t] + 15) & ~15LL;
  }
  MC_REQUIRE(total > 0, "nothing to merge");
  std::vector<char*> d_src(n_src, nullptr);
  char* d_dst = nullptr;
  mc_ties_plan_t* plan = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  for (int k = 0; k < n_src && e == cudaSuccess; ++k) e = cudaMalloc(&d_src[k], (size_t)total * ss);
  if (e == cudaSuccess) e = cudaMalloc(&d_dst, (size_t)total * ds);
  int rc = MC_OK;
  if (e == cudaSuccess) {
    for (int k = 0; k < n_src && e == cudaSuccess; ++k)
      for (int t = 0; t < n_tensors && e == cudaSuccess; ++t)
        if (numel[t])
          e = cudaMemcpyAsync(d_src[k] + off[t] * ss, h_src[(size_t)k * n_tensors + t], (size_t)numel[t] * ss, cudaMemcpyHostToDevice, s);
  }
  if (e == cudaSuccess) {
    std::vector<const void*> src_tab((size_t)n_src * n_tensors);
    std::vector<void*> dst_tab(n_tensors);
    for (int t = 0; t < n_tensors; ++t) {
      for (int k = 0; k < n_src; ++k) src_tab[(size_t)k * n_tensors + t] = d_src[k] + off[t] * ss;
      dst_tab[t] = d_dst + off[t] * ds;
    }
    rc = mc_ties_plan_create(&plan, n_tensors, n_src, src_tab.data(), dst_tab.data(), numel, src_dtype, dst_dtype);
    if (rc == MC_OK) rc = mc_ties_plan_run(plan, kth, func, s);
    if (rc == MC_OK && stats) rc = mc_ties_plan_stats(plan, stats, s);
    for (int t = 0; t < n_tensors && rc == MC_OK && e == cudaSuccess; ++t)
      if (numel[t]) e = cudaMemcpyAsync(h_dst[t], d_dst + off[t] * ds, (size_t)numel[t] * ds, cudaMemcpyDeviceToHost, s);
  }
  if (s) {
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
  }
  if (plan) mc_ties_plan_destroy(plan);
  for (int k = 0; k < n_src; ++k) cudaFree(d_src[k]);
  cudaFree(d_dst);
  if (s) cudaStreamDestroy(s);
  if (rc != MC_OK) return rc;
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "mc_ties_host failed: %s", cudaGetErrorString(e));
  return MC_OK;
}
