// Kernel templates of the N-source parameter merge (see mc_merge.cu for the contract).
#pragma once
#include "mc_common.cuh"

namespace mc {

struct MergeSeg {
  const void* src[MC_MERGE_MAX_SRC];
  void* dst;
  long long numel;
  int aligned;  // every pointer is 32-byte aligned → vector path
  int pad;
};
struct MergeChunk {
  int seg;
  int idx;  // chunk index inside the segment; element offset = idx * CHUNK
};
struct MergeArgs {
  float w[MC_MERGE_MAX_SRC];
  float n_float;
  int mode;
};

template <int NSRC, typename S, typename D>
__device__ __forceinline__ D merge_one(const S (&in)[NSRC], const MergeArgs& a) {
  if (a.mode == MC_MERGE_WEIGHTED) {
    float acc = __fmul_rn(a.w[0], to_f32<S>(in[0]));
#pragma unroll
    for (int s = 1; s < NSRC; ++s) acc = __fadd_rn(acc, __fmul_rn(a.w[s], to_f32<S>(in[s])));
    return from_f32<D>(acc);
  }
  // reference `sum`: ((0 + t0) + t1) + ...; every add rounded in the storage dtype
  float acc = to_f32<D>(from_f32<D>(__fadd_rn(0.0f, to_f32<S>(in[0]))));
#pragma unroll
  for (int s = 1; s < NSRC; ++s) acc = to_f32<D>(from_f32<D>(__fadd_rn(acc, to_f32<S>(in[s]))));
  if (a.mode == MC_MERGE_REF_MEAN) acc = __fdiv_rn(acc, a.n_float);
  return from_f32<D>(acc);
}

// E elements per thread per vector; loads are E*sizeof(S) bytes, stores E*sizeof(D) bytes.
template <int NSRC, typename S, typename D, int E, int UNROLL, int THREADS>
__global__ void __launch_bounds__(THREADS)
merge_kernel(const MergeSeg* __restrict__ segs, const MergeChunk* __restrict__ chunks, int nchunks, MergeArgs a) {
  constexpr int CHUNK = THREADS * UNROLL * E;
  using VS = Vec<E * sizeof(S)>;
  using VD = Vec<E * sizeof(D)>;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const MergeChunk ch = chunks[c];
    const MergeSeg* sg = segs + ch.seg;
    const long long base = (long long)ch.idx * CHUNK;
    const long long rem = sg->numel - base;
    if (sg->aligned && rem >= CHUNK) {
      VS v[NSRC][UNROLL];
#pragma unroll
      for (int s = 0; s < NSRC; ++s) {
        const VS* p = reinterpret_cast<const VS*>(reinterpret_cast<const S*>(sg->src[s]) + base) + threadIdx.x;
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) v[s][j] = ld_stream(p + j * THREADS);
      }
      VD* q = reinterpret_cast<VD*>(reinterpret_cast<D*>(sg->dst) + base) + threadIdx.x;
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        VD o;
        D* oe = reinterpret_cast<D*>(&o);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          S in[NSRC];
#pragma unroll
          for (int s = 0; s < NSRC; ++s) in[s] = reinterpret_cast<const S*>(&v[s][j])[e];
          oe[e] = merge_one<NSRC, S, D>(in, a);
        }
        st_stream(q + j * THREADS, o);
      }
    } else {
      // segment tail or unaligned segment: element-granular, still coalesced
      const long long n = rem < CHUNK ? rem : CHUNK;
      for (long long i = threadIdx.x; i < n; i += THREADS) {
        S in[NSRC];
#pragma unroll
        for (int s = 0; s < NSRC; ++s) in[s] = reinterpret_cast<const S*>(sg->src[s])[base + i];
        reinterpret_cast<D*>(sg->dst)[base + i] = merge_one<NSRC, S, D>(in, a);
      }
    }
  }
}

// ---- variant table ---------------------------------------------------------------------------
struct Variant {
  int vec_bytes, unroll, threads;
};
static const Variant kVariants[] __attribute__((unused)) = {
    {16, 4, 256},  // 0 (default)
    {32, 2, 256},  // 1
    {16, 2, 512},  // 2
    {32, 1, 512},  // 3
    {32, 4, 128},  // 4
    {16, 8, 128},  // 5
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

typedef void (*merge_fn_t)(const MergeSeg*, const MergeChunk*, int, MergeArgs);

template <int NSRC, typename S, typename D, int VB, int UNROLL, int THREADS>
static merge_fn_t kernel_ptr() {
  constexpr int W = sizeof(S) > sizeof(D) ? sizeof(S) : sizeof(D);
  return merge_kernel<NSRC, S, D, VB / W, UNROLL, THREADS>;
}

template <int NSRC, typename S, typename D>
static merge_fn_t pick_variant(int variant) {
  switch (variant) {
    case 0: return kernel_ptr<NSRC, S, D, 16, 4, 256>();
    case 1: return kernel_ptr<NSRC, S, D, 32, 2, 256>();
    case 2: return kernel_ptr<NSRC, S, D, 16, 2, 512>();
    case 3: return kernel_ptr<NSRC, S, D, 32, 1, 512>();
    case 4: return kernel_ptr<NSRC, S, D, 32, 4, 128>();
    case 5: return kernel_ptr<NSRC, S, D, 16, 8, 128>();
  }
  return nullptr;
}

template <typename S, typename D>
static merge_fn_t pick_nsrc(int n_src, int variant) {
  switch (n_src) {
    case 1: return pick_variant<1, S, D>(variant);
    case 2: return pick_variant<2, S, D>(variant);
    case 3: return pick_variant<3, S, D>(variant);
    case 4: return pick_variant<4, S, D>(variant);
    case 5: return pick_variant<5, S, D>(variant);
    case 6: return pick_variant<6, S, D>(variant);
    case 7: return pick_variant<7, S, D>(variant);
    case 8: return pick_variant<8, S, D>(variant);
  }
  return nullptr;
}

// one translation unit per (src, dst) dtype pair instantiates these (mc_merge_inst.cu, -DMC_PAIR=k)
merge_fn_t pick_merge_bf16_bf16(int n_src, int variant);
merge_fn_t pick_merge_f16_f16(int n_src, int variant);
merge_fn_t pick_merge_f32_f32(int n_src, int variant);
merge_fn_t pick_merge_bf16_f32(int n_src, int variant);
merge_fn_t pick_merge_f16_f32(int n_src, int variant);
merge_fn_t pick_merge_f32_bf16(int n_src, int variant);
merge_fn_t pick_merge_f32_f16(int n_src, int variant);

}  // namespace mc
