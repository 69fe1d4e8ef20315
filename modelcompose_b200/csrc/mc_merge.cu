// N-source parameter merge — one fused pass over all sources, multi-tensor, persistent grid.
//
// Replaces (reference paths): scripts/model_composition/merge_unimodal_modelcompose.py:105-112
// (`sum`/`mean`), and materialises the online-merge-reset blend that the reference evaluates per
// forward in modelcompose/model/language_model/multimodal_llama.py:130-149.
//
// Roofline: HBM.  Algorithmic bytes per element = n_src*sizeof(src) + sizeof(dst); there is no
// reuse, so every byte crosses HBM exactly once: loads are 128/256-bit, L1::no_allocate +
// L2::evict_first; stores 128/256-bit L1::no_allocate.  Arithmetic is fp32 with separate
// multiply and add (__fmul_rn/__fadd_rn are never contracted into FMA) so results are
// bit-identical to the reference's unfused torch ops.
#include <algorithm>
#include <vector>

#include "mc_merge_kernels.cuh"

namespace mc {

static merge_fn_t pick_kernel(int sdt, int ddt, int n_src, int variant) {
  if (sdt == MC_BF16 && ddt == MC_BF16) return pick_merge_bf16_bf16(n_src, variant);
  if (sdt == MC_F16 && ddt == MC_F16) return pick_merge_f16_f16(n_src, variant);
  if (sdt == MC_F32 && ddt == MC_F32) return pick_merge_f32_f32(n_src, variant);
  if (sdt == MC_BF16 && ddt == MC_F32) return pick_merge_bf16_f32(n_src, variant);
  if (sdt == MC_F16 && ddt == MC_F32) return pick_merge_f16_f32(n_src, variant);
  if (sdt == MC_F32 && ddt == MC_BF16) return pick_merge_f32_bf16(n_src, variant);
  if (sdt == MC_F32 && ddt == MC_F16) return pick_merge_f32_f16(n_src, variant);
  return nullptr;
}

}  // namespace mc

using namespace mc;

struct mc_merge_plan {
  int n_src, src_dtype, dst_dtype, variant, device;
  int nsegs, nchunks, grid, threads;
  long long total_elems;
  merge_fn_t fn;
  MergeSeg* d_segs;
  MergeChunk* d_chunks;
};

extern "C" int mc_merge_plan_create(mc_merge_plan_t** out, int n_tensors, int n_src, const void* const* src,
                                    void* const* dst, const int64_t* numel, int src_dtype, int dst_dtype, int tuning) {
  MC_REQUIRE(out != nullptr, "plan out-pointer is NULL");
  *out = nullptr;
  MC_REQUIRE(n_tensors >= 0, "n_tensors < 0");
  MC_REQUIRE(n_src >= 1 && n_src <= MC_MERGE_MAX_SRC, "n_src %d outside [1, %d]", n_src, MC_MERGE_MAX_SRC);
  MC_REQUIRE(dtype_valid(src_dtype) && dtype_valid(dst_dtype), "bad dtype");
  MC_REQUIRE(n_tensors == 0 || (src && dst && numel), "NULL table");
  // tuning == 0 → library default, chosen by the sweep in profiles/r01_merge_sweep.txt: 128-bit accesses,
  // 2 vectors per source in flight per thread, 512-thread CTAs, one CTA per chunk (7.04 TB/s on B200).
  // Explicit codes set bit 24: bits 0-7 variant, 8-15 CTAs/SM cap, 16 grid = #chunks instead of persistent.
  if (tuning == 0) tuning = (1 << 24) | (1 << 16) | 2;
  const int variant = tuning & 0xff;
  const int ctas_per_sm_override = (tuning >> 8) & 0xff;
  const bool non_persistent = (tuning >> 16) & 1;
  MC_REQUIRE(variant < kNumVariants, "tuning variant %d unknown", variant);
  merge_fn_t fn = pick_kernel(src_dtype, dst_dtype, n_src, variant);
  if (!fn) return fail(MC_ERR_UNSUPPORTED, "dtype pair (%d -> %d) not built", src_dtype, dst_dtype);

  const size_t ss = dtype_size(src_dtype), ds = dtype_size(dst_dtype);
  const Variant& V = kVariants[variant];
  const int E = V.vec_bytes / (int)std::max(ss, ds);
  const long long CHUNK = (long long)V.threads * V.unroll * E;

  // fuse tensors that are back-to-back in every source and in dst
  std::vector<MergeSeg> segs;
  for (int t = 0; t < n_tensors; ++t) {
    MC_REQUIRE(numel[t] >= 0, "numel[%d] < 0", t);
    if (numel[t] == 0) continue;
    MergeSeg s{};
    for (int k = 0; k < n_src; ++k) {
      s.src[k] = src[(size_t)k * n_tensors + t];
      MC_REQUIRE(s.src[k] != nullptr, "src[%d][%d] is NULL", k, t);
    }
    s.dst = dst[t];
    MC_REQUIRE(s.dst != nullptr, "dst[%d] is NULL", t);
    s.numel = numel[t];
    bool fused = false;
    if (!segs.empty()) {
      MergeSeg& p = segs.back();
      bool contig = (const char*)p.dst + p.numel * ds == (const char*)s.dst && p.numel + s.numel < (1LL << 40);
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
    long long n = (s.numel + CHUNK - 1) / CHUNK;
    MC_REQUIRE(n < (1LL << 31) && (long long)chunks.size() + n < (1LL << 31), "too many chunks");
    for (long long j = 0; j < n; ++j) chunks.push_back(MergeChunk{(int)i, (int)j});
    total += s.numel;
  }

  mc_merge_plan* p = new (std::nothrow) mc_merge_plan();
  if (!p) return fail(MC_ERR_NOMEM, "host allocation failed");
  p->n_src = n_src;
  p->src_dtype = src_dtype;
  p->dst_dtype = dst_dtype;
  p->variant = variant;
  p->nsegs = (int)segs.size();
  p->nchunks = (int)chunks.size();
  p->threads = V.threads;
  p->total_elems = total;
  p->fn = fn;
  p->d_segs = nullptr;
  p->d_chunks = nullptr;
  p->grid = 0;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess && p->nchunks > 0) {
    e = cudaMalloc(&p->d_segs, segs.size() * sizeof(MergeSeg));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_chunks, chunks.size() * sizeof(MergeChunk));
    if (e == cudaSuccess) e = cudaMemcpy(p->d_segs, segs.data(), segs.size() * sizeof(MergeSeg), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
      e = cudaMemcpy(p->d_chunks, chunks.data(), chunks.size() * sizeof(MergeChunk), cudaMemcpyHostToDevice);
    int per_sm = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, V.threads, 0);
    if (e == cudaSuccess) {
      if (ctas_per_sm_override) per_sm = std::min(per_sm, ctas_per_sm_override);
      per_sm = std::max(per_sm, 1);
      int sms = sm_count();
      if (sms <= 0) e = cudaErrorNoDevice;
      long long g = (long long)sms * per_sm;
      p->grid = (int)(non_persistent ? p->nchunks : std::min<long long>(g, p->nchunks));
    }
  }
  if (e != cudaSuccess) {
    cudaFree(p->d_segs);
    cudaFree(p->d_chunks);
    delete p;
    return fail(MC_ERR_CUDA, "merge plan setup failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return MC_OK;
}

extern "C" int mc_merge_plan_run(const mc_merge_plan_t* p, const float* weights, int mode, mc_stream_t stream) {
  MC_REQUIRE(p != nullptr, "plan is NULL");
  MC_REQUIRE(mode == MC_MERGE_WEIGHTED || mode == MC_MERGE_REF_SUM || mode == MC_MERGE_REF_MEAN, "bad mode %d", mode);
  MC_REQUIRE(mode != MC_MERGE_WEIGHTED || weights != nullptr, "weights is NULL");
  MC_REQUIRE(mode == MC_MERGE_WEIGHTED || p->src_dtype == p->dst_dtype,
             "reference sum/mean modes need src dtype == dst dtype");
  if (p->nchunks == 0) return MC_OK;
  MergeArgs a{};
  for (int s = 0; s < p->n_src; ++s) a.w[s] = weights ? weights[s] : 1.0f;
  a.n_float = (float)p->n_src;
  a.mode = mode;
  p->fn<<<p->grid, p->threads, 0, (cudaStream_t)stream>>>(p->d_segs, p->d_chunks, p->nchunks, a);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}

extern "C" int64_t mc_merge_plan_bytes(const mc_merge_plan_t* p) {
  if (!p) return 0;
  return (int64_t)p->total_elems * (int64_t)(p->n_src * dtype_size(p->src_dtype) + dtype_size(p->dst_dtype));
}

extern "C" int mc_merge_plan_destroy(mc_merge_plan_t* p) {
  if (!p) return MC_OK;
  cudaFree(p->d_segs);
  cudaFree(p->d_chunks);
  delete p;
  return MC_OK;
}

extern "C" int mc_merge_tensors(int n_tensors, int n_src, const void* const* src, void* const* dst,
                                const int64_t* numel, const float* weights, int mode, int src_dtype, int dst_dtype,
                                mc_stream_t stream) {
  mc_merge_plan_t* plan = nullptr;
  int rc = mc_merge_plan_create(&plan, n_tensors, n_src, src, dst, numel, src_dtype, dst_dtype, 0);
  if (rc != MC_OK) return rc;
  rc = mc_merge_plan_run(plan, weights, mode, stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  mc_merge_plan_destroy(plan);
  if (rc != MC_OK) return rc;
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "merge failed: %s", cudaGetErrorString(e));
  return MC_OK;
}

// ---- host-buffer streaming merge -----------------------------------------------------------------
// Tensors are packed into slabs (each tensor start padded to 32 B) and pipelined:
//   copy-in stream : H2D of slab i+1         (n_src copies per tensor)
//   compute stream : merge kernel on slab i   (one launch per slab)
//   copy-out stream: D2H of slab i-1
extern "C" int mc_merge_host(int n_tensors, int n_src, const void* const* h_src, void* const* h_dst,
                             const int64_t* numel, const float* weights, int mode, int src_dtype, int dst_dtype,
                             size_t staging_bytes) {
  MC_REQUIRE(n_tensors >= 0, "n_tensors < 0");
  MC_REQUIRE(n_src >= 1 && n_src <= MC_MERGE_MAX_SRC, "n_src %d outside [1, %d]", n_src, MC_MERGE_MAX_SRC);
  MC_REQUIRE(dtype_valid(src_dtype) && dtype_valid(dst_dtype), "bad dtype");
  MC_REQUIRE(n_tensors == 0 || (h_src && h_dst && numel), "NULL table");
  MC_REQUIRE(mode == MC_MERGE_WEIGHTED || src_dtype == dst_dtype, "reference sum/mean modes need src dtype == dst dtype");
  const size_t ss = dtype_size(src_dtype), ds = dtype_size(dst_dtype);
  if (staging_bytes == 0) staging_bytes = 64u << 20;
  const long long slab_elems = (long long)(staging_bytes / std::max(ss, ds)) & ~15LL;
  MC_REQUIRE(slab_elems >= 4096, "staging_bytes too small");

  // work items: (tensor, element offset, count) pieces no longer than a slab
  struct Piece { int t; long long off, n; };
  std::vector<std::vector<Piece>> slabs(1);
  long long fill = 0;
  for (int t = 0; t < n_tensors; ++t) {
    MC_REQUIRE(numel[t] >= 0, "numel[%d] < 0", t);
    long long off = 0;
    while (off < numel[t]) {
      long long room = slab_elems - fill;
      if (room < 16) {
        slabs.emplace_back();
        fill = 0;
        room = slab_elems;
      }
      long long n = std::min(room, numel[t] - off);
      slabs.back().push_back(Piece{t, off, n});
      off += n;
      fill = (fill + n + 15) & ~15LL;
    }
  }
  if (slabs.back().empty()) slabs.pop_back();
  if (slabs.empty()) return MC_OK;

  constexpr int NBUF = 2;
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t in_done[NBUF] = {}, k_done[NBUF] = {}, out_done[NBUF] = {};
  char* d_in[NBUF][MC_MERGE_MAX_SRC] = {};
  char* d_out[NBUF] = {};
  std::vector<mc_merge_plan_t*> plans;
  int rc = MC_OK;
  cudaError_t e = cudaSuccess;
#define MC_TRY(expr)                   \
  do {                                 \
    if (e == cudaSuccess) e = (expr);  \
  } while (0)
  MC_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
  MC_TRY(cudaStreamCreateWithFlags(&s_k, cudaStreamNonBlocking));
  MC_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  for (int b = 0; b < NBUF; ++b) {
    MC_TRY(cudaEventCreateWithFlags(&in_done[b], cudaEventDisableTiming));
    MC_TRY(cudaEventCreateWithFlags(&k_done[b], cudaEventDisableTiming));
    MC_TRY(cudaEventCreateWithFlags(&out_done[b], cudaEventDisableTiming));
    for (int k = 0; k < n_src; ++k) MC_TRY(cudaMalloc(&d_in[b][k], (size_t)slab_elems * ss));
    MC_TRY(cudaMalloc(&d_out[b], (size_t)slab_elems * ds));
  }
  // One plan per staging slot, not per slab: the pieces of a slab sit back to back in the slot (each padded to 16
  // elements), the merge is elementwise, so the slot is merged as ONE tensor as long as its longest slab; the pad
  // elements are computed and never copied back (the staging is zeroed once so they are defined).  54 GB through
  // 64 MB slots used to build 211 plans (allocation + upload + synchronise each) before the first byte moved.
  long long slot_len[NBUF] = {};
  for (size_t i = 0; i < slabs.size(); ++i) {
    long long pos = 0;
    for (const Piece& pc : slabs[i]) pos = (pos + pc.n + 15) & ~15LL;
    slot_len[i % NBUF] = std::max(slot_len[i % NBUF], std::min(pos, slab_elems));
  }
  for (int b = 0; b < NBUF && rc == MC_OK && e == cudaSuccess; ++b) {
    if (slot_len[b] == 0) break;
    const void* src_tab[MC_MERGE_MAX_SRC];
    for (int k = 0; k < n_src; ++k) {
      src_tab[k] = d_in[b][k];
      MC_TRY(cudaMemsetAsync(d_in[b][k], 0, (size_t)slab_elems * ss, s_in));
    }
    void* dst_tab[1] = {d_out[b]};
    const int64_t n_tab[1] = {slot_len[b]};
    mc_merge_plan_t* plan = nullptr;
    rc = mc_merge_plan_create(&plan, 1, n_src, src_tab, dst_tab, n_tab, src_dtype, dst_dtype, 0);
    if (rc == MC_OK) plans.push_back(plan);
  }
  for (size_t i = 0; i < slabs.size() && e == cudaSuccess && rc == MC_OK; ++i) {
    const int b = (int)(i % NBUF);
    const auto& pcs = slabs[i];
    // slot b's input staging is free once the kernel of slab i-NBUF has consumed it
    if (i >= NBUF) MC_TRY(cudaStreamWaitEvent(s_in, k_done[b], 0));
    long long pos = 0;
    for (size_t j = 0; j < pcs.size(); ++j) {
      const Piece& pc = pcs[j];
      for (int k = 0; k < n_src; ++k) {
        const char* hp = (const char*)h_src[(size_t)k * n_tensors + pc.t] + pc.off * ss;
        MC_TRY(cudaMemcpyAsync(d_in[b][k] + pos * ss, hp, (size_t)pc.n * ss, cudaMemcpyHostToDevice, s_in));
      }
      pos = (pos + pc.n + 15) & ~15LL;
    }
    MC_TRY(cudaEventRecord(in_done[b], s_in));
    MC_TRY(cudaStreamWaitEvent(s_k, in_done[b], 0));
    if (i >= NBUF) MC_TRY(cudaStreamWaitEvent(s_k, out_done[b], 0));  // slot b's output staging drained
    if (e != cudaSuccess) break;
    rc = mc_merge_plan_run(plans[b], weights, mode, s_k);
    if (rc != MC_OK) break;
    MC_TRY(cudaEventRecord(k_done[b], s_k));
    MC_TRY(cudaStreamWaitEvent(s_out, k_done[b], 0));
    pos = 0;
    for (size_t j = 0; j < pcs.size(); ++j) {
      const Piece& pc = pcs[j];
      char* hp = (char*)h_dst[pc.t] + pc.off * ds;
      MC_TRY(cudaMemcpyAsync(hp, d_out[b] + pos * ds, (size_t)pc.n * ds, cudaMemcpyDeviceToHost, s_out));
      pos = (pos + pc.n + 15) & ~15LL;
    }
    MC_TRY(cudaEventRecord(out_done[b], s_out));
  }
  if (s_in) cudaStreamSynchronize(s_in);
  if (s_k) cudaStreamSynchronize(s_k);
  if (s_out) {
    cudaError_t e2 = cudaStreamSynchronize(s_out);
    if (e == cudaSuccess) e = e2;
  }
#undef MC_TRY
  for (auto* p : plans) mc_merge_plan_destroy(p);
  for (int b = 0; b < NBUF; ++b) {
    for (int k = 0; k < n_src; ++k) cudaFree(d_in[b][k]);
    cudaFree(d_out[b]);
    if (in_done[b]) cudaEventDestroy(in_done[b]);
    if (k_done[b]) cudaEventDestroy(k_done[b]);
    if (out_done[b]) cudaEventDestroy(out_done[b]);
  }
  if (s_in) cudaStreamDestroy(s_in);
  if (s_k) cudaStreamDestroy(s_k);
  if (s_out) cudaStreamDestroy(s_out);
  if (rc != MC_OK) return rc;
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "mc_merge_host failed: %s", cudaGetErrorString(e));
  return MC_OK;
}
