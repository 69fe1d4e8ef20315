// Modality-token splice — the integer gather/scatter of prepare_inputs_labels_for_multimodal.
//
// Replaces (reference paths): modelcompose/model/multimodal_arch.py:270-285 (modal_token_match),
// :287-459 (prepare_inputs_labels_for_multimodal) and the prefix/suffix torch.cat of
// encode_modal_inputs (:244-253).  SURVEY.md §8 rows A11-A13.
//
// The reference walks every sample on the host (a device sync per sentinel search) and issues
// O(batch x segments) tiny embed/cat/full kernels.  Here the work is three launches:
// (mc_splice_plan_scan = launches 1 + 2 and one stream sync; mc_splice_run = launch 3; a plan is reusable across
// batches of the same [B, S] and modality geometry and allocates nothing in steady state)
//   1. splice_scan_kernel    one CTA per sample: output offset of every input token (block scan),
//                            per-sample sentinel counts; the last CTA to finish turns the counts into
//                            the batch-global per-modality cursors (:302,:365), the padded length and
//                            the error flags (ragged, unknown sentinel, cursor overrun, ...).
//   2. splice_expand_kernel  one warp per input token: writes a 16-byte row descriptor for every
//                            output row it produces (1 for a text token, prefix+n+suffix for a sentinel).
//   3. splice_gather_kernel  one warp per output row: 128-bit streaming copy of the H-element row from
//                            the embedding table / feature block / prefix / suffix (or zero fill for
//                            padding), plus modal id, labels, attention mask and the reference-shaped
//                            per-modality masks.
// Roofline: HBM.  Algorithmic bytes per output row = 2*H*sizeof(dtype) (+ a few bytes of ids/masks);
// rows are copied bit-exactly.
#include <algorithm>
#include <cstring>
#include <vector>

#include "mc_common.cuh"

namespace mc {

constexpr int kMaxModal = MC_SPLICE_MAX_MODAL;
constexpr int kScanThreads = 256;
constexpr int kSearchLimit = 10000;  // multimodal_arch.py:278

enum RowKind : int { ROW_PAD = 0, ROW_EMBED = 1, ROW_PREFIX = 2, ROW_FEATURE = 3, ROW_SUFFIX = 4 };

struct SpliceModalDev {
  long long sentinel;
  int n_blocks, n_rows, n_prefix, n_suffix;
};

struct SpliceHeader {  // device + pinned-host mirror
  int max_len;
  int min_len;
  int error;        // mc_splice_error bits
  int error_b;      // first sample that raised it
  int total_blocks[kMaxModal];
};

struct SplicePlanDev {
  int B, S, n_modal, vocab;
  SpliceModalDev modal[kMaxModal];
};

// ---- 1. scan ---------------------------------------------------------------------------------------
__device__ __forceinline__ int modal_of(const SplicePlanDev& p, long long id) {
  for (int m = 0; m < p.n_modal; ++m)
    if (id == p.modal[m].sentinel) return m;
  return -1;
}

// tok_off[b*S+pos] : output offset of the token inside its sample
// tok_blk[b*S+pos] : for a sentinel: modal | (index of this sentinel among the sample's sentinels of that modality) << 8; else -1
__global__ void __launch_bounds__(kScanThreads)
splice_scan_kernel(SplicePlanDev p, const long long* __restrict__ ids, int* __restrict__ tok_off,
                   int* __restrict__ tok_blk, int* __restrict__ out_len, int* __restrict__ cnt /*[B][kMaxModal]*/,
                   int* __restrict__ base /*[B][kMaxModal]*/, SpliceHeader* __restrict__ hdr,
                   unsigned int* __restrict__ done_counter) {
  __shared__ int s_warp[kScanThreads / 32][1 + kMaxModal];
  __shared__ int s_carry[1 + kMaxModal];
  __shared__ int s_err;
  __shared__ bool s_last;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid <= kMaxModal) s_carry[tid] = 0;
  if (tid == 0) s_err = 0;
  __syncthreads();
  for (int start = 0; start < p.S; start += kScanThreads) {
    const int pos = start + tid;
    int v[1 + kMaxModal];  // v[0] = rows this token emits, v[1+m] = 1 if sentinel of modality m
#pragma unroll
    for (int k = 0; k <= kMaxModal; ++k) v[k] = 0;
    int modal = -1;
    if (pos < p.S) {
      const long long id = ids[(long long)b * p.S + pos];
      modal = modal_of(p, id);
      if (modal >= 0) {
        v[0] = p.modal[modal].n_prefix + p.modal[modal].n_rows + p.modal[modal].n_suffix;
        v[1 + modal] = 1;
      } else {
        v[0] = 1;
        // reference: embed_tokens(out-of-range id) raises IndexError, a sentinel without features KeyError
        if (id < 0 || id >= p.vocab) atomicOr(&s_err, MC_SPLICE_ERR_BAD_TOKEN);
      }
    }
    int incl[1 + kMaxModal];
#pragma unroll
    for (int k = 0; k <= kMaxModal; ++k) {
      int x = v[k];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
      }
      incl[k] = x;
      if (lane == 31) s_warp[warp][k] = x;
    }
    __syncthreads();
    int woff[1 + kMaxModal];
#pragma unroll
    for (int k = 0; k <= kMaxModal; ++k) {
      int acc = s_carry[k];
      for (int w = 0; w < warp; ++w) acc += s_warp[w][k];
      woff[k] = acc;
    }
    if (pos < p.S) {
      tok_off[(long long)b * p.S + pos] = woff[0] + incl[0] - v[0];
      tok_blk[(long long)b * p.S + pos] = modal >= 0 ? (modal | ((woff[1 + modal] + incl[1 + modal] - 1) << 8)) : -1;
    }
    __syncthreads();
    if (tid == kScanThreads - 1) {
#pragma unroll
      for (int k = 0; k <= kMaxModal; ++k) s_carry[k] = woff[k] + incl[k];
    }
    __syncthreads();
  }
  if (tid == 0) {
    out_len[b] = s_carry[0];
    for (int m = 0; m < kMaxModal; ++m) cnt[b * kMaxModal + m] = s_carry[1 + m];
    if (s_err) {
      atomicOr(&hdr->error, s_err);
      atomicMin(&hdr->error_b, b);
    }
    __threadfence();
    s_last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA: batch-global cursors (exclusive scan over samples), padded length, error flags
  __threadfence();
  if (tid < kMaxModal) {
    int acc = 0;
    for (int i = 0; i < p.B; ++i) {
      base[i * kMaxModal + tid] = acc;
      acc += cnt[i * kMaxModal + tid];
    }
    hdr->total_blocks[tid] = acc;
    if (tid < p.n_modal && acc > p.modal[tid].n_blocks) {  // reference: IndexError on modal_features[modal][cur]
      atomicOr(&hdr->error, MC_SPLICE_ERR_CURSOR);
    }
  }
  if (tid == 32) {
    int mx = 0, mn = 0x7fffffff;
    for (int i = 0; i < p.B; ++i) {
      mx = max(mx, out_len[i]);
      mn = min(mn, out_len[i]);
    }
    hdr->max_len = mx;
    hdr->min_len = p.B ? mn : 0;
  }
}

// ---- 2. expand: row descriptors ---------------------------------------------------------------------
// desc = {kind | modal << 8, source row, input position, 0}
__global__ void __launch_bounds__(256)
splice_expand_kernel(SplicePlanDev p, const long long* __restrict__ ids, const int* __restrict__ tok_off,
                     const int* __restrict__ tok_blk, const int* __restrict__ out_len, const int* __restrict__ base,
                     const SpliceHeader* __restrict__ hdr, int4* __restrict__ desc) {
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)p.B * p.S) return;
  const int b = (int)(gw / p.S), pos = (int)(gw % p.S);
  const int max_len = hdr->max_len;
  int4* row = desc + (long long)b * max_len;
  const int off = tok_off[gw], blk = tok_blk[gw];
  if (blk < 0) {
    if (lane == 0) row[off] = make_int4(ROW_EMBED, (int)ids[gw], pos, 0);
  } else {
    const int m = blk & 0xff;
    const SpliceModalDev md = p.modal[m];
    const int block = base[b * kMaxModal + m] + (blk >> 8);
    const int L = md.n_prefix + md.n_rows + md.n_suffix;
    for (int r = lane; r < L; r += 32) {
      int4 d;
      if (r < md.n_prefix) d = make_int4(ROW_PREFIX | (m << 8), r, pos, 0);
      else if (r < md.n_prefix + md.n_rows) d = make_int4(ROW_FEATURE | (m << 8), block * md.n_rows + (r - md.n_prefix), pos, 0);
      else d = make_int4(ROW_SUFFIX | (m << 8), r - md.n_prefix - md.n_rows, pos, 0);
      row[off + r] = d;
    }
  }
  if (pos == p.S - 1) {  // right padding of a ragged batch (:390-430)
    for (int j = out_len[b] + lane; j < max_len; j += 32) row[j] = make_int4(ROW_PAD, 0, -1, 0);
  }
}

// ---- 3. gather ---------------------------------------------------------------------------------------
struct SpliceRunArgs {
  const void* embed;
  const void* features[kMaxModal];
  const void* prefix[kMaxModal];
  const void* suffix[kMaxModal];
  void* mask_out[kMaxModal];
  void* default_mask_out;     // bool [B, max_len] (:452-453) or NULL
  const void* attn_in;        // [B, S] mask_elem_size bytes per element, or NULL
  void* attn_out;             // [B, max_len]
  const long long* labels_in; // [B, S] or NULL
  long long* labels_out;      // [B, max_len]
  void* embeds_out;           // [B, max_len, H]
  unsigned char* modal_id_out;  // [B, max_len] 0 = default/padding, 1+m = modality m
  int mask_elem_size;         // 1 (bool) or 8 (int64)
  int row_bytes;              // H * sizeof(dtype), multiple of 16
};

__device__ __forceinline__ void store_mask(void* base, long long idx, int elem, int v) {
  if (elem == 1) reinterpret_cast<unsigned char*>(base)[idx] = (unsigned char)v;
  else reinterpret_cast<long long*>(base)[idx] = v;
}

template <int UNROLL>
__global__ void __launch_bounds__(256)
splice_gather_kernel(SplicePlanDev p, SpliceRunArgs a, const int4* __restrict__ desc, const int* __restrict__ out_len,
                     int max_len, long long n_rows) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int n_vec = a.row_bytes >> 4;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
    const int4 d = desc[row];
    const int kind = d.x & 0xff, m = d.x >> 8;
    const int b = (int)(row / max_len), j = (int)(row % max_len);
    const char* src = nullptr;
    if (kind == ROW_EMBED) src = (const char*)a.embed + (long long)d.y * a.row_bytes;
    else if (kind == ROW_FEATURE) src = (const char*)a.features[m] + (long long)d.y * a.row_bytes;
    else if (kind == ROW_PREFIX) src = (const char*)a.prefix[m] + (long long)d.y * a.row_bytes;
    else if (kind == ROW_SUFFIX) src = (const char*)a.suffix[m] + (long long)d.y * a.row_bytes;
    Vec<16>* dst = reinterpret_cast<Vec<16>*>((char*)a.embeds_out + row * a.row_bytes);
    if (src) {
      const Vec<16>* s = reinterpret_cast<const Vec<16>*>(src);
      int i = lane;
      for (; i + (UNROLL - 1) * 32 < n_vec; i += UNROLL * 32) {
        Vec<16> v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ld_stream(s + i + u * 32);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) st_stream(dst + i + u * 32, v[u]);
      }
      for (; i < n_vec; i += 32) st_stream(dst + i, ld_stream(s + i));
    } else {
      Vec<16> z;
      z.w[0] = z.w[1] = z.w[2] = z.w[3] = 0u;
      for (int i = lane; i < n_vec; i += 32) st_stream(dst + i, z);
    }
    if (lane == 0) {
      const bool modal_row = kind >= ROW_PREFIX;
      a.modal_id_out[row] = modal_row ? (unsigned char)(1 + m) : 0;
      if (a.default_mask_out) reinterpret_cast<unsigned char*>(a.default_mask_out)[row] = modal_row ? 0 : 1;
      if (a.labels_out) a.labels_out[row] = (kind == ROW_EMBED) ? a.labels_in[(long long)b * p.S + d.z] : -100;
      if (a.attn_out) {
        // :445-448 / :418-426 — left-extend with True by the added length, right-pad with False
        const int len = out_len[b], added = len - p.S;
        int v;
        if (j < added) v = 1;
        else if (j < len) {
          const long long src_i = (long long)b * p.S + (j - added);
          v = a.mask_elem_size == 1 ? (reinterpret_cast<const unsigned char*>(a.attn_in)[src_i] != 0)
                                    : (int)reinterpret_cast<const long long*>(a.attn_in)[src_i];
        } else v = 0;
        store_mask(a.attn_out, row, a.mask_elem_size, v);
      }
    } else if (lane <= p.n_modal) {
      const int mm = lane - 1;
      if (a.mask_out[mm]) store_mask(a.mask_out[mm], row, a.mask_elem_size, (kind >= ROW_PREFIX && m == mm) ? 1 : 0);
    }
  }
}

}  // namespace mc

using namespace mc;

struct mc_splice_plan {
  SplicePlanDev dev;
  int max_len, min_len, device;
  int* d_tok_off;
  int* d_tok_blk;
  int* d_out_len;
  int* d_cnt;
  int* d_base;
  SpliceHeader* d_hdr;
  unsigned int* d_done;
  int4* d_desc;
  char* d_arena;     // one allocation behind the seven small tables above
  size_t desc_rows;  // capacity of d_desc (grow-only)
  bool scanned;
  std::vector<int> out_len;
  int total_blocks[kMaxModal];
};

static void splice_plan_free(mc_splice_plan* p) {
  if (!p) return;
  cudaFree(p->d_arena);
  cudaFree(p->d_desc);
  delete p;
}

extern "C" int mc_splice_plan_create(mc_splice_plan_t** out, int B, int S, int vocab, const mc_splice_modal_t* modals,
                                     int n_modal) {
  MC_REQUIRE(out != nullptr, "plan out-pointer is NULL");
  *out = nullptr;
  MC_REQUIRE(B >= 1 && S >= 1, "empty batch (B=%d, S=%d)", B, S);
  MC_REQUIRE(S < kSearchLimit, "S=%d: the reference stops matching sentinels at index %d (multimodal_arch.py:278)", S,
             kSearchLimit);
  MC_REQUIRE(vocab >= 1, "vocab < 1");
  MC_REQUIRE(n_modal >= 0 && n_modal <= kMaxModal && (n_modal == 0 || modals), "n_modal %d outside [0, %d]", n_modal,
             kMaxModal);
  mc_splice_plan* p = new (std::nothrow) mc_splice_plan();
  if (!p) return fail(MC_ERR_NOMEM, "host allocation failed");
  memset(&p->dev, 0, sizeof(p->dev));
  p->dev.B = B;
  p->dev.S = S;
  p->dev.n_modal = n_modal;
  p->dev.vocab = vocab;
  for (int m = 0; m < n_modal; ++m) {
    const mc_splice_modal_t& s = modals[m];
    if (s.sentinel >= 0 || s.n_blocks < 0 || s.n_rows < 0 || s.n_prefix < 0 || s.n_suffix < 0) {
      delete p;
      return fail(MC_ERR_INVALID, "modal[%d]: sentinel must be negative and counts non-negative", m);
    }
    p->dev.modal[m] = SpliceModalDev{(long long)s.sentinel, s.n_blocks, s.n_rows, s.n_prefix, s.n_suffix};
  }
  p->d_tok_off = p->d_tok_blk = p->d_out_len = p->d_cnt = p->d_base = nullptr;
  p->d_hdr = nullptr;
  p->d_done = nullptr;
  p->d_desc = nullptr;
  p->d_arena = nullptr;
  p->desc_rows = 0;
  p->scanned = false;
  p->max_len = p->min_len = 0;
  cudaError_t e = cudaGetDevice(&p->device);
  const size_t nt = (size_t)B * S;
  auto up16 = [](size_t n) { return (n + 15) & ~(size_t)15; };
  const size_t o_off = 0, o_blk = o_off + up16(nt * sizeof(int)), o_len = o_blk + up16(nt * sizeof(int)),
               o_cnt = o_len + up16((size_t)B * sizeof(int)), o_base = o_cnt + up16((size_t)B * kMaxModal * sizeof(int)),
               o_hdr = o_base + up16((size_t)B * kMaxModal * sizeof(int)), o_done = o_hdr + up16(sizeof(SpliceHeader)),
               arena_bytes = o_done + 16;
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_arena, arena_bytes);
  if (e != cudaSuccess) {
    splice_plan_free(p);
    return fail(MC_ERR_CUDA, "splice plan allocation failed: %s", cudaGetErrorString(e));
  }
  p->d_tok_off = reinterpret_cast<int*>(p->d_arena + o_off);
  p->d_tok_blk = reinterpret_cast<int*>(p->d_arena + o_blk);
  p->d_out_len = reinterpret_cast<int*>(p->d_arena + o_len);
  p->d_cnt = reinterpret_cast<int*>(p->d_arena + o_cnt);
  p->d_base = reinterpret_cast<int*>(p->d_arena + o_base);
  p->d_hdr = reinterpret_cast<SpliceHeader*>(p->d_arena + o_hdr);
  p->d_done = reinterpret_cast<unsigned int*>(p->d_arena + o_done);
  p->out_len.assign(B, 0);
  *out = p;
  return MC_OK;
}

extern "C" int mc_splice_plan_scan(mc_splice_plan_t* p, const int64_t* d_input_ids, mc_stream_t stream_) {
  MC_REQUIRE(p != nullptr, "plan is NULL");
  MC_REQUIRE(d_input_ids != nullptr, "input_ids is NULL");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int B = p->dev.B, S = p->dev.S;
  const size_t nt = (size_t)B * S;
  p->scanned = false;
  SpliceHeader h0;
  memset(&h0, 0, sizeof(h0));
  h0.error_b = 0x7fffffff;
  MC_CUDA_OK(cudaMemcpyAsync(p->d_hdr, &h0, sizeof(h0), cudaMemcpyHostToDevice, stream));
  MC_CUDA_OK(cudaMemsetAsync(p->d_done, 0, sizeof(unsigned int), stream));
  splice_scan_kernel<<<B, kScanThreads, 0, stream>>>(p->dev, (const long long*)d_input_ids, p->d_tok_off, p->d_tok_blk,
                                                     p->d_out_len, p->d_cnt, p->d_base, p->d_hdr, p->d_done);
  MC_CUDA_OK(cudaGetLastError());
  SpliceHeader h;
  MC_CUDA_OK(cudaMemcpyAsync(&h, p->d_hdr, sizeof(h), cudaMemcpyDeviceToHost, stream));
  MC_CUDA_OK(cudaMemcpyAsync(p->out_len.data(), p->d_out_len, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, stream));
  MC_CUDA_OK(cudaStreamSynchronize(stream));  // the output shape depends on the data: one sync per batch
  if (h.error & MC_SPLICE_ERR_BAD_TOKEN)
    return fail(MC_ERR_INVALID, "sample %d holds a token id that is neither in [0, vocab) nor a configured modality sentinel", h.error_b);
  if (h.error & MC_SPLICE_ERR_CURSOR) return fail(MC_ERR_INVALID, "more modality sentinels in the batch than feature blocks supplied");
  p->max_len = h.max_len;
  p->min_len = h.min_len;
  memcpy(p->total_blocks, h.total_blocks, sizeof(p->total_blocks));
  const size_t n_rows = (size_t)B * p->max_len;
  if (n_rows > p->desc_rows) {  // grow-only: steady-state batches of one shape never allocate
    if (p->d_desc) MC_CUDA_OK(cudaFree(p->d_desc));
    p->d_desc = nullptr;
    p->desc_rows = 0;
    MC_CUDA_OK(cudaMalloc((void**)&p->d_desc, n_rows * sizeof(int4)));
    p->desc_rows = n_rows;
  }
  if (n_rows) {
    const long long threads = (long long)nt * 32;
    splice_expand_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(
        p->dev, (const long long*)d_input_ids, p->d_tok_off, p->d_tok_blk, p->d_out_len, p->d_base, p->d_hdr, p->d_desc);
    MC_CUDA_OK(cudaGetLastError());
  }
  p->scanned = true;
  return MC_OK;
}

extern "C" int mc_splice_plan_info(const mc_splice_plan_t* p, int* max_len, int* min_len, int32_t* out_len,
                                   int32_t* blocks_used) {
  MC_REQUIRE(p != nullptr, "plan is NULL");
  if (max_len) *max_len = p->max_len;
  if (min_len) *min_len = p->min_len;
  if (out_len) memcpy(out_len, p->out_len.data(), p->out_len.size() * sizeof(int));
  if (blocks_used) memcpy(blocks_used, p->total_blocks, sizeof(int) * kMaxModal);
  return MC_OK;
}

extern "C" int64_t mc_splice_plan_bytes(const mc_splice_plan_t* p, int row_bytes) {
  if (!p) return 0;
  return (int64_t)p->dev.B * p->max_len * 2 * (int64_t)row_bytes;
}

extern "C" int mc_splice_run(const mc_splice_plan_t* p, const mc_splice_io_t* io, const mc_splice_modal_t* modals,
                             mc_stream_t stream) {
  MC_REQUIRE(p != nullptr && io != nullptr, "NULL plan / io");
  MC_REQUIRE(p->scanned, "mc_splice_plan_scan has not succeeded on this plan");
  MC_REQUIRE(dtype_valid(io->dtype), "bad dtype");
  MC_REQUIRE(io->hidden >= 1, "hidden < 1");
  const long long row_bytes = (long long)io->hidden * (long long)dtype_size(io->dtype);
  MC_REQUIRE(row_bytes % 16 == 0, "hidden*sizeof(dtype)=%lld must be a multiple of 16 bytes", row_bytes);
  MC_REQUIRE(io->embed_table && io->out_embeds && io->out_modal_id, "embed_table / out_embeds / out_modal_id is NULL");
  MC_REQUIRE(((uintptr_t)io->embed_table & 15) == 0 && ((uintptr_t)io->out_embeds & 15) == 0, "embed/out not 16-byte aligned");
  MC_REQUIRE(io->mask_elem_size == 1 || io->mask_elem_size == 8, "mask_elem_size must be 1 (bool) or 8 (int64)");
  MC_REQUIRE((io->attention_mask_in == nullptr) == (io->out_attention_mask == nullptr), "attention mask in/out must both be given or both NULL");
  MC_REQUIRE((io->labels_in == nullptr) == (io->out_labels == nullptr), "labels in/out must both be given or both NULL");
  MC_REQUIRE(p->dev.n_modal == 0 || modals, "modals is NULL");
  SpliceRunArgs a;
  memset(&a, 0, sizeof(a));
  for (int m = 0; m < p->dev.n_modal; ++m) {
    const mc_splice_modal_t& s = modals[m];
    const SpliceModalDev& d = p->dev.modal[m];
    MC_REQUIRE(s.sentinel == d.sentinel && s.n_blocks == d.n_blocks && s.n_rows == d.n_rows && s.n_prefix == d.n_prefix &&
                   s.n_suffix == d.n_suffix, "modal[%d] differs from the one the plan was built with", m);
    const bool used = p->total_blocks[m] > 0;
    MC_REQUIRE(!used || d.n_rows == 0 || s.features, "modal[%d].features is NULL", m);
    MC_REQUIRE(!used || d.n_prefix == 0 || s.prefix, "modal[%d].prefix is NULL", m);
    MC_REQUIRE(!used || d.n_suffix == 0 || s.suffix, "modal[%d].suffix is NULL", m);
    MC_REQUIRE((((uintptr_t)s.features | (uintptr_t)s.prefix | (uintptr_t)s.suffix) & 15) == 0, "modal[%d] pointers not 16-byte aligned", m);
    a.features[m] = s.features;
    a.prefix[m] = s.prefix;
    a.suffix[m] = s.suffix;
    a.mask_out[m] = s.mask_out;
  }
  a.embed = io->embed_table;
  a.default_mask_out = io->out_default_mask;
  a.attn_in = io->attention_mask_in;
  a.attn_out = io->out_attention_mask;
  a.labels_in = (const long long*)io->labels_in;
  a.labels_out = (long long*)io->out_labels;
  a.embeds_out = io->out_embeds;
  a.modal_id_out = (unsigned char*)io->out_modal_id;
  a.mask_elem_size = io->mask_elem_size;
  a.row_bytes = (int)row_bytes;
  const long long n_rows = (long long)p->dev.B * p->max_len;
  if (n_rows == 0) return MC_OK;
  // one warp per row, 8 rows per CTA; cap the grid at 16 CTAs per SM worth of rows-in-flight (grid-stride beyond)
  const int sms = sm_count();
  MC_REQUIRE(sms > 0, "no CUDA device");
  long long ctas = (n_rows + 7) / 8;
  ctas = std::min<long long>(ctas, (long long)sms * 64);
  splice_gather_kernel<4><<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(p->dev, a, p->d_desc, p->d_out_len, p->max_len, n_rows);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}

extern "C" int mc_splice_plan_destroy(mc_splice_plan_t* p) {
  splice_plan_free(p);
  return MC_OK;
}
