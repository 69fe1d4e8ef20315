/*
 * modelcompose_b200 — C ABI of the B200-native ModelCompose composition hot path.
 *
 * The reference (THUNLP-MT/ModelCompose) is pure Python: it has no FFI / plugin interface.
 * The seams this library sits behind are the Python call signatures listed in SURVEY.md §8(b);
 * every entry point below names the reference code it replaces (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes stubs a reference maintainer would add.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all pointers are caller-owned.
 *   - "device pointer" arguments must be valid on the CURRENT CUDA device; kernels are
 *     enqueued on the given stream (cudaStream_t passed as void*; NULL = legacy default
 *     stream) and the call returns without synchronising unless documented otherwise.
 *   - return value: MC_OK (0) or a negative mc_status; mc_last_error() gives a
 *     thread-local human-readable message for the last failure on the calling thread.
 *   - no hidden global state except a per-device attribute cache (SM count); thread-safe
 *     across streams.  Plan objects own a few KB/MB of device memory for pointer tables.
 *   - there is NO CPU fallback: with no usable sm_100 device every compute entry point
 *     returns MC_ERR_CUDA.
 */
#ifndef MODELCOMPOSE_B200_H
#define MODELCOMPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MC_ABI_VERSION 3

#if defined(__GNUC__)
#define MC_API __attribute__((visibility("default")))
#else
#define MC_API
#endif

typedef void* mc_stream_t; /* cudaStream_t */

typedef enum mc_status {
  MC_OK = 0,
  MC_ERR_INVALID = -1,     /* bad argument (NULL pointer, size, dtype, count out of range) */
  MC_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed (see mc_last_error) */
  MC_ERR_UNSUPPORTED = -3, /* valid request this build does not cover */
  MC_ERR_NOMEM = -4
} mc_status;

typedef enum mc_dtype { MC_F32 = 0, MC_F16 = 1, MC_BF16 = 2 } mc_dtype;

MC_API int mc_abi_version(void);
MC_API const char* mc_last_error(void);
/* Fills name (<= cap bytes), SM count and compute capability (major*10+minor) of the current device. */
MC_API int mc_device_info(char* name, size_t cap, int* sm_count, int* cc);

/* ------------------------------------------------------------------------------------------------
 * N-source parameter merge
 *
 * Replaces the elementwise arithmetic of the reference merge:
 *   - scripts/model_composition/merge_unimodal_modelcompose.py:105-112  (`sum` / `mean` strategies:
 *     Python sum() of N same-shape tensors, every add rounded in the storage dtype, then `/ N`)
 *   - the materialised form of the online-merge-reset blend executed per forward in
 *     modelcompose/model/language_model/multimodal_llama.py:130-149 with the coefficients of :93-106,
 *     i.e. W_eff = (1 - Σ w_m)·W_base + Σ w_m·ckpt_m   (reference's own full-weight
 *     formula: scripts/convert_to_multimodal.py:111-113, scripts/model_composition/delta_weights_compare.py:61)
 *
 * MC_MERGE_WEIGHTED : dst = rn_dst( ((w0*s0) + w1*s1) + ... )   fp32 products, fp32 left-to-right adds,
 *                     separate multiply and add (never contracted to FMA), ONE rounding to dst dtype.
 *                     Bit-identical to torch `(w0*t0.float() + w1*t1.float() + ...).to(dst)` on CPU.
 * MC_MERGE_REF_SUM  : dst = (((0 + s0) + s1) + ...) with every add rounded to the storage dtype
 *                     (src dtype == dst dtype required; weights ignored) — the reference `sum` strategy.
 * MC_MERGE_REF_MEAN : REF_SUM followed by an IEEE division by n_src, rounded to the storage dtype.
 * ---------------------------------------------------------------------------------------------- */
typedef enum mc_merge_mode { MC_MERGE_WEIGHTED = 0, MC_MERGE_REF_SUM = 1, MC_MERGE_REF_MEAN = 2 } mc_merge_mode;

#define MC_MERGE_MAX_SRC 8

typedef struct mc_merge_plan mc_merge_plan_t;

/* Builds the device-side pointer / chunk tables for merging `n_tensors` parameter tensors from `n_src`
 * checkpoints.  src[s * n_tensors + t] is the device pointer of tensor t in source s, dst[t] its output,
 * numel[t] its element count (may be 0).  Tensors that are contiguous in every source and in dst are fused
 * into one segment.  `tuning` selects a kernel variant (0 = library default; explicit codes: bit 24 set, bits 0-7 variant,
 * bits 8-15 CTAs/SM cap, bit 16 one CTA per chunk instead of a persistent grid).  Synchronous. */
MC_API int mc_merge_plan_create(mc_merge_plan_t** plan, int n_tensors, int n_src, const void* const* src,
                         void* const* dst, const int64_t* numel, int src_dtype, int dst_dtype, int tuning);
/* Enqueues ONE kernel launch that merges every tensor of the plan.  weights: n_src host floats. */
MC_API int mc_merge_plan_run(const mc_merge_plan_t* plan, const float* weights, int mode, mc_stream_t stream);
/* Algorithmic bytes one run moves: (n_src * sizeof(src) + sizeof(dst)) * total elements. */
MC_API int64_t mc_merge_plan_bytes(const mc_merge_plan_t* plan);
MC_API int mc_merge_plan_destroy(mc_merge_plan_t* plan);

/* One-shot form (create + run + stream-synchronise + destroy) for device-resident tensors. */
MC_API int mc_merge_tensors(int n_tensors, int n_src, const void* const* src, void* const* dst, const int64_t* numel,
                     const float* weights, int mode, int src_dtype, int dst_dtype, mc_stream_t stream);

/* Host-buffer form used by the merge CLI: tensors live in HOST memory (pinned or pageable); the call
 * streams them through device staging buffers (H2D, merge kernel, D2H overlapped on three streams),
 * and returns after the last output byte has landed in h_dst.  staging_bytes = size of ONE staging
 * slab per stream buffer (0 = 64 MiB).  Allocates and frees its own staging memory. */
MC_API int mc_merge_host(int n_tensors, int n_src, const void* const* h_src, void* const* h_dst, const int64_t* numel,
                  const float* weights, int mode, int src_dtype, int dst_dtype, size_t staging_bytes);


/* ------------------------------------------------------------------------------------------------
 * TIES merge (trim, elect sign, disjoint merge) — the `ties-{sum,mean,max}`, `convert-drop-*` strategies
 *
 * Replaces scripts/model_composition/ties_merging.py:88-179 (topk_values_mask, resolve_sign, resolve_zero_signs,
 * disjoint_merge, ties_merging) as called by do_merging (:182-222) from
 * scripts/model_composition/merge_unimodal_modelcompose.py:59-64,75-85.  The reference flattens every source into one
 * vector (sorted keys) and runs torch ops over the [n_src, d] matrix; here the tensors stay where they are (pointer
 * table, as the merge plan) because every step is either elementwise or a global statistic:
 *   trim   thr_s = k-th smallest |x| of source s over ALL its tensors, exact: bf16 / fp16 — a 1/32 sample brackets the
 *          rank, one streaming pass counts the keys below the bracket and histograms the few bins inside it (a miss
 *          falls back to a full 2^15-bin histogram pass); fp32 — radix select, three histogram passes of 11+10+10 bits;
 *          m = x * (|x| >= thr_s)
 *   elect  sgn = sign(round_src(sum_s m_s)); elements whose sum is 0 take the majority sign = sign(#pos - #neg), one
 *          scalar over all elements
 *   merge  entries whose sign agrees with sgn are kept, then  SUM: round_src(sum kept)
 *          MEAN: float32( round_src(sum kept) ) / max(#kept != 0, 1)   (dst is float32: torch promotes bf16 / float32)
 *          MAX : round_src(max |kept|) * sgn                           (-0 where sgn = -1 and nothing is kept)
 * The majority sign is only known after a full pass, so the merge pass runs with the speculative majority +1 while
 * taking the census and listing the elements whose survivors cancel exactly (the only ones that depend on it); if the
 * majority turns out different those elements are recomputed from the list.  A dense second merge pass runs only when
 * the list overflows (2^20 entries) or for MAX with a negative majority (every element without survivors becomes -0).
 * Everything is enqueued on the stream; nothing synchronises.  Bit-identical to the reference's CPU torch result for finite inputs.
 * dst dtype: float32 for MC_TIES_MEAN, the source dtype otherwise.
 * ---------------------------------------------------------------------------------------------- */
typedef enum mc_ties_func { MC_TIES_SUM = 0, MC_TIES_MEAN = 1, MC_TIES_MAX = 2 } mc_ties_func;

typedef struct mc_ties_stats {
  float threshold[MC_MERGE_MAX_SRC]; /* k-th smallest |x| per source */
  int64_t n_pos, n_neg;              /* elements whose elected sign is + / - before zeros are resolved */
  int64_t n_zero;                    /* elements where no source survives the trim */
  int64_t n_ambiguous;               /* elements whose survivors cancel exactly */
  int32_t majority;                  /* sign(n_pos - n_neg) */
  int32_t full_select_ran;           /* 1: thresholds came from the full-range histogram passes (fp32, small inputs, or the
                                        sampled bracket missed); 0: from the sampled bracket + one counting pass */
  int32_t fix_pass_ran;              /* 0: speculative outputs were final; 1: the listed majority-dependent elements were
                                        recomputed; 2: dense second merge pass */
} mc_ties_stats_t;

typedef struct mc_ties_plan mc_ties_plan_t;

/* Pointer / chunk tables as mc_merge_plan_create (src[s * n_tensors + t], dst[t], numel[t]); dst_dtype must be MC_F32
 * when the plan will run MC_TIES_MEAN and src_dtype otherwise (checked at run).  dst may be NULL for a statistics-only
 * plan (mc_ties_plan_metrics).  Synchronous. */
MC_API int mc_ties_plan_create(mc_ties_plan_t** plan, int n_tensors, int n_src, const void* const* src, void* const* dst,
                        const int64_t* numel, int src_dtype, int dst_dtype);
/* Enqueues the whole TIES merge.  kth = 1-based rank (ascending magnitude) of the smallest kept element among the
 * plan's total element count d: the reference's `d - int(d * K)` (ties_merging.py:89-96); 1 <= kth <= d.
 * Six launches, decisions taken on the device, nothing synchronises.  The plan owns the run's device state (thresholds,
 * census, arrival counters): runs of ONE plan must be ordered (same stream, or events between streams). */
MC_API int mc_ties_plan_run(const mc_ties_plan_t* plan, int64_t kth, int func, mc_stream_t stream);
/* Synchronises `stream` and reports the thresholds / census of the last run. */
MC_API int mc_ties_plan_stats(const mc_ties_plan_t* plan, mc_ties_stats_t* out, mc_stream_t stream);
/* Algorithmic bytes of one run without the fix pass: (select passes + 1) reads of every source + one write of dst. */
MC_API int64_t mc_ties_plan_bytes(const mc_ties_plan_t* plan);
MC_API int64_t mc_ties_plan_elements(const mc_ties_plan_t* plan);
MC_API int mc_ties_plan_destroy(mc_ties_plan_t* plan);
/* Parameter-interference metrics of the reference's scripts/model_composition/calculate_metrics.py:26-37,53-64 over the
 * sources of a plan (dst may be NULL at plan creation for this use): L2 distance and cosine distance (1 - cos) between
 * the first two sources, and the soft sign dissimilarity 1 - mean(|sum_s x_s| / sum_s |x_s|) over the elements where the
 * denominator is non-zero, before (ssd) and after (tssd) the top-k magnitude trim of rank `kth` (as mc_ties_plan_run).
 * Per-element arithmetic is fp32 as the reference's; the reductions over elements accumulate in fp64.  Synchronises. */
typedef struct mc_interference_metrics {
  double l2, cosine, ssd, tssd;
  int64_t ssd_elements, tssd_elements; /* elements that entered the two means */
  float threshold[MC_MERGE_MAX_SRC];
} mc_interference_metrics_t;
MC_API int mc_ties_plan_metrics(const mc_ties_plan_t* plan, int64_t kth, mc_interference_metrics_t* out, mc_stream_t stream);
/* Host-buffer form: copies the sources to the device, measures, frees. */
MC_API int mc_interference_host(int n_tensors, int n_src, const void* const* h_src, const int64_t* numel, int64_t kth,
                         int src_dtype, mc_interference_metrics_t* out);
/* Host-buffer form used by the merge CLI: copies every source to the device (all of them must be resident for the two
 * global statistics), runs the plan and copies the result back; returns when h_dst is complete.  stats may be NULL. */
MC_API int mc_ties_host(int n_tensors, int n_src, const void* const* h_src, void* const* h_dst, const int64_t* numel,
                 int64_t kth, int func, int src_dtype, mc_ties_stats_t* stats);

/* ------------------------------------------------------------------------------------------------
 * Modality-token splice
 *
 * Replaces modelcompose/model/multimodal_arch.py:287-459 (prepare_inputs_labels_for_multimodal), its helper
 * modal_token_match (:270-285, sentinel ids from modelcompose/constants.py:23-30) and the prefix/suffix
 * torch.cat of encode_modal_inputs (:244-253).  Every sentinel id in input_ids is replaced by the next
 * feature block of its modality (cursor global across the batch, :302,:365) framed by that modality's
 * prefix/suffix rows; text ids are looked up in the embedding table.  Rows are copied bit-exactly.
 * Outputs have B x max_len rows; samples shorter than max_len are right-padded as the reference does for
 * ragged batches (:390-430: zero rows, IGNORE_INDEX labels, False masks).
 * ---------------------------------------------------------------------------------------------- */
#define MC_SPLICE_MAX_MODAL 6
#define MC_SPLICE_ERR_BAD_TOKEN 1 /* negative id that is no configured sentinel, or id >= vocab */
#define MC_SPLICE_ERR_CURSOR 2    /* more sentinels of a modality than feature blocks supplied */

typedef struct mc_splice_modal {
  int64_t sentinel;      /* token id marking this modality (negative), e.g. -200 vision */
  int32_t n_blocks;      /* feature blocks available: features is [n_blocks, n_rows, hidden] */
  int32_t n_rows;
  int32_t n_prefix;      /* rows of `prefix` spliced before every block (0 = none) */
  int32_t n_suffix;
  const void* features;  /* device; ignored by mc_splice_plan_create */
  const void* prefix;    /* device [n_prefix, hidden] or NULL */
  const void* suffix;    /* device [n_suffix, hidden] or NULL */
  void* mask_out;        /* device [B, max_len] of mask_elem_size bytes: 1 where the row belongs to this modality; NULL = skip */
} mc_splice_modal_t;

typedef struct mc_splice_io {
  int32_t dtype;                 /* mc_dtype of embed_table / features / out_embeds */
  int32_t hidden;                /* row length in elements; hidden * sizeof(dtype) must be a multiple of 16 */
  const void* embed_table;       /* device [vocab, hidden] */
  const void* attention_mask_in; /* device [B, S] of mask_elem_size bytes, or NULL */
  int32_t mask_elem_size;        /* 1 (torch.bool) or 8 (torch.int64): dtype of every mask in and out */
  const int64_t* labels_in;      /* device [B, S] or NULL */
  void* out_embeds;              /* device [B, max_len, hidden] */
  uint8_t* out_modal_id;         /* device [B, max_len]: 0 = text ("default") or padding, 1 + m = modality m */
  void* out_attention_mask;      /* device [B, max_len] (left-extended by the added length, :445-448), or NULL */
  int64_t* out_labels;           /* device [B, max_len] (IGNORE_INDEX = -100 over spliced rows), or NULL */
  uint8_t* out_default_mask;     /* device bool [B, max_len]: rows of no modality (:452-453), or NULL */
} mc_splice_io_t;

typedef struct mc_splice_plan mc_splice_plan_t;

/* Allocates the plan's tables for batches of shape [B, S] with this modality geometry (pointers in `modals` are
 * ignored here).  Reusable: scan + run it for every batch of that shape; nothing is allocated in steady state. */
MC_API int mc_splice_plan_create(mc_splice_plan_t** plan, int B, int S, int vocab, const mc_splice_modal_t* modals, int n_modal);
/* Scans input_ids (device, int64 [B, S]) on the GPU: output offsets, batch-global block cursors, padded length, row
 * descriptors.  Synchronises `stream` once (the output shape depends on the data).  Fails with MC_ERR_INVALID where the
 * reference raises (unknown negative id, id >= vocab, sentinel without a feature block left). */
MC_API int mc_splice_plan_scan(mc_splice_plan_t* plan, const int64_t* d_input_ids, mc_stream_t stream);
/* max_len / min_len over the batch (ragged iff they differ), per-sample lengths (B ints, may be NULL) and
 * feature blocks consumed per modality (MC_SPLICE_MAX_MODAL ints, may be NULL). */
MC_API int mc_splice_plan_info(const mc_splice_plan_t* plan, int* max_len, int* min_len, int32_t* out_len,
                        int32_t* blocks_used);
/* Algorithmic bytes of one mc_splice_run: B * max_len rows, each read once and written once. */
MC_API int64_t mc_splice_plan_bytes(const mc_splice_plan_t* plan, int row_bytes);
/* Enqueues ONE gather launch on `stream`.  `modals` must repeat the plan's geometry and carry the device pointers. */
MC_API int mc_splice_run(const mc_splice_plan_t* plan, const mc_splice_io_t* io, const mc_splice_modal_t* modals,
                  mc_stream_t stream);
MC_API int mc_splice_plan_destroy(mc_splice_plan_t* plan);

/* ------------------------------------------------------------------------------------------------
 * Grouped / modality-routed linear (tcgen05 tensor cores, TMA-fed, fp32 accumulation in TMEM)
 *
 * Replaces the GEMMs of the composed-model prefill:
 *   - modelcompose/model/language_model/multimodal_llama.py:120-160 (LocalLoraLinear.forward) as routed by
 *     :262-268,:335-336 (attention q/k/v/o) and :380-390 (MLP gate/up/down): the reference evaluates every
 *     adapter on every token and mask-sums; here a token only runs the adapter its modality mask selects
 *   - modelcompose/model/multimodal_projector/builder.py:202-219 (mlp2x_gelu / linear projectors), one
 *     problem per modality in a single launch
 * Each problem computes  C[M,N] = epilogue( A0[M,K0]·B0[N,K0]^T + A1[M,K1]·B1[N,K1]^T ), all operands
 * row-major 16-bit (bf16 or fp16, K contiguous), C in the same dtype.
 *
 * Routing (optional).  Rows carry a group id (row_group[m]: 0 = text/"default", 1 + i = modality i; the
 * splice writes it) and mtile_mask[t] is the OR of (1 << group) over the rows of 128-row tile t
 * (mc_route_tile_masks).  group_cols[0..n_groups] (HOST array) assigns a contiguous column range to every group:
 *   - epilogue MC_LINEAR_EPI_ROWMASK (the LoRA down-projection T = x·A_all^T): the ranges partition N;
 *     C[m,n] = acc * col_scale[n] if n lies in the range of row m's group, else 0; N tiles whose groups are
 *     all absent from the M tile are skipped (their C is left untouched and is never read by the next step);
 *   - K1 > 0 with n_groups > 0 (the up-projection y = x·W^T + T·B_all^T): the ranges partition K1 (boundaries
 *     multiples of 64); 64-wide K1 blocks of groups absent from the M tile are skipped.
 * ---------------------------------------------------------------------------------------------- */
#define MC_LINEAR_MAX_PROBLEMS 4
#define MC_LINEAR_MAX_SEGMENTS 8

typedef enum mc_linear_epilogue {
  MC_LINEAR_EPI_NONE = 0,
  MC_LINEAR_EPI_BIAS = 1,      /* + bias[n] */
  MC_LINEAR_EPI_BIAS_GELU = 2, /* gelu_erf(acc + bias[n])  (projector builder.py:214-217) */
  MC_LINEAR_EPI_ROWMASK = 3,   /* see above */
  MC_LINEAR_EPI_RESIDUAL = 4,  /* + residual[m,n] (decoder residual adds, multimodal_llama.py:448,461) */
  MC_LINEAR_EPI_SILU_MUL = 5,  /* silu(residual[m,n]) * acc: `residual` holds the stored gate_proj output, the product
                                  is the up_proj problem's result (multimodal_llama.py:381-388); may run in place */
  MC_LINEAR_EPI_ROPE = 6       /* rotary embedding of the q / k projection outputs (multimodal_llama.py:281-282), same
                                  rounding points as mc_rope; head_dim in {64,128,256} dividing N and the N tile; C, rope_cos and
                                  rope_sin 32-byte aligned and ldc a multiple of 16 (the epilogue moves 32-byte vectors) */
} mc_linear_epilogue;

typedef struct mc_linear_desc {
  int32_t M, N, K0, K1;       /* K1 = 0: no second product */
  const void* A0; int64_t lda0; /* device [M, K0], leading dimension in elements (multiple of 8) */
  const void* B0; int64_t ldb0; /* device [N, K0] (nn.Linear weight layout) */
  const void* A1; int64_t lda1; /* device [M, K1] or NULL */
  const void* B1; int64_t ldb1; /* device [N, K1] or NULL */
  void* C; int64_t ldc;         /* device [M, N] */
  const void* bias;             /* device [N], same dtype as C, or NULL */
  const void* residual; int64_t ldr; /* device [M, N] or NULL (RESIDUAL: addend; SILU_MUL: gate operand) */
  const float* col_scale;       /* device fp32 [N] (ROWMASK) or NULL */
  const uint8_t* row_group;     /* device [M] (ROWMASK) or NULL */
  const uint32_t* mtile_mask;   /* device [ceil(M/128)] or NULL (= every group present) */
  const int32_t* group_cols;    /* HOST [n_groups + 1] or NULL */
  int32_t n_groups;
  int32_t epilogue;             /* mc_linear_epilogue */
  const void* rope_cos;         /* ROPE: device cos / sin tables [positions, rope_head_dim], same dtype as C */
  const void* rope_sin;
  const int32_t* rope_pos;      /* ROPE: device scalar, position of the first row of every sequence (NULL = 0) */
  int32_t rope_seq_len;         /* ROPE: row m is token (rope_pos + m % rope_seq_len) of its sequence */
  int32_t rope_head_dim;
  const int32_t* c_rowmap;      /* device [M] or NULL: row m of the problem is written to row c_rowmap[m] of C (a permutation:
                                   activations kept in modality-major row order scatter back to sequence order for
                                   attention / the logits); ROPE takes the token position from the mapped row */
  /* Segmented problem (grouped GEMM over per-group weights — the materialised form W_eff,g = W + sum_a s_a B_a A_a of the
   * reference's per-forward blend, multimodal_llama.py:130-149; full-weight formula scripts/convert_to_multimodal.py:111-113):
   * rows [seg_start[g], seg_start[g+1]) of A0 / C multiply with B0_seg[g] instead of B0.  seg_start is a DEVICE array of
   * n_seg + 1 ascending row offsets (seg_start[0] >= 0, seg_start[n_seg] <= M), read by the kernel at run time: the caller
   * rewrites it per batch (rows sorted by routing group) without touching the plan.  n_seg = 0: ordinary problem.
   * Every problem of a launch must use the same seg_start / n_seg; K1 must be 0; not for ROWMASK launches or tuning 4. */
  const int32_t* seg_start;
  int32_t n_seg;                /* 0 .. MC_LINEAR_MAX_SEGMENTS */
  const void* const* B0_seg;    /* HOST array of n_seg device pointers, each [N, K0] with leading dimension ldb0 */
} mc_linear_desc_t;

typedef struct mc_linear_plan mc_linear_plan_t;

/* Encodes the TMA descriptors and tile schedule of 1..MC_LINEAR_MAX_PROBLEMS problems that run as ONE launch.
 * dtype: MC_BF16 or MC_F16.  tuning: bits 0-7 tile (0 = default, 1 = 128x128, 2 = 128x256, 3 = CTA pair 512x256,
 * 4 = CTA pair 256x256 with overlapped epilogue; 3 and 4 not for ROWMASK launches), bits 8-15 rasterisation group
 * override, bit 16 disables the compacted tile schedule of routed-N launches.  The plan stays valid while the pointers
 * in `desc` do (activations are normally static per-shape buffers, so plans are built once and reused). */
MC_API int mc_linear_plan_create(mc_linear_plan_t** plan, const mc_linear_desc_t* desc, int n_problems, int dtype, int tuning);
MC_API int mc_linear_plan_run(const mc_linear_plan_t* plan, mc_stream_t stream);
/* Nominal FLOPs of one run: sum of 2*M*N*(K0+K1) (skipped K1 blocks / N tiles are NOT subtracted). */
MC_API double mc_linear_plan_flops(const mc_linear_plan_t* plan);
MC_API int mc_linear_plan_destroy(mc_linear_plan_t* plan);
/* mtile_mask[t] = OR over rows of 128-row tile t of (1 << row_group[row]). */
MC_API int mc_route_tile_masks(const uint8_t* d_row_group, int M, uint32_t* d_mtile_mask, mc_stream_t stream);
/* Same with every tile of an aligned run of `coarsen` (1, 2 or 4) tiles carrying the union of the run: the CTA-pair kernels skip LoRA
 * k-blocks per 256 / 512 rows, so the down-projection must have produced (zeros in) the rank columns of every group of the run. */
MC_API int mc_route_tile_masks_coarse(const uint8_t* d_row_group, int M, uint32_t* d_mtile_mask, int coarsen, mc_stream_t stream);
/* Modality-major row order of a batch (the order the routed / grouped linears run in: every 128-row tile holds one adapter
 * group): a STABLE counting sort of the T = B * S' rows by routing group.  d_modal_id: device uint8 [T], the ids the splice emits
 * (mc_splice_run: 0 = text, 1 + i = i-th spliced modality); lut: HOST uint8 [n_lut <= 16] mapping those ids to routing groups of
 * the model (NULL = identity); outputs (device): perm[i] = sequence-order row held at buffer row i, inv_perm[t] = buffer row of
 * sequence row t, row_group[i] = group of buffer row i, seg_start[0 .. n_groups] = first buffer row of every group (the
 * segment table of the grouped GEMM), group_seq[t] (or NULL) = routing group of sequence row t.  One launch, no host sync. */
MC_API int mc_route_permutation(const uint8_t* d_modal_id, int T, const uint8_t* lut, int n_lut, int n_groups, int32_t* d_perm,
                         int32_t* d_inv_perm, uint8_t* d_row_group, int32_t* d_seg_start, uint8_t* d_group_seq, mc_stream_t stream);
/* out = silu(gate) * up, elementwise over [rows, cols] (multimodal_llama.py:381-388; act rounded to the storage
 * dtype before the product, as the reference's separate ops do). */
MC_API int mc_silu_mul(const void* gate, const void* up, void* out, int64_t rows, int cols, int64_t ld_gate, int64_t ld_up,
                int64_t ld_out, int dtype, mc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise glue of the decoder layer (transformers==4.31.0 semantics, which the reference star-imports at
 * modelcompose/model/language_model/multimodal_llama.py:66):
 *   mc_rmsnorm : LlamaRMSNorm (used at :405-406,:441,:455,:603) — fp32 variance, x*rsqrt(var+eps) rounded to the
 *                storage dtype, then weight * that, rounded again.
 *   mc_rope    : apply_rotary_pos_emb (:281-282) in place on q and k viewed as [tokens, n_heads, head_dim];
 *                cos/sin tables [>= pos_offset + seq_len, head_dim] in the storage dtype; position of token t is
 *                pos_offset + t % seq_len (prefill: position_ids = arange(seq_len), :526-533; decode step: seq_len 1,
 *                pos_offset = past length).
 * ---------------------------------------------------------------------------------------------- */
/* Causal self-attention of the prefill, flash-style on tcgen05 (replaces the eager attention of
 * modelcompose/model/language_model/multimodal_llama.py:295-312: QK^T / sqrt(d) + causal mask, fp32 softmax, PV — without
 * materialising the [B, heads, S, S] scores).  q / k / v: device [batch * seq_len, ld_qkv] with head h in columns
 * [h * head_dim, (h + 1) * head_dim) (the projection outputs as they are, RoPE already applied to q and k); out likewise with
 * ld_out.  out_rowmap (device int32 [batch * seq_len] or NULL): the output of token t is written to row out_rowmap[t] — the
 * modality-major buffer order of the routed linears, so no separate gather pass is needed.  head_dim must be 128; every
 * sequence has seq_len tokens and attends causally to itself (no padding mask, no past keys: those cases stay on the library). */
MC_API int mc_attention_causal(const void* q, const void* k, const void* v, void* out, int64_t ld_qkv, int64_t ld_out,
                        const int32_t* out_rowmap, int batch, int seq_len, int n_heads, int head_dim, float softmax_scale,
                        int dtype, mc_stream_t stream);
/* Same, with a tuning word (development / profiling aid; 0 = the library default, what mc_attention_causal runs):
 * bits 4-7: 1 + number of element pairs out of every 4 whose exp2 runs as a polynomial on the FMA pipe instead of the MUFU pipe
 * (0 = default: none — with persistent CTAs the all-MUFU form measured fastest); bit 8: rescale on every new row maximum instead of lazily. */
MC_API int mc_attention_causal_tuned(const void* q, const void* k, const void* v, void* out, int64_t ld_qkv, int64_t ld_out,
                              const int32_t* out_rowmap, int batch, int seq_len, int n_heads, int head_dim, float softmax_scale,
                              int dtype, int tuning, mc_stream_t stream);
/* dst[i, :] = src[index[i], :] for i < rows (bit-exact row copy, 128-bit accesses; row_bytes % 16 == 0).  Used to put the
 * spliced embeddings / the attention output into the modality-major row order the routed linears run in, and back. */
MC_API int mc_gather_rows(const void* src, int64_t ld_src_bytes, void* dst, int64_t ld_dst_bytes, const int32_t* index,
                   int64_t rows, int row_bytes, mc_stream_t stream);
MC_API int mc_rmsnorm(const void* x, const void* weight, void* out, int64_t rows, int hidden, int64_t ldx, int64_t ldo,
               float eps, int dtype, mc_stream_t stream);
MC_API int mc_rope(void* q, void* k, const void* cos_table, const void* sin_table, int64_t tokens, int seq_len, int pos_offset,
            int n_heads, int head_dim, int64_t ldq, int64_t ldk, int dtype, mc_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Decode step (one new token per sequence against a key/value cache): the generation loop behind
 *   modelcompose/model/multimodal_arch.py:290-293          (cache present: no splice, mask rebuilt over past + 1)
 *   modelcompose/model/language_model/multimodal_llama.py:436-438  (modality masks dropped: every row takes the default
 *                                                           adapter), :274-312 (attention with past_key_value), :747-767
 *   (prepare_inputs_for_generation).  With M = batch <= 64 rows every linear is a stream over its weight matrix (HBM-bound),
 * so the step runs on its own kernels instead of the 128-row tcgen05 tiles of the prefill.
 * ---------------------------------------------------------------------------------------------- */
#define MC_SKINNY_MAX_PROBLEMS 4
#define MC_SKINNY_MAX_M 64

typedef enum mc_skinny_epilogue {
  MC_SKINNY_EPI_NONE = 0,
  MC_SKINNY_EPI_RESIDUAL = 1, /* + residual[m,n] (fp32 add, one rounding; may run in place) */
  MC_SKINNY_EPI_COLSCALE = 2, /* acc * col_scale[n] in fp32 before rounding: the LoRA down-projection T = s * (x A^T) */
  MC_SKINNY_EPI_SILU_MUL = 3  /* dual problem: C = silu(gate) * up with gate = A0 B0^T + A1 B1^T, up = A0 B0u^T + A1u B1u^T;
                                 gate, silu(gate) and up each rounded to the storage dtype before the product — the rounding
                                 points of multimodal_llama.py:381-388 and of the prefill's MC_LINEAR_EPI_SILU_MUL */
} mc_skinny_epilogue;

typedef struct mc_skinny_desc {
  int32_t M, N, K0, K1;          /* M <= MC_SKINNY_MAX_M; K0, K1, N multiples of 8; K1 = 0: no second product */
  const void* A0; int64_t lda0;  /* device [M, K0] activations */
  const void* B0; int64_t ldb0;  /* device [N, K0] weights (nn.Linear layout) */
  const void* A1; int64_t lda1;  /* device [M, K1] rank-space activations of the default adapter group, or NULL */
  const void* B1; int64_t ldb1;  /* device [N, K1] (columns of B_all of that group) */
  const void* B0u;               /* SILU_MUL only: second weight matrix [N, K0] (leading dimension ldb0) */
  const void* A1u;               /* SILU_MUL only, K1 > 0: [M, K1] (leading dimension lda1) */
  const void* B1u;               /* SILU_MUL only, K1 > 0: [N, K1] (leading dimension ldb1) */
  void* C; int64_t ldc;          /* device [M, N] */
  const void* residual; int64_t ldr;
  const float* col_scale;        /* device fp32 [N] (COLSCALE) */
  int32_t epilogue;              /* mc_skinny_epilogue */
} mc_skinny_desc_t;

typedef struct mc_skinny_plan mc_skinny_plan_t;

/* C = epilogue(A0 B0^T + A1 B1^T) for 1..MC_SKINNY_MAX_PROBLEMS problems in ONE launch (q/k/v share a launch).
 * Default kernel: persistent stream-K — every SM owns an equal span of (block of 64 weight rows, 128-element K chunk)
 * iterations whatever N and K are; weight and activation boxes are staged by TMA (cp.async.bulk.tensor + mbarrier ring, ~170 KB
 * in flight per SM; the plan holds the tensor maps), fragments are read with ldmatrix, mma.sync.m16n8k16 accumulates in fp32;
 * row blocks shared by several CTAs are combined through the workspace by the last CTA to arrive, in CTA order (deterministic).
 * tuning (development / A-B): bit 4 selects the register kernel instead (weight rows streamed with 128-bit loads straight into
 * MMA fragments, K split over the warps of a CTA), bit 5 32-row blocks, bit 7 forces 128-element K chunks (default: 256 when the
 * ring keeps three stages), bits 8-11 cap the ring depth, bit 6 makes the
 * consumers skip the arithmetic (pipeline ceiling; results are garbage), bits 0-3 = 1: 16-row CTAs of the register kernel.
 * The plan stays valid while the pointers in `desc` do (decode buffers are static, plans are built once per cache).
 * workspace: device memory of mc_skinny_workspace_bytes() bytes, ZEROED once by the caller and then left alone (the kernel
 * restores it); launches that may run concurrently need separate workspaces. */
MC_API size_t mc_skinny_workspace_bytes(void);
MC_API int mc_skinny_plan_create(mc_skinny_plan_t** plan, const mc_skinny_desc_t* desc, int n_problems, int dtype, int tuning);
MC_API int mc_skinny_plan_run(const mc_skinny_plan_t* plan, void* workspace, size_t workspace_bytes, mc_stream_t stream);
/* Fuses the RMSNorm that PRODUCES the plan's activations into its launch: before the products, dst[rows, hidden] =
 * rmsnorm(src) * weight is computed (one warp per row, exactly the arithmetic of mc_rmsnorm) while the first ring of weight tiles is
 * already in flight; the problems of the plan are expected to read dst as their A0.  Stream-K kernel only.  Removes one launch
 * (>= 5.5 us inside a captured graph) per norm of the decode step. */
MC_API int mc_skinny_plan_set_norm(mc_skinny_plan_t* plan, const void* src, int64_t ld_src, const void* weight, void* dst, int64_t ld_dst,
                            int rows, int hidden, float eps);
/* weight bytes one run streams (the HBM roofline's numerator) */
MC_API int64_t mc_skinny_plan_bytes(const mc_skinny_plan_t* plan);
MC_API int mc_skinny_plan_destroy(mc_skinny_plan_t* plan);

/* Rotary embedding of the new token's q and k (rounding points of mc_rope) and append of k / v to the cache.
 * q, k_new, v_new: device [batch, n_heads * head_dim] (projection outputs, q is rotated in place); caches: device
 * [batch, n_heads, capacity, head_dim] (the transformers-4.31 past_key_value layout with room to grow: the keys of one
 * (sequence, head) are one contiguous stream); d_pos: DEVICE int32 scalar = number of tokens already cached = position of the new
 * token (read at run time, so a captured CUDA graph replays across steps); cos / sin tables [positions, head_dim]. */
MC_API int mc_decode_rope_append(void* q, const void* k_new, const void* v_new, int64_t ld_qkv, void* k_cache, void* v_cache,
                          int64_t capacity, const int32_t* d_pos, const void* cos_table, const void* sin_table, int batch,
                          int n_heads, int head_dim, int dtype, mc_stream_t stream);

/* Attention of one query token per sequence over the cached keys [0, *d_pos] (the new token included — call after
 * mc_decode_rope_append): softmax(q k^T * softmax_scale) v in fp32 (multimodal_llama.py:295-312 with q_len = 1).
 * key_mask: device uint8 [batch, ld_mask] or NULL; key j of sequence b takes part iff key_mask[b, j] != 0 (padded prompts).
 * The keys are cut into n_splits ranges per (sequence, head) that run as separate CTAs and are combined by the last one to
 * finish, in split order (deterministic); scratch: device fp32 [batch * n_heads * n_splits * (head_dim + 2)], counters: device
 * int32 [batch * n_heads], zero before the first call (the kernel leaves them zero).  head_dim must be 128. */
MC_API int mc_decode_attention(const void* q, const void* k_cache, const void* v_cache, int64_t capacity, const int32_t* d_pos,
                        const uint8_t* key_mask, int64_t ld_mask, void* out, int64_t ld_q, int64_t ld_out, int batch,
                        int n_heads, int head_dim, float softmax_scale, int n_splits, float* scratch, int32_t* counters,
                        int dtype, mc_stream_t stream);

/* mc_decode_rope_append + mc_decode_attention in ONE launch: q / k_new / v_new are the raw projection outputs [batch, ld_qkv] (q is
 * NOT modified); every CTA rotates q and k_new of its (sequence, head) itself (rounding points of mc_rope), the CTA whose key range
 * holds the new position *d_pos appends k / v to the caches, and the new key is taken from shared memory instead of being read back.
 * Same result as the two separate calls, bit for bit; one launch less per layer of a decode step. */
MC_API int mc_decode_attention_fused(const void* q, const void* k_new, const void* v_new, int64_t ld_qkv, void* k_cache, void* v_cache,
                              int64_t capacity, const int32_t* d_pos, const void* cos_table, const void* sin_table,
                              const uint8_t* key_mask, int64_t ld_mask, void* out, int64_t ld_out, int batch, int n_heads,
                              int head_dim, float softmax_scale, int n_splits, float* scratch, int32_t* counters, int dtype,
                              mc_stream_t stream);

/* Launch mode of the CALLING THREAD for the decode-chain entry points (mc_gather_rows, mc_rmsnorm, mc_skinny_plan_run,
 * mc_decode_rope_append, mc_decode_attention, mc_decode_attention_fused, mc_argmax_rows): bit 0 = programmatic dependent launch — each kernel may become
 * resident while its predecessor in the stream still runs, prefetches what does not depend on it (the skinny-linear kernel its
 * first ring of weight tiles) and waits (griddepcontrol.wait) before touching anything a predecessor writes.  Only for a chain
 * in which EVERY kernel is one of the above (the decode step); returns the previous mode.  Captured CUDA graphs keep the edges. */
MC_API int mc_set_launch_mode(int flags);

/* out[m] = index of the largest logit of row m (first index on ties), as int32 and / or int64: the greedy sampler of
 * generate() (HF greedy_search behind modelcompose/eval/model_multimodal_qa_loader.py:93-102), kept on the device so a decode
 * step needs no host round trip.  d_counter (device int32 or NULL) is incremented by one: the position counter of a captured
 * decode loop (mc_decode_rope_append / mc_decode_attention read it as d_pos in the next step). */
MC_API int mc_argmax_rows(const void* logits, int64_t ld, int rows, int cols, int32_t* out_i32, int64_t* out_i64,
                   int32_t* d_counter, int dtype, mc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MODELCOMPOSE_B200_H */
