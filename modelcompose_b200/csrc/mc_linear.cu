// Grouped / modality-routed linear on tcgen05 tensor cores (TMA -> smem -> tcgen05.mma -> TMEM -> epilogue).
//
// Replaces (reference paths):
//   modelcompose/model/language_model/multimodal_llama.py:120-160  LocalLoraLinear.forward (base + per-adapter LoRA)
//   :262-268,:335-336 (attention projections) and :380-390 (MLP): evaluate every adapter on every token, then
//                     mask-and-sum; here each token only runs its own adapter (bit-identical routing)
//   modelcompose/model/multimodal_projector/builder.py:202-219     mlp2x_gelu / linear projectors (grouped by modality)
//
// One kernel computes, for every problem p of a launch (problems = modalities for the projector, 1 otherwise):
//     C[M,N] = epilogue( A0[M,K0] · B0[N,K0]^T  +  A1[M,K1] · B1[N,K1]^T )
// The second product is the K-extension that carries the LoRA update: A1 = T (the rank-space activations,
// one column block per adapter group, zero outside a token's own group) and B1 = [s·B_g | ...].  64-wide K1
// blocks whose group has no token in the 128-row tile are skipped (per-tile group bitmask), so a tile that is
// all-image only pays rank 128, a text tile pays the concatenated default rank (N_modal·128).
// The same kernel produces T (epilogue ROWMASK: columns of group g keep only rows of group g, scaled by the
// adapter scaling; whole N tiles whose groups are absent from the M tile are skipped).
//
// Tile: 128 x BN (BN = 256 or 128) x 64, cta_group::1, bf16/fp16 operands (K-major, 128B swizzle), fp32
// accumulation in TMEM (2 accumulator stages x BN columns, so the epilogue of tile i overlaps the MMAs of
// tile i+1).  Persistent grid (one CTA per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-7 epilogue.  Roofline: tensor pipe (bf16 dense).
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "mc_tc.cuh"

namespace mc {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kMaxProb = MC_LINEAR_MAX_PROBLEMS;
constexpr int kThreads = 256;
constexpr int kMaxOwnTiles = 256;    // compact schedule: tiles one CTA may own (plan falls back to the static schedule beyond)
constexpr int kMaxMaskTiles = 2048;  // compact schedule: M tiles whose group masks are staged in shared memory

constexpr int kMaxSeg = MC_LINEAR_MAX_SEGMENTS;  // row segments (routing groups) of a segmented launch

struct alignas(64) LinProblem {
  CUtensorMap tmA0, tmB0, tmA1, tmB1;
  const CUtensorMap* tmB_seg;  // segmented launch: device array of one B0 map per row segment (the group's own weights)
  void* C;
  const void* bias;
  const float* col_scale;
  const unsigned char* row_group;
  const unsigned char* col_group;
  const unsigned int* mtile_mask;
  const unsigned int* ntile_mask;
  const unsigned int* kb1_mask;
  const void* residual;
  const void* rope_cos;  // ROPE epilogue: cos / sin tables [positions, head_dim] in the storage dtype
  const void* rope_sin;
  const int* rope_pos;   // device scalar: position of the first row of every sequence (NULL = 0)
  const int* c_rowmap;   // output row of problem row m (NULL = m): scatters a modality-major row order back to sequence order
  long long ldc, ldr;
  int M, N, nkb0, nkb1, tiles_m, tiles_n, tile_end, epilogue;
  int rope_seq_len, rope_head_dim;
};

struct alignas(64) LinParams {
  LinProblem prob[kMaxProb];
  int n_prob, total_tiles, is_f16;
  unsigned int idesc;
  int group_m;  // tile rasterisation: group_m M-tiles share a sweep over N (keeps the operand working set in L2)
  int epi_staged;  // 1: NONE / RESIDUAL / SILU_MUL epilogues go through the shared-memory transpose (coalesced global accesses)
  int compact;  // 1: every CTA first compacts the (M tile, N tile) pairs that survive routing and owns every grid-th of them
  // Segmented launch (grouped GEMM over materialised per-group weights): the rows of every problem are cut into n_seg
  // consecutive segments whose boundaries live in DEVICE memory (seg_start[0 .. n_seg], written per batch by the host code
  // that sorts the rows by routing group — no host synchronisation, no re-planning); segment g multiplies with tmB_seg[g].
  // M tiles never straddle a segment: tile counts are derived in the kernel prologue.
  const int* seg_start;
  int n_seg;                        // 0 = ordinary launch
  int bm;                           // rows of an M tile (128; 512 / 256 for the CTA-pair kernels)
  int cum_tiles_n[kMaxProb + 1];    // prefix sums of tiles_n over the problems (segmented launches share the M tiling)
};

struct SegTable {  // shared memory, filled in the kernel prologue of a segmented launch
  int row[kMaxSeg + 1];   // first row of segment g
  int tile[kMaxSeg + 1];  // first M tile of segment g; tile[n_seg] = number of M tiles
};

__device__ __forceinline__ void seg_table_fill(const LinParams& P, SegTable* seg) {  // one thread
  int acc = 0;
  for (int g = 0; g <= P.n_seg; ++g) {
    const int r = P.seg_start[g];
    seg->row[g] = r;
    seg->tile[g] = acc;
    if (g < P.n_seg) acc += (P.seg_start[g + 1] - r + P.bm - 1) / P.bm;
  }
}

// ---- tile scheduling (identical in all three roles) -----------------------------------------------------
struct Tile {
  int p, mt, nt;
  int m0, m_end;       // first row of the tile, and the row bound of its problem / segment (rows >= m_end are not stored)
  int seg;             // row segment (segmented launches), else -1
  unsigned int gmask;  // groups present in the M tile (all ones when unrouted)
  bool skip;
};

__device__ __forceinline__ int total_tiles(const LinParams& P, const SegTable* seg) {
  return P.n_seg ? seg->tile[P.n_seg] * P.cum_tiles_n[P.n_prob] : P.total_tiles;
}

__device__ __forceinline__ Tile decode_tile(const LinParams& P, int tile, const SegTable* seg) {
  Tile t;
  int p = 0, begin = 0, tiles_m;
  if (P.n_seg) {
    tiles_m = seg->tile[P.n_seg];
    while (p < P.n_prob - 1 && tile >= tiles_m * P.cum_tiles_n[p + 1]) ++p;
    begin = tiles_m * P.cum_tiles_n[p];
  } else {
    while (p < P.n_prob - 1 && tile >= P.prob[p].tile_end) {
      begin = P.prob[p].tile_end;
      ++p;
    }
    tiles_m = P.prob[p].tiles_m;
  }
  const LinProblem& pr = P.prob[p];
  const int local = tile - begin;
  const int per_group = P.group_m * pr.tiles_n;
  const int g = local / per_group, within = local % per_group;
  const int gsize = min(P.group_m, tiles_m - g * P.group_m);
  t.p = p;
  t.mt = g * P.group_m + within % gsize;
  t.nt = within / gsize;
  if (P.n_seg) {
    int sg = 0;
    while (sg < P.n_seg - 1 && t.mt >= seg->tile[sg + 1]) ++sg;
    t.seg = sg;
    t.m0 = seg->row[sg] + (t.mt - seg->tile[sg]) * P.bm;
    t.m_end = seg->row[sg + 1];
    t.gmask = 0xffffffffu;
    t.skip = false;
    return t;
  }
  t.seg = -1;
  t.m0 = t.mt * P.bm;
  t.m_end = pr.M;
  t.gmask = pr.mtile_mask ? pr.mtile_mask[t.mt] : 0xffffffffu;
  t.skip = pr.ntile_mask != nullptr && (pr.ntile_mask[t.nt] & t.gmask) == 0u;
  return t;
}

// Walks the tiles this CTA owns: the static round-robin schedule, or the compacted list built in the prologue.
struct TileIter {
  const LinParams& P;
  const int* own;
  const unsigned int* mm;
  const SegTable* seg;
  int n_own, it, total;
  __device__ __forceinline__ TileIter(const LinParams& P_, const int* own_, int n_own_, const unsigned int* mm_, const SegTable* seg_)
      : P(P_), own(own_), mm(mm_), seg(seg_), n_own(n_own_), it(0), total(total_tiles(P_, seg_)) {}
  __device__ __forceinline__ bool next(Tile& t) {
    if (P.compact) {
      if (it >= n_own) return false;
      const int packed = own[it++];
      t.p = packed >> 28;
      t.mt = (packed >> 10) & 0x3ffff;
      t.nt = packed & 0x3ff;
      t.m0 = t.mt * kBM;
      t.m_end = P.prob[t.p].M;
      t.seg = -1;
      t.gmask = mm[t.mt];
      t.skip = false;
      return true;
    }
    for (;;) {
      const int tile = blockIdx.x + (it++) * gridDim.x;
      if (tile >= total) return false;
      t = decode_tile(P, tile, seg);
      if (!t.skip) return true;
    }
  }
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ void unpack8(const uint4& u, bool is_f16, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (is_f16) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      f[2 * i] = __low2float(h);
      f[2 * i + 1] = __high2float(h);
    } else {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      f[2 * i] = __low2float(h);
      f[2 * i + 1] = __high2float(h);
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], bool is_f16) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (is_f16) {
      const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// One 128-row x BN-column accumulator tile: TMEM -> registers -> epilogue -> global.  Executed by the 4 epilogue warps;
// `row` is this thread's output row, `taddr` the TMEM address of its lane quadrant and accumulator stage.
// The RESIDUAL / SILU_MUL operand is fetched one 32-column chunk ahead so its latency hides behind the TMEM load.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const LinProblem& pr, bool is_f16, uint32_t taddr, int row, int m_end, int n0) {
  const bool row_ok = row < m_end;
  const int epi = pr.epilogue;
  const bool has_aux = (epi == MC_LINEAR_EPI_RESIDUAL || epi == MC_LINEAR_EPI_SILU_MUL) && row_ok;
  const int rg = (epi == MC_LINEAR_EPI_ROWMASK && row_ok) ? (int)pr.row_group[row] : -1;
  const int orow = (pr.c_rowmap && row_ok) ? pr.c_rowmap[row] : row;  // where this row lands in C (and its token index)
  char* crow = reinterpret_cast<char*>(pr.C) + (long long)orow * pr.ldc * 2;
  const char* rrow = reinterpret_cast<const char*>(pr.residual) + (long long)row * pr.ldr * 2;
  if (epi == MC_LINEAR_EPI_ROPE) {
    // apply_rotary_pos_emb fused behind the q / k projections (multimodal_llama.py:281-282): the projection output is
    // rounded to the storage dtype first, then q*cos + rotate_half(q)*sin with every product and the sum rounded, exactly
    // the op sequence of rope_kernel / the eager reference.  A tile holds BN / head_dim whole heads.
    const int D = pr.rope_head_dim, half = D >> 1;
    const int pos = (pr.rope_pos ? *pr.rope_pos : 0) + orow % pr.rope_seq_len;
    const char* cosr = reinterpret_cast<const char*>(pr.rope_cos) + (long long)pos * D * 2;
    const char* sinr = reinterpret_cast<const char*>(pr.rope_sin) + (long long)pos * D * 2;
#pragma unroll 1
    for (int h0 = 0; h0 < BN; h0 += D) {
      if (n0 + h0 >= pr.N) break;  // warp-uniform (N is a multiple of head_dim)
#pragma unroll 1
      for (int c = 0; c < half; c += 32) {
        uint32_t lo[32], hi[32];
        tmem_ld_32x32(taddr + (uint32_t)(h0 + c), lo);
        tmem_ld_32x32(taddr + (uint32_t)(h0 + half + c), hi);
        // cos / sin and the two output halves move as 32-byte vectors (one full sector per thread and instruction: the rows of a warp
        // are 32 different lines either way, 16-byte accesses took two instructions and two L2 requests per sector)
        Vec<32> cv[2], sv[2];
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            cv[j] = *reinterpret_cast<const Vec<32>*>(cosr + (c + 16 * j) * 2);
            sv[j] = *reinterpret_cast<const Vec<32>*>(sinr + (c + 16 * j) * 2);
          }
        }
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            // Packed 16-bit arithmetic reproduces the eager ops bit for bit: a product of two 16-bit floats is exact in fp32, so
            // "fp32 product rounded to the storage dtype" is the single rounding of HMUL2; a sum of two 16-bit floats is exact in
            // fp32 unless the exponents differ by more than 16, where both roundings return the larger operand, so HADD2 equals
            // "fp32 sum rounded" (the _rn intrinsics keep ptxas from contracting product and sum into an FMA, which would skip the
            // product's rounding).  6 packed instructions per element pair instead of ~20 fp32 operations and six conversions.
            Vec<32> o1, o2;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              const int e = 16 * j + 2 * w;
              uint32_t aw, bw;  // the linear's output, rounded to the storage dtype
              if (is_f16) {
                const __half2 ha = __floats2half2_rn(__uint_as_float(lo[e]), __uint_as_float(lo[e + 1]));
                const __half2 hb = __floats2half2_rn(__uint_as_float(hi[e]), __uint_as_float(hi[e + 1]));
                aw = *reinterpret_cast<const uint32_t*>(&ha);
                bw = *reinterpret_cast<const uint32_t*>(&hb);
                const __half2 xa = *reinterpret_cast<const __half2*>(&aw), xb = *reinterpret_cast<const __half2*>(&bw);
                const __half2 c2 = *reinterpret_cast<const __half2*>(&cv[j].w[w]), s2 = *reinterpret_cast<const __half2*>(&sv[j].w[w]);
                const __half2 r1 = __hadd2_rn(__hmul2_rn(xa, c2), __hneg2(__hmul2_rn(xb, s2)));  // x1*cos + (-x2)*sin; _rn: never contracted into an FMA
                const __half2 r2 = __hadd2_rn(__hmul2_rn(xb, c2), __hmul2_rn(xa, s2));           // x2*cos + x1*sin
                o1.w[w] = *reinterpret_cast<const uint32_t*>(&r1);
                o2.w[w] = *reinterpret_cast<const uint32_t*>(&r2);
              } else {
                const __nv_bfloat162 ha = __floats2bfloat162_rn(__uint_as_float(lo[e]), __uint_as_float(lo[e + 1]));
                const __nv_bfloat162 hb = __floats2bfloat162_rn(__uint_as_float(hi[e]), __uint_as_float(hi[e + 1]));
                aw = *reinterpret_cast<const uint32_t*>(&ha);
                bw = *reinterpret_cast<const uint32_t*>(&hb);
                const __nv_bfloat162 xa = *reinterpret_cast<const __nv_bfloat162*>(&aw), xb = *reinterpret_cast<const __nv_bfloat162*>(&bw);
                const __nv_bfloat162 c2 = *reinterpret_cast<const __nv_bfloat162*>(&cv[j].w[w]), s2 = *reinterpret_cast<const __nv_bfloat162*>(&sv[j].w[w]);
                const __nv_bfloat162 r1 = __hadd2_rn(__hmul2_rn(xa, c2), __hneg2(__hmul2_rn(xb, s2)));
                const __nv_bfloat162 r2 = __hadd2_rn(__hmul2_rn(xb, c2), __hmul2_rn(xa, s2));
                o1.w[w] = *reinterpret_cast<const uint32_t*>(&r1);
                o2.w[w] = *reinterpret_cast<const uint32_t*>(&r2);
              }
            }
            const int col = n0 + h0 + c + 16 * j;
            *reinterpret_cast<Vec<32>*>(crow + (long long)col * 2) = o1;
            *reinterpret_cast<Vec<32>*>(crow + (long long)(col + half) * 2) = o2;
          }
        }
      }
    }
    return;
  }
  uint4 aux[4], aux_next[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    aux[j] = make_uint4(0u, 0u, 0u, 0u);
    aux_next[j] = make_uint4(0u, 0u, 0u, 0u);
    const int col = n0 + 8 * j;
    if (has_aux && col < pr.N) aux[j] = *reinterpret_cast<const uint4*>(rrow + (long long)col * 2);
  }
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    if (n0 + c >= pr.N) break;  // warp-uniform
    uint32_t v[32];
    tmem_ld_32x32(taddr + (uint32_t)c, v);
    if (has_aux) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + c + 32 + 8 * j;
        if (c + 32 < BN && col < pr.N) aux_next[j] = *reinterpret_cast<const uint4*>(rrow + (long long)col * 2);
      }
    }
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + c + 8 * j;
        if (col < pr.N) {  // N % 8 == 0: the whole 8-column vector is in range
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]);
          if (epi == MC_LINEAR_EPI_BIAS || epi == MC_LINEAR_EPI_BIAS_GELU) {
            float b[8];
            unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(pr.bias) + (long long)col * 2), is_f16, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += b[e];
            if (epi == MC_LINEAR_EPI_BIAS_GELU) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = gelu_erf(f[e]);
            }
          } else if (epi == MC_LINEAR_EPI_ROWMASK) {
            const uint2 cg = *reinterpret_cast<const uint2*>(pr.col_group + col);
            const float4 s0 = *reinterpret_cast<const float4*>(pr.col_scale + col);
            const float4 s1 = *reinterpret_cast<const float4*>(pr.col_scale + col + 4);
            const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int g = (int)(((e < 4 ? cg.x : cg.y) >> (8 * (e & 3))) & 0xffu);
              f[e] = (g == rg) ? f[e] * s[e] : 0.0f;
            }
          } else if (epi == MC_LINEAR_EPI_RESIDUAL) {
            float r[8];
            unpack8(aux[j], is_f16, r);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += r[e];
          } else if (epi == MC_LINEAR_EPI_SILU_MUL) {
            // C = silu(aux) * acc with the reference's rounding points (multimodal_llama.py:381-388): aux is the stored
            // gate_proj output; silu(gate) and up_proj are each rounded to the storage dtype before the product
            float g[8], sg[8], up[8];
            unpack8(aux[j], is_f16, g);
#pragma unroll
            for (int e = 0; e < 8; ++e) sg[e] = __fdividef(g[e], 1.0f + __expf(-g[e]));
            const uint4 sgr = pack8(sg, is_f16), upr = pack8(f, is_f16);
            unpack8(sgr, is_f16, sg);
            unpack8(upr, is_f16, up);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = sg[e] * up[e];
          }
          *reinterpret_cast<uint4*>(crow + (long long)col * 2) = pack8(f, is_f16);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) aux[j] = aux_next[j];
  }
}

constexpr int kEpiStageBytes = 32 * 32 * 4;  // per epilogue warp: one 32-row x 32-column fp32 chunk

// NONE / RESIDUAL / SILU_MUL epilogues with COALESCED global accesses.  In epilogue_tile a thread owns an output row, so every
// warp-level 16-byte load / store touches 32 different rows (32 lines): at 512 x 256 pair tiles the address-divergent accesses
// of the eight epilogue warps, not the TMEM reads or the arithmetic, are what the non-overlapped epilogue costs (~17 k cycles per
// tile for a plain store, ~38 k with the SiLU(gate) operand, against 65 k cycles of MMAs at K = 4096).  Here every 32 x 32 fp32
// chunk is transposed through 4 KB of shared memory per warp (16-byte units XOR-swizzled by row: conflict-free both ways): a
// pass then covers 8 rows x 64 contiguous bytes, i.e. 8 lines per instruction instead of 32, for the operand loads (issued one
// chunk ahead) and the stores.  Arithmetic and rounding points are those of epilogue_tile (bit-identical results).
template <int BN>
__device__ __forceinline__ void epilogue_tile_staged(const LinProblem& pr, bool is_f16, uint32_t taddr, int row0, int lane, int m_end,
                                                     int n0, float* stage) {
  const int epi = pr.epilogue;
  const bool has_aux = epi == MC_LINEAR_EPI_RESIDUAL || epi == MC_LINEAR_EPI_SILU_MUL;
  const int piece = lane & 3;
  char* cptr[4];
  const char* rptr[4];
  bool ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + 8 * i + (lane >> 2);
    ok[i] = r < m_end;
    const int orow = (pr.c_rowmap && ok[i]) ? pr.c_rowmap[r] : r;
    cptr[i] = reinterpret_cast<char*>(pr.C) + ((long long)orow * pr.ldc + n0 + piece * 8) * 2;
    rptr[i] = reinterpret_cast<const char*>(pr.residual) + ((long long)r * pr.ldr + n0 + piece * 8) * 2;
  }
  uint4 aux[4], aux_next[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    aux[i] = make_uint4(0u, 0u, 0u, 0u);
    aux_next[i] = make_uint4(0u, 0u, 0u, 0u);
    if (has_aux && ok[i] && n0 + piece * 8 < pr.N) aux[i] = *reinterpret_cast<const uint4*>(rptr[i]);
  }
  float* my_row = stage + lane * 32;
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    if (n0 + c >= pr.N) break;  // warp-uniform
    uint32_t v[32];
    tmem_ld_32x32(taddr + (uint32_t)c, v);
    if (has_aux) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c + 32 < BN && ok[i] && n0 + c + 32 + piece * 8 < pr.N) aux_next[i] = *reinterpret_cast<const uint4*>(rptr[i] + (c + 32) * 2);
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(my_row + ((j ^ (lane & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = 8 * i + (lane >> 2);
      const float4 lo = *reinterpret_cast<const float4*>(stage + rr * 32 + (((2 * piece) ^ (rr & 7)) << 2));
      const float4 hi = *reinterpret_cast<const float4*>(stage + rr * 32 + (((2 * piece + 1) ^ (rr & 7)) << 2));
      if (ok[i] && n0 + c + piece * 8 < pr.N) {
        float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        if (epi == MC_LINEAR_EPI_RESIDUAL) {
          float r[8];
          unpack8(aux[i], is_f16, r);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += r[e];
        } else if (epi == MC_LINEAR_EPI_SILU_MUL) {
          float g[8], sg[8], up[8];
          unpack8(aux[i], is_f16, g);
#pragma unroll
          for (int e = 0; e < 8; ++e) sg[e] = __fdividef(g[e], 1.0f + __expf(-g[e]));
          const uint4 sgr = pack8(sg, is_f16), upr = pack8(f, is_f16);
          unpack8(sgr, is_f16, sg);
          unpack8(upr, is_f16, up);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = sg[e] * up[e];
        }
        *reinterpret_cast<uint4*>(cptr[i] + c * 2) = pack8(f, is_f16);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) aux[i] = aux_next[i];
  }
}

// (A coalesced form of the ROPE epilogue — both halves of a rotation pair through the transpose buffer, cos / sin read coalesced —
// was tried in round 2: bit-identical, but no gain on the q/k/v launch inside the step and, through register pressure in the
// CTA-pair kernel (168 registers, spills), a loss on its other epilogues; profiles/r02_epilogue_ab.txt.  ROPE stays row-per-thread.)
__device__ __forceinline__ bool epilogue_is_staged(const LinParams& P, const LinProblem& pr) {
  return P.epi_staged && (pr.epilogue == MC_LINEAR_EPI_NONE || pr.epilogue == MC_LINEAR_EPI_RESIDUAL || pr.epilogue == MC_LINEAR_EPI_SILU_MUL);
}

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = kBM * kBK * 2;
  static constexpr int B_BYTES = BN * kBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFFSET = (BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 127) & ~127;  // 4 epilogue warps x 4 KB transpose buffers
  static constexpr int TOTAL = EPI_OFFSET + 4 * kEpiStageBytes;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for the manual 1024-byte alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1) linear_kernel(const __grid_constant__ LinParams P) {
  using L = SmemLayout<BN, STAGES>;
  constexpr int TMEM_COLS = 2 * BN;  // power of two: 256 or 512
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.n_prob; ++p) {
      tma_prefetch_desc(&P.prob[p].tmA0);
      tma_prefetch_desc(&P.prob[p].tmB0);
      if (P.prob[p].nkb1) {
        tma_prefetch_desc(&P.prob[p].tmA1);
        tma_prefetch_desc(&P.prob[p].tmB1);
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  __shared__ int s_own[kMaxOwnTiles];
  __shared__ unsigned int s_mm[kMaxMaskTiles];
  __shared__ int s_n_own;
  __shared__ SegTable s_seg;
  if (P.n_seg && threadIdx.x == 96) seg_table_fill(P, &s_seg);
  if (P.compact) {
    // Routed-N launch (LoRA down-projection): most (M tile, N tile) pairs are skipped, and a static round-robin over
    // the full grid of pairs leaves the survivors unevenly spread.  Every CTA compacts the surviving pairs (identical
    // order everywhere, M-major so concurrent tiles share A rows), interleaves the problems of the launch, and keeps
    // every gridDim.x-th entry.  All problems of a compact launch share M, the N tiling and the masks (plan-checked).
    const LinProblem& p0 = P.prob[0];
    for (int i = threadIdx.x; i < p0.tiles_m; i += kThreads) s_mm[i] = p0.mtile_mask[i];
    __syncthreads();
    if (warp == 3) {
      const int total = p0.tiles_m * p0.tiles_n;
      int count = 0, mine = 0;
      for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        const int mt = idx / p0.tiles_n, nt = idx - mt * p0.tiles_n;
        const bool active = idx < total && (p0.ntile_mask[nt] & s_mm[mt]) != 0u;
        const unsigned int b = __ballot_sync(0xffffffffu, active);
        const int j = count + __popc(b & ((1u << lane) - 1u));
        for (int p = 0; p < P.n_prob; ++p) {
          const bool own = active && ((j * P.n_prob + p) % (int)gridDim.x) == (int)blockIdx.x;
          const unsigned int b2 = __ballot_sync(0xffffffffu, own);
          if (own) {
            const int slot = mine + __popc(b2 & ((1u << lane) - 1u));
            if (slot < kMaxOwnTiles) s_own[slot] = (p << 28) | (mt << 10) | nt;
          }
          mine += __popc(b2);
        }
        count += __popc(b);
      }
      if (lane == 0) s_n_own = min(mine, kMaxOwnTiles);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_own = P.compact ? s_n_own : 0;

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the schedule (converged), one elected lane issues the copies =====
    {
      int stage = 0;
      uint32_t phase = 0;
      TileIter tiles(P, s_own, n_own, s_mm, &s_seg);
      Tile t;
      while (tiles.next(t)) {
        const LinProblem& pr = P.prob[t.p];
        const int m0 = t.m0, n0 = t.nt * BN;
        const CUtensorMap* tmB0 = t.seg >= 0 ? &pr.tmB_seg[t.seg] : &pr.tmB0;
        for (int kb = 0; kb < pr.nkb0 + pr.nkb1; ++kb) {
          const bool ext = kb >= pr.nkb0;
          const int k = ext ? kb - pr.nkb0 : kb;
          if (ext && pr.kb1_mask && (pr.kb1_mask[k] & t.gmask) == 0u) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            tma_load_2d(ext ? &pr.tmA1 : &pr.tmA0, &full_bar[stage], sa, k * kBK, m0);
            tma_load_2d(ext ? &pr.tmB1 : tmB0, &full_bar[stage], sa + L::A_BYTES, k * kBK, n0);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp converged, one elected lane issues (under a plain `lane == 0` branch ptxas wraps every
    // tcgen05.mma in an election loop of ~10 instructions, which at BN = 128 costs as much as the 64-cycle MMA itself) =====
    {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      TileIter tiles(P, s_own, n_own, s_mm, &s_seg);
      Tile t;
      while (tiles.next(t)) {
        const LinProblem& pr = P.prob[t.p];
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < pr.nkb0 + pr.nkb1; ++kb) {
          const bool ext = kb >= pr.nkb0;
          if (ext && pr.kb1_mask && (pr.kb1_mask[kb - pr.nkb0] & t.gmask) == 0u) continue;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t a_desc = umma_smem_desc(sa), b_desc = umma_smem_desc(sa + L::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // +32 bytes (16 elements) along K inside the 128-byte swizzle row: start-address field += 2
              umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), P.idesc, (accumulate | (uint32_t)k) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const bool is_f16 = P.is_f16 != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    TileIter tiles(P, s_own, n_own, s_mm, &s_seg);
    Tile t;
    while (tiles.next(t)) {
      const LinProblem& pr = P.prob[t.p];
      const int row = t.m0 + q * 32 + lane;
      const int n0 = t.nt * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if (epilogue_is_staged(P, pr))
        epilogue_tile_staged<BN>(pr, is_f16, taddr, t.m0 + q * 32, lane, t.m_end, n0, reinterpret_cast<float*>(smem + L::EPI_OFFSET + q * kEpiStageBytes));
      else
        epilogue_tile<BN>(pr, is_f16, taddr, row, t.m_end, n0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}


// =====================================================================================================================
// 2-CTA variant (cta_group::2): a CTA pair (thread-block cluster of 2 = the two SMs of a TPC) computes a 512 x 256 tile.
// The leader CTA issues tcgen05.mma.cta_group::2 (M = 256: 128 rows from each CTA), which reads both CTAs' shared memory
// (each CTA holds its own A rows and HALF of the B tile) and writes both CTAs' TMEM; each CTA drains its own accumulator
// rows.  Barriers: TMA bytes of both CTAs complete on the leader's `full` barrier, tcgen05.commit multicasts to both
// CTAs' `empty` / `tfull` barriers, both epilogues arrive on the leader's `tempty`.  For un-routed-N launches.
// =====================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (.release.cta) semantics: the explicit .release.cluster form costs a GPU-scope MEMBAR per k-block, and nothing
  // this thread wrote needs to be visible to the peer (the operands travel through the async proxy / tcgen05 fences)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(COLS) : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the barrier at the same smem offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  const unsigned short mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

constexpr int kThreads2 = 384;  // 4 control warps + 8 epilogue warps (two per TMEM lane quadrant, one per accumulator half)

template <int STAGES>
struct SmemLayout2 {
  static constexpr int A_HALF = kBM * kBK * 2;     // 128 rows x 64
  static constexpr int A_BYTES = 2 * A_HALF;       // this CTA's 256 rows of the 512-row pair tile
  static constexpr int B_BYTES = 128 * kBK * 2;    // this CTA's 128 rows (N) of the 256-column B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFFSET = (BAR_OFFSET + (2 * STAGES + 2) * 8 + 16 + 127) & ~127;  // 8 epilogue warps x 4 KB transpose buffers
  static constexpr int TOTAL = EPI_OFFSET + 8 * kEpiStageBytes;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

// groups present in the rows the pair's MMA for accumulator half h touches: 128 rows of the leader and 128 of the peer
__device__ __forceinline__ unsigned int pair_half_mask(const LinProblem& pr, int mt, int h) {
  if (!pr.mtile_mask) return 0xffffffffu;
  const int n128 = (pr.M + kBM - 1) / kBM;
  const int t0 = mt * 4 + h, t1 = mt * 4 + 2 + h;
  unsigned int m = 0;
  if (t0 < n128) m |= pr.mtile_mask[t0];
  if (t1 < n128) m |= pr.mtile_mask[t1];
  return m;
}

// Pair tile 512 x 256: each CTA owns 256 rows (two 128-row accumulator halves, 2 x 256 TMEM columns = all of TMEM) and
// stages A[256 x 64] + B[128 x 64] = 48 KB per k-block; per k-block the leader issues 2 x 4 tcgen05.mma.cta_group::2
// (M 256 = 128 rows of each CTA, N 256, K 16).  Shared-memory traffic per SM drops to ~94 B/cycle at full tensor rate
// (single-CTA 128x256 tiles need ~190 B/cycle), which is what lifts the tensor pipe above the ~80 % the other tilings
// plateau at.  The accumulator is single-buffered (TMEM is full): the epilogue of a tile is not overlapped with MMAs.
template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1) linear2_kernel(const __grid_constant__ LinParams P) {
  using L = SmemLayout2<STAGES>;
  constexpr int BN = 256, TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.n_prob; ++p) {
      tma_prefetch_desc(&P.prob[p].tmA0);
      tma_prefetch_desc(&P.prob[p].tmB0);
      if (P.prob[p].nkb1) {
        tma_prefetch_desc(&P.prob[p].tmA1);
        tma_prefetch_desc(&P.prob[p].tmB1);
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer's remote arrive (the leader's copy is the one used)
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit from the leader
    }
    mbar_init(tfull_bar, 1);     // multicast tcgen05.commit from the leader
    mbar_init(tempty_bar, 16);   // 8 epilogue warps x 2 CTAs arrive on the leader's copy
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg2<TMEM_COLS>(tmem_slot);
  __shared__ SegTable s_seg;
  if (P.n_seg && threadIdx.x == 96) seg_table_fill(P, &s_seg);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_tiles = total_tiles(P, &s_seg);

  // the pair walks tiles pair, pair + n_pairs, ...; both CTAs decode identically
  if (warp == 0) {
    // whole warp converged, one elected lane issues (see linear_kernel)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const Tile t = decode_tile(P, tile, &s_seg);
        const LinProblem& pr = P.prob[t.p];
        const unsigned int gmask = pair_half_mask(pr, t.mt, 0) | pair_half_mask(pr, t.mt, 1);
        const int m0 = t.m0 + (int)rank * 256, n0 = t.nt * BN + (int)rank * 128;
        const CUtensorMap* tmB0 = t.seg >= 0 ? &pr.tmB_seg[t.seg] : &pr.tmB0;
        for (int kb = 0; kb < pr.nkb0 + pr.nkb1; ++kb) {
          const bool ext = kb >= pr.nkb0;
          const int k = ext ? kb - pr.nkb0 : kb;
          if (ext && pr.kb1_mask && (pr.kb1_mask[k] & gmask) == 0u) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);  // both CTAs' bytes land on this barrier
            else mbar_arrive_cluster(full_leader);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            tma_load_2d_cg2(ext ? &pr.tmA1 : &pr.tmA0, full_leader, sa, k * kBK, m0);
            tma_load_2d_cg2(ext ? &pr.tmB1 : tmB0, full_leader, sa + L::A_BYTES, k * kBK, n0);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0, t_phase = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const Tile t = decode_tile(P, tile, &s_seg);
        const LinProblem& pr = P.prob[t.p];
        const unsigned int hmask[2] = {pair_half_mask(pr, t.mt, 0), pair_half_mask(pr, t.mt, 1)};
        mbar_wait(tempty_bar, t_phase ^ 1u);  // both CTAs' epilogues have drained the previous tile
        tc_fence_after();
        uint32_t accumulate[2] = {0u, 0u};
        for (int kb = 0; kb < pr.nkb0 + pr.nkb1; ++kb) {
          const bool ext = kb >= pr.nkb0;
          const unsigned int kbm = (ext && pr.kb1_mask) ? pr.kb1_mask[kb - pr.nkb0] : 0xffffffffu;
          if ((kbm & (hmask[0] | hmask[1])) == 0u) continue;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t b_desc = umma_smem_desc(sa + L::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if ((kbm & hmask[h]) == 0u) continue;  // this half's rows carry none of the k-block's adapter group
              const uint64_t a_desc = umma_smem_desc(sa + h * L::A_HALF);
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                umma_f16_cg2(tmem_base + (uint32_t)(h * BN), a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), P.idesc,
                             (accumulate[h] | (uint32_t)k) ? 1u : 0u);
            }
            umma_commit_cg2(&empty_bar[stage]);
          }
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h)
            if ((kbm & hmask[h]) != 0u) accumulate[h] = 1;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma_commit_cg2(tfull_bar);
        __syncwarp();
        t_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, h = (warp - 4) >> 2;
    const bool is_f16 = P.is_f16 != 0;
    uint32_t t_phase = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const Tile t = decode_tile(P, tile, &s_seg);
      const LinProblem& pr = P.prob[t.p];
      const int row = t.m0 + (int)rank * 256 + h * 128 + q * 32 + lane;
      mbar_wait(tfull_bar, t_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * BN);
      if (epilogue_is_staged(P, pr))
        epilogue_tile_staged<BN>(pr, is_f16, taddr, row - lane, lane, t.m_end, t.nt * BN,
                                 reinterpret_cast<float*>(smem + L::EPI_OFFSET + (warp - 4) * kEpiStageBytes));
      else
        epilogue_tile<BN>(pr, is_f16, taddr, row, t.m_end, t.nt * BN);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(tempty_bar);
        else mbar_arrive_cluster(mapa_u32(smem_u32(tempty_bar), 0));
      }
      t_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer can still signal its barriers or read its memory
  tc_fence_after();
  if (warp == 2) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
}

// =====================================================================================================================
// CTA-pair variant with overlapped epilogue: a pair computes a 256 x 256 tile (cta_group::2, M = 256: 128 rows per CTA).
// Each CTA stages its own A rows [128 x 64] and HALF of the B tile [128 x 64] = 32 KB per k-block (6-stage ring), so the
// tensor cores read 2/3 of the shared-memory bytes per MMA of the single-CTA kernel and B crosses L2 -> SM once per
// pair instead of once per CTA.  One accumulator is 256 TMEM columns: the 512 columns hold TWO, and the epilogue of
// tile i (4 warps per CTA) overlaps the MMAs of tile i + 1 exactly as in linear_kernel.  LoRA k-blocks are skipped on the
// union of the groups of the pair's two 128-row tiles (pure tiles in the modality-major row order).
// =====================================================================================================================
template <int STAGES>
struct SmemLayout3 {
  static constexpr int A_BYTES = kBM * kBK * 2;   // this CTA's 128 rows
  static constexpr int B_BYTES = 128 * kBK * 2;   // this CTA's 128 (of 256) B rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int EPI_OFFSET = (BAR_OFFSET + (2 * STAGES + 4) * 8 + 16 + 127) & ~127;
  static constexpr int TOTAL = EPI_OFFSET + 4 * kEpiStageBytes;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

__device__ __forceinline__ unsigned int pair256_mask(const LinProblem& pr, int mt) {
  if (!pr.mtile_mask) return 0xffffffffu;
  const int n128 = (pr.M + kBM - 1) / kBM;
  unsigned int m = pr.mtile_mask[2 * mt];
  if (2 * mt + 1 < n128) m |= pr.mtile_mask[2 * mt + 1];
  return m;
}

template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) linear3_kernel(const __grid_constant__ LinParams P) {
  using L = SmemLayout3<STAGES>;
  constexpr int BN = 256, TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.n_prob; ++p) {
      tma_prefetch_desc(&P.prob[p].tmA0);
      tma_prefetch_desc(&P.prob[p].tmB0);
      if (P.prob[p].nkb1) {
        tma_prefetch_desc(&P.prob[p].tmA1);
        tma_prefetch_desc(&P.prob[p].tmB1);
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer's remote arrive (the leader's copy is the one used)
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit from the leader
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);   // multicast tcgen05.commit from the leader
      mbar_init(&tempty_bar[a], 8);  // 4 epilogue warps x 2 CTAs arrive on the leader's copy
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_cg2<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one elected lane per CTA): own A rows + own half of B, bytes complete on the leader's barrier =====
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < P.total_tiles; tile += n_pairs) {
        const Tile t = decode_tile(P, tile, nullptr);
        const LinProblem& pr = P.prob[t.p];
        const unsigned int gmask = pair256_mask(pr, t.mt);
        const int m0 = t.mt * 256 + (int)rank * 128, n0 = t.nt * BN + (int)rank * 128;
        for (int kb = 0; kb < pr.nkb0 + pr.nkb1; ++kb) {
          const bool ext = kb >= pr.nkb0;
          const int k = ext ? kb - pr.nkb0 : kb;
          if (ext && pr.kb1_mask && (pr.kb1_mask[k] & gmask) == 0u) continue;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
            else mbar_arrive_cluster(full_leader);
            uint8_t* sa = smem + stage * L::STAGE_BYTES;
            tma_load_2d_cg2(ext ? &pr.tmA1 : &pr.tmA0, full_leader, sa, k * kBK, m0);
            tma_load_2d_cg2(ext ? &pr.tmB1 : &pr.tmB0, full_leader, sa + L::A_BYTES, k * kBK, n0);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane of the leader CTA) =====
    if (rank == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = pair; tile < P.total_tiles; tile += n_pairs) {
        const Tile t = decode_tile(P, tile, nullptr);
        const LinProblem& pr = P.prob[t.p];
        const unsigned int gmask = pair256_mask(pr, t.mt);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < pr.nkb0 + pr.nkb1; ++kb) {
          const bool ext = kb >= pr.nkb0;
          if (ext && pr.kb1_mask && (pr.kb1_mask[kb - pr.nkb0] & gmask) == 0u) continue;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t a_desc = umma_smem_desc(sa), b_desc = umma_smem_desc(sa + L::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_f16_cg2(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), P.idesc, (accumulate | (uint32_t)k) ? 1u : 0u);
            umma_commit_cg2(&empty_bar[stage]);  // frees the stage in both CTAs once these MMAs have read it
          }
          __syncwarp();
          accumulate = 1;
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) umma_commit_cg2(&tfull_bar[acc]);  // accumulator complete -> both CTAs' epilogues
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: each CTA drains its own 128 accumulator rows =====
    const int q = warp & 3;
    const bool is_f16 = P.is_f16 != 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < P.total_tiles; tile += n_pairs) {
      const Tile t = decode_tile(P, tile, nullptr);
      const LinProblem& pr = P.prob[t.p];
      const int row = t.mt * 256 + (int)rank * 128 + q * 32 + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if (epilogue_is_staged(P, pr))
        epilogue_tile_staged<BN>(pr, is_f16, taddr, row - lane, lane, pr.M, t.nt * BN,
                                 reinterpret_cast<float*>(smem + L::EPI_OFFSET + (warp & 3) * kEpiStageBytes));
      else
        epilogue_tile<BN>(pr, is_f16, taddr, row, pr.M, t.nt * BN);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer can still signal its barriers or read its memory
  tc_fence_after();
  if (warp == 2) tmem_dealloc_cg2<TMEM_COLS>(tmem_base);
}

// ---- routing helpers ------------------------------------------------------------------------------------
// mask[t] = OR over the rows of 128-row tile t of (1 << group[row])
// coarsen = 2 / 4: every tile of an aligned run of 2 / 4 tiles gets the union of the run (the CTA-pair kernels skip LoRA k-blocks per
// 256 / 512 rows)
__global__ void route_tile_mask_kernel(const unsigned char* __restrict__ group, int M, unsigned int* __restrict__ mask, int n_tiles, int coarsen) {
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  unsigned int m = 0;
  const int first = tile / coarsen * coarsen;
  for (int r = first * kBM + lane; r < min(M, (first + coarsen) * kBM); r += 32) m |= 1u << (group[r] & 31);
#pragma unroll
  for (int d = 16; d; d >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, d);
  if (lane == 0) mask[tile] = m;
}

// Modality-major row order of a batch: a STABLE counting sort of the T rows by routing group, one CTA.  Every thread owns a
// contiguous run of rows (so the order inside a group is the sequence order), counts its rows per group, the per-group counts are
// scanned over the threads, and a second walk over the run places the rows.  Replaces a library radix sort + bincount + cumsum +
// two index kernels per batch.  lut (in parameter space) maps the splice's modality ids to routing groups.
constexpr int kRouteThreads = 1024;
constexpr int kRouteMaxGroups = MC_LINEAR_MAX_SEGMENTS;
struct RouteLut {
  unsigned char g[16];
};
__global__ void __launch_bounds__(kRouteThreads) route_permutation_kernel(const unsigned char* __restrict__ modal_id, int T, RouteLut lut, int n_groups,
                                                                          int* __restrict__ perm, int* __restrict__ inv_perm,
                                                                          unsigned char* __restrict__ row_group, int* __restrict__ seg_start,
                                                                          unsigned char* __restrict__ group_seq) {
  __shared__ int warp_tot[kRouteMaxGroups][32];
  __shared__ int group_base[kRouteMaxGroups + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (T + kRouteThreads - 1) / kRouteThreads;
  const int t0 = min(tid * per, T), t1 = min(t0 + per, T);
  int cnt[kRouteMaxGroups];
#pragma unroll
  for (int g = 0; g < kRouteMaxGroups; ++g) cnt[g] = 0;
  for (int t = t0; t < t1; ++t) {
    const int g = lut.g[modal_id[t] & 15];
#pragma unroll
    for (int k = 0; k < kRouteMaxGroups; ++k) cnt[k] += (g == k);
  }
  int start[kRouteMaxGroups];  // exclusive prefix of this thread's count inside the group
#pragma unroll
  for (int g = 0; g < kRouteMaxGroups; ++g) {
    int v = cnt[g];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    if (lane == 31) warp_tot[g][warp] = v;
    start[g] = v - cnt[g];
  }
  __syncthreads();
  if (warp < kRouteMaxGroups) {  // warp g scans the 32 warp totals of group g
    int v = warp_tot[warp][lane];
    const int own = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += o;
    }
    warp_tot[warp][lane] = v - own;
    if (lane == 31) group_base[warp + 1] = v;  // total of the group
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int g = 0; g < kRouteMaxGroups; ++g) {
      const int n = group_base[g + 1];
      group_base[g] = acc;
      acc += n;
    }
    group_base[kRouteMaxGroups] = acc;
    for (int g = 0; g <= n_groups; ++g) seg_start[g] = group_base[min(g, kRouteMaxGroups)];
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < kRouteMaxGroups; ++g) start[g] += group_base[g] + warp_tot[g][warp];
  for (int t = t0; t < t1; ++t) {
    const int g = lut.g[modal_id[t] & 15];
    int pos = 0;
#pragma unroll
    for (int k = 0; k < kRouteMaxGroups; ++k) {
      if (g == k) pos = start[k]++;
    }
    perm[pos] = t;
    inv_perm[t] = pos;
    row_group[pos] = (unsigned char)g;
    if (group_seq) group_seq[t] = (unsigned char)g;
  }
}

// out[i] = silu(gate[i]) * up[i]   (multimodal_llama.py:381-388: act_fn(gate_proj(x)) * up_proj(x)); rounding points
// as the reference: silu result rounded to the storage dtype, then the product rounded.
template <typename T>
__global__ void __launch_bounds__(256) silu_mul_kernel(const T* __restrict__ gate, const T* __restrict__ up, T* __restrict__ out,
                                                       long long rows, int cols, long long ldg, long long ldu, long long ldo) {
  const long long vec_per_row = cols / 8;
  const long long total = rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row, c = (i % vec_per_row) * 8;
    const uint4 g = *reinterpret_cast<const uint4*>(gate + r * ldg + c);
    const uint4 u = *reinterpret_cast<const uint4*>(up + r * ldu + c);
    const T* ge = reinterpret_cast<const T*>(&g);
    const T* ue = reinterpret_cast<const T*>(&u);
    uint4 o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = to_f32<T>(ge[e]);
      const T s = from_f32<T>(__fdividef(x, 1.0f + __expf(-x)));
      oe[e] = from_f32<T>(to_f32<T>(s) * to_f32<T>(ue[e]));
    }
    *reinterpret_cast<uint4*>(out + r * ldo + c) = o;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode_fn() {
  static std::mutex mu;
  static encode_tiled_fn fn = nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}

// 2-D row-major [rows, cols] 16-bit tensor, box = box_rows x 64 columns, 128B swizzle, zero fill out of bounds
int encode_operand(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows, int dtype) {
  encode_tiled_fn fn = get_encode_fn();
  if (!fn) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dtype == MC_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MC_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) for [%lld x %lld] ld %lld", (int)r, rows, cols, ld);
  return MC_OK;
}

}  // namespace mc

using namespace mc;

struct mc_linear_plan {
  LinParams params;
  int bn, grid, dtype;
  int two_cta;  // 1: linear2_kernel (512 x 256 pair tiles); 2: linear3_kernel (256 x 256 pair tiles, overlapped epilogue)
  size_t smem_bytes;
  double flops;
  std::vector<void*> owned;  // device arrays built by the plan (group tables)
};

template <int BN, int STAGES>
static cudaError_t launch_linear(const mc_linear_plan* p, cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(linear_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  linear_kernel<BN, STAGES><<<p->grid, kThreads, L::DYN_BYTES, stream>>>(p->params);
  return cudaGetLastError();
}

template <int STAGES>
static cudaError_t launch_linear2(const mc_linear_plan* p, cudaStream_t stream) {
  using L = SmemLayout2<STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(linear2_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  // the kernel carries __cluster_dims__(2, 1, 1); the grid is even
  linear2_kernel<STAGES><<<p->grid, kThreads2, L::DYN_BYTES, stream>>>(p->params);
  return cudaGetLastError();
}

template <int STAGES>
static cudaError_t launch_linear3(const mc_linear_plan* p, cudaStream_t stream) {
  using L = SmemLayout3<STAGES>;
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(linear3_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  linear3_kernel<STAGES><<<p->grid, kThreads, L::DYN_BYTES, stream>>>(p->params);
  return cudaGetLastError();
}

static void linear_plan_free(mc_linear_plan* p) {
  if (!p) return;
  for (void* d : p->owned) cudaFree(d);
  delete p;
}

template <typename T>
static cudaError_t upload(mc_linear_plan* p, const std::vector<T>& host, const T** dev_out) {
  T* d = nullptr;
  cudaError_t e = cudaMalloc(&d, std::max<size_t>(host.size(), 1) * sizeof(T));
  if (e != cudaSuccess) return e;
  p->owned.push_back(d);
  if (!host.empty()) e = cudaMemcpy(d, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
  *dev_out = d;
  return e;
}

extern "C" int mc_linear_plan_create(mc_linear_plan_t** out, const mc_linear_desc_t* desc, int n_problems, int dtype, int tuning) {
  MC_REQUIRE(out != nullptr, "plan out-pointer is NULL");
  *out = nullptr;
  MC_REQUIRE(desc != nullptr && n_problems >= 1 && n_problems <= kMaxProb, "n_problems %d outside [1, %d]", n_problems, kMaxProb);
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "linear: dtype must be bf16 or fp16");
  // tuning bits 0-7: 0 = default (BN 256 when every problem has N >= 256, else 128); 1 = force BN 128; 2 = force BN 256
  // bits 8-15: rasterisation group override; bit 16: disable the compact schedule of routed-N launches
  int bn = 256;
  for (int i = 0; i < n_problems; ++i)
    if (desc[i].N < 256) bn = 128;
  if ((tuning & 0xff) == 1) bn = 128;
  if ((tuning & 0xff) == 2) bn = 256;
  // 3 = CTA-pair kernel (cta_group::2, 512 x 256 pair tiles), 4 = CTA-pair kernel with 256 x 256 pair tiles and an
  // overlapped epilogue; neither takes ROWMASK (routed-N) launches
  const int pair_mode = (tuning & 0xff) == 3 ? 1 : ((tuning & 0xff) == 4 ? 2 : 0);
  const bool two = pair_mode != 0;
  if (two) bn = 256;
  const int bm = pair_mode == 1 ? 512 : (pair_mode == 2 ? 256 : kBM);
  const int a_box = pair_mode == 1 ? 256 : kBM;
  mc_linear_plan* p = new (std::nothrow) mc_linear_plan();
  if (!p) return fail(MC_ERR_NOMEM, "host allocation failed");
  memset(&p->params, 0, sizeof(p->params));
  p->bn = bn;
  p->two_cta = pair_mode;
  p->dtype = dtype;
  p->flops = 0;
  int tile_end = 0;
  int rc = MC_OK;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < n_problems && rc == MC_OK && e == cudaSuccess; ++i) {
    const mc_linear_desc_t& d = desc[i];
    LinProblem& pr = p->params.prob[i];
#define PLAN_REQUIRE(cond, ...)                    \
  if (!(cond)) {                                   \
    rc = fail(MC_ERR_INVALID, __VA_ARGS__);        \
    break;                                         \
  }
    PLAN_REQUIRE(d.M >= 1 && d.N >= 8 && d.K0 >= 8, "problem %d: need M >= 1, N >= 8, K0 >= 8", i);
    PLAN_REQUIRE(d.N % 8 == 0 && d.K0 % 8 == 0 && d.K1 % 8 == 0, "problem %d: N, K0, K1 must be multiples of 8", i);
    PLAN_REQUIRE(d.A0 && d.B0 && d.C, "problem %d: A0 / B0 / C is NULL", i);
    PLAN_REQUIRE(d.lda0 >= d.K0 && d.ldb0 >= d.K0 && d.ldc >= d.N && d.lda0 % 8 == 0 && d.ldb0 % 8 == 0 && d.ldc % 8 == 0,
                 "problem %d: leading dimensions must cover the row and be multiples of 8 elements", i);
    PLAN_REQUIRE((((uintptr_t)d.A0 | (uintptr_t)d.B0 | (uintptr_t)d.C | (uintptr_t)d.A1 | (uintptr_t)d.B1 | (uintptr_t)d.bias |
                   (uintptr_t)d.residual) & 15) == 0, "problem %d: operand pointers must be 16-byte aligned", i);
    PLAN_REQUIRE(d.K1 == 0 || (d.A1 && d.B1 && d.lda1 >= d.K1 && d.ldb1 >= d.K1 && d.lda1 % 8 == 0 && d.ldb1 % 8 == 0),
                 "problem %d: K1 > 0 needs A1 / B1 with valid leading dimensions", i);
    PLAN_REQUIRE(d.epilogue >= MC_LINEAR_EPI_NONE && d.epilogue <= MC_LINEAR_EPI_ROPE, "problem %d: bad epilogue %d", i, d.epilogue);
    PLAN_REQUIRE(d.epilogue != MC_LINEAR_EPI_ROPE ||
                     (d.rope_cos && d.rope_sin && d.rope_seq_len >= 1 && d.rope_head_dim >= 64 && d.rope_head_dim % 64 == 0 &&
                      bn % d.rope_head_dim == 0 && d.N % d.rope_head_dim == 0 &&
                      (((uintptr_t)d.rope_cos | (uintptr_t)d.rope_sin | (uintptr_t)d.C) & 31) == 0 && d.ldc % 16 == 0),
                 "problem %d: ROPE epilogue needs cos/sin tables, seq_len >= 1, a head_dim in {64, 128, 256} dividing N and the tile, and "
                 "32-byte aligned C / cos / sin with ldc %% 16 == 0", i);
    PLAN_REQUIRE((d.epilogue != MC_LINEAR_EPI_BIAS && d.epilogue != MC_LINEAR_EPI_BIAS_GELU) || d.bias, "problem %d: bias is NULL", i);
    PLAN_REQUIRE((d.epilogue != MC_LINEAR_EPI_RESIDUAL && d.epilogue != MC_LINEAR_EPI_SILU_MUL) ||
                     (d.residual && d.ldr >= d.N && d.ldr % 8 == 0), "problem %d: residual / gate operand missing", i);
    const bool routed_n = d.epilogue == MC_LINEAR_EPI_ROWMASK;
    PLAN_REQUIRE(!(two && routed_n), "problem %d: the CTA-pair kernel does not take ROWMASK launches", i);
    const bool routed_k = d.K1 > 0 && d.n_groups > 0;
    PLAN_REQUIRE(d.n_seg >= 0 && d.n_seg <= kMaxSeg, "problem %d: n_seg %d outside [0, %d]", i, d.n_seg, kMaxSeg);
    PLAN_REQUIRE(d.n_seg == desc[0].n_seg && d.seg_start == desc[0].seg_start, "problem %d: every problem of a launch shares one segment table", i);
    PLAN_REQUIRE(d.n_seg == 0 || (d.seg_start && d.B0_seg && d.K1 == 0 && !routed_n && pair_mode != 2),
                 "problem %d: a segmented problem needs seg_start and B0_seg, K1 = 0, no ROWMASK epilogue and not tuning 4", i);
    PLAN_REQUIRE(!routed_n || (d.n_groups >= 1 && d.n_groups <= 32 && d.group_cols && d.row_group && d.col_scale),
                 "problem %d: ROWMASK epilogue needs n_groups in [1,32], group_cols, row_group and col_scale", i);
    PLAN_REQUIRE(!routed_k || (d.n_groups <= 32 && d.group_cols && d.mtile_mask), "problem %d: routed K1 needs group_cols and mtile_mask", i);
    pr.M = d.M;
    pr.N = d.N;
    pr.nkb0 = (d.K0 + kBK - 1) / kBK;
    pr.nkb1 = (d.K1 + kBK - 1) / kBK;
    pr.tiles_m = (d.M + bm - 1) / bm + d.n_seg;  // segmented: upper bound (every segment may end in a partial tile)
    pr.tiles_n = (d.N + bn - 1) / bn;
    p->params.cum_tiles_n[i + 1] = p->params.cum_tiles_n[i] + pr.tiles_n;
    tile_end += pr.tiles_m * pr.tiles_n;
    pr.tile_end = tile_end;
    pr.epilogue = d.epilogue;
    pr.C = d.C;
    pr.ldc = d.ldc;
    pr.bias = d.bias;
    pr.residual = d.residual;
    pr.ldr = d.ldr;
    pr.rope_cos = d.rope_cos;
    pr.rope_sin = d.rope_sin;
    pr.rope_pos = d.rope_pos;
    pr.rope_seq_len = d.rope_seq_len;
    pr.rope_head_dim = d.rope_head_dim;
    pr.c_rowmap = d.c_rowmap;
    pr.col_scale = d.col_scale;
    pr.row_group = d.row_group;
    pr.mtile_mask = d.mtile_mask;
    if (routed_n || routed_k) {
      const int extent = routed_n ? d.N : d.K1;
      PLAN_REQUIRE(d.group_cols[0] == 0 && d.group_cols[d.n_groups] == extent, "problem %d: group_cols must span [0, %d]", i, extent);
      std::vector<unsigned char> cg(extent);
      for (int g = 0; g < d.n_groups; ++g) {
        PLAN_REQUIRE(d.group_cols[g] <= d.group_cols[g + 1], "problem %d: group_cols not ascending", i);
        for (int c = d.group_cols[g]; c < d.group_cols[g + 1]; ++c) cg[c] = (unsigned char)g;
      }
      if (rc != MC_OK) break;
      const int blk = routed_n ? bn : kBK;
      std::vector<unsigned int> bm((extent + blk - 1) / blk, 0u);
      for (int c = 0; c < extent; ++c) bm[c / blk] |= 1u << cg[c];
      if (routed_n) {
        e = upload(p, cg, &pr.col_group);
        if (e == cudaSuccess) e = upload(p, bm, &pr.ntile_mask);
      } else {
        for (size_t b = 0; b < bm.size(); ++b)
          PLAN_REQUIRE((bm[b] & (bm[b] - 1)) == 0, "problem %d: K1 group boundaries must be multiples of %d", i, kBK);
        if (rc != MC_OK) break;
        e = upload(p, bm, &pr.kb1_mask);
      }
      if (e != cudaSuccess) break;
    }
    rc = encode_operand(&pr.tmA0, d.A0, d.M, d.K0, d.lda0, a_box, dtype);
    if (rc == MC_OK) rc = encode_operand(&pr.tmB0, d.B0, d.N, d.K0, d.ldb0, two ? 128 : bn, dtype);
    if (rc == MC_OK && d.n_seg > 0) {
      std::vector<CUtensorMap> maps(d.n_seg);
      for (int g = 0; g < d.n_seg && rc == MC_OK; ++g) {
        if (!d.B0_seg[g] || ((uintptr_t)d.B0_seg[g] & 15)) rc = fail(MC_ERR_INVALID, "problem %d: B0_seg[%d] is NULL or misaligned", i, g);
        else rc = encode_operand(&maps[g], d.B0_seg[g], d.N, d.K0, d.ldb0, two ? 128 : bn, dtype);
      }
      if (rc == MC_OK) {
        CUtensorMap* dm = nullptr;  // cudaMalloc returns 256-byte aligned memory: fine for 64-byte aligned tensor maps
        e = cudaMalloc(&dm, sizeof(CUtensorMap) * d.n_seg);
        if (e == cudaSuccess) {
          p->owned.push_back(dm);
          e = cudaMemcpy(dm, maps.data(), sizeof(CUtensorMap) * d.n_seg, cudaMemcpyHostToDevice);
          pr.tmB_seg = dm;
        }
      }
    }
    if (rc == MC_OK && d.K1 > 0) rc = encode_operand(&pr.tmA1, d.A1, d.M, d.K1, d.lda1, a_box, dtype);
    if (rc == MC_OK && d.K1 > 0) rc = encode_operand(&pr.tmB1, d.B1, d.N, d.K1, d.ldb1, two ? 128 : bn, dtype);
    p->flops += 2.0 * d.M * (double)d.N * (double)(d.K0 + d.K1);
#undef PLAN_REQUIRE
  }
  if (rc == MC_OK && e != cudaSuccess) rc = fail(MC_ERR_CUDA, "linear plan setup failed: %s", cudaGetErrorString(e));
  if (rc != MC_OK) {
    linear_plan_free(p);
    return rc;
  }
  p->params.n_prob = n_problems;
  p->params.total_tiles = tile_end;
  p->params.seg_start = desc[0].seg_start;
  p->params.n_seg = desc[0].n_seg;
  p->params.bm = bm;
  // rasterisation: when B (weights) alone overflows a good part of L2, sweep N over 32 M-tiles at a time so B streams
  // from HBM once per 4096 rows instead of once per 1024; tuning bits 8-15 override
  {
    double b_bytes = 0;
    for (int i = 0; i < n_problems; ++i) b_bytes = std::max(b_bytes, (double)desc[i].N * (desc[i].K0 + desc[i].K1) * 2.0);
    p->params.group_m = b_bytes > 48e6 ? 32 : 8;
    // the group is meant in ROWS (1024 / 4096): the pair kernels' tiles are 512 / 256 rows tall, and 32 of those per group
    // put 360 MB of down_proj activations between two visits of a weight tile
    p->params.group_m = std::max(1, p->params.group_m * kBM / bm);
    if ((tuning >> 8) & 0xff) p->params.group_m = (tuning >> 8) & 0xff;
  }
  // compact schedule for routed-N launches whose problems share M, N tiling and routing
  {
    const LinProblem& p0 = p->params.prob[0];
    bool ok = p0.ntile_mask != nullptr && p0.mtile_mask != nullptr && p0.tiles_m <= kMaxMaskTiles && p0.tiles_n <= 1023 &&
              !((tuning >> 16) & 1);
    for (int i = 1; i < n_problems && ok; ++i) {
      const LinProblem& pi = p->params.prob[i];
      ok = pi.ntile_mask != nullptr && pi.mtile_mask == p0.mtile_mask && pi.M == p0.M && pi.N == p0.N &&
           desc[i].n_groups == desc[0].n_groups &&
           memcmp(desc[i].group_cols, desc[0].group_cols, sizeof(int32_t) * (desc[0].n_groups + 1)) == 0;
    }
    const int sms_ = sm_count();
    if (ok && sms_ > 0 && (long long)n_problems * p0.tiles_m * p0.tiles_n > (long long)kMaxOwnTiles * sms_) ok = false;
    p->params.compact = ok ? 1 : 0;
  }
  p->params.is_f16 = dtype == MC_F16;
  p->params.epi_staged = ((tuning >> 17) & 1) ? 0 : 1;  // tuning bit 17: row-per-thread epilogue everywhere (the round-1 form)
  // instruction descriptor: D = F32 (bits 4-5 = 1), A/B format (bits 7-9, 10-12: 0 = F16, 1 = BF16), both K-major,
  // N >> 3 at bit 17, M >> 4 at bit 24
  const unsigned int fmt = dtype == MC_F16 ? 0u : 1u;
  p->params.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(bn >> 3) << 17) | ((unsigned)((two ? 256 : kBM) >> 4) << 24);
  const int sms = sm_count();
  if (sms <= 0) {
    linear_plan_free(p);
    return fail(MC_ERR_CUDA, "no CUDA device");
  }
  p->grid = p->params.compact ? sms : std::min(sms, tile_end);
  if (two) p->grid = std::min(sms & ~1, 2 * tile_end);
  *out = p;
  return MC_OK;
}

extern "C" int mc_linear_plan_run(const mc_linear_plan_t* p, mc_stream_t stream) {
  MC_REQUIRE(p != nullptr, "plan is NULL");
  cudaError_t e = p->two_cta == 1 ? launch_linear2<4>(p, (cudaStream_t)stream)
                  : p->two_cta == 2 ? launch_linear3<6>(p, (cudaStream_t)stream)
                  : p->bn == 256 ? launch_linear<256, 4>(p, (cudaStream_t)stream)
                                 : launch_linear<128, 6>(p, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(MC_ERR_CUDA, "linear launch failed: %s", cudaGetErrorString(e));
  return MC_OK;
}

extern "C" double mc_linear_plan_flops(const mc_linear_plan_t* p) { return p ? p->flops : 0.0; }

extern "C" int mc_linear_plan_destroy(mc_linear_plan_t* p) {
  linear_plan_free(p);
  return MC_OK;
}

extern "C" int mc_route_tile_masks_coarse(const uint8_t* d_row_group, int M, uint32_t* d_mtile_mask, int coarsen, mc_stream_t stream) {
  MC_REQUIRE(d_row_group && d_mtile_mask && M >= 1, "route masks: NULL pointer or M < 1");
  MC_REQUIRE(coarsen == 1 || coarsen == 2 || coarsen == 4, "route masks: coarsen must be 1, 2 or 4");
  const int n_tiles = (M + kBM - 1) / kBM;
  route_tile_mask_kernel<<<(n_tiles + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_row_group, M, d_mtile_mask, n_tiles, coarsen);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}

extern "C" int mc_route_tile_masks(const uint8_t* d_row_group, int M, uint32_t* d_mtile_mask, mc_stream_t stream) {
  return mc_route_tile_masks_coarse(d_row_group, M, d_mtile_mask, 1, stream);
}

extern "C" int mc_route_permutation(const uint8_t* d_modal_id, int T, const uint8_t* lut, int n_lut, int n_groups, int32_t* d_perm,
                                    int32_t* d_inv_perm, uint8_t* d_row_group, int32_t* d_seg_start, uint8_t* d_group_seq, mc_stream_t stream) {
  MC_REQUIRE(d_modal_id && d_perm && d_inv_perm && d_row_group && d_seg_start && T >= 1, "route permutation: NULL pointer or T < 1");
  MC_REQUIRE(n_groups >= 1 && n_groups <= kRouteMaxGroups && n_lut >= 0 && n_lut <= 16, "route permutation: n_groups outside [1, %d] or n_lut > 16",
             kRouteMaxGroups);
  RouteLut l;
  for (int i = 0; i < 16; ++i) {
    int g = (lut != nullptr && i < n_lut) ? (int)lut[i] : (lut == nullptr ? i : 0);
    MC_REQUIRE(!(lut != nullptr && i < n_lut) || g < n_groups, "route permutation: lut[%d] = %d is not a routing group (< %d)", i, g, n_groups);
    l.g[i] = (unsigned char)(g < n_groups ? g : 0);  // ids the splice never produces for this model
  }
  route_permutation_kernel<<<1, kRouteThreads, 0, (cudaStream_t)stream>>>(d_modal_id, T, l, n_groups, d_perm, d_inv_perm, d_row_group, d_seg_start,
                                                                          d_group_seq);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}

extern "C" int mc_silu_mul(const void* gate, const void* up, void* out, int64_t rows, int cols, int64_t ld_gate, int64_t ld_up,
                           int64_t ld_out, int dtype, mc_stream_t stream) {
  MC_REQUIRE(gate && up && out, "silu_mul: NULL pointer");
  MC_REQUIRE(dtype == MC_BF16 || dtype == MC_F16, "silu_mul: dtype must be bf16 or fp16");
  MC_REQUIRE(rows >= 0 && cols >= 8 && cols % 8 == 0 && ld_gate % 8 == 0 && ld_up % 8 == 0 && ld_out % 8 == 0,
             "silu_mul: cols and leading dimensions must be multiples of 8");
  MC_REQUIRE((((uintptr_t)gate | (uintptr_t)up | (uintptr_t)out) & 15) == 0, "silu_mul: pointers must be 16-byte aligned");
  if (rows == 0) return MC_OK;
  const long long total = rows * (cols / 8);
  const int sms = sm_count();
  MC_REQUIRE(sms > 0, "no CUDA device");
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
  if (dtype == MC_BF16)
    silu_mul_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)gate, (const __nv_bfloat16*)up,
                                                                          (__nv_bfloat16*)out, rows, cols, ld_gate, ld_up, ld_out);
  else
    silu_mul_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)gate, (const __half*)up, (__half*)out, rows, cols,
                                                                    ld_gate, ld_up, ld_out);
  MC_CUDA_OK(cudaGetLastError());
  return MC_OK;
}
