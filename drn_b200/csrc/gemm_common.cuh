// Definitions shared by the tcgen05 contraction kernels (gemm.cu: one tile per CTA; gemm2.cu: persistent CTA pairs).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace drn {

struct PlanesView {
  const __nv_bfloat16* ptr;
  long long plane_stride;
  int B, T, P, C;
};

struct GemmKParams {
  int form, b_mn;
  int B, T, N, K, M;
  int Rm, Bbm, tiles_per_sample;                // ROWS: M-tile = Rm time slots x Bbm samples (=128 rows)
  int Rk, Bbk, kblocks_per_sample, num_kblocks;  // WGRAD: K-block = Rk time slots x Bbk samples (=64 rows)
  int ntaps;
  int tap_shift[DRN_MAX_TAPS], tap_par[DRN_MAX_TAPS], tap_w[DRN_MAX_TAPS];
  int a_c0, b_c0;
  int nprod, split_k;
  float* out;
  long long out_ld;
  int out_col0, out_mode;
  long long out_tap_stride;
  long long out_split_stride;  // WGRAD: != 0 -> K-split s stores its partial sum at out + s*out_split_stride (no atomics)
  int out_T, out_t_mul, out_t_add;
  const float* bias;
  const float* rowscale;
  int rowscale_ld;
  float* out2;
  long long out2_ld;
  __nv_bfloat16* outp;
  long long outp_ld;
  int outp_col0;
  long long outp_plane_stride;
  float* stats;  // ROWS, optional: per 32-row block partial column sums [blk][2][N] (BatchNorm statistics, see drn_gemm_t)
  int vec_ok;   // 16-byte vector accesses are aligned
  int vec8_ok;  // 32-byte (full-sector) stores are aligned
  PlanesView a, b;
  int dbg_lbo, dbg_sbo, dbg_kadv;
};

// A group of independent problems walked by ONE persistent launch of the CTA-pair kernel (gemm2.cu)
constexpr int GROUP_MAX = 6;
constexpr int LPT_MAX_PAIRS = 80, LPT_MAX_TILES = 16;
struct GroupParams {
  int nprob;
  int raster_gm;                  // tile-rows per rasterisation group (1 = plain row-major tile order)
  int tile_start[GROUP_MAX + 1];  // prefix sums of the per-problem 256 x 256 tile counts
  int n_tiles[GROUP_MAX], m_tiles[GROUP_MAX];
  // stream-K schedule (gemm2.cu): the k-iterations of all tiles form one sequence (tile-major); SM pair c runs iterations
  // [c*sk_quota, (c+1)*sk_quota).  A pair whose range starts inside a tile stores that partial accumulator in its workspace
  // slot; the pair that ran the tile's first iterations folds the partial tiles in (in pair order: deterministic) and runs
  // the epilogue.  sk_quota == 0: static round-robin over whole tiles.
  int sk_quota;
  // hybrid schedule: tiles [0, sk_static_tiles) -- the full waves -- are walked whole, round-robin; only the k-iterations of
  // the remaining tiles (from iteration sk_it0 on) are cut into per-pair ranges.  Full stream-K: both zero.
  int sk_static_tiles, sk_it0;
  int it_start[GROUP_MAX + 1];    // prefix sums of tiles x k-iterations per problem
  int nk_tile[GROUP_MAX];         // k-iterations of one tile (uniform inside a problem)
  float* sk_ws;                   // [pairs][2 CTAs][8 column chunks][128 rows][32] fp32
  unsigned* sk_flags;             // [pairs][2 CTAs][8 epilogue warps], zero between launches
  // Static schedule balanced on the host (launch_group): the problems of a group have tiles of different lengths (a weight
  // gradient tile runs 32 k-iterations, a lateral 1x1 conv tile 4), and plain round-robin piles the long ones onto the same
  // pairs.  lpt != 0: SM pair c walks lpt_tiles[c][0 .. lpt_count[c]) -- longest-processing-time-first assignment, tiles of
  // one problem handed out in rounds so that concurrently running pairs still work on neighbouring tiles (L2 sharing).
  int epi_t;                      // transposed epilogue (epilogue_chunk_t) for full, vector-aligned chunks
  unsigned long long* trace;      // diagnostic (drn_gemm_trace): [CTA][8] %globaltimer stamps of this launch, or null
  int lpt;
  unsigned char lpt_count[LPT_MAX_PAIRS];
  unsigned short lpt_tiles[LPT_MAX_PAIRS][LPT_MAX_TILES];
  GemmKParams p[GROUP_MAX];
};
constexpr size_t SK_SLOT_FLOATS = 2 * 128 * 256;  // one 256 x 256 fp32 partial tile per SM pair
constexpr size_t SK_FLAG_BYTES = 8192;
struct GroupMaps {
  CUtensorMap a[GROUP_MAX];
  CUtensorMap b[GROUP_MAX];
};

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;               // 64 bf16 = one 128-byte swizzle row
constexpr uint32_t A_PLANE = BLOCK_M * 128;  // bytes of one A plane tile
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N>
struct TileCfg {
  static constexpr uint32_t B_PLANE = BLOCK_N * 128;
  static constexpr uint32_t STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int STAGES = (BLOCK_N >= 256) ? 2 : 3;
  static constexpr uint32_t SMEM = STAGES * STAGE + 1024;
};

// 32-byte store: one full L2 sector per thread, so a row-per-thread epilogue never leaves half-written sectors behind
// (16-byte stores made L2 fetch the other half from DRAM: 2x the output size in extra reads, ncu r01 v4).
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// Epilogue for one 32-column chunk held in registers (one row per thread).
// ------------------------------------------------------------------------------------------------
// Column sums over the 32 lanes of a warp of 32 per-lane values: lane j ends up with sum_lanes a[j] in a[0].  Recursive
// halving -- 31 shuffles instead of 32 x 5.
__device__ __forceinline__ void warp_col_reduce(float (&a)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? a[i] : a[i + s];
      const float keep = up ? a[i + s] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

// `stats_blk` = index of this warp's 32-row block in p.stats (< 0: no statistics).  Must be called by all 32 lanes.
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, float* v, bool valid, long long orow, int bb,
                                               int ncol0, float* out_base, int stats_blk = -1) {
  const int nleft = p.N - ncol0;
  const bool full = nleft >= 32;
  const int cnt = full ? 32 : nleft;
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < cnt) v[j] += __ldg(p.bias + ncol0 + j);
  }
  if (stats_blk >= 0) {  // warp-uniform: BatchNorm statistics of the conv output, padding rows excluded
    float s1[32], s2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s1[j] = valid ? v[j] : 0.f;
      s2[j] = s1[j] * s1[j];
    }
    const int lane = threadIdx.x & 31;
    warp_col_reduce(s1, lane);
    warp_col_reduce(s2, lane);
    if (lane < cnt) {
      float* st = p.stats + static_cast<long long>(stats_blk) * 2 * p.N + ncol0 + lane;
      st[0] = s1[0];
      st[p.N] = s2[0];
    }
  }
  if (!valid) return;
  if (p.out2) {
    float* o2 = p.out2 + orow * p.out2_ld + ncol0;
    if (full && p.vec8_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) st_global_v8(o2 + j, v + j);
    } else if (full && p.vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o2 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) o2[j] = v[j];
    }
  }
  if (p.rowscale) {
    const float* rs = p.rowscale + static_cast<long long>(bb) * p.rowscale_ld + ncol0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < cnt) v[j] *= __ldg(rs + j);
  }
  if (out_base) {
    float* o = out_base + orow * p.out_ld + p.out_col0 + ncol0;
    if (p.out_mode == DRN_OUT_ATOMIC) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) atomicAdd(o + j, v[j]);
    } else if (full && p.vec8_ok) {
      if (p.out_mode == DRN_OUT_ADD) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 t0 = *reinterpret_cast<const float4*>(o + j), t1 = *reinterpret_cast<const float4*>(o + j + 4);
          float w[8] = {t0.x + v[j],     t0.y + v[j + 1], t0.z + v[j + 2], t0.w + v[j + 3],
                        t1.x + v[j + 4], t1.y + v[j + 5], t1.z + v[j + 6], t1.w + v[j + 7]};
          st_global_v8(o + j, w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8) st_global_v8(o + j, v + j);
      }
    } else if (full && p.vec_ok) {
      if (p.out_mode == DRN_OUT_ADD) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 t = *reinterpret_cast<const float4*>(o + j);
          t.x += v[j]; t.y += v[j + 1]; t.z += v[j + 2]; t.w += v[j + 3];
          *reinterpret_cast<float4*>(o + j) = t;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) o[j] = (p.out_mode == DRN_OUT_ADD) ? o[j] + v[j] : v[j];
    }
  }
  if (p.outp) {
    __nv_bfloat16* oh = p.outp + orow * p.outp_ld + p.outp_col0 + ncol0;
    __nv_bfloat16* ol = oh + p.outp_plane_stride;
    if (full && p.vec_ok) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __nv_bfloat16 h0, l0, h1, l1;
          split_bf16(v[j + 2 * q], h0, l0);
          split_bf16(v[j + 2 * q + 1], h1, l1);
          h[q] = pack_bf16x2(h0, h1);
          l[q] = pack_bf16x2(l0, l1);
        }
        *reinterpret_cast<uint4*>(oh + j) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(ol + j) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < cnt) {
          __nv_bfloat16 h0, l0;
          split_bf16(v[j], h0, l0);
          oh[j] = h0;
          ol[j] = l0;
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Transposed epilogue of one FULL 32 x 32 chunk (persistent pair kernel).  tcgen05.ld hands every lane one ROW of the chunk, so
// a row-per-thread store instruction touches 32 different 128-byte lines (32 LSU wavefronts for 1 KB) -- measured 4-25 us of
// exposed drain at the end of every launch (scripts/gemm_trace.py) and the bound of the short-K layers.  Here the warp passes
// the chunk through a 4 KB XOR-swizzled shared-memory tile and comes back with lane = (row l/8 of a group of 4, columns
// 4 (l%8) ..+3): a store instruction then writes four complete 128-byte lines (4 wavefronts for 512 B), bias / gate values
// are one 16-byte quantity per lane, and the BatchNorm column sums fall out of 8 adds + 2 shuffles instead of 62 shuffles.
// Arithmetic and its order per element are those of epilogue_chunk (bit-identical outputs; only the order of the fp32 partial
// sums inside a 32-row statistics block differs).  Requires p.vec_ok, a full chunk and a non-atomic output mode.
struct EpiRows {      // the 8 rows of its 32-row block a lane serves: row 4 i + lane / 8
  int orow[8];        // output row (< 2^31: rows of one problem)
  int bb[8];          // sample (row gate)
  unsigned valid;     // bit i
};

__device__ __forceinline__ void epilogue_chunk_t(const GemmKParams& p, const float* v, uint32_t stage, const EpiRows& er, int ncol0,
                                                 float* out_base, int stats_blk, int lane) {
  const int cg = lane & 7, rsub = lane >> 3;
  const int col = ncol0 + 4 * cg;
  const bool add = out_base && p.out_mode == DRN_OUT_ADD;
  // Everything that comes from global memory is requested up front (bias, row gate; accumulate mode: the old values of all 8
  // rows, once this lane's accumulator row has left its registers), so that the loads overlap the shared-memory transposition
  // instead of forming a load -> add -> store chain per row.
  float b4[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int k = 0; k < 4; ++k) b4[k] = __ldg(p.bias + col + k);
  }
  float rs4[4] = {1.f, 1.f, 1.f, 1.f};
  int rs_bb = -1;
  if (p.rowscale && (er.valid & 1u)) {  // (a padding row's sample index lies outside the gate tensor)
    rs_bb = er.bb[0];
#pragma unroll
    for (int k = 0; k < 4; ++k) rs4[k] = __ldg(p.rowscale + static_cast<long long>(rs_bb) * p.rowscale_ld + col + k);
  }
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const uint32_t a = stage + lane * 128 + ((j4 ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * j4]), "f"(v[4 * j4 + 1]), "f"(v[4 * j4 + 2]),
                 "f"(v[4 * j4 + 3])
                 : "memory");
  }
  float4 old[8];
  if (add) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      old[i] = ((er.valid >> i) & 1u) ? __ldcg(reinterpret_cast<const float4*>(out_base + static_cast<long long>(er.orow[i]) * p.out_ld + p.out_col0 + col))
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = 4 * i + rsub;
    float x[4];
    const uint32_t a = stage + rl * 128 + ((cg ^ (rl & 7)) << 4);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]) : "r"(a));
    if (!((er.valid >> i) & 1u)) continue;  // padding rows: no output, no statistics
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] += b4[k];
    if (stats_blk >= 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        s1[k] += x[k];
        s2[k] += x[k] * x[k];
      }
    }
    const long long orow = er.orow[i];
    if (p.out2) *reinterpret_cast<float4*>(p.out2 + orow * p.out2_ld + col) = make_float4(x[0], x[1], x[2], x[3]);
    if (p.rowscale) {
      if (er.bb[i] != rs_bb) {  // the block straddles two samples (short sequences only)
        rs_bb = er.bb[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) rs4[k] = __ldg(p.rowscale + static_cast<long long>(rs_bb) * p.rowscale_ld + col + k);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k] *= rs4[k];
    }
    if (out_base) {
      float4* o = reinterpret_cast<float4*>(out_base + orow * p.out_ld + p.out_col0 + col);
      if (add) *o = make_float4(old[i].x + x[0], old[i].y + x[1], old[i].z + x[2], old[i].w + x[3]);
      else *o = make_float4(x[0], x[1], x[2], x[3]);
    }
    if (p.outp) {
      __nv_bfloat16* oh = p.outp + orow * p.outp_ld + p.outp_col0 + col;
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split_bf16(x[k], h[k], l[k]);
      *reinterpret_cast<uint2*>(oh) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
      *reinterpret_cast<uint2*>(oh + p.outp_plane_stride) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
  }
  if (stats_blk >= 0) {  // warp-uniform
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 8);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 8);
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], 16);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], 16);
    }
    if (rsub == 0) {
      float* st = p.stats + static_cast<long long>(stats_blk) * 2 * p.N + col;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        st[k] = s1[k];
        st[p.N + k] = s2[k];
      }
    }
  }
  __syncwarp();  // the staging tile is free for the next chunk
}

}  // namespace drn
