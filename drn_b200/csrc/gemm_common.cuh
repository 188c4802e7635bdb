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
  int it_start[GROUP_MAX + 1];    // prefix sums of tiles x k-iterations per problem
  int nk_tile[GROUP_MAX];         // k-iterations of one tile (uniform inside a problem)
  float* sk_ws;                   // [pairs][2 CTAs][8 column chunks][128 rows][32] fp32
  unsigned* sk_flags;             // [pairs][2 CTAs][8 epilogue warps], zero between launches
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

}  // namespace drn
