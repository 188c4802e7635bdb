// drn_gemm / drn_gemm_group, engine 0 for large problems: PERSISTENT kernel on CTA PAIRS (tcgen05 cta_group::2) that walks the
// 256 x 256 output tiles of up to GROUP_MAX independent problems in ONE launch.
//
//   * one cluster of 2 CTAs per SM pair (74 pairs on a B200); each CTA stages its own 128 rows of A and its own 128
//     columns of B (hi and lo planes) by TMA, the leader CTA's single MMA thread issues 256 x 256 x 16 UMMAs that read both
//     CTAs' shared memory, so every staged byte feeds twice the math of the 1-CTA kernel (gemm.cu);
//   * 3-stage TMA->MMA pipeline that runs across tile AND problem boundaries;
//   * the fp32 accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i (TMEM -> registers ->
//     global) overlaps the MMAs of tile i+1;
//   * grouping: the DRN path is dominated by ~40 small contractions per step (FPN / head convs of three pyramid levels,
//     their data- and weight-gradients).  Launched one by one each pays pipeline fill/drain and leaves most of a wave
//     idle (5-27 % tensor-pipe utilisation measured); as one tile list they fill the waves and share one fill/drain.
//     ROWS (conv forward / dgrad) and WGRAD problems mix freely in a group.
// Operand forms, tensor maps and epilogue semantics are those of gemm.cu (include/drn_b200.h).
#include <cuda.h>
#include <stdlib.h>

#include "gemm_common.cuh"

namespace drn {

constexpr int P2_STAGES = 3;
constexpr uint32_t P2_HALF = 128 * 128;                 // one plane of one operand half: 128 rows x 128 B
constexpr uint32_t P2_STAGE = 4 * P2_HALF;              // A hi, A lo, B hi, B lo
constexpr uint32_t P2_SMEM = P2_STAGES * P2_STAGE + 1024;
constexpr int P2_THREADS = 192;
constexpr int P2_TILE = 256;

struct PairTile {
  int prob;       // problem of the group
  int b0, t0;     // ROWS: first (sample, time slot) of this CTA's 128 rows
  int ms;         // ROWS: index of this CTA's 128-row sub-tile
  int m0;         // WGRAD: first A channel (= output row) of this CTA
  int nb;         // first column of B staged by this CTA
  int n0;         // first output column of the pair's tile
  int tap;        // WGRAD: tap handled by this tile
  int split;      // WGRAD: K-split index
  int it_begin, nk;
};

__device__ __forceinline__ PairTile decode_tile(const GroupParams& gp, int tile, int rank) {
  PairTile t{};
  int pr = 0;
#pragma unroll
  for (int i = 1; i < GROUP_MAX; ++i)
    if (i < gp.nprob && tile >= gp.tile_start[i]) pr = i;
  t.prob = pr;
  const GemmKParams& p = gp.p[pr];
  const int local = tile - gp.tile_start[pr];
  const int n_tiles = gp.n_tiles[pr], m_tiles = gp.m_tiles[pr];
  const bool wgrad = (p.form == DRN_GEMM_WGRAD);
  // Optional rasterisation: tiles walked in groups of raster_gm tile-rows, n-major inside a group, so concurrently running
  // clusters share a few A row-blocks.  Measured neutral on B200 (prop_fc forward: 627 -> 611 MB of DRAM reads for 201 MB of
  // operands, no change in time: the kernel sits at ~90 % tensor pipe / ~18 % DRAM), so the default is 1 = row-major order.
  const int per = m_tiles * n_tiles;
  const int idx = local % per;
  const int z = local / per;  // WGRAD: (tap, split); ROWS: split
  const int RASTER_GM = gp.raster_gm;
  const int grp = idx / (RASTER_GM * n_tiles);
  const int first_m = grp * RASTER_GM;
  const int gm = min(RASTER_GM, m_tiles - first_m);
  const int within = idx - grp * RASTER_GM * n_tiles;
  const int mt = first_m + within % gm;
  const int nt = within / gm;
  t.n0 = nt * P2_TILE;
  t.nb = t.n0 + rank * 128;
  if (!wgrad) {
    t.split = z;                     // K-split of a ROWS problem (conv0 forward: few tiles, very long K)
    const int ms = 2 * mt + rank;    // 128-row sub-tile of this CTA
    t.ms = ms;
    if (p.Bbm == 1) {
      t.b0 = ms / p.tiles_per_sample;
      t.t0 = (ms % p.tiles_per_sample) * p.Rm;
    } else {
      t.b0 = ms * p.Bbm;
      t.t0 = 0;
    }
    const int its = p.ntaps * (p.K / BLOCK_K);
    t.it_begin = static_cast<int>(static_cast<long long>(its) * t.split / p.split_k);
    t.nk = static_cast<int>(static_cast<long long>(its) * (t.split + 1) / p.split_k) - t.it_begin;
  } else {
    t.m0 = mt * P2_TILE + rank * 128;
    t.tap = z / p.split_k;
    t.split = z % p.split_k;
    t.it_begin = static_cast<int>(static_cast<long long>(p.num_kblocks) * t.split / p.split_k);
    t.nk = static_cast<int>(static_cast<long long>(p.num_kblocks) * (t.split + 1) / p.split_k) - t.it_begin;
  }
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ GroupParams gp, const __grid_constant__ GroupMaps gm, int num_tiles) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[P2_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[P2_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_holder;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < gp.nprob; ++i) {
      tma_prefetch_desc(&gm.a[i]);
      tma_prefetch_desc(&gm.b[i]);
    }
    for (int s = 0; s < P2_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), 8);  // 4 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(smem_u32(&tmem_base_holder), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;
  pdl_wait();  // prologue done (barriers, TMEM, descriptor prefetch): wait here for the kernel that produces our operands

  if (warp == 0) {
    // ===== TMA producer (one thread per CTA; completion is signalled on the LEADER's full barrier) =====
    if (lane == 0) {
      int g = 0;  // global k-iteration counter (pipeline runs across tiles and problems)
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const PairTile t = decode_tile(gp, tile, rank);
        const GemmKParams& p = gp.p[t.prob];
        const CUtensorMap* tma_a = &gm.a[t.prob];
        const CUtensorMap* tma_b = &gm.b[t.prob];
        const bool wgrad = (p.form == DRN_GEMM_WGRAD);
        const bool b_mn = wgrad || (p.b_mn != 0);
        const int nplanes = (p.nprod == 1) ? 1 : 2;
        const int kpt = wgrad ? 1 : p.K / BLOCK_K;
        const uint32_t tx = 2u * nplanes * 2u * P2_HALF;  // both CTAs' bytes land on the leader's barrier
        for (int i = 0; i < t.nk; ++i, ++g) {
          const int s = g % P2_STAGES;
          const uint32_t ph = (g / P2_STAGES) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          if (leader) mbar_arrive_expect_tx(smem_u32(&full_bar[s]), tx);
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
          const uint32_t sa = smem_base + s * P2_STAGE;
          const uint32_t sb = sa + 2 * P2_HALF;
          const int it = t.it_begin + i;
          if (!wgrad) {
            const int tap = it / kpt, kb = it % kpt;
            for (int pl = 0; pl < nplanes; ++pl) {
              tma_load_5d_pair(sa + pl * P2_HALF, tma_a, fb, p.a_c0 + kb * BLOCK_K, p.tap_par[tap], t.t0 + p.tap_shift[tap],
                               t.b0, pl);
              if (!b_mn) {
                tma_load_5d_pair(sb + pl * P2_HALF, tma_b, fb, p.b_c0 + kb * BLOCK_K, 0, t.nb, p.tap_w[tap], pl);
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  tma_load_5d_pair(sb + pl * P2_HALF + j * 8192, tma_b, fb, p.b_c0 + t.nb + j * 64, 0, kb * BLOCK_K,
                                   p.tap_w[tap], pl);
              }
            }
          } else {
            int bk, tk;
            if (p.Bbk == 1) {
              bk = it / p.kblocks_per_sample;
              tk = (it % p.kblocks_per_sample) * p.Rk;
            } else {
              bk = it * p.Bbk;
              tk = 0;
            }
            for (int pl = 0; pl < nplanes; ++pl) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                tma_load_5d_pair(sa + pl * P2_HALF + j * 8192, tma_a, fb, p.a_c0 + t.m0 + j * 64, 0, tk, bk, pl);
                tma_load_5d_pair(sb + pl * P2_HALF + j * 8192, tma_b, fb, p.b_c0 + t.nb + j * 64, p.tap_par[t.tap],
                                 tk + p.tap_shift[t.tap], bk, pl);
              }
            }
          }
        }
      }
      // drain: every commit multicast to this CTA's empty barriers must have landed before the CTA may exit
      for (int d = 0; d < P2_STAGES && d < g; ++d) {
        const int gi = g - 1 - d;
        mbar_wait(smem_u32(&empty_bar[gi % P2_STAGES]), (gi / P2_STAGES) & 1);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA =====
    if (leader && lane == 0) {
      int g = 0, lt = 0;  // lt counts the tiles that actually use an accumulator stage
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const PairTile t = decode_tile(gp, tile, rank);
        if (t.nk <= 0) continue;
        const GemmKParams& p = gp.p[t.prob];
        const bool wgrad = (p.form == DRN_GEMM_WGRAD);
        const bool a_mn = wgrad;
        const bool b_mn = wgrad || (p.b_mn != 0);
        const uint32_t idesc = umma_idesc_bf16(P2_TILE, P2_TILE, a_mn, b_mn);
        const uint32_t a_lbo = a_mn ? 8192u : 0u, a_kadv = a_mn ? 2048u : 32u;
        const uint32_t b_lbo = b_mn ? 8192u : 0u, b_kadv = b_mn ? 2048u : 32u;
        const int nprod = p.nprod;
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        ++lt;
        mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1);  // both CTAs' epilogues drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * P2_TILE;
        uint32_t accumulate = 0;
        for (int i = 0; i < t.nk; ++i, ++g) {
          const int s = g % P2_STAGES;
          const uint32_t ph = (g / P2_STAGES) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          const uint32_t sa = smem_base + s * P2_STAGE;
          const uint32_t sb = sa + 2 * P2_HALF;
          for (int prod = 0; prod < nprod; ++prod) {
            const uint32_t pa = (prod >> 1) & 1, pb = prod & 1;  // (hi,hi) (hi,lo) (lo,hi) (lo,lo)
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t ad = umma_smem_desc(sa + pa * P2_HALF + k * a_kadv, a_lbo, 1024u);
              const uint64_t bd = umma_smem_desc(sb + pb * P2_HALF + k * b_kadv, b_lbo, 1024u);
              umma_bf16_pair(d_tmem, ad, bd, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit_pair(smem_u32(&empty_bar[s]));  // frees this stage in BOTH CTAs
        }
        umma_commit_pair(smem_u32(&tmem_full_bar[as]));
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..5 (both CTAs): TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int lt = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const PairTile t = decode_tile(gp, tile, rank);
      if (t.nk <= 0) continue;
      const GemmKParams& p = gp.p[t.prob];
      const bool wgrad = (p.form == DRN_GEMM_WGRAD);
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      ++lt;
      mbar_wait(smem_u32(&tmem_full_bar[as]), aph);
      tc_fence_after();
      bool valid;
      long long orow;
      int bb = 0;
      float* out_base = p.out;
      if (!wgrad) {
        bb = t.b0 + row / p.Rm;
        const int tt = t.t0 + row % p.Rm;
        valid = (bb < p.B) && (tt < p.T);
        orow = static_cast<long long>(bb) * p.out_T + static_cast<long long>(tt) * p.out_t_mul + p.out_t_add;
        if (out_base) out_base += t.split * p.out_split_stride;
      } else {
        valid = (t.m0 + row) < p.M;
        orow = t.m0 + row;
        if (out_base) out_base += p.tap_w[t.tap] * p.out_tap_stride + t.split * p.out_split_stride;
      }
      const uint32_t taddr = tmem_base + as * P2_TILE + (static_cast<uint32_t>(q * 32) << 16);
      const int stats_blk = (!wgrad && p.stats) ? t.ms * 4 + q : -1;  // 32-row block of the output (BatchNorm partial sums)
#pragma unroll 1
      for (int c0 = 0; c0 < P2_TILE; c0 += 32) {
        if (t.n0 + c0 >= p.N) break;
        float v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        epilogue_chunk(p, v, valid, orow, bb, t.n0 + c0, out_base, stats_blk);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// Process-wide cap on the SM pairs the persistent kernel occupies (0 = none), set by the caller around individual launches:
// the data-parallel schedule leaves 4 of the 74 pairs to NCCL while the prop_fc weight gradient runs (drn_b200/dense.py).
static int g_pair_clusters = 0;
void set_pair_clusters(int n) { g_pair_clusters = n > 0 ? n : 0; }

int launch_group(const GroupParams& gp, const GroupMaps& gm, int sm_count, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaFuncSetAttribute(gemm_pair): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int num_tiles = gp.tile_start[gp.nprob];
  int clusters = sm_count / 2;
  static int cap = -1;  // DRN_PAIR_CLUSTERS: leave SM pairs free for kernels of other streams (tuning / probing knob)
  if (cap < 0) {
    const char* e = getenv("DRN_PAIR_CLUSTERS");
    cap = e ? atoi(e) : 0;
  }
  if (cap > 0 && clusters > cap) clusters = cap;
  if (g_pair_clusters > 0 && clusters > g_pair_clusters) clusters = g_pair_clusters;
  if (clusters > num_tiles) clusters = num_tiles;
  if (clusters < 1) clusters = 1;
  launch_k(gemm_pair_kernel, 2 * clusters, P2_THREADS, P2_SMEM, st, gp, gm, num_tiles);
  return check_launch("gemm_pair_kernel");
}

}  // namespace drn
