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
//   * stream-K schedule (when the caller supplies a workspace, drn_gemm_group_ws): instead of whole tiles round-robin, the
//     k-iterations of ALL tiles of the launch form one sequence that is cut into equal contiguous ranges, one per SM pair.
//     224 tower tiles on 74 pairs are 3.03 waves of tiles but 72.6 (of 73) iterations per pair; 56 long tiles of the conv2
//     backward are one 32-iteration wave but 20.8 iterations per pair.  At most the first segment of a pair starts inside a
//     tile: its accumulator goes to the pair's workspace slot (fp32, plain stores) and a flag; the pair that ran the first
//     iterations of that tile -- always its LAST segment in time, so the partial tiles are long there -- adds them in pair
//     order (deterministic) and runs the normal epilogue (bias, gate, statistics, planes).
// Operand forms, tensor maps and epilogue semantics are those of gemm.cu (include/drn_b200.h).
#include <cuda.h>
#include <stdlib.h>

#include "gemm_common.cuh"

namespace drn {

constexpr int P2_STAGES = 3;
constexpr uint32_t P2_HALF = 128 * 128;                 // one plane of one operand half: 128 rows x 128 B
constexpr uint32_t P2_STAGE = 4 * P2_HALF;              // A hi, A lo, B hi, B lo
constexpr int P2_EPI_WARPS = 8;                         // 2 warps per TMEM lane quarter, each drains half of the 256 columns
constexpr uint32_t P2_EPI_STAGE = 4096;                 // per epilogue warp: one 32 x 32 fp32 chunk (transposed epilogue)
constexpr uint32_t P2_SMEM = P2_STAGES * P2_STAGE + P2_EPI_WARPS * P2_EPI_STAGE + 1024;
constexpr int P2_THREADS = 64 + 32 * P2_EPI_WARPS;
constexpr int P2_TILE = 256;
constexpr int TRACE_CTAS = 160;

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

// Per-role cursor over the work of one SM pair: whole tiles (static round-robin) or the segments of its stream-K range.
struct PairSched {
  int tile, stride, num_tiles;  // static
  int pair;                     // static, host-balanced (gp.lpt): this pair's row of gp.lpt_tiles
  int it, it_end;               // stream-K: global k-iteration cursor / end of this pair's range
  int seg_it;                   // stream-K: global iteration at which the current segment starts
};

__device__ __forceinline__ PairSched sched_init(const GroupParams& gp, int cluster_id, int num_clusters, int num_tiles) {
  PairSched s{};
  s.tile = gp.lpt ? 0 : cluster_id;
  s.stride = num_clusters;
  s.num_tiles = num_tiles;
  s.pair = cluster_id;
  if (gp.sk_quota > 0) {
    const int total = gp.it_start[gp.nprob];
    s.it = min(total, gp.sk_it0 + cluster_id * gp.sk_quota);
    s.it_end = min(total, s.it + gp.sk_quota);
  }
  return s;
}

// Next segment of this pair: tile `t`, k-iterations [k0, k0 + kn) of the tile's t.nk.  False when the pair is done.
__device__ __forceinline__ bool next_seg(const GroupParams& gp, PairSched& s, int rank, PairTile& t, int& k0, int& kn) {
  if (gp.sk_quota <= 0) {
    if (gp.lpt) {  // host-balanced tile list of this pair (s.tile = cursor into it)
      if (s.tile >= gp.lpt_count[s.pair]) return false;
      t = decode_tile(gp, gp.lpt_tiles[s.pair][s.tile], rank);
      ++s.tile;
    } else {
      if (s.tile >= s.num_tiles) return false;
      t = decode_tile(gp, s.tile, rank);
      s.tile += s.stride;
    }
    k0 = 0;
    kn = t.nk;
    return true;
  }
  if (s.tile < gp.sk_static_tiles) {  // hybrid: the full waves first, whole tiles
    t = decode_tile(gp, s.tile, rank);
    s.tile += s.stride;
    k0 = 0;
    kn = t.nk;
    return true;
  }
  if (s.it >= s.it_end) return false;
  int pr = 0;
#pragma unroll
  for (int i = 1; i < GROUP_MAX; ++i)
    if (i < gp.nprob && s.it >= gp.it_start[i]) pr = i;
  const int loc = s.it - gp.it_start[pr];
  const int nkp = gp.nk_tile[pr];
  const int local = loc / nkp;
  k0 = loc - local * nkp;
  t = decode_tile(gp, gp.tile_start[pr] + local, rank);
  kn = min(nkp - k0, s.it_end - s.it);
  s.seg_it = s.it;
  s.it += kn;
  return true;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Diagnostic stamps (drn_gemm_trace): where a launch spends its time -- entry, prologue done, first operands landed, last MMA
// issued, first / last accumulator drained, exit -- per CTA, from the GPU's nanosecond timer.
__device__ __forceinline__ void trace_stamp(const GroupParams& gp, int slot) {
  if (gp.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    gp.trace[blockIdx.x * 8 + slot] = t;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ GroupParams gp, const __grid_constant__ GroupMaps gm, int num_tiles) {
  pdl_trigger();
  if (threadIdx.x == 0) trace_stamp(gp, 0);
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
      mbar_init(smem_u32(&tmem_empty_bar[a]), 2 * P2_EPI_WARPS);  // epilogue warps x 2 CTAs
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
  if (threadIdx.x == 0) trace_stamp(gp, 1);

  if (warp == 0) {
    // ===== TMA producer (one thread per CTA; completion is signalled on the LEADER's full barrier) =====
    if (lane == 0) {
      int g = 0;  // global k-iteration counter (pipeline runs across tiles and problems)
      PairSched sc = sched_init(gp, cluster_id, num_clusters, num_tiles);
      PairTile t;
      int k0, kn;
      while (next_seg(gp, sc, rank, t, k0, kn)) {
        const GemmKParams& p = gp.p[t.prob];
        const CUtensorMap* tma_a = &gm.a[t.prob];
        const CUtensorMap* tma_b = &gm.b[t.prob];
        const bool wgrad = (p.form == DRN_GEMM_WGRAD);
        const bool b_mn = wgrad || (p.b_mn != 0);
        const int nplanes = (p.nprod == 1) ? 1 : 2;
        const int kpt = wgrad ? 1 : p.K / BLOCK_K;
        const uint32_t tx = 2u * nplanes * 2u * P2_HALF;  // both CTAs' bytes land on the leader's barrier
        for (int i = 0; i < kn; ++i, ++g) {
          const int s = g % P2_STAGES;
          const uint32_t ph = (g / P2_STAGES) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          if (leader) mbar_arrive_expect_tx(smem_u32(&full_bar[s]), tx);
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
          const uint32_t sa = smem_base + s * P2_STAGE;
          const uint32_t sb = sa + 2 * P2_HALF;
          const int it = t.it_begin + k0 + i;
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
      int g = 0, lt = 0;  // lt counts the segments that actually use an accumulator stage
      PairSched sc = sched_init(gp, cluster_id, num_clusters, num_tiles);
      PairTile t;
      int k0, kn;
      while (next_seg(gp, sc, rank, t, k0, kn)) {
        if (kn <= 0) continue;
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
        for (int i = 0; i < kn; ++i, ++g) {
          const int s = g % P2_STAGES;
          const uint32_t ph = (g / P2_STAGES) & 1;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tc_fence_after();
          if (g == 0) trace_stamp(gp, 2);
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
        trace_stamp(gp, 3);  // overwritten tile after tile: the last one stays
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2.. (both CTAs): TMEM lane quarter = warp % 4 (hardware rule); with 8 warps the two warps of a
    // quarter split the 256 accumulator columns, which halves the drain time of a tile -- the short-K layers (FPN laterals,
    // stride-2 data gradients: 4-16 k-iterations a tile) are bound by it, and every launch ends with one exposed drain =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    constexpr int COLS_PER_WARP = P2_TILE / (P2_EPI_WARPS / 4);
    const int cbeg = ((warp - 2) >> 2) * COLS_PER_WARP, cend = cbeg + COLS_PER_WARP;
    const uint32_t epi_stage = smem_base + P2_STAGES * P2_STAGE + (warp - 2) * P2_EPI_STAGE;
    const bool epi_transposed = gp.epi_t != 0;
    int lt = 0;
    PairSched sc = sched_init(gp, cluster_id, num_clusters, num_tiles);
    PairTile t;
    int k0, kn;
    while (next_seg(gp, sc, rank, t, k0, kn)) {
      if (kn <= 0) continue;
      const GemmKParams& p = gp.p[t.prob];
      const bool wgrad = (p.form == DRN_GEMM_WGRAD);
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      ++lt;
      const uint32_t taddr = tmem_base + as * P2_TILE + (static_cast<uint32_t>(q * 32) << 16);
      // stream-K / hybrid: partial tiles live at [pair][rank][32-column chunk][float4 j of the chunk][128 rows][4] of the
      // workspace -- rows innermost, so that the 32 lanes (= rows) of a store or load instruction touch 512 contiguous bytes
      // (row-major chunks made every instruction touch 32 lines: the fold of 5 partial tiles took ~40 us)
      const size_t ws_off = static_cast<size_t>(rank) * (128 * 256) + static_cast<size_t>(row) * 4;
      int ncontrib = 0;
      if (k0 == 0 && kn < t.nk) {
        // owner of a tile other pairs finish: pairs cluster_id+1 ... whose ranges start before the end of this tile
        const int tile_end = sc.seg_it + t.nk;
        while (gp.sk_it0 + (cluster_id + 1 + ncontrib) * gp.sk_quota < tile_end) ++ncontrib;
        if (lane == 0) {
          for (int j = 0; j < ncontrib; ++j) {
            const unsigned* f = gp.sk_flags + (cluster_id + 1 + j) * 16 + rank * 8 + (warp - 2);
            const long long t0 = clock64();
            while (ld_acquire_gpu(f) == 0u) {
              if (clock64() - t0 > 4000000000LL) __trap();
            }
          }
        }
        __syncwarp();
      }
      mbar_wait(smem_u32(&tmem_full_bar[as]), aph);
      tc_fence_after();
      if (warp == 2 && lane == 0) trace_stamp(gp, lt == 1 ? 4 : 5);  // first / last accumulator complete
      if (k0 > 0) {
        // not the owner: this pair's range starts inside the tile -> partial accumulator to the workspace slot + flag
        float* slot = gp.sk_ws + static_cast<size_t>(cluster_id) * SK_SLOT_FLOATS + ws_off;
#pragma unroll 1
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
          if (t.n0 + c0 >= p.N) break;
          float v[32];
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          float* d = slot + (c0 >> 5) * (128 * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *reinterpret_cast<float4*>(d + j4 * 512) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));
          __threadfence();
          st_release_gpu(gp.sk_flags + cluster_id * 16 + rank * 8 + (warp - 2), 1u);
        }
        continue;
      }
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
      const int stats_blk = (!wgrad && p.stats) ? t.ms * 4 + q : -1;  // 32-row block of the output (BatchNorm partial sums)
      // transposed epilogue (gemm_common.cuh): this lane's 8 rows of the block, fetched from the lanes that own them
      // (plain stores -- weight-gradient slices, K-split slices -- stay row-per-thread: nothing to load, and the final drain of
      // a launch, bound by the ~4 TB/s at which 148 SMs can write, measured 5 us that way against 8 us through shared memory)
      const bool fast = (p.vec_ok != 0) && (p.out_mode != DRN_OUT_ATOMIC) && epi_transposed &&
                        (p.bias || p.stats || p.out2 || p.rowscale || p.outp || p.out_mode == DRN_OUT_ADD);
      EpiRows er;
      er.valid = 0u;
      if (fast) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int src = 4 * i + (lane >> 3);
          er.orow[i] = __shfl_sync(0xffffffffu, static_cast<int>(orow), src);
          er.bb[i] = __shfl_sync(0xffffffffu, bb, src);
          er.valid |= (__shfl_sync(0xffffffffu, valid ? 1u : 0u, src) & 1u) << i;
        }
      }
#pragma unroll 1
      for (int c0 = cbeg; c0 < cend; c0 += 32) {
        if (t.n0 + c0 >= p.N) break;
        float v[32];
        tmem_ld_32x32(taddr + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < ncontrib; ++j) {  // fold the partial tiles, in pair order
          const float* src = gp.sk_ws + static_cast<size_t>(cluster_id + 1 + j) * SK_SLOT_FLOATS + ws_off + (c0 >> 5) * (128 * 32);
          float4 x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = __ldcg(reinterpret_cast<const float4*>(src + i * 512));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[4 * i] += x[i].x; v[4 * i + 1] += x[i].y; v[4 * i + 2] += x[i].z; v[4 * i + 3] += x[i].w;
          }
        }
        if (fast && t.n0 + c0 + 32 <= p.N) epilogue_chunk_t(p, v, epi_stage, er, t.n0 + c0, out_base, stats_blk, lane);
        else epilogue_chunk(p, v, valid, orow, bb, t.n0 + c0, out_base, stats_blk);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[as]), 0));
        for (int j = 0; j < ncontrib; ++j) gp.sk_flags[(cluster_id + 1 + j) * 16 + rank * 8 + (warp - 2)] = 0u;  // re-armed for the next launch
      }
      if (warp == 2 && lane == 0) trace_stamp(gp, 6);  // accumulator drained (the last one stays)
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
  if (threadIdx.x == 0) trace_stamp(gp, 7);
}

// Process-wide cap on the SM pairs the persistent kernel occupies (0 = none), set by the caller around individual launches:
// the data-parallel schedule leaves 4 of the 74 pairs to NCCL while the prop_fc weight gradient runs (drn_b200/dense.py).
static int g_pair_clusters = 0;
void set_pair_clusters(int n) { g_pair_clusters = n > 0 ? n : 0; }
// Schedule of launches that are given a workspace: 0 = static only, 1 = hybrid (static full waves + k-split last wave; default),
// 2 = full stream-K (every tile boundary ignored; measured slower on the DRN layers, kept for tests and A/B).
static int g_schedule = -1;
void set_schedule(int mode) { g_schedule = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }

// drn_gemm_trace(buf, launches): the next `launches` launches of the pair kernel (eager or captured into a CUDA graph: the slot
// is baked into the captured launch) write their stamps to buf[launch][TRACE_CTAS][8]; each returns its CTA count in info.
static unsigned long long* g_trace = nullptr;
static int g_trace_left = 0, g_trace_next = 0;
static int g_trace_info[256][3];
void gemm_trace(unsigned long long* buf, int launches) {
  g_trace = buf;
  g_trace_left = buf ? (launches > 256 ? 256 : launches) : 0;
  g_trace_next = 0;
}
int gemm_trace_info(int launch, int* ctas, int* tiles, int* lpt) {
  if (launch < 0 || launch >= g_trace_next) return -1;
  *ctas = g_trace_info[launch][0];
  *tiles = g_trace_info[launch][1];
  *lpt = g_trace_info[launch][2];
  return 0;
}

// Stream-K runs when the caller passes a workspace (drn_gemm_group_ws).  DRN_SK_MIN (environment, read once) = fewest
// k-iterations worth giving an SM pair (a range shorter than that costs more in the fold than it saves).
static int sk_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Host-side planning of one launch (pure: no CUDA call): fills the schedule fields of `gp` (static round-robin, host-balanced
// tile lists, hybrid or full stream-K ranges) and returns the number of SM pairs (clusters) to launch.  Shared by launch_group and
// by drn_gemm_schedule_probe, which lets the CPU tests check the planner without a GPU.
int plan_group(GroupParams& gp, const int* nk_tile, int sm_count, void* ws, size_t ws_bytes) {
  if (g_schedule < 0) {
    const char* e = getenv("DRN_SCHEDULE");  // static | hybrid | streamk
    g_schedule = (e && e[0] == 's' && e[1] == 't' && e[2] == 'a') ? 0 : ((e && e[0] == 's' && e[1] == 't' && e[2] == 'r') ? 2 : 1);
  }
  const int num_tiles = gp.tile_start[gp.nprob];
  int clusters = sm_count / 2;
  static int cap = -1;  // DRN_PAIR_CLUSTERS: leave SM pairs free for kernels of other streams (tuning / probing knob)
  static int sk_min = 4, hybrid_min_saved = 24, hybrid_min_quota = 8;
  if (cap < 0) {
    cap = sk_env("DRN_PAIR_CLUSTERS", 0);
    sk_min = sk_env("DRN_SK_MIN", 4);
    if (sk_min < 1) sk_min = 1;
    hybrid_min_saved = sk_env("DRN_HYBRID_MIN_SAVED", 24);
    hybrid_min_quota = sk_env("DRN_HYBRID_MIN_QUOTA", 8);
    if (hybrid_min_quota < 1) hybrid_min_quota = 1;
  }
  if (cap > 0 && clusters > cap) clusters = cap;
  if (g_pair_clusters > 0 && clusters > g_pair_clusters) clusters = g_pair_clusters;
  if (clusters < 1) clusters = 1;
  // ---- stream-K: equal contiguous ranges of k-iterations per SM pair (needs uniform tiles inside every problem) ------------
  gp.sk_quota = 0;
  gp.sk_static_tiles = 0;
  gp.sk_it0 = 0;
  gp.sk_ws = nullptr;
  gp.sk_flags = nullptr;
  bool uniform = true;
  long long total = 0;
  gp.it_start[0] = 0;
  for (int k = 0; k < gp.nprob; ++k) {
    if (nk_tile[k] <= 0) uniform = false;
    gp.nk_tile[k] = nk_tile[k] > 0 ? nk_tile[k] : 1;
    total += static_cast<long long>(gp.tile_start[k + 1] - gp.tile_start[k]) * gp.nk_tile[k];
    gp.it_start[k + 1] = static_cast<int>(total);
  }
  for (int k = gp.nprob; k < GROUP_MAX; ++k) gp.it_start[k + 1] = gp.it_start[gp.nprob];
  const bool ws_ok = ws && uniform && total > 0 && total < (1ll << 30) &&
                     ws_bytes >= SK_FLAG_BYTES + static_cast<size_t>(clusters) * SK_SLOT_FLOATS * sizeof(float) &&
                     static_cast<size_t>(clusters) * 16 * sizeof(unsigned) <= SK_FLAG_BYTES;
  bool same_nk = true;
  for (int k = 1; k < gp.nprob; ++k) same_nk = same_nk && gp.nk_tile[k] == gp.nk_tile[0];
  if (ws_ok && g_schedule == 2) {  // full stream-K
    int quota = static_cast<int>((total + clusters - 1) / clusters);
    if (quota < sk_min) quota = sk_min;
    gp.sk_quota = quota;
    gp.sk_flags = static_cast<unsigned*>(ws);
    gp.sk_ws = reinterpret_cast<float*>(static_cast<char*>(ws) + SK_FLAG_BYTES);
    clusters = static_cast<int>((total + quota - 1) / quota);
  } else if (ws_ok && g_schedule == 1 && same_nk && num_tiles > 0) {
    // Hybrid: the full waves stay on the static round-robin (neighbouring pairs on neighbouring tiles: L2 sharing, one epilogue
    // per tile); only the LAST, partial wave -- R < pairs tiles that would keep R pairs busy for a whole tile while the others
    // idle -- is cut into equal k-ranges over all pairs.  Worth it only when it saves clearly more than a fold costs (partial
    // tile written, flag, partial tile read back: ~10 us on the critical path, measured: FPN layer forward, 11 iterations saved,
    // got 7 us slower; prop_fc weight gradient, 69 saved, 30 us faster) -> at least 24 iterations (DRN_HYBRID_MIN_SAVED).
    const int nk = gp.nk_tile[0];
    const int full = (num_tiles / clusters) * clusters, R = num_tiles - full;
    if (R > 0) {
      int quota = static_cast<int>((static_cast<long long>(R) * nk + clusters - 1) / clusters);
      if (quota < hybrid_min_quota) quota = hybrid_min_quota;  // every extra segment is one more partial tile to write and fold
      if (nk - quota >= hybrid_min_saved) {
        gp.sk_quota = quota;
        gp.sk_static_tiles = full;
        gp.sk_it0 = full * nk;
        gp.sk_flags = static_cast<unsigned*>(ws);
        gp.sk_ws = reinterpret_cast<float*>(static_cast<char*>(ws) + SK_FLAG_BYTES);
      }
    }
  } else if (clusters > num_tiles) {
    clusters = num_tiles;
  }
  if (gp.sk_quota == 0 && clusters > num_tiles) clusters = num_tiles;
  if (clusters < 1) clusters = 1;
  // ---- host-balanced static schedule (longest tiles first onto the least-loaded pair) --------------------------------------
  gp.lpt = 0;
  static int lpt_on = -1;
  if (lpt_on < 0) lpt_on = sk_env("DRN_LPT", 1);
  if (lpt_on && uniform && gp.sk_quota == 0 && gp.nprob > 1 && clusters <= LPT_MAX_PAIRS && num_tiles < 65536 && num_tiles > clusters) {
    int load[LPT_MAX_PAIRS] = {0};
    static thread_local unsigned short cnt[GROUP_MAX][LPT_MAX_PAIRS];
    bool differ = false, fits = true;
    for (int k = 0; k < gp.nprob; ++k) {
      if (gp.nk_tile[k] != gp.nk_tile[0]) differ = true;
      for (int c = 0; c < clusters; ++c) cnt[k][c] = 0;
    }
    if (differ) {  // uniform tile lengths: round-robin is already the balanced assignment
      int total_cnt[LPT_MAX_PAIRS] = {0};
      for (int k = 0; k < gp.nprob && fits; ++k) {  // problems arrive sorted by decreasing tile length (drn_gemm_group_ws)
        const int nt = gp.tile_start[k + 1] - gp.tile_start[k];
        int next = 0;  // ties go round-robin from the pair after the previous pick
        for (int i = 0; i < nt; ++i) {
          int best = -1;
          for (int j = 0; j < clusters; ++j) {
            const int c = (next + j) % clusters;
            if (best < 0 || load[c] < load[best]) best = c;
          }
          load[best] += gp.nk_tile[k];
          ++cnt[k][best];
          if (++total_cnt[best] > LPT_MAX_TILES) { fits = false; break; }
          next = (best + 1) % clusters;
        }
      }
      if (fits) {
        for (int c = 0; c < clusters; ++c) gp.lpt_count[c] = 0;
        for (int k = 0; k < gp.nprob; ++k) {
          int tile = gp.tile_start[k];
          for (int round = 0; tile < gp.tile_start[k + 1]; ++round)
            for (int c = 0; c < clusters; ++c)
              if (cnt[k][c] > round) gp.lpt_tiles[c][gp.lpt_count[c]++] = static_cast<unsigned short>(tile++);
        }
        gp.lpt = 1;
      }
    }
  }
  return clusters;
}

int launch_group(GroupParams& gp, const GroupMaps& gm, const int* nk_tile, int sm_count, void* ws, size_t ws_bytes, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaFuncSetAttribute(gemm_pair): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int clusters = plan_group(gp, nk_tile, sm_count, ws, ws_bytes);
  const int num_tiles = gp.tile_start[gp.nprob];
  static int epi_t = -1;  // DRN_EPI_T=0: row-per-thread epilogue stores (A/B)
  if (epi_t < 0) epi_t = sk_env("DRN_EPI_T", 1);
  gp.epi_t = epi_t;
  gp.trace = nullptr;
  if (g_trace_left > 0) {
    gp.trace = g_trace + static_cast<size_t>(g_trace_next) * TRACE_CTAS * 8;
    g_trace_info[g_trace_next][0] = 2 * clusters;
    g_trace_info[g_trace_next][1] = num_tiles;
    g_trace_info[g_trace_next][2] = gp.lpt;
    ++g_trace_next;
    --g_trace_left;
  }
  launch_k(gemm_pair_kernel, 2 * clusters, P2_THREADS, P2_SMEM, st, gp, gm, num_tiles);
  return check_launch("gemm_pair_kernel");
}


// Planner probe (no GPU needed): `nprob` problems with tiles[k] tiles of nk[k] k-iterations each, in the order given (the
// launcher sorts by decreasing nk), on `pairs` SM pairs, with (has_ws != 0) or without a workspace, under schedule `mode`
// (0 static / 1 hybrid / 2 stream-K; < 0 = leave the process setting).  Returns the pairs to launch; kind = 0 round-robin,
// 1 host-balanced lists, 2 hybrid, 3 stream-K; quota / static_tiles = range length and whole-tile prefix of the k-split
// schedules; counts[pair] and lists[pair * 16 + i] = the tile list of a pair (kind 1).
int schedule_probe(int nprob, const int* tiles, const int* nk, int pairs, int has_ws, int mode, int* kind, int* quota,
                   int* static_tiles, unsigned char* counts, unsigned short* lists) {
  if (nprob < 1 || nprob > GROUP_MAX || pairs < 1) return -1;
  static thread_local GroupParams gp;
  gp.nprob = nprob;
  gp.tile_start[0] = 0;
  for (int k = 0; k < nprob; ++k) gp.tile_start[k + 1] = gp.tile_start[k] + tiles[k];
  for (int k = nprob; k < GROUP_MAX; ++k) gp.tile_start[k + 1] = gp.tile_start[nprob];
  const int saved = g_schedule;
  if (mode >= 0) set_schedule(mode);
  static char fake_ws[64];
  const int clusters = plan_group(gp, nk, 2 * pairs, has_ws ? fake_ws : nullptr, has_ws ? (static_cast<size_t>(1) << 40) : 0);
  const int effective = g_schedule;  // (plan_group resolves the environment default on first use)
  if (mode >= 0) g_schedule = saved;
  if (kind) *kind = gp.lpt ? 1 : (gp.sk_quota > 0 ? (effective == 2 ? 3 : 2) : 0);
  if (quota) *quota = gp.sk_quota;
  if (static_tiles) *static_tiles = gp.sk_static_tiles;
  if (gp.lpt && counts && lists) {
    for (int c = 0; c < clusters && c < LPT_MAX_PAIRS; ++c) {
      counts[c] = gp.lpt_count[c];
      for (int i = 0; i < gp.lpt_count[c]; ++i) lists[c * LPT_MAX_TILES + i] = gp.lpt_tiles[c][i];
    }
  }
  return clusters;
}

}  // namespace drn
