// drn_gemm: split-BF16 (hi/lo planes, 3 products) tensor-core contraction on tcgen05 / TMEM fed by 5-D TMA tiles.
//
// One kernel serves every dense layer of the DRN hot path (include/drn_b200.h):
//   ROWS  : temporal conv / linear forward and data gradient.  A tile = 128 output positions x 64 channels, fetched
//           per tap with a time shift; conv zero padding and ragged tiles come from TMA out-of-bounds zero fill on the
//           (c, parity, t, b, plane) tensor map, so samples never leak into each other.
//   WGRAD : weight gradient; both operands are "MN-major" (the contraction runs over (b,t) rows), read straight from
//           the channels-last planes with MN-major UMMA descriptors -- no transposes are materialised.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = epilogue (TMEM -> registers -> global, one accumulator row per thread).
#include <cuda.h>
#include <stdlib.h>
#include <cudaTypedefs.h>

#include "gemm_common.cuh"

namespace drn {

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel: one 128 x BLOCK_N output tile per CTA.
// ------------------------------------------------------------------------------------------------
template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmKParams p, const __grid_constant__ CUtensorMap tma_a,
               const __grid_constant__ CUtensorMap tma_b) {
  pdl_trigger();
  using Cfg = TileCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t empty_bar[Cfg::STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_holder;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool wgrad = (p.form == DRN_GEMM_WGRAD);
  const bool a_mn = wgrad;
  const bool b_mn = wgrad || (p.b_mn != 0);
  const int nplanes = (p.nprod == 1) ? 1 : 2;

  // ---- tile coordinates -----------------------------------------------------------------------
  const int n0 = blockIdx.y * BLOCK_N;
  int b0 = 0, t0 = 0, m0 = 0, tap_fixed = 0, it_begin = 0, it_end = 0;
  const int kpt = p.K / BLOCK_K;  // k-blocks per tap (ROWS)
  if (!wgrad) {
    const int mt = blockIdx.x;
    if (p.Bbm == 1) {
      b0 = mt / p.tiles_per_sample;
      t0 = (mt % p.tiles_per_sample) * p.Rm;
    } else {
      b0 = mt * p.Bbm;
    }
    it_end = p.ntaps * kpt;
  } else {
    m0 = blockIdx.x * BLOCK_M;
    tap_fixed = blockIdx.z / p.split_k;
    const int split = blockIdx.z % p.split_k;
    it_begin = static_cast<int>(static_cast<long long>(p.num_kblocks) * split / p.split_k);
    it_end = static_cast<int>(static_cast<long long>(p.num_kblocks) * (split + 1) / p.split_k);
  }
  const int nk = it_end - it_begin;

  // ---- one-time setup -------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(&tmem_base_holder), BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // barriers initialised, TMEM allocated: only now wait for the kernel before this one (its output = our operands)
  const uint32_t tmem_base = tmem_base_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && nk > 0) {
      const uint32_t tx = nplanes * (A_PLANE + Cfg::B_PLANE);
      for (int i = 0; i < nk; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1;
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_arrive_expect_tx(fb, tx);
        const uint32_t sa = smem_base + s * Cfg::STAGE;
        const uint32_t sb = sa + 2 * A_PLANE;
        const int it = it_begin + i;
        if (!wgrad) {
          const int tap = it / kpt;
          const int kb = it % kpt;
          for (int pl = 0; pl < nplanes; ++pl) {
            tma_load_5d(sa + pl * A_PLANE, &tma_a, fb, p.a_c0 + kb * BLOCK_K, p.tap_par[tap], t0 + p.tap_shift[tap], b0,
                        pl);
            if (!b_mn) {
              tma_load_5d(sb + pl * Cfg::B_PLANE, &tma_b, fb, p.b_c0 + kb * BLOCK_K, 0, n0, p.tap_w[tap], pl);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_N / 64; ++j)
                tma_load_5d(sb + pl * Cfg::B_PLANE + j * 8192, &tma_b, fb, p.b_c0 + n0 + j * 64, 0, kb * BLOCK_K,
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
            for (int j = 0; j < 2; ++j)
              tma_load_5d(sa + pl * A_PLANE + j * 8192, &tma_a, fb, p.a_c0 + m0 + j * 64, 0, tk, bk, pl);
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_5d(sb + pl * Cfg::B_PLANE + j * 8192, &tma_b, fb, p.b_c0 + n0 + j * 64, p.tap_par[tap_fixed],
                          tk + p.tap_shift[tap_fixed], bk, pl);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && nk > 0) {
      const uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, a_mn, b_mn);
      const uint32_t mn_lbo = p.dbg_lbo ? p.dbg_lbo : 8192u;
      const uint32_t mn_sbo = p.dbg_sbo ? p.dbg_sbo : 1024u;
      const uint32_t mn_kadv = p.dbg_kadv ? p.dbg_kadv : 2048u;
      const uint32_t a_lbo = a_mn ? mn_lbo : 0u, a_sbo = a_mn ? mn_sbo : 1024u, a_kadv = a_mn ? mn_kadv : 32u;
      const uint32_t b_lbo = b_mn ? mn_lbo : 0u, b_sbo = b_mn ? mn_sbo : 1024u, b_kadv = b_mn ? mn_kadv : 32u;
      uint32_t accumulate = 0;
      for (int i = 0; i < nk; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (i / Cfg::STAGES) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * Cfg::STAGE;
        const uint32_t sb = sa + 2 * A_PLANE;
        for (int prod = 0; prod < p.nprod; ++prod) {
          // products: (hi,hi) (hi,lo) (lo,hi) (lo,lo)
          const uint32_t pa = (prod >> 1) & 1, pb = prod & 1;
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t ad = umma_smem_desc(sa + pa * A_PLANE + k * a_kadv, a_lbo, a_sbo);
            const uint64_t bd = umma_smem_desc(sb + pb * Cfg::B_PLANE + k * b_kadv, b_lbo, b_sbo);
            umma_bf16(tmem_base, ad, bd, idesc, accumulate);
            accumulate = 1;
          }
        }
        umma_commit(smem_u32(&empty_bar[s]));  // frees the smem slot when these MMAs retire
      }
      umma_commit(smem_u32(&tmem_full_bar));
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
    if (nk > 0) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      mbar_wait(smem_u32(&tmem_full_bar), 0);
      tc_fence_after();
      bool valid;
      long long orow;
      int bb = 0;
      float* out_base = p.out;
      if (!wgrad) {
        bb = b0 + row / p.Rm;
        const int tt = t0 + row % p.Rm;
        valid = (bb < p.B) && (tt < p.T);
        orow = static_cast<long long>(bb) * p.out_T + static_cast<long long>(tt) * p.out_t_mul + p.out_t_add;
      } else {
        valid = (m0 + row) < p.M;
        orow = m0 + row;
        if (out_base) out_base += p.tap_w[tap_fixed] * p.out_tap_stride;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        if (n0 + c0 >= p.N) break;
        float v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        epilogue_chunk(p, v, valid, orow, bb, n0 + c0, out_base);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core checker kernel (engine 1): same operands, same epilogue, plain loops.  Used by tests to validate
// the tensor-core path on device at sizes where a host reference would be slow.  Never used by the model.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fetch(const PlanesView& v, int b, int t, int par, int c, bool use_lo) {
  if (b < 0 || b >= v.B || t < 0 || t >= v.T || c < 0 || c >= v.C) return 0.f;
  const long long idx = ((static_cast<long long>(b) * v.T + t) * v.P + par) * v.C + c;
  float x = __bfloat162float(v.ptr[idx]);
  if (use_lo) x += __bfloat162float(v.ptr[idx + v.plane_stride]);
  return x;
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmKParams p) {
  pdl_sync();
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const bool wgrad = (p.form == DRN_GEMM_WGRAD);
  const bool use_lo = p.nprod > 1;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m_base = blockIdx.x * 64, n_base = blockIdx.y * 64;
  const int tap_fixed = blockIdx.z;
  const int Mtot = wgrad ? p.M : p.B * p.T;
  const long long Ktot = wgrad ? static_cast<long long>(p.B) * p.T : static_cast<long long>(p.ntaps) * p.K;
  float acc[4][4] = {};
  for (long long k0 = 0; k0 < Ktot; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      const int kk = e >> 6, mm = e & 63;
      const long long k = k0 + kk;
      float av = 0.f, bv = 0.f;
      if (k < Ktot) {
        if (!wgrad) {
          const int tap = static_cast<int>(k / p.K), kc = static_cast<int>(k % p.K);
          const int m = m_base + mm;
          if (m < Mtot) av = fetch(p.a, m / p.T, m % p.T + p.tap_shift[tap], p.tap_par[tap], p.a_c0 + kc, use_lo);
          const int n = n_base + mm;
          if (n < p.N)
            bv = p.b_mn ? fetch(p.b, p.tap_w[tap], kc, 0, p.b_c0 + n, use_lo)
                        : fetch(p.b, p.tap_w[tap], n, 0, p.b_c0 + kc, use_lo);
        } else {
          const int b = static_cast<int>(k / p.T), t = static_cast<int>(k % p.T);
          const int m = m_base + mm;
          if (m < Mtot) av = fetch(p.a, b, t, 0, p.a_c0 + m, use_lo);
          const int n = n_base + mm;
          if (n < p.N) bv = fetch(p.b, b, t + p.tap_shift[tap_fixed], p.tap_par[tap_fixed], p.b_c0 + n, use_lo);
        }
      }
      As[kk][mm] = av;
      Bs[kk][mm] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int m = m_base + ty * 4 + i;
    if (m >= Mtot) continue;
    long long orow;
    int bb = 0;
    float* out_base = p.out;
    if (!wgrad) {
      bb = m / p.T;
      const int tt = m % p.T;
      orow = static_cast<long long>(bb) * p.out_T + static_cast<long long>(tt) * p.out_t_mul + p.out_t_add;
    } else {
      orow = m;
      if (out_base) out_base += p.tap_w[tap_fixed] * p.out_tap_stride;
    }
    for (int j = 0; j < 4; ++j) {
      const int n = n_base + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.out2) p.out2[orow * p.out2_ld + n] = v;
      if (p.rowscale) v *= p.rowscale[static_cast<long long>(bb) * p.rowscale_ld + n];
      if (out_base) {
        float* o = out_base + orow * p.out_ld + p.out_col0 + n;
        if (p.out_mode == DRN_OUT_ATOMIC) atomicAdd(o, v);
        else if (p.out_mode == DRN_OUT_ADD) *o += v;
        else *o = v;
      }
      if (p.outp) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        __nv_bfloat16* oh = p.outp + orow * p.outp_ld + p.outp_col0 + n;
        oh[0] = h;
        oh[p.outp_plane_stride] = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  }
  return fn;
}

// 5-D map over (c, parity, t, b, plane) with a {64, 1, box_t, box_b, 1} box and 128-byte swizzle.
static int make_map(CUtensorMap* m, const drn_planes_t& v, int box_t, int box_b) {
  auto enc = get_encode();
  if (!enc) return fail(DRN_EDRIVER, "cuTensorMapEncodeTiled entry point not available");
  if (v.C % 8 != 0 || (reinterpret_cast<uintptr_t>(v.ptr) & 15) != 0 || (v.plane_stride % 8) != 0)
    return fail(DRN_EINVAL, "planes tensor must have C %% 8 == 0 and 16-byte aligned planes (C=%d)", v.C);
  cuuint64_t dims[5] = {static_cast<cuuint64_t>(v.C), static_cast<cuuint64_t>(v.P), static_cast<cuuint64_t>(v.T),
                        static_cast<cuuint64_t>(v.B), 2};
  cuuint64_t strides[4] = {static_cast<cuuint64_t>(v.C) * 2, static_cast<cuuint64_t>(v.P) * v.C * 2,
                           static_cast<cuuint64_t>(v.T) * v.P * v.C * 2, static_cast<cuuint64_t>(v.plane_stride) * 2};
  cuuint32_t box[5] = {64, 1, static_cast<cuuint32_t>(box_t), static_cast<cuuint32_t>(box_b), 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(DRN_EINVAL, "cuTensorMapEncodeTiled failed (%d) for C=%d P=%d T=%d B=%d box_t=%d box_b=%d", (int)r, v.C,
                v.P, v.T, v.B, box_t, box_b);
  return 0;
}

static int pow2_ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

template <int BLOCK_N>
static int launch_tc(const GemmKParams& kp, const CUtensorMap& ma, const CUtensorMap& mb, dim3 grid, cudaStream_t st) {
  using Cfg = TileCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_k(gemm_tc_kernel<BLOCK_N>, grid, GEMM_THREADS, Cfg::SMEM, st, kp, ma, mb);
  return check_launch("gemm_tc_kernel");
}

int launch_group(GroupParams& gp, const GroupMaps& gm, const int* nk_tile, int sm_count, void* ws, size_t ws_bytes,
                 cudaStream_t st);  // gemm2.cu

static int sm_count_cached() {
  static int sm_count = 0;
  if (!sm_count) {
    sm_count = drn_sm_count();
    if (sm_count <= 0) sm_count = 148;
  }
  return sm_count;
}

struct Prepared {
  GemmKParams kp;
  bool wgrad;
  int m_sub;                                 // 128-row sub-tiles (ROWS) / 128-channel sub-tiles (WGRAD)
  int pair_n_tiles, pair_m_tiles, pair_tiles;  // 256 x 256 tiling of the CTA-pair kernel
  int cost;                                  // k-iterations per pair tile (load-balance key)
  int nk_uniform;                            // k-iterations of EVERY tile, or 0 when the K-split does not divide evenly
};

// Validate a descriptor and derive the kernel-side parameters (no tensor maps yet).
static int prepare(const drn_gemm_t* g, Prepared* out) {
  if (!g) return fail(DRN_EINVAL, "drn_gemm: null descriptor");
  const bool wgrad = g->form == DRN_GEMM_WGRAD;
  if (g->form != DRN_GEMM_ROWS && !wgrad) return fail(DRN_EINVAL, "drn_gemm: unknown form %d", g->form);
  if (g->nprod != 1 && g->nprod != 3 && g->nprod != 4) return fail(DRN_EINVAL, "drn_gemm: nprod must be 1, 3 or 4");
  if (g->ntaps < 1 || g->ntaps > DRN_MAX_TAPS) return fail(DRN_EINVAL, "drn_gemm: ntaps %d", g->ntaps);
  if (g->B < 1 || g->T < 1 || g->N < 1) return fail(DRN_EINVAL, "drn_gemm: empty problem");
  if (!wgrad && (g->K < 64 || g->K % 64 != 0)) return fail(DRN_EINVAL, "drn_gemm: K=%d must be a multiple of 64", g->K);
  if (wgrad && g->M < 1) return fail(DRN_EINVAL, "drn_gemm: wgrad needs M");
  if (!g->out && !g->out2 && !g->outp) return fail(DRN_EINVAL, "drn_gemm: no output");

  GemmKParams kp{};
  kp.form = g->form;
  kp.b_mn = g->b_mn;
  kp.B = g->B; kp.T = g->T; kp.N = g->N; kp.K = g->K; kp.M = g->M;
  kp.ntaps = g->ntaps;
  for (int i = 0; i < DRN_MAX_TAPS; ++i) {
    kp.tap_shift[i] = g->tap_shift[i];
    kp.tap_par[i] = g->tap_par[i];
    kp.tap_w[i] = g->tap_w[i];
  }
  kp.a_c0 = g->a_c0; kp.b_c0 = g->b_c0;
  kp.nprod = g->nprod;
  kp.split_k = g->split_k < 1 ? 1 : g->split_k;
  kp.out = g->out; kp.out_ld = g->out_ld; kp.out_col0 = g->out_col0; kp.out_mode = g->out_mode;
  kp.out_tap_stride = g->out_tap_stride;
  kp.out_split_stride = g->out_split_stride;
  kp.out_T = g->out_T; kp.out_t_mul = g->out_t_mul; kp.out_t_add = g->out_t_add;
  kp.bias = g->bias; kp.rowscale = g->rowscale; kp.rowscale_ld = g->rowscale_ld;
  kp.out2 = g->out2; kp.out2_ld = g->out2_ld;
  kp.outp = static_cast<__nv_bfloat16*>(g->outp); kp.outp_ld = g->outp_ld; kp.outp_col0 = g->outp_col0;
  kp.outp_plane_stride = g->outp_plane_stride;
  kp.a = PlanesView{static_cast<const __nv_bfloat16*>(g->a.ptr), g->a.plane_stride, g->a.B, g->a.T, g->a.P, g->a.C};
  kp.b = PlanesView{static_cast<const __nv_bfloat16*>(g->b.ptr), g->b.plane_stride, g->b.B, g->b.T, g->b.P, g->b.C};
  kp.dbg_lbo = g->dbg_lbo; kp.dbg_sbo = g->dbg_sbo; kp.dbg_kadv = g->dbg_kadv;
  kp.stats = g->stats;
  if (kp.stats && (wgrad || kp.split_k > 1))
    return fail(DRN_EINVAL, "drn_gemm: stats (fused BatchNorm partial sums) need the ROWS form without a K-split");
  if (!wgrad && kp.split_k > 1 && kp.out_split_stride == 0)
    return fail(DRN_EINVAL, "drn_gemm: a ROWS problem splits K only into slices (out_split_stride)");
  if (kp.split_k > 1 && kp.out_split_stride == 0 && kp.out_mode != DRN_OUT_ATOMIC)
    return fail(DRN_EINVAL, "drn_gemm: split_k needs an atomic output or out_split_stride");
  if (kp.split_k > 1 && (kp.out2 || kp.outp || kp.bias || kp.rowscale))
    return fail(DRN_EINVAL, "drn_gemm: split_k supports only the plain output");

  bool vec = true, vec8 = true;
  auto al = [](const void* p, uintptr_t a) { return reinterpret_cast<uintptr_t>(p) % a == 0; };
  if (kp.out) {
    vec = vec && al(kp.out, 16) && (kp.out_ld % 4 == 0) && (kp.out_col0 % 4 == 0) && (kp.out_tap_stride % 4 == 0) &&
          (kp.out_split_stride % 4 == 0);
    vec8 = vec8 && al(kp.out, 32) && (kp.out_ld % 8 == 0) && (kp.out_col0 % 8 == 0) && (kp.out_tap_stride % 8 == 0) &&
           (kp.out_split_stride % 8 == 0);
  }
  if (kp.out2) {
    vec = vec && al(kp.out2, 16) && (kp.out2_ld % 4 == 0);
    vec8 = vec8 && al(kp.out2, 32) && (kp.out2_ld % 8 == 0);
  }
  if (kp.outp) vec = vec && al(kp.outp, 16) && (kp.outp_ld % 8 == 0) && (kp.outp_col0 % 8 == 0) && (kp.outp_plane_stride % 8 == 0);
  kp.vec_ok = vec ? 1 : 0;
  kp.vec8_ok = (vec && vec8) ? 1 : 0;

  if (!wgrad) {
    if (g->T >= 128) { kp.Rm = 128; kp.Bbm = 1; }
    else { kp.Rm = pow2_ceil(g->T); kp.Bbm = 128 / kp.Rm; }
    kp.tiles_per_sample = ceil_div(g->T, kp.Rm);
    if (kp.split_k > g->ntaps * (g->K / BLOCK_K)) return fail(DRN_EINVAL, "drn_gemm: split_k %d exceeds the K-blocks", kp.split_k);
  } else {
    if (g->T >= 64) { kp.Rk = 64; kp.Bbk = 1; }
    else { kp.Rk = pow2_ceil(g->T); kp.Bbk = 64 / kp.Rk; }
    kp.kblocks_per_sample = ceil_div(g->T, kp.Rk);
    kp.num_kblocks = (kp.Bbk == 1) ? g->B * kp.kblocks_per_sample : ceil_div(g->B, kp.Bbk);
    if (kp.split_k > kp.num_kblocks) {
      if (kp.out_split_stride != 0) return fail(DRN_EINVAL, "drn_gemm: split_k %d exceeds the %d K-blocks (slices would stay unwritten)", kp.split_k, kp.num_kblocks);
      kp.split_k = kp.num_kblocks;
    }
  }
  out->kp = kp;
  out->wgrad = wgrad;
  out->m_sub = wgrad ? ceil_div(g->M, BLOCK_M) : ((kp.Bbm == 1) ? g->B * kp.tiles_per_sample : ceil_div(g->B, kp.Bbm));
  out->pair_n_tiles = ceil_div(g->N, 256);
  out->pair_m_tiles = ceil_div(out->m_sub, 2);
  out->pair_tiles = out->pair_m_tiles * out->pair_n_tiles * (wgrad ? g->ntaps * kp.split_k : kp.split_k);
  const int its = wgrad ? kp.num_kblocks : g->ntaps * (g->K / BLOCK_K);
  out->cost = ceil_div(its, kp.split_k);
  out->nk_uniform = (its % kp.split_k == 0) ? its / kp.split_k : 0;
  return 0;
}

static int pair_maps(const drn_gemm_t* g, const Prepared& pr, CUtensorMap* ma, CUtensorMap* mb) {
  int rc;
  if (!pr.wgrad) {
    if ((rc = make_map(ma, g->a, pr.kp.Rm, pr.kp.Bbm)) != 0) return rc;
    return make_map(mb, g->b, g->b_mn ? 64 : 128, 1);
  }
  if ((rc = make_map(ma, g->a, pr.kp.Rk, pr.kp.Bbk)) != 0) return rc;
  return make_map(mb, g->b, pr.kp.Rk, pr.kp.Bbk);
}

}  // namespace drn

using namespace drn;

// One launch for up to GROUP_MAX independent problems (persistent CTA-pair kernel).  Tiles are ordered by decreasing
// k-iterations per tile so the static round-robin over the 74 SM pairs ends with the cheapest tiles.
extern "C" size_t drn_gemm_workspace_bytes(void) {
  return SK_FLAG_BYTES + static_cast<size_t>(sm_count_cached() / 2) * SK_SLOT_FLOATS * sizeof(float);
}

extern "C" int drn_gemm_group(int n, const drn_gemm_t* descs, void* stream) {
  return drn_gemm_group_ws(n, descs, nullptr, 0, stream);
}

extern "C" int drn_gemm_group_ws(int n, const drn_gemm_t* descs, void* workspace, size_t workspace_bytes, void* stream) {
  if (n < 1 || n > GROUP_MAX) return fail(DRN_EINVAL, "drn_gemm_group: 1..%d problems (got %d)", GROUP_MAX, n);
  if (!descs) return fail(DRN_EINVAL, "drn_gemm_group: null descriptors");
  Prepared pr[GROUP_MAX];
  int order[GROUP_MAX];
  int rc;
  for (int i = 0; i < n; ++i) {
    if (descs[i].engine == 1 || descs[i].engine == 3) return fail(DRN_EINVAL, "drn_gemm_group: problems run on the CTA-pair engine only");
    if ((rc = prepare(&descs[i], &pr[i])) != 0) return rc;
    order[i] = i;
  }
  for (int i = 1; i < n; ++i)  // insertion sort, descending cost, stable
    for (int j = i; j > 0 && pr[order[j]].cost > pr[order[j - 1]].cost; --j) {
      const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t;
    }
  static thread_local GroupParams gp;
  static thread_local GroupMaps gm;
  gp.nprob = n;
  {
    static int gm = -1;  // DRN_RASTER_GM: tuning override
    if (gm < 0) {
      const char* e = getenv("DRN_RASTER_GM");
      gm = e ? atoi(e) : 0;
    }
    gp.raster_gm = gm > 0 ? gm : 1;  // A/B-measured r01: 1 / 4 / 8 / 16 within noise (DRAM is at ~18 % while the tensor pipe is at ~90 %)
  }
  int nk_tile[GROUP_MAX];
  gp.tile_start[0] = 0;
  for (int k = 0; k < n; ++k) {
    const int i = order[k];
    nk_tile[k] = pr[i].nk_uniform;
    gp.p[k] = pr[i].kp;
    gp.n_tiles[k] = pr[i].pair_n_tiles;
    gp.m_tiles[k] = pr[i].pair_m_tiles;
    gp.tile_start[k + 1] = gp.tile_start[k] + pr[i].pair_tiles;
    if ((rc = pair_maps(&descs[i], pr[i], &gm.a[k], &gm.b[k])) != 0) return rc;
  }
  for (int k = n; k < GROUP_MAX; ++k) gp.tile_start[k + 1] = gp.tile_start[n];
  if (workspace && (reinterpret_cast<uintptr_t>(workspace) & 31) != 0) return fail(DRN_EINVAL, "drn_gemm_group_ws: workspace must be 32-byte aligned");
  return launch_group(gp, gm, nk_tile, sm_count_cached(), workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

namespace drn {
void set_pair_clusters(int n);
void set_schedule(int mode);
int schedule_probe(int nprob, const int* tiles, const int* nk, int pairs, int has_ws, int mode, int* kind, int* quota,
                   int* static_tiles, unsigned char* counts, unsigned short* lists);
void gemm_trace(unsigned long long* buf, int launches);
int gemm_trace_info(int launch, int* ctas, int* tiles, int* lpt);
}
extern "C" void drn_set_pair_clusters(int n) { drn::set_pair_clusters(n); }
extern "C" void drn_gemm_set_schedule(int mode) { drn::set_schedule(mode); }
extern "C" int drn_gemm_schedule_probe(int nprob, const int* tiles, const int* nk, int pairs, int has_ws, int mode, int* kind,
                                       int* quota, int* static_tiles, unsigned char* counts, unsigned short* lists) {
  return drn::schedule_probe(nprob, tiles, nk, pairs, has_ws, mode, kind, quota, static_tiles, counts, lists);
}
extern "C" void drn_gemm_trace(uint64_t* buf, int launches) { drn::gemm_trace(reinterpret_cast<unsigned long long*>(buf), launches); }
extern "C" int drn_gemm_trace_info(int launch, int* ctas, int* tiles, int* lpt) { return drn::gemm_trace_info(launch, ctas, tiles, lpt); }

extern "C" int drn_gemm_stats_rows(const drn_gemm_t* g) {
  Prepared pr;
  int rc = prepare(g, &pr);
  if (rc) return rc;
  if (pr.wgrad) return fail(DRN_EINVAL, "drn_gemm_stats_rows: ROWS form only");
  return 8 * pr.pair_m_tiles;  // 2 CTAs x 4 epilogue warps x 32 rows per 256-row pair tile
}

extern "C" int drn_gemm(const drn_gemm_t* g, void* stream) { return drn_gemm_ws(g, nullptr, 0, stream); }

extern "C" int drn_gemm_ws(const drn_gemm_t* g, void* workspace, size_t workspace_bytes, void* stream) {
  Prepared pr;
  int rc = prepare(g, &pr);
  if (rc) return rc;
  if (g->stats && g->engine != 2) return fail(DRN_EINVAL, "drn_gemm: stats are written by the CTA-pair kernel only (engine 2 / drn_gemm_group)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GemmKParams& kp = pr.kp;
  const bool wgrad = pr.wgrad;

  if (g->engine == 1) {
    if (kp.out_split_stride != 0 && kp.split_k > 1) return fail(DRN_EINVAL, "drn_gemm: the checker engine has no K-split");
    const int Mtot = wgrad ? g->M : g->B * g->T;
    dim3 grid(ceil_div(Mtot, 64), ceil_div(g->N, 64), wgrad ? g->ntaps : 1);
    launch_k(gemm_simt_kernel, grid, 256, 0, st, kp);
    return check_launch("gemm_simt_kernel");
  }

  // ---- engine selection: 0 = auto, 2 = persistent CTA-pair kernel (gemm2.cu), 3 = one-tile-per-CTA kernel -------------
  const int sm_count = sm_count_cached();
  bool use_pair = (g->engine == 2) || (g->engine == 0 && pr.pair_tiles >= sm_count / 4);
  if (g->engine == 3) use_pair = false;
  if (g->dbg_lbo || g->dbg_sbo || g->dbg_kadv) use_pair = false;
  if (use_pair) return drn_gemm_group_ws(1, g, workspace, workspace_bytes, stream);
  if (kp.split_k > 1 && (kp.out_split_stride != 0 || !wgrad)) return drn_gemm_group_ws(1, g, workspace, workspace_bytes, stream);  // slices: pair kernel only

  // one tile per CTA: 128 x 256 tiles unless that leaves most SMs idle, then 128 x 128
  int block_n = (g->N > 128) ? 256 : 128;
  if (block_n == 256 && pr.m_sub * ceil_div(g->N, 256) * (wgrad ? g->ntaps * kp.split_k : 1) < (sm_count * 2) / 3) block_n = 128;
  CUtensorMap ma, mb;
  dim3 grid;
  if (!wgrad) {
    if ((rc = make_map(&ma, g->a, kp.Rm, kp.Bbm)) != 0) return rc;
    if ((rc = make_map(&mb, g->b, g->b_mn ? 64 : block_n, 1)) != 0) return rc;
    grid = dim3(pr.m_sub, ceil_div(g->N, block_n), 1);
  } else {
    if ((rc = make_map(&ma, g->a, kp.Rk, kp.Bbk)) != 0) return rc;
    if ((rc = make_map(&mb, g->b, kp.Rk, kp.Bbk)) != 0) return rc;
    grid = dim3(ceil_div(g->M, BLOCK_M), ceil_div(g->N, block_n), g->ntaps * kp.split_k);
  }
  if (block_n == 256) return launch_tc<256>(kp, ma, mb, grid, st);
  return launch_tc<128>(kp, ma, mb, grid, st);
}
