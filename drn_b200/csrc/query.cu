// Query encoder of the DRN path (reference model/language_module.py:27-62, model/ops.py:16-25,74-85), forward and
// hand-derived backward, plus the small fp32 CUDA-core contraction (drn_sgemm) it and the query gates are built from.
//
// The encoder is 0.6 % of the path's FLOPs but, as ~300 library launches (cuDNN RNN + cuBLAS + ATen), it was 27 % of the
// step.  Here it is ~20 launches forward and ~45 backward, all graph-capturable, all exact fp32 FMA arithmetic:
//   embedding gather -> input projection (one contraction for all time steps) -> L recurrent steps of a masked
//   ("packed") BiLSTM, one launch per step for both directions -> q_vector gather -> qInput / qInput0..2 -> attention
//   over the words (mask -1e30, softmax) -> the three command vectors; backward = the same chain reversed (BPTT).
// Layouts: R = B*L rows (b-major); xg / G / dG [R][2 dirs][4 gates i,f,g,o][H]; Hout [R][2H]; Cst [R][2][H].
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace drn {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------------------
// Small fp32 contraction: C[m][n] (=|+=) sum_k A(m,k) * B(k,n) (+ bias[n] + bias2[n]) with arbitrary element strides, so the
// same kernel serves x W^T (Linear forward), dy W (data gradient) and dy^T x (weight gradient).  64(32) x 64 x 16 tiles,
// 256 threads, 4(2) x 4 outputs per thread; split-K over blockIdx.z with atomic accumulation.
// ------------------------------------------------------------------------------------------------------------------------
struct SgemmP {
  const float* A;
  long long sam, sak;
  const float* B;
  long long sbk, sbn;
  float* C;
  long long ldc;
  int M, N, K;
  const float* bias;
  const float* bias2;
  int relu, atomic, kchunk;
};

template <int BM>
__device__ __forceinline__ void sgemm_body(const SgemmP& p, int bx, int by, int bz) {
  constexpr int BN = 64, BK = 16, TM = BM / 16;
  __shared__ float As[BK][BM + 1];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = bx * BM, n0 = by * BN;
  const int kbeg = bz * p.kchunk;
  const int kend = min(p.K, kbeg + p.kchunk);
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (p.sak == 1);
  const bool b_nfast = (p.sbn == 1);
  constexpr int AE = (BM * BK) / 256, BE = (BN * BK) / 256;
  float areg[AE], breg[BE];
  // global -> registers for the tile starting at k0 (issued one tile ahead of the math: software pipelining)
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < AE; ++e) {
      const int idx = tid + e * 256;
      int m, k;
      if (a_kfast) { k = idx % BK; m = idx / BK; }
      else { m = idx % BM; k = idx / BM; }
      const int gm = m0 + m, gk = k0 + k;
      areg[e] = (gm < p.M && gk < kend) ? __ldg(p.A + gm * p.sam + gk * p.sak) : 0.f;
    }
#pragma unroll
    for (int e = 0; e < BE; ++e) {
      const int idx = tid + e * 256;
      int n, k;
      if (b_nfast) { n = idx % BN; k = idx / BN; }
      else { k = idx % BK; n = idx / BK; }
      const int gn = n0 + n, gk = k0 + k;
      breg[e] = (gn < p.N && gk < kend) ? __ldg(p.B + gk * p.sbk + gn * p.sbn) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int e = 0; e < AE; ++e) {
      const int idx = tid + e * 256;
      int m, k;
      if (a_kfast) { k = idx % BK; m = idx / BK; }
      else { m = idx % BM; k = idx / BM; }
      As[k][m] = areg[e];
    }
#pragma unroll
    for (int e = 0; e < BE; ++e) {
      const int idx = tid + e * 256;
      int n, k;
      if (b_nfast) { n = idx % BN; k = idx / BN; }
      else { k = idx % BK; n = idx / BK; }
      Bs[k][n] = breg[e];
    }
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        acc[i][0] = fmaf(a[i], b4.x, acc[i][0]);
        acc[i][1] = fmaf(a[i], b4.y, acc[i][1]);
        acc[i][2] = fmaf(a[i], b4.z, acc[i][2]);
        acc[i][3] = fmaf(a[i], b4.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= p.N) continue;
      float v = acc[i][j];
      if (bz == 0) {
        if (p.bias) v += __ldg(p.bias + gn);
        if (p.bias2) v += __ldg(p.bias2 + gn);
      }
      float* o = p.C + gm * p.ldc + gn;
      if (p.atomic) atomicAdd(o, v);
      else *o = p.relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// Rank-K update with K <= 32 (the weight gradients of every nn.Linear of the query encoder and of the gates: K = the batch):
// C[M][N] (=, +=) sum_k A(m,k) B(k,n).  The general body above walks K in 16-deep tiles behind a global-memory round trip each
// and writes a 64 x 64 tile per CTA: for K = 32 it is pure latency (50 us for the 17 MB of gate weight gradients, ncu r01).
// Here both operand slices (32 x 64 and 32 x 128) are staged ONCE, a thread owns a 4 x 8 patch and the CTA streams out a
// 64 x 128 tile with 16-byte stores: HBM-write bound.
constexpr int OU_BM = 64, OU_BN = 128, OU_K = 32;
__device__ __forceinline__ void outer_body(const SgemmP& p, int bx, int by) {
  __shared__ __align__(16) float As[OU_K][OU_BM];
  __shared__ __align__(16) float Bs[OU_K][OU_BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = bx * OU_BM, n0 = by * OU_BN;
  const bool a_kfast = (p.sak == 1), b_nfast = (p.sbn == 1);
#pragma unroll
  for (int e = 0; e < (OU_BM * OU_K) / 256; ++e) {
    const int idx = tid + e * 256;
    int m, k;
    if (a_kfast) { k = idx % OU_K; m = idx / OU_K; }
    else { m = idx % OU_BM; k = idx / OU_BM; }
    As[k][m] = (m0 + m < p.M && k < p.K) ? __ldg(p.A + (m0 + m) * p.sam + k * p.sak) : 0.f;
  }
#pragma unroll
  for (int e = 0; e < (OU_BN * OU_K) / 256; ++e) {
    const int idx = tid + e * 256;
    int n, k;
    if (b_nfast) { n = idx % OU_BN; k = idx / OU_BN; }
    else { k = idx % OU_K; n = idx / OU_K; }
    Bs[k][n] = (n0 + n < p.N && k < p.K) ? __ldg(p.B + k * p.sbk + (n0 + n) * p.sbn) : 0.f;
  }
  __syncthreads();
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int k = 0; k < OU_K; ++k) {
    const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
  const bool vec = (p.ldc % 4 == 0) && (reinterpret_cast<uintptr_t>(p.C) % 16 == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gn = n0 + h * 64 + tx * 4;
      float* o = p.C + gm * p.ldc + gn;
      float v[4] = {acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]};
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < p.N) v[j] += __ldg(p.bias + gn + j);
      }
      if (!p.atomic && vec && gn + 3 < p.N) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < p.N) {
            if (p.atomic) atomicAdd(o + j, v[j]);
            else o[j] = v[j];
          }
      }
    }
  }
}

// The query-encoder kernels are meant to run BESIDE the persistent tcgen05 GEMM (which configures every SM for the maximum
// shared-memory carve-out); a kernel preferring a different L1/shared split cannot become co-resident on such an SM.
template <typename F>
static void prefer_max_smem(F* fn) {
  cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int BM>
__global__ void __launch_bounds__(256) sgemm_kernel(const SgemmP p) {
  pdl_sync();
  sgemm_body<BM>(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Several small contractions in ONE launch (the query encoder / gate backward is a chain of ~20 of them, each far too small
// to fill the GPU and each paying a launch): CTA -> job by prefix sums, then the single-problem body.
constexpr int SG_MAX_JOBS = 12;
struct SgemmJobs {
  int n;
  int cta_start[SG_MAX_JOBS + 1];
  int gx[SG_MAX_JOBS], gy[SG_MAX_JOBS], bm[SG_MAX_JOBS];
  SgemmP p[SG_MAX_JOBS];
};
__global__ void __launch_bounds__(256) sgemm_multi_kernel(const SgemmJobs jobs) {
  pdl_sync();
  int j = 0;
#pragma unroll
  for (int i = 1; i < SG_MAX_JOBS; ++i)
    if (i < jobs.n && static_cast<int>(blockIdx.x) >= jobs.cta_start[i]) j = i;
  const int local = blockIdx.x - jobs.cta_start[j];
  const int bx = local % jobs.gx[j];
  const int by = (local / jobs.gx[j]) % jobs.gy[j];
  const int bz = local / (jobs.gx[j] * jobs.gy[j]);
  if (jobs.bm[j] == 0) outer_body(jobs.p[j], bx, by);
  else if (jobs.bm[j] == 32) sgemm_body<32>(jobs.p[j], bx, by, bz);
  else sgemm_body<64>(jobs.p[j], bx, by, bz);
}

// Host-side batch: add() problems, launch() once.  A problem ACCUMULATES into its output with atomics (C must hold valid data;
// small problems are K-split to fill the GPU) unless store = true: then it overwrites C with plain stores, unsplit.
struct SgemmBatch {
  SgemmJobs jobs;
  SgemmBatch() { jobs.n = 0; jobs.cta_start[0] = 0; }
  int add(const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn, float* C, long long ldc, int M,
          int N, int K, const float* bias = nullptr, bool store = false) {
    if (jobs.n >= SG_MAX_JOBS) return fail(DRN_EINVAL, "sgemm batch: more than %d problems", SG_MAX_JOBS);
    if (M < 1 || N < 1 || K < 1) return fail(DRN_EINVAL, "sgemm batch: empty problem (%d,%d,%d)", M, N, K);
    const int i = jobs.n++;
    if (K <= OU_K && M >= OU_BM && N >= OU_BN && bias == nullptr) {  // rank-K update: staged once, HBM-write bound (outer_body)
      jobs.p[i] = SgemmP{A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, nullptr, nullptr, 0, store ? 0 : 1, K};
      jobs.gx[i] = ceil_div(M, OU_BM);
      jobs.gy[i] = ceil_div(N, OU_BN);
      jobs.bm[i] = 0;
      jobs.cta_start[i + 1] = jobs.cta_start[i] + jobs.gx[i] * jobs.gy[i];
      return 0;
    }
    const int bm = (M <= 32) ? 32 : 64;
    const int tiles = ceil_div(M, bm) * ceil_div(N, 64);
    // K-split (atomic accumulation) until the GPU is full.  A CTA walks its k-range as a chain of dependent global-memory round
    // trips (16-deep tiles, one tile of prefetch), so the M <= 32 problems with a long contraction (data gradients of the gates
    // and of the command projections: 16-32 tiles, K up to 4096) are split four times finer: ~4 round trips instead of ~26.
    int splits = 1;
    if (!store && tiles < 74) splits = max(1, min(ceil_div(bm == 32 ? 592 : 148, tiles), ceil_div(K, 64)));
    const int kchunk = ceil_div(ceil_div(K, splits), 16) * 16;
    splits = ceil_div(K, kchunk);
    jobs.p[i] = SgemmP{A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, nullptr, 0, store ? 0 : 1, kchunk};
    jobs.gx[i] = ceil_div(M, bm);
    jobs.gy[i] = ceil_div(N, 64);
    jobs.bm[i] = bm;
    jobs.cta_start[i + 1] = jobs.cta_start[i] + jobs.gx[i] * jobs.gy[i] * splits;
    return 0;
  }
  int launch(cudaStream_t st) {
    if (jobs.n == 0) return 0;
    for (int i = jobs.n; i < SG_MAX_JOBS; ++i) jobs.cta_start[i + 1] = jobs.cta_start[jobs.n];
    launch_k(sgemm_multi_kernel, jobs.cta_start[jobs.n], 256, 0, st, jobs);
    jobs.n = 0;
    jobs.cta_start[0] = 0;
    return check_launch("sgemm_multi");
  }
};

// accumulate = 0: C is overwritten; 1: C += (C must hold valid data, e.g. a zeroed gradient buffer)
static int sgemm(cudaStream_t st, const float* A, long long sam, long long sak, const float* B, long long sbk, long long sbn,
                 float* C, long long ldc, int M, int N, int K, const float* bias, const float* bias2, int relu, int accumulate,
                 bool allow_split = true) {
  if (M < 1 || N < 1 || K < 1) return fail(DRN_EINVAL, "drn_sgemm: empty problem (%d,%d,%d)", M, N, K);
  if (relu && accumulate) return fail(DRN_EINVAL, "drn_sgemm: relu cannot be combined with accumulation");
  const int bm = (M <= 32) ? 32 : 64;
  const int tiles = ceil_div(M, bm) * ceil_div(N, 64);
  int splits = 1;
  if (!relu && allow_split && tiles < 148) {  // fill the 148 SMs by splitting the contraction
    splits = min(ceil_div(296, tiles), ceil_div(K, 64));
    if (splits < 1) splits = 1;
  }
  int kchunk = ceil_div(ceil_div(K, splits), 16) * 16;
  splits = ceil_div(K, kchunk);
  SgemmP p{A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, bias2, relu, (splits > 1 || accumulate) ? 1 : 0, kchunk};
  if (splits > 1 && !accumulate) {
    cudaError_t e = cudaMemset2DAsync(C, ldc * sizeof(float), 0, N * sizeof(float), M, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "drn_sgemm memset: %s", cudaGetErrorString(e));
  }
  static bool carve = false;
  if (!carve) {
    prefer_max_smem(sgemm_kernel<32>);
    prefer_max_smem(sgemm_kernel<64>);
    carve = true;
  }
  dim3 grid(ceil_div(M, bm), ceil_div(N, 64), splits);
  if (bm == 32) launch_k(sgemm_kernel<32>, grid, 256, 0, st, p);
  else launch_k(sgemm_kernel<64>, grid, 256, 0, st, p);
  return check_launch("sgemm");
}

// ------------------------------------------------------------------------------------------------------------------------
// Deterministic small-batch nn.Linear forward: out[b][n] = act(bias[n] + sum_k x[b][k] W[n][k]).  The forward of the path must
// be run-to-run reproducible (train-mode BatchNorm amplifies 1e-7 input noise ~100x into the early-layer gradients): every
// output is a FIXED-order sum -- no atomics.  This is a weight-streaming problem (B <= 32 rows of x against N x K weights,
// 30 MB per step for the command and gate linears) whose fp32-FMA form was issue-bound (32 FMAs per weight element, 117 us per
// step); it now runs on the warp-level tensor cores with the same split-BF16 products as the dense path
// (hi*hi + hi*lo + lo*hi, fp32 accumulate: ~2^-17 relative per product), 0.4 instructions per weight element:
//   * a CTA owns LIN_ROWS = 16 output rows (one m16 tile, W = the row-major A operand) x one 32-sample chunk (four n8 tiles,
//     x = the col-major B operand); its 8 warps split K into contiguous slices and the 8 partial tiles are folded through
//     shared memory in warp order;
//   * fragments are loaded straight from global memory with 16-byte loads: the contraction index inside a k16 block is
//     permuted (logical k {2t, 2t+1, 2t+8, 2t+9} <- physical {4t .. 4t+3}, the same for both operands), so a thread's four
//     A (or B) values of one row are ONE float4; fp32 -> (hi, lo) bf16x2 happens in registers.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int LIN_WARPS = 8, LIN_ROWS = 16, LIN_RED_LD = 33;
struct LinP {
  const float* x; long long ldx;
  const float* W; long long ldw;
  const float* bias;
  float* out; long long ldo;
  int Bn, N, K, relu;
};
constexpr int LIN_MAX_JOBS = 4;
struct LinJobs {
  int n;
  int cta_start[LIN_MAX_JOBS + 1];
  int gx[LIN_MAX_JOBS];
  LinP p[LIN_MAX_JOBS];
};
// (a, b) -> hi = {bf16(b) : bf16(a)} (a in the low half = the lower contraction index), lo = the bf16 of the remainders
__device__ __forceinline__ void lin_split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}
__device__ __forceinline__ void lin_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four consecutive values p[k .. k+3] of one operand row, zero beyond kend; 16-byte load when the whole k16 block is inside
__device__ __forceinline__ float4 lin_load4(const float* __restrict__ p, int k, int kend, bool fast) {
  if (fast) return __ldg(reinterpret_cast<const float4*>(p + k));
  float4 v;
  v.x = (k + 0 < kend) ? __ldg(p + k + 0) : 0.f;
  v.y = (k + 1 < kend) ? __ldg(p + k + 1) : 0.f;
  v.z = (k + 2 < kend) ? __ldg(p + k + 2) : 0.f;
  v.w = (k + 3 < kend) ? __ldg(p + k + 3) : 0.f;
  return v;
}
__global__ void __launch_bounds__(LIN_WARPS * 32) linear_small_kernel(const LinJobs jobs) {
  pdl_sync();
  int jb = 0;
#pragma unroll
  for (int i = 1; i < LIN_MAX_JOBS; ++i)
    if (i < jobs.n && static_cast<int>(blockIdx.x) >= jobs.cta_start[i]) jb = i;
  const LinP& P = jobs.p[jb];
  const int local = blockIdx.x - jobs.cta_start[jb];
  const int bxx = local % jobs.gx[jb], byy = local / jobs.gx[jb];
  const int Bn = P.Bn, N = P.N, K = P.K;
  __shared__ float red[LIN_WARPS][LIN_ROWS][LIN_RED_LD];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const int b0 = byy * 32, n0 = bxx * LIN_ROWS;
  const int kslice = (K + LIN_WARPS * 16 - 1) / (LIN_WARPS * 16) * 16;  // per-warp K slice, whole k16 blocks
  const int kbeg = w * kslice, kend = min(K, kbeg + kslice);
  const bool vec = (P.ldw % 4 == 0) && (P.ldx % 4 == 0) && (((reinterpret_cast<uintptr_t>(P.W) | reinterpret_cast<uintptr_t>(P.x)) & 15) == 0);
  // rows / samples beyond the problem are clamped for the loads and dropped at the store
  const float* __restrict__ wr0 = P.W + static_cast<long long>(min(n0 + g, N - 1)) * P.ldw;
  const float* __restrict__ wr1 = P.W + static_cast<long long>(min(n0 + g + 8, N - 1)) * P.ldw;
  const float* __restrict__ xr[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) xr[j] = P.x + static_cast<long long>(min(b0 + 8 * j + g, Bn - 1)) * P.ldx;
  float c[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
#pragma unroll 2
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
    const bool fast = vec && (k0 + 16 <= kend);
    const int k = k0 + 4 * t;
    const float4 wa = lin_load4(wr0, k, kend, fast), wb = lin_load4(wr1, k, kend, fast);
    float4 xv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xv[j] = lin_load4(xr[j], k, kend, fast);
    uint32_t ah[4], al[4];
    lin_split_pack(wa.x, wa.y, ah[0], al[0]);  // (row g,     logical k 2t, 2t+1)
    lin_split_pack(wb.x, wb.y, ah[1], al[1]);  // (row g + 8, logical k 2t, 2t+1)
    lin_split_pack(wa.z, wa.w, ah[2], al[2]);  // (row g,     logical k 2t+8, 2t+9)
    lin_split_pack(wb.z, wb.w, ah[3], al[3]);  // (row g + 8, logical k 2t+8, 2t+9)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t bh0, bl0, bh1, bl1;
      lin_split_pack(xv[j].x, xv[j].y, bh0, bl0);
      lin_split_pack(xv[j].z, xv[j].w, bh1, bl1);
      lin_mma(c[j], al, bh0, bh1);  // small terms first
      lin_mma(c[j], ah, bl0, bl1);
      lin_mma(c[j], ah, bh0, bh1);
    }
  }
  // accumulator fragment: c[j][0..1] = (row g, samples 8j + 2t, +1), c[j][2..3] = (row g + 8, same samples)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[w][g][8 * j + 2 * t] = c[j][0];
    red[w][g][8 * j + 2 * t + 1] = c[j][1];
    red[w][g + 8][8 * j + 2 * t] = c[j][2];
    red[w][g + 8][8 * j + 2 * t + 1] = c[j][3];
  }
  __syncthreads();
  const int row = tid & (LIN_ROWS - 1);
  if (n0 + row < N) {
    const float bv = P.bias ? __ldg(P.bias + n0 + row) : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int sm = (tid >> 4) + 16 * h;
      if (b0 + sm >= Bn) continue;
      float v = red[0][row][sm];
#pragma unroll
      for (int q = 1; q < LIN_WARPS; ++q) v += red[q][row][sm];
      v += bv;
      P.out[static_cast<long long>(b0 + sm) * P.ldo + n0 + row] = P.relu ? fmaxf(v, 0.f) : v;
    }
  }
}
static int linear_launch(cudaStream_t st, LinJobs& jobs) {
  for (int i = jobs.n; i < LIN_MAX_JOBS; ++i) jobs.cta_start[i + 1] = jobs.cta_start[jobs.n];
  launch_k(linear_small_kernel, jobs.cta_start[jobs.n], LIN_WARPS * 32, 0, st, jobs);
  return check_launch("linear_small");
}
static int linear_add(LinJobs& jobs, const float* x, long long ldx, const float* W, long long ldw, const float* bias, float* out,
                      long long ldo, int Bn, int N, int K, int relu) {
  if (jobs.n >= LIN_MAX_JOBS) return fail(DRN_EINVAL, "drn_linear_fwd: more than %d problems in a batch", LIN_MAX_JOBS);
  if (Bn < 1 || N < 1 || K < 1) return fail(DRN_EINVAL, "drn_linear_fwd: empty problem (%d,%d,%d)", Bn, N, K);
  const int i = jobs.n++;
  if (i == 0) jobs.cta_start[0] = 0;
  jobs.p[i] = LinP{x, ldx, W, ldw, bias, out, ldo, Bn, N, K, relu};
  jobs.gx[i] = ceil_div(N, LIN_ROWS);
  jobs.cta_start[i + 1] = jobs.cta_start[i] + jobs.gx[i] * ceil_div(Bn, 32);
  return 0;
}
static int linear_small(cudaStream_t st, const float* x, long long ldx, const float* W, long long ldw, const float* bias,
                        float* out, long long ldo, int Bn, int N, int K, int relu) {
  LinJobs jobs;
  jobs.n = 0;
  int rc = linear_add(jobs, x, ldx, W, ldw, bias, out, ldo, Bn, N, K, relu);
  if (rc) return rc;
  return linear_launch(st, jobs);
}

// ------------------------------------------------------------------------------------------------------------------------
// Query encoder device view
// ------------------------------------------------------------------------------------------------------------------------
constexpr int QE_MAX_L = 64;
constexpr int DE_SLICES = 8;  // dE = dG [W_ih ; W_ih_r] contracts over 8H = 4096 gate rows with only 4 output tiles: K-split slices
struct QeDev {
  int B, L, H, E, tok_ld, BC;  // BC = ceil(B / 32) sample chunks
  const long long* tokens;
  const long long* lengths;
  const float* w_hh[2];
  int EP;  // embedding width rounded up to the 64-element K-block of the tensor-core contraction (zero padded)
  float *xg, *G, *Cst, *Hout, *v, *hid, *c3, *alpha, *bias_sum;
  __nv_bfloat16 *E_pl, *Wih_pl, *dG_pl, *Hprev_pl;  // split-BF16 planes (hi, lo) of the operands of the big projections
  float *dH, *dc3, *dhid, *dhid_pre, *dv, *dG, *dcarry, *part, *dE, *HT, *dGT, *dr, *one;
  unsigned* cnt;
  unsigned* bar;  // [2 directions][QE_MAX_L] arrival counters of the per-direction step barrier (lstm_fwd2_kernel), zeroed by qe_embed_kernel
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---- embedding gather (language_module.py:41) / scatter-add of its gradient (row 0 = padding_idx gets none) ------------
__global__ void __launch_bounds__(256) qe_embed_kernel(QeDev q, const float* __restrict__ emb) {
  pdl_sync();
  const int r = blockIdx.x;
  const int b = r / q.L, t = r % q.L;
  const long long tok = q.tokens[static_cast<long long>(b) * q.tok_ld + t];
  const long long ps = static_cast<long long>(q.B) * q.L * q.EP;
  if (r == 0 && threadIdx.x == 0) q.one[0] = 1.f;  // the constant the column-sum contractions multiply by
  if (r == 0 && threadIdx.x < 2 * QE_MAX_L) q.bar[threadIdx.x] = 0u;  // step-barrier counters of the recurrence that follows
  for (int e = threadIdx.x; e < q.EP; e += blockDim.x) {
    __nv_bfloat16 h, l;
    split_bf16(e < q.E ? emb[tok * q.E + e] : 0.f, h, l);
    q.E_pl[static_cast<long long>(r) * q.EP + e] = h;
    q.E_pl[ps + static_cast<long long>(r) * q.EP + e] = l;
  }
}
// W_ih of both directions -> one [8H][EP] planes operand (rows = dir, gate, unit; zero-padded columns), bias_sum = b_ih + b_hh
__global__ void __launch_bounds__(256) qe_pack_wih_kernel(QeDev q, const float* __restrict__ w0, const float* __restrict__ w1,
                                                          const float* __restrict__ bi0, const float* __restrict__ bh0,
                                                          const float* __restrict__ bi1, const float* __restrict__ bh1) {
  pdl_sync();
  // one warp per packed row (8H rows): 32-bit indexing, no divisions (a flat 64-bit-indexed grid-stride loop over 148 CTAs was a
  // 25 us latency chain for 5 MB)
  const int H4 = 4 * q.H, EP = q.EP, E = q.E;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= 2 * H4) return;
  const int rr = row < H4 ? row : row - H4;
  const float* __restrict__ w = (row < H4 ? w0 : w1) + static_cast<long long>(rr) * E;
  const long long ps = 2LL * H4 * EP, o = static_cast<long long>(row) * EP;
  for (int e = lane; e < EP; e += 32) {
    __nv_bfloat16 h, l;
    split_bf16(e < E ? __ldg(w + e) : 0.f, h, l);
    q.Wih_pl[o + e] = h;
    q.Wih_pl[ps + o + e] = l;
  }
  if (lane == 0) q.bias_sum[row] = row < H4 ? bi0[rr] + bh0[rr] : bi1[rr] + bh1[rr];
}
__global__ void __launch_bounds__(256) qe_embed_bwd_kernel(QeDev q, float* __restrict__ g_emb) {
  pdl_sync();
  const int r = blockIdx.x;
  const int b = r / q.L, t = r % q.L;
  const long long tok = q.tokens[static_cast<long long>(b) * q.tok_ld + t];
  if (tok == 0 || t >= q.lengths[b]) return;
  const long long slice = static_cast<long long>(q.B) * q.L * q.E;
  for (int e = threadIdx.x; e < q.E; e += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int sl = 0; sl < DE_SLICES; ++sl) v += q.dE[sl * slice + static_cast<long long>(r) * q.E + e];  // K-split slices of dE
    atomicAdd(g_emb + tok * q.E + e, v);
  }
}


// ---- the recurrence, both directions (language_module.py:42-46; torch.nn.LSTM gate order i,f,g,o) ---------------------
// grid (H/8, 2, BC); warp = one hidden unit (its 4 gate rows of W_hh), lane = sample.  Packed-sequence semantics: a sample
// is live at time t iff t < length; the reverse direction starts at t = length-1 from the zero state, which falls out of
// the state being zero at every non-live position.  The hidden state is also kept TRANSPOSED ([unit][32 samples],
// double-buffered over steps) so that a step stages it, like its 64 KB slice of W_hh, with straight 16-byte cp.async
// copies.  PERSIST: ONE cooperative launch walks all L steps with the W_hh slice resident in shared memory and a grid
// barrier between steps (used when the grid fits the GPU: B <= 32); otherwise one launch per step.
constexpr int LSTM_THREADS = 512;  // 16 warps: two per hidden unit (each takes half of the contraction), 8 units per CTA
template <bool PERSIST>
__global__ void __launch_bounds__(LSTM_THREADS) lstm_fwd_kernel(QeDev q, int s0, int s1) {
  pdl_sync();
  extern __shared__ __align__(16) float smem[];
  const int H = q.H, L = q.L;
  float* ws = smem;                 // [32][H]   rows (unit, gate) -> unit*4+g
  float* hs = smem + 32 * H;        // [H][32]   previous hidden state of this sample chunk, unit-major
  float* comb = smem + 64 * H;      // [8][4][32] partial gate sums of the upper K half
  const int ug = blockIdx.x, dir = blockIdx.y, bc = blockIdx.z, b0 = bc * 32;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int ul = w >> 1, kh = w & 1;  // local unit, K half
  const long long ht_sz = 2LL * q.BC * H * 32;
  const int b = b0 + lane;
  const int unit = ug * 8 + ul;
  const bool inb = b < q.B;
  const int len = inb ? static_cast<int>(q.lengths[b]) : 0;
  bool w_loaded = false;
  float c_carry = 0.f;  // cell state of (sample, unit) written by the previous step of THIS launch (same thread every step)
  for (int s = s0; s < s1; ++s) {
    const int t = dir == 0 ? s : L - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    // input projection of this step (independent of the recurrence): issued before the wait on the staged state below
    const long long r = static_cast<long long>(inb ? b : 0) * L + t;
    float acc[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[g] = (inb && kh == 0) ? q.xg[((r * 2 + dir) * 4 + g) * H + unit] : 0.f;
    if (s > 0) {
      if (!w_loaded) {
        const float* W = q.w_hh[dir];
        const int c4 = H / 4;  // 16-byte chunks per W row
        for (int idx = tid; idx < 32 * c4; idx += LSTM_THREADS) {
          const int row = idx / c4, k4 = idx % c4;
          const int u = ug * 8 + (row >> 2), g = row & 3;
          cp_async16(ws + row * H + k4 * 4, W + (static_cast<long long>(g) * H + u) * H + k4 * 4);
        }
        w_loaded = true;
      }
      const float* hprev = q.HT + ((s - 1) & 1) * ht_sz + (static_cast<long long>(dir) * q.BC + bc) * H * 32;
      for (int idx = tid; idx < H * 8; idx += LSTM_THREADS) cp_async16(hs + idx * 4, hprev + idx * 4);
      cp_async_wait_all();
      __syncthreads();
    }
    if (s > 0) {
      const float* wr = ws + (ul * 4) * H;
      const int kbeg = kh * (H / 2), kend = kbeg + H / 2;
#pragma unroll 2
      for (int k = kbeg; k < kend; k += 4) {
        const float h0 = hs[(k + 0) * 32 + lane], h1 = hs[(k + 1) * 32 + lane], h2 = hs[(k + 2) * 32 + lane],
                    h3 = hs[(k + 3) * 32 + lane];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 wv = *reinterpret_cast<const float4*>(wr + g * H + k);
          acc[g] = fmaf(h0, wv.x, acc[g]);
          acc[g] = fmaf(h1, wv.y, acc[g]);
          acc[g] = fmaf(h2, wv.z, acc[g]);
          acc[g] = fmaf(h3, wv.w, acc[g]);
        }
      }
      if (kh == 1) {
#pragma unroll
        for (int g = 0; g < 4; ++g) comb[(ul * 4 + g) * 32 + lane] = acc[g];
      }
      __syncthreads();
      if (kh == 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[g] += comb[(ul * 4 + g) * 32 + lane];
      }
    }
    if (kh == 0) {
      const bool live = inb && t < len;
      float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, c = 0.f, h = 0.f;
      if (live) {
        gi = sigmoidf_(acc[0]);
        gf = sigmoidf_(acc[1]);
        gg = tanhf(acc[2]);
        go = sigmoidf_(acc[3]);
        // previous cell state: zero at non-live positions, so the register copy is exact for both directions
        const float cp = (s > s0) ? c_carry
                                  : ((tp >= 0 && tp < L) ? q.Cst[((static_cast<long long>(b) * L + tp) * 2 + dir) * H + unit] : 0.f);
        c = gf * cp + gi * gg;
        h = go * tanhf(c);
      }
      c_carry = c;
      q.HT[(s & 1) * ht_sz + ((static_cast<long long>(dir) * q.BC + bc) * H + unit) * 32 + lane] = h;
      if (inb) {
        float* G = q.G + ((r * 2 + dir) * 4) * H + unit;
        G[0] = gi; G[H] = gf; G[2 * H] = gg; G[3 * H] = go;
        q.Cst[(r * 2 + dir) * H + unit] = c;
        q.Hout[r * 2 * H + dir * H + unit] = h;
      }
    }
    if (PERSIST && s + 1 < s1) {
      __threadfence();
      cg::this_grid().sync();
    } else if (!PERSIST) {
      __syncthreads();
    }
  }
}

// ---- the recurrence, second form (B <= 32, H % 128 == 0: the DRN configuration): W_hh in REGISTERS -------------------------
// The kernel above keeps both operands of h W_hh^T in shared memory (lane = sample: one broadcast 16-byte load of W per 4
// FMAs plus four 4-byte loads of h per 16): 8 shared-memory instructions per 16 FMAs, LSU-bound at ~4.3 us of the 11.7 us a
// step took (r01 ncu: 117 us for L = 10).  Here a warp owns ONE hidden unit and a lane owns a K-SLICE of its four gate rows
// (k = 4 lane + 128 j: 64 weights in registers for H = 512, loaded once per launch); the previous hidden state is staged
// sample-major, so one conflict-free 16-byte load feeds 16 FMAs; the 32 lanes' partial sums of (gate, sample) are folded by
// recursive halving (16 shuffles per gate and 16 samples) which leaves lane l with the four gate pre-activations of sample l --
// the thread that then runs the cell update and keeps the cell state in a register for the whole sequence.  Steps are separated
// by a per-DIRECTION barrier (64 CTAs, one arrival counter per step) instead of the grid-wide cooperative barrier.
// grid (H/8, 2), 8 warps; launched cooperatively (all CTAs must be co-resident for the barrier).
constexpr int LSTM2_THREADS = 256;
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Arrival at a step barrier: fire-and-forget add with release semantics (orders this thread's earlier writes and, through the
// __syncthreads() before it, those of the whole CTA) -- no round trip for the returned value, no separate fence.
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float fold16(float (&a)[16], int lane) {
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? a[i] : a[i + s];
      const float keep = up ? a[i + s] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return a[0] + __shfl_xor_sync(0xffffffffu, a[0], 16);  // lane l: sum over all lanes of entry (l & 15)
}
__device__ __forceinline__ float fold8(float (&a)[8], int lane) {
#pragma unroll
  for (int s = 4; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? a[i] : a[i + s];
      const float keep = up ? a[i + s] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  float v = a[0] + __shfl_xor_sync(0xffffffffu, a[0], 8);
  return v + __shfl_xor_sync(0xffffffffu, v, 16);  // lane l: sum over all lanes of entry (l & 7)
}
__global__ void __launch_bounds__(LSTM2_THREADS, 1) lstm_fwd2_kernel(QeDev q) {
  pdl_sync();
  extern __shared__ __align__(16) float hs[];  // [32 samples][H]: hidden state of the previous step
  const int H = q.H, L = q.L, NJ = H >> 7;
  const int ug = blockIdx.x, dir = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int unit = ug * 8 + w;
  const int b = lane;
  const bool inb = b < q.B;
  const int len = inb ? static_cast<int>(q.lengths[b]) : 0;
  float4 wr[4][4];
#pragma unroll
  for (int g = 0; g < 4; ++g)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      wr[g][j] = (j < NJ) ? __ldg(reinterpret_cast<const float4*>(q.w_hh[dir] + (static_cast<long long>(g) * H + unit) * H + 4 * lane + 128 * j))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
  const long long ht_sz = 2LL * 32 * H;  // one buffer: [2 directions][32][H]
  float c_carry = 0.f;
  for (int s = 0; s < L; ++s) {
    const int t = dir == 0 ? s : L - 1 - s;
    const long long r = static_cast<long long>(inb ? b : 0) * L + t;
    float pre[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) pre[g] = inb ? q.xg[((r * 2 + dir) * 4 + g) * H + unit] : 0.f;  // issued before the barrier wait
    if (s > 0) {
      if (tid == 0) {  // per-direction barrier: every CTA of this direction has published its h of step s - 1 (arrival below)
        unsigned* cnt = q.bar + dir * QE_MAX_L + (s - 1);
        const long long t0 = clock64();
        while (ld_acquire_u32(cnt) < gridDim.x) {
          if (clock64() - t0 > 4000000000LL) __trap();
        }
      }
      __syncthreads();
      const float* hprev = q.HT + ((s - 1) & 1) * ht_sz + static_cast<long long>(dir) * 32 * H;
      for (int idx = tid; idx < 8 * H; idx += LSTM2_THREADS) cp_async16(hs + idx * 4, hprev + idx * 4);
      cp_async_wait_all();
      __syncthreads();
      // 8 samples a pass; the dot products run as packed FFMA2 (two k-lanes per instruction: 1024 instead of 2048 issue slots
      // per lane and step), the two halves of every accumulator pair are added before the fold
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {
        float2 acc[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[g][i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float* hrow = hs + (pass * 8 + i) * H + 4 * lane;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < NJ) {
              const float4 h4 = *reinterpret_cast<const float4*>(hrow + 128 * j);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                acc[g][i] = __ffma2_rn(make_float2(wr[g][j].x, wr[g][j].y), make_float2(h4.x, h4.y), acc[g][i]);
                acc[g][i] = __ffma2_rn(make_float2(wr[g][j].z, wr[g][j].w), make_float2(h4.z, h4.w), acc[g][i]);
              }
            }
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) a8[i] = acc[g][i].x + acc[g][i].y;
          const float v = fold8(a8, lane);
          if ((lane >> 3) == pass) pre[g] += v;  // lane l keeps sample l
        }
      }
    }
    const bool live = inb && t < len;
    float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, c = 0.f, h = 0.f;
    if (live) {
      gi = sigmoidf_(pre[0]);
      gf = sigmoidf_(pre[1]);
      gg = tanhf(pre[2]);
      go = sigmoidf_(pre[3]);
      c = gf * c_carry + gi * gg;  // non-live positions carry c = 0: exact for both directions (packed-sequence semantics)
      h = go * tanhf(c);
    }
    c_carry = c;
    q.HT[(s & 1) * ht_sz + (static_cast<long long>(dir) * 32 + b) * H + unit] = h;
    __syncthreads();  // all of this CTA's h values are written (and hs is free): arrive at the barrier of this step at once,
    if (tid == 0 && s + 1 < L) red_release_add(q.bar + dir * QE_MAX_L + s, 1u);  // the stores nobody waits for come after
    if (inb) {
      float* G = q.G + ((r * 2 + dir) * 4) * H + unit;
      G[0] = gi; G[H] = gf; G[2 * H] = gg; G[3 * H] = go;
      q.Cst[(r * 2 + dir) * H + unit] = c;
      q.Hout[r * 2 * H + dir * H + unit] = h;
      // Hprev planes [dir][b][t][unit] = the hidden state the recurrence consumes at time t (operand of the W_hh weight
      // gradient): this step's h is the input of the next one, the first step consumed zeros (was a separate kernel)
      const long long ps = 2LL * q.B * L * H;
      const long long row0 = (static_cast<long long>(dir) * q.B + b) * L;
      if (s == 0) {
        const __nv_bfloat16 z = __float2bfloat16(0.f);
        q.Hprev_pl[(row0 + t) * H + unit] = z;
        q.Hprev_pl[ps + (row0 + t) * H + unit] = z;
      }
      if (s + 1 < L) {
        const int tn = dir == 0 ? t + 1 : t - 1;
        __nv_bfloat16 hh, ll;
        split_bf16(h, hh, ll);
        q.Hprev_pl[(row0 + tn) * H + unit] = hh;
        q.Hprev_pl[ps + (row0 + tn) * H + unit] = ll;
      }
    }
  }
}

// Hprev planes [dir][b][t] = hidden state the recurrence consumed at step t (operand of the W_hh weight gradient)
__global__ void __launch_bounds__(256) qe_hprev_kernel(QeDev q) {
  pdl_sync();
  const int H = q.H, L = q.L;
  const long long total = 2LL * q.B * L * H, ps = total;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const int k = static_cast<int>(i % H);
    const long long rr = i / H;
    const int t = static_cast<int>(rr % L);
    const long long db = rr / L;
    const int b = static_cast<int>(db % q.B), dir = static_cast<int>(db / q.B);
    const int tp = dir == 0 ? t - 1 : t + 1;
    float v = 0.f;
    if (tp >= 0 && tp < L) v = q.Hout[(static_cast<long long>(b) * L + tp) * 2 * H + dir * H + k];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    q.Hprev_pl[i] = h;
    q.Hprev_pl[ps + i] = l;
  }
}

// ---- BPTT, both directions ------------------------------------------------------------------------------------------------
// grid (H/32, nq, 2*BC).  Phase A: the CTA (unit group ug, quarter jq of the 4H gate rows) forms its part of
//   dh_rec[b][u] = sum_j dG_next[b][j] * W_hh[j][u]   (lane = unit u, 32 accumulators = samples) and writes it to `part`.
// Phase B: the parts are summed and the cell backward produces dG of this time step and the carried dc.
//   per-step launches: Phase B runs in the LAST of the nq CTAs of a unit group (device counter, no spinning); nq = 1 on the
//     first step (no recurrent gradient yet);
//   PERSIST: one cooperative launch, W_hh tile resident, grid barriers between Phase A, Phase B and the next step; the
//     jq = 0 CTA of each unit group runs Phase B.
// dG is also kept transposed ([gate row][32 samples], double-buffered) so Phase A stages both operands with cp.async.
// Barrier over the CTAs of ONE direction of a cooperative LSTM launch: one arrival counter per use (zeroed before the launch),
// about a third of the cost of the grid-wide cooperative barrier, and the two directions never wait for each other.
__device__ __forceinline__ void dir_barrier(unsigned* cnt, unsigned expected) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_add(cnt, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(cnt) < expected) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();
}
constexpr int BWD_JQ = 4;
struct BwdIn {
  float gi, gf, gg, go, c, cp, dh;
  bool live;
};
template <bool PERSIST>
__global__ void __launch_bounds__(LSTM_THREADS) lstm_bwd_kernel(QeDev q, int s0, int s1, int nq_launch) {
  pdl_sync();
  extern __shared__ __align__(16) float smem[];
  const int H = q.H, L = q.L;
  const int ug = blockIdx.x, jq = blockIdx.y, dir = blockIdx.z / q.BC, bc = blockIdx.z % q.BC, b0 = bc * 32;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int UG = H / 32;
  const long long slot = (static_cast<long long>(dir) * q.BC + bc) * UG + ug;
  const long long gt_sz = 2LL * q.BC * 4 * H * 32;
  float* part = q.part + slot * BWD_JQ * 1024;
  float* dgs = smem;               // [H][32]  dG_next rows of this quarter, (gate row j, sample b)
  float* wsm = smem + H * 32;      // [H][32]  W_hh[jq*H + j][ug*32 + u]
  float* red = smem + 2 * H * 32;  // [16][32][33]; reused as the [4][32][33] transposition tile of Phase B
  constexpr int NW = LSTM_THREADS / 32;
  __shared__ bool is_last;
  bool w_loaded = false;
  float carry_reg[2] = {0.f, 0.f};
  for (int s = s0; s < s1; ++s) {
    const int t = dir == 0 ? L - 1 - s : s;
    const int tp = dir == 0 ? t - 1 : t + 1;  // recurrence predecessor (source of c_prev)
    const int nq = PERSIST ? (s == 0 ? 1 : BWD_JQ) : nq_launch;
    bool do_b = PERSIST ? (jq == 0) : true;
    // Phase B operands that do not depend on Phase A (gates, cell states, upstream dH of this time step): the persistent
    // kernel's Phase-B CTAs request them BEFORE Phase A and its grid barrier, which hides one global-memory round trip per step
    auto load_in = [&](int idx) -> BwdIn {
      BwdIn in;
      in.live = false;
      in.gi = in.gf = in.gg = in.go = in.c = in.cp = in.dh = 0.f;
      const int b = b0 + (idx >> 5), unit = ug * 32 + (idx & 31);
      if (b < q.B && t < q.lengths[b]) {
        const long long r = static_cast<long long>(b) * L + t;
        const float* G = q.G + ((r * 2 + dir) * 4) * H + unit;
        in.live = true;
        in.gi = G[0]; in.gf = G[H]; in.gg = G[2 * H]; in.go = G[3 * H];
        in.c = q.Cst[(r * 2 + dir) * H + unit];
        in.cp = (tp >= 0 && tp < L) ? q.Cst[((static_cast<long long>(b) * L + tp) * 2 + dir) * H + unit] : 0.f;
        in.dh = q.dH[r * 2 * H + dir * H + unit];
      }
      return in;
    };
    BwdIn pre[2];
    const bool prefetched = PERSIST && do_b;
    if (prefetched) {
#pragma unroll
      for (int e = 0; e < 2; ++e) pre[e] = load_in(tid + e * LSTM_THREADS);
    }
    if (nq > 1) {
      const float* src = q.dGT + ((s - 1) & 1) * gt_sz + ((static_cast<long long>(dir) * q.BC + bc) * 4 * H + jq * H) * 32;
      for (int idx = tid; idx < H * 8; idx += LSTM_THREADS) cp_async16(dgs + idx * 4, src + idx * 4);
      if (!w_loaded) {
        const float* W = q.w_hh[dir] + static_cast<long long>(jq) * H * H + ug * 32;
        for (int idx = tid; idx < H * 8; idx += LSTM_THREADS) {
          const int j = idx >> 3, c = idx & 7;
          cp_async16(wsm + j * 32 + c * 4, W + static_cast<long long>(j) * H + c * 4);
        }
        w_loaded = PERSIST;
      }
      cp_async_wait_all();
      __syncthreads();
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      const int jw = H / NW;  // rows per warp
#pragma unroll 2
      for (int j = w * jw; j < (w + 1) * jw; ++j) {
        const float wv = wsm[j * 32 + lane];
        const float2 wv2 = make_float2(wv, wv);
        const float4* d4 = reinterpret_cast<const float4*>(dgs + j * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // packed FFMA2: two samples per instruction (same products, same order per sample)
          const float4 d = d4[i];
          const float2 r0 = __ffma2_rn(make_float2(d.x, d.y), wv2, make_float2(acc[4 * i + 0], acc[4 * i + 1]));
          const float2 r1 = __ffma2_rn(make_float2(d.z, d.w), wv2, make_float2(acc[4 * i + 2], acc[4 * i + 3]));
          acc[4 * i + 0] = r0.x; acc[4 * i + 1] = r0.y; acc[4 * i + 2] = r1.x; acc[4 * i + 3] = r1.y;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) red[(w * 32 + i) * 33 + lane] = acc[i];
      __syncthreads();
      for (int idx = tid; idx < 1024; idx += LSTM_THREADS) {
        const int bb = idx >> 5, u = idx & 31;
        float sum = 0.f;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) sum += red[(ww * 32 + bb) * 33 + u];
        part[jq * 1024 + idx] = sum;
      }
      if (!PERSIST) __threadfence();  // PERSIST: the barrier's release (after its __syncthreads) publishes the CTA's writes
      if (PERSIST) {
        dir_barrier(q.bar + 2 * QE_MAX_L + dir * 2 * QE_MAX_L + 2 * s, gridDim.x * gridDim.y * q.BC);
      } else {
        __syncthreads();
        if (tid == 0) is_last = (atomicAdd(q.cnt + slot, 1u) == static_cast<unsigned>(nq - 1));
        __syncthreads();
        do_b = is_last;
        __threadfence();
      }
    }
    if (do_b) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {  // 1024 (sample, unit) pairs, two per thread -- the same two in every step
        const int idx = tid + e * LSTM_THREADS;
        const int bl = idx >> 5, u = idx & 31;
        const int b = b0 + bl, unit = ug * 32 + u;
        float d_i = 0.f, d_f = 0.f, d_g = 0.f, d_o = 0.f, carry = 0.f;
        if (b < q.B) {
          float dhrec = 0.f;
          if (nq > 1) {
#pragma unroll
            for (int qq = 0; qq < BWD_JQ; ++qq) dhrec += __ldcg(part + qq * 1024 + idx);
          }
          const long long r = static_cast<long long>(b) * L + t;
          const BwdIn in = prefetched ? pre[e] : load_in(idx);
          if (in.live) {
            const float gi = in.gi, gf = in.gf, gg = in.gg, go = in.go;
            const float dh = in.dh + dhrec;
            const float tc = tanhf(in.c);
            float dc = dh * go * (1.f - tc * tc);
            // cell-state gradient carried from the previous step: a register in the persistent kernel
            if (s > 0) dc += PERSIST ? carry_reg[e] : q.dcarry[(static_cast<long long>(dir) * q.B + b) * H + unit];
            d_i = dc * gg * gi * (1.f - gi);
            d_f = dc * in.cp * gf * (1.f - gf);
            d_g = dc * gi * (1.f - gg * gg);
            d_o = dh * tc * go * (1.f - go);
            carry = dc * gf;
          }
          const long long go_ = ((r * 2 + dir) * 4) * H + unit;
          float* dG = q.dG + go_;
          dG[0] = d_i; dG[H] = d_f; dG[2 * H] = d_g; dG[3 * H] = d_o;
          const long long gps = static_cast<long long>(q.B) * L * 8 * H;
          const float dv4[4] = {d_i, d_f, d_g, d_o};
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            __nv_bfloat16 hh, ll;
            split_bf16(dv4[g4], hh, ll);
            q.dG_pl[go_ + g4 * H] = hh;
            q.dG_pl[gps + go_ + g4 * H] = ll;
          }
          if (!PERSIST) q.dcarry[(static_cast<long long>(dir) * q.B + b) * H + unit] = carry;
        }
        carry_reg[e] = carry;
        red[(0 * 32 + u) * 33 + bl] = d_i;
        red[(1 * 32 + u) * 33 + bl] = d_f;
        red[(2 * 32 + u) * 33 + bl] = d_g;
        red[(3 * 32 + u) * 33 + bl] = d_o;
      }
      __syncthreads();
      float* dgt = q.dGT + (s & 1) * gt_sz + (static_cast<long long>(dir) * q.BC + bc) * 4 * H * 32;
      for (int idx = tid; idx < 4096; idx += LSTM_THREADS) {  // (gate, unit) rows x 32 samples, coalesced over samples
        const int row = idx >> 5, bl = idx & 31;
        const int g = row >> 5, u = row & 31;
        dgt[(static_cast<long long>(g) * H + ug * 32 + u) * 32 + bl] = red[row * 33 + bl];
      }
      if (!PERSIST && tid == 0 && nq > 1) q.cnt[slot] = 0u;
    }
    if (PERSIST && s + 1 < s1) {
      dir_barrier(q.bar + 2 * QE_MAX_L + dir * 2 * QE_MAX_L + 2 * s + 1, gridDim.x * gridDim.y * q.BC);
    }
  }
}

// ---- q_vector = [H[b,0] | H[b,len-1]] (language_module.py:50-55) and the scatter of its gradient ------------------------
__global__ void __launch_bounds__(256) qe_vgather_kernel(QeDev q) {
  pdl_sync();
  const int b = blockIdx.x, D = 2 * q.H;
  const long long last = q.lengths[b] - 1;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    q.v[static_cast<long long>(b) * 2 * D + d] = q.Hout[(static_cast<long long>(b) * q.L) * D + d];
    q.v[static_cast<long long>(b) * 2 * D + D + d] = q.Hout[(static_cast<long long>(b) * q.L + last) * D + d];
  }
}
__global__ void __launch_bounds__(256) qe_vscatter_kernel(QeDev q) {
  pdl_sync();
  const int b = blockIdx.x, D = 2 * q.H;
  const long long last = q.lengths[b] - 1;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    q.dH[(static_cast<long long>(b) * q.L) * D + d] += q.dv[static_cast<long long>(b) * 2 * D + d];
    q.dH[(static_cast<long long>(b) * q.L + last) * D + d] += q.dv[static_cast<long long>(b) * 2 * D + D + d];
  }
}
// ReLU backward from the activation itself (relu(x) > 0 <=> x > 0)
__global__ void relu_mask_kernel(const float* __restrict__ g, const float* __restrict__ act, float* __restrict__ y, int n) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = act[i] > 0.f ? g[i] : 0.f;
}

// ---- attention over the words (language_module.py:27-36): grid (B, 3) --------------------------------------------------
__global__ void __launch_bounds__(256) qe_attn_fwd_kernel(QeDev q, const float* __restrict__ wa, const float* __restrict__ ba,
                                                          float* cmd0, float* cmd1, float* cmd2) {
  pdl_sync();
  extern __shared__ __align__(16) float smem[];
  const int D = 2 * q.H, L = q.L;
  float* wc = smem;       // [D]
  float* raw = smem + D;  // [L]
  const int b = blockIdx.x, t = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* c = q.c3 + (static_cast<long long>(t) * q.B + b) * D;
  const float* Hb = q.Hout + static_cast<long long>(b) * L * D;
  const int len = static_cast<int>(q.lengths[b]);
  for (int d = tid; d < D; d += 256) wc[d] = c[d] * wa[d];
  __syncthreads();
  for (int j = w; j < L; j += 8) {
    float sdot = 0.f;
    const float4* h4 = reinterpret_cast<const float4*>(Hb + static_cast<long long>(j) * D);
#pragma unroll 8
    for (int d = lane; d < D / 4; d += 32) {
      const float4 h = __ldg(h4 + d);
      sdot += wc[4 * d] * h.x + wc[4 * d + 1] * h.y + wc[4 * d + 2] * h.z + wc[4 * d + 3] * h.w;
    }
    sdot = warp_sum(sdot);
    if (lane == 0) raw[j] = (j < len) ? sdot + ba[0] : -1e30f;  // ops.py:74-85
  }
  __syncthreads();
  if (w == 0) {
    float m = -INFINITY;
    for (int j = lane; j < L; j += 32) m = fmaxf(m, raw[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float e = expf(raw[j] - m);
      raw[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int j = lane; j < L; j += 32) {
      const float a = raw[j] / sum;
      raw[j] = a;
      q.alpha[(static_cast<long long>(t) * q.B + b) * L + j] = a;
    }
  }
  __syncthreads();
  float* cmd = (t == 0 ? cmd0 : (t == 1 ? cmd1 : cmd2)) + static_cast<long long>(b) * D;
  for (int d = tid; d < D; d += 256) {
    float acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < L; ++j) acc = fmaf(raw[j], Hb[static_cast<long long>(j) * D + d], acc);
    cmd[d] = acc;
  }
}

// Backward of the attention, two launches.  (1) grid (B, 3): softmax backward scalars dr[t][b][j] = alpha_j (dalpha_j - sum_i
// alpha_i dalpha_i), dalpha_j = dcmd . H[b,j]; zero at masked positions (alpha = 0).
__global__ void __launch_bounds__(256) qe_attn_bwd_scalars_kernel(QeDev q, const float* dcmd0, const float* dcmd1,
                                                                  const float* dcmd2, float* __restrict__ g_ba) {
  pdl_sync();
  __shared__ float da[QE_MAX_L];
  const int D = 2 * q.H, L = q.L;
  const int b = blockIdx.x, t = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* Hb = q.Hout + static_cast<long long>(b) * L * D;
  const float* dcmd = (t == 0 ? dcmd0 : (t == 1 ? dcmd1 : dcmd2)) + static_cast<long long>(b) * D;
  const float* alpha = q.alpha + (static_cast<long long>(t) * q.B + b) * L;
  for (int j = w; j < L; j += 8) {  // D = 2H is a multiple of 64: 16-byte loads, all of a row's loads in flight (the scalar
    float sdot = 0.f;               // loop was a chain of ~64 dependent memory round trips: 33 us for 4 MB, ncu r01 v12)
    const float4* h4 = reinterpret_cast<const float4*>(Hb + static_cast<long long>(j) * D);
    const float4* c4 = reinterpret_cast<const float4*>(dcmd);
#pragma unroll 8
    for (int d = lane; d < D / 4; d += 32) {
      const float4 a = __ldg(c4 + d), h = __ldg(h4 + d);
      sdot += a.x * h.x + a.y * h.y + a.z * h.z + a.w * h.w;
    }
    sdot = warp_sum(sdot);
    if (lane == 0) da[j] = sdot;
  }
  __syncthreads();
  if (w == 0) {
    float sdot = 0.f;
    for (int j = lane; j < L; j += 32) sdot = fmaf(alpha[j], da[j], sdot);
    sdot = warp_sum(sdot);
    float dba = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float dr = alpha[j] * (da[j] - sdot);
      q.dr[(static_cast<long long>(t) * q.B + b) * L + j] = dr;
      dba += dr;
    }
    dba = warp_sum(dba);
    if (lane == 0 && g_ba) atomicAdd(g_ba, dba);
  }
}
// (2) grid (B, 2H/256), one thread per channel d: dH[b,j,d] = sum_t alpha_tj dcmd_t[d] + dr_tj wa[d] c_t[d];
// dc_t[b,d] = wa[d] sum_j dr_tj H[b,j,d]; d cmd_inter2logits.weight[d] += sum_t c_t[d] sum_j dr_tj H[b,j,d].
__global__ void __launch_bounds__(256) qe_attn_bwd_apply_kernel(QeDev q, const float* __restrict__ wa, const float* dcmd0,
                                                                const float* dcmd1, const float* dcmd2,
                                                                float* __restrict__ g_wa) {
  pdl_sync();
  __shared__ float al[3][QE_MAX_L], dr[3][QE_MAX_L];
  const int D = 2 * q.H, L = q.L;
  const int b = blockIdx.x, d = blockIdx.y * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < 3 * L; i += 256) {
    const int t = i / L, j = i % L;
    al[t][j] = q.alpha[(static_cast<long long>(t) * q.B + b) * L + j];
    dr[t][j] = q.dr[(static_cast<long long>(t) * q.B + b) * L + j];
  }
  __syncthreads();
  if (d >= D) return;
  const float wad = wa[d];
  float cd[3], dcd[3], hsum[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    cd[t] = q.c3[(static_cast<long long>(t) * q.B + b) * D + d];
    dcd[t] = (t == 0 ? dcmd0 : (t == 1 ? dcmd1 : dcmd2))[static_cast<long long>(b) * D + d];
  }
  const float* Hb = q.Hout + static_cast<long long>(b) * L * D + d;
  float* dHb = q.dH + static_cast<long long>(b) * L * D + d;
  for (int j = 0; j < L; ++j) {
    const float hv = Hb[static_cast<long long>(j) * D];
    float g = 0.f;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      hsum[t] = fmaf(dr[t][j], hv, hsum[t]);
      g += al[t][j] * dcd[t] + dr[t][j] * wad * cd[t];
    }
    dHb[static_cast<long long>(j) * D] = g;
  }
  float gw = 0.f;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    q.dc3[(static_cast<long long>(t) * q.B + b) * D + d] = wad * hsum[t];
    gw = fmaf(cd[t], hsum[t], gw);
  }
  if (g_wa) atomicAdd(g_wa + d, gw);
}

// ---- workspace carving ------------------------------------------------------------------------------------------------
static size_t carve(const drn_qe_t* a, QeDev* q) {
  const size_t B = a->B, L = a->L, H = a->H, E = a->E, R = B * L;
  const size_t BC = (B + 31) / 32;
  size_t off = 0;
  char* base = static_cast<char*>(a->workspace);
  auto take = [&](size_t nfloat) -> float* {
    float* p = q ? reinterpret_cast<float*>(base + off) : nullptr;
    off += ((nfloat * sizeof(float) + 255) / 256) * 256;
    return p;
  };
  const size_t EP = (E + 63) / 64 * 64;
  float* E_pl = take(R * EP);          // 2 planes of bf16 = R*EP floats
  float* Wih_pl = take(8 * H * EP);    // [8H][EP] x 2 planes of bf16
  float* bias_sum = take(8 * H);
  float* dG_pl = take(R * 8 * H);
  float* Hprev_pl = take(2 * R * H);
  float* xg = take(R * 8 * H);
  float* G = take(R * 8 * H);
  float* Cst = take(R * 2 * H);
  float* Hout = take(R * 2 * H);
  float* v = take(B * 4 * H);
  float* hid = take(B * H);
  float* c3 = take(3 * B * 2 * H);
  float* alpha = take(3 * B * L);
  float* dH = take(R * 2 * H);
  float* dc3 = take(3 * B * 2 * H);
  float* dhid = take(B * H);
  float* dhid_pre = take(B * H);
  float* dv = take(B * 4 * H);
  float* dG = take(R * 8 * H);
  float* dcarry = take(2 * B * H);
  float* part = take(2 * BC * (H / 32) * BWD_JQ * 1024);
  float* dE = take(DE_SLICES * R * E);
  float* cnt = take(2 * BC * (H / 32));
  float* HT = take(2 * 2 * BC * H * 32);
  float* dGT = take(2 * 2 * BC * 4 * H * 32);
  float* dr = take(3 * B * L);
  float* one = take(1);
  float* bar = take(6 * QE_MAX_L);  // forward: [2][L]; backward: [2][2 L] (two barriers per step)
  if (q) {
    q->EP = static_cast<int>(EP);
    q->E_pl = reinterpret_cast<__nv_bfloat16*>(E_pl); q->Wih_pl = reinterpret_cast<__nv_bfloat16*>(Wih_pl);
    q->dG_pl = reinterpret_cast<__nv_bfloat16*>(dG_pl); q->Hprev_pl = reinterpret_cast<__nv_bfloat16*>(Hprev_pl);
    q->bias_sum = bias_sum;
    q->xg = xg; q->G = G; q->Cst = Cst; q->Hout = Hout; q->v = v;
    q->hid = hid; q->c3 = c3; q->alpha = alpha; q->dH = dH; q->dc3 = dc3; q->dhid = dhid; q->dhid_pre = dhid_pre; q->dv = dv;
    q->dG = dG; q->dcarry = dcarry; q->part = part; q->dE = dE; q->cnt = reinterpret_cast<unsigned*>(cnt);
    q->HT = HT; q->dGT = dGT; q->dr = dr; q->one = one; q->bar = reinterpret_cast<unsigned*>(bar);
  }
  return off;
}

static int make_dev(const drn_qe_t* a, QeDev* q, const char* who) {
  if (!a) return fail(DRN_EINVAL, "%s: null descriptor", who);
  if (a->B < 1 || a->L < 1 || a->L > QE_MAX_L) return fail(DRN_EINVAL, "%s: need B >= 1 and 1 <= L <= %d (B=%d, L=%d)", who, QE_MAX_L, a->B, a->L);
  if (a->H < 32 || a->H % 32 || a->H > 512) return fail(DRN_EINVAL, "%s: hidden size must be a multiple of 32, <= 512 (H=%d)", who, a->H);
  if (a->E < 1 || a->tok_ld < a->L) return fail(DRN_EINVAL, "%s: bad embedding width / token stride", who);
  const size_t need = carve(a, nullptr);
  if (!a->workspace || a->workspace_bytes < need)
    return fail(DRN_EINVAL, "%s: workspace too small (%zu < %zu bytes)", who, a->workspace_bytes, need);
  q->B = a->B; q->L = a->L; q->H = a->H; q->E = a->E; q->tok_ld = a->tok_ld; q->BC = (a->B + 31) / 32;
  q->tokens = reinterpret_cast<const long long*>(a->tokens);
  q->lengths = reinterpret_cast<const long long*>(a->lengths);
  q->w_hh[0] = a->w_hh[0];
  q->w_hh[1] = a->w_hh[1];
  carve(a, q);
  return 0;
}

// A cooperative (grid-barrier) launch needs every CTA resident at once.  DRN_QE_PERSIST=0 forces the per-step launches.
static bool fits_cooperative(const void* fn, dim3 grid, size_t smem, int threads = 512) {
  static int allow = -1;
  if (allow < 0) {
    const char* e = getenv("DRN_QE_PERSIST");
    allow = (e && e[0] == '0') ? 0 : 1;
  }
  if (!allow) return false;
  int dev = 0, coop = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem) != cudaSuccess) return false;
  return static_cast<long long>(grid.x) * grid.y * grid.z <= static_cast<long long>(sms) * per_sm;
}

// drn_gemm_t helpers: a [rows][C] planes matrix as the (c, parity, t, b, plane) view of the tensor-core contraction
static drn_planes_t planes_view(const __nv_bfloat16* p, long long plane_stride, int rows, int C) {
  drn_planes_t v;
  v.ptr = const_cast<__nv_bfloat16*>(p);
  v.plane_stride = plane_stride;
  v.B = 1; v.T = rows; v.P = 1; v.C = C;
  return v;
}
static drn_gemm_t qe_gemm_base(int form, int rows) {
  drn_gemm_t g{};
  g.form = form;
  g.B = 1; g.T = rows;
  g.ntaps = 1;
  g.nprod = 3;
  g.split_k = 1;
  g.out_mode = DRN_OUT_STORE;
  g.out_T = rows; g.out_t_mul = 1;
  g.engine = 2;
  return g;
}

static int set_smem(const void* fn, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return fail(static_cast<int>(e), "cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace drn

using namespace drn;
#define ST(s) static_cast<cudaStream_t>(s)
#define TRY(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

extern "C" int drn_sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
                         int M, int N, int K, const float* bias, int relu, int accumulate, void* stream) {
  return sgemm(ST(stream), A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, nullptr, relu, accumulate);
}

extern "C" int drn_linear_fwd(const float* x, int64_t ldx, const float* W, int64_t ldw, const float* bias, float* out, int64_t ldo,
                              int B, int N, int K, int relu, void* stream) {
  return linear_small(ST(stream), x, ldx, W, ldw, bias, out, ldo, B, N, K, relu);
}

extern "C" int drn_sgemm_batch(int n, const drn_sgemm_job_t* jobs, void* stream) {
  if (n < 1 || n > SG_MAX_JOBS || !jobs) return fail(DRN_EINVAL, "drn_sgemm_batch: 1..%d jobs (got %d)", SG_MAX_JOBS, n);
  SgemmBatch sb;
  for (int i = 0; i < n; ++i) {
    const drn_sgemm_job_t& j = jobs[i];
    TRY(sb.add(j.A, j.sam, j.sak, j.B, j.sbk, j.sbn, j.C, j.ldc, j.M, j.N, j.K, j.bias, j.store != 0));
  }
  return sb.launch(ST(stream));
}

extern "C" int drn_linear_fwd_batch(int n, const drn_linear_job_t* jobs, void* stream) {
  if (n < 1 || n > LIN_MAX_JOBS || !jobs) return fail(DRN_EINVAL, "drn_linear_fwd_batch: 1..%d jobs (got %d)", LIN_MAX_JOBS, n);
  LinJobs lj;
  lj.n = 0;
  for (int i = 0; i < n; ++i) {
    const drn_linear_job_t& j = jobs[i];
    TRY(linear_add(lj, j.x, j.ldx, j.W, j.ldw, j.bias, j.out, j.ldo, j.B, j.N, j.K, j.relu));
  }
  return linear_launch(ST(stream), lj);
}

// Kernels one drn_qe_forward / drn_qe_backward call launches (gpu_launches accounting of bench.py): the recurrence is ONE
// cooperative launch when its grid fits the GPU, else one launch per time step.
// The register-resident recurrence (lstm_fwd2_kernel, which also writes the Hprev planes) runs when the whole batch is one
// 32-sample chunk, H is a multiple of 128 and its cooperative grid fits the GPU.  DRN_QE_FWD2=0: shared-memory form (A/B).
static bool use_fwd2(int BC, int H, int L) {
  static int fwd2 = -1;
  if (fwd2 < 0) {
    const char* e = getenv("DRN_QE_FWD2");
    fwd2 = e ? atoi(e) : 1;
  }
  const size_t smem_f2 = static_cast<size_t>(32) * H * sizeof(float);
  return fwd2 && BC == 1 && H % 128 == 0 && L <= QE_MAX_L &&
         set_smem(reinterpret_cast<const void*>(lstm_fwd2_kernel), smem_f2, "lstm_fwd2") == 0 &&
         fits_cooperative(reinterpret_cast<const void*>(lstm_fwd2_kernel), dim3(H / 8, 2, 1), smem_f2, LSTM2_THREADS);
}

extern "C" int drn_qe_launch_count(int B, int L, int H, int backward) {
  const int BC = (B + 31) / 32;
  if (!backward) {
    if (use_fwd2(BC, H, L)) return 8;  // embed, pack_wih, xg contraction, recurrence (+ Hprev), vgather, qInput, qInput0-2, attention
    const size_t smem_f = (32 * H + H * 32 + 8 * 4 * 32) * sizeof(float);
    set_smem(reinterpret_cast<const void*>(lstm_fwd_kernel<true>), smem_f, "lstm_fwd");
    const bool coop = fits_cooperative(reinterpret_cast<const void*>(lstm_fwd_kernel<true>), dim3(H / 8, 2, BC), smem_f);
    return 8 + (coop ? 1 : L);  // embed, pack_wih, xg contraction, hprev, vgather, qInput, qInput0-2, attention + recurrence
  }
  const size_t smem_b = (2 * H * 32 + (LSTM_THREADS / 32) * 32 * 33) * sizeof(float);
  set_smem(reinterpret_cast<const void*>(lstm_bwd_kernel<true>), smem_b, "lstm_bwd");
  const bool coop = fits_cooperative(reinterpret_cast<const void*>(lstm_bwd_kernel<true>), dim3(H / 32, BWD_JQ, 2 * BC), smem_b);
  return 9 + (coop ? 1 : L);  // 2 attention, 2 small-contraction batches, relu mask, vscatter, grouped contraction, bias sums, embedding + BPTT
}

extern "C" size_t drn_qe_workspace_bytes(int B, int L, int H, int E) {
  drn_qe_t a{};
  a.B = B; a.L = L; a.H = H; a.E = E;
  return carve(&a, nullptr);
}

// parts: bit 0 = everything before the recurrence (embedding, W_ih packing, input projection), bit 1 = the recurrence and
// everything after it.  The split lets the caller fork an independent HBM-bound branch (weight packing) at the point where
// the latency-bound recurrence starts (drn_b200/dense.py: forward_pre).
static int qe_forward_parts(const drn_qe_t* a, int parts, void* stream) {
  QeDev q;
  TRY(make_dev(a, &q, "drn_qe_forward"));
  cudaStream_t st = ST(stream);
  const int B = q.B, L = q.L, H = q.H, R = B * L, D = 2 * H;
  if (parts & 1) {
  cudaError_t e = cudaMemsetAsync(q.cnt, 0, sizeof(unsigned) * 2 * q.BC * (H / 32), st);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "drn_qe_forward memset: %s", cudaGetErrorString(e));
  launch_k(qe_embed_kernel, R, 256, 0, st, q, a->emb);
  TRY(check_launch("qe_embed"));
  // xg = E [W_ih ; W_ih_reverse]^T + b_ih + b_hh for all time steps and both directions: ONE tensor-core contraction
  // (split-BF16, deterministic) on the persistent CTA-pair kernel
  launch_k(qe_pack_wih_kernel, ceil_div(8 * H, 8), 256, 0, st, q, a->w_ih[0], a->w_ih[1], a->b_ih[0], a->b_hh[0], a->b_ih[1], a->b_hh[1]);
  TRY(check_launch("qe_pack_wih"));
  {
    drn_gemm_t g = qe_gemm_base(DRN_GEMM_ROWS, R);
    g.a = planes_view(q.E_pl, static_cast<long long>(R) * q.EP, R, q.EP);
    g.b = planes_view(q.Wih_pl, 8LL * H * q.EP, 8 * H, q.EP);
    g.N = 8 * H; g.K = q.EP;
    g.out = q.xg; g.out_ld = 8 * H; g.bias = q.bias_sum;
    TRY(drn_gemm_group(1, &g, stream));
  }
  }
  if (!(parts & 2)) return 0;
  const size_t smem_f = (32 * H + H * 32 + 8 * 4 * 32) * sizeof(float);
  TRY(set_smem(reinterpret_cast<const void*>(lstm_fwd_kernel<true>), smem_f, "lstm_fwd"));
  TRY(set_smem(reinterpret_cast<const void*>(lstm_fwd_kernel<false>), smem_f, "lstm_fwd"));
  const dim3 grid_f(H / 8, 2, q.BC);
  const size_t smem_f2 = static_cast<size_t>(32) * H * sizeof(float);
  const bool fwd2 = use_fwd2(q.BC, H, L);
  if (fwd2) {
    void* args[] = {&q};
    cudaError_t ce = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lstm_fwd2_kernel), dim3(H / 8, 2, 1), dim3(LSTM2_THREADS), args, smem_f2, st);
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "lstm_fwd2 (cooperative): %s", cudaGetErrorString(ce));
  } else if (fits_cooperative(reinterpret_cast<const void*>(lstm_fwd_kernel<true>), grid_f, smem_f)) {
    int s0 = 0, s1 = L;
    void* args[] = {&q, &s0, &s1};
    cudaError_t ce = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lstm_fwd_kernel<true>), grid_f, dim3(LSTM_THREADS), args, smem_f, st);
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "lstm_fwd (cooperative): %s", cudaGetErrorString(ce));
  } else {
    for (int s = 0; s < L; ++s) {
      launch_k(lstm_fwd_kernel<false>, grid_f, LSTM_THREADS, smem_f, st, q, s, s + 1);
      TRY(check_launch("lstm_fwd_step"));
    }
  }
  if (!fwd2) {
    launch_k(qe_hprev_kernel, 148, 256, 0, st, q);
    TRY(check_launch("qe_hprev"));
  }
  launch_k(qe_vgather_kernel, B, 256, 0, st, q);
  TRY(check_launch("qe_vgather"));
  TRY(linear_small(st, q.v, 4 * H, a->w1, 4 * H, a->b1, q.hid, H, B, H, 4 * H, 1));
  {
    LinJobs lj;
    lj.n = 0;
    for (int t = 0; t < 3; ++t)
      TRY(linear_add(lj, q.hid, H, a->w2[t], H, a->b2[t], q.c3 + static_cast<long long>(t) * B * D, D, B, D, H, 0));
    TRY(linear_launch(st, lj));
  }
  launch_k(qe_attn_fwd_kernel, dim3(B, 3), 256, (D + L) * sizeof(float), st, q, a->wa, a->ba, a->cmd[0], a->cmd[1], a->cmd[2]);
  return check_launch("qe_attn_fwd");
}
extern "C" int drn_qe_forward(const drn_qe_t* a, void* stream) { return qe_forward_parts(a, 3, stream); }
extern "C" int drn_qe_forward_part(const drn_qe_t* a, int part, void* stream) {
  if (part != 1 && part != 2) return fail(DRN_EINVAL, "drn_qe_forward_part: part must be 1 or 2 (got %d)", part);
  return qe_forward_parts(a, part, stream);
}

// parts: bit 0 = attention / command / q_vector backward (before the BPTT), bit 1 = the BPTT and everything after it
static int qe_backward_parts(const drn_qe_t* a, int parts, void* stream) {
  QeDev q;
  TRY(make_dev(a, &q, "drn_qe_backward"));
  cudaStream_t st = ST(stream);
  const int B = q.B, L = q.L, H = q.H, E = q.E, R = B * L, D = 2 * H;
  if (!a->dcmd[0] || !a->dcmd[1] || !a->dcmd[2]) return fail(DRN_EINVAL, "drn_qe_backward: dcmd missing");
  if (parts & 1) {
  launch_k(qe_attn_bwd_scalars_kernel, dim3(B, 3), 256, 0, st, q, a->dcmd[0], a->dcmd[1], a->dcmd[2], a->g_ba);
  TRY(check_launch("qe_attn_bwd_scalars"));
  launch_k(qe_attn_bwd_apply_kernel, dim3(B, ceil_div(D, 256)), 256, 0, st, q, a->wa, a->dcmd[0], a->dcmd[1], a->dcmd[2], a->g_wa);
  TRY(check_launch("qe_attn_bwd_apply"));
  // dhid / dv accumulate from several contractions: one zero-fill of the contiguous [dhid | dhid_pre | dv] scratch
  {
    const size_t bytes = reinterpret_cast<char*>(q.dv + static_cast<size_t>(B) * 4 * H) - reinterpret_cast<char*>(q.dhid);
    cudaError_t e = cudaMemsetAsync(q.dhid, 0, bytes, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "drn_qe_backward memset: %s", cudaGetErrorString(e));
  }
  SgemmBatch sb;
  for (int t = 0; t < 3; ++t) {  // qInput0..2: d hid, d weight, d bias -- nine small contractions, one launch
    const float* dc = q.dc3 + static_cast<long long>(t) * B * D;
    TRY(sb.add(dc, D, 1, a->w2[t], H, 1, q.dhid, H, B, H, D));
    if (a->g_w2[t]) TRY(sb.add(dc, 1, D, q.hid, H, 1, a->g_w2[t], H, D, H, B));
    if (a->g_b2[t]) TRY(sb.add(dc, 1, D, q.one, 0, 0, a->g_b2[t], 1, D, 1, B));
  }
  TRY(sb.launch(st));
  launch_k(relu_mask_kernel, ceil_div(B * H, 256), 256, 0, st, q.dhid, q.hid, q.dhid_pre, B * H);
  TRY(check_launch("qe relu_mask"));
  TRY(sb.add(q.dhid_pre, H, 1, a->w1, 4 * H, 1, q.dv, 4 * H, B, 4 * H, H));
  if (a->g_w1) TRY(sb.add(q.dhid_pre, 1, H, q.v, 4 * H, 1, a->g_w1, 4 * H, H, 4 * H, B));
  if (a->g_b1) TRY(sb.add(q.dhid_pre, 1, H, q.one, 0, 0, a->g_b1, 1, H, 1, B));
  TRY(sb.launch(st));
  launch_k(qe_vscatter_kernel, B, 256, 0, st, q);
  TRY(check_launch("qe_vscatter"));
  }
  if (!(parts & 2)) return 0;
  const size_t smem_b = (2 * H * 32 + (LSTM_THREADS / 32) * 32 * 33) * sizeof(float);
  TRY(set_smem(reinterpret_cast<const void*>(lstm_bwd_kernel<true>), smem_b, "lstm_bwd"));
  TRY(set_smem(reinterpret_cast<const void*>(lstm_bwd_kernel<false>), smem_b, "lstm_bwd"));
  const dim3 grid_b(H / 32, BWD_JQ, 2 * q.BC);
  if (fits_cooperative(reinterpret_cast<const void*>(lstm_bwd_kernel<true>), grid_b, smem_b)) {
    cudaError_t me = cudaMemsetAsync(q.bar + 2 * QE_MAX_L, 0, 4 * QE_MAX_L * sizeof(unsigned), st);  // the step barriers' counters
    if (me != cudaSuccess) return fail(static_cast<int>(me), "drn_qe_backward memset: %s", cudaGetErrorString(me));
    int s0 = 0, s1 = L, nq = BWD_JQ;
    void* args[] = {&q, &s0, &s1, &nq};
    cudaError_t ce = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(lstm_bwd_kernel<true>), grid_b, dim3(LSTM_THREADS), args, smem_b, st);
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "lstm_bwd (cooperative): %s", cudaGetErrorString(ce));
  } else {
    for (int s = 0; s < L; ++s) {
      const int nq = s == 0 ? 1 : BWD_JQ;
      launch_k(lstm_bwd_kernel<false>, dim3(H / 32, nq, 2 * q.BC), LSTM_THREADS, smem_b, st, q, s, s + 1, nq);
      TRY(check_launch("lstm_bwd_step"));
    }
  }
  // The five big projections of the LSTM backward in ONE grouped tensor-core launch (split-BF16 planes written by the
  // kernels above): dW_ih, dW_hh of both directions (weight-gradient form over the R = B*L rows) and dE = dG [W_ih ; W_ih_r].
  {
    drn_gemm_t g[5];
    int n = 0;
    const drn_planes_t dGp = planes_view(q.dG_pl, static_cast<long long>(R) * 8 * H, R, 8 * H);
    for (int dir = 0; dir < 2; ++dir) {
      if (a->g_w_ih[dir]) {
        drn_gemm_t& w = g[n++] = qe_gemm_base(DRN_GEMM_WGRAD, R);
        w.a = dGp; w.a_c0 = dir * 4 * H; w.M = 4 * H;
        w.b = planes_view(q.E_pl, static_cast<long long>(R) * q.EP, R, q.EP);
        w.N = E;
        w.out = a->g_w_ih[dir]; w.out_ld = E; w.out_mode = DRN_OUT_ADD;
      }
      if (a->g_w_hh[dir]) {
        drn_gemm_t& w = g[n++] = qe_gemm_base(DRN_GEMM_WGRAD, R);
        w.a = dGp; w.a_c0 = dir * 4 * H; w.M = 4 * H;
        w.b = planes_view(q.Hprev_pl + static_cast<long long>(dir) * R * H, 2LL * R * H, R, H);
        w.N = H;
        w.out = a->g_w_hh[dir]; w.out_ld = H; w.out_mode = DRN_OUT_ADD;
      }
    }
    if (a->g_emb) {
      drn_gemm_t& d = g[n++] = qe_gemm_base(DRN_GEMM_ROWS, R);
      d.a = dGp;
      d.b = planes_view(q.Wih_pl, 8LL * H * q.EP, 8 * H, q.EP);
      d.b_mn = 1; d.N = E; d.K = 8 * H;
      d.out = q.dE; d.out_ld = E;
      d.split_k = DE_SLICES; d.out_split_stride = static_cast<long long>(R) * E;
    }
    if (n) TRY(drn_gemm_group(n, g, stream));
  }
  {
    SgemmBatch cb;  // LSTM bias gradients = column sums of dG (both biases of a direction receive the same sum)
    for (int dir = 0; dir < 2; ++dir) {
      const float* dG = q.dG + dir * 4 * H;  // [R] rows of stride 8H
      if (a->g_b_ih[dir]) TRY(cb.add(dG, 1, 8 * H, q.one, 0, 0, a->g_b_ih[dir], 1, 4 * H, 1, R));
      if (a->g_b_hh[dir]) TRY(cb.add(dG, 1, 8 * H, q.one, 0, 0, a->g_b_hh[dir], 1, 4 * H, 1, R));
    }
    TRY(cb.launch(st));
  }
  if (a->g_emb) {
    launch_k(qe_embed_bwd_kernel, R, 256, 0, st, q, a->g_emb);
    TRY(check_launch("qe_embed_bwd"));
  }
  return 0;
}
extern "C" int drn_qe_backward(const drn_qe_t* a, void* stream) { return qe_backward_parts(a, 3, stream); }
extern "C" int drn_qe_backward_part(const drn_qe_t* a, int part, void* stream) {
  if (part != 1 && part != 2) return fail(DRN_EINVAL, "drn_qe_backward_part: part must be 1 or 2 (got %d)", part);
  return qe_backward_parts(a, part, stream);
}


// ---- staging of the caller's query tensors into the static buffers the graphs read, with validation -----------------------
// The kernels index the embedding table with the token ids and the LSTM output with length-1 (qe_embed_kernel,
// qe_embed_bwd_kernel's atomicAdd into the gradient buffer, qe_vgather / qe_vscatter): an id outside [0, vocab) or a length
// outside [1, L] must never reach them.  The reference raises (IndexError in nn.Embedding, pack_padded_sequence's length check:
// language_module.py:41-42); here an invalid id is staged as 0 (= padding: zero embedding row semantics aside, no gradient), an
// invalid length is clamped, `err` gets the sticky bits 1 (token) / 2 (length) and `poison` = NaN for THIS batch (else 0): the
// caller adds it to the losses, so a bad batch is loud without a device synchronisation.  One CTA: B * L <= a few thousand.
__global__ void __launch_bounds__(256) qe_stage_kernel(const long long* __restrict__ tok, long long tok_ld, int ncols,
                                                       const long long* __restrict__ len, int B, int L, int vocab,
                                                       long long* __restrict__ tok_out, long long* __restrict__ len_out,
                                                       int* __restrict__ err, float* __restrict__ poison) {
  pdl_sync();
  int bad = 0;
  for (int i = threadIdx.x; i < B * L; i += blockDim.x) {
    const int b = i / L, t = i % L;
    long long v = (t < ncols) ? tok[static_cast<long long>(b) * tok_ld + t] : 0;
    if (v < 0 || v >= vocab) {
      bad |= 1;
      v = 0;
    }
    tok_out[i] = v;
  }
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    long long v = len[b];
    if (v < 1 || v > L || v > ncols) {
      bad |= 2;
      v = v < 1 ? 1 : (L < ncols ? L : ncols);
    }
    len_out[b] = v;
  }
  const int any1 = __syncthreads_or(bad & 1), any2 = __syncthreads_or(bad & 2);
  if (threadIdx.x == 0) {
    const int e = (any1 ? 1 : 0) | (any2 ? 2 : 0);
    if (e) atomicOr(err, e);
    poison[0] = e ? __int_as_float(0x7fc00000) : 0.f;
  }
}

extern "C" int drn_qe_stage(const int64_t* tokens, int64_t tok_ld, int ncols, const int64_t* lengths, int B, int L, int vocab,
                            int64_t* tokens_out, int64_t* lengths_out, int32_t* err, float* poison, void* stream) {
  if (!tokens || !lengths || !tokens_out || !lengths_out || !err || !poison) return fail(DRN_EINVAL, "drn_qe_stage: null pointer");
  if (B < 1 || L < 1 || ncols < 1 || tok_ld < ncols || vocab < 1)
    return fail(DRN_EINVAL, "drn_qe_stage: bad shape (B=%d L=%d ncols=%d tok_ld=%lld vocab=%d)", B, L, ncols, (long long)tok_ld, vocab);
  launch_k(qe_stage_kernel, 1, 256, 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const long long*>(tokens),
           static_cast<long long>(tok_ld), ncols, reinterpret_cast<const long long*>(lengths), B, L, vocab,
           reinterpret_cast<long long*>(tokens_out), reinterpret_cast<long long*>(lengths_out), reinterpret_cast<int*>(err), poison);
  return check_launch("qe_stage_kernel");
}
