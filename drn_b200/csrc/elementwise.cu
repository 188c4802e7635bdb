// HBM-bound kernels of the path: plane splitting / weight packing, train-mode BatchNorm (+ReLU, +FPN upsample-add,
// +query gate) forward and backward, gate reductions.  All are coalesced 16-byte-vectorised streaming kernels; the
// per-channel reductions use per-thread fp32 partials -> shared-memory tree -> one fp64 atomic per channel per CTA.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace drn {

constexpr int EW_THREADS = 256;

__device__ __forceinline__ void load8(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// 8 bf16 hi + 8 bf16 lo -> 8 floats
__device__ __forceinline__ void load8_planes(const __nv_bfloat16* hi, long long plane_stride, float* v) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi);
  const uint4 l = *reinterpret_cast<const uint4*>(hi + plane_stride);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    v[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8_planes(__nv_bfloat16* hi, long long plane_stride, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v[2 * i], h0, l0);
    split_bf16(v[2 * i + 1], h1, l1);
    h[i] = pack_bf16x2(h0, h1);
    l[i] = pack_bf16x2(l0, l1);
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(hi + plane_stride) = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- split fp32 rows into planes ---------------------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) split_planes_kernel(const float* __restrict__ src, long long rows, int C8,
                                                                  long long src_ld, __nv_bfloat16* __restrict__ dst,
                                                                  long long dst_ld, int dst_col0, long long plane_stride) {
  pdl_sync();
  const long long total = rows * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C8;
    const int c = static_cast<int>(i % C8) * 8;
    float v[8];
    load8(src + r * src_ld + c, v);
    store8_planes(dst + r * dst_ld + dst_col0 + c, plane_stride, v);
  }
}

// ---- query gate on fp32 rows (model/backbone.py:28-30, level 0): dst planes[b,t,col0+c] = q[b,c] * x[b,t,c] ---------------
__global__ void __launch_bounds__(EW_THREADS) gate_planes_kernel(const float* __restrict__ x, const float* __restrict__ q, int T,
                                                                 int C8, long long total, __nv_bfloat16* __restrict__ dst,
                                                                 long long dst_ld, int dst_col0, long long plane_stride) {
  pdl_sync();
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C8;
    const int c = static_cast<int>(i % C8) * 8;
    const long long b = r / T;
    float v[8], g[8];
    load8(x + r * (C8 * 8LL) + c, v);
    load8(q + b * (C8 * 8LL) + c, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= g[j];
    store8_planes(dst + r * dst_ld + dst_col0 + c, plane_stride, v);
  }
}

// ---- conv / linear weights [O][C][k] fp32 -> tap-major planes [k][Ototal][C]; one launch packs a whole table ------------
constexpr int PACK_MAX_ITEMS = 24;
struct PackItem {
  const float* w;
  __nv_bfloat16* dst;
  float* grad;  // unpack only
  int O, C, k, Ototal, o0;
  long long plane_stride;
  int nslices;              // unpack only: partial sums to add up (K-splits x pyramid levels)
  long long slice_stride;
  int vec;                  // every pointer / stride of the item allows the 16-byte forms below (set by fill_table)
};
struct PackTable {
  int n;
  PackItem it[PACK_MAX_ITEMS];
};

__global__ void __launch_bounds__(EW_THREADS) pack_conv_weight_kernel(const PackTable tab) {
  pdl_sync();
  const PackItem& e = tab.it[blockIdx.y];
  const long long total = static_cast<long long>(e.O) * e.C;
  if (e.k == 1) {  // nn.Linear / 1x1 conv: straight split, 8 elements per thread
    const long long total8 = total >> 3;  // C % 8 == 0
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      float v[8];
      load8(e.w + i * 8, v);
      const long long o = (i * 8) / e.C, c = (i * 8) % e.C;
      store8_planes(e.dst + (e.o0 + o) * e.C + c, e.plane_stride, v);
    }
    return;
  }
  if (e.k == 3 && e.vec) {
    // A thread owns 4 consecutive input channels of one output channel: its 12 consecutive fp32 weights [c..c+3][tap] are three
    // 16-byte loads, the 4 hi + 4 lo bf16 of a tap two 8-byte stores (256 contiguous bytes per warp) -- the scalar form below
    // issues 12 strided 4-byte loads and 24 2-byte stores for the same elements.  Same values.
    const int C4 = e.C >> 2;
    const long long total4 = static_cast<long long>(e.O) * C4;
    for (long long i4 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i4 < total4;
         i4 += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int o = static_cast<int>(i4 / C4), c = static_cast<int>(i4 % C4) * 4;
      const float4* src = reinterpret_cast<const float4*>(e.w + (static_cast<long long>(o) * e.C + c) * 3);
      const float4 a = __ldg(src), b = __ldg(src + 1), d = __ldg(src + 2);
      const float w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w};  // [channel][tap]
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int cl = 0; cl < 4; ++cl) split_bf16(w[cl * 3 + r], h[cl], l[cl]);
        __nv_bfloat16* dst = e.dst + (static_cast<long long>(r) * e.Ototal + e.o0 + o) * e.C + c;
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        *reinterpret_cast<uint2*>(dst + e.plane_stride) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      }
    }
    return;
  }
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i / e.C), c = static_cast<int>(i % e.C);
    for (int r = 0; r < e.k; ++r) {
      __nv_bfloat16 h, l;
      split_bf16(e.w[i * e.k + r], h, l);
      const long long d = (static_cast<long long>(r) * e.Ototal + e.o0 + o) * e.C + c;
      e.dst[d] = h;
      e.dst[d + e.plane_stride] = l;
    }
  }
}

// ---- weight-gradient workspaces [slice][k][Ototal][C] -> parameter gradients [O][C][k], summing the slices ------------------
// (the K-splits of drn_gemm WGRAD and the three pyramid levels of a shared head conv store their partial sums side by side:
// deterministic, no atomics, no zero-fill).  Reads are coalesced over C; gridDim.y = table item.
// Vector form (k = 1 or 3, 16-byte aligned): a thread owns 4 consecutive input channels of one output channel -- 16-byte loads of
// the slices (same slice order: bit-identical sums), K 16-byte stores of the 4 K contiguous gradient values.
template <int K>
__device__ __forceinline__ void unpack_vec(const PackItem& e) {
  const int C4 = e.C >> 2;
  const long long total4 = static_cast<long long>(e.O) * C4;
  for (long long i4 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i4 < total4;
       i4 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i4 / C4), c = static_cast<int>(i4 % C4) * 4;
    float out[4 * K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const float* src = e.w + (static_cast<long long>(r) * e.Ototal + e.o0 + o) * e.C + c;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int sl = 0; sl < e.nslices; ++sl) {
        const float4 x = __ldcs(reinterpret_cast<const float4*>(src + sl * e.slice_stride));
        s4.x += x.x; s4.y += x.y; s4.z += x.z; s4.w += x.w;
      }
      out[0 * K + r] = s4.x; out[1 * K + r] = s4.y; out[2 * K + r] = s4.z; out[3 * K + r] = s4.w;
    }
    float4* g = reinterpret_cast<float4*>(e.grad + (static_cast<long long>(o) * e.C + c) * K);
#pragma unroll
    for (int j = 0; j < K; ++j) g[j] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
  }
}

__global__ void __launch_bounds__(EW_THREADS) unpack_conv_wgrad_kernel(const PackTable tab) {
  pdl_sync();
  const PackItem& e = tab.it[blockIdx.y];
  if (e.vec && e.k == 3) { unpack_vec<3>(e); return; }
  if (e.vec && e.k == 1) { unpack_vec<1>(e); return; }
  const long long total = static_cast<long long>(e.O) * e.C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i / e.C), c = static_cast<int>(i % e.C);
    for (int r = 0; r < e.k; ++r) {
      const float* src = e.w + (static_cast<long long>(r) * e.Ototal + e.o0 + o) * e.C + c;
      float acc = 0.f;
      for (int sl = 0; sl < e.nslices; ++sl) acc += src[sl * e.slice_stride];
      e.grad[i * e.k + r] = acc;
    }
  }
}

// ---- position feature (model/main_model.py:53-55): Linear(3 -> Cp) of (s, e, e-s), written as planes ----------------
__global__ void __launch_bounds__(EW_THREADS) pos_feature_kernel(const double* __restrict__ pse, const float* __restrict__ Wp,
                                                                 const float* __restrict__ bp, long long rows, int Cp,
                                                                 __nv_bfloat16* __restrict__ dst, long long dst_ld,
                                                                 int dst_col0, long long plane_stride,
                                                                 float* __restrict__ pos_in) {
  pdl_sync();
  const long long total = rows * Cp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / Cp;
    const int c = static_cast<int>(i % Cp);
    const double s = pse[2 * r], e = pse[2 * r + 1];
    const float f0 = static_cast<float>(s), f1 = static_cast<float>(e), f2 = static_cast<float>(e - s);
    if (c == 0 && pos_in) {
      pos_in[3 * r] = f0;
      pos_in[3 * r + 1] = f1;
      pos_in[3 * r + 2] = f2;
    }
    float v = bp[c];
    v = fmaf(f0, Wp[3 * c], v);
    v = fmaf(f1, Wp[3 * c + 1], v);
    v = fmaf(f2, Wp[3 * c + 2], v);
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    __nv_bfloat16* d = dst + r * dst_ld + dst_col0 + c;
    d[0] = h;
    d[plane_stride] = l;
  }
}

// ---- column statistics: block = 32 x 8 threads, 128 columns x ROWS_PER_BLOCK rows -------------------------------------
// The LAST block to finish (device-wide counter) finalises: MODE 0 -> BatchNorm coefficients + running statistics,
// MODE 1 -> mean(g), mean(g*xhat) + dgamma/dbeta accumulation.  It also re-zeroes the fp64 sums and the counter, so
// no memset / finalize launches are needed between uses.
constexpr int BN_MAX_PARTS = 2;

struct BnParts {  // the fused cls|bbox tower output carries two BatchNorm modules side by side
  int nparts;
  int c0[BN_MAX_PARTS], n[BN_MAX_PARTS];
  const float* gamma[BN_MAX_PARTS];
  const float* beta[BN_MAX_PARTS];
  float* running_mean[BN_MAX_PARTS];
  float* running_var[BN_MAX_PARTS];
  long long* nbt[BN_MAX_PARTS];
  float* dgamma[BN_MAX_PARTS];
  float* dbeta[BN_MAX_PARTS];
};

// One BatchNorm application (a conv block on one pyramid level).  Up to BN_MAX_JOBS independent applications -- the three
// levels of a shared head / FPN block -- are served by ONE launch (blockIdx.z / blockIdx.y = job).
constexpr int BN_MAX_JOBS = 3;
struct BnJob {
  float* y;
  const float* y2;  // optional second K-split slice of the conv output: the statistics pass folds it into y
  const float* da;
  long long rows;
  int B, T, C;
  float* coef;
  double* sums;
  unsigned* counter;
  float* bcoef;
  BnParts parts;
  const __nv_bfloat16* up;
  long long up_ps;
  const float* gate;
  __nv_bfloat16* out_a;
  long long a_ps;
  __nv_bfloat16* out_qa;
  long long qa_ps;
  __nv_bfloat16* dy;
  long long dy_ps;
  const float* partials;  // MODE 2: [prows][2][C] partial column sums written by the contraction epilogue
  long long prows;
};
struct BnJobs {
  int n;
  BnJob j[BN_MAX_JOBS];
};

// coef rows: 0 scale, 1 shift, 2 batch mean, 3 invstd, 4 unbiased batch variance (what the running average absorbs)
__device__ __forceinline__ void bn_coef_from_stats(double mean, double var, float gamma, float beta, float eps, int c, int C,
                                                   float* coef) {
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float scale = gamma * invstd;
  coef[c] = scale;
  coef[C + c] = beta - static_cast<float>(mean) * scale;
  coef[2 * C + c] = static_cast<float>(mean);
  coef[3 * C + c] = invstd;
}

// MODE 0: sum y, sum y^2 ; 1: BN backward sums (sum g, sum g*xhat) with g = relu-masked da ; 2: as 0, but reduced from the
// per-32-row partial sums the contraction epilogue wrote (J.partials: 1/16 of the bytes of y)
template <int MODE>
__global__ void __launch_bounds__(256) col_stats_kernel(const BnJobs jobs, float momentum, float eps, int update_running,
                                                        int STAT_ROWS) {
  pdl_sync();
  __shared__ float red[2][8][128];
  __shared__ bool is_last;
  const BnJob& J = jobs.j[blockIdx.z];
  float* __restrict__ y = J.y;
  const float* __restrict__ da = J.da;
  const long long rows = (MODE == 2) ? J.prows : J.rows;  // rows this kernel walks (MODE 2: rows of the partial-sum buffer)
  const int C = J.C;
  float* __restrict__ coef = J.coef;
  double* __restrict__ sums = J.sums;
  unsigned* __restrict__ counter = J.counter;
  float* __restrict__ bcoef = J.bcoef;
  const BnParts& parts = J.parts;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 128 + tx * 4;
  const long long r0 = static_cast<long long>(blockIdx.y) * STAT_ROWS;
  if (blockIdx.x * 128 >= C || r0 >= rows) return;  // block outside this job's extent (grid is sized for the largest job)
  float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  if (c < C) {
    float sc[4], sh[4], mu[4], is[4];
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sc[j] = coef[c + j];
        sh[j] = coef[C + c + j];
        mu[j] = coef[2 * C + c + j];
        is[j] = coef[3 * C + c + j];
      }
    }
    constexpr int U = 4;  // rows in flight per thread
    for (int i0 = ty; i0 < STAT_ROWS; i0 += 8 * U) {
      float4 v4[U], g4[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + i0 + 8 * u;
        ok[u] = (i0 + 8 * u < STAT_ROWS) && (r < rows);
        v4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        g4[u] = v4[u];
        if (ok[u] && MODE == 2) {
          v4[u] = *reinterpret_cast<const float4*>(J.partials + r * 2 * C + c);
          g4[u] = *reinterpret_cast<const float4*>(J.partials + r * 2 * C + C + c);
        } else if (ok[u]) {
          v4[u] = *reinterpret_cast<const float4*>(y + r * C + c);
          if (MODE == 0 && J.y2) {  // y <- y + y2 (K-split slices of the contraction), summed once, here
            const float4 w4 = *reinterpret_cast<const float4*>(J.y2 + r * C + c);
            v4[u].x += w4.x; v4[u].y += w4.y; v4[u].z += w4.z; v4[u].w += w4.w;
            *reinterpret_cast<float4*>(y + r * C + c) = v4[u];
          }
          if (MODE == 1) g4[u] = *reinterpret_cast<const float4*>(da + r * C + c);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const float v[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
        if (MODE == 2) {
          s0[0] += v4[u].x; s0[1] += v4[u].y; s0[2] += v4[u].z; s0[3] += v4[u].w;
          s1[0] += g4[u].x; s1[1] += g4[u].y; s1[2] += g4[u].z; s1[3] += g4[u].w;
        } else if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s0[j] += v[j];
            s1[j] = fmaf(v[j], v[j], s1[j]);
          }
        } else {
          const float g[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float gm = (fmaf(v[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
            s0[j] += gm;
            s1[j] = fmaf(gm, (v[j] - mu[j]) * is[j], s1[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][ty][tx * 4 + j] = s0[j];
    red[1][ty][tx * 4 + j] = s1[j];
  }
  __syncthreads();
  const int t = ty * 32 + tx;  // 256 threads: 2 x 128 outputs
  const int which = t >> 7, col = t & 127;
  if (blockIdx.x * 128 + col < C) {
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += static_cast<double>(red[which][i][col]);
    atomicAdd(sums + static_cast<long long>(which) * C + blockIdx.x * 128 + col, acc);
  }
  // ---- last block finalises ----
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned total = static_cast<unsigned>((C + 127) / 128) * static_cast<unsigned>((rows + STAT_ROWS - 1) / STAT_ROWS);
    is_last = (atomicAdd(counter, 1u) == total - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double n = static_cast<double>(J.rows);
  const long long nrows = J.rows;
  for (int p = 0; p < parts.nparts; ++p) {
    for (int i = t; i < parts.n[p]; i += 256) {
      const int cc = parts.c0[p] + i;
      const double a0 = __ldcg(sums + cc), a1 = __ldcg(sums + C + cc);
      sums[cc] = 0.0;
      sums[C + cc] = 0.0;
      if (MODE != 1) {
        const double mean = a0 / n;
        double var = a1 / n - mean * mean;
        if (var < 0) var = 0;
        const double unbiased = (nrows > 1) ? var * n / (n - 1.0) : var;
        if (update_running) {
          parts.running_mean[p][i] = (1.f - momentum) * parts.running_mean[p][i] + momentum * static_cast<float>(mean);
          parts.running_var[p][i] = (1.f - momentum) * parts.running_var[p][i] + momentum * static_cast<float>(unbiased);
        }
        coef[4 * C + cc] = static_cast<float>(unbiased);
        bn_coef_from_stats(mean, var, parts.gamma[p][i], parts.beta[p][i], eps, cc, C, coef);
      } else {
        bcoef[cc] = static_cast<float>(a0 / n);
        bcoef[C + cc] = static_cast<float>(a1 / n);
        if (parts.dgamma[p]) {  // atomics: the levels of a shared head block finalise concurrently in one launch
          atomicAdd(parts.dbeta[p] + i, static_cast<float>(a0));
          atomicAdd(parts.dgamma[p] + i, static_cast<float>(a1));
        }
      }
    }
    if (MODE != 1 && update_running && t == 0 && parts.nbt[p]) parts.nbt[p][0] += 1;
  }
  if (t == 0) *counter = 0u;
}

// ---- ordered running-statistics update of jobs that SHARE BatchNorm modules (the head applied to three pyramid levels,
// model/fcos.py:93-102): running <- (1-m) running + m stat, job after job, exactly the order of the reference's level loop.
__global__ void bn_running_update_kernel(const BnJobs jobs, float momentum) {
  pdl_sync();
  const int p = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < jobs.j[0].parts.n[p]; i += gridDim.x * blockDim.x) {
    for (int k = 0; k < jobs.n; ++k) {
      const BnJob& J = jobs.j[k];
      if (p >= J.parts.nparts) continue;
      const int cc = J.parts.c0[p] + i;
      float* rm = J.parts.running_mean[p] + i;
      float* rv = J.parts.running_var[p] + i;
      *rm = (1.f - momentum) * *rm + momentum * J.coef[2 * J.C + cc];
      *rv = (1.f - momentum) * *rv + momentum * J.coef[4 * J.C + cc];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int k = 0; k < jobs.n; ++k)
      if (p < jobs.j[k].parts.nparts && jobs.j[k].parts.nbt[p]) jobs.j[k].parts.nbt[p][0] += 1;
}

// ---- eval-mode BN: coefficients from the running statistics ----------------------------------------------------------------
__global__ void bn_eval_coef_kernel(const BnJobs jobs, float eps) {
  pdl_sync();
  const BnJob& J = jobs.j[blockIdx.y];
  const int C = J.C;
  const BnParts& parts = J.parts;
  float* __restrict__ coef = J.coef;
  for (int p = 0; p < parts.nparts; ++p)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < parts.n[p]; i += gridDim.x * blockDim.x)
      bn_coef_from_stats(parts.running_mean[p][i], parts.running_var[p][i], parts.gamma[p][i], parts.beta[p][i], eps,
                         parts.c0[p] + i, C, coef);
}

// ---- BN apply + ReLU (+ nearest x2 upsample add, + query gate) -> planes ---------------------------------------------
__global__ void __launch_bounds__(EW_THREADS) bn_relu_apply_kernel(const BnJobs jobs) {
  pdl_sync();
  const BnJob& J = jobs.j[blockIdx.y];
  const float* __restrict__ y = J.y;
  const int B = J.B, T = J.T, C = J.C;
  const float* __restrict__ coef = J.coef;
  const __nv_bfloat16* __restrict__ up = J.up;
  const long long up_ps = J.up_ps;
  const float* __restrict__ gate = J.gate;
  __nv_bfloat16* __restrict__ out_a = J.out_a;
  const long long a_ps = J.a_ps;
  __nv_bfloat16* __restrict__ out_qa = J.out_qa;
  const long long qa_ps = J.qa_ps;
  const int C8 = C >> 3;
  const long long total = static_cast<long long>(B) * T * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / C8;
    const int c = static_cast<int>(i % C8) * 8;
    const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
    float v[8], sc[8], sh[8];
    load8(y + row * C + c, v);
    load8(coef + c, sc);
    load8(coef + C + c, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
    if (up) {
      float u[8];
      load8_planes(up + (static_cast<long long>(b) * (T >> 1) + (t >> 1)) * C + c, up_ps, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += u[j];
    }
    if (out_a) store8_planes(out_a + row * C + c, a_ps, v);
    if (out_qa) {
      float q[8];
      load8(gate + static_cast<long long>(b) * C + c, q);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] *= v[j];
      store8_planes(out_qa + row * C + c, qa_ps, q);
    }
  }
}

// ---- BN backward apply: dy = scale * (g - mean(g) - xhat * mean(g*xhat)) -> planes ------------------------------------
__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_kernel(const BnJobs jobs) {
  pdl_sync();
  const BnJob& J = jobs.j[blockIdx.y];
  const float* __restrict__ da = J.da;
  const float* __restrict__ y = J.y;
  const long long rows = J.rows;
  const int C = J.C;
  const float* __restrict__ coef = J.coef;
  const float* __restrict__ bcoef = J.bcoef;
  __nv_bfloat16* __restrict__ dy = J.dy;
  const long long dy_ps = J.dy_ps;
  const int C8 = C >> 3;
  const long long total = rows * C8;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / C8;
    const int c = static_cast<int>(i % C8) * 8;
    float v[8], g[8], sc[8], sh[8], mu[8], is[8], mg[8], mgx[8];
    load8(y + row * C + c, v);
    load8(da + row * C + c, g);
    load8(coef + c, sc);
    load8(coef + C + c, sh);
    load8(coef + 2 * C + c, mu);
    load8(coef + 3 * C + c, is);
    load8(bcoef + c, mg);
    load8(bcoef + C + c, mgx);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gm = (fmaf(v[j], sc[j], sh[j]) > 0.f) ? g[j] : 0.f;
      const float xh = (v[j] - mu[j]) * is[j];
      v[j] = sc[j] * (gm - mg[j] - xh * mgx[j]);
    }
    store8_planes(dy + row * C + c, dy_ps, v);
  }
}

// ---- row-walking forms of the two BatchNorm apply kernels ---------------------------------------------------------------
// The flat forms above re-load the per-channel coefficients (2 resp. 6 vectors) for every 8 elements, which made them L1-bound
// (ncu r01 v12: L1/TEX 78 % / 90 % busy at 3.5 / 4.1 TB/s of DRAM traffic) and spend a quarter of their instructions on 64-bit
// index divisions.  Here a thread owns 8 CHANNELS (C/8 a power of two <= 256), keeps their coefficients in registers and walks
// `rpt` rows (256 / (C/8) rows per pass of the CTA), two rows in flight.
struct RowWalk {
  int c, rs, RS;
  long long r0;
};
__device__ __forceinline__ RowWalk row_walk(int C, int rpt) {
  const int C8 = C >> 3;
  RowWalk w;
  w.c = (threadIdx.x & (C8 - 1)) * 8;
  w.rs = threadIdx.x / C8;
  w.RS = EW_THREADS / C8;
  w.r0 = static_cast<long long>(blockIdx.x) * (w.RS * rpt);
  return w;
}
__global__ void __launch_bounds__(EW_THREADS, 3) bn_relu_apply_rows_kernel(const BnJobs jobs, int rpt) {
  pdl_sync();
  const BnJob& J = jobs.j[blockIdx.y];
  const int T = J.T, C = J.C;
  const long long rows = static_cast<long long>(J.B) * T;
  const RowWalk w = row_walk(C, rpt);
  if (w.r0 >= rows) return;
  const float* __restrict__ y = J.y;
  const __nv_bfloat16* __restrict__ up = J.up;
  const float* __restrict__ gate = J.gate;
  __nv_bfloat16* __restrict__ out_a = J.out_a;
  __nv_bfloat16* __restrict__ out_qa = J.out_qa;
  const int c = w.c;
  float sc[8], sh[8], q[8];
  load8(J.coef + c, sc);
  load8(J.coef + C + c, sh);
  int qb = -1;
  for (int i = 0; i < rpt; i += 2) {
    long long row[2];
    bool ok[2];
    float v[2][8], u[2][8];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      row[e] = w.r0 + static_cast<long long>(i + e) * w.RS + w.rs;
      ok[e] = (i + e < rpt) && (row[e] < rows);
      if (ok[e]) {
        load8(y + row[e] * C + c, v[e]);
        if (up) {
          const int b = static_cast<int>(row[e] / T), t = static_cast<int>(row[e] - static_cast<long long>(b) * T);
          load8_planes(up + (static_cast<long long>(b) * (T >> 1) + (t >> 1)) * C + c, J.up_ps, u[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (!ok[e]) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[e][j] = fmaxf(fmaf(v[e][j], sc[j], sh[j]), 0.f);
      if (up) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[e][j] += u[e][j];
      }
      if (out_a) store8_planes(out_a + row[e] * C + c, J.a_ps, v[e]);
      if (out_qa) {
        const int b = static_cast<int>(row[e] / T);
        if (b != qb) {  // the gate is per sample: re-read only when the walk crosses into the next sample
          load8(gate + static_cast<long long>(b) * C + c, q);
          qb = b;
        }
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = q[j] * v[e][j];
        store8_planes(out_qa + row[e] * C + c, J.qa_ps, g);
      }
    }
  }
}
__global__ void __launch_bounds__(EW_THREADS) bn_bwd_apply_rows_kernel(const BnJobs jobs, int rpt) {
  pdl_sync();
  const BnJob& J = jobs.j[blockIdx.y];
  const int C = J.C;
  const long long rows = J.rows;
  const RowWalk w = row_walk(C, rpt);
  if (w.r0 >= rows) return;
  const float* __restrict__ da = J.da;
  const float* __restrict__ y = J.y;
  __nv_bfloat16* __restrict__ dy = J.dy;
  const int c = w.c;
  // dy = scale * (g - mean(g) - xhat * mean(g xhat)), xhat = (y - mu) * invstd: k1 = invstd * mean(g xhat) per channel
  float sc[8], sh[8], mu[8], mg[8], k1[8];
  load8(J.coef + c, sc);
  load8(J.coef + C + c, sh);
  load8(J.coef + 2 * C + c, mu);
  load8(J.bcoef + c, mg);
  {
    float is[8], mgx[8];
    load8(J.coef + 3 * C + c, is);
    load8(J.bcoef + C + c, mgx);
#pragma unroll
    for (int j = 0; j < 8; ++j) k1[j] = is[j] * mgx[j];
  }
  for (int i = 0; i < rpt; i += 2) {
    long long row[2];
    bool ok[2];
    float v[2][8], g[2][8];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      row[e] = w.r0 + static_cast<long long>(i + e) * w.RS + w.rs;
      ok[e] = (i + e < rpt) && (row[e] < rows);
      if (ok[e]) {
        load8(y + row[e] * C + c, v[e]);
        load8(da + row[e] * C + c, g[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (!ok[e]) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gm = (fmaf(v[e][j], sc[j], sh[j]) > 0.f) ? g[e][j] : 0.f;
        v[e][j] = sc[j] * (gm - mg[j] - (v[e][j] - mu[j]) * k1[j]);
      }
      store8_planes(dy + row[e] * C + c, J.dy_ps, v[e]);
    }
  }
}
// rows per thread of the row-walking kernels (0: some job's channel count does not fit them -> flat kernels); *gx = grid.x
static int rows_walk_plan(const BnJobs& t, unsigned* gx) {
  long long gmax = 0;
  static int rpt_env = -1;  // DRN_BN_RPT: rows per thread (tuning; cold-cache ncu favours fewer rows = more CTAs, r01 v17)
  if (rpt_env < 0) {
    const char* e = getenv("DRN_BN_RPT");
    rpt_env = e ? atoi(e) : 0;
  }
  const int rpt = (rpt_env >= 1 && rpt_env <= 64) ? rpt_env : 8;
  for (int i = 0; i < t.n; ++i) {
    const int C8 = t.j[i].C / 8;
    if (t.j[i].C % 8 || C8 < 1 || C8 > EW_THREADS || (C8 & (C8 - 1))) return 0;
    const long long per_cta = static_cast<long long>(EW_THREADS / C8) * rpt;
    const long long rows = t.j[i].rows > 0 ? t.j[i].rows : static_cast<long long>(t.j[i].B) * t.j[i].T;
    const long long g = (rows + per_cta - 1) / per_cta;
    gmax = g > gmax ? g : gmax;
  }
  *gx = static_cast<unsigned>(gmax > 0 ? gmax : 1);
  return rpt;
}

// ---- FPN backward of nearest x2 upsample: dst[b,j,:] += src[b,2j,:] + src[b,2j+1,:] ------------------------------------
__global__ void __launch_bounds__(EW_THREADS) pair_sum_add_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                                  long long rows_half, int C) {
  pdl_sync();
  const int C4 = C >> 2;
  const long long total = rows_half * C4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C4;
    const int c = static_cast<int>(i % C4) * 4;
    float4 d = *reinterpret_cast<const float4*>(dst + r * C + c);
    const float4 a = *reinterpret_cast<const float4*>(src + (2 * r) * C + c);
    const float4 b = *reinterpret_cast<const float4*>(src + (2 * r + 1) * C + c);
    d.x += a.x + b.x; d.y += a.y + b.y; d.z += a.z + b.z; d.w += a.w + b.w;
    *reinterpret_cast<float4*>(dst + r * C + c) = d;
  }
}

// ---- gate gradient: dq[b,c] (+)= sum_t g[b,t,c] * a[b,t,c]; block = 32x8 threads, 128 columns of one sample -----------
// A_PLANES: a is a planes tensor (hi+lo) else fp32.  Optionally also writes dP planes = q[b,c]*g and column sums of dP.
template <bool A_PLANES>
__global__ void __launch_bounds__(256) gate_reduce_kernel(const float* __restrict__ g, long long g_ld,
                                                          const void* __restrict__ a, long long a_ld, long long a_ps, int T,
                                                          int C, int t_chunk, float* __restrict__ dq,
                                                          const float* __restrict__ q, __nv_bfloat16* __restrict__ dp,
                                                          long long dp_ps, float* __restrict__ dbias) {
  pdl_sync();
  __shared__ float red[2][8][128];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 128 + tx * 4;
  const int b = blockIdx.z;
  const int t0 = blockIdx.y * t_chunk;
  float s[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
  if (c < C) {
    float qv[4] = {0, 0, 0, 0};
    if (dp) {
#pragma unroll
      for (int j = 0; j < 4; ++j) qv[j] = q[static_cast<long long>(b) * C + c + j];
    }
    for (int i = ty; i < t_chunk; i += 8) {
      const int t = t0 + i;
      if (t >= T) break;
      const long long row = static_cast<long long>(b) * T + t;
      const float4 g4 = *reinterpret_cast<const float4*>(g + row * g_ld + c);
      const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
      float av[4];
      if (A_PLANES) {
        const __nv_bfloat16* ah = static_cast<const __nv_bfloat16*>(a) + row * a_ld + c;
        const uint2 h = *reinterpret_cast<const uint2*>(ah);
        const uint2 l = *reinterpret_cast<const uint2*>(ah + a_ps);
        av[0] = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
        av[1] = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
        av[2] = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
        av[3] = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
      } else {
        const float4 a4 = *reinterpret_cast<const float4*>(static_cast<const float*>(a) + row * a_ld + c);
        av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] = fmaf(gv[j], av[j], s[j]);
      if (dp) {
        float pv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pv[j] = gv[j] * qv[j];
          sb[j] += pv[j];
        }
        __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
        split_bf16(pv[0], h0, l0); split_bf16(pv[1], h1, l1); split_bf16(pv[2], h2, l2); split_bf16(pv[3], h3, l3);
        __nv_bfloat16* d = dp + row * C + c;
        *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16x2(h0, h1), pack_bf16x2(h2, h3));
        *reinterpret_cast<uint2*>(d + dp_ps) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][ty][tx * 4 + j] = s[j];
    red[1][ty][tx * 4 + j] = sb[j];
  }
  __syncthreads();
  const int t = ty * 32 + tx;
  const int which = t >> 7, col = t & 127;
  const int cc = blockIdx.x * 128 + col;
  if (cc < C) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += red[which][i][col];
    if (which == 0) atomicAdd(dq + static_cast<long long>(b) * C + cc, acc);
    else if (dbias) atomicAdd(dbias + cc, acc);
  }
}

// ---- position-feature backward: dWp[c][j] += sum_rows dpos[row,c]*pos_in[row,j], dbp[c] += sum_rows dpos[row,c] ----------
__global__ void __launch_bounds__(256) pos_bwd_kernel(const float* __restrict__ dx, long long dx_ld, int col0,
                                                      const float* __restrict__ pos_in, long long rows, int Cp,
                                                      int rows_per_block, float* __restrict__ dWp, float* __restrict__ dbp) {
  pdl_sync();
  const int c = threadIdx.x;  // Cp <= 256 threads
  if (c >= Cp) return;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  float a0 = 0, a1 = 0, a2 = 0, ab = 0;
  constexpr int U = 8;  // rows in flight per thread
  for (int i0 = 0; i0 < rows_per_block; i0 += U) {
    float g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + i0 + u;
      g[u] = (i0 + u < rows_per_block && r < rows) ? dx[r * dx_ld + col0 + c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + i0 + u;
      if (i0 + u >= rows_per_block || r >= rows) break;
      a0 = fmaf(g[u], pos_in[3 * r], a0);
      a1 = fmaf(g[u], pos_in[3 * r + 1], a1);
      a2 = fmaf(g[u], pos_in[3 * r + 2], a2);
      ab += g[u];
    }
  }
  atomicAdd(dWp + 3 * c, a0);
  atomicAdd(dWp + 3 * c + 1, a1);
  atomicAdd(dWp + 3 * c + 2, a2);
  atomicAdd(dbp + c, ab);
}

// ---- column sums (bias gradients of the small Linear layers): out[c] += sum_rows x[row,c] ------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long rows, int C, long long ld,
                                                     float* __restrict__ out) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) acc += x[r * ld + c];
  atomicAdd(out + c, acc);
}

static inline int ew_grid(long long total) {
  long long g = (total + EW_THREADS - 1) / EW_THREADS;
  const long long cap = 148LL * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace drn

using namespace drn;

#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int drn_split_planes(const float* src, int64_t rows, int C, int64_t src_ld, void* dst, int64_t dst_ld,
                                int dst_col0, int64_t dst_plane_stride, void* stream) {
  if (C % 8 || src_ld % 4 || dst_ld % 8 || dst_col0 % 8 || dst_plane_stride % 8) return fail(DRN_EINVAL, "drn_split_planes: alignment (C=%d)", C);
  if (rows <= 0) return 0;
  launch_k(split_planes_kernel, ew_grid(rows * (C / 8)), EW_THREADS, 0, ST(stream), src, rows, C / 8, src_ld, static_cast<__nv_bfloat16*>(dst),
                                                                              dst_ld, dst_col0, dst_plane_stride);
  return check_launch("split_planes");
}

extern "C" int drn_gate_planes(const float* x, const float* q, int B, int T, int C, void* dst, int64_t dst_ld, int dst_col0,
                               int64_t dst_plane_stride, void* stream) {
  if (C % 8 || dst_ld % 8 || dst_col0 % 8 || dst_plane_stride % 8) return fail(DRN_EINVAL, "drn_gate_planes: alignment (C=%d)", C);
  const long long total = static_cast<long long>(B) * T * (C / 8);
  if (total <= 0) return 0;
  launch_k(gate_planes_kernel, ew_grid(total), EW_THREADS, 0, ST(stream), x, q, T, C / 8, total, static_cast<__nv_bfloat16*>(dst), dst_ld,
                                                                    dst_col0, dst_plane_stride);
  return check_launch("gate_planes");
}

static int fill_table(PackTable* t, int n, const drn_pack_item_t* items) {
  if (n < 1 || n > PACK_MAX_ITEMS) return fail(DRN_EINVAL, "pack table: 1..%d items (got %d)", PACK_MAX_ITEMS, n);
  t->n = n;
  for (int i = 0; i < n; ++i) {
    if (items[i].C % 8) return fail(DRN_EINVAL, "pack table: C %% 8 (item %d, C=%d)", i, items[i].C);
    t->it[i] = PackItem{items[i].src, static_cast<__nv_bfloat16*>(items[i].planes), items[i].grad, items[i].O, items[i].C,
                        items[i].k, items[i].Ototal, items[i].o0, items[i].plane_stride,
                        items[i].nslices < 1 ? 1 : items[i].nslices, items[i].slice_stride, 0};
    static int v2 = -1;  // DRN_PACK_V2=0: scalar forms only (A/B)
    if (v2 < 0) {
      const char* e = getenv("DRN_PACK_V2");
      v2 = e ? atoi(e) : 1;
    }
    auto al16 = [](const void* p) { return p == nullptr || reinterpret_cast<uintptr_t>(p) % 16 == 0; };
    t->it[i].vec = (v2 != 0 && al16(items[i].src) && al16(items[i].planes) && al16(items[i].grad) &&
                    items[i].plane_stride % 8 == 0 && items[i].slice_stride % 4 == 0) ? 1 : 0;
  }
  return 0;
}

extern "C" int drn_pack_conv_weights(int n, const drn_pack_item_t* items, void* stream) {
  PackTable t;
  int rc = fill_table(&t, n, items);
  if (rc) return rc;
  launch_k(pack_conv_weight_kernel, dim3(148, n), EW_THREADS, 0, ST(stream), t);
  return check_launch("pack_conv_weights");
}

extern "C" int drn_unpack_conv_wgrads(int n, const drn_pack_item_t* items, void* stream) {
  PackTable t;
  int rc = fill_table(&t, n, items);
  if (rc) return rc;
  launch_k(unpack_conv_wgrad_kernel, dim3(148 * 2, n), EW_THREADS, 0, ST(stream), t);
  return check_launch("unpack_conv_wgrads");
}

extern "C" int drn_pos_feature(const double* pse, const float* Wp, const float* bp, int64_t rows, int Cp, void* dst,
                               int64_t dst_ld, int dst_col0, int64_t plane_stride, float* pos_in, void* stream) {
  launch_k(pos_feature_kernel, ew_grid(rows * Cp), EW_THREADS, 0, ST(stream), pse, Wp, bp, rows, Cp, static_cast<__nv_bfloat16*>(dst),
                                                                        dst_ld, dst_col0, plane_stride, pos_in);
  return check_launch("pos_feature");
}

static int fill_parts(BnParts* bp, int nparts, const drn_bn_part_t* parts) {
  if (nparts < 1 || nparts > BN_MAX_PARTS) return fail(DRN_EINVAL, "BatchNorm: 1..%d parameter parts supported (got %d)", BN_MAX_PARTS, nparts);
  bp->nparts = nparts;
  for (int i = 0; i < nparts; ++i) {
    bp->c0[i] = parts[i].c0;
    bp->n[i] = parts[i].n;
    bp->gamma[i] = parts[i].gamma;
    bp->beta[i] = parts[i].beta;
    bp->running_mean[i] = parts[i].running_mean;
    bp->running_var[i] = parts[i].running_var;
    bp->nbt[i] = reinterpret_cast<long long*>(parts[i].num_batches_tracked);
    bp->dgamma[i] = parts[i].dgamma;
    bp->dbeta[i] = parts[i].dbeta;
  }
  return 0;
}

static int fill_jobs(BnJobs* t, int n, const drn_bn_job_t* jobs, const char* who) {
  if (n < 1 || n > BN_MAX_JOBS) return fail(DRN_EINVAL, "%s: 1..%d jobs (got %d)", who, BN_MAX_JOBS, n);
  t->n = n;
  for (int i = 0; i < n; ++i) {
    const drn_bn_job_t& s = jobs[i];
    BnJob& d = t->j[i];
    if (s.C % 8 || s.B < 1 || s.T < 1) return fail(DRN_EINVAL, "%s: job %d needs C %% 8 == 0 and B, T >= 1 (C=%d)", who, i, s.C);
    d.y = const_cast<float*>(s.y); d.y2 = s.y2; d.da = s.da;
    d.rows = static_cast<long long>(s.B) * s.T;
    d.B = s.B; d.T = s.T; d.C = s.C;
    d.coef = s.coef; d.sums = s.sums; d.counter = s.counter; d.bcoef = s.bcoef;
    int rc = fill_parts(&d.parts, s.nparts, s.parts);
    if (rc) return rc;
    d.up = static_cast<const __nv_bfloat16*>(s.up); d.up_ps = s.up_plane_stride;
    d.gate = s.gate;
    d.out_a = static_cast<__nv_bfloat16*>(s.out_a); d.a_ps = s.a_plane_stride;
    d.out_qa = static_cast<__nv_bfloat16*>(s.out_qa); d.qa_ps = s.qa_plane_stride;
    d.dy = static_cast<__nv_bfloat16*>(s.dy); d.dy_ps = s.dy_plane_stride;
    d.partials = s.partials; d.prows = s.partial_rows;
  }
  return 0;
}

// Rows per statistics block.  Measured on B200 (scripts/ab_bench.sh, r01): 64 rows (more, smaller blocks) beats 256 by ~1.5 % of the
// step -- the fp64 atomics that end each block are not the bottleneck.  DRN_STAT_ROWS = 64 | 128 | 256 overrides (tuning).
static int stats_grid(const BnJobs& t, dim3* grid) {
  int cmax = 0;
  long long rmax = 0, elems = 0;
  for (int i = 0; i < t.n; ++i) {
    cmax = t.j[i].C > cmax ? t.j[i].C : cmax;
    rmax = t.j[i].rows > rmax ? t.j[i].rows : rmax;
    elems += t.j[i].rows * t.j[i].C;
  }
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("DRN_STAT_ROWS");
    forced = e ? atoi(e) : 0;
  }
  int sr = 64;
  (void)elems;
  if (forced == 64 || forced == 128 || forced == 256) sr = forced;
  *grid = dim3(ceil_div(cmax, 128), static_cast<unsigned>((rmax + sr - 1) / sr), t.n);
  return sr;
}
static int ew_grid_jobs(const BnJobs& t) {
  long long m = 0;
  for (int i = 0; i < t.n; ++i) {
    const long long tot = t.j[i].rows * (t.j[i].C / 8);
    m = tot > m ? tot : m;
  }
  return ew_grid(m);
}

extern "C" int drn_bn_stats_multi(int n, const drn_bn_job_t* jobs, float momentum, float eps, int training, void* stream) {
  BnJobs t{};
  int rc = fill_jobs(&t, n, jobs, "drn_bn_stats_multi");
  if (rc) return rc;
  if (!training) {
    int cmax = 0;
    for (int i = 0; i < n; ++i) cmax = t.j[i].C > cmax ? t.j[i].C : cmax;
    launch_k(bn_eval_coef_kernel, dim3(ceil_div(cmax, 256), n), 256, 0, ST(stream), t, eps);
    return check_launch("bn_eval_coef");
  }
  int with_partials = 0;
  for (int i = 0; i < n; ++i) with_partials += t.j[i].partials ? 1 : 0;
  if (with_partials) {  // statistics from the partial sums of the contraction epilogue (drn_gemm_t.stats): y is not read
    if (with_partials != n) return fail(DRN_EINVAL, "drn_bn_stats_multi: partial sums must be given for all jobs of a launch or for none");
    int cmax = 0;
    long long pmax = 0;
    for (int i = 0; i < n; ++i) {
      if (t.j[i].y2 || t.j[i].prows < 1) return fail(DRN_EINVAL, "drn_bn_stats_multi: job %d: partial sums exclude y2 and need partial_rows >= 1", i);
      cmax = t.j[i].C > cmax ? t.j[i].C : cmax;
      pmax = t.j[i].prows > pmax ? t.j[i].prows : pmax;
    }
    const int sr = 64;
    launch_k(col_stats_kernel<2>, dim3(ceil_div(cmax, 128), static_cast<unsigned>((pmax + sr - 1) / sr), n), dim3(32, 8), 0, ST(stream), t,
             momentum, eps, training == 1 ? 1 : 0, sr);
    return check_launch("bn_stats(partials)");
  }
  dim3 grid;
  const int sr = stats_grid(t, &grid);
  launch_k(col_stats_kernel<0>, grid, dim3(32, 8), 0, ST(stream), t, momentum, eps, training == 1 ? 1 : 0, sr);
  return check_launch("bn_stats");
}

extern "C" int drn_bn_running_update(int n, const drn_bn_job_t* jobs, float momentum, void* stream) {
  BnJobs t{};
  int rc = fill_jobs(&t, n, jobs, "drn_bn_running_update");
  if (rc) return rc;
  for (int i = 1; i < n; ++i)
    if (t.j[i].parts.nparts != t.j[0].parts.nparts) return fail(DRN_EINVAL, "drn_bn_running_update: jobs must share their BatchNorm modules");
  int nmax = 0;
  for (int p = 0; p < t.j[0].parts.nparts; ++p) nmax = t.j[0].parts.n[p] > nmax ? t.j[0].parts.n[p] : nmax;
  launch_k(bn_running_update_kernel, dim3(ceil_div(nmax, 256), t.j[0].parts.nparts), 256, 0, ST(stream), t, momentum);
  return check_launch("bn_running_update");
}

extern "C" int drn_bn_relu_apply_multi(int n, const drn_bn_job_t* jobs, void* stream) {
  BnJobs t{};
  int rc = fill_jobs(&t, n, jobs, "drn_bn_relu_apply_multi");
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    if (t.j[i].up && (t.j[i].T % 2)) return fail(DRN_EINVAL, "drn_bn_relu_apply: upsample-add needs even T");
    if (t.j[i].out_qa && !t.j[i].gate) return fail(DRN_EINVAL, "drn_bn_relu_apply: gated output without gate");
  }
  unsigned gx = 0;
  const int rpt = rows_walk_plan(t, &gx);
  if (rpt) launch_k(bn_relu_apply_rows_kernel, dim3(gx, n), EW_THREADS, 0, ST(stream), t, rpt);
  else launch_k(bn_relu_apply_kernel, dim3(ew_grid_jobs(t), n), EW_THREADS, 0, ST(stream), t);
  return check_launch("bn_relu_apply");
}

extern "C" int drn_bn_bwd_reduce_multi(int n, const drn_bn_job_t* jobs, void* stream) {
  BnJobs t{};
  int rc = fill_jobs(&t, n, jobs, "drn_bn_bwd_reduce_multi");
  if (rc) return rc;
  dim3 grid;
  const int sr = stats_grid(t, &grid);
  launch_k(col_stats_kernel<1>, grid, dim3(32, 8), 0, ST(stream), t, 0.f, 0.f, 0, sr);
  return check_launch("bn_bwd_reduce");
}

extern "C" int drn_bn_bwd_apply_multi(int n, const drn_bn_job_t* jobs, void* stream) {
  BnJobs t{};
  int rc = fill_jobs(&t, n, jobs, "drn_bn_bwd_apply_multi");
  if (rc) return rc;
  unsigned gx = 0;
  const int rpt = rows_walk_plan(t, &gx);
  if (rpt) launch_k(bn_bwd_apply_rows_kernel, dim3(gx, n), EW_THREADS, 0, ST(stream), t, rpt);
  else launch_k(bn_bwd_apply_kernel, dim3(ew_grid_jobs(t), n), EW_THREADS, 0, ST(stream), t);
  return check_launch("bn_bwd_apply");
}

// single-application forms
static drn_bn_job_t one_job(const float* y, int64_t rows, int B, int T, int C, int nparts, const drn_bn_part_t* parts) {
  drn_bn_job_t j{};
  j.y = y;
  j.B = B > 0 ? B : 1;
  j.T = B > 0 ? T : static_cast<int>(rows);
  j.C = C;
  j.nparts = nparts;
  for (int i = 0; i < nparts && i < 2; ++i) j.parts[i] = parts[i];
  return j;
}
extern "C" int drn_bn_stats(const float* y, int64_t rows, int C, int nparts, const drn_bn_part_t* parts, float momentum, float eps,
                            int training, float* coef, double* sums, unsigned* counter, void* stream) {
  if (nparts < 1 || nparts > 2) return fail(DRN_EINVAL, "drn_bn_stats: 1..2 parts");
  drn_bn_job_t j = one_job(y, rows, 0, 0, C, nparts, parts);
  j.coef = coef; j.sums = sums; j.counter = counter;
  return drn_bn_stats_multi(1, &j, momentum, eps, training, stream);
}
extern "C" int drn_bn_relu_apply(const float* y, int B, int T, int C, const float* coef, const void* up, int64_t up_plane_stride,
                                 const float* gate, void* out_a, int64_t a_plane_stride, void* out_qa, int64_t qa_plane_stride,
                                 void* stream) {
  static const drn_bn_part_t none{};
  drn_bn_job_t j = one_job(y, 0, B, T, C, 1, &none);
  j.coef = const_cast<float*>(coef);
  j.up = up; j.up_plane_stride = up_plane_stride; j.gate = gate;
  j.out_a = out_a; j.a_plane_stride = a_plane_stride; j.out_qa = out_qa; j.qa_plane_stride = qa_plane_stride;
  return drn_bn_relu_apply_multi(1, &j, stream);
}
extern "C" int drn_bn_bwd_reduce(const float* da, const float* y, int64_t rows, int C, float* coef, int nparts,
                                 const drn_bn_part_t* parts, double* sums, unsigned* counter, float* bcoef, void* stream) {
  if (nparts < 1 || nparts > 2) return fail(DRN_EINVAL, "drn_bn_bwd_reduce: 1..2 parts");
  drn_bn_job_t j = one_job(y, rows, 0, 0, C, nparts, parts);
  j.da = da; j.coef = coef; j.sums = sums; j.counter = counter; j.bcoef = bcoef;
  return drn_bn_bwd_reduce_multi(1, &j, stream);
}
extern "C" int drn_bn_bwd_apply(const float* da, const float* y, int64_t rows, int C, const float* coef, const float* bcoef,
                                void* dy, int64_t dy_plane_stride, void* stream) {
  static const drn_bn_part_t none{};
  drn_bn_job_t j = one_job(y, rows, 0, 0, C, 1, &none);
  j.da = da; j.coef = const_cast<float*>(coef); j.bcoef = const_cast<float*>(bcoef);
  j.dy = dy; j.dy_plane_stride = dy_plane_stride;
  return drn_bn_bwd_apply_multi(1, &j, stream);
}

extern "C" int drn_pair_sum_add(float* dst, const float* src, int64_t rows_half, int C, void* stream) {
  if (C % 4) return fail(DRN_EINVAL, "drn_pair_sum_add: C %% 4");
  launch_k(pair_sum_add_kernel, ew_grid(rows_half * (C / 4)), EW_THREADS, 0, ST(stream), dst, src, rows_half, C);
  return check_launch("pair_sum_add");
}

extern "C" int drn_gate_reduce(const float* g, int64_t g_ld, const void* a, int64_t a_ld, int64_t a_plane_stride, int a_is_planes,
                               int B, int T, int C, float* dq, const float* q, void* dp, int64_t dp_plane_stride, float* dbias,
                               void* stream) {
  if (C % 4 || g_ld % 4 || a_ld % 4) return fail(DRN_EINVAL, "drn_gate_reduce: alignment");
  const int t_chunk = 64;
  dim3 grid(ceil_div(C, 128), ceil_div(T, t_chunk), B);
  if (a_is_planes)
    launch_k(gate_reduce_kernel<true>, grid, dim3(32, 8), 0, ST(stream), g, g_ld, a, a_ld, a_plane_stride, T, C, t_chunk, dq, q,
                                                                  static_cast<__nv_bfloat16*>(dp), dp_plane_stride, dbias);
  else
    launch_k(gate_reduce_kernel<false>, grid, dim3(32, 8), 0, ST(stream), g, g_ld, a, a_ld, a_plane_stride, T, C, t_chunk, dq, q,
                                                                   static_cast<__nv_bfloat16*>(dp), dp_plane_stride, dbias);
  return check_launch("gate_reduce");
}

extern "C" int drn_pos_bwd(const float* dx, int64_t dx_ld, int col0, const float* pos_in, int64_t rows, int Cp, float* dWp,
                           float* dbp, void* stream) {
  if (Cp > 256) return fail(DRN_EINVAL, "drn_pos_bwd: Cp > 256");
  const int rpb = 32;
  launch_k(pos_bwd_kernel, static_cast<unsigned>((rows + rpb - 1) / rpb), 256, 0, ST(stream), dx, dx_ld, col0, pos_in, rows, Cp, rpb, dWp,
                                                                                       dbp);
  return check_launch("pos_bwd");
}

extern "C" int drn_colsum(const float* x, int64_t rows, int C, int64_t ld, float* out, void* stream) {
  dim3 grid(ceil_div(C, 256), static_cast<unsigned>(rows < 64 ? rows : 64));
  launch_k(colsum_kernel, grid, 256, 0, ST(stream), x, rows, C, ld, out);
  return check_launch("colsum");
}
