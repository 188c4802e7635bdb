// Language-guided pooling (reference model/LGP.py:29-51) as a standalone fused op, forward and backward.  The reference
// module is dead code (imported by nothing) but part of the query-video fusion named by the task; SURVEY.md section 8f-4.
//
//   q' = BN(conv1x1(query tiled over t))          -> constant along t: computed once per (sample, channel)
//   s[b,j,p] = sum_c x[b,c,2j+p] * q'[b,c]        (j < t/2, p in {0,1})
//   a = softmax_p(s);   out[b,c,j] = a[b,j,0] x[b,c,2j] + a[b,j,1] x[b,c,2j+1]
//
// Layout is the reference's channels-first [B, C, t] (t contiguous): one thread per (sample, pair j), float2 loads coalesced
// over j, the channel loop strides by t.  HBM-bound: x is read twice (second pass from L2), out written once.
#include "common.cuh"
#include "ptx.cuh"

namespace drn {

// BatchNorm1d over the tiled query: statistics over (B, t) of a tensor constant along t = statistics over B; the running
// variance absorbs the UNBIASED variance with n = B*t elements (what torch computes on the tiled tensor).
__global__ void lgp_bn_kernel(const float* __restrict__ z, int B, int C, int t, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* running_mean, float* running_var, long long* nbt,
                              float momentum, float eps, int training, float* __restrict__ qn, float* __restrict__ xhat,
                              float* __restrict__ invstd_out) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    double s = 0.0, ss = 0.0;
    for (int b = 0; b < B; ++b) {
      const double v = z[static_cast<long long>(b) * C + c];
      s += v;
      ss += v * v;
    }
    const double m = s / B;
    double vb = ss / B - m * m;
    if (vb < 0) vb = 0;
    mean = static_cast<float>(m);
    var = static_cast<float>(vb);
    const double n = static_cast<double>(B) * t;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(n > 1 ? vb * n / (n - 1.0) : vb);
    if (c == 0 && nbt) nbt[0] += 1;
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float is = 1.f / sqrtf(var + eps);
  invstd_out[c] = is;
  for (int b = 0; b < B; ++b) {
    const float xh = (z[static_cast<long long>(b) * C + c] - mean) * is;
    xhat[static_cast<long long>(b) * C + c] = xh;
    qn[static_cast<long long>(b) * C + c] = xh * gamma[c] + beta[c];
  }
}

// grid (ceil(t/2 / 128), B), block 128: thread = pair j
__global__ void __launch_bounds__(128) lgp_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ qn, int C, int t,
                                                           float* __restrict__ att, float* __restrict__ out) {
  pdl_sync();
  extern __shared__ float qs[];  // [C]
  const int b = blockIdx.y, h = t >> 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) qs[c] = qn[static_cast<long long>(b) * C + c];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= h) return;
  const float* xb = x + static_cast<long long>(b) * C * t + 2 * j;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float2 v = *reinterpret_cast<const float2*>(xb + static_cast<long long>(c) * t);
    s0 = fmaf(v.x, qs[c], s0);
    s1 = fmaf(v.y, qs[c], s1);
  }
  const float m = fmaxf(s0, s1);
  const float e0 = expf(s0 - m), e1 = expf(s1 - m);
  const float a0 = e0 / (e0 + e1), a1 = e1 / (e0 + e1);
  att[(static_cast<long long>(b) * h + j) * 2] = a0;
  att[(static_cast<long long>(b) * h + j) * 2 + 1] = a1;
  float* ob = out + static_cast<long long>(b) * C * h + j;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float2 v = *reinterpret_cast<const float2*>(xb + static_cast<long long>(c) * t);
    ob[static_cast<long long>(c) * h] = a0 * v.x + a1 * v.y;
  }
}

// backward of the pooling: dx, and dqn[b,c] += sum_j ds[b,j,p] x[b,c,2j+p] (block partial sums over its 128 pairs -> atomics)
__global__ void __launch_bounds__(128) lgp_pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ qn,
                                                           const float* __restrict__ att, const float* __restrict__ dout, int C,
                                                           int t, float* __restrict__ dx, float* __restrict__ dqn) {
  pdl_sync();
  extern __shared__ float qs[];  // [C]
  __shared__ float red[4];
  const int b = blockIdx.y, h = t >> 1;
  for (int c = threadIdx.x; c < C; c += blockDim.x) qs[c] = qn[static_cast<long long>(b) * C + c];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < h;
  const int jj = live ? j : 0;
  const float* xb = x + static_cast<long long>(b) * C * t + 2 * jj;
  const float* db = dout + static_cast<long long>(b) * C * h + jj;
  float a0 = 0.f, a1 = 0.f, da0 = 0.f, da1 = 0.f;
  if (live) {
    a0 = att[(static_cast<long long>(b) * h + j) * 2];
    a1 = att[(static_cast<long long>(b) * h + j) * 2 + 1];
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      const float2 v = *reinterpret_cast<const float2*>(xb + static_cast<long long>(c) * t);
      const float d = db[static_cast<long long>(c) * h];
      da0 = fmaf(d, v.x, da0);
      da1 = fmaf(d, v.y, da1);
    }
  }
  const float dot = a0 * da0 + a1 * da1;
  const float ds0 = a0 * (da0 - dot), ds1 = a1 * (da1 - dot);  // softmax backward
  float* dxb = dx + static_cast<long long>(b) * C * t + 2 * jj;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = 0; c < C; ++c) {
    float part = 0.f;
    if (live) {
      const float2 v = *reinterpret_cast<const float2*>(xb + static_cast<long long>(c) * t);
      const float d = db[static_cast<long long>(c) * h];
      *reinterpret_cast<float2*>(dxb + static_cast<long long>(c) * t) = make_float2(a0 * d + ds0 * qs[c], a1 * d + ds1 * qs[c]);
      part = ds0 * v.x + ds1 * v.y;
    }
    part = warp_sum(part);
    if (lane == 0) red[w] = part;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(dqn + static_cast<long long>(b) * C + c, red[0] + red[1] + red[2] + red[3]);
    __syncthreads();
  }
}

// BatchNorm backward over the batch applied to dqn (the tiling over t cancels): dz, dgamma, dbeta
__global__ void lgp_bn_bwd_kernel(const float* __restrict__ dqn, const float* __restrict__ xhat, const float* __restrict__ invstd,
                                  const float* __restrict__ gamma, int B, int C, int training, float* __restrict__ dz,
                                  float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f, sx = 0.f;
  for (int b = 0; b < B; ++b) {
    const float g = dqn[static_cast<long long>(b) * C + c];
    s += g;
    sx = fmaf(g, xhat[static_cast<long long>(b) * C + c], sx);
  }
  if (dgamma) atomicAdd(dgamma + c, sx);
  if (dbeta) atomicAdd(dbeta + c, s);
  const float k = gamma[c] * invstd[c];
  for (int b = 0; b < B; ++b) {
    const float g = dqn[static_cast<long long>(b) * C + c];
    dz[static_cast<long long>(b) * C + c] =
        training ? k * (g - s / B - xhat[static_cast<long long>(b) * C + c] * sx / B) : k * g;
  }
}

}  // namespace drn

using namespace drn;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int drn_lgp_bn(const float* z, int B, int C, int t, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, int64_t* num_batches_tracked, float momentum, float eps, int training, float* qn,
                          float* xhat, float* invstd, void* stream) {
  if (B < 1 || C < 1 || t < 2) return fail(DRN_EINVAL, "drn_lgp_bn: bad shape");
  launch_k(lgp_bn_kernel, ceil_div(C, 128), 128, 0, ST(stream), z, B, C, t, gamma, beta, running_mean, running_var,
                                                         reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, training, qn,
                                                         xhat, invstd);
  return check_launch("lgp_bn");
}

extern "C" int drn_lgp_pool_fwd(const float* x, const float* qn, int B, int C, int t, float* att, float* out, void* stream) {
  if (B < 1 || C < 1 || t < 2 || t % 2) return fail(DRN_EINVAL, "drn_lgp_pool_fwd: t must be even (t=%d)", t);
  if (C * sizeof(float) > 48 * 1024) return fail(DRN_EINVAL, "drn_lgp_pool_fwd: at most 12288 channels");
  launch_k(lgp_pool_fwd_kernel, dim3(ceil_div(t / 2, 128), B), 128, C * sizeof(float), ST(stream), x, qn, C, t, att, out);
  return check_launch("lgp_pool_fwd");
}

extern "C" int drn_lgp_pool_bwd(const float* x, const float* qn, const float* att, const float* dout, int B, int C, int t, float* dx,
                                float* dqn, void* stream) {
  if (B < 1 || C < 1 || t < 2 || t % 2) return fail(DRN_EINVAL, "drn_lgp_pool_bwd: t must be even (t=%d)", t);
  if (C * sizeof(float) > 48 * 1024) return fail(DRN_EINVAL, "drn_lgp_pool_bwd: at most 12288 channels");
  cudaError_t e = cudaMemsetAsync(dqn, 0, sizeof(float) * B * C, ST(stream));
  if (e != cudaSuccess) return fail(static_cast<int>(e), "drn_lgp_pool_bwd memset: %s", cudaGetErrorString(e));
  launch_k(lgp_pool_bwd_kernel, dim3(ceil_div(t / 2, 128), B), 128, C * sizeof(float), ST(stream), x, qn, att, dout, C, t, dx, dqn);
  return check_launch("lgp_pool_bwd");
}

extern "C" int drn_lgp_bn_bwd(const float* dqn, const float* xhat, const float* invstd, const float* gamma, int B, int C, int training,
                              float* dz, float* dgamma, float* dbeta, void* stream) {
  if (B < 1 || C < 1) return fail(DRN_EINVAL, "drn_lgp_bn_bwd: bad shape");
  launch_k(lgp_bn_bwd_kernel, ceil_div(C, 128), 128, 0, ST(stream), dqn, xhat, invstd, gamma, B, C, training, dz, dgamma, dbeta);
  return check_launch("lgp_bn_bwd");
}
