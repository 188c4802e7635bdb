// Fused gradient-norm clipping + Adam (reference main.py:239-244: torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
// followed by torch.optim.Adam(lr).step()).  After the hot path every step touches all 45.5 M
// parameters: ~1.1 GB of HBM traffic as ~10 foreach kernels + a host-synchronising norm in stock PyTorch.  Here: one
// norm kernel + one update kernel over a chunk table (multi-tensor apply), no host synchronisation.  HBM-bound.
#include "common.cuh"
#include "ptx.cuh"

namespace drn {

constexpr int OPT_CHUNK = 16384;  // elements per CTA

struct AdamItemDev {
  float* param;
  const float* grad;  // null: no gradient this step (skipped by the norm and by the update)
  float* exp_avg;
  float* exp_avg_sq;
  long long numel;
  int update;         // 0: contributes to the norm only (a parameter outside the optimizer, main.py:124-138 stage 2)
};

// scratch[0] = sum of squares (double), scratch[1] = total norm, scratch[2] = clip coefficient (written by the update kernel)
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const AdamItemDev* __restrict__ items, const int* __restrict__ chunk_item,
                                                          const long long* __restrict__ chunk_off, double* __restrict__ scratch,
                                                          int* __restrict__ steps) {
  pdl_sync();
  const AdamItemDev it = items[chunk_item[blockIdx.x]];
  if (!it.grad) return;
  const long long o = chunk_off[blockIdx.x];
  // torch.optim.Adam counts steps PER PARAMETER (a parameter without gradient is skipped and keeps its count): the first chunk of
  // every updated tensor advances its counter here, one launch before the update kernel reads it
  if (o == 0 && threadIdx.x == 0 && it.update) steps[chunk_item[blockIdx.x]] += 1;
  const long long n = min(static_cast<long long>(OPT_CHUNK), it.numel - o);
  const float* g = it.grad + o;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const long long n4 = n >> 2;
    for (long long i = threadIdx.x; i < n4; i += 256) {
      const float4 v = reinterpret_cast<const float4*>(g)[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 256) acc += g[i] * g[i];
  } else {
    for (long long i = threadIdx.x; i < n; i += 256) acc += g[i] * g[i];
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += static_cast<double>(red[w]);
    atomicAdd(scratch, s);
  }
}

__global__ void __launch_bounds__(256) clip_adam_kernel(const AdamItemDev* __restrict__ items, const int* __restrict__ chunk_item,
                                                        const long long* __restrict__ chunk_off, double* __restrict__ scratch,
                                                        float max_norm, float lr, float beta1, float beta2, float eps,
                                                        const int* __restrict__ steps) {
  pdl_sync();
  const AdamItemDev it = items[chunk_item[blockIdx.x]];
  // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
  const float total_norm = static_cast<float>(sqrt(scratch[0]));
  float coef = max_norm > 0.f ? max_norm / (total_norm + 1e-6f) : 1.f;
  coef = fminf(coef, 1.f);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    scratch[1] = total_norm;
    scratch[2] = coef;
  }
  if (!it.grad || !it.update) return;
  const long long o = chunk_off[blockIdx.x];
  const long long n = min(static_cast<long long>(OPT_CHUNK), it.numel - o);
  const int step = steps[chunk_item[blockIdx.x]];
  const float bias_c1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
  const float bias_c2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
  const float step_size = lr / bias_c1;
  for (long long i = threadIdx.x; i < n; i += 256) {
    const float g = it.grad[o + i] * coef;
    // torch.optim.Adam (no amsgrad, no weight decay): exp_avg.lerp_(g, 1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2);
    // denom = exp_avg_sq.sqrt() / sqrt(bias_c2) + eps; param.addcdiv_(exp_avg, denom, value = -lr / bias_c1)
    const float m0 = it.exp_avg[o + i];
    const float m = m0 + (g - m0) * (1.f - beta1);
    const float v = it.exp_avg_sq[o + i] * beta2 + (1.f - beta2) * g * g;
    it.exp_avg[o + i] = m;
    it.exp_avg_sq[o + i] = v;
    const float denom = sqrtf(v) / bias_c2_sqrt + eps;
    it.param[o + i] = it.param[o + i] - step_size * (m / denom);
  }
}

}  // namespace drn

using namespace drn;

extern "C" int drn_clip_adam(int nchunks, const void* items, const int32_t* chunk_item, const int64_t* chunk_off, double* scratch,
                             int32_t* steps, float max_norm, float lr, float beta1, float beta2, float eps, void* stream) {
  if (nchunks < 1 || !items || !chunk_item || !chunk_off || !scratch || !steps) return fail(DRN_EINVAL, "drn_clip_adam: null table");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(scratch, 0, 3 * sizeof(double), st);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "drn_clip_adam memset: %s", cudaGetErrorString(e));
  const AdamItemDev* it = static_cast<const AdamItemDev*>(items);
  launch_k(grad_sqnorm_kernel, nchunks, 256, 0, st, it, chunk_item, reinterpret_cast<const long long*>(chunk_off), scratch, steps);
  int rc = check_launch("grad_sqnorm");
  if (rc) return rc;
  launch_k(clip_adam_kernel, nchunks, 256, 0, st, it, chunk_item, reinterpret_cast<const long long*>(chunk_off), scratch, max_norm, lr,
                                            beta1, beta2, eps, steps);
  return check_launch("clip_adam");
}

extern "C" int drn_clip_adam_chunk(void) { return OPT_CHUNK; }
