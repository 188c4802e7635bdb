// Proposal feature pooling + padding: the step immediately BEFORE the hot path (reference dataset.py:105-155
// CharadesSTA.get_data, 180-206 collate_data).  For every proposal of every video of the batch: the
// contiguous range of feature windows it covers (integer arithmetic of dataset.py:126-145, clamped to the windows the file
// really has), an element-wise max over those rows, zero padding up to the longest proposal list.  Integer / max arithmetic
// only: bit-exact with the reference.  HBM-bound: one CTA per (proposal, video), 16-byte loads, every thread keeps 4 columns.
#include "common.cuh"

namespace drn {

__global__ void __launch_bounds__(256) pool_proposals_kernel(const float* __restrict__ feats, const long long* __restrict__ win_off,
                                                             const double* __restrict__ p_start, const int* __restrict__ p_end,
                                                             const int* __restrict__ nprops, const int* __restrict__ num_frames,
                                                             int P, int D, int window, int interval, float* __restrict__ out,
                                                             double* __restrict__ pse) {
  pdl_sync();
  const int p = blockIdx.x, b = blockIdx.y;
  float* o = out + (static_cast<long long>(b) * P + p) * D;
  const int D4 = D >> 2;
  if (p >= nprops[b]) {  // padding rows (dataset.py:188-189: torch.zeros)
    for (int c = threadIdx.x; c < D4; c += blockDim.x) reinterpret_cast<float4*>(o)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
      pse[(static_cast<long long>(b) * P + p) * 2] = 0.0;
      pse[(static_cast<long long>(b) * P + p) * 2 + 1] = 0.0;
    }
    return;
  }
  const double s = p_start[static_cast<long long>(b) * P + p];
  const int e = p_end[static_cast<long long>(b) * P + p];
  const int n_win = static_cast<int>(win_off[b + 1] - win_off[b]);
  const int ft_start_index = (static_cast<int>(s) / interval) * interval;  // dataset.py:126
  int lo = ft_start_index / interval, hi = lo;
  if (static_cast<double>(e) - s > static_cast<double>(window)) {          // dataset.py:128-139
    const int span = e - ft_start_index;
    const int count = span > 0 ? (span + interval - 1) / interval : 0;       // len(range(ft_start_index, e, interval))
    hi = lo + count - 1;
  }
  lo = min(n_win - 1, lo);                                                   // dataset.py:145
  hi = min(n_win - 1, hi);
  const float* f = feats + (win_off[b] + lo) * D;
  for (int c = threadIdx.x; c < D4; c += blockDim.x) {
    float4 m = reinterpret_cast<const float4*>(f)[c];
    for (int r = 1; r <= hi - lo; ++r) {
      const float4 v = reinterpret_cast<const float4*>(f + static_cast<long long>(r) * D)[c];
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
    reinterpret_cast<float4*>(o)[c] = m;
  }
  if (threadIdx.x == 0) {  // dataset.py:124: (start / num_frames, end / num_frames) in float64
    pse[(static_cast<long long>(b) * P + p) * 2] = s / static_cast<double>(num_frames[b]);
    pse[(static_cast<long long>(b) * P + p) * 2 + 1] = static_cast<double>(e) / static_cast<double>(num_frames[b]);
  }
}

}  // namespace drn

using namespace drn;

extern "C" int drn_pool_proposals(const float* feats, const int64_t* win_off, const double* p_start, const int32_t* p_end,
                                  const int32_t* nprops, const int32_t* num_frames, int B, int P, int D, int window, int interval,
                                  float* out_feats, double* out_pse, void* stream) {
  if (B < 1 || P < 1 || D < 4 || D % 4) return fail(DRN_EINVAL, "drn_pool_proposals: need B, P >= 1 and D %% 4 == 0 (D=%d)", D);
  if (window < 1 || interval < 1) return fail(DRN_EINVAL, "drn_pool_proposals: window / interval must be positive");
  if (reinterpret_cast<uintptr_t>(feats) % 16 || reinterpret_cast<uintptr_t>(out_feats) % 16)
    return fail(DRN_EINVAL, "drn_pool_proposals: feature buffers must be 16-byte aligned");
  launch_k(pool_proposals_kernel, dim3(P, B), 256, 0, static_cast<cudaStream_t>(stream), 
      feats, reinterpret_cast<const long long*>(win_off), p_start, p_end, nprops, num_frames, P, D, window, interval, out_feats,
      out_pse);
  return check_launch("pool_proposals");
}
