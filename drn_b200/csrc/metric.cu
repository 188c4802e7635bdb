// Batched temporal NMS + recall@k: the step immediately AFTER the hot path in evaluation (reference
// utils/evaluate_utils.py:132-190 compute_IoU_recall_top_n_ours, 192-215 nms_temporal, 232-236 calculate_IoU; driven by
// main.py:450-454 with iou 0.5, topk {1, 5}, temporal_nms=True).  The reference runs it as pure-Python loops over Python
// floats, query by query, O(n^2) per query; here one warp owns one query and the whole test set is one launch.
//
// Arithmetic is IEEE double, operation by operation as the reference evaluates it on Python floats (fp32 detections widen
// exactly; no expression here contains a multiply-add, so nothing can be contracted): the picks are BIT-EXACT.
//   order     stable ascending sort by score, best = last (evaluate_utils.py:200) == argmax of (score, index): among equal
//             scores the LATER candidate is picked first.  (The descending pre-sort of _postprocess_raw_results_no_merge, line 97,
//             is stable too, so it does not change which detections are picked, only how they are numbered.)
//   suppress  candidate j survives pick i iff inter / (len_i + len_j - inter) <= overlap, inter = max(0, min(e) - max(s))
//   hit       any of the first k picks has (min(e) - max(s)) / (max(e) - min(s)) >= iou against the ground truth (unclamped)
// Zero-length detections make the reference divide by zero (ZeroDivisionError); they are dropped before the NMS, as in
// oracle/metrics.py.  A query without any candidate gets the reference's fallback detection (0, 1) with score 1
// (model/inference.py:192-197) when `empty_fallback` is set.
#include <math_constants.h>

#include "common.cuh"

namespace drn {

constexpr int NMS_MAX_N = 256;  // candidates per query (DRN: 3 levels x 32)
constexpr int NMS_PER_LANE = NMS_MAX_N / 32;
constexpr int NMS_WARPS = 4;

__global__ void __launch_bounds__(NMS_WARPS * 32) nms_recall_kernel(
    const float* __restrict__ det, const float* __restrict__ score, const int* __restrict__ count, const double* __restrict__ gt,
    int Q, int G, int K, int nms, double overlap, double iou_thr, const int* __restrict__ topk, int ntopk, int empty_fallback,
    int* __restrict__ picks, int* __restrict__ npicks, int* __restrict__ hits, int* __restrict__ correct) {
  pdl_sync();
  __shared__ double sx1[NMS_WARPS][NMS_MAX_N], sx2[NMS_WARPS][NMS_MAX_N];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * NMS_WARPS + w;
  if (q >= Q) return;
  const int N = G * K;
  double x1[NMS_PER_LANE], x2[NMS_PER_LANE];
  float s[NMS_PER_LANE];
  bool alive[NMS_PER_LANE];
  int total = 0;
#pragma unroll
  for (int r = 0; r < NMS_PER_LANE; ++r) {
    const int i = r * 32 + lane;
    bool v = i < N;
    if (v) v = (i % K) < count[q * G + i / K];
    x1[r] = v ? static_cast<double>(det[(static_cast<long long>(q) * N + i) * 2]) : 0.0;
    x2[r] = v ? static_cast<double>(det[(static_cast<long long>(q) * N + i) * 2 + 1]) : 0.0;
    s[r] = v ? score[static_cast<long long>(q) * N + i] : 0.f;
    total += v ? 1 : 0;
    alive[r] = v && (x2[r] - x1[r] > 0.0);
    sx1[w][i] = x1[r];
    sx2[w][i] = x2[r];
  }
  total = __reduce_add_sync(0xffffffffu, total);
  if (total == 0 && empty_fallback) {  // inference.py:192-197: one detection (0, 1), score 1
    if (lane == 0) {
      x1[0] = 0.0; x2[0] = 1.0; s[0] = 1.f; alive[0] = true;
      sx1[w][0] = 0.0; sx2[w][0] = 1.0;
    }
  }
  __syncwarp();
  const double gs = gt[2 * q], ge = gt[2 * q + 1];
  int np = 0, first_hit = 0x7fffffff;
  int* my_picks = picks ? picks + static_cast<long long>(q) * N : nullptr;
  while (true) {
    // best live candidate: highest score; among equal scores the LATER one with NMS (nms_temporal takes the last element of a
    // stable ascending sort), the EARLIER one without (picks = the stable descending order itself, evaluate_utils.py:97,165)
    float bs = -CUDART_INF_F;
    int bi = -1;
#pragma unroll
    for (int r = 0; r < NMS_PER_LANE; ++r) {
      const int i = r * 32 + lane;
      if (alive[r] && (bi < 0 || s[r] > bs || (s[r] == bs && (nms ? i > bi : i < bi)))) {
        bs = s[r];
        bi = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi >= 0 && (bi < 0 || os > bs || (os == bs && (nms ? oi > bi : oi < bi)))) {
        bs = os;
        bi = oi;
      }
    }
    if (bi < 0) break;
    const double px1 = sx1[w][bi], px2 = sx2[w][bi], plen = px2 - px1;
    if (lane == 0 && my_picks) my_picks[np] = bi;
    const double giou = (fmin(ge, px2) - fmax(gs, px1)) / (fmax(ge, px2) - fmin(gs, px1));
    if (giou >= iou_thr && np < first_hit) first_hit = np;
    ++np;
#pragma unroll
    for (int r = 0; r < NMS_PER_LANE; ++r) {
      if (!alive[r]) continue;
      if (r * 32 + lane == bi) {
        alive[r] = false;
        continue;
      }
      if (!nms) continue;
      const double inter = fmax(0.0, fmin(px2, x2[r]) - fmax(px1, x1[r]));
      const double o = inter / (plen + (x2[r] - x1[r]) - inter);
      if (!(o <= overlap)) alive[r] = false;
    }
  }
  if (lane == 0) {
    if (my_picks)
      for (int i = np; i < N; ++i) my_picks[i] = -1;
    if (npicks) npicks[q] = np;
    for (int t = 0; t < ntopk; ++t) {
      const int h = first_hit < topk[t] ? 1 : 0;
      if (hits) hits[q * ntopk + t] = h;
      if (h && correct) atomicAdd(correct + t, 1);
    }
  }
}

}  // namespace drn

using namespace drn;

extern "C" int drn_nms_recall(const float* det, const float* score, const int32_t* count, const double* gt, int Q, int G, int K,
                              int nms, double overlap, double iou_thr, const int32_t* topk, int ntopk, int empty_fallback, int32_t* picks,
                              int32_t* npicks, int32_t* hits, int32_t* correct, void* stream) {
  if (!det || !score || !count || !gt || !topk) return fail(DRN_EINVAL, "drn_nms_recall: null input");
  if (Q < 1 || G < 1 || K < 1 || G * K > NMS_MAX_N)
    return fail(DRN_EINVAL, "drn_nms_recall: need Q >= 1 and 1 <= groups x slots <= %d (G=%d, K=%d)", NMS_MAX_N, G, K);
  if (ntopk < 1 || ntopk > 8) return fail(DRN_EINVAL, "drn_nms_recall: 1..8 top-k values (got %d)", ntopk);
  launch_k(nms_recall_kernel, (Q + NMS_WARPS - 1) / NMS_WARPS, NMS_WARPS * 32, 0, static_cast<cudaStream_t>(stream), det, score,
           count, gt, Q, G, K, nms ? 1 : 0, overlap, iou_thr, topk, ntopk, empty_fallback, picks, npicks, hits, correct);
  return check_launch("nms_recall_kernel");
}
