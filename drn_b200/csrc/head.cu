// FCOS head projections (the N = 1 / 2 / 1 skinny convs of model/fcos.py:41-69, on CUDA cores), closed-form target
// assignment (model/loss.py:90-127), sigmoid focal loss (model/layers/sigmoid_focal_loss.py:40-52, stable log-sigmoid),
// IoU regression loss (model/layers/iou_loss.py:5-24), the stage-2/3 IoU-score branch (model/loss.py:168-198) and
// all of their backward passes.  One thread per location; scalars are reduced block-wise and accumulated in fp64.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace drn {

constexpr int MAX_LEVELS = 4;

struct LevelGeom {
  int nlevels;
  int B;
  int T[MAX_LEVELS];      // locations per sample on each level
  int off[MAX_LEVELS];    // cumulative locations per sample before this level
  float stride[MAX_LEVELS];
  float lo[MAX_LEVELS], hi[MAX_LEVELS];  // size-of-interest bands (model/loss.py:47-51)
  int P;                  // total locations per sample
};

// ---- skinny conv forward: out[b,t,o] = bias[o] + sum_r sum_c W[o][c][r] * X[b, t+r-pad, c0+c]; one warp per (b,t) ------
template <int NOUT, int K>
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long x_ps, int x_ld, int c0,
                                                         int Cw, int B, int T, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ out) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * T) return;
  const int b = warp / T, t = warp % T;
  constexpr int PAD = (K - 1) / 2;
  float acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = 0.f;
#pragma unroll
  for (int r = 0; r < K; ++r) {
    const int ts = t + r - PAD;
    if (ts < 0 || ts >= T) continue;
    const __nv_bfloat16* row = x + (static_cast<long long>(b) * T + ts) * x_ld + c0;
    for (int c = lane * 8; c < Cw; c += 256) {
      const uint4 h = *reinterpret_cast<const uint4*>(row + c);
      const uint4 l = *reinterpret_cast<const uint4*>(row + c + x_ps);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
      float v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
        v[2 * i + 1] = __uint_as_float(hw[i] & 0xffff0000u) + __uint_as_float(lw[i] & 0xffff0000u);
      }
#pragma unroll
      for (int o = 0; o < NOUT; ++o)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o] = fmaf(v[j], __ldg(W + (static_cast<long long>(o) * Cw + c + j) * K + r), acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const float s = warp_sum(acc[o]);
    if (lane == 0) out[static_cast<long long>(warp) * NOUT + o] = s + bias[o];
  }
}

// ---- skinny conv backward: dX[b,t,c0+c] = sum_o sum_r d[b,t+pad-r,o] W[o][c][r];  dW[o][c][r] += sum X[b,t,c] d[b,t+pad-r,o]
// one thread per channel PAIR (4-byte plane loads, 8-byte stores), a block walks a chunk of rows.
template <int NOUT, int K>
__device__ __forceinline__ void skinny_bwd_body(const float* __restrict__ d, const __nv_bfloat16* __restrict__ x, long long x_ps,
                                                int x_ld, int c0, int Cw, int B, int T, const float* __restrict__ W,
                                                int rows_per_block, float* __restrict__ dx, int dx_ld, int dx_accumulate,
                                                float* __restrict__ dW, int cblock, int rblock) {
  const int c = (cblock * blockDim.x + threadIdx.x) * 2;
  if (c >= Cw) return;
  constexpr int PAD = (K - 1) / 2;
  float w[2][NOUT][K], gw[2][NOUT][K];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
      for (int r = 0; r < K; ++r) {
        w[h][o][r] = W[(static_cast<long long>(o) * Cw + c + h) * K + r];
        gw[h][o][r] = 0.f;
      }
  const long long rows = static_cast<long long>(B) * T;
  const long long r0 = static_cast<long long>(rblock) * rows_per_block;
  if (r0 >= rows) return;
  constexpr int U = 8;  // rows in flight per thread: the activation loads of U rows are issued before any of them is used
  for (int i0 = 0; i0 < rows_per_block; i0 += U) {
    uint32_t xh[U], xl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = r0 + i0 + u;
      xh[u] = xl[u] = 0u;
      if (i0 + u < rows_per_block && row < rows) {
        const __nv_bfloat16* xp = x + row * x_ld + c0 + c;
        xh[u] = *reinterpret_cast<const uint32_t*>(xp);
        xl[u] = *reinterpret_cast<const uint32_t*>(xp + x_ps);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long row = r0 + i0 + u;
      if (i0 + u >= rows_per_block || row >= rows) break;
      const int t = static_cast<int>(row % T);
      const float xv[2] = {__uint_as_float(xh[u] << 16) + __uint_as_float(xl[u] << 16),
                           __uint_as_float(xh[u] & 0xffff0000u) + __uint_as_float(xl[u] & 0xffff0000u)};
      float g[2] = {0.f, 0.f};
#pragma unroll
      for (int r = 0; r < K; ++r) {
        const int td = t + PAD - r;
        if (td < 0 || td >= T) continue;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
          const float e = __ldg(d + (row + PAD - r) * NOUT + o);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            g[h] = fmaf(e, w[h][o][r], g[h]);
            gw[h][o][r] = fmaf(e, xv[h], gw[h][o][r]);
          }
        }
      }
      float2* dp = reinterpret_cast<float2*>(dx + row * dx_ld + c0 + c);
      if (dx_accumulate) {
        const float2 old = *dp;
        *dp = make_float2(old.x + g[0], old.y + g[1]);
      } else {
        *dp = make_float2(g[0], g[1]);
      }
    }
  }
  if (dW) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int o = 0; o < NOUT; ++o)
#pragma unroll
        for (int r = 0; r < K; ++r) atomicAdd(dW + (static_cast<long long>(o) * Cw + c + h) * K + r, gw[h][o][r]);
  }
}

template <int NOUT, int K>
__global__ void __launch_bounds__(256) skinny_bwd_kernel(const float* __restrict__ d, const __nv_bfloat16* __restrict__ x,
                                                         long long x_ps, int x_ld, int c0, int Cw, int B, int T,
                                                         const float* __restrict__ W, int rows_per_block,
                                                         float* __restrict__ dx, int dx_ld, int dx_accumulate,
                                                         float* __restrict__ dW) {
  pdl_sync();
  skinny_bwd_body<NOUT, K>(d, x, x_ps, x_ld, c0, Cw, B, T, W, rows_per_block, dx, dx_ld, dx_accumulate, dW, blockIdx.x, blockIdx.y);
}

// ---- all pyramid levels, cls_logits + bbox_pred (+ iou_scores.3) in ONE launch -----------------------------------------------
struct HeadLevels {
  int nlevels, B, F;              // F = tower channels per branch (512); the tower tensor carries [cls | bbox] = 2F channels
  int T[MAX_LEVELS], off[MAX_LEVELS];
  const __nv_bfloat16* tw[MAX_LEVELS];
  long long tw_ps[MAX_LEVELS];
  const __nv_bfloat16* hi[MAX_LEVELS];  // iou_scores hidden activations (F/2 channels); null = skip the IoU projection
  long long hi_ps[MAX_LEVELS];
  float* dtw[MAX_LEVELS];         // backward: gradient w.r.t. the tower tensor, fp32 [B*T][2F]
};

// forward: one warp per location; weights staged tap-major in shared memory once per CTA.  NCH = 2F / 256 chunks of 256
// channels per row (a lane owns 8 consecutive channels of every chunk): compile-time, so that the plane loads of a whole tap
// (2 x NCH 16-byte loads per lane) are in flight together and the cls / bbox branch of a chunk is resolved statically.
__device__ __forceinline__ void unpack8(const uint4& h, const uint4& l, float (&v)[8]) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    v[2 * k] = __uint_as_float(hw[k] << 16) + __uint_as_float(lw[k] << 16);
    v[2 * k + 1] = __uint_as_float(hw[k] & 0xffff0000u) + __uint_as_float(lw[k] & 0xffff0000u);
  }
}
__device__ __forceinline__ float dot8(const float (&v)[8], const float4& w0, const float4& w1) {
  return v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z + v[7] * w1.w;
}
template <int NCH>
__global__ void __launch_bounds__(256) head_proj_fwd_kernel(HeadLevels g, const float* __restrict__ Wc, const float* __restrict__ bc,
                                                            const float* __restrict__ Wb, const float* __restrict__ bb,
                                                            const float* __restrict__ Wi, const float* __restrict__ bi,
                                                            float* __restrict__ cls_raw, float* __restrict__ box_raw,
                                                            float* __restrict__ iou_raw, long long total) {
  pdl_sync();
  extern __shared__ __align__(16) float hsm[];
  constexpr int F = NCH * 128;
  float* wc = hsm;              // [3][F]
  float* wb = hsm + 3 * F;      // [2][3][F]
  float* wi = hsm + 9 * F;      // [F/2]
  // weights tap-major and PERMUTED so that the warp's float4 reads below are lane-contiguous (conflict-free): channel
  // c = chunk*256 + lane*8 + q*4 + e lives at ((chunk*2 + q)*32 + lane)*4 + e.  (The natural order cost 2.3 wavefronts per
  // shared load and made this kernel shared-memory bound, ncu r01.)
  for (int i = threadIdx.x; i < 3 * F; i += 256) {
    const int r = i / F, c = i % F;
    const int pidx = r * F + (((c >> 8) * 2 + ((c & 7) >> 2)) * 32 + ((c & 255) >> 3)) * 4 + (c & 3);
    wc[pidx] = Wc[c * 3 + r];
    wb[pidx] = Wb[c * 3 + r];
    wb[3 * F + pidx] = Wb[(F + c) * 3 + r];
  }
  for (int i = threadIdx.x; i < F / 2; i += 256) wi[i] = Wi[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nwarp = static_cast<long long>(gridDim.x) * 8;
  for (long long i = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); i < total; i += nwarp) {
    int lvl = 0;
#pragma unroll
    for (int l = 1; l < MAX_LEVELS; ++l)
      if (l < g.nlevels && i >= static_cast<long long>(g.B) * g.off[l]) lvl = l;
    const int j = static_cast<int>(i - static_cast<long long>(g.B) * g.off[lvl]);  // row inside the level (< B * T)
    const int T = g.T[lvl];
    const int t = j % T;
    const long long ps = g.tw_ps[lvl];
    float a_cls = 0.f, a_b0 = 0.f, a_b1 = 0.f, a_iou = 0.f;
    const bool has_iou = g.hi[lvl] != nullptr;
    uint4 ih = make_uint4(0u, 0u, 0u, 0u), il = ih;
    const bool iou_lane = has_iou && lane * 8 < F / 2;  // F/2 <= 256 channels: one chunk
    if (iou_lane) {
      const __nv_bfloat16* row = g.hi[lvl] + static_cast<long long>(j) * (F / 2) + lane * 8;
      ih = __ldg(reinterpret_cast<const uint4*>(row));
      il = __ldg(reinterpret_cast<const uint4*>(row + g.hi_ps[lvl]));
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ts = t + r - 1;
      if (ts < 0 || ts >= T) continue;
      const __nv_bfloat16* row = g.tw[lvl] + (static_cast<long long>(j) + r - 1) * (2 * F) + lane * 8;
      uint4 h[NCH], l[NCH];
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        h[k] = __ldg(reinterpret_cast<const uint4*>(row + k * 256));
        l[k] = __ldg(reinterpret_cast<const uint4*>(row + k * 256 + ps));
      }
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        float v[8];
        unpack8(h[k], l[k], v);
        constexpr int HALF = NCH / 2;               // chunks [0, HALF) = cls tower, [HALF, NCH) = bbox tower
        const int kk = k < HALF ? k : k - HALF;     // chunk inside its branch
        const int p0 = (kk * 2) * 128 + lane * 4;   // permuted offsets of the lane's two float4s
        const int p1 = p0 + 128;
        if (k < HALF) {
          a_cls += dot8(v, *reinterpret_cast<const float4*>(wc + r * F + p0), *reinterpret_cast<const float4*>(wc + r * F + p1));
        } else {
          a_b0 += dot8(v, *reinterpret_cast<const float4*>(wb + r * F + p0), *reinterpret_cast<const float4*>(wb + r * F + p1));
          a_b1 += dot8(v, *reinterpret_cast<const float4*>(wb + (3 + r) * F + p0), *reinterpret_cast<const float4*>(wb + (3 + r) * F + p1));
        }
      }
    }
    if (iou_lane) {
      float v[8];
      unpack8(ih, il, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) a_iou = fmaf(v[k], wi[lane * 8 + k], a_iou);
    }
    a_cls = warp_sum(a_cls);
    a_b0 = warp_sum(a_b0);
    a_b1 = warp_sum(a_b1);
    a_iou = warp_sum(a_iou);
    if (lane == 0) {
      cls_raw[i] = a_cls + bc[0];
      box_raw[2 * i] = a_b0 + bb[0];
      box_raw[2 * i + 1] = a_b1 + bb[1];
      if (has_iou) iou_raw[i] = a_iou + bi[0];
    }
  }
}

// backward of cls_logits (blockIdx.x = 0: channels [0,F)) and bbox_pred (blockIdx.x = 1: channels [F,2F)) on every level
// (blockIdx.z); blockIdx.y = chunk of rows
// backward of the N = 1 / 2 projections on all levels in one launch: d_tower[b,t,c] = sum_{o,r} d[b,t+1-r,o] W[o][c][r] and
// dW[o][c][r] += sum_{b,t} tower[b,t,c] d[b,t+1-r,o].  Memory-bound (planes read once, fp32 gradient written once: 8 bytes
// per (location, channel)).  blockIdx.y = branch (0: cls_logits on channels [0,F), 1: bbox_pred on [F,2F)); a CTA walks row
// blocks (HB_ROWS consecutive locations of ONE sample, so no per-row index arithmetic) round-robin over all levels with the
// weights and the dW partial sums in registers; a thread owns 4 channels (8-byte plane loads, 16-byte stores) of every
// rsub-th row; the upstream gradients of a row block (+1 halo row each side, zero outside the sample) sit in shared memory.
// dW leaves the CTA once: partials of the row sub-groups are folded through shared memory, one atomic per weight per CTA.
constexpr int HB_ROWS = 32;
template <int NOUT, int HB_U>
__device__ __forceinline__ void head_bwd_branch(const HeadLevels& g, const float* __restrict__ dall, const float* __restrict__ W,
                                                float* __restrict__ dW, int c0, int nblk_total, float* dsm, float* wred) {
  const int F = g.F, nq = F >> 2;
  const int cq = threadIdx.x % nq, rs = threadIdx.x / nq, rsub = blockDim.x / nq;
  const int c = cq * 4;
  float w[4][NOUT][3], gw[4][NOUT][3];
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        w[h][o][r] = __ldg(W + (static_cast<long long>(o) * F + c + h) * 3 + r);
        gw[h][o][r] = 0.f;
      }
  for (int blk = blockIdx.x; blk < nblk_total; blk += gridDim.x) {
    int lvl = 0, rem = blk;
#pragma unroll
    for (int l = 0; l < MAX_LEVELS - 1; ++l) {
      const int nb = g.B * ((g.T[l] + HB_ROWS - 1) / HB_ROWS);
      if (lvl == l && l + 1 < g.nlevels && rem >= nb) {
        rem -= nb;
        lvl = l + 1;
      }
    }
    const int T = g.T[lvl], bps = (T + HB_ROWS - 1) / HB_ROWS;
    const int b = rem / bps, t0 = (rem - b * bps) * HB_ROWS;
    const int nrow = min(HB_ROWS, T - t0);
    const long long row0 = static_cast<long long>(b) * T + t0;
    const float* __restrict__ dl = dall + static_cast<long long>(g.B) * g.off[lvl] * NOUT;
    __syncthreads();  // the previous row block's readers are done with dsm
    for (int i = threadIdx.x; i < (HB_ROWS + 2) * NOUT; i += blockDim.x) {
      const int j = i / NOUT, o = i % NOUT;
      const int t = t0 - 1 + j;  // dsm row j = time t0 - 1 + j
      dsm[i] = (t >= 0 && t < T) ? __ldg(dl + (static_cast<long long>(b) * T + t) * NOUT + o) : 0.f;
    }
    __syncthreads();
    const __nv_bfloat16* __restrict__ xb = g.tw[lvl] + row0 * (2 * F) + c0 + c;
    const long long ps = g.tw_ps[lvl];
    float* __restrict__ dxb = g.dtw[lvl] + row0 * (2 * F) + c0 + c;
    for (int i0 = rs; i0 < nrow; i0 += rsub * HB_U) {
      uint2 xh[HB_U], xl[HB_U];
#pragma unroll
      for (int u = 0; u < HB_U; ++u) {  // all plane loads of HB_U rows are issued before any of them is used
        const int i = i0 + u * rsub;
        xh[u] = xl[u] = make_uint2(0u, 0u);
        if (i < nrow) {
          const __nv_bfloat16* xp = xb + static_cast<long long>(i) * (2 * F);
          xh[u] = __ldg(reinterpret_cast<const uint2*>(xp));
          xl[u] = __ldg(reinterpret_cast<const uint2*>(xp + ps));
        }
      }
#pragma unroll
      for (int u = 0; u < HB_U; ++u) {
        const int i = i0 + u * rsub;
        if (i < nrow) {
          const float xv[4] = {__uint_as_float(xh[u].x << 16) + __uint_as_float(xl[u].x << 16),
                               __uint_as_float(xh[u].x & 0xffff0000u) + __uint_as_float(xl[u].x & 0xffff0000u),
                               __uint_as_float(xh[u].y << 16) + __uint_as_float(xl[u].y << 16),
                               __uint_as_float(xh[u].y & 0xffff0000u) + __uint_as_float(xl[u].y & 0xffff0000u)};
          float ga[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
              const float e = dsm[(i + 2 - r) * NOUT + o];  // d[t + 1 - r]
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                ga[h] = fmaf(e, w[h][o][r], ga[h]);
                gw[h][o][r] = fmaf(e, xv[h], gw[h][o][r]);
              }
            }
          *reinterpret_cast<float4*>(dxb + static_cast<long long>(i) * (2 * F)) = make_float4(ga[0], ga[1], ga[2], ga[3]);
        }
      }
    }
  }
  if (!dW) return;
  const int nw = F * NOUT * 3;
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
#pragma unroll
      for (int r = 0; r < 3; ++r) wred[rs * nw + (o * F + c + h) * 3 + r] = gw[h][o][r];
  __syncthreads();
  for (int i = threadIdx.x; i < nw; i += blockDim.x) {
    float v = wred[i];
    for (int q = 1; q < rsub; ++q) v += wred[q * nw + i];
    atomicAdd(dW + i, v);
  }
}
__global__ void __launch_bounds__(256, 3) head_proj_bwd_kernel(HeadLevels g, const float* __restrict__ dcls,
                                                            const float* __restrict__ dbox, const float* __restrict__ Wc,
                                                            const float* __restrict__ Wb, int nblk_total,
                                                            float* __restrict__ dWc, float* __restrict__ dWb) {
  pdl_sync();
  extern __shared__ __align__(16) float hb_sm[];  // [(HB_ROWS + 2) * 2] upstream gradients, then [rsub][F * NOUT * 3] dW partials
  float* dsm = hb_sm;
  float* wred = hb_sm + (HB_ROWS + 2) * 2 + 4;
  if (blockIdx.y == 0) head_bwd_branch<1, 4>(g, dcls, Wc, dWc, 0, nblk_total, dsm, wred);
  else head_bwd_branch<2, 2>(g, dbox, Wb, dWb, g.F, nblk_total, dsm, wred);
}

// ---- loss ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

struct LocInfo {
  int lvl, b, t;
  float loc;
};
__device__ __forceinline__ LocInfo locate(const LevelGeom& g, long long i) {
  LocInfo r;
  r.lvl = 0;
#pragma unroll
  for (int l = 1; l < MAX_LEVELS; ++l)
    if (l < g.nlevels && i >= static_cast<long long>(g.B) * g.off[l]) r.lvl = l;
  const long long j = i - static_cast<long long>(g.B) * g.off[r.lvl];
  r.b = static_cast<int>(j / g.T[r.lvl]);
  r.t = static_cast<int>(j % g.T[r.lvl]);
  r.loc = g.stride[r.lvl] * r.t + g.stride[r.lvl] * 0.5f;  // model/fcos.py:204-211
  return r;
}

struct IouBranch {
  float tiou, dt_dpl, dt_dpr;  // tIoU of the decoded prediction vs GT and its derivative w.r.t. the (left,right) distances
};
// model/loss.py:176-187 + segment_tiou (241-256).  `first` = first location of the sample in level-major order, the only
// one the reference clamps to [0,1] (the loss.py:180-181 quirk).
__device__ __forceinline__ IouBranch iou_branch(float loc, float pl, float pr, float gs, float ge, bool first) {
  float ps = (loc - pl) * (1.f / 32.f), pe = (loc + pr) * (1.f / 32.f);
  float cs = 1.f, ce = 1.f;  // clamp pass-through
  if (first) {
    if (ps < 0.f || ps > 1.f) cs = 0.f;
    if (pe < 0.f || pe > 1.f) ce = 0.f;
    ps = fminf(fmaxf(ps, 0.f), 1.f);
    pe = fminf(fmaxf(pe, 0.f), 1.f);
  }
  const float iraw = fminf(pe, ge) - fmaxf(ps, gs);
  const float uraw = fmaxf(pe, ge) - fminf(ps, gs);
  const float inter = fmaxf(iraw, 0.f), uni = fmaxf(uraw, 0.f);
  const float den = uni + 1e-6f;
  IouBranch r;
  r.tiou = inter / den;
  const float ia = iraw > 0.f ? 1.f : 0.f, ua = uraw > 0.f ? 1.f : 0.f;
  const float di_dpe = ia * (pe < ge ? 1.f : 0.f), di_dps = -ia * (ps > gs ? 1.f : 0.f);
  const float du_dpe = ua * (pe > ge ? 1.f : 0.f), du_dps = -ua * (ps < gs ? 1.f : 0.f);
  const float dt_dpe = (di_dpe * den - inter * du_dpe) / (den * den);
  const float dt_dps = (di_dps * den - inter * du_dps) / (den * den);
  r.dt_dpl = dt_dps * cs * (-1.f / 32.f);
  r.dt_dpr = dt_dpe * ce * (1.f / 32.f);
  return r;
}

// acc: [0] focal sum, [1] n_pos, [2] sum of -log IoU over positives, [3] smooth-L1 sum, [4] IoU-branch count
__global__ void __launch_bounds__(256) fcos_loss_fwd_kernel(LevelGeom g, const float* __restrict__ cls_raw,
                                                            const float* __restrict__ box_raw, const float* __restrict__ iou_raw,
                                                            const float* __restrict__ scales, const float* __restrict__ gt,
                                                            float gamma, float alpha, int iou_branch_on,
                                                            float* __restrict__ bbox_out, double* __restrict__ acc) {
  pdl_sync();
  __shared__ double red[5][8];
  const long long total = static_cast<long long>(g.B) * g.P;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  double v[5] = {0, 0, 0, 0, 0};
  if (i < total) {
    const LocInfo L = locate(g, i);
    const float s = scales[L.lvl];
    const float pl = expf(box_raw[2 * i] * s), pr = expf(box_raw[2 * i + 1] * s);
    bbox_out[2 * i] = pl;
    bbox_out[2 * i + 1] = pr;
    const float gs = gt[2 * L.b], ge = gt[2 * L.b + 1];
    const float tl = L.loc - gs * 32.f, tr = ge * 32.f - L.loc;
    const float m = fmaxf(tl, tr);
    const bool pos = (fminf(tl, tr) > 0.f) && (m >= g.lo[L.lvl]) && (m <= g.hi[L.lvl]);
    const float x = cls_raw[i];
    const float p = 1.f / (1.f + expf(-x));
    if (pos) {
      v[0] = -alpha * powf(1.f - p, gamma) * log_sigmoid(x);
      v[1] = 1.0;
      const float inter = fminf(pl, tl) + fminf(pr, tr);
      const float uni = tl + tr + pl + pr - inter;
      v[2] = -logf((inter + 1e-8f) / (uni + 1e-8f));
    } else {
      v[0] = -(1.f - alpha) * powf(p, gamma) * log_sigmoid(-x);
    }
    if (iou_branch_on) {
      const IouBranch ib = iou_branch(L.loc, pl, pr, gs, ge, L.lvl == 0 && L.t == 0);
      if (ib.tiou > 0.9f) {
        const float d = 1.f / (1.f + expf(-iou_raw[i])) - ib.tiou;
        const float ad = fabsf(d);
        v[3] = ad < 1.f ? 0.5f * d * d : ad - 0.5f;
        v[4] = 1.0;
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double s = warp_sum(v[k]);
    if (lane == 0) red[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
    if (s != 0.0) atomicAdd(acc + threadIdx.x, s);
  }
}

// losses: [0] loss_cls, [1] loss_reg, [2] loss_iou, [3] n_pos, [4] IoU-branch count
__global__ void fcos_loss_finalize_kernel(const double* __restrict__ acc, int B, float* __restrict__ losses) {
  pdl_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double npos = acc[1], cnt = acc[4];
  losses[0] = static_cast<float>(acc[0] / (npos + B));          // model/loss.py:210-213
  losses[1] = npos > 0 ? static_cast<float>(acc[2] / npos) : 0.f;  // loss.py:215-231
  losses[2] = cnt > 0 ? static_cast<float>(acc[3] / cnt) : 0.f;    // loss.py:189-197
  losses[3] = static_cast<float>(npos);
  losses[4] = static_cast<float>(cnt);
}

// upstream: [0] d loss_cls, [1] d loss_reg, [2] d loss_iou.  Outputs gradients w.r.t. the raw conv outputs.
// pgrad: [0] d cls_logits.bias, [1..2] d bbox_pred.bias, [3] d iou_scores.3.bias, [4..4+nlevels) d scales
__global__ void __launch_bounds__(256) fcos_loss_bwd_kernel(LevelGeom g, const float* __restrict__ cls_raw,
                                                            const float* __restrict__ box_raw, const float* __restrict__ iou_raw,
                                                            const float* __restrict__ scales, const float* __restrict__ gt,
                                                            float gamma, float alpha, int iou_branch_on,
                                                            const double* __restrict__ acc, const float* __restrict__ upstream,
                                                            float* __restrict__ dcls, float* __restrict__ dbox,
                                                            float* __restrict__ diou, float* __restrict__ pgrad) {
  pdl_sync();
  __shared__ float red[4 + MAX_LEVELS][8];
  const long long total = static_cast<long long>(g.B) * g.P;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  float pg[4 + MAX_LEVELS];
#pragma unroll
  for (int k = 0; k < 4 + MAX_LEVELS; ++k) pg[k] = 0.f;
  if (i < total) {
    const LocInfo L = locate(g, i);
    const float npos = static_cast<float>(acc[1]), cnt = static_cast<float>(acc[4]);
    const float g_cls = upstream[0] / (npos + g.B);
    const float g_reg = npos > 0.f ? upstream[1] / npos : 0.f;
    const float g_iou = (iou_branch_on && cnt > 0.f) ? upstream[2] / cnt : 0.f;
    const float s = scales[L.lvl];
    const float r0 = box_raw[2 * i], r1 = box_raw[2 * i + 1];
    const float pl = expf(r0 * s), pr = expf(r1 * s);
    const float gs = gt[2 * L.b], ge = gt[2 * L.b + 1];
    const float tl = L.loc - gs * 32.f, tr = ge * 32.f - L.loc;
    const float m = fmaxf(tl, tr);
    const bool pos = (fminf(tl, tr) > 0.f) && (m >= g.lo[L.lvl]) && (m <= g.hi[L.lvl]);
    const float x = cls_raw[i];
    const float p = 1.f / (1.f + expf(-x));
    float dx, dpl = 0.f, dpr = 0.f, du = 0.f;
    if (pos) {
      dx = -alpha * powf(1.f - p, gamma) * ((1.f - p) - gamma * p * log_sigmoid(x));
      const float il = pl < tl ? 1.f : 0.f, ir = pr < tr ? 1.f : 0.f;
      const float inter = fminf(pl, tl) + fminf(pr, tr);
      const float uni = tl + tr + pl + pr - inter;
      dpl = g_reg * (-il / (inter + 1e-8f) + (1.f - il) / (uni + 1e-8f));
      dpr = g_reg * (-ir / (inter + 1e-8f) + (1.f - ir) / (uni + 1e-8f));
    } else {
      dx = -(1.f - alpha) * powf(p, gamma) * (gamma * (1.f - p) * log_sigmoid(-x) - p);
    }
    dx *= g_cls;
    if (g_iou != 0.f) {
      const IouBranch ib = iou_branch(L.loc, pl, pr, gs, ge, L.lvl == 0 && L.t == 0);
      if (ib.tiou > 0.9f) {
        const float sg = 1.f / (1.f + expf(-iou_raw[i]));
        const float d = sg - ib.tiou;
        const float dl = (fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f)) * g_iou;
        du = dl * sg * (1.f - sg);
        dpl -= dl * ib.dt_dpl;
        dpr -= dl * ib.dt_dpr;
      }
    }
    const float d0 = dpl * pl * s, d1 = dpr * pr * s;  // through exp(scale * raw), model/fcos.py:98-100
    dcls[i] = dx;
    dbox[2 * i] = d0;
    dbox[2 * i + 1] = d1;
    diou[i] = du;
    pg[0] = dx;
    pg[1] = d0;
    pg[2] = d1;
    pg[3] = du;
    const float ds = dpl * pl * r0 + dpr * pr * r1;
#pragma unroll
    for (int l = 0; l < MAX_LEVELS; ++l)
      if (l == L.lvl) pg[4 + l] = ds;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4 + MAX_LEVELS; ++k) {
    const float sm = warp_sum(pg[k]);
    if (lane == 0) red[k][warp] = sm;
  }
  __syncthreads();
  if (threadIdx.x < 4 + g.nlevels) {
    float sm = 0;
    for (int w = 0; w < 8; ++w) sm += red[threadIdx.x][w];
    if (sm != 0.f) atomicAdd(pgrad + threadIdx.x, sm);
  }
}

// ---- eval-mode post-processing (model/inference.py:49-136 per level, 167-215 concatenation): one CTA per (level, sample) ----
// sigmoid -> threshold on the CLASS score (before the IoU product, inference.py:71) -> keep the top_n by score ->
// decode (loc -/+ reg)/32, clamp to [0,1] -> sqrt score.  Fixed-shape outputs [B][nlevels][top_n], kept candidates in
// location order, so the host needs ONE small device->host copy and no per-image synchronisation.
constexpr int POST_MAX_T = 2048;
__global__ void __launch_bounds__(256) postprocess_kernel(LevelGeom g, const float* __restrict__ cls_raw,
                                                          const float* __restrict__ bbox, const float* __restrict__ iou_raw,
                                                          float thr, int top_n, int use_iou, float* __restrict__ out_det,
                                                          float* __restrict__ out_score, float* __restrict__ out_loc,
                                                          int* __restrict__ out_count) {
  pdl_sync();
  __shared__ float sc[POST_MAX_T];
  __shared__ unsigned char keep[POST_MAX_T];
  __shared__ int ncand_s;
  const int lvl = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int T = g.T[lvl];
  const long long base = static_cast<long long>(g.B) * g.off[lvl] + static_cast<long long>(b) * T;
  if (tid == 0) ncand_s = 0;
  __syncthreads();
  int mine = 0;
  for (int t = tid; t < T; t += 256) {
    const float c = 1.f / (1.f + expf(-cls_raw[base + t]));
    const bool cand = c > thr;
    float s = c;
    if (use_iou) s = c * (1.f / (1.f + expf(-iou_raw[base + t])));
    sc[t] = cand ? s : -1.f;
    mine += cand ? 1 : 0;
  }
  if (mine) atomicAdd(&ncand_s, mine);
  __syncthreads();
  const int k = min(ncand_s, top_n);
  for (int t = tid; t < T; t += 256) {
    const float s = sc[t];
    int rank = 0;
    if (s >= 0.f)
      for (int u = 0; u < T; ++u) rank += (sc[u] > s || (sc[u] == s && u < t)) ? 1 : 0;
    keep[t] = (s >= 0.f && rank < k) ? 1 : 0;
  }
  __syncthreads();
  const long long o = (static_cast<long long>(b) * g.nlevels + lvl) * top_n;
  for (int t = tid; t < T; t += 256) {
    if (!keep[t]) continue;
    int pos = 0;
    for (int u = 0; u < t; ++u) pos += keep[u];
    const float loc = g.stride[lvl] * t + g.stride[lvl] * 0.5f;
    const float l = bbox[2 * (base + t)], r = bbox[2 * (base + t) + 1];
    out_det[2 * (o + pos)] = fminf(fmaxf((loc - l) / 32.f, 0.f), 1.f);
    out_det[2 * (o + pos) + 1] = fminf(fmaxf((loc + r) / 32.f, 0.f), 1.f);
    out_score[o + pos] = sqrtf(sc[t]);
    out_loc[o + pos] = loc / 32.f;
  }
  if (tid == 0) out_count[b * g.nlevels + lvl] = k;
}

}  // namespace drn

using namespace drn;
#define ST(s) static_cast<cudaStream_t>(s)

static int make_geom(LevelGeom* g, int nlevels, int B, const int* T, const float* strides) {
  if (nlevels < 1 || nlevels > 3) return fail(DRN_EINVAL, "fcos: the reference defines size bands for exactly 3 levels (got %d)", nlevels);
  const float lo[3] = {-1.f, 5.6f, 11.f}, hi[3] = {6.f, 11.f, 100000000.f};  // model/loss.py:47-51
  g->nlevels = nlevels;
  g->B = B;
  int off = 0;
  for (int l = 0; l < nlevels; ++l) {
    g->T[l] = T[l];
    g->off[l] = off;
    off += T[l];
    g->stride[l] = strides[l];
    g->lo[l] = lo[l];
    g->hi[l] = hi[l];
  }
  g->P = off;
  return 0;
}

extern "C" int drn_skinny_conv_fwd(const void* x, int64_t x_plane_stride, int x_ld, int c0, int Cw, int B, int T, int nout, int k,
                                   const float* W, const float* bias, float* out, void* stream) {
  if (Cw % 8 || x_ld % 8 || c0 % 8) return fail(DRN_EINVAL, "drn_skinny_conv_fwd: alignment");
  const long long warps = static_cast<long long>(B) * T;
  const unsigned grid = static_cast<unsigned>((warps * 32 + 255) / 256);
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  if (nout == 1 && k == 3) launch_k(skinny_fwd_kernel<1, 3>, grid, 256, 0, ST(stream), xp, x_plane_stride, x_ld, c0, Cw, B, T, W, bias, out);
  else if (nout == 2 && k == 3) launch_k(skinny_fwd_kernel<2, 3>, grid, 256, 0, ST(stream), xp, x_plane_stride, x_ld, c0, Cw, B, T, W, bias, out);
  else if (nout == 1 && k == 1) launch_k(skinny_fwd_kernel<1, 1>, grid, 256, 0, ST(stream), xp, x_plane_stride, x_ld, c0, Cw, B, T, W, bias, out);
  else return fail(DRN_EINVAL, "drn_skinny_conv_fwd: unsupported (nout=%d,k=%d)", nout, k);
  return check_launch("skinny_conv_fwd");
}

extern "C" int drn_skinny_conv_bwd(const float* d, const void* x, int64_t x_plane_stride, int x_ld, int c0, int Cw, int B, int T,
                                   int nout, int k, const float* W, float* dx, int dx_ld, int dx_accumulate, float* dW,
                                   void* stream) {
  if (Cw % 2 || c0 % 2 || dx_ld % 2 || x_ld % 2) return fail(DRN_EINVAL, "drn_skinny_conv_bwd: alignment");
  const long long rows_total = static_cast<long long>(B) * T;
  int rpb = 16;  // ~2 waves of CTAs: fewer, longer CTAs cut the dW atomics (one set per CTA) without starving the SMs
  while (rpb < 64 && rows_total / rpb > 2 * 148) rpb *= 2;
  dim3 grid(ceil_div(Cw, 512), static_cast<unsigned>((rows_total + rpb - 1) / rpb));
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  if (nout == 1 && k == 3) launch_k(skinny_bwd_kernel<1, 3>, grid, 256, 0, ST(stream), d, xp, x_plane_stride, x_ld, c0, Cw, B, T, W, rpb, dx, dx_ld, dx_accumulate, dW);
  else if (nout == 2 && k == 3) launch_k(skinny_bwd_kernel<2, 3>, grid, 256, 0, ST(stream), d, xp, x_plane_stride, x_ld, c0, Cw, B, T, W, rpb, dx, dx_ld, dx_accumulate, dW);
  else if (nout == 1 && k == 1) launch_k(skinny_bwd_kernel<1, 1>, grid, 256, 0, ST(stream), d, xp, x_plane_stride, x_ld, c0, Cw, B, T, W, rpb, dx, dx_ld, dx_accumulate, dW);
  else return fail(DRN_EINVAL, "drn_skinny_conv_bwd: unsupported (nout=%d,k=%d)", nout, k);
  return check_launch("skinny_conv_bwd");
}

extern "C" int drn_fcos_loss_fwd(int nlevels, int B, const int* T, const float* strides, const float* cls_raw, const float* box_raw,
                                 const float* iou_raw, const float* scales, const float* gt, float gamma, float alpha,
                                 int iou_branch_on, float* bbox_out, double* acc, float* losses, void* stream) {
  LevelGeom g;
  int rc = make_geom(&g, nlevels, B, T, strides);
  if (rc) return rc;
  cudaError_t e = cudaMemsetAsync(acc, 0, 8 * sizeof(double), ST(stream));
  if (e != cudaSuccess) return fail(static_cast<int>(e), "memset: %s", cudaGetErrorString(e));
  const long long total = static_cast<long long>(B) * g.P;
  launch_k(fcos_loss_fwd_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, ST(stream), g, cls_raw, box_raw, iou_raw, scales, gt, gamma,
                                                                                         alpha, iou_branch_on, bbox_out, acc);
  launch_k(fcos_loss_finalize_kernel, 1, 32, 0, ST(stream), acc, B, losses);
  return check_launch("fcos_loss_fwd");
}

extern "C" int drn_fcos_loss_bwd(int nlevels, int B, const int* T, const float* strides, const float* cls_raw, const float* box_raw,
                                 const float* iou_raw, const float* scales, const float* gt, float gamma, float alpha,
                                 int iou_branch_on, const double* acc, const float* upstream, float* dcls, float* dbox, float* diou,
                                 float* pgrad, void* stream) {
  LevelGeom g;
  int rc = make_geom(&g, nlevels, B, T, strides);
  if (rc) return rc;
  const long long total = static_cast<long long>(B) * g.P;
  launch_k(fcos_loss_bwd_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, ST(stream), 
      g, cls_raw, box_raw, iou_raw, scales, gt, gamma, alpha, iou_branch_on, acc, upstream, dcls, dbox, diou, pgrad);
  return check_launch("fcos_loss_bwd");
}

extern "C" int drn_postprocess(int nlevels, int B, const int* T, const float* strides, const float* cls_raw, const float* bbox,
                               const float* iou_raw, float thr, int top_n, int use_iou, float* out_det, float* out_score,
                               float* out_loc, int* out_count, void* stream) {
  LevelGeom g;
  int rc = make_geom(&g, nlevels, B, T, strides);
  if (rc) return rc;
  if (top_n < 1) return fail(DRN_EINVAL, "drn_postprocess: top_n must be positive");
  for (int l = 0; l < nlevels; ++l)
    if (T[l] > POST_MAX_T) return fail(DRN_EINVAL, "drn_postprocess: at most %d locations per level (got %d)", POST_MAX_T, T[l]);
  launch_k(postprocess_kernel, dim3(nlevels, B), 256, 0, ST(stream), g, cls_raw, bbox, iou_raw, thr, top_n, use_iou, out_det, out_score,
                                                             out_loc, out_count);
  return check_launch("postprocess");
}

static int make_head_levels(HeadLevels* g, const drn_head_levels_t* h, const char* who) {
  if (!h || h->nlevels < 1 || h->nlevels > 3) return fail(DRN_EINVAL, "%s: 1..3 levels", who);
  if (h->F < 16 || h->F % 16 || h->B < 1) return fail(DRN_EINVAL, "%s: tower channels must be a multiple of 16 (F=%d)", who, h->F);
  g->nlevels = h->nlevels;
  g->B = h->B;
  g->F = h->F;
  int off = 0;
  for (int l = 0; l < h->nlevels; ++l) {
    g->T[l] = h->T[l];
    g->off[l] = off;
    off += h->T[l];
    g->tw[l] = static_cast<const __nv_bfloat16*>(h->tower[l]);
    g->tw_ps[l] = h->tower_plane_stride[l];
    g->hi[l] = static_cast<const __nv_bfloat16*>(h->iou_hidden[l]);
    g->hi_ps[l] = h->iou_hidden_plane_stride[l];
    g->dtw[l] = h->d_tower[l];
    if (!g->tw[l] || g->tw_ps[l] % 8) return fail(DRN_EINVAL, "%s: level %d tower planes missing / misaligned", who, l);
  }
  return off;
}

extern "C" int drn_head_proj_fwd(const drn_head_levels_t* h, const float* Wc, const float* bc, const float* Wb, const float* bb,
                                 const float* Wi, const float* bi, float* cls_raw, float* box_raw, float* iou_raw, void* stream) {
  HeadLevels g{};
  const int P = make_head_levels(&g, h, "drn_head_proj_fwd");
  if (P < 0) return P;
  const long long total = static_cast<long long>(g.B) * P;
  const size_t smem = (9 * g.F + g.F / 2) * sizeof(float);
  if (g.F != 256 && g.F != 512) return fail(DRN_EINVAL, "drn_head_proj_fwd: tower channels per branch must be 256 or 512 (F=%d)", g.F);
  long long ctas = (total + 7) / 8;
  if (ctas > 148 * 3) ctas = 148 * 3;  // 70 registers x 256 threads: three CTAs per SM, one wave (grid-stride loop over locations)
  if (g.F == 512)
    launch_k(head_proj_fwd_kernel<4>, static_cast<unsigned>(ctas), 256, smem, ST(stream), g, Wc, bc, Wb, bb, Wi, bi, cls_raw, box_raw, iou_raw, total);
  else
    launch_k(head_proj_fwd_kernel<2>, static_cast<unsigned>(ctas), 256, smem, ST(stream), g, Wc, bc, Wb, bb, Wi, bi, cls_raw, box_raw, iou_raw, total);
  return check_launch("head_proj_fwd");
}

extern "C" int drn_head_proj_bwd(const drn_head_levels_t* h, const float* dcls, const float* dbox, const float* Wc, const float* Wb,
                                 float* dWc, float* dWb, void* stream) {
  HeadLevels g{};
  const int P = make_head_levels(&g, h, "drn_head_proj_bwd");
  if (P < 0) return P;
  if (g.F > 512) return fail(DRN_EINVAL, "drn_head_proj_bwd: at most 512 tower channels per branch (F=%d)", g.F);
  for (int l = 0; l < g.nlevels; ++l)
    if (!g.dtw[l]) return fail(DRN_EINVAL, "drn_head_proj_bwd: d_tower[%d] missing", l);
  if (g.F != 256 && g.F != 512) return fail(DRN_EINVAL, "drn_head_proj_bwd: tower channels per branch must be 256 or 512 (F=%d)", g.F);
  long long nblk = 0;
  for (int l = 0; l < g.nlevels; ++l) nblk += static_cast<long long>(g.B) * ((g.T[l] + HB_ROWS - 1) / HB_ROWS);
  // 3 CTAs per SM over the two branches; whole rounds of row blocks per CTA (448 blocks at B = 32, T = 256 -> 150 CTAs x 3 rounds)
  const long long gmax = 148 * 3 / 2;
  const long long rounds = (nblk + gmax - 1) / gmax;
  const long long gx = (nblk + rounds - 1) / rounds;
  const int rsub = 256 / (g.F / 4);
  const size_t smem = ((HB_ROWS + 2) * 2 + 4 + static_cast<size_t>(rsub) * g.F * 2 * 3) * sizeof(float);
  dim3 grid(static_cast<unsigned>(gx), 2, 1);
  launch_k(head_proj_bwd_kernel, grid, 256, smem, ST(stream), g, dcls, dbox, Wc, Wb, static_cast<int>(nblk), dWc, dWb);
  return check_launch("head_proj_bwd");
}
