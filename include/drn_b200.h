/* drn_b200.h -- C ABI of libdrn_sm100.so: the B200 (sm_100a) kernels behind DRN's dense-regression hot path.
 *
 * Caller = the Python drop-in `model/` package through ctypes (drn_b200/lib.py).  Conventions
 * (SURVEY.md section 8b):
 *   - every entry point returns 0 on success, a positive cudaError_t, or a negative DRN_E* code; the
 *     text of the last error of the calling thread is drn_last_error().  Nothing throws / exits.
 *   - all buffers are device memory owned by the caller (torch caching allocator); the library
 *     allocates nothing persistent and never synchronises the device.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it.
 *   - activations are channels-last: fp32 [B, T, C] row-major, or "split planes": two bf16 tensors
 *     hi = bf16(x), lo = bf16(x - hi) stored as [2][B][T][C] (plane stride given in elements).  The
 *     tensor-core kernels contract hi/lo planes with 3 BF16 MMAs per product (fp32 TMEM accumulate),
 *     which reproduces fp32 results to ~2^-16 (the 1e-3 parity contract cannot be met by one
 *     BF16/TF32 pass -- SURVEY.md section 7 hard part 1).
 *
 * Each function cites the reference code (paths relative to the DRN repository) it replaces.
 */
#ifndef DRN_B200_H
#define DRN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRN_VERSION 100

#define DRN_EINVAL (-1)   /* bad argument / unsupported shape */
#define DRN_EARCH (-2)    /* device is not compute capability 10.x */
#define DRN_EDRIVER (-3)  /* driver entry point (cuTensorMapEncodeTiled) unavailable */

int drn_version(void);
const char* drn_last_error(void);
/* 0 iff the current device is a CC 10.x part (B200); DRN_EARCH otherwise. */
int drn_device_check(void);
int drn_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * Split-plane tensor view: bf16 [2 planes][B][T][P][C]; element (plane,b,t,p,c) lives at
 *   ptr + plane*plane_stride + ((b*T + t)*P + p)*C + c.
 * P is a "parity" split of the time axis used by the stride-2 temporal convs (P=2: time index
 * 2*t+p); P=1 otherwise.  C must be a multiple of 8, ptr 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  void* ptr;
  int64_t plane_stride; /* elements between the hi and the lo plane */
  int32_t B, T, P, C;
} drn_planes_t;

#define DRN_MAX_TAPS 4

/* ------------------------------------------------------------------------------------------------
 * drn_gemm: the one tensor-core contraction behind every dense layer of the path, forward and backward.
 *
 *   form DRN_GEMM_ROWS (temporal conv / linear forward and data-gradient):
 *     D[(b,t), n] = sum_taps sum_k A[b, t + shift_tap, par_tap, a_c0 + k] * W[wtap_tap][..]
 *       b_mn = 0: W element = Bm[b=wtap][t=n][c=b_c0+k]   (weights stored [tap][N][K]: conv/linear forward)
 *       b_mn = 1: W element = Bm[b=wtap][t=k][c=n]        (weights stored [tap][K][N]: data gradient)
 *     replaces nn.Linear / nn.Conv1d forward+dgrad at model/main_model.py:59 (prop_fc),
 *     model/basic_blocks.py:9-31 (backbone + FPN convs), model/fcos.py:31-69 (head towers, mix_fc, iou_scores).
 *     Rows outside [0,T) read as zero (= Conv1d zero padding) through TMA out-of-bounds fill.
 *   form DRN_GEMM_WGRAD (weight gradient):
 *     D[m, n] = sum_{b,t} A[b, t, 0, m] * Bm[b, t + shift_0, par_0, n]      (A = dY planes, Bm = X planes)
 *
 * Epilogue (per output element v = acc): v += bias[n]; out2[row,n] = v (optional);
 *   v *= rowscale[b*rowscale_ld + n] (optional); out[row, out_col0+n] (=, +=, atomic +=) v (optional);
 *   outp planes [row, outp_col0+n] = split(v) (optional).
 *   ROWS: row = b*out_T + t*out_t_mul + out_t_add.  WGRAD: row = m.
 * ---------------------------------------------------------------------------------------------- */
#define DRN_GEMM_ROWS 0
#define DRN_GEMM_WGRAD 2

#define DRN_OUT_STORE 0
#define DRN_OUT_ADD 1
#define DRN_OUT_ATOMIC 2

typedef struct {
  int32_t form;
  int32_t b_mn;          /* ROWS only: weight operand layout, see above */
  drn_planes_t a, b;
  int32_t B, T;          /* ROWS: output rows = B samples x T time slots (in units of A's t axis) */
  int32_t N, K;          /* ROWS: N output channels, K contraction length per tap (multiple of 64).
                            WGRAD: M = a.C-range [m0..), N = columns of Bm used, K ignored */
  int32_t M;             /* WGRAD only: rows of D (channels of A used, starting at a_c0) */
  int32_t ntaps;
  int32_t tap_shift[DRN_MAX_TAPS], tap_par[DRN_MAX_TAPS], tap_w[DRN_MAX_TAPS];
  int32_t a_c0, b_c0;    /* channel offsets into A / Bm */
  int32_t nprod;         /* 1: hi*hi only (bf16 speed mode); 3: hi*hi+hi*lo+lo*hi (parity mode); 4: + lo*lo */
  int32_t split_k;       /* WGRAD: number of K splits (>1 requires out_mode ATOMIC on a zeroed buffer) */
  /* epilogue */
  float* out; int64_t out_ld; int32_t out_col0; int32_t out_mode;
  int64_t out_tap_stride; /* WGRAD: tap i is written at out + tap_w[i]*out_tap_stride */
  int32_t out_T, out_t_mul, out_t_add;
  const float* bias;
  const float* rowscale; int32_t rowscale_ld;
  float* out2; int64_t out2_ld;
  void* outp; int64_t outp_ld; int32_t outp_col0; int64_t outp_plane_stride;
  /* engine: 0 = tcgen05 tensor cores (product path), 1 = fp32 CUDA-core checker kernel (tests only) */
  int32_t engine;
  int32_t dbg_lbo, dbg_sbo, dbg_kadv; /* 0 = defaults; descriptor overrides used by the bring-up sweep test */
} drn_gemm_t;

int drn_gemm(const drn_gemm_t* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRN_B200_H */
