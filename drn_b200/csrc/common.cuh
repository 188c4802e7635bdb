// Shared host-side helpers of libdrn_sm100.so (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/drn_b200.h"

namespace drn {

char* err_buf();  // thread-local, 512 bytes

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------------
// A step is ~100 dependent kernel launches replayed from CUDA graphs; with plain stream order kernel i+1 is not scheduled
// before kernel i has drained and flushed.  Every kernel of the library is launched through launch_k(), which marks it
// "programmatic stream serialization allowed": its CTAs may become resident as soon as all CTAs of the kernel before it have
// STARTED (every kernel executes pdl_sync() first: griddepcontrol.launch_dependents, then griddepcontrol.wait), run their
// prologue, and block in griddepcontrol.wait until the previous kernel has completed and its memory is visible.  No kernel
// touches global memory before its wait, so the data dependences are exactly those of the stream order.  Captured into a
// CUDA graph these become programmatic dependency edges.  Opt-in (DRN_PDL=1): replayed from the two CUDA graphs of a step it
// measured neutral on B200 (profiles/r01_ab_pdl_v14.log), so by default kernels are launched without the attribute and the
// two instructions are no-ops.
bool pdl_enabled();  // api.cu: reads DRN_PDL once
// DRN_CARVEOUT=1 (api.cu): every kernel of the library asks for the SAME L1 / shared-memory split as the persistent contraction
// kernel (maximum shared memory).  An SM has to drain before its split can change, so a step that alternates between a
// 197 KB-shared-memory contraction and default-split element-wise kernels pays a reconfiguration on every transition.
void carveout_once(const void* kernel);
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  carveout_once(reinterpret_cast<const void*>(kernel));
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface through check_launch (cudaGetLastError)
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_trigger();
  pdl_wait();
}
#endif

}  // namespace drn
