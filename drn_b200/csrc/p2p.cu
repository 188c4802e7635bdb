// Gradient exchange of the data-parallel path over NVLink peer memory (include/drn_b200.h: drn_p2p_allreduce_avg).
//
// The reference's multi-GPU mode is nn.DataParallel (main.py:99): replicas compute gradients of their batch shard, the
// gradients are summed.  Here every rank owns one flat fp32 gradient buffer (model/main_model.py:_run_backward); the buffers
// of all ranks are mapped into every process through CUDA IPC, and ONE kernel per region does the whole exchange:
//
//   ready exchange   rank r stores its epoch into flag word READY+r of every peer (st.release.sys) and waits until its own
//                    READY words carry the epoch of all peers: every rank has reached the call, so -- stream order -- every
//                    rank's gradients of the region are complete and visible;
//   reduce + publish the region is cut into `world` slices; rank r loads slice r from all `world` buffers (16-byte loads over
//                    NVLink, all issued before the first use), adds them in rank order, scales by 1/world and stores the
//                    result into all `world` buffers.  An element is read and then overwritten by ONE thread of ONE rank, so
//                    the in-place update needs no intermediate barrier and every rank ends with bit-identical values;
//   done exchange    the last CTA to finish (device counter) publishes DONE+r to every peer and waits for theirs: when the
//                    kernel completes, every peer's stores into this rank's buffer have landed.
//
// Per GPU the NVLink traffic is (world-1)/world of the region in each direction -- what a ring moves in 2(world-1) dependent
// steps moves here in one, which is what the NVSwitch topology is for (every peer at full bandwidth).  Spins are bounded
// (~30 s) and trap, so a missing peer fails the step instead of hanging the GPU.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace drn {

constexpr int P2P_THREADS = 512;
constexpr int p2p_unroll(int world) { return world >= 8 ? 2 : 16 / world; }  // ~16 x 16-byte loads in flight per thread
constexpr int FLAG_EPOCH = 0, FLAG_COUNTER = 1, FLAG_READY = 16, FLAG_DONE = 32;
static_assert(FLAG_DONE + DRN_P2P_MAX_RANKS <= DRN_P2P_FLAG_WORDS, "flag block too small");

struct P2PComm {
  int world, rank;
  float* buf[DRN_P2P_MAX_RANKS];
  unsigned* flags[DRN_P2P_MAX_RANKS];
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Peer data: written by another GPU's kernels before its READY store, read exactly once by this launch and only after the
// acquire of that flag (L1 is invalidated at every launch, so no line of it can be stale here): plain loads, kept out of L1.
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// made visible to the owner by the fence.sys + st.release.sys(DONE) that follow
__device__ __forceinline__ void st_peer(float4* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// epoch comparison that survives the 32-bit wrap
__device__ __forceinline__ bool reached(unsigned flag, unsigned epoch) { return static_cast<int>(flag - epoch) >= 0; }

__device__ __forceinline__ void wait_flag(const unsigned* f, unsigned epoch) {
  const long long t0 = clock64();
  while (!reached(ld_acquire_sys(f), epoch)) {
    if (clock64() - t0 > 60000000000LL) __trap();  // ~30 s: a peer never arrived (ranks may be seconds apart at start-up)
  }
}

template <int WORLD>
__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_kernel(const P2PComm c, long long off4, long long n4, float scale) {
  __shared__ int is_last;
  unsigned* mine = c.flags[c.rank];
  const unsigned epoch = ld_relaxed_u32(mine + FLAG_EPOCH) + 1u;  // bumped by the last CTA of this launch, after everyone read it
  const int tid = threadIdx.x;
  // ---- ready exchange -------------------------------------------------------------------------------------------------
  if (blockIdx.x == 0 && tid < WORLD && tid != c.rank) st_release_sys(c.flags[tid] + FLAG_READY + c.rank, epoch);
  if (tid < WORLD && tid != c.rank) wait_flag(mine + FLAG_READY + tid, epoch);
  __syncthreads();
  // ---- reduce slice `rank` of the region, publish it to every buffer ---------------------------------------------------
  const long long per = (n4 + WORLD - 1) / WORLD;
  const long long beg = off4 + per * c.rank;
  long long end = beg + per;
  if (end > off4 + n4) end = off4 + n4;
  constexpr int P2P_UNROLL = p2p_unroll(WORLD);
  const long long stride = static_cast<long long>(gridDim.x) * P2P_THREADS;
  for (long long i0 = beg + static_cast<long long>(blockIdx.x) * P2P_THREADS + tid; i0 < end; i0 += stride * P2P_UNROLL) {
    float4 v[P2P_UNROLL][WORLD];
#pragma unroll
    for (int u = 0; u < P2P_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < end) {
#pragma unroll
        for (int p = 0; p < WORLD; ++p) v[u][p] = ld_peer(reinterpret_cast<const float4*>(c.buf[p]) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < P2P_UNROLL; ++u) {
      const long long i = i0 + u * stride;
      if (i < end) {
        float4 s = v[u][0];
#pragma unroll
        for (int p = 1; p < WORLD; ++p) {
          s.x += v[u][p].x; s.y += v[u][p].y; s.z += v[u][p].z; s.w += v[u][p].w;
        }
        s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
#pragma unroll
        for (int p = 0; p < WORLD; ++p) st_peer(reinterpret_cast<float4*>(c.buf[p]) + i, s);
      }
    }
  }
  // ---- done exchange: the last CTA of this rank speaks for all of them ----------------------------------------------------
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(mine + FLAG_COUNTER, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence_system();  // the other CTAs' stores (fenced before their counter increments) precede the DONE stores below
  if (tid < WORLD && tid != c.rank) {
    st_release_sys(c.flags[tid] + FLAG_DONE + c.rank, epoch);
    wait_flag(mine + FLAG_DONE + tid, epoch);
  }
  __syncthreads();
  if (tid == 0) {
    mine[FLAG_COUNTER] = 0u;
    mine[FLAG_EPOCH] = epoch;
  }
}

template <int WORLD>
static void launch_p2p(const P2PComm& c, long long off4, long long n4, int ctas, cudaStream_t st) {
  // Plain stream order (never a programmatic dependent launch: the kernel must not start before the gradients are complete).
  // An even grid is launched as clusters of 2 CTAs, i.e. on whole TPCs: the exchange runs beside the CTA-pair contraction
  // kernel (prop_fc weight-gradient chunks on 64 of the 74 TPCs), whose clusters need BOTH SMs of a TPC -- 20 single CTAs
  // scattered over 20 TPCs would leave it 54 pairs and a second wave, 10 clusters leave it the 64 it asks for.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(P2P_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (ctas % 2 == 0) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, p2p_allreduce_kernel<WORLD>, c, off4, n4, 1.0f / WORLD);
}

typedef CUresult (*cuMemGetAddressRange_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
static cuMemGetAddressRange_fn address_range_fn() {
  static cuMemGetAddressRange_fn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<cuMemGetAddressRange_fn>(f);
  }
  return fn;
}

}  // namespace drn

using namespace drn;

extern "C" int drn_ipc_export(const void* ptr, unsigned char* handle64, int64_t* offset) {
  if (!ptr || !handle64 || !offset) return fail(DRN_EINVAL, "drn_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cuMemGetAddressRange_fn fn = address_range_fn();
  if (!fn) return fail(DRN_EINVAL, "drn_ipc_export: cuMemGetAddressRange unavailable");
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = fn(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) return fail(static_cast<int>(r), "drn_ipc_export: cuMemGetAddressRange failed (%d)", static_cast<int>(r));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(static_cast<int>(e), "drn_ipc_export: cudaIpcGetMemHandle: %s (cudaMalloc memory only: no expandable segments)",
                cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return 0;
}

extern "C" int drn_ipc_open(const unsigned char* handle64, int64_t offset, void** out) {
  if (!handle64 || !out) return fail(DRN_EINVAL, "drn_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(static_cast<int>(e), "drn_ipc_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  *out = static_cast<char*>(base) + offset;
  return 0;
}

extern "C" int drn_ipc_close(void* ptr, int64_t offset) {
  if (!ptr) return 0;
  cudaError_t e = cudaIpcCloseMemHandle(static_cast<char*>(ptr) - offset);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(static_cast<int>(e), "drn_ipc_close: %s", cudaGetErrorString(e));
  }
  return 0;
}

extern "C" int drn_p2p_allreduce_avg(const drn_p2p_t* comm, int64_t offset, int64_t n, int ctas, void* stream) {
  if (!comm) return fail(DRN_EINVAL, "drn_p2p_allreduce_avg: null comm");
  const int W = comm->world;
  if (W < 2 || W > DRN_P2P_MAX_RANKS || comm->rank < 0 || comm->rank >= W)
    return fail(DRN_EINVAL, "drn_p2p_allreduce_avg: world %d / rank %d out of range (2..%d ranks)", W, comm->rank, DRN_P2P_MAX_RANKS);
  if (offset < 0 || n < 0 || (offset & 3) || (n & 3))
    return fail(DRN_EINVAL, "drn_p2p_allreduce_avg: offset %lld and n %lld must be non-negative multiples of 4 floats",
                static_cast<long long>(offset), static_cast<long long>(n));
  P2PComm c{};
  c.world = W;
  c.rank = comm->rank;
  for (int r = 0; r < W; ++r) {
    if (!comm->buf[r] || !comm->flags[r]) return fail(DRN_EINVAL, "drn_p2p_allreduce_avg: rank %d buffer / flags not mapped", r);
    if (reinterpret_cast<uintptr_t>(comm->buf[r]) & 15) return fail(DRN_EINVAL, "drn_p2p_allreduce_avg: rank %d buffer not 16-byte aligned", r);
    c.buf[r] = comm->buf[r];
    c.flags[r] = comm->flags[r];
  }
  if (n == 0) return 0;  // every rank skips the same call: the epochs stay aligned
  const long long n4 = n >> 2, off4 = offset >> 2;
  // Default grid = the 20 SMs (10 TPCs) the chunked prop_fc weight gradient leaves free; ~128 KB of loads in flight per CTA
  // cover the NVLink latency.  No CTA waits for another one of this grid, so a larger grid would simply run in waves.
  if (ctas <= 0) ctas = 20;
  const long long per = (n4 + W - 1) / W;
  const int unroll = p2p_unroll(W);
  const long long want = (per + P2P_THREADS * unroll - 1) / (P2P_THREADS * unroll);
  if (ctas > want) ctas = static_cast<int>(want);
  if (ctas < 1) ctas = 1;
  if (ctas > 2 && (ctas & 1)) --ctas;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (W) {
    case 2: launch_p2p<2>(c, off4, n4, ctas, st); break;
    case 3: launch_p2p<3>(c, off4, n4, ctas, st); break;
    case 4: launch_p2p<4>(c, off4, n4, ctas, st); break;
    case 5: launch_p2p<5>(c, off4, n4, ctas, st); break;
    case 6: launch_p2p<6>(c, off4, n4, ctas, st); break;
    case 7: launch_p2p<7>(c, off4, n4, ctas, st); break;
    default: launch_p2p<8>(c, off4, n4, ctas, st); break;
  }
  return check_launch("p2p_allreduce_kernel");
}
