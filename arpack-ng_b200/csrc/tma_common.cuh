// tma_common.cuh -- mbarrier / TMA plumbing and deterministic grid reduction shared by the TMA-tiled kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <vector>

#include "peer.cuh"

namespace ab200 {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  uint32_t spins = 0;
  long long t0 = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    // a lost TMA completion must surface as an error, never as a hung GPU: trap after ~10 s
    if (!ok && ((++spins & 0xFFFu) == 0)) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000LL) __trap();
    }
  } while (!ok);
}
// 2-D tiled load: box at (c0 = fastest coordinate, c1) of the tensor described by tmap -> shared memory
__device__ __forceinline__ void load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void load_1d(uint32_t smem_dst, const CUtensorMap* tmap, int c0, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(smem_u32(bar))
      : "memory");
}

// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl() may be scheduled while its predecessor
// on the stream is still running -- as soon as every CTA of the predecessor has executed pdl_trigger() and SM
// resources allow -- and runs its prologue (shared-memory carve-up, mbarrier initialisation) there.  It must execute
// pdl_wait() BEFORE ITS FIRST GLOBAL-MEMORY ACCESS: that returns when the predecessor has completed and flushed.  The
// per-kernel launch gap and prologue move off the critical path.  Without the launch attribute both calls are no-ops.
// Measured (B200, config 2, A/B on one box): 500.7-502.7 vs 500.4-500.6 steps/s on one GPU, 917.9-919.0 vs 917.8 on
// two -- within noise: with the sweeps enqueued far ahead of the device the launch gap is not what limits a step.  The
// attribute is therefore OFF by default (AB200_PDL=1 switches it on); the kernels keep the two calls.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Combine per-CTA partials deterministically: the CTA that takes the last ticket sums partial[b*pcols + c] over
// b (lane-strided, then an xor tree) and writes out[c].  Must be reached by every thread of every CTA.
// Multi-GPU (pub != nullptr && pub->nranks > 0): the rank's sums do not go to out[] but straight into every peer's
// slot of the fused reduction *pub (peer.cuh) -- the all-reduce that used to follow this kernel as a launch of its
// own starts here and ends in the prologue of the kernel that consumes the values.
template <typename T>
__device__ void finish_grid_reduce(T* partial, int pcols, int ncols, T* out, unsigned int* ticket,
                                   const PeerReduce* pub = nullptr) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const bool fused = pub != nullptr && pub->nranks > 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  for (int c = warp; c < ncols; c += nwarps) {
    T s = T(0);
    for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(partial + (size_t)b * pcols + c);
    s = warp_sum(s);
    if (lane == 0) {
      if (fused) peer_put<T>(*pub, c, s);
      else out[c] = s;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
  if (fused) peer_publish(*pub);
}

// per-function launch attributes (dynamic shared memory size) belong to a (function, device) pair: a process that
// drives several GPUs must set them once per device, not once per process
constexpr int kMaxDevices = 16;
inline int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  return (dev < 0 ? 0 : dev) % kMaxDevices;
}

// launch with the programmatic-stream-serialization attribute when AB200_PDL=1 (default: a plain launch)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args&&... args) {
  static const bool on = getenv("AB200_PDL") && getenv("AB200_PDL")[0] == '1';
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = on ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- host: tensor-map descriptors, cached (the encode call is a driver round trip of ~0.4 ms) ----
// 2-D map over a column-major n x ncols matrix (leading dimension ldv), box = box_rows x box_cols;
// ncols == 0 -> 1-D map over a vector of n elements, box = box_rows.
bool get_tensor_map(CUtensorMap* map, const void* base, int elem_size, int64_t n, int64_t ldv, int ncols,
                    int box_rows, int box_cols);
bool tensor_maps_available();

}  // namespace tma
}  // namespace ab200
