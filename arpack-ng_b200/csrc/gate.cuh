// gate.cuh -- device side of StepGate (vecops.hpp): the scale of the next Lanczos/Arnoldi step and the reference's
// rare-path tests, evaluated by a gated kernel from the previous step's mailbox slot.
#pragma once
#include "peer.cuh"
#include "vecops.hpp"

namespace ab200 {

template <typename T>
__device__ __forceinline__ bool stopped(const T* stop) {
  return stop != nullptr && *reinterpret_cast<const volatile T*>(stop) != T(0);
}

// Block-cooperative (every thread of the CTA must call it; it synchronises the CTA).
// true: the step may run with v_j = inv * resid.  false: a rare path of dsaitr.f / dnaitr.f is due (third DGKS pass,
// :768-780; norm below safmin or exactly zero, :378-453) -- the caller writes g.stop_code to *g.stop from ONE thread
// of ONE block and returns without touching memory.  Same arithmetic as IrlBase::finish_orth on the host (IEEE sqrt
// and division, the REAL literal 0.717), so both sides always agree.
// Multi-GPU: when the DGKS pass ran, its norm ||r'||^2 may still be spread over the ranks' peer slots (g.peer); the
// gate waits for all ranks, adds the slots in rank order and block 0 leaves the sum in g.c_log[0] for the host.
template <typename T>
__device__ __forceinline__ bool gate_eval_block(const StepGate<T>& g, T& inv) {
  __shared__ T s_inv;
  __shared__ int s_ok;
  if (threadIdx.x < 32) {   // the first warp decides (every kernel that uses the gate has at least one full warp)
    const int lane = (int)threadIdx.x;
    const T wn = sqrt(g.A[g.prev_j]);
    T rn = sqrt(g.B[g.prev_j]);
    bool ok = true;
    if (!(rn > T(0.717f) * wn)) {  // the DGKS pass ran (dsaitr.f:656); uniform across the warp
      T c0;
      if (g.peer.nranks > 0) {
        if (lane < g.peer.nranks) peer_wait_rank(g.peer, lane);   // one lane per rank: the waits overlap
        __syncwarp();
        c0 = peer_sum<T>(g.peer, 0);
        if (blockIdx.x == 0 && lane == 0 && g.c_log != nullptr) g.c_log[0] = c0;
      } else {
        c0 = g.C[0];
      }
      const T rn1 = sqrt(c0);
      if (rn1 > T(0.717f) * rn) rn = rn1;
      else ok = false;
    }
    if (!(rn >= g.tiny) || !(rn > T(0))) ok = false;
    if (lane == 0) {
      s_inv = T(1) / rn;
      s_ok = ok ? 1 : 0;
    }
  }
  __syncthreads();
  inv = s_inv;
  return s_ok != 0;
}

}  // namespace ab200
