// gate.cuh -- device side of StepGate (vecops.hpp): the scale of the next Lanczos/Arnoldi step and the reference's
// rare-path tests, evaluated by every thread of a gated kernel from the previous step's mailbox slot.
#pragma once
#include "vecops.hpp"

namespace ab200 {

// true: the step may run with v_j = inv * resid.  false: a rare path of dsaitr.f / dnaitr.f is due (third DGKS pass,
// :768-780; norm below safmin or exactly zero, :378-453) -- the caller writes g.stop_code to *g.stop from ONE thread
// and returns without touching memory.  Same arithmetic as IrlBase::finish_orth on the host (IEEE sqrt and division,
// the REAL literal 0.717), so both sides always agree.
template <typename T>
__device__ __forceinline__ bool gate_eval(const StepGate<T>& g, T& inv) {
  const T wn = sqrt(g.A[g.prev_j]);
  T rn = sqrt(g.B[g.prev_j]);
  bool ok = true;
  if (!(rn > T(0.717f) * wn)) {  // the DGKS pass ran (dsaitr.f:656)
    const T rn1 = sqrt(g.C[0]);
    if (rn1 > T(0.717f) * rn) rn = rn1;
    else ok = false;
  }
  if (!(rn >= g.tiny) || !(rn > T(0))) ok = false;
  inv = T(1) / rn;
  return ok;
}

template <typename T>
__device__ __forceinline__ bool stopped(const T* stop) {
  return stop != nullptr && *reinterpret_cast<const volatile T*>(stop) != T(0);
}

}  // namespace ab200
