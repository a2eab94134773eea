// driver.hpp -- C++ view of the driver-layer operators for the solver's registered-operator mode.
#pragma once
#include "vecops.hpp"
namespace ab200 {

// A square CSR operator resident in HBM (int32 indices), registered for a solve with ab200_register_csr_op_*.
template <typename T>
struct CsrOpDesc {
  int nrows = 0;
  long long nnz = 0;
  const int* rowptr = nullptr;
  const int* col = nullptr;
  const T* val = nullptr;
  // row-partitioned operator under a communicator (PARPACK layout): columns >= nrows address halo[], which holds
  // the last halo_lo entries of the lower neighbour's x followed by the first halo_hi entries of the upper one's
  int comm = 0;
  int halo_lo = 0, halo_hi = 0;
  T* halo = nullptr;
};

// y = A x on the library stream
template <typename T>
int csr_op_apply(const CsrOpDesc<T>& op, const T* x, T* y);
// Fused K1+K2+K3: v_j = inv*resid, y = A v_j, dots_out = {v_j^T y, y^T y}; returns 1 when the operator's row
// lengths do not suit the fused kernel (the caller then runs start_step + csr_op_apply)
// whether csr_op_apply_fused would take the operator at all (short rows only)
template <typename T>
inline bool csr_op_fusable(const CsrOpDesc<T>& op) {
  return op.nrows > 0 && (op.nnz > 0 ? (double)op.nnz / op.nrows : 8.0) <= 7.9;
}
// gate != nullptr: the scale is formed on the device from the previous step's mailbox slot (device-resident sweep),
// and the kernel exits at once when the sweep's stop flag is set or the gate trips (see StepGate, vecops.hpp)
template <typename T>
int csr_op_apply_fused(const CsrOpDesc<T>& op, T inv, const StepGate<T>* gate, const T* stop, const T* resid, T* vj,
                       T* y, T* partial, T* dots_out, unsigned int* ticket);

}  // namespace ab200
