// vecops_tma.cu -- TMA-tiled fast path of the tall-skinny kernels (placeholder until the first GPU
// validation of the generic kernels has passed; every entry returns false => generic kernels run).
#include "vecops_cuda.cuh"

namespace ab200 {

template <typename T>
struct CudaVecOps<T>::TmaCache {};

template <typename T>
bool CudaVecOps<T>::fast_path_ok(int64_t, int, const T*, int64_t) const {
  return false;
}
template <typename T>
bool CudaVecOps<T>::orth_step_tma(int64_t, int, const T*, int64_t, const T*, T*, T*, T*, T*) {
  return false;
}
template <typename T>
bool CudaVecOps<T>::dots_tma(int64_t, int, const T*, int64_t, const T*, const T*, T*) {
  return false;
}
template <typename T>
bool CudaVecOps<T>::vq_tma(int64_t, int, int, const T*, int64_t, const T*, T*, int64_t, bool, T, T, int, T*, T*) {
  return false;
}
template <typename T>
void CudaVecOps<T>::tma_release() {}

template struct CudaVecOps<double>::TmaCache;
template struct CudaVecOps<float>::TmaCache;
template bool CudaVecOps<double>::fast_path_ok(int64_t, int, const double*, int64_t) const;
template bool CudaVecOps<float>::fast_path_ok(int64_t, int, const float*, int64_t) const;
template bool CudaVecOps<double>::orth_step_tma(int64_t, int, const double*, int64_t, const double*, double*, double*, double*, double*);
template bool CudaVecOps<float>::orth_step_tma(int64_t, int, const float*, int64_t, const float*, float*, float*, float*, float*);
template bool CudaVecOps<double>::dots_tma(int64_t, int, const double*, int64_t, const double*, const double*, double*);
template bool CudaVecOps<float>::dots_tma(int64_t, int, const float*, int64_t, const float*, const float*, float*);
template bool CudaVecOps<double>::vq_tma(int64_t, int, int, const double*, int64_t, const double*, double*, int64_t, bool, double, double, int, double*, double*);
template bool CudaVecOps<float>::vq_tma(int64_t, int, int, const float*, int64_t, const float*, float*, int64_t, bool, float, float, int, float*, float*);
template void CudaVecOps<double>::tma_release();
template void CudaVecOps<float>::tma_release();

}  // namespace ab200
