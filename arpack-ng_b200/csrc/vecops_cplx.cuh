// vecops_cplx.cuh -- CUDA (sm_100a) implementation of the VecOps device boundary for complex vectors
// (znaupd/zneupd, cnaupd/cneupd).
#pragma once
#include <complex>

#include "vecops_cuda.cuh"

namespace ab200 {

// VecOps<std::complex<R>>: Hermitian reductions (dot(x, y) = sum conj(x_i) y_i, dots(V, x) = V^H x), complex
// tall-skinny updates.  Everything that is type-agnostic on interleaved (re, im) pairs -- memory, mailbox, copies,
// real scalings, the xLARNV stream -- is delegated to a CudaVecOps<R> working on 2n reals.
template <typename R>
class CudaVecOpsZ final : public VecOps<std::complex<R>> {
 public:
  using T = std::complex<R>;
  explicit CudaVecOpsZ(cudaStream_t stream, NcclComm* comm = nullptr) : real_(stream, comm), stream_(stream) {
    int dev = 0;
    AB200_CUDA_CHECK(cudaGetDevice(&dev));
    AB200_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev));
  }
  ~CudaVecOpsZ() override { cudaFree(qbuf_); }

  T* alloc(size_t count) override { return reinterpret_cast<T*>(real_.alloc(2 * count)); }
  void release(T* p) override { real_.release(reinterpret_cast<R*>(p)); }
  void upload(T* d, const T* h, size_t c) override { real_.upload((R*)d, (const R*)h, 2 * c); }
  void download(T* h, const T* d, size_t c) override { real_.download((R*)h, (const R*)d, 2 * c); }
  void upload2d(T* d, size_t ldd, const T* h, size_t lds, size_t rows, size_t cols) override {
    real_.upload2d((R*)d, 2 * ldd, (const R*)h, 2 * lds, 2 * rows, cols);
  }
  void download2d(T* h, size_t ldd, const T* d, size_t lds, size_t rows, size_t cols) override {
    real_.download2d((R*)h, 2 * ldd, (const R*)d, 2 * lds, 2 * rows, cols);
  }
  void sync() override { real_.sync(); }
  bool is_device_pointer(const void* p) override { return real_.is_device_pointer(p); }

  T* mailbox(size_t count) override { return reinterpret_cast<T*>(real_.mailbox(2 * count)); }
  void fetch(T* host_dst, const T* mb, size_t count) override { real_.fetch((R*)host_dst, (const R*)mb, 2 * count); }
  void post(T* mb, const T* host_src, size_t count) override { real_.post((R*)mb, (const R*)host_src, 2 * count); }
  // PARPACK: MPI_ALLREDUCE(MPI_DOUBLE_COMPLEX, SUM) of pznaitr.f:437-449 == a sum of 2*count reals, in stream order
  void allreduce_sum(T* mb, size_t count) override { real_.allreduce_sum((R*)mb, 2 * count); }
  int rank() const override { return real_.rank(); }
  int nranks() const override { return real_.nranks(); }

  void copy(int64_t n, const T* x, T* y) override { real_.copy(2 * n, (const R*)x, (R*)y); }
  void zero(int64_t n, T* x) override { real_.zero(2 * n, (R*)x); }
  void scal(int64_t n, T alpha, T* x) override;
  void axpby_norm(int64_t n, T a, T b, const T* x, T* y, T* mb_nrm2_out) override;
  void dot(int64_t n, const T* x, const T* y, T* mb_out) override;
  // zlarnv(idist = 2) == dlarnv on the 2n interleaved reals (SRC/zgetv0.f:229-230)
  void larnv_uniform_m1_1(int64_t n, int iseed[4], T* x) override { real_.larnv_uniform_m1_1(2 * n, iseed, (R*)x); }
  void start_step(int64_t n, T inv_rnorm, const T* resid, T* vj, T* out_x, T* bx, bool bx_from_resid) override {
    real_.start_step(2 * n, inv_rnorm.real(), (const R*)resid, (R*)vj, (R*)out_x, (R*)bx, bx_from_resid);
  }
  void ger(int64_t n, int k, const T* resid, const T* w_host, T* z, int64_t ldz) override;

  void dots(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* mb_out) override;
  void update(int64_t n, int j, const T* v, int64_t ldv, const T* mb_coef, const T* src, T* dst,
              T* mb_nrm2) override;
  void orth_step(int64_t, int, const T*, int64_t, const T*, T*, T*, T*, T*) override {
    throw CudaError("CudaVecOpsZ::orth_step is not used by the complex solver");
  }
  void vq_update(int64_t n, int kin, int kout, T* v, int64_t ldv, const T* q_host, int ldq, bool with_resid,
                 T sigma, T beta, int beta_col, T* resid, T* mb_nrm2) override;
  void vq_out(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* m_host, int ldm, T* out,
              int64_t ldo) override;
  void copy2d(int64_t n, int cols, const T* src, int64_t lds, T* dst, int64_t ldd) override {
    real_.copy2d(2 * n, cols, (const R*)src, 2 * lds, (R*)dst, 2 * ldd);
  }

 private:
  CudaVecOps<R> real_;
  cudaStream_t stream_;
  int num_sms_ = 148;
  T* qbuf_ = nullptr;  // device copy of the small coefficient matrices
  size_t qbuf_count_ = 0;

  int reduce_grid(int64_t n) const;
  T* partial(size_t count) { return reinterpret_cast<T*>(real_.reduction_scratch(2 * count)); }
  T* stage_matrix(const T* host, int rows, int cols, int ld);
  void vq(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* qdev, T* out, int64_t ldo,
          bool with_resid, T sigma, T beta, int beta_col, T* resid, T* mb_nrm2);
};

}  // namespace ab200
