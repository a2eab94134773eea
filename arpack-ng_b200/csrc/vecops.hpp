// vecops.hpp -- the device boundary of the IRLM/IRAM host control code.
//
// Every n-length operation of the reference's hot path is expressed
// through this interface.  The shipped library has exactly one implementation, CudaVecOps
// (vecops_cuda.cu: hand-written sm_100a kernels).  There is NO CPU implementation in the product;
// tests/hostdouble/ carries a test double so that the host state machine can be exercised by the
// CPU test-suite, and it is never linked into libarpack_b200.so.
//
// All calls are asynchronous on the backend's stream except fetch()/sync().  "Mailbox" pointers
// (mb) address a small device buffer owned by the backend; reductions deposit their results there
// and the host reads them once per Lanczos step with fetch().
#pragma once
#include <cstddef>
#include <cstdint>

namespace ab200 {

template <typename T>
struct VecOps {
  virtual ~VecOps() {}

  // ---- memory -------------------------------------------------------------------------------
  virtual T* alloc(size_t count) = 0;  // device allocation (count elements)
  virtual void release(T* p) = 0;
  virtual void upload(T* dst_dev, const T* src_host, size_t count) = 0;    // H2D, async
  virtual void download(T* dst_host, const T* src_dev, size_t count) = 0;  // D2H, async
  virtual void upload2d(T* dst_dev, size_t ld_dst, const T* src_host, size_t ld_src, size_t rows,
                        size_t cols) = 0;
  virtual void download2d(T* dst_host, size_t ld_dst, const T* src_dev, size_t ld_src, size_t rows,
                          size_t cols) = 0;
  virtual void sync() = 0;
  // true when p is memory the kernels can address directly (device / managed / registered-mapped)
  virtual bool is_device_pointer(const void* p) = 0;

  // ---- mailbox ------------------------------------------------------------------------------
  virtual T* mailbox(size_t count) = 0;  // (re)allocates the device mailbox, returns its base
  // copy count mailbox entries starting at mb to host and wait for them
  virtual void fetch(T* host_dst, const T* mb, size_t count) = 0;
  // host -> mailbox (coefficients computed on the host, e.g. none in the common path)
  virtual void post(T* mb, const T* host_src, size_t count) = 0;
  // PARPACK: sum the mailbox segment across ranks, in stream order (replaces MPI_ALLREDUCE,
  // PARPACK/SRC/MPI/pdsaitr.f:604,720, pdnorm2.f:72-80).  No-op for a single rank.
  virtual void allreduce_sum(T* mb, size_t count) = 0;
  virtual int rank() const = 0;
  virtual int nranks() const = 0;

  // ---- BLAS-1 shaped kernels ----------------------------------------------------------------
  virtual void copy(int64_t n, const T* x, T* y) = 0;
  virtual void zero(int64_t n, T* x) = 0;
  // x := alpha * x
  virtual void scal(int64_t n, T alpha, T* x) = 0;
  // y := a*y + b*x ; if nrm2_out != nullptr also *nrm2_out = sum(y_new^2)   (dsapps.f:491-493 + dsaup2.f:807)
  virtual void axpby_norm(int64_t n, T a, T b, const T* x, T* y, T* mb_nrm2_out) = 0;
  // mb_out[0] = sum x_i*y_i
  virtual void dot(int64_t n, const T* x, const T* y, T* mb_out) = 0;
  // LAPACK xLARNV(idist=2) stream (dgetv0.f:236): x_i uniform(-1,1); iseed is advanced on the host
  virtual void larnv_uniform_m1_1(int64_t n, int iseed[4], T* x) = 0;
  // K1+K2 (dsaitr.f:438-442,464): vj = resid*inv ; out_x = vj ; if bx != nullptr: bx *= inv
  // (for bmat='I' pass bx_from_resid=true to write bx = resid*inv instead of scaling in place)
  virtual void start_step(int64_t n, T inv_rnorm, const T* resid, T* vj, T* out_x, T* bx,
                          bool bx_from_resid) = 0;
  // Speculative K1+K2 of the NEXT step, enqueued before the host has read the mailbox of the orth_step just issued:
  // the scale is formed on the device, inv = 1/sqrt(mbC[1] != 0 ? mbC[0] : mbB[j]) -- the value the host will derive
  // in the common path -- and nothing is written when that norm is below `tiny`.  Lets the device work through the
  // host round trip of fetch_marked().  Returns false when the backend does not support it.
  virtual bool start_step_speculative(int64_t /*n*/, int /*j*/, const T* /*mbB*/, const T* /*mbC*/, T /*tiny*/,
                                      const T* /*resid*/, T* /*vj*/, T* /*out_x*/, T* /*bx*/) {
    return false;
  }
  // mark_fetch_point(): remember the current end of the stream; fetch_marked(): like fetch(), but waits only for
  // the work enqueued before the mark (so kernels issued after it keep running during the copy)
  virtual void mark_fetch_point() {}
  virtual void fetch_marked(T* host_dst, const T* mb, size_t count) { fetch(host_dst, mb, count); }
  // rank-1 purification Z(:,0:k) += resid * w^T  (dseupd.f:857); w is a host vector
  virtual void ger(int64_t n, int k, const T* resid, const T* w_host, T* z, int64_t ldz) = 0;

  // ---- tall-skinny kernels on V(n x j), column-major, leading dimension ldv ------------------
  // K6 (+K5): mb_out[0..j) = V_j^T x ; mb_out[j] = sum x_i*y_i     (y may alias x)
  virtual void dots(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* mb_out) = 0;
  // K7 (+K8): dst = src - V_j * mb_coef[0..j) ; if mb_nrm2 != nullptr: *mb_nrm2 = sum dst^2 (src may alias dst)
  virtual void update(int64_t n, int j, const T* v, int64_t ldv, const T* mb_coef, const T* src, T* dst,
                      T* mb_nrm2) = 0;
  // The fused Lanczos/Arnoldi orthogonalisation of one step for bmat='I' (K4..K10):
  //   mbA[0..j) = h = V_j^T w,  mbA[j] = ||w||^2
  //   resid = w - V_j h,        mbB[j] = ||resid||^2,  mbB[0..j) = s = V_j^T resid  (speculative)
  //   if sqrt(mbB[j]) <= 0.717*sqrt(mbA[j])   (DGKS test, dsaitr.f:656, decided on the device):
  //        resid -= V_j s,      mbC[0] = ||resid||^2 ; mbC[1] = 1   else mbC[1] = 0
  // all-reduces are inserted between the stages when nranks() > 1.
  virtual void orth_step(int64_t n, int j, const T* v, int64_t ldv, const T* w, T* resid, T* mbA, T* mbB,
                         T* mbC) = 0;
  // K12..K15 (+K16): in place V(:,0:kout) = V(:,0:kin) * Q(0:kin,0:kout)  (q_host column-major, ld ldq);
  // optional fused residual update resid = sigma*resid + beta*Vnew(:,beta_col) and its squared norm.
  virtual void vq_update(int64_t n, int kin, int kout, T* v, int64_t ldv, const T* q_host, int ldq,
                         bool with_resid, T sigma, T beta, int beta_col, T* resid, T* mb_nrm2) = 0;
  // out(:,0:kout) = V(:,0:kin) * M  without touching V (dneupd: Z = V*M)
  virtual void vq_out(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* m_host, int ldm, T* out,
                      int64_t ldo) = 0;
  // strided 2-D copy on the device: dst(:,0:cols) = src(:,0:cols)
  virtual void copy2d(int64_t n, int cols, const T* src, int64_t lds, T* dst, int64_t ldd) = 0;
};

}  // namespace ab200
