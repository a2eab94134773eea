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

// One reduction across the ranks of a communicator that is FUSED into the kernels on both sides of it (multi-GPU,
// peer memory over NVLink; comm_nccl.cpp, peer.cuh): the last CTA of the producing kernel stores this rank's partial
// sums into every peer's slot and publishes `seq`; the consuming kernel waits for every rank's `seq` and adds the
// slots in rank order (bit-identical on all ranks).  nranks == 0: not in use (single rank, or reduced by a separate
// all-reduce).  Replaces MPI_ALLREDUCE of pdsaitr.f:604,720 / pdnorm2.f:72-80 without a launch of its own.
struct PeerReduce {
  unsigned char* base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int rank = 0, nranks = 0;
  int kind = 0;                  // which slot family: 0 = [h, ||w||^2], 1 = [s, ||r||^2], 2 = [||r'||^2]
  unsigned long long seq = 0;
  long long timeout_cycles = 0;  // > 0: trap instead of waiting for ever (AB200_P2P_TIMEOUT_S)
};

// How a gated step kernel decides, on the device, whether and with which scale to run: the mailbox segments of
// the step that precedes it (prev_j columns) hold ||w||^2 = A[prev_j], ||r||^2 = B[prev_j], ||r'||^2 = C[0] and
// the DGKS flag C[1].  The kernel reproduces the host logic of IrlBase::finish_orth bit for bit:
//   rn = sqrt(B[j]);  if the DGKS pass ran: rn1 = sqrt(C[0]); rn1 > 0.717*rn ? rn = rn1 : TRIP (third pass)
//   rn < tiny or rn == 0 -> TRIP (restart / dlascl path);  otherwise scale = 1/rn.
// On a trip *stop = stop_code and nothing is written.
template <typename T>
struct StepGate {
  const T* A = nullptr;
  const T* B = nullptr;
  const T* C = nullptr;
  int prev_j = 0;
  T tiny = 0;
  T* stop = nullptr;
  T stop_code = 0;
  // multi-GPU: C[0] of the previous step is still spread over the ranks' peer slots; the gate waits for it, adds it up
  // and leaves the sum in c_log[0] for the host (peer.nranks == 0: C[0] is already reduced)
  PeerReduce peer;
  T* c_log = nullptr;
};

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
  // mb_out[0] = max |x_i| (local rows).  false: not provided by this backend.
  virtual bool absmax(int64_t /*n*/, const T* /*x*/, T* /*mb_out*/) { return false; }
  // LAPACK xLARNV(idist=2) stream (dgetv0.f:236): x_i uniform(-1,1); iseed is advanced on the host
  virtual void larnv_uniform_m1_1(int64_t n, int iseed[4], T* x) = 0;
  // K1+K2 (dsaitr.f:438-442,464): vj = resid*inv ; out_x = vj ; if bx != nullptr: bx *= inv
  // (for bmat='I' pass bx_from_resid=true to write bx = resid*inv instead of scaling in place)
  virtual void start_step(int64_t n, T inv_rnorm, const T* resid, T* vj, T* out_x, T* bx,
                          bool bx_from_resid) = 0;
  // ---- device-resident sweep (IrlBase::extend, deferred mode) ---------------------------------
  // A whole sweep of Lanczos/Arnoldi steps is enqueued without a host round trip: every step deposits its
  // reductions in its own mailbox slot, the scale 1/||r|| of the next step and the reference's rare-path tests
  // (third DGKS pass dsaitr.f:768-780, tiny / zero norm dsaitr.f:378-453) are evaluated on the device by the
  // gated start of step, and a sticky stop flag turns every later kernel of the sweep into an early exit, so
  // that the host can resume from exactly the state the reference would be in.
  virtual bool deferred_ok() const { return false; }
  // every step kernel (start_step, start_step_gated, orth_step, and a registered operator's fused product)
  // issued from now on exits at once when *stop != 0; nullptr switches the check off
  virtual void set_stop_flag(T* /*stop*/) {}
  // K1+K2 with the scale formed on the device from the previous step's mailbox slot (see StepGate)
  virtual void start_step_gated(int64_t /*n*/, const StepGate<T>& /*g*/, const T* /*resid*/, T* /*vj*/, T* /*out_x*/,
                                T* /*bx*/) {}
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
