// vecops_cuda.cuh -- CUDA (sm_100a) implementation of the VecOps device boundary.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "vecops.hpp"

namespace ab200 {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};
#define AB200_CUDA_CHECK(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      throw ::ab200::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + \
                               __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
  } while (0)

// NCCL communicator wrapper (comm_nccl.cpp); opaque here
struct NcclComm;
void nccl_allreduce_sum(NcclComm* c, void* buf, size_t count, bool is_double, cudaStream_t s);
int nccl_rank(const NcclComm* c);
int nccl_nranks(const NcclComm* c);
// reductions fused into the kernels on both sides (peer memory; comm_nccl.cpp, peer.cuh)
bool nccl_peer_reduce_begin(NcclComm* c, int kind, PeerReduce* out);
void nccl_peer_reduce_finalize(NcclComm* c, const PeerReduce& pr, int count, void* log, const void* pred_w2,
                               const void* pred_r2, const void* stop, bool is_double, cudaStream_t s);

// process-wide launch statistics (bench.py reports gpu_launches from here)
struct LaunchStats {
  unsigned long long kernels = 0;
  unsigned long long allreduces = 0;
  unsigned long long fast_path = 0;   // launches of the TMA-tiled kernels
  unsigned long long fallback = 0;    // launches of the generic kernels on the tall-skinny ops
  unsigned long long fetches = 0;     // blocking device->host mailbox reads (host round trips)
};
LaunchStats& launch_stats();

// Per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).  Disabled by
// default; when enabled every profiled launch is bracketed by two event records and the elapsed times
// are accumulated per kernel name together with the algorithmic bytes the launch was charged with.
struct KernelProfile {
  const char* name;
  unsigned long long launches;
  double ms;
  double bytes;
};
class Profiler {
 public:
  bool enabled = false;
  void begin(cudaStream_t s, const char* name, double bytes);
  void end(cudaStream_t s);
  void flush();  // synchronises the pending events and accumulates
  void reset();
  const std::vector<KernelProfile>& table() { flush(); return table_; }

 private:
  struct Pending { int idx; cudaEvent_t e0, e1; };
  std::vector<KernelProfile> table_;
  std::vector<Pending> pending_;
  std::vector<cudaEvent_t> pool_;
  cudaEvent_t cur0_ = nullptr;
  int cur_idx_ = -1;
  cudaEvent_t get_event();
};
Profiler& profiler();
struct ProfScope {
  cudaStream_t s;
  bool on;
  ProfScope(cudaStream_t stream, const char* name, double bytes) : s(stream), on(profiler().enabled) {
    if (on) profiler().begin(s, name, bytes);
  }
  ~ProfScope() {
    if (on) profiler().end(s);
  }
};

template <typename T>
class CudaVecOps final : public VecOps<T> {
 public:
  explicit CudaVecOps(cudaStream_t stream, NcclComm* comm);
  ~CudaVecOps() override;

  T* alloc(size_t count) override;
  void release(T* p) override;
  void upload(T* dst_dev, const T* src_host, size_t count) override;
  void download(T* dst_host, const T* src_dev, size_t count) override;
  void upload2d(T* dst_dev, size_t ld_dst, const T* src_host, size_t ld_src, size_t rows, size_t cols) override;
  void download2d(T* dst_host, size_t ld_dst, const T* src_dev, size_t ld_src, size_t rows, size_t cols) override;
  void sync() override;
  bool is_device_pointer(const void* p) override;

  T* mailbox(size_t count) override;
  void fetch(T* host_dst, const T* mb, size_t count) override;
  void post(T* mb, const T* host_src, size_t count) override;
  void allreduce_sum(T* mb, size_t count) override;
  int rank() const override;
  int nranks() const override;

  void copy(int64_t n, const T* x, T* y) override;
  void zero(int64_t n, T* x) override;
  void scal(int64_t n, T alpha, T* x) override;
  void axpby_norm(int64_t n, T a, T b, const T* x, T* y, T* mb_nrm2_out) override;
  void dot(int64_t n, const T* x, const T* y, T* mb_out) override;
  bool absmax(int64_t n, const T* x, T* mb_out) override;
  void larnv_uniform_m1_1(int64_t n, int iseed[4], T* x) override;
  void start_step(int64_t n, T inv_rnorm, const T* resid, T* vj, T* out_x, T* bx, bool bx_from_resid) override;
  bool deferred_ok() const override { return true; }
  void set_stop_flag(T* stop) override { stop_ = stop; }
  void start_step_gated(int64_t n, const StepGate<T>& g, const T* resid, T* vj, T* out_x, T* bx) override;
  void ger(int64_t n, int k, const T* resid, const T* w_host, T* z, int64_t ldz) override;

  void dots(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* mb_out) override;
  void update(int64_t n, int j, const T* v, int64_t ldv, const T* mb_coef, const T* src, T* dst,
              T* mb_nrm2) override;
  void orth_step(int64_t n, int j, const T* v, int64_t ldv, const T* w, T* resid, T* mbA, T* mbB,
                 T* mbC) override;
  void vq_update(int64_t n, int kin, int kout, T* v, int64_t ldv, const T* q_host, int ldq, bool with_resid,
                 T sigma, T beta, int beta_col, T* resid, T* mb_nrm2) override;
  void vq_out(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* m_host, int ldm, T* out,
              int64_t ldo) override;
  void copy2d(int64_t n, int cols, const T* src, int64_t lds, T* dst, int64_t ldd) override;

  cudaStream_t stream() const { return stream_; }
  void set_stream(cudaStream_t s) { stream_ = s; }
  // a recycled back-end starts a new solve: nothing of the previous one may leak into it
  void reset_for_reuse() {
    stop_ = nullptr;
    has_pending_ = false;
    agreed_v_ = nullptr;
    agreed_ldv_ = -1;
    agreed_fuse_ = false;
  }
  // 0 = auto (TMA-tiled kernels when the layout allows), 1 = force the generic kernels
  void set_kernel_mode(int m) { kernel_mode_ = m; }
  // scratch of the deterministic two-stage reductions, shared with driver-layer kernels that run on the same stream
  T* reduction_scratch(size_t count) {
    ensure_partial(count);
    return partial_;
  }
  unsigned int* reduction_ticket() { return ticket_; }
  const T* stop_flag() const { return stop_; }
  // Multi-GPU: the norm ||r'||^2 of the last orth_step may still sit in the ranks' peer slots (its reduction is fused
  // into whichever kernel consumes it).  A gated consumer takes it over with attach_pending(); every other use of the
  // mailbox goes through resolve_pending(), which finishes the reduction with one small kernel.
  void attach_pending(StepGate<T>& g) {
    if (has_pending_ && g.C == pending_log_) {
      g.peer = pending_;
      g.c_log = pending_log_;
      has_pending_ = false;
    } else {
      resolve_pending();
    }
  }
  void resolve_pending();

 private:
  cudaStream_t stream_;
  NcclComm* comm_;
  int kernel_mode_ = 0;
  int num_sms_ = 148;
  // mailbox
  T* mb_dev_ = nullptr;
  size_t mb_count_ = 0;
  T* mb_pinned_ = nullptr;
  // sticky stop flag of a device-resident sweep (mailbox entry, 0 = run); checked by every step kernel
  T* stop_ = nullptr;
  // fused reduction of ||r'||^2 that no kernel has consumed yet (see attach_pending)
  bool has_pending_ = false;
  PeerReduce pending_;
  T* pending_log_ = nullptr;
  const T* pending_w2_ = nullptr;
  const T* pending_r2_ = nullptr;
  const T* pending_stop_ = nullptr;   // the sweep's stop flag at the time the reduction was started
  // Fused reductions are a protocol between the ranks' kernels: either every rank runs the TMA-tiled kernels for this
  // solve's V (alignment, leading dimension -- rank-local properties) or nobody fuses.  Agreed once per (V, ldv).
  const void* agreed_v_ = nullptr;
  int64_t agreed_ldv_ = -1;
  bool agreed_fuse_ = false;
  bool ranks_agree_on_fusing(int64_t n, int j, const T* v, int64_t ldv);
  // two-stage reduction scratch: partial_[grid][pcols_] and a ticket counter
  T* partial_ = nullptr;
  size_t partial_count_ = 0;
  unsigned int* ticket_ = nullptr;
  // small device buffer for Q / coefficient matrices
  T* qbuf_ = nullptr;
  size_t qbuf_count_ = 0;
  T* qpinned_ = nullptr;
  size_t qpinned_count_ = 0;

  void ensure_partial(size_t count);
  T* stage_matrix(const T* host, int rows, int cols, int ld);  // -> device, packed rows x cols, column-major
  bool fast_path_ok(int64_t n, int j, const T* v, int64_t ldv) const;
  int reduce_grid(int64_t n) const;
  void vq_smem_attr(int kin);

  // generic kernels (vecops_cuda.cu)
  void dots_generic(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* mb_out);
  void update_generic(int64_t n, int j, const T* v, int64_t ldv, const T* coef, const T* src, T* dst, T* nrm2,
                      const T* pred_w2, const T* pred_r2, T* flag_out);
  // TMA-tiled kernels (vecops_tma.cu)
  bool orth_step_tma(int64_t n, int j, const T* v, int64_t ldv, const T* w, T* resid, T* mbA, T* mbB, T* mbC);
  bool dots_tma(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* mb_out);
  bool vq_tma(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* qdev, T* out, int64_t ldo,
              bool with_resid, T sigma, T beta, int beta_col, T* resid, T* mb_nrm2);
  // FP64 tensor-core form of the same update (vq_mma.cu); false: not applicable (FP32, layout) -> vq_tma / generic
  bool vq_mma(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* qdev, const T* q_host, int ldq, T* out,
              int64_t ldo, bool with_resid, T sigma, T beta, int beta_col, T* resid, T* mb_nrm2);
};

}  // namespace ab200
