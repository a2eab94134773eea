// tests/hostdouble/hostdouble.cpp -- TEST DOUBLE, never part of the product.
//
// The product's host control code (arpack-ng_b200/csrc/irl_*.hpp) talks to the device through the
// VecOps interface, whose only shipped implementation is CUDA.  To let the CPU test-suite exercise
// that control code (state machine, ncv-sized host math, PARPACK semantics) without a GPU, this file
// provides a plain-loop VecOps and a small C API around the solvers.  It is compiled into
// tests/_build/libab200_hostdouble.so by tests/conftest.py and is not linked into, loaded by, or
// reachable from libarpack_b200.so.  It does not use anything under oracle/.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "../../arpack-ng_b200/csrc/irl_nonsym.hpp"
#include "../../arpack-ng_b200/csrc/irl_sym.hpp"

extern "C" {
void scipy_dlarnv_(const int*, int*, const int*, double*);
void scipy_slarnv_(const int*, int*, const int*, float*);
}

namespace {

using namespace ab200;

typedef void (*allreduce_fn)(void* user, void* buf, int count, int is_double, int op);
typedef void (*op_fn)(const void* x, void* y, int n);

template <typename T>
struct HostVecOps final : VecOps<T> {
  int rank_ = 0, nranks_ = 1;
  allreduce_fn ar_ = nullptr;
  std::vector<T> mb_;
  unsigned long long calls = 0;

  T* alloc(size_t c) override { return (T*)std::calloc(c ? c : 1, sizeof(T)); }
  void release(T* p) override { std::free(p); }
  void upload(T* d, const T* s, size_t c) override { std::memcpy(d, s, sizeof(T) * c); }
  void download(T* d, const T* s, size_t c) override { std::memcpy(d, s, sizeof(T) * c); }
  void upload2d(T* d, size_t ldd, const T* s, size_t lds, size_t rows, size_t cols) override {
    for (size_t c = 0; c < cols; ++c) std::memcpy(d + c * ldd, s + c * lds, sizeof(T) * rows);
  }
  void download2d(T* d, size_t ldd, const T* s, size_t lds, size_t rows, size_t cols) override {
    upload2d(d, ldd, s, lds, rows, cols);
  }
  void sync() override {}
  bool is_device_pointer(const void*) override { return true; }
  T* mailbox(size_t c) override {
    mb_.assign(c, T(0));
    return mb_.data();
  }
  void fetch(T* h, const T* mb, size_t c) override { std::memcpy(h, mb, sizeof(T) * c); }
  void post(T* mb, const T* h, size_t c) override { std::memcpy(mb, h, sizeof(T) * c); }
  void allreduce_sum(T* mb, size_t c) override {
    if (ar_ && c) ar_(nullptr, mb, (int)c, sizeof(T) == 8, 0);
  }
  int rank() const override { return rank_; }
  int nranks() const override { return nranks_; }

  void copy(int64_t n, const T* x, T* y) override {
    if (x != y) std::memmove(y, x, sizeof(T) * (size_t)n);
  }
  void zero(int64_t n, T* x) override { std::memset(x, 0, sizeof(T) * (size_t)n); }
  void scal(int64_t n, T a, T* x) override {
    for (int64_t i = 0; i < n; ++i) x[i] *= a;
  }
  void axpby_norm(int64_t n, T a, T b, const T* x, T* y, T* out) override {
    T s = 0;
    for (int64_t i = 0; i < n; ++i) {
      T t = a * y[i];
      if (x) t += b * x[i];
      y[i] = t;
      s += t * t;
    }
    if (out) *out = s;
  }
  void dot(int64_t n, const T* x, const T* y, T* out) override {
    T s = 0;
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    *out = s;
  }
  bool absmax(int64_t n, const T* x, T* out) override {
    T m = 0;
    for (int64_t i = 0; i < n; ++i) m = std::max(m, std::fabs(x[i]));
    *out = m;
    return true;
  }
  void larnv_uniform_m1_1(int64_t n, int iseed[4], T* x) override;
  void start_step(int64_t n, T inv, const T* resid, T* vj, T* outx, T* bx, bool from_resid) override {
    if (halted()) return;
    for (int64_t i = 0; i < n; ++i) {
      const T t = resid[i] * inv;
      vj[i] = t;
      outx[i] = t;
      if (bx) bx[i] = from_resid ? t : bx[i] * inv;
    }
  }
  // device-resident sweep emulation: kernels "run" at once, so the sticky stop flag is simply read here
  T* stop_ = nullptr;
  unsigned long long gated_calls = 0;
  bool halted() const { return stop_ != nullptr && *stop_ != T(0); }
  bool deferred_ok() const override { return true; }
  void set_stop_flag(T* stop) override { stop_ = stop; }
  void start_step_gated(int64_t n, const StepGate<T>& g, const T* resid, T* vj, T* outx, T* bx) override {
    ++gated_calls;
    if (g.stop != nullptr && *g.stop != T(0)) return;
    const T wn = std::sqrt(g.A[g.prev_j]);
    T rn = std::sqrt(g.B[g.prev_j]);
    bool ok = true;
    if (!(rn > (T)0.717f * wn)) {
      const T rn1 = std::sqrt(g.C[0]);
      if (rn1 > (T)0.717f * rn) rn = rn1;
      else ok = false;
    }
    if (!(rn >= g.tiny) || !(rn > T(0))) ok = false;
    if (!ok) { *g.stop = g.stop_code; return; }
    const T inv = T(1) / rn;
    for (int64_t i = 0; i < n; ++i) {
      const T t = resid[i] * inv;
      vj[i] = t;
      outx[i] = t;
      if (bx) bx[i] = t;
    }
  }
  void ger(int64_t n, int k, const T* resid, const T* w, T* z, int64_t ldz) override {
    for (int c = 0; c < k; ++c)
      for (int64_t i = 0; i < n; ++i) z[i + c * ldz] += resid[i] * w[c];
  }
  void dots(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* out) override {
    calls++;
    for (int k = 0; k < j; ++k) {
      T s = 0;
      for (int64_t i = 0; i < n; ++i) s += v[i + k * ldv] * x[i];
      out[k] = s;
    }
    T s = 0;
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    out[j] = s;
  }
  void update(int64_t n, int j, const T* v, int64_t ldv, const T* coef, const T* src, T* dst, T* nrm2) override {
    calls++;
    T s = 0;
    for (int64_t i = 0; i < n; ++i) {
      T a = 0;
      for (int k = 0; k < j; ++k) a += v[i + k * ldv] * coef[k];
      const T d = src[i] - a;
      dst[i] = d;
      s += d * d;
    }
    if (nrm2) *nrm2 = s;
  }
  void orth_step(int64_t n, int j, const T* v, int64_t ldv, const T* w, T* resid, T* A, T* B, T* C) override {
    if (halted()) return;
    dots(n, j, v, ldv, w, w, A);
    allreduce_sum(A, (size_t)j + 1);
    update(n, j, v, ldv, A, w, resid, nullptr);
    dots(n, j, v, ldv, resid, resid, B);
    allreduce_sum(B, (size_t)j + 1);
    if (!(std::sqrt(B[j]) > (T)0.717f * std::sqrt(A[j]))) {
      update(n, j, v, ldv, B, resid, resid, C);
      C[1] = 1;
      allreduce_sum(C, 1);
    } else {
      C[1] = 0;
    }
  }
  void vq_core(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* q, int ldq, T* out, int64_t ldo,
               bool with_resid, T sigma, T beta, int beta_col, T* resid, T* nrm2) {
    std::vector<T> row((size_t)kin);
    T s = 0;
    for (int64_t i = 0; i < n; ++i) {
      for (int k = 0; k < kin; ++k) row[k] = v[i + k * ldv];
      T bval = 0;
      for (int c = 0; c < kout; ++c) {
        T a = 0;
        for (int k = 0; k < kin; ++k) a += row[k] * q[k + (size_t)c * ldq];
        out[i + c * ldo] = a;
        if (c == beta_col) bval = a;
      }
      if (with_resid) {
        const T t = sigma * resid[i] + (beta_col >= 0 ? beta * bval : T(0));
        resid[i] = t;
        s += t * t;
      }
    }
    if (with_resid && nrm2) *nrm2 = s;
  }
  void vq_update(int64_t n, int kin, int kout, T* v, int64_t ldv, const T* q, int ldq, bool with_resid, T sigma,
                 T beta, int beta_col, T* resid, T* nrm2) override {
    vq_core(n, kin, kout, v, ldv, q, ldq, v, ldv, with_resid, sigma, beta, beta_col, resid, nrm2);
  }
  void vq_out(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* m, int ldm, T* out,
              int64_t ldo) override {
    vq_core(n, kin, kout, v, ldv, m, ldm, out, ldo, false, 0, 0, -1, nullptr, nullptr);
  }
  void copy2d(int64_t n, int cols, const T* src, int64_t lds, T* dst, int64_t ldd) override {
    for (int c = 0; c < cols; ++c) std::memmove(dst + c * ldd, src + c * lds, sizeof(T) * (size_t)n);
  }
};
template <>
void HostVecOps<double>::larnv_uniform_m1_1(int64_t n, int iseed[4], double* x) {
  const int idist = 2, nn = (int)n;
  scipy_dlarnv_(&idist, iseed, &nn, x);
}
template <>
void HostVecOps<float>::larnv_uniform_m1_1(int64_t n, int iseed[4], float* x) {
  const int idist = 2, nn = (int)n;
  scipy_slarnv_(&idist, iseed, &nn, x);
}

template <typename T>
struct Proc {  // one "process": SAVE'd seed etc.
  HostVecOps<T> ops;
  SeedState seed;
  T smlnum_first = T(-1);
  bool par = false;
  std::unique_ptr<IrlSym<T>> sym;
  std::unique_ptr<IrlNonsym<T>> nonsym;
  // registered-operator mode (IrlBase::set_registered_op): y = OP x through a C callback, no ido = +-1 hand-offs
  bool defer = true;
  op_fn reg_op = nullptr;
  int reg_fused = 0;
  int reg_n = 0;

  template <typename Solver>
  void attach(Solver* s) {
    if (!reg_op) return;
    op_fn f = reg_op;
    const int n = reg_n;
    auto plain = [f, n](const T* x, T* y) { f((const void*)x, (void*)y, n); };
    if (!reg_fused) {
      s->set_registered_op(plain, nullptr);
      return;
    }
    // host emulation of the fused SpMV: vj = inv*resid, y = OP vj, dots {vj.y, y.y}
    HostVecOps<T>* o = &ops;
    s->set_registered_op(plain, [f, n, o](T inv, const StepGate<T>* g, const T* resid, T* vj, T* y, T* mb_dots) -> bool {
      if (o->halted()) return true;
      if (g != nullptr) {
        // the gate of the fused kernel: same tests as start_step_gated, on a scratch pass that writes nothing
        const T wn = std::sqrt(g->A[g->prev_j]);
        T rn = std::sqrt(g->B[g->prev_j]);
        bool ok = true;
        if (!(rn > (T)0.717f * wn)) {
          const T rn1 = std::sqrt(g->C[0]);
          if (rn1 > (T)0.717f * rn) rn = rn1;
          else ok = false;
        }
        if (!(rn >= g->tiny) || !(rn > T(0))) ok = false;
        if (!ok) { *g->stop = g->stop_code; return true; }
        inv = T(1) / rn;
      }
      for (int i = 0; i < n; ++i) vj[i] = inv * resid[i];
      f((const void*)vj, (void*)y, n);
      T a = 0, b = 0;
      for (int i = 0; i < n; ++i) { a += vj[i] * y[i]; b += y[i] * y[i]; }
      mb_dots[0] = a;
      mb_dots[1] = b;
      return true;
    });
  }
};

}  // namespace

extern "C" {

void* hd_new(int is_double) { return is_double ? (void*)new Proc<double>() : (void*)new Proc<float>(); }
void hd_free(void* p, int is_double) {
  if (is_double) delete (Proc<double>*)p;
  else delete (Proc<float>*)p;
}
void hd_set_comm(void* p, int is_double, int rank, int nranks, allreduce_fn fn) {
  if (is_double) { auto* q = (Proc<double>*)p; q->par = true; q->ops.rank_ = rank; q->ops.nranks_ = nranks; q->ops.ar_ = fn; }
  else { auto* q = (Proc<float>*)p; q->par = true; q->ops.rank_ = rank; q->ops.nranks_ = nranks; q->ops.ar_ = fn; }
}
void hd_set_registered_op(void* p, int is_double, op_fn fn, int n, int fused) {
  if (is_double) { auto* q = (Proc<double>*)p; q->reg_op = fn; q->reg_n = n; q->reg_fused = fused; }
  else { auto* q = (Proc<float>*)p; q->reg_op = fn; q->reg_n = n; q->reg_fused = fused; }
}
double hd_fused_dot_maxdiff(void* p, int is_double, int fam_sym) {
  if (is_double) { auto* q = (Proc<double>*)p; return fam_sym ? q->sym->fused_dot_maxdiff : q->nonsym->fused_dot_maxdiff; }
  auto* q = (Proc<float>*)p;
  return fam_sym ? q->sym->fused_dot_maxdiff : q->nonsym->fused_dot_maxdiff;
}
// out3 = {steps that ran inside device-resident batches, batches cut short by a rare path, host round trips}
void hd_deferred_stats(void* p, int is_double, int fam_sym, long long* out3) {
#define HD_DS(Q) do { if (fam_sym) { out3[0] = Q->sym->deferred_steps(); out3[1] = Q->sym->deferred_trips(); out3[2] = Q->sym->host_round_trips(); } \
                      else { out3[0] = Q->nonsym->deferred_steps(); out3[1] = Q->nonsym->deferred_trips(); out3[2] = Q->nonsym->host_round_trips(); } } while (0)
  if (is_double) { auto* q = (Proc<double>*)p; HD_DS(q); }
  else { auto* q = (Proc<float>*)p; HD_DS(q); }
#undef HD_DS
}
// deferral needs hand-off slots the "device" reaches by itself: the double's arrays always qualify
void hd_set_deferral(void* p, int is_double, int on) {
  if (is_double) ((Proc<double>*)p)->defer = on != 0;
  else ((Proc<float>*)p)->defer = on != 0;
}
// COMMON /debug/ of the control code under test (what debug_c does in the product, api.cu)
void hd_debug(const int* levels24) { std::memcpy(static_cast<void*>(&trace_levels()), levels24, sizeof(TraceLevels)); }
void hd_stats(void* p, int is_double, int fam_sym, int* out5) {
  const Counters* c = nullptr;
  if (is_double) { auto* q = (Proc<double>*)p; c = fam_sym ? &q->sym->counters() : &q->nonsym->counters(); }
  else { auto* q = (Proc<float>*)p; c = fam_sym ? &q->sym->counters() : &q->nonsym->counters(); }
  out5[0] = c->nopx; out5[1] = c->nbx; out5[2] = c->nrorth; out5[3] = c->nitref; out5[4] = c->nrstrt;
}

#define HD_AUPD(NAME, T, ISSYM)                                                                                  \
  void NAME(void* p, int* ido, const char* bmat, int n, const char* which, int nev, T* tol, T* resid, int ncv,   \
            T* v, int ldv, int* iparam, int* ipntr, T* workd, T* workl, int lworkl, int* info) {                 \
    auto* q = (Proc<T>*)p;                                                                                       \
    if (*ido == 0) {                                                                                             \
      if (ISSYM) q->sym.reset(new IrlSym<T>(&q->ops, q->par, &q->seed));                                         \
      else q->nonsym.reset(new IrlNonsym<T>(&q->ops, q->par, &q->seed, &q->smlnum_first));                       \
      if (ISSYM) q->attach(q->sym.get());                                                                        \
      else q->attach(q->nonsym.get());                                                                           \
      if (ISSYM) q->sym->set_deferral(q->defer);                                                                 \
      else q->nonsym->set_deferral(q->defer);                                                                    \
    }                                                                                                            \
    if (ISSYM) q->sym->aupd(ido, bmat[0], n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl,   \
                            lworkl, info);                                                                       \
    else q->nonsym->aupd(ido, bmat[0], n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl,      \
                         lworkl, info);                                                                          \
  }
HD_AUPD(hd_dsaupd, double, 1)
HD_AUPD(hd_ssaupd, float, 1)
HD_AUPD(hd_dnaupd, double, 0)
HD_AUPD(hd_snaupd, float, 0)

#define HD_SEUPD(NAME, T)                                                                                        \
  void NAME(void* p, int rvec, const char* howmny, int* select, T* d, T* z, int ldz, T sigma, const char* bmat,  \
            int n, const char* which, int nev, T tol, T* resid, int ncv, T* v, int ldv, int* iparam, int* ipntr, \
            T* workd, T* workl, int lworkl, int* info) {                                                         \
    auto* q = (Proc<T>*)p;                                                                                       \
    if (!q->sym) q->sym.reset(new IrlSym<T>(&q->ops, q->par, &q->seed));                                         \
    q->sym->ensure_mailbox(ncv);                                                                                 \
    q->sym->eupd(rvec != 0, howmny[0], select, d, z, ldz, sigma, bmat[0], n, which, nev, tol, resid, ncv, v,     \
                 ldv, iparam, ipntr, workd, workl, lworkl, info);                                                \
  }
HD_SEUPD(hd_dseupd, double)
HD_SEUPD(hd_sseupd, float)

#define HD_NEUPD(NAME, T)                                                                                        \
  void NAME(void* p, int rvec, const char* howmny, int* select, T* dr, T* di, T* z, int ldz, T sigmar, T sigmai, \
            T* workev, const char* bmat, int n, const char* which, int nev, T tol, T* resid, int ncv, T* v,      \
            int ldv, int* iparam, int* ipntr, T* workd, T* workl, int lworkl, int* info) {                       \
    auto* q = (Proc<T>*)p;                                                                                       \
    if (!q->nonsym) q->nonsym.reset(new IrlNonsym<T>(&q->ops, q->par, &q->seed, &q->smlnum_first));              \
    q->nonsym->ensure_mailbox(ncv);                                                                              \
    q->nonsym->eupd(rvec != 0, howmny[0], select, dr, di, z, ldz, sigmar, sigmai, workev, bmat[0], n, which,     \
                    nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);                    \
  }
HD_NEUPD(hd_dneupd, double)
HD_NEUPD(hd_sneupd, float)
}
