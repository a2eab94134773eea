// tests/hostdouble/hostdouble_cplx.cpp -- TEST DOUBLE, never part of the product (see hostdouble.cpp).
//
// Plain-loop VecOps<std::complex<R>> so that the CPU test-suite can drive the product's complex host control code
// (arpack-ng_b200/csrc/irl_complex.hpp) without a GPU.  Uses nothing under oracle/.
#include <algorithm>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../../arpack-ng_b200/csrc/irl_complex.hpp"

extern "C" {
void scipy_zlarnv_(const int*, int*, const int*, std::complex<double>*);
void scipy_clarnv_(const int*, int*, const int*, std::complex<float>*);
}

namespace {

using namespace ab200;

typedef void (*allreduce_fn)(void* user, void* buf, int count, int is_double, int op);

template <typename R>
struct HostVecOpsZ final : VecOps<std::complex<R>> {
  using T = std::complex<R>;
  std::vector<T> mb_;
  int rank_ = 0, nranks_ = 1;
  allreduce_fn ar_ = nullptr;

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
  void allreduce_sum(T* mb, size_t c) override {  // complex sum == sum of 2c reals
    if (ar_ && c) ar_(nullptr, mb, (int)(2 * c), sizeof(R) == 8, 0);
  }
  int rank() const override { return rank_; }
  int nranks() const override { return nranks_; }

  void copy(int64_t n, const T* x, T* y) override {
    if (x != y) std::memmove(y, x, sizeof(T) * (size_t)n);
  }
  void zero(int64_t n, T* x) override { std::fill(x, x + n, T(0)); }
  void scal(int64_t n, T a, T* x) override {
    for (int64_t i = 0; i < n; ++i) x[i] *= a;
  }
  void axpby_norm(int64_t n, T a, T b, const T* x, T* y, T* out) override {
    R s = 0;
    for (int64_t i = 0; i < n; ++i) {
      T t = a * y[i];
      if (x) t += b * x[i];
      y[i] = t;
      s += std::norm(t);
    }
    if (out) *out = T(s);
  }
  void dot(int64_t n, const T* x, const T* y, T* out) override {  // sum conj(x) y
    T s = 0;
    for (int64_t i = 0; i < n; ++i) s += std::conj(x[i]) * y[i];
    *out = s;
  }
  void larnv_uniform_m1_1(int64_t n, int iseed[4], T* x) override;
  void start_step(int64_t n, T inv, const T* resid, T* vj, T* outx, T* bx, bool from_resid) override {
    const R s = inv.real();
    for (int64_t i = 0; i < n; ++i) {
      const T t = resid[i] * s;
      vj[i] = t;
      outx[i] = t;
      if (bx) bx[i] = from_resid ? t : bx[i] * s;
    }
  }
  void ger(int64_t n, int k, const T* resid, const T* w, T* z, int64_t ldz) override {  // zgeru: no conjugate
    for (int c = 0; c < k; ++c)
      for (int64_t i = 0; i < n; ++i) z[i + c * ldz] += resid[i] * w[c];
  }
  void dots(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* out) override {  // V^H x
    for (int k = 0; k < j; ++k) {
      T s = 0;
      for (int64_t i = 0; i < n; ++i) s += std::conj(v[i + k * ldv]) * x[i];
      out[k] = s;
    }
    T s = 0;
    for (int64_t i = 0; i < n; ++i) s += std::conj(x[i]) * y[i];
    out[j] = s;
  }
  void update(int64_t n, int j, const T* v, int64_t ldv, const T* coef, const T* src, T* dst, T* nrm2) override {
    R s = 0;
    for (int64_t i = 0; i < n; ++i) {
      T a = 0;
      for (int k = 0; k < j; ++k) a += v[i + k * ldv] * coef[k];
      const T d = src[i] - a;
      dst[i] = d;
      s += std::norm(d);
    }
    if (nrm2) *nrm2 = T(s);
  }
  void orth_step(int64_t, int, const T*, int64_t, const T*, T*, T*, T*, T*) override { std::abort(); }
  void vq_core(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* q, int ldq, T* out, int64_t ldo,
               bool with_resid, T sigma, T beta, int beta_col, T* resid, T* nrm2) {
    std::vector<T> row((size_t)kin);
    R s = 0;
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
        s += std::norm(t);
      }
    }
    if (with_resid && nrm2) *nrm2 = T(s);
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
void HostVecOpsZ<double>::larnv_uniform_m1_1(int64_t n, int iseed[4], std::complex<double>* x) {
  const int idist = 2, nn = (int)n;
  scipy_zlarnv_(&idist, iseed, &nn, x);
}
template <>
void HostVecOpsZ<float>::larnv_uniform_m1_1(int64_t n, int iseed[4], std::complex<float>* x) {
  const int idist = 2, nn = (int)n;
  scipy_clarnv_(&idist, iseed, &nn, x);
}

template <typename R>
struct ProcZ {
  HostVecOpsZ<R> ops;
  SeedState seed;
  R smlnum_first = R(-1);
  bool par = false;
  std::unique_ptr<IrlComplex<R>> slv;
};

}  // namespace

extern "C" {

void* hdz_new(int is_double) { return is_double ? (void*)new ProcZ<double>() : (void*)new ProcZ<float>(); }
void hdz_free(void* p, int is_double) {
  if (is_double) delete (ProcZ<double>*)p;
  else delete (ProcZ<float>*)p;
}
void hdz_set_comm(void* p, int is_double, int rank, int nranks, allreduce_fn fn) {
  if (is_double) { auto* q = (ProcZ<double>*)p; q->par = true; q->ops.rank_ = rank; q->ops.nranks_ = nranks; q->ops.ar_ = fn; }
  else { auto* q = (ProcZ<float>*)p; q->par = true; q->ops.rank_ = rank; q->ops.nranks_ = nranks; q->ops.ar_ = fn; }
}
void hdz_stats(void* p, int is_double, int* out5) {
  const Counters* c = is_double ? &((ProcZ<double>*)p)->slv->counters() : &((ProcZ<float>*)p)->slv->counters();
  out5[0] = c->nopx; out5[1] = c->nbx; out5[2] = c->nrorth; out5[3] = c->nitref; out5[4] = c->nrstrt;
}

#define HDZ_AUPD(NAME, R)                                                                                          \
  void NAME(void* p, int* ido, const char* bmat, int n, const char* which, int nev, R* tol, void* resid, int ncv,  \
            void* v, int ldv, int* iparam, int* ipntr, void* workd, void* workl, int lworkl, R* rwork, int* info) { \
    using Z = std::complex<R>;                                                                                     \
    auto* q = (ProcZ<R>*)p;                                                                                        \
    if (*ido == 0) q->slv.reset(new IrlComplex<R>(&q->ops, &q->seed, &q->smlnum_first, q->par));                           \
    q->slv->aupd(ido, bmat[0], n, which, nev, tol, (Z*)resid, ncv, (Z*)v, ldv, iparam, ipntr, (Z*)workd,           \
                 (Z*)workl, lworkl, rwork, info);                                                                  \
  }
HDZ_AUPD(hd_znaupd, double)
HDZ_AUPD(hd_cnaupd, float)

#define HDZ_EUPD(NAME, R)                                                                                          \
  void NAME(void* p, int rvec, const char* howmny, int* select, void* d, void* z, int ldz, R sre, R sim,           \
            void* workev, const char* bmat, int n, const char* which, int nev, R tol, void* resid, int ncv,        \
            void* v, int ldv, int* iparam, int* ipntr, void* workd, void* workl, int lworkl, R* rwork,             \
            int* info) {                                                                                           \
    using Z = std::complex<R>;                                                                                     \
    auto* q = (ProcZ<R>*)p;                                                                                        \
    if (!q->slv) q->slv.reset(new IrlComplex<R>(&q->ops, &q->seed, &q->smlnum_first, q->par));                             \
    q->slv->ensure_mailbox(ncv);                                                                                   \
    q->slv->eupd(rvec != 0, howmny[0], select, (Z*)d, (Z*)z, ldz, Z(sre, sim), (Z*)workev, bmat[0], n, which,      \
                 nev, tol, (Z*)resid, ncv, (Z*)v, ldv, iparam, ipntr, (Z*)workd, (Z*)workl, lworkl, rwork, info);  \
  }
HDZ_EUPD(hd_zneupd_ri, double)
HDZ_EUPD(hd_cneupd_ri, float)
}
