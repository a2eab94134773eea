// hostmath.hpp -- ncv-sized host arithmetic of the projected problem.
//
// north_star keeps the O(ncv^2) control work on the host "in Fortran/LAPACK".  The image has no
// Fortran compiler, so the control code is C++ and calls the LAPACK that is present (OpenBLAS as
// bundled by SciPy, symbols scipy_<name>_), i.e. the same routines the reference calls from
// SRC/dseigt.f, dsapps.f, dneigh.f, dnapps.f, dseupd.f, dneupd.f.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <vector>

extern "C" {
#define AB200_DECL(P, R)                                                                                     \
  R scipy_##P##lamch_(const char*, size_t);                                                                  \
  void scipy_##P##lartg_(const R*, const R*, R*, R*, R*);                                                    \
  void scipy_##P##steqr_(const char*, const int*, R*, R*, R*, const int*, R*, int*, size_t);                 \
  void scipy_##P##lahqr_(const int*, const int*, const int*, const int*, const int*, R*, const int*, R*, R*, \
                         const int*, const int*, R*, const int*, int*);                                      \
  void scipy_##P##trevc_(const char*, const char*, int*, const int*, const R*, const int*, R*, const int*,   \
                         R*, const int*, const int*, int*, R*, int*, size_t, size_t);                        \
  void scipy_##P##trsen_(const char*, const char*, const int*, const int*, R*, const int*, R*, const int*,   \
                         R*, R*, int*, R*, R*, R*, const int*, int*, const int*, int*, size_t, size_t);      \
  void scipy_##P##geqr2_(const int*, const int*, R*, const int*, R*, R*, int*);                              \
  void scipy_##P##orm2r_(const char*, const char*, const int*, const int*, const int*, const R*, const int*, \
                         const R*, R*, const int*, R*, int*, size_t, size_t);                                \
  void scipy_##P##larfg_(const int*, R*, R*, const int*, R*);                                                \
  void scipy_##P##larf_(const char*, const int*, const int*, const R*, const int*, const R*, R*, const int*, \
                        R*, size_t);                                                                         \
  R scipy_##P##lanhs_(const char*, const int*, const R*, const int*, R*, size_t);                            \
  R scipy_##P##lapy2_(const R*, const R*);                                                                   \
  void scipy_##P##labad_(R*, R*);                                                                            \
  R scipy_##P##nrm2_(const int*, const R*, const int*);                                                      \
  void scipy_##P##gemv_(const char*, const int*, const int*, const R*, const R*, const int*, const R*,       \
                        const int*, const R*, R*, const int*, size_t);                                       \
  void scipy_##P##trmm_(const char*, const char*, const char*, const char*, const int*, const int*, const R*, \
                        const R*, const int*, R*, const int*, size_t, size_t, size_t, size_t);
AB200_DECL(d, double)
AB200_DECL(s, float)
#undef AB200_DECL
}

namespace ab200 {

template <typename T>
struct Lapack;

#define AB200_LAPACK(P, R)                                                                                    \
  template <>                                                                                                 \
  struct Lapack<R> {                                                                                          \
    static R lamch(const char* c) { return scipy_##P##lamch_(c, 1); }                                         \
    static void lartg(R f, R g, R& c, R& s, R& r) { scipy_##P##lartg_(&f, &g, &c, &s, &r); }                  \
    static int steqr_I(int n, R* d, R* e, R* z, int ldz, R* work) {                                           \
      int info = 0;                                                                                           \
      scipy_##P##steqr_("I", &n, d, e, z, &ldz, work, &info, 1);                                              \
      return info;                                                                                            \
    }                                                                                                         \
    static int lahqr(bool wantt, bool wantz, int n, int ilo, int ihi, R* h, int ldh, R* wr, R* wi, int iloz,  \
                     int ihiz, R* z, int ldz) {                                                               \
      int info = 0, wt = wantt, wz = wantz;                                                                   \
      scipy_##P##lahqr_(&wt, &wz, &n, &ilo, &ihi, h, &ldh, wr, wi, &iloz, &ihiz, z, &ldz, &info);             \
      return info;                                                                                            \
    }                                                                                                         \
    static int trevc(const char* side, const char* howmny, int* select, int n, const R* t, int ldt, R* vl,    \
                     int ldvl, R* vr, int ldvr, int mm, int* m, R* work) {                                    \
      int info = 0;                                                                                           \
      scipy_##P##trevc_(side, howmny, select, &n, t, &ldt, vl, &ldvl, vr, &ldvr, &mm, m, work, &info, 1, 1);  \
      return info;                                                                                            \
    }                                                                                                         \
    static int trsen_NV(const int* select, int n, R* t, int ldt, R* q, int ldq, R* wr, R* wi, int* m,         \
                        R* work, int lwork) {                                                                 \
      int info = 0, iwork[1], liwork = 1;                                                                     \
      R s, sep;                                                                                               \
      scipy_##P##trsen_("N", "V", select, &n, t, &ldt, q, &ldq, wr, wi, m, &s, &sep, work, &lwork, iwork,     \
                        &liwork, &info, 1, 1);                                                                \
      return info;                                                                                            \
    }                                                                                                         \
    static int geqr2(int m, int n, R* a, int lda, R* tau, R* work) {                                          \
      int info = 0;                                                                                           \
      scipy_##P##geqr2_(&m, &n, a, &lda, tau, work, &info);                                                   \
      return info;                                                                                            \
    }                                                                                                         \
    static int orm2r(const char* side, const char* trans, int m, int n, int k, const R* a, int lda,           \
                     const R* tau, R* c, int ldc, R* work) {                                                  \
      int info = 0;                                                                                           \
      scipy_##P##orm2r_(side, trans, &m, &n, &k, a, &lda, tau, c, &ldc, work, &info, 1, 1);                   \
      return info;                                                                                            \
    }                                                                                                         \
    static void larfg(int n, R& alpha, R* x, int incx, R& tau) { scipy_##P##larfg_(&n, &alpha, x, &incx, &tau); } \
    static void larf(const char* side, int m, int n, const R* v, int incv, R tau, R* c, int ldc, R* work) {   \
      scipy_##P##larf_(side, &m, &n, v, &incv, &tau, c, &ldc, work, 1);                                       \
    }                                                                                                         \
    static R lanhs1(int n, const R* a, int lda, R* work) { return scipy_##P##lanhs_("1", &n, a, &lda, work, 1); } \
    static R lapy2(R x, R y) { return scipy_##P##lapy2_(&x, &y); }                                            \
    static void labad(R& small_, R& large_) { scipy_##P##labad_(&small_, &large_); }                          \
    static R nrm2(int n, const R* x, int incx) { return scipy_##P##nrm2_(&n, x, &incx); }                     \
    static void gemvT(int m, int n, const R* a, int lda, const R* x, R* y) {                                  \
      const R one = 1, zero = 0;                                                                              \
      const int i1 = 1;                                                                                       \
      scipy_##P##gemv_("T", &m, &n, &one, a, &lda, x, &i1, &zero, y, &i1, 1);                                 \
    }                                                                                                         \
    static void trmm_RUNN(int m, int n, const R* a, int lda, R* b, int ldb) {                                 \
      const R one = 1;                                                                                        \
      scipy_##P##trmm_("R", "U", "N", "N", &m, &n, &one, a, &lda, b, &ldb, 1, 1, 1, 1);                       \
    }                                                                                                         \
  };
AB200_LAPACK(d, double)
AB200_LAPACK(s, float)
#undef AB200_LAPACK

// The reference's DGKS constant is the default-REAL literal 0.717 (SRC/dsaitr.f:656, dgetv0.f:375).
template <typename T>
inline T dgks_threshold() {
  return (T)0.717f;
}

template <typename T>
inline T eps23_of(T eps, bool parpack) {
  // dsaup2.f:272-273 uses a DOUBLE exponent, pdsaup2.f:294 a default-REAL one
  if (sizeof(T) == 8) return (T)std::pow((double)eps, parpack ? (double)(2.0f / 3.0f) : 2.0 / 3.0);
  return (T)std::pow((float)eps, 2.0f / 3.0f);
}

// Ordering predicates of the reference's Shell sorts.  The sorts themselves (gap n/2, n/4, ...,
// insertion by swaps; not stable) must be reproduced exactly because tie order decides which of two
// equal Ritz values is kept (SRC/dsortr.f:59-218, dsesrt.f:68-217, dsortc.f:66-344).
enum class Key { LM, SM, LA, SA, LR, SR, LI, SI, NONE };
inline Key key_of(const char* w) {
  const char a = w[0], b = w[1];
  if (a == 'L' && b == 'M') return Key::LM;
  if (a == 'S' && b == 'M') return Key::SM;
  if (a == 'L' && b == 'A') return Key::LA;
  if (a == 'S' && b == 'A') return Key::SA;
  if (a == 'L' && b == 'R') return Key::LR;
  if (a == 'S' && b == 'R') return Key::SR;
  if (a == 'L' && b == 'I') return Key::LI;
  if (a == 'S' && b == 'I') return Key::SI;
  return Key::NONE;
}

// Generic gapped insertion sort: out_of_order(j, j+gap) decides, swap_fn(j, j+gap) exchanges.
template <typename Pred, typename Swap>
inline void shell_sort(int n, Pred out_of_order, Swap swap_fn) {
  for (int gap = n / 2; gap > 0; gap /= 2)
    for (int i = gap; i < n; ++i)
      for (int j = i - gap; j >= 0 && out_of_order(j, j + gap); j -= gap) swap_fn(j, j + gap);
}

// dsortr: real keys (LM/SM by magnitude, LA/SA algebraic); "which" names the end that sorts LAST
template <typename T>
inline void sort_real(Key k, int n, T* x1, T* x2 /*may be null*/) {
  auto ooo = [&](int a, int b) -> bool {
    switch (k) {
      case Key::SA: return x1[a] < x1[b];
      case Key::SM: return std::fabs(x1[a]) < std::fabs(x1[b]);
      case Key::LA: return x1[a] > x1[b];
      case Key::LM: return std::fabs(x1[a]) > std::fabs(x1[b]);
      default: return false;
    }
  };
  auto sw = [&](int a, int b) {
    T t = x1[a]; x1[a] = x1[b]; x1[b] = t;
    if (x2) { t = x2[a]; x2[a] = x2[b]; x2[b] = t; }
  };
  shell_sort(n, ooo, sw);
}
// dsesrt: as sort_real, the columns of a(lda, n) (na rows) follow x
template <typename T>
inline void sort_real_cols(Key k, int n, T* x, int na, T* a, int lda) {
  auto ooo = [&](int p, int q) -> bool {
    switch (k) {
      case Key::SA: return x[p] < x[q];
      case Key::SM: return std::fabs(x[p]) < std::fabs(x[q]);
      case Key::LA: return x[p] > x[q];
      case Key::LM: return std::fabs(x[p]) > std::fabs(x[q]);
      default: return false;
    }
  };
  auto sw = [&](int p, int q) {
    T t = x[p]; x[p] = x[q]; x[q] = t;
    for (int r = 0; r < na; ++r) {
      T u = a[(size_t)p * lda + r]; a[(size_t)p * lda + r] = a[(size_t)q * lda + r]; a[(size_t)q * lda + r] = u;
    }
  };
  shell_sort(n, ooo, sw);
}
// dsortc: complex keys held as (xr, xi); y follows
template <typename T>
inline void sort_cplx(Key k, int n, T* xr, T* xi, T* y /*may be null*/) {
  auto mag = [&](int i) { return Lapack<T>::lapy2(xr[i], xi[i]); };
  auto ooo = [&](int a, int b) -> bool {
    switch (k) {
      case Key::LM: return mag(a) > mag(b);
      case Key::SM: return mag(a) < mag(b);
      case Key::LR: return xr[a] > xr[b];
      case Key::SR: return xr[a] < xr[b];
      case Key::LI: return std::fabs(xi[a]) > std::fabs(xi[b]);
      case Key::SI: return std::fabs(xi[a]) < std::fabs(xi[b]);
      default: return false;
    }
  };
  auto sw = [&](int a, int b) {
    T t = xr[a]; xr[a] = xr[b]; xr[b] = t;
    t = xi[a]; xi[a] = xi[b]; xi[b] = t;
    if (y) { t = y[a]; y[a] = y[b]; y[b] = t; }
  };
  shell_sort(n, ooo, sw);
}

}  // namespace ab200
