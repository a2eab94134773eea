// hostmath_cplx.hpp -- ncv-sized host arithmetic of the complex projected problem.
//
// Complex LAPACK (OpenBLAS as bundled by SciPy, symbols scipy_z*/scipy_c*): the routines the reference calls from
// SRC/zneigh.f (zlahqr, ztrevc), znapps.f (zlartg, zlanhs), znaitr.f (zlanhs), zneupd.f (zlahqr, ztrsen, zgeqr2,
// zunm2r, ztrevc, ztrmm).  std::complex<R> is layout-compatible with Fortran COMPLEX*16 / COMPLEX.
#pragma once
#include <complex>

#include "hostmath.hpp"

extern "C" {
#define AB200_DECLZ(P, RP, R, Z)                                                                                  \
  void scipy_##P##lartg_(const Z*, const Z*, R*, Z*, Z*);                                                         \
  void scipy_##P##lahqr_(const int*, const int*, const int*, const int*, const int*, Z*, const int*, Z*,          \
                         const int*, const int*, Z*, const int*, int*);                                           \
  void scipy_##P##trevc_(const char*, const char*, int*, const int*, Z*, const int*, Z*, const int*, Z*,          \
                         const int*, const int*, int*, Z*, R*, int*, size_t, size_t);                             \
  void scipy_##P##trsen_(const char*, const char*, const int*, const int*, Z*, const int*, Z*, const int*, Z*,    \
                         int*, R*, R*, Z*, const int*, int*, size_t, size_t);                                     \
  void scipy_##P##geqr2_(const int*, const int*, Z*, const int*, Z*, Z*, int*);                                   \
  void scipy_##P##unm2r_(const char*, const char*, const int*, const int*, const int*, const Z*, const int*,      \
                         const Z*, Z*, const int*, Z*, int*, size_t, size_t);                                     \
  R scipy_##P##lanhs_(const char*, const int*, const Z*, const int*, R*, size_t);                                 \
  R scipy_##RP##nrm2_(const int*, const Z*, const int*);
AB200_DECLZ(z, dz, double, std::complex<double>)
AB200_DECLZ(c, sc, float, std::complex<float>)
#undef AB200_DECLZ
}

namespace ab200 {

template <typename R>
struct LapackZ;

#define AB200_LAPACKZ(P, RP, R)                                                                                   \
  template <>                                                                                                     \
  struct LapackZ<R> {                                                                                             \
    using Z = std::complex<R>;                                                                                    \
    static void lartg(Z f, Z g, R& c, Z& s, Z& r) { scipy_##P##lartg_(&f, &g, &c, &s, &r); }                      \
    static int lahqr(bool wantt, bool wantz, int n, int ilo, int ihi, Z* h, int ldh, Z* w, int iloz, int ihiz,    \
                     Z* z, int ldz) {                                                                             \
      int info = 0, wt = wantt, wz = wantz;                                                                       \
      scipy_##P##lahqr_(&wt, &wz, &n, &ilo, &ihi, h, &ldh, w, &iloz, &ihiz, z, &ldz, &info);                      \
      return info;                                                                                                \
    }                                                                                                             \
    static int trevc(const char* side, const char* howmny, int* select, int n, Z* t, int ldt, Z* vl, int ldvl,    \
                     Z* vr, int ldvr, int mm, int* m, Z* work, R* rwork) {                                        \
      int info = 0;                                                                                               \
      scipy_##P##trevc_(side, howmny, select, &n, t, &ldt, vl, &ldvl, vr, &ldvr, &mm, m, work, rwork, &info, 1,   \
                        1);                                                                                       \
      return info;                                                                                                \
    }                                                                                                             \
    static int trsen_NV(const int* select, int n, Z* t, int ldt, Z* q, int ldq, Z* w, int* m, Z* work,            \
                        int lwork) {                                                                              \
      int info = 0;                                                                                               \
      R s, sep;                                                                                                   \
      scipy_##P##trsen_("N", "V", select, &n, t, &ldt, q, &ldq, w, m, &s, &sep, work, &lwork, &info, 1, 1);       \
      return info;                                                                                                \
    }                                                                                                             \
    static int geqr2(int m, int n, Z* a, int lda, Z* tau, Z* work) {                                              \
      int info = 0;                                                                                               \
      scipy_##P##geqr2_(&m, &n, a, &lda, tau, work, &info);                                                       \
      return info;                                                                                                \
    }                                                                                                             \
    static int unm2r(const char* side, const char* trans, int m, int n, int k, const Z* a, int lda,               \
                     const Z* tau, Z* c, int ldc, Z* work) {                                                      \
      int info = 0;                                                                                               \
      scipy_##P##unm2r_(side, trans, &m, &n, &k, a, &lda, tau, c, &ldc, work, &info, 1, 1);                       \
      return info;                                                                                                \
    }                                                                                                             \
    static R lanhs1(int n, const Z* a, int lda, R* work) { return scipy_##P##lanhs_("1", &n, a, &lda, work, 1); } \
    static R nrm2(int n, const Z* x, int incx) { return scipy_##RP##nrm2_(&n, x, &incx); }                        \
    /* dlapy2(dble(z), aimag(z)) -- the modulus as the reference forms it */                                      \
    static R abs(Z z) { return Lapack<R>::lapy2(z.real(), z.imag()); }                                            \
  };
AB200_LAPACKZ(z, dz, double)
AB200_LAPACKZ(c, sc, float)
#undef AB200_LAPACKZ

// zsortc (SRC/zsortc.f:1-322): Shell sort of complex x by modulus / real part / imaginary part; y follows
template <typename R>
inline void sort_z(Key k, int n, std::complex<R>* x, std::complex<R>* y /*may be null*/) {
  using Z = std::complex<R>;
  auto ooo = [&](int a, int b) -> bool {
    switch (k) {
      case Key::LM: return LapackZ<R>::abs(x[a]) > LapackZ<R>::abs(x[b]);
      case Key::SM: return LapackZ<R>::abs(x[a]) < LapackZ<R>::abs(x[b]);
      case Key::LR: return x[a].real() > x[b].real();
      case Key::SR: return x[a].real() < x[b].real();
      case Key::LI: return x[a].imag() > x[b].imag();
      case Key::SI: return x[a].imag() < x[b].imag();
      default: return false;
    }
  };
  auto sw = [&](int a, int b) {
    Z t = x[a]; x[a] = x[b]; x[b] = t;
    if (y) { t = y[a]; y[a] = y[b]; y[b] = t; }
  };
  shell_sort(n, ooo, sw);
}

}  // namespace ab200
