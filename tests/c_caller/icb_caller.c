/* icb_caller.c -- a plain C program written against the ICB prototypes (include/arpack_b200.h declares them exactly
 * as ICB/arpack.h:12-21 does) with HOST arrays and a CPU operator: what an existing arpack-ng user has.  Linked
 * against libarpack_b200.so it must give the reference's answers without any source change (the drop-in claim of
 * INTEGRATION.md section 1), through the bind(c) names, the gfortran-ABI names and the stat/debug accessors.
 *
 * exit code: 0 all checks passed; 3 the library reported "no CUDA device" (info = -9990); 1 a check failed.      */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arpack_b200.h"

static int fail(const char* what, double got, double want) {
  fprintf(stderr, "icb_caller: %s: got %.15g, want %.15g\n", what, got, want);
  return 1;
}

/* y = A x for the nx x nx 5-point Laplacian scaled by (nx+1)^2 (the operator of EXAMPLES/SIMPLE/dssimp.f:484-540) */
static void lap2d(int nx, const double* x, double* y) {
  const double h2 = (double)(nx + 1) * (nx + 1);
  for (int j = 0; j < nx; ++j)
    for (int i = 0; i < nx; ++i) {
      const int k = j * nx + i;
      double s = 4.0 * x[k];
      if (i > 0) s -= x[k - 1];
      if (i < nx - 1) s -= x[k + 1];
      if (j > 0) s -= x[k - nx];
      if (j < nx - 1) s -= x[k + nx];
      y[k] = h2 * s;
    }
}

static int cmp_double(const void* a, const void* b) {
  const double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}

/* 1. dsaupd_c / dseupd_c, BASELINE config 1: nx = 10, nev = 4, ncv = 20, 'LM', tol = 0 */
static int sym_bind_c(void) {
  enum { NX = 10, N = NX * NX, NEV = 4, NCV = 20, LWORKL = NCV * NCV + 8 * NCV };
  static double resid[N], v[N * NCV], workd[3 * N], workl[LWORKL], d[NEV], z[N * NEV];
  a_int iparam[11] = {0}, ipntr[11] = {0}, select[NCV];
  a_int ido = 0, info = 0;
  iparam[0] = 1; iparam[2] = 300; iparam[3] = 1; iparam[6] = 1;
  int handoffs = 0;
  for (;;) {
    dsaupd_c(&ido, "I", N, "LM", NEV, 0.0, resid, NCV, v, N, iparam, ipntr, workd, workl, LWORKL, &info);
    if (ido == -1 || ido == 1) { lap2d(NX, workd + ipntr[0] - 1, workd + ipntr[1] - 1); ++handoffs; }
    else break;
  }
  if (info == -9990) return 3;
  if (info != 0 || ido != 99) return fail("dsaupd_c info", info, 0);
  if (iparam[4] != NEV) return fail("dsaupd_c nconv", iparam[4], NEV);
  if (iparam[8] != handoffs) return fail("dsaupd_c nopx", iparam[8], handoffs);
  dseupd_c(1, "A", select, d, z, N, 0.0, "I", N, "LM", NEV, 0.0, resid, NCV, v, N, iparam, ipntr, workd, workl, LWORKL,
           &info);
  if (info != 0) return fail("dseupd_c info", info, 0);
  /* analytic spectrum: (nx+1)^2 (4 - 2cos(i pi/(nx+1)) - 2cos(j pi/(nx+1))); the four largest */
  double all[N];
  const double pi = 3.14159265358979323846, h2 = (NX + 1.0) * (NX + 1.0);
  for (int i = 1; i <= NX; ++i)
    for (int j = 1; j <= NX; ++j)
      all[(i - 1) * NX + j - 1] = h2 * (4.0 - 2.0 * cos(i * pi / (NX + 1)) - 2.0 * cos(j * pi / (NX + 1)));
  qsort(all, N, sizeof(double), cmp_double);
  for (int k = 0; k < NEV; ++k)
    if (fabs(d[k] - all[N - NEV + k]) > 1e-9 * all[N - 1]) return fail("dssimp eigenvalue", d[k], all[N - NEV + k]);
  /* residuals || A z - d z || with the caller's own operator */
  static double az[N];
  for (int k = 0; k < NEV; ++k) {
    lap2d(NX, z + k * N, az);
    double r = 0.0;
    for (int i = 0; i < N; ++i) r += (az[i] - d[k] * z[k * N + i]) * (az[i] - d[k] * z[k * N + i]);
    if (sqrt(r) > 1e-8 * fabs(d[k])) return fail("dssimp residual", sqrt(r), 0.0);
  }
  /* the counters of COMMON /timing/ through ICB/stat_c.h */
  a_int nopx = 0, nbx = 0, nrorth = 0, nitref = 0, nrstrt = 0;
  float t[26];
  stat_c(&nopx, &nbx, &nrorth, &nitref, &nrstrt, t, t + 1, t + 2, t + 3, t + 4, t + 5, t + 6, t + 7, t + 8, t + 9, t + 10,
         t + 11, t + 12, t + 13, t + 14, t + 15, t + 16, t + 17, t + 18, t + 19, t + 20, t + 21, t + 22, t + 23, t + 24,
         t + 25);
  if (nopx != handoffs) return fail("stat_c nopx", nopx, handoffs);
  printf("icb_caller: dsaupd_c/dseupd_c OK (%d OP*x, %d restarts)\n", handoffs, (int)iparam[2]);
  return 0;
}

/* 2. the gfortran-ABI names: every argument by reference, CHARACTER lengths appended by value */
static int sym_fortran_abi(void) {
  enum { N = 400, NEV = 5, NCV = 15, LWORKL = NCV * NCV + 8 * NCV };
  static double resid[N], v[N * NCV], workd[3 * N], workl[LWORKL], d[NEV], z[N * NEV];
  a_int iparam[11] = {0}, ipntr[11] = {0}, select[NCV];
  a_int ido = 0, info = 0, n = N, nev = NEV, ncv = NCV, ldv = N, lworkl = LWORKL, rvec = 1;
  double tol = 1e-10, sigma = 0.0;
  iparam[0] = 1; iparam[2] = 1000; iparam[3] = 1; iparam[6] = 1;
  for (;;) {
    dsaupd_(&ido, "I", &n, "LA", &nev, &tol, resid, &ncv, v, &ldv, iparam, ipntr, workd, workl, &lworkl, &info, 1, 2);
    if (ido != -1 && ido != 1) break;
    const double* x = workd + ipntr[0] - 1;
    double* y = workd + ipntr[1] - 1;
    for (int i = 0; i < N; ++i) y[i] = (i + 1.0) * x[i]; /* A = diag(1..N) */
  }
  if (info == -9990) return 3;
  if (info != 0) return fail("dsaupd_ info", info, 0);
  dseupd_(&rvec, "A", select, d, z, &ldv, &sigma, "I", &n, "LA", &nev, &tol, resid, &ncv, v, &ldv, iparam, ipntr, workd,
          workl, &lworkl, &info, 1, 1, 2);
  if (info != 0) return fail("dseupd_ info", info, 0);
  for (int k = 0; k < NEV; ++k)
    if (fabs(d[k] - (N - NEV + 1.0 + k)) > 1e-8) return fail("dsaupd_ eigenvalue", d[k], N - NEV + 1.0 + k);
  printf("icb_caller: dsaupd_/dseupd_ (Fortran ABI) OK\n");
  return 0;
}

/* 3. dnaupd_c / dneupd_c on a non-symmetric bidiagonal-plus-corner matrix with a known real spectrum 1..N */
static int nonsym_bind_c(void) {
  enum { N = 300, NEV = 4, NCV = 20, LWORKL = 3 * NCV * NCV + 6 * NCV };
  static double resid[N], v[N * NCV], workd[3 * N], workl[LWORKL], dr[NEV + 1], di[NEV + 1], z[N * (NEV + 1)],
      workev[3 * NCV];
  a_int iparam[11] = {0}, ipntr[14] = {0}, select[NCV];
  a_int ido = 0, info = 0;
  iparam[0] = 1; iparam[2] = 3000; iparam[3] = 1; iparam[6] = 1;
  for (;;) {
    dnaupd_c(&ido, "I", N, "LM", NEV, 1e-10, resid, NCV, v, N, iparam, ipntr, workd, workl, LWORKL, &info);
    if (ido != -1 && ido != 1) break;
    const double* x = workd + ipntr[0] - 1;
    double* y = workd + ipntr[1] - 1;
    for (int i = 0; i < N; ++i) y[i] = (i + 1.0) * x[i] + (i + 1 < N ? 0.5 * x[i + 1] : 0.0); /* upper bidiagonal */
  }
  if (info == -9990) return 3;
  if (info != 0) return fail("dnaupd_c info", info, 0);
  dneupd_c(1, "A", select, dr, di, z, N, 0.0, 0.0, workev, "I", N, "LM", NEV, 1e-10, resid, NCV, v, N, iparam, ipntr, workd,
           workl, LWORKL, &info);
  if (info != 0) return fail("dneupd_c info", info, 0);
  double got[NEV];
  for (int k = 0; k < NEV; ++k) {
    if (di[k] != 0.0) return fail("dneupd_c imaginary part", di[k], 0.0);
    got[k] = dr[k];
  }
  qsort(got, NEV, sizeof(double), cmp_double);
  for (int k = 0; k < NEV; ++k)
    if (fabs(got[k] - (N - NEV + 1.0 + k)) > 1e-7) return fail("dnaupd_c eigenvalue", got[k], N - NEV + 1.0 + k);
  printf("icb_caller: dnaupd_c/dneupd_c OK\n");
  return 0;
}

/* 5. znaupd_c / zneupd_c from C99 (a_dcomplex = double _Complex here, a struct on the library's C++ side):
 *    (a) TESTS/icb_arpack_c.c:98-165 -- A = diag((i+1)(1+i)), nev=9, ncv=19, 'LM', tol=1e-6, rvec=0;
 *    (b) shift-invert, mode 3, on the same diagonal matrix: sigma is passed BY VALUE as a C complex and must arrive
 *        intact, otherwise the back-transformed eigenvalues 1/theta + sigma are wrong. */
static int complex_bind_c(void) {
  enum { N = 1000, NEV = 9, NCV = 19, LWORKL = NCV * (3 * NCV + 5) };
  static a_dcomplex resid[N], v[N * NCV], workd[3 * N], workl[LWORKL], d[NEV + 1], z[N * NEV], workev[2 * NCV];
  static double rwork[NCV];
  a_int iparam[11] = {0}, ipntr[14] = {0}, select[NCV];
  a_int ido = 0, info = 0;
  iparam[0] = 1; iparam[2] = 10 * N; iparam[3] = 1; iparam[6] = 1;
  do {
    znaupd_c(&ido, "I", N, "LM", NEV, 1e-6, resid, NCV, v, N, iparam, ipntr, workd, workl, LWORKL, rwork, &info);
    if (ido == 1 || ido == -1) {
      const a_dcomplex* x = workd + ipntr[0] - 1;
      a_dcomplex* y = workd + ipntr[1] - 1;
      for (int i = 0; i < N; ++i) y[i] = x[i] * CMPLX(i + 1.0, i + 1.0);
    }
  } while (ido == 1 || ido == -1);
  if (info == -9990) return 3;
  if (info < 0 || iparam[4] < NEV) return fail("znaupd_c info/nconv", info, iparam[4]);
  zneupd_c(0, "A", select, d, z, N, CMPLX(0.0, 0.0), workev, "I", N, "LM", NEV, 1e-6, resid, NCV, v, N, iparam, ipntr,
           workd, workl, LWORKL, rwork, &info);
  if (info < 0) return fail("zneupd_c info", info, 0);
  for (int i = 0; i < NEV; ++i) {
    const double ref = N - (NEV - 1) + i;
    if (fabs(creal(d[i]) - ref) > 1e-5 || fabs(cimag(d[i]) - ref) > 1e-5) return fail("znaupd_c eigenvalue", creal(d[i]), ref);
  }
  /* (b) mode 3 around sigma = 500.3 + 500.2i: the nearest eigenvalues are (500+k)(1+i) */
  const a_dcomplex sigma = CMPLX(500.3, 500.2);
  ido = 0; info = 0;
  iparam[0] = 1; iparam[2] = 10 * N; iparam[3] = 1; iparam[6] = 3;
  do {
    znaupd_c(&ido, "I", N, "LM", 4, 1e-10, resid, NCV, v, N, iparam, ipntr, workd, workl, LWORKL, rwork, &info);
    if (ido == 1 || ido == -1) {
      const a_dcomplex* x = workd + ipntr[0] - 1;
      a_dcomplex* y = workd + ipntr[1] - 1;
      for (int i = 0; i < N; ++i) y[i] = x[i] / (CMPLX(i + 1.0, i + 1.0) - sigma);
    }
  } while (ido == 1 || ido == -1);
  if (info != 0 || iparam[4] < 4) return fail("znaupd_c mode 3 info/nconv", info, iparam[4]);
  zneupd_c(1, "A", select, d, z, N, sigma, workev, "I", N, "LM", 4, 1e-10, resid, NCV, v, N, iparam, ipntr, workd, workl,
           LWORKL, rwork, &info);
  if (info != 0) return fail("zneupd_c mode 3 info", info, 0);
  for (int k = 0; k < 4; ++k) {
    /* every returned value must be a diagonal entry m(1+i) with m in 499..502 */
    const double m = floor(creal(d[k]) + 0.5);
    if (m < 499 || m > 502 || fabs(creal(d[k]) - m) > 1e-7 || fabs(cimag(d[k]) - m) > 1e-7)
      return fail("zneupd_c mode 3 eigenvalue (sigma by value)", creal(d[k]), cimag(d[k]));
    /* and z(:,k) the matching unit vector up to a phase */
    if (fabs(cabs(z[(size_t)k * N + (int)m - 1]) - 1.0) > 1e-7) return fail("zneupd_c mode 3 eigenvector", cabs(z[(size_t)k * N + (int)m - 1]), 1.0);
  }
  printf("icb_caller: znaupd_c/zneupd_c OK\n");
  return 0;
}

/* 4. argument errors come back as info < 0 with ido = 99, never as an abort (dsaupd.f:539-543) */
static int argument_errors(void) {
  enum { N = 50, NCV = 10, LWORKL = NCV * NCV + 8 * NCV };
  static double resid[N], v[N * NCV], workd[3 * N], workl[LWORKL];
  a_int iparam[11] = {0}, ipntr[11] = {0};
  iparam[0] = 1; iparam[2] = 10; iparam[3] = 1; iparam[6] = 1;
  a_int ido = 0, info = 0;
  dsaupd_c(&ido, "I", N, "LM", 12, 0.0, resid, NCV, v, N, iparam, ipntr, workd, workl, LWORKL, &info); /* ncv <= nev */
  if (info == -9990) return 3;
  if (info != -3 || ido != 99) return fail("ncv <= nev must give info -3", info, -3);
  ido = 0; info = 0;
  dsaupd_c(&ido, "I", N, "XX", 3, 0.0, resid, NCV, v, N, iparam, ipntr, workd, workl, LWORKL, &info);
  if (info != -5 || ido != 99) return fail("bad which must give info -5", info, -5);
  printf("icb_caller: argument errors OK\n");
  return 0;
}

int main(void) {
  debug_c(6, -3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  sstats_c();
  int (*parts[])(void) = {sym_bind_c, sym_fortran_abi, nonsym_bind_c, argument_errors, complex_bind_c};
  for (unsigned k = 0; k < sizeof(parts) / sizeof(parts[0]); ++k) {
    const int rc = parts[k]();
    if (rc == 3) {
      fprintf(stderr, "icb_caller: the library reports no usable CUDA device (info = -9990)\n");
      return 3;
    }
    if (rc != 0) return 1;
  }
  printf("icb_caller: all checks passed\n");
  return 0;
}
