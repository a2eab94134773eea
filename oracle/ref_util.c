/* oracle/ref_util.c -- TEST INFRASTRUCTURE ONLY (see ref_arpack.h).
 * Context management, the LAPACK dlarnv stream, a threaded CSR SpMV and a whole-solve
 * driver used to time the CPU baseline without Python in the loop. */
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "ref_ctx.h"

extern void scipy_openblas_set_num_threads(int);
extern int scipy_openblas_get_num_threads(void);
extern void scipy_dlarnv_(const int* idist, int* iseed, const int* n, double* x);

ref_ctx* ref_ctx_new(void) { return (ref_ctx*)calloc(1, sizeof(ref_ctx)); }
void ref_ctx_free(ref_ctx* c) {
  if (!c) return;
  free(c->dstate);
  free(c->sstate);
  free(c->zstate);
  free(c->cstate);
  free(c);
}
void ref_ctx_set_comm(ref_ctx* c, int rank, int nranks, ref_allreduce_fn fn, void* user) {
  c->par = 1;
  c->rank = rank;
  c->nranks = nranks;
  c->ar = fn;
  c->ar_user = user;
}
void ref_ctx_stats(const ref_ctx* c, int* nopx, int* nbx, int* nrorth, int* nitref, int* nrstrt) {
  if (nopx) *nopx = c->nopx;
  if (nbx) *nbx = c->nbx;
  if (nrorth) *nrorth = c->nrorth;
  if (nitref) *nitref = c->nitref;
  if (nrstrt) *nrstrt = c->nrstrt;
}
void ref_set_blas_threads(int nthreads) { scipy_openblas_set_num_threads(nthreads); }
int ref_get_blas_threads(void) { return scipy_openblas_get_num_threads(); }

void ref_dlarnv2(int* iseed4, int n, double* x) {
  int idist = 2;
  scipy_dlarnv_(&idist, iseed4, &n, x);
}

void ref_csr_spmv(int nrows, const int* rowptr, const int* col, const double* val, const double* x,
                  double* y, int nthreads) {
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int i = 0; i < nrows; ++i) {
    double acc = 0.0;
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += val[p] * x[col[p]];
    y[i] = acc;
  }
}

/* 2-D 5-point Laplacian (4, -1)*scale in CSR, row-major natural ordering (BASELINE config 2 operator),
 * built on the CPU for the reference arm of bench.py.  Returns nnz; arrays sized by the caller
 * (nnz = 5*nx*ny - 2*nx - 2*ny). */
long long ref_gen_laplace2d(int nx, int ny, double scale, int* rowptr, int* col, double* val) {
  long long p = 0;
  for (int iy = 0; iy < ny; ++iy) {
    for (int ix = 0; ix < nx; ++ix) {
      const long long r = (long long)iy * nx + ix;
      rowptr[r] = (int)p;
      if (iy > 0) { col[p] = (int)(r - nx); val[p] = -scale; ++p; }
      if (ix > 0) { col[p] = (int)(r - 1); val[p] = -scale; ++p; }
      col[p] = (int)r; val[p] = 4.0 * scale; ++p;
      if (ix < nx - 1) { col[p] = (int)(r + 1); val[p] = -scale; ++p; }
      if (iy < ny - 1) { col[p] = (int)(r + nx); val[p] = -scale; ++p; }
    }
  }
  rowptr[(long long)nx * ny] = (int)p;
  return p;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int ref_dsaupd_csr_solve(ref_ctx* c, int n, const int* rowptr, const int* col, const double* val,
                         const char* which, int nev, int ncv, double tol, int mxiter, int info_in,
                         double* resid, double* v, double* d, double* z, double* workl_out,
                         int* out_counts, int spmv_threads, double* seconds_total, double* seconds_op) {
  int lworkl = ncv * ncv + 8 * ncv;
  double* workl = (double*)calloc((size_t)lworkl, sizeof(double));
  double* workd = (double*)calloc((size_t)3 * n, sizeof(double));
  int iparam[11] = {0}, ipntr[11] = {0};
  iparam[0] = 1;
  iparam[2] = mxiter;
  iparam[3] = 1;
  iparam[6] = 1;
  int ido = 0, info = info_in;
  double t0 = now_s(), top = 0.0;
  double tolv = tol;
  for (;;) {
    ref_dsaupd(c, &ido, "I", n, which, nev, &tolv, resid, ncv, v, n, iparam, ipntr, workd, workl, lworkl, &info);
    if (ido == -1 || ido == 1) {
      double a = now_s();
      ref_csr_spmv(n, rowptr, col, val, workd + ipntr[0] - 1, workd + ipntr[1] - 1, spmv_threads);
      top += now_s() - a;
    } else {
      break;
    }
  }
  int info_aupd = info;
  if (d && info >= 0 && iparam[4] > 0) {
    int* select = (int*)calloc((size_t)ncv, sizeof(int));
    int ierr = 0;
    /* Fortran semantics: dseupd sees the tol that dsaupd updated in place */
    ref_dseupd(c, z != NULL, "A", select, d, z ? z : v, n, 0.0, "I", n, which, nev, tolv, resid, ncv, v, n,
               iparam, ipntr, workd, workl, lworkl, &ierr);
    free(select);
    if (ierr != 0) info_aupd = 1000 - ierr; /* flag an eupd failure distinctly */
  }
  if (seconds_total) *seconds_total = now_s() - t0;
  if (seconds_op) *seconds_op = top;
  if (out_counts) {
    out_counts[0] = iparam[2];
    out_counts[1] = iparam[4];
    out_counts[2] = iparam[8];
    out_counts[3] = iparam[9];
    out_counts[4] = iparam[10];
  }
  if (workl_out) memcpy(workl_out, workl, sizeof(double) * (size_t)lworkl);
  free(workl);
  free(workd);
  return info_aupd;
}
