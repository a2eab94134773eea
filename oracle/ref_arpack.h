/*
 * oracle/ref_arpack.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the implicitly restarted Lanczos/Arnoldi path of
 * ARPACK-NG 3.9.x, used as the parity oracle for the CUDA implementation.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (arpack-ng_b200/csrc) never links,
 * includes or calls anything in this directory.
 *
 * The reference is Fortran 77; it cannot be compiled in this image (no Fortran
 * compiler, no MPI), so `oracle/_ref` does not exist.  Parity is pinned against
 * the reference's own known-answer tests instead (tests/test_oracle_golden.py):
 *   TESTS/icb_arpack_c.c:31-91, TESTS/bug_1315_double.c:23-84,
 *   PARPACK/TESTS/MPI/icb_parpack_c.c:30-102, EXAMPLES/SIMPLE/dssimp.f:180-287,
 *   EXAMPLES/SIMPLE/dnsimp.f, and LAPACK dlarnv streams.
 *
 * All BLAS/LAPACK arithmetic is delegated to the OpenBLAS that SciPy bundles
 * (symbols scipy_<name>_), i.e. the same routines the Fortran calls
 * (un-vendored dependency of the reference: CMakeLists.txt:284,296).
 *
 * SAVE/COMMON state of the Fortran lives in an explicit context (ref_ctx); one
 * context == one "process" of the reference (e.g. the dgetv0 seed persists
 * across solves made with the same context, SRC/dgetv0.f:164,202-208).
 *
 * PARPACK semantics (PARPACK/SRC/MPI/pd*.f) are selected with
 * ref_ctx_set_comm(): every MPI_ALLREDUCE of the reference becomes a call to
 * the user-supplied all-reduce callback.
 */
#ifndef REF_ARPACK_H
#define REF_ARPACK_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_ctx ref_ctx;

/* op: 0 = SUM, 1 = MAX, 2 = MIN.  In-place on buf (count elements, double or float). */
typedef void (*ref_allreduce_fn)(void* user, void* buf, int count, int is_double, int op);

ref_ctx* ref_ctx_new(void);
void ref_ctx_free(ref_ctx*);
/* Switch the context to PARPACK semantics (pdsaupd/pdnaupd ...). */
void ref_ctx_set_comm(ref_ctx*, int rank, int nranks, ref_allreduce_fn fn, void* user);
/* COMMON /timing/ counters (stat.h:11) */
void ref_ctx_stats(const ref_ctx*, int* nopx, int* nbx, int* nrorth, int* nitref, int* nrstrt);
/* number of OpenBLAS threads used by the oracle's BLAS (cpu_baseline reporting) */
void ref_set_blas_threads(int nthreads);
int ref_get_blas_threads(void);

/* ---- double precision (SRC/dsaupd.f, dseupd.f, dnaupd.f, dneupd.f; PARPACK pd* when comm set).
 * Fortran calling semantics: every scalar by pointer, tol is in/out.               */
void ref_dsaupd(ref_ctx*, int* ido, const char* bmat, int n, const char* which, int nev, double* tol,
                double* resid, int ncv, double* v, int ldv, int* iparam, int* ipntr, double* workd,
                double* workl, int lworkl, int* info);
void ref_dseupd(ref_ctx*, int rvec, const char* howmny, int* select, double* d, double* z, int ldz,
                double sigma, const char* bmat, int n, const char* which, int nev, double tol,
                double* resid, int ncv, double* v, int ldv, int* iparam, int* ipntr, double* workd,
                double* workl, int lworkl, int* info);
void ref_dnaupd(ref_ctx*, int* ido, const char* bmat, int n, const char* which, int nev, double* tol,
                double* resid, int ncv, double* v, int ldv, int* iparam, int* ipntr, double* workd,
                double* workl, int lworkl, int* info);
void ref_dneupd(ref_ctx*, int rvec, const char* howmny, int* select, double* dr, double* di, double* z,
                int ldz, double sigmar, double sigmai, double* workev, const char* bmat, int n,
                const char* which, int nev, double tol, double* resid, int ncv, double* v, int ldv,
                int* iparam, int* ipntr, double* workd, double* workl, int lworkl, int* info);

/* ---- single precision twins (SRC/ssaupd.f ...) */
void ref_ssaupd(ref_ctx*, int* ido, const char* bmat, int n, const char* which, int nev, float* tol,
                float* resid, int ncv, float* v, int ldv, int* iparam, int* ipntr, float* workd,
                float* workl, int lworkl, int* info);
void ref_sseupd(ref_ctx*, int rvec, const char* howmny, int* select, float* d, float* z, int ldz,
                float sigma, const char* bmat, int n, const char* which, int nev, float tol,
                float* resid, int ncv, float* v, int ldv, int* iparam, int* ipntr, float* workd,
                float* workl, int lworkl, int* info);
void ref_snaupd(ref_ctx*, int* ido, const char* bmat, int n, const char* which, int nev, float* tol,
                float* resid, int ncv, float* v, int ldv, int* iparam, int* ipntr, float* workd,
                float* workl, int lworkl, int* info);
void ref_sneupd(ref_ctx*, int rvec, const char* howmny, int* select, float* dr, float* di, float* z,
                int ldz, float sigmar, float sigmai, float* workev, const char* bmat, int n,
                const char* which, int nev, float tol, float* resid, int ncv, float* v, int ldv,
                int* iparam, int* ipntr, float* workd, float* workl, int lworkl, int* info);

/* ---- helpers used by tests / bench (plain CPU, reference semantics) ---- */
/* LAPACK dlarnv(idist=2) stream exactly as SRC/dgetv0.f:236 draws it */
void ref_dlarnv2(int* iseed4, int n, double* x);
/* y = A x for CSR (int32 indices); threaded over rows with nthreads pthreads (cpu_baseline OP) */
void ref_csr_spmv(int nrows, const int* rowptr, const int* col, const double* val, const double* x,
                  double* y, int nthreads);

/* 2-D 5-point Laplacian (4,-1)*scale in CSR built on the CPU (operator of BASELINE config 2); returns nnz */
long long ref_gen_laplace2d(int nx, int ny, double scale, int* rowptr, int* col, double* val);

/* Run a whole symmetric solve (dsaupd loop + optional dseupd) on a CSR operator, mode 1, bmat='I'.
 * Used for timing the CPU baseline without Python in the loop.  Returns info of dsaupd.
 * out_counts = {iparam(3), iparam(5), nopx, nbx, nrorth}; d (nev) and z (n*nev, may be NULL). */
int ref_dsaupd_csr_solve(ref_ctx*, int n, const int* rowptr, const int* col, const double* val,
                         const char* which, int nev, int ncv, double tol, int mxiter, int info_in,
                         double* resid, double* v, double* d, double* z, double* workl_out,
                         int* out_counts, int spmv_threads, double* seconds_total, double* seconds_op);

#ifdef __cplusplus
}
#endif
#endif
