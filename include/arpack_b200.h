/*
 * arpack_b200.h -- C ABI of libarpack_b200.so, the B200-native drop-in for ARPACK-NG's implicitly
 * restarted Lanczos/Arnoldi hot path (dsaupd/dseupd, dnaupd/dneupd, PARPACK pdsaupd/pdnaupd).
 *
 * Every entry point below is what a binding of the reference for this path binds; the reference
 * interface each one replaces is cited as file:line relative to the arpack-ng source tree.
 * Plain pointers and sizes only; no CUDA or torch types appear in any signature.
 *
 * Array arguments resid, v, workd (and z) may be HOST pointers (an unmodified CPU caller: the
 * library mirrors them in HBM and copies the hand-off vectors across PCIe) or DEVICE pointers
 * (used in place; on ido = -1/1/2 the operand workd + ipntr[0] - 1 and the result slot
 * workd + ipntr[1] - 1 are device addresses, ordered on the stream of ab200_get_stream()).
 * workl, iparam, ipntr, select, d/dr/di, workev are always host memory.
 *
 * Error behaviour follows the reference (info < 0 with ido = 99 for argument errors, info = 1 max
 * iterations, 3 no shifts, -8/-9/-9999 as in SRC/dsaupd.f:243-276), plus info = -9990 when no CUDA
 * device is usable or a CUDA call failed (message on stderr).  There is no CPU fallback.
 */
#ifndef ARPACK_B200_H
#define ARPACK_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int a_int;  /* arpackdef.h.in:8-14 (LP64 build) */
typedef int a_fint; /* MPI_Fint of ICB/parpack.h:7-8; here: handle from ab200_comm_create() */
/* a_fcomplex / a_dcomplex of arpackdef.h.in:19-41 (`float _Complex` / `double _Complex`).  In C they are exactly
 * those types; a C++ translation unit sees a two-member struct, which the SysV x86-64 and AArch64 ABIs lay out AND pass
 * by value like the C99 complex types (ICB/arpack.hpp:213-216 relies on the same equivalence for std::complex). */
#ifdef __cplusplus
typedef struct { float re, im; } a_fcomplex;
typedef struct { double re, im; } a_dcomplex;
#else
#include <complex.h>
typedef float _Complex a_fcomplex;
typedef double _Complex a_dcomplex;
#endif

/* ---- ICB/arpack.h:14-21 (bind(c) shims SRC/icbads.F90:3-92, icbadn.F90:3-97, icbass.F90, icbasn.F90) ---- */
void dsaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid,
              a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
              a_int lworkl, a_int* info);
void dseupd_c(a_int rvec, char const* howmny, a_int const* select, double* d, double* z, a_int ldz, double sigma,
              char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid, a_int ncv,
              double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl, a_int lworkl,
              a_int* info);
void dnaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid,
              a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
              a_int lworkl, a_int* info);
void dneupd_c(a_int rvec, char const* howmny, a_int const* select, double* dr, double* di, double* z, a_int ldz,
              double sigmar, double sigmai, double* workev, char const* bmat, a_int n, char const* which,
              a_int nev, double tol, double* resid, a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr,
              double* workd, double* workl, a_int lworkl, a_int* info);
void ssaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid,
              a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl,
              a_int lworkl, a_int* info);
void sseupd_c(a_int rvec, char const* howmny, a_int const* select, float* d, float* z, a_int ldz, float sigma,
              char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid, a_int ncv, float* v,
              a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl, a_int lworkl, a_int* info);
void snaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid,
              a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl,
              a_int lworkl, a_int* info);
void sneupd_c(a_int rvec, char const* howmny, a_int const* select, float* dr, float* di, float* z, a_int ldz,
              float sigmar, float sigmai, float* workev, char const* bmat, a_int n, char const* which, a_int nev,
              float tol, float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd,
              float* workl, a_int lworkl, a_int* info);

/* ---- complex Arnoldi: ICB/arpack.h:10-11,20-21 (bind(c) shims SRC/icbacn.F90, icbazn.F90:3-94 over SRC/cnaupd.f,
 * cneupd.f, znaupd.f:384, zneupd.f:248).  workl holds 3*ncv^2 + 5*ncv complex entries, rwork ncv reals, d nev+1,
 * workev 2*ncv; which is one of LM SM LR SR LI SI; modes 1-3.  resid, v, workd, z: host or device, as above. */
void znaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol, a_dcomplex* resid,
              a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_dcomplex* workd, a_dcomplex* workl,
              a_int lworkl, double* rwork, a_int* info);
void zneupd_c(a_int rvec, char const* howmny, a_int const* select, a_dcomplex* d, a_dcomplex* z, a_int ldz,
              a_dcomplex sigma, a_dcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev,
              double tol, a_dcomplex* resid, a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr,
              a_dcomplex* workd, a_dcomplex* workl, a_int lworkl, double* rwork, a_int* info);
void cnaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol, a_fcomplex* resid,
              a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_fcomplex* workd, a_fcomplex* workl,
              a_int lworkl, float* rwork, a_int* info);
void cneupd_c(a_int rvec, char const* howmny, a_int const* select, a_fcomplex* d, a_fcomplex* z, a_int ldz,
              a_fcomplex sigma, a_fcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev, float tol,
              a_fcomplex* resid, a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_fcomplex* workd,
              a_fcomplex* workl, a_int lworkl, float* rwork, a_int* info);
/* the same with sigma as two reals, for bindings that cannot pass a C complex by value (Python ctypes) */
void ab200_zneupd_ri(a_int rvec, char const* howmny, a_int const* select, void* d, void* z, a_int ldz, double sigma_re,
                     double sigma_im, void* workev, char const* bmat, a_int n, char const* which, a_int nev, double tol,
                     void* resid, a_int ncv, void* v, a_int ldv, a_int* iparam, a_int* ipntr, void* workd, void* workl,
                     a_int lworkl, double* rwork, a_int* info);
void ab200_cneupd_ri(a_int rvec, char const* howmny, a_int const* select, void* d, void* z, a_int ldz, float sigma_re,
                     float sigma_im, void* workev, char const* bmat, a_int n, char const* which, a_int nev, float tol,
                     void* resid, a_int ncv, void* v, a_int ldv, a_int* iparam, a_int* ipntr, void* workd, void* workl,
                     a_int lworkl, float* rwork, a_int* info);
/* PARPACK twins, ICB/parpack.h:28-33 (PARPACK/SRC/MPI/pznaupd.f, pzneupd.f, pcnaupd.f, pcneupd.f): n = local rows,
 * comm = handle from ab200_comm_create(); the MPI_ALLREDUCEs of pznaitr.f:437-449, pzgetv0.f:322-328 and pdznorm2.f
 * are all-reduces of the device mailbox on the solve's stream. */
void pznaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol,
               a_dcomplex* resid, a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_dcomplex* workd,
               a_dcomplex* workl, a_int lworkl, double* rwork, a_int* info);
void pzneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, a_dcomplex* d, a_dcomplex* z,
               a_int ldz, a_dcomplex sigma, a_dcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev,
               double tol, a_dcomplex* resid, a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr,
               a_dcomplex* workd, a_dcomplex* workl, a_int lworkl, double* rwork, a_int* info);
void pcnaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol,
               a_fcomplex* resid, a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_fcomplex* workd,
               a_fcomplex* workl, a_int lworkl, float* rwork, a_int* info);
void pcneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, a_fcomplex* d, a_fcomplex* z,
               a_int ldz, a_fcomplex sigma, a_fcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev,
               float tol, a_fcomplex* resid, a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr,
               a_fcomplex* workd, a_fcomplex* workl, a_int lworkl, float* rwork, a_int* info);
void ab200_pzneupd_ri(a_fint comm, a_int rvec, char const* howmny, a_int const* select, void* d, void* z, a_int ldz,
                      double sigma_re, double sigma_im, void* workev, char const* bmat, a_int n, char const* which,
                      a_int nev, double tol, void* resid, a_int ncv, void* v, a_int ldv, a_int* iparam, a_int* ipntr,
                      void* workd, void* workl, a_int lworkl, double* rwork, a_int* info);
/* legacy Fortran ABI of the same (SRC/znaupd.f:384, zneupd.f:248) */
void znaupd_(a_int* ido, const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, a_dcomplex* resid,
             a_int* ncv, a_dcomplex* v, a_int* ldv, a_int* iparam, a_int* ipntr, a_dcomplex* workd, a_dcomplex* workl,
             a_int* lworkl, double* rwork, a_int* info, size_t bmat_len, size_t which_len);
void zneupd_(a_int* rvec, const char* howmny, a_int* select, a_dcomplex* d, a_dcomplex* z, a_int* ldz,
             a_dcomplex* sigma, a_dcomplex* workev, const char* bmat, a_int* n, const char* which, a_int* nev,
             double* tol, a_dcomplex* resid, a_int* ncv, a_dcomplex* v, a_int* ldv, a_int* iparam, a_int* ipntr,
             a_dcomplex* workd, a_dcomplex* workl, a_int* lworkl, double* rwork, a_int* info, size_t howmny_len,
             size_t bmat_len, size_t which_len);

/* ---- ICB/parpack.h:12-27 (PARPACK/SRC/MPI/icbpds.F90, icbpdn.F90, icbpss.F90, icbpsn.F90).
 * n is the LOCAL row count; all ranks call in lock-step.  MPI_ALLREDUCE of the reference
 * (pdsaitr.f:604,720; pdnorm2.f:72-80; pdgetv0.f:369) is an NCCL all-reduce on the solve's stream. */
void pdsaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol,
               double* resid, a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd,
               double* workl, a_int lworkl, a_int* info);
void pdseupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, double* d, double* z, a_int ldz,
               double sigma, char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid,
               a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
               a_int lworkl, a_int* info);
void pdnaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol,
               double* resid, a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd,
               double* workl, a_int lworkl, a_int* info);
void pdneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, double* dr, double* di,
               double* z, a_int ldz, double sigmar, double sigmai, double* workev, char const* bmat, a_int n,
               char const* which, a_int nev, double tol, double* resid, a_int ncv, double* v, a_int ldv,
               a_int* iparam, a_int* ipntr, double* workd, double* workl, a_int lworkl, a_int* info);
void pssaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol,
               float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd,
               float* workl, a_int lworkl, a_int* info);
void psseupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, float* d, float* z, a_int ldz,
               float sigma, char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid,
               a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl,
               a_int lworkl, a_int* info);
void psnaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol,
               float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd,
               float* workl, a_int lworkl, a_int* info);
void psneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, float* dr, float* di, float* z,
               a_int ldz, float sigmar, float sigmai, float* workev, char const* bmat, a_int n, char const* which,
               a_int nev, float tol, float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr,
               float* workd, float* workl, a_int lworkl, a_int* info);

/* ---- legacy Fortran ABI (VISUAL_STUDIO/arpack-ng_exports.def; SRC/dsaupd.f:408, dseupd.f:218,
 * dnaupd.f:406, dneupd.f:302): all arguments by reference, CHARACTER lengths appended by value ---- */
void dsaupd_(a_int* ido, const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, double* resid,
             a_int* ncv, double* v, a_int* ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
             a_int* lworkl, a_int* info, size_t bmat_len, size_t which_len);
void dseupd_(a_int* rvec, const char* howmny, a_int* select, double* d, double* z, a_int* ldz, double* sigma,
             const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, double* resid, a_int* ncv,
             double* v, a_int* ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl, a_int* lworkl,
             a_int* info, size_t howmny_len, size_t bmat_len, size_t which_len);
void dnaupd_(a_int* ido, const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, double* resid,
             a_int* ncv, double* v, a_int* ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
             a_int* lworkl, a_int* info, size_t bmat_len, size_t which_len);
void dneupd_(a_int* rvec, const char* howmny, a_int* select, double* dr, double* di, double* z, a_int* ldz,
             double* sigmar, double* sigmai, double* workev, const char* bmat, a_int* n, const char* which,
             a_int* nev, double* tol, double* resid, a_int* ncv, double* v, a_int* ldv, a_int* iparam,
             a_int* ipntr, double* workd, double* workl, a_int* lworkl, a_int* info, size_t howmny_len,
             size_t bmat_len, size_t which_len);

/* ---- ICB/debug_c.h:7 (ICB/debug_icb.F90), ICB/stat_c.h:7-14 (ICB/stat_icb.F90) ---- */
void debug_c(a_int logfil, a_int ndigit, a_int mgetv0, a_int msaupd, a_int msaup2, a_int msaitr, a_int mseigt,
             a_int msapps, a_int msgets, a_int mseupd, a_int mnaupd, a_int mnaup2, a_int mnaitr, a_int mneigh,
             a_int mnapps, a_int mngets, a_int mneupd, a_int mcaupd, a_int mcaup2, a_int mcaitr, a_int mceigh,
             a_int mcapps, a_int mcgets, a_int mceupd);
void sstats_c(void);
void sstatn_c(void);
void stat_c(a_int* nopx, a_int* nbx, a_int* nrorth, a_int* nitref, a_int* nrstrt, float* tsaupd, float* tsaup2,
            float* tsaitr, float* tseigt, float* tsgets, float* tsapps, float* tsconv, float* tnaupd, float* tnaup2,
            float* tnaitr, float* tneigh, float* tngets, float* tnapps, float* tnconv, float* tcaupd, float* tcaup2,
            float* tcaitr, float* tceigh, float* tcgets, float* tcapps, float* tcconv, float* tmvopx, float* tmvbx,
            float* tgetv0, float* titref, float* trvec);

/* ---- extensions (no counterpart in the reference) ---- */
/* CUDA stream (cudaStream_t as void*) on which the library enqueues and on which the hand-off is ordered */
void ab200_set_stream(void* cuda_stream);
void* ab200_get_stream(void);
/* 0 = automatic (TMA-tiled kernels when the layout allows), 1 = generic kernels only */
void ab200_set_kernel_mode(int mode);
/* free the device mirrors / solver state keyed to this workl (also done when workl is reused with ido = 0) */
/* Compatibility switches for callers that rely on side effects of the reference the hot path does not need (default 0):
 *   bit 0: maintain workd(ipntr(3)) = B*x (= x) at every ido = 1 hand-off of mode 1 as SRC/dsaitr.f:517 / dnaitr.f:505
 *          does; by default that slot is only written where the protocol documents it (bmat = 'G', modes 2-5), which
 *          saves one n-vector store per Lanczos step.  Applies to solves started (ido = 0) after the call. */
void ab200_set_compat(int flags);
void ab200_release(const void* workl);
void ab200_release_all(void);
/* forget the per-address nnz cache of ab200_csr_spmv_f64/f32 (also done by ab200_release_all) */
void ab200_forget_csr_cache(void);
/* out4 = {kernels launched, all-reduces issued, TMA-path launches, generic-path launches} since load */
void ab200_launch_stats(unsigned long long* out4);
/* blocking device->host mailbox reads since load: one per Lanczos/Arnoldi step in the synchronous mode, one per sweep
   when the sweep runs device-resident (IrlBase::extend) */
unsigned long long ab200_host_round_trips(void);
/* forget the SAVE'd dgetv0 seed / dnaitr smlnum, as if the process had just started */
void ab200_reset_seed(void);
/* Registered-operator mode (opt-in; strict RCI stays the default): the library applies y = A x (square CSR matrix in
 * HBM, int32 indices) itself for the solve keyed to `workl` (mode 1, bmat 'I'), so one *aupd_c call runs the whole solve
 * -- the loop of EXAMPLES/MATRIX_MARKET/arpackSolver.hpp:787-846 without a host round trip per step.  K1+K2 are folded
 * into the SpMV and alpha = v^T OP v, ||OP v||^2 come out of its epilogue.  The registration is ONE-SHOT: it is consumed
 * by the next ido = 0 call with this workl (and dropped there if the solve is not mode 1 / bmat 'I' / sequential), so a
 * later solve that happens to reuse the address runs the plain protocol.  nrows = 0 unregisters.
 * rowptr/col/val may each be DEVICE arrays (used in place, must outlive the solve) or HOST arrays (copied to HBM once at
 * ido = 0 and owned by the solve's context; they must stay valid until that call): a caller whose matrix, resid, v and
 * workd all live in host memory -- every caller of the reference -- runs the whole solve on the GPU with one upload of
 * the matrix and one download of V/resid, instead of four PCIe crossings of n values per Lanczos step. */
int ab200_register_csr_op_f64(const void* workl, int nrows, long long nnz, const int* rowptr, const int* col,
                              const double* val);
int ab200_register_csr_op_f32(const void* workl, int nrows, long long nnz, const int* rowptr, const int* col,
                              const float* val);
/* The same under a communicator (p*aupd_c): nloc local rows of a row-partitioned matrix; columns >= nloc address
 * halo_buf = [last halo_lo entries of the lower neighbour's x | first halo_hi of the upper neighbour's], filled by the
 * library with an NCCL neighbour exchange before every product (PARPACK/EXAMPLES/MPI/pdsdrv1.f:463-483). */
int ab200_register_csr_halo_op_f64(const void* workl, int comm, int nloc, long long nnz, const int* rowptr,
                                   const int* col, const double* val, int halo_lo, int halo_hi, double* halo_buf);
double ab200_fused_dot_maxdiff(const void* workl);
/* per-kernel CUDA-event timing on the launching stream (bench.py's roofline): enable, run, read the table */
void ab200_profile_enable(int on);
void ab200_profile_reset(void);
int ab200_profile_get(int idx, char* name64, double* ms, unsigned long long* launches, double* bytes);
/* kernel unit-test hooks: one fused orthogonalisation step (K4..K10) / one restart update (K12..K16) on caller-supplied
 * DEVICE arrays; out_host gets [h(0..j-1), ||w||^2, s(0..j-1), ||r||^2, ||r'||^2, dgks_flag] */
int ab200_debug_orth_f64(long long n, int j, const double* v, long long ldv, const double* w, double* resid,
                         double* out_host);
int ab200_debug_vq_f64(long long n, int kin, int kout, double* v, long long ldv, const double* q_host, double sigma,
                       double beta, int beta_col, double* resid, double* nrm2_host);
/* kernel micro-benchmark on synthetic data (tools/kernel_sweep.py): what = 0 orth step, 1 multi-dots, 2 V*Q update */
int ab200_kernel_probe_f64(long long n, int j, int ncv, int iters, int what, int kout);
/* the same hooks for the complex kernels (device arrays of interleaved complex128; see api_cplx.cu) */
int ab200_debug_zorth_f64(long long n, int j, const void* v, long long ldv, const void* w, void* resid, void* out_host);
int ab200_debug_zvq_f64(long long n, int kin, int kout, const void* v, long long ldv, const void* q_host, void* out,
                        long long ldo, double sigma_re, double sigma_im, double beta_re, double beta_im, int beta_col,
                        void* resid, double* nrm2_host);
int ab200_device_count(void);
const char* ab200_version(void);

/* NCCL communicators standing in for PARPACK's MPI communicator */
int ab200_nccl_unique_id(void* out128);                        /* rank 0: 128-byte id to broadcast */
int ab200_comm_create(const void* id128, int rank, int nranks); /* collective; returns the comm handle (>= 1) */
void ab200_comm_destroy(int handle);
int ab200_comm_rank(int handle);
/* 1 when the per-step all-reduces of this communicator run through peer memory (CUDA IPC over NVLink: one small kernel
 * per all-reduce, sums in rank order), 0 when they use ncclAllReduce (IPC unavailable, > 8 ranks, AB200_P2P=0) */
int ab200_comm_uses_p2p(int handle);
int ab200_comm_size(int handle);
/* Collective: the communicator's own halo receive buffer for a row-partitioned operator whose lower / upper neighbour
 * planes hold halo_lo / halo_hi elements of elem_size bytes (8 or 4).  Passed as halo_buf to ab200_csr_spmv_halo_f64 /
 * ab200_register_csr_halo_op_f64 it makes the plane exchange of PARPACK/EXAMPLES/MPI/pdsdrv1.f:463-483 run through peer
 * memory (the neighbours store their planes into it over NVLink).  NULL: peer memory is unavailable -- allocate an
 * ordinary device buffer of halo_lo + halo_hi elements instead (exchange by ncclSend/ncclRecv). */
void* ab200_comm_halo_buffer(int handle, long long halo_lo, long long halo_hi, int elem_size);

/* ---- driver layer: the user's OP for the arpackmm-style tool (EXAMPLES/MATRIX_MARKET/arpackSolver.hpp:806-841
 * does these products with Eigen on the CPU).  All pointers are DEVICE pointers unless named *_host. ---- */
/* y = A x, CSR with int32 indices (K3 of SURVEY.md §2.3) */
int ab200_csr_spmv_f64(int nrows, const int* rowptr, const int* col, const double* val, const double* x, double* y);
/* SpMV kernel for short-row matrices: 0 = CSR-bulk (cp.async.bulk ring, row-owner consumers; default), 1 = CSR-stream,
 * 2 = row per sub-warp, 4 = CSR-bulk with parked products (the round-1 consumers); all return the same bits */
void ab200_set_spmv_variant(int variant);
int ab200_csr_spmv_f32(int nrows, const int* rowptr, const int* col, const float* val, const float* x, float* y);
/* as above with host vectors: H2D(x), SpMV, D2H(y) -- the OP of an unmodified host RCI loop */
int ab200_csr_spmv_hostvec_f64(int nrows, int ncols, const int* rowptr, const int* col, const double* val,
                               const double* x_host, double* y_host);
/* row-partitioned SpMV with neighbour halos (PARPACK/EXAMPLES/MPI/pdsdrv1.f:463-483): columns >= nloc address
 * the halo buffer [lower | upper]; the planes are exchanged with ncclSend/ncclRecv on the solve's stream */
int ab200_csr_spmv_halo_f64(int comm, int nloc, int halo_lo, int halo_hi, const int* rowptr, const int* col,
                            const double* val, const double* x, double* y, double* halo_buf);
/* synthetic operators of BASELINE.json's configs, generated on the device; each returns nnz (or < 0).
 * Call with rowptr = NULL to query nnz only. */
long long ab200_gen_laplace2d(int nx, int ny, double scale, int* rowptr, int* col, double* val);
/* 7-point stencil (diag, -1) on nx x ny x nz, z-slab [z0, z0+nzloc) with halo columns; ny = 1, diag = 4 gives the
 * y-slab of the 2-D 5-point Laplacian */
long long ab200_gen_laplace3d(int nx, int ny, int nz, int z0, int nzloc, double diag, int* rowptr, int* col,
                              double* val);
long long ab200_gen_convdiff2d(int nx, double rho, int* rowptr, int* col, double* val);
/* ---- OP = A^T A on a row-sharded sparse A (BASELINE config 5; the caller's av + atv of EXAMPLES/SVD/dsvd.f:342-343,
 * 400-470, singular values = sqrt of the Ritz values, :416) ---- */
/* rows [row0, row0 + nrows) of the synthetic matrix of SURVEY.md 8d: exactly per_row entries per row, entry k of row r at
 * column splitmix64(seed + per_row*r + k) mod ncols with value 2u-1, u = top 53 bits of splitmix64(that hash) / 2^53.
 * rowptr = NULL queries nnz. */
long long ab200_gen_randsparse(long long row0, int nrows, int ncols, int per_row, unsigned long long seed, int* rowptr,
                               int* col, double* val);
/* CSR of A^T (ncols x nrows) from the CSR of A, on the device; entries of a row of A^T in ascending row order of A */
int ab200_csr_transpose_f64(int nrows, int ncols, long long nnz, const int* rowptr, const int* col, const double* val,
                            int* t_rowptr, int* t_col, double* t_val);
/* handle of an operator z = sum_s A_s^T (A_s x) over the shards added below; comm = 0: one process, x and z hold ncols
 * entries; comm = handle of ab200_comm_create(): x and z are this rank's slice of ncols / nranks entries (all-gather of
 * x, local products, reduce-scatter of the partial sums).  Device arrays stay owned by the caller. */
int ab200_gram_create(int comm, int ncols);
int ab200_gram_add_shard(int handle, int nrows, long long nnz, const int* rowptr, const int* col, const double* val,
                         const int* t_rowptr, const int* t_col, const double* t_val);
int ab200_gram_apply(int handle, const double* x_loc, double* z_loc);
void ab200_gram_destroy(int handle);
/* let dsaupd_c / pdsaupd_c apply the operator themselves for the solve keyed to workl (one call runs the whole solve) */
int ab200_register_gram_op_f64(const void* workl, int gram_handle);
/* start vector of SURVEY.md §8(d): resid[i] = 2 u(i0+i) - 1, u = top 53 bits of splitmix64(seed + i) / 2^53 */
int ab200_fill_hash_f64(long long n, long long i0, unsigned long long seed, double* x);
/* residual check of arpackSolver.hpp:297-352: out[k] = || A z_k - d_k z_k ||_2 (device z, host d/out) */
int ab200_residuals_f64(int n, const int* rowptr, const int* col, const double* val, int k, const double* z,
                        long long ldz, const double* d_host, double* out_host);


/* ---- on-disk formats of the tool layer (host only, no CUDA; EXAMPLES/MATRIX_MARKET/arpackSolver.hpp) ---- */
/* Coordinate file with the reader semantics of arpackSolver.hpp:361-416: '%' comments and blank lines skipped, header
 * "n m [nnz]", body "i j value", 1-based when max(i) == n or max(j) == m else 0-based; returned as CSR (malloc'ed,
 * release with ab200_mm_free) with column-sorted rows and duplicates summed (Eigen setFromTriplets, :417-424).
 * Returns 0, 1 cannot open, 2 bad header, 3 bad line, 4 index out of range, 5 too large for int32, 6 out of memory. */
int ab200_mm_read_csr(const char* path, int* nrows, int* ncols, long long* nnz, int** rowptr_host, int** col_host,
                      double** val_host);
/* the same for complex coordinate files ("i j (re, im)", the reference's Az.mtx / Bz.mtx; arpackSolver.hpp:398-400 reads
 * the value with the stream extraction of std::complex): val_host holds nnz interleaved (re, im) pairs */
int ab200_mm_read_csr_z(const char* path, int* nrows, int* ncols, long long* nnz, int** rowptr_host, int** col_host,
                        double** val_host);
void ab200_mm_free(void* p);
/* the --restart dump of arpackSolver.hpp:664-704: "count" then one value per line; load fails (2) on a count
 * mismatch and, with allow_zero = 0, replaces |value| < 1e-6 by machine epsilon (resid must not be zero) */
int ab200_restart_save_f64(const char* path, long long count, const double* values);
int ab200_restart_load_f64(const char* path, long long count, double* values, int allow_zero);

#ifdef __cplusplus
}
#endif
#endif
