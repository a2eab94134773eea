// arpackmm_b200 -- the B200 twin of the reference's Matrix-Market driver (EXAMPLES/MATRIX_MARKET/arpackmm.cpp,
// arpackSolver.hpp).  Same command-line options, same input files, same "OUT:" lines, same residual check
// (||A v - lambda B v|| <= sqrt(tol), arpackSolver.hpp:297-352) and the same --restart dump files; the eigen-solve
// runs through libarpack_b200's C-ABI with device-resident arrays and the matrix products are CSR SpMV kernels.
//
// What is and is not carried over from the reference tool:
//   * real problems in double or single precision (--simplePrec), symmetric (ds*upd) or not (--nonSymPb, dn*upd);
//   * standard problems: mode 1, a real shift is applied as A - sigma I and undone afterwards (arpackSolver.hpp:255-262);
//   * generalised problems (--genPb/--B): mode 2 (OP = B^-1 A) and, with --shiftReal, mode 3 (OP = (A - sigma B)^-1 B),
//     the inner systems solved on the GPU by CG (--slv CG) or BiCGSTAB (--slv BiCG), un-preconditioned;
//   * NOT built: complex problems (--cpxPb), dense matrices (--dense), the Eigen direct solvers (LU QR LLT LDLT) and
//     preconditioners (--slvItrPC): these options are recognised and rejected with a message, exit code 1.
//   * extension: --registered hands the CSR matrix to the library (ab200_register_csr_op_*), one *aupd call per solve.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/arpack_b200.h"

namespace {

struct Options {
  std::string fileA = "A.mtx", fileB = "N.A.";
  int nbEV = 1, nbCV = 3;
  bool stdPb = true, symPb = true, cpxPb = false, simplePrec = false, dense = false;
  std::string mag = "LM";
  bool shiftReal = false, shiftImag = false, invert = false;
  double sigmaReal = 0.0, sigmaImag = 0.0, tol = 1.e-6;
  int maxIt = 100;
  bool schur = false;
  std::string slv = "BiCG", slvItrPC = "Diagonal";
  bool slvPCGiven = false;
  double slvItrTol = 1.e-6;
  int slvItrMaxIt = 100;
  bool check = true, restart = false, registered = false;
  int verbose = 0;
};

int usage(int rc = 1) {
  std::cout << "Usage: running arpack (B200) with matrix market files to check for eigen values/vectors.\n\n"
               "  --A F:            file name of matrix A such that A X = lambda X. (standard)   default: A.mtx\n"
               "  --B F:            file name of matrix B such that A X = lambda B X. (generalized) default: B.mtx with --genPb\n"
               "  --nbEV:           number of eigen values/vectors to compute.                 default: 1\n"
               "  --nbCV:           number of columns of the matrix V.                         default: 2*nbEV+1\n"
               "  --genPb:          generalized problem.                                       default: standard problem\n"
               "  --nonSymPb:       non symmetric problem (<=> use dn[ae]upd).                 default: symmetric (ds[ae]upd)\n"
               "  --simplePrec:     use simple precision ([s]*upd).                            default: double precision\n"
               "  --mag M:          LM, SM, LR, SR, LA, SA, LI, SI.                            default: LM\n"
               "  --shiftReal S:    real shift sigma = S.                                      default: 0\n"
               "  --shiftImag S:    imaginary shift (complex problems only: not built).\n"
               "  --invert:         invert mode (accepted; as in the reference it only shows in the OPT line).\n"
               "  --tol T:          tolerance T.                                               default: 1.e-06\n"
               "  --maxIt M:        maximum iterations M.                                      default: 100\n"
               "  --schur:          compute Schur vectors (howmny = 'P').\n"
               "  --slv S:          inner solver for modes 2/3: BiCG (BiCGSTAB) or CG.         default: BiCG\n"
               "                    LU QR LLT LDLT (Eigen direct solvers) are not built.\n"
               "  --slvItrTol T:    solver tolerance.                                          default: 1.e-6\n"
               "  --slvItrMaxIt M:  solver maximum iterations.                                 default: 100\n"
               "  --slvItrPC PC:    preconditioners are not built (only the default, none, is available).\n"
               "  --noCheck:        do not check the eigen pairs.\n"
               "  --verbose N:      verbosity.\n"
               "  --restart:        restart from arpackSolver.resid.out / arpackSolver.v.out.\n"
               "  --registered:     (extension) register the CSR operator with the library: one *aupd call per solve.\n"
               "  --cpxPb, --dense: not built.\n";
  return rc;
}

#define CK(expr)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      std::cerr << "Error: " #expr ": " << cudaGetErrorString(e_) << std::endl;                   \
      std::exit(1);                                                                               \
    }                                                                                             \
  } while (0)

// ---- host CSR ------------------------------------------------------------------------------------
struct Csr {
  int n = 0, m = 0;
  std::vector<int> rowptr, col;
  std::vector<double> val;
};

int read_csr(const std::string& file, Csr& A) {
  int n = 0, m = 0;
  long long nnz = 0;
  int *rp = nullptr, *co = nullptr;
  double* va = nullptr;
  if (ab200_mm_read_csr(file.c_str(), &n, &m, &nnz, &rp, &co, &va) != 0) return 1;
  A.n = n;
  A.m = m;
  A.rowptr.assign(rp, rp + n + 1);
  A.col.assign(co, co + nnz);
  A.val.assign(va, va + nnz);
  ab200_mm_free(rp);
  ab200_mm_free(co);
  ab200_mm_free(va);
  return 0;
}

// C = A + alpha * B (same shape), merged row by row
Csr csr_add(const Csr& A, double alpha, const Csr& B) {
  Csr C;
  C.n = A.n;
  C.m = A.m;
  C.rowptr.assign(A.n + 1, 0);
  for (int r = 0; r < A.n; ++r) {
    int p = A.rowptr[r], q = B.rowptr[r];
    const int pe = A.rowptr[r + 1], qe = B.rowptr[r + 1];
    while (p < pe || q < qe) {
      if (q >= qe || (p < pe && A.col[p] < B.col[q])) { C.col.push_back(A.col[p]); C.val.push_back(A.val[p]); ++p; }
      else if (p >= pe || B.col[q] < A.col[p]) { C.col.push_back(B.col[q]); C.val.push_back(alpha * B.val[q]); ++q; }
      else { C.col.push_back(A.col[p]); C.val.push_back(A.val[p] + alpha * B.val[q]); ++p; ++q; }
    }
    C.rowptr[r + 1] = (int)C.col.size();
  }
  return C;
}

Csr csr_identity(int n) {
  Csr I;
  I.n = I.m = n;
  I.rowptr.resize(n + 1);
  I.col.resize(n);
  I.val.assign(n, 1.0);
  for (int r = 0; r <= n; ++r) I.rowptr[r] = r;
  for (int r = 0; r < n; ++r) I.col[r] = r;
  return I;
}

void host_spmv(const Csr& A, const std::vector<std::complex<double>>& x, std::vector<std::complex<double>>& y) {
  y.assign(A.n, 0.0);
  for (int r = 0; r < A.n; ++r) {
    std::complex<double> s = 0.0;
    for (int p = A.rowptr[r]; p < A.rowptr[r + 1]; ++p) s += A.val[p] * x[A.col[p]];
    y[r] = s;
  }
}

// ---- device CSR + the few vector kernels the inner solvers need --------------------------------------
template <typename T>
struct DevCsr {
  int n = 0;
  long long nnz = 0;
  int *rowptr = nullptr, *col = nullptr;
  T* val = nullptr;
  void upload(const Csr& A) {
    n = A.n;
    nnz = (long long)A.val.size();
    std::vector<T> v(A.val.begin(), A.val.end());
    CK(cudaMalloc(&rowptr, sizeof(int) * (n + 1)));
    CK(cudaMalloc(&col, sizeof(int) * (nnz ? nnz : 1)));
    CK(cudaMalloc(&val, sizeof(T) * (nnz ? nnz : 1)));
    CK(cudaMemcpy(rowptr, A.rowptr.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(col, A.col.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(val, v.data(), sizeof(T) * nnz, cudaMemcpyHostToDevice));
  }
  void apply(const T* x, T* y) const {
    int rc;
    if (std::is_same<T, double>::value)
      rc = ab200_csr_spmv_f64(n, rowptr, col, (const double*)val, (const double*)x, (double*)y);
    else
      rc = ab200_csr_spmv_f32(n, rowptr, col, (const float*)val, (const float*)x, (float*)y);
    if (rc != 0) { std::cerr << "Error: SpMV KO" << std::endl; std::exit(1); }
  }
};

template <typename T>
__global__ void k_axpby(int n, T a, const T* x, T b, T* y) {  // y = a x + b y
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i] + b * y[i];
}
template <typename T>
__global__ void k_dot_partial(int n, const T* x, const T* y, double* partial) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (double)x[i] * (double)y[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

template <typename T>
struct Vec {
  cudaStream_t s;
  double* partial = nullptr;
  double* partial_h = nullptr;
  static constexpr int kGrid = 296;
  explicit Vec(cudaStream_t st) : s(st) {
    CK(cudaMalloc(&partial, sizeof(double) * kGrid));
    CK(cudaMallocHost(&partial_h, sizeof(double) * kGrid));
  }
  int grid(int n) const { int g = (n + 255) / 256; return g > kGrid ? kGrid : (g < 1 ? 1 : g); }
  double dot(int n, const T* x, const T* y) {
    const int g = grid(n);
    k_dot_partial<T><<<g, 256, 0, s>>>(n, x, y, partial);
    CK(cudaMemcpyAsync(partial_h, partial, sizeof(double) * g, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    double t = 0.0;
    for (int i = 0; i < g; ++i) t += partial_h[i];
    return t;
  }
  void axpby(int n, T a, const T* x, T b, T* y) { k_axpby<T><<<grid(n), 256, 0, s>>>(n, a, x, b, y); }
  void copy(int n, const T* x, T* y) { CK(cudaMemcpyAsync(y, x, sizeof(T) * n, cudaMemcpyDeviceToDevice, s)); }
  void zero(int n, T* x) { CK(cudaMemsetAsync(x, 0, sizeof(T) * n, s)); }
};

// x = M^-1 b by CG (symmetric positive definite M) or BiCGSTAB; x0 = 0 (Eigen's solve() starts from zero too)
template <typename T>
struct InnerSolver {
  const DevCsr<T>* M = nullptr;
  bool cg = false;
  double tol = 1e-6;
  int maxit = 100;
  Vec<T>* vec = nullptr;
  T *r = nullptr, *p = nullptr, *q = nullptr, *rh = nullptr, *sv = nullptr, *t = nullptr;
  long long iterations = 0;
  void init(const DevCsr<T>* m, bool use_cg, double tolerance, int maxIt, Vec<T>* v) {
    M = m; cg = use_cg; tol = tolerance; maxit = maxIt; vec = v;
    for (T** b : {&r, &p, &q, &rh, &sv, &t}) CK(cudaMalloc(b, sizeof(T) * (M->n ? M->n : 1)));
  }
  void solve(const T* b, T* x) {
    const int n = M->n;
    const double bnorm = std::sqrt(vec->dot(n, b, b));
    vec->zero(n, x);
    if (bnorm == 0.0) return;
    vec->copy(n, b, r);
    if (cg) {
      vec->copy(n, r, p);
      double rr = bnorm * bnorm;
      for (int it = 0; it < maxit; ++it) {
        M->apply(p, q);
        const double alpha = rr / vec->dot(n, p, q);
        vec->axpby(n, (T)alpha, p, (T)1, x);
        vec->axpby(n, (T)-alpha, q, (T)1, r);
        const double rr1 = vec->dot(n, r, r);
        ++iterations;
        if (std::sqrt(rr1) <= tol * bnorm) break;
        vec->axpby(n, (T)1, r, (T)(rr1 / rr), p);
        rr = rr1;
      }
    } else {
      vec->copy(n, r, rh);
      double rho = 1.0, alpha = 1.0, omega = 1.0;
      vec->zero(n, p);
      vec->zero(n, q);  // q = v of the usual notation
      for (int it = 0; it < maxit; ++it) {
        const double rho1 = vec->dot(n, rh, r);
        if (rho1 == 0.0) break;
        const double beta = (rho1 / rho) * (alpha / omega);
        vec->axpby(n, (T)-omega, q, (T)1, p);     // p = p - omega v
        vec->axpby(n, (T)1, r, (T)beta, p);       // p = r + beta p
        M->apply(p, q);
        alpha = rho1 / vec->dot(n, rh, q);
        vec->copy(n, r, sv);
        vec->axpby(n, (T)-alpha, q, (T)1, sv);    // s = r - alpha v
        vec->axpby(n, (T)alpha, p, (T)1, x);
        ++iterations;
        if (std::sqrt(vec->dot(n, sv, sv)) <= tol * bnorm) break;
        M->apply(sv, t);
        const double tt = vec->dot(n, t, t);
        omega = tt > 0.0 ? vec->dot(n, t, sv) / tt : 0.0;
        vec->axpby(n, (T)omega, sv, (T)1, x);
        vec->copy(n, sv, r);
        vec->axpby(n, (T)-omega, t, (T)1, r);
        if (std::sqrt(vec->dot(n, r, r)) <= tol * bnorm) break;
        if (omega == 0.0) break;
        rho = rho1;
      }
    }
  }
};

// ---- the ICB entry points by precision ------------------------------------------------------------
void aupd(bool sym, int* ido, const char* bmat, int n, const char* which, int nev, double tol, double* resid, int ncv,
          double* v, int ldv, int* iparam, int* ipntr, double* workd, double* workl, int lworkl, int* info) {
  if (sym) dsaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
  else dnaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void aupd(bool sym, int* ido, const char* bmat, int n, const char* which, int nev, float tol, float* resid, int ncv,
          float* v, int ldv, int* iparam, int* ipntr, float* workd, float* workl, int lworkl, int* info) {
  if (sym) ssaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
  else snaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void seupd(int rvec, const char* howmny, const int* select, double* d, double* z, int ldz, double sigma,
           const char* bmat, int n, const char* which, int nev, double tol, double* resid, int ncv, double* v, int ldv,
           int* iparam, int* ipntr, double* workd, double* workl, int lworkl, int* info) {
  dseupd_c(rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd,
           workl, lworkl, info);
}
void seupd(int rvec, const char* howmny, const int* select, float* d, float* z, int ldz, float sigma, const char* bmat,
           int n, const char* which, int nev, float tol, float* resid, int ncv, float* v, int ldv, int* iparam,
           int* ipntr, float* workd, float* workl, int lworkl, int* info) {
  sseupd_c(rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd,
           workl, lworkl, info);
}
void neupd(int rvec, const char* howmny, const int* select, double* dr, double* di, double* z, int ldz, double sr,
           double si, double* workev, const char* bmat, int n, const char* which, int nev, double tol, double* resid,
           int ncv, double* v, int ldv, int* iparam, int* ipntr, double* workd, double* workl, int lworkl, int* info) {
  dneupd_c(rvec, howmny, select, dr, di, z, ldz, sr, si, workev, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam,
           ipntr, workd, workl, lworkl, info);
}
void neupd(int rvec, const char* howmny, const int* select, float* dr, float* di, float* z, int ldz, float sr, float si,
           float* workev, const char* bmat, int n, const char* which, int nev, float tol, float* resid, int ncv, float* v,
           int ldv, int* iparam, int* ipntr, float* workd, float* workl, int lworkl, int* info) {
  sneupd_c(rvec, howmny, select, dr, di, z, ldz, sr, si, workev, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam,
           ipntr, workd, workl, lworkl, info);
}

struct Output {
  int nbVal = 0, mode = 0, nbIt = 0;
  double imsTime = 0.0, rciTime = 0.0;
};

double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
  return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - t0).count() /
         1.e6;
}

template <typename T>
int run(Options& opt, Output& out) {
  // ---- matrices (arpackmm.cpp:625-660) ----
  Csr A, B;
  auto t0 = std::chrono::high_resolution_clock::now();
  if (read_csr(opt.fileA, A) != 0) { std::cerr << "Error: read A KO" << std::endl; return 1; }
  std::cout << "\nINP: create A " << seconds_since(t0) << " s" << std::endl;
  if (A.n != A.m) { std::cerr << "Error: A must be square" << std::endl; return 1; }
  if (!opt.stdPb) {
    t0 = std::chrono::high_resolution_clock::now();
    if (read_csr(opt.fileB, B) != 0) { std::cerr << "Error: read B KO" << std::endl; return 1; }
    std::cout << "\nINP: create B " << seconds_since(t0) << " s" << std::endl;
    if (A.n != B.n) { std::cerr << "Error: A.rows() != B.rows()" << std::endl; return 1; }
    if (A.m != B.m) { std::cerr << "Error: A.cols() != B.cols()" << std::endl; return 1; }
  }
  const int n = A.n;
  int nbCV = opt.nbCV;
  if (nbCV > n) nbCV = n;  // cut-off arpack workspace dim (arpackSolver.hpp:235)
  const int nev = opt.nbEV;

  // ---- problem transformation (arpackSolver.hpp:248-270) ----
  const double eps = std::numeric_limits<T>::epsilon();
  const bool shiftReal = std::fabs(opt.sigmaReal) > eps, shiftImag = std::fabs(opt.sigmaImag) > eps;
  if (shiftImag) { std::cerr << "Error: an imaginary shift needs a complex problem (--cpxPb): not built" << std::endl; return 1; }
  bool backTransform = false;
  int mode;
  Csr Aop = A;
  if (opt.stdPb) {
    mode = 1;
    if (shiftReal) { Aop = csr_add(A, -opt.sigmaReal, csr_identity(n)); backTransform = true; }
  } else {
    mode = shiftReal ? 3 : 2;
  }
  out.mode = mode;
  if (opt.verbose >= 1) std::cout << "\narpackSolver:\n\nmode " << mode << ", backTransform " << (backTransform ? "yes" : "no") << std::endl;

  cudaStream_t stream = (cudaStream_t)ab200_get_stream();
  DevCsr<T> dA, dB, dS;
  dA.upload(Aop);
  Vec<T> vec(stream);
  InnerSolver<T> solver;
  t0 = std::chrono::high_resolution_clock::now();
  if (mode >= 2) {
    dB.upload(B);
    if (mode == 2) {
      solver.init(&dB, opt.slv == "CG", opt.slvItrTol, opt.slvItrMaxIt, &vec);
    } else {
      dS.upload(csr_add(A, -opt.sigmaReal, B));
      solver.init(&dS, opt.slv == "CG", opt.slvItrTol, opt.slvItrMaxIt, &vec);
    }
  }
  out.imsTime = seconds_since(t0);

  // ---- workspace: resid, v, workd, z in HBM; workl and the small arrays on the host ----
  const int ldv = n;
  T *resid = nullptr, *v = nullptr, *workd = nullptr, *z = nullptr, *scratch = nullptr;
  CK(cudaMalloc(&resid, sizeof(T) * n));
  CK(cudaMalloc(&scratch, sizeof(T) * n));
  CK(cudaMalloc(&v, sizeof(T) * (size_t)ldv * nbCV));
  CK(cudaMalloc(&workd, sizeof(T) * 3 * (size_t)n));
  CK(cudaMalloc(&z, sizeof(T) * (size_t)n * (nev + 1)));  // nbEV+1 for dneupd
  CK(cudaMemset(workd, 0, sizeof(T) * 3 * (size_t)n));
  CK(cudaMemset(z, 0, sizeof(T) * (size_t)n * (nev + 1)));
  {
    std::vector<T> r0(n, (T)eps), v0((size_t)ldv * nbCV, (T)(10. * eps));  // close to, but not, zero (:737-747)
    CK(cudaMemcpy(resid, r0.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v, v0.data(), sizeof(T) * v0.size(), cudaMemcpyHostToDevice));
  }
  const int lworkl = opt.symPb ? nbCV * nbCV + 8 * nbCV : 3 * nbCV * nbCV + 6 * nbCV;
  std::vector<T> workl(lworkl, (T)0);
  int iparam[11] = {0}, ipntr[14] = {0};
  iparam[0] = 1;
  iparam[2] = opt.maxIt;
  iparam[3] = 1;
  iparam[6] = mode;
  int ido = 0, info = 0;
  const char* bmat = (mode == 1) ? "I" : "G";
  const char* which = opt.mag.c_str();

  if (opt.restart) {  // arpackSolver.hpp:768-773
    info = 1;
    std::vector<double> rr(n), vv((size_t)ldv * nbCV);
    if (ab200_restart_load_f64("arpackSolver.resid.out", n, rr.data(), 0) != 0) { std::cerr << "Error: bad restart (resid)" << std::endl; return 1; }
    if (ab200_restart_load_f64("arpackSolver.v.out", (long long)ldv * nbCV, vv.data(), 1) != 0) { std::cerr << "Error: bad restart (v)" << std::endl; return 1; }
    std::vector<T> rt(rr.begin(), rr.end()), vt(vv.begin(), vv.end());
    CK(cudaMemcpy(resid, rt.data(), sizeof(T) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v, vt.data(), sizeof(T) * vt.size(), cudaMemcpyHostToDevice));
    if (opt.verbose >= 1) std::cout << "\narpackSolver:\n\narpackSolver.resid.out: restart OK\n\narpackSolver.v.out: restart OK" << std::endl;
  }
  if (opt.registered && mode == 1) {
    int rc;
    if (std::is_same<T, double>::value)
      rc = ab200_register_csr_op_f64(workl.data(), n, dA.nnz, dA.rowptr, dA.col, (const double*)dA.val);
    else
      rc = ab200_register_csr_op_f32(workl.data(), n, dA.nnz, dA.rowptr, dA.col, (const float*)dA.val);
    if (rc != 0) { std::cerr << "Error: operator registration KO" << std::endl; return 1; }
  }

  // ---- reverse communication loop (arpackSolver.hpp:787-846) ----
  long long handoffs = 0;
  do {
    aupd(opt.symPb, &ido, bmat, n, which, nev, (T)opt.tol, resid, nbCV, v, ldv, iparam, ipntr, workd, workl.data(), lworkl,
         &info);
    if (info == 1) std::cerr << "Error: [dz][sn]aupd - KO: maximum number of iterations taken. Increase --maxIt..." << std::endl;
    if (info == 3) std::cerr << "Error: [dz][sn]aupd - KO: no shifts could be applied. Increase --nbCV..." << std::endl;
    if (info == -9) std::cerr << "Error: [dz][sn]aupd - KO: starting vector is zero. Retry: play with shift..." << std::endl;
    if (info < 0) { std::cerr << "Error: [dz][sn]aupd - KO with info " << info << ", nbIt " << iparam[2] << std::endl; return 1; }
    auto t1 = std::chrono::high_resolution_clock::now();
    T* X = workd + ipntr[0] - 1;
    T* Y = workd + ipntr[1] - 1;
    if (ido == -1 || ido == 1) {
      ++handoffs;
      if (mode == 1) {
        dA.apply(X, Y);
      } else if (mode == 2) {
        dA.apply(X, scratch);                      // A x
        if (ido == 1 && opt.symPb) vec.copy(n, scratch, X);  // remark 5 of dsaupd: x <- A x
        solver.solve(scratch, Y);                  // y = B^-1 A x
      } else {
        if (ido == -1) {
          dB.apply(X, scratch);
          solver.solve(scratch, Y);                // y = (A - sigma B)^-1 B x
        } else {
          solver.solve(workd + ipntr[2] - 1, Y);   // B x is provided
        }
      }
    } else if (ido == 2) {
      if (mode == 1) vec.copy(n, X, Y);
      else dB.apply(X, Y);
    } else if (ido != 99) {
      std::cerr << "Error: unexpected ido " << ido << " - KO" << std::endl;
      return 1;
    }
    out.rciTime += seconds_since(t1);
  } while (ido != 99);
  out.nbIt = iparam[2];

  // ---- eigen pairs (arpackSolver.hpp:848-866) ----
  const char* howmny = opt.schur ? "P" : "A";
  std::vector<int> select(nbCV, 1);
  std::vector<std::complex<double>> vals;
  int nconv = iparam[4];
  int ierr = 0;
  std::vector<T> d(nev + 1, 0), di(nev + 1, 0), workev(3 * nbCV, 0);
  if (opt.symPb) {
    seupd(1, howmny, select.data(), d.data(), z, n, (T)opt.sigmaReal, bmat, n, which, nev, (T)opt.tol, resid, nbCV, v, ldv,
          iparam, ipntr, workd, workl.data(), lworkl, &ierr);
  } else {
    neupd(1, howmny, select.data(), d.data(), di.data(), z, n, (T)opt.sigmaReal, (T)0, workev.data(), bmat, n, which, nev,
          (T)opt.tol, resid, nbCV, v, ldv, iparam, ipntr, workd, workl.data(), lworkl, &ierr);
  }
  if (ierr < 0) { std::cerr << "Error: [dz][sn]eupd - KO with info " << ierr << std::endl; std::cerr << "Error: bad arpack eupd" << std::endl; return 1; }
  nconv = iparam[4];
  if (nconv > nev + (opt.symPb ? 0 : 1)) nconv = nev + (opt.symPb ? 0 : 1);
  for (int k = 0; k < nconv; ++k) vals.emplace_back((double)d[k], opt.symPb ? 0.0 : (double)di[k]);
  if (backTransform)
    for (auto& l : vals) l += opt.sigmaReal;
  out.nbVal = (int)vals.size();

  // ---- dump for a later --restart (arpackmm always dumps: arpackmm.cpp:614) ----
  {
    std::vector<T> rt(n), vt((size_t)ldv * nbCV);
    CK(cudaMemcpy(rt.data(), resid, sizeof(T) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vt.data(), v, sizeof(T) * vt.size(), cudaMemcpyDeviceToHost));
    std::vector<double> rr(rt.begin(), rt.end()), vv(vt.begin(), vt.end());
    ab200_restart_save_f64("arpackSolver.resid.out", n, rr.data());
    ab200_restart_save_f64("arpackSolver.v.out", (long long)ldv * nbCV, vv.data());
  }

  // ---- check (arpackSolver.hpp:297-352): || A v - lambda B v || <= sqrt(tol) ----
  const std::string rs = opt.schur ? "Schur" : "Ritz";
  if (vals.empty() && opt.check) { std::cerr << "Error: no " << rs << " value / vector found" << std::endl; return 1; }
  for (size_t i = 0; i < vals.size() && opt.verbose >= 1; ++i)
    std::cout << "\narpackSolver:\n\n" << rs << " value " << i << ": " << vals[i] << std::endl;
  if (opt.check && opt.schur) {
    std::cout << "\narpackSolver:\n\ncheck skipped: " << (opt.symPb ? "dseupd returns no vectors for howmny = 'P'"
                                                                    : "Schur vectors are not eigenvectors") << std::endl;
  } else if (opt.check) {
    std::vector<T> zh((size_t)n * (nev + 1));
    CK(cudaMemcpy(zh.data(), z, sizeof(T) * zh.size(), cudaMemcpyDeviceToHost));
    const double dTol = std::sqrt(opt.tol);
    std::vector<std::complex<double>> V(n), left, right;
    auto check_one = [&](size_t i, size_t kre, bool cplx, double sgn) -> bool {
      for (int r = 0; r < n; ++r) {
        const double re = (double)zh[kre * n + r];
        const double im = cplx ? sgn * (double)zh[(kre + 1) * n + r] : 0.0;
        V[r] = {re, im};
      }
      host_spmv(A, V, left);
      if (opt.stdPb) right = V;
      else host_spmv(B, V, right);
      double diff = 0.0, vn = 0.0;
      for (int r = 0; r < n; ++r) { diff += std::norm(left[r] - vals[i] * right[r]); vn += std::norm(V[r]); }
      diff = std::sqrt(diff);
      if (!(diff <= dTol)) {
        std::cerr << "\nError: bad vector " << i << " (norm " << std::sqrt(vn) << "):\n\nError: diff (norm " << diff
                  << ", tol " << dTol << ")" << std::endl;
        std::cerr << "Error: check KO" << std::endl;
        return false;
      }
      if (opt.verbose >= 1)
        std::cout << "\narpackSolver:\n\n" << rs << " value/vector " << i << ": check OK, diff (norm " << diff << ", tol "
                  << dTol << ")" << std::endl;
      return true;
    };
    // real eigenvalue: column i; conjugate pair (i, i+1): columns hold (Re, Im) of the first member's vector
    // (dneupd.f:84-96), the second member's vector is its conjugate
    for (size_t i = 0; i < vals.size();) {
      const bool cplx = !opt.symPb && vals[i].imag() != 0.0;
      if (!cplx) {
        if (!check_one(i, i, false, 1.0)) return 1;
        ++i;
      } else if (i + 1 < vals.size()) {
        if (!check_one(i, i, true, 1.0) || !check_one(i + 1, i, true, -1.0)) return 1;
        i += 2;
      } else {
        ++i;  // second member of the pair was not returned: nothing to pair the column with
      }
    }
  }
  if (opt.verbose >= 1)
    std::cout << "\nOUT: OP*x hand-offs " << handoffs << ", inner solver iterations " << solver.iterations << std::endl;
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Options opt;
  bool nbCVGiven = false;
  auto need = [&](int& a, const std::string& clo) -> const char* {
    if (++a >= argc) { std::cerr << "Error: bad " << clo << " - need argument" << std::endl; std::exit(usage()); }
    return argv[a];
  };
  auto num = [&](const char* s, const std::string& clo, auto& outv) {
    std::stringstream ss(s);
    ss >> outv;
    if (!ss) { std::cerr << "Error: bad " << clo << " - bad argument" << std::endl; std::exit(usage()); }
  };
  for (int a = 1; a < argc; a++) {
    const std::string clo = argv[a];
    if (clo == "--help" || clo == "-h") return usage(0);
    else if (clo == "--A") opt.fileA = need(a, clo);
    else if (clo == "--B") opt.fileB = need(a, clo);
    else if (clo == "--dense") { need(a, clo); opt.dense = true; }
    else if (clo == "--nbEV") { num(need(a, clo), clo, opt.nbEV); if (!nbCVGiven) opt.nbCV = 2 * opt.nbEV + 1; }
    else if (clo == "--nbCV") { num(need(a, clo), clo, opt.nbCV); nbCVGiven = true; }
    else if (clo == "--genPb") { opt.stdPb = false; if (opt.fileB == "N.A.") opt.fileB = "B.mtx"; }
    else if (clo == "--nonSymPb") opt.symPb = false;
    else if (clo == "--cpxPb") { opt.symPb = false; opt.cpxPb = true; }
    else if (clo == "--simplePrec") opt.simplePrec = true;
    else if (clo == "--mag") {
      opt.mag = need(a, clo);
      const char* ok[] = {"LM", "SM", "LR", "SR", "LA", "SA", "LI", "SI"};
      bool good = false;
      for (const char* m : ok) good = good || opt.mag == m;
      if (!good) { std::cerr << "Error: bad " << clo << " - bad argument" << std::endl; return usage(); }
    }
    else if (clo == "--shiftReal") { num(need(a, clo), clo, opt.sigmaReal); opt.shiftReal = true; }
    else if (clo == "--shiftImag") { num(need(a, clo), clo, opt.sigmaImag); opt.shiftImag = true; }
    else if (clo == "--invert") opt.invert = true;
    else if (clo == "--tol") num(need(a, clo), clo, opt.tol);
    else if (clo == "--maxIt") num(need(a, clo), clo, opt.maxIt);
    else if (clo == "--schur") opt.schur = true;
    else if (clo == "--slv") opt.slv = need(a, clo);
    else if (clo == "--slvItrTol") num(need(a, clo), clo, opt.slvItrTol);
    else if (clo == "--slvItrMaxIt") num(need(a, clo), clo, opt.slvItrMaxIt);
    else if (clo == "--slvItrPC") { opt.slvItrPC = need(a, clo); opt.slvPCGiven = true; }
    else if (clo == "--slvDrtPivot" || clo == "--slvDrtOffset" || clo == "--slvDrtScale") need(a, clo);
    else if (clo == "--noCheck") opt.check = false;
    else if (clo == "--verbose") num(need(a, clo), clo, opt.verbose);
    else if (clo == "--debug") { int lvl = 0; num(need(a, clo), clo, lvl); }
    else if (clo == "--restart") opt.restart = true;
    else if (clo == "--registered") opt.registered = true;
    else { std::cerr << "Error: unknown option " << clo << std::endl; return usage(); }
  }
  if (opt.cpxPb) { std::cerr << "Error: --cpxPb (zn[ae]upd) is not built in arpackmm_b200" << std::endl; return 1; }
  if (opt.dense) { std::cerr << "Error: --dense is not built in arpackmm_b200 (sparse CSR only)" << std::endl; return 1; }
  if (opt.slv != "BiCG" && opt.slv != "CG") { std::cerr << "Error: --slv " << opt.slv << " is not built in arpackmm_b200 (BiCG or CG)" << std::endl; return 1; }
  if (opt.slvPCGiven) { std::cerr << "Error: --slvItrPC is not built in arpackmm_b200 (no preconditioner)" << std::endl; return 1; }
  if (opt.symPb && (opt.mag == "LR" || opt.mag == "SR" || opt.mag == "LI" || opt.mag == "SI")) {
    std::cerr << "Error: bad --mag for a symmetric problem" << std::endl;
    return 1;
  }
  if (ab200_device_count() <= 0) { std::cerr << "Error: no CUDA device; arpackmm_b200 has no CPU path" << std::endl; return 1; }

  std::cout << "OPT: A " << opt.fileA << ", B " << opt.fileB << ", nbEV " << opt.nbEV << ", nbCV " << opt.nbCV << ", stdPb "
            << (opt.stdPb ? "yes" : "no") << ", symPb " << (opt.symPb ? "yes" : "no") << ", simplePrec "
            << (opt.simplePrec ? "yes" : "no") << ", mag " << opt.mag << ", shiftReal " << opt.sigmaReal << ", invert "
            << (opt.invert ? "yes" : "no") << ", tol " << opt.tol << ", maxIt " << opt.maxIt << ", "
            << (opt.schur ? "Schur" : "Ritz") << " vectors, slv " << opt.slv << ", check " << (opt.check ? "yes" : "no")
            << ", restart " << (opt.restart ? "yes" : "no") << ", registered " << (opt.registered ? "yes" : "no") << std::endl;

  std::cout.precision(15);  // the reference prints 6 digits; more are harmless and let scripts compare values
  auto start = std::chrono::high_resolution_clock::now();
  Output out;
  const int rc = opt.simplePrec ? run<float>(opt, out) : run<double>(opt, out);
  if (rc != 0) { std::cerr << "Error: arpack solve KO" << std::endl; return rc; }
  std::cout << "\nOUT: mode " << out.mode << ", nb EV found " << out.nbVal << ", nb iterations " << out.nbIt << std::endl;
  std::cout << "OUT: init mode solver " << out.imsTime << " s, RCI time " << out.rciTime << " s" << std::endl;
  std::cout << "OUT: full time " << seconds_since(start) << " s" << std::endl;
  return 0;
}
