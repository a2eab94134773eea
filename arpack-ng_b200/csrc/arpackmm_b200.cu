// arpackmm_b200 -- the B200 twin of the reference's Matrix-Market driver (EXAMPLES/MATRIX_MARKET/arpackmm.cpp,
// arpackSolver.hpp).  Same command-line options, same input files, same "OUT:" / "STAT:" lines, same residual check
// (||A v - lambda B v|| <= sqrt(tol), arpackSolver.hpp:297-352) and the same --restart dump files; the eigen-solve
// runs through libarpack_b200's C-ABI with device-resident arrays, every matrix product and every inner solve runs on
// the GPU.
//
// What is carried over from the reference tool, and how:
//   * real problems in double or single precision (--simplePrec), symmetric (ds*upd) or not (--nonSymPb, dn*upd), and
//     complex problems (--cpxPb, zn*upd / cn*upd; files with "(re, im)" entries);
//   * standard problems: mode 1, a real shift is applied as A - sigma I and undone afterwards (arpackSolver.hpp:255-262);
//   * generalised problems (--genPb/--B): mode 2 (OP = B^-1 A) and, with a shift, mode 3 (OP = (A - sigma B)^-1 B);
//   * --slv BiCG / CG: BiCGSTAB and conjugate gradients with the iteration of Eigen's solvers (right-preconditioned
//     BiCGSTAB with its restart rule; same stopping test ||r|| <= tol ||b||), preconditioned by --slvItrPC Diag (Jacobi,
//     the default) or ILU#D#F (dual-threshold incomplete LU: drop tolerance D, fill factor F; set up on the host like
//     any other input, applied on the GPU as two level-scheduled sparse triangular solves);
//   * --slv LU / QR / LLT / LDLT, sparse: cuSOLVER's device sparse factorisations -- QR (csrqr) for LU and QR, Cholesky
//     (csrchol) for LLT and LDLT, factorised once, solved once per OP*x; --slvDrtPivot is the singularity threshold,
//     --slvDrtOffset / --slvDrtScale adjust the diagonal before a Cholesky factorisation (d_ii <- offset + scale d_ii,
//     Eigen's SimplicialCholesky::setShift).  cuSOLVER has no device sparse LU and no sparse LDL^T: LU shares the QR
//     factorisation, LDLT the Cholesky one (falling back to QR when a non-positive pivot shows up);
//   * --dense RR: A and B as dense matrices; products are dense matrix-vector kernels, the solvers cuSOLVER's dense
//     getrf (LU), geqrf (QR) and potrf (LLT, LDLT: falls back to getrf when the matrix is not positive definite).
//     cuSOLVER pivots partially: RR = true and RR = false select the same factorisations;
//   * extension: --registered hands the CSR matrix to the library (ab200_register_csr_op_*), one *aupd call per solve.
// Unknown options are ignored like the reference's parser does (arpackmm.sh passes a bare "LA"), with a warning.
#include <cublas_v2.h>
#include <cuda/std/complex>
#include <cuda_runtime.h>
#include <cusolverDn.h>
#include <cusolverSp.h>
#include <cusolverSp_LOWLEVEL_PREVIEW.h>
#include <cusparse.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <set>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/arpack_b200.h"

namespace {

using cfloat = cuda::std::complex<float>;
using cdouble = cuda::std::complex<double>;
using hcomplex = std::complex<double>;

struct Options {
  std::string fileA = "A.mtx", fileB = "N.A.";
  int nbEV = 1, nbCV = 3;
  bool stdPb = true, symPb = true, cpxPb = false, simplePrec = false, dense = false, denseRR = true;
  std::string mag = "LM";
  bool shiftReal = false, shiftImag = false, invert = false;
  double sigmaReal = 0.0, sigmaImag = 0.0, tol = 1.e-6;
  int maxIt = 100;
  bool schur = false;
  std::string slv = "BiCG", slvItrPC = "Diag";
  double slvItrTol = 1.e-6;
  int slvItrMaxIt = 100;
  double slvDrtPivot = 1.e-6, slvDrtOffset = 0.0, slvDrtScale = 1.0;
  bool check = true, restart = false, registered = false;
  int verbose = 0, debug = 0;
  bool direct() const {
    return slv.find("LU") != std::string::npos || slv.find("QR") != std::string::npos ||
           slv.find("LLT") != std::string::npos || slv.find("LDLT") != std::string::npos;
  }
};

int usage(int rc = 1) {
  std::cout << "Usage: running arpack (B200) with matrix market files to check for eigen values/vectors.\n\n"
               "  --A F:            file name of matrix A such that A X = lambda X. (standard)   default: A.mtx\n"
               "  --B F:            file name of matrix B such that A X = lambda B X. (generalized) default: B.mtx with --genPb\n"
               "  --dense RR:       consider A and B as dense matrices (RR = true | false; direct solvers only).\n"
               "  --nbEV:           number of eigen values/vectors to compute.                 default: 1\n"
               "  --nbCV:           number of columns of the matrix V.                         default: 2*nbEV+1\n"
               "  --genPb:          generalized problem.                                       default: standard problem\n"
               "  --nonSymPb:       non symmetric problem (<=> use dn[ae]upd).                 default: symmetric (ds[ae]upd)\n"
               "  --cpxPb:          complex (non symmetric) problem (<=> use zn[ae]upd).\n"
               "  --simplePrec:     use simple precision ([sc]*upd).                           default: double precision\n"
               "  --mag M:          LM, SM, LR, SR, LA, SA, LI, SI.                            default: LM\n"
               "  --shiftReal S:    real shift sigma = S.                                      default: 0\n"
               "  --shiftImag S:    imaginary shift sigma = S.                                 default: 0\n"
               "  --invert:         invert mode (accepted; as in the reference it only shows in the OPT line).\n"
               "  --tol T:          tolerance T.                                               default: 1.e-06\n"
               "  --maxIt M:        maximum iterations M.                                      default: 100\n"
               "  --schur:          compute Schur vectors (howmny = 'P').\n"
               "  --slv S:          solver (needed if arpack mode > 1).                        default: BiCG\n"
               "                      BiCG: iterative, any matrices          CG:   iterative, sym matrices only\n"
               "                      LU:   direct, any matrices             QR:   direct, any matrices\n"
               "                      LLT:  direct, SPD matrices only        LDLT: direct, symmetric positive matrices\n"
               "  --slvItrTol T:    solver tolerance (iterative solvers).                      default: 1.e-6\n"
               "  --slvItrMaxIt M:  solver maximum iterations (iterative solvers).             default: 100\n"
               "  --slvItrPC PC:    preconditioner: Diag (Jacobi) or ILU#D#F (drop tolerance D, fill factor F). default: Diag\n"
               "  --slvDrtPivot P:  singularity threshold of the direct solvers.                default: 1.e-06\n"
               "  --slvDrtOffset O: Cholesky diagonal offset.                                  default: 0.\n"
               "  --slvDrtScale S:  Cholesky diagonal scale.                                   default: 1.\n"
               "  --noCheck:        do not check the eigen pairs.\n"
               "  --verbose N:      verbosity.\n"
               "  --debug D:        debug level (up to 3).\n"
               "  --restart:        restart from arpackSolver.resid.out / arpackSolver.v.out.\n"
               "  --registered:     (extension) register the CSR operator with the library: one *aupd call per solve.\n";
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
// cuSOLVER / cuBLAS / cuSPARSE status codes are all "0 = success"
#define CKS(expr)                                                                                 \
  do {                                                                                            \
    const int s_ = (int)(expr);                                                                   \
    if (s_ != 0) {                                                                                \
      std::cerr << "Error: " #expr ": status " << s_ << std::endl;                                \
      std::exit(1);                                                                               \
    }                                                                                             \
  } while (0)

// ---- scalar plumbing: S = float | double | cfloat | cdouble on the device; double | std::complex<double> for the
// host-side matrix arithmetic, the inner solvers' scalars and the check -------------------------------------------
template <typename S> struct ScalarOf { using Real = S; using Host = double; static constexpr bool cplx = false; };
template <> struct ScalarOf<cfloat> { using Real = float; using Host = hcomplex; static constexpr bool cplx = true; };
template <> struct ScalarOf<cdouble> { using Real = double; using Host = hcomplex; static constexpr bool cplx = true; };
template <typename S> using RealOf = typename ScalarOf<S>::Real;
template <typename S> using HostOf = typename ScalarOf<S>::Host;

template <typename S>
S from_host(const HostOf<S>& h) {
  if constexpr (ScalarOf<S>::cplx) return S((RealOf<S>)h.real(), (RealOf<S>)h.imag());
  else return (S)h;
}
template <typename S>
HostOf<S> to_host(const S& s) {
  if constexpr (ScalarOf<S>::cplx) return hcomplex((double)s.real(), (double)s.imag());
  else return (double)s;
}
inline double hreal(double x) { return x; }
inline double hreal(const hcomplex& x) { return x.real(); }
template <typename H> H make_host(double re, double im) {
  if constexpr (std::is_same<H, double>::value) return re;
  else return H(re, im);
}

// pointer / value views for the cuSOLVER and cuBLAS prototypes
inline float* cu(float* p) { return p; }
inline const float* cu(const float* p) { return p; }
inline double* cu(double* p) { return p; }
inline const double* cu(const double* p) { return p; }
inline cuComplex* cu(cfloat* p) { return reinterpret_cast<cuComplex*>(p); }
inline const cuComplex* cu(const cfloat* p) { return reinterpret_cast<const cuComplex*>(p); }
inline cuDoubleComplex* cu(cdouble* p) { return reinterpret_cast<cuDoubleComplex*>(p); }
inline const cuDoubleComplex* cu(const cdouble* p) { return reinterpret_cast<const cuDoubleComplex*>(p); }
inline float cuv(float x) { return x; }
inline double cuv(double x) { return x; }
inline cuComplex cuv(cfloat x) { return make_cuComplex(x.real(), x.imag()); }
inline cuDoubleComplex cuv(cdouble x) { return make_cuDoubleComplex(x.real(), x.imag()); }

// cuSOLVER, cuSPARSE and cuBLAS are only needed by the direct solvers; together they are 1.5 GB of shared objects whose
// device code is registered when a process loads them.  They are therefore opened on first use (dlopen), never linked:
// every call goes through DL(name), which resolves the symbol with the prototype of the header's declaration.
void* dl_lookup(const char* name) {
  static void* libs[3] = {nullptr, nullptr, nullptr};
  static bool opened = false;
  if (!opened) {
    opened = true;
    const char* names[3][2] = {{"libcublas.so.12", "libcublas.so"}, {"libcusparse.so.12", "libcusparse.so"},
                               {"libcusolver.so.11", "libcusolver.so"}};
    for (int k = 0; k < 3; ++k) {
      libs[k] = dlopen(names[k][0], RTLD_NOW | RTLD_GLOBAL);
      if (!libs[k]) libs[k] = dlopen(names[k][1], RTLD_NOW | RTLD_GLOBAL);
      if (!libs[k]) { std::cerr << "Error: direct solvers need " << names[k][0] << ": " << dlerror() << std::endl; std::exit(1); }
    }
  }
  for (void* lib : libs)
    if (void* f = dlsym(lib, name)) return f;
  std::cerr << "Error: symbol " << name << " not found in cuBLAS / cuSPARSE / cuSOLVER" << std::endl;
  std::exit(1);
}
#define DL_STR2(x) #x
#define DL_STR(x) DL_STR2(x)  /* cublasCreate is a macro for cublasCreate_v2: stringify the expansion */
#define DL(f) (reinterpret_cast<decltype(&f)>(dl_lookup(DL_STR(f))))

template <typename S, typename FS, typename FD, typename FC, typename FZ>
auto pick(FS s, FD d, FC c, FZ z) {
  if constexpr (std::is_same<S, float>::value) return s;
  else if constexpr (std::is_same<S, double>::value) return d;
  else if constexpr (std::is_same<S, cfloat>::value) return c;
  else return z;
}
#define SP_FN(S, fn) pick<S>(DL(cusolverSpS##fn), DL(cusolverSpD##fn), DL(cusolverSpC##fn), DL(cusolverSpZ##fn))
#define DN_FN(S, fn) pick<S>(DL(cusolverDnS##fn), DL(cusolverDnD##fn), DL(cusolverDnC##fn), DL(cusolverDnZ##fn))

// ---- host CSR ------------------------------------------------------------------------------------
template <typename H>
struct Csr {
  int n = 0, m = 0;
  std::vector<int> rowptr, col;
  std::vector<H> val;
};

int read_csr(const std::string& file, Csr<double>& A) {
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
int read_csr(const std::string& file, Csr<hcomplex>& A) {
  int n = 0, m = 0;
  long long nnz = 0;
  int *rp = nullptr, *co = nullptr;
  double* va = nullptr;
  if (ab200_mm_read_csr_z(file.c_str(), &n, &m, &nnz, &rp, &co, &va) != 0) return 1;
  A.n = n;
  A.m = m;
  A.rowptr.assign(rp, rp + n + 1);
  A.col.assign(co, co + nnz);
  A.val.resize((size_t)nnz);
  for (long long k = 0; k < nnz; ++k) A.val[(size_t)k] = hcomplex(va[2 * k], va[2 * k + 1]);
  ab200_mm_free(rp);
  ab200_mm_free(co);
  ab200_mm_free(va);
  return 0;
}

// C = A + alpha * B (same shape), merged row by row
template <typename H>
Csr<H> csr_add(const Csr<H>& A, H alpha, const Csr<H>& B) {
  Csr<H> C;
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

template <typename H>
Csr<H> csr_identity(int n) {
  Csr<H> I;
  I.n = I.m = n;
  I.rowptr.resize(n + 1);
  I.col.resize(n);
  I.val.assign(n, H(1.0));
  for (int r = 0; r <= n; ++r) I.rowptr[r] = r;
  for (int r = 0; r < n; ++r) I.col[r] = r;
  return I;
}

template <typename H>
void host_spmv(const Csr<H>& A, const std::vector<hcomplex>& x, std::vector<hcomplex>& y) {
  y.assign(A.n, 0.0);
  for (int r = 0; r < A.n; ++r) {
    hcomplex s = 0.0;
    for (int p = A.rowptr[r]; p < A.rowptr[r + 1]; ++p) s += A.val[p] * x[A.col[p]];
    y[r] = s;
  }
}

// ---- device kernels of the tool: complex CSR and dense products, the vector operations of the inner solvers, the
// level-scheduled triangular solve of the ILU preconditioner ------------------------------------------------------
template <typename S>
__device__ inline S warp_sum(S v) {
  if constexpr (ScalarOf<S>::cplx) {
    auto re = v.real(), im = v.imag();
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_down_sync(0xffffffffu, re, o);
      im += __shfl_down_sync(0xffffffffu, im, o);
    }
    return S(re, im);
  } else {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
  }
}

// y = A x, one warp per row (tool-sized problems; the real CSR products go through the library's SpMV instead)
template <typename S>
__global__ void k_csr_rows(int n, const int* __restrict__ rp, const int* __restrict__ ci, const S* __restrict__ v,
                           const S* __restrict__ x, S* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n; r += nwarps) {
    S s = S(0);
    for (int p = rp[r] + lane; p < rp[r + 1]; p += 32) s += v[p] * x[ci[p]];
    s = warp_sum(s);
    if (lane == 0) y[r] = s;
  }
}
// y = A x for a dense row-major n x n matrix, one warp per row
template <typename S>
__global__ void k_dense_rows(int n, const S* __restrict__ a, const S* __restrict__ x, S* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n; r += nwarps) {
    const S* row = a + (size_t)r * n;
    S s = S(0);
    for (int c = lane; c < n; c += 32) s += row[c] * x[c];
    s = warp_sum(s);
    if (lane == 0) y[r] = s;
  }
}
template <typename S>
__global__ void k_axpby(int n, S a, const S* x, S b, S* y) {  // y = a x + b y
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i] + b * y[i];
}
template <typename S>
__global__ void k_mul(int n, const S* d, const S* x, S* y) {  // y = d .* x
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = d[i] * x[i];
}
// partial[2 b], partial[2 b + 1] = real and imaginary part of block b's share of conj(x) . y, in double
template <typename S>
__global__ void k_dot_partial(int n, const S* x, const S* y, double* partial) {
  __shared__ double red_re[256], red_im[256];
  double re = 0.0, im = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if constexpr (ScalarOf<S>::cplx) {
      const double xr = x[i].real(), xi = x[i].imag(), yr = y[i].real(), yi = y[i].imag();
      re += xr * yr + xi * yi;
      im += xr * yi - xi * yr;
    } else {
      re += (double)x[i] * (double)y[i];
    }
  }
  red_re[threadIdx.x] = re;
  red_im[threadIdx.x] = im;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { red_re[threadIdx.x] += red_re[threadIdx.x + o]; red_im[threadIdx.x] += red_im[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = red_re[0]; partial[2 * blockIdx.x + 1] = red_im[0]; }
}
// x = T^-1 b for a triangular CSR matrix T = (strict part) + diagonal, level by level inside ONE thread block: the
// rows of a level only need rows of earlier levels.  dinv = inverse diagonal, or null for a unit diagonal.
template <typename S>
__global__ void __launch_bounds__(1024) k_sptrsv_levels(int nlevels, const int* __restrict__ level_ptr, const int* __restrict__ level_rows,
                                const int* __restrict__ rp, const int* __restrict__ ci, const S* __restrict__ v,
                                const S* __restrict__ dinv, const S* __restrict__ b, S* x) {
  for (int l = 0; l < nlevels; ++l) {
    for (int k = level_ptr[l] + (int)threadIdx.x; k < level_ptr[l + 1]; k += (int)blockDim.x) {
      const int r = level_rows[k];
      S s = b[r];
      for (int p = rp[r]; p < rp[r + 1]; ++p) s -= v[p] * x[ci[p]];
      x[r] = dinv ? s * dinv[r] : s;
    }
    __syncthreads();
  }
}

template <typename S>
struct Vec {
  using H = HostOf<S>;
  cudaStream_t s;
  double* partial = nullptr;
  double* partial_h = nullptr;
  static constexpr int kGrid = 296;
  explicit Vec(cudaStream_t st) : s(st) {
    CK(cudaMalloc(&partial, sizeof(double) * 2 * kGrid));
    CK(cudaMallocHost(&partial_h, sizeof(double) * 2 * kGrid));
  }
  int grid(int n) const { int g = (n + 255) / 256; return g > kGrid ? kGrid : (g < 1 ? 1 : g); }
  H dot(int n, const S* x, const S* y) {  // conj(x) . y
    const int g = grid(n);
    k_dot_partial<S><<<g, 256, 0, s>>>(n, x, y, partial);
    CK(cudaMemcpyAsync(partial_h, partial, sizeof(double) * 2 * g, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    double re = 0.0, im = 0.0;
    for (int i = 0; i < g; ++i) { re += partial_h[2 * i]; im += partial_h[2 * i + 1]; }
    return make_host<H>(re, im);
  }
  double sqnorm(int n, const S* x) { return hreal(dot(n, x, x)); }
  void axpby(int n, H a, const S* x, H b, S* y) { k_axpby<S><<<grid(n), 256, 0, s>>>(n, from_host<S>(a), x, from_host<S>(b), y); }
  void mul(int n, const S* d, const S* x, S* y) { k_mul<S><<<grid(n), 256, 0, s>>>(n, d, x, y); }
  void copy(int n, const S* x, S* y) { CK(cudaMemcpyAsync(y, x, sizeof(S) * n, cudaMemcpyDeviceToDevice, s)); }
  void zero(int n, S* x) { CK(cudaMemsetAsync(x, 0, sizeof(S) * n, s)); }
};

template <typename T>
T* upload(const std::vector<T>& h) {
  T* d = nullptr;
  CK(cudaMalloc(&d, sizeof(T) * (h.empty() ? 1 : h.size())));
  if (!h.empty()) CK(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return d;
}

// ---- a matrix on the device: CSR or dense, y = M x ---------------------------------------------------------
template <typename S>
struct DevMat {
  using H = HostOf<S>;
  int n = 0;
  long long nnz = 0;
  bool dense = false;
  int *rowptr = nullptr, *col = nullptr;
  S* val = nullptr;        // CSR values, or the dense matrix row-major (products)
  S* colmajor = nullptr;   // dense only: column-major copy (what a factorisation takes)
  cudaStream_t stream = nullptr;
  void upload_from(const Csr<H>& A, bool as_dense, cudaStream_t st) {
    n = A.n;
    dense = as_dense;
    stream = st;
    if (!dense) {
      nnz = (long long)A.val.size();
      std::vector<S> v(A.val.size());
      for (size_t k = 0; k < v.size(); ++k) v[k] = from_host<S>(A.val[k]);
      rowptr = upload(A.rowptr);
      col = upload(A.col);
      val = upload(v);
    } else {
      nnz = (long long)n * n;
      std::vector<S> rm((size_t)n * n, S(0)), cm((size_t)n * n, S(0));  // setZero first (arpackSolver.hpp:196)
      for (int r = 0; r < n; ++r)
        for (int p = A.rowptr[r]; p < A.rowptr[r + 1]; ++p) {
          rm[(size_t)r * n + A.col[p]] = from_host<S>(A.val[p]);
          cm[(size_t)A.col[p] * n + r] = from_host<S>(A.val[p]);
        }
      val = upload(rm);
      colmajor = upload(cm);
    }
  }
  void apply(const S* x, S* y) const {
    if (dense) {
      const int g = std::min(1184, std::max(1, (n + 7) / 8));
      k_dense_rows<S><<<g, 256, 0, stream>>>(n, val, x, y);
      CK(cudaGetLastError());
    } else if constexpr (std::is_same<S, double>::value) {
      if (ab200_csr_spmv_f64(n, rowptr, col, val, x, y) != 0) { std::cerr << "Error: SpMV KO" << std::endl; std::exit(1); }
    } else if constexpr (std::is_same<S, float>::value) {
      if (ab200_csr_spmv_f32(n, rowptr, col, val, x, y) != 0) { std::cerr << "Error: SpMV KO" << std::endl; std::exit(1); }
    } else {
      const int g = std::min(1184, std::max(1, (n + 7) / 8));
      k_csr_rows<S><<<g, 256, 0, stream>>>(n, rowptr, col, val, x, y);
      CK(cudaGetLastError());
    }
  }
};

// ---- preconditioners (arpackmm.cpp:790-833: "Diag" or "ILU#D#F") -----------------------------------------
// Dual-threshold incomplete LU in the manner of Eigen's IncompleteLUT (Saad's ILUT): row by row, entries below
// droptol * ||row|| are dropped, at most fill = nnz * fillfactor / n + 1 entries are kept in the L part and in the U
// part of every row.  No fill-reducing permutation.  L is unit lower (strict part stored), U = diag + strict upper.
template <typename H>
struct IlutFactors {
  Csr<H> L, U;
  std::vector<H> dinv;
};
template <typename H>
IlutFactors<H> ilut(const Csr<H>& A, double droptol, int fillfactor) {
  const int n = A.n;
  IlutFactors<H> f;
  f.L.n = f.L.m = f.U.n = f.U.m = n;
  f.L.rowptr.assign(n + 1, 0);
  f.U.rowptr.assign(n + 1, 0);
  f.dinv.assign(n, H(1.0));
  std::vector<H> udiag(n, H(0.0));
  const long long fill = (long long)A.val.size() * fillfactor / std::max(n, 1) + 1;
  std::vector<H> w(n, H(0.0));
  std::vector<char> in_w(n, 0);
  for (int i = 0; i < n; ++i) {
    std::set<int> lower;          // pending columns k < i, ascending
    std::vector<int> touched;
    double rownorm = 0.0;
    for (int p = A.rowptr[i]; p < A.rowptr[i + 1]; ++p) {
      const int c = A.col[p];
      w[c] = A.val[p];
      in_w[c] = 1;
      touched.push_back(c);
      if (c < i) lower.insert(c);
      rownorm += std::norm(hcomplex(A.val[p]));
    }
    rownorm = std::sqrt(rownorm);
    if (rownorm == 0.0) { std::cerr << "Error: ILU - zero row " << i << std::endl; std::exit(1); }
    std::vector<std::pair<int, H>> lrow;
    while (!lower.empty()) {
      const int k = *lower.begin();
      lower.erase(lower.begin());
      const H fact = w[k] * f.dinv[k];
      w[k] = H(0.0);
      if (std::abs(fact) <= droptol * rownorm) continue;  // dropped
      for (int p = f.U.rowptr[k]; p < f.U.rowptr[k + 1]; ++p) {
        const int j = f.U.col[p];
        if (!in_w[j]) { in_w[j] = 1; touched.push_back(j); w[j] = H(0.0); if (j < i) lower.insert(j); }
        w[j] -= fact * f.U.val[p];
      }
      lrow.emplace_back(k, fact);
    }
    std::vector<std::pair<int, H>> urow;
    for (int c : touched)
      if (c > i && std::abs(w[c]) > droptol * rownorm) urow.emplace_back(c, w[c]);
    H d = in_w[i] ? w[i] : H(0.0);
    if (d == H(0.0)) d = H(std::sqrt(droptol) * rownorm);  // IncompleteLUT's rule for a vanished pivot
    if (d == H(0.0)) d = H(rownorm);
    auto keep_largest = [&](std::vector<std::pair<int, H>>& row) {
      if ((long long)row.size() > fill) {
        std::nth_element(row.begin(), row.begin() + fill, row.end(),
                         [](const auto& a, const auto& b) { return std::abs(a.second) > std::abs(b.second); });
        row.resize((size_t)fill);
      }
      std::sort(row.begin(), row.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    };
    keep_largest(lrow);
    keep_largest(urow);
    for (auto& e : lrow) { f.L.col.push_back(e.first); f.L.val.push_back(e.second); }
    for (auto& e : urow) { f.U.col.push_back(e.first); f.U.val.push_back(e.second); }
    f.L.rowptr[i + 1] = (int)f.L.col.size();
    f.U.rowptr[i + 1] = (int)f.U.col.size();
    udiag[i] = d;
    f.dinv[i] = H(1.0) / d;
    for (int c : touched) { w[c] = H(0.0); in_w[c] = 0; }
  }
  return f;
}

// rows grouped by dependency depth: a row's level is one more than the deepest row it reads
template <typename H>
void level_sets(const Csr<H>& T, bool lower, std::vector<int>& level_ptr, std::vector<int>& level_rows) {
  const int n = T.n;
  std::vector<int> depth(n, 0);
  int nlev = 0;
  for (int q = 0; q < n; ++q) {
    const int r = lower ? q : n - 1 - q;
    int d = 0;
    for (int p = T.rowptr[r]; p < T.rowptr[r + 1]; ++p) d = std::max(d, depth[T.col[p]] + 1);
    depth[r] = d;
    nlev = std::max(nlev, d + 1);
  }
  level_ptr.assign(nlev + 1, 0);
  for (int r = 0; r < n; ++r) level_ptr[depth[r] + 1]++;
  for (int l = 0; l < nlev; ++l) level_ptr[l + 1] += level_ptr[l];
  level_rows.resize(n);
  std::vector<int> fillp(level_ptr.begin(), level_ptr.end() - 1);
  for (int r = 0; r < n; ++r) level_rows[fillp[depth[r]]++] = r;
}

template <typename S>
struct Preconditioner {
  using H = HostOf<S>;
  int kind = 0;  // 1 = Jacobi, 2 = ILU
  int n = 0;
  cudaStream_t stream = nullptr;
  S* dinv = nullptr;  // Jacobi: 1 / a_ii (1 where a_ii = 0, Eigen's DiagonalPreconditioner); ILU: 1 / u_ii
  struct Tri { int nlevels = 0; int *level_ptr = nullptr, *level_rows = nullptr, *rp = nullptr, *ci = nullptr; S* v = nullptr; } L, U;
  S* tmp = nullptr;
  void upload_tri(const Csr<H>& T, bool lower, Tri& t) {
    std::vector<int> lp, lr;
    level_sets(T, lower, lp, lr);
    t.nlevels = (int)lp.size() - 1;
    t.level_ptr = upload(lp);
    t.level_rows = upload(lr);
    t.rp = upload(T.rowptr);
    t.ci = upload(T.col);
    std::vector<S> v(T.val.size());
    for (size_t k = 0; k < v.size(); ++k) v[k] = from_host<S>(T.val[k]);
    t.v = upload(v);
  }
  int init(const Csr<H>& M, const std::string& spec, cudaStream_t st, int verbose) {
    n = M.n;
    stream = st;
    std::stringstream clo(spec);
    std::string name;
    std::getline(clo, name, '#');
    std::vector<S> d(n);
    if (name == "Diag") {
      kind = 1;
      for (int r = 0; r < n; ++r) {
        H a = H(0.0);
        for (int p = M.rowptr[r]; p < M.rowptr[r + 1]; ++p)
          if (M.col[p] == r) a = M.val[p];
        d[r] = from_host<S>(a != H(0.0) ? H(1.0) / a : H(1.0));
      }
      dinv = upload(d);
    } else if (name == "ILU") {
      kind = 2;
      double droptol = 1.0;   // arpackmm.cpp:799, 808
      int fillfactor = 2;
      std::string s;
      if (std::getline(clo, s, '#')) { std::stringstream t(s); t >> droptol; }
      if (std::getline(clo, s)) { std::stringstream t(s); t >> fillfactor; }
      IlutFactors<H> f = ilut(M, droptol, fillfactor);
      for (int r = 0; r < n; ++r) d[r] = from_host<S>(f.dinv[r]);
      dinv = upload(d);
      upload_tri(f.L, true, L);
      upload_tri(f.U, false, U);
      CK(cudaMalloc(&tmp, sizeof(S) * (n ? n : 1)));
      if (verbose >= 1)
        std::cout << "\narpackItrSolver:\n\nILU: drop tolerance " << droptol << ", fill factor " << fillfactor << ", nnz(L) "
                  << f.L.val.size() << ", nnz(U) " << f.U.val.size() + n << ", levels " << L.nlevels << " / " << U.nlevels
                  << std::endl;
    } else {
      std::cerr << "Error: bad --slvItrPC - bad argument (Diag or ILU#D#F)" << std::endl;
      return 1;
    }
    return 0;
  }
  void apply(Vec<S>& vec, const S* r, S* z) const {  // z = M^-1 r
    if (kind == 1) {
      vec.mul(n, dinv, r, z);
    } else {
      k_sptrsv_levels<S><<<1, 1024, 0, stream>>>(L.nlevels, L.level_ptr, L.level_rows, L.rp, L.ci, L.v, (const S*)nullptr, r, tmp);
      k_sptrsv_levels<S><<<1, 1024, 0, stream>>>(U.nlevels, U.level_ptr, U.level_rows, U.rp, U.ci, U.v, dinv, tmp, z);
      CK(cudaGetLastError());
    }
  }
};

// ---- the inner solver of modes 2 and 3: x = M^-1 b -------------------------------------------------------------
template <typename S>
struct InnerSolver {
  using H = HostOf<S>;
  using R = RealOf<S>;
  const DevMat<S>* M = nullptr;
  Vec<S>* vec = nullptr;
  cudaStream_t stream = nullptr;
  int n = 0;
  std::string name;
  long long iterations = 0, solves = 0;
  // iterative
  bool iterative = false, cg = false;
  double tol = 1e-6;
  int maxit = 100;
  Preconditioner<S> pc;
  S *r = nullptr, *r0 = nullptr, *p = nullptr, *v = nullptr, *y = nullptr, *z = nullptr, *sv = nullptr, *t = nullptr;
  // direct
  cusolverSpHandle_t sp = nullptr;
  cusparseMatDescr_t descr = nullptr;
  csrqrInfo_t qr = nullptr;
  csrcholInfo_t chol = nullptr;
  void* spbuf = nullptr;
  int *f_rp = nullptr, *f_ci = nullptr;
  S* f_val = nullptr;
  cusolverDnHandle_t dn = nullptr;
  cublasHandle_t blas = nullptr;
  S *fac = nullptr, *tau = nullptr, *dnwork = nullptr, *rhs = nullptr;
  int* ipiv = nullptr;
  int* devinfo = nullptr;
  int dnlwork = 0;
  int dense_kind = 0;  // 1 getrf, 2 geqrf, 3 potrf

  int devinfo_value() {
    int h = 0;
    CK(cudaMemcpyAsync(&h, devinfo, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return h;
  }

  int init_iterative(const DevMat<S>* m, const Csr<H>& host, const Options& opt, Vec<S>* vv) {
    M = m; vec = vv; stream = vv->s; n = m->n; iterative = true; cg = (opt.slv == "CG");
    name = opt.slv;
    tol = opt.slvItrTol;
    maxit = opt.slvItrMaxIt;
    if (opt.slv != "BiCG" && opt.slv != "CG") { std::cerr << "Error: bad --slv - bad argument" << std::endl; return 1; }
    if (pc.init(host, opt.slvItrPC, stream, opt.verbose) != 0) return 1;
    for (S** b : {&r, &r0, &p, &v, &y, &z, &sv, &t}) CK(cudaMalloc(b, sizeof(S) * (n ? n : 1)));
    return 0;
  }

  double maxabs = 1.0;  // largest |entry| of the matrix being factorised

  int factor_sparse_qr(const Options& opt) {
    CKS(DL(cusolverSpCreateCsrqrInfo)(&qr));
    CKS(DL(cusolverSpXcsrqrAnalysis)(sp, n, n, (int)M->nnz, descr, f_rp, f_ci, qr));
    size_t internal = 0, workspace = 0;
    CKS(SP_FN(S, csrqrBufferInfo)(sp, n, n, (int)M->nnz, descr, cu((const S*)f_val), f_rp, f_ci, qr, &internal, &workspace));
    CK(cudaMalloc(&spbuf, workspace ? workspace : 1));
    CKS(SP_FN(S, csrqrSetup)(sp, n, n, (int)M->nnz, descr, cu((const S*)f_val), f_rp, f_ci, cuv(S(0)), qr));
    CKS(SP_FN(S, csrqrFactor)(sp, n, n, (int)M->nnz, (decltype(cu((S*)nullptr)))nullptr, (decltype(cu((S*)nullptr)))nullptr, qr, spbuf));
    int position = -1;
    // numerically zero diagonal of R: below pivot threshold x machine precision x the largest entry of the matrix
    CKS(SP_FN(S, csrqrZeroPivot)(sp, qr, (R)(opt.slvDrtPivot * std::numeric_limits<R>::epsilon() * maxabs), &position));
    if (position >= 0) { std::cerr << "Error: " << opt.slv << " - singular matrix (zero pivot at " << position << ")" << std::endl; return 1; }
    if (opt.verbose >= 1)
      std::cout << "\narpackDrtSolver:\n\nsparse QR on the GPU: internal " << internal << " B, workspace " << workspace << " B" << std::endl;
    return 0;
  }

  int init_direct(const DevMat<S>* m, const Csr<H>& host, const Options& opt, Vec<S>* vv) {
    M = m; vec = vv; stream = vv->s; n = m->n; iterative = false;
    name = opt.slv;
    const bool want_chol = (opt.slv == "LLT" || opt.slv == "LDLT");
    if (opt.slv != "LU" && opt.slv != "QR" && !want_chol) { std::cerr << "Error: bad --slv - bad argument" << std::endl; return 1; }
    CK(cudaMalloc(&rhs, sizeof(S) * (n ? n : 1)));
    if (!m->dense) {
      CKS(DL(cusolverSpCreate)(&sp));
      CKS(DL(cusolverSpSetStream)(sp, stream));
      CKS(DL(cusparseCreateMatDescr)(&descr));
      CKS(DL(cusparseSetMatType)(descr, CUSPARSE_MATRIX_TYPE_GENERAL));
      CKS(DL(cusparseSetMatIndexBase)(descr, CUSPARSE_INDEX_BASE_ZERO));
      // the factorisations take their own copy of the matrix: Cholesky with the adjusted diagonal (setShift)
      std::vector<S> fv(host.val.size());
      maxabs = 0.0;
      for (const H& a : host.val) maxabs = std::max(maxabs, (double)std::abs(a));
      for (int rr = 0; rr < n; ++rr)
        for (int q = host.rowptr[rr]; q < host.rowptr[rr + 1]; ++q) {
          H a = host.val[q];
          if (want_chol && host.col[q] == rr) a = H(opt.slvDrtOffset) + H(opt.slvDrtScale) * a;
          fv[q] = from_host<S>(a);
        }
      f_rp = upload(host.rowptr);
      f_ci = upload(host.col);
      f_val = upload(fv);
      if (want_chol) {
        CKS(DL(cusolverSpCreateCsrcholInfo)(&chol));
        CKS(DL(cusolverSpXcsrcholAnalysis)(sp, n, (int)M->nnz, descr, f_rp, f_ci, chol));
        size_t internal = 0, workspace = 0;
        CKS(SP_FN(S, csrcholBufferInfo)(sp, n, (int)M->nnz, descr, cu((const S*)f_val), f_rp, f_ci, chol, &internal, &workspace));
        CK(cudaMalloc(&spbuf, workspace ? workspace : 1));
        CKS(SP_FN(S, csrcholFactor)(sp, n, (int)M->nnz, descr, cu((const S*)f_val), f_rp, f_ci, chol, spbuf));
        int position = -1;
        CKS(SP_FN(S, csrcholZeroPivot)(sp, chol, (R)0, &position));
        if (position >= 0) {
          if (opt.slv == "LLT") { std::cerr << "Error: LLT - matrix is not positive definite (pivot " << position << ")" << std::endl; return 1; }
          std::cerr << "Warning: LDLT - non-positive pivot at " << position << ": using the sparse QR factorisation" << std::endl;
          CKS(DL(cusolverSpDestroyCsrcholInfo)(chol));
          chol = nullptr;
          CK(cudaFree(spbuf));
          spbuf = nullptr;
          return factor_sparse_qr(opt);
        }
        if (opt.verbose >= 1)
          std::cout << "\narpackDrtSolver:\n\nsparse Cholesky on the GPU: internal " << internal << " B, workspace " << workspace << " B" << std::endl;
        return 0;
      }
      return factor_sparse_qr(opt);
    }
    // dense: factorise a copy of the column-major matrix
    CKS(DL(cusolverDnCreate)(&dn));
    CKS(DL(cusolverDnSetStream)(dn, stream));
    CKS(DL(cublasCreate)(&blas));
    CKS(DL(cublasSetStream)(blas, stream));
    CK(cudaMalloc(&fac, sizeof(S) * (size_t)n * n));
    CK(cudaMalloc(&devinfo, sizeof(int)));
    CK(cudaMemcpyAsync(fac, m->colmajor, sizeof(S) * (size_t)n * n, cudaMemcpyDeviceToDevice, stream));
    if (want_chol) {
      if (opt.slvDrtOffset != 0.0 || opt.slvDrtScale != 1.0)
        std::cerr << "Warning: --slvDrtOffset / --slvDrtScale apply to the sparse Cholesky factorisations only" << std::endl;
      CKS(DN_FN(S, potrf_bufferSize)(dn, CUBLAS_FILL_MODE_LOWER, n, cu(fac), n, &dnlwork));
      CK(cudaMalloc(&dnwork, sizeof(S) * (dnlwork ? dnlwork : 1)));
      CKS(DN_FN(S, potrf)(dn, CUBLAS_FILL_MODE_LOWER, n, cu(fac), n, cu(dnwork), dnlwork, devinfo));
      const int info = devinfo_value();
      if (info == 0) { dense_kind = 3; return 0; }
      if (opt.slv == "LLT") { std::cerr << "Error: LLT - matrix is not positive definite (minor " << info << ")" << std::endl; return 1; }
      std::cerr << "Warning: LDLT - matrix is not positive definite (minor " << info << "): using LU with partial pivoting" << std::endl;
      CK(cudaFree(dnwork));
      dnwork = nullptr;
      CK(cudaMemcpyAsync(fac, m->colmajor, sizeof(S) * (size_t)n * n, cudaMemcpyDeviceToDevice, stream));
    }
    if (opt.slv == "QR") {
      CK(cudaMalloc(&tau, sizeof(S) * (n ? n : 1)));
      CKS(DN_FN(S, geqrf_bufferSize)(dn, n, n, cu(fac), n, &dnlwork));
      int lw2 = 0;
      CKS((pick<S>(DL(cusolverDnSormqr_bufferSize), DL(cusolverDnDormqr_bufferSize), DL(cusolverDnCunmqr_bufferSize), DL(cusolverDnZunmqr_bufferSize))(
          dn, CUBLAS_SIDE_LEFT, ScalarOf<S>::cplx ? CUBLAS_OP_C : CUBLAS_OP_T, n, 1, n, cu((const S*)fac), n, cu((const S*)tau),
          cu((const S*)rhs), n, &lw2)));
      dnlwork = std::max(dnlwork, lw2);
      CK(cudaMalloc(&dnwork, sizeof(S) * (dnlwork ? dnlwork : 1)));
      CKS(DN_FN(S, geqrf)(dn, n, n, cu(fac), n, cu(tau), cu(dnwork), dnlwork, devinfo));
      if (devinfo_value() != 0) { std::cerr << "Error: QR - geqrf KO" << std::endl; return 1; }
      dense_kind = 2;
      return 0;
    }
    CK(cudaMalloc(&ipiv, sizeof(int) * (n ? n : 1)));
    CKS(DN_FN(S, getrf_bufferSize)(dn, n, n, cu(fac), n, &dnlwork));
    CK(cudaMalloc(&dnwork, sizeof(S) * (dnlwork ? dnlwork : 1)));
    CKS(DN_FN(S, getrf)(dn, n, n, cu(fac), n, cu(dnwork), ipiv, devinfo));
    const int info = devinfo_value();
    if (info != 0) { std::cerr << "Error: " << opt.slv << " - singular matrix (U(" << info << "," << info << ") = 0)" << std::endl; return 1; }
    dense_kind = 1;
    return 0;
  }

  void solve(const S* b, S* x) {
    ++solves;
    if (!iterative) {
      if (!M->dense) {
        vec->copy(n, b, rhs);  // the factorisations may scribble on the right-hand side
        if (chol) CKS(SP_FN(S, csrcholSolve)(sp, n, cu((const S*)rhs), cu(x), chol, spbuf));
        else CKS(SP_FN(S, csrqrSolve)(sp, n, n, cu(rhs), cu(x), qr, spbuf));
      } else {
        vec->copy(n, b, x);
        if (dense_kind == 1) {
          CKS(DN_FN(S, getrs)(dn, CUBLAS_OP_N, n, 1, cu((const S*)fac), n, ipiv, cu(x), n, devinfo));
        } else if (dense_kind == 3) {
          CKS(DN_FN(S, potrs)(dn, CUBLAS_FILL_MODE_LOWER, n, 1, cu((const S*)fac), n, cu(x), n, devinfo));
        } else {  // x = R^-1 Q^H b
          CKS((pick<S>(DL(cusolverDnSormqr), DL(cusolverDnDormqr), DL(cusolverDnCunmqr), DL(cusolverDnZunmqr))(
              dn, CUBLAS_SIDE_LEFT, ScalarOf<S>::cplx ? CUBLAS_OP_C : CUBLAS_OP_T, n, 1, n, cu((const S*)fac), n, cu((const S*)tau), cu(x),
              n, cu(dnwork), dnlwork, devinfo)));
          CKS((pick<S>(DL(cublasStrsv), DL(cublasDtrsv), DL(cublasCtrsv), DL(cublasZtrsv))(blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n,
                                                                            cu((const S*)fac), n, cu(x), 1)));
        }
      }
      return;
    }
    // iterative, from x0 = 0 like Eigen's solve()
    vec->zero(n, x);
    const double rhs2 = vec->sqnorm(n, b);
    if (rhs2 == 0.0) return;
    const double tol2 = tol * tol * rhs2;
    vec->copy(n, b, r);
    if (cg) {  // Eigen's conjugate_gradient()
      const double threshold = std::max(tol2, std::numeric_limits<double>::min());
      if (rhs2 < threshold) return;
      pc.apply(*vec, r, p);
      double absNew = hreal(vec->dot(n, r, p));
      for (int it = 0; it < maxit; ++it) {
        M->apply(p, t);
        const H alpha = H(absNew) / vec->dot(n, p, t);
        vec->axpby(n, alpha, p, H(1.0), x);
        vec->axpby(n, -alpha, t, H(1.0), r);
        ++iterations;
        if (vec->sqnorm(n, r) < threshold) break;
        pc.apply(*vec, r, z);
        const double absOld = absNew;
        absNew = hreal(vec->dot(n, r, z));
        vec->axpby(n, H(1.0), z, H(absNew / absOld), p);
      }
      return;
    }
    // Eigen's bicgstab(): right-preconditioned, restarted when rho collapses
    vec->copy(n, r, r0);
    double r0_2 = rhs2;
    H rho = H(1.0), alpha = H(1.0), w = H(1.0);
    vec->zero(n, v);
    vec->zero(n, p);
    const double eps2 = std::numeric_limits<R>::epsilon() * (double)std::numeric_limits<R>::epsilon();
    int i = 0, restarts = 0;
    while (vec->sqnorm(n, r) > tol2 && i < maxit) {
      const H rho_old = rho;
      rho = vec->dot(n, r0, r);
      if (std::abs(rho) < eps2 * r0_2) {
        M->apply(x, t);                       // r = b - M x
        vec->copy(n, b, r);
        vec->axpby(n, H(-1.0), t, H(1.0), r);
        vec->copy(n, r, r0);
        r0_2 = vec->sqnorm(n, r);
        rho = H(r0_2);
        if (restarts++ == 0) i = 0;
      }
      const H beta = (rho / rho_old) * (alpha / w);
      vec->axpby(n, -w, v, H(1.0), p);        // p = r + beta (p - w v)
      vec->axpby(n, H(1.0), r, beta, p);
      pc.apply(*vec, p, y);
      M->apply(y, v);
      alpha = rho / vec->dot(n, r0, v);
      vec->copy(n, r, sv);
      vec->axpby(n, -alpha, v, H(1.0), sv);   // s = r - alpha v
      pc.apply(*vec, sv, z);
      M->apply(z, t);
      const double tt = vec->sqnorm(n, t);
      w = tt > 0.0 ? vec->dot(n, t, sv) / H(tt) : H(0.0);
      vec->axpby(n, alpha, y, H(1.0), x);
      vec->axpby(n, w, z, H(1.0), x);
      vec->copy(n, sv, r);
      vec->axpby(n, -w, t, H(1.0), r);
      ++i;
      ++iterations;
      if (w == H(0.0)) break;
    }
  }
};

// ---- the ICB entry points by scalar type -----------------------------------------------------------------
void aupd(bool sym, int* ido, const char* bmat, int n, const char* which, int nev, double tol, double* resid, int ncv,
          double* v, int ldv, int* iparam, int* ipntr, double* workd, double* workl, int lworkl, double*, int* info) {
  if (sym) dsaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
  else dnaupd_c(ido, bmat, n, which, nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void aupd(bool sym, int* ido, const char* bmat, int n, const char* which, int nev, double tol, float* resid, int ncv,
          float* v, int ldv, int* iparam, int* ipntr, float* workd, float* workl, int lworkl, float*, int* info) {
  if (sym) ssaupd_c(ido, bmat, n, which, nev, (float)tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
  else snaupd_c(ido, bmat, n, which, nev, (float)tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void aupd(bool, int* ido, const char* bmat, int n, const char* which, int nev, double tol, cdouble* resid, int ncv,
          cdouble* v, int ldv, int* iparam, int* ipntr, cdouble* workd, cdouble* workl, int lworkl, double* rwork, int* info) {
  znaupd_c(ido, bmat, n, which, nev, tol, (a_dcomplex*)resid, ncv, (a_dcomplex*)v, ldv, iparam, ipntr, (a_dcomplex*)workd,
           (a_dcomplex*)workl, lworkl, rwork, info);
}
void aupd(bool, int* ido, const char* bmat, int n, const char* which, int nev, double tol, cfloat* resid, int ncv,
          cfloat* v, int ldv, int* iparam, int* ipntr, cfloat* workd, cfloat* workl, int lworkl, float* rwork, int* info) {
  cnaupd_c(ido, bmat, n, which, nev, (float)tol, (a_fcomplex*)resid, ncv, (a_fcomplex*)v, ldv, iparam, ipntr, (a_fcomplex*)workd,
           (a_fcomplex*)workl, lworkl, rwork, info);
}

// eigenvalues (complex in general) after *eupd; z receives the vectors (arpackSolver.hpp:520-640)
struct EupdArgs {
  const Options* opt;
  int rvec;
  const char *howmny, *bmat, *which;
  const int* select;
  int n, nev, ncv, ldv, lworkl;
  int *iparam, *ipntr;
};
int eupd(const EupdArgs& a, double* z, double* resid, double* v, double* workd, double* workl, double*, std::vector<hcomplex>& vals) {
  int ierr = 0;
  std::vector<double> d(a.nev + 1, 0.0), di(a.nev + 1, 0.0), workev(3 * a.ncv, 0.0);
  if (a.opt->symPb)
    dseupd_c(a.rvec, a.howmny, a.select, d.data(), z, a.n, a.opt->sigmaReal, a.bmat, a.n, a.which, a.nev, a.opt->tol, resid, a.ncv, v,
             a.ldv, a.iparam, a.ipntr, workd, workl, a.lworkl, &ierr);
  else
    dneupd_c(a.rvec, a.howmny, a.select, d.data(), di.data(), z, a.n, a.opt->sigmaReal, a.opt->sigmaImag, workev.data(), a.bmat, a.n,
             a.which, a.nev, a.opt->tol, resid, a.ncv, v, a.ldv, a.iparam, a.ipntr, workd, workl, a.lworkl, &ierr);
  for (int k = 0; k <= a.nev; ++k) vals.emplace_back(d[k], a.opt->symPb ? 0.0 : di[k]);
  return ierr;
}
int eupd(const EupdArgs& a, float* z, float* resid, float* v, float* workd, float* workl, float*, std::vector<hcomplex>& vals) {
  int ierr = 0;
  std::vector<float> d(a.nev + 1, 0.f), di(a.nev + 1, 0.f), workev(3 * a.ncv, 0.f);
  if (a.opt->symPb)
    sseupd_c(a.rvec, a.howmny, a.select, d.data(), z, a.n, (float)a.opt->sigmaReal, a.bmat, a.n, a.which, a.nev, (float)a.opt->tol,
             resid, a.ncv, v, a.ldv, a.iparam, a.ipntr, workd, workl, a.lworkl, &ierr);
  else
    sneupd_c(a.rvec, a.howmny, a.select, d.data(), di.data(), z, a.n, (float)a.opt->sigmaReal, (float)a.opt->sigmaImag, workev.data(),
             a.bmat, a.n, a.which, a.nev, (float)a.opt->tol, resid, a.ncv, v, a.ldv, a.iparam, a.ipntr, workd, workl, a.lworkl, &ierr);
  for (int k = 0; k <= a.nev; ++k) vals.emplace_back((double)d[k], a.opt->symPb ? 0.0 : (double)di[k]);
  return ierr;
}
int eupd(const EupdArgs& a, cdouble* z, cdouble* resid, cdouble* v, cdouble* workd, cdouble* workl, double* rwork,
         std::vector<hcomplex>& vals) {
  int ierr = 0;
  std::vector<hcomplex> d(a.nev + 1, 0.0), workev(2 * a.ncv, 0.0);
  ab200_zneupd_ri(a.rvec, a.howmny, a.select, d.data(), z, a.n, a.opt->sigmaReal, a.opt->sigmaImag, workev.data(), a.bmat, a.n, a.which,
                  a.nev, a.opt->tol, resid, a.ncv, v, a.ldv, a.iparam, a.ipntr, workd, workl, a.lworkl, rwork, &ierr);
  for (int k = 0; k <= a.nev; ++k) vals.push_back(d[k]);
  return ierr;
}
int eupd(const EupdArgs& a, cfloat* z, cfloat* resid, cfloat* v, cfloat* workd, cfloat* workl, float* rwork,
         std::vector<hcomplex>& vals) {
  int ierr = 0;
  std::vector<std::complex<float>> d(a.nev + 1, 0.f), workev(2 * a.ncv, 0.f);
  ab200_cneupd_ri(a.rvec, a.howmny, a.select, d.data(), z, a.n, (float)a.opt->sigmaReal, (float)a.opt->sigmaImag, workev.data(), a.bmat,
                  a.n, a.which, a.nev, (float)a.opt->tol, resid, a.ncv, v, a.ldv, a.iparam, a.ipntr, workd, workl, a.lworkl, rwork, &ierr);
  for (int k = 0; k <= a.nev; ++k) vals.emplace_back((double)d[k].real(), (double)d[k].imag());
  return ierr;
}

// ---- the --restart dump files (arpackSolver.hpp:664-704): element count, then one value per line; complex values
// in the "(re,im)" form the C++ stream operators of the reference write and read ------------------------------------
template <typename H>
int restart_save(const char* path, const std::vector<H>& values) {
  if constexpr (std::is_same<H, double>::value) {
    return ab200_restart_save_f64(path, (long long)values.size(), values.data());
  } else {
    std::ofstream ofs(path, std::ofstream::trunc);
    if (!ofs.is_open()) return 1;
    ofs.precision(17);
    ofs << values.size() << "\n";
    for (const H& x : values) ofs << x << "\n";
    return ofs.good() ? 0 : 2;
  }
}
template <typename H>
int restart_load(const char* path, std::vector<H>& values, bool allow_zero, double eps) {
  if constexpr (std::is_same<H, double>::value) {
    return ab200_restart_load_f64(path, (long long)values.size(), values.data(), allow_zero ? 1 : 0);
  } else {
    std::ifstream ifs(path);
    if (!ifs.is_open()) return 1;
    long long have = 0;
    ifs >> have;
    if (!ifs || have != (long long)values.size()) { std::cerr << "arpack_b200: " << path << ": bad dim - restart KO" << std::endl; return 2; }
    for (H& x : values) {
      H val(0.0, 0.0);
      ifs >> val;
      if (!ifs) return 3;
      if (std::abs(val) < 1.e-6 && !allow_zero) val = H(eps, eps);  // makeConstant(epsilon): never a zero residual
      x = val;
    }
    return 0;
  }
}

struct Output {
  int nbVal = 0, mode = 0, nbIt = 0;
  double imsTime = 0.0, rciTime = 0.0;
};

double seconds_since(std::chrono::high_resolution_clock::time_point t0) {
  return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - t0).count() /
         1.e6;
}

template <typename S>
int run(Options& opt, Output& out) {
  using H = HostOf<S>;
  using R = RealOf<S>;
  constexpr bool cplx = ScalarOf<S>::cplx;
  // ---- matrices (arpackmm.cpp:625-660) ----
  Csr<H> A, B;
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
  if (opt.dense && (long long)n * n * (long long)sizeof(S) > (8LL << 30)) {
    std::cerr << "Error: --dense needs " << (double)n * n * sizeof(S) / 1e9 << " GB per copy: use the sparse path" << std::endl;
    return 1;
  }
  int nbCV = opt.nbCV;
  if (nbCV > n) nbCV = n;  // cut-off arpack workspace dim (arpackSolver.hpp:235)
  const int nev = opt.nbEV;

  // ---- problem transformation (arpackSolver.hpp:248-270) ----
  const double eps = std::numeric_limits<R>::epsilon();
  const bool shiftReal = std::fabs(opt.sigmaReal) > eps, shiftImag = std::fabs(opt.sigmaImag) > eps;
  const H sigma = make_host<H>(opt.sigmaReal, opt.sigmaImag);  // makeSigma: a real problem only sees the real part
  bool backTransform = false;
  int mode;
  Csr<H> Aop = A;
  if (opt.stdPb) {
    mode = 1;
    if (shiftReal && !shiftImag) { Aop = csr_add(A, -sigma, csr_identity<H>(n)); backTransform = true; }
  } else {
    mode = (shiftReal || shiftImag) ? 3 : 2;
  }
  out.mode = mode;
  if (opt.verbose >= 1) std::cout << "\narpackSolver:\n\nmode " << mode << ", backTransform " << (backTransform ? "yes" : "no") << std::endl;

  cudaStream_t stream = (cudaStream_t)ab200_get_stream();
  DevMat<S> dA, dB, dS;
  dA.upload_from(Aop, opt.dense, stream);
  Vec<S> vec(stream);
  InnerSolver<S> solver;
  t0 = std::chrono::high_resolution_clock::now();
  if (mode >= 2) {
    dB.upload_from(B, opt.dense, stream);
    const DevMat<S>* target = &dB;
    Csr<H> Shost;
    if (mode == 3) {
      Shost = csr_add(A, -sigma, B);
      dS.upload_from(Shost, opt.dense, stream);
      target = &dS;
    }
    const Csr<H>& host = (mode == 3) ? Shost : B;
    const int rc = opt.direct() ? solver.init_direct(target, host, opt, &vec) : solver.init_iterative(target, host, opt, &vec);
    if (rc != 0) { std::cerr << "Error: initialize solver KO" << std::endl; return 1; }
    CK(cudaStreamSynchronize(stream));
  }
  out.imsTime = seconds_since(t0);

  // ---- workspace: resid, v, workd, z in HBM; workl and the small arrays on the host ----
  const int ldv = n;
  S *resid = nullptr, *v = nullptr, *workd = nullptr, *z = nullptr, *scratch = nullptr;
  CK(cudaMalloc(&resid, sizeof(S) * (n ? n : 1)));
  CK(cudaMalloc(&scratch, sizeof(S) * (n ? n : 1)));
  CK(cudaMalloc(&v, sizeof(S) * (size_t)ldv * nbCV));
  CK(cudaMalloc(&workd, sizeof(S) * 3 * (size_t)n));
  CK(cudaMalloc(&z, sizeof(S) * (size_t)n * (nev + 1)));  // nbEV+1 for dneupd
  CK(cudaMemset(workd, 0, sizeof(S) * 3 * (size_t)n));
  CK(cudaMemset(z, 0, sizeof(S) * (size_t)n * (nev + 1)));
  {
    // close to, but not, zero (:737-747); makeConstant gives a complex constant equal real and imaginary parts
    const S r0v = from_host<S>(make_host<H>(eps, eps)), v0v = from_host<S>(make_host<H>(10., 10.) * make_host<H>(eps, eps));
    std::vector<S> r0(n, r0v), v0((size_t)ldv * nbCV, v0v);
    CK(cudaMemcpy(resid, r0.data(), sizeof(S) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v, v0.data(), sizeof(S) * v0.size(), cudaMemcpyHostToDevice));
  }
  const int lworkl = opt.symPb ? nbCV * nbCV + 8 * nbCV : 3 * nbCV * nbCV + 6 * nbCV;
  std::vector<S> workl(lworkl, S(0));
  std::vector<R> rwork(nbCV, R(0));
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
    std::vector<H> rr(n), vv((size_t)ldv * nbCV);
    if (restart_load("arpackSolver.resid.out", rr, false, eps) != 0) { std::cerr << "Error: bad restart (resid)" << std::endl; return 1; }
    if (restart_load("arpackSolver.v.out", vv, true, eps) != 0) { std::cerr << "Error: bad restart (v)" << std::endl; return 1; }
    std::vector<S> rt(rr.size()), vt(vv.size());
    for (size_t k = 0; k < rr.size(); ++k) rt[k] = from_host<S>(rr[k]);
    for (size_t k = 0; k < vv.size(); ++k) vt[k] = from_host<S>(vv[k]);
    CK(cudaMemcpy(resid, rt.data(), sizeof(S) * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(v, vt.data(), sizeof(S) * vt.size(), cudaMemcpyHostToDevice));
    if (opt.verbose >= 1) std::cout << "\narpackSolver:\n\narpackSolver.resid.out: restart OK\n\narpackSolver.v.out: restart OK" << std::endl;
  }
  if (opt.registered) {
    if constexpr (!cplx) {
      if (mode == 1 && !opt.dense) {
        int rc;
        if constexpr (std::is_same<S, double>::value)
          rc = ab200_register_csr_op_f64(workl.data(), n, dA.nnz, dA.rowptr, dA.col, dA.val);
        else
          rc = ab200_register_csr_op_f32(workl.data(), n, dA.nnz, dA.rowptr, dA.col, dA.val);
        if (rc != 0) { std::cerr << "Error: operator registration KO" << std::endl; return 1; }
      }
    }
  }

  // ---- reverse communication loop (arpackSolver.hpp:787-846) ----
  long long handoffs = 0;
  do {
    aupd(opt.symPb, &ido, bmat, n, which, nev, opt.tol, resid, nbCV, v, ldv, iparam, ipntr, workd, workl.data(), lworkl,
         rwork.data(), &info);
    if (info == 1) std::cerr << "Error: [dz][sn]aupd - KO: maximum number of iterations taken. Increase --maxIt..." << std::endl;
    if (info == 3) std::cerr << "Error: [dz][sn]aupd - KO: no shifts could be applied. Increase --nbCV..." << std::endl;
    if (info == -9) std::cerr << "Error: [dz][sn]aupd - KO: starting vector is zero. Retry: play with shift..." << std::endl;
    if (info < 0) { std::cerr << "Error: [dz][sn]aupd - KO with info " << info << ", nbIt " << iparam[2] << std::endl; return 1; }
    auto t1 = std::chrono::high_resolution_clock::now();
    S* X = workd + ipntr[0] - 1;
    S* Y = workd + ipntr[1] - 1;
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
  std::vector<int> select(nbCV, 1);
  std::vector<hcomplex> all, vals;
  EupdArgs ea{&opt, 1, opt.schur ? "P" : "A", bmat, which, select.data(), n, nev, nbCV, ldv, lworkl, iparam, ipntr};
  const int ierr = eupd(ea, z, resid, v, workd, workl.data(), rwork.data(), all);
  if (ierr == -14) std::cerr << "Error: [dz][sn]eupd - KO: [dz][sn]aupd did not find any eigenvalues to sufficient accuracy" << std::endl;
  if (ierr < 0 && ierr != -14) { std::cerr << "Error: [dz][sn]eupd - KO with info " << ierr << std::endl; std::cerr << "Error: bad arpack eupd" << std::endl; return 1; }
  int nconv = iparam[4];
  const int most = nev + ((opt.symPb || cplx) ? 0 : 1);
  if (nconv > most) nconv = most;
  if (ierr == -14) nconv = 0;
  for (int k = 0; k < nconv; ++k) vals.push_back(all[k]);
  if (backTransform)
    for (auto& l : vals) l += hcomplex(sigma);
  out.nbVal = (int)vals.size();

  // ---- dump for a later --restart (arpackmm always dumps: arpackmm.cpp:614) ----
  {
    std::vector<S> rt(n), vt((size_t)ldv * nbCV);
    CK(cudaMemcpy(rt.data(), resid, sizeof(S) * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vt.data(), v, sizeof(S) * vt.size(), cudaMemcpyDeviceToHost));
    std::vector<H> rr(rt.size()), vv(vt.size());
    for (size_t k = 0; k < rt.size(); ++k) rr[k] = to_host(rt[k]);
    for (size_t k = 0; k < vt.size(); ++k) vv[k] = to_host(vt[k]);
    restart_save("arpackSolver.resid.out", rr);
    restart_save("arpackSolver.v.out", vv);
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
    std::vector<S> zh((size_t)n * (nev + 1));
    CK(cudaMemcpy(zh.data(), z, sizeof(S) * zh.size(), cudaMemcpyDeviceToHost));
    const double dTol = std::sqrt(opt.tol);
    std::vector<hcomplex> V(n), left, right;
    auto check_one = [&](size_t i, size_t kre, bool pair, double sgn) -> bool {
      for (int r = 0; r < n; ++r) {
        if constexpr (cplx) {
          V[r] = to_host(zh[kre * n + r]);
        } else {
          const double re = (double)zh[kre * n + r];
          const double im = pair ? sgn * (double)zh[(kre + 1) * n + r] : 0.0;
          V[r] = {re, im};
        }
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
    // complex problems: column i is the vector of value i.  Real problems: a real eigenvalue has column i; for a
    // conjugate pair (i, i+1) the columns hold (Re, Im) of the first member's vector (dneupd.f:84-96), the second
    // member's vector is its conjugate
    for (size_t i = 0; i < vals.size();) {
      const bool pair = !cplx && !opt.symPb && vals[i].imag() != 0.0;
      if (!pair) {
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
    std::cout << "\nOUT: OP*x hand-offs " << handoffs << ", inner solver " << (mode >= 2 ? solver.name : std::string("none")) << ": "
              << solver.solves << " solves, " << solver.iterations << " iterations" << std::endl;
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
    else if (clo == "--dense") {
      const std::string rr = need(a, clo);
      if (rr != "true" && rr != "false") { std::cerr << "Error: bad " << clo << " - bad argument" << std::endl; return usage(); }
      opt.dense = true;
      opt.denseRR = (rr == "true");
    }
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
    else if (clo == "--slvItrPC") opt.slvItrPC = need(a, clo);
    else if (clo == "--slvDrtPivot") num(need(a, clo), clo, opt.slvDrtPivot);
    else if (clo == "--slvDrtOffset") num(need(a, clo), clo, opt.slvDrtOffset);
    else if (clo == "--slvDrtScale") num(need(a, clo), clo, opt.slvDrtScale);
    else if (clo == "--noCheck") opt.check = false;
    else if (clo == "--verbose") num(need(a, clo), clo, opt.verbose);
    else if (clo == "--debug") {
      num(need(a, clo), clo, opt.debug);
      if (opt.debug > 3) opt.debug = 3;
      const int d = opt.debug;
      debug_c(6, -6, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d, d);  // arpackmm.cpp:296-298
    }
    else if (clo == "--restart") opt.restart = true;
    else if (clo == "--registered") opt.registered = true;
    else std::cerr << "Warning: unknown option " << clo << " ignored" << std::endl;  // the reference's parser skips them silently
  }
  if (!opt.stdPb && opt.fileB.empty()) {
    std::cerr << "Error: generalized problem without B matrix, specify --B XX with XX being a matrix market file" << std::endl;
    return usage();
  }
  if (opt.dense && !opt.direct()) {
    std::cerr << "Error: dense matrices does not support iterative solvers, specify --slv XX with XX being a direct solver" << std::endl;
    return 1;
  }
  if (opt.symPb && (opt.mag == "LR" || opt.mag == "SR" || opt.mag == "LI" || opt.mag == "SI")) {
    std::cerr << "Error: bad --mag for a symmetric problem" << std::endl;
    return 1;
  }
  if (ab200_device_count() <= 0) { std::cerr << "Error: no CUDA device; arpackmm_b200 has no CPU path" << std::endl; return 1; }

  std::cout << "OPT: A " << opt.fileA << ", B " << opt.fileB << ", dense "
            << (opt.dense ? (opt.denseRR ? "yes (RR true)" : "yes (RR false)") : "no") << ", nbEV " << opt.nbEV << ", nbCV "
            << opt.nbCV << ", stdPb " << (opt.stdPb ? "yes" : "no") << ", symPb " << (opt.symPb ? "yes" : "no") << ", cpxPb "
            << (opt.cpxPb ? "yes" : "no") << ", simplePrec " << (opt.simplePrec ? "yes" : "no") << ", mag " << opt.mag << std::endl;
  std::cout << "OPT: shiftReal " << (opt.shiftReal ? "yes" : "no") << ", sigmaReal " << opt.sigmaReal << ", shiftImag "
            << (opt.shiftImag ? "yes" : "no") << ", sigmaImag " << opt.sigmaImag << ", invert " << (opt.invert ? "yes" : "no")
            << ", tol " << opt.tol << ", maxIt " << opt.maxIt << ", " << (opt.schur ? "Schur" : "Ritz") << " vectors" << std::endl;
  std::cout << "OPT: slv " << opt.slv;
  if (!opt.direct())
    std::cout << ", slvItrPC " << opt.slvItrPC << ", slvItrTol " << opt.slvItrTol << ", slvItrMaxIt " << opt.slvItrMaxIt;
  else
    std::cout << ", slvDrtPivot " << opt.slvDrtPivot << ", slvDrtOffset " << opt.slvDrtOffset << ", slvDrtScale " << opt.slvDrtScale;
  std::cout << std::endl;
  std::cout << "OPT: check " << (opt.check ? "yes" : "no") << ", verbose " << opt.verbose << ", debug " << opt.debug << ", restart "
            << (opt.restart ? "yes" : "no") << ", registered " << (opt.registered ? "yes" : "no") << std::endl;

  std::cout.precision(15);  // the reference prints 6 digits; more are harmless and let scripts compare values
  sstats_c();  // reset the counters (arpackmm.cpp:846-848)
  sstatn_c();
  auto start = std::chrono::high_resolution_clock::now();
  Output out;
  int rc;
  if (opt.cpxPb) rc = opt.simplePrec ? run<cfloat>(opt, out) : run<cdouble>(opt, out);
  else rc = opt.simplePrec ? run<float>(opt, out) : run<double>(opt, out);
  if (rc != 0) { std::cerr << "Error: arpack solve KO" << std::endl; return rc; }
  std::cout << "\nOUT: mode " << out.mode << ", nb EV found " << out.nbVal << ", nb iterations " << out.nbIt << std::endl;
  std::cout << "OUT: init mode solver " << out.imsTime << " s, RCI time " << out.rciTime << " s" << std::endl;
  std::cout << "OUT: full time " << seconds_since(start) << " s" << std::endl;

  int nopx = 0, nbx = 0, nrorth = 0, nitref = 0, nrstrt = 0;
  float t[26] = {0};
  stat_c(&nopx, &nbx, &nrorth, &nitref, &nrstrt, &t[0], &t[1], &t[2], &t[3], &t[4], &t[5], &t[6], &t[7], &t[8], &t[9], &t[10],
         &t[11], &t[12], &t[13], &t[14], &t[15], &t[16], &t[17], &t[18], &t[19], &t[20], &t[21], &t[22], &t[23], &t[24], &t[25]);
  std::cout << "\nSTAT: total number of user OP*x operation                         " << nopx << std::endl;
  std::cout << "STAT: total number of user  B*x operation                         " << nbx << std::endl;
  std::cout << "STAT: total number of reorthogonalization steps taken             " << nrorth << std::endl;
  std::cout << "STAT: total number of it. refinement steps in reorthogonalization " << nitref << std::endl;
  std::cout << "STAT: total number of restart steps                               " << nrstrt << std::endl;
  return 0;
}
