// api_cplx.cu -- C-ABI of the complex Arnoldi entry points (declared in include/arpack_b200.h):
//   znaupd_c zneupd_c cnaupd_c cneupd_c   ICB/arpack.h:10-11,20-21 (bind(c) shims SRC/icbazn.F90, icbacn.F90)
//   znaupd_  zneupd_                      legacy Fortran ABI (SRC/znaupd.f:384, zneupd.f:248)
// Same conventions as api.cu: resid, v, workd and z may be HOST or DEVICE pointers (classified at ido = 0; host arrays
// get HBM mirrors owned by a context keyed to the workl address), workl/rwork/iparam/ipntr/select/d/workev are host
// memory, no CPU fallback (info = -9990 without a usable device).
#include <cuda_runtime.h>

#include <complex>
#include <cstdio>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "../../include/arpack_b200.h"
#include "irl_complex.hpp"
#include "vecops_cplx.cuh"

namespace ab200 {
void set_last_counters(const Counters& c);  // api.cu (stat_c)
NcclComm* comm_from_handle(int handle);     // comm_nccl.cpp

namespace {

constexpr int kInfoDeviceError = -9990;
std::mutex g_mu_z;

template <typename R> struct GlobalsZ {
  static SeedState seed;      // zgetv0's SAVE'd iseed (zgetv0.f:157-162)
  static SeedState seed_par;  // pzgetv0's
  static R smlnum_first;      // znaitr/znapps SAVE'd smlnum (znaitr.f:303-317)
};
template <typename R> SeedState GlobalsZ<R>::seed;
template <typename R> SeedState GlobalsZ<R>::seed_par;
template <typename R> R GlobalsZ<R>::smlnum_first = R(-1);

template <typename R>
struct CtxZ {
  using Z = std::complex<R>;
  std::unique_ptr<CudaVecOpsZ<R>> ops;
  std::unique_ptr<IrlComplex<R>> slv;
  bool par = false;
  int n = 0, ncv = 0, mode = 1;
  char bmat = 'I';
  Z *resid_u = nullptr, *v_u = nullptr, *workd_u = nullptr;
  bool resid_host = false, v_host = false, workd_host = false;
  Z *resid_d = nullptr, *v_d = nullptr, *workd_d = nullptr;
  int64_t ldv_d = 0;
  Z* z_mirror = nullptr;
  int last_ido = 0;
  int last_ipntr[3] = {0, 0, 0};
  bool finished = false;
  ~CtxZ() {
    if (ops) {
      try { ops->sync(); } catch (...) {}
      if (resid_host) ops->release(resid_d);
      if (v_host) ops->release(v_d);
      if (workd_host) ops->release(workd_d);
      ops->release(z_mirror);
    }
  }
};

template <typename R>
std::unordered_map<const void*, std::unique_ptr<CtxZ<R>>>& ztable() {
  static std::unordered_map<const void*, std::unique_ptr<CtxZ<R>>> t;
  return t;
}

void require_device_z() {
  int cnt = 0;
  const cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0) {
    cudaGetLastError();
    throw CudaError(std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); arpack_b200 has no CPU path");
  }
}

template <typename R>
CtxZ<R>* make_ctx_z(const void* key, bool par, int comm_handle, int n, int ncv, std::complex<R>* resid,
                    std::complex<R>* v, int ldv, std::complex<R>* workd, bool upload_all) {
  require_device_z();
  auto c = std::make_unique<CtxZ<R>>();
  NcclComm* comm = nullptr;
  if (par) {
    comm = comm_from_handle(comm_handle);
    if (!comm) throw CudaError("p[cz]naupd_c: comm is not a handle returned by ab200_comm_create()");
  }
  c->par = par;
  c->ops = std::make_unique<CudaVecOpsZ<R>>((cudaStream_t)ab200_get_stream(), comm);
  c->n = n;
  c->ncv = ncv;
  c->resid_u = resid; c->v_u = v; c->workd_u = workd;
  c->resid_host = !c->ops->is_device_pointer(resid);
  c->v_host = !c->ops->is_device_pointer(v);
  c->workd_host = !c->ops->is_device_pointer(workd);
  if (n > 0 && ncv > 0) {
    c->resid_d = c->resid_host ? c->ops->alloc((size_t)n) : resid;
    c->ldv_d = c->v_host ? (int64_t)n : (int64_t)ldv;
    c->v_d = c->v_host ? c->ops->alloc((size_t)c->ldv_d * ncv) : v;
    c->workd_d = c->workd_host ? c->ops->alloc((size_t)3 * n) : workd;
    if (upload_all) {
      if (c->resid_host) c->ops->upload(c->resid_d, resid, (size_t)n);
      if (c->v_host) c->ops->upload2d(c->v_d, (size_t)c->ldv_d, v, (size_t)ldv, (size_t)n, (size_t)ncv);
      if (c->workd_host) c->ops->upload(c->workd_d, workd, (size_t)3 * n);
    }
  }
  CtxZ<R>* raw = c.get();
  std::lock_guard<std::mutex> lk(g_mu_z);
  ztable<R>()[key] = std::move(c);
  return raw;
}

template <typename R>
CtxZ<R>* find_ctx_z(const void* key) {
  std::lock_guard<std::mutex> lk(g_mu_z);
  auto it = ztable<R>().find(key);
  return it == ztable<R>().end() ? nullptr : it->second.get();
}

template <typename R>
std::unique_ptr<IrlComplex<R>> make_solver_z(CtxZ<R>* c) {
  return std::make_unique<IrlComplex<R>>(c->ops.get(), c->par ? &GlobalsZ<R>::seed_par : &GlobalsZ<R>::seed,
                                         &GlobalsZ<R>::smlnum_first, c->par);
}

template <typename R>
void zaupd_entry(bool par, int comm_handle, int* ido, const char* bmat, int n, const char* which, int nev, R* tol, std::complex<R>* resid,
                 int ncv, std::complex<R>* v, int ldv, int* iparam, int* ipntr, std::complex<R>* workd,
                 std::complex<R>* workl, int lworkl, R* rwork, int* info) {
  try {
    CtxZ<R>* c = nullptr;
    if (*ido == 0) {
      c = make_ctx_z<R>(workl, par, comm_handle, n, ncv, resid, v, ldv, workd, false);
      c->bmat = bmat[0];
      c->mode = iparam[6];
      c->slv = make_solver_z<R>(c);
      if (*info != 0 && c->resid_host && n > 0) c->ops->upload(c->resid_d, resid, (size_t)n);
    } else {
      c = find_ctx_z<R>(workl);
      if (!c || c->finished) throw CudaError("[cz]naupd_c re-entered with ido != 0 but no active solve is keyed to this workl");
      // the user's result of the previous hand-off
      if (c->workd_host && (c->last_ido == -1 || c->last_ido == 1 || c->last_ido == 2))
        c->ops->upload(c->workd_d + c->last_ipntr[1] - 1, c->workd_u + c->last_ipntr[1] - 1, (size_t)c->n);
    }
    c->slv->aupd(ido, bmat[0], n, which, nev, tol, c->resid_d, ncv, c->v_d, c->ldv_d, iparam, ipntr, c->workd_d, workl,
                 lworkl, rwork, info);
    c->last_ido = *ido;
    c->last_ipntr[0] = ipntr[0]; c->last_ipntr[1] = ipntr[1]; c->last_ipntr[2] = ipntr[2];
    if (*ido == -1 || *ido == 1 || *ido == 2) {
      if (c->workd_host) {
        c->ops->download(c->workd_u + ipntr[0] - 1, c->workd_d + ipntr[0] - 1, (size_t)c->n);
        if (*ido == 1 && (c->mode >= 3 || c->bmat == 'G'))
          c->ops->download(c->workd_u + ipntr[2] - 1, c->workd_d + ipntr[2] - 1, (size_t)c->n);
        c->ops->sync();
      }
    } else if (*ido == 99) {
      set_last_counters(c->slv->counters());
      if (c->ops && c->n > 0 && (*info >= 0 || *info == -8 || *info == -9 || *info == -9999)) {
        if (c->resid_host) c->ops->download(resid, c->resid_d, (size_t)c->n);
        if (c->v_host) c->ops->download2d(v, (size_t)ldv, c->v_d, (size_t)c->ldv_d, (size_t)c->n, (size_t)c->ncv);
        if (c->workd_host && c->bmat == 'G') c->ops->download(workd, c->workd_d, (size_t)3 * c->n);
      }
      c->ops->sync();
      c->finished = true;
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: [cz]naupd_c: %s\n", e.what());
    *info = kInfoDeviceError;
    *ido = 99;
  }
}

template <typename R>
void zeupd_entry(bool par, int comm_handle, int rvec, const char* howmny, const int* select, std::complex<R>* d, std::complex<R>* z, int ldz,
                 std::complex<R> sigma, std::complex<R>* workev, const char* bmat, int n, const char* which, int nev,
                 R tol, std::complex<R>* resid, int ncv, std::complex<R>* v, int ldv, int* iparam, int* ipntr,
                 std::complex<R>* workd, std::complex<R>* workl, int lworkl, R* rwork, int* info) {
  using Z = std::complex<R>;
  try {
    CtxZ<R>* c = find_ctx_z<R>(workl);
    if (!(c && c->finished && c->v_u == v && c->n == n && c->ncv == ncv)) {
      // zneupd without a preceding znaupd in this process (or with other arrays): rebuild the device view
      c = make_ctx_z<R>(workl, par, comm_handle, n, ncv, resid, v, ldv, workd, true);
      c->finished = true;
    }
    if (!c->slv) c->slv = make_solver_z<R>(c);
    c->slv->ensure_mailbox(ncv);
    // map z: alias of v, device array, or host array with an HBM mirror
    Z* zdev = nullptr;
    int64_t zld = 0;
    bool zhost = false;
    if (rvec) {
      if (z == c->v_u) {
        zdev = c->v_d; zld = c->ldv_d;
      } else if (!c->ops->is_device_pointer(z)) {
        zhost = true;
        zld = n;
        c->ops->release(c->z_mirror);
        c->z_mirror = c->ops->alloc((size_t)zld * nev);
        zdev = c->z_mirror;
      } else {
        zdev = z; zld = ldz;
      }
    }
    std::vector<int> sel(select, select + (ncv > 0 ? ncv : 0));
    c->slv->eupd(rvec != 0, howmny[0], sel.data(), d, zdev, zld, sigma, workev, bmat[0], n, which, nev, tol, c->resid_d,
                 ncv, c->v_d, c->ldv_d, iparam, ipntr, c->workd_d, workl, lworkl, rwork, info);
    if (rvec && *info == 0) {
      const int nconv = std::min(iparam[4], nev);
      if (zhost) c->ops->download2d(z, (size_t)ldz, zdev, (size_t)zld, (size_t)n, (size_t)nconv);
      if (c->v_host) c->ops->download2d(v, (size_t)ldv, c->v_d, (size_t)c->ldv_d, (size_t)n, (size_t)ncv);
    }
    c->ops->sync();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: [cz]neupd_c: %s\n", e.what());
    *info = kInfoDeviceError;
  }
}

}  // namespace

// hooks for api.cu (ab200_release, ab200_release_all, ab200_reset_seed)
void release_cplx(const void* workl) {
  std::lock_guard<std::mutex> lk(g_mu_z);
  ztable<double>().erase(workl);
  ztable<float>().erase(workl);
}
void release_all_cplx() {
  std::lock_guard<std::mutex> lk(g_mu_z);
  ztable<double>().clear();
  ztable<float>().clear();
}
void reset_seed_cplx() {
  GlobalsZ<double>::seed = SeedState(); GlobalsZ<float>::seed = SeedState();
  GlobalsZ<double>::seed_par = SeedState(); GlobalsZ<float>::seed_par = SeedState();
  GlobalsZ<double>::smlnum_first = -1.0; GlobalsZ<float>::smlnum_first = -1.0f;
}

}  // namespace ab200

using namespace ab200;
using zd = std::complex<double>;
using zf = std::complex<float>;

extern "C" {

void znaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol, a_dcomplex* resid,
              a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_dcomplex* workd, a_dcomplex* workl,
              a_int lworkl, double* rwork, a_int* info) {
  zaupd_entry<double>(false, 0, ido, bmat, n, which, nev, &tol, (zd*)resid, ncv, (zd*)v, ldv, iparam, ipntr, (zd*)workd,
                      (zd*)workl, lworkl, rwork, info);
}
void zneupd_c(a_int rvec, char const* howmny, a_int const* select, a_dcomplex* d, a_dcomplex* z, a_int ldz,
              a_dcomplex sigma, a_dcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev,
              double tol, a_dcomplex* resid, a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr,
              a_dcomplex* workd, a_dcomplex* workl, a_int lworkl, double* rwork, a_int* info) {
  zeupd_entry<double>(false, 0, rvec, howmny, select, (zd*)d, (zd*)z, ldz, zd(sigma.re, sigma.im), (zd*)workev, bmat, n, which,
                      nev, tol, (zd*)resid, ncv, (zd*)v, ldv, iparam, ipntr, (zd*)workd, (zd*)workl, lworkl, rwork,
                      info);
}
void cnaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol, a_fcomplex* resid,
              a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_fcomplex* workd, a_fcomplex* workl,
              a_int lworkl, float* rwork, a_int* info) {
  zaupd_entry<float>(false, 0, ido, bmat, n, which, nev, &tol, (zf*)resid, ncv, (zf*)v, ldv, iparam, ipntr, (zf*)workd,
                     (zf*)workl, lworkl, rwork, info);
}
void cneupd_c(a_int rvec, char const* howmny, a_int const* select, a_fcomplex* d, a_fcomplex* z, a_int ldz,
              a_fcomplex sigma, a_fcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev, float tol,
              a_fcomplex* resid, a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_fcomplex* workd,
              a_fcomplex* workl, a_int lworkl, float* rwork, a_int* info) {
  zeupd_entry<float>(false, 0, rvec, howmny, select, (zf*)d, (zf*)z, ldz, zf(sigma.re, sigma.im), (zf*)workev, bmat, n, which,
                     nev, tol, (zf*)resid, ncv, (zf*)v, ldv, iparam, ipntr, (zf*)workd, (zf*)workl, lworkl, rwork,
                     info);
}
// ---- ICB/parpack.h:28-33 (PARPACK/SRC/MPI/icbpzn.F90, icbpcn.F90); comm = handle from ab200_comm_create ----
void pznaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol,
               a_dcomplex* resid, a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_dcomplex* workd,
               a_dcomplex* workl, a_int lworkl, double* rwork, a_int* info) {
  zaupd_entry<double>(true, comm, ido, bmat, n, which, nev, &tol, (zd*)resid, ncv, (zd*)v, ldv, iparam, ipntr,
                      (zd*)workd, (zd*)workl, lworkl, rwork, info);
}
void pzneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, a_dcomplex* d, a_dcomplex* z,
               a_int ldz, a_dcomplex sigma, a_dcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev,
               double tol, a_dcomplex* resid, a_int ncv, a_dcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr,
               a_dcomplex* workd, a_dcomplex* workl, a_int lworkl, double* rwork, a_int* info) {
  zeupd_entry<double>(true, comm, rvec, howmny, select, (zd*)d, (zd*)z, ldz, zd(sigma.re, sigma.im), (zd*)workev, bmat,
                      n, which, nev, tol, (zd*)resid, ncv, (zd*)v, ldv, iparam, ipntr, (zd*)workd, (zd*)workl, lworkl,
                      rwork, info);
}
void pcnaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol,
               a_fcomplex* resid, a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr, a_fcomplex* workd,
               a_fcomplex* workl, a_int lworkl, float* rwork, a_int* info) {
  zaupd_entry<float>(true, comm, ido, bmat, n, which, nev, &tol, (zf*)resid, ncv, (zf*)v, ldv, iparam, ipntr,
                     (zf*)workd, (zf*)workl, lworkl, rwork, info);
}
void pcneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, a_fcomplex* d, a_fcomplex* z,
               a_int ldz, a_fcomplex sigma, a_fcomplex* workev, char const* bmat, a_int n, char const* which, a_int nev,
               float tol, a_fcomplex* resid, a_int ncv, a_fcomplex* v, a_int ldv, a_int* iparam, a_int* ipntr,
               a_fcomplex* workd, a_fcomplex* workl, a_int lworkl, float* rwork, a_int* info) {
  zeupd_entry<float>(true, comm, rvec, howmny, select, (zf*)d, (zf*)z, ldz, zf(sigma.re, sigma.im), (zf*)workev, bmat, n,
                     which, nev, tol, (zf*)resid, ncv, (zf*)v, ldv, iparam, ipntr, (zf*)workd, (zf*)workl, lworkl, rwork,
                     info);
}
// pzneupd_c with sigma as two reals (Python ctypes)
void ab200_pzneupd_ri(a_fint comm, a_int rvec, char const* howmny, a_int const* select, void* d, void* z, a_int ldz,
                      double sigma_re, double sigma_im, void* workev, char const* bmat, a_int n, char const* which,
                      a_int nev, double tol, void* resid, a_int ncv, void* v, a_int ldv, a_int* iparam, a_int* ipntr,
                      void* workd, void* workl, a_int lworkl, double* rwork, a_int* info) {
  zeupd_entry<double>(true, comm, rvec, howmny, select, (zd*)d, (zd*)z, ldz, zd(sigma_re, sigma_im), (zd*)workev, bmat,
                      n, which, nev, tol, (zd*)resid, ncv, (zd*)v, ldv, iparam, ipntr, (zd*)workd, (zd*)workl, lworkl,
                      rwork, info);
}

// ctypes-friendly twins of z/cneupd_c: sigma as two reals (ctypes cannot pass a C complex by value)
void ab200_zneupd_ri(a_int rvec, char const* howmny, a_int const* select, void* d, void* z, a_int ldz, double sigma_re,
                     double sigma_im, void* workev, char const* bmat, a_int n, char const* which, a_int nev, double tol,
                     void* resid, a_int ncv, void* v, a_int ldv, a_int* iparam, a_int* ipntr, void* workd, void* workl,
                     a_int lworkl, double* rwork, a_int* info) {
  zeupd_entry<double>(false, 0, rvec, howmny, select, (zd*)d, (zd*)z, ldz, zd(sigma_re, sigma_im), (zd*)workev, bmat, n, which,
                      nev, tol, (zd*)resid, ncv, (zd*)v, ldv, iparam, ipntr, (zd*)workd, (zd*)workl, lworkl, rwork,
                      info);
}
void ab200_cneupd_ri(a_int rvec, char const* howmny, a_int const* select, void* d, void* z, a_int ldz, float sigma_re,
                     float sigma_im, void* workev, char const* bmat, a_int n, char const* which, a_int nev, float tol,
                     void* resid, a_int ncv, void* v, a_int ldv, a_int* iparam, a_int* ipntr, void* workd, void* workl,
                     a_int lworkl, float* rwork, a_int* info) {
  zeupd_entry<float>(false, 0, rvec, howmny, select, (zf*)d, (zf*)z, ldz, zf(sigma_re, sigma_im), (zf*)workev, bmat, n, which,
                     nev, tol, (zf*)resid, ncv, (zf*)v, ldv, iparam, ipntr, (zf*)workd, (zf*)workl, lworkl, rwork,
                     info);
}

// Kernel unit-test hooks (tests/test_gpu_complex.py): ONE orthogonalisation pass / ONE restart update of the complex
// path on caller-supplied DEVICE arrays (interleaved complex128), results handed back to host arrays, so that each
// kernel can be compared with a plain numpy statement of the same op independently of the solver.
//   out_host (complex, j + 2 entries) = { h = V_j^H w (j), sum conj(w) w, sum |r|^2 } with r = w - V_j h written to resid
int ab200_debug_zorth_f64(long long n, int j, const void* v, long long ldv, const void* w, void* resid, void* out_host) {
  try {
    require_device_z();
    CudaVecOpsZ<double> ops((cudaStream_t)ab200_get_stream());
    zd* mb = ops.mailbox((size_t)j + 4);
    ops.dots(n, j, (const zd*)v, ldv, (const zd*)w, (const zd*)w, mb);
    ops.update(n, j, (const zd*)v, ldv, mb, (const zd*)w, (zd*)resid, mb + j + 1);
    ops.fetch((zd*)out_host, mb, (size_t)j + 2);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: debug_zorth: %s\n", e.what());
    return -1;
  }
}
// out(:,0:kout) = V(:,0:kin)*Q (q_host column-major kin x kout, interleaved complex128); out == v is allowed (in place);
// resid <- sigma*resid + beta*out(:,beta_col) (beta_col < 0: no beta term); nrm2_host[0] = sum |resid|^2
int ab200_debug_zvq_f64(long long n, int kin, int kout, const void* v, long long ldv, const void* q_host, void* out,
                        long long ldo, double sigma_re, double sigma_im, double beta_re, double beta_im, int beta_col,
                        void* resid, double* nrm2_host) {
  try {
    require_device_z();
    CudaVecOpsZ<double> ops((cudaStream_t)ab200_get_stream());
    zd* mb = ops.mailbox(4);
    if (out == v) {
      ops.vq_update(n, kin, kout, (zd*)out, ldv, (const zd*)q_host, kin, resid != nullptr, zd(sigma_re, sigma_im),
                    zd(beta_re, beta_im), beta_col, (zd*)resid, mb);
    } else {
      ops.vq_out(n, kin, kout, (const zd*)v, ldv, (const zd*)q_host, kin, (zd*)out, ldo);
    }
    zd h(0);
    ops.fetch(&h, mb, 1);
    nrm2_host[0] = h.real();
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: debug_zvq: %s\n", e.what());
    return -1;
  }
}

// legacy Fortran ABI (gfortran: everything by reference, CHARACTER lengths appended)
void znaupd_(a_int* ido, const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, a_dcomplex* resid,
             a_int* ncv, a_dcomplex* v, a_int* ldv, a_int* iparam, a_int* ipntr, a_dcomplex* workd, a_dcomplex* workl,
             a_int* lworkl, double* rwork, a_int* info, size_t, size_t) {
  zaupd_entry<double>(false, 0, ido, bmat, *n, which, *nev, tol, (zd*)resid, *ncv, (zd*)v, *ldv, iparam, ipntr, (zd*)workd,
                      (zd*)workl, *lworkl, rwork, info);
}
void zneupd_(a_int* rvec, const char* howmny, a_int* select, a_dcomplex* d, a_dcomplex* z, a_int* ldz,
             a_dcomplex* sigma, a_dcomplex* workev, const char* bmat, a_int* n, const char* which, a_int* nev,
             double* tol, a_dcomplex* resid, a_int* ncv, a_dcomplex* v, a_int* ldv, a_int* iparam, a_int* ipntr,
             a_dcomplex* workd, a_dcomplex* workl, a_int* lworkl, double* rwork, a_int* info, size_t, size_t, size_t) {
  zeupd_entry<double>(false, 0, *rvec, howmny, select, (zd*)d, (zd*)z, *ldz, zd(sigma->re, sigma->im), (zd*)workev, bmat, *n,
                      which, *nev, *tol, (zd*)resid, *ncv, (zd*)v, *ldv, iparam, ipntr, (zd*)workd, (zd*)workl, *lworkl,
                      rwork, info);
}
}
