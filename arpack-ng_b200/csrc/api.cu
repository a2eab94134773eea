// api.cu -- the C-ABI of libarpack_b200.so (declared in include/arpack_b200.h).
//
// Exports the reference's ISO_C_BINDING entry points with unchanged signatures
// (ICB/arpack.h:12-21, ICB/parpack.h:20-27, SRC/icbads.F90, SRC/icbadn.F90), the legacy Fortran-ABI
// names, the debug/stat accessors (ICB/debug_c.h, ICB/stat_c.h) and a few extensions.
//
// resid, v, workd and z may be HOST or DEVICE pointers, classified per array with
// cudaPointerGetAttributes at ido = 0:
//   * device pointers are used in place; the ido=+-1/2 hand-off then passes device addresses
//     workd + ipntr[k] - 1 to the caller's OP kernel, ordered on ab200_get_stream();
//   * host pointers get a device mirror owned by the solve context; the operand handed to the user
//     is copied D2H before returning and the user's result H2D on re-entry, V/resid are copied back
//     at ido = 99.  An unmodified CPU caller therefore works as is.
// workl, iparam, ipntr, select, d/dr/di and workev are always host memory.
//
// There is no CPU fallback: without a usable CUDA device every entry point reports info = -9990
// and prints the reason on stderr.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/arpack_b200.h"
#include "driver.hpp"
#include "irl_nonsym.hpp"
#include "irl_sym.hpp"
#include "vecops_cuda.cuh"

namespace ab200 {
NcclComm* comm_from_handle(int handle);
int gram_apply(int handle, const double* x_loc, double* z_loc);  // gram.cu
// complex entry points live in api_cplx.cu
void release_cplx(const void* workl);
void release_all_cplx();
void reset_seed_cplx();

namespace {

constexpr int kInfoDeviceError = -9990;

// process-wide settings: read at ido = 0 of a solve, written by the ab200_set_* calls (possibly from another thread)
std::atomic<cudaStream_t> g_stream{nullptr};
std::atomic<int> g_kernel_mode{0};
std::atomic<int> g_compat{0};  // ab200_set_compat(): bit 0 = maintain workd(ipntr(3)) in mode 1
std::mutex g_mu;

// COMMON /debug/ (debug.h) and the counters of COMMON /timing/ (stat.h) of the last solve
struct DebugLevels { int v[24]; };
DebugLevels g_debug = {{6, -3, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};
Counters g_last_counters;

template <typename T> struct Globals {
  static SeedState seed;       // dgetv0's SAVE'd iseed (dgetv0.f:164)
  static SeedState seed_par;   // pdgetv0's
  static T smlnum_first;       // dnaitr/dnapps SAVE'd smlnum from the n of the first call
};
template <typename T> SeedState Globals<T>::seed;
template <typename T> SeedState Globals<T>::seed_par;
template <typename T> T Globals<T>::smlnum_first = T(-1);

// Device back-ends are recycled between solves: creating one costs a handful of cudaMalloc / cudaMallocHost calls and
// destroying it as many cudaFree / cudaFreeHost (each an implicit device synchronisation) -- tens of milliseconds per
// solve on a busy host, all of it outside any kernel.  A finished solve parks its back-end here; the next solve on
// the same device, stream and communicator takes it over (mailbox, reduction scratch, pinned staging buffers).
template <typename T>
struct OpsPool {
  struct Item { int dev; cudaStream_t stream; NcclComm* comm; std::unique_ptr<CudaVecOps<T>> ops; };
  std::vector<Item> items;
  std::mutex mu;
  std::unique_ptr<CudaVecOps<T>> take(cudaStream_t stream, NcclComm* comm) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t i = 0; i < items.size(); ++i)
        if (items[i].dev == dev && items[i].stream == stream && items[i].comm == comm) {
          std::unique_ptr<CudaVecOps<T>> o = std::move(items[i].ops);
          items.erase(items.begin() + (long)i);
          o->reset_for_reuse();
          return o;
        }
    }
    return std::make_unique<CudaVecOps<T>>(stream, comm);
  }
  void give(std::unique_ptr<CudaVecOps<T>> o, NcclComm* comm) {
    if (!o) return;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (items.size() >= 4) items.erase(items.begin());   // oldest out (its destructor frees the device memory)
    items.push_back(Item{dev, o->stream(), comm, std::move(o)});
  }
  void clear() {
    std::lock_guard<std::mutex> lk(mu);
    items.clear();
  }
};
template <typename T>
OpsPool<T>& ops_pool() {
  static OpsPool<T> p;
  return p;
}

template <typename T>
struct Ctx {
  std::unique_ptr<CudaVecOps<T>> ops;
  NcclComm* comm_ptr = nullptr;
  std::unique_ptr<IrlSym<T>> sym;
  std::unique_ptr<IrlNonsym<T>> nonsym;
  bool par = false;
  int n = 0, ncv = 0, mode = 1;
  char bmat = 'I';
  // user arrays
  T *resid_u = nullptr, *v_u = nullptr, *workd_u = nullptr;
  int64_t ldv_u = 0;
  bool resid_host = false, v_host = false, workd_host = false;
  // device views (== user pointers when those are device memory)
  T *resid_d = nullptr, *v_d = nullptr, *workd_d = nullptr;
  int64_t ldv_d = 0;
  T* z_mirror = nullptr;
  void* csr_mirror[3] = {nullptr, nullptr, nullptr};  // HBM copies of a registered operator given as host arrays
  int last_ido = 0;
  int last_ipntr[3] = {0, 0, 0};
  bool finished = false;

  ~Ctx() {
    if (ops) {
      try { ops->sync(); } catch (...) {}
      if (resid_host) ops->release(resid_d);
      if (v_host) ops->release(v_d);
      if (workd_host) ops->release(workd_d);
      ops->release(z_mirror);
      sym.reset();       // the solvers hold a pointer to ops: gone before it changes hands
      nonsym.reset();
      ops_pool<T>().give(std::move(ops), comm_ptr);
    }
    for (void* p : csr_mirror)
      if (p) cudaFree(p);
  }
};

template <typename T>
std::unordered_map<const void*, std::unique_ptr<Ctx<T>>>& table() {
  static std::unordered_map<const void*, std::unique_ptr<Ctx<T>>> t;
  return t;
}

void require_device() {
  int cnt = 0;
  const cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0) {
    cudaGetLastError();
    throw CudaError(std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                    "); arpack_b200 has no CPU path");
  }
}

template <typename T>
Ctx<T>* make_ctx(const void* key, bool par, int comm_handle, int n, int ncv, T* resid, T* v, int ldv, T* workd,
                 bool upload_all) {
  require_device();
  auto c = std::make_unique<Ctx<T>>();
  NcclComm* comm = nullptr;
  if (par) {
    comm = comm_from_handle(comm_handle);
    if (!comm) throw CudaError("p*aupd_c: comm is not a handle returned by ab200_comm_create()");
  }
  c->ops = ops_pool<T>().take(g_stream.load(), comm);
  c->comm_ptr = comm;
  c->ops->set_kernel_mode(g_kernel_mode.load());
  c->par = par;
  c->n = n;
  c->ncv = ncv;
  c->resid_u = resid; c->v_u = v; c->workd_u = workd; c->ldv_u = ldv;
  c->resid_host = !c->ops->is_device_pointer(resid);
  c->v_host = !c->ops->is_device_pointer(v);
  c->workd_host = !c->ops->is_device_pointer(workd);
  if (n > 0 && ncv > 0) {
    c->resid_d = c->resid_host ? c->ops->alloc((size_t)n) : resid;
    if (c->v_host) {
      c->ldv_d = ((int64_t)n + 1) & ~int64_t(1);  // even leading dimension: 16-byte aligned columns
      c->v_d = c->ops->alloc((size_t)c->ldv_d * ncv);
    } else {
      c->ldv_d = ldv;
      c->v_d = v;
    }
    c->workd_d = c->workd_host ? c->ops->alloc((size_t)3 * n) : workd;
    if (upload_all) {
      if (c->resid_host) c->ops->upload(c->resid_d, resid, (size_t)n);
      if (c->v_host) c->ops->upload2d(c->v_d, (size_t)c->ldv_d, v, (size_t)ldv, (size_t)n, (size_t)ncv);
      if (c->workd_host) c->ops->upload(c->workd_d, workd, (size_t)3 * n);
    }
  }
  Ctx<T>* raw = c.get();
  std::lock_guard<std::mutex> lk(g_mu);
  // an unmodified ICB caller never calls ab200_release(): keep the table bounded by dropping finished solves
  if (table<T>().size() >= 16) {
    for (auto it = table<T>().begin(); it != table<T>().end();) {
      if (it->first != key && it->second->finished) it = table<T>().erase(it);
      else ++it;
    }
  }
  table<T>()[key] = std::move(c);
  return raw;
}

// operators registered for the solve keyed to a workl address (picked up at ido = 0)
template <typename T>
std::unordered_map<const void*, CsrOpDesc<T>>& registered_ops() {
  static std::unordered_map<const void*, CsrOpDesc<T>> t;
  return t;
}

template <typename T, typename Solver>
void attach_registered_op(Ctx<T>* c, Solver* solver, const void* key, int n) {
  CsrOpDesc<T> d;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = registered_ops<T>().find(key);
    if (it == registered_ops<T>().end()) return;
    d = it->second;
    registered_ops<T>().erase(it);  // one-shot: a later solve that reuses this workl address starts unregistered
  }
  if (d.nrows != n) throw CudaError("registered CSR operator has a different row count than the solve");
  if ((d.comm != 0) != c->par) throw CudaError("registered CSR operator: halo registration needs the p*aupd_c entry "
                                               "points (and plain registration the sequential ones)");
  CudaVecOps<T>* ops = c->ops.get();
  // a caller that owns its matrix on the host (arpackSolver.hpp reads A into host memory, :361-424) registers those
  // arrays as they are: they are copied to HBM once, here, and the copies live as long as the solve's context
  auto stage = [&](const void* host, size_t bytes, int slot) -> const void* {
    if (ops->is_device_pointer(host)) return host;
    if (d.comm != 0) throw CudaError("registered CSR operator: the halo variant takes device arrays");
    AB200_CUDA_CHECK(cudaMalloc(&c->csr_mirror[slot], bytes));
    AB200_CUDA_CHECK(cudaMemcpyAsync(c->csr_mirror[slot], host, bytes, cudaMemcpyHostToDevice, ops->stream()));
    return c->csr_mirror[slot];
  };
  d.rowptr = static_cast<const int*>(stage(d.rowptr, sizeof(int) * ((size_t)d.nrows + 1), 0));
  d.col = static_cast<const int*>(stage(d.col, sizeof(int) * (size_t)d.nnz, 1));
  d.val = static_cast<const T*>(stage(d.val, sizeof(T) * (size_t)d.nnz, 2));
  solver->set_registered_op(
      [d](const T* x, T* y) {
        if (csr_op_apply<T>(d, x, y) != 0) throw CudaError("registered CSR operator: SpMV launch failed");
      },
      [d, ops](T inv, const StepGate<T>* gate, const T* resid, T* vj, T* y, T* mb_dots) -> bool {
        if (!csr_op_fusable<T>(d)) return false;  // long rows: start_step + plain SpMV
        // multi-GPU: a gated product finishes the fused reduction of ||r'||^2 itself (CudaVecOps::attach_pending)
        StepGate<T> g;
        if (gate != nullptr) {
          g = *gate;
          ops->attach_pending(g);
          gate = &g;
        } else {
          ops->resolve_pending();
        }
        const int rc = csr_op_apply_fused<T>(d, inv, gate, ops->stop_flag(), resid, vj, y,
                                             ops->reduction_scratch(2 * 148 * 16), mb_dots, ops->reduction_ticket());
        if (rc < 0) throw CudaError("registered CSR operator: fused SpMV launch failed");
        return rc == 0;
      });
}

// A^T A operators (gram.cu) registered for the solve keyed to a workl address: handle from ab200_gram_create()
std::unordered_map<const void*, int>& registered_grams() {
  static std::unordered_map<const void*, int> t;
  return t;
}
template <typename T, typename Solver>
bool attach_gram_op(Solver* solver, const void* key) {
  int h = 0;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = registered_grams().find(key);
    if (it == registered_grams().end()) return false;
    h = it->second;
    registered_grams().erase(it);  // one-shot, like the CSR registration
  }
  if (sizeof(T) != 8) throw CudaError("registered A^T A operator: FP64 only");
  solver->set_registered_op(
      [h](const T* x, T* y) {
        if (gram_apply(h, reinterpret_cast<const double*>(x), reinterpret_cast<double*>(y)) != 0)
          throw CudaError("registered A^T A operator: product failed");
      },
      nullptr);
  return true;
}

template <typename T>
Ctx<T>* find_ctx(const void* key) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = table<T>().find(key);
  return it == table<T>().end() ? nullptr : it->second.get();
}

template <typename T>
void report(const char* where, const std::exception& e) {
  std::fprintf(stderr, "arpack_b200: %s: %s\n", where, e.what());
}

// ---- one reverse-communication call, symmetric or nonsymmetric -------------------------------
template <typename T, bool SYM>
void aupd_entry(bool par, int comm_handle, int* ido, const char* bmat, int n, const char* which, int nev, T* tol,
                T* resid, int ncv, T* v, int ldv, int* iparam, int* ipntr, T* workd, T* workl, int lworkl,
                int* info) {
  try {
    Ctx<T>* c = nullptr;
    if (*ido == 0) {
      c = make_ctx<T>(workl, par, comm_handle, n, ncv, resid, v, ldv, workd, false);
      c->bmat = bmat[0];
      c->mode = iparam[6];
      SeedState* seed = par ? &Globals<T>::seed_par : &Globals<T>::seed;
      if (SYM) c->sym = std::make_unique<IrlSym<T>>(c->ops.get(), par, seed);
      else c->nonsym = std::make_unique<IrlNonsym<T>>(c->ops.get(), par, seed, &Globals<T>::smlnum_first);
      if (iparam[6] == 1 && bmat[0] == 'I') {
        const bool gram = SYM ? attach_gram_op<T>(c->sym.get(), workl) : attach_gram_op<T>(c->nonsym.get(), workl);
        if (!gram) {
          if (SYM) attach_registered_op<T>(c, c->sym.get(), workl, n);
          else attach_registered_op<T>(c, c->nonsym.get(), workl, n);
        }
      } else {
        std::lock_guard<std::mutex> lk(g_mu);
        registered_ops<T>().erase(workl);  // not applicable to this solve (bmat='G', modes 2-5)
        registered_grams().erase(workl);
      }
      if (SYM) c->sym->set_keep_bx((g_compat & 1) != 0);
      else c->nonsym->set_keep_bx((g_compat & 1) != 0);
      // device-resident sweeps (no host round trip per step) need hand-off slots the device reaches by itself: a
      // device-resident workd, or a registered operator (then no hand-off happens at all)
      {
        const bool has_op = SYM ? c->sym->has_registered_op() : c->nonsym->has_registered_op();
        if (SYM) c->sym->set_deferral(!c->workd_host || has_op);
        else c->nonsym->set_deferral(!c->workd_host || has_op);
      }
      if (*info != 0 && c->resid_host && n > 0) c->ops->upload(c->resid_d, resid, (size_t)n);
    } else {
      c = find_ctx<T>(workl);
      if (!c || c->finished) throw CudaError("*aupd_c re-entered with ido != 0 but no active solve is keyed to this workl");
      // the user's result of the previous hand-off
      if (c->workd_host && (c->last_ido == -1 || c->last_ido == 1 || c->last_ido == 2)) {
        c->ops->upload(c->workd_d + c->last_ipntr[1] - 1, c->workd_u + c->last_ipntr[1] - 1, (size_t)c->n);
        if (SYM && c->mode == 2 && c->last_ido == 1)  // dsaupd mode 2: x was overwritten with A*x (dsaupd.f:309-313)
          c->ops->upload(c->workd_d + c->last_ipntr[0] - 1, c->workd_u + c->last_ipntr[0] - 1, (size_t)c->n);
      }
    }
    if (SYM)
      c->sym->aupd(ido, bmat[0], n, which, nev, tol, c->resid_d, ncv, c->v_d, c->ldv_d, iparam, ipntr, c->workd_d,
                   workl, lworkl, info);
    else
      c->nonsym->aupd(ido, bmat[0], n, which, nev, tol, c->resid_d, ncv, c->v_d, c->ldv_d, iparam, ipntr,
                      c->workd_d, workl, lworkl, info);
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
      {
        std::lock_guard<std::mutex> lk(g_mu);
        g_last_counters = SYM ? c->sym->counters() : c->nonsym->counters();
      }
      // argument errors return before anything was built; every other exit leaves V/resid meaningful
      if (c->ops && c->n > 0 && (*info >= 0 || *info == -8 || *info == -9 || *info == -9999)) {
        if (c->resid_host) c->ops->download(resid, c->resid_d, (size_t)c->n);
        if (c->v_host) c->ops->download2d(v, (size_t)ldv, c->v_d, (size_t)c->ldv_d, (size_t)c->n, (size_t)c->ncv);
        if (c->workd_host && c->bmat == 'G') c->ops->download(workd, c->workd_d, (size_t)3 * c->n);
      }
      c->ops->sync();
      c->finished = true;
    }
  } catch (const std::exception& e) {
    report<T>(SYM ? "[ds]saupd_c" : "[ds]naupd_c", e);
    *info = kInfoDeviceError;
    *ido = 99;
  }
}

// *eupd is the last call of a solve (it destroys V, dseupd.f:730-746): a context that owns HBM mirrors of the
// caller's host arrays -- n*(ncv+4) elements -- gives them back now instead of waiting for an ab200_release() that an
// unmodified caller of the reference never makes.  (Device-resident callers own their arrays; nothing to free.)
template <typename T>
void drop_host_mirrors(Ctx<T>* c, const void* key) {
  if (!(c->resid_host || c->v_host || c->workd_host || c->z_mirror != nullptr)) return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = table<T>().find(key);
  if (it != table<T>().end() && it->second.get() == c) table<T>().erase(it);
}

template <typename T>
struct ZView {
  T* dev = nullptr;
  int64_t ld = 0;
  bool host = false;
};

template <typename T>
ZView<T> map_z(Ctx<T>* c, T* z, int ldz, int cols, bool rvec) {
  ZView<T> zv;
  if (!rvec) return zv;
  if (z == c->v_u) {  // Z aliases V
    zv.dev = c->v_d; zv.ld = c->ldv_d; zv.host = false;
    return zv;
  }
  zv.host = !c->ops->is_device_pointer(z);
  if (zv.host) {
    zv.ld = ((int64_t)c->n + 1) & ~int64_t(1);
    c->ops->release(c->z_mirror);
    c->z_mirror = c->ops->alloc((size_t)zv.ld * cols);
    zv.dev = c->z_mirror;
  } else {
    zv.dev = z; zv.ld = ldz;
  }
  return zv;
}

template <typename T>
Ctx<T>* ctx_for_eupd(bool par, int comm_handle, int n, int ncv, T* resid, T* v, int ldv, T* workd, T* workl) {
  Ctx<T>* c = find_ctx<T>(workl);
  if (c && c->finished && c->v_u == v && c->n == n && c->ncv == ncv) return c;
  // *eupd without a preceding *aupd in this process (or with other arrays): rebuild the device view
  c = make_ctx<T>(workl, par, comm_handle, n, ncv, resid, v, ldv, workd, true);
  c->finished = true;
  return c;
}

template <typename T>
void seupd_entry(bool par, int comm_handle, int rvec, const char* howmny, const int* select, T* d, T* z, int ldz,
                 T sigma, const char* bmat, int n, const char* which, int nev, T tol, T* resid, int ncv, T* v,
                 int ldv, int* iparam, int* ipntr, T* workd, T* workl, int lworkl, int* info) {
  try {
    Ctx<T>* c = ctx_for_eupd<T>(par, comm_handle, n, ncv, resid, v, ldv, workd, workl);
    if (!c->sym) c->sym = std::make_unique<IrlSym<T>>(c->ops.get(), par, par ? &Globals<T>::seed_par : &Globals<T>::seed);
    c->sym->ensure_mailbox(ncv);
    ZView<T> zv = map_z(c, z, ldz, nev, rvec != 0);
    std::vector<int> sel(select, select + (ncv > 0 ? ncv : 0));
    c->sym->eupd(rvec != 0, howmny[0], sel.data(), d, zv.dev, zv.ld, sigma, bmat[0], n, which, nev, tol,
                 c->resid_d, ncv, c->v_d, c->ldv_d, iparam, ipntr, c->workd_d, workl, lworkl, info);
    if (rvec && *info == 0) {
      const int nconv = iparam[4];
      if (zv.host && z != c->v_u) c->ops->download2d(z, (size_t)ldz, zv.dev, (size_t)zv.ld, (size_t)n, (size_t)nconv);
      if (c->v_host) c->ops->download2d(v, (size_t)ldv, c->v_d, (size_t)c->ldv_d, (size_t)n, (size_t)ncv);
    }
    c->ops->sync();
    drop_host_mirrors<T>(c, workl);
  } catch (const std::exception& e) {
    report<T>("[ds]seupd_c", e);
    *info = kInfoDeviceError;
  }
}

template <typename T>
void neupd_entry(bool par, int comm_handle, int rvec, const char* howmny, const int* select, T* dr, T* di, T* z,
                 int ldz, T sigmar, T sigmai, T* workev, const char* bmat, int n, const char* which, int nev,
                 T tol, T* resid, int ncv, T* v, int ldv, int* iparam, int* ipntr, T* workd, T* workl,
                 int lworkl, int* info) {
  try {
    Ctx<T>* c = ctx_for_eupd<T>(par, comm_handle, n, ncv, resid, v, ldv, workd, workl);
    if (!c->nonsym)
      c->nonsym = std::make_unique<IrlNonsym<T>>(c->ops.get(), par, par ? &Globals<T>::seed_par : &Globals<T>::seed,
                                                 &Globals<T>::smlnum_first);
    c->nonsym->ensure_mailbox(ncv);
    ZView<T> zv = map_z(c, z, ldz, nev + 1, rvec != 0);
    std::vector<int> sel(select, select + (ncv > 0 ? ncv : 0));
    c->nonsym->eupd(rvec != 0, howmny[0], sel.data(), dr, di, zv.dev, zv.ld, sigmar, sigmai, workev, bmat[0], n,
                    which, nev, tol, c->resid_d, ncv, c->v_d, c->ldv_d, iparam, ipntr, c->workd_d, workl, lworkl,
                    info);
    if (rvec && *info == 0) {
      const int nconv = iparam[4];
      if (zv.host && z != c->v_u)
        c->ops->download2d(z, (size_t)ldz, zv.dev, (size_t)zv.ld, (size_t)n, (size_t)std::min(nconv, nev + 1));
      if (c->v_host) c->ops->download2d(v, (size_t)ldv, c->v_d, (size_t)c->ldv_d, (size_t)n, (size_t)ncv);
    }
    c->ops->sync();
    drop_host_mirrors<T>(c, workl);
  } catch (const std::exception& e) {
    report<T>("[ds]neupd_c", e);
    *info = kInfoDeviceError;
  }
}

}  // namespace

void set_last_counters(const Counters& c) { g_last_counters = c; }  // COMMON /timing/ as seen by stat_c
}  // namespace ab200

using namespace ab200;

extern "C" {

// ---- ICB/arpack.h ------------------------------------------------------------------------------
void dsaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid,
              a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
              a_int lworkl, a_int* info) {
  aupd_entry<double, true>(false, 0, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                           workl, lworkl, info);
}
void dseupd_c(a_int rvec, char const* howmny, a_int const* select, double* d, double* z, a_int ldz, double sigma,
              char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid, a_int ncv,
              double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl, a_int lworkl,
              a_int* info) {
  seupd_entry<double>(false, 0, rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v,
                      ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void dnaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid,
              a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
              a_int lworkl, a_int* info) {
  aupd_entry<double, false>(false, 0, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                            workl, lworkl, info);
}
void dneupd_c(a_int rvec, char const* howmny, a_int const* select, double* dr, double* di, double* z, a_int ldz,
              double sigmar, double sigmai, double* workev, char const* bmat, a_int n, char const* which,
              a_int nev, double tol, double* resid, a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr,
              double* workd, double* workl, a_int lworkl, a_int* info) {
  neupd_entry<double>(false, 0, rvec, howmny, select, dr, di, z, ldz, sigmar, sigmai, workev, bmat, n, which, nev,
                      tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void ssaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid,
              a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl,
              a_int lworkl, a_int* info) {
  aupd_entry<float, true>(false, 0, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                          workl, lworkl, info);
}
void sseupd_c(a_int rvec, char const* howmny, a_int const* select, float* d, float* z, a_int ldz, float sigma,
              char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid, a_int ncv, float* v,
              a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl, a_int lworkl, a_int* info) {
  seupd_entry<float>(false, 0, rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v,
                     ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void snaupd_c(a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid,
              a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl,
              a_int lworkl, a_int* info) {
  aupd_entry<float, false>(false, 0, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                           workl, lworkl, info);
}
void sneupd_c(a_int rvec, char const* howmny, a_int const* select, float* dr, float* di, float* z, a_int ldz,
              float sigmar, float sigmai, float* workev, char const* bmat, a_int n, char const* which, a_int nev,
              float tol, float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd,
              float* workl, a_int lworkl, a_int* info) {
  neupd_entry<float>(false, 0, rvec, howmny, select, dr, di, z, ldz, sigmar, sigmai, workev, bmat, n, which, nev,
                     tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}

// ---- ICB/parpack.h (comm = handle from ab200_comm_create) --------------------------------------
void pdsaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol,
               double* resid, a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd,
               double* workl, a_int lworkl, a_int* info) {
  aupd_entry<double, true>(true, comm, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                           workl, lworkl, info);
}
void pdseupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, double* d, double* z, a_int ldz,
               double sigma, char const* bmat, a_int n, char const* which, a_int nev, double tol, double* resid,
               a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
               a_int lworkl, a_int* info) {
  seupd_entry<double>(true, comm, rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v,
                      ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void pdnaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, double tol,
               double* resid, a_int ncv, double* v, a_int ldv, a_int* iparam, a_int* ipntr, double* workd,
               double* workl, a_int lworkl, a_int* info) {
  aupd_entry<double, false>(true, comm, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                            workl, lworkl, info);
}
void pdneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, double* dr, double* di,
               double* z, a_int ldz, double sigmar, double sigmai, double* workev, char const* bmat, a_int n,
               char const* which, a_int nev, double tol, double* resid, a_int ncv, double* v, a_int ldv,
               a_int* iparam, a_int* ipntr, double* workd, double* workl, a_int lworkl, a_int* info) {
  neupd_entry<double>(true, comm, rvec, howmny, select, dr, di, z, ldz, sigmar, sigmai, workev, bmat, n, which,
                      nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void pssaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol,
               float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd,
               float* workl, a_int lworkl, a_int* info) {
  aupd_entry<float, true>(true, comm, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                          workl, lworkl, info);
}
void psseupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, float* d, float* z, a_int ldz,
               float sigma, char const* bmat, a_int n, char const* which, a_int nev, float tol, float* resid,
               a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd, float* workl,
               a_int lworkl, a_int* info) {
  seupd_entry<float>(true, comm, rvec, howmny, select, d, z, ldz, sigma, bmat, n, which, nev, tol, resid, ncv, v,
                     ldv, iparam, ipntr, workd, workl, lworkl, info);
}
void psnaupd_c(a_fint comm, a_int* ido, char const* bmat, a_int n, char const* which, a_int nev, float tol,
               float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr, float* workd,
               float* workl, a_int lworkl, a_int* info) {
  aupd_entry<float, false>(true, comm, ido, bmat, n, which, nev, &tol, resid, ncv, v, ldv, iparam, ipntr, workd,
                           workl, lworkl, info);
}
void psneupd_c(a_fint comm, a_int rvec, char const* howmny, a_int const* select, float* dr, float* di, float* z,
               a_int ldz, float sigmar, float sigmai, float* workev, char const* bmat, a_int n, char const* which,
               a_int nev, float tol, float* resid, a_int ncv, float* v, a_int ldv, a_int* iparam, a_int* ipntr,
               float* workd, float* workl, a_int lworkl, a_int* info) {
  neupd_entry<float>(true, comm, rvec, howmny, select, dr, di, z, ldz, sigmar, sigmai, workev, bmat, n, which,
                     nev, tol, resid, ncv, v, ldv, iparam, ipntr, workd, workl, lworkl, info);
}

// ---- legacy Fortran ABI (gfortran: everything by reference, CHARACTER lengths appended) ---------
void dsaupd_(a_int* ido, const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, double* resid,
             a_int* ncv, double* v, a_int* ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
             a_int* lworkl, a_int* info, size_t, size_t) {
  aupd_entry<double, true>(false, 0, ido, bmat, *n, which, *nev, tol, resid, *ncv, v, *ldv, iparam, ipntr, workd,
                           workl, *lworkl, info);
}
void dseupd_(a_int* rvec, const char* howmny, a_int* select, double* d, double* z, a_int* ldz, double* sigma,
             const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, double* resid, a_int* ncv,
             double* v, a_int* ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl, a_int* lworkl,
             a_int* info, size_t, size_t, size_t) {
  seupd_entry<double>(false, 0, *rvec, howmny, select, d, z, *ldz, *sigma, bmat, *n, which, *nev, *tol, resid,
                      *ncv, v, *ldv, iparam, ipntr, workd, workl, *lworkl, info);
}
void dnaupd_(a_int* ido, const char* bmat, a_int* n, const char* which, a_int* nev, double* tol, double* resid,
             a_int* ncv, double* v, a_int* ldv, a_int* iparam, a_int* ipntr, double* workd, double* workl,
             a_int* lworkl, a_int* info, size_t, size_t) {
  aupd_entry<double, false>(false, 0, ido, bmat, *n, which, *nev, tol, resid, *ncv, v, *ldv, iparam, ipntr, workd,
                            workl, *lworkl, info);
}
void dneupd_(a_int* rvec, const char* howmny, a_int* select, double* dr, double* di, double* z, a_int* ldz,
             double* sigmar, double* sigmai, double* workev, const char* bmat, a_int* n, const char* which,
             a_int* nev, double* tol, double* resid, a_int* ncv, double* v, a_int* ldv, a_int* iparam,
             a_int* ipntr, double* workd, double* workl, a_int* lworkl, a_int* info, size_t, size_t, size_t) {
  neupd_entry<double>(false, 0, *rvec, howmny, select, dr, di, z, *ldz, *sigmar, *sigmai, workev, bmat, *n, which,
                      *nev, *tol, resid, *ncv, v, *ldv, iparam, ipntr, workd, workl, *lworkl, info);
}

// ---- ICB/debug_c.h, ICB/stat_c.h ---------------------------------------------------------------
void debug_c(a_int logfil, a_int ndigit, a_int mgetv0, a_int msaupd, a_int msaup2, a_int msaitr, a_int mseigt,
             a_int msapps, a_int msgets, a_int mseupd, a_int mnaupd, a_int mnaup2, a_int mnaitr, a_int mneigh,
             a_int mnapps, a_int mngets, a_int mneupd, a_int mcaupd, a_int mcaup2, a_int mcaitr, a_int mceigh,
             a_int mcapps, a_int mcgets, a_int mceupd) {
  const int vals[24] = {logfil, ndigit, mgetv0, msaupd, msaup2, msaitr, mseigt, msapps, msgets, mseupd, mnaupd, mnaup2,
                        mnaitr, mneigh, mnapps, mngets, mneupd, mcaupd, mcaup2, mcaitr, mceigh, mcapps, mcgets, mceupd};
  std::memcpy(g_debug.v, vals, sizeof(vals));
  // COMMON /debug/ as the host control code sees it (trace.hpp): same 24 integers, same order (debug.h:8-16)
  static_assert(sizeof(TraceLevels) == sizeof(vals), "TraceLevels mirrors COMMON /debug/");
  std::memcpy(static_cast<void*>(&trace_levels()), vals, sizeof(vals));
}
void sstats_c(void) { g_last_counters = Counters(); }
void sstatn_c(void) { g_last_counters = Counters(); }
void stat_c(a_int* nopx, a_int* nbx, a_int* nrorth, a_int* nitref, a_int* nrstrt, float* tsaupd, float* tsaup2,
            float* tsaitr, float* tseigt, float* tsgets, float* tsapps, float* tsconv, float* tnaupd, float* tnaup2,
            float* tnaitr, float* tneigh, float* tngets, float* tnapps, float* tnconv, float* tcaupd, float* tcaup2,
            float* tcaitr, float* tceigh, float* tcgets, float* tcapps, float* tcconv, float* tmvopx, float* tmvbx,
            float* tgetv0, float* titref, float* trvec) {
  *nopx = g_last_counters.nopx; *nbx = g_last_counters.nbx; *nrorth = g_last_counters.nrorth;
  *nitref = g_last_counters.nitref; *nrstrt = g_last_counters.nrstrt;
  // the reference's timers are dead in arpack-ng builds (UTIL/second_NONE.f:29-31): always 0
  float* ts[] = {tsaupd, tsaup2, tsaitr, tseigt, tsgets, tsapps, tsconv, tnaupd, tnaup2, tnaitr, tneigh, tngets, tnapps,
                 tnconv, tcaupd, tcaup2, tcaitr, tceigh, tcgets, tcapps, tcconv, tmvopx, tmvbx, tgetv0, titref, trvec};
  for (float* t : ts)
    if (t) *t = 0.0f;
}

// ---- extensions --------------------------------------------------------------------------------
void ab200_set_stream(void* cuda_stream) { g_stream = (cudaStream_t)cuda_stream; }
void* ab200_get_stream(void) { return (void*)g_stream.load(); }
void ab200_set_kernel_mode(int mode) { g_kernel_mode = mode; }
void ab200_set_compat(int flags) { g_compat = flags; }
void ab200_release(const void* workl) {
  release_cplx(workl);
  std::lock_guard<std::mutex> lk(g_mu);
  table<double>().erase(workl);
  table<float>().erase(workl);
  registered_ops<double>().erase(workl);
  registered_ops<float>().erase(workl);
  registered_grams().erase(workl);
}
void ab200_release_all(void) {
  release_all_cplx();
  std::lock_guard<std::mutex> lk(g_mu);
  table<double>().clear();
  table<float>().clear();
  ab200_forget_csr_cache();
  ops_pool<double>().clear();
  ops_pool<float>().clear();
  registered_ops<double>().clear();  // descriptors hold the caller's pointers: none may outlive this call
  registered_ops<float>().clear();
  registered_grams().clear();
}
// the same registration for an A^T A operator built with ab200_gram_create()/ab200_gram_add_shard() (gram.cu): the
// library applies it inside dsaupd_c / pdsaupd_c, one call runs the whole solve
int ab200_register_gram_op_f64(const void* workl, int gram_handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (gram_handle <= 0) { registered_grams().erase(workl); return 0; }
  registered_grams()[workl] = gram_handle;
  return 0;
}
void ab200_launch_stats(unsigned long long* out4) {
  const LaunchStats& s = launch_stats();
  out4[0] = s.kernels; out4[1] = s.allreduces; out4[2] = s.fast_path; out4[3] = s.fallback;
}
unsigned long long ab200_host_round_trips(void) { return launch_stats().fetches; }
void ab200_reset_seed(void) {
  Globals<double>::seed = SeedState(); Globals<double>::seed_par = SeedState();
  Globals<float>::seed = SeedState(); Globals<float>::seed_par = SeedState();
  Globals<double>::smlnum_first = -1.0; Globals<float>::smlnum_first = -1.0f;
  reset_seed_cplx();
}
// Registered-operator mode: y = A x for a square CSR matrix resident in HBM is applied by the
// library itself for the solve keyed to `workl` (mode 1, bmat = 'I'), so *aupd_c never returns ido = +-1 and one call
// runs the whole solve; start-of-step scaling and the alpha / ||w||^2 dots are fused into the SpMV kernel.
// nrows = 0 removes the registration.
int ab200_register_csr_op_f64(const void* workl, int nrows, long long nnz, const int* rowptr, const int* col,
                              const double* val) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (nrows <= 0) { registered_ops<double>().erase(workl); return 0; }
  CsrOpDesc<double> d;
  d.nrows = nrows; d.nnz = nnz; d.rowptr = rowptr; d.col = col; d.val = val;
  registered_ops<double>()[workl] = d;
  return 0;
}
int ab200_register_csr_op_f32(const void* workl, int nrows, long long nnz, const int* rowptr, const int* col,
                              const float* val) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (nrows <= 0) { registered_ops<float>().erase(workl); return 0; }
  CsrOpDesc<float> d;
  d.nrows = nrows; d.nnz = nnz; d.rowptr = rowptr; d.col = col; d.val = val;
  registered_ops<float>()[workl] = d;
  return 0;
}
// The same for a row-partitioned operator under a communicator (pdsaupd_c / pdnaupd_c): nloc local rows, columns >=
// nloc address halo_buf = [last halo_lo entries of the lower neighbour's x | first halo_hi of the upper one's], which
// the library fills with an NCCL neighbour exchange before every product (pdsdrv1.f:463-483 inside the library).
int ab200_register_csr_halo_op_f64(const void* workl, int comm, int nloc, long long nnz, const int* rowptr,
                                   const int* col, const double* val, int halo_lo, int halo_hi, double* halo_buf) {
  if (comm_from_handle(comm) == nullptr || nloc <= 0) return -1;
  std::lock_guard<std::mutex> lk(g_mu);
  CsrOpDesc<double> d;
  d.nrows = nloc; d.nnz = nnz; d.rowptr = rowptr; d.col = col; d.val = val;
  d.comm = comm; d.halo_lo = halo_lo; d.halo_hi = halo_hi; d.halo = halo_buf;
  registered_ops<double>()[workl] = d;
  return 0;
}
// largest relative disagreement between the SpMV-epilogue dots and the CGS sweep in the last registered-op solve
double ab200_fused_dot_maxdiff(const void* workl) {
  if (Ctx<double>* c = find_ctx<double>(workl)) {
    if (c->sym) return (double)c->sym->fused_dot_maxdiff;
    if (c->nonsym) return (double)c->nonsym->fused_dot_maxdiff;
  }
  if (Ctx<float>* c = find_ctx<float>(workl)) {
    if (c->sym) return (double)c->sym->fused_dot_maxdiff;
    if (c->nonsym) return (double)c->nonsym->fused_dot_maxdiff;
  }
  return -1.0;
}
void ab200_profile_enable(int on) { profiler().enabled = (on != 0); }
void ab200_profile_reset(void) { profiler().reset(); }
// returns the number of profiled kernels; entry idx (if valid) is copied out
int ab200_profile_get(int idx, char* name64, double* ms, unsigned long long* launches, double* bytes) {
  const auto& t = profiler().table();
  if (idx >= 0 && idx < (int)t.size()) {
    std::strncpy(name64, t[idx].name, 63);
    name64[63] = 0;
    *ms = t[idx].ms; *launches = t[idx].launches; *bytes = t[idx].bytes;
  }
  return (int)t.size();
}
// Kernel unit-test hooks (tests/test_gpu_kernels.py): run ONE fused orthogonalisation step / restart update on
// caller-supplied DEVICE arrays and hand the mailbox back, so that each kernel can be compared with a plain
// reference of the same op independently of the solver.
//   out_host[0..j]      = h = V_j^T w, ||w||^2
//   out_host[j+1..2j+1] = s = V_j^T r, ||r||^2     (r = w - V_j h, written to resid)
//   out_host[2j+2]      = ||r'||^2 (r' = r - V_j s), out_host[2j+3] = 1 if the DGKS pass ran (then resid = r')
int ab200_debug_orth_f64(long long n, int j, const double* v, long long ldv, const double* w, double* resid,
                         double* out_host) {
  try {
    require_device();
    CudaVecOps<double> ops(g_stream.load(), nullptr);
    ops.set_kernel_mode(g_kernel_mode.load());
    const int seg = j + 2;
    double* mb = ops.mailbox((size_t)3 * seg);
    ops.orth_step(n, j, v, ldv, w, resid, mb, mb + seg, mb + 2 * seg);
    std::vector<double> h((size_t)3 * seg);
    ops.fetch(h.data(), mb, (size_t)3 * seg);
    for (int i = 0; i <= j; ++i) out_host[i] = h[i];
    for (int i = 0; i <= j; ++i) out_host[j + 1 + i] = h[seg + i];
    out_host[2 * j + 2] = h[2 * seg];
    out_host[2 * j + 3] = h[2 * seg + 1];
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: debug_orth: %s\n", e.what());
    return -1;
  }
}
// V(:,0:kout) <- V(:,0:kin)*Q (q_host column-major kin x kout), resid <- sigma*resid + beta*Vnew(:,beta_col);
// *nrm2_host = ||resid||^2
int ab200_debug_vq_f64(long long n, int kin, int kout, double* v, long long ldv, const double* q_host, double sigma,
                       double beta, int beta_col, double* resid, double* nrm2_host) {
  try {
    require_device();
    CudaVecOps<double> ops(g_stream.load(), nullptr);
    ops.set_kernel_mode(g_kernel_mode.load());
    double* mb = ops.mailbox(8);
    ops.vq_update(n, kin, kout, v, ldv, q_host, kin, true, sigma, beta, beta_col, resid, mb);
    ops.fetch(nrm2_host, mb, 1);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: debug_vq: %s\n", e.what());
    return -1;
  }
}
// Kernel micro-benchmark hook (tools/kernel_sweep.py): run `iters` orthogonalisation steps (what = 0), multi-dots
// (1) or restart updates V <- V*Q with kout columns (2) on synthetic data of the given shape.  Timings are read back
// through the profiler table.  Returns 0 on success.
int ab200_kernel_probe_f64(long long n, int j, int ncv, int iters, int what, int kout) {
  try {
    require_device();
    CudaVecOps<double> ops(g_stream.load(), nullptr);
    ops.set_kernel_mode(g_kernel_mode.load());
    const int64_t ldv = (n + 1) & ~1LL;
    double* v = ops.alloc((size_t)ldv * ncv);
    double* w = ops.alloc((size_t)n);
    double* r = ops.alloc((size_t)n);
    double* mb = ops.mailbox((size_t)3 * (ncv + 2));
    ab200_fill_hash_f64(ldv * (long long)ncv, 0, 1234567ULL, v);
    ab200_fill_hash_f64(n, 0, 7654321ULL, w);
    ops.scal((int64_t)ldv * ncv, 1e-3, v);
    std::vector<double> q((size_t)ncv * ncv, 0.0);
    for (int c = 0; c < ncv; ++c)
      for (int k = 0; k < ncv; ++k) q[(size_t)c * ncv + k] = (k == c) ? 1.0 : 1e-3 / (1 + k + c);
    ops.sync();
    for (int it = 0; it < iters; ++it) {
      if (what == 0) ops.orth_step(n, j, v, ldv, w, r, mb, mb + ncv + 2, mb + 2 * (ncv + 2));
      else if (what == 1) ops.dots(n, j, v, ldv, w, w, mb);
      else ops.vq_update(n, ncv, kout, v, ldv, q.data(), ncv, true, 0.5, 0.25, kout - 1, r, mb);
    }
    ops.sync();
    ops.release(v); ops.release(w); ops.release(r);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: kernel_probe: %s\n", e.what());
    return -1;
  }
}
int ab200_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
  return cnt;
}
const char* ab200_version(void) { return "arpack_b200 0.1 (sm_100a)"; }
}
