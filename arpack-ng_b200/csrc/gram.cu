// gram.cu -- OP = A^T A on a row-sharded sparse A: the operator of BASELINE config 5 (SVD through dsaupd).
//
// In the reference the operator belongs to the caller: EXAMPLES/SVD/dsvd.f:342-343 applies `av` (w = A x) and then
// `atv` (y = A^T w) at every ido = 1, and the singular values are the square roots of the Ritz values (:416).  This is
// the B200 form of that pair of routines for a matrix that is too tall for one pass to be cache-friendly, or that is
// spread over several GPUs:
//
//   * A is split into row shards A_s (on one GPU: processed one after the other; on G GPUs: each rank owns some);
//     y = A^T A x = sum_s A_s^T (A_s x).  A shard is sized so that the two vectors gathered at random -- x (ncols
//     values) for A_s x and w_s = A_s x (rows of the shard) for A_s^T w_s -- stay resident in the 126 MB L2
//     (an access-policy window pins them while the CSR streams pass through).
//   * A_s^T is kept as an explicit CSR (built once on the device by a STABLE sort of the entries by column, so every
//     row of A_s^T lists its entries in ascending row order): A_s^T w as a gather is deterministic, a scatter with
//     floating-point atomics would not be, and bit-reproducible sums are what makes the solver's counts reproducible.
//   * under a communicator the eigenproblem vectors are sharded too (ncols / G entries per rank, PARPACK's layout):
//     all-gather x, local products, reduce-scatter the partial results (NCCL over NVLink).
//
// Generator: row r has exactly per_row entries, entry k at column
// splitmix64(seed + per_row*r + k) mod ncols with value 2u - 1, u = top 53 bits of splitmix64(that hash) / 2^53;
// duplicate columns in a row simply add up.  tests/problems.py holds the numpy twin used on the oracle side.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/arpack_b200.h"
#include "vecops_cuda.cuh"

namespace ab200 {
NcclComm* comm_from_handle(int handle);
void nccl_allgather(NcclComm* c, const void* send, void* recv, size_t count_per_rank, bool is_double, cudaStream_t s);
void nccl_reducescatter_sum(NcclComm* c, const void* send, void* recv, size_t count_per_rank, bool is_double,
                            cudaStream_t s);

namespace {

inline cudaStream_t cur_stream() { return (cudaStream_t)ab200_get_stream(); }

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

__global__ void k_gen_randsparse(long long row0, int nrows, int ncols, int per_row, unsigned long long seed,
                                 int* __restrict__ rowptr, int* __restrict__ col, double* __restrict__ val) {
  const long long nnz = (long long)nrows * per_row;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / per_row;
    const int k = (int)(e - r * per_row);
    const unsigned long long h1 = splitmix64(seed + (unsigned long long)per_row * (unsigned long long)(row0 + r) + k);
    const unsigned long long h2 = splitmix64(h1);
    col[e] = (int)(h1 % (unsigned long long)ncols);
    val[e] = 2.0 * ((double)(h2 >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
  }
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += (long long)gridDim.x * blockDim.x)
    rowptr[r] = (int)(r * per_row);
}

// row index of every entry of a CSR matrix
__global__ void k_expand_rows(int nrows, const int* __restrict__ rowptr, int* __restrict__ rows) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (long long)gridDim.x * blockDim.x)
    for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) rows[p] = (int)r;
}
__global__ void k_iota(long long n, int* __restrict__ x) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = (int)i;
}
// t_rowptr[c] = first position of a key >= c in the sorted key array (c = 0..ncols)
__global__ void k_lower_bounds(int ncols, long long nnz, const int* __restrict__ keys, int* __restrict__ t_rowptr) {
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c <= ncols; c += (long long)gridDim.x * blockDim.x) {
    long long lo = 0, hi = nnz;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (keys[mid] < (int)c) lo = mid + 1;
      else hi = mid;
    }
    t_rowptr[c] = (int)lo;
  }
}
__global__ void k_permute(long long nnz, const int* __restrict__ perm, const int* __restrict__ rows,
                          const double* __restrict__ val, int* __restrict__ t_col, double* __restrict__ t_val) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (long long)gridDim.x * blockDim.x) {
    const int e = perm[p];
    t_col[p] = rows[e];
    t_val[p] = val[e];
  }
}

// y (+)= A x, LPR lanes per row (power of two); entries of a row are summed lane-strided, then an xor tree: a fixed
// order, so the product is bit-reproducible.  The vector gathered at random is pinned in L2 by the launch attribute.
// gather flavours of x (AB200_GRAM_LD): 0 = plain load, 1 = read-only path (ld.global.nc), 2 = read-only and not
// allocated in L1 (every gathered sector is used once per SM: keeping it in L1 only evicts the streams)
template <int LD>
__device__ __forceinline__ double gather(const double* p) {
  if (LD == 1) return __ldg(p);
  if (LD == 2) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
  }
  return *p;
}

template <int LPR, int LD>
__global__ void __launch_bounds__(256) k_spmv_rows(int nrows, const int* __restrict__ rowptr,
                                                   const int* __restrict__ col, const double* __restrict__ val,
                                                   const double* __restrict__ x, double* __restrict__ y, int acc) {
  const int lane = threadIdx.x & (LPR - 1);
  const long long sub = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const long long nsub = ((long long)gridDim.x * blockDim.x) / LPR;
  for (long long row = sub; row < nrows; row += nsub) {
    const int p0 = rowptr[row], p1 = rowptr[row + 1];
    double a = 0.0;
    for (int p = p0 + lane; p < p1; p += LPR) a += __ldcs(val + p) * gather<LD>(x + __ldcs(col + p));
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o, LPR);
    if (lane == 0) y[row] = acc ? y[row] + a : a;
  }
}

struct L2Window {
  size_t max_window = 0;
  bool ready = false;
};
L2Window& l2window() {
  static L2Window w;
  if (!w.ready) {
    w.ready = true;
    static const bool off = getenv("AB200_L2_WINDOW") && std::strcmp(getenv("AB200_L2_WINDOW"), "0") == 0;
    (void)off;
    int dev = 0, max_persist = 0, max_win = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (!off && max_persist > 0 && max_win > 0 &&
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess)
      w.max_window = (size_t)(max_win < max_persist ? max_win : max_persist);
    cudaGetLastError();
  }
  return w;
}

// Measured on the full-size config-5 shards (B200, 2.5 M x 5 M, 40 M entries; tools/gram_probe.py): the product is
// bound by the L2 -> SM sector traffic of the 8-byte gathers (ncu: DRAM bytes = 1.00 x algorithmic, every gathered
// sector delivers 8 of its 32 bytes, lts throughput 55 %, l1tex 65 %), not by DRAM.  Few lanes per row win (A x, 16
// entries per row: 8 lanes 2.7 TB/s, 16 lanes 2.2, 32 lanes 1.5, 2 lanes 2.0; A^T w, ~8 per row: 4 lanes 2.5, 8 lanes
// 1.9) and gathers that bypass L1 help a little -- but only together with the L2 access-policy window on the gathered
// vector: once the device has a persisting set-aside, un-windowed gathers of the other product lose half their rate
// (1.2 TB/s inside a solve), so both products pin their vector (2.6 / 2.2 TB/s).
int launch_rows(cudaStream_t s, int nrows, long long nnz, const int* rowptr, const int* col, const double* val,
                const double* x, size_t x_bytes, double* y, int acc, const char* name, bool pin_x) {
  if (nrows <= 0) return 0;
  const double avg = (double)nnz / nrows;
  int lpr = avg <= 12.0 ? 4 : avg <= 24.0 ? 8 : avg <= 48.0 ? 16 : 32;
  static const int lpr_env = getenv("AB200_GRAM_LPR") ? atoi(getenv("AB200_GRAM_LPR")) : 0;   // tuning knobs
  static const int ld_env = getenv("AB200_GRAM_LD") ? atoi(getenv("AB200_GRAM_LD")) : 2;
  if (lpr_env == 1 || lpr_env == 2 || lpr_env == 4 || lpr_env == 8 || lpr_env == 16 || lpr_env == 32) lpr = lpr_env;
  long long g = ((long long)nrows * lpr + 255) / 256;
  const long long cap = 148LL * 32;
  g = g > cap ? cap : (g < 1 ? 1 : g);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)g);
  cfg.blockDim = dim3(256);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int nattr = 0;
  const size_t win = l2window().max_window;
  if (win > 0 && x_bytes > 0 && pin_x) {
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow.base_ptr = const_cast<double*>(x);
    attr[0].val.accessPolicyWindow.num_bytes = x_bytes < win ? x_bytes : win;
    attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    nattr = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  ProfScope ps(s, name, (double)nnz * 12.0 + (nrows + 1) * 4.0 + (double)x_bytes + (acc ? 16.0 : 8.0) * nrows);
  cudaError_t e;
#define AB200_ROWS(LPR_)                                                                                      \
  (ld_env == 2   ? cudaLaunchKernelEx(&cfg, k_spmv_rows<LPR_, 2>, nrows, rowptr, col, val, x, y, acc)          \
   : ld_env == 1 ? cudaLaunchKernelEx(&cfg, k_spmv_rows<LPR_, 1>, nrows, rowptr, col, val, x, y, acc)          \
                 : cudaLaunchKernelEx(&cfg, k_spmv_rows<LPR_, 0>, nrows, rowptr, col, val, x, y, acc))
  switch (lpr) {
    case 1: e = AB200_ROWS(1); break;
    case 2: e = AB200_ROWS(2); break;
    case 4: e = AB200_ROWS(4); break;
    case 8: e = AB200_ROWS(8); break;
    case 16: e = AB200_ROWS(16); break;
    default: e = AB200_ROWS(32); break;
  }
#undef AB200_ROWS
  launch_stats().kernels++;
  return e == cudaSuccess ? 0 : -1;
}

struct Shard {
  int nrows = 0;
  long long nnz = 0;
  const int *rowptr = nullptr, *col = nullptr;
  const double* val = nullptr;
  const int *t_rowptr = nullptr, *t_col = nullptr;
  const double* t_val = nullptr;
};
struct Gram {
  int comm = 0, ncols = 0, nranks = 1;
  std::vector<Shard> shards;
  double *xfull = nullptr, *zfull = nullptr, *w = nullptr;
  size_t w_cap = 0;
  ~Gram() {
    cudaFree(xfull);
    cudaFree(zfull);
    cudaFree(w);
  }
};
std::mutex g_mu;
std::vector<std::unique_ptr<Gram>> g_grams;  // handle = index + 1

Gram* gram_from_handle(int h) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (h < 1 || h > (int)g_grams.size()) return nullptr;
  return g_grams[h - 1].get();
}

inline int gen_grid(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g > 148LL * 16 ? 148LL * 16 : (g < 1 ? 1 : g));
}

}  // namespace

// used by api.cu for the registered form (the library applies the operator inside *aupd_c)
int gram_apply(int handle, const double* x_loc, double* z_loc) {
  Gram* G = gram_from_handle(handle);
  if (!G || G->shards.empty()) return -1;
  cudaStream_t s = cur_stream();
  try {
    NcclComm* c = G->comm ? comm_from_handle(G->comm) : nullptr;
    if (G->comm && !c) return -2;
    const double* x = x_loc;
    double* z = z_loc;
    const size_t per_rank = (size_t)G->ncols / (size_t)G->nranks;
    if (c && G->nranks > 1) {
      nccl_allgather(c, x_loc, G->xfull, per_rank, true, s);  // every rank needs all of x for its rows of A
      x = G->xfull;
      z = G->zfull;
    }
    for (size_t i = 0; i < G->shards.size(); ++i) {
      const Shard& sh = G->shards[i];
      if (launch_rows(s, sh.nrows, sh.nnz, sh.rowptr, sh.col, sh.val, x, sizeof(double) * G->ncols, G->w, 0,
                      "gram_Ax", true) != 0)
        return -3;
      if (launch_rows(s, G->ncols, sh.nnz, sh.t_rowptr, sh.t_col, sh.t_val, G->w, sizeof(double) * sh.nrows, z,
                      i > 0 ? 1 : 0, "gram_ATw", true) != 0)
        return -4;
    }
    if (c && G->nranks > 1) nccl_reducescatter_sum(c, G->zfull, z_loc, per_rank, true, s);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: gram operator: %s\n", e.what());
    return -5;
  }
  return 0;
}

}  // namespace ab200

using namespace ab200;

extern "C" {

long long ab200_gen_randsparse(long long row0, int nrows, int ncols, int per_row, unsigned long long seed, int* rowptr,
                               int* col, double* val) {
  const long long nnz = (long long)nrows * per_row;
  if (nnz > 2147483647LL || ncols <= 0 || per_row <= 0) return -2;
  if (!rowptr) return nnz;
  k_gen_randsparse<<<gen_grid(nnz), 256, 0, cur_stream()>>>(row0, nrows, ncols, per_row, seed, rowptr, col, val);
  launch_stats().kernels++;
  return cudaGetLastError() == cudaSuccess ? nnz : -1;
}

int ab200_csr_transpose_f64(int nrows, int ncols, long long nnz, const int* rowptr, const int* col, const double* val,
                            int* t_rowptr, int* t_col, double* t_val) {
  if (nnz > 2147483647LL) return -2;
  cudaStream_t s = cur_stream();
  int *rows = nullptr, *idx = nullptr, *keys_out = nullptr, *idx_out = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = 0;
  int end_bit = 1;
  while (end_bit < 31 && (1LL << end_bit) < (long long)ncols) ++end_bit;
  if (cudaMalloc(&rows, 4 * (size_t)nnz) != cudaSuccess || cudaMalloc(&idx, 4 * (size_t)nnz) != cudaSuccess ||
      cudaMalloc(&keys_out, 4 * (size_t)nnz) != cudaSuccess || cudaMalloc(&idx_out, 4 * (size_t)nnz) != cudaSuccess) {
    rc = -1;
  }
  if (rc == 0) {
    k_expand_rows<<<gen_grid(nrows), 256, 0, s>>>(nrows, rowptr, rows);
    k_iota<<<gen_grid(nnz), 256, 0, s>>>(nnz, idx);
    // stable LSD radix sort of (column, entry index): entries of one column keep their row order
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, col, keys_out, idx, idx_out, (int)nnz, 0, end_bit, s);
    if (cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1) != cudaSuccess) rc = -1;
  }
  if (rc == 0) {
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, col, keys_out, idx, idx_out, (int)nnz, 0, end_bit, s);
    k_lower_bounds<<<gen_grid((long long)ncols + 1), 256, 0, s>>>(ncols, nnz, keys_out, t_rowptr);
    k_permute<<<gen_grid(nnz), 256, 0, s>>>(nnz, idx_out, rows, val, t_col, t_val);
    launch_stats().kernels += 5;
    if (cudaStreamSynchronize(s) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = -3;
  }
  cudaFree(rows); cudaFree(idx); cudaFree(keys_out); cudaFree(idx_out); cudaFree(tmp);
  if (rc != 0) cudaGetLastError();
  return rc;
}

int ab200_gram_create(int comm, int ncols) {
  auto g = std::make_unique<Gram>();
  g->comm = comm;
  g->ncols = ncols;
  if (comm != 0) {
    NcclComm* c = comm_from_handle(comm);
    if (!c) return -1;
    g->nranks = nccl_nranks(c);
    if (ncols % g->nranks != 0) return -2;  // equal slices of the eigenproblem vectors per rank
    if (g->nranks > 1) {
      if (cudaMalloc(&g->xfull, sizeof(double) * (size_t)ncols) != cudaSuccess ||
          cudaMalloc(&g->zfull, sizeof(double) * (size_t)ncols) != cudaSuccess) {
        cudaGetLastError();
        return -3;
      }
    }
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_grams.push_back(std::move(g));
  return (int)g_grams.size();
}

int ab200_gram_add_shard(int handle, int nrows, long long nnz, const int* rowptr, const int* col, const double* val,
                         const int* t_rowptr, const int* t_col, const double* t_val) {
  Gram* G = gram_from_handle(handle);
  if (!G || nrows <= 0) return -1;
  Shard sh;
  sh.nrows = nrows; sh.nnz = nnz; sh.rowptr = rowptr; sh.col = col; sh.val = val;
  sh.t_rowptr = t_rowptr; sh.t_col = t_col; sh.t_val = t_val;
  if ((size_t)nrows > G->w_cap) {
    cudaFree(G->w);
    G->w = nullptr;
    if (cudaMalloc(&G->w, sizeof(double) * (size_t)nrows) != cudaSuccess) { cudaGetLastError(); return -2; }
    G->w_cap = (size_t)nrows;
  }
  G->shards.push_back(sh);
  return 0;
}

int ab200_gram_apply(int handle, const double* x_loc, double* z_loc) { return gram_apply(handle, x_loc, z_loc); }

void ab200_gram_destroy(int handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (handle >= 1 && handle <= (int)g_grams.size()) g_grams[handle - 1].reset();
}
}
