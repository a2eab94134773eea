// vecops_cplx.cu -- sm_100a kernels behind VecOps<std::complex<R>> (complex Arnoldi).
//
// Reference call sites replaced (SRC/znaitr.f, znapps.f, zgetv0.f, zneupd.f):
//   k_zdots    zgemv('C') h = V_j^H w + zzdotc / dznrm2     (znaitr.f:529-560, :637-640; zgetv0.f:326-327)
//   k_zupdate  zgemv('N') r = w - V_j h + dznrm2             (znaitr.f:561-562, :641-642, :587-597)
//   k_zvq      kev x (zgemv + zcopy), zlacpy, zscal + zaxpy, dznrm2   (znapps.f:443-485, znaup2.f:742-747);
//              zunm2r / ztrmm on the n x ncv arrays          (zneupd.f:666-670, :760-763)
//   k_zger     zgeru purification                            (zneupd.f:868)
// One complex element is one 16-byte (8-byte for single) vector load; rows are contiguous across the threads of a
// warp, so every access to V is fully coalesced.  Reductions are deterministic: fixed grid for a given n, shuffle
// tree inside the warp, per-CTA partials summed by the last CTA in index order -- no floating-point atomics.
// These are the first, generic kernels of the complex path (any ldv, any ncv); the TMA-tiled forms the real path
// uses (vecops_tma.cu) are not instantiated for complex data yet.
#include <algorithm>

#include "vecops_cplx.cuh"

namespace ab200 {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <typename R> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

template <typename C, typename R>
__device__ __forceinline__ C mk(R x, R y) {
  C c;
  c.x = x;
  c.y = y;
  return c;
}
// acc += conj(a) * b
template <typename C>
__device__ __forceinline__ void fma_conj(C& acc, const C a, const C b) {
  acc.x += a.x * b.x + a.y * b.y;
  acc.y += a.x * b.y - a.y * b.x;
}
// acc += a * b
template <typename C>
__device__ __forceinline__ void fma_cplx(C& acc, const C a, const C b) {
  acc.x += a.x * b.x - a.y * b.y;
  acc.y += a.x * b.y + a.y * b.x;
}
template <typename C>
__device__ __forceinline__ C warp_sum2(C v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// The CTA that takes the last ticket sums partial[b*pcols + c] over b in a fixed order and writes out[c].
template <typename C>
__device__ void finish_grid_reduce2(C* partial, int pcols, int ncols, C* out, unsigned int* ticket) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  for (int c = warp; c < ncols; c += nwarps) {
    C s = mk<C>(decltype(s.x)(0), decltype(s.x)(0));
    for (int b = lane; b < (int)gridDim.x; b += 32) {
      const C p = __ldcg(partial + (size_t)b * pcols + c);
      s.x += p.x;
      s.y += p.y;
    }
    s = warp_sum2(s);
    if (lane == 0) out[c] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// out[0..j) = V_j^H x, out[j] = sum conj(x_i) y_i.  Column chunks of CC accumulate in registers while the CTA streams
// its contiguous row range; x is re-read once per chunk (from L2).
template <typename R, int CC>
__global__ void __launch_bounds__(kThreads) k_zdots(int64_t n, int j, const typename Vec2<R>::type* __restrict__ v,
                                                    int64_t ldv, const typename Vec2<R>::type* __restrict__ x,
                                                    const typename Vec2<R>::type* __restrict__ y,
                                                    typename Vec2<R>::type* __restrict__ partial, int pcols,
                                                    typename Vec2<R>::type* __restrict__ out, unsigned int* ticket) {
  using C = typename Vec2<R>::type;
  __shared__ C red[kWarps][CC + 1];
  const int64_t rpc = ((n + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
  const int64_t r0 = (int64_t)blockIdx.x * rpc;
  const int64_t r1 = (r0 + rpc < n) ? r0 + rpc : n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  C* mine = partial + (size_t)blockIdx.x * pcols;
  for (int c0 = 0; c0 < j || c0 == 0; c0 += CC) {
    C acc[CC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) acc[cc] = mk<C>(R(0), R(0));
    C accxy = mk<C>(R(0), R(0));
    const C* vc = v + (int64_t)c0 * ldv;
    const int ncol = (j - c0 < CC) ? (j - c0) : CC;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += kThreads) {
      const C xv = x[r];
      if (c0 == 0) fma_conj(accxy, xv, y[r]);
#pragma unroll
      for (int cc = 0; cc < CC; ++cc)
        if (cc < ncol) fma_conj(acc[cc], vc[r + (int64_t)cc * ldv], xv);
    }
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      const C s = warp_sum2(acc[cc]);
      if (lane == 0) red[warp][cc] = s;
    }
    {
      const C s = warp_sum2(accxy);
      if (lane == 0) red[warp][CC] = s;
    }
    __syncthreads();
    if (threadIdx.x <= CC) {
      C s = mk<C>(R(0), R(0));
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        s.x += red[w][threadIdx.x].x;
        s.y += red[w][threadIdx.x].y;
      }
      if (threadIdx.x < CC) {
        if (c0 + (int)threadIdx.x < j) mine[c0 + threadIdx.x] = s;
      } else if (c0 == 0) {
        mine[j] = s;
      }
    }
    __syncthreads();
    if (j == 0) break;
  }
  finish_grid_reduce2(partial, pcols, j + 1, out, ticket);
}

// dst = src - V_j*coef ; nrm2_out = (sum |dst|^2, 0)
template <typename R>
__global__ void __launch_bounds__(kThreads) k_zupdate(int64_t n, int j, const typename Vec2<R>::type* __restrict__ v,
                                                      int64_t ldv, const typename Vec2<R>::type* __restrict__ coef,
                                                      const typename Vec2<R>::type* src, typename Vec2<R>::type* dst,
                                                      typename Vec2<R>::type* __restrict__ partial,
                                                      typename Vec2<R>::type* __restrict__ nrm2_out,
                                                      unsigned int* ticket) {
  using C = typename Vec2<R>::type;
  extern __shared__ unsigned char smem_raw[];
  C* cs = reinterpret_cast<C*>(smem_raw);
  __shared__ R red[kWarps];
  for (int k = threadIdx.x; k < j; k += kThreads) cs[k] = coef[k];
  __syncthreads();
  R nrm = R(0);
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    const C* vr = v + r;
    C a0 = mk<C>(R(0), R(0)), a1 = a0, a2 = a0, a3 = a0;
    int k = 0;
    for (; k + 4 <= j; k += 4) {
      const C v0 = vr[(int64_t)(k + 0) * ldv], v1 = vr[(int64_t)(k + 1) * ldv];
      const C v2 = vr[(int64_t)(k + 2) * ldv], v3 = vr[(int64_t)(k + 3) * ldv];
      fma_cplx(a0, v0, cs[k + 0]);
      fma_cplx(a1, v1, cs[k + 1]);
      fma_cplx(a2, v2, cs[k + 2]);
      fma_cplx(a3, v3, cs[k + 3]);
    }
    for (; k < j; ++k) fma_cplx(a0, vr[(int64_t)k * ldv], cs[k]);
    const C s = src[r];
    const C d = mk<C>(s.x - ((a0.x + a1.x) + (a2.x + a3.x)), s.y - ((a0.y + a1.y) + (a2.y + a3.y)));
    dst[r] = d;
    nrm += d.x * d.x + d.y * d.y;
  }
  if (nrm2_out == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  if (lane == 0) red[warp] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    R s = R(0);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w];
    partial[blockIdx.x] = mk<C>(s, R(0));
  }
  finish_grid_reduce2(partial, 1, 1, nrm2_out, ticket);
}

// out = sum conj(x_i) y_i
template <typename R>
__global__ void __launch_bounds__(kThreads) k_zdotc(int64_t n, const typename Vec2<R>::type* __restrict__ x,
                                                    const typename Vec2<R>::type* __restrict__ y,
                                                    typename Vec2<R>::type* __restrict__ partial,
                                                    typename Vec2<R>::type* __restrict__ out, unsigned int* ticket) {
  using C = typename Vec2<R>::type;
  __shared__ C red[kWarps];
  const int64_t rpc = ((n + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
  const int64_t r0 = (int64_t)blockIdx.x * rpc;
  const int64_t r1 = (r0 + rpc < n) ? r0 + rpc : n;
  C acc = mk<C>(R(0), R(0));
  for (int64_t r = r0 + threadIdx.x; r < r1; r += kThreads) fma_conj(acc, x[r], y[r]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  acc = warp_sum2(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    C s = mk<C>(R(0), R(0));
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      s.x += red[w].x;
      s.y += red[w].y;
    }
    partial[blockIdx.x] = s;
  }
  finish_grid_reduce2(partial, 1, 1, out, ticket);
}

// y = a*y + b*x (+ sum |y|^2)
template <typename R>
__global__ void __launch_bounds__(kThreads) k_zaxpby_norm(int64_t n, typename Vec2<R>::type a,
                                                          typename Vec2<R>::type b,
                                                          const typename Vec2<R>::type* __restrict__ x,
                                                          typename Vec2<R>::type* y,
                                                          typename Vec2<R>::type* __restrict__ partial,
                                                          typename Vec2<R>::type* __restrict__ out,
                                                          unsigned int* ticket) {
  using C = typename Vec2<R>::type;
  __shared__ R red[kWarps];
  const int64_t rpc = ((n + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
  const int64_t r0 = (int64_t)blockIdx.x * rpc;
  const int64_t r1 = (r0 + rpc < n) ? r0 + rpc : n;
  R acc = R(0);
  for (int64_t r = r0 + threadIdx.x; r < r1; r += kThreads) {
    C t = mk<C>(R(0), R(0));
    fma_cplx(t, a, y[r]);
    if (x != nullptr) fma_cplx(t, b, x[r]);
    y[r] = t;
    acc += t.x * t.x + t.y * t.y;
  }
  if (out == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    R s = R(0);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w];
    partial[blockIdx.x] = mk<C>(s, R(0));
  }
  finish_grid_reduce2(partial, 1, 1, out, ticket);
}

template <typename R>
__global__ void k_zscal(int64_t n, typename Vec2<R>::type alpha, typename Vec2<R>::type* x) {
  using C = typename Vec2<R>::type;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    C t = mk<C>(R(0), R(0));
    fma_cplx(t, alpha, x[r]);
    x[r] = t;
  }
}

// Z(:,0:k) += resid * w^T (no conjugate: zgeru)
template <typename R>
__global__ void k_zger(int64_t n, int k, const typename Vec2<R>::type* __restrict__ resid,
                       const typename Vec2<R>::type* __restrict__ w, typename Vec2<R>::type* z, int64_t ldz) {
  using C = typename Vec2<R>::type;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const C rv = resid[r];
    for (int c = 0; c < k; ++c) {
      C t = z[r + (int64_t)c * ldz];
      fma_cplx(t, rv, w[c]);
      z[r + (int64_t)c * ldz] = t;
    }
  }
}

// out(:,0:kout) = V(:,0:kin) * Q (kin x kout, column-major, packed), `rows` (32 or 64) rows per tile.  The whole
// tile of V is staged in shared memory (tile[k][r]: a warp reads 32 consecutive rows of one column, conflict-free)
// before anything is written, so out may alias V.  The 256 threads of a CTA form 256/rows column groups; a thread
// owns one row and computes two output columns per pass (one shared-memory read feeds two complex FMAs).  The tile is
// small (kin KB at 64 rows), so several CTAs are resident per SM and one CTA's loads overlap another's arithmetic.
// Optional fused resid = sigma*resid + beta*out(:,beta_col) and its squared norm (done by the thread that owns
// column beta_col of the row; by column group 0 when there is no beta term).
template <typename R>
__global__ void __launch_bounds__(kThreads) k_zvq(int64_t n, int kin, int kout, int rows,
                                                  const typename Vec2<R>::type* v, int64_t ldv,
                                                  const typename Vec2<R>::type* __restrict__ q,
                                                  typename Vec2<R>::type* out, int64_t ldo, int with_resid,
                                                  typename Vec2<R>::type sigma, typename Vec2<R>::type beta,
                                                  int beta_col, typename Vec2<R>::type* resid,
                                                  typename Vec2<R>::type* __restrict__ partial,
                                                  typename Vec2<R>::type* __restrict__ nrm2_out,
                                                  unsigned int* ticket) {
  using C = typename Vec2<R>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* tile = reinterpret_cast<C*>(smem_raw);
  __shared__ R red[kWarps];
  R nrm = R(0);
  const int ngroups = kThreads / rows;        // 4 (rows = 64) or 8 (rows = 32)
  const int r = threadIdx.x % rows;           // my row inside the tile
  const int g = threadIdx.x / rows;           // my column group
  const int owner_col = (beta_col >= 0) ? beta_col : 0;
  const int64_t ntiles = (n + rows - 1) / rows;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t r0 = t * rows;
    const int nr = (int)((n - r0 < rows) ? (n - r0) : rows);
    __syncthreads();  // the previous tile has been consumed
    for (int idx = threadIdx.x; idx < kin * rows; idx += kThreads) {
      const int k = idx / rows, rr = idx - k * rows;
      if (rr < nr) tile[idx] = v[r0 + rr + (int64_t)k * ldv];
    }
    __syncthreads();
    if (r < nr) {
      for (int c0 = 2 * g; c0 < kout; c0 += 2 * ngroups) {
        const bool two = (c0 + 1 < kout);
        const C* q0 = q + (size_t)c0 * kin;
        const C* q1 = q0 + (two ? kin : 0);
        C a0 = mk<C>(R(0), R(0)), a1 = a0;
        for (int k = 0; k < kin; ++k) {
          const C tv = tile[k * rows + r];
          fma_cplx(a0, tv, __ldg(q0 + k));
          fma_cplx(a1, tv, __ldg(q1 + k));
        }
        out[r0 + r + (int64_t)c0 * ldo] = a0;
        if (two) out[r0 + r + (int64_t)(c0 + 1) * ldo] = a1;
        if (with_resid && (c0 == owner_col || (two && c0 + 1 == owner_col))) {
          const C bval = (c0 == owner_col) ? a0 : a1;
          C tr = mk<C>(R(0), R(0));
          fma_cplx(tr, sigma, resid[r0 + r]);
          if (beta_col >= 0) fma_cplx(tr, beta, bval);
          resid[r0 + r] = tr;
          nrm += tr.x * tr.x + tr.y * tr.y;
        }
      }
    }
  }
  if (!with_resid || nrm2_out == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
  if (lane == 0) red[warp] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    R s2 = R(0);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s2 += red[w];
    partial[blockIdx.x] = mk<C>(s2, R(0));
  }
  finish_grid_reduce2(partial, 1, 1, nrm2_out, ticket);
}

#define AB200_LAUNCHED()                  \
  do {                                    \
    launch_stats().kernels++;             \
    AB200_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

template <typename R>
typename Vec2<R>::type to_c(std::complex<R> z) {
  typename Vec2<R>::type c;
  c.x = z.real();
  c.y = z.imag();
  return c;
}

}  // namespace

template <typename R>
int CudaVecOpsZ<R>::reduce_grid(int64_t n) const {
  // a pure function of n (and the SM count): keeps reductions bit-reproducible
  const int64_t want = (n + 4 * kThreads - 1) / (4 * kThreads);
  const int64_t cap = (int64_t)num_sms_ * 4;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template <typename R>
std::complex<R>* CudaVecOpsZ<R>::stage_matrix(const T* host, int rows, int cols, int ld) {
  const size_t cnt = (size_t)rows * cols;
  if (cnt > qbuf_count_) {
    AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
    cudaFree(qbuf_);
    AB200_CUDA_CHECK(cudaMalloc(&qbuf_, sizeof(T) * cnt));
    qbuf_count_ = cnt;
  }
  // the host matrix lives in caller-owned (pageable) memory that may change right after this call: pack it and copy
  // synchronously with respect to the host (cudaMemcpyAsync from pageable memory returns after staging)
  std::vector<T> packed(cnt);
  for (int c = 0; c < cols; ++c) std::copy(host + (size_t)c * ld, host + (size_t)c * ld + rows, packed.begin() + (size_t)c * rows);
  AB200_CUDA_CHECK(cudaMemcpyAsync(qbuf_, packed.data(), sizeof(T) * cnt, cudaMemcpyHostToDevice, stream_));
  AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
  return qbuf_;
}

template <typename R>
void CudaVecOpsZ<R>::scal(int64_t n, T alpha, T* x) {
  using C = typename Vec2<R>::type;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  k_zscal<R><<<grid, 256, 0, stream_>>>(n, to_c<R>(alpha), reinterpret_cast<C*>(x));
  AB200_LAUNCHED();
}

template <typename R>
void CudaVecOpsZ<R>::axpby_norm(int64_t n, T a, T b, const T* x, T* y, T* out) {
  using C = typename Vec2<R>::type;
  const int grid = reduce_grid(n);
  C* part = reinterpret_cast<C*>(partial((size_t)grid));
  ProfScope ps(stream_, "zaxpby_norm", (double)sizeof(T) * n * (x ? 3.0 : 2.0));
  k_zaxpby_norm<R><<<grid, kThreads, 0, stream_>>>(n, to_c<R>(a), to_c<R>(b), reinterpret_cast<const C*>(x),
                                                  reinterpret_cast<C*>(y), part, reinterpret_cast<C*>(out),
                                                  real_.reduction_ticket());
  AB200_LAUNCHED();
}

template <typename R>
void CudaVecOpsZ<R>::dot(int64_t n, const T* x, const T* y, T* out) {
  using C = typename Vec2<R>::type;
  const int grid = reduce_grid(n);
  C* part = reinterpret_cast<C*>(partial((size_t)grid));
  ProfScope ps(stream_, "zdotc", (double)sizeof(T) * n * (x == y ? 1.0 : 2.0));
  k_zdotc<R><<<grid, kThreads, 0, stream_>>>(n, reinterpret_cast<const C*>(x), reinterpret_cast<const C*>(y), part,
                                            reinterpret_cast<C*>(out), real_.reduction_ticket());
  AB200_LAUNCHED();
}

template <typename R>
void CudaVecOpsZ<R>::ger(int64_t n, int k, const T* resid, const T* w_host, T* z, int64_t ldz) {
  using C = typename Vec2<R>::type;
  if (k <= 0) return;
  const T* wdev = stage_matrix(w_host, k, 1, k);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  k_zger<R><<<grid, 256, 0, stream_>>>(n, k, reinterpret_cast<const C*>(resid), reinterpret_cast<const C*>(wdev),
                                       reinterpret_cast<C*>(z), ldz);
  AB200_LAUNCHED();
}

template <typename R>
void CudaVecOpsZ<R>::dots(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* out) {
  using C = typename Vec2<R>::type;
  constexpr int CC = 8;  // columns per register chunk: x is re-read once per chunk (12.5 % of the V traffic, from L2)
  const int grid = reduce_grid(n);
  const int pcols = j + 1;
  C* part = reinterpret_cast<C*>(partial((size_t)grid * pcols));
  ProfScope ps(stream_, "zdots", (double)sizeof(T) * n * (j + (x == y ? 1.0 : 2.0)));
  k_zdots<R, CC><<<grid, kThreads, 0, stream_>>>(n, j, reinterpret_cast<const C*>(v), ldv,
                                                 reinterpret_cast<const C*>(x), reinterpret_cast<const C*>(y), part,
                                                 pcols, reinterpret_cast<C*>(out), real_.reduction_ticket());
  AB200_LAUNCHED();
}

template <typename R>
void CudaVecOpsZ<R>::update(int64_t n, int j, const T* v, int64_t ldv, const T* coef, const T* src, T* dst, T* nrm2) {
  using C = typename Vec2<R>::type;
  const int grid = reduce_grid(n);
  C* part = reinterpret_cast<C*>(partial((size_t)grid));
  ProfScope ps(stream_, "zupdate", (double)sizeof(T) * n * (j + 2.0));
  k_zupdate<R><<<grid, kThreads, sizeof(T) * (size_t)std::max(j, 1), stream_>>>(
      n, j, reinterpret_cast<const C*>(v), ldv, reinterpret_cast<const C*>(coef), reinterpret_cast<const C*>(src),
      reinterpret_cast<C*>(dst), part, reinterpret_cast<C*>(nrm2), real_.reduction_ticket());
  AB200_LAUNCHED();
}

template <typename R>
void CudaVecOpsZ<R>::vq(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* qdev, T* out, int64_t ldo,
                        bool with_resid, T sigma, T beta, int beta_col, T* resid, T* nrm2) {
  using C = typename Vec2<R>::type;
  if (kin <= 0 || kout <= 0) {
    if (with_resid) axpby_norm(n, sigma, T(0), nullptr, resid, nrm2);
    return;
  }
  // 64 rows per tile (32 beyond 200 columns): the tile takes kin KB of shared memory, so up to 8 CTAs share an SM
  const size_t budget = 200 * 1024;
  int rows = 64;
  if (sizeof(T) * (size_t)kin * rows > budget) rows = 32;
  const size_t smem = sizeof(T) * (size_t)kin * rows;
  if (smem > budget) throw CudaError("complex V*Q: ncv too large for the shared-memory tile (ncv <= 400 supported)");
  static size_t attr_set[2] = {0, 0};
  size_t& cur = attr_set[sizeof(R) == 8 ? 0 : 1];
  if (smem > 48 * 1024 && cur < budget) {
    AB200_CUDA_CHECK(cudaFuncSetAttribute(k_zvq<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    cur = budget;
  }
  const int64_t ntiles = (n + rows - 1) / rows;
  int per_sm = (int)std::min<size_t>(8, (220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<int64_t>(ntiles, (int64_t)num_sms_ * per_sm);
  C* part = reinterpret_cast<C*>(partial((size_t)grid));
  ProfScope ps(stream_, "zvq", (double)sizeof(T) * n * (kin + kout + (with_resid ? 2.0 : 0.0)));
  k_zvq<R><<<grid, kThreads, smem, stream_>>>(n, kin, kout, rows, reinterpret_cast<const C*>(v), ldv,
                                              reinterpret_cast<const C*>(qdev), reinterpret_cast<C*>(out), ldo,
                                              with_resid ? 1 : 0, to_c<R>(sigma), to_c<R>(beta), beta_col,
                                              reinterpret_cast<C*>(resid), part, reinterpret_cast<C*>(nrm2),
                                              real_.reduction_ticket());
  AB200_LAUNCHED();
}

template <typename R>
void CudaVecOpsZ<R>::vq_update(int64_t n, int kin, int kout, T* v, int64_t ldv, const T* q_host, int ldq,
                               bool with_resid, T sigma, T beta, int beta_col, T* resid, T* nrm2) {
  const T* qdev = (kin > 0 && kout > 0) ? stage_matrix(q_host, kin, kout, ldq) : nullptr;
  vq(n, kin, kout, v, ldv, qdev, v, ldv, with_resid, sigma, beta, beta_col, resid, nrm2);
}

template <typename R>
void CudaVecOpsZ<R>::vq_out(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* m_host, int ldm, T* out,
                            int64_t ldo) {
  const T* qdev = stage_matrix(m_host, kin, kout, ldm);
  vq(n, kin, kout, v, ldv, qdev, out, ldo, false, T(0), T(0), -1, nullptr, nullptr);
}

template class CudaVecOpsZ<double>;
template class CudaVecOpsZ<float>;

}  // namespace ab200
