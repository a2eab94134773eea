// vecops_cuda.cu -- generic sm_100a kernels behind the VecOps boundary.
//
// These kernels accept any alignment / leading dimension and any ncv.  They are the correctness
// baseline and the fallback of the TMA-tiled fast path in vecops_tma.cu.  All reductions are
// deterministic: fixed grid for a given n, fixed intra-block tree, per-CTA partials combined by the
// last CTA in a fixed order (no floating-point atomics), so a solve is bit-reproducible run to run.
//
// Reference call sites replaced: K1/K2 start_step, K5/K6 dots, K7/K8 update,
// K9/K10 conditional re-orthogonalisation, K12-K16 vq, K17 larnv, K21 ger.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "gate.cuh"
#include "tma_common.cuh"
#include "vecops_cuda.cuh"

namespace ab200 {

LaunchStats& launch_stats() {
  static LaunchStats s;
  return s;
}

Profiler& profiler() {
  static Profiler p;
  return p;
}
cudaEvent_t Profiler::get_event() {
  if (!pool_.empty()) {
    cudaEvent_t e = pool_.back();
    pool_.pop_back();
    return e;
  }
  cudaEvent_t e;
  AB200_CUDA_CHECK(cudaEventCreate(&e));
  return e;
}
void Profiler::begin(cudaStream_t s, const char* name, double bytes) {
  int idx = -1;
  for (size_t i = 0; i < table_.size(); ++i)
    if (table_[i].name == name || std::strcmp(table_[i].name, name) == 0) { idx = (int)i; break; }
  if (idx < 0) {
    table_.push_back(KernelProfile{name, 0ULL, 0.0, 0.0});
    idx = (int)table_.size() - 1;
  }
  table_[idx].launches++;
  table_[idx].bytes += bytes;
  cur_idx_ = idx;
  cur0_ = get_event();
  AB200_CUDA_CHECK(cudaEventRecord(cur0_, s));
}
void Profiler::end(cudaStream_t s) {
  cudaEvent_t e1 = get_event();
  AB200_CUDA_CHECK(cudaEventRecord(e1, s));
  pending_.push_back(Pending{cur_idx_, cur0_, e1});
  if (pending_.size() >= 8192) flush();
}
void Profiler::flush() {
  // "(between launches)": device time from the end of one profiled launch to the start of the next one on the same
  // stream -- copies, memsets, un-profiled kernels and genuine idle time (the host not keeping the queue fed)
  int gap_idx = -1;
  if (pending_.size() > 1) {
    static const char* kGap = "(between launches)";
    for (size_t i = 0; i < table_.size(); ++i)
      if (table_[i].name == kGap) gap_idx = (int)i;
    if (gap_idx < 0) {
      table_.push_back(KernelProfile{kGap, 0ULL, 0.0, 0.0});
      gap_idx = (int)table_.size() - 1;
    }
  }
  cudaEvent_t prev_end = nullptr;
  for (auto& p : pending_) {
    float ms = 0.f;
    cudaEventSynchronize(p.e1);
    if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) table_[p.idx].ms += ms;
    if (prev_end != nullptr && gap_idx >= 0 && cudaEventElapsedTime(&ms, prev_end, p.e0) == cudaSuccess && ms < 50.f) {
      table_[gap_idx].ms += ms;   // (gaps above 50 ms are pauses between solves, not part of one)
      table_[gap_idx].launches++;
    }
    prev_end = p.e1;
  }
  for (auto& p : pending_) {
    pool_.push_back(p.e0);
    pool_.push_back(p.e1);
  }
  pending_.clear();
}
void Profiler::reset() {
  flush();
  table_.clear();
}

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ T ld_cg(const T* p) {
  return __ldcg(p);
}

// Combine the per-CTA partials: the CTA that takes the last ticket sums partial[b*pcols + c] over b
// in a fixed order (lane-strided, then an xor tree) and writes out[c].
template <typename T>
__device__ void finish_grid_reduce(T* partial, int pcols, int ncols, T* out, unsigned int* ticket) {
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
    T s = T(0);
    for (int b = lane; b < (int)gridDim.x; b += 32) s += ld_cg(partial + (size_t)b * pcols + c);
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// ---------------------------------------------------------------------------------------------
// K5/K6: out[0..j) = V_j^T x, out[j] = x.y.  Column chunks of CC accumulate in registers while the
// CTA streams its contiguous row range; x is re-read once per chunk (mostly from L2).
// ---------------------------------------------------------------------------------------------
template <typename T, int CC>
__global__ void __launch_bounds__(kThreads) k_dots(int64_t n, int j, const T* __restrict__ v, int64_t ldv,
                                                   const T* __restrict__ x, const T* __restrict__ y,
                                                   T* __restrict__ partial, int pcols, T* __restrict__ out,
                                                   unsigned int* ticket, const T* stop) {
  __shared__ T red[kWarps][CC + 1];
  if (stopped(stop)) return;
  const int64_t rpc = ((n + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
  const int64_t r0 = (int64_t)blockIdx.x * rpc;
  const int64_t r1 = (r0 + rpc < n) ? r0 + rpc : n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* mine = partial + (size_t)blockIdx.x * pcols;
  for (int c0 = 0; c0 < j || c0 == 0; c0 += CC) {
    T acc[CC];
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) acc[cc] = T(0);
    T accxy = T(0);
    const T* vc = v + (int64_t)c0 * ldv;
    const int ncol = (j - c0 < CC) ? (j - c0) : CC;
    for (int64_t r = r0 + threadIdx.x; r < r1; r += kThreads) {
      const T xv = x[r];
      if (c0 == 0) accxy += xv * y[r];
      if (ncol == CC) {
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) acc[cc] += vc[r + (int64_t)cc * ldv] * xv;
      } else {
#pragma unroll
        for (int cc = 0; cc < CC; ++cc)
          if (cc < ncol) acc[cc] += vc[r + (int64_t)cc * ldv] * xv;
      }
    }
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      const T s = warp_sum(acc[cc]);
      if (lane == 0) red[warp][cc] = s;
    }
    {
      const T s = warp_sum(accxy);
      if (lane == 0) red[warp][CC] = s;
    }
    __syncthreads();
    if (threadIdx.x <= CC) {
      T s = T(0);
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
      if (threadIdx.x < CC) {
        if (c0 + (int)threadIdx.x < j) mine[c0 + threadIdx.x] = s;
      } else if (c0 == 0) {
        mine[j] = s;
      }
    }
    __syncthreads();
    if (j == 0) break;
  }
  finish_grid_reduce(partial, pcols, j + 1, out, ticket);
}

// ---------------------------------------------------------------------------------------------
// K7/K8 (+K9/K10 when predicated): dst = src - V_j*coef ; nrm2 = sum dst^2.
// With pred_* set the kernel first evaluates the reference's DGKS test on the device
// (dsaitr.f:656: rnorm > 0.717*wnorm -> nothing to do) and exits early when no pass is needed.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_update(int64_t n, int j, const T* __restrict__ v, int64_t ldv,
                                                     const T* __restrict__ coef, const T* src, T* dst,
                                                     T* __restrict__ partial, T* __restrict__ nrm2_out,
                                                     unsigned int* ticket, const T* pred_w2,
                                                     const T* pred_r2, T* flag_out, const T* stop) {
  extern __shared__ unsigned char smem_raw[];
  T* cs = reinterpret_cast<T*>(smem_raw);
  __shared__ T red[kWarps];
  if (stopped(stop)) return;
  if (pred_w2 != nullptr) {
    const T wn = sqrt(*pred_w2), rn = sqrt(*pred_r2);
    const bool needed = !(rn > T(0.717f) * wn);
    if (!needed) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && flag_out) *flag_out = T(0);
      return;
    }
  }
  for (int k = threadIdx.x; k < j; k += kThreads) cs[k] = coef[k];
  __syncthreads();
  T nrm = T(0);
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    const T* vr = v + r;
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
    int k = 0;
    for (; k + 4 <= j; k += 4) {
      a0 += vr[(int64_t)(k + 0) * ldv] * cs[k + 0];
      a1 += vr[(int64_t)(k + 1) * ldv] * cs[k + 1];
      a2 += vr[(int64_t)(k + 2) * ldv] * cs[k + 2];
      a3 += vr[(int64_t)(k + 3) * ldv] * cs[k + 3];
    }
    for (; k < j; ++k) a0 += vr[(int64_t)k * ldv] * cs[k];
    const T d = src[r] - ((a0 + a1) + (a2 + a3));
    dst[r] = d;
    nrm += d * d;
  }
  if (nrm2_out == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  nrm = warp_sum(nrm);
  if (lane == 0) red[warp] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w];
    partial[blockIdx.x] = s;
    if (blockIdx.x == 0 && flag_out) *flag_out = T(1);
  }
  finish_grid_reduce(partial, 1, 1, nrm2_out, ticket);
}

// two-vector dot
template <typename T>
__global__ void __launch_bounds__(kThreads) k_dot(int64_t n, const T* __restrict__ x, const T* __restrict__ y,
                                                  T* __restrict__ partial, T* __restrict__ out,
                                                  unsigned int* ticket) {
  __shared__ T red[kWarps];
  const int64_t rpc = ((n + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
  const int64_t r0 = (int64_t)blockIdx.x * rpc;
  const int64_t r1 = (r0 + rpc < n) ? r0 + rpc : n;
  T acc = T(0);
  for (int64_t r = r0 + threadIdx.x; r < r1; r += kThreads) acc += x[r] * y[r];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
  finish_grid_reduce(partial, 1, 1, out, ticket);
}

// max |x_i|: per-CTA maxima, combined by the CTA that takes the last ticket (max is order-independent)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_absmax(int64_t n, const T* __restrict__ x, T* __restrict__ partial,
                                                     T* __restrict__ out, unsigned int* ticket) {
  __shared__ T red[kWarps];
  __shared__ bool s_last;
  T m = T(0);
  for (int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x; r < n; r += (int64_t)gridDim.x * kThreads) {
    const T a = fabs(x[r]);
    m = (a > m || a != a) ? a : m;   // a NaN entry propagates
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T other = __shfl_xor_sync(0xffffffffu, m, o);
    m = (other > m || other != other) ? other : m;
  }
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kWarps; ++w) m = (red[w] > m || red[w] != red[w]) ? red[w] : m;
    partial[blockIdx.x] = m;
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    if (s_last) {
      __threadfence();
      T g = T(0);
      for (unsigned int b = 0; b < gridDim.x; ++b) {
        const T v = ld_cg(partial + b);
        g = (v > g || v != v) ? v : g;
      }
      *out = g;
      *ticket = 0u;
    }
  }
}

// y = a*y + b*x (+ ||y||^2)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_axpby_norm(int64_t n, T a, T b, const T* __restrict__ x, T* y,
                                                         T* __restrict__ partial, T* __restrict__ out,
                                                         unsigned int* ticket) {
  __shared__ T red[kWarps];
  const int64_t rpc = ((n + gridDim.x - 1) / gridDim.x + kThreads - 1) / kThreads * kThreads;
  const int64_t r0 = (int64_t)blockIdx.x * rpc;
  const int64_t r1 = (r0 + rpc < n) ? r0 + rpc : n;
  T acc = T(0);
  for (int64_t r = r0 + threadIdx.x; r < r1; r += kThreads) {
    T t = a * y[r];
    if (x != nullptr) t += b * x[r];
    y[r] = t;
    acc += t * t;
  }
  if (out == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
  finish_grid_reduce(partial, 1, 1, out, ticket);
}

template <typename T>
__global__ void k_scal(int64_t n, T alpha, T* x) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    x[r] *= alpha;
}

// K1+K2: v_j = resid*inv, x = v_j, Bx scaled (dsaitr.f:438-442,464)
template <typename T>
__global__ void k_start_step(int64_t n, T inv, const T* __restrict__ resid, T* __restrict__ vj,
                             T* __restrict__ outx, T* bx, bool bx_from_resid, const T* stop) {
  tma::pdl_trigger();
  tma::pdl_wait();
  if (stopped(stop)) return;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const T t = resid[r] * inv;
    vj[r] = t;
    outx[r] = t;
    if (bx != nullptr) bx[r] = bx_from_resid ? t : bx[r] * inv;
  }
}

// gated K1+K2 of a device-resident sweep: scale and rare-path tests from the previous step's mailbox slot
template <typename T>
__global__ void k_start_step_gated(int64_t n, const StepGate<T> g, const T* __restrict__ resid, T* __restrict__ vj,
                                   T* __restrict__ outx, T* bx) {
  tma::pdl_trigger();
  tma::pdl_wait();
  if (stopped(g.stop)) return;
  T inv;
  if (!gate_eval_block(g, inv)) {
    // every block takes the same decision from the same values; one thread records it.  No block of THIS kernel can
    // observe the store before its own entry check only if none is scheduled later -- which is why the check above
    // treats "already stopped" and "stops now" alike: nothing is written either way.
    if (blockIdx.x == 0 && threadIdx.x == 0) *g.stop = g.stop_code;
    return;
  }
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const T t = resid[r] * inv;
    vj[r] = t;
    outx[r] = t;
    if (bx != nullptr) bx[r] = t;
  }
}

// K21: Z(:,0:k) += resid * w^T
template <typename T>
__global__ void k_ger(int64_t n, int k, const T* __restrict__ resid, const T* __restrict__ w, T* z, int64_t ldz) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    const T rv = resid[r];
    for (int c = 0; c < k; ++c) z[r + (int64_t)c * ldz] += rv * w[c];
  }
}

template <typename T>
__global__ void k_copy2d(int64_t n, int cols, const T* __restrict__ src, int64_t lds, T* __restrict__ dst,
                         int64_t ldd) {
  for (int c = blockIdx.y; c < cols; c += gridDim.y)
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
      dst[r + (int64_t)c * ldd] = src[r + (int64_t)c * lds];
}

// ---------------------------------------------------------------------------------------------
// K17: LAPACK xLARNV(idist=2).  xLARUV is the multiplicative congruential generator
// s_{i} = a * s_{i-1} mod 2^48 with a = 33952834046453 (the table MM(i,:) holds a^i); element i
// (1-based) of the stream is therefore seed * a^i mod 2^48 and the stream can be generated in
// parallel with bit-identical results.  u = s / 2^48, x = 2u - 1.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kLaruvA = 33952834046453ULL;
constexpr unsigned long long kMask48 = (1ULL << 48) - 1ULL;

__host__ __device__ inline unsigned long long mulmod48(unsigned long long a, unsigned long long b) {
  return (a * b) & kMask48;
}
__host__ __device__ inline unsigned long long powmod48(unsigned long long base, unsigned long long e) {
  unsigned long long r = 1ULL;
  while (e) {
    if (e & 1ULL) r = mulmod48(r, base);
    base = mulmod48(base, base);
    e >>= 1;
  }
  return r;
}

template <typename T>
__device__ inline T laruv_to_unit(unsigned long long s);
template <>
__device__ inline double laruv_to_unit<double>(unsigned long long s) {
  // exact: 48 bits fit a double mantissa (dlaruv.f evaluates r*(it1 + r*(it2 + r*(it3 + r*it4))))
  return (double)s * (1.0 / 281474976710656.0);
}
template <>
__device__ inline float laruv_to_unit<float>(unsigned long long s) {
  // slaruv.f rounds at every level of the nested evaluation
  const float r = 1.0f / 4096.0f;
  const float it1 = (float)((s >> 36) & 4095ULL), it2 = (float)((s >> 24) & 4095ULL);
  const float it3 = (float)((s >> 12) & 4095ULL), it4 = (float)(s & 4095ULL);
  return __fmul_rn(r, __fadd_rn(it1, __fmul_rn(r, __fadd_rn(it2, __fmul_rn(r, __fadd_rn(it3, __fmul_rn(r, it4)))))));
}

template <typename T, int PER_THREAD>
__global__ void k_larnv(int64_t n, unsigned long long seed, T* __restrict__ x, unsigned int* hit_one) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // thread t produces elements i = t + m*stride (0-based); s_i = seed * a^(i+1)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (t >= n) return;
  unsigned long long s = mulmod48(seed, powmod48(kLaruvA, (unsigned long long)(t + 1)));
  const unsigned long long astride = powmod48(kLaruvA, (unsigned long long)stride);
  for (int64_t i = t; i < n; i += stride) {
    const T u = laruv_to_unit<T>(s);
    if (u == T(1)) atomicAdd(hit_one, 1u);  // slaruv would re-draw (never happens in double)
    x[i] = T(2) * u - T(1);
    s = mulmod48(s, astride);
  }
}

// ---------------------------------------------------------------------------------------------
// K12-K16: out(:,0:kout) = V(:,0:kin) * Q, safe when out aliases V: each CTA owns 32-row tiles,
// stages the tile in shared memory, synchronises, then writes.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_vq(int64_t n, int kin, int kout, const T* v, int64_t ldv,
                                                 const T* __restrict__ q, T* out, int64_t ldo, bool with_resid,
                                                 T sigma, T beta, int beta_col, T* resid,
                                                 T* __restrict__ partial, T* __restrict__ nrm2_out,
                                                 unsigned int* ticket) {
  extern __shared__ unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);  // [kin][33]
  __shared__ T red[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (n + 31) / 32;
  T nrm = T(0);
  for (int64_t tix = blockIdx.x; tix < ntiles; tix += gridDim.x) {
    const int64_t row = tix * 32 + lane;
    const bool ok = row < n;
    for (int k = warp; k < kin; k += kWarps) tile[k * 33 + lane] = ok ? v[row + (int64_t)k * ldv] : T(0);
    __syncthreads();
    for (int c = warp; c < kout; c += kWarps) {
      const T* qc = q + (size_t)c * kin;
      T a0 = T(0), a1 = T(0);
      int k = 0;
      for (; k + 2 <= kin; k += 2) {
        a0 += tile[k * 33 + lane] * __ldg(qc + k);
        a1 += tile[(k + 1) * 33 + lane] * __ldg(qc + k + 1);
      }
      if (k < kin) a0 += tile[k * 33 + lane] * __ldg(qc + k);
      const T o = a0 + a1;
      if (ok) out[row + (int64_t)c * ldo] = o;
      if (with_resid && c == beta_col && ok) {
        const T t = sigma * resid[row] + beta * o;
        resid[row] = t;
        nrm += t * t;
      }
    }
    if (with_resid && beta_col < 0 && warp == 0 && ok) {
      const T t = sigma * resid[row];
      resid[row] = t;
      nrm += t * t;
    }
    __syncthreads();
  }
  if (!with_resid || nrm2_out == nullptr) return;
  nrm = warp_sum(nrm);
  if (lane == 0) red[warp] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    T s = T(0);
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
  finish_grid_reduce(partial, 1, 1, nrm2_out, ticket);
}

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
template <typename T>
CudaVecOps<T>::CudaVecOps(cudaStream_t stream, NcclComm* comm) : stream_(stream), comm_(comm) {
  int dev = 0;
  AB200_CUDA_CHECK(cudaGetDevice(&dev));
  AB200_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev));
  AB200_CUDA_CHECK(cudaMalloc(&ticket_, sizeof(unsigned int) * 4));
  AB200_CUDA_CHECK(cudaMemset(ticket_, 0, sizeof(unsigned int) * 4));
  const char* km = getenv("AB200_KERNELS");
  if (km && std::strcmp(km, "generic") == 0) kernel_mode_ = 1;
}

template <typename T>
CudaVecOps<T>::~CudaVecOps() {
  cudaFree(mb_dev_);
  cudaFreeHost(mb_pinned_);
  cudaFree(partial_);
  cudaFree(ticket_);
  cudaFree(qbuf_);
  cudaFreeHost(qpinned_);
}

template <typename T>
T* CudaVecOps<T>::alloc(size_t count) {
  T* p = nullptr;
  AB200_CUDA_CHECK(cudaMalloc(&p, sizeof(T) * (count ? count : 1)));
  return p;
}
template <typename T>
void CudaVecOps<T>::release(T* p) {
  if (p) cudaFree(p);
}
template <typename T>
void CudaVecOps<T>::upload(T* d, const T* h, size_t count) {
  AB200_CUDA_CHECK(cudaMemcpyAsync(d, h, sizeof(T) * count, cudaMemcpyHostToDevice, stream_));
}
template <typename T>
void CudaVecOps<T>::download(T* h, const T* d, size_t count) {
  AB200_CUDA_CHECK(cudaMemcpyAsync(h, d, sizeof(T) * count, cudaMemcpyDeviceToHost, stream_));
}
template <typename T>
void CudaVecOps<T>::upload2d(T* d, size_t ldd, const T* h, size_t lds, size_t rows, size_t cols) {
  AB200_CUDA_CHECK(cudaMemcpy2DAsync(d, ldd * sizeof(T), h, lds * sizeof(T), rows * sizeof(T), cols,
                                     cudaMemcpyHostToDevice, stream_));
}
template <typename T>
void CudaVecOps<T>::download2d(T* h, size_t ldd, const T* d, size_t lds, size_t rows, size_t cols) {
  AB200_CUDA_CHECK(cudaMemcpy2DAsync(h, ldd * sizeof(T), d, lds * sizeof(T), rows * sizeof(T), cols,
                                     cudaMemcpyDeviceToHost, stream_));
}
template <typename T>
void CudaVecOps<T>::sync() {
  resolve_pending();
  AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
}
template <typename T>
bool CudaVecOps<T>::is_device_pointer(const void* p) {
  cudaPointerAttributes a;
  const cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

template <typename T>
T* CudaVecOps<T>::mailbox(size_t count) {
  if (count > mb_count_) {
    cudaFree(mb_dev_);
    cudaFreeHost(mb_pinned_);
    AB200_CUDA_CHECK(cudaMalloc(&mb_dev_, sizeof(T) * count));
    AB200_CUDA_CHECK(cudaMallocHost(&mb_pinned_, sizeof(T) * count));
    mb_count_ = count;
  }
  AB200_CUDA_CHECK(cudaMemsetAsync(mb_dev_, 0, sizeof(T) * mb_count_, stream_));
  // size the reduction scratch once for the whole solve (largest grid x one mailbox segment), so that
  // no step ever pays a synchronising cudaFree/cudaMalloc
  ensure_partial((size_t)num_sms_ * 8 * 72);
  return mb_dev_;
}
template <typename T>
bool CudaVecOps<T>::ranks_agree_on_fusing(int64_t n, int j, const T* v, int64_t ldv) {
  if (comm_ == nullptr) return false;
  if (agreed_v_ == (const void*)v && agreed_ldv_ == ldv) return agreed_fuse_;
  // one tiny all-reduce + host read per solve: does EVERY rank take the TMA-tiled path for this V?
  const T mine = (kernel_mode_ == 0 && fast_path_ok(n, j, v, ldv)) ? T(1) : T(0);
  T* slot = partial_;  // reduction scratch: free between kernels
  ensure_partial(8);
  slot = partial_;
  AB200_CUDA_CHECK(cudaMemcpyAsync(slot, &mine, sizeof(T), cudaMemcpyHostToDevice, stream_));
  AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
  nccl_allreduce_sum(comm_, slot, 1, sizeof(T) == 8, stream_);
  T sum = T(0);
  AB200_CUDA_CHECK(cudaMemcpyAsync(&sum, slot, sizeof(T), cudaMemcpyDeviceToHost, stream_));
  AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
  agreed_v_ = v;
  agreed_ldv_ = ldv;
  agreed_fuse_ = (sum == T(nccl_nranks(comm_)));
  return agreed_fuse_;
}
template <typename T>
void CudaVecOps<T>::resolve_pending() {
  if (!has_pending_) return;
  has_pending_ = false;
  nccl_peer_reduce_finalize(comm_, pending_, 1, pending_log_, pending_w2_, pending_r2_, pending_stop_, sizeof(T) == 8,
                            stream_);
}
template <typename T>
void CudaVecOps<T>::fetch(T* host_dst, const T* mb, size_t count) {
  resolve_pending();
  const size_t off = (size_t)(mb - mb_dev_);
  AB200_CUDA_CHECK(cudaMemcpyAsync(mb_pinned_ + off, mb, sizeof(T) * count, cudaMemcpyDeviceToHost, stream_));
  AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
  launch_stats().fetches++;
  std::memcpy(host_dst, mb_pinned_ + off, sizeof(T) * count);
}
template <typename T>
void CudaVecOps<T>::post(T* mb, const T* host_src, size_t count) {
  resolve_pending();
  const size_t off = (size_t)(mb - mb_dev_);
  AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
  std::memcpy(mb_pinned_ + off, host_src, sizeof(T) * count);
  AB200_CUDA_CHECK(cudaMemcpyAsync(mb, mb_pinned_ + off, sizeof(T) * count, cudaMemcpyHostToDevice, stream_));
}
template <typename T>
void CudaVecOps<T>::allreduce_sum(T* mb, size_t count) {
  if (comm_ == nullptr || count == 0) return;
  resolve_pending();
  nccl_allreduce_sum(comm_, mb, count, sizeof(T) == 8, stream_);
  launch_stats().allreduces++;
}
template <typename T>
int CudaVecOps<T>::rank() const {
  return comm_ ? nccl_rank(comm_) : 0;
}
template <typename T>
int CudaVecOps<T>::nranks() const {
  return comm_ ? nccl_nranks(comm_) : 1;
}

template <typename T>
void CudaVecOps<T>::ensure_partial(size_t count) {
  if (count > partial_count_) {
    AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
    cudaFree(partial_);
    AB200_CUDA_CHECK(cudaMalloc(&partial_, sizeof(T) * count));
    partial_count_ = count;
  }
}
template <typename T>
int CudaVecOps<T>::reduce_grid(int64_t n) const {
  // a pure function of n (and the SM count): keeps reductions bit-reproducible
  const int64_t want = (n + 4 * kThreads - 1) / (4 * kThreads);
  const int64_t cap = (int64_t)num_sms_ * 4;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

template <typename T>
T* CudaVecOps<T>::stage_matrix(const T* host, int rows, int cols, int ld) {
  const size_t cnt = (size_t)rows * cols;
  if (cnt > qbuf_count_) {
    AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
    cudaFree(qbuf_);
    AB200_CUDA_CHECK(cudaMalloc(&qbuf_, sizeof(T) * cnt));
    qbuf_count_ = cnt;
  }
  if (cnt > qpinned_count_) {
    AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
    cudaFreeHost(qpinned_);
    AB200_CUDA_CHECK(cudaMallocHost(&qpinned_, sizeof(T) * cnt));
    qpinned_count_ = cnt;
  }
  // the pinned staging buffer is reused: make sure the previous upload has drained
  AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
  for (int c = 0; c < cols; ++c) std::memcpy(qpinned_ + (size_t)c * rows, host + (size_t)c * ld, sizeof(T) * rows);
  AB200_CUDA_CHECK(cudaMemcpyAsync(qbuf_, qpinned_, sizeof(T) * cnt, cudaMemcpyHostToDevice, stream_));
  return qbuf_;
}

#define AB200_LAUNCHED()                        \
  do {                                          \
    launch_stats().kernels++;                   \
    AB200_CUDA_CHECK(cudaGetLastError());       \
  } while (0)

template <typename T>
void CudaVecOps<T>::copy(int64_t n, const T* x, T* y) {
  if (x == y || n <= 0) return;
  AB200_CUDA_CHECK(cudaMemcpyAsync(y, x, sizeof(T) * (size_t)n, cudaMemcpyDeviceToDevice, stream_));
}
template <typename T>
void CudaVecOps<T>::zero(int64_t n, T* x) {
  AB200_CUDA_CHECK(cudaMemsetAsync(x, 0, sizeof(T) * (size_t)n, stream_));
}
template <typename T>
void CudaVecOps<T>::scal(int64_t n, T alpha, T* x) {
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  k_scal<T><<<grid, 256, 0, stream_>>>(n, alpha, x);
  AB200_LAUNCHED();
}
template <typename T>
void CudaVecOps<T>::axpby_norm(int64_t n, T a, T b, const T* x, T* y, T* out) {
  const int grid = reduce_grid(n);
  ensure_partial((size_t)grid);
  ProfScope ps(stream_, "axpby_norm", (double)sizeof(T) * n * (x ? 3.0 : 2.0));
  k_axpby_norm<T><<<grid, kThreads, 0, stream_>>>(n, a, b, x, y, partial_, out, ticket_);
  AB200_LAUNCHED();
}
template <typename T>
void CudaVecOps<T>::dot(int64_t n, const T* x, const T* y, T* out) {
  const int grid = reduce_grid(n);
  ensure_partial((size_t)grid);
  ProfScope ps(stream_, "dot", (double)sizeof(T) * n * (x == y ? 1.0 : 2.0));
  k_dot<T><<<grid, kThreads, 0, stream_>>>(n, x, y, partial_, out, ticket_);
  AB200_LAUNCHED();
}

template <typename T>
bool CudaVecOps<T>::absmax(int64_t n, const T* x, T* out) {
  const int grid = reduce_grid(n);
  ensure_partial((size_t)grid);
  k_absmax<T><<<grid, kThreads, 0, stream_>>>(n, x, partial_, out, ticket_);
  AB200_LAUNCHED();
  return true;
}

template <typename T>
void CudaVecOps<T>::larnv_uniform_m1_1(int64_t n, int iseed[4], T* x) {
  const unsigned long long seed = ((unsigned long long)iseed[0] << 36) | ((unsigned long long)iseed[1] << 24) |
                                  ((unsigned long long)iseed[2] << 12) | (unsigned long long)iseed[3];
  unsigned int* hit = ticket_ + 1;
  AB200_CUDA_CHECK(cudaMemsetAsync(hit, 0, sizeof(unsigned int), stream_));
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  k_larnv<T, 1><<<grid, 256, 0, stream_>>>(n, seed, x, hit);
  AB200_LAUNCHED();
  if (sizeof(T) == 4) {
    unsigned int h = 0;
    AB200_CUDA_CHECK(cudaMemcpyAsync(&h, hit, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream_));
    AB200_CUDA_CHECK(cudaStreamSynchronize(stream_));
    if (h != 0)
      throw CudaError("slarnv: a draw rounded to 1.0f (slaruv re-draw path, probability 2^-24 per element) -- "
                      "pass a start vector with info=1");
  }
  const unsigned long long s = mulmod48(seed, powmod48(kLaruvA, (unsigned long long)n));
  iseed[0] = (int)((s >> 36) & 4095ULL);
  iseed[1] = (int)((s >> 24) & 4095ULL);
  iseed[2] = (int)((s >> 12) & 4095ULL);
  iseed[3] = (int)(s & 4095ULL);
}

template <typename T>
void CudaVecOps<T>::start_step(int64_t n, T inv, const T* resid, T* vj, T* outx, T* bx, bool from_resid) {
  resolve_pending();
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  ProfScope ps(stream_, "start_step", (double)sizeof(T) * n * (bx ? (from_resid ? 4.0 : 5.0) : 3.0));
  AB200_CUDA_CHECK(tma::launch_pdl(k_start_step<T>, grid, 256, 0, stream_, n, inv, resid, vj, outx, bx, from_resid,
                                   (const T*)stop_));
  launch_stats().kernels++;
}
template <typename T>
void CudaVecOps<T>::start_step_gated(int64_t n, const StepGate<T>& g0, const T* resid, T* vj, T* outx, T* bx) {
  StepGate<T> g = g0;
  attach_pending(g);  // multi-GPU: the gate finishes the reduction of ||r'||^2 itself
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  ProfScope ps(stream_, "start_step", (double)sizeof(T) * n * (bx ? 4.0 : 3.0));
  AB200_CUDA_CHECK(tma::launch_pdl(k_start_step_gated<T>, grid, 256, 0, stream_, n, g, resid, vj, outx, bx));
  launch_stats().kernels++;
}
template <typename T>
void CudaVecOps<T>::ger(int64_t n, int k, const T* resid, const T* w_host, T* z, int64_t ldz) {
  T* wdev = stage_matrix(w_host, k, 1, k);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  k_ger<T><<<grid, 256, 0, stream_>>>(n, k, resid, wdev, z, ldz);
  AB200_LAUNCHED();
}
template <typename T>
void CudaVecOps<T>::copy2d(int64_t n, int cols, const T* src, int64_t lds, T* dst, int64_t ldd) {
  if (cols <= 0) return;
  dim3 grid((unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 2), (unsigned)std::min(cols, 64));
  k_copy2d<T><<<grid, 256, 0, stream_>>>(n, cols, src, lds, dst, ldd);
  AB200_LAUNCHED();
}

template <typename T>
void CudaVecOps<T>::dots_generic(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* out) {
  constexpr int CC = 8;
  const int grid = reduce_grid(n);
  const int pcols = j + 1;
  ensure_partial((size_t)grid * pcols);
  ProfScope ps(stream_, "dots_generic", (double)sizeof(T) * n * (j + (x == y ? 1.0 : 2.0)));
  k_dots<T, CC><<<grid, kThreads, 0, stream_>>>(n, j, v, ldv, x, y, partial_, pcols, out, ticket_, stop_);
  AB200_LAUNCHED();
  launch_stats().fallback++;
}
template <typename T>
void CudaVecOps<T>::update_generic(int64_t n, int j, const T* v, int64_t ldv, const T* coef, const T* src,
                                   T* dst, T* nrm2, const T* pw2, const T* pr2, T* flag) {
  const int grid = (int)std::min<int64_t>((n + kThreads - 1) / kThreads, (int64_t)num_sms_ * 8);
  ensure_partial((size_t)grid);
  ProfScope ps(stream_, pw2 ? "reorth_generic" : "update_generic", (double)sizeof(T) * n * (j + 2.0));
  k_update<T><<<grid, kThreads, sizeof(T) * (size_t)std::max(j, 1), stream_>>>(n, j, v, ldv, coef, src, dst,
                                                                               partial_, nrm2, ticket_, pw2, pr2,
                                                                               flag, stop_);
  AB200_LAUNCHED();
  launch_stats().fallback++;
}

template <typename T>
void CudaVecOps<T>::dots(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* out) {
  resolve_pending();
  if (kernel_mode_ == 0 && fast_path_ok(n, j, v, ldv) && dots_tma(n, j, v, ldv, x, y, out)) return;
  dots_generic(n, j, v, ldv, x, y, out);
}
template <typename T>
void CudaVecOps<T>::update(int64_t n, int j, const T* v, int64_t ldv, const T* coef, const T* src, T* dst,
                           T* nrm2) {
  resolve_pending();
  update_generic(n, j, v, ldv, coef, src, dst, nrm2, nullptr, nullptr, nullptr);
}

template <typename T>
void CudaVecOps<T>::orth_step(int64_t n, int j, const T* v, int64_t ldv, const T* w, T* resid, T* mbA, T* mbB,
                              T* mbC) {
  resolve_pending();
  if (comm_ != nullptr) ranks_agree_on_fusing(n, j, v, ldv);  // collective, once per solve
  if (kernel_mode_ == 0 && fast_path_ok(n, j, v, ldv) && orth_step_tma(n, j, v, ldv, w, resid, mbA, mbB, mbC))
    return;
  // generic composition: 4 sweeps over V_j (the reference's own dependency order, K6 K7 K9 K9)
  dots_generic(n, j, v, ldv, w, w, mbA);
  allreduce_sum(mbA, (size_t)j + 1);
  update_generic(n, j, v, ldv, mbA, w, resid, nullptr, nullptr, nullptr, nullptr);
  dots_generic(n, j, v, ldv, resid, resid, mbB);
  allreduce_sum(mbB, (size_t)j + 1);
  update_generic(n, j, v, ldv, mbB, resid, resid, mbC, mbA + j, mbB + j, mbC + 1);
  allreduce_sum(mbC, 1);
}

template <typename T>
void CudaVecOps<T>::vq_smem_attr(int kin) {
  const size_t bytes = sizeof(T) * 33 * (size_t)kin;
  if (bytes > 48 * 1024)
    AB200_CUDA_CHECK(cudaFuncSetAttribute(k_vq<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

template <typename T>
void CudaVecOps<T>::vq_update(int64_t n, int kin, int kout, T* v, int64_t ldv, const T* q_host, int ldq,
                              bool with_resid, T sigma, T beta, int beta_col, T* resid, T* mb_nrm2) {
  if (kout <= 0) {
    if (with_resid) axpby_norm(n, sigma, T(0), nullptr, resid, mb_nrm2);
    return;
  }
  resolve_pending();
  T* qdev = stage_matrix(q_host, kin, kout, ldq);
  if (kernel_mode_ == 0 && vq_mma(n, kin, kout, v, ldv, qdev, q_host, ldq, v, ldv, with_resid, sigma, beta, beta_col,
                                  resid, mb_nrm2))
    return;
  if (kernel_mode_ == 0 && vq_tma(n, kin, kout, v, ldv, qdev, v, ldv, with_resid, sigma, beta, beta_col, resid,
                                  mb_nrm2))
    return;
  const int grid = (int)std::min<int64_t>((n + 31) / 32, (int64_t)num_sms_ * 6);
  ensure_partial((size_t)grid);
  ProfScope ps(stream_, "vq_generic", (double)sizeof(T) * n * (kin + kout + (with_resid ? 2.0 : 0.0)));
  vq_smem_attr(kin);
  k_vq<T><<<grid, kThreads, sizeof(T) * 33 * (size_t)kin, stream_>>>(n, kin, kout, v, ldv, qdev, v, ldv,
                                                                      with_resid, sigma, beta, beta_col, resid,
                                                                      partial_, mb_nrm2, ticket_);
  AB200_LAUNCHED();
  launch_stats().fallback++;
}
template <typename T>
void CudaVecOps<T>::vq_out(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* m_host, int ldm,
                           T* out, int64_t ldo) {
  if (kout <= 0) return;
  resolve_pending();
  T* qdev = stage_matrix(m_host, kin, kout, ldm);
  if (kernel_mode_ == 0 &&
      vq_mma(n, kin, kout, v, ldv, qdev, m_host, ldm, out, ldo, false, T(0), T(0), -1, nullptr, nullptr))
    return;
  if (kernel_mode_ == 0 &&
      vq_tma(n, kin, kout, v, ldv, qdev, out, ldo, false, T(0), T(0), -1, nullptr, nullptr))
    return;
  const int grid = (int)std::min<int64_t>((n + 31) / 32, (int64_t)num_sms_ * 6);
  ensure_partial((size_t)grid);
  ProfScope ps(stream_, "vq_generic", (double)sizeof(T) * n * (kin + kout));
  vq_smem_attr(kin);
  k_vq<T><<<grid, kThreads, sizeof(T) * 33 * (size_t)kin, stream_>>>(n, kin, kout, v, ldv, qdev, out, ldo, false,
                                                                      T(0), T(0), -1, nullptr, partial_, nullptr,
                                                                      ticket_);
  AB200_LAUNCHED();
  launch_stats().fallback++;
}

template class CudaVecOps<double>;
template class CudaVecOps<float>;

}  // namespace ab200
