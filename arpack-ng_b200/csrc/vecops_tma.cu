// vecops_tma.cu -- TMA-tiled fast path of the tall-skinny kernels (sm_100a).
//
// One persistent CTA per SM streams V(n x j) through shared memory in tiles of R = 128 rows x all
// j columns.  A producer warp issues one 2-D TMA box copy (cp.async.bulk.tensor) per 8 columns, plus a
// 1-D box for the vector operand, into a ring of mbarrier-guarded stages; two groups of eight consumer
// warps take alternate tiles (so that one group's latency-bound phases overlap the other's):
//
//   DOTS      out[c] = sum_i V[i,c]*x[i]                      (K6, + x.y for K5)
//   UPD_SPEC  r = w - V*h ; ||r||^2 ; s[c] = sum_i V[i,c]*r[i]  (K7+K8 and, speculatively, the K9 dots:
//             the tile is still in shared memory when r is known, so DGKS costs no extra pass over V)
//   UPD       r -= V*s ; ||r||^2, predicated on the reference's test rnorm <= 0.717*wnorm (K9+K10)
//             (both in two forms: k_upd, rows per lane, for j <= 32 and all UPD launches; k_orth, a column per
//             warp, for UPD_SPEC beyond 32 columns -- see launch_upd)
//   VQ        out(:,0:kout) = V(:,0:kin)*Q, optionally r = sigma*r + beta*out(:,c), ||r||^2 (K12-K16, K20)
//
// so a Lanczos/Arnoldi step reads V_j three times instead of the reference's four and
// a restart reads V once.  Rows beyond n and columns beyond j are zero-filled by the TMA unit
// (out-of-bounds boxes), the neutral element of every sum here.  Reductions are deterministic (fixed
// grid, fixed trees, last-CTA combine).
#include <cstdlib>
#include <cstring>

#include "gate.cuh"
#include "tma_common.cuh"
#include "vecops_cuda.cuh"

namespace ab200 {

namespace {

using namespace ab200::tma;

constexpr int R = 128;            // rows per tile == TMA box height
constexpr int CB = 8;             // columns per TMA box == warps per consumer group
constexpr int NCW = 8;            // warps per consumer group
constexpr int NG = 2;             // consumer groups (alternate tiles)
constexpr int MAXB = 8;           // column boxes per tile -> j <= 64
constexpr int MAXST = 12;         // ring depth limit
constexpr int kThreadsTma = (NG * NCW + 1) * 32;
constexpr uint32_t kTileBudget = 192 * 1024;
constexpr uint32_t kTileBudgetVq = 188 * 1024;

enum Mode { DOTS = 0, UPD = 1, UPD_SPEC = 2 };

template <typename T>
struct OrthParams {
  int64_t n;
  int j, nboxes, nstages;
  uint32_t stage_bytes, box_bytes;
  int x_tma;      // vector operand arrives through the TMA ring (needs a 16-byte aligned base)
  const T* x;     // DOTS: x            UPD*: src
  const T* y;     // DOTS: y (may alias x)
  T* dst;         // UPD*: destination (may alias src)
  const T* coef;  // UPD*: j coefficients (device mailbox)
  T* partial;
  int pcols;
  T* out;         // DOTS/UPD_SPEC: out[0..j], UPD: out[0] (may be null)
  unsigned int* ticket;
  const T* pred_w2;
  const T* pred_r2;
  T* flag_out;
  const T* stop;  // sticky stop flag of a device-resident sweep (may be null)
  // multi-GPU: reductions fused into the kernels (peer.cuh).  pub: this kernel's sums go to the peers' slots instead
  // of `out`; sub: the coefficients (and the value behind them) come from the peers' slots instead of `coef` /
  // `pred_r2`, and CTA 0 leaves the reduced values in sub_log for the host.  nranks == 0: not in use.
  PeerReduce pub, sub;
  T* sub_log;
};

// Consumer side of a fused reduction: wait for every rank, add the slots in rank order into cs[0..j) (zero-padded up
// to MAXB*CB) and return value j; CTA 0 leaves all j+1 sums in p.sub_log.  Called by ALL threads of the CTA, before any
// early exit that depends on the values.
template <typename T>
__device__ __forceinline__ T fused_coefs(const OrthParams<T>& p, T* cs) {
  __shared__ T s_tail;
  peer_wait_all(p.sub);
  for (int k = threadIdx.x; k < MAXB * CB + 1; k += blockDim.x) {
    if (k <= p.j) {
      const T v = peer_sum<T>(p.sub, k);
      if (k < p.j) cs[k] = v;
      else s_tail = v;
      if (blockIdx.x == 0 && p.sub_log != nullptr) p.sub_log[k] = v;
    } else {
      cs[k - 1] = T(0);
    }
  }
  __syncthreads();
  return s_tail;
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreadsTma, 1)
k_orth(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap xmap, const OrthParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_trigger();
  T* tiles = reinterpret_cast<T*>(smem);
  unsigned char* aux = smem + (size_t)p.nstages * p.stage_bytes;
  T* ps = reinterpret_cast<T*>(aux);                     // [NG groups][2 buffers][2 halves][R]
  T* cs = ps + NG * 4 * R;                               // [MAXB*CB] coefficients
  T* gsum = cs + MAXB * CB;                              // [MAXB*CB + 8] partials of consumer group 1
  {
    // prologue that touches no global memory: may overlap the tail of the previous kernel (launch_pdl)
    uint64_t* full0 = reinterpret_cast<uint64_t*>(gsum + MAXB * CB + 8);
    if (threadIdx.x == 0) {
      for (int s = 0; s < p.nstages; ++s) {
        mbar_init(full0 + s, 1);
        mbar_init(full0 + MAXST + s, NCW);
      }
      mbar_fence_init();
    }
  }
  pdl_wait();
  if (stopped(p.stop)) return;
  const bool sub = MODE != DOTS && p.sub.nranks > 0;
  T tail = T(0);
  if (sub) tail = fused_coefs(p, cs);
  if (MODE == UPD && p.pred_w2 != nullptr) {
    // the reference's DGKS test (dsaitr.f:656), evaluated identically by every thread
    const T wn = sqrt(*p.pred_w2), rn = sqrt(sub ? tail : *p.pred_r2);
    if (rn > T(0.717f) * wn) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && p.flag_out) *p.flag_out = T(0);
      return;
    }
  }
  uint64_t* full = reinterpret_cast<uint64_t*>(gsum + MAXB * CB + 8);
  uint64_t* empty = full + MAXST;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (MODE != DOTS && !sub)
    for (int k = tid; k < MAXB * CB; k += kThreadsTma) cs[k] = (k < p.j) ? p.coef[k] : T(0);
  __syncthreads();

  const int64_t ntiles = (p.n + R - 1) / R;
  const uint32_t stage_elems = p.stage_bytes / sizeof(T);
  const uint32_t v_bytes = (uint32_t)p.nboxes * p.box_bytes;  // the x tile sits right behind the V boxes

  if (warp == NG * NCW) {
    // ---------------- producer: one elected lane feeds the ring ----------------
    // stage / phase are kept incrementally: no division on the critical path
    // (issuing the boxes from several lanes was measured slower: UTMALDG takes uniform-register operands, so the
    // compiler serialises per-lane issues; one elected lane issuing back to back is the fast form)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = v_bytes + (p.x_tma ? (uint32_t)(R * sizeof(T)) : 0u);
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(empty + s, ph ^ 1u);
        mbar_expect_tx(full + s, tx_bytes);
        const uint32_t dst = smem_u32(tiles) + (uint32_t)s * p.stage_bytes;
        for (int b = 0; b < p.nboxes; ++b) load_2d(dst + (uint32_t)b * p.box_bytes, &tmap, (int)(t * R), b * CB, full + s);
        if (p.x_tma) load_1d(dst + v_bytes, &xmap, (int)(t * R), full + s);
        if (++s == p.nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ---------------- consumers ----------------
    const int g = warp / NCW, gw = warp - g * NCW, gtid = tid - g * NCW * 32;
    T acc[MAXB];
#pragma unroll
    for (int b = 0; b < MAXB; ++b) acc[b] = T(0);
    T accn = T(0);
    const bool y_is_x = (p.y == p.x);
    // without TMA for the vector operand it is software-prefetched one tile ahead in registers
    T xn[4] = {T(0), T(0), T(0), T(0)};
    if (!p.x_tma) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int64_t r = ((int64_t)blockIdx.x + (int64_t)g * gridDim.x) * R + lane + 32 * m;
        xn[m] = (r < p.n) ? p.x[r] : T(0);
      }
    }
    int s = g, kt = 0;  // stage of my current tile (nstages is a multiple of NG), count of my tiles
    uint32_t ph = 0;
    for (int64_t t = (int64_t)blockIdx.x + (int64_t)g * gridDim.x; t < ntiles; t += (int64_t)NG * gridDim.x, ++kt) {
      const int64_t row0 = t * R;
      T xv[4];
      if (!p.x_tma) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          xv[m] = xn[m];
          const int64_t r = row0 + (int64_t)NG * gridDim.x * R + lane + 32 * m;
          xn[m] = (r < p.n) ? p.x[r] : T(0);
        }
      }
      mbar_wait(full + s, ph);
      const T* tile = tiles + (size_t)s * stage_elems;
      if (p.x_tma) {
        const T* xs = tile + v_bytes / sizeof(T);
#pragma unroll
        for (int m = 0; m < 4; ++m) xv[m] = xs[lane + 32 * m];
      }
      if (MODE == DOTS && gw == 0) {
        if (y_is_x) {
#pragma unroll
          for (int m = 0; m < 4; ++m) accn += xv[m] * xv[m];
        } else {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int64_t r = row0 + lane + 32 * m;
            accn += xv[m] * ((r < p.n) ? p.y[r] : T(0));
          }
        }
      }
      if (MODE != DOTS) {
        // phase 1: row-local product V(i,:)*coef, columns split in two interleaved halves
        const int row = gtid & (R - 1), half = gtid >> 7;
        const T* tr = tile + row;
        T a0 = T(0), a1 = T(0);
        int c = half;
#pragma unroll 4
        for (; c + 2 < p.j; c += 4) {
          a0 += tr[c * R] * cs[c];
          a1 += tr[(c + 2) * R] * cs[c + 2];
        }
        if (c < p.j) a0 += tr[c * R] * cs[c];
        T* psb = ps + g * 4 * R + (kt & 1) * 2 * R;
        psb[half * R + row] = a0 + a1;
        if (g == 0) asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory");
        else asm volatile("bar.sync 2, %0;" ::"n"(NCW * 32) : "memory");
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int rr = lane + 32 * m;
          xv[m] = xv[m] - (psb[rr] + psb[R + rr]);
        }
        if (gw < 4) {
          const T rq = (gw == 0) ? xv[0] : (gw == 1) ? xv[1] : (gw == 2) ? xv[2] : xv[3];
          const int64_t r = row0 + lane + 32 * gw;
          if (r < p.n) p.dst[r] = rq;
        }
        if (gw == 0) {
#pragma unroll
          for (int m = 0; m < 4; ++m) accn += xv[m] * xv[m];
        }
      }
      if (MODE != UPD) {
        // dots of my column of every box with the (updated) vector
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
          if (b < p.nboxes) {
            const T* col = tile + (size_t)(b * CB + gw) * R + lane;
            acc[b] += (col[0] * xv[0] + col[32] * xv[1]) + (col[64] * xv[2] + col[96] * xv[3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      s += NG;
      if (s >= p.nstages) { s -= p.nstages; ph ^= 1u; }
    }
    // per-CTA partials: group 1 parks its warp sums in shared memory, group 0 adds its own and publishes
    T* mine = p.partial + (size_t)blockIdx.x * p.pcols;
    if (MODE != UPD) {
#pragma unroll
      for (int b = 0; b < MAXB; ++b) acc[b] = warp_sum(acc[b]);
    }
    accn = warp_sum(accn);
    if (g == 1 && lane == 0) {
      if (MODE != UPD) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) gsum[b * CB + gw] = acc[b];
      }
      if (gw == 0) gsum[MAXB * CB] = accn;
    }
    asm volatile("bar.sync 3, %0;" ::"n"(NG * NCW * 32) : "memory");
    if (g == 0 && lane == 0) {
      if (MODE != UPD) {
#pragma unroll
        for (int b = 0; b < MAXB; ++b) {
          const int c = b * CB + gw;
          if (b < p.nboxes && c < p.j) mine[c] = acc[b] + gsum[c];
        }
      }
      if (gw == 0 && p.out != nullptr) mine[(MODE == UPD) ? 0 : p.j] = accn + gsum[MAXB * CB];
    }
  }
  if (p.out == nullptr) return;
  if (MODE == UPD && blockIdx.x == 0 && tid == 0 && p.flag_out) *p.flag_out = T(1);
  finish_grid_reduce(p.partial, p.pcols, (MODE == UPD) ? 1 : p.j + 1, p.out, p.ticket, &p.pub);
}

// ---------------------------------------------------------------------------------------------
// UPD / UPD_SPEC, second form: rows belong to lanes.  Warp gw of a group owns rows 16*gw .. 16*gw+15 of the
// tile; lane l works on row (l & 15) and on the columns of parity (l >> 4).  The row-local product V(i,:)*coef is
// finished inside the warp with one shuffle, so r(i) is known without a group barrier or a trip through shared
// memory, and the speculative dots s[c] += V(i,c)*r(i) accumulate in registers (KB = ceil(j/2) per lane), reduced
// across lanes and warps once per kernel.  Same ring, same producer, same aux layout as k_orth.
// ---------------------------------------------------------------------------------------------
template <typename T, bool SPEC, int KB>
__global__ void __launch_bounds__(kThreadsTma, 1)
k_upd(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap xmap, const OrthParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_trigger();
  T* tiles = reinterpret_cast<T*>(smem);
  unsigned char* aux = smem + (size_t)p.nstages * p.stage_bytes;
  T* wsum = reinterpret_cast<T*>(aux);                   // [NG*NCW warps][MAXB*CB columns]  (== NG*4*R elements)
  T* cs = wsum + NG * 4 * R;                             // [MAXB*CB] coefficients
  {
    // prologue that touches no global memory: may overlap the tail of the previous kernel (launch_pdl)
    uint64_t* full0 = reinterpret_cast<uint64_t*>(cs + MAXB * CB + MAXB * CB + 8);
    if (threadIdx.x == 0) {
      for (int s = 0; s < p.nstages; ++s) {
        mbar_init(full0 + s, 1);
        mbar_init(full0 + MAXST + s, NCW);
      }
      mbar_fence_init();
    }
  }
  pdl_wait();
  if (stopped(p.stop)) return;
  const bool sub = p.sub.nranks > 0;
  T tail = T(0);
  if (sub) tail = fused_coefs(p, cs);
  if (!SPEC && p.pred_w2 != nullptr) {
    // the reference's DGKS test (dsaitr.f:656), evaluated identically by every thread
    const T wn = sqrt(*p.pred_w2), rn = sqrt(sub ? tail : *p.pred_r2);
    if (rn > T(0.717f) * wn) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && p.flag_out) *p.flag_out = T(0);
      return;
    }
  }
  T* wnrm = cs + MAXB * CB;                              // [NG*NCW] per-warp ||r||^2
  uint64_t* full = reinterpret_cast<uint64_t*>(wnrm + MAXB * CB + 8);
  uint64_t* empty = full + MAXST;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!sub)
    for (int k = tid; k < MAXB * CB; k += kThreadsTma) cs[k] = (k < p.j) ? p.coef[k] : T(0);
  __syncthreads();

  const int64_t ntiles = (p.n + R - 1) / R;
  const uint32_t stage_elems = p.stage_bytes / sizeof(T);
  const uint32_t v_bytes = (uint32_t)p.nboxes * p.box_bytes;

  if (warp == NG * NCW) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = v_bytes + (p.x_tma ? (uint32_t)(R * sizeof(T)) : 0u);
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(empty + s, ph ^ 1u);
        mbar_expect_tx(full + s, tx_bytes);
        const uint32_t dst = smem_u32(tiles) + (uint32_t)s * p.stage_bytes;
        for (int b = 0; b < p.nboxes; ++b) load_2d(dst + (uint32_t)b * p.box_bytes, &tmap, (int)(t * R), b * CB, full + s);
        if (p.x_tma) load_1d(dst + v_bytes, &xmap, (int)(t * R), full + s);
        if (++s == p.nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    const int g = warp / NCW, gw = warp - g * NCW;
    const int half = lane >> 4, row = gw * 16 + (lane & 15);
    T acc[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) acc[k] = T(0);
    T accn = T(0);
    T xn = T(0);
    if (!p.x_tma) {
      const int64_t r = ((int64_t)blockIdx.x + (int64_t)g * gridDim.x) * R + row;
      xn = (r < p.n) ? p.x[r] : T(0);
    }
    int s = g;
    uint32_t ph = 0;
    const int j = p.j;
    for (int64_t t = (int64_t)blockIdx.x + (int64_t)g * gridDim.x; t < ntiles; t += (int64_t)NG * gridDim.x) {
      const int64_t row0 = t * R;
      T xv = xn;
      if (!p.x_tma) {
        const int64_t r = row0 + (int64_t)NG * gridDim.x * R + row;
        xn = (r < p.n) ? p.x[r] : T(0);
      }
      mbar_wait(full + s, ph);
      const T* tr = tiles + (size_t)s * stage_elems + row;
      if (p.x_tma) xv = tr[v_bytes / sizeof(T)];
      // phase 1: my half of the columns of my row, two chains; the other half comes over one shuffle
      T a0 = T(0), a1 = T(0);
      constexpr bool kKeep = SPEC && KB <= 16;  // few enough columns: the row stays in registers for phase 2
      T keep[kKeep ? KB : 1];
      if (kKeep) {
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const int cc = 2 * k + half;
          const T vv = (cc < j) ? tr[cc * R] : T(0);
          keep[k] = vv;
          if (k & 1) a1 += vv * cs[cc];   // cs is zero-padded up to MAXB*CB
          else a0 += vv * cs[cc];
        }
      } else {
        int c = half;
#pragma unroll 4
        for (; c + 2 < j; c += 4) {
          a0 += tr[c * R] * cs[c];
          a1 += tr[(c + 2) * R] * cs[c + 2];
        }
        if (c < j) a0 += tr[c * R] * cs[c];
      }
      T a = a0 + a1;
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      const T rv = xv - a;
      if (half == 0) {
        const int64_t r = row0 + row;
        if (r < p.n) p.dst[r] = rv;
        accn += rv * rv;
      }
      if (SPEC) {
        // phase 2: s[c] += V(i,c) * r(i) for my columns (registers, or the tile that is still in shared memory)
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          const int cc = 2 * k + half;
          if (kKeep) acc[k] += keep[k] * rv;
          else if (cc < j) acc[k] += tr[cc * R] * rv;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      s += NG;
      if (s >= p.nstages) { s -= p.nstages; ph ^= 1u; }
    }
    // once per kernel: lanes -> warp (fixed xor tree over the 16 rows), warps -> CTA (fixed order), CTA partial
    accn = warp_sum(accn);
    if (lane == 0) wnrm[warp] = accn;
    if (SPEC) {
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        T v = acc[k];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        if ((lane & 15) == 0 && 2 * k + half < MAXB * CB) wsum[warp * (MAXB * CB) + 2 * k + half] = v;
      }
    }
    asm volatile("bar.sync 3, %0;" ::"n"(NG * NCW * 32) : "memory");
    T* mine = p.partial + (size_t)blockIdx.x * p.pcols;
    if (SPEC && tid < j) {
      T sum = T(0);
#pragma unroll
      for (int w = 0; w < NG * NCW; ++w) sum += wsum[w * (MAXB * CB) + tid];
      mine[tid] = sum;
    }
    if (tid == NG * NCW * 32 - 1 && p.out != nullptr) {
      T sum = T(0);
#pragma unroll
      for (int w = 0; w < NG * NCW; ++w) sum += wnrm[w];
      mine[SPEC ? j : 0] = sum;
    }
  }
  if (p.out == nullptr) return;
  if (!SPEC && blockIdx.x == 0 && tid == 0 && p.flag_out) *p.flag_out = T(1);
  finish_grid_reduce(p.partial, p.pcols, SPEC ? p.j + 1 : 1, p.out, p.ticket, &p.pub);
}

// ---------------------------------------------------------------------------------------------
// VQ: out(:,0:kout) = V(:,0:kin) * Q.  Each consumer group owns a 128-row tile; a warp computes a
// register block of (32*RPL rows) x (4 columns): RPL V values per lane and four Q entries (broadcast
// from shared memory) feed 4*RPL FP64 FMAs per k, which keeps the shared-memory traffic below the HBM
// time of the tile.  out may alias V (the tile is complete in shared memory before anything is written).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct VqParams {
  int64_t n, ldo;
  int kin, kout, kp;  // kp = kout rounded up to a multiple of 4 (zero-padded Q columns)
  int nboxes, nstages, rpl_items;  // rpl_items = number of (row slice, column group) work items
  uint32_t stage_bytes, box_bytes;
  const T* q;  // device, packed column-major kin x kout
  T* out;
  int with_resid, beta_col;
  T sigma, beta;
  T* resid;
  T* partial;
  T* nrm2_out;
  unsigned int* ticket;
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, T (&q)[4]);
template <>
__device__ __forceinline__ void load4<double>(const double* p, double (&q)[4]) {
  const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
  q[0] = a.x; q[1] = a.y; q[2] = b.x; q[3] = b.y;
}
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&q)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w;
}

template <typename T, int RPL>
__global__ void __launch_bounds__(kThreadsTma, 1) k_vq_tma(const __grid_constant__ CUtensorMap tmap, const VqParams<T> p) {
  extern __shared__ __align__(128) unsigned char smem[];
  T* tiles = reinterpret_cast<T*>(smem);
  unsigned char* aux = smem + (size_t)p.nstages * p.stage_bytes;
  T* qs = reinterpret_cast<T*>(aux);                          // [kin][kp]
  T* red = qs + (size_t)MAXB * CB * MAXB * CB;                // [NG*NCW]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + NG * NCW + 8);
  uint64_t* empty = full + MAXST;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NCW);
    }
    mbar_fence_init();
  }
  for (int i = tid; i < p.kin * p.kp; i += kThreadsTma) {
    const int k = i / p.kp, c = i - k * p.kp;
    qs[i] = (c < p.kout) ? p.q[(size_t)c * p.kin + k] : T(0);
  }
  if (tid < NG * NCW + 8) red[tid] = T(0);
  __syncthreads();
  const int64_t ntiles = (p.n + R - 1) / R;
  const uint32_t stage_elems = p.stage_bytes / sizeof(T);
  T nrm = T(0);
  if (warp == NG * NCW) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(empty + s, ph ^ 1u);
        mbar_expect_tx(full + s, (uint32_t)p.nboxes * p.box_bytes);
        const uint32_t dst = smem_u32(tiles) + (uint32_t)s * p.stage_bytes;
        for (int b = 0; b < p.nboxes; ++b) load_2d(dst + (uint32_t)b * p.box_bytes, &tmap, (int)(t * R), b * CB, full + s);
        if (++s == p.nstages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    const int g = warp / NCW, gw = warp - g * NCW;
    constexpr int NRS = 4 / RPL;  // row slices of 32*RPL rows per tile
    int s = g;
    uint32_t ph = 0;
    for (int64_t t = (int64_t)blockIdx.x + (int64_t)g * gridDim.x; t < ntiles; t += (int64_t)NG * gridDim.x) {
      const int64_t row0 = t * R;
      mbar_wait(full + s, ph);
      const T* tile = tiles + (size_t)s * stage_elems;
      for (int item = gw; item < p.rpl_items; item += NCW) {
        const int rs = item % NRS, cg = item / NRS;
        const T* tv = tile + rs * 32 * RPL + lane;
        const T* qrow = qs + cg * 4;
        T acc[RPL][4];
#pragma unroll
        for (int m = 0; m < RPL; ++m)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[m][c] = T(0);
#pragma unroll 2
        for (int k = 0; k < p.kin; ++k) {
          T qv[4];
          load4<T>(qrow + (size_t)k * p.kp, qv);
          T vv[RPL];
#pragma unroll
          for (int m = 0; m < RPL; ++m) vv[m] = tv[(size_t)k * R + 32 * m];
#pragma unroll
          for (int m = 0; m < RPL; ++m)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[m][c] += vv[m] * qv[c];
        }
        // every warp of the group must have finished reading the tile before anyone overwrites V when
        // out aliases V: writes only touch this tile's rows, which no other tile reads, and the tile
        // itself lives in shared memory, so no further synchronisation is needed
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
          const int64_t r = row0 + rs * 32 * RPL + lane + 32 * m;
          if (r < p.n) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int col = cg * 4 + c;
              if (col < p.kout) p.out[r + (int64_t)col * p.ldo] = acc[m][c];
              if (p.with_resid && col == p.beta_col) {
                const T v = p.sigma * p.resid[r] + p.beta * acc[m][c];
                p.resid[r] = v;
                nrm += v * v;
              }
            }
            if (p.with_resid && p.beta_col < 0 && cg == 0) {
              const T v = p.sigma * p.resid[r];
              p.resid[r] = v;
              nrm += v * v;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      s += NG;
      if (s >= p.nstages) { s -= p.nstages; ph ^= 1u; }
    }
    nrm = warp_sum(nrm);
    if (lane == 0) red[warp] = nrm;
  }
  if (!p.with_resid || p.nrm2_out == nullptr) return;
  __syncthreads();
  if (tid == 0) {
    T sum = T(0);
#pragma unroll
    for (int w = 0; w < NG * NCW; ++w) sum += red[w];
    p.partial[blockIdx.x] = sum;
  }
  finish_grid_reduce(p.partial, 1, 1, p.nrm2_out, p.ticket);
}

// ---- host side -------------------------------------------------------------------------------
template <typename T>
size_t aux_bytes() {
  return sizeof(T) * (NG * 4 * R + 2 * MAXB * CB + 8) + sizeof(uint64_t) * 2 * MAXST;
}
template <typename T>
size_t aux_bytes_vq() {
  return sizeof(T) * ((size_t)MAXB * CB * MAXB * CB + NG * NCW + 8) + sizeof(uint64_t) * 2 * MAXST;
}

template <typename T>
void geometry(int j, bool x_tma, OrthParams<T>& p) {
  p.j = j;
  p.nboxes = (j + CB - 1) / CB;
  p.box_bytes = (uint32_t)(R * CB * sizeof(T));
  static const bool allow_xtma = !(getenv("AB200_XTMA") && std::strcmp(getenv("AB200_XTMA"), "0") == 0);
  x_tma = x_tma && allow_xtma;
  p.x_tma = x_tma ? 1 : 0;
  // the vector tile (R elements) rides behind the V boxes; stages stay 128-byte aligned
  p.stage_bytes = (uint32_t)p.nboxes * p.box_bytes + (x_tma ? (uint32_t)(R * sizeof(T)) : 0u);
  int ns = (int)(kTileBudget / p.stage_bytes);
  p.nstages = ns > MAXST ? MAXST : (ns < 2 ? 2 : ns);
  static const int force_max = getenv("AB200_MAXST") ? atoi(getenv("AB200_MAXST")) : 0;
  if (force_max > 1 && p.nstages > force_max) p.nstages = force_max;
  // The ring depth must be a multiple of the number of consumer groups: tile `it` goes to group it % NG
  // and to stage it % nstages, so only then does every stage (and its pair of mbarriers) belong to ONE
  // group, which therefore observes every phase of it.  With an odd depth a group would see every
  // other phase of a barrier and the parity test of mbarrier.try_wait could be satisfied by the
  // previous-but-one completion (observed on the B200: sporadic launch failures).
  p.nstages -= p.nstages % NG;
}

template <typename T, int MODE>
bool launch_orth(cudaStream_t stream, int num_sms, const T* v, int64_t ldv, OrthParams<T>& p, const char* name,
                 double bytes) {
  CUtensorMap map, xmap;
  if (!get_tensor_map(&map, v, (int)sizeof(T), p.n, ldv, p.j, R, CB)) return false;
  if (p.x_tma) {
    if (!get_tensor_map(&xmap, p.x, (int)sizeof(T), p.n, 0, 0, R, 0)) return false;
  } else {
    xmap = map;
  }
  static bool attr_set[kMaxDevices] = {};   // the attribute belongs to (function, device)
  const int dev = current_device_slot();
  const size_t smem = (size_t)p.nstages * p.stage_bytes + aux_bytes<T>();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(k_orth<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(kTileBudget + aux_bytes<T>())) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set[dev] = true;
  }
  const int64_t ntiles = (p.n + R - 1) / R;
  const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
  ProfScope ps(stream, name, bytes);
  AB200_CUDA_CHECK(launch_pdl(k_orth<T, MODE>, grid, kThreadsTma, smem, stream, map, xmap, p));
  launch_stats().kernels++;
  launch_stats().fast_path++;
  return true;
}

template <typename T, bool SPEC, int KB>
bool launch_upd_kb(cudaStream_t stream, int grid, size_t smem, const CUtensorMap& map, const CUtensorMap& xmap,
                   const OrthParams<T>& p) {
  static bool attr_set[kMaxDevices] = {};
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    if (cudaFuncSetAttribute(k_upd<T, SPEC, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(kTileBudget + aux_bytes<T>())) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set[dev] = true;
  }
  return launch_pdl(k_upd<T, SPEC, KB>, grid, kThreadsTma, smem, stream, map, xmap, p) == cudaSuccess;
}

// UPD / UPD_SPEC through k_upd (rows per lane); AB200_UPD=old keeps the k_orth form for A/B measurements
template <typename T, bool SPEC>
bool launch_upd(cudaStream_t stream, int num_sms, const T* v, int64_t ldv, OrthParams<T>& p, const char* name,
                double bytes) {
  static const bool old_form = getenv("AB200_UPD") && std::strcmp(getenv("AB200_UPD"), "old") == 0;
  // beyond 32 columns the row no longer fits in registers between the two phases and the column-per-warp form of
  // k_orth is the faster one (measured on config 3, ncv = 64: 6.0 vs 5.3 TB/s)
  if (old_form || (SPEC && p.j > 32)) return launch_orth<T, SPEC ? UPD_SPEC : UPD>(stream, num_sms, v, ldv, p, name, bytes);
  CUtensorMap map, xmap;
  if (!get_tensor_map(&map, v, (int)sizeof(T), p.n, ldv, p.j, R, CB)) return false;
  if (p.x_tma) {
    if (!get_tensor_map(&xmap, p.x, (int)sizeof(T), p.n, 0, 0, R, 0)) return false;
  } else {
    xmap = map;
  }
  const size_t smem = (size_t)p.nstages * p.stage_bytes + aux_bytes<T>();
  const int64_t ntiles = (p.n + R - 1) / R;
  const int grid = (int)(ntiles < num_sms ? ntiles : num_sms);
  ProfScope ps(stream, name, bytes);
  bool ok;
  if (!SPEC) ok = launch_upd_kb<T, false, 1>(stream, grid, smem, map, xmap, p);
  else if (p.j <= 16) ok = launch_upd_kb<T, true, 8>(stream, grid, smem, map, xmap, p);
  else ok = launch_upd_kb<T, true, 16>(stream, grid, smem, map, xmap, p);
  if (!ok) return false;
  launch_stats().kernels++;
  launch_stats().fast_path++;
  AB200_CUDA_CHECK(cudaGetLastError());
  return true;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

template <typename T>
bool CudaVecOps<T>::fast_path_ok(int64_t n, int j, const T* v, int64_t ldv) const {
  if (j < 1 || j > MAXB * CB || n < 1) return false;
  if (!aligned16(v)) return false;                                     // TMA: 16-byte aligned base ...
  if (((size_t)ldv * sizeof(T)) % 16 != 0) return false;               // ... and column stride
  if (n > 2147483647LL - R) return false;                              // TMA coordinates are int32
  return tensor_maps_available();
}

template <typename T>
bool CudaVecOps<T>::dots_tma(int64_t n, int j, const T* v, int64_t ldv, const T* x, const T* y, T* out) {
  OrthParams<T> p{};
  p.n = n;
  geometry<T>(j, aligned16(x), p);
  const int grid = (int)std::min<int64_t>((n + R - 1) / R, num_sms_);
  p.pcols = j + 1;
  ensure_partial((size_t)grid * p.pcols);
  p.x = x; p.y = y; p.partial = partial_; p.out = out; p.ticket = ticket_; p.stop = stop_;
  return launch_orth<T, DOTS>(stream_, num_sms_, v, ldv, p, "dots_tma",
                              (double)sizeof(T) * n * (j + (x == y ? 1.0 : 2.0)));
}

template <typename T>
bool CudaVecOps<T>::orth_step_tma(int64_t n, int j, const T* v, int64_t ldv, const T* w, T* resid, T* mbA, T* mbB,
                                  T* mbC) {
  const int grid = (int)std::min<int64_t>((n + R - 1) / R, num_sms_);
  ensure_partial((size_t)grid * (j + 1));
  // multi-GPU with peer memory: the three reductions of the step are fused into the kernels -- the last CTA of the
  // producer stores this rank's sums into every peer's slot, the consumer's prologue adds the slots in rank order
  // (peer.cuh).  Otherwise a separate all-reduce (one small kernel, or ncclAllReduce) follows each sweep.
  PeerReduce prA, prB, prC;
  const bool fuse = comm_ != nullptr && j + 1 <= kPeerMaxCount && agreed_fuse_ &&
                    nccl_peer_reduce_begin(comm_, 0, &prA) &&
                    nccl_peer_reduce_begin(comm_, 1, &prB) && nccl_peer_reduce_begin(comm_, 2, &prC);
  // sweep A: h = V^T w, ||w||^2
  {
    OrthParams<T> p{};
    p.n = n;
    geometry<T>(j, aligned16(w), p);
    p.pcols = j + 1;
    p.x = w; p.y = w; p.partial = partial_; p.out = mbA; p.ticket = ticket_; p.stop = stop_;
    if (fuse) p.pub = prA;
    if (!launch_orth<T, DOTS>(stream_, num_sms_, v, ldv, p, "dots_tma", (double)sizeof(T) * n * (j + 1.0))) {
      if (fuse) throw CudaError("dots_tma launch failed with a fused reduction in flight");
      return false;
    }
  }
  if (!fuse) allreduce_sum(mbA, (size_t)j + 1);
  // sweep B: r = w - V h, ||r||^2, s = V^T r (speculative)
  {
    OrthParams<T> p{};
    p.n = n;
    geometry<T>(j, aligned16(w), p);
    p.pcols = j + 1;
    p.x = w; p.dst = resid; p.coef = mbA; p.partial = partial_; p.out = mbB; p.ticket = ticket_; p.stop = stop_;
    if (fuse) { p.sub = prA; p.sub_log = mbA; p.pub = prB; }
    if (!launch_upd<T, true>(stream_, num_sms_, v, ldv, p, "update_spec_tma", (double)sizeof(T) * n * (j + 2.0)))
      throw CudaError("update_spec_tma launch failed after dots_tma succeeded");
  }
  if (!fuse) allreduce_sum(mbB, (size_t)j + 1);
  // sweep C (only if the DGKS test fires, decided on the device): r -= V s, ||r||^2
  {
    OrthParams<T> p{};
    p.n = n;
    geometry<T>(j, aligned16(resid), p);
    p.pcols = 1;
    p.x = resid; p.dst = resid; p.coef = mbB; p.partial = partial_; p.out = mbC; p.ticket = ticket_;
    p.pred_w2 = mbA + j; p.pred_r2 = mbB + j; p.flag_out = mbC + 1; p.stop = stop_;
    if (fuse) { p.sub = prB; p.sub_log = mbB; p.pub = prC; }
    if (!launch_upd<T, false>(stream_, num_sms_, v, ldv, p, "reorth_tma", (double)sizeof(T) * n * (j + 2.0)))
      throw CudaError("reorth_tma launch failed after dots_tma succeeded");
  }
  if (fuse) {
    // ||r'||^2 stays in the peer slots until somebody needs it: the gated start of the next step (attach_pending)
    // or, failing that, a one-block finalize kernel in front of the next mailbox read (resolve_pending)
    has_pending_ = true;
    pending_ = prC;
    pending_log_ = mbC;
    pending_w2_ = mbA + j;
    pending_r2_ = mbB + j;
    pending_stop_ = stop_;
  } else {
    allreduce_sum(mbC, 1);
  }
  return true;
}

template <typename T>
bool CudaVecOps<T>::vq_tma(int64_t n, int kin, int kout, const T* v, int64_t ldv, const T* qdev, T* out, int64_t ldo,
                           bool with_resid, T sigma, T beta, int beta_col, T* resid, T* mb_nrm2) {
  if (kin < 1 || kin > MAXB * CB || kout < 1 || kout > MAXB * CB) return false;
  if (!fast_path_ok(n, kin, v, ldv)) return false;
  VqParams<T> p{};
  p.n = n; p.ldo = ldo; p.kin = kin; p.kout = kout; p.kp = (kout + 3) & ~3;
  p.nboxes = (kin + CB - 1) / CB;
  p.box_bytes = (uint32_t)(R * CB * sizeof(T));
  p.stage_bytes = (uint32_t)p.nboxes * p.box_bytes;
  int ns = (int)(kTileBudgetVq / p.stage_bytes);
  p.nstages = ns > MAXST ? MAXST : (ns < 2 ? 2 : ns);
  p.nstages -= p.nstages % NG;  // see geometry(): one consumer group per stage
  p.q = qdev; p.out = out;
  p.with_resid = with_resid ? 1 : 0; p.beta_col = beta_col; p.sigma = sigma; p.beta = beta; p.resid = resid;
  p.nrm2_out = mb_nrm2; p.ticket = ticket_;
  const int ncg = p.kp / 4;
  // rows per lane: enough (row slice x column group) items to occupy the 8 warps of a group
  const int rpl = (ncg >= 8) ? 4 : (ncg >= 4) ? 2 : 1;
  p.rpl_items = (4 / rpl) * ncg;
  CUtensorMap map;
  if (!get_tensor_map(&map, v, (int)sizeof(T), n, ldv, kin, R, CB)) return false;
  const int64_t ntiles = (n + R - 1) / R;
  const int grid = (int)(ntiles < num_sms_ ? ntiles : num_sms_);
  ensure_partial((size_t)grid);
  p.partial = partial_;
  const size_t smem = (size_t)p.nstages * p.stage_bytes + aux_bytes_vq<T>();
  const int smem_max = (int)(kTileBudgetVq + aux_bytes_vq<T>());
  static bool attr_set_dev[kMaxDevices][3] = {};
  bool* attr_set = attr_set_dev[current_device_slot()];
  const int ai = rpl == 4 ? 2 : (rpl == 2 ? 1 : 0);
  if (!attr_set[ai]) {
    cudaError_t e = cudaSuccess;
    if (rpl == 4) e = cudaFuncSetAttribute(k_vq_tma<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    else if (rpl == 2) e = cudaFuncSetAttribute(k_vq_tma<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    else e = cudaFuncSetAttribute(k_vq_tma<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    attr_set[ai] = true;
  }
  ProfScope ps(stream_, "vq_tma", (double)sizeof(T) * n * (kin + kout + (with_resid ? 2.0 : 0.0)));
  if (rpl == 4) k_vq_tma<T, 4><<<grid, kThreadsTma, smem, stream_>>>(map, p);
  else if (rpl == 2) k_vq_tma<T, 2><<<grid, kThreadsTma, smem, stream_>>>(map, p);
  else k_vq_tma<T, 1><<<grid, kThreadsTma, smem, stream_>>>(map, p);
  launch_stats().kernels++;
  launch_stats().fast_path++;
  AB200_CUDA_CHECK(cudaGetLastError());
  return true;
}
template bool CudaVecOps<double>::fast_path_ok(int64_t, int, const double*, int64_t) const;
template bool CudaVecOps<float>::fast_path_ok(int64_t, int, const float*, int64_t) const;
template bool CudaVecOps<double>::orth_step_tma(int64_t, int, const double*, int64_t, const double*, double*, double*, double*, double*);
template bool CudaVecOps<float>::orth_step_tma(int64_t, int, const float*, int64_t, const float*, float*, float*, float*, float*);
template bool CudaVecOps<double>::dots_tma(int64_t, int, const double*, int64_t, const double*, const double*, double*);
template bool CudaVecOps<float>::dots_tma(int64_t, int, const float*, int64_t, const float*, const float*, float*);
template bool CudaVecOps<double>::vq_tma(int64_t, int, int, const double*, int64_t, const double*, double*, int64_t, bool, double, double, int, double*, double*);
template bool CudaVecOps<float>::vq_tma(int64_t, int, int, const float*, int64_t, const float*, float*, int64_t, bool, float, float, int, float*, float*);

}  // namespace ab200
