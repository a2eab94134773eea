// driver.cu -- the caller's side of the reverse-communication loop, on the device.
//
// In the reference the matrix belongs to the user: EXAMPLES/SIMPLE/dssimp.f:484-540 (av/tv),
// EXAMPLES/NONSYM/dndrv1.f:453-470, PARPACK/EXAMPLES/MPI/pdsdrv1.f:418-483 (av with halo send/recv),
// EXAMPLES/MATRIX_MARKET/arpackSolver.hpp:806-841 (Eigen products).  These kernels are the B200
// equivalents used by the arpackmm-style driver, the tests and bench.py: CSR SpMV (K3), synthetic
// operator generators for BASELINE.json's configs, the hashed start vector, and the residual check.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/arpack_b200.h"
#include "driver.hpp"
#include "gate.cuh"
#include "tma_common.cuh"
#include "vecops_cuda.cuh"

namespace ab200 {
NcclComm* comm_from_handle(int handle);
void nccl_halo_exchange(NcclComm* c, const void* send_lo, void* recv_lo, size_t n_lo, const void* send_hi,
                        void* recv_hi, size_t n_hi, bool is_double, cudaStream_t s);
bool nccl_halo_is_peer_buffer(const NcclComm* c, const void* buf);
const void* nccl_halo_exchange_peer(NcclComm* c, const void* send_lo, const void* send_hi, cudaStream_t s);

namespace {

inline cudaStream_t cur_stream() { return (cudaStream_t)ab200_get_stream(); }

// SpMV kernel choice for short-row matrices: 0 = CSR-bulk (default), 1 = CSR-stream, 2 = row per sub-warp.
// Initialised from AB200_SPMV={bulk,stream,subwarp}; ab200_set_spmv_variant() overrides it (tests, tuning).
int& spmv_variant() {
  static int v = [] {
    const char* e = getenv("AB200_SPMV");
    if (e && std::strcmp(e, "stream") == 0) return 1;
    if (e && std::strcmp(e, "subwarp") == 0) return 2;
    return 0;
  }();
  return v;
}

// ---------------------------------------------------------------------------------------------
// CSR SpMV, LPR lanes per row (power of two <= 32).  Rows of the target operators are short
// (5..16 nnz): a sub-warp per row keeps the val/col streams coalesced across neighbouring rows
// while x is gathered through L1/L2.  Algorithmic traffic: nnz*(w+4) + (n+1)*4 + 2*n*w bytes.
// xh: optional second source for columns >= nloc (halo buffer of the row-partitioned operator).
// ---------------------------------------------------------------------------------------------
template <typename T, int LPR>
__global__ void __launch_bounds__(256) k_csr_spmv(int nrows, const int* __restrict__ rowptr,
                                                  const int* __restrict__ col, const T* __restrict__ val,
                                                  const T* __restrict__ x, T* __restrict__ y, int nloc,
                                                  const T* __restrict__ xh) {
  const int lane = threadIdx.x & (LPR - 1);
  const long long sub = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
  const long long nsub = ((long long)gridDim.x * blockDim.x) / LPR;
  for (long long row = sub; row < nrows; row += nsub) {
    const int p0 = rowptr[row], p1 = rowptr[row + 1];
    T acc = T(0);
    for (int p = p0 + lane; p < p1; p += LPR) {
      const int c = col[p];
      const T xv = (xh != nullptr && c >= nloc) ? xh[c - nloc] : x[c];
      acc += val[p] * xv;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, LPR);
    if (lane == 0) y[row] = acc;
  }
}

// "CSR-stream" variant for short rows: a CTA owns ROWS consecutive rows, multiplies their non-zeros
// in the coalesced order of the val/col streams (thread i takes entries i, i+ROWS, ...; four
// independent col/val loads and four x gathers in flight per thread), parks the products in shared
// memory, then every thread sums its own row in order.  Only two dependent memory levels per entry
// instead of the three of the row-per-sub-warp kernel, and deterministic row sums.
//
// FUSED (registered-operator mode of the solver, square A): the operand is xs*x with x = the
// unnormalised residual and xs = 1/||r|| (K1/K2 folded into K3: v_j = xs*x is written on the way, the
// scaled copies of x are never materialised), and the epilogue accumulates alpha = v_j^T (A v_j) and
// ||A v_j||^2 into dots_out[0..1] (north_star: "alpha = v^T w fused into the SpMV epilogue").
template <typename T, int ROWS, int CAP, bool FUSED>
__global__ void __launch_bounds__(ROWS) k_csr_spmv_stream(int nrows, const int* __restrict__ rowptr,
                                                          const int* __restrict__ col, const T* __restrict__ val,
                                                          const T* __restrict__ x, T* __restrict__ y, int nloc,
                                                          const T* __restrict__ xh, T xs, T* __restrict__ vj_out,
                                                          T* __restrict__ partial, T* __restrict__ dots_out,
                                                          unsigned int* ticket, const StepGate<T> gate,
                                                          const T* stop) {
  __shared__ T prod[CAP];
  if (FUSED) {
    if (stopped(stop)) return;
    if (gate.stop != nullptr && !gate_eval_block(gate, xs)) {
      if (blockIdx.x == 0 && threadIdx.x == 0) *gate.stop = gate.stop_code;
      return;
    }
  }
  __shared__ int srow[ROWS + 1];
  __shared__ T red[2][ROWS / 32];
  const int tid = threadIdx.x;
  T dxy = T(0), dyy = T(0);
  for (long long blk = blockIdx.x; blk * ROWS < nrows; blk += gridDim.x) {
    const int r0 = (int)(blk * ROWS);
    const int nr = (nrows - r0 < ROWS) ? (nrows - r0) : ROWS;
    if (tid < nr) srow[tid] = rowptr[r0 + tid];
    if (tid == 0) srow[nr] = rowptr[r0 + nr];
    T xrow = T(0);
    if (FUSED && tid < nr) xrow = xs * x[r0 + tid];
    __syncthreads();
    const int p0 = srow[0], p1 = srow[nr];
    const int rs = (tid < nr) ? srow[tid] : p1, re = (tid < nr) ? srow[tid + 1] : p1;
    T acc = T(0);
    for (int c0 = p0; c0 < p1; c0 += CAP) {
      const int c1 = (c0 + CAP < p1) ? c0 + CAP : p1;
      int i = c0 + tid;
      for (; i + 3 * ROWS < c1; i += 4 * ROWS) {
        const int ca = col[i], cb = col[i + ROWS], cc = col[i + 2 * ROWS], cd = col[i + 3 * ROWS];
        const T va = val[i], vb = val[i + ROWS], vc = val[i + 2 * ROWS], vd = val[i + 3 * ROWS];
        const T xa = (xh != nullptr && ca >= nloc) ? xh[ca - nloc] : x[ca];
        const T xb = (xh != nullptr && cb >= nloc) ? xh[cb - nloc] : x[cb];
        const T xc = (xh != nullptr && cc >= nloc) ? xh[cc - nloc] : x[cc];
        const T xd = (xh != nullptr && cd >= nloc) ? xh[cd - nloc] : x[cd];
        prod[i - c0] = va * xa;
        prod[i + ROWS - c0] = vb * xb;
        prod[i + 2 * ROWS - c0] = vc * xc;
        prod[i + 3 * ROWS - c0] = vd * xd;
      }
      for (; i < c1; i += ROWS) {
        const int ca = col[i];
        const T xa = (xh != nullptr && ca >= nloc) ? xh[ca - nloc] : x[ca];
        prod[i - c0] = val[i] * xa;
      }
      __syncthreads();
      const int a = (rs > c0) ? rs : c0, b = (re < c1) ? re : c1;
      for (int q = a; q < b; ++q) acc += prod[q - c0];
      __syncthreads();
    }
    if (FUSED) acc *= xs;  // A*(xs*x) = xs*(A*x): one multiply per row instead of one per entry
    if (tid < nr) {
      y[r0 + tid] = acc;
      if (FUSED) {
        if (vj_out != nullptr) vj_out[r0 + tid] = xrow;
        dxy += xrow * acc;
        dyy += acc * acc;
      }
    }
  }
  if (!FUSED || dots_out == nullptr) return;
  dxy = tma::warp_sum(dxy);
  dyy = tma::warp_sum(dyy);
  if ((tid & 31) == 0) { red[0][tid >> 5] = dxy; red[1][tid >> 5] = dyy; }
  __syncthreads();
  if (tid < 2) {
    T sum = T(0);
#pragma unroll
    for (int w = 0; w < ROWS / 32; ++w) sum += red[tid][w];
    partial[(size_t)blockIdx.x * 2 + tid] = sum;
  }
  tma::finish_grid_reduce(partial, 2, 2, dots_out, ticket);
}

// "CSR-bulk" variant: the same row-block scheme fed by the copy engine.  A persistent CTA walks row blocks of ROWS
// rows; one producer thread streams each block's slice of the col/val arrays (contiguous in CSR) and its ROWS+1 row
// pointers into a shared-memory ring with cp.async.bulk (1-D bulk copies, 16-byte granules, mbarrier complete_tx), so the
// DRAM streams stay in flight while the ROWS consumer threads are busy with the x gathers of earlier blocks.  The
// consumers are software-pipelined one block deep: the x gathers of block b+1 are issued (into registers) before the
// barrier / row sums / y store of block b.  Products are parked in the staged val slice, consumers synchronise among
// themselves (named barrier; the producer warp never joins) and every thread sums its own row in entry order -- the
// row sums are bit-identical to the CSR-stream kernel's.
// Bulk copies need 16-byte aligned sources: a block's slice starts at its first entry rounded down to a multiple of
// 4 entries (the pad entries belong to the previous block and are ignored) and never extends past nnz & ~3; the <= 3
// tail entries of the matrix are read directly.  A block with more than CAP entries is not staged: its consumers
// stream it from global memory chunk by chunk (the CSR-stream scheme) with the slot's val array as scratch.
template <typename T, int ROWS, int CAP>
struct SpmvBulkStage {
  T val[CAP];
  int col[CAP];
  int rp[ROWS + 4];
};

__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(tma::smem_u32(bar))
               : "memory");
}

// OWN (default): the consumers do not park products at all -- thread t reads the entries of ITS row straight from the
// staged slice (row lengths of the target operators are <= EPT, consecutive rows start ~EPT entries apart, an odd
// stride for the stencils: conflict-free), gathers x and sums in entry order.  Per entry that is two shared-memory
// loads and one gather instead of two loads, one store, one more load and a named barrier per block: the kernel was
// bound by L1/LSU throughput (ncu: l1tex 82 %, DRAM 69 %), not by DRAM.  Products and sums are rounded separately
// (no FMA contraction) so that every variant returns the same bits.
template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }

template <typename T, int ROWS, int EPT, int NST, bool FUSED, bool OWN>
__global__ void __launch_bounds__(ROWS + 32) k_csr_spmv_bulk(int nrows, const int* __restrict__ rowptr,
                                                             const int* __restrict__ col, const T* __restrict__ val,
                                                             const T* __restrict__ x, T* __restrict__ y, int nloc,
                                                             const T* __restrict__ xh, T xs, T* __restrict__ vj_out,
                                                             T* __restrict__ partial, T* __restrict__ dots_out,
                                                             unsigned int* ticket, const StepGate<T> gate,
                                                             const T* stop) {
  constexpr int CAP = ROWS * EPT;  // entries per ring slot, alignment pad included
  using Stage = SpmvBulkStage<T, ROWS, CAP>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Stage* stages = reinterpret_cast<Stage*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage) * NST);
  uint64_t* empty = full + NST;
  __shared__ T red[2][ROWS / 32];
  const int tid = threadIdx.x;
  const int nblk = (nrows + ROWS - 1) / ROWS;
  const int gstep = (int)gridDim.x;
  // prologue without global-memory accesses: may overlap the tail of the previous kernel (tma::launch_pdl)
  tma::pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < NST; ++i) {
      tma::mbar_init(&full[i], 1);
      tma::mbar_init(&empty[i], ROWS / 32);
    }
    tma::mbar_fence_init();
  }
  tma::pdl_wait();
  if (FUSED) {
    if (stopped(stop)) return;
    if (gate.stop != nullptr && !gate_eval_block(gate, xs)) {
      if (blockIdx.x == 0 && threadIdx.x == 0) *gate.stop = gate.stop_code;
      return;
    }
  }
  __syncthreads();
  const int nnz4 = __ldg(rowptr + nrows) & ~3;
  T dxy = T(0), dyy = T(0);
  if (tid >= ROWS) {
    // ---------------- producer warp ----------------
    const int lane = tid - ROWS;
    int stage = 0;
    uint32_t phase = 0;
    for (int base = blockIdx.x; base < nblk; base += 32 * gstep) {
      // row-pointer bounds of the next 32 blocks of this CTA, one per lane (one load latency per 32 blocks)
      const int myblk = base + lane * gstep;
      int q0 = 0, q1 = 0;
      if (myblk < nblk) {
        const int r0 = myblk * ROWS;
        q0 = __ldg(rowptr + r0);
        q1 = __ldg(rowptr + ((nrows - r0 > ROWS) ? r0 + ROWS : nrows));
      }
      for (int l = 0; l < 32; ++l) {
        const int blk = base + l * gstep;
        if (blk >= nblk) break;
        const int p0 = __shfl_sync(0xffffffffu, q0, l), p1 = __shfl_sync(0xffffffffu, q1, l);
        if (lane == 0) {
          const int r0 = blk * ROWS;
          const bool rp_bulk = (nrows + 1 - r0 >= ROWS + 4);
          const int sidx = p0 & ~3;
          int ce = (p1 + 3) & ~3;
          if (ce > nnz4) ce = nnz4;
          int cnt = ce > sidx ? ce - sidx : 0;
          if (p1 - sidx > CAP) cnt = 0;  // oversized block: consumers stream it themselves
          tma::mbar_wait(&empty[stage], phase ^ 1u);
          Stage* st = &stages[stage];
          const uint32_t bytes = (uint32_t)cnt * (uint32_t)(sizeof(T) + 4) + (rp_bulk ? (ROWS + 4) * 4u : 0u);
          if (bytes) {
            tma::mbar_expect_tx(&full[stage], bytes);
            if (cnt) {
              bulk_load(tma::smem_u32(st->val), val + sidx, (uint32_t)cnt * (uint32_t)sizeof(T), &full[stage]);
              bulk_load(tma::smem_u32(st->col), col + sidx, (uint32_t)cnt * 4u, &full[stage]);
            }
            if (rp_bulk) bulk_load(tma::smem_u32(st->rp), rowptr + r0, (ROWS + 4) * 4u, &full[stage]);
          } else {
            tma::mbar_arrive(&full[stage]);
          }
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (OWN) {
    // ---------------- consumers, row-owner form: thread t reads row r0 + t from the staged slice ----------------
    int stage = 0;
    uint32_t phase = 0;
    for (int blk = blockIdx.x; blk < nblk; blk += gstep) {
      tma::mbar_wait(&full[stage], phase);
      Stage* st = &stages[stage];
      const int r0 = blk * ROWS;
      const int nr = (nrows - r0 < ROWS) ? (nrows - r0) : ROWS;
      int p0, p1, rs, re;
      if (nrows + 1 - r0 >= ROWS + 4) {
        p0 = st->rp[0]; p1 = st->rp[ROWS]; rs = st->rp[tid]; re = st->rp[tid + 1];
      } else {
        p0 = __ldg(rowptr + r0);
        p1 = __ldg(rowptr + r0 + nr);
        rs = (tid < nr) ? __ldg(rowptr + r0 + tid) : p1;
        re = (tid < nr) ? __ldg(rowptr + r0 + tid + 1) : p1;
      }
      T xrow = T(0);
      if (FUSED && tid < nr) xrow = xs * x[r0 + tid];
      const int sidx = p0 & ~3;
      int clen = ((p1 + 3) & ~3);
      clen = (clen > nnz4 ? nnz4 : clen) - sidx;   // entries of the slice that were staged
      if (p1 - sidx > CAP) clen = 0;               // oversized block: nothing was staged, everything comes from global
      T acc = T(0);
      // the first EPT entries of my row: columns and values first, then all gathers in flight at once
      int cc[EPT];
      T vv[EPT], xx[EPT];
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int q = rs + k, i = q - sidx;
        cc[k] = 0;
        vv[k] = T(0);
        if (q < re) {
          if (i < clen) { cc[k] = st->col[i]; vv[k] = st->val[i]; }
          else { cc[k] = __ldg(col + q); vv[k] = __ldg(val + q); }
        }
      }
#pragma unroll
      for (int k = 0; k < EPT; ++k)
        xx[k] = (rs + k < re) ? ((xh != nullptr && cc[k] >= nloc) ? xh[cc[k] - nloc] : x[cc[k]]) : T(0);
#pragma unroll
      for (int k = 0; k < EPT; ++k)
        if (rs + k < re) acc = add_rn(acc, mul_rn(vv[k], xx[k]));
      for (int q = rs + EPT; q < re; ++q) {   // rows longer than EPT entries
        const int i = q - sidx;
        const int c = (i < clen) ? st->col[i] : __ldg(col + q);
        const T v = (i < clen) ? st->val[i] : __ldg(val + q);
        const T xv = (xh != nullptr && c >= nloc) ? xh[c - nloc] : x[c];
        acc = add_rn(acc, mul_rn(v, xv));
      }
      __syncwarp();
      if ((tid & 31) == 0) tma::mbar_arrive(&empty[stage]);
      if (FUSED) acc *= xs;  // A*(xs*x) = xs*(A*x): one multiply per row instead of one per entry
      if (tid < nr) {
        y[r0 + tid] = acc;
        if (FUSED) {
          if (vj_out != nullptr) vj_out[r0 + tid] = xrow;
          dxy += xrow * acc;
          dyy += acc * acc;
        }
      }
      if (++stage == NST) { stage = 0; phase ^= 1u; }
    }
  } else {
    // ---------------- consumers: thread t owns row r0 + t ----------------
    // registers of one block: its entries' columns, values and gathered x, and the owner row's bounds
    struct Item {
      T xx[EPT];
      int rs, re, sidx, p1;
      T xrow;
      bool direct;
    };
    auto load_item = [&](Item& it, int blk, Stage* st) {
      const int r0 = blk * ROWS;
      const int nr = (nrows - r0 < ROWS) ? (nrows - r0) : ROWS;
      int p0;
      if (nrows + 1 - r0 >= ROWS + 4) {
        p0 = st->rp[0];
        it.p1 = st->rp[ROWS];
        it.rs = st->rp[tid];
        it.re = st->rp[tid + 1];
      } else {
        p0 = __ldg(rowptr + r0);
        it.p1 = __ldg(rowptr + r0 + nr);
        it.rs = (tid < nr) ? __ldg(rowptr + r0 + tid) : it.p1;
        it.re = (tid < nr) ? __ldg(rowptr + r0 + tid + 1) : it.p1;
      }
      it.xrow = T(0);
      if (FUSED && tid < nr) it.xrow = xs * x[r0 + tid];
      it.sidx = p0 & ~3;
      const int len = it.p1 - it.sidx;
      it.direct = len > CAP;
      if (it.direct) return;
      int clen = ((it.p1 + 3) & ~3);
      clen = (clen > nnz4 ? nnz4 : clen) - it.sidx;
      // columns of my entries (the <= 3 tail entries of the matrix were not staged: fetch them into the slot),
      // then all of my x gathers at once; the values stay in the slot until the products are formed
      int cc[EPT];
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = tid + k * ROWS;
        cc[k] = 0;
        if (i < clen) cc[k] = st->col[i];
        else if (i < len) { cc[k] = __ldg(col + it.sidx + i); st->val[i] = __ldg(val + it.sidx + i); }
      }
#pragma unroll
      for (int k = 0; k < EPT; ++k)
        it.xx[k] = (xh != nullptr && cc[k] >= nloc) ? xh[cc[k] - nloc] : x[cc[k]];
    };
    int stage = 0;
    uint32_t phase = 0;
    int blk = blockIdx.x;
    Item cur, nxt;
    if (blk < nblk) {
      tma::mbar_wait(&full[0], 0u);
      load_item(cur, blk, &stages[0]);
    }
    while (blk < nblk) {
      Stage* st = &stages[stage];
      const int r0 = blk * ROWS;
      const int nr = (nrows - r0 < ROWS) ? (nrows - r0) : ROWS;
      T acc = T(0);
      if (!cur.direct) {
        const int len = cur.p1 - cur.sidx;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          const int i = tid + k * ROWS;
          if (i < len) st->val[i] = st->val[i] * cur.xx[k];
        }
      } else {
        // oversized block, streamed from global memory in chunks of CAP entries
        for (int c0 = cur.sidx; c0 < cur.p1; c0 += CAP) {
          const int c1 = (cur.p1 - c0 > CAP) ? c0 + CAP : cur.p1;
          for (int i = c0 + tid; i < c1; i += ROWS) {
            const int c = __ldg(col + i);
            const T xv = (xh != nullptr && c >= nloc) ? xh[c - nloc] : x[c];
            st->val[i - c0] = __ldg(val + i) * xv;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(ROWS) : "memory");
          const int a = (cur.rs > c0) ? cur.rs : c0, b = (cur.re < c1) ? cur.re : c1;
          for (int q = a; q < b; ++q) acc += st->val[q - c0];
          asm volatile("bar.sync 1, %0;" ::"n"(ROWS) : "memory");
        }
      }
      // the next block's gathers go out before this block's barrier and row sums
      const int nblk_next = blk + gstep;
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == NST) { nstage = 0; nphase ^= 1u; }
      if (nblk_next < nblk) {
        tma::mbar_wait(&full[nstage], nphase);
        load_item(nxt, nblk_next, &stages[nstage]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ROWS) : "memory");
      if (!cur.direct)
        for (int q = cur.rs; q < cur.re; ++q) acc += st->val[q - cur.sidx];
      __syncwarp();
      if ((tid & 31) == 0) tma::mbar_arrive(&empty[stage]);
      if (FUSED) acc *= xs;  // A*(xs*x) = xs*(A*x): one multiply per row instead of one per entry
      if (tid < nr) {
        y[r0 + tid] = acc;
        if (FUSED) {
          if (vj_out != nullptr) vj_out[r0 + tid] = cur.xrow;
          dxy += cur.xrow * acc;
          dyy += acc * acc;
        }
      }
      cur = nxt;
      stage = nstage;
      phase = nphase;
      blk = nblk_next;
    }
  }
  if (!FUSED || dots_out == nullptr) return;
  if (tid < ROWS) {
    dxy = tma::warp_sum(dxy);
    dyy = tma::warp_sum(dyy);
    if ((tid & 31) == 0) { red[0][tid >> 5] = dxy; red[1][tid >> 5] = dyy; }
  }
  __syncthreads();
  if (tid < 2) {
    T sum = T(0);
#pragma unroll
    for (int w = 0; w < ROWS / 32; ++w) sum += red[tid][w];
    partial[(size_t)blockIdx.x * 2 + tid] = sum;
  }
  tma::finish_grid_reduce(partial, 2, 2, dots_out, ticket);
}

template <typename T, int ROWS, int EPT, int NST, bool FUSED, bool OWN>
int launch_spmv_bulk_cfg(cudaStream_t s, int ctas_per_sm, int nrows, long long nnz, const int* rowptr, const int* col,
                         const T* val, const T* x, T* y, int nloc, const T* xh, T xs, T* vj_out, T* partial,
                         T* dots_out, unsigned int* ticket, const StepGate<T>& gate, const T* stop) {
  using Stage = SpmvBulkStage<T, ROWS, ROWS * EPT>;
  constexpr size_t smem = sizeof(Stage) * NST + 2 * NST * sizeof(uint64_t);
  auto kern = k_csr_spmv_bulk<T, ROWS, EPT, NST, FUSED, OWN>;
  static bool attr_done[tma::kMaxDevices] = {};   // (function, device) attribute
  static int sms_of[tma::kMaxDevices] = {};
  const int slot = tma::current_device_slot();
  if (!attr_done[slot]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    attr_done[slot] = true;
  }
  if (!sms_of[slot]) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms_of[slot], cudaDevAttrMultiProcessorCount, dev);
    if (sms_of[slot] <= 0) sms_of[slot] = 148;
  }
  const int sms = sms_of[slot];
  const long long nblk = ((long long)nrows + ROWS - 1) / ROWS;
  long long g = (long long)sms * ctas_per_sm;
  if (g > nblk) g = nblk;
  (void)nnz;
  const cudaError_t e = tma::launch_pdl(kern, (int)g, ROWS + 32, smem, s, nrows, rowptr, col, val, x, y, nloc, xh, xs,
                                        vj_out, partial, dots_out, ticket, gate, stop);
  launch_stats().kernels++;
  if (e != cudaSuccess) cudaGetLastError();
  return e == cudaSuccess ? 0 : -1;
}

// the bulk kernel needs 16-byte aligned CSR arrays (cp.async.bulk) and the true nnz
inline bool spmv_bulk_ok(const int* rowptr, const int* col, const void* val, long long nnz) {
  return (spmv_variant() == 0 || spmv_variant() == 4) && nnz > 0 && (((uintptr_t)rowptr | (uintptr_t)col | (uintptr_t)val) & 15u) == 0;
}

template <typename T, bool FUSED>
int launch_spmv_bulk(cudaStream_t s, int nrows, long long nnz, const int* rowptr, const int* col, const T* val,
                     const T* x, T* y, int nloc, const T* xh, T xs, T* vj_out, T* partial, T* dots_out,
                     unsigned int* ticket, const StepGate<T>& gate = StepGate<T>(), const T* stop = nullptr) {
  static const int variant = getenv("AB200_SPMV_BULK") ? atoi(getenv("AB200_SPMV_BULK")) : 0;
  // row-owner consumers (default) or the product-parking form (AB200_SPMV_OWN=0 / ab200_set_spmv_variant(4))
  // measured (B200, SpMV alone, L2 flushed): 5-point 2-D stencil 6.07 TB/s row-owner vs 5.63 parked; 7-point 3-D stencil
  // 6.46 vs 5.25 with three CTAs per SM (the 8 entries per thread cost 72 registers: a fourth CTA does not fit, and a
  // grid of 4 per SM then runs in two uneven waves: 4.84).  AB200_SPMV_OWN=0 switches back to the parked products.
  static const int own_env = getenv("AB200_SPMV_OWN") ? atoi(getenv("AB200_SPMV_OWN")) : 1;
  const bool own = own_env != 0 && spmv_variant() != 4;
#define AB200_BULK_CFG(ROWS_, EPT_, NST_, CTAS_)                                                                    \
  return own ? launch_spmv_bulk_cfg<T, ROWS_, EPT_, NST_, FUSED, true>(s, CTAS_, nrows, nnz, rowptr, col, val, x, y, nloc, \
                                                                       xh, xs, vj_out, partial, dots_out, ticket, gate,  \
                                                                       stop)                                             \
             : launch_spmv_bulk_cfg<T, ROWS_, EPT_, NST_, FUSED, false>(s, CTAS_, nrows, nnz, rowptr, col, val, x, y,     \
                                                                        nloc, xh, xs, vj_out, partial, dots_out, ticket, \
                                                                        gate, stop)
  // entries per thread and ring slot: a 256-row block of a matrix with avg entries per row holds ~256*avg (+3 of
  // alignment pad); 6 covers 5-point stencils, 8 covers 7-point stencils
  const double avg = (double)nnz / (double)(nrows > 0 ? nrows : 1);
  if (avg <= 5.9) {
    switch (variant) {
      case 1: AB200_BULK_CFG(128, 6, 3, 6);
      case 2: AB200_BULK_CFG(256, 6, 2, 4);
      case 3: AB200_BULK_CFG(256, 6, 4, 2);
      default: AB200_BULK_CFG(256, 6, 3, 3);
    }
  }
  switch (variant) {
    case 1: AB200_BULK_CFG(128, 8, 3, 4);
    case 2: AB200_BULK_CFG(256, 8, 2, 3);
    case 3: AB200_BULK_CFG(256, 8, 3, 2);
    default:
      if (own) AB200_BULK_CFG(256, 8, 2, 3);
      AB200_BULK_CFG(256, 8, 2, 4);  // parked products: measured best on the 7-point Laplacian (5.25 TB/s alone)
  }
#undef AB200_BULK_CFG
}

template <typename T, bool FUSED>
int launch_spmv_stream(cudaStream_t s, int nrows, const int* rowptr, const int* col, const T* val, const T* x, T* y,
                       int nloc, const T* xh, T xs, T* vj_out, T* partial, T* dots_out, unsigned int* ticket,
                       const StepGate<T>& gate = StepGate<T>(), const T* stop = nullptr) {
  static const int rows_env = getenv("AB200_SPMV_ROWS") ? atoi(getenv("AB200_SPMV_ROWS")) : 256;
  const long long cap = 148LL * 8;
  if (rows_env == 512) {
    constexpr int ROWS = 512, CAP = 9 * 512;
    long long g = ((long long)nrows + ROWS - 1) / ROWS;
    const int grid = (int)(g > cap / 2 ? cap / 2 : g);
    k_csr_spmv_stream<T, ROWS, CAP, FUSED><<<grid, ROWS, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh, xs, vj_out,
                                                               partial, dots_out, ticket, gate, stop);
  } else if (rows_env == 128) {
    constexpr int ROWS = 128, CAP = 9 * 128;
    long long g = ((long long)nrows + ROWS - 1) / ROWS;
    const int grid = (int)(g > cap * 2 ? cap * 2 : g);
    k_csr_spmv_stream<T, ROWS, CAP, FUSED><<<grid, ROWS, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh, xs, vj_out,
                                                               partial, dots_out, ticket, gate, stop);
  } else {
    constexpr int ROWS = 256, CAP = 9 * 256;
    long long g = ((long long)nrows + ROWS - 1) / ROWS;
    const int grid = (int)(g > cap ? cap : g);
    k_csr_spmv_stream<T, ROWS, CAP, FUSED><<<grid, ROWS, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh, xs, vj_out,
                                                               partial, dots_out, ticket, gate, stop);
  }
  launch_stats().kernels++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template <typename T>
int launch_spmv(int nrows, const int* rowptr, const int* col, const T* val, const T* x, T* y, int nloc,
                const T* xh, long long nnz_hint) {
  if (nrows <= 0) return 0;
  const double avg = nnz_hint > 0 ? (double)nnz_hint / nrows : 8.0;
  const int threads = 256;
  cudaStream_t s = cur_stream();
  const long long nnz = nnz_hint > 0 ? nnz_hint : 0;
  ProfScope ps(s, "csr_spmv", (double)nnz * (sizeof(T) + 4.0) + (nrows + 1) * 4.0 + 2.0 * nrows * sizeof(T));
  const bool use_stream = spmv_variant() != 2;
  if (use_stream && avg <= 7.9 && spmv_bulk_ok(rowptr, col, val, nnz))
    return launch_spmv_bulk<T, false>(s, nrows, nnz, rowptr, col, val, x, y, nloc, xh, T(1), nullptr, nullptr, nullptr,
                                      nullptr);
  if (use_stream && avg <= 8.5)
    return launch_spmv_stream<T, false>(s, nrows, rowptr, col, val, x, y, nloc, xh, T(1), nullptr, nullptr, nullptr,
                                        nullptr);
  auto grid_for = [&](int lpr) {
    long long g = ((long long)nrows * lpr + threads - 1) / threads;
    const long long cap = 148LL * 32;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
  };
  if (avg <= 3.0) k_csr_spmv<T, 2><<<grid_for(2), threads, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh);
  else if (avg <= 6.0) k_csr_spmv<T, 4><<<grid_for(4), threads, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh);
  else if (avg <= 12.0) k_csr_spmv<T, 8><<<grid_for(8), threads, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh);
  else if (avg <= 24.0) k_csr_spmv<T, 16><<<grid_for(16), threads, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh);
  else k_csr_spmv<T, 32><<<grid_for(32), threads, 0, s>>>(nrows, rowptr, col, val, x, y, nloc, xh);
  launch_stats().kernels++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// nnz of a CSR matrix whose rowptr lives on the device, cached per (rowptr address, rows).  The value only steers the
// choice of kernel and its tile shape (every kernel reads the true row bounds on the device and handles any row
// length), so a stale entry -- the arrays were freed and others of the same size allocated at the same address --
// costs performance at worst, never correctness.  ab200_release_all() forgets the cache.
struct NnzEntry { const int* ptr; int rows; long long nnz; };
NnzEntry g_nnz_cache[8] = {};
long long nnz_of(int nrows, const int* rowptr) {
  typedef NnzEntry Entry;
  Entry* cache = g_nnz_cache;
  static int next = 0;
  for (int i = 0; i < 8; ++i)
    if (cache[i].ptr == rowptr && cache[i].rows == nrows) return cache[i].nnz;
  int h = 0;
  if (cudaMemcpyAsync(&h, rowptr + nrows, sizeof(int), cudaMemcpyDeviceToHost, cur_stream()) != cudaSuccess) return 0;
  cudaStreamSynchronize(cur_stream());
  cache[next] = Entry{rowptr, nrows, (long long)h};
  next = (next + 1) % 8;
  return h;
}

// ---------------------------------------------------------------------------------------------
// generators (row-major natural ordering, ascending column order inside a row)
// ---------------------------------------------------------------------------------------------
// 2-D 5-point Laplacian scaled by `scale` (dssimp.f:484-540 uses scale = (nx+1)^2; BASELINE config 2: 1)
__global__ void k_gen_laplace2d(int nx, int ny, double scale, int* rowptr, int* col, double* val) {
  const long long n = (long long)nx * ny;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (long long)gridDim.x * blockDim.x) {
    // entries before row r: 5 per row minus the missing neighbours
    const long long iy = r / nx, ix = r % nx;  // row r = (iy, ix)
    // rows 0..r-1: missing left = number of rows with ix==0, etc.
    long long before = 5 * r;
    before -= (iy + (ix > 0 ? 1 : 0));            // rows with ix == 0 among 0..r-1
    before -= (iy);                               // rows with ix == nx-1 among 0..r-1 (complete grid lines only)
    before -= (r < nx ? r : nx);                  // rows with iy == 0
    before -= (iy == ny ? nx : 0) + ((iy == ny - 1) ? ix : 0);  // rows with iy == ny-1
    if (r == n) {
      rowptr[n] = (int)before;
      return;
    }
    rowptr[r] = (int)before;
    long long p = before;
    if (iy > 0) { col[p] = (int)(r - nx); val[p] = -scale; ++p; }
    if (ix > 0) { col[p] = (int)(r - 1); val[p] = -scale; ++p; }
    col[p] = (int)r; val[p] = 4.0 * scale; ++p;
    if (ix < nx - 1) { col[p] = (int)(r + 1); val[p] = -scale; ++p; }
    if (iy < ny - 1) { col[p] = (int)(r + nx); val[p] = -scale; ++p; }
  }
}

// 2-D convection-diffusion of dndrv1.f:453-470 / dnsimp.f:570 on the unit square, h = 1/(nx+1)
__global__ void k_gen_convdiff2d(int nx, double rho, int* rowptr, int* col, double* val) {
  const long long n = (long long)nx * nx;
  const double h = 1.0 / (double)(nx + 1), h2 = h * h;
  const double dd = 4.0 / h2, dl = -1.0 / h2 - 0.5 * rho / h, du = -1.0 / h2 + 0.5 * rho / h, ob = -1.0 / h2;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += (long long)gridDim.x * blockDim.x) {
    const long long iy = r / nx, ix = r % nx;
    long long before = 5 * r;
    before -= (iy + (ix > 0 ? 1 : 0));
    before -= iy;
    before -= (r < nx ? r : nx);
    before -= (iy == nx ? nx : 0) + ((iy == nx - 1) ? ix : 0);
    if (r == n) {
      rowptr[n] = (int)before;
      return;
    }
    rowptr[r] = (int)before;
    long long p = before;
    if (iy > 0) { col[p] = (int)(r - nx); val[p] = ob; ++p; }
    if (ix > 0) { col[p] = (int)(r - 1); val[p] = dl; ++p; }
    col[p] = (int)r; val[p] = dd; ++p;
    if (ix < nx - 1) { col[p] = (int)(r + 1); val[p] = du; ++p; }
    if (iy < nx - 1) { col[p] = (int)(r + nx); val[p] = ob; ++p; }
  }
}

// 3-D 7-point Laplacian (6, -1) on nx x ny x nz, rows of the z-slab [z0, z0+nzloc).  Local column
// numbering of the row-partitioned operator: own rows [0, nloc), then the lower halo plane
// (z0-1) at [nloc, nloc+nx*ny) if z0 > 0, then the upper halo plane (z0+nzloc).
__device__ inline long long lap3d_before(long long r, int nx, int ny, int nzloc, bool has_lo, bool has_hi) {
  const long long plane = (long long)nx * ny;
  const long long iz = r / plane, rem = r % plane, iy = rem / nx, ix = rem % nx;
  long long before = 7 * r;
  // missing x-neighbours: one per grid line end
  const long long lines = iz * ny + iy;  // complete lines before r
  before -= lines + (ix > 0 ? 1 : 0);    // ix == 0
  before -= lines;                       // ix == nx-1
  // missing y-neighbours: rows with iy == 0 / iy == ny-1
  before -= iz * nx + (iy > 0 ? nx : ix);                       // iy == 0
  before -= iz * nx + (iy == ny - 1 ? ix : 0);                  // iy == ny-1
  // missing z-neighbours only at the global boundary (halo planes supply the others)
  if (!has_lo) before -= (iz > 0 ? plane : rem);
  if (!has_hi) before -= (iz == nzloc - 1 ? rem : 0) + (iz >= nzloc ? plane : 0);
  return before;
}
__global__ void k_gen_laplace3d(int nx, int ny, int nz, int z0, int nzloc, double diag, int* rowptr, int* col,
                                double* val) {
  const long long plane = (long long)nx * ny, nloc = plane * nzloc;
  const bool has_lo = z0 > 0, has_hi = (z0 + nzloc) < nz;
  const long long lo_base = nloc, hi_base = nloc + (has_lo ? plane : 0);
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r <= nloc; r += (long long)gridDim.x * blockDim.x) {
    const long long before = lap3d_before(r, nx, ny, nzloc, has_lo, has_hi);
    if (r == nloc) {
      rowptr[nloc] = (int)before;
      return;
    }
    rowptr[r] = (int)before;
    const long long iz = r / plane, rem = r % plane, iy = rem / nx, ix = rem % nx;
    long long p = before;
    if (iz > 0) { col[p] = (int)(r - plane); val[p] = -1.0; ++p; }
    else if (has_lo) { col[p] = (int)(lo_base + rem); val[p] = -1.0; ++p; }
    if (iy > 0) { col[p] = (int)(r - nx); val[p] = -1.0; ++p; }
    if (ix > 0) { col[p] = (int)(r - 1); val[p] = -1.0; ++p; }
    col[p] = (int)r; val[p] = diag; ++p;
    if (ix < nx - 1) { col[p] = (int)(r + 1); val[p] = -1.0; ++p; }
    if (iy < ny - 1) { col[p] = (int)(r + nx); val[p] = -1.0; ++p; }
    if (iz < nzloc - 1) { col[p] = (int)(r + plane); val[p] = -1.0; ++p; }
    else if (has_hi) { col[p] = (int)(hi_base + rem); val[p] = -1.0; ++p; }
  }
}

__device__ inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__global__ void k_fill_hash(long long n, long long i0, unsigned long long seed, double* x) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long h = splitmix64(seed + (unsigned long long)(i0 + i));
    const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    x[i] = 2.0 * u - 1.0;
  }
}

// || A z - d z ||^2 per column, one block-wide deterministic reduction per column (small k)
__global__ void __launch_bounds__(256) k_residual_sq(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                                                     const double* __restrict__ val, const double* __restrict__ z,
                                                     double d, double* __restrict__ partial) {
  __shared__ double red[8];
  double acc = 0.0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int p = rowptr[r]; p < rowptr[r + 1]; ++p) s += val[p] * z[col[p]];
    s -= d * z[r];
    acc += s * s;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

inline int gen_grid(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g > 148LL * 16 ? 148LL * 16 : (g < 1 ? 1 : g));
}

}  // namespace
}  // namespace ab200

namespace ab200 {
// C++ entry points used by api.cu for the registered-operator mode (see driver.hpp)
// halo planes of x for a row-partitioned operator: my first halo_lo entries go down, my last halo_hi go up
// *xh: where the SpMV that follows must read its halo columns from (the communicator's own buffer has two parities)
template <typename T>
int exchange_halo_planes(int comm, int nloc, int halo_lo, int halo_hi, const T* x, T* halo, const T** xh) {
  *xh = halo;
  if (comm == 0 || (halo_lo == 0 && halo_hi == 0)) return 0;
  try {
    NcclComm* c = comm_from_handle(comm);
    if (!c) return -1;
    // my first plane goes down, my last plane goes up; halo = [plane from below | plane from above]
    if (nccl_halo_is_peer_buffer(c, halo))
      *xh = static_cast<const T*>(nccl_halo_exchange_peer(c, x, x + (nloc - halo_hi), cur_stream()));
    else
      nccl_halo_exchange(c, x, halo, (size_t)halo_lo, x + (nloc - halo_hi), halo + halo_lo, (size_t)halo_hi,
                         sizeof(T) == 8, cur_stream());
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: halo exchange: %s\n", e.what());
    return -1;
  }
  return 0;
}
template <typename T>
int exchange_halo(const CsrOpDesc<T>& op, const T* x, const T** xh) {
  return exchange_halo_planes<T>(op.comm, op.nrows, op.halo_lo, op.halo_hi, x, op.halo, xh);
}

template <typename T>
int csr_op_apply(const CsrOpDesc<T>& op, const T* x, T* y) {
  if (op.comm != 0) {
    const T* xh = nullptr;
    if (exchange_halo(op, x, &xh) != 0) return -1;
    return launch_spmv<T>(op.nrows, op.rowptr, op.col, op.val, x, y, op.nrows, xh, op.nnz);
  }
  return launch_spmv<T>(op.nrows, op.rowptr, op.col, op.val, x, y, 0, nullptr, op.nnz);
}
template <typename T>
int csr_op_apply_fused(const CsrOpDesc<T>& op, T inv, const StepGate<T>* gate, const T* stop, const T* resid, T* vj,
                       T* y, T* partial, T* dots_out, unsigned int* ticket) {
  if (op.nrows <= 0) return 0;
  const double avg = op.nnz > 0 ? (double)op.nnz / op.nrows : 8.0;
  if (avg > 7.9) return 1;  // caller falls back to start_step + plain SpMV
  cudaStream_t s = cur_stream();
  ProfScope ps(s, "csr_spmv_fused",
               (double)op.nnz * (sizeof(T) + 4.0) + (op.nrows + 1) * 4.0 + 3.0 * op.nrows * sizeof(T));
  // under a communicator the neighbours' planes of the UNSCALED residual are exchanged: the kernel applies 1/||r|| to
  // the row sums, halo entries included (every rank uses the same global norm); the epilogue dots would be local
  // partial sums there, so they are not produced
  const int nloc = op.comm != 0 ? op.nrows : 0;
  const T* xh = nullptr;
  if (op.comm != 0) {
    if (exchange_halo(op, resid, &xh) != 0) return -1;
    dots_out = nullptr;
  }
  const StepGate<T> g = gate ? *gate : StepGate<T>();
  if (spmv_bulk_ok(op.rowptr, op.col, op.val, op.nnz))
    return launch_spmv_bulk<T, true>(s, op.nrows, op.nnz, op.rowptr, op.col, op.val, resid, y, nloc, xh, inv, vj,
                                     partial, dots_out, ticket, g, stop);
  return launch_spmv_stream<T, true>(s, op.nrows, op.rowptr, op.col, op.val, resid, y, nloc, xh, inv, vj, partial,
                                     dots_out, ticket, g, stop);
}
template int csr_op_apply<double>(const CsrOpDesc<double>&, const double*, double*);
template int csr_op_apply<float>(const CsrOpDesc<float>&, const float*, float*);
template int csr_op_apply_fused<double>(const CsrOpDesc<double>&, double, const StepGate<double>*, const double*,
                                        const double*, double*, double*, double*, double*, unsigned int*);
template int csr_op_apply_fused<float>(const CsrOpDesc<float>&, float, const StepGate<float>*, const float*,
                                       const float*, float*, float*, float*, float*, unsigned int*);
}  // namespace ab200

using namespace ab200;

extern "C" {


void ab200_forget_csr_cache(void) {
  for (auto& e : g_nnz_cache) e = NnzEntry{nullptr, 0, 0};
}
void ab200_set_spmv_variant(int variant) { spmv_variant() = (variant >= 0 && variant <= 4) ? variant : 0; }

int ab200_csr_spmv_f64(int nrows, const int* rowptr, const int* col, const double* val, const double* x, double* y) {
  return launch_spmv<double>(nrows, rowptr, col, val, x, y, 0, nullptr, nnz_of(nrows, rowptr));
}
int ab200_csr_spmv_f32(int nrows, const int* rowptr, const int* col, const float* val, const float* x, float* y) {
  return launch_spmv<float>(nrows, rowptr, col, val, x, y, 0, nullptr, nnz_of(nrows, rowptr));
}

int ab200_csr_spmv_hostvec_f64(int nrows, int ncols, const int* rowptr, const int* col, const double* val,
                               const double* x_host, double* y_host) {
  static double* xd = nullptr;
  static double* yd = nullptr;
  static size_t xcap = 0, ycap = 0;
  cudaStream_t s = cur_stream();
  if ((size_t)ncols > xcap) { cudaFree(xd); if (cudaMalloc(&xd, sizeof(double) * ncols) != cudaSuccess) return -1; xcap = ncols; }
  if ((size_t)nrows > ycap) { cudaFree(yd); if (cudaMalloc(&yd, sizeof(double) * nrows) != cudaSuccess) return -1; ycap = nrows; }
  if (cudaMemcpyAsync(xd, x_host, sizeof(double) * ncols, cudaMemcpyHostToDevice, s) != cudaSuccess) return -2;
  if (launch_spmv<double>(nrows, rowptr, col, val, xd, yd, 0, nullptr, nnz_of(nrows, rowptr)) != 0) return -3;
  if (cudaMemcpyAsync(y_host, yd, sizeof(double) * nrows, cudaMemcpyDeviceToHost, s) != cudaSuccess) return -4;
  return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -5;
}

int ab200_csr_spmv_halo_f64(int comm, int nloc, int halo_lo, int halo_hi, const int* rowptr, const int* col,
                            const double* val, const double* x, double* y, double* halo_buf) {
  const double* xh = halo_buf;
  if (exchange_halo_planes<double>(comm, nloc, halo_lo, halo_hi, x, halo_buf, &xh) != 0) return -1;
  return launch_spmv<double>(nloc, rowptr, col, val, x, y, nloc, xh, nnz_of(nloc, rowptr));
}

long long ab200_gen_laplace2d(int nx, int ny, double scale, int* rowptr, int* col, double* val) {
  const long long n = (long long)nx * ny, nnz = 5 * n - 2LL * nx - 2LL * ny;
  if (nnz > 2147483647LL) return -2;
  if (!rowptr) return nnz;
  k_gen_laplace2d<<<gen_grid(n + 1), 256, 0, cur_stream()>>>(nx, ny, scale, rowptr, col, val);
  launch_stats().kernels++;
  return cudaGetLastError() == cudaSuccess ? nnz : -1;
}
long long ab200_gen_convdiff2d(int nx, double rho, int* rowptr, int* col, double* val) {
  const long long n = (long long)nx * nx, nnz = 5 * n - 4LL * nx;
  if (nnz > 2147483647LL) return -2;
  if (!rowptr) return nnz;
  k_gen_convdiff2d<<<gen_grid(n + 1), 256, 0, cur_stream()>>>(nx, rho, rowptr, col, val);
  launch_stats().kernels++;
  return cudaGetLastError() == cudaSuccess ? nnz : -1;
}
long long ab200_gen_laplace3d(int nx, int ny, int nz, int z0, int nzloc, double diag, int* rowptr, int* col,
                              double* val) {
  const long long plane = (long long)nx * ny, nloc = plane * nzloc;
  const bool has_lo = z0 > 0, has_hi = (z0 + nzloc) < nz;
  long long nnz = 7 * nloc - 2LL * ny * nzloc - 2LL * nx * nzloc;
  if (!has_lo) nnz -= plane;
  if (!has_hi) nnz -= plane;
  if (nnz > 2147483647LL) return -2;
  if (!rowptr) return nnz;
  k_gen_laplace3d<<<gen_grid(nloc + 1), 256, 0, cur_stream()>>>(nx, ny, nz, z0, nzloc, diag, rowptr, col, val);
  launch_stats().kernels++;
  return cudaGetLastError() == cudaSuccess ? nnz : -1;
}
int ab200_fill_hash_f64(long long n, long long i0, unsigned long long seed, double* x) {
  k_fill_hash<<<gen_grid(n), 256, 0, cur_stream()>>>(n, i0, seed, x);
  launch_stats().kernels++;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
int ab200_residuals_f64(int n, const int* rowptr, const int* col, const double* val, int k, const double* z,
                        long long ldz, const double* d_host, double* out_host) {
  const int grid = gen_grid(n);
  double* partial = nullptr;
  if (cudaMalloc(&partial, sizeof(double) * grid) != cudaSuccess) return -1;
  std::vector<double> h((size_t)grid);
  cudaStream_t s = cur_stream();
  for (int c = 0; c < k; ++c) {
    k_residual_sq<<<grid, 256, 0, s>>>(n, rowptr, col, val, z + (size_t)c * ldz, d_host[c], partial);
    launch_stats().kernels++;
    cudaMemcpyAsync(h.data(), partial, sizeof(double) * grid, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    double acc = 0.0;
    for (int b = 0; b < grid; ++b) acc += h[b];
    out_host[c] = sqrt(acc);
  }
  cudaFree(partial);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
}
