// vq_mma.cu -- out(:,0:kout) = V(:,0:kin) * Q on the FP64 tensor cores (DMMA), FP64 only.
//
// Replaces K12-K16 of the restart (SRC/dsapps.f:450-493: kev x (dgemv + dcopy), dscal + daxpy; SRC/dnapps.f the same
// with dgemv columns) and K20 (dseupd.f:742 dorm2r): ONE pass over V, in place.
//
// Why tensor cores here (north_star: "only if ncu shows this update compute-bound"): at ncv = 64, kout ~ 30 the update
// needs 2*kin*kout flops per 8*(kin+kout) bytes = 5 flop/B, i.e. 33 TFLOP/s at the HBM roofline -- 90 % of the FP64
// peak of a B200 -- and the SIMT kernel it replaces (k_vq_tma, scalar DFMA fed by one shared-memory load per 2.7 FMAs)
// ran at 10 TFLOP/s = 0.30 of the HBM roofline, bound by issue slots and shared-memory wavefronts, not by DRAM
// (profiles/README.md).  tcgen05 has no FP64 kind: mma.sync.m8n8k4.f64 (DMMA) is Blackwell's FP64 tensor-core path.
// One DMMA = 256 FMAs for two 8-byte shared-memory loads per lane; six loads feed eight DMMAs below.
//
// Layout: a CTA per SM walks tiles of R = 128 rows.  A producer thread streams the tile column by column with 1-D bulk
// copies (cp.async.bulk, 1 KB each, mbarrier complete_tx) into a ring of stages whose COLUMN STRIDE IS R + 4 DOUBLES:
// the A fragment of a DMMA is a[row = lane/4][k = lane%4], so a half-warp reads 4 rows x 4 columns, and with a stride
// = 4 (mod 16) doubles those 16 addresses fall into 16 different 8-byte bank pairs (a dense stride of 128 would make
// it a 4-way conflict).  A 2-D tensor-map box cannot be padded like that (128-byte destination alignment), hence the
// bulk copies; the last, partial tile of the matrix is copied by the producer warp with ordinary loads.
// Two groups of eight warps take alternate tiles; warp w of a group owns rows 16w..16w+15 (two 8-row blocks) and all
// n-blocks of 8 output columns: per k-step of 4 it loads 2 A fragments and NB B fragments and issues 2*NB DMMAs.
// Q sits in shared memory with row stride kp + 4 (same bank argument for b[k = lane%4][n = lane/4]).
// Structural zeros of Q are skipped: after dsapps' QR sweeps column c of Q is zero below row np + c
// (dsapps.f:461 multiplies only kplusp-i+1 entries); the host passes, per n-block, how many k-steps hold non-zeros.
#include <cstdlib>
#include <cstring>

#include "tma_common.cuh"
#include "vecops_cuda.cuh"

namespace ab200 {

namespace {

using namespace ab200::tma;

constexpr int R = 128;             // rows per tile
constexpr int RS = R + 4;          // shared-memory column stride of a tile (doubles)
constexpr int NCW = 8;             // warps per consumer group
constexpr int NG = 2;              // consumer groups
constexpr int QN = 4;              // a stage is filled and consumed in QN column groups: own mbarrier, own producer warp
constexpr int kThreads = (NG * NCW + QN) * 32;
constexpr int MAXK = 64;           // kin, kout <= 64
constexpr int MAXST = 8;
constexpr uint32_t kSmemMax = 226 * 1024;   // dynamic shared memory of one CTA on sm_100 (227 KB) minus the static part

struct VqMmaParams {
  int64_t n, ldv, ldo;
  int kin, kin4, kout, kp, qs;     // kin4 = kin rounded up to 4; kp = kout rounded up to 8; qs = kp + 4
  int nstages;
  uint32_t q_elems;                // doubles reserved for Q in shared memory (kin4 * qs, rounded up)
  int qsteps;                      // k-steps (4 columns each) per column group of a stage
  uint32_t stage_elems;            // kin4 * RS
  const double* v;
  const double* q;                 // device, packed column-major kin x kout
  double* out;
  int with_resid, beta_col;
  double sigma, beta;
  double* resid;
  double* partial;
  double* nrm2_out;
  unsigned int* ticket;
  unsigned char ksteps[8];         // per n-block: k-steps (of 4 rows of Q) that hold non-zeros
};

__device__ __forceinline__ void bulk_col(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

template <int NB>
__global__ void __launch_bounds__(kThreads, 1) k_vq_mma(const VqMmaParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* tiles = reinterpret_cast<double*>(smem);
  double* qsm = tiles + (size_t)p.nstages * p.stage_elems;               // [kin4][qs]
  double* red = qsm + p.q_elems;                                          // [NG*NCW]
  // full[s][q]: column group q of stage s has landed -- the consumers start their DMMAs on the first columns while
  // the later ones are still in flight, so that loading and computing overlap inside a stage as well as across stages
  // Every (stage, consumer group) pair has its OWN set of full barriers: with a ring depth that is not a multiple of
  // NG a stage serves the groups alternately, and a group that met a barrier only every other phase could take the
  // completion of phase k-2 for that of phase k (mbarrier waits see one parity bit).  With a barrier per pair each
  // group observes every phase of the barriers it waits on; tile number `it` of the CTA uses stage it % nstages,
  // group it % NG and phase it / lcm(nstages, NG) of that pair.  The empty barrier of a stage is shared: its only
  // waiter, the producer, sees all of its phases.
  uint64_t* full = reinterpret_cast<uint64_t*>(red + NG * NCW + 8);
  uint64_t* empty = full + MAXST * NG * QN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < p.nstages; ++s) {
      for (int q = 0; q < NG * QN; ++q) mbar_init(full + s * NG * QN + q, 1);
      mbar_init(empty + s, NCW);
    }
    mbar_fence_init();
  }
  const int period = (p.nstages % NG == 0) ? p.nstages : p.nstages * NG;   // lcm(nstages, NG) for NG = 2
  // Q^T-free layout: qsm[k][c], zero-padded to kin4 x kp
  for (int i = tid; i < p.kin4 * p.qs; i += kThreads) {
    const int k = i / p.qs, c = i - k * p.qs;
    qsm[i] = (k < p.kin && c < p.kout) ? p.q[(size_t)c * p.kin + k] : 0.0;
  }
  // the padding columns kin..kin4 of every stage are never written by the copies: zero them once
  for (int s = 0; s < p.nstages; ++s)
    for (int i = tid; i < (p.kin4 - p.kin) * RS; i += kThreads) tiles[(size_t)s * p.stage_elems + (size_t)p.kin * RS + i] = 0.0;
  if (tid < NG * NCW + 8) red[tid] = 0.0;
  __syncthreads();
  const int64_t ntiles = (p.n + R - 1) / R;
  double nrm = 0.0;
  if (warp >= NG * NCW) {
    // ---------------- producer warps: warp q feeds column group q of every stage ----------------
    // (one thread issuing all 1 KB copies of a tile sustains ~20 GB/s per SM -- the copies are cheap but each is an
    // instruction; QN issuers lift that ceiling above the SM's share of the HBM bandwidth)
    const int q = warp - NG * NCW;
    const int c0 = 4 * p.qsteps * q;
    int c1 = c0 + 4 * p.qsteps;
    c1 = c1 > p.kin ? p.kin : c1;
    int s = 0, it = 0;
    uint32_t ph = 0;   // phase of the stage's empty barrier
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int64_t row0 = t * R;
      const int rows = (int)((p.n - row0 < R) ? (p.n - row0) : R);
      double* dst = tiles + (size_t)s * p.stage_elems;
      uint64_t* bar = full + (s * NG + it % NG) * QN + q;
      if (lane == 0) mbar_wait(empty + s, ph ^ 1u);
      __syncwarp();
      if (c1 <= c0) {
        if (lane == 0) mbar_arrive(bar);   // nothing to load for this group (kin small): complete its phase all the same
      } else if (rows == R) {
        if (lane == 0) {
          mbar_expect_tx(bar, (uint32_t)(c1 - c0) * (uint32_t)(R * sizeof(double)));
          for (int c = c0; c < c1; ++c)
            bulk_col(smem_u32(dst + (size_t)c * RS), p.v + row0 + (int64_t)c * p.ldv, (uint32_t)(R * sizeof(double)),
                     bar);
        }
      } else {
        // last, partial tile: ordinary loads, zero fill (bulk copies need multiples of 16 bytes)
        for (int i = lane; i < (c1 - c0) * R; i += 32) {
          const int c = c0 + i / R, r = i % R;
          dst[(size_t)c * RS + r] = (r < rows) ? p.v[row0 + r + (int64_t)c * p.ldv] : 0.0;
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar);
      }
      if (++s == p.nstages) { s = 0; ph ^= 1u; }
    }
  } else {
    // ---------------- consumers ----------------
    const int g = warp / NCW, gw = warp - g * NCW;
    const int qr = lane >> 2, qc = lane & 3;   // DMMA fragment coordinates
    const int nks = p.kin4 >> 2;
    int it = g;   // CTA-local tile counter of my tiles: g, g + NG, ...
    for (int64_t t = (int64_t)blockIdx.x + (int64_t)g * gridDim.x; t < ntiles; t += (int64_t)NG * gridDim.x, it += NG) {
      const int s = it % p.nstages;
      const uint32_t ph = (uint32_t)(it / period) & 1u;
      uint64_t* fullg = full + (s * NG + g) * QN;
      const int64_t row0 = t * R;
      const double* tile = tiles + (size_t)s * p.stage_elems;
      const double* ap = tile + (size_t)qc * RS + 16 * gw + qr;   // a[row = qr][k = qc] of k-step 0, m-block 0
      const double* bp = qsm + (size_t)qc * p.qs + qr;            // b[k = qc][n = qr] of k-step 0, n-block 0
      double acc[2][NB][2];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) acc[m][nb][0] = acc[m][nb][1] = 0.0;
      for (int q = 0; q < QN; ++q) {
        mbar_wait(fullg + q, ph);
        const int k0 = q * p.qsteps;
        int k1 = k0 + p.qsteps;
        k1 = k1 > nks ? nks : k1;
#pragma unroll 4
        for (int ks = k0; ks < k1; ++ks) {
          const double a0 = ap[(size_t)(4 * ks) * RS], a1 = ap[(size_t)(4 * ks) * RS + 8];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            if (ks < (int)p.ksteps[nb]) {   // uniform: structural zeros of Q below the band are skipped
              const double b = bp[(size_t)(4 * ks) * p.qs + 8 * nb];
              dmma(acc[0][nb], a0, b);
              dmma(acc[1][nb], a1, b);
            }
          }
        }
      }
      // the tile is in registers now: hand the stage back before the stores
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      // c[row = qr][col = 2*qc + {0,1}]
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int64_t r = row0 + 16 * gw + 8 * m + qr;
        if (r < p.n) {
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = 8 * nb + 2 * qc + e;
              if (col < p.kout) {
                p.out[r + (int64_t)col * p.ldo] = acc[m][nb][e];
                if (p.with_resid && col == p.beta_col) {
                  const double v = p.sigma * p.resid[r] + p.beta * acc[m][nb][e];
                  p.resid[r] = v;
                  nrm += v * v;
                }
              }
            }
          }
          if (p.with_resid && p.beta_col < 0 && qc == 0) {
            const double v = p.sigma * p.resid[r];
            p.resid[r] = v;
            nrm += v * v;
          }
        }
      }
    }
    nrm = warp_sum(nrm);
    if (lane == 0) red[warp] = nrm;
  }
  if (!p.with_resid || p.nrm2_out == nullptr) return;
  __syncthreads();
  if (tid == 0) {
    double sum = 0.0;
#pragma unroll
    for (int w = 0; w < NG * NCW; ++w) sum += red[w];
    p.partial[blockIdx.x] = sum;
  }
  finish_grid_reduce(p.partial, 1, 1, p.nrm2_out, p.ticket);
}

size_t aux_bytes() { return sizeof(double) * (NG * NCW + 8) + sizeof(uint64_t) * (MAXST * NG * QN + MAXST); }

template <int NB>
cudaError_t launch(int grid, size_t smem, cudaStream_t s, const VqMmaParams& p) {
  static bool attr_set[kMaxDevices] = {};   // (function, device) attribute
  const int dev = current_device_slot();
  if (!attr_set[dev]) {
    const cudaError_t e = cudaFuncSetAttribute(k_vq_mma<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  k_vq_mma<NB><<<grid, kThreads, smem, s>>>(p);
  return cudaGetLastError();
}

}  // namespace

// q_host is needed next to the device copy for the zero-structure scan (kin x kout, column-major, leading dim ldq)
template <>
bool CudaVecOps<double>::vq_mma(int64_t n, int kin, int kout, const double* v, int64_t ldv, const double* qdev,
                                const double* q_host, int ldq, double* out, int64_t ldo, bool with_resid, double sigma,
                                double beta, int beta_col, double* resid, double* mb_nrm2) {
  static const bool off = getenv("AB200_VQ") && std::strcmp(getenv("AB200_VQ"), "simt") == 0;
  if (off || kin < 1 || kin > MAXK || kout < 1 || kout > MAXK) return false;
  if (!fast_path_ok(n, kin, v, ldv)) return false;   // 16-byte aligned base and column stride
  VqMmaParams p{};
  p.n = n; p.ldv = ldv; p.ldo = ldo; p.kin = kin; p.kin4 = (kin + 3) & ~3; p.kout = kout; p.kp = (kout + 7) & ~7;
  p.qs = p.kp + 4;
  p.stage_elems = (uint32_t)p.kin4 * RS;
  p.qsteps = ((p.kin4 >> 2) + QN - 1) / QN;
  p.q_elems = ((uint32_t)(p.kin4 * p.qs) + 15u) & ~15u;
  // small updates stay with the SIMT kernel (measured: kin*kout = 240 -> 5.5 vs 4.9 TB/s, 100 -> 4.7 vs 3.7)
  static const int mma_min = getenv("AB200_VQ_MMA_MIN") ? atoi(getenv("AB200_VQ_MMA_MIN")) : 300;
  if (kin * kout < mma_min) return false;
  const size_t fixed = (size_t)p.q_elems * sizeof(double) + aux_bytes() + 128;
  int ns = (int)((kSmemMax - fixed) / (p.stage_elems * sizeof(double)));
  ns = ns > MAXST ? MAXST : ns;
  if (ns < 2) return false;
  p.nstages = ns;
  p.v = v; p.q = qdev; p.out = out;
  p.with_resid = with_resid ? 1 : 0; p.beta_col = beta_col; p.sigma = sigma; p.beta = beta; p.resid = resid;
  p.nrm2_out = mb_nrm2; p.ticket = ticket_;
  const int nb = p.kp / 8;
  for (int b = 0; b < 8; ++b) {
    int last = -1;   // last row of Q with a non-zero in columns 8b .. 8b+7
    if (b < nb)
      for (int c = 8 * b; c < 8 * b + 8 && c < kout; ++c)
        for (int k = kin - 1; k > last; --k)
          if (q_host[(size_t)c * ldq + k] != 0.0) { last = k; break; }
    p.ksteps[b] = (unsigned char)((last + 4) / 4);
  }
  const int64_t ntiles = (n + R - 1) / R;
  const int grid = (int)(ntiles < num_sms_ ? ntiles : num_sms_);
  ensure_partial((size_t)grid);
  p.partial = partial_;
  const size_t smem = (size_t)p.nstages * p.stage_elems * sizeof(double) + fixed;
  ProfScope ps(stream_, "vq_mma", (double)sizeof(double) * n * (kin + kout + (with_resid ? 2.0 : 0.0)));
  cudaError_t e;
  switch (nb) {
    case 1: e = launch<1>(grid, smem, stream_, p); break;
    case 2: e = launch<2>(grid, smem, stream_, p); break;
    case 3: e = launch<3>(grid, smem, stream_, p); break;
    case 4: e = launch<4>(grid, smem, stream_, p); break;
    case 5: e = launch<5>(grid, smem, stream_, p); break;
    case 6: e = launch<6>(grid, smem, stream_, p); break;
    case 7: e = launch<7>(grid, smem, stream_, p); break;
    default: e = launch<8>(grid, smem, stream_, p); break;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  launch_stats().kernels++;
  launch_stats().fast_path++;
  return true;
}
template <>
bool CudaVecOps<float>::vq_mma(int64_t, int, int, const float*, int64_t, const float*, const float*, int, float*, int64_t,
                               bool, float, float, int, float*, float*) {
  return false;   // FP32 keeps the SIMT kernel (k_vq_tma): no FP32 shape of mma.sync keeps full FP32 precision
}

}  // namespace ab200
