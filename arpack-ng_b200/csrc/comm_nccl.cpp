// comm_nccl.cpp -- NCCL communicator handles standing in for PARPACK's MPI communicator, and the peer-memory
// (CUDA IPC over NVLink) fast paths for the latency-bound exchanges of a Lanczos/Arnoldi step.
//
// The p*_c entry points keep the reference signature (ICB/parpack.h:20-27): their first argument is
// an MPI_Fint.  Here that integer is a handle into a table of NCCL communicators created with
// ab200_comm_create().  NCCL is bound at run time with dlopen so that the library (a) uses the very
// libnccl.so.2 a host process such as PyTorch has already loaded and (b) still loads on a box
// without NCCL for single-GPU use.
//
// What replaces what (reference: PARPACK/SRC/MPI):
//   * MPI_ALLREDUCE of the CGS/DGKS coefficients and norms (pdsaitr.f:604,720, pdnaitr.f:592,699, pdnorm2.f:72-80):
//       - fused into the producing and consuming kernels (peer.cuh, vecops_tma.cu): no launch of its own;
//       - k_p2p_allreduce: one small kernel, for the reductions outside the fused step (start vector, restart norm,
//         generic kernels);
//       - ncclAllReduce when peer memory is not available (AB200_P2P=0, IPC refused, more than 8 ranks).
//   * the neighbour exchange of the row-partitioned SpMV (PARPACK/EXAMPLES/MPI/pdsdrv1.f:463-483, MPI_SEND/MPI_RECV
//     of one grid plane): k_halo_exchange stores the planes into the neighbours' halo buffers over NVLink and waits
//     for theirs; grouped ncclSend/ncclRecv otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "peer.cuh"
#include "vecops_cuda.cuh"

namespace ab200 {

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclSum = 0, kNcclFloat32 = 7, kNcclFloat64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& api() {
  static NcclApi a;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      a.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) {
      a.error = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
      return;
    }
#define AB200_SYM(field, name)                                        \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, name)); \
  if (!a.field) a.error = std::string("missing NCCL symbol ") + name;
    AB200_SYM(GetUniqueId, "ncclGetUniqueId")
    AB200_SYM(CommInitRank, "ncclCommInitRank")
    AB200_SYM(CommDestroy, "ncclCommDestroy")
    AB200_SYM(AllReduce, "ncclAllReduce")
    AB200_SYM(AllGather, "ncclAllGather")
    AB200_SYM(ReduceScatter, "ncclReduceScatter")
    AB200_SYM(Send, "ncclSend")
    AB200_SYM(Recv, "ncclRecv")
    AB200_SYM(GroupStart, "ncclGroupStart")
    AB200_SYM(GroupEnd, "ncclGroupEnd")
    AB200_SYM(GetErrorString, "ncclGetErrorString")
#undef AB200_SYM
  });
  return a;
}

void check(ncclResult_t r, const char* what) {
  if (r != 0) {
    const char* s = api().GetErrorString ? api().GetErrorString(r) : "?";
    throw CudaError(std::string(what) + " failed: " + s);
  }
}

// in-kernel waits give up (trap -> a CUDA error on the host) after this many SM cycles; 0 = wait for ever.
// Ranks of a reverse-communication solver legitimately drift apart between calls (a slow user OP, I/O on one rank),
// so the default is long: 300 s.  AB200_P2P_TIMEOUT_S overrides it (0 = unbounded, like an MPI collective).
long long wait_budget_cycles() {
  static const long long v = [] {
    double s = 300.0;
    if (const char* e = getenv("AB200_P2P_TIMEOUT_S")) s = atof(e);
    return s <= 0.0 ? 0LL : (long long)(s * 1.9e9);
  }();
  return v;
}
}  // namespace

struct P2pPeers {
  unsigned char* base[kPeerMaxRanks];
};

struct NcclComm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  bool destroyed = false;
  // peer-memory reductions
  bool p2p = false;
  unsigned char* p2p_local = nullptr;
  P2pPeers peers{};
  unsigned long long seq = 0;                    // stand-alone all-reduce
  unsigned long long fseq[kPeerKinds] = {0, 0, 0};  // fused reductions, one counter per kind
  // peer-memory halo exchange: my receive buffer (two parities of [from below | from above] + flags), the two
  // neighbours' buffers mapped here, and everybody's plane sizes
  unsigned char* halo_local = nullptr;
  unsigned char* halo_below = nullptr;  // rank-1's buffer
  unsigned char* halo_above = nullptr;  // rank+1's buffer
  size_t halo_lo = 0, halo_hi = 0, halo_stride = 0;  // my plane sizes (elements) and the bytes of one parity
  size_t below_lo = 0, below_stride = 0, above_stride = 0;
  int halo_es = 8;
  unsigned long long halo_seq = 0;
  unsigned int* halo_ticket = nullptr;
};

// ---- stand-alone all-reduce of a few values through peer memory --------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kPeerMaxCount) k_p2p_allreduce(T* mb, int count, int rank, int nranks, P2pPeers peers,
                                                                 unsigned long long seq, long long budget) {
  const int tid = threadIdx.x;
  const int par = (int)(seq & 1ull);
  if (tid < count) {
    const T v = mb[tid];
    for (int p = 0; p < nranks; ++p) {
      T* slot = reinterpret_cast<T*>(peers.base[p]) + ((size_t)par * kPeerMaxRanks + rank) * kPeerMaxCount + tid;
      *reinterpret_cast<volatile T*>(slot) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < nranks) {
    unsigned long long* f =
        reinterpret_cast<unsigned long long*>(peers.base[tid] + kPeerAloneData + (size_t)rank * kPeerFlagStride);
    peer_store_release(f, seq);
    // wait for rank tid's contribution to land in MY buffer
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(peers.base[rank] + kPeerAloneData + (size_t)tid * kPeerFlagStride);
    long long t0 = 0;
    unsigned int spins = 0;
    while (peer_load_acquire(mine) < seq) {
      if (budget > 0 && ((++spins & 0x3FFu) == 0)) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > budget) __trap();
      }
    }
  }
  __syncthreads();
  if (tid < count) {
    const T* slots = reinterpret_cast<const T*>(peers.base[rank]) + (size_t)par * kPeerMaxRanks * kPeerMaxCount + tid;
    T sum = T(0);
    for (int p = 0; p < nranks; ++p) sum += *reinterpret_cast<const volatile T*>(slots + (size_t)p * kPeerMaxCount);
    mb[tid] = sum;
  }
}

// end of a fused reduction whose consumer is not a kernel of ours (the host is about to read the value): add the
// ranks' slots into log[0..count).  With a predicate the reduction only exists when the DGKS pass ran (dsaitr.f:656).
// A sweep whose stop flag is up produced nothing after the step that tripped it (the producers exited at entry, the
// mailbox slots of those steps still hold values of an earlier sweep): nothing to wait for then.
template <typename T>
__global__ void __launch_bounds__(kPeerMaxCount) k_peer_finalize(PeerReduce pr, int count, T* log, const T* pred_w2,
                                                                 const T* pred_r2, const T* stop) {
  if (stop != nullptr && *reinterpret_cast<const volatile T*>(stop) != T(0)) return;
  if (pred_w2 != nullptr) {
    const T wn = sqrt(*pred_w2), rn = sqrt(*pred_r2);
    if (rn > T(0.717f) * wn) return;
  }
  peer_wait_all(pr);
  if ((int)threadIdx.x < count) log[threadIdx.x] = peer_sum<T>(pr, (int)threadIdx.x);
}

// ---- neighbour exchange of the halo planes through peer memory ----------------------------------------------------
// Every block copies a share of both planes into the neighbours' buffers; the block that takes the last ticket
// publishes the sequence number to both neighbours and waits for theirs, so that the kernel -- and with it the stream
// -- only moves on when this rank's halo buffer is complete.
template <typename T>
__global__ void __launch_bounds__(256) k_halo_exchange(const T* send_lo, size_t n_lo, T* dst_below, const T* send_hi,
                                                       size_t n_hi, T* dst_above, unsigned long long* flag_below,
                                                       unsigned long long* flag_above,
                                                       const unsigned long long* my_flag_from_below,
                                                       const unsigned long long* my_flag_from_above,
                                                       unsigned long long seq, unsigned int* ticket, long long budget) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (dst_below != nullptr)
    for (size_t i = t; i < n_lo; i += stride) dst_below[i] = send_lo[i];
  if (dst_above != nullptr)
    for (size_t i = t; i < n_hi; i += stride) dst_above[i] = send_hi[i];
  __shared__ bool s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) {
    *ticket = 0u;
    if (flag_below != nullptr) peer_store_release(flag_below, seq);
    if (flag_above != nullptr) peer_store_release(flag_above, seq);
  }
  if (threadIdx.x < 2) {
    const unsigned long long* f = threadIdx.x == 0 ? my_flag_from_below : my_flag_from_above;
    if (f != nullptr) {
      long long t0 = 0;
      unsigned int spins = 0;
      while (peer_load_acquire(f) < seq) {
        if (budget > 0 && ((++spins & 0x3FFu) == 0)) {
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          else if (now - t0 > budget) __trap();
        }
      }
    }
  }
}

namespace {
std::mutex g_mu;
std::vector<NcclComm*> g_comms;  // handle = index + 1
}  // namespace

NcclComm* comm_from_handle(int handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (handle < 1 || handle > (int)g_comms.size()) return nullptr;
  NcclComm* c = g_comms[handle - 1];
  return (c == nullptr || c->destroyed) ? nullptr : c;
}

// returns false when the peer-memory path does not apply (the caller then uses NCCL); a launch failure throws: the
// peers are already committed to this path for this sequence number, silently switching collectives would hang them
bool p2p_allreduce_sum(NcclComm* c, void* buf, size_t count, bool is_double, cudaStream_t s) {
  if (!c->p2p || count > (size_t)kPeerMaxCount) return false;
  const unsigned long long seq = ++c->seq;
  if (is_double)
    k_p2p_allreduce<double><<<1, kPeerMaxCount, 0, s>>>((double*)buf, (int)count, c->rank, c->nranks, c->peers, seq,
                                                        wait_budget_cycles());
  else
    k_p2p_allreduce<float><<<1, kPeerMaxCount, 0, s>>>((float*)buf, (int)count, c->rank, c->nranks, c->peers, seq,
                                                       wait_budget_cycles());
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw CudaError(std::string("peer-memory all-reduce launch failed: ") + cudaGetErrorString(e));
  launch_stats().kernels++;
  return true;
}

// A fused reduction of the given kind starts: hand out the buffers and the next sequence number (every rank calls
// this in the same order, so the numbers agree).  false: no peer memory -> reduce with nccl_allreduce_sum instead.
bool nccl_peer_reduce_begin(NcclComm* c, int kind, PeerReduce* out) {
  static const bool off = getenv("AB200_FUSED_REDUCE") && std::strcmp(getenv("AB200_FUSED_REDUCE"), "0") == 0;
  if (c == nullptr || !c->p2p || off || kind < 0 || kind >= kPeerKinds) return false;
  for (int p = 0; p < kPeerMaxRanks; ++p) out->base[p] = p < c->nranks ? c->peers.base[p] : nullptr;
  out->rank = c->rank;
  out->nranks = c->nranks;
  out->kind = kind;
  out->seq = ++c->fseq[kind];
  out->timeout_cycles = wait_budget_cycles();
  return true;
}
void nccl_peer_reduce_finalize(NcclComm* c, const PeerReduce& pr, int count, void* log, const void* pred_w2,
                               const void* pred_r2, const void* stop, bool is_double, cudaStream_t s) {
  (void)c;
  if (is_double)
    k_peer_finalize<double><<<1, kPeerMaxCount, 0, s>>>(pr, count, (double*)log, (const double*)pred_w2,
                                                        (const double*)pred_r2, (const double*)stop);
  else
    k_peer_finalize<float><<<1, kPeerMaxCount, 0, s>>>(pr, count, (float*)log, (const float*)pred_w2,
                                                       (const float*)pred_r2, (const float*)stop);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw CudaError(std::string("peer-reduce finalize launch failed: ") + cudaGetErrorString(e));
  launch_stats().kernels++;
}

namespace {
// gather `item` bytes from every rank (host memory in, host memory out) with the NCCL communicator
void host_allgather(NcclComm* c, const void* mine, void* all, size_t item) {
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  if (cudaMalloc(&d_send, item) != cudaSuccess || cudaMalloc(&d_recv, item * c->nranks) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(d_send);
    cudaFree(d_recv);
    throw CudaError("peer-memory setup: cudaMalloc of the exchange buffers failed");
  }
  cudaMemcpy(d_send, mine, item, cudaMemcpyHostToDevice);
  check(api().AllGather(d_send, d_recv, item, /*ncclChar*/ 0, c->comm, 0), "ncclAllGather(setup)");
  cudaStreamSynchronize(0);
  cudaMemcpy(all, d_recv, item * c->nranks, cudaMemcpyDeviceToHost);
  cudaFree(d_send);
  cudaFree(d_recv);
}

// collective: map every rank's buffer into every other rank (CUDA IPC); all ranks agree on the outcome
void p2p_setup(NcclComm* c) {
  const char* e = getenv("AB200_P2P");
  if ((e && std::strcmp(e, "0") == 0) || c->nranks < 2 || c->nranks > kPeerMaxRanks) return;
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&c->p2p_local, kPeerBytes) != cudaSuccess || cudaMemset(c->p2p_local, 0, kPeerBytes) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine, c->p2p_local) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  cudaDeviceSynchronize();
  const size_t item = sizeof(cudaIpcMemHandle_t) + 8;
  std::vector<unsigned char> h_send(item, 0), h_recv(item * c->nranks, 0);
  std::memcpy(h_send.data(), &mine, sizeof(mine));
  h_send[sizeof(mine)] = (unsigned char)ok;
  host_allgather(c, h_send.data(), h_recv.data(), item);
  for (int p = 0; p < c->nranks; ++p) ok = ok && h_recv[p * item + sizeof(mine)] == 1;
  std::vector<void*> opened;
  if (ok) {
    for (int p = 0; p < c->nranks; ++p) {
      if (p == c->rank) { c->peers.base[p] = c->p2p_local; continue; }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, h_recv.data() + p * item, sizeof(h));
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
        break;
      }
      c->peers.base[p] = (unsigned char*)ptr;
      opened.push_back(ptr);
    }
  }
  // second round: everybody must have mapped everybody, otherwise nobody uses the path
  unsigned char st = (unsigned char)ok;
  std::vector<unsigned char> all((size_t)c->nranks, 0);
  host_allgather(c, &st, all.data(), 1);
  for (int p = 0; p < c->nranks; ++p) ok = ok && all[p] == 1;
  c->p2p = ok != 0;
  if (!c->p2p) {
    // a partially mapped set is of no use: close what was opened, free the local buffer
    for (void* ptr : opened) cudaIpcCloseMemHandle(ptr);
    for (int p = 0; p < kPeerMaxRanks; ++p) c->peers.base[p] = nullptr;
    if (c->p2p_local) cudaFree(c->p2p_local);
    c->p2p_local = nullptr;
    if (getenv("AB200_DEBUG"))
      std::fprintf(stderr, "arpack_b200: rank %d: peer-memory reductions unavailable, using NCCL\n", c->rank);
  }
}

void halo_teardown(NcclComm* c) {
  if (c->halo_below) cudaIpcCloseMemHandle(c->halo_below);
  if (c->halo_above) cudaIpcCloseMemHandle(c->halo_above);
  c->halo_below = c->halo_above = nullptr;
  if (c->halo_local) cudaFree(c->halo_local);
  c->halo_local = nullptr;
  if (c->halo_ticket) cudaFree(c->halo_ticket);
  c->halo_ticket = nullptr;
}

void p2p_teardown(NcclComm* c) {
  halo_teardown(c);
  if (c->p2p)
    for (int p = 0; p < c->nranks; ++p)
      if (p != c->rank && c->peers.base[p]) cudaIpcCloseMemHandle(c->peers.base[p]);
  c->p2p = false;
  if (c->p2p_local) cudaFree(c->p2p_local);
  c->p2p_local = nullptr;
}

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
}  // namespace

// Collective.  This rank's halo receive buffer for planes of halo_lo (from the rank below) and halo_hi (from the rank
// above) elements: returns the device address the row-partitioned SpMV reads its halo columns from, laid out
// [from below | from above] like the buffer of ab200_csr_spmv_halo_f64.  The buffer belongs to the communicator (two
// parities and the arrival flags live behind the returned address' first copy); nullptr when peer memory is not
// available -- the caller then allocates an ordinary buffer and the exchange uses ncclSend/ncclRecv.
void* nccl_halo_alloc(NcclComm* c, size_t halo_lo, size_t halo_hi, int elem_size) {
  if (c == nullptr || !c->p2p) return nullptr;
  // same geometry as before (the usual case: one operator, many solves): the buffer is kept, and no collective runs
  if (c->halo_local != nullptr && c->halo_lo == halo_lo && c->halo_hi == halo_hi && c->halo_es == elem_size)
    return c->halo_local;
  halo_teardown(c);
  c->halo_lo = halo_lo; c->halo_hi = halo_hi; c->halo_es = elem_size;
  c->halo_stride = round_up((halo_lo + halo_hi) * (size_t)elem_size, 256) + 256;  // data, then two 128-byte flags
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&c->halo_local, 2 * c->halo_stride) != cudaSuccess ||
      cudaMemset(c->halo_local, 0, 2 * c->halo_stride) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine, c->halo_local) != cudaSuccess || cudaMalloc(&c->halo_ticket, 4) != cudaSuccess ||
      cudaMemset(c->halo_ticket, 0, 4) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  cudaDeviceSynchronize();
  struct Item { cudaIpcMemHandle_t h; unsigned long long lo, hi, stride; int ok; int pad; };
  Item me{};
  me.h = mine; me.lo = halo_lo; me.hi = halo_hi; me.stride = c->halo_stride; me.ok = ok;
  std::vector<Item> all((size_t)c->nranks);
  host_allgather(c, &me, all.data(), sizeof(Item));
  for (int p = 0; p < c->nranks; ++p) ok = ok && all[p].ok == 1;
  // the planes must match across each rank boundary
  if (ok && c->rank > 0 && all[c->rank - 1].hi != halo_lo) ok = 0;
  if (ok && c->rank < c->nranks - 1 && all[c->rank + 1].lo != halo_hi) ok = 0;
  if (ok && c->rank > 0 && halo_lo > 0) {
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[c->rank - 1].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    c->halo_below = (unsigned char*)ptr;
    c->below_lo = (size_t)all[c->rank - 1].lo;
    c->below_stride = (size_t)all[c->rank - 1].stride;
  }
  if (ok && c->rank < c->nranks - 1 && halo_hi > 0) {
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[c->rank + 1].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    c->halo_above = (unsigned char*)ptr;
    c->above_stride = (size_t)all[c->rank + 1].stride;
  }
  unsigned char st = (unsigned char)ok;
  std::vector<unsigned char> sts((size_t)c->nranks, 0);
  host_allgather(c, &st, sts.data(), 1);
  for (int p = 0; p < c->nranks; ++p) ok = ok && sts[p] == 1;
  if (!ok) {
    halo_teardown(c);
    return nullptr;
  }
  c->halo_seq = 0;
  return c->halo_local;
}

// true when buf is the communicator's own halo buffer (then the exchange runs through peer memory)
bool nccl_halo_is_peer_buffer(const NcclComm* c, const void* buf) {
  return c != nullptr && c->halo_local != nullptr && buf == (const void*)c->halo_local;
}

// Exchange through peer memory; returns the address of the parity buffer the SpMV of THIS exchange must read.
const void* nccl_halo_exchange_peer(NcclComm* c, const void* send_lo, const void* send_hi, cudaStream_t s) {
  const unsigned long long seq = ++c->halo_seq;
  const size_t par = (size_t)(seq & 1ull);
  const size_t es = (size_t)c->halo_es;
  unsigned char* mine = c->halo_local + par * c->halo_stride;
  const size_t my_data = round_up((c->halo_lo + c->halo_hi) * es, 256);
  // I am "above" rank-1: my first plane lands behind its from-below plane, and I raise its from-above flag
  unsigned char* below = c->halo_below ? c->halo_below + par * c->below_stride : nullptr;
  unsigned char* above = c->halo_above ? c->halo_above + par * c->above_stride : nullptr;
  void* dst_below = below ? below + c->below_lo * es : nullptr;
  void* dst_above = above ? above : nullptr;
  unsigned long long* flag_below = below ? reinterpret_cast<unsigned long long*>(below + (c->below_stride - 256) + 128) : nullptr;
  unsigned long long* flag_above = above ? reinterpret_cast<unsigned long long*>(above + (c->above_stride - 256)) : nullptr;
  const unsigned long long* from_below = c->halo_below ? reinterpret_cast<const unsigned long long*>(mine + my_data) : nullptr;
  const unsigned long long* from_above = c->halo_above ? reinterpret_cast<const unsigned long long*>(mine + my_data + 128) : nullptr;
  const size_t nmax = c->halo_lo > c->halo_hi ? c->halo_lo : c->halo_hi;
  int grid = (int)((nmax + 1023) / 1024);
  grid = grid < 1 ? 1 : (grid > 64 ? 64 : grid);
  if (es == 8)
    k_halo_exchange<double><<<grid, 256, 0, s>>>((const double*)send_lo, c->halo_lo, (double*)dst_below,
                                                 (const double*)send_hi, c->halo_hi, (double*)dst_above, flag_below,
                                                 flag_above, from_below, from_above, seq, c->halo_ticket,
                                                 wait_budget_cycles());
  else
    k_halo_exchange<float><<<grid, 256, 0, s>>>((const float*)send_lo, c->halo_lo, (float*)dst_below,
                                                (const float*)send_hi, c->halo_hi, (float*)dst_above, flag_below,
                                                flag_above, from_below, from_above, seq, c->halo_ticket,
                                                wait_budget_cycles());
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw CudaError(std::string("peer-memory halo exchange launch failed: ") + cudaGetErrorString(e));
  launch_stats().kernels++;
  return mine;
}

void nccl_allreduce_sum(NcclComm* c, void* buf, size_t count, bool is_double, cudaStream_t s) {
  if (p2p_allreduce_sum(c, buf, count, is_double, s)) return;
  check(api().AllReduce(buf, buf, count, is_double ? kNcclFloat64 : kNcclFloat32, kNcclSum, c->comm, s),
        "ncclAllReduce");
}
int nccl_rank(const NcclComm* c) { return c->rank; }
int nccl_nranks(const NcclComm* c) { return c->nranks; }

// neighbour exchange used by the driver's row-partitioned SpMV (halo planes), all in one group
void nccl_halo_exchange(NcclComm* c, const void* send_lo, void* recv_lo, size_t n_lo, const void* send_hi,
                        void* recv_hi, size_t n_hi, bool is_double, cudaStream_t s) {
  const int dt = is_double ? kNcclFloat64 : kNcclFloat32;
  check(api().GroupStart(), "ncclGroupStart");
  if (c->rank > 0 && n_lo) {
    check(api().Send(send_lo, n_lo, dt, c->rank - 1, c->comm, s), "ncclSend");
    check(api().Recv(recv_lo, n_lo, dt, c->rank - 1, c->comm, s), "ncclRecv");
  }
  if (c->rank < c->nranks - 1 && n_hi) {
    check(api().Send(send_hi, n_hi, dt, c->rank + 1, c->comm, s), "ncclSend");
    check(api().Recv(recv_hi, n_hi, dt, c->rank + 1, c->comm, s), "ncclRecv");
  }
  check(api().GroupEnd(), "ncclGroupEnd");
}
// the two collectives of a row-sharded A^T A operator (BASELINE config 5; EXAMPLES/SVD/dsvd.f:342-343 distributed):
// gather the ranks' slices of x, reduce-scatter the partial products
void nccl_allgather(NcclComm* c, const void* send, void* recv, size_t count_per_rank, bool is_double,
                    cudaStream_t s) {
  check(api().AllGather(send, recv, count_per_rank, is_double ? kNcclFloat64 : kNcclFloat32, c->comm, s),
        "ncclAllGather");
}
void nccl_reducescatter_sum(NcclComm* c, const void* send, void* recv, size_t count_per_rank, bool is_double,
                            cudaStream_t s) {
  check(api().ReduceScatter(send, recv, count_per_rank, is_double ? kNcclFloat64 : kNcclFloat32, kNcclSum,
                            c->comm, s),
        "ncclReduceScatter");
}

}  // namespace ab200

extern "C" {

// rank 0 creates the id (128 bytes) and shares it with the other ranks (e.g. torch.distributed broadcast)
int ab200_nccl_unique_id(void* out128) {
  auto& a = ab200::api();
  if (!a.error.empty()) {
    std::fprintf(stderr, "arpack_b200: %s\n", a.error.c_str());
    return -1;
  }
  ab200::ncclUniqueId id;
  if (a.GetUniqueId(&id) != 0) return -2;
  std::memcpy(out128, &id, 128);
  return 0;
}

// collective: every rank calls it with the same id; returns the handle to pass as `comm`
int ab200_comm_create(const void* id128, int rank, int nranks) {
  auto& a = ab200::api();
  if (!a.error.empty()) {
    std::fprintf(stderr, "arpack_b200: %s\n", a.error.c_str());
    return -1;
  }
  ab200::ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  auto* c = new ab200::NcclComm();
  c->rank = rank;
  c->nranks = nranks;
  const int r = a.CommInitRank(&c->comm, nranks, id, rank);
  if (r != 0) {
    std::fprintf(stderr, "arpack_b200: ncclCommInitRank failed: %s\n", a.GetErrorString(r));
    delete c;
    return -3;
  }
  try {
    ab200::p2p_setup(c);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: %s\n", e.what());
    a.CommDestroy(c->comm);
    delete c;
    return -4;
  }
  std::lock_guard<std::mutex> lk(ab200::g_mu);
  ab200::g_comms.push_back(c);
  return (int)ab200::g_comms.size();
}

void ab200_comm_destroy(int handle) {
  ab200::NcclComm* c = ab200::comm_from_handle(handle);
  if (!c) return;
  cudaDeviceSynchronize();
  ab200::p2p_teardown(c);
  if (c->comm) ab200::api().CommDestroy(c->comm);
  c->comm = nullptr;
  c->destroyed = true;  // the handle stays reserved; comm_from_handle() no longer resolves it
}

// 1 when the per-step reductions of this communicator run through peer memory (CUDA IPC over NVLink: fused into the
// kernels / one small kernel, sums in rank order), 0 when they use ncclAllReduce
int ab200_comm_uses_p2p(int handle) {
  ab200::NcclComm* c = ab200::comm_from_handle(handle);
  return (c && c->p2p) ? 1 : 0;
}
int ab200_comm_rank(int handle) {
  ab200::NcclComm* c = ab200::comm_from_handle(handle);
  return c ? c->rank : -1;
}
int ab200_comm_size(int handle) {
  ab200::NcclComm* c = ab200::comm_from_handle(handle);
  return c ? c->nranks : -1;
}
// Collective: the communicator's halo receive buffer for a row-partitioned operator whose lower / upper neighbour
// planes hold halo_lo / halo_hi elements of elem_size bytes (8 or 4).  Pass the returned address as halo_buf of
// ab200_csr_spmv_halo_f64 / ab200_register_csr_halo_op_f64: the plane exchange then runs through peer memory.
// NULL: peer memory is not available, allocate an ordinary device buffer instead (exchange by ncclSend/ncclRecv).
void* ab200_comm_halo_buffer(int handle, long long halo_lo, long long halo_hi, int elem_size) {
  ab200::NcclComm* c = ab200::comm_from_handle(handle);
  if (!c || halo_lo < 0 || halo_hi < 0 || (elem_size != 8 && elem_size != 4)) return nullptr;
  static const bool off = getenv("AB200_PEER_HALO") && std::strcmp(getenv("AB200_PEER_HALO"), "0") == 0;
  if (off) return nullptr;
  try {
    return ab200::nccl_halo_alloc(c, (size_t)halo_lo, (size_t)halo_hi, elem_size);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "arpack_b200: halo buffer: %s\n", e.what());
    return nullptr;
  }
}
}
