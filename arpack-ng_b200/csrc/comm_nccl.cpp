// comm_nccl.cpp -- NCCL communicator handles standing in for PARPACK's MPI communicator.
//
// The p*_c entry points keep the reference signature (ICB/parpack.h:20-27): their first argument is
// an MPI_Fint.  Here that integer is a handle into a table of NCCL communicators created with
// ab200_comm_create().  NCCL is bound at run time with dlopen so that the library (a) uses the very
// libnccl.so.2 a host process such as PyTorch has already loaded and (b) still loads on a box
// without NCCL for single-GPU use.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

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
}  // namespace

// Peer-memory all-reduce for the tiny per-step reductions (<= kP2pMaxCount values): every rank owns a buffer that its
// peers map through CUDA IPC; one small kernel per all-reduce writes this rank's contribution into every peer's slot
// (stores over NVLink), publishes a sequence number, waits for the peers' numbers and sums the slots in rank order
// (so every rank obtains bit-identical results).  Latency: one kernel launch + one NVLink round trip instead of an
// NCCL collective.  Falls back to NCCL when IPC is unavailable, for more than 8 ranks, or with AB200_P2P=0.
constexpr int kP2pMaxRanks = 8;
constexpr int kP2pMaxCount = 160;
constexpr size_t kP2pDataBytes = sizeof(double) * 2 * kP2pMaxRanks * kP2pMaxCount;  // [parity][rank][value]
constexpr size_t kP2pFlagStride = 128;                                              // one cache line per flag
constexpr size_t kP2pBytes = kP2pDataBytes + kP2pFlagStride * kP2pMaxRanks;

struct P2pPeers {
  unsigned char* base[kP2pMaxRanks];
};

struct NcclComm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  bool p2p = false;
  unsigned char* p2p_local = nullptr;
  P2pPeers peers{};
  unsigned long long seq = 0;
};

template <typename T>
__global__ void __launch_bounds__(kP2pMaxCount) k_p2p_allreduce(T* mb, int count, int rank, int nranks, P2pPeers peers,
                                                                unsigned long long seq) {
  const int tid = threadIdx.x;
  const int par = (int)(seq & 1ull);
  if (tid < count) {
    const T v = mb[tid];
    for (int p = 0; p < nranks; ++p) {
      T* slot = reinterpret_cast<T*>(peers.base[p]) + ((size_t)par * kP2pMaxRanks + rank) * kP2pMaxCount + tid;
      *reinterpret_cast<volatile T*>(slot) = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < nranks) {
    unsigned long long* f =
        reinterpret_cast<unsigned long long*>(peers.base[tid] + kP2pDataBytes + (size_t)rank * kP2pFlagStride);
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
    // wait for rank tid's contribution to land in MY buffer
    const unsigned long long* mine =
        reinterpret_cast<const unsigned long long*>(peers.base[rank] + kP2pDataBytes + (size_t)tid * kP2pFlagStride);
    unsigned long long got = 0;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(mine) : "memory");
      if (got >= seq) break;
      // a peer never arrived: fail loudly instead of hanging.  The bound (~60 s) is far above any start-up skew
      // between ranks (module loading, allocations), which an MPI or NCCL collective would simply wait out.
      if (clock64() - t0 > 120000000000LL) __trap();
    } while (true);
  }
  __syncthreads();
  if (tid < count) {
    const T* slots = reinterpret_cast<const T*>(peers.base[rank]) + (size_t)par * kP2pMaxRanks * kP2pMaxCount + tid;
    T sum = T(0);
    for (int p = 0; p < nranks; ++p) sum += *reinterpret_cast<const volatile T*>(slots + (size_t)p * kP2pMaxCount);
    mb[tid] = sum;
  }
}

namespace {
std::mutex g_mu;
std::vector<NcclComm*> g_comms;  // handle = index + 1
}  // namespace

NcclComm* comm_from_handle(int handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (handle < 1 || handle > (int)g_comms.size()) return nullptr;
  return g_comms[handle - 1];
}

// returns false when the peer-memory path does not apply (the caller then uses NCCL)
bool p2p_allreduce_sum(NcclComm* c, void* buf, size_t count, bool is_double, cudaStream_t s) {
  if (!c->p2p || count > (size_t)kP2pMaxCount) return false;
  const unsigned long long seq = ++c->seq;
  if (is_double)
    k_p2p_allreduce<double><<<1, kP2pMaxCount, 0, s>>>((double*)buf, (int)count, c->rank, c->nranks, c->peers, seq);
  else
    k_p2p_allreduce<float><<<1, kP2pMaxCount, 0, s>>>((float*)buf, (int)count, c->rank, c->nranks, c->peers, seq);
  return cudaGetLastError() == cudaSuccess;
}

namespace {
// collective: map every rank's buffer into every other rank (CUDA IPC); all ranks agree on the outcome
void p2p_setup(NcclComm* c) {
  const char* e = getenv("AB200_P2P");
  if ((e && std::strcmp(e, "0") == 0) || c->nranks < 2 || c->nranks > kP2pMaxRanks) return;
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&c->p2p_local, kP2pBytes) != cudaSuccess || cudaMemset(c->p2p_local, 0, kP2pBytes) != cudaSuccess ||
      cudaIpcGetMemHandle(&mine, c->p2p_local) != cudaSuccess) {
    cudaGetLastError();
    ok = 0;
  }
  // exchange the handles (and the per-rank status) with the communicator we already have
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  const size_t item = sizeof(cudaIpcMemHandle_t) + 8;
  std::vector<unsigned char> h_send(item, 0), h_recv(item * c->nranks, 0);
  std::memcpy(h_send.data(), &mine, sizeof(mine));
  h_send[sizeof(mine)] = (unsigned char)ok;
  if (cudaMalloc(&d_send, item) != cudaSuccess || cudaMalloc(&d_recv, item * c->nranks) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(d_send);
    cudaFree(d_recv);
    throw CudaError("p2p_setup: cudaMalloc of the exchange buffers failed");
  }
  cudaMemcpy(d_send, h_send.data(), item, cudaMemcpyHostToDevice);
  check(api().AllGather(d_send, d_recv, item, /*ncclChar*/ 0, c->comm, 0), "ncclAllGather(ipc handles)");
  cudaStreamSynchronize(0);
  cudaMemcpy(h_recv.data(), d_recv, item * c->nranks, cudaMemcpyDeviceToHost);
  for (int p = 0; p < c->nranks; ++p) ok = ok && h_recv[p * item + sizeof(mine)] == 1;
  int opened = 0;
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
      ++opened;
    }
  }
  // second round: everybody must have mapped everybody, otherwise nobody uses the path
  h_send[0] = (unsigned char)ok;
  cudaMemcpy(d_send, h_send.data(), 1, cudaMemcpyHostToDevice);
  check(api().AllGather(d_send, d_recv, 1, 0, c->comm, 0), "ncclAllGather(p2p status)");
  cudaStreamSynchronize(0);
  cudaMemcpy(h_recv.data(), d_recv, c->nranks, cudaMemcpyDeviceToHost);
  for (int p = 0; p < c->nranks; ++p) ok = ok && h_recv[p] == 1;
  cudaFree(d_send);
  cudaFree(d_recv);
  c->p2p = ok != 0;
  if (!c->p2p && getenv("AB200_DEBUG"))
    std::fprintf(stderr, "arpack_b200: rank %d: peer-memory all-reduce unavailable, using NCCL\n", c->rank);
  (void)opened;
}

void p2p_teardown(NcclComm* c) {
  if (c->p2p)
    for (int p = 0; p < c->nranks; ++p)
      if (p != c->rank && c->peers.base[p]) cudaIpcCloseMemHandle(c->peers.base[p]);
  c->p2p = false;
  if (c->p2p_local) cudaFree(c->p2p_local);
  c->p2p_local = nullptr;
}
}  // namespace

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
}

// 1 when the small all-reduces of this communicator go through peer memory, 0 when they use NCCL
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
}
