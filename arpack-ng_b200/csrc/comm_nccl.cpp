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

struct NcclComm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

namespace {
std::mutex g_mu;
std::vector<NcclComm*> g_comms;  // handle = index + 1
}  // namespace

NcclComm* comm_from_handle(int handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (handle < 1 || handle > (int)g_comms.size()) return nullptr;
  return g_comms[handle - 1];
}

void nccl_allreduce_sum(NcclComm* c, void* buf, size_t count, bool is_double, cudaStream_t s) {
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
  std::lock_guard<std::mutex> lk(ab200::g_mu);
  ab200::g_comms.push_back(c);
  return (int)ab200::g_comms.size();
}

void ab200_comm_destroy(int handle) {
  ab200::NcclComm* c = ab200::comm_from_handle(handle);
  if (!c) return;
  if (c->comm) ab200::api().CommDestroy(c->comm);
  c->comm = nullptr;
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
