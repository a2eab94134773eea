// peer.cuh -- device side of the peer-memory reductions (multi-GPU, one process per GPU, CUDA IPC over NVLink).
//
// Every rank owns one buffer that all its peers map (comm_nccl.cpp).  Layout, in bytes from the base:
//   [stand-alone all-reduce: 2 parities x 8 ranks x 160 values x 8 B][its flags: 8 x 128 B]
//   [fused reductions:  3 kinds x 2 parities x 8 ranks x 160 values x 8 B][their flags: 3 kinds x 8 ranks x 128 B]
// A value slot is addressed by (kind, parity of the sequence number, SOURCE rank); a flag by (kind, source rank).
// Protocol of one reduction with sequence number s: the producer stores its values into slot (kind, s & 1, me) of
// EVERY rank's buffer, fences at system scope, then stores s with release semantics into flag (kind, me) of every
// buffer; a consumer spins (acquire, system scope) on the flags of its OWN buffer until all of them are >= s and adds
// the slots in rank order -- every rank adds the same numbers in the same order.  Two parities suffice: a rank can
// only produce reduction s+2 of a kind after it consumed s+1, which needs every rank to have produced s+1, which each
// rank does after consuming s (stream order).
#pragma once
#include <cstddef>
#include <cstdint>

#include "vecops.hpp"

namespace ab200 {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerMaxCount = 160;
constexpr int kPeerKinds = 3;
constexpr size_t kPeerFlagStride = 128;  // one cache line per flag
constexpr size_t kPeerAloneData = 8ull * 2 * kPeerMaxRanks * kPeerMaxCount;
constexpr size_t kPeerAloneFlags = kPeerFlagStride * kPeerMaxRanks;
constexpr size_t kPeerFusedOff = kPeerAloneData + kPeerAloneFlags;
constexpr size_t kPeerFusedData = 8ull * kPeerKinds * 2 * kPeerMaxRanks * kPeerMaxCount;
constexpr size_t kPeerFusedFlagsOff = kPeerFusedOff + kPeerFusedData;
constexpr size_t kPeerBytes = kPeerFusedFlagsOff + kPeerFlagStride * kPeerKinds * kPeerMaxRanks;

#ifdef __CUDACC__
template <typename T>
__device__ __forceinline__ T* peer_slot(unsigned char* base, int kind, int par, int src) {
  return reinterpret_cast<T*>(base + kPeerFusedOff +
                              8ull * ((((size_t)kind * 2 + par) * kPeerMaxRanks + src) * kPeerMaxCount));
}
__device__ __forceinline__ unsigned long long* peer_flag(unsigned char* base, int kind, int src) {
  return reinterpret_cast<unsigned long long*>(base + kPeerFusedFlagsOff +
                                               kPeerFlagStride * ((size_t)kind * kPeerMaxRanks + src));
}
__device__ __forceinline__ void peer_store_release(unsigned long long* f, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_load_acquire(const unsigned long long* f) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
  return v;
}
// wait until rank src has published sequence number pr.seq (or later) into MY buffer
__device__ __forceinline__ void peer_wait_rank(const PeerReduce& pr, int src) {
  const unsigned long long* f = peer_flag(pr.base[pr.rank], pr.kind, src);
  long long t0 = 0;
  unsigned int spins = 0;
  while (peer_load_acquire(f) < pr.seq) {
    // a peer that never arrives must not hang the GPU for ever unless the user asked for that: trap after the
    // configured time (default 300 s, AB200_P2P_TIMEOUT_S; 0 = wait like MPI would)
    if (pr.timeout_cycles > 0 && ((++spins & 0x3FFu) == 0)) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > pr.timeout_cycles) __trap();
    }
  }
}
// my partial value c of the reduction -> slot (kind, parity, me) of every rank's buffer
template <typename T>
__device__ __forceinline__ void peer_put(const PeerReduce& pr, int c, T v) {
  const int par = (int)(pr.seq & 1ull);
  for (int p = 0; p < pr.nranks; ++p)
    *reinterpret_cast<volatile T*>(peer_slot<T>(pr.base[p], pr.kind, par, pr.rank) + c) = v;
}
// after every thread of the CTA has done its peer_put calls: make them visible, then publish the sequence number.
// Must be reached by all threads of the CTA.
__device__ __forceinline__ void peer_publish(const PeerReduce& pr) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pr.nranks) peer_store_release(peer_flag(pr.base[threadIdx.x], pr.kind, pr.rank), pr.seq);
}
// consumer: all threads of the CTA call this; afterwards peer_sum() may be used by any thread
__device__ __forceinline__ void peer_wait_all(const PeerReduce& pr) {
  if ((int)threadIdx.x < pr.nranks) peer_wait_rank(pr, (int)threadIdx.x);
  __syncthreads();
}
template <typename T>
__device__ __forceinline__ T peer_sum(const PeerReduce& pr, int c) {
  const int par = (int)(pr.seq & 1ull);
  T s = T(0);
  for (int p = 0; p < pr.nranks; ++p)
    s += *reinterpret_cast<const volatile T*>(peer_slot<T>(pr.base[pr.rank], pr.kind, par, p) + c);
  return s;
}
#endif  // __CUDACC__

}  // namespace ab200
