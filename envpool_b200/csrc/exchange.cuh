// Peer exchange: the one data-path collective of the sharded step.
//
// The reference has no counterpart (it is single-process); north_star asks that after every
// step each GPU holds the outputs of ALL envs.  Every rank owns one "gather" allocation
//
//     gather[2][world][slab_bytes] | flags[world] (u64) | ctl {blocks_done, seq, error}
//
// mapped into every peer (CUDA IPC between processes, plain pointers inside one process).
// A step writes its packed output slab straight into gather[parity][rank] of the LOCAL
// allocation (no staging copy); `push_kernel` then stores that slice into gather[parity][rank]
// of every peer over NVLink (16-byte stores, one read of the L2-hot slice per peer set) and
// publishes the step's sequence number in flags[rank] of every rank with a system-scope
// release.  `wait_kernel` (one warp) acquires flags[0..world) >= seq before the consumer of
// the gathered batch runs.  The two parities make step t+1's stores land in the other half
// while step t is still being consumed: a peer cannot start pushing step t+2 before it has
// seen this rank's flag for t+1, which this rank raises only after (stream order) its consumer
// of step t -- so no acknowledgement traffic is needed.
//
// Why not ncclAllGather: measured 35 us per call for the 2.75 MB CartPole slab on 2 GPUs
// (protocol latency), against ~4 us for the step itself; the peer stores cost the NVLink time
// of the payload and one flag round trip.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace epb {

constexpr int kMaxPeers = 16;

struct PeerView {
  char* slice[kMaxPeers];                // gather[parity][rank] in the allocation of rank g
  unsigned long long* flag[kMaxPeers];   // &flags[rank] in the allocation of rank g
  int world;
  int rank;
};

struct ExchangeCtl {
  unsigned int blocks_done;
  unsigned int pad;
  unsigned long long seq;   // steps pushed by this rank
  int error;                // 1 = a wait timed out
  int pad2;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Copy n16 16-byte units of the local slice to every peer, then signal.
__global__ void __launch_bounds__(256)
push_kernel(PeerView pv, int64_t n16, ExchangeCtl* ctl) {
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(pv.slice[pv.rank]);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint4 v = src[i];
#pragma unroll 1
    for (int g = 0; g < pv.world; ++g)
      if (g != pv.rank) reinterpret_cast<uint4*>(pv.slice[g])[i] = v;
  }
  // last-block-done: every block fences its peer stores at system scope before it counts
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int ticket = atomicAdd(&ctl->blocks_done, 1u);
    if (ticket == gridDim.x - 1) {
      ctl->blocks_done = 0;
      unsigned long long s = ctl->seq + 1;
      ctl->seq = s;
      __threadfence_system();
      for (int g = 0; g < pv.world; ++g) st_release_sys(pv.flag[g], s);
    }
  }
}

// One warp: lane g waits until rank g's slice of step `ctl->seq` has landed here.
// Bounded: ~4e9 cycles (about 2 s) without progress sets ctl->error instead of hanging.
__global__ void wait_kernel(const unsigned long long* flags, int world, ExchangeCtl* ctl) {
  const unsigned long long want = ctl->seq;
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + threadIdx.x) < want) {
      if (clock64() - t0 > 4000000000LL) {
        atomicExch(&ctl->error, 1);
        break;
      }
      __nanosleep(64);
    }
  }
}

}  // namespace epb
