// Peer exchange: the one data-path collective of the sharded step.
//
// The reference has no counterpart (it is single-process); north_star asks that after every
// step each GPU holds the outputs of ALL envs.  Every rank owns one "gather" allocation
//
//     gather[2][world][slab_bytes] | flags[world] (u64) | ctl {blocks_done, seq, error}
//
// mapped into every peer (CUDA IPC between processes, plain pointers inside one process).
// A step writes its packed output slab straight into gather[parity][rank] of the LOCAL
// allocation (no staging copy) and the same bytes go into gather[parity][rank] of every peer
// over NVLink with 16-byte stores, after which the step's sequence number is published in
// flags[rank] of every rank with a system-scope release.  Two producers of the peer stores:
//   * fused (classic_control / toy_text): the step kernel's own epilogue -- each CTA forwards
//     the rows it has just written (`peer_forward_rows`), so the transfer rides inside the
//     step launch and no second kernel sits on the critical path;
//   * `push_kernel` (HalfCheetah, or ENVPOOL_B200_EXCHANGE=push): a copy kernel behind the
//     step kernel (capi.cu).
// `wait_kernel` (one warp, capi.cu) acquires flags[0..world) >= seq before the consumer of
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

constexpr int kMaxCols = 13;  // 8 common state keys + at most 5 env keys (OutView::env)

struct ExchangeCtl {
  unsigned int blocks_done;
  unsigned int pad;
  unsigned long long seq;   // steps pushed by this rank
  int error;                // 1 = a wait timed out
  int pad2;
};

struct PeerView {
  char* slice[kMaxPeers];                // gather[parity][rank] in the allocation of rank g
  unsigned long long* flag[kMaxPeers];   // &flags[rank] in the allocation of rank g
  ExchangeCtl* ctl;                      // this rank's control block
  int world;
  int rank;
  // packed-slab columns (fused epilogue): byte offset in the slab and bytes per row
  int ncols;
  int col_rb[kMaxCols];
  int64_t col_off[kMaxCols];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Last-block-done publication: every CTA fences its peer stores at system scope before it
// counts itself; the CTA that completes the count bumps the sequence number and raises
// flags[rank] on every rank.  Call with all threads of the CTA.
__device__ __forceinline__ void peer_publish(const PeerView* __restrict__ pv) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    ExchangeCtl* ctl = pv->ctl;
    unsigned int ticket = atomicAdd(&ctl->blocks_done, 1u);
    if (ticket == gridDim.x - 1) {
      ctl->blocks_done = 0;
      unsigned long long s = ctl->seq + 1;
      ctl->seq = s;
      __threadfence_system();
      const int world = pv->world;
      for (int g = 0; g < world; ++g) st_release_sys(pv->flag[g], s);
    }
  }
}

// Fused epilogue of the step kernels: the CTA forwards the output rows [row0, row0 + kB)
// it has just written into its local gather slice to every peer, column by column, then
// publishes.  Requires identity row<->env mapping (sync step of all envs) and kB % 16 == 0,
// so that every per-column chunk starts 16-byte aligned; the tail CTA may copy up to 15
// bytes past its last row, which stays inside the column's 256-byte padding.
template <int kB>
__device__ __forceinline__ void peer_forward_rows(const PeerView* __restrict__ pv, int64_t row0,
                                                  int n) {
  static_assert(kB % 16 == 0, "CTA rows must keep 1-byte columns 16-byte aligned");
  __syncthreads();  // all rows of this CTA are written (by this CTA)
  const int64_t left = (int64_t)n - row0;
  const int rows = left < kB ? (int)left : kB;
  const int world = pv->world, rank = pv->rank, ncols = pv->ncols;
  const char* local = pv->slice[rank];
  for (int k = 0; k < ncols; ++k) {
    const int rb = pv->col_rb[k];
    const int64_t off = pv->col_off[k] + row0 * rb;
    const int n16 = (rows * rb + 15) >> 4;
    for (int i = threadIdx.x; i < n16; i += kB) {
      const uint4 v = *reinterpret_cast<const uint4*>(local + off + 16 * (int64_t)i);
#pragma unroll 1
      for (int g = 0; g < world; ++g)
        if (g != rank) *reinterpret_cast<uint4*>(pv->slice[g] + off + 16 * (int64_t)i) = v;
    }
  }
  peer_publish(pv);
}

}  // namespace epb
