// envpool_b200 C ABI (include/envpool_b200.h): pool lifetime, key tables, the host-buffer
// send/recv path (pinned staging + one packed D2H per batch) and the device-resident
// step / rollout path.  This file is the GPU-side replacement for the reference's
// AsyncEnvPool + ActionBufferQueue + StateBufferQueue (envpool/core/async_envpool.h,
// action_buffer_queue.h, state_buffer_queue.h): the "queue" is a CUDA stream, the
// "state buffer" is a packed output slab in HBM mirrored into recycled pinned host slabs.
#include <cuda_runtime.h>

#include <climits>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/envpool_b200.h"
#include "common.cuh"
#include "mujoco.cuh"

namespace epb {

__global__ void seed_kernel(StateView sv, int base_seed, const int32_t* env_seed) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= sv.n_envs) return;
  // Env::ResolveSeed (core/env.h:101-111): env_seed[env_id] or seed + env_id
  uint32_t s = env_seed ? (uint32_t)env_seed[e] : (uint32_t)(base_seed + sv.env_id_offset + e);
  // chunked layout (common.cuh): chunk c of env e is the 8 words at ((c*N + e)*8)
  const int64_t N = sv.n_envs;
  uint32_t w[8];
  for (int c = 0; c < kMtN / 8; ++c) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = c * 8 + k;
      if (i > 0) s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i;
      w[k] = s;
    }
    uint4* dst = reinterpret_cast<uint4*>(sv.mt + ((int64_t)c * N + e) * 8);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  sv.mt_idx[e] = 0;  // std::mt19937 starts exhausted: the first draw regenerates word 0
  sv.flags[e] = -1;  // current_step_ = -1 (env.h:81), done_ = true (cartpole.h:67)
}

thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define EPB_CUDA(expr)                                                            \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess)                                                        \
      return fail(EPB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// Every entry point runs on the pool's device but leaves the caller's current device as it
// found it (a host process may drive several pools / use torch on another GPU).
struct DeviceGuard {
  int prev = -1;
  cudaError_t status;
  explicit DeviceGuard(int dev) {
    status = cudaGetDevice(&prev);
    if (status == cudaSuccess && prev != dev) status = cudaSetDevice(dev);
    else if (status == cudaSuccess) prev = -1;  // nothing to restore
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct Key {
  const char* name;
  int dtype, ndim, shape[3], row_bytes;
  int64_t off;
};

static int dtype_size(int d) { return d == EPB_F64 ? 8 : d == EPB_BOOL ? 1 : 4; }

struct Pending {
  void* slab;
  int n;       // rows written by this send/reset
  int row0;    // rows already handed out (async mode hands out batch_size rows at a time)
  cudaEvent_t ev;
  bool ready;  // event already waited for
};

}  // namespace epb

using namespace epb;

struct epb_pool {
  int kind = 0;
  epb_config cfg{};
  int N = 0;
  int precision = 0;
  std::vector<Key> keys;
  Key act{};
  int64_t slab_bytes = 0;
  int NR = 0, NI = 0, real_size = 8;
  StateView sv{};
  // device allocations
  void* d_state_blob = nullptr;  // flags | mt_idx | istate | rstate | mt | (mujoco extras)
  int64_t state_bytes = 0;
  char* d_slab = nullptr;
  char* d_last = nullptr;  // slab the most recent *_device launch wrote (d_slab or a gather slice)
  void* d_action = nullptr;
  int32_t* d_ids = nullptr;
  // pinned host staging
  void* h_action[2] = {nullptr, nullptr};
  int32_t* h_ids[2] = {nullptr, nullptr};
  cudaEvent_t h_stage_ev[2] = {nullptr, nullptr};
  int stage_flip = 0;
  std::vector<int32_t> arange;  // 0..N-1, for the identity-ids fast path
  int batch = 0;                // rows per recv; < N = async mode (async_envpool.h:93-97)
  std::vector<std::pair<void*, int>> leases;  // slab -> outstanding recv leases + queue refs
  std::vector<void*> free_slabs;
  std::vector<void*> all_slabs;
  std::deque<Pending> pending;
  std::vector<cudaEvent_t> free_events;
  std::mutex mu;
  cudaStream_t stream = nullptr;
  launch_fn step_fn = nullptr, rollout_fn = nullptr;
  MjcPool* mjc = nullptr;
  // cached CUDA graphs of K-step chains (epb_step_many_device), most recent first
  struct GraphEntry {
    cudaGraphExec_t exec;
    const void* actions;
    int T, t0, K;
    cudaStream_t stream;
  };
  std::vector<GraphEntry> graphs;
  int64_t launches = 0;
  int bytes_per_step = 0;
  // peer exchange (exchange.cuh): gather[2][world][slab] | flags[world] | ctl
  char* x_base = nullptr;
  int x_world = 0, x_rank = 0;
  int64_t x_flags_off = 0, x_ctl_off = 0, x_bytes = 0;
  char* x_peer[kMaxPeers] = {};
  bool x_ipc[kMaxPeers] = {};
  bool x_attached = false;
  long long x_timeout_ns = 10000000000LL;
  bool x_fused = false;  // peer stores issued by the step kernel's epilogue (else push_kernel)
  uint64_t x_steps = 0;  // host count of exchanged steps; parity picks the gather half

  int64_t x_mine(int parity) const {
    return ((int64_t)parity * x_world + x_rank) * slab_bytes;
  }
  ExchangeCtl* x_ctl() const { return reinterpret_cast<ExchangeCtl*>(x_base + x_ctl_off); }
  // the two per-parity PeerViews live behind the control block, in device memory
  PeerView* x_view(int parity) const {
    return reinterpret_cast<PeerView*>(x_base + x_ctl_off + 256) + parity;
  }

  OutView slab_view(char* base) const {
    OutView ov{};
    auto col = [&](int k) { return static_cast<void*>(base + keys[k].off); };
    ov.env_id = static_cast<int32_t*>(col(0));
    ov.players_id = static_cast<int32_t*>(col(1));
    ov.elapsed = static_cast<int32_t*>(col(2));
    ov.done = static_cast<uint8_t*>(col(3));
    ov.reward = static_cast<float*>(col(4));
    ov.discount = static_cast<float*>(col(5));
    ov.step_type = static_cast<int32_t*>(col(6));
    ov.trunc = static_cast<uint8_t*>(col(7));
    for (size_t k = 8; k < keys.size(); ++k) ov.env[k - 8] = col((int)k);
    ov.t_stride_rows = N;
    return ov;
  }
};

namespace {

void add_key(epb_pool* p, const char* name, int dtype, std::initializer_list<int> shape) {
  Key k{};
  k.name = name;
  k.dtype = dtype;
  k.ndim = (int)shape.size();
  int elems = 1, i = 0;
  for (int s : shape) {
    k.shape[i++] = s;
    elems *= s;
  }
  k.row_bytes = elems * dtype_size(dtype);
  p->keys.push_back(k);
}

int build_keys(epb_pool* p) {
  // common_state_spec, envpool/core/env_spec.h:37-43
  add_key(p, "info:env_id", EPB_I32, {});
  add_key(p, "info:players.env_id", EPB_I32, {});
  add_key(p, "elapsed_step", EPB_I32, {});
  add_key(p, "done", EPB_BOOL, {});
  add_key(p, "reward", EPB_F32, {});
  add_key(p, "discount", EPB_F32, {});
  add_key(p, "step_type", EPB_I32, {});
  add_key(p, "trunc", EPB_BOOL, {});
  Key& a = p->act;
  a = Key{};
  a.name = "action";
  a.dtype = EPB_I32;
  a.ndim = 0;
  a.row_bytes = 4;
  switch (p->kind) {
    case EPB_CARTPOLE: add_key(p, "obs", EPB_F32, {4}); p->NR = 4; break;
    case EPB_PENDULUM:
      add_key(p, "obs", EPB_F32, {3}); p->NR = 2;
      a.dtype = EPB_F32; a.ndim = 1; a.shape[0] = 1;
      break;
    case EPB_ACROBOT:
      add_key(p, "obs", EPB_F32, {6});
      add_key(p, "info:state", EPB_F32, {2});
      p->NR = 4;
      break;
    case EPB_MOUNTAIN_CAR: add_key(p, "obs", EPB_F32, {2}); p->NR = 2; break;
    case EPB_MOUNTAIN_CAR_CONTINUOUS:
      add_key(p, "obs", EPB_F32, {2}); p->NR = 2;
      a.dtype = EPB_F32; a.ndim = 1; a.shape[0] = 1;
      break;
    case EPB_FROZEN_LAKE: case EPB_TAXI: case EPB_NCHAIN:
      add_key(p, "obs", EPB_I32, {}); p->NI = 1; break;
    case EPB_CATCH: add_key(p, "obs", EPB_F32, {10, 5}); p->NI = 1; break;
    case EPB_CLIFF_WALKING:
      add_key(p, "obs", EPB_I32, {});
      add_key(p, "info:prob", EPB_F32, {});
      p->NI = 1;
      break;
    case EPB_BLACKJACK: add_key(p, "obs", EPB_I32, {3}); p->NI = 2; break;
    case EPB_HALF_CHEETAH:
      // mujoco/gym/half_cheetah.h:44-62
      add_key(p, "obs", EPB_F64, {17});
      add_key(p, "info:reward_run", EPB_F64, {});
      add_key(p, "info:reward_ctrl", EPB_F64, {});
      add_key(p, "info:x_position", EPB_F64, {});
      add_key(p, "info:x_velocity", EPB_F64, {});
      a.dtype = EPB_F64; a.ndim = 1; a.shape[0] = 6; a.row_bytes = 48;
      break;
    default: return -1;
  }
  int64_t off = 0;
  for (Key& k : p->keys) {
    k.off = off;
    off += ((int64_t)k.row_bytes * p->N + 255) / 256 * 256;
  }
  p->slab_bytes = off;
  return 0;
}

int get_slab(epb_pool* p, void** out) {
  std::lock_guard<std::mutex> lk(p->mu);
  if (!p->free_slabs.empty()) {
    *out = p->free_slabs.back();
    p->free_slabs.pop_back();
    return EPB_OK;
  }
  void* h = nullptr;
  cudaError_t e = cudaHostAlloc(&h, (size_t)p->slab_bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(EPB_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  p->all_slabs.push_back(h);
  *out = h;
  return EPB_OK;
}

int get_event(epb_pool* p, cudaEvent_t* ev) {
  std::lock_guard<std::mutex> lk(p->mu);
  if (!p->free_events.empty()) {
    *ev = p->free_events.back();
    p->free_events.pop_back();
    return EPB_OK;
  }
  EPB_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  return EPB_OK;
}

// Copy n16 16-byte units of the local slice to every peer, then publish.
__global__ void __launch_bounds__(256)
push_kernel(const PeerView* __restrict__ pv, int64_t n16) {
  const int world = pv->world, rank = pv->rank;
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(pv->slice[rank]);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    uint4 v = src[i];
#pragma unroll 1
    for (int g = 0; g < world; ++g)
      if (g != rank) reinterpret_cast<uint4*>(pv->slice[g])[i] = v;
  }
  peer_publish(pv);
}

// One warp: lane g waits until rank g's slice of step `ctl->seq` has landed here.
// Bounded: `timeout_ns` without progress (default 10 s, ENVPOOL_B200_EXCHANGE_TIMEOUT_S) sets
// ctl->error instead of hanging the GPU when a peer has died.
__global__ void wait_kernel(const unsigned long long* flags, int world, ExchangeCtl* ctl,
                            long long timeout_ns) {
  const unsigned long long want = ctl->seq;
  if ((int)threadIdx.x < world) {
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(flags + threadIdx.x) < want) {
      long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        atomicExch(&ctl->error, 1);
        break;
      }
      __nanosleep(64);
    }
  }
}

// Build the two per-parity PeerViews from the attached peer bases and copy them to the device.
int upload_views(epb_pool* p) {
  PeerView v[2];
  memset(v, 0, sizeof(v));
  for (int parity = 0; parity < 2; ++parity) {
    PeerView& pv = v[parity];
    pv.world = p->x_world;
    pv.rank = p->x_rank;
    pv.ctl = p->x_ctl();
    for (int g = 0; g < p->x_world; ++g) {
      pv.slice[g] = p->x_peer[g] + p->x_mine(parity);
      pv.flag[g] =
          reinterpret_cast<unsigned long long*>(p->x_peer[g] + p->x_flags_off) + p->x_rank;
    }
    pv.ncols = (int)p->keys.size();
    for (int k = 0; k < pv.ncols; ++k) {
      pv.col_rb[k] = p->keys[k].row_bytes;
      pv.col_off[k] = p->keys[k].off;
    }
  }
  EPB_CUDA(cudaMemcpy(p->x_view(0), v, sizeof(v), cudaMemcpyHostToDevice));
  p->x_attached = true;
  return EPB_OK;
}

// Launch one batch step on `stream`.  d_action/d_ids are device pointers.
int launch_batch(epb_pool* p, const void* d_action, const int32_t* d_ids, int n,
                 int force_reset, char* d_slab, cudaStream_t stream,
                 const PeerView* peers = nullptr) {
  p->d_last = d_slab;
  if (p->kind == EPB_HALF_CHEETAH) {
    EPB_CUDA(mjc_launch_step(p->mjc, p->sv, p->slab_view(d_slab),
                             static_cast<const double*>(d_action), d_ids, n, force_reset,
                             stream));
    ++p->launches;
    return EPB_OK;
  }
  LaunchArgs a{};
  a.sv = p->sv;
  a.ov = p->slab_view(d_slab);
  a.action = d_action;
  a.env_ids = d_ids;
  a.n = n;
  a.force_reset = force_reset;
  a.stream = stream;
  a.peers = peers;
  EPB_CUDA(p->step_fn(a));
  ++p->launches;
  return EPB_OK;
}

// Host path shared by send and reset.
int host_submit(epb_pool* p, const void* action, const int32_t* env_ids, int n,
                int force_reset) {
  if (n <= 0 || n > p->N) return fail(EPB_ERR_INVALID, "batch rows must be in [1, num_envs]");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  const int f = p->stage_flip;
  p->stage_flip ^= 1;
  // the staging pair alternates; wait until the copy that last used this half is done
  EPB_CUDA(cudaEventSynchronize(p->h_stage_ev[f]));
  bool identity = (n == p->N);
  if (env_ids) {
    // fast path: the usual sync-mode call passes env_id == arange(N) (python/envpool.py
    // all_env_ids); one memcmp against a cached arange settles it
    if (identity && memcmp(env_ids, p->arange.data(), sizeof(int32_t) * n) == 0) {
      // identity gather, nothing to upload
    } else {
      bool ok = true;
      identity = false;
      const unsigned un = (unsigned)p->N;
      for (int i = 0; i < n; ++i) ok &= (unsigned)env_ids[i] < un;
      if (!ok) return fail(EPB_ERR_INVALID, "env_id out of range");
    }
  } else if (n != p->N) {
    identity = false;  // rows 0..n-1 of a partial batch: ids are 0..n-1
  }
  const int32_t* d_ids = nullptr;
  if (!identity) {
    if (env_ids) {
      memcpy(p->h_ids[f], env_ids, sizeof(int32_t) * n);
    } else {
      for (int i = 0; i < n; ++i) p->h_ids[f][i] = i;
    }
    EPB_CUDA(cudaMemcpyAsync(p->d_ids, p->h_ids[f], sizeof(int32_t) * n,
                             cudaMemcpyHostToDevice, p->stream));
    d_ids = p->d_ids;
  }
  if (!force_reset) {
    if (!action) return fail(EPB_ERR_INVALID, "action is NULL");
    size_t bytes = (size_t)p->act.row_bytes * n;
    memcpy(p->h_action[f], action, bytes);
    EPB_CUDA(cudaMemcpyAsync(p->d_action, p->h_action[f], bytes, cudaMemcpyHostToDevice,
                             p->stream));
  }
  EPB_CUDA(cudaEventRecord(p->h_stage_ev[f], p->stream));
  int rc = launch_batch(p, p->d_action, d_ids, n, force_reset, p->d_slab, p->stream);
  if (rc != EPB_OK) return rc;
  void* slab = nullptr;
  rc = get_slab(p, &slab);
  if (rc != EPB_OK) return rc;
  if (n == p->N) {
    EPB_CUDA(cudaMemcpyAsync(slab, p->d_slab, (size_t)p->slab_bytes, cudaMemcpyDeviceToHost,
                             p->stream));
  } else {
    for (const Key& k : p->keys) {
      EPB_CUDA(cudaMemcpyAsync(static_cast<char*>(slab) + k.off, p->d_slab + k.off,
                               (size_t)k.row_bytes * n, cudaMemcpyDeviceToHost, p->stream));
    }
  }
  cudaEvent_t ev;
  rc = get_event(p, &ev);
  if (rc != EPB_OK) return rc;
  EPB_CUDA(cudaEventRecord(ev, p->stream));
  {
    std::lock_guard<std::mutex> lk(p->mu);
    p->pending.push_back(Pending{slab, n, 0, ev, false});
    p->leases.emplace_back(slab, 1);  // the queue's own reference
  }
  return EPB_OK;
}

}  // namespace

extern "C" {

const char* epb_last_error(void) { return g_err.c_str(); }
int epb_abi_version(void) { return EPB_ABI_VERSION; }

int epb_create(int kind, const epb_config* cfg, epb_pool** out) {
  if (!cfg || !out) return fail(EPB_ERR_INVALID, "null argument");
  if (kind < 0 || kind >= EPB_NUM_KINDS) return fail(EPB_ERR_INVALID, "unknown env kind");
  if (cfg->num_envs <= 0) return fail(EPB_ERR_INVALID, "num_envs must be positive");
  // EnvSpec ctor check, envpool/core/env_spec.h:75-80
  if (cfg->batch_size > cfg->num_envs)
    return fail(EPB_ERR_INVALID,
                "It is required that batch_size <= num_envs, got num_envs = " +
                    std::to_string(cfg->num_envs) +
                    ", batch_size = " + std::to_string(cfg->batch_size));
  if (cfg->batch_size < 0) return fail(EPB_ERR_INVALID, "batch_size must be >= 0");
  int ndev = 0;
  EPB_CUDA(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(EPB_ERR_INVALID, "bad device ordinal");
  DeviceGuard guard(cfg->device);
  EPB_CUDA(guard.status);

  epb_pool* p = new epb_pool();
  p->kind = kind;
  p->cfg = *cfg;
  p->cfg.env_seed = nullptr;
  p->N = cfg->num_envs;
  p->precision = cfg->precision == EPB_PREC_F32 ? 1 : 0;
  p->real_size = p->precision ? 4 : 8;
  if (build_keys(p) != 0) {
    delete p;
    return fail(EPB_ERR_INVALID, "unknown env kind");
  }
  int iopt = cfg->iopt;
  if (iopt < 0) iopt = kind == EPB_FROZEN_LAKE ? 4 : kind == EPB_BLACKJACK ? 2 : 0;
  if (kind == EPB_FROZEN_LAKE && iopt != 4 && iopt != 8) {
    delete p;
    return fail(EPB_ERR_INVALID, "FrozenLake size must be 4 or 8");
  }
  if (kind == EPB_HALF_CHEETAH) {
    p->precision = 0;  // HalfCheetah physics is fp64 only
    p->real_size = 8;
    p->mjc = mjc_pool_create(p->N, p->precision, cfg->frame_skip > 0 ? cfg->frame_skip : 5,
                             cfg->ctrl_cost_weight >= 0 ? cfg->ctrl_cost_weight : 0.1,
                             cfg->forward_reward_weight >= 0 ? cfg->forward_reward_weight : 1.0,
                             cfg->reset_noise_scale >= 0 ? cfg->reset_noise_scale : 0.1);
    if (!p->mjc) {
      delete p;
      return fail(EPB_ERR_CUDA, "HalfCheetah model setup failed");
    }
    p->NR = mjc_state_reals(p->mjc);
    p->NI = 0;
  }
  const int64_t N = p->N;
  auto al = [](int64_t b) { return (b + 255) / 256 * 256; };
  int64_t o_flags = 0;
  int64_t o_idx = o_flags + al(4 * N);
  int64_t o_ist = o_idx + al(4 * N);
  int64_t o_rst = o_ist + al(4 * N * (p->NI > 0 ? p->NI : 1));
  int64_t o_mt = o_rst + al((int64_t)p->real_size * N * (p->NR > 0 ? p->NR : 1));
  p->state_bytes = o_mt + al(4 * N * kMtN);
  cudaError_t e = cudaMalloc(&p->d_state_blob, (size_t)p->state_bytes);
  if (e == cudaSuccess) e = cudaMemset(p->d_state_blob, 0, (size_t)p->state_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_slab, (size_t)p->slab_bytes);
  if (e == cudaSuccess) e = cudaMemset(p->d_slab, 0, (size_t)p->slab_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_action, (size_t)p->act.row_bytes * N);
  if (e == cudaSuccess) e = cudaMalloc(&p->d_ids, 4 * (size_t)N);
  for (int f = 0; f < 2 && e == cudaSuccess; ++f) {
    e = cudaHostAlloc(&p->h_action[f], (size_t)p->act.row_bytes * N, cudaHostAllocDefault);
    if (e == cudaSuccess)
      e = cudaHostAlloc(reinterpret_cast<void**>(&p->h_ids[f]), 4 * (size_t)N,
                        cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->h_stage_ev[f], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaStreamCreate(&p->stream);  // blocking: ordered with the legacy default stream (torch interop)
  if (e != cudaSuccess) {
    std::string msg = std::string("device allocation: ") + cudaGetErrorString(e);
    epb_destroy(p);
    return fail(EPB_ERR_CUDA, msg);
  }
  p->batch = cfg->batch_size > 0 ? cfg->batch_size : p->N;
  p->arange.resize(p->N);
  for (int i = 0; i < p->N; ++i) p->arange[i] = i;
  char* blob = static_cast<char*>(p->d_state_blob);
  p->sv.n_envs = p->N;
  p->sv.max_steps = cfg->max_episode_steps > 0 ? cfg->max_episode_steps : INT_MAX;
  p->sv.env_id_offset = cfg->env_id_offset;
  p->sv.iopt = iopt;
  p->sv.flags = reinterpret_cast<int32_t*>(blob + o_flags);
  p->sv.mt_idx = reinterpret_cast<int32_t*>(blob + o_idx);
  p->sv.istate = reinterpret_cast<int32_t*>(blob + o_ist);
  p->sv.rstate = blob + o_rst;
  p->sv.mt = reinterpret_cast<uint32_t*>(blob + o_mt);

  if (kind <= EPB_MOUNTAIN_CAR_CONTINUOUS) {
    p->step_fn = classic_step_fn(kind, p->precision);
    p->rollout_fn = classic_rollout_fn(kind, p->precision);
  } else if (kind <= EPB_BLACKJACK) {
    p->step_fn = toytext_step_fn(kind, iopt);
    p->rollout_fn = toytext_rollout_fn(kind, iopt);
  }

  // seed on device
  int32_t* d_env_seed = nullptr;
  if (cfg->env_seed) {
    e = cudaMalloc(reinterpret_cast<void**>(&d_env_seed), 4 * (size_t)N);
    if (e == cudaSuccess)
      e = cudaMemcpy(d_env_seed, cfg->env_seed, 4 * (size_t)N, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    seed_kernel<<<(p->N + 127) / 128, 128, 0, p->stream>>>(p->sv, cfg->seed, d_env_seed);
    e = cudaGetLastError();
    ++p->launches;
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
  if (d_env_seed) cudaFree(d_env_seed);
  if (e != cudaSuccess) {
    std::string msg = std::string("seeding: ") + cudaGetErrorString(e);
    epb_destroy(p);
    return fail(EPB_ERR_CUDA, msg);
  }
  // algorithmic bytes per env-step of the single-step kernel (identity env_ids):
  // action + 2 x (flags + env state) + every output column (+ RNG traffic for per-step RNG)
  int b = p->act.row_bytes + 2 * (4 + p->NR * p->real_size + p->NI * 4);
  for (const Key& k : p->keys) b += k.row_bytes;
  if (kind == EPB_FROZEN_LAKE || (kind == EPB_CLIFF_WALKING && iopt)) b += 16 + 8;
  if (kind == EPB_NCHAIN) b += 32 + 8;
  if (kind == EPB_HALF_CHEETAH) b -= 2 * (32 - 27) * 8;  // 27 of the 32-double record are live
  p->bytes_per_step = b;
  *out = p;
  return EPB_OK;
}

int epb_destroy(epb_pool* p) {
  if (!p) return EPB_OK;
  DeviceGuard guard(p->cfg.device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  for (Pending& pd : p->pending)
    if (!pd.ready) cudaEventDestroy(pd.ev);
  for (cudaEvent_t ev : p->free_events) cudaEventDestroy(ev);
  for (void* s : p->all_slabs) cudaFreeHost(s);
  for (int f = 0; f < 2; ++f) {
    if (p->h_action[f]) cudaFreeHost(p->h_action[f]);
    if (p->h_ids[f]) cudaFreeHost(p->h_ids[f]);
    if (p->h_stage_ev[f]) cudaEventDestroy(p->h_stage_ev[f]);
  }
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.exec);
  if (p->mjc) mjc_pool_destroy(p->mjc);
  for (int g = 0; g < kMaxPeers; ++g)
    if (p->x_ipc[g] && p->x_peer[g]) cudaIpcCloseMemHandle(p->x_peer[g]);
  if (p->x_base) cudaFree(p->x_base);
  if (p->d_state_blob) cudaFree(p->d_state_blob);
  if (p->d_slab) cudaFree(p->d_slab);
  if (p->d_action) cudaFree(p->d_action);
  if (p->d_ids) cudaFree(p->d_ids);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return EPB_OK;
}

int epb_num_state_keys(const epb_pool* p) { return p ? (int)p->keys.size() : 0; }
int epb_num_envs(const epb_pool* p) { return p ? p->N : 0; }
int64_t epb_slab_bytes(const epb_pool* p) { return p ? p->slab_bytes : 0; }

static void fill_info(const Key& k, epb_key_info* out) {
  out->name = k.name;
  out->dtype = k.dtype;
  out->ndim = k.ndim;
  for (int i = 0; i < 3; ++i) out->shape[i] = k.shape[i];
  out->row_bytes = k.row_bytes;
  out->slab_offset = k.off;
}
int epb_state_key(const epb_pool* p, int k, epb_key_info* out) {
  if (!p || !out || k < 0 || k >= (int)p->keys.size()) return fail(EPB_ERR_INVALID, "bad key index");
  fill_info(p->keys[k], out);
  return EPB_OK;
}
int epb_action_key(const epb_pool* p, epb_key_info* out) {
  if (!p || !out) return fail(EPB_ERR_INVALID, "null argument");
  fill_info(p->act, out);
  return EPB_OK;
}

int epb_send(epb_pool* p, const void* action, const int32_t* env_ids, int n) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  return host_submit(p, action, env_ids, n, 0);
}
int epb_reset(epb_pool* p, const int32_t* env_ids, int n) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  return host_submit(p, nullptr, env_ids, n, 1);
}

namespace {
// slab reference counting: one reference held by the pending queue while rows remain, one
// per outstanding recv lease.  Caller holds p->mu.
void slab_ref(epb_pool* p, void* slab, int delta) {
  for (size_t i = 0; i < p->leases.size(); ++i) {
    if (p->leases[i].first == slab) {
      p->leases[i].second += delta;
      if (p->leases[i].second <= 0) {
        p->leases.erase(p->leases.begin() + i);
        p->free_slabs.push_back(slab);
      }
      return;
    }
  }
  if (delta > 0) p->leases.emplace_back(slab, delta);
}
}  // namespace

// Sync mode (batch == num_envs): hands out the oldest send/reset as a whole (its n rows,
// partial-id sends included).  Async mode: exactly `batch` rows per call, in submission
// order -- on the GPU every env of a send finishes together, so "the first batch_size envs to
// finish" (state_buffer_queue.h:148-163) is the submission order.
int epb_recv_slab_ex(epb_pool* p, void** slab, int* row0, int* n_rows) {
  if (!p || !slab || !n_rows || !row0) return fail(EPB_ERR_INVALID, "null argument");
  const bool async = p->batch < p->N;
  std::unique_lock<std::mutex> lk(p->mu);
  if (p->pending.empty()) return fail(EPB_ERR_STATE, "recv without an outstanding send/reset");
  auto wait_head = [&](Pending& pd) -> cudaError_t {
    if (pd.ready) return cudaSuccess;
    cudaEvent_t ev = pd.ev;
    lk.unlock();
    cudaError_t e = cudaEventSynchronize(ev);
    lk.lock();
    pd.ready = true;
    p->free_events.push_back(ev);
    return e;
  };
  Pending& head = p->pending.front();
  cudaError_t e = wait_head(head);
  if (e != cudaSuccess) return fail(EPB_ERR_CUDA, std::string("recv: ") + cudaGetErrorString(e));
  const int want = async ? p->batch : head.n - head.row0;
  if (head.n - head.row0 >= want) {
    *slab = head.slab;
    *row0 = head.row0;
    *n_rows = want;
    slab_ref(p, head.slab, +1);
    head.row0 += want;
    if (head.row0 == head.n) {
      slab_ref(p, head.slab, -1);  // queue reference
      p->pending.pop_front();
    }
    return EPB_OK;
  }
  // async batch straddles several sends: assemble it in a fresh slab (host memcpy)
  int have = 0;
  for (const Pending& pd : p->pending) have += pd.n - pd.row0;
  if (have < want) return fail(EPB_ERR_STATE, "recv: fewer than batch_size envs outstanding");
  void* dst = nullptr;
  lk.unlock();
  int rc = get_slab(p, &dst);
  lk.lock();
  if (rc != EPB_OK) return rc;
  int filled = 0;
  while (filled < want) {
    Pending& pd = p->pending.front();
    e = wait_head(pd);
    if (e != cudaSuccess) {
      p->free_slabs.push_back(dst);  // nothing leased yet: hand the assembly slab back
      return fail(EPB_ERR_CUDA, std::string("recv: ") + cudaGetErrorString(e));
    }
    int take = pd.n - pd.row0;
    if (take > want - filled) take = want - filled;
    for (const Key& k : p->keys)
      memcpy(static_cast<char*>(dst) + k.off + (size_t)filled * k.row_bytes,
             static_cast<char*>(pd.slab) + k.off + (size_t)pd.row0 * k.row_bytes,
             (size_t)take * k.row_bytes);
    pd.row0 += take;
    filled += take;
    if (pd.row0 == pd.n) {
      slab_ref(p, pd.slab, -1);
      p->pending.pop_front();
    }
  }
  slab_ref(p, dst, +1);
  *slab = dst;
  *row0 = 0;
  *n_rows = want;
  return EPB_OK;
}
int epb_recv_slab(epb_pool* p, void** slab, int* n_rows) {
  int row0 = 0;
  int rc = epb_recv_slab_ex(p, slab, &row0, n_rows);
  if (rc == EPB_OK && row0 != 0) {
    // plain variant cannot express a row offset: only valid in sync mode
    epb_release_slab(p, *slab);
    return fail(EPB_ERR_STATE, "use epb_recv_slab_ex in async mode");
  }
  return rc;
}
int epb_release_slab(epb_pool* p, void* slab) {
  if (!p || !slab) return fail(EPB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(p->mu);
  slab_ref(p, slab, -1);
  return EPB_OK;
}
int epb_recv(epb_pool* p, void* const* cols, int* n_rows) {
  void* slab = nullptr;
  int n = 0, row0 = 0;
  int rc = epb_recv_slab_ex(p, &slab, &row0, &n);
  if (rc != EPB_OK) return rc;
  if (cols) {
    for (size_t k = 0; k < p->keys.size(); ++k) {
      if (cols[k]) memcpy(cols[k], static_cast<char*>(slab) + p->keys[k].off +
                                       (size_t)row0 * p->keys[k].row_bytes,
                          (size_t)p->keys[k].row_bytes * n);
    }
  }
  if (n_rows) *n_rows = n;
  return epb_release_slab(p, slab);
}

int epb_step_device(epb_pool* p, const void* d_action, const int32_t* d_env_ids, int n,
                    void* stream) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  if (n <= 0 || n > p->N) return fail(EPB_ERR_INVALID, "batch rows must be in [1, num_envs]");
  if (!d_action) return fail(EPB_ERR_INVALID, "action is NULL");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->stream;
  return launch_batch(p, d_action, d_env_ids, n, 0, p->d_slab, s);
}
int epb_reset_device(epb_pool* p, const int32_t* d_env_ids, int n, void* stream) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  if (n <= 0 || n > p->N) return fail(EPB_ERR_INVALID, "batch rows must be in [1, num_envs]");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->stream;
  return launch_batch(p, nullptr, d_env_ids, n, 1, p->d_slab, s);
}
int epb_outputs_device(const epb_pool* p, void** d_slab) {
  if (!p || !d_slab) return fail(EPB_ERR_INVALID, "null argument");
  *d_slab = p->d_last ? p->d_last : p->d_slab;
  return EPB_OK;
}

int epb_rollout_device(epb_pool* p, const void* d_actions, int T, void* const* d_cols,
                       void* stream) {
  if (!p || !d_actions || !d_cols) return fail(EPB_ERR_INVALID, "null argument");
  if (T <= 0) return fail(EPB_ERR_INVALID, "T must be positive");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->stream;
  OutView ov{};
  ov.env_id = static_cast<int32_t*>(d_cols[0]);
  ov.players_id = static_cast<int32_t*>(d_cols[1]);
  ov.elapsed = static_cast<int32_t*>(d_cols[2]);
  ov.done = static_cast<uint8_t*>(d_cols[3]);
  ov.reward = static_cast<float*>(d_cols[4]);
  ov.discount = static_cast<float*>(d_cols[5]);
  ov.step_type = static_cast<int32_t*>(d_cols[6]);
  ov.trunc = static_cast<uint8_t*>(d_cols[7]);
  for (size_t k = 8; k < p->keys.size(); ++k) ov.env[k - 8] = d_cols[k];
  ov.t_stride_rows = p->N;
  if (p->kind == EPB_HALF_CHEETAH) {
    EPB_CUDA(mjc_launch_rollout(p->mjc, p->sv, ov, static_cast<const double*>(d_actions), T, s));
    ++p->launches;
    return EPB_OK;
  }
  LaunchArgs a{};
  a.sv = p->sv;
  a.ov = ov;
  a.action = d_actions;
  a.n = p->N;
  a.T = T;
  a.stream = s;
  EPB_CUDA(p->rollout_fn(a));
  ++p->launches;
  return EPB_OK;
}

int epb_step_many_device(epb_pool* p, const void* d_actions, int T_stream, int t0, int K,
                         int use_graph, void* stream) {
  if (!p || !d_actions) return fail(EPB_ERR_INVALID, "null argument");
  if (T_stream <= 0 || K <= 0 || t0 < 0) return fail(EPB_ERR_INVALID, "bad step-chain shape");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->stream;
  const size_t row = (size_t)p->act.row_bytes * p->N;
  auto chain = [&](cudaStream_t st) -> int {
    for (int k = 0; k < K; ++k) {
      const char* a = static_cast<const char*>(d_actions) + row * ((t0 + k) % T_stream);
      int rc = launch_batch(p, a, nullptr, p->N, 0, p->d_slab, st);
      if (rc != EPB_OK) return rc;
    }
    return EPB_OK;
  };
  if (!use_graph) return chain(s);
  cudaGraphExec_t exec = nullptr;
  for (const auto& g : p->graphs)
    if (g.actions == d_actions && g.T == T_stream && g.t0 == t0 && g.K == K && g.stream == s)
      exec = g.exec;
  if (!exec) {
    if (p->graphs.size() >= 4) {
      cudaGraphExecDestroy(p->graphs.back().exec);
      p->graphs.pop_back();
    }
    cudaGraph_t g = nullptr;
    EPB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int64_t before = p->launches;
    int rc = chain(s);
    p->launches = before;  // capture records, it does not launch
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (rc != EPB_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (e != cudaSuccess) return fail(EPB_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(EPB_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    p->graphs.insert(p->graphs.begin(), epb_pool::GraphEntry{exec, d_actions, T_stream, t0, K, s});
  }
  EPB_CUDA(cudaGraphLaunch(exec, s));
  p->launches += K;
  return EPB_OK;
}

// ---- peer exchange ---------------------------------------------------------------------
int epb_exchange_init(epb_pool* p, int world, int rank, void* ipc_handle_out) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(EPB_ERR_INVALID, "exchange: world must be in [1,16] and rank in [0,world)");
  if (p->x_base) return fail(EPB_ERR_STATE, "exchange already initialised");
  static_assert(sizeof(cudaIpcMemHandle_t) == EPB_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  p->x_flags_off = 2 * (int64_t)world * p->slab_bytes;
  p->x_ctl_off = p->x_flags_off + 256;
  p->x_bytes = p->x_ctl_off + 256 + ((2 * (int64_t)sizeof(PeerView) + 255) / 256) * 256;
  EPB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->x_base), (size_t)p->x_bytes));
  EPB_CUDA(cudaMemset(p->x_base, 0, (size_t)p->x_bytes));
  EPB_CUDA(cudaDeviceSynchronize());
  p->x_world = world;
  p->x_rank = rank;
  p->x_peer[rank] = p->x_base;
  // HalfCheetah's kernels have no forwarding epilogue; ENVPOOL_B200_EXCHANGE=push is the A/B switch
  const char* mode = getenv("ENVPOOL_B200_EXCHANGE");
  p->x_fused = p->kind != EPB_HALF_CHEETAH && !(mode && strcmp(mode, "push") == 0);
  if (const char* to = getenv("ENVPOOL_B200_EXCHANGE_TIMEOUT_S")) {
    double sec = atof(to);
    if (sec > 0) p->x_timeout_ns = (long long)(sec * 1e9);
  }
  if (world == 1) {
    int rc = upload_views(p);
    if (rc != EPB_OK) return rc;
  }
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    EPB_CUDA(cudaIpcGetMemHandle(&h, p->x_base));
    memcpy(ipc_handle_out, &h, sizeof(h));
  }
  return EPB_OK;
}
int epb_exchange_base(const epb_pool* p, void** base, int64_t* bytes) {
  if (!p || !base) return fail(EPB_ERR_INVALID, "null argument");
  if (!p->x_base) return fail(EPB_ERR_STATE, "exchange not initialised");
  *base = p->x_base;
  if (bytes) *bytes = p->x_bytes;
  return EPB_OK;
}
int epb_exchange_attach(epb_pool* p, void* const* peer_bases) {
  if (!p || !peer_bases) return fail(EPB_ERR_INVALID, "null argument");
  if (!p->x_base) return fail(EPB_ERR_STATE, "exchange not initialised");
  for (int g = 0; g < p->x_world; ++g) {
    if (g == p->x_rank) continue;
    if (!peer_bases[g]) return fail(EPB_ERR_INVALID, "exchange: null peer base");
    p->x_peer[g] = static_cast<char*>(peer_bases[g]);
  }
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  return upload_views(p);
}
int epb_exchange_attach_ipc(epb_pool* p, const void* ipc_handles) {
  if (!p || !ipc_handles) return fail(EPB_ERR_INVALID, "null argument");
  if (!p->x_base) return fail(EPB_ERR_STATE, "exchange not initialised");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  for (int g = 0; g < p->x_world; ++g) {
    if (g == p->x_rank || p->x_ipc[g]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(ipc_handles) + (size_t)g * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    EPB_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->x_peer[g] = static_cast<char*>(ptr);
    p->x_ipc[g] = true;
  }
  return upload_views(p);
}
int epb_step_exchange_device(epb_pool* p, const void* d_action, void* stream) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  if (!p->x_attached) return fail(EPB_ERR_STATE, "exchange: peers not attached");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->stream;
  const int parity = (int)(p->x_steps & 1);
  char* mine = p->x_base + p->x_mine(parity);
  const int force = d_action ? 0 : 1;
  if (p->x_fused) {
    int rc = launch_batch(p, d_action, nullptr, p->N, force, mine, s, p->x_view(parity));
    if (rc != EPB_OK) return rc;
  } else {
    int rc = launch_batch(p, d_action, nullptr, p->N, force, mine, s);
    if (rc != EPB_OK) return rc;
    const int64_t n16 = p->slab_bytes / 16;
    int64_t blocks = (n16 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    push_kernel<<<(unsigned)blocks, 256, 0, s>>>(p->x_view(parity), n16);
    EPB_CUDA(cudaGetLastError());
    ++p->launches;
  }
  ++p->x_steps;
  return EPB_OK;
}
int epb_exchange_wait(epb_pool* p, void* stream, void** d_gathered) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  if (!p->x_attached || p->x_steps == 0)
    return fail(EPB_ERR_STATE, "exchange: nothing has been exchanged yet");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : p->stream;
  wait_kernel<<<1, 32, 0, s>>>(
      reinterpret_cast<const unsigned long long*>(p->x_base + p->x_flags_off), p->x_world,
      p->x_ctl(), p->x_timeout_ns);
  EPB_CUDA(cudaGetLastError());
  ++p->launches;
  if (d_gathered)
    *d_gathered = p->x_base + (int64_t)((p->x_steps - 1) & 1) * p->x_world * p->slab_bytes;
  return EPB_OK;
}
int epb_exchange_status(epb_pool* p, int64_t* steps_pushed, int* timed_out) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  if (!p->x_base) return fail(EPB_ERR_STATE, "exchange not initialised");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  ExchangeCtl c{};
  EPB_CUDA(cudaMemcpy(&c, p->x_base + p->x_ctl_off, sizeof(c), cudaMemcpyDeviceToHost));
  if (steps_pushed) *steps_pushed = (int64_t)c.seq;
  if (timed_out) *timed_out = c.error;
  return EPB_OK;
}

int epb_sync(epb_pool* p) {
  if (!p) return fail(EPB_ERR_INVALID, "null pool");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  EPB_CUDA(cudaStreamSynchronize(p->stream));
  return EPB_OK;
}
void* epb_stream(epb_pool* p) { return p ? static_cast<void*>(p->stream) : nullptr; }

int64_t epb_state_bytes(const epb_pool* p) { return p ? p->state_bytes : 0; }
int epb_state_layout(const epb_pool* p, int64_t* out) {
  if (!p || !out) return fail(EPB_ERR_INVALID, "null argument");
  const char* blob = static_cast<const char*>(p->d_state_blob);
  out[0] = reinterpret_cast<const char*>(p->sv.flags) - blob;
  out[1] = reinterpret_cast<const char*>(p->sv.mt_idx) - blob;
  out[2] = reinterpret_cast<const char*>(p->sv.istate) - blob;
  out[3] = static_cast<const char*>(p->sv.rstate) - blob;
  out[4] = reinterpret_cast<const char*>(p->sv.mt) - blob;
  out[5] = p->NI;
  out[6] = p->NR;
  out[7] = p->real_size;
  return EPB_OK;
}
int epb_state_export(epb_pool* p, void* host_dst) {
  if (!p || !host_dst) return fail(EPB_ERR_INVALID, "null argument");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  EPB_CUDA(cudaStreamSynchronize(p->stream));
  EPB_CUDA(cudaMemcpy(host_dst, p->d_state_blob, (size_t)p->state_bytes, cudaMemcpyDeviceToHost));
  return EPB_OK;
}
int epb_state_import(epb_pool* p, const void* host_src) {
  if (!p || !host_src) return fail(EPB_ERR_INVALID, "null argument");
  DeviceGuard guard(p->cfg.device);
  EPB_CUDA(guard.status);
  EPB_CUDA(cudaStreamSynchronize(p->stream));
  EPB_CUDA(cudaMemcpy(p->d_state_blob, host_src, (size_t)p->state_bytes, cudaMemcpyHostToDevice));
  return EPB_OK;
}

int64_t epb_launch_count(const epb_pool* p) { return p ? p->launches : 0; }
int epb_bytes_per_env_step(const epb_pool* p) { return p ? p->bytes_per_step : 0; }

}  // extern "C"
