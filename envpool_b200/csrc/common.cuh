// envpool_b200 device-side core: SoA pool views, the device std::mt19937 + libstdc++
// distribution recipes, and the generic thread-per-env step / rollout kernels that every
// classic_control and toy_text env plugs into.
//
// What this replaces in the reference (paths relative to /root/reference/envpool/):
//   core/async_envpool.h:59-82   Send: batch action -> per-env ActionSlice   (action gather)
//   core/async_envpool.h:118-131 worker loop: auto-reset decision + EnvStep
//   core/env.h:184-256           EnvStep / PreProcess / Allocate (common columns)
//   core/state_buffer.h:81-131   row reservation + completion counting         (obs scatter)
// On the GPU all of it is one kernel: row i of the batch is one thread, the env's state is
// a column of SoA arrays in HBM, outputs are written straight into the packed output slab.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace epb {

constexpr int kMtN = 624;
constexpr int kMtM = 397;
constexpr int kBlock = 128;  // threads per CTA for thread-per-env kernels

// Persistent per-env state, structure-of-arrays over the local env index.
struct StateView {
  int32_t n_envs;         // N: SoA stride (envs owned by this pool)
  int32_t max_steps;      // config max_episode_steps (INT_MAX when unset)
  int32_t env_id_offset;  // global id of local env 0
  int32_t iopt;           // env option (size / version / is_slippery / natural|sab<<1)
  int32_t* flags;         // (current_step_ << 1) | done_      (env.h:81, cartpole.h:67)
  uint32_t* mt;           // [624][N] mt19937 words
  int32_t* mt_idx;        // [N] index of the next word to regenerate, 0..623
  void* rstate;           // [NR][N] real state (double or float)
  int32_t* istate;        // [NI][N] integer state
};

// Output columns for one batch (pointers into a packed slab or caller arrays).
// Order = the reference's state-key order (env_spec.h:37-43), then env keys.
struct OutView {
  int32_t* env_id;       // info:env_id
  int32_t* players_id;   // info:players.env_id
  int32_t* elapsed;      // elapsed_step
  uint8_t* done;         // done
  float* reward;         // reward
  float* discount;       // discount
  int32_t* step_type;    // step_type
  uint8_t* trunc;        // trunc
  void* env[5];          // env-specific keys in declaration order
  int64_t t_stride_rows; // rollout: rows between consecutive time steps (= N)
};

// ---------------------------------------------------------------------------------------
// std::mt19937 on the device.  The table lives in HBM as mt[k*N + env] so that envs that
// are in lockstep read/write coalesced rows.  Words are regenerated one at a time (the
// block "twist" of the textbook implementation unrolled in time -- identical sequence,
// 12 B read + 4 B written per draw instead of a 2.5 KB burst).
struct Mt {
  uint32_t* base;  // &mt[eid]
  int64_t stride;  // N
  int idx;
  __device__ __forceinline__ Mt(const StateView& sv, int eid)
      : base(sv.mt + eid), stride(sv.n_envs), idx(sv.mt_idx[eid]) {}
  // idx already loaded by the caller (issued together with the env-state loads so the
  // draw does not pay a second dependent round trip)
  __device__ __forceinline__ Mt(const StateView& sv, int eid, int idx_)
      : base(sv.mt + eid), stride(sv.n_envs), idx(idx_) {}
  __device__ __forceinline__ uint32_t next() {
    int i = idx;
    int i1 = (i + 1 == kMtN) ? 0 : i + 1;
    int im = (i + kMtM >= kMtN) ? i + kMtM - kMtN : i + kMtM;
    uint32_t a = base[(int64_t)i * stride];
    uint32_t b = base[(int64_t)i1 * stride];
    uint32_t c = base[(int64_t)im * stride];
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    uint32_t v = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    base[(int64_t)i * stride] = v;
    idx = i1;
    v ^= (v >> 11);
    v ^= (v << 7) & 0x9d2c5680u;
    v ^= (v << 15) & 0xefc60000u;
    v ^= (v >> 18);
    return v;
  }
  __device__ __forceinline__ void save(const StateView& sv, int eid) { sv.mt_idx[eid] = idx; }

  // K consecutive draws in ONE memory round trip.  Regenerating word i needs the current
  // words i, i+1 and i+397; for K <= 226 none of those is itself regenerated earlier in the
  // same batch except word i+1, whose *old* value is what the recurrence wants -- so all
  // 2K+1 loads can be issued together (memory-level parallelism instead of K dependent
  // DRAM latencies), then the K new words are stored.  Same sequence as K next() calls.
  template <int K>
  __device__ __forceinline__ void next_batch(uint32_t (&out)[K]) {
    static_assert(K >= 1 && K <= 64, "batch too large");
    uint32_t w[K + 1], m[K];
    const int i = idx;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
      int j = i + k;
      j = j >= kMtN ? j - kMtN : j;
      w[k] = base[(int64_t)j * stride];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      int j = i + k + kMtM;
      j = j >= kMtN ? j - kMtN : j;
      j = j >= kMtN ? j - kMtN : j;
      m[k] = base[(int64_t)j * stride];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      uint32_t y = (w[k] & 0x80000000u) | (w[k + 1] & 0x7fffffffu);
      uint32_t v = m[k] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      int j = i + k;
      j = j >= kMtN ? j - kMtN : j;
      base[(int64_t)j * stride] = v;
      v ^= (v >> 11);
      v ^= (v << 7) & 0x9d2c5680u;
      v ^= (v << 15) & 0xefc60000u;
      v ^= (v >> 18);
      out[k] = v;
    }
    int j = i + K;
    idx = j >= kMtN ? j - kMtN : j;
  }

  // std::generate_canonical<double,53> (libstdc++ 13 bits/random.tcc:3349-3381) from two
  // engine outputs.  Explicit _rn intrinsics: never contracted into an FMA, whatever the
  // TU's flags.
  static __device__ __forceinline__ double canonical_from(uint32_t d1, uint32_t d2) {
    double sum = __dadd_rn((double)d1, __dmul_rn((double)d2, 4294967296.0));
    double ret = __dmul_rn(sum, 5.421010862427522170037e-20);  // exact: / 2^64
    if (ret >= 1.0) ret = 0.99999999999999988897769753748;     // nextafter(1,0)
    return ret;
  }
  __device__ __forceinline__ double canonical() {
    uint32_t d[2];
    next_batch<2>(d);
    return canonical_from(d[0], d[1]);
  }
  // std::uniform_real_distribution<double>(a,b)
  __device__ __forceinline__ double uniform_real(double a, double b) {
    return __dadd_rn(__dmul_rn(canonical(), __dsub_rn(b, a)), a);
  }
  // NC consecutive uniform_real(a,b) draws, one memory round trip
  template <int NC>
  __device__ __forceinline__ void uniform_real_batch(double a, double b, double (&out)[NC]) {
    uint32_t d[2 * NC];
    next_batch<2 * NC>(d);
#pragma unroll
    for (int k = 0; k < NC; ++k)
      out[k] = __dadd_rn(__dmul_rn(canonical_from(d[2 * k], d[2 * k + 1]), __dsub_rn(b, a)), a);
  }
  // std::uniform_int_distribution<int>(a,b): Lemire (bits/uniform_int_dist.h:252-282)
  __device__ __forceinline__ int uniform_int(int a, int b) {
    uint32_t range = (uint32_t)b - (uint32_t)a + 1u;
    uint64_t product = (uint64_t)next() * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
      uint32_t threshold = (0u - range) % range;
      while (low < threshold) {
        product = (uint64_t)next() * (uint64_t)range;
        low = (uint32_t)product;
      }
    }
    return a + (int)(product >> 32);
  }
};

// K uniform_int draws with the engine words fetched in one round trip.  Lemire's method
// almost never rejects (probability range/2^32 per draw); when it does, the extra words
// come from sequential next() calls after the batch, so the word order is unchanged.
template <int K>
struct MtIntBatch {
  Mt& rng;
  uint32_t words[K];
  int cursor;
  __device__ __forceinline__ explicit MtIntBatch(Mt& r) : rng(r), cursor(0) {
    rng.template next_batch<K>(words);
  }
  __device__ __forceinline__ uint32_t word() {
    uint32_t v = 0;
    bool hit = false;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (k == cursor) {
        v = words[k];
        hit = true;
      }
    ++cursor;
    return hit ? v : rng.next();
  }
  __device__ __forceinline__ int uniform_int(int a, int b) {
    uint32_t range = (uint32_t)b - (uint32_t)a + 1u;
    uint64_t product = (uint64_t)word() * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
      uint32_t threshold = (0u - range) % range;
      while (low < threshold) {
        product = (uint64_t)word() * (uint64_t)range;
        low = (uint32_t)product;
      }
    }
    return a + (int)(product >> 32);
  }
};

// Seeds every env's table: std::mt19937(seed) == init_genrand (env.h:113).  One thread per
// env; each of the 624 steps is a coalesced row write.
__global__ void seed_kernel(StateView sv, int base_seed, const int32_t* env_seed);

// ---------------------------------------------------------------------------------------
// Common output columns: Env::Allocate, core/env.h:224-256.
__device__ __forceinline__ void write_common(const OutView& ov, int64_t row, int global_eid,
                                             int cur, int done, float reward,
                                             int max_steps) {
  int step_type = (cur == 0) ? 0 : (done ? 2 : 1);
  if (ov.env_id) ov.env_id[row] = global_eid;
  if (ov.players_id) ov.players_id[row] = global_eid;
  if (ov.elapsed) ov.elapsed[row] = cur;
  if (ov.done) ov.done[row] = (uint8_t)done;
  if (ov.reward) ov.reward[row] = reward;
  if (ov.discount) ov.discount[row] = done ? 0.0f : 1.0f;
  if (ov.step_type) ov.step_type[row] = step_type;
  if (ov.trunc) ov.trunc[row] = (uint8_t)(done && (cur >= max_steps));
}

// Per-env result of one EnvStep, kept in registers until the output write.
struct StepOut {
  float reward;
  float extra;  // env-specific scalar (CliffWalking info:prob)
};

// The Env concept every family member implements:
//   using Act = <action scalar type>;  struct State {...};
//   static void load(const StateView&, int eid, State&);
//   static void store(const StateView&, int eid, const State&);
//   static void reset(const StateView&, State&, Mt*, StepOut&);            // XxxEnv::Reset
//   static void step(const StateView&, State&, Act, int cur, int& done, Mt*, StepOut&);
//   static void write_obs(const StateView&, const OutView&, int64_t row, const State&,
//                         const StepOut&);
//   static constexpr bool kRngInReset, kRngInStep;  (Mt* is NULL when false)
//   static constexpr bool kBlockObs;                (block-cooperative obs write)
//
// One EnvStep (env.h:184-222) incl. the auto-reset decision (async_envpool.h:127).
template <class Env>
__device__ __forceinline__ void env_step(const StateView& sv, int eid, int& flags,
                                         typename Env::State& s, typename Env::Act a,
                                         bool force_reset, StepOut& so, int& mt_idx) {
  int done = flags & 1;
  int cur = flags >> 1;
  bool reset = force_reset || done;
  if (reset) {
    cur = 0;
    done = 0;
    if (Env::kRngInReset) {
      Mt rng(sv, eid, mt_idx);
      Env::reset(sv, s, &rng, so);
      mt_idx = rng.idx;
    } else {
      Env::reset(sv, s, nullptr, so);
    }
  } else {
    ++cur;
    if (Env::kRngInStep) {
      Mt rng(sv, eid, mt_idx);
      Env::step(sv, s, a, cur, done, &rng, so);
      mt_idx = rng.idx;
    } else {
      Env::step(sv, s, a, cur, done, nullptr, so);
    }
  }
  flags = (cur << 1) | done;
}

// Single sync step of a batch: thread `row` handles env env_ids[row] (identity if NULL).
template <class Env>
__global__ void __launch_bounds__(kBlock)
step_kernel(StateView sv, OutView ov, const typename Env::Act* __restrict__ action,
            const int32_t* __restrict__ env_ids, int n, int force_reset) {
  int row = blockIdx.x * kBlock + threadIdx.x;
  bool active = row < n;
  typename Env::State s;
  StepOut so;
  so.reward = 0.f;
  so.extra = 0.f;
  int eid = 0, flags = 0;
  typename Env::Act a = typename Env::Act();
  if (active) {
    // inputs no earlier kernel of the stream writes: fetch them before the grid dependency
    eid = env_ids ? env_ids[row] : row;
    if (!force_reset) a = action[row];
  }
  // Programmatic dependent launch: let the next step's grid start launching now, and wait
  // here for the previous step's grid (it owns the state and the output slab until done).
  // Both are no-ops when the kernel is launched without the PDL attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (active) {
    constexpr bool kRng = Env::kRngInReset || Env::kRngInStep;
    flags = sv.flags[eid];
    int mt_idx = kRng ? sv.mt_idx[eid] : 0;
    const int mt_idx0 = mt_idx;
    Env::load(sv, eid, s);
    env_step<Env>(sv, eid, flags, s, a, force_reset != 0, so, mt_idx);
    Env::store(sv, eid, s);
    sv.flags[eid] = flags;
    if (kRng && mt_idx != mt_idx0) sv.mt_idx[eid] = mt_idx;
    write_common(ov, row, eid + sv.env_id_offset, flags >> 1, flags & 1, so.reward,
                 sv.max_steps);
  }
  if constexpr (Env::kBlockObs) {
    Env::block_write_obs(ov, (int64_t)blockIdx.x * kBlock, n, active, s, so);
  } else if (active) {
    Env::write_obs(sv, ov, row, s, so);
  }
}

// Fused rollout: T sync steps of all N envs in one launch; state stays in registers, the
// action stream [T,N] is read and the outputs [T,N,...] written once each.
template <class Env>
__global__ void __launch_bounds__(kBlock)
rollout_kernel(StateView sv, OutView ov, const typename Env::Act* __restrict__ actions,
               int T) {
  int eid = blockIdx.x * kBlock + threadIdx.x;
  const int n = sv.n_envs;
  bool active = eid < n;
  typename Env::State s;
  int flags = 0, mt_idx = 0;
  constexpr bool kRng = Env::kRngInReset || Env::kRngInStep;
  if (active) {
    flags = sv.flags[eid];
    if (kRng) mt_idx = sv.mt_idx[eid];
    Env::load(sv, eid, s);
  }
  typename Env::Act a_next = typename Env::Act();
  if (active && T > 0) a_next = actions[eid];
  for (int t = 0; t < T; ++t) {
    StepOut so;
    so.reward = 0.f;
    so.extra = 0.f;
    int64_t row = (int64_t)t * ov.t_stride_rows + eid;
    typename Env::Act a = a_next;
    if (active && t + 1 < T) a_next = actions[(int64_t)(t + 1) * n + eid];  // prefetch
    if (active) {
      env_step<Env>(sv, eid, flags, s, a, false, so, mt_idx);
      write_common(ov, row, eid + sv.env_id_offset, flags >> 1, flags & 1, so.reward,
                   sv.max_steps);
    }
    if constexpr (Env::kBlockObs) {
      Env::block_write_obs(ov, (int64_t)t * ov.t_stride_rows + (int64_t)blockIdx.x * kBlock,
                           (int64_t)t * ov.t_stride_rows + n, active, s, so);
    } else if (active) {
      Env::write_obs(sv, ov, row, s, so);
    }
  }
  if (active) {
    Env::store(sv, eid, s);
    sv.flags[eid] = flags;
    if (kRng) sv.mt_idx[eid] = mt_idx;
  }
}

// Host-side launch table filled by each family's translation unit.
struct LaunchArgs {
  StateView sv;
  OutView ov;
  const void* action;
  const int32_t* env_ids;
  int n;
  int force_reset;
  int T;  // rollout only
  cudaStream_t stream;
};
typedef cudaError_t (*launch_fn)(const LaunchArgs&);

// ENVPOOL_B200_PDL=0 turns the programmatic-dependent-launch attribute off (A/B switch).
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("ENVPOOL_B200_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

template <class Env>
cudaError_t launch_step(const LaunchArgs& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((a.n + kBlock - 1) / kBlock);
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, step_kernel<Env>, a.sv, a.ov,
                            static_cast<const typename Env::Act*>(a.action), a.env_ids, a.n,
                            a.force_reset);
}
template <class Env>
cudaError_t launch_rollout(const LaunchArgs& a) {
  int grid = (a.sv.n_envs + kBlock - 1) / kBlock;
  rollout_kernel<Env><<<grid, kBlock, 0, a.stream>>>(
      a.sv, a.ov, static_cast<const typename Env::Act*>(a.action), a.T);
  return cudaGetLastError();
}

// family entry points (classic.cu / toytext.cu / mujoco.cu)
launch_fn classic_step_fn(int kind, int precision);
launch_fn classic_rollout_fn(int kind, int precision);
launch_fn toytext_step_fn(int kind, int iopt);
launch_fn toytext_rollout_fn(int kind, int iopt);

}  // namespace epb
