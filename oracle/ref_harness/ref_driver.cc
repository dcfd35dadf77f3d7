// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Thin C driver around the *unmodified* reference headers under /root/reference
// (compiled where they lie; no reference source is copied into this repo).  It
// instantiates the reference's own AsyncEnvPool<Env> (envpool/core/async_envpool.h)
// for the classic_control and toy_text envs and exposes Reset / Send / Recv through a
// flat C ABI so that Python (ctypes) can (1) dump golden trajectories into
// tests/golden/ and (2) time the reference CPU thread pool for bench.py's
// `--impl reference` arm and `cpu_baseline` (kind "reference").
//
// The drive pattern follows the reference's own C++ tests
// (envpool/dummy/dummy_envpool_test.cc:114-160,
//  envpool/mujoco/gym/mujoco_gym_envpool_test.cc:27-54).
//
// MuJoCo envs are NOT built here: MuJoCo 3.6.0 is an un-vendored third-party
// dependency (envpool/workspace0.bzl) and is absent from this image.
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "envpool/classic_control/acrobot.h"
#include "envpool/classic_control/cartpole.h"
#include "envpool/classic_control/mountain_car.h"
#include "envpool/classic_control/mountain_car_continuous.h"
#include "envpool/classic_control/pendulum.h"
#include "envpool/core/async_envpool.h"
#include "envpool/toy_text/blackjack.h"
#include "envpool/toy_text/catch.h"
#include "envpool/toy_text/cliffwalking.h"
#include "envpool/toy_text/frozen_lake.h"
#include "envpool/toy_text/nchain.h"
#include "envpool/toy_text/taxi.h"

// The step path never renders; the reference keeps the OpenCV drawing code in
// render_utils.cc, which is not compiled here.  Provide inert definitions.
namespace classic_control::rendering {
void RenderCartPole(double, double, int, int, unsigned char*) {}
void RenderPendulum(double, bool, double, int, int, unsigned char*) {}
void RenderMountainCar(double, double, int, int, unsigned char*) {}
void RenderAcrobot(double, double, int, int, unsigned char*) {}
}  // namespace classic_control::rendering

namespace {

struct RefPoolBase {
  virtual ~RefPoolBase() = default;
  virtual void Reset() = 0;
  virtual void Step(const void* action) = 0;
  virtual int NumKeys() const = 0;
  virtual std::size_t KeyBytes(int k) const = 0;
  virtual void Copy(int k, void* dst) const = 0;
  virtual double Bench(const void* actions, int steps_in_stream, int warmup,
                       int steps) = 0;
  int num_envs = 0;
};

template <typename Pool, typename ActT, int kActDim>
struct RefPool : RefPoolBase {
  using Spec = typename Pool::Spec;
  std::unique_ptr<Spec> spec;
  std::unique_ptr<Pool> pool;
  std::vector<Array> last;
  Array ids;

  static Array MakeIds(int n) {
    ::Spec<int> s(std::vector<int>{n});
    return Array(s);
  }
  template <typename Config>
  explicit RefPool(const Config& config)
      : ids(MakeIds(config["num_envs"_])) {
    spec = std::make_unique<Spec>(config);
    pool = std::make_unique<Pool>(*spec);
    num_envs = config["num_envs"_];
    for (int i = 0; i < num_envs; ++i) ids[i] = i;
  }
  void Reset() override {
    pool->Reset(ids);
    last = pool->Recv();
  }
  std::vector<Array> MakeAction(const void* action) {
    std::vector<int> shape = {num_envs};
    if (kActDim > 0) shape.push_back(kActDim);
    ::Spec<ActT> act_spec(shape);
    Array act(act_spec);
    std::memcpy(act.Data(), action,
                sizeof(ActT) * num_envs * (kActDim > 0 ? kActDim : 1));
    return {ids, ids, act};
  }
  void Step(const void* action) override {
    pool->Send(MakeAction(action));
    last = pool->Recv();
  }
  int NumKeys() const override { return static_cast<int>(last.size()); }
  std::size_t KeyBytes(int k) const override {
    return last[k].size * last[k].element_size;
  }
  void Copy(int k, void* dst) const override {
    std::memcpy(dst, last[k].Data(), KeyBytes(k));
  }
  // Protocol of benchmark/test_envpool.py:94-107 in sync mode: a fresh action
  // batch per step (cycled from a [steps_in_stream, N] stream), auto-reset on.
  double Bench(const void* actions, int steps_in_stream, int warmup,
               int steps) override {
    const std::size_t row =
        sizeof(ActT) * num_envs * (kActDim > 0 ? kActDim : 1);
    const char* base = static_cast<const char*>(actions);
    Reset();
    for (int t = 0; t < warmup; ++t) {
      pool->Send(MakeAction(base + row * (t % steps_in_stream)));
      last = pool->Recv();
    }
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < steps; ++t) {
      pool->Send(MakeAction(base + row * ((warmup + t) % steps_in_stream)));
      last = pool->Recv();
    }
    std::chrono::duration<double> dt = std::chrono::steady_clock::now() - t0;
    return dt.count();
  }
};

template <typename Pool, typename ActT, int kActDim, typename Tweak>
RefPoolBase* Make(int num_envs, int num_threads, int seed,
                  int max_episode_steps, Tweak tweak) {
  auto config = Pool::Spec::kDefaultConfig;
  config["num_envs"_] = num_envs;
  config["batch_size"_] = num_envs;
  config["num_threads"_] = num_threads;
  config["seed"_] = seed;
  if (max_episode_steps > 0) config["max_episode_steps"_] = max_episode_steps;
  tweak(config);
  return new RefPool<Pool, ActT, kActDim>(config);
}

}  // namespace

extern "C" {

// `task`: reference env class name without the "Env" suffix.  `iopt` carries the one
// integer option some envs have (FrozenLake size, Pendulum version, CliffWalking
// is_slippery, Blackjack natural | sab<<1); pass -1 for the reference default.
void* ref_create(const char* task, int num_envs, int num_threads, int seed,
                 int max_episode_steps, int iopt) {
  std::string t(task);
  auto none = [](auto&) {};
  try {
    if (t == "CartPole")
      return Make<classic_control::CartPoleEnvPool, int, 0>(
          num_envs, num_threads, seed, max_episode_steps, none);
    if (t == "Pendulum")
      return Make<classic_control::PendulumEnvPool, float, 1>(
          num_envs, num_threads, seed, max_episode_steps, [&](auto& c) {
            if (iopt >= 0) c["version"_] = iopt;
          });
    if (t == "Acrobot")
      return Make<classic_control::AcrobotEnvPool, int, 0>(
          num_envs, num_threads, seed, max_episode_steps, none);
    if (t == "MountainCar")
      return Make<classic_control::MountainCarEnvPool, int, 0>(
          num_envs, num_threads, seed, max_episode_steps, none);
    if (t == "MountainCarContinuous")
      return Make<classic_control::MountainCarContinuousEnvPool, float, 1>(
          num_envs, num_threads, seed, max_episode_steps, none);
    if (t == "FrozenLake")
      return Make<toy_text::FrozenLakeEnvPool, int, 0>(
          num_envs, num_threads, seed, max_episode_steps, [&](auto& c) {
            if (iopt >= 0) c["size"_] = iopt;
          });
    if (t == "Catch")
      return Make<toy_text::CatchEnvPool, int, 0>(
          num_envs, num_threads, seed, max_episode_steps, none);
    if (t == "Taxi")
      return Make<toy_text::TaxiEnvPool, int, 0>(num_envs, num_threads, seed,
                                                 max_episode_steps, none);
    if (t == "NChain")
      return Make<toy_text::NChainEnvPool, int, 0>(num_envs, num_threads, seed,
                                                   max_episode_steps, none);
    if (t == "CliffWalking")
      return Make<toy_text::CliffWalkingEnvPool, int, 0>(
          num_envs, num_threads, seed, max_episode_steps, [&](auto& c) {
            if (iopt >= 0) c["is_slippery"_] = (iopt != 0);
          });
    if (t == "Blackjack")
      return Make<toy_text::BlackjackEnvPool, int, 0>(
          num_envs, num_threads, seed, max_episode_steps, [&](auto& c) {
            if (iopt >= 0) {
              c["natural"_] = (iopt & 1) != 0;
              c["sab"_] = (iopt & 2) != 0;
            }
          });
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_create(%s): %s\n", task, e.what());
    return nullptr;
  }
  return nullptr;
}

void ref_destroy(void* h) { delete static_cast<RefPoolBase*>(h); }
void ref_reset(void* h) { static_cast<RefPoolBase*>(h)->Reset(); }
void ref_step(void* h, const void* action) {
  static_cast<RefPoolBase*>(h)->Step(action);
}
int ref_num_keys(void* h) { return static_cast<RefPoolBase*>(h)->NumKeys(); }
std::uint64_t ref_key_bytes(void* h, int k) {
  return static_cast<RefPoolBase*>(h)->KeyBytes(k);
}
void ref_copy(void* h, int k, void* dst) {
  static_cast<RefPoolBase*>(h)->Copy(k, dst);
}
// Returns elapsed seconds for `steps` timed Send/Recv pairs after `warmup`.
double ref_bench(void* h, const void* actions, int steps_in_stream, int warmup,
                 int steps) {
  return static_cast<RefPoolBase*>(h)->Bench(actions, steps_in_stream, warmup,
                                             steps);
}
int ref_hardware_concurrency() {
  return static_cast<int>(std::thread::hardware_concurrency());
}

}  // extern "C"
