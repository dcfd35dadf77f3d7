// TEST INFRASTRUCTURE ONLY.  The toolchain's own libstdc++ <random> behind a C API: the
// reference draws all of its randomness through std::mt19937 + std::uniform_int_distribution /
// std::uniform_real_distribution / std::normal_distribution (envpool/core/env.h:75,113 and the
// env headers), so the library itself is the ground truth for the RNG recipes that
// oracle/ep_oracle.c restates and the CUDA kernels implement.  tests/
// test_oracle_rng_vs_libstdcxx.py loads crafted engine states (operator>>) into both and
// compares draw by draw (state loaded through the object representation), including the corner cases random sampling never reaches (Lemire
// rejections, generate_canonical's >= 1 clamp).
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <random>

namespace {
struct StdRng {
  std::mt19937 gen;
  std::normal_distribution<double> normal;
};
}  // namespace

extern "C" {

void* stdrng_create(uint32_t seed) {
  auto* r = new StdRng();
  r->gen.seed(seed);
  return r;
}
void stdrng_destroy(void* h) { delete static_cast<StdRng*>(h); }

// Load the 624 state words and the read position (what operator<< prints).  Done through the
// object representation rather than operator>>: this .so carries a static libstdc++ and its
// iostreams must not be used inside a host process that has its own.  libstdc++'s
// mersenne_twister_engine is { _UIntType _M_x[624]; size_t _M_p; } (bits/random.h); the
// static_assert and the seeded-stream test (engine output after a load == the standard's
// sequence) guard the assumption.
struct MtLayout {
  std::mt19937::result_type x[624];
  std::size_t p;
};
static_assert(sizeof(std::mt19937) == sizeof(MtLayout), "unexpected std::mt19937 layout");

int stdrng_set(void* h, const uint32_t* mt624, int idx) {
  auto* r = static_cast<StdRng*>(h);
  MtLayout l;
  for (int i = 0; i < 624; ++i) l.x[i] = mt624[i];
  l.p = static_cast<std::size_t>(idx);
  std::memcpy(static_cast<void*>(&r->gen), &l, sizeof(l));
  r->normal.reset();
  return 0;
}
uint32_t stdrng_next(void* h) { return static_cast<StdRng*>(h)->gen(); }
int stdrng_uniform_int(void* h, int a, int b) {
  std::uniform_int_distribution<int> d(a, b);
  return d(static_cast<StdRng*>(h)->gen);
}
double stdrng_uniform_real(void* h, double a, double b) {
  std::uniform_real_distribution<double> d(a, b);
  return d(static_cast<StdRng*>(h)->gen);
}
double stdrng_normal(void* h, double mean, double stddev) {
  auto* r = static_cast<StdRng*>(h);
  return r->normal(r->gen, std::normal_distribution<double>::param_type(mean, stddev));
}

}  // extern "C"
