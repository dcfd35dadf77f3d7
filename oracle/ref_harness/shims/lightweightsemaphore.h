// Shim with the interface of moodycamel::LightweightSemaphore (cameron314/
// concurrentqueue, un-vendored third-party dependency of the reference): the
// same published design -- an atomic count that absorbs uncontended traffic, a
// bounded spin, then a kernel semaphore -- written from scratch.  The CPU
// baseline is very sensitive to this primitive, so the
// spin phase is kept: a mutex/condvar stand-in would under-report the reference.
#ifndef EPB200_SHIM_LIGHTWEIGHTSEMAPHORE_H_
#define EPB200_SHIM_LIGHTWEIGHTSEMAPHORE_H_
#include <semaphore.h>

#include <atomic>
#include <cerrno>
#include <cstddef>
#include <cstdint>

namespace moodycamel {

class LightweightSemaphore {
 public:
  using ssize_t = std::int64_t;
  explicit LightweightSemaphore(ssize_t initial = 0, int max_spins = 10000)
      : count_(initial), max_spins_(max_spins) {
    sem_init(&sem_, 0, 0);
  }
  ~LightweightSemaphore() { sem_destroy(&sem_); }
  LightweightSemaphore(const LightweightSemaphore&) = delete;
  LightweightSemaphore& operator=(const LightweightSemaphore&) = delete;

  bool tryWait() {
    ssize_t old = count_.load(std::memory_order_relaxed);
    while (old > 0) {
      if (count_.compare_exchange_weak(old, old - 1, std::memory_order_acquire,
                                       std::memory_order_relaxed)) {
        return true;
      }
    }
    return false;
  }

  bool wait() {
    if (tryWait()) return true;
    for (int spin = max_spins_; spin > 0; --spin) {
      ssize_t old = count_.load(std::memory_order_relaxed);
      if (old > 0 && count_.compare_exchange_strong(
                         old, old - 1, std::memory_order_acquire,
                         std::memory_order_relaxed)) {
        return true;
      }
      std::atomic_signal_fence(std::memory_order_acquire);  // keep the loop
    }
    ssize_t old = count_.fetch_sub(1, std::memory_order_acquire);
    if (old > 0) return true;
    int rc;
    do {
      rc = sem_wait(&sem_);
    } while (rc == -1 && errno == EINTR);
    return rc == 0;
  }

  void signal(ssize_t n = 1) {
    ssize_t old = count_.fetch_add(n, std::memory_order_release);
    ssize_t to_release = -old < n ? -old : n;
    while (to_release-- > 0) sem_post(&sem_);
  }

  std::size_t availableApprox() const {
    ssize_t c = count_.load(std::memory_order_relaxed);
    return c > 0 ? static_cast<std::size_t>(c) : 0;
  }

 private:
  std::atomic<ssize_t> count_;
  int max_spins_;
  sem_t sem_;
};

}  // namespace moodycamel
#endif  // EPB200_SHIM_LIGHTWEIGHTSEMAPHORE_H_
