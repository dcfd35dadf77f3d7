// Shim for abseil's check macros (abseil is not vendored under /root/reference and
// is not installed in this image).  Test infrastructure only: lets the reference's
// own headers compile unmodified for oracle/_ref.  CHECK_* abort with a message,
// DCHECK_* compile to nothing (the reference is built -DNDEBUG for timing).
#ifndef EPB200_SHIM_ABSL_LOG_CHECK_H_
#define EPB200_SHIM_ABSL_LOG_CHECK_H_
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace epb200_shim {
class FatalStream {
 public:
  FatalStream(const char* file, int line, const char* expr) {
    ss_ << file << ":" << line << " CHECK failed: " << expr << " ";
  }
  [[noreturn]] ~FatalStream() {
    std::cerr << ss_.str() << std::endl;
    std::abort();
  }
  template <typename T>
  FatalStream& operator<<(const T& v) {
    ss_ << v;
    return *this;
  }

 private:
  std::ostringstream ss_;
};
struct NullStream {
  template <typename T>
  NullStream& operator<<(const T&) {
    return *this;
  }
};
struct Voidify {
  void operator&(const FatalStream&) {}
  void operator&(const NullStream&) {}
};
}  // namespace epb200_shim

#define EPB200_CHECK_IMPL(cond, text) \
  (cond) ? (void)0                    \
         : ::epb200_shim::Voidify() & \
               ::epb200_shim::FatalStream(__FILE__, __LINE__, text)
#define CHECK(c) EPB200_CHECK_IMPL((c), #c)
#define CHECK_EQ(a, b) EPB200_CHECK_IMPL((a) == (b), #a " == " #b)
#define CHECK_NE(a, b) EPB200_CHECK_IMPL((a) != (b), #a " != " #b)
#define CHECK_LE(a, b) EPB200_CHECK_IMPL((a) <= (b), #a " <= " #b)
#define CHECK_LT(a, b) EPB200_CHECK_IMPL((a) < (b), #a " < " #b)
#define CHECK_GE(a, b) EPB200_CHECK_IMPL((a) >= (b), #a " >= " #b)
#define CHECK_GT(a, b) EPB200_CHECK_IMPL((a) > (b), #a " > " #b)
#define EPB200_DCHECK_IMPL \
  true ? (void)0 : ::epb200_shim::Voidify() & ::epb200_shim::NullStream()
#define DCHECK(c) EPB200_DCHECK_IMPL
#define DCHECK_EQ(a, b) EPB200_DCHECK_IMPL
#define DCHECK_NE(a, b) EPB200_DCHECK_IMPL
#define DCHECK_LE(a, b) EPB200_DCHECK_IMPL
#define DCHECK_LT(a, b) EPB200_DCHECK_IMPL
#define DCHECK_GE(a, b) EPB200_DCHECK_IMPL
#define DCHECK_GT(a, b) EPB200_DCHECK_IMPL
#endif  // EPB200_SHIM_ABSL_LOG_CHECK_H_
