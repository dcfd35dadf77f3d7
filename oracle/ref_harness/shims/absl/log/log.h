// Shim for abseil's LOG/DLOG (see check.h in this directory): messages are dropped.
#ifndef EPB200_SHIM_ABSL_LOG_LOG_H_
#define EPB200_SHIM_ABSL_LOG_LOG_H_
#include "absl/log/check.h"
#define LOG(severity) ::epb200_shim::NullStream()
#define DLOG(severity) ::epb200_shim::NullStream()
#endif  // EPB200_SHIM_ABSL_LOG_LOG_H_
