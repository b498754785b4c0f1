// api_common.cuh — shared host-side helpers of the C ABI (error string, launch counter,
// workspace carving, launch wrappers).
#pragma once
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "../../include/carel_head.h"
#include "common.cuh"
#include "tmap.h"

namespace carel {

char* error_buffer();                 // thread-local, defined in api_common.cu
std::atomic<long long>& launch_counter();
int sm_count();

inline int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return 1;
}

#define CAREL_CHECK_CUDA(expr)                                                          \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                       __LINE__);                                                       \
  } while (0)

#define CAREL_REQUIRE(cond, ...)                  \
  do {                                            \
    if (!(cond)) return set_error(__VA_ARGS__);   \
  } while (0)

// Optional per-kernel timing (eager profiling passes only — never enable under graph capture):
// when enabled, the launch helpers bracket each tensor-core kernel with CUDA events on the launch
// stream; carel_timing_report() synchronises and sums them per tag.
bool timing_enabled();
void timing_begin(const char* tag, cudaStream_t st);
void timing_end(cudaStream_t st);

// count + check a kernel launch
#define CAREL_LAUNCHED()                               \
  do {                                                 \
    launch_counter().fetch_add(1);                     \
    CAREL_CHECK_CUDA(cudaGetLastError());              \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int64_t round_up64(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct WsCarver {
  uint8_t* base;
  size_t off = 0;
  explicit WsCarver(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <class T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base ? base + off : nullptr);
    off += n * sizeof(T);
    return r;
  }
  size_t bytes() const { return align_up(off, 256); }
};

}  // namespace carel
