// api_common.cuh — shared host-side helpers of the C ABI (error string, launch counter,
// workspace carving, launch wrappers).
#pragma once
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

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

// A library-owned second stream per device for work that is independent of the caller's stream inside ONE entry point
// (carel_bow_fwd: the sparse-target branch).  fork(): the side stream waits for everything enqueued on `st` so far;
// join(k): `st` waits for everything enqueued on the side stream so far (k = which of the join events to use: a later
// fork / join of the same call must not re-record an event an earlier wait still refers to under graph capture).
// Event record / wait pairs are capturable: under CUDA-graph capture the side work becomes a parallel branch.
// nullptr when disabled (carel_set_side_branches / CAREL_SIDE_BRANCHES) or when stream / event creation failed.
struct SideBranch {
  std::mutex mu;      // held by an entry point from its fork to its last join: the events are per device, not per caller
  cudaStream_t s = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  cudaError_t fork(cudaStream_t st) {
    cudaError_t e = cudaEventRecord(ev_fork, st);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(s, ev_fork, 0);
  }
  cudaError_t join(cudaStream_t st, int k) {
    cudaError_t e = cudaEventRecord(ev_join[k], s);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(st, ev_join[k], 0);
  }
};
SideBranch* side_branch();
int side_branch_mask();      // bit 0: carel_bow_fwd's sparse-target branch, bit 1: the two O(nnz) gathers of carel_bow_bwd

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
