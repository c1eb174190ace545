// Shared host-side plumbing for the CUDA translation units: error reporting and small RAII helpers.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string>
#include <vector>

namespace b200 {

// thread-local message of the last failing call (returned by b200_last_error())
std::string &last_error();
int set_error(int code, const char *fmt, ...);

#define B200_CUDA_CHECK(expr)                                                                             \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return ::b200::set_error(-100 - (int)_e, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,       \
                               cudaGetErrorString(_e));                                                   \
  } while (0)

#define B200_CHECK(expr)        \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)

// device buffer with automatic release (host-side helper)
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  int alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      p = nullptr;
      cudaGetLastError();  // clear the (non-sticky) error so that other users of the context are not affected
      return set_error(-100 - (int)e, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
    }
    bytes = n;
    return 0;
  }
  // grow-only: keeps the allocation when it is already large enough (scratch reuse across calls)
  int reserve(size_t n) {
    if (p && bytes >= n) return 0;
    return alloc(n);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T *as() const { return (T *)p; }
};

struct Timer {  // CUDA-event timer on one stream
  cudaEvent_t a, b;
  cudaStream_t st;
  explicit Timer(cudaStream_t s = 0) : st(s) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
  }
  ~Timer() {
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
  void start() { cudaEventRecord(a, st); }
  void stop_async() { cudaEventRecord(b, st); }  // read later with elapsed() after the stream has drained
  float elapsed() {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  }
  float stop() {
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
  }
};

// number of kernels this library launched (bench.py reports it as gpu_launches)
std::atomic<unsigned long long> &launch_counter();
static inline void note_launch(unsigned n = 1) { launch_counter().fetch_add(n, std::memory_order_relaxed); }

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

}  // namespace b200
