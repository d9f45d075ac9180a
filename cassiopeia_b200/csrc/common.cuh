// Shared helpers for the cas_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cas_b200.h"

namespace cas {

// thread-local error message / launch counter behind cas_last_error() / cas_last_launch_count()
void set_error(const char *fmt, ...);
void count_launch(int64_t n = 1);
void reset_launch_count();

#define CAS_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::cas::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                       cudaGetErrorString(_e));                                         \
      return CAS_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define CAS_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t _e = cudaPeekAtLastError();                                             \
    if (_e != cudaSuccess) {                                                            \
      ::cas::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,         \
                       cudaGetErrorString(_e));                                         \
      return CAS_ERR_CUDA;                                                              \
    }                                                                                   \
    ::cas::count_launch();                                                              \
  } while (0)

#define CAS_REQUIRE(cond, ...)                                                          \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      ::cas::set_error(__VA_ARGS__);                                                    \
      return CAS_ERR_INVALID;                                                           \
    }                                                                                   \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// carve aligned sub-buffers out of a caller-provided workspace
struct Carver {
  char *base;
  size_t off = 0;
  explicit Carver(void *p) : base(static_cast<char *>(p)) {}
  template <typename T>
  T *take(size_t count) {
    off = align_up(off, 256);
    T *p = reinterpret_cast<T *>(base + off);
    off += count * sizeof(T);
    return p;
  }
  size_t used() const { return align_up(off, 256); }
};

int sm_count();

}  // namespace cas
