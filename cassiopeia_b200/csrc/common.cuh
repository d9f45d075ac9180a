// Shared helpers for the cas_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

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


// Programmatic dependent launch (sm_90+): the kernel may be scheduled while its predecessor in the stream is
// still draining; it must execute pdl_wait() before touching anything the predecessor wrote, and both call
// pdl_launch_dependents() early so that the NEXT kernel's launch latency overlaps their execution.  Used by the
// three small kernels of a pruned NJ iteration, whose back-to-back launch latencies are a visible part of the
// ~45 us iteration.  CAS_PDL=0 falls back to ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

// Development aid (build with -DCAS_TIMELINE, see scripts/timeline.py): per-CTA clock64 stamps at the phase
// boundaries of the loop kernels, for a window of iterations.  Compiled out of the product library.
#ifdef CAS_TIMELINE
struct Timeline {
  unsigned long long *buf;  // [window][3 kernels][TL_CTAS][TL_SLOTS]
  int t0, t1;
};
constexpr int TL_CTAS = 512, TL_SLOTS = 12;
static __device__ Timeline g_timeline;  // one copy per translation unit; agglom.cu's is set by timeline_set()
__device__ __forceinline__ void tl_stamp(int t, int kernel, int slot) {
  const Timeline tl = g_timeline;
  if (tl.buf && t >= tl.t0 && t < tl.t1 && blockIdx.x < TL_CTAS) {
    unsigned long long c = clock64();
    unsigned long long g;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g));
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    unsigned long long *p = tl.buf + ((((size_t)(t - tl.t0) * 3 + kernel) * TL_CTAS + blockIdx.x) * TL_SLOTS);
    p[slot] = c;
    if (slot == 0) { p[TL_SLOTS - 1] = g; p[TL_SLOTS - 2] = smid; }
  }
}
#define TL_STAMP(t, kernel, slot) do { if (threadIdx.x == 0) ::cas::tl_stamp((t), (kernel), (slot)); } while (0)
#else
#define TL_STAMP(t, kernel, slot) do { } while (0)
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t stream,
                                     Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace cas
