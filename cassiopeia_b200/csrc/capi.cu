// C-ABI entry points of libcas_b200.so (declared in include/cas_b200.h).
#include <mutex>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace cas {

static thread_local char g_err[1024] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int64_t n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

bool pdl_enabled() {
  const char *e = getenv("CAS_PDL");
  return !(e && e[0] == '0');
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
      cached = sms;
    else
      cached = 148;
  }
  return cached;
}

// implemented in whd_exact.cu / whd_gemm.cu / agglom.cu
size_t whd_exact_workspace_bytes(int64_t n, int32_t m, int32_t weighted);
int whd_exact_map(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w,
                  int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0, int64_t row1,
                  void *workspace, size_t workspace_bytes, cudaStream_t stream, int cyc_G = 0,
                  int cyc_r = 0);
int pairwise_exact_map(int32_t fn, int32_t flags, const int32_t *cm, int64_t n, int32_t m, int32_t missing,
                       const double *w, int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0,
                       int64_t row1, void *workspace, size_t workspace_bytes, cudaStream_t stream, int cyc_G = 0,
                       int cyc_r = 0);
int whd_read_error_flag(const void *workspace, cudaStream_t stream, int *flag);
size_t whd_gemm_workspace_bytes(int64_t n, int32_t m, int32_t n_states, int32_t weighted);
int int8_mma_peak(int iters, cudaStream_t stream, double *ops_per_s);
int whd_gemm_map(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w,
                 int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0, int64_t row1,
                 void *workspace, size_t workspace_bytes, cudaStream_t stream, int world = 1, int rank = 0);
size_t agglom_workspace_bytes(int64_t n, bool upgma);
int agglom_read_stats(const void *workspace, int64_t n, cudaStream_t stream, int64_t *near_ties, int64_t *exact_ties);
size_t smj_workspace_bytes(int64_t n, int32_t m);
int smj_solve(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w, int32_t n_states,
              int32_t fn, double *D, int64_t ld, int64_t max_iters, int32_t *joins, int32_t *last2,
              void *workspace, size_t workspace_bytes, cudaStream_t stream);
int agglom_read_scanned(const void *workspace, int64_t n, cudaStream_t stream, int64_t *elements);
int agglom_read_exact_stats(const void *workspace, int64_t n, cudaStream_t stream, int64_t *out4);
int agglom_read_exact_log(const void *workspace, int64_t n, cudaStream_t stream, int32_t *out, int64_t cap_rows);
int loop_launch_floor(int64_t k, int64_t iters, int32_t kernels_per_iter, void *workspace, cudaStream_t stream);
#ifdef CAS_TIMELINE
int timeline_set(void *buf, int t0, int t1);
#endif
int agglom_read_prune_diag(const void *workspace, int64_t n, cudaStream_t stream, int64_t *rows_listed, double *drift);
int agglom_solve(void *D, int64_t n, int64_t ld, int32_t dtype, int32_t algo, int32_t mode,
                 int64_t max_iters, int32_t *joins, int32_t *last2, void *workspace,
                 size_t workspace_bytes, cudaStream_t stream);

}  // namespace cas

using namespace cas;

extern "C" {

int cas_abi_version(void) { return CAS_ABI_VERSION; }

const char *cas_last_error(void) { return g_err; }

int64_t cas_last_launch_count(void) { return g_launches; }

int cas_device_info(int32_t *sms, int32_t *cc_major, int32_t *cc_minor, size_t *free_bytes,
                    size_t *total_bytes) {
  int dev = 0;
  CAS_CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CAS_CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (sms) *sms = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  size_t f = 0, t = 0;
  CAS_CUDA_TRY(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return CAS_OK;
}

size_t cas_whd_workspace_bytes(int64_t n, int32_t m, int32_t n_states, int32_t weighted,
                               int32_t method) {
  if (method == CAS_MAP_TENSOR) return whd_gemm_workspace_bytes(n, m, n_states, weighted);
  return whd_exact_workspace_bytes(n, m, weighted);
}

int cas_whd_map(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w,
                int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0, int64_t row1,
                int32_t method, void *workspace, size_t workspace_bytes, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(cm && D && workspace, "null pointer argument");
  CAS_REQUIRE(n >= 1 && m >= 1, "empty character matrix (n=%lld, m=%d)", (long long)n, m);
  CAS_REQUIRE(0 <= row0 && row0 <= row1 && row1 <= n, "bad row block [%lld,%lld)",
              (long long)row0, (long long)row1);
  CAS_REQUIRE(ld >= n, "ld=%lld < n=%lld", (long long)ld, (long long)n);
  CAS_REQUIRE(dtype == CAS_F32 || dtype == CAS_F64, "bad dtype %d", dtype);
  if (method == CAS_MAP_TENSOR)
    return whd_gemm_map(cm, n, m, missing, w, n_states, D, ld, dtype, row0, row1, workspace,
                        workspace_bytes, (cudaStream_t)stream);
  CAS_REQUIRE(method == CAS_MAP_EXACT, "bad map method %d", method);
  return whd_exact_map(cm, n, m, missing, w, n_states, D, ld, dtype, row0, row1, workspace,
                       workspace_bytes, (cudaStream_t)stream);
}

int cas_pairwise_map(int32_t fn, int32_t flags, const int32_t *cm, int64_t n, int32_t m, int32_t missing,
                     const double *w, int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0,
                     int64_t row1, void *workspace, size_t workspace_bytes, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(cm && D && workspace, "null pointer argument");
  CAS_REQUIRE(n >= 1 && m >= 1, "empty character matrix (n=%lld, m=%d)", (long long)n, m);
  CAS_REQUIRE(0 <= row0 && row0 <= row1 && row1 <= n, "bad row block [%lld,%lld)",
              (long long)row0, (long long)row1);
  CAS_REQUIRE(ld >= n, "ld=%lld < n=%lld", (long long)ld, (long long)n);
  CAS_REQUIRE(dtype == CAS_F32 || dtype == CAS_F64, "bad dtype %d", dtype);
  return pairwise_exact_map(fn, flags, cm, n, m, missing, w, n_states, D, ld, dtype, row0, row1, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int cas_whd_map_cyclic(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w,
                       int32_t n_states, void *D_local, int64_t ld, int32_t dtype, int32_t rank,
                       int32_t world, int32_t method, void *workspace, size_t workspace_bytes,
                       void *stream) {
  reset_launch_count();
  CAS_REQUIRE(cm && D_local && workspace, "null pointer argument");
  CAS_REQUIRE(n >= 1 && m >= 1 && ld >= n, "bad sizes");
  CAS_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
  CAS_REQUIRE(dtype == CAS_F32 || dtype == CAS_F64, "bad dtype %d", dtype);
  CAS_REQUIRE(method == CAS_MAP_EXACT || method == CAS_MAP_TENSOR, "bad map method %d", method);
  if (method == CAS_MAP_TENSOR && world > 1)
    return whd_gemm_map(cm, n, m, missing, w, n_states, D_local, ld, dtype, 0, n, workspace, workspace_bytes,
                        (cudaStream_t)stream, world, rank);
  if (method == CAS_MAP_TENSOR)  // one rank: the shard is the whole map
    return whd_gemm_map(cm, n, m, missing, w, n_states, D_local, ld, dtype, 0, n, workspace, workspace_bytes,
                        (cudaStream_t)stream);
  return whd_exact_map(cm, n, m, missing, w, n_states, D_local, ld, dtype, 0, n, workspace,
                       workspace_bytes, (cudaStream_t)stream, world, rank);
}

size_t cas_agglom_workspace_bytes(int64_t n, int32_t algo, int32_t mode) {
  (void)mode;
  return agglom_workspace_bytes(n, algo == CAS_ALGO_UPGMA);
}

int cas_agglom_stats(const void *workspace, int64_t n, void *stream, int64_t *near_ties, int64_t *exact_ties) {
  CAS_REQUIRE(workspace, "null pointer argument");
  return agglom_read_stats(workspace, n, (cudaStream_t)stream, near_ties, exact_ties);
}

size_t cas_smj_workspace_bytes(int64_t n, int32_t m) { return smj_workspace_bytes(n, m); }

int cas_smj_solve(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w, int32_t n_states,
                  int32_t fn, double *S_map, int64_t ld, int64_t max_iters, int32_t *joins, int32_t *last2,
                  void *workspace, size_t workspace_bytes, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(cm && S_map && joins && last2 && workspace, "null pointer argument");
  CAS_REQUIRE(n >= 1 && m >= 1 && n < (int64_t)1 << 30, "bad sizes (n=%lld, m=%d)", (long long)n, m);
  CAS_REQUIRE(ld >= n && ld % 32 == 0, "ld=%lld must be >= n and a multiple of 32", (long long)ld);
  CAS_REQUIRE(((uintptr_t)S_map & 15) == 0, "the map must be 16-byte aligned");
  return smj_solve(cm, n, m, missing, w, n_states, fn, S_map, ld, max_iters, joins, last2, workspace,
                   workspace_bytes, (cudaStream_t)stream);
}

#ifdef CAS_TIMELINE
__attribute__((visibility("default"))) int cas_debug_timeline(void *buf, int t0, int t1) {
  return cas::timeline_set(buf, t0, t1);
}
#endif

int cas_loop_launch_floor(int64_t k, int64_t iters, int32_t kernels_per_iter, void *workspace, size_t workspace_bytes,
                          void *stream) {
  reset_launch_count();
  CAS_REQUIRE(workspace && k >= 1 && iters >= 0 && kernels_per_iter >= 1 && kernels_per_iter <= 3, "bad argument");
  CAS_REQUIRE(workspace_bytes >= agglom_workspace_bytes(k, false), "workspace too small");
  return loop_launch_floor(k, iters, kernels_per_iter, workspace, (cudaStream_t)stream);
}

int cas_int8_mma_peak(int32_t iters, void *stream, double *ops_per_s) {
  CAS_REQUIRE(ops_per_s && iters >= 1, "bad argument");
  return int8_mma_peak(iters, (cudaStream_t)stream, ops_per_s);
}

int cas_agglom_exact_stats(const void *workspace, int64_t n, void *stream, int64_t *out4) {
  CAS_REQUIRE(workspace && out4, "null pointer argument");
  return agglom_read_exact_stats(workspace, n, (cudaStream_t)stream, out4);
}

int cas_agglom_exact_log(const void *workspace, int64_t n, void *stream, int32_t *out, int64_t cap_rows) {
  CAS_REQUIRE(workspace && out && cap_rows >= 0, "bad argument");
  return agglom_read_exact_log(workspace, n, (cudaStream_t)stream, out, cap_rows);
}

int cas_agglom_scanned_elements(const void *workspace, int64_t n, void *stream, int64_t *elements) {
  CAS_REQUIRE(workspace, "null pointer argument");
  return agglom_read_scanned(workspace, n, (cudaStream_t)stream, elements);
}

int cas_agglom_prune_diag(const void *workspace, int64_t n, void *stream, int64_t *rows_listed, double *drift) {
  CAS_REQUIRE(workspace, "null pointer argument");
  return agglom_read_prune_diag(workspace, n, (cudaStream_t)stream, rows_listed, drift);
}

int cas_nj_solve(void *D, int64_t n, int64_t ld, int32_t dtype, int32_t mode, int64_t max_iters,
                 int32_t *joins, int32_t *last2, void *workspace, size_t workspace_bytes,
                 void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D && joins && last2 && workspace, "null pointer argument");
  return agglom_solve(D, n, ld, dtype, CAS_ALGO_NJ, mode, max_iters, joins, last2, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}

int cas_upgma_solve(void *D, int64_t n, int64_t ld, int32_t dtype, int32_t mode,
                    int64_t max_iters, int32_t *joins, int32_t *last2, void *workspace,
                    size_t workspace_bytes, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D && joins && last2 && workspace, "null pointer argument");
  return agglom_solve(D, n, ld, dtype, CAS_ALGO_UPGMA, mode, max_iters, joins, last2, workspace,
                      workspace_bytes, (cudaStream_t)stream);
}

// RAII device buffer for the host-buffer entry point
// Device buffers of cas_solve_host are kept between calls (a cudaMalloc / cudaFree pair of a 20 GB map was measured
// at anything between 10 and 400 ms): a call re-uses a cached buffer that is large enough, cas_solve_host_release()
// frees them.  One call at a time (the cache is guarded by a mutex held for the whole call).
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t alloc(size_t bytes) {
    bytes = bytes ? bytes : 1;
    if (p && cap >= bytes) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    const cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes; else p = nullptr;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
static DevBuf g_host_bufs[7];
static std::mutex g_host_mu;

int cas_solve_host_release(void) {
  std::lock_guard<std::mutex> lock(g_host_mu);
  for (auto &b : g_host_bufs) b.release();
  return CAS_OK;
}

int cas_solve_host(const int32_t *cm_host, int64_t n, int32_t m, int32_t missing,
                   const double *w_host, int32_t n_states, int32_t algo, int32_t mode,
                   int32_t map_method, int32_t *joins_host, int32_t *last2_host,
                   double *timings_ms) {
  reset_launch_count();
  CAS_REQUIRE(cm_host && joins_host && last2_host, "null pointer argument");
  CAS_REQUIRE(n >= 2 && m >= 1, "need at least two samples (n=%lld, m=%d)", (long long)n, m);
  CAS_REQUIRE(algo == CAS_ALGO_NJ || algo == CAS_ALGO_UPGMA, "bad algo %d", algo);
  const int dtype = (mode == CAS_MODE_FAST) ? CAS_F32 : CAS_F64;
  const size_t esz = dtype == CAS_F64 ? 8 : 4;
  const int64_t ld = (int64_t)align_up((size_t)n, 32);
  const bool weighted = w_host != nullptr;

  cudaStream_t stream;
  CAS_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } sg{stream};
  cudaEvent_t ev[5];
  for (auto &e : ev) CAS_CUDA_TRY(cudaEventCreate(&e));
  struct EvGuard { cudaEvent_t *e; ~EvGuard() { for (int i = 0; i < 5; ++i) cudaEventDestroy(e[i]); } } eg{ev};

  std::lock_guard<std::mutex> lock(g_host_mu);
  DevBuf &cm_d = g_host_bufs[0], &w_d = g_host_bufs[1], &D_d = g_host_bufs[2], &ws_map = g_host_bufs[3],
         &ws_loop = g_host_bufs[4], &joins_d = g_host_bufs[5], &last2_d = g_host_bufs[6];
  const size_t njoin = n > 2 ? (size_t)(n - 2) : 0;
  CAS_CUDA_TRY(cm_d.alloc((size_t)n * m * sizeof(int32_t)));
  if (weighted) CAS_CUDA_TRY(w_d.alloc((size_t)m * (n_states + 1) * sizeof(double)));
  CAS_CUDA_TRY(D_d.alloc((size_t)n * ld * esz));
  const size_t wsm = cas_whd_workspace_bytes(n, m, n_states, weighted, map_method);
  const size_t wsl = cas_agglom_workspace_bytes(n, algo, mode);
  CAS_CUDA_TRY(ws_map.alloc(wsm));
  CAS_CUDA_TRY(ws_loop.alloc(wsl));
  CAS_CUDA_TRY(joins_d.alloc((njoin + 1) * 2 * sizeof(int32_t)));
  CAS_CUDA_TRY(last2_d.alloc(2 * sizeof(int32_t)));

  CAS_CUDA_TRY(cudaEventRecord(ev[0], stream));
  CAS_CUDA_TRY(cudaMemcpyAsync(cm_d.p, cm_host, (size_t)n * m * sizeof(int32_t),
                               cudaMemcpyHostToDevice, stream));
  if (weighted)
    CAS_CUDA_TRY(cudaMemcpyAsync(w_d.p, w_host, (size_t)m * (n_states + 1) * sizeof(double),
                                 cudaMemcpyHostToDevice, stream));
  CAS_CUDA_TRY(cudaEventRecord(ev[1], stream));
  int rc;
  if (map_method == CAS_MAP_TENSOR)
    rc = whd_gemm_map((const int32_t *)cm_d.p, n, m, missing, (const double *)(weighted ? w_d.p : nullptr), n_states, D_d.p,
                      ld, dtype, 0, n, ws_map.p, wsm, stream);
  else
    rc = whd_exact_map((const int32_t *)cm_d.p, n, m, missing, (const double *)(weighted ? w_d.p : nullptr), n_states,
                       D_d.p, ld, dtype, 0, n, ws_map.p, wsm, stream);
  if (rc != CAS_OK) return rc;
  CAS_CUDA_TRY(cudaEventRecord(ev[2], stream));
  rc = agglom_solve(D_d.p, n, ld, dtype, algo, mode, -1, (int32_t *)joins_d.p,
                    (int32_t *)last2_d.p, ws_loop.p, wsl, stream);
  if (rc != CAS_OK) return rc;
  CAS_CUDA_TRY(cudaEventRecord(ev[3], stream));
  if (njoin)
    CAS_CUDA_TRY(cudaMemcpyAsync(joins_host, joins_d.p, njoin * 2 * sizeof(int32_t),
                                 cudaMemcpyDeviceToHost, stream));
  CAS_CUDA_TRY(cudaMemcpyAsync(last2_host, last2_d.p, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                               stream));
  CAS_CUDA_TRY(cudaEventRecord(ev[4], stream));
  CAS_CUDA_TRY(cudaStreamSynchronize(stream));
  // both map kernels' encode passes set the first int of their workspace when they meet a state outside
  // [0, n_states] that is not the missing indicator: the join list would come from a corrupt map
  int flag = 0;
  rc = whd_read_error_flag(ws_map.p, stream, &flag);
  if (rc != CAS_OK) return rc;
  CAS_REQUIRE(flag == 0, "character matrix holds a state outside [0, n_states] that is not the missing indicator");
  if (njoin && last2_host[0] == -2) {
    set_error("the loop found no finite entry to join (is the dissimilarity map all NaN / inf?)");
    return CAS_ERR_INVALID;
  }
  if (timings_ms) {
    for (int i = 0; i < 4; ++i) {
      float ms = 0.f;
      CAS_CUDA_TRY(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      timings_ms[i] = ms;
    }
  }
  return CAS_OK;
}

}  // extern "C"
