// Single-step entry points behind the reference's overridable hooks:
//   cas_nj_q        <- NeighborJoiningSolver.compute_q        (NeighborJoiningSolver.py:142-171)
//   cas_update_map  <- {NeighborJoining,UPGMA}Solver.update_dissimilarity_map
//                      (NeighborJoiningSolver.py:173-260, UPGMASolver.py:140-238)
// float64, reference operation order, no FMA.  O(n^2) utilities, not the hot loop.
#include "common.cuh"

namespace cas {

__global__ void q_rowsum_kernel(const double *__restrict__ D, int64_t ld, int n, double *__restrict__ rs) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  // row sum of row c, accumulated left to right (numba d.sum(axis=1)); D symmetric is NOT
  // assumed here: read row c itself.
  const double *row = D + (int64_t)c * ld;
  double s = 0.0;
  for (int j = 0; j < n; ++j) s = __dadd_rn(s, row[j]);
  rs[c] = s;
}

__global__ void q_fill_kernel(const double *__restrict__ D, int64_t ld, int n,
                              const double *__restrict__ rs, double tinv, double *__restrict__ Q) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n) return;
  double q = 0.0;
  if (i != j) {
    const int hi = max(i, j), lo = min(i, j);
    q = __dsub_rn(D[(int64_t)hi * ld + lo], __dmul_rn(tinv, __dadd_rn(rs[hi], rs[lo])));
  }
  Q[(int64_t)i * n + j] = q;
}

__global__ void update_map_kernel(const double *__restrict__ D, int64_t ld, int n, int algo, int ci,
                                  int cj, double zi, double zj, double *__restrict__ out, int64_t ldo) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  const int no = n - 1;
  if (c >= no) return;
  // r-th surviving source index (ci < cj)
  auto src = [&](int x) {
    int s = x;
    if (s >= ci) ++s;
    if (s >= cj) ++s;
    return s;
  };
  double v;
  const bool rnew = (r == no - 1), cnew = (c == no - 1);
  if (rnew && cnew) {
    v = 0.0;
  } else if (rnew || cnew) {
    const int s = src(rnew ? c : r);
    const double a = D[(int64_t)s * ld + ci], b = D[(int64_t)s * ld + cj];
    if (algo == CAS_ALGO_NJ)
      v = __dmul_rn(0.5, __dsub_rn(__dadd_rn(a, b), D[(int64_t)ci * ld + cj]));
    else
      v = __ddiv_rn(__dadd_rn(__dmul_rn(zi, a), __dmul_rn(zj, b)), __dadd_rn(zi, zj));
  } else {
    v = D[(int64_t)src(r) * ld + src(c)];
  }
  out[(int64_t)r * ldo + c] = v;
}

}  // namespace cas

using namespace cas;

extern "C" {

int cas_nj_q(const double *D, int64_t n, int64_t ld, double *Q, double *rowsum_scratch, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D && Q && rowsum_scratch, "null pointer argument");
  CAS_REQUIRE(n >= 3 && ld >= n, "compute_q needs n >= 3 (n=%lld)", (long long)n);
  cudaStream_t s = (cudaStream_t)stream;
  q_rowsum_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(D, ld, (int)n, rowsum_scratch);
  CAS_LAUNCH_CHECK();
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)n);
  q_fill_kernel<<<grid, 256, 0, s>>>(D, ld, (int)n, rowsum_scratch, 1.0 / (double)(n - 2), Q);
  CAS_LAUNCH_CHECK();
  return CAS_OK;
}

int cas_update_map(const double *D, int64_t n, int64_t ld, int32_t algo, int64_t i, int64_t j,
                   int64_t size_i, int64_t size_j, double *out, int64_t ld_out, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D && out, "null pointer argument");
  CAS_REQUIRE(n >= 3 && ld >= n && ld_out >= n - 1, "bad sizes");
  CAS_REQUIRE(0 <= i && i < n && 0 <= j && j < n && i != j, "bad cherry (%lld, %lld)", (long long)i, (long long)j);
  CAS_REQUIRE(algo == CAS_ALGO_NJ || algo == CAS_ALGO_UPGMA, "bad algo %d", algo);
  const int ci = (int)(i < j ? i : j), cj = (int)(i < j ? j : i);
  const double zi = (double)(i < j ? size_i : size_j), zj = (double)(i < j ? size_j : size_i);
  dim3 grid((unsigned)((n - 1 + 255) / 256), (unsigned)(n - 1));
  update_map_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(D, ld, (int)n, algo, ci, cj, zi, zj, out, ld_out);
  CAS_LAUNCH_CHECK();
  return CAS_OK;
}

}  // extern "C"
