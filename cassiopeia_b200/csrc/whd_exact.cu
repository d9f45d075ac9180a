// Exact pairwise weighted-Hamming dissimilarity map on CUDA cores.
//
// Restates cassiopeia/solver/dissimilarity_functions.py:46-72 (per pair) and the pair
// enumeration of cassiopeia/data/utilities.py:358-368, producing the square symmetric
// map the reference obtains through scipy squareform (CassiopeiaTree.py:1935).
//
// Bit-exactness: the numerator is accumulated in float64 in character order exactly
// as the reference does (`d += w[a] + w[b]`: the two weights are added first, then
// added to d; an uncut state contributes weight 0.0 and x + 0.0 == x), the unweighted
// numerator is an exact integer, and the only other rounding is the IEEE float64
// division by the shared-present count.  No multiplications => nothing for the
// compiler to contract; the adds use __dadd_rn anyway.
//
// Layout: the int32 character matrix is first packed (K1 `encode`) into 16-bit codes
// [n, m_pad] (missing -> 0xFFFF) and, when weighted, a per-cell weight row
// [n, m_pad] float64 (weight of the cell's own state, 0.0 for uncut/missing), so the
// pair kernel needs no table lookups.  Pair kernel: 64x64 output tile per CTA,
// 4x4 pairs per thread, characters staged through shared memory in chunks of 32.
//
// weighted_hamming_distance (and its exponential) only adds something at a character where both
// cells are present and their states differ, which requires at least one of them to be cut.
// K1b `mask_kernel` therefore packs, per cell and 32-character chunk, a PRESENT and a CUT bit mask;
// the pair kernel forms (cut_a | cut_b) & present_a & present_b with three bit operations per chunk
// and visits only those characters, in ascending order (= the reference's accumulation order), the
// shared-present count being a popcount.  On lineage-tracing data (~10 % of the sites cut) that is
// ~10 character visits per pair instead of m: the ncu profile of the dense loop showed the kernel
// issue-bound at 15 thread instructions per (pair, character) (profiles/r2_whd_exact_kernel.md).
#include "common.cuh"

namespace cas {

constexpr int TILE = 64;      // pairs per CTA edge
constexpr int CK = 32;        // characters per shared-memory chunk
constexpr int CPAD = TILE + 4;  // u16 row pitch (8-byte aligned 4-code reads)
constexpr int WPAD = TILE + 2;  // f64 row pitch (16-byte aligned 2-weight reads)
constexpr uint16_t MISS = 0xFFFFu;

__global__ void encode_kernel(const int32_t *__restrict__ cm, int64_t n, int m, int m_pad,
                              int32_t missing, const double *__restrict__ w, int n_states,
                              uint16_t *__restrict__ codes, double *__restrict__ wrow,
                              int *__restrict__ err_flag) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = n * (int64_t)m_pad;
  if (idx >= total) return;
  int64_t r = idx / m_pad;
  int c = (int)(idx - r * m_pad);
  uint16_t code = MISS;
  double wt = 0.0;
  if (c < m) {
    int32_t s = cm[r * m + c];
    if (s == missing) {
      code = MISS;
    } else if (s < 0 || s > n_states) {
      atomicExch(err_flag, 1);  // state outside [0, n_states]: host validation was skipped
      code = MISS;
    } else {
      code = (uint16_t)s;
      if (w != nullptr && s != 0) wt = w[(int64_t)c * (n_states + 1) + s];
    }
  }
  codes[idx] = code;
  if (wrow != nullptr) wrow[idx] = wt;
}

// one thread per (cell, 32-character chunk): bit c of `pres` = character present, of `cut` = present and cut
__global__ void mask_kernel(const uint16_t *__restrict__ codes, int64_t n, int m, int m_pad, int nch,
                            uint32_t *__restrict__ pres, uint32_t *__restrict__ cut) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * (int64_t)nch) return;
  const int64_t r = idx / nch;
  const int ch = (int)(idx - r * nch);
  uint32_t p = 0, c = 0;
  for (int b = 0; b < 32; ++b) {
    const int cc = ch * 32 + b;
    if (cc >= m) break;
    const uint16_t code = codes[r * m_pad + cc];
    if (code != MISS) {
      p |= 1u << b;
      if (code != 0) c |= 1u << b;
    }
  }
  pres[idx] = p;
  cut[idx] = c;
}

template <typename T>
__device__ __forceinline__ T cvt_out(double v);
template <>
__device__ __forceinline__ double cvt_out<double>(double v) { return v; }
template <>
__device__ __forceinline__ float cvt_out<float>(double v) { return __double2float_rn(v); }

// grid.x enumerates tiles.  tri != 0: lower-triangular tiles of the full map, mirrored.
// tri == 0: all tiles of the row block [row0,row1) x [0,n).
//
// FN selects the pair function (CAS_FN_* of cas_b200.h), all of
// cassiopeia/solver/dissimilarity_functions.py, all symmetric, all accumulated in character order:
//   0 weighted_hamming_distance (:12-72)            differing present pairs: w(a)+w(b) | 1 | 2, / #present
//   1 hamming_similarity_without_missing (:75-113)   equal cut states: w(a) | 1, not normalised
//   2 hamming_similarity_normalized_over_missing (:116-160)   same numerator / #present
//   3 hamming_distance (:163-194)                    a != b (missing is a state); ignore_missing skips
//                                                    positions where either is missing; integer
//   4 weighted_hamming_similarity (:197-240)         equal cut: 2 w(a) | 2; equal uncut: 1 (unweighted only)
//   5 exponential_negative_hamming_distance (:243-300)   exp(-(0)), 0 when nothing is shared
template <typename T, bool WEIGHTED, int FN>
__global__ void __launch_bounds__(256)
whd_exact_kernel(const uint16_t *__restrict__ codes, const double *__restrict__ wrow, int64_t n,
                 int m, int m_pad, T *__restrict__ D, int64_t ld, int64_t row0, int64_t row1,
                 int tri, int tiles_x, int cyc_G, int cyc_r, int ignore_missing,
                 const uint32_t *__restrict__ pres, const uint32_t *__restrict__ cut, int nch) {
  __shared__ __align__(16) uint16_t sA[CK][CPAD];
  __shared__ __align__(16) uint16_t sB[CK][CPAD];
  __shared__ __align__(16) double sWA[WEIGHTED ? CK : 1][WPAD];
  __shared__ __align__(16) double sWB[WEIGHTED ? CK : 1][WPAD];
  // SPARSE: the function only adds where both are present and at least one is cut; when few characters of the
  // tile's cells are cut (< 1/8: lineage data at low mutation rates) only those are visited, else the dense loop,
  // whose shared-memory reads are amortised over the thread's 4 x 4 pairs, is faster (measured: sparse 36 ms vs
  // dense 41 ms at 7 % cut, 85 ms vs 41 ms at 38 % cut, 50k x 60)
  constexpr bool SPARSE = (FN == 0 || FN == 5);
  __shared__ uint32_t sPA[SPARSE ? TILE : 1], sCA[SPARSE ? TILE : 1], sPB[SPARSE ? TILE : 1], sCB[SPARSE ? TILE : 1];
  __shared__ int sCutBits[2];
  static_assert(CK == 32, "one 32-bit mask per chunk");
  if (threadIdx.x < 2) sCutBits[threadIdx.x] = 0;
  __syncthreads();

  int64_t tr, tc;  // tile row / col indices
  if (tri) {
    // linear index -> (tr, tc) with tc <= tr
    int64_t t = blockIdx.x;
    int64_t r = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= t) ++r;
    while (r * (r + 1) / 2 > t) --r;
    tr = r;
    tc = t - r * (r + 1) / 2;
  } else {
    tr = blockIdx.x / tiles_x;
    tc = blockIdx.x % tiles_x;
  }
  // first row (global index) of this tile; block-cyclic shards (cyc_G > 0) own the 64-row
  // blocks cyc_r, cyc_r + cyc_G, ... and store them contiguously in local order
  const int64_t i0 = cyc_G > 0 ? (tr * cyc_G + cyc_r) * TILE : row0 + tr * TILE;
  const int64_t o0 = cyc_G > 0 ? tr * TILE : tr * TILE;  // first output row of this tile
  const int64_t j0 = tc * TILE;         // first column
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  double acc[4][4];
  int npres[4][4];
  int inum[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { acc[a][b] = 0.0; npres[a][b] = 0; inum[a][b] = 0; }

  for (int c0 = 0; c0 < m; c0 += CK) {
    // stage [64 rows x CK chars] of both operands, transposed to [char][row]
    for (int e = threadIdx.x; e < TILE * CK; e += 256) {
      int row = e / CK, c = e % CK;
      int64_t gi = i0 + row, gj = j0 + row;
      int cc = c0 + c;
      uint16_t ca = MISS, cb = MISS;
      double wa = 0.0, wb = 0.0;
      if (cc < m_pad) {
        if (gi < n) { ca = codes[gi * m_pad + cc]; if (WEIGHTED) wa = wrow[gi * m_pad + cc]; }
        if (gj < n) { cb = codes[gj * m_pad + cc]; if (WEIGHTED) wb = wrow[gj * m_pad + cc]; }
      }
      sA[c][row] = ca;
      sB[c][row] = cb;
      if (WEIGHTED) { sWA[c][row] = wa; sWB[c][row] = wb; }
    }
    if (SPARSE && threadIdx.x < 2 * TILE) {
      const int row = threadIdx.x & (TILE - 1);
      const bool bside = threadIdx.x >= TILE;
      const int64_t g = (bside ? j0 : i0) + row;
      const uint32_t pm = g < n ? pres[g * nch + c0 / CK] : 0u, cm = g < n ? cut[g * nch + c0 / CK] : 0u;
      if (bside) { sPB[row] = pm; sCB[row] = cm; } else { sPA[row] = pm; sCA[row] = cm; }
      const int bits = __popc(cm);
      const int wsum = __reduce_add_sync(0xffffffffu, bits);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sCutBits[(c0 / CK) & 1], wsum);
    }
    __syncthreads();
    const int cend = min(CK, m - c0);
    const bool sparse_now = SPARSE && sCutBits[(c0 / CK) & 1] * 8 < 2 * TILE * cend;
    if (threadIdx.x == 0) sCutBits[((c0 / CK) & 1) ^ 1] = 0;  // the other parity is next chunk's counter
    if (sparse_now) {
      uint32_t pa[4], ca4[4], pb[4], cb4[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { pa[x] = sPA[ty * 4 + x]; ca4[x] = sCA[ty * 4 + x]; }
#pragma unroll
      for (int y = 0; y < 4; ++y) { pb[y] = sPB[tx * 4 + y]; cb4[y] = sCB[tx * 4 + y]; }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const uint32_t both = pa[x] & pb[y];
          npres[x][y] += __popc(both);
          uint32_t cand = (ca4[x] | cb4[y]) & both;  // both present, at least one cut: the states MAY differ
          while (cand) {
            const int c = __ffs(cand) - 1;  // ascending: the reference's accumulation order
            cand &= cand - 1;
            const uint16_t a = sA[c][ty * 4 + x], b = sB[c][tx * 4 + y];
            if (a != b) {
              if (WEIGHTED) {
                // d += w(a) + w(b): weights summed first, then accumulated (reference order); an uncut
                // state carries weight 0.0 and x + 0.0 == x, which is the reference's one-sided branch
                acc[x][y] = __dadd_rn(acc[x][y], __dadd_rn(sWA[c][ty * 4 + x], sWB[c][tx * 4 + y]));
              } else {
                inum[x][y] += (a != 0 ? 1 : 0) + (b != 0 ? 1 : 0);
              }
            }
          }
        }
      }
    }
    for (int c = 0; !sparse_now && c < cend; ++c) {
      uint16_t a[4], b[4];
      *reinterpret_cast<uint2 *>(a) = *reinterpret_cast<const uint2 *>(&sA[c][ty * 4]);
      *reinterpret_cast<uint2 *>(b) = *reinterpret_cast<const uint2 *>(&sB[c][tx * 4]);
      double wa[4], wb[4];
      if (WEIGHTED) {
        *reinterpret_cast<double2 *>(&wa[0]) = *reinterpret_cast<const double2 *>(&sWA[c][ty * 4]);
        *reinterpret_cast<double2 *>(&wa[2]) = *reinterpret_cast<const double2 *>(&sWA[c][ty * 4 + 2]);
        *reinterpret_cast<double2 *>(&wb[0]) = *reinterpret_cast<const double2 *>(&sWB[c][tx * 4]);
        *reinterpret_cast<double2 *>(&wb[2]) = *reinterpret_cast<const double2 *>(&sWB[c][tx * 4 + 2]);
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const bool pres = (a[x] != MISS) && (b[y] != MISS);
          npres[x][y] += pres ? 1 : 0;
          if (FN == 0 || FN == 5) {
            const bool neq = pres && (a[x] != b[y]);
            if (WEIGHTED) {
              // d += w(a) + w(b): weights summed first, then accumulated (reference order)
              const double t = __dadd_rn(wa[x], wb[y]);
              if (neq) acc[x][y] = __dadd_rn(acc[x][y], t);
            } else {
              if (neq) inum[x][y] += (a[x] != 0 ? 1 : 0) + (b[y] != 0 ? 1 : 0);
            }
          } else if (FN == 1 || FN == 2) {
            const bool hit = pres && (a[x] == b[y]) && (a[x] != 0);
            if (WEIGHTED) {
              if (hit) acc[x][y] = __dadd_rn(acc[x][y], wa[x]);
            } else {
              inum[x][y] += hit ? 1 : 0;
            }
          } else if (FN == 3) {
            const bool differ = (a[x] != b[y]) && (pres || !ignore_missing);
            inum[x][y] += differ ? 1 : 0;
          } else {
            const bool eq = pres && (a[x] == b[y]);
            if (WEIGHTED) {
              if (eq && a[x] != 0) acc[x][y] = __dadd_rn(acc[x][y], __dmul_rn(2.0, wa[x]));
            } else {
              if (eq) inum[x][y] += (a[x] != 0) ? 2 : 1;
            }
          }
        }
      }
    }
    __syncthreads();
  }

  const bool mirror = tri && (tr != tc);
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int64_t gi = i0 + ty * 4 + x;
    if (gi >= n || gi >= row1) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int64_t gj = j0 + tx * 4 + y;
      if (gj >= n) continue;
      double v = 0.0;
      const double num = WEIGHTED ? acc[x][y] : (double)inum[x][y];
      if (FN == 1 || FN == 3) {
        v = num;
      } else if (npres[x][y] > 0) {
        v = __ddiv_rn(num, (double)npres[x][y]);
        if (FN == 5) v = exp(-v);
      }
      if (gi == gj) v = 0.0;
      const T o = cvt_out<T>(v);
      D[(o0 + ty * 4 + x) * ld + gj] = o;
      if (mirror) D[gj * ld + gi] = o;
    }
  }
}

static inline int pad_m(int m) { return (m + 7) / 8 * 8; }

size_t whd_exact_workspace_bytes(int64_t n, int32_t m, int32_t weighted) {
  Carver c(nullptr);
  size_t total = (size_t)n * pad_m(m);
  c.take<int>(64);
  c.take<uint16_t>(total);
  if (weighted) c.take<double>(total);
  const size_t nch = (size_t)(m + CK - 1) / CK;
  c.take<uint32_t>((size_t)n * nch);
  c.take<uint32_t>((size_t)n * nch);
  return c.used();
}

int pairwise_exact_map(int32_t fn, int32_t flags, const int32_t *cm, int64_t n, int32_t m, int32_t missing,
                       const double *w, int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0,
                       int64_t row1, void *workspace, size_t workspace_bytes, cudaStream_t stream, int cyc_G,
                       int cyc_r);

int whd_exact_map(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w,
                  int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0, int64_t row1,
                  void *workspace, size_t workspace_bytes, cudaStream_t stream, int cyc_G, int cyc_r) {
  return pairwise_exact_map(CAS_FN_WEIGHTED_HAMMING_DISTANCE, 0, cm, n, m, missing, w, n_states, D, ld, dtype,
                            row0, row1, workspace, workspace_bytes, stream, cyc_G, cyc_r);
}

int pairwise_exact_map(int32_t fn, int32_t flags, const int32_t *cm, int64_t n, int32_t m, int32_t missing,
                       const double *w, int32_t n_states, void *D, int64_t ld, int32_t dtype, int64_t row0,
                       int64_t row1, void *workspace, size_t workspace_bytes, cudaStream_t stream, int cyc_G,
                       int cyc_r) {
  CAS_REQUIRE(n_states >= 0 && n_states <= 65534, "n_states=%d outside [0, 65534]", n_states);
  CAS_REQUIRE(fn >= 0 && fn <= 5, "unknown pair function %d", fn);
  // hamming_distance takes no weights (reference signature :163-168)
  const bool weighted = (w != nullptr) && fn != CAS_FN_HAMMING_DISTANCE;
  const int ignore_missing = (flags & CAS_FN_FLAG_IGNORE_MISSING) ? 1 : 0;
  CAS_REQUIRE(workspace_bytes >= whd_exact_workspace_bytes(n, m, weighted),
              "workspace too small for whd map");
  Carver carve(workspace);
  const int mp = pad_m(m);
  int *err_flag = carve.take<int>(64);
  uint16_t *codes = carve.take<uint16_t>((size_t)n * mp);
  double *wrow = weighted ? carve.take<double>((size_t)n * mp) : nullptr;
  const int nch = (m + CK - 1) / CK;
  uint32_t *pres = carve.take<uint32_t>((size_t)n * nch);
  uint32_t *cut = carve.take<uint32_t>((size_t)n * nch);

  CAS_CUDA_TRY(cudaMemsetAsync(err_flag, 0, sizeof(int), stream));
  {
    int64_t total = n * (int64_t)mp;
    int64_t blocks = (total + 255) / 256;
    encode_kernel<<<(unsigned)blocks, 256, 0, stream>>>(cm, n, m, mp, missing, w, n_states, codes,
                                                        wrow, err_flag);
    CAS_LAUNCH_CHECK();
    if (fn == CAS_FN_WEIGHTED_HAMMING_DISTANCE || fn == CAS_FN_EXPONENTIAL_NEGATIVE_HAMMING_DISTANCE) {
      const int64_t items = n * (int64_t)nch;
      mask_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(codes, n, m, mp, nch, pres, cut);
      CAS_LAUNCH_CHECK();
    }
  }
  const bool full = (row0 == 0 && row1 == n) && cyc_G <= 0;
  int64_t tiles_r = (row1 - row0 + TILE - 1) / TILE;
  if (cyc_G > 0) {
    // 64-row blocks owned by cyc_r among ceil(n / 64)
    const int64_t nblocks = (n + TILE - 1) / TILE;
    tiles_r = nblocks > cyc_r ? (nblocks - cyc_r + cyc_G - 1) / cyc_G : 0;
    row0 = 0;
    row1 = n;
  }
  const int64_t tiles_c = (n + TILE - 1) / TILE;
  const int64_t ntiles = full ? tiles_r * (tiles_r + 1) / 2 : tiles_r * tiles_c;
  CAS_REQUIRE(ntiles < (int64_t)2147483647, "too many tiles");
  if (ntiles == 0) return CAS_OK;
  dim3 grid((unsigned)ntiles);
#define LAUNCH(T, W, F)                                                                        \
  whd_exact_kernel<T, W, F><<<grid, 256, 0, stream>>>(codes, wrow, n, m, mp, (T *)D, ld, row0, \
                                                      row1, full ? 1 : 0, (int)tiles_c, cyc_G, \
                                                      cyc_r, ignore_missing, pres, cut, nch)
#define LAUNCH_TW(F)                                                                           \
  do {                                                                                         \
    if (dtype == CAS_F64) {                                                                    \
      if (weighted) LAUNCH(double, true, F); else LAUNCH(double, false, F);                    \
    } else {                                                                                   \
      if (weighted) LAUNCH(float, true, F); else LAUNCH(float, false, F);                      \
    }                                                                                          \
  } while (0)
  switch (fn) {
    case 0: LAUNCH_TW(0); break;
    case 1: LAUNCH_TW(1); break;
    case 2: LAUNCH_TW(2); break;
    case 3:
      if (dtype == CAS_F64) LAUNCH(double, false, 3); else LAUNCH(float, false, 3);
      break;
    case 4: LAUNCH_TW(4); break;
    default: LAUNCH_TW(5); break;
  }
#undef LAUNCH_TW
#undef LAUNCH
  CAS_LAUNCH_CHECK();
  return CAS_OK;
}

// device flag written by encode_kernel (first int of the workspace)
int whd_read_error_flag(const void *workspace, cudaStream_t stream, int *flag) {
  CAS_CUDA_TRY(cudaMemcpyAsync(flag, workspace, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CAS_CUDA_TRY(cudaStreamSynchronize(stream));
  return CAS_OK;
}

}  // namespace cas
