// Persistent NJ / UPGMA loop: ALL iterations in one cooperative kernel launch.
//
// The multi-kernel loop of agglom.cu pays two kernel launches per iteration; below k ~ 10^4 the
// map is L2-resident and the loop is launch/latency-bound (~20 us per iteration).  This kernel keeps
// a co-resident grid (one 1024-thread CTA per SM) alive for the whole loop and separates the phases of an
// iteration with two grid-wide barriers:
//
//   P1 scan   : the same screened scan as scan_kernel (agglom_common.cuh), map loads through L2
//               (ld.global.cg: other CTAs rewrite the map between phases); each CTA publishes its
//               candidate together with everything the merge needs (d_ij, cluster sizes), so that ...
//   -- barrier --
//   P2 select : ... every CTA reduces all candidates redundantly (identical result, no second
//               barrier, and no read of state that P3 mutates);
//   P3 merge  : the arithmetic of merge_kernel, 256-wide chunks of v dealt to CTAs; per-chunk partial
//               sums of the new row;
//   -- barrier --
//   every CTA adds the partial sums in the fixed chunk order (same value everywhere) and stores the
//   merged node's row sum, then starts the next scan.
//
// Results are bit-identical to the multi-kernel loop (same per-element arithmetic, same
// deterministic reductions); tests compare both.  Used for fast-mode NJ and for UPGMA; strict NJ
// keeps the multi-kernel loop (its sequential row-sum pass has its own kernel).
#include "agglom_common.cuh"

namespace cas {

struct CandX {
  double q;
  unsigned long long key;
  double d;      // d(si, sj)
  int si, sj;    // slots (row, column of the lower triangle)
  int zlo, zhi;  // cluster sizes of the lower / higher slot
};

struct LoopArgs {
  void *D;
  int64_t ld;
  int n;
  int iters;          // number of joins to perform
  double *R;
  int32_t *ids;
  int32_t *sizes;     // UPGMA, else nullptr
  CandX *cand;        // [gridDim]
  double *partial;    // [ceil(n/256)]
  unsigned long long *rmax_bits;
  unsigned *bar_count;
  unsigned *bar_gen;
  int *error;
  int32_t *joins;
  int target_ctas;    // CTA count the tile planner aims for (== gridDim.x)
};

constexpr int PT = 1024;           // threads per CTA: one CTA per SM keeps the grid barriers cheap
constexpr int PWARPS = PT / 32;
constexpr int PGROUPS = PT / MERGE_THREADS;  // 256-thread groups; each owns one merge chunk at a time

// sense-reversing grid barrier on two global words; gives up after ~2 s (error flag) instead of hanging
__device__ __forceinline__ void grid_barrier(unsigned *count, unsigned *gen, int *error) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned g = *((volatile unsigned *)gen);
    __threadfence();
    if (atomicAdd(count, 1u) == gridDim.x - 1) {
      atomicExch(count, 0u);
      __threadfence();
      atomicAdd(gen, 1u);
    } else {
      const long long t0 = clock64();
      while (*((volatile unsigned *)gen) == g) {
        if (clock64() - t0 > 4000000000LL) { atomicExch(error, 1); break; }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ bool candx_better(const CandX &a, const CandX &b) {
  return (a.q < b.q) || (a.q == b.q && a.key < b.key);
}

template <typename T, bool NJ>
__global__ void __launch_bounds__(PT, 1) loop_kernel(LoopArgs a) {
  __shared__ __align__(16) float sU[NJ ? TC : 4];
  __shared__ CandX sX[PWARPS];
  __shared__ double sPart[PWARPS];
  __shared__ unsigned long long sMax[PWARPS];
  __shared__ double sChunk[NJ ? 1024 : 1];  // staging of the per-chunk partial sums
  __shared__ int sPlan[4];

  T *D = static_cast<T *>(a.D);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t ld = a.ld;

  for (int t = 0; t < a.iters; ++t) {
    const int k = a.n - t;
    const double tinv = 1.0 / (double)(k - 2);
    // ---- plan the tiling of this iteration (rows per tile shrink while there are too few tiles)
    if (threadIdx.x == 0) {
      const int ncb = (k + TC - 1) / TC;
      int tr = 32, nrb = 0, ntiles = 0;
      for (;;) {
        nrb = (k + tr - 1) / tr;
        long long cnt = 0;
        for (int cb = 0; cb < ncb; ++cb) {
          const int first = (cb * TC + 1) / tr;
          if (nrb > first) cnt += nrb - first;
        }
        ntiles = (int)cnt;
        if (ntiles >= 2 * (int)gridDim.x || tr == 8) break;
        tr >>= 1;
      }
      sPlan[0] = tr; sPlan[1] = nrb; sPlan[2] = ntiles;
    }
    __syncthreads();
    const int tr = sPlan[0], nrb = sPlan[1], ntiles = sPlan[2];

    // ---- P1: scan
    const double umax = NJ ? tinv * __longlong_as_double(__ldcg(a.rmax_bits)) : 0.0;
    double bq = INFINITY, b2 = INFINITY;
    int bi = -1, bj = -1;
    int staged_cb = -1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int cb = 0, rem = tile;
      for (;;) {
        const int first = (cb * TC + 1) / tr;
        const int cnt = max(nrb - first, 0);
        if (rem < cnt) { rem += first; break; }
        rem -= cnt;
        ++cb;
      }
      const int c0 = cb * TC;
      const int c1 = min(c0 + TC, k);
      if (NJ && cb != staged_cb) {
        __syncthreads();
        for (int x = threadIdx.x; x < TC; x += PT)
          sU[x] = (c0 + x < c1) ? __double2float_rn(__dmul_rn(tinv, __ldcg(a.R + c0 + x))) : 0.f;
        staged_cb = cb;
        __syncthreads();
      }
      const int r0 = rem * tr;
      const int r1 = min(r0 + tr, k);
      // row sums come from L2 (no L1: another CTA rewrote them); fetch the next row's while this one streams
      double Rnext = (NJ && r0 + warp < r1) ? __ldcg(a.R + r0 + warp) : 0.0;
      for (int i = r0 + warp; i < r1; i += PWARPS) {
        const double Ri = Rnext;
        if (NJ && i + PWARPS < r1) Rnext = __ldcg(a.R + i + PWARPS);
        const int jend = min(c1, i);
        if (jend <= c0) continue;
        scan_row_segment<T, NJ, true>(D + (int64_t)i * ld, i, c0, jend, lane, sU, a.R, Ri, tinv, umax, a.ids,
                                      bq, bi, bj, b2);
      }
    }
    // CTA candidate, published with everything the merge needs
    CandX best;
    best.q = bq; best.si = bi; best.sj = bj; best.d = 0.0; best.zlo = 1; best.zhi = 1;
    best.key = (bi >= 0) ? make_key(__ldcg(a.ids + bi), __ldcg(a.ids + bj)) : ~0ull;
    auto warp_reduce = [&](CandX &c) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        CandX o;
        o.q = __shfl_down_sync(0xffffffffu, c.q, off);
        o.key = __shfl_down_sync(0xffffffffu, c.key, off);
        o.d = __shfl_down_sync(0xffffffffu, c.d, off);
        o.si = __shfl_down_sync(0xffffffffu, c.si, off);
        o.sj = __shfl_down_sync(0xffffffffu, c.sj, off);
        o.zlo = __shfl_down_sync(0xffffffffu, c.zlo, off);
        o.zhi = __shfl_down_sync(0xffffffffu, c.zhi, off);
        if (candx_better(o, c)) c = o;
      }
    };
    warp_reduce(best);
    if (lane == 0) sX[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      CandX x = sX[0];
      for (int w = 1; w < PWARPS; ++w)
        if (candx_better(sX[w], x)) x = sX[w];
      if (x.si >= 0) {
        const int lo = min(x.si, x.sj), hi = max(x.si, x.sj);
        x.d = (double)__ldcg(D + (int64_t)lo * ld + hi);
        if (a.sizes) { x.zlo = __ldcg(a.sizes + lo); x.zhi = __ldcg(a.sizes + hi); }
      }
      a.cand[blockIdx.x] = x;
    }
    grid_barrier(a.bar_count, a.bar_gen, a.error);

    // ---- P2: every CTA selects the pair (nothing read here is written in P3)
    best.q = INFINITY; best.key = ~0ull; best.d = 0.0; best.si = -1; best.sj = -1; best.zlo = 1; best.zhi = 1;
    for (int x = threadIdx.x; x < (int)gridDim.x; x += PT) {
      CandX o;
      o.q = __ldcg(&a.cand[x].q); o.key = __ldcg(&a.cand[x].key); o.d = __ldcg(&a.cand[x].d);
      o.si = __ldcg(&a.cand[x].si); o.sj = __ldcg(&a.cand[x].sj);
      o.zlo = __ldcg(&a.cand[x].zlo); o.zhi = __ldcg(&a.cand[x].zhi);
      if (candx_better(o, best)) best = o;
    }
    warp_reduce(best);
    if (lane == 0) sX[warp] = best;
    __syncthreads();
    best = sX[0];
    for (int w = 1; w < PWARPS; ++w)
      if (candx_better(sX[w], best)) best = sX[w];
    __syncthreads();
    const int lo = min(best.si, best.sj), hi = max(best.si, best.sj), last = k - 1;
    const double dij = best.d;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      a.joins[2 * t] = (int)(best.key >> 32);  // key = (min id << 32) | max id
      a.joins[2 * t + 1] = (int)(best.key & 0xffffffffu);
    }

    // ---- P3: merge (arithmetic of merge_kernel); every 256-thread group takes one 256-slot chunk
    const int nchunks = (k + MERGE_THREADS - 1) / MERGE_THREADS;
    const int group = threadIdx.x / MERGE_THREADS, gtid = threadIdx.x % MERGE_THREADS;
    for (int chunk0 = blockIdx.x * PGROUPS; chunk0 < nchunks; chunk0 += gridDim.x * PGROUPS) {
      const int chunk = chunk0 + group;
      const int v = chunk * MERGE_THREADS + gtid;
      double contrib = 0.0, mine = 0.0;
      if (chunk < nchunks && v < k) {
        if (v == lo) {
          D[(int64_t)lo * ld + lo] = (T)0;
          a.ids[lo] = a.n + t;
          if (a.sizes) a.sizes[lo] = best.zlo + best.zhi;
        } else if (v == hi) {
          D[(int64_t)hi * ld + hi] = (T)0;
        } else {
          const double da = (double)__ldcg(D + (int64_t)lo * ld + v);
          const double db = (double)__ldcg(D + (int64_t)hi * ld + v);
          double nv;
          if (NJ) {
            nv = __dmul_rn(0.5, __dsub_rn(__dadd_rn(da, db), dij));
          } else {
            const double zi = (double)best.zlo, zj = (double)best.zhi;
            nv = __ddiv_rn(__dadd_rn(__dmul_rn(zi, da), __dmul_rn(zj, db)), (double)(best.zlo + best.zhi));
          }
          const T nvs = round_to<T>(nv);
          contrib = (double)nvs;
          double Rv = 0.0;
          if (NJ) Rv = __dadd_rn(__dsub_rn(__dsub_rn(__ldcg(a.R + v), da), db), contrib);
          if (hi != last && v != last) {
            const T l = __ldcg(D + (int64_t)last * ld + v);
            D[(int64_t)hi * ld + v] = l;
            D[(int64_t)v * ld + hi] = l;
          }
          D[(int64_t)lo * ld + v] = nvs;
          D[(int64_t)v * ld + lo] = nvs;
          if (v == last && hi != last) {
            D[(int64_t)hi * ld + lo] = nvs;
            D[(int64_t)lo * ld + hi] = nvs;
            D[(int64_t)hi * ld + hi] = (T)0;
            a.ids[hi] = __ldcg(a.ids + last);
            if (a.sizes) a.sizes[hi] = __ldcg(a.sizes + last);
            if (NJ) { a.R[hi] = Rv; mine = Rv; }
          } else if (NJ) {
            a.R[v] = Rv;
            mine = Rv;
          }
        }
      }
      if (NJ) {
        // max |row sum| of the group's slots: shuffle tree, then one atomic per 256-slot chunk (thousands
        // of same-address atomics per iteration would serialise in L2)
        unsigned long long mb = (unsigned long long)__double_as_longlong(fabs(mine));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          contrib = __dadd_rn(contrib, __shfl_down_sync(0xffffffffu, contrib, off));
          const unsigned long long ob = __shfl_down_sync(0xffffffffu, mb, off);
          mb = ob > mb ? ob : mb;
        }
        if (lane == 0) { sPart[warp] = contrib; sMax[warp] = mb; }
        __syncthreads();
        if (gtid == 0 && chunk < nchunks) {
          double sum = 0.0;  // the 8 warps of this group, in order (same as merge_kernel)
          unsigned long long gm = 0;
          for (int w = 0; w < MERGE_THREADS / 32; ++w) {
            sum = __dadd_rn(sum, sPart[group * (MERGE_THREADS / 32) + w]);
            const unsigned long long o = sMax[group * (MERGE_THREADS / 32) + w];
            gm = o > gm ? o : gm;
          }
          a.partial[chunk] = sum;
          atomicMax(a.rmax_bits, gm);
        }
        __syncthreads();
      }
    }
    grid_barrier(a.bar_count, a.bar_gen, a.error);

    // ---- row sum of the merged node: every CTA adds the chunk partials in the same fixed order
    if (NJ) {
      for (int c = threadIdx.x; c < nchunks; c += PT) sChunk[c] = __ldcg(a.partial + c);
      __syncthreads();
      if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int c = 0; c < nchunks; ++c) sum = __dadd_rn(sum, sChunk[c]);
        a.R[lo] = sum;  // identical value from every CTA
        atomicMax(a.rmax_bits, (unsigned long long)__double_as_longlong(fabs(sum)));
        __threadfence();
      }
      __syncthreads();
    }
  }
}

// ---- host -------------------------------------------------------------------------------------
template <typename T, bool NJ>
static int launch_loop(LoopArgs &args, int grid, cudaStream_t stream) {
  void *params[] = {&args};
  CAS_CUDA_TRY(cudaLaunchCooperativeKernel((const void *)loop_kernel<T, NJ>, dim3(grid), dim3(PT), params, 0,
                                           stream));
  count_launch();
  return CAS_OK;
}

// largest co-resident grid for the loop kernel (0 if cooperative launch is unsupported)
template <typename T, bool NJ>
static int coop_grid() {
  int dev = 0, coop = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, loop_kernel<T, NJ>, PT, 0) != cudaSuccess) return 0;
  if (per_sm < 1) return 0;
  return sm_count();  // one 1024-thread CTA per SM
}

int persistent_loop(void *D, int64_t n, int64_t ld, int32_t dtype, int32_t algo, int64_t iters, double *R,
                    int32_t *ids, int32_t *sizes, void *cand_scratch, double *partial,
                    unsigned long long *rmax_bits, unsigned *bar_words, int *error, int32_t *joins,
                    cudaStream_t stream) {
  if (n > 1024 * (int64_t)MERGE_THREADS) {  // sChunk staging holds 1024 chunk partials
    set_error("persistent loop supports n <= 262144");
    return CAS_ERR_UNSUPPORTED;
  }
  LoopArgs a;
  a.D = D; a.ld = ld; a.n = (int)n; a.iters = (int)iters; a.R = R; a.ids = ids;
  a.sizes = (algo == CAS_ALGO_UPGMA) ? sizes : nullptr;
  a.cand = static_cast<CandX *>(cand_scratch); a.partial = partial; a.rmax_bits = rmax_bits;
  a.bar_count = bar_words; a.bar_gen = bar_words + 1; a.error = error; a.joins = joins;
  int grid;
#define RUN(T, NJ)                                                          \
  do {                                                                      \
    grid = coop_grid<T, NJ>();                                              \
    if (grid <= 0) { set_error("cooperative launch unavailable"); return CAS_ERR_UNSUPPORTED; } \
    a.target_ctas = grid;                                                   \
    return launch_loop<T, NJ>(a, grid, stream);                             \
  } while (0)
  if (algo == CAS_ALGO_NJ) {
    if (dtype == CAS_F64) RUN(double, true); else RUN(float, true);
  } else {
    if (dtype == CAS_F64) RUN(double, false); else RUN(float, false);
  }
#undef RUN
}

size_t persistent_cand_bytes() { return sizeof(CandX) * (size_t)MAX_SCAN_CTAS; }

}  // namespace cas
