// "Exact" (adaptive strict) Neighbor-Joining: the reference's join list without re-accumulating every row sum.
//
// The reference evaluates q_ij = d_ij - (1/(k-2)) (rs_i + rs_j) with rs = d.sum(axis=1), a left-to-right float64
// accumulation in row (= id) order (cassiopeia/solver/NeighborJoiningSolver.py:158-171), and takes the first
// row-major minimum (:138-140).  Re-accumulating all k sums costs 8 k^2 bytes per iteration (the strict mode).
// Here the sums are maintained incrementally with rigorous error bounds (agglom_common.cuh, "exact mode"), the scan
// runs on the maintained criterion q~, and only when a second pair lies within the error window W of the minimum
// -- an exact tie or a near-tie the maintained sums cannot decide -- the LAST CTA of the scan
//   A. collects every pair with q~ <= q~min + W and marks its two rows
//      (pruned scan: only the rows on this iteration's work list, guided by the bounds the scan just refreshed;
//       dense scan: only the tiles of the CTAs whose own best lies inside the window),
//   B. evaluates the reference's sequential sum for the marked rows (8 rows per pass: all threads gather
//      d[row][order[p]] into shared memory 256 positions at a time while 8 lanes walk the dependent add chains),
//   C. re-evaluates the collected pairs with those sums in the reference's expression and applies the
//      (value, id_lo, id_hi) rule.
// On simulated lineage data 0.1-0.3 % of the iterations are ambiguous and 3-7 rows are involved
// (oracle/oracle.c: oracle_nj_adaptive is the CPU model, checked against the reference loop).
#pragma once

#include "agglom_pruned.cuh"

namespace cas {

struct ExactArgs {
  double *Eb;                  // [n] by slot: |R_v - real row sum| <= Eb_v
  double *Ab;                  // [n] by slot: sum_j |d_vj| <= Ab_v
  unsigned long long *xw;      // running maxima (agglom_common.cuh: XW_*), = the rmax_bits array
  int32_t *mark;               // [n] epoch tags (iteration + 1) of marked slots
  int32_t *mlist;              // [n] marked slots of the iteration being resolved
  double *Rex;                 // [n] by slot: sequential sums of the marked rows
  unsigned *mcount;            // marked rows (left at 0)
  unsigned long long *xstats;  // [0] ambiguous iterations, [1] rows resolved, [2] largest row set,
                               // [3] iterations whose pair differs from arg-min q~
  const int32_t *order;        // live slots in id order, this iteration
  int32_t *xlog;               // [XLOG_CAP][4] diagnostics of the first ambiguous iterations: t, k, rows, window
  // Strict tail.  Towards the end of a solve on low-information data the remaining nodes form a star: all q are
  // equal in exact arithmetic, every pair lies inside any error window, and the reference's choice is decided by
  // the rounding of its sequential sums -- of ALL rows.  The first ambiguous iteration that involves more than
  // XR_HEAVY rows therefore raises `strict_flag`; from then on (for live sizes <= STRICT_TAIL_K, where the gated
  // launch exists) rowsum_seq_kernel re-accumulates every row sum sequentially before the scan, stamps
  // `strict_epoch` with the iteration, and the scan runs with a zero window and no resolver: the strict mode's
  // arithmetic, at its cost (8 k^2 bytes per iteration), only where it is needed.
  unsigned *strict_flag;
  int *strict_epoch;
};
constexpr int XLOG_CAP = 4096;
constexpr int XR_HEAVY = 64;
constexpr int STRICT_TAIL_K = 16384;

constexpr int XR_ROWS = 8;     // rows whose sequential sums are evaluated together
constexpr int XR_POS = 256;    // positions staged per step (= threads of the CTA)

__device__ __forceinline__ void exact_mark(const ExactArgs &x, int slot, int epoch) {
  if (atomicExch(x.mark + slot, epoch) != epoch) x.mlist[atomicAdd(x.mcount, 1u)] = slot;
}

// the maintained criterion, bit for bit the expression of the scans
__device__ __forceinline__ double exact_q(double d, double tinv, double ri, double rj) {
  return __dsub_rn(d, __dmul_rn(tinv, __dadd_rn(ri, rj)));
}

// one row: columns [j0, j1) of row i (all below the diagonal), one warp, lane-strided
__device__ __forceinline__ void exact_collect_span(const double *__restrict__ row, int i, int j0, int j1, int lane,
                                                   const double *R, double Ri, double tinv, double capq,
                                                   const ExactArgs &x, int epoch) {
  for (int j = j0 + lane; j < j1; j += 32) {
    const double q = exact_q(row[j], tinv, Ri, R[j]);
    if (q <= capq) {
      exact_mark(x, i, epoch);
      exact_mark(x, j, epoch);
    }
  }
}

// B: sum = ((0 + d[s][o_0]) + d[s][o_1]) + ... for M rows.  rowptr(m) = address of the m-th row's elements (indexed
// by slot), out(m, sum) stores the result.  All threads of a 256-thread CTA; `order` = live slots in id order.
template <typename RowPtr, typename Out>
__device__ __forceinline__ void exact_rowsums_generic(int k, const int32_t *order, int M, double (*tile)[XR_POS + 1],
                                                      RowPtr rowptr, Out out) {
  const int tid = threadIdx.x;
  for (int p0 = 0; p0 < M; p0 += XR_ROWS) {
    const int nr = min(XR_ROWS, M - p0);
    const double *rows[XR_ROWS];
#pragma unroll
    for (int r = 0; r < XR_ROWS; ++r) rows[r] = rowptr(p0 + min(r, nr - 1));
    double pre[XR_ROWS];
    auto fetch = [&](int pos0) {
      const int p = pos0 + tid;
      const int slot = p < k ? order[p] : -1;
#pragma unroll
      for (int r = 0; r < XR_ROWS; ++r) pre[r] = (r < nr && slot >= 0) ? rows[r][slot] : 0.0;
    };
    fetch(0);
    double acc = 0.0;
    for (int pos0 = 0; pos0 < k; pos0 += XR_POS) {
#pragma unroll
      for (int r = 0; r < XR_ROWS; ++r) tile[r][tid] = pre[r];
      __syncthreads();
      if (pos0 + XR_POS < k) fetch(pos0 + XR_POS);  // in flight while the chains advance
      if (tid < nr) {
        const double *t = tile[tid];
        const int cnt = min(XR_POS, k - pos0);
        int e = 0;
        for (; e + 8 <= cnt; e += 8) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = t[e + u];
#pragma unroll
          for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, v[u]);
        }
        for (; e < cnt; ++e) acc = __dadd_rn(acc, t[e]);
      }
      __syncthreads();
    }
    if (tid < nr) out(p0 + tid, acc);
  }
  __syncthreads();
}

__device__ __forceinline__ void exact_rowsums(const double *__restrict__ D, int64_t ld, int k, const ExactArgs &x,
                                              int M, double (*tile)[XR_POS + 1]) {
  exact_rowsums_generic(
      k, x.order, M, tile, [&](int m) { return D + (int64_t)x.mlist[m] * ld; },
      [&](int m, double sum) { x.Rex[x.mlist[m]] = sum; });
}

// C: the reference's pair among the marked rows.  Returns it in (si, sj) (unchanged when nothing qualifies, which
// cannot happen: the arg-min of q~ always does).  All threads; the result is identical in every thread.
__device__ __forceinline__ void exact_decide(const double *__restrict__ D, int64_t ld, const double *R,
                                             const int32_t *ids, double tinv, double capq, const ExactArgs &x, int M,
                                             int &si, int &sj) {
  __shared__ double sQ[8];
  __shared__ unsigned long long sK[8];
  __shared__ int sI[8], sJ[8];
  double bq = INFINITY;
  unsigned long long bk = ~0ull;
  int bi = -1, bj = -1;
  for (int a = 1; a < M; ++a) {
    const int sa = x.mlist[a];
    for (int b = threadIdx.x; b < a; b += blockDim.x) {
      const int sb = x.mlist[b];
      const int i = max(sa, sb), j = min(sa, sb);
      const double d = D[(int64_t)i * ld + j];
      if (exact_q(d, tinv, R[i], R[j]) > capq) continue;
      const double q = exact_q(d, tinv, x.Rex[i], x.Rex[j]);
      const unsigned long long key = make_key(ids[i], ids[j]);
      if (cand_better(q, key, bq, bk)) { bq = q; bk = key; bi = i; bj = j; }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double oq = __shfl_xor_sync(0xffffffffu, bq, off);
    const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off), oj = __shfl_xor_sync(0xffffffffu, bj, off);
    if (cand_better(oq, ok, bq, bk)) { bq = oq; bk = ok; bi = oi; bj = oj; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sQ[warp] = bq; sK[warp] = bk; sI[warp] = bi; sJ[warp] = bj; }
  __syncthreads();
  bq = sQ[0]; bk = sK[0]; bi = sI[0]; bj = sJ[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
    if (cand_better(sQ[w], sK[w], bq, bk)) { bq = sQ[w]; bk = sK[w]; bi = sI[w]; bj = sJ[w]; }
  __syncthreads();
  if (bi >= 0) { si = bi; sj = bj; }
}

// B + C + statistics, after the rows have been marked.  All threads of the last CTA.
static __device__ __noinline__ void exact_finish(const double *D, int64_t ld, int k, const double *R, const int32_t *ids,
                                          double tinv, double capq, ExactArgs x, int *psi, int *psj, int t_iter) {
  __shared__ double tile[XR_ROWS][XR_POS + 1];
  __syncthreads();  // every mark of this CTA has been issued
  const int M = (int)__ldcg(x.mcount);
  exact_rowsums(D, ld, k, x, M, tile);
  int si = *psi, sj = *psj;
  const int si0 = max(si, sj), sj0 = min(si, sj);
  exact_decide(D, ld, R, ids, tinv, capq, x, M, si, sj);
  if (threadIdx.x == 0) {
    *x.mcount = 0u;
    const unsigned long long slot = x.xstats[0];
    if (x.xlog && slot < XLOG_CAP) {
      x.xlog[4 * slot + 0] = t_iter;
      x.xlog[4 * slot + 1] = k;
      x.xlog[4 * slot + 2] = M;
      x.xlog[4 * slot + 3] = __float_as_int((float)(capq));
    }
    if (M > XR_HEAVY && x.strict_flag) *x.strict_flag = 1u;
    atomicAdd(x.xstats + 0, 1ull);
    atomicAdd(x.xstats + 1, (unsigned long long)M);
    atomicMax(x.xstats + 2, (unsigned long long)M);
    if (max(si, sj) != si0 || min(si, sj) != sj0) atomicAdd(x.xstats + 3, 1ull);
  }
  *psi = si;
  *psj = sj;
}

// A for the pruned scan.  Rows that are not on this iteration's work list were proven unable to hold a pair below
// gbest + window (>= capq) by the row test; listed rows carry bounds the scan has just refreshed.  E and beta were
// written by other CTAs: read through L2.  Warps gw, gw + nw, ... of the caller share the listed rows;
// rowof(ridx, i, row) maps a work-list entry (the row's index into E / beta) to its slot and its elements.
template <typename RowOf>
__device__ __forceinline__ void exact_collect_listed(const double *R, double tinv, int k, const PruneArgs &p,
                                                     const ExactArgs &x, double capq, int epoch, unsigned nfront,
                                                     unsigned nback, int gw, int nw, RowOf rowof) {
  const int lane = threadIdx.x & 31;
  const float cnow = prune_cnow(p);
  const float umaxf = prune_umaxf(p, tinv);
  const float capf = __double2float_ru(capq);
  auto impossible = [&](float e, float uif) {
    const float mag = fabsf(capf) + fabsf(uif) + fabsf(e) + cnow + umaxf;
    return (e - cnow) > (capf + uif) + 2e-6f * mag;
  };
  const unsigned nrows = nfront + nback;
  for (unsigned r = gw; r < nrows; r += nw) {
    const int ridx = r < nfront ? p.worklist[r] : p.worklist[p.wlast - (int)(r - nfront)];
    int i = -1;
    const double *row = nullptr;
    rowof(ridx, i, row);
    if (i < 1 || i >= k) continue;
    const double Ri = R[i];
    const float uif = (float)__dmul_rn(tinv, Ri);
    if (impossible(dec_f32(__ldcg(p.beta + ridx)), uif)) continue;
    const unsigned *erow = p.E + (int64_t)ridx * p.stride;
    const int nsub = (i + SW - 1) / SW;
    for (int sb = 0; sb < nsub; sb += 32) {
      const int sgm = sb + lane;
      const bool need = sgm < nsub && !impossible(dec_f32(__ldcg(erow + sgm)), uif);
      unsigned flags = __ballot_sync(0xffffffffu, need);
      while (flags) {
        const int s = sb + __ffs(flags) - 1;
        flags &= flags - 1;
        exact_collect_span(row, i, s * SW, min(s * SW + SW, i), lane, R, Ri, tinv, capq, x, epoch);
      }
    }
  }
}

static __device__ __noinline__ void exact_collect_pruned(const double *D, int64_t ld, int k, const double *R, double tinv,
                                                  PruneArgs p, ExactArgs x, double capq, int epoch, unsigned nfront,
                                                  unsigned nback) {
  exact_collect_listed(R, tinv, k, p, x, capq, epoch, nfront, nback, (int)(threadIdx.x >> 5), (int)(blockDim.x >> 5),
                       [&](int ridx, int &i, const double *&row) {
                         i = ridx;
                         row = D + (int64_t)ridx * ld;
                       });
}

}  // namespace cas
