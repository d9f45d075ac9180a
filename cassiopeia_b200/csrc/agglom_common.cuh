// Device-side pieces shared by the single-GPU loop (agglom.cu) and the row-sharded loop (sharded.cu).
#pragma once

#include <math.h>

#include "common.cuh"

namespace cas {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;
constexpr int TC = 2048;  // columns per tile (R_j block staged in shared memory: 16 KB)
constexpr int MERGE_THREADS = 256;
constexpr int MAX_SCAN_CTAS = 2048;

struct Cand {
  double q;
  unsigned long long key;  // (id_lo << 32) | id_hi ; ~0 = empty
  int si, sj;              // slots, si > sj (row, column of the lower triangle)
  double q2;               // runner-up: smallest exactly evaluated value of any other pair (+inf if none);
                           // exact whenever it lies within the screening margin of q (near-tie diagnostics)
};

struct Sel {
  int lo, hi;       // slots, lo < hi
  int size_lo, size_hi;
  double dij;
  int id_lo, id_hi;
};

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<double> { static constexpr int N = 2; };

// Streaming 128-bit loads of the map.  COH = false: the map is read-only for the whole kernel
// (non-coherent path, no L1 allocation).  COH = true: the persistent loop kernel rewrites the map
// between phases, so loads must be served from L2 (ld.global.cg) to see other CTAs' stores.
template <bool COH = false>
__device__ __forceinline__ void ld_stream(const float *p, float (&v)[4]) {
  if (COH)
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p) : "memory");
  else
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}
template <bool COH = false>
__device__ __forceinline__ void ld_stream(const double *p, double (&v)[2]) {
  if (COH)
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p) : "memory");
  else
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(v[0]), "=d"(v[1]) : "l"(p));
}
template <bool COH, typename X>
__device__ __forceinline__ X ld_state(const X *p) { return COH ? __ldcg(p) : *p; }

__device__ __forceinline__ unsigned long long make_key(int ida, int idb) {
  unsigned lo = (unsigned)min(ida, idb), hi = (unsigned)max(ida, idb);
  return ((unsigned long long)lo << 32) | hi;
}

// rare path: candidate ties (or beats) the thread's running best
template <bool COH = false>
__device__ __forceinline__ void tie_path(double q, int i, int j, double &bq, int &bi, int &bj, double &b2,
                                         const int32_t *ids) {
  if (bi < 0) {
    bq = q; bi = i; bj = j;
    return;
  }
  if (q < bq) {
    b2 = fmin(b2, bq);  // the old best becomes a runner-up
    bq = q; bi = i; bj = j;
    return;
  }
  b2 = q;  // exact tie: runner-up equals the best
  unsigned long long kn = make_key(ld_state<COH>(ids + i), ld_state<COH>(ids + j));
  unsigned long long ko = make_key(ld_state<COH>(ids + bi), ld_state<COH>(ids + bj));
  if (kn < ko) { bi = i; bj = j; }
}

__device__ __forceinline__ bool cand_better(double qa, unsigned long long ka, double qb,
                                            unsigned long long kb) {
  // empty candidates carry key ~0 and q = +inf
  return (qa < qb) || (qa == qb && ka < kb);
}

// Screening arithmetic.  The exact criterion of a pair is the reference's float64 expression
//   q = d_ij - ( (1/(k-2)) * (rs_i + rs_j) )          (NJ)        q = d_ij  (UPGMA)
// Evaluating it for every element costs ~30 instructions per element and makes the scan
// issue-bound (ncu, profiles/r1_scan_before.md).  Instead each element is screened in
// float32: qf = fl32(d) - fl32(u_j) with u_j = (1/(k-2)) * rs_j staged in shared memory, against
// a per-row threshold thr >= (best + u_i) + E, where E = 2^-20 * (max|u| + |best + u_i|)
// bounds every rounding of the screen (proof in DESIGN.md "scan screening").  Only elements
// that pass are re-evaluated with the exact float64 expression and the lexicographic
// (value, id_lo, id_hi) rule, so the selected pair is identical to evaluating everything
// exactly.
__device__ __forceinline__ float screen_value(float d) { return d; }
__device__ __forceinline__ float screen_value(double d) { return __double2float_rd(d); }

// running maximum of |row sum| (non-negative doubles order like their bit patterns)
__device__ __forceinline__ void warp_atomic_absmax(unsigned long long *dst, double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(fabs(x));
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, b, off);
    b = o > b ? o : b;
  }
  // the maximum only ever grows: almost every warp can skip the (single-address) atomic
  if ((threadIdx.x & 31) == 0 && b > __ldcg(dst)) atomicMax(dst, b);
}

// Order-preserving unsigned encodings of float / double (unsigned compare == numeric compare), so that
// atomicMin / atomicMax maintain lower / upper bounds that may be negative.
__device__ __forceinline__ unsigned enc_f32(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned e) {
  return __uint_as_float((e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e);
}
__device__ __forceinline__ unsigned long long enc_f64(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long e) {
  return __longlong_as_double((long long)((e >> 63) ? (e & 0x7FFFFFFFFFFFFFFFull) : ~e));
}

// ---- "exact" (adaptive strict) NJ mode -----------------------------------------------------------------
// The reference re-accumulates every row sum sequentially each iteration (NeighborJoiningSolver.py:160).  The
// exact mode MAINTAINS the sums (R_v += new - d_vi - d_vj) together with rigorous bounds
//     |R_v - S_v| <= Eb_v      (S_v the real-number row sum; Eb_v accumulates u'|x| for every rounded result x)
//     sum_j |d_vj| <= Ab_v     (kept with upward-biased arithmetic)
// so that the reference's sequential sum rs_v (k - 1 rounded additions) obeys |rs_v - R_v| <= Ebmax + k u' Abmax
// =: delta, and the reference criterion q and the maintained one q~ differ by at most
//     Delta = 2 t delta + 8 u' (dmax + 2 t (Rmax + delta)),      t = 1 / (k - 2).
// A pair whose q~ exceeds the smallest q~ by more than W = 2 Delta can neither win nor tie under the reference's
// arithmetic; if exactly one pair lies inside the window it is the reference's pair, otherwise the sequential sums
// of just the rows involved are evaluated (exact_resolve) and the reference's rule decides.  The running maxima
// live next to the |row sum| maximum: xw[0] = Rmax, xw[5] = Ebmax, xw[6] = Abmax, xw[7] = max |d| (bit patterns of
// non-negative doubles, maintained with atomicMax).  oracle/oracle.c: oracle_nj_adaptive is the CPU model.
constexpr double EXACT_U1 = 1.1102230246251565e-16 * (1.0 + 9.5367431640625e-07);  // 2^-53 (1 + 2^-20)
constexpr int XW_EBMAX = 5, XW_ABMAX = 6, XW_DABS = 7;

__device__ __forceinline__ double exact_window(const unsigned long long *xw, int k, double tinv) {
  const double rmax = __longlong_as_double((long long)__ldcg(xw + 0));
  const double eb = __longlong_as_double((long long)__ldcg(xw + XW_EBMAX));
  const double ab = __longlong_as_double((long long)__ldcg(xw + XW_ABMAX));
  const double dm = __longlong_as_double((long long)__ldcg(xw + XW_DABS));
  const double delta = eb + (double)k * EXACT_U1 * ab;
  const double Delta = 2.0 * tinv * delta + 8.0 * EXACT_U1 * (dm + 2.0 * tinv * (rmax + delta));
  return 2.0 * Delta * 1.0001;
}
// near-tie window of the diagnostics (1e-6 relative, absolute below 1) around x, never narrower than `wfloor`
// (exact mode: twice the error window, so that every pair the resolver may need has been evaluated)
__device__ __forceinline__ double near_window(double x, double wfloor) {
  return fmax(1.0000001e-6 * fmax(1.0, fabs(x)), wfloor);
}
__device__ __forceinline__ double up30(double x) { return x * (1.0 + 9.313225746154785e-10); }  // x (1 + 2^-30), x >= 0
// all lanes: raise the running maximum *dst to the largest non-negative x of the warp
__device__ __forceinline__ void warp_atomic_max_nonneg(unsigned long long *dst, double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned long long o = __shfl_xor_sync(0xffffffffu, b, off);
    b = o > b ? o : b;
  }
  if ((threadIdx.x & 31) == 0 && b > __ldcg(dst)) atomicMax(dst, b);
}

template <typename T> __device__ __forceinline__ T round_to(double v);
template <> __device__ __forceinline__ float round_to<float>(double v) { return __double2float_rn(v); }
template <> __device__ __forceinline__ double round_to<double>(double v) { return v; }


// Scans columns [c0, jend) of one matrix row (slot i) with one warp: float32 screen against the
// staged u_j block `sU`, exact float64 re-evaluation of the survivors.  (bq, bi, bj) is the
// lane's running best (value, row slot, column slot).
template <typename T, bool NJ, bool COH = false>
__device__ __forceinline__ void scan_row_segment(const T *row, int i, int c0, int jend,
                                                 int lane, const float *sU, const double *R,
                                                 double Ri, double tinv, double umax,
                                                 const int32_t *ids, double &bq, int &bi,
                                                 int &bj, double &b2, double wfloor = 0.0) {
  constexpr int V = VecOf<T>::N;
  constexpr int U = 4;          // independent 128-bit loads in flight per lane
  constexpr int STEP = 32 * V;  // elements per warp-wide load
  const double ui = NJ ? __dmul_rn(tinv, Ri) : 0.0;
  // screening threshold of this row for the lane's current best
  auto threshold = [&]() -> float {
    if (bi < 0) return INFINITY;
    if (!NJ) return __double2float_ru(bq);
    const double B = bq + ui;
    return __double2float_ru(B + 9.5367431640625e-07 * (umax + fabs(B)) + wfloor + 1e-300);
  };
  float thr = threshold();
  for (int jb = c0; jb < jend; jb += STEP * U) {
    T v[U][V];
    const bool whole = (jb + STEP * U <= jend);
    if (whole) {
#pragma unroll
      for (int u = 0; u < U; ++u) ld_stream<COH>(row + jb + u * STEP + lane * V, v[u]);
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = jb + u * STEP + lane * V;
        if (j < jend) {
          ld_stream<COH>(row + j, v[u]);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) v[u][e] = (T)INFINITY;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = jb + u * STEP + lane * V;
      float qf[V];
      if (NJ) {
        float su[V];
        if (V == 4) *reinterpret_cast<float4 *>(su) = *reinterpret_cast<const float4 *>(&sU[j - c0]);
        else *reinterpret_cast<float2 *>(su) = *reinterpret_cast<const float2 *>(&sU[j - c0]);
#pragma unroll
        for (int e = 0; e < V; ++e) qf[e] = __fsub_rn(screen_value(v[u][e]), su[e]);
      } else {
#pragma unroll
        for (int e = 0; e < V; ++e) qf[e] = screen_value(v[u][e]);
      }
      float m = qf[0];
#pragma unroll
      for (int e = 1; e < V; ++e) m = fminf(m, qf[e]);
      if (m <= thr) {
        // rare: exact float64 evaluation of the elements that passed the screen
#pragma unroll
        for (int e = 0; e < V; ++e) {
          if (qf[e] <= thr && (whole || j + e < jend)) {
            double q = (double)v[u][e];
            if (NJ) {
              // q = d_ij - ( (1/(k-2)) * (rs_i + rs_j) ), no FMA  (NeighborJoiningSolver.py:163-170)
              q = __dsub_rn(q, __dmul_rn(tinv, __dadd_rn(Ri, ld_state<COH>(R + j + e))));
            }
            if (q <= bq || bi < 0) {
              tie_path<COH>(q, i, j + e, bq, bi, bj, b2, ids);
              thr = threshold();
            } else {
              b2 = fmin(b2, q);  // passed the screen but lost: a runner-up
            }
          }
        }
      }
    }
  }
}

// merges candidate o into c: the better one wins, the runner-up is the smallest of the rest
__device__ __forceinline__ void cand_merge(Cand &c, const Cand &o) {
  if (cand_better(o.q, o.key, c.q, c.key)) {
    const double r = fmin(o.q2, fmin(c.q2, c.si >= 0 ? c.q : INFINITY));
    c = o;
    c.q2 = r;
  } else {
    c.q2 = fmin(c.q2, fmin(o.q2, o.si >= 0 ? o.q : INFINITY));
  }
}
__device__ __forceinline__ Cand cand_shfl_down(const Cand &c, int off) {
  Cand o;
  o.q = __shfl_down_sync(0xffffffffu, c.q, off);
  o.key = __shfl_down_sync(0xffffffffu, c.key, off);
  o.si = __shfl_down_sync(0xffffffffu, c.si, off);
  o.sj = __shfl_down_sync(0xffffffffu, c.sj, off);
  o.q2 = __shfl_down_sync(0xffffffffu, c.q2, off);
  return o;
}

// Sum of the merge kernels' per-CTA partial sums (and, in exact mode, of the partial sums of absolute values) by ONE
// warp in a FIXED order: lane-strided, then a shuffle tree.  Deterministic and independent of everything but `count`,
// hence identical in the single-GPU and the sharded merge (whose ranks all compute the merged row redundantly).
// Lane 0 returns the sums.  Any element passes through at most count / 32 + 1 + 5 rounded additions here.
__device__ __forceinline__ void sum_partials_one_warp(const double *partial, const double *apartial, int count,
                                                      bool with_abs, double &sum, double &asum) {
  const int lane = threadIdx.x & 31;
  sum = 0.0;
  asum = 0.0;
  // eight loads per lane in flight; out-of-range slots contribute +0.0 (x + 0.0 == x)
  for (int b0 = 0; b0 < count; b0 += 256) {
    double pv[8], av[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int b = b0 + u * 32 + lane;
      pv[u] = b < count ? __ldcg(partial + b) : 0.0;
      av[u] = (with_abs && b < count) ? __ldcg(apartial + b) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      sum = __dadd_rn(sum, pv[u]);
      asum = __dadd_rn(asum, av[u]);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sum = __dadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, off));
    asum = __dadd_rn(asum, __shfl_down_sync(0xffffffffu, asum, off));
  }
}
// bound on the rounded additions any element of the new row passes through: 5 + 8 inside its CTA, then the above
__device__ __forceinline__ double new_row_sum_depth(int count) { return (double)(13 + count / 32 + 8 + 5 + 2); }

// Warp-wide candidate reduction, order-free and therefore identical to any sequence of cand_merge calls: the winner
// is the lexicographic minimum of (q, key), the runner-up the smallest value among everything else (the other lanes'
// winners and every lane's own runner-up).  Three butterfly minima instead of five rounds of full merges.  Every
// lane returns the result.
__device__ __forceinline__ Cand warp_cand_reduce(Cand c) {
  const int lane = threadIdx.x & 31;
  const bool has = c.si >= 0;
  double qm = has ? c.q : INFINITY;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) qm = fmin(qm, __shfl_xor_sync(0xffffffffu, qm, off));
  unsigned long long km = (has && c.q == qm) ? c.key : ~0ull;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, km, off);
    km = o < km ? o : km;
  }
  const unsigned wb = __ballot_sync(0xffffffffu, has && c.q == qm && c.key == km);
  const int wl = wb ? __ffs(wb) - 1 : 0;
  double r = (wb && lane == wl) ? c.q2 : fmin(c.q2, has ? c.q : INFINITY);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) r = fmin(r, __shfl_xor_sync(0xffffffffu, r, off));
  Cand o;
  o.q = wb ? qm : INFINITY;
  o.key = wb ? km : ~0ull;
  o.si = __shfl_sync(0xffffffffu, c.si, wl);
  o.sj = __shfl_sync(0xffffffffu, c.sj, wl);
  if (!wb) { o.si = -1; o.sj = -1; }
  o.q2 = r;
  return o;
}

// Reduces the lanes' bests to one candidate per CTA (every thread of warp 0 returns it; the other warps return
// their own warp's candidate).  Contains one __syncthreads.
template <bool COH = false>
__device__ __forceinline__ Cand block_best(double bq, int bi, int bj, double b2, const int32_t *ids, Cand *sCand) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Cand c;
  c.q = bq; c.si = bi; c.sj = bj; c.q2 = b2;
  c.key = (bi >= 0) ? make_key(ld_state<COH>(ids + bi), ld_state<COH>(ids + bj)) : ~0ull;
  c = warp_cand_reduce(c);
  if (lane == 0) sCand[warp] = c;
  __syncthreads();
  if (warp == 0) {
    Cand e;
    e.q = INFINITY; e.key = ~0ull; e.si = -1; e.sj = -1; e.q2 = INFINITY;
    if (lane < (int)(blockDim.x >> 5)) e = sCand[lane];
    c = warp_cand_reduce(e);
  }
  return c;
}

// All threads of one CTA reduce `count` candidates written by other CTAs / ranks (read through L2).
__device__ __forceinline__ Cand reduce_candidates(const Cand *cand, int count, Cand *sCand) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Cand c;
  c.q = INFINITY; c.key = ~0ull; c.si = -1; c.sj = -1; c.q2 = INFINITY;
  for (int x0 = 0; x0 < count; x0 += 2 * blockDim.x) {  // two candidates per thread in flight
    Cand o[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int x = x0 + u * blockDim.x + threadIdx.x;
      o[u].q = INFINITY; o[u].key = ~0ull; o[u].si = -1; o[u].sj = -1; o[u].q2 = INFINITY;
      if (x < count) {
        o[u].q = __ldcg(&cand[x].q);
        o[u].key = __ldcg(&cand[x].key);
        o[u].si = __ldcg(&cand[x].si);
        o[u].sj = __ldcg(&cand[x].sj);
        o[u].q2 = __ldcg(&cand[x].q2);
      }
    }
    cand_merge(c, o[0]);
    cand_merge(c, o[1]);
  }
  c = warp_cand_reduce(c);
  __syncthreads();
  if (lane == 0) sCand[warp] = c;
  __syncthreads();
  Cand e;
  e.q = INFINITY; e.key = ~0ull; e.si = -1; e.sj = -1; e.q2 = INFINITY;
  if (lane < (int)(blockDim.x >> 5)) e = sCand[lane];
  return warp_cand_reduce(e);  // identical in every thread
}

}  // namespace cas
