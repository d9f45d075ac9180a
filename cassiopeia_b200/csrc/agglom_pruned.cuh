// Building blocks of the exact pruned scan, shared by the single-GPU loop (agglom.cu) and the row-sharded
// loop (sharded.cu).
//
// Bounds.  For NJ the criterion of a pair is q = d_ij - u_i - u_j with u = tinv * R (UPGMA: q = d, u = 0).
// For every row i and every 128-column sub-segment s of its lower-triangle part the scan keeps
//     E[i][s]  <=  min_j ( d_ij - uf_j ) + C          (uf_j = fl32(u_j), j live, j < i, j in s)
// where C is the cumulative UPWARD drift of the uf_j: every merge measures the largest increase of any
// column's uf and adds it to C, so a bound written at drift C0 as (min + C0) stays valid for ever as
// (E - C): the columns' uf can only have grown by C - C0 since then.  The two columns a merge rewrites (the
// new node in slot lo, the node that moved into slot hi) are excluded from the drift and folded into the
// bounds with their exact values by the next iteration's row-test kernel; the two rewritten rows are marked
// unknown (-inf) and rescanned once.  beta[i] = min_s E[i][s] gives a one-load test per row.
//
// A pair can only win, tie or be a near-tie runner-up if q <= gbest + window, where gbest is the exact
// criterion of some real pair (warm-started by the previous merge from a persistent pool of good pairs,
// lowered by every warp that finds a better one).  Rows / sub-segments whose bound says "impossible" are
// never read; everything else goes through the same float32 screen + exact float64 evaluation +
// (value, id_lo, id_hi) rule as the dense scan.  Hence the selected pair and the near-tie diagnostics are
// IDENTICAL to the dense scan's, whatever the timing of the gbest updates.
#pragma once

#include "agglom_common.cuh"

namespace cas {

constexpr int SW = 128;        // columns per bound (sub-segment)
constexpr int PRUNE_UNIT = 8;  // sub-segments per work unit (1024 columns): one warp, one bound test (lanes
                               // 0..7), at most one group of eight 128-bit loads per lane -- the critical
                               // path of an iteration is one unit, so units are kept short
constexpr int PRUNE_WIDE = 32; // sub-segments per work unit of a row that is known to be almost all skipped:
                               // one 128-byte load of bounds per step instead of four 32-byte ones
constexpr int POOL = 512;      // pairs kept to warm-start gbest (>= CTAs of the pruned scan)

struct PruneArgs {
  unsigned *E;                // [rows][stride] encoded float bounds (see above)
  int64_t stride;             // sub-segments per row, multiple of 32
  unsigned *beta;             // [rows] encoded float, min over the row's E
  const float *uf;            // [n] fl32(tinv * R_j) by slot (NJ; unused for UPGMA)
  float *hdr;                 // [0] = C as of the previous scan
  unsigned *dmax;             // encoded float: largest uf increase measured by the last merge
  unsigned long long *gbest;  // encoded double
  int32_t *worklist;          // rows that failed this iteration's row test (row indices)
  unsigned *wcount;           // [0] rows queued from the front of the list, [1] rows queued from its back (split)
  int split;                  // 0: one list (front).  1: two lists -- rows scanned in short units (PRUNE_UNIT
                              // sub-segments: good for rows that are read in full, the critical path) are queued
                              // from the back (slot wlast downwards), rows scanned in PRUNE_WIDE-sub-segment
                              // units (one coalesced bound load per step: good for rows of which almost nothing
                              // is read) from the front
  int wide;                   // split only: 1 = only rows whose bound is unknown go to the back, 0 = all rows do
  int wlast;                  // split only: index of the last slot of the work list
  int32_t *sugg;              // [gridDim][2] per-CTA suggestion (slots) for the pool
  unsigned long long *scanned;   // statistics: map elements actually read
  unsigned long long *cta_read;  // [gridDim] per-CTA counts of this launch (summed by the last CTA)
  const unsigned long long *xw;  // exact NJ mode: running maxima behind exact_window() (agglom_common.cuh); else null
  const unsigned long long *rmaxb;  // NJ: bits of the running max |row sum| (umax = tinv * that bounds every |u_j|)
  const int *strict_epoch;       // exact NJ mode: iteration whose row sums are the reference's sequential sums (the
                                 // strict tail, agglom_exact.cuh), or null
  int t;                         // iteration index (only read with strict_epoch)
  const void *newcol;            // [2][newcol_stride] map elements: the two columns the previous merge rewrote, by
  int64_t newcol_stride;         // slot -- d(v, lo) and d(v, hi) as the merge computed them.  The row test reads these
                                 // contiguous copies instead of two map elements per row (one per 2 MB page: the
                                 // scattered reads were most of the row test's time)
};

// Per-iteration scalars that the previous kernels left in global memory (drift, gbest, running maxima, the previous
// selection, work-list lengths).  Every warp of every CTA needs them; loading each word once per CTA (a different
// thread each, then shared memory) instead of once per warp keeps thousands of requests for the same few words from
// queueing up at their L2 slices, which the timeline (profiles/r2_timeline_*.txt) showed as ~0.5 us per dependent step.
struct IterRaw {
  unsigned long long rmax, eb, ab, dm, gbest;
  unsigned dmaxw, wc0, wc1;
  float hdr, uf_lo, uf_hi;
  int plo, phi;
  int strict;  // 1: this iteration's row sums ARE the reference's sequential sums (strict tail): no error window
};

__device__ __forceinline__ void iter_raw_load(IterRaw *s, const PruneArgs &p, const void *prev_sel, bool want_prev) {
  const int t = threadIdx.x;
  if (t == 0) s->rmax = p.rmaxb ? __ldcg(p.rmaxb) : 0ull;
  if (t == 32) s->eb = p.xw ? __ldcg(p.xw + XW_EBMAX) : 0ull;
  if (t == 64) s->ab = p.xw ? __ldcg(p.xw + XW_ABMAX) : 0ull;
  if (t == 96) s->dm = p.xw ? __ldcg(p.xw + XW_DABS) : 0ull;
  if (t == 128) s->gbest = __ldcg(p.gbest);
  if (t == 160) { s->dmaxw = __ldcg(p.dmax); s->hdr = __ldcg(p.hdr); }
  if (t == 192) {
    s->wc0 = __ldcg(p.wcount);
    s->wc1 = __ldcg(p.wcount + 1);
    s->strict = (p.strict_epoch && __ldcg(p.strict_epoch) == p.t) ? 1 : 0;
  }
  if (t == 224) {
    int plo = -1, phi = -1;
    if (want_prev && prev_sel) {
      const int *ps = static_cast<const int *>(prev_sel);  // Sel: lo, hi first
      plo = __ldcg(ps);
      phi = __ldcg(ps + 1);
    }
    s->plo = plo;
    s->phi = phi;
    s->uf_lo = (plo >= 0 && p.uf) ? __ldcg(p.uf + plo) : 0.f;
    s->uf_hi = (phi >= 0 && p.uf) ? __ldcg(p.uf + phi) : 0.f;
  }
  __syncthreads();
}
__device__ __forceinline__ float iter_cnow(const IterRaw &s) {
  const float d = fmaxf(dec_f32(s.dmaxw), 0.f);
  return __fadd_ru(s.hdr, __fmul_ru(d, 1.000001f));
}
__device__ __forceinline__ double iter_wfloor(const IterRaw &s, bool exact, int k, double tinv) {
  if (!exact || s.strict) return 0.0;
  const double rmax = __longlong_as_double((long long)s.rmax), eb = __longlong_as_double((long long)s.eb);
  const double ab = __longlong_as_double((long long)s.ab), dm = __longlong_as_double((long long)s.dm);
  const double delta = eb + (double)k * EXACT_U1 * ab;
  const double Delta = 2.0 * tinv * delta + 8.0 * EXACT_U1 * (dm + 2.0 * tinv * (rmax + delta));
  return 2.0 * (2.0 * Delta * 1.0001);
}

// largest |u_j| of any column (NJ), as a float rounded up; 0 for UPGMA
__device__ __forceinline__ float prune_umaxf(const PruneArgs &p, double tinv) {
  return p.rmaxb ? __double2float_ru(tinv * __longlong_as_double((long long)__ldcg(p.rmaxb))) : 0.f;
}

// floor of the near-tie window: twice the exact mode's error window (0 in the other modes)
__device__ __forceinline__ double prune_wfloor(const PruneArgs &p, int k, double tinv) {
  return p.xw ? 2.0 * exact_window(p.xw, k, tinv) : 0.0;
}

// C as of THIS scan: identical in every thread of every kernel of the iteration (two scalar loads)
__device__ __forceinline__ float prune_cnow(const PruneArgs &p) {
  // directed rounding: C must never under-estimate the true cumulative drift
  const float d = fmaxf(dec_f32(__ldcg(p.dmax)), 0.f);
  return __fadd_ru(p.hdr[0], __fmul_ru(d, 1.000001f));
}

// Bound of a sub-segment whose smallest fl32(d) - uf_j is dm, in the drifting frame, rounded down by more
// than the float operations can have rounded up.  A lane without a valid column carries dm = +inf: its
// inf - inf would be a NaN, and the encoding of a NaN is not ordered (the compiler builds enc_f32 from
// abs/neg, which may canonicalise the NaN to a pattern that decodes as -0.0 and wins the warp minimum --
// a valid but uselessly low bound that kept every row's diagonal sub-segment on the work list).
__device__ __forceinline__ unsigned prune_bound(float dm, float cnow) {
  return dm == INFINITY ? enc_f32(INFINITY) : enc_f32((dm + cnow) - 1e-6f * (fabsf(dm) + cnow));
}

template <typename T, bool NJ>
struct RowScan {
  static constexpr int V = VecOf<T>::N;
  const double *R;
  const int32_t *ids;
  double tinv, umax;
  double bq, b2;
  int bi, bj;
  int i;
  double Ri, ui;
  double cap;  // gbest + near-tie window: pairs above it cannot win, tie or be a near-tie runner-up, so the
               // (slow, serial) exact evaluation stops there
  float thr, uif;
  double wfloor;  // floor of the near-tie window (exact mode; else 0)
  // cheap "suggestion" for the pool: the lane's smallest float32-screened criterion over everything it read
  float sbest;
  int ssi, ssj;
  __device__ __forceinline__ float threshold() const {
    const double best = bi >= 0 ? fmin(bq, cap) : cap;
    if (best == INFINITY) return INFINITY;
    if (!NJ) return __double2float_ru(best);
    const double B = best + ui;
    return __double2float_ru(B + 9.5367431640625e-07 * (umax + fabs(B)) + wfloor + 1e-300);
  }
  __device__ __forceinline__ void set_cap(double gb) {
    const double c = gb + near_window(gb, wfloor) + 1e-9 * fabs(gb);
    if (c < cap) { cap = c; thr = threshold(); }
  }
  __device__ __forceinline__ void set_row(int row, double ri) {
    i = row; Ri = ri;
    ui = NJ ? __dmul_rn(tinv, Ri) : 0.0;
    uif = (float)ui;
    thr = threshold();
  }
  // one 128-bit vector of row i starting at column j; columns >= jend are not part of the triangle
  // (GUARD = false: all columns are known to be valid).  Returns min (fl32(d) - uf_j) over the valid columns.
  template <bool GUARD>
  __device__ __forceinline__ float vec(const T (&v)[V], const float (&su)[V], int j, int jend) {
    float qf[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float sv = (!GUARD || j + e < jend) ? screen_value(v[e]) : INFINITY;
      qf[e] = NJ ? __fsub_rn(sv, su[e]) : sv;
    }
    float m = qf[0];
#pragma unroll
    for (int e = 1; e < V; ++e) m = fminf(m, qf[e]);
    if (m - uif < sbest) {
      sbest = m - uif;
      ssi = i;
#pragma unroll
      for (int e = V - 1; e >= 0; --e)
        if (qf[e] == m) ssj = j + e;
    }
    if (m <= thr) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        if (qf[e] <= thr && (!GUARD || j + e < jend)) {
          double q = (double)v[e];
          if (NJ) q = __dsub_rn(q, __dmul_rn(tinv, __dadd_rn(Ri, R[j + e])));
          if (q <= bq || bi < 0) {
            tie_path<false>(q, i, j + e, bq, bi, bj, b2, ids);
            thr = threshold();
          } else {
            b2 = fmin(b2, q);
          }
        }
      }
    }
    return m;
  }
};

// "a row / sub-segment with bound b (already minus C) and own u cannot hold a competitive pair":
// false while gbest = +inf or the bound is -inf (unknown)
// The slack covers every float32 rounding on either side of the comparison, including that of the column term
// itself (the bounds are on d - fl32(u_j), the criterion uses u_j: up to 2^-24 umax apart).
template <typename T, bool NJ>
__device__ __forceinline__ bool prune_impossible(float e, float cnow, float uif, double gb, double wfloor,
                                                 float umaxf) {
  const float gbw = __double2float_ru(gb + near_window(gb, wfloor));
  const float mag = fabsf(gbw) + fabsf(uif) + fabsf(e) + cnow + umaxf;
  return (e - cnow) > (gbw + uif) + 2e-6f * mag;
}

// Row-test kernel body (thread = row): fold the two columns the previous merge rewrote into the row's
// bounds, test the row, queue it if it may hold a competitive pair.  `row` = the row's map elements,
// prev_* = slots of the previous join (prev_lo < 0: none yet; prev_moved: a node moved into slot prev_hi).
template <typename T, bool NJ>
__device__ __forceinline__ void prune_rowtest(const PruneArgs &p, const IterRaw &sc, bool valid, int64_t ridx, int i,
                                              double Ri, double tinv, int k_prev_last, double wfloor, unsigned beta0,
                                              T nc_lo, T nc_hi) {
  const float cnow = iter_cnow(sc);
  const int prev_lo = sc.plo, prev_hi = sc.phi;
  const bool prev_moved = prev_hi != k_prev_last;  // a node moved into slot hi unless hi was the last live slot
  bool unsafe = false, unknown = false;
  if (valid) {
    unsigned b = beta0;
    unknown = b == enc_f32(-INFINITY);  // the whole row will be read
    if (prev_lo >= 0 && i != prev_lo && i != prev_hi) {
      unsigned *erow = p.E + ridx * p.stride;
      // both loads first: one round trip for the two read-modify-writes
      const bool fold_lo = i > prev_lo, fold_hi = prev_moved && i > prev_hi;
      const unsigned e_lo_old = fold_lo ? erow[prev_lo / SW] : 0u, e_hi_old = fold_hi ? erow[prev_hi / SW] : 0u;
      const float x_lo = fold_lo ? screen_value(nc_lo) : 0.f;
      const float x_hi = fold_hi ? screen_value(nc_hi) : 0.f;
      if (fold_lo) {
        const float q = x_lo - (NJ ? sc.uf_lo : 0.f);
        const unsigned e = enc_f32((q + cnow) - 1e-6f * (fabsf(x_lo) + fabsf(q) + cnow));
        unsigned cur = min(e_lo_old, e);
        erow[prev_lo / SW] = cur;
        b = min(b, e);
        if (fold_hi && prev_hi / SW == prev_lo / SW) {
          const float q2 = x_hi - (NJ ? sc.uf_hi : 0.f);
          const unsigned e2 = enc_f32((q2 + cnow) - 1e-6f * (fabsf(x_hi) + fabsf(q2) + cnow));
          erow[prev_hi / SW] = min(cur, e2);
          b = min(b, e2);
        }
      }
      if (fold_hi && !(fold_lo && prev_hi / SW == prev_lo / SW)) {
        const float q = x_hi - (NJ ? sc.uf_hi : 0.f);
        const unsigned e = enc_f32((q + cnow) - 1e-6f * (fabsf(x_hi) + fabsf(q) + cnow));
        erow[prev_hi / SW] = min(e_hi_old, e);
        b = min(b, e);
      }
    }
    const double gb = dec_f64(sc.gbest);
    const float umaxf = NJ ? __double2float_ru(tinv * __longlong_as_double((long long)sc.rmax)) : 0.f;
    unsafe = !prune_impossible<T, NJ>(dec_f32(b), cnow, NJ ? (float)(tinv * Ri) : 0.f, gb, wfloor, umaxf);
    // an unsafe row's beta is rebuilt by its work units (atomicMin)
    p.beta[ridx] = unsafe ? enc_f32(INFINITY) : b;
  }
  const int lane = threadIdx.x & 31;
  const bool back = unsafe && p.split && (!p.wide || unknown);
  const bool front = unsafe && !back;
  const unsigned m = __ballot_sync(0xffffffffu, front);
  if (m) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(p.wcount, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (front) p.worklist[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)ridx;
  }
  const unsigned mb = __ballot_sync(0xffffffffu, back);
  if (mb) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(p.wcount + 1, (unsigned)__popc(mb));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (back) p.worklist[p.wlast - (int)(base + __popc(mb & ((1u << lane) - 1u)))] = (int32_t)ridx;
  }
}

// Scan state of one warp that persists over the work units it handles in a launch.
struct PruneWarp {
  double gb_seen;             // latest value of gbest this warp has read or written
  unsigned long long nread;   // map elements read
};

// One work unit = sub-segments [sb, sb + unit) of one row (unit <= 32), handled by one warp: decide from the bounds
// which sub-segments can hold a competitive pair, read exactly those (eight 128-bit loads in flight per
// lane), refresh their bounds, lower the row bound, publish improvements of gbest.  `row` / `erow` point at
// the row's map elements / bounds, `i` is its slot (columns [0, i) are the lower triangle), `list` is the
// warp's shared-memory scratch (`unit` entries).
template <typename T, bool NJ>
__device__ __forceinline__ void prune_unit(RowScan<T, NJ> &rs, PruneWarp &pw, const T *__restrict__ row,
                                           unsigned *erow, int64_t ridx, int i, int sb, float cnow,
                                           const PruneArgs &p, unsigned short *list, int lane,
                                           int unit = PRUNE_UNIT) {
  constexpr int V = VecOf<T>::N;
  constexpr int LOADS = SW / (32 * V);  // 128-bit loads per lane and sub-segment (1 float, 2 double)
  constexpr int G = 8 / LOADS;          // sub-segments in flight
  const int nsub = (i + SW - 1) / SW;   // sub-segments holding a column < i; the last one ends at the diagonal
  // ---- decide: lane l looks at sub-segments sb + 32 c + l (coalesced 128-byte loads of the bounds)
  const double gb = dec_f64(__ldcg(p.gbest));
  pw.gb_seen = gb;
  rs.set_cap(gb);
  unsigned kept = enc_f32(INFINITY);  // lane: the bound it looked at, if it skipped the sub-segment
  bool need = false;
  {
    const int sgm = sb + lane;
    const bool exists = lane < unit && sgm < nsub;
    const unsigned e = exists ? erow[sgm] : enc_f32(INFINITY);
    const bool skip = prune_impossible<T, NJ>(dec_f32(e), cnow, rs.uif, gb, rs.wfloor, (float)rs.umax * 1.0000002f);
    need = exists && !skip;
    if (exists && skip) kept = e;
  }
  // the diagonal sub-segment needs per-column guards: it is handled apart from the interior ones
  unsigned flags = __ballot_sync(0xffffffffu, need);
  bool diag_needed = false;
  {
    const int dsub = nsub - 1 - sb;
    if (dsub >= 0 && dsub < unit) {
      diag_needed = (flags >> dsub) & 1u;
      flags &= ~(1u << dsub);
    }
  }
  // ---- the needed interior sub-segments, in lane order
  const int total = __popc(flags);
  if (total > 0) {
    if (need && ((flags >> lane) & 1u)) list[__popc(flags & ((1u << lane) - 1u))] = (unsigned short)lane;
    __syncwarp();
  }
  unsigned fresh = enc_f32(INFINITY);  // warp: smallest refreshed bound
  // ---- read them, G at a time (all columns valid: no guards)
  for (int base_i = 0; base_i < total; base_i += G) {
    int sg[G];
    T v[G][LOADS][V];
    float su[G][LOADS][V];  // fl32(u_j) of the same columns: loaded in the same wave as the map elements (issued one
                            // by one between the evaluations they were eight dependent L2 round trips per unit)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sg[g] = (base_i + g < total) ? sb + (int)list[base_i + g] : -1;
#pragma unroll
      for (int l = 0; l < LOADS; ++l) {
#pragma unroll
        for (int e = 0; e < V; ++e) su[g][l][e] = 0.f;
      }
      if (sg[g] >= 0) {
#pragma unroll
        for (int l = 0; l < LOADS; ++l) {
          const int j = sg[g] * SW + l * 32 * V + lane * V;
          ld_stream<false>(row + j, v[g][l]);
          if (NJ) {
            if (V == 4) *reinterpret_cast<float4 *>(su[g][l]) = __ldg(reinterpret_cast<const float4 *>(p.uf + j));
            else *reinterpret_cast<float2 *>(su[g][l]) = __ldg(reinterpret_cast<const float2 *>(p.uf + j));
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (sg[g] < 0) continue;  // warp-uniform
      float dm = INFINITY;
#pragma unroll
      for (int l = 0; l < LOADS; ++l) {
        const int j = sg[g] * SW + l * 32 * V + lane * V;
        dm = fminf(dm, rs.template vec<false>(v[g][l], su[g][l], j, i));
      }
      const unsigned m = __reduce_min_sync(0xffffffffu, prune_bound(dm, cnow));
      if (lane == 0) erow[sg[g]] = m;
      fresh = min(fresh, m);
    }
  }
  pw.nread += (unsigned long long)total * SW;
  if (diag_needed) {
    const int sgd = nsub - 1;
    float dm = INFINITY;
#pragma unroll
    for (int l = 0; l < LOADS; ++l) {
      const int j = sgd * SW + l * 32 * V + lane * V;
      T v[V];
      float su[V];
#pragma unroll
      for (int e = 0; e < V; ++e) { v[e] = (T)INFINITY; su[e] = 0.f; }
      if (j < i) {
        ld_stream<false>(row + j, v);
        if (NJ) {
          if (V == 4) *reinterpret_cast<float4 *>(su) = __ldg(reinterpret_cast<const float4 *>(p.uf + j));
          else *reinterpret_cast<float2 *>(su) = __ldg(reinterpret_cast<const float2 *>(p.uf + j));
        }
      }
      dm = fminf(dm, rs.template vec<true>(v, su, j, i));
    }
    const unsigned m = __reduce_min_sync(0xffffffffu, prune_bound(dm, cnow));
    if (lane == 0) erow[sgd] = m;
    fresh = min(fresh, m);
    pw.nread += (unsigned long long)(i - sgd * SW);
  }
  // the row bound: every sub-segment of the unit was either kept or refreshed
  {
    const unsigned bm = min(__reduce_min_sync(0xffffffffu, kept), fresh);
    if (lane == 0) atomicMin(p.beta + ridx, bm);
  }
  if (total > 0 || diag_needed) {
    double wb = rs.bq;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wb = fmin(wb, __shfl_xor_sync(0xffffffffu, wb, off));
    // publish only real improvements of the global bound (one address: keep the atomics rare)
    if (wb < pw.gb_seen) {
      if (lane == 0) atomicMin(p.gbest, enc_f64(wb));
      pw.gb_seen = wb;
      rs.set_cap(wb);
    }
    __syncwarp();  // the list is rewritten by the warp's next unit
  }
}

// Per-CTA suggestion = the smallest float-screened value any lane saw, with its pair.  All threads call it.
template <typename T, bool NJ>
__device__ __forceinline__ void prune_publish_suggestion(const RowScan<T, NJ> &rs, const PruneArgs &p) {
  __shared__ float sSv[SCAN_WARPS];
  __shared__ int sSi[SCAN_WARPS], sSj[SCAN_WARPS];
  float v = rs.sbest;
  int si = rs.ssi, sj = rs.ssj;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, si, off), oj = __shfl_xor_sync(0xffffffffu, sj, off);
    if (ov < v) { v = ov; si = oi; sj = oj; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sSv[warp] = v; sSi[warp] = si; sSj[warp] = sj; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < SCAN_WARPS; ++w)
      if (sSv[w] < v) { v = sSv[w]; si = sSi[w]; sj = sSj[w]; }
    p.sugg[2 * blockIdx.x] = v < INFINITY ? si : -1;
    p.sugg[2 * blockIdx.x + 1] = v < INFINITY ? sj : -1;
  }
}

// Warm start of gbest for the next iteration (called by every thread of the merge's last CTA, after all
// other CTAs' writes are fenced): entry x of the persistent pool keeps the best of its old pair, CTA x's exact
// winner and CTA x's float32 suggestion of the scan that just ended, all re-evaluated exactly under the NEW
// row sums; pairs touching the joined slots are dropped, the moved slot is renamed.  The smallest value is
// an upper bound of the next minimum.  `addr(si, sj)` returns the address of d(si, sj) (si > sj) or null when
// this rank cannot evaluate the pair.  The loads are issued in three independent waves (indices, values,
// nothing dependent in between) so that the tail of the merge costs two memory round trips, not a dozen.
// `rot` rotates which CTA feeds which entry from iteration to iteration: in steady state only the first few
// CTAs of the scan have any work unit, and with a fixed mapping only their entries would ever be refreshed
// while every join kills entries -- the pool would shrink to a handful of pairs and the bound would go loose.
// Called by whole warps: NTHREADS threads with indices tid = 0 .. NTHREADS - 1 (a multiple of 32).
template <typename T, bool NJ, int NTHREADS, typename Addr>
__device__ __forceinline__ void prune_warm_start(int32_t *pool, const Cand *cand, int ncand, const int32_t *sugg,
                                                 unsigned long long *gbest, const double *R, double tinv_next,
                                                 int lo, int hi, int last, int rot, Addr addr, int tid,
                                                 bool rename_moved = true) {
  constexpr int EPT = (POOL + NTHREADS - 1) / NTHREADS;
  int px[EPT][3], py[EPT][3];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int x = tid + e * NTHREADS;
#pragma unroll
    for (int c = 0; c < 3; ++c) { px[e][c] = -1; py[e][c] = -1; }
    if (x < POOL) {
      px[e][0] = pool[2 * x]; py[e][0] = pool[2 * x + 1];
      const int y = (x + rot) % POOL;  // the CTA whose winner / suggestion competes for this entry
      if (y < ncand) {
        px[e][1] = __ldcg(&cand[y].si); py[e][1] = __ldcg(&cand[y].sj);
        px[e][2] = __ldcg(sugg + 2 * y); py[e][2] = __ldcg(sugg + 2 * y + 1);
      }
    }
  }
  const T *pd[EPT][3];
  double dv[EPT][3], ri[EPT][3], rj[EPT][3];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int x = px[e][c], y = py[e][c];
      pd[e][c] = nullptr;
      dv[e][c] = INFINITY; ri[e][c] = 0.0; rj[e][c] = 0.0;
      if (x < 0 || y < 0 || x == lo || y == lo || x == hi || y == hi) continue;
      // the node of the last slot moved into slot hi (hi != last here): follow it -- or, when the caller has not yet
      // written the moved row and column (merge_kernel defers its map stores past the grid reduction), drop the pair
      if ((x == last || y == last) && !rename_moved) continue;
      if (x == last) x = hi;
      if (y == last) y = hi;
      if (x >= last || y >= last) continue;
      px[e][c] = max(x, y); py[e][c] = min(x, y);
      pd[e][c] = addr(px[e][c], py[e][c]);
      if (pd[e][c]) {
        dv[e][c] = (double)__ldcg(pd[e][c]);
        if (NJ) { ri[e][c] = __ldcg(R + px[e][c]); rj[e][c] = __ldcg(R + py[e][c]); }
      }
    }
  }
  double best = INFINITY;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int x = tid + e * NTHREADS;
    double qb = INFINITY;
    int bi = -1, bj = -1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (!pd[e][c]) continue;
      double q = dv[e][c];
      if (NJ) q = __dsub_rn(q, __dmul_rn(tinv_next, __dadd_rn(ri[e][c], rj[e][c])));
      if (q < qb) { qb = q; bi = px[e][c]; bj = py[e][c]; }
    }
    if (x < POOL) { pool[2 * x] = bi; pool[2 * x + 1] = bj; }
    best = fmin(best, qb);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, off));
  if ((threadIdx.x & 31) == 0 && best < INFINITY) atomicMin(gbest, enc_f64(best));
}

// After every CTA of a scan has finished: commit the drift (C := C as of this scan, dmax := 0).  One thread.
__device__ __forceinline__ void prune_commit(const PruneArgs &p) {
  p.hdr[0] = prune_cnow(p);
  *p.dmax = enc_f32(0.f);
}

}  // namespace cas
