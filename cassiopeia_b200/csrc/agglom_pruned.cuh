// Building blocks of the exact pruned scan, shared by the single-GPU loop (agglom.cu) and the row-sharded
// loop (sharded.cu).  See the comment above scan_pruned_kernel in agglom.cu for the invariants.
#pragma once

#include "agglom_common.cuh"

namespace cas {

constexpr int PRUNE_CH = 8;  // chunks of 32 sub-segments a warp decides together

constexpr int SW = 128;  // columns per lower bound
constexpr int POOL = 512;  // pairs kept to warm-start the bound (>= CTAs of the pruned scan)

struct PruneArgs {
  unsigned *LB;               // [n][stride]
  int64_t stride;             // sub-segments per row, multiple of 32
  unsigned *Rsub;             // [stride] encoded float, NJ only
  const float *uf;            // [n] fl32(tinv * R_j), NJ only
  unsigned long long *gbest;  // encoded double
  unsigned long long *scanned;  // statistics: map elements actually read
  unsigned long long *cta_read; // [gridDim] per-CTA element counts of this launch (summed by the last CTA)
};

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
  double cap;  // gbest + margin: exact evaluation (and candidate recording) stops above it.  The margin is
               // the near-tie window or 1 % of |gbest|, whichever is larger: pairs above the window cannot
               // win, tie or be a near-tie runner-up; the extra 1 % keeps the per-CTA winners that warm-start
               // the next iteration's bound plentiful at negligible cost
  float thr;
  // near-tie window of the diagnostics (1e-6 relative, absolute below 1) around the value x
  static __device__ __forceinline__ double window(double x) { return 1.0000001e-6 * fmax(1.0, fabs(x)); }
  // the lane's own best is usually far looser than the global bound (a lane sees few elements per launch)
  __device__ __forceinline__ float threshold() const {
    const double best = bi >= 0 ? fmin(bq, cap) : cap;
    if (best == INFINITY) return INFINITY;
    if (!NJ) return __double2float_ru(best);
    const double B = best + ui;
    return __double2float_ru(B + 9.5367431640625e-07 * (umax + fabs(B)) + 1e-300);
  }
  __device__ __forceinline__ void set_cap(double gb) {
    const double c = gb + fmax(window(gb), 0.01 * fabs(gb)) + 1e-9 * fabs(gb);
    if (c < cap) { cap = c; thr = threshold(); }
  }
  __device__ __forceinline__ void set_row(int row, double ri) {
    i = row; Ri = ri;
    ui = NJ ? __dmul_rn(tinv, Ri) : 0.0;
    thr = threshold();
  }
  // one 128-bit vector of row i starting at column j; columns >= jend are not part of the triangle.
  // Returns the minimum of the float32 (rounded-down) dissimilarities of the valid columns.
  // GUARD = false: every column of the vector is known to be < jend (interior sub-segments).
  template <bool GUARD = true>
  __device__ __forceinline__ float vec(const T (&v)[V], const float (&su)[V], int j, int jend) {
    float sv[V], qf[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      sv[e] = (!GUARD || j + e < jend) ? screen_value(v[e]) : INFINITY;
      qf[e] = NJ ? __fsub_rn(sv[e], su[e]) : sv[e];
    }
    float m = qf[0], dm = sv[0];
#pragma unroll
    for (int e = 1; e < V; ++e) { m = fminf(m, qf[e]); dm = fminf(dm, sv[e]); }
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
    return dm;
  }
};

// Scan state of one warp that persists over the rows it handles in a launch.
struct PruneWarp {
  double gb_seen;             // latest value of gbest this warp has read or written
  unsigned long long nread;   // map elements read
};

// One row of the pruned scan, handled by one warp: decide from the row's lower bounds which sub-segments
// can hold the minimum (or a near-tie), read exactly those, refresh their bounds, publish improvements of
// the global bound.  `row` / `lbrow` point at the row's map elements / bounds, `i` is its slot (columns
// [0, i) belong to the lower triangle), `list` is the warp's shared-memory scratch (CH * 32 entries).
template <typename T, bool NJ>
__device__ __forceinline__ void prune_row(RowScan<T, NJ> &rs, PruneWarp &pw, const T *__restrict__ row,
                                          unsigned *lbrow, int i, double tinv_, const PruneArgs &p,
                                          unsigned short *list, int lane) {
  constexpr int V = VecOf<T>::N;
  constexpr int LOADS = SW / (32 * V);  // 128-bit loads per lane and sub-segment (1 float, 2 double)
  constexpr int G = 8 / LOADS;          // sub-segments in flight: eight 128-bit loads per lane
  constexpr int CH = PRUNE_CH;          // chunks of 32 sub-segments decided together (32768 columns)
  const int nsub = (i + SW - 1) / SW;   // sub-segments holding a column < i; the last one ends at the diagonal
  double &gb_seen = pw.gb_seen;
  bool scanned_any = false;
  for (int sb = 0; sb < nsub; sb += CH * 32) {
    // ---- decide: lane l looks at sub-segments sb + 32 c + l, c < CH (coalesced 128-byte loads).
    // Float arithmetic; every rounding is covered by the relative slack, which only makes the test
    // scan a little more:  skip  <=>  lb - us > (gbest + window) + u_i + slack.
    const double gb = dec_f64(__ldcg(p.gbest));
    gb_seen = gb;
    rs.set_cap(gb);
    const float gbw = __double2float_ru(gb + RowScan<T, NJ>::window(gb));
    const float uif = NJ ? __double2float_ru(rs.ui) : 0.f;
    const float tinvf = (float)tinv_;
    const float base = gbw + uif;
    const float mag0 = fabsf(gbw) + fabsf(uif);
    unsigned lbv[CH], rsv[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int sgm = sb + c * 32 + lane;
      lbv[c] = sgm < nsub ? lbrow[sgm] : enc_f32(INFINITY);
      rsv[c] = (NJ && sgm < nsub) ? p.Rsub[sgm] : enc_f32(0.f);
    }
    unsigned flags = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (sb + c * 32 >= nsub) break;  // warp-uniform: no sub-segment of this chunk exists
      const float lb = dec_f32(lbv[c]);
      const float us = NJ ? tinvf * dec_f32(rsv[c]) : 0.f;
      const float mag = mag0 + fabsf(lb) + fabsf(us);
      // false while gbest = +inf or lb = -inf (unknown): then the sub-segment is read
      const bool skip = (lb - us) > base + 1e-6f * mag;
      const bool nd = (sb + c * 32 + lane < nsub) && !skip;
      flags |= (nd ? 1u : 0u) << c;
    }
    // the diagonal sub-segment needs per-column guards: it is handled apart from the interior ones
    bool diag_needed = false;
    {
      const int dsub = nsub - 1 - sb;  // position of the diagonal sub-segment in this round, if any
      if (dsub >= 0 && dsub < CH * 32) {
        const unsigned bit = (lane == (dsub & 31)) ? (1u << (dsub >> 5)) : 0u;
        diag_needed = __any_sync(0xffffffffu, (flags & bit) != 0u);
        flags &= ~bit;
      }
    }
    // ---- compact the needed interior sub-segments into the warp's list
    const int cnt = __popc(flags);
    int total = 0;
    if (__any_sync(0xffffffffu, cnt != 0)) {
      int pre = cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, pre, off);
        if (lane >= off) pre += o;
      }
      total = __shfl_sync(0xffffffffu, pre, 31);
      int wpos = pre - cnt;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if ((flags >> c) & 1u) list[wpos++] = (unsigned short)(c * 32 + lane);
      __syncwarp();
    }
    if (total == 0 && !diag_needed) continue;
    // ---- read them, G at a time (all columns valid: no guards)
    for (int base_i = 0; base_i < total; base_i += G) {
      int sg[G];
      T v[G][LOADS][V];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        sg[g] = (base_i + g < total) ? sb + (int)list[base_i + g] : -1;
        if (sg[g] >= 0) {
#pragma unroll
          for (int l = 0; l < LOADS; ++l) ld_stream<false>(row + sg[g] * SW + l * 32 * V + lane * V, v[g][l]);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (sg[g] < 0) continue;  // warp-uniform
        float dm = INFINITY;
#pragma unroll
        for (int l = 0; l < LOADS; ++l) {
          const int j = sg[g] * SW + l * 32 * V + lane * V;
          float su[V];
#pragma unroll
          for (int e = 0; e < V; ++e) su[e] = 0.f;
          if (NJ) {
            if (V == 4) *reinterpret_cast<float4 *>(su) = __ldg(reinterpret_cast<const float4 *>(p.uf + j));
            else *reinterpret_cast<float2 *>(su) = __ldg(reinterpret_cast<const float2 *>(p.uf + j));
          }
          dm = fminf(dm, rs.template vec<false>(v[g][l], su, j, i));
        }
        const unsigned m = __reduce_min_sync(0xffffffffu, enc_f32(dm));
        if (lane == 0) lbrow[sg[g]] = m;
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
      const unsigned m = __reduce_min_sync(0xffffffffu, enc_f32(dm));
      if (lane == 0) lbrow[sgd] = m;
      pw.nread += (unsigned long long)(i - sgd * SW);
    }
    scanned_any = true;
    __syncwarp();  // the list is rewritten by the next round
  }
  if (scanned_any) {
    double wb = rs.bq;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wb = fmin(wb, __shfl_xor_sync(0xffffffffu, wb, off));
    // publish only real improvements of the global bound (one address: keep the atomics rare)
    if (wb < gb_seen) {
      if (lane == 0) atomicMin(p.gbest, enc_f64(wb));
      gb_seen = wb;
      rs.set_cap(wb);
    }
  }
}

}  // namespace cas
