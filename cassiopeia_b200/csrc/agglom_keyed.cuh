// UPGMA pruned scan with tie-aware ("keyed") bounds -- single-GPU loop (agglom.cu).
//
// UPGMA's criterion is the stored distance itself (UPGMASolver.py:118-170: first row-major arg-min of the map =
// lexicographic minimum of (d, id_lo, id_hi)).  Lineage data is tie-heavy: duplicated cells are at distance 0 from
// each other and unweighted Hamming distances take few distinct values, so at most iterations thousands of pairs
// TIE with the minimum.  Bounds on the value alone (agglom_pruned.cuh) cannot exclude a tied sub-segment, and the
// value-only pruned scan re-read every tied pair every iteration (slower than the dense scan on such inputs).
//
// Here every bound is a lexicographic pair.  For row i and 128-column sub-segment s
//     (Ev[i][s], key(id_i, Ep[i][s]))  <=  min over live columns j < i in s of (d_ij, key(id_i, id_j))
// where Ev is the EXACT minimum value in the map's own precision (no arithmetic is involved, so no slack) and Ep the
// node id of the partner that attained it when the sub-segment was last read.  Distances between two surviving
// nodes never change and node ids are never re-used, so a stored pair can only become stale by DISAPPEARING (its
// partner was merged away): the bound then under-estimates the true minimum -- still valid, and repaired the next
// time the sub-segment is read.  The two columns a merge rewrites (new node in slot lo, moved node in slot hi) are
// folded in with their exact (value, id) by the next row test, as in the value-only scheme; the two rewritten rows
// are marked unknown and read once.  (betaV[i], betaP[i]) is the row-level minimum of the row's sub-segment bounds.
//
// A row / sub-segment can hold the winner only if its bound is <= (gv0, gk0), the (value, key) of a REAL live pair
// chosen by the previous merge from a persistent pool (warm start), or if its value is <= the running value bound
// gbest that the scan lowers as it finds better pairs.  Everything else is skipped without touching the map.  The
// selected pair is therefore the dense scan's pair, bit for bit; the tie / near-tie COUNTERS are computed from the
// pairs actually read and may be lower than the dense scan's (they are diagnostics, not part of the result).
//
// Row-level bounds of rows that were read cannot be rebuilt with atomics (a 64-bit value plus a 64-bit key), so a
// listed row is flagged DIRTY and the NEXT row-test kernel rebuilds its row-level bound from its sub-segment bounds
// (one warp per dirty row, extra CTAs of the same launch) before testing it.  The work list is double-buffered for
// that purpose.
#pragma once

#include "agglom_common.cuh"
#include "agglom_pruned.cuh"

namespace cas {

// betaP of a row that was listed at iteration t: its row-level bound must be rebuilt from Ev / Ep by the row test of
// iteration t + 1.  The mark carries the parity of t, so that a row the thread path lists NOW is not mistaken, by a
// dirty-path warp of the same launch that finds it in the previous list (the two rows the last merge rewrote always
// are), for a row listed at t - 1 -- it would be queued, and read, twice.
__device__ __forceinline__ int keyed_dirty_mark(int t) { return -2 - (t & 1); }
constexpr int KEYED_NOPID = -1;   // Ep / betaP: no partner known (key = 0, the smallest)

template <typename T> struct KeyedEnc;
template <> struct KeyedEnc<float> {
  using BT = unsigned;
  static __device__ __forceinline__ BT enc(double v) { return enc_f32((float)v); }  // v is a float value: exact
  static __device__ __forceinline__ double dec(BT e) { return (double)dec_f32(e); }
};
template <> struct KeyedEnc<double> {
  using BT = unsigned long long;
  static __device__ __forceinline__ BT enc(double v) { return enc_f64(v); }
  static __device__ __forceinline__ double dec(BT e) { return dec_f64(e); }
};

struct KeyedArgs {
  void *Ev;                // [rows][stride] encoded exact sub-segment minima (KeyedEnc<T>::BT)
  int32_t *Ep;             // [rows][stride] partner node id of that minimum (KEYED_NOPID: unknown)
  int64_t stride;
  void *betaV;             // [rows] encoded row-level minimum value
  int32_t *betaP;          // [rows] partner id / KEYED_NOPID / keyed_dirty_mark(t)
  int32_t *list_new;       // work list this iteration's row test builds and its scan consumes
  unsigned *cnt_new;       // [0] rows queued from the front (known bounds: wide units), [1] from the back (unknown)
  const int32_t *list_old; // the previous iteration's list = the dirty rows
  unsigned *cnt_old;       // (reset by this iteration's scan tail: the next row test builds into it)
  int wlast;               // index of the last slot of a list
  unsigned long long *gbest;   // encoded double: running bound on the minimum VALUE (a real pair's value)
  unsigned long long *gstart;  // [0] encoded double gv0, [1] key gk0: the warm-start pair, fixed during a scan
  const int32_t *ids;
  const void *newcol; int64_t newcol_stride;  // contiguous copies of the two columns the previous merge rewrote
  unsigned long long *scanned; unsigned long long *cta_read;
};

__device__ __forceinline__ unsigned long long keyed_key(int idi, int pid) {
  return pid < 0 ? 0ull : make_key(idi, pid);
}
__device__ __forceinline__ bool vk_less(double va, unsigned long long ka, double vb, unsigned long long kb) {
  return va < vb || (va == vb && ka < kb);
}
// "nothing at or above the bound (bv, bk) can be the winner": its value exceeds the running value bound, or it ties
// with the warm-start pair's value and loses on the key.  (-inf, *) = unknown is never skipped.
__device__ __forceinline__ bool keyed_skip(double bv, unsigned long long bk, double gb, double gv0,
                                           unsigned long long gk0) {
  return bv > gb || (bv == gv0 && bk > gk0);
}

// warp-wide lexicographic minimum of (v, k) with a payload; every lane returns the result
__device__ __forceinline__ void warp_vk_min(double &v, unsigned long long &k, int &payload) {
  double vm = v;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) vm = fmin(vm, __shfl_xor_sync(0xffffffffu, vm, off));
  unsigned long long km = (v == vm) ? k : ~0ull;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, km, off);
    km = o < km ? o : km;
  }
  const unsigned wb = __ballot_sync(0xffffffffu, v == vm && k == km);
  const int src = wb ? __ffs(wb) - 1 : 0;
  payload = __shfl_sync(0xffffffffu, payload, src);
  v = vm;
  k = wb ? km : ~0ull;
}

// Staged per-iteration scalars of the keyed kernels (one load each per CTA).
struct KeyedIter {
  double gv0; unsigned long long gk0;
  unsigned c_old0, c_old1, c_new0, c_new1;
  int plo, phi, id_moved;
};

__device__ __forceinline__ void keyed_iter_load(KeyedIter *s, const KeyedArgs &p, const void *prev_sel, bool want_prev) {
  const int t = threadIdx.x;
  if (t == 0) s->gv0 = dec_f64(__ldcg(p.gstart));
  if (t == 32) s->gk0 = __ldcg(p.gstart + 1);
  if (t == 64) { s->c_old0 = __ldcg(p.cnt_old); s->c_old1 = __ldcg(p.cnt_old + 1); }
  if (t == 96) { s->c_new0 = __ldcg(p.cnt_new); s->c_new1 = __ldcg(p.cnt_new + 1); }
  if (t == 128) {
    int plo = -1, phi = -1, idm = -1;
    if (want_prev && prev_sel) {
      const int *ps = static_cast<const int *>(prev_sel);  // Sel: lo, hi first
      plo = __ldcg(ps);
      phi = __ldcg(ps + 1);
      if (phi >= 0) idm = __ldcg(p.ids + phi);  // after the merge: the id of the node that moved into slot hi
    }
    s->plo = plo; s->phi = phi; s->id_moved = idm;
  }
  __syncthreads();
}

// Fold the candidate (cv, cpid) into the bound (ev, ep) of row with id idi: lexicographic minimum.
__device__ __forceinline__ bool keyed_fold(double &ev, int &ep, double cv, int cpid, int idi) {
  if (vk_less(cv, keyed_key(idi, cpid), ev, keyed_key(idi, ep))) { ev = cv; ep = cpid; return true; }
  return false;
}

// queue row `ridx` (warp-aggregated): front list when its bounds are known, back list when the whole row is read
__device__ __forceinline__ void keyed_queue(const KeyedArgs &p, bool front, bool back, int ridx) {
  const int lane = threadIdx.x & 31;
  const unsigned m = __ballot_sync(0xffffffffu, front);
  if (m) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(p.cnt_new, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (front) p.list_new[base + __popc(m & ((1u << lane) - 1u))] = ridx;
  }
  const unsigned mb = __ballot_sync(0xffffffffu, back);
  if (mb) {
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(p.cnt_new + 1, (unsigned)__popc(mb));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (back) p.list_new[p.wlast - (int)(base + __popc(mb & ((1u << lane) - 1u)))] = ridx;
  }
}

// Row test.  CTAs [0, nrow_ctas): thread = row (rows that are not dirty).  CTAs [nrow_ctas, gridDim): one warp per
// dirty row of the previous iteration's list: rebuild the row-level bound from the sub-segment bounds, then test.
template <typename T>
__global__ void __launch_bounds__(256) keyed_rowtest_kernel(KeyedArgs p, int k, const Sel *prev, int has_prev,
                                                            int prev_new_id, int nrow_ctas, int t) {
  using KE = KeyedEnc<T>;
  using BT = typename KE::BT;
  pdl_wait();
  pdl_launch_dependents();
  __shared__ KeyedIter sc;
  BT *Ev = static_cast<BT *>(p.Ev);
  BT *betaV = static_cast<BT *>(p.betaV);
  const T *newcol = static_cast<const T *>(p.newcol);
  const int lane = threadIdx.x & 31;
  const BT NINF = KE::enc(-INFINITY);
  const int mark_now = keyed_dirty_mark(t), mark_prev = keyed_dirty_mark(t - 1);
  if ((int)blockIdx.x < nrow_ctas) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i >= 1 && i < k;
    // the row's own loads are in flight while the scalars are staged
    const int idi = valid ? p.ids[i] : 0;
    BT bve = valid ? betaV[i] : NINF;
    int bp = valid ? p.betaP[i] : KEYED_NOPID;
    const T nc_lo = (valid && has_prev) ? newcol[i] : (T)0;
    const T nc_hi = (valid && has_prev) ? newcol[p.newcol_stride + i] : (T)0;
    keyed_iter_load(&sc, p, prev, has_prev != 0);
    const int plo = sc.plo, phi = sc.phi;
    const bool prev_moved = phi != k;  // k = the previous live size - 1 = the slot that was last: hi == last => no move
    bool front = false, back = false;
    if (valid && bp != mark_prev) {
      double bv = KE::dec(bve);
      const bool unknown = bve == NINF;
      bool changed = false;
      if (!unknown && plo >= 0 && i != plo && i != phi) {
        const bool fold_lo = i > plo, fold_hi = prev_moved && i > phi;
        BT *erow = Ev + (int64_t)i * p.stride;
        int32_t *prow = p.Ep + (int64_t)i * p.stride;
        const int s_lo = plo / SW, s_hi = phi / SW;
        // both loads first: one round trip for the two read-modify-writes
        const BT e_lo = fold_lo ? erow[s_lo] : NINF, e_hi = fold_hi ? erow[s_hi] : NINF;
        const int q_lo = fold_lo ? prow[s_lo] : KEYED_NOPID, q_hi = fold_hi ? prow[s_hi] : KEYED_NOPID;
        if (fold_lo) {
          double ev = KE::dec(e_lo);
          int ep = q_lo;
          bool ch = keyed_fold(ev, ep, (double)nc_lo, prev_new_id, idi);
          if (fold_hi && s_hi == s_lo) ch |= keyed_fold(ev, ep, (double)nc_hi, sc.id_moved, idi);
          if (ch) { erow[s_lo] = KE::enc(ev); prow[s_lo] = ep; }
          changed |= keyed_fold(bv, bp, (double)nc_lo, prev_new_id, idi);
        }
        if (fold_hi) {
          if (!(fold_lo && s_hi == s_lo)) {
            double ev = KE::dec(e_hi);
            int ep = q_hi;
            if (keyed_fold(ev, ep, (double)nc_hi, sc.id_moved, idi)) { erow[s_hi] = KE::enc(ev); prow[s_hi] = ep; }
          }
          changed |= keyed_fold(bv, bp, (double)nc_hi, sc.id_moved, idi);
        }
      }
      const bool competitive = unknown || !keyed_skip(bv, keyed_key(idi, bp), sc.gv0, sc.gv0, sc.gk0);
      if (competitive) {
        back = unknown;
        front = !unknown;
        p.betaP[i] = mark_now;  // rebuilt by the next row test, after this iteration's scan has refreshed the row
        if (changed && !unknown) betaV[i] = KE::enc(bv);
      } else if (changed) {
        betaV[i] = KE::enc(bv);
        p.betaP[i] = bp;
      }
    }
    keyed_queue(p, front, back, i);
    return;
  }
  // ---- dirty rows: one warp each
  keyed_iter_load(&sc, p, prev, has_prev != 0);
  const int plo = sc.plo, phi = sc.phi;
  const bool prev_moved = phi != k;
  const unsigned ndirty = sc.c_old0 + sc.c_old1;
  const int W = ((int)gridDim.x - nrow_ctas) * (int)(blockDim.x >> 5);
  const int gw = ((int)blockIdx.x - nrow_ctas) * (int)(blockDim.x >> 5) + (int)(threadIdx.x >> 5);
  for (unsigned idx = gw; idx < ndirty; idx += W) {
    const int r = idx < sc.c_old0 ? p.list_old[idx] : p.list_old[p.wlast - (int)(idx - sc.c_old0)];
    bool front = false, back = false;
    if (r >= 1 && r < k && p.betaP[r] == mark_prev) {  // warp-uniform
      const int idr = p.ids[r];
      const int nsub = (r + SW - 1) / SW;
      BT *erow = Ev + (int64_t)r * p.stride;
      int32_t *prow = p.Ep + (int64_t)r * p.stride;
      const bool foldable = plo >= 0 && r != plo && r != phi;
      const int s_lo = (foldable && r > plo) ? plo / SW : -1;
      const int s_hi = (foldable && prev_moved && r > phi) ? phi / SW : -1;
      const double c_lo = s_lo >= 0 ? (double)newcol[r] : 0.0;
      const double c_hi = s_hi >= 0 ? (double)newcol[p.newcol_stride + r] : 0.0;
      double bv = INFINITY;
      int bp = KEYED_NOPID;
      unsigned long long bk = ~0ull;
      bool any_unknown = false;
      for (int s0 = 0; s0 < nsub; s0 += 128) {  // four sub-segment bounds per lane in flight
        BT e[4];
        int q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int s = s0 + u * 32 + lane;
          e[u] = s < nsub ? erow[s] : KE::enc(INFINITY);
          q[u] = s < nsub ? prow[s] : KEYED_NOPID;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int s = s0 + u * 32 + lane;
          if (s >= nsub) continue;
          if (e[u] == NINF) { any_unknown = true; continue; }
          double ev = KE::dec(e[u]);
          int ep = q[u];
          bool ch = false;
          if (s == s_lo) ch |= keyed_fold(ev, ep, c_lo, prev_new_id, idr);
          if (s == s_hi) ch |= keyed_fold(ev, ep, c_hi, sc.id_moved, idr);
          if (ch) { erow[s] = KE::enc(ev); prow[s] = ep; }
          const unsigned long long ek = keyed_key(idr, ep);
          if (vk_less(ev, ek, bv, bk)) { bv = ev; bk = ek; bp = ep; }
        }
      }
      any_unknown = __any_sync(0xffffffffu, any_unknown);
      warp_vk_min(bv, bk, bp);
      const bool competitive = any_unknown || !keyed_skip(bv, bk, sc.gv0, sc.gv0, sc.gk0);
      if (lane == 0) {
        if (competitive) {
          back = any_unknown;
          front = !any_unknown;
          p.betaP[r] = mark_now;
        } else {
          betaV[r] = KE::enc(bv);
          p.betaP[r] = bp;
        }
      }
    }
    keyed_queue(p, front, back, r);
  }
}

// Lane state of the keyed scan: running best pair and runner-up value over everything the lane has read.
struct KeyedLane {
  double bq, b2;
  unsigned long long bk;
  int bi, bj;
};

// One element d(i, j) = v with column id idj: winner tracking + the sub-segment's lane minimum.
__device__ __forceinline__ void keyed_elem(KeyedLane &ls, double v, int i, int j, int idi, int idj, double &sv,
                                           unsigned long long &sk, int &spid) {
  const unsigned long long kk = make_key(idi, idj);
  if (vk_less(v, kk, sv, sk)) { sv = v; sk = kk; spid = idj; }
  if (vk_less(v, kk, ls.bq, ls.bk)) {
    ls.b2 = fmin(ls.b2, ls.bq);
    ls.bq = v; ls.bk = kk; ls.bi = i; ls.bj = j;
  } else if (v == v) {
    ls.b2 = fmin(ls.b2, v);
  }
}

// One work unit = sub-segments [sb, sb + unit) of row i (unit <= 32), one warp.
template <typename T>
__device__ __forceinline__ void keyed_unit(KeyedLane &ls, PruneWarp &pw, const T *__restrict__ row, int64_t ridx, int i,
                                           int idi, int sb, const KeyedArgs &p, const KeyedIter &sc,
                                           unsigned short *list, int lane, int unit) {
  using KE = KeyedEnc<T>;
  using BT = typename KE::BT;
  constexpr int V = VecOf<T>::N;
  constexpr int LOADS = SW / (32 * V);
  constexpr int G = 8 / LOADS;
  BT *erow = static_cast<BT *>(p.Ev) + ridx * p.stride;
  int32_t *prow = p.Ep + ridx * p.stride;
  const int nsub = (i + SW - 1) / SW;
  const double gb = dec_f64(__ldcg(p.gbest));
  pw.gb_seen = gb;
  bool need = false;
  {
    const int sgm = sb + lane;
    const bool exists = lane < unit && sgm < nsub;
    const BT e = exists ? erow[sgm] : KE::enc(INFINITY);
    const int q = exists ? prow[sgm] : KEYED_NOPID;
    const bool skip = e != KE::enc(-INFINITY) && keyed_skip(KE::dec(e), keyed_key(idi, q), gb, sc.gv0, sc.gk0);
    need = exists && !skip;
  }
  unsigned flags = __ballot_sync(0xffffffffu, need);
  bool diag_needed = false;
  {
    const int dsub = nsub - 1 - sb;
    if (dsub >= 0 && dsub < unit) {
      diag_needed = (flags >> dsub) & 1u;
      flags &= ~(1u << dsub);
    }
  }
  const int total = __popc(flags);
  if (total > 0) {
    if (need && ((flags >> lane) & 1u)) list[__popc(flags & ((1u << lane) - 1u))] = (unsigned short)lane;
    __syncwarp();
  }
  for (int base_i = 0; base_i < total; base_i += G) {
    int sg[G];
    T v[G][LOADS][V];
    int idv[G][LOADS][V];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sg[g] = (base_i + g < total) ? sb + (int)list[base_i + g] : -1;
      if (sg[g] >= 0) {
#pragma unroll
        for (int l = 0; l < LOADS; ++l) {
          const int j = sg[g] * SW + l * 32 * V + lane * V;
          ld_stream<false>(row + j, v[g][l]);
          if (V == 4) *reinterpret_cast<int4 *>(idv[g][l]) = __ldcg(reinterpret_cast<const int4 *>(p.ids + j));
          else *reinterpret_cast<int2 *>(idv[g][l]) = __ldcg(reinterpret_cast<const int2 *>(p.ids + j));
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      if (sg[g] < 0) continue;  // warp-uniform
      double sv = INFINITY;
      unsigned long long sk = ~0ull;
      int spid = KEYED_NOPID;
#pragma unroll
      for (int l = 0; l < LOADS; ++l) {
        const int j = sg[g] * SW + l * 32 * V + lane * V;
#pragma unroll
        for (int e = 0; e < V; ++e) keyed_elem(ls, (double)v[g][l][e], i, j + e, idi, idv[g][l][e], sv, sk, spid);
      }
      warp_vk_min(sv, sk, spid);
      if (lane == 0) { erow[sg[g]] = KE::enc(sv); prow[sg[g]] = sv < INFINITY ? spid : KEYED_NOPID; }
    }
  }
  pw.nread += (unsigned long long)total * SW;
  if (diag_needed) {
    const int sgd = nsub - 1;
    double sv = INFINITY;
    unsigned long long sk = ~0ull;
    int spid = KEYED_NOPID;
#pragma unroll
    for (int l = 0; l < LOADS; ++l) {
      const int j = sgd * SW + l * 32 * V + lane * V;
      if (j < i) {  // rows are padded to 32 elements and ids to n + 32: the vector loads stay in bounds
        T v[V];
        int idv[V];
        ld_stream<false>(row + j, v);
        if (V == 4) *reinterpret_cast<int4 *>(idv) = __ldcg(reinterpret_cast<const int4 *>(p.ids + j));
        else *reinterpret_cast<int2 *>(idv) = __ldcg(reinterpret_cast<const int2 *>(p.ids + j));
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (j + e < i) keyed_elem(ls, (double)v[e], i, j + e, idi, idv[e], sv, sk, spid);
      }
    }
    warp_vk_min(sv, sk, spid);
    if (lane == 0) { erow[sgd] = KE::enc(sv); prow[sgd] = sv < INFINITY ? spid : KEYED_NOPID; }
    pw.nread += (unsigned long long)(i - sgd * SW);
  }
  if (total > 0 || diag_needed) {
    double wb = ls.bq;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wb = fmin(wb, __shfl_xor_sync(0xffffffffu, wb, off));
    if (wb < pw.gb_seen) {
      if (lane == 0) atomicMin(p.gbest, enc_f64(wb));
      pw.gb_seen = wb;
    }
    __syncwarp();  // the list is rewritten by the warp's next unit
  }
}

// Warm start (all threads of the merge's last CTA, after every other CTA's stores are fenced): entry x of the pool
// keeps the lexicographically best of its old pair and CTA x's winner of the scan that just ended, re-read from the
// updated map; pairs touching the joined slots are dropped, the moved slot is renamed.  The best entry becomes
// (gv0, gk0) and the initial value bound of the next scan.
template <typename T, int NTHREADS>
__device__ __forceinline__ void keyed_warm_start(const T *D, int64_t ld, const int32_t *ids, int32_t *pool,
                                                 const Cand *cand, int ncand, unsigned long long *gbest,
                                                 unsigned long long *gstart, int lo, int hi, int last, int rot) {
  constexpr int EPT = (POOL + NTHREADS - 1) / NTHREADS;
  __shared__ double sV[NTHREADS / 32];
  __shared__ unsigned long long sK[NTHREADS / 32];
  const int tid = threadIdx.x;
  int px[EPT][2], py[EPT][2];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int x = tid + e * NTHREADS;
#pragma unroll
    for (int c = 0; c < 2; ++c) { px[e][c] = -1; py[e][c] = -1; }
    if (x < POOL) {
      px[e][0] = pool[2 * x]; py[e][0] = pool[2 * x + 1];
      const int y = (x + rot) % POOL;
      if (y < ncand) { px[e][1] = __ldcg(&cand[y].si); py[e][1] = __ldcg(&cand[y].sj); }
    }
  }
  double dv[EPT][2];
  int ia[EPT][2], ib[EPT][2];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      int x = px[e][c], y = py[e][c];
      dv[e][c] = INFINITY; ia[e][c] = 0; ib[e][c] = 0;
      px[e][c] = -1;
      if (x < 0 || y < 0 || x == lo || y == lo || x == hi || y == hi) continue;
      if (x == last) x = hi;
      if (y == last) y = hi;
      if (x >= last || y >= last || x == y) continue;
      const int a = max(x, y), b = min(x, y);
      px[e][c] = a; py[e][c] = b;
      dv[e][c] = (double)__ldcg(D + (int64_t)a * ld + b);
      ia[e][c] = __ldcg(ids + a); ib[e][c] = __ldcg(ids + b);
    }
  }
  double bv = INFINITY;
  unsigned long long bk = ~0ull;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int x = tid + e * NTHREADS;
    double qv = INFINITY;
    unsigned long long qk = ~0ull;
    int bi = -1, bj = -1;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (px[e][c] < 0) continue;
      const unsigned long long kk = make_key(ia[e][c], ib[e][c]);
      if (vk_less(dv[e][c], kk, qv, qk)) { qv = dv[e][c]; qk = kk; bi = px[e][c]; bj = py[e][c]; }
    }
    if (x < POOL) { pool[2 * x] = bi; pool[2 * x + 1] = bj; }
    if (vk_less(qv, qk, bv, bk)) { bv = qv; bk = qk; }
  }
  int dummy = 0;
  warp_vk_min(bv, bk, dummy);
  if ((tid & 31) == 0) { sV[tid >> 5] = bv; sK[tid >> 5] = bk; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < NTHREADS / 32; ++w)
      if (vk_less(sV[w], sK[w], bv, bk)) { bv = sV[w]; bk = sK[w]; }
    const unsigned long long e = enc_f64(bv);
    *gbest = e;
    gstart[0] = e;
    gstart[1] = bv < INFINITY ? bk : ~0ull;
  }
}

}  // namespace cas
