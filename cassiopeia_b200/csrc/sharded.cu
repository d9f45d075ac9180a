// Row-sharded Neighbor-Joining / UPGMA loop for maps larger than one GPU.
//
// No reference counterpart (the reference is single-host); same arithmetic as the single-GPU
// fast loop in agglom.cu, so 1-GPU and N-GPU runs produce the identical join list.
//
// Partition: slots are dealt to ranks in blocks of SHB = 64 rows, block-cyclically
// (owner(s) = (s / 64) % G); a rank stores its rows contiguously in local-row order
// (local(s) = (s / 64 / G) * 64 + s % 64), full length ld, so scans stay coalesced and the
// triangular work is balanced.  Row sums R, ids and cluster sizes are replicated.
//
// One iteration at live size k (host knows k, the pair lives on the device):
//   scan   : each rank scans the lower-triangle part of its own rows -> one candidate
//   C1     : all-gather of the G candidates (24 B each); every rank reduces them identically
//            (lexicographic (value, id_lo, id_hi): exact, order-free => root-free arg-min)
//   cols   : each rank extracts, for its rows v, d(v,lo), d(v,hi), d(v,last) -- by symmetry these
//            are the full rows lo, hi and last
//   C2     : all-gather of those three vectors (3 * k values)
//   merge  : every rank computes the merged row redundantly (bit-identical), updates the
//            replicated R / ids / sizes, the two columns of its own rows and, if it owns them,
//            rows lo and hi (swap compaction as in agglom.cu).
// Collectives: NCCL all-gather on the caller's stream (dlopen'ed libnccl.so.2, the one torch
// already loaded), or none at all in the single-process emulation used to test the kernels on
// one GPU (all virtual ranks share the gather buffers).
#include <dlfcn.h>
#include <stddef.h>
#include <string.h>

#include <vector>

#include "agglom_exact.cuh"

namespace cas {

bool prune_enabled_sharded(bool nj);  // agglom.cu
int prune_min_k();             // agglom.cu

constexpr int SHB = 64;  // rows per ownership block

__host__ __device__ __forceinline__ int sh_owner(int slot, int G) { return (slot / SHB) % G; }
__host__ __device__ __forceinline__ int sh_local(int slot, int G) {
  return (slot / SHB / G) * SHB + slot % SHB;
}
__host__ __device__ __forceinline__ int sh_slot(int local, int r, int G) {
  return ((local / SHB) * G + r) * SHB + local % SHB;
}
// number of local rows of rank r whose slot is < k
__host__ __device__ __forceinline__ int sh_count(int k, int r, int G) {
  const int nb = k / SHB, rem = k % SHB;
  int rows = (nb / G + (r < nb % G ? 1 : 0)) * SHB;
  if (rem > 0 && nb % G == r) rows += rem;
  return rows;
}
// stride (rows) of one rank's chunk in the gather buffers at live size k
__host__ __device__ __forceinline__ int sh_chunk(int k, int G) {
  const int nblocks = (k + SHB - 1) / SHB;
  return ((nblocks + G - 1) / G) * SHB;
}

// ---- direct NVLink exchange (P2P): every rank owns an exchange buffer that its peers map
// (cudaIpc) and write into with plain stores; iteration-tagged flags replace the collectives.
constexpr int SH_MAXG = 16;
constexpr int XT = 1024;  // exact mode: most slots whose sequential sums one ambiguous iteration may need (the first
                          // iteration that involves more than XR_SH_HEAVY switches the strict tail on)
struct ShExHeader {
  unsigned long long flag1[SH_MAXG];  // flag1[r] = tag: rank r's candidate for that iteration has landed
  unsigned long long flag2[SH_MAXG];  // flag2[r] = tag: rank r's rows (lo, hi, last) chunk has landed
  unsigned long long flagA[SH_MAXG];  // exact mode: rank r's list of marked slots has landed
  unsigned long long flagB[SH_MAXG];  // exact mode: rank r's sequential sums (of the marked slots it owns) have landed
  unsigned long long flagS[SH_MAXG];  // exact mode, strict tail: rank r's sequential sums of ALL its rows have landed
  Cand cand[SH_MAXG];
  int countA[SH_MAXG];                // exact mode: entries of listA[r]
  int error;                          // 1: a spin-wait timed out, 2: more than XT marked slots
  int pad[15];
  // exact mode, one ambiguous iteration at a time:
  int listA[SH_MAXG][XT];             // marked slots found by rank r among its rows
  double xsum[XT];                    // sequential row sum of the u-th slot of the sorted union (written by its owner)
  double xd[XT][XT];                  // xd[u][v] = d(slot_u, slot_v), written by the owner of slot_u
};
static_assert(sizeof(ShExHeader) % 256 == 0, "exchange header must keep the gather area aligned");
struct ShPeers {
  char *ex[SH_MAXG];  // exchange buffers of all ranks (own one included); ex[0] == nullptr => no P2P
};
__device__ __forceinline__ ShExHeader *ex_header(char *ex) { return reinterpret_cast<ShExHeader *>(ex); }
__device__ __forceinline__ char *ex_gath(char *ex) { return ex + sizeof(ShExHeader); }

// spin until *flag >= tag (written by a peer GPU); gives up after ~2 s and records the failure
// Once any wait has failed the flag stays set and every later wait returns at once: the pre-enqueued rest of the
// loop drains in microseconds instead of spinning two seconds per wait, and the host raises (PeerGroup.check).
__device__ __forceinline__ void wait_flag(const unsigned long long *flag, unsigned long long tag, int *error) {
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= tag) return;
    if (*(volatile int *)error != 0) return;
    if (clock64() - t0 > 4000000000LL) { atomicCAS(error, 0, 1); return; }
  }
}
__device__ __forceinline__ void post_flag(unsigned long long *flag, unsigned long long tag) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(tag) : "memory");
}

struct ShScanArgs {
  const void *Dloc;
  int64_t ld;
  int k, G, r;
  int nlrb;   // local row blocks of 32 rows
  int ncb, ntiles;
  const double *R;
  const int32_t *ids;
  double tinv;
  const unsigned long long *rmax_bits;
  Cand *cand;       // per-CTA scratch
  unsigned *counter;
  Cand *out;        // this rank's slot in the candidate gather buffer (NCCL / shared-buffer modes)
  ShPeers peers;    // P2P mode: publish the candidate to every rank
  unsigned long long tag;
};

constexpr int SH_TR = 32;  // rows per scan tile (divides SHB, so a tile's slots are contiguous)

template <typename T, bool NJ>
__global__ void __launch_bounds__(SCAN_THREADS, 4) sh_scan_kernel(ShScanArgs a) {
  __shared__ __align__(16) float sU[NJ ? TC : 4];
  __shared__ Cand sCand[SCAN_WARPS];
  __shared__ bool sLast;
  const T *__restrict__ D = static_cast<const T *>(a.Dloc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = a.k;
  const double umax = NJ ? a.tinv * __longlong_as_double(*a.rmax_bits) : 0.0;
  double bq = INFINITY, b2 = INFINITY;
  int bi = -1, bj = -1;
  int staged_cb = -1;

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    int cb = 0, rem = tile;
    for (;;) {
      // first local row block holding a row with slot > c0 (rows with slot <= c0 have no
      // element in this column block)
      int first = sh_count(cb * TC + 1, a.r, a.G) / SH_TR;
      int cnt = max(a.nlrb - first, 0);
      if (rem < cnt) { rem += first; break; }
      rem -= cnt;
      ++cb;
    }
    const int c0 = cb * TC;
    const int c1 = min(c0 + TC, k);
    if (NJ && cb != staged_cb) {
      __syncthreads();
      for (int x = threadIdx.x; x < TC; x += SCAN_THREADS)
        sU[x] = (c0 + x < c1) ? __double2float_rn(__dmul_rn(a.tinv, a.R[c0 + x])) : 0.f;
      staged_cb = cb;
      __syncthreads();
    }
    const int l0 = rem * SH_TR;
    for (int lr = l0 + warp; lr < l0 + SH_TR; lr += SCAN_WARPS) {
      const int i = sh_slot(lr, a.r, a.G);
      if (i >= k) continue;
      const int jend = min(c1, i);
      if (jend <= c0) continue;
      const T *__restrict__ row = D + (int64_t)lr * a.ld;
      const double Ri = NJ ? a.R[i] : 0.0;
      scan_row_segment<T, NJ>(row, i, c0, jend, lane, sU, a.R, Ri, a.tinv, umax, a.ids, bq, bi, bj, b2);
    }
  }

  Cand c = block_best(bq, bi, bj, b2, a.ids, sCand);
  if (threadIdx.x == 0) {
    a.cand[blockIdx.x] = c;
    __threadfence();
    unsigned prev = atomicInc(a.counter, gridDim.x - 1);
    sLast = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!sLast) return;
  __threadfence();
  c = reduce_candidates(a.cand, (int)gridDim.x, sCand);
  if (a.peers.ex[0] != nullptr) {
    if (threadIdx.x < a.G) {
      ShExHeader *h = ex_header(a.peers.ex[threadIdx.x]);
      h->cand[a.r] = c;
      __threadfence_system();
      post_flag(&h->flag1[a.r], a.tag);
    }
  } else if (threadIdx.x == 0) {
    *a.out = c;
  }
}

// Pruned scan of the local rows (agglom_pruned.cuh): bounds, gbest and the pool are per rank and never cross
// the link -- a rank's bound is the value of one of its own real pairs, an upper bound of the global minimum
// as well.  Row indices in the work list, E and beta are LOCAL rows.
template <typename T, bool NJ>
__global__ void __launch_bounds__(256) sh_rowtest_kernel(PruneArgs p, const T *__restrict__ Dloc, int64_t ld, int k,
                                                         int G, int r, const double *__restrict__ R, double tinv,
                                                         const Sel *prev, int has_prev) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ IterRaw sc;
  const int lr = blockIdx.x * blockDim.x + threadIdx.x;
  const int nloc = sh_count(k, r, G);
  const int i = lr < nloc ? sh_slot(lr, r, G) : 0;
  const bool valid = lr < nloc && i >= 1;
  const double Ri = (valid && NJ) ? R[i] : 0.0;
  const unsigned beta0 = valid ? p.beta[lr] : 0u;
  const T *newcol = static_cast<const T *>(p.newcol);
  const T nc_lo = (valid && has_prev) ? newcol[i] : (T)0, nc_hi = (valid && has_prev) ? newcol[p.newcol_stride + i] : (T)0;
  iter_raw_load(&sc, p, prev, has_prev != 0);
  prune_rowtest<T, NJ>(p, sc, valid, lr, i, Ri, tinv, k, iter_wfloor(sc, p.xw != nullptr, k, tinv), beta0, nc_lo,
                       nc_hi);
}

template <typename T, bool NJ>
__global__ void __launch_bounds__(SCAN_THREADS, 2) sh_scan_pruned_kernel(ShScanArgs a, PruneArgs p) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ Cand sCand[SCAN_WARPS];
  __shared__ bool sLast;
  __shared__ unsigned long long sRead[SCAN_WARPS];
  __shared__ unsigned short sList[SCAN_WARPS][PRUNE_UNIT];
  const T *__restrict__ D = static_cast<const T *>(a.Dloc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = a.k;
  __shared__ IterRaw sc;
  iter_raw_load(&sc, p, nullptr, false);
  RowScan<T, NJ> rs;
  rs.R = a.R; rs.ids = a.ids; rs.tinv = a.tinv;
  rs.umax = NJ ? a.tinv * __longlong_as_double((long long)sc.rmax) : 0.0;
  rs.bq = INFINITY; rs.b2 = INFINITY; rs.bi = -1; rs.bj = -1; rs.cap = INFINITY;
  rs.sbest = INFINITY; rs.ssi = -1; rs.ssj = -1; rs.uif = 0.f;
  rs.wfloor = iter_wfloor(sc, p.xw != nullptr, k, a.tinv);
  PruneWarp pw;
  pw.gb_seen = INFINITY; pw.nread = 0;
  const int gw = warp * gridDim.x + blockIdx.x, W = gridDim.x * SCAN_WARPS;  // consecutive units on different SMs
  const float cnow = iter_cnow(sc);
  {
    const long long upr = (((long long)k - 1 + SW - 1) / SW + PRUNE_UNIT - 1) / PRUNE_UNIT;  // units per row
    const long long nunits = (long long)sc.wc0 * upr;
    for (long long u = gw; u < nunits; u += W) {
      const int lr = p.worklist[u / upr];
      const int i = sh_slot(lr, a.r, a.G);
      const int sb = (int)(u % upr) * PRUNE_UNIT;
      if ((long long)sb * SW >= i) continue;
      rs.set_row(i, NJ ? a.R[i] : 0.0);
      prune_unit<T, NJ>(rs, pw, D + (int64_t)lr * a.ld, p.E + (int64_t)lr * p.stride, lr, i, sb, cnow, p,
                        sList[warp], lane);
    }
  }
  if (lane == 0) sRead[warp] = pw.nread;
  prune_publish_suggestion<T, NJ>(rs, p);
  Cand c = block_best(rs.bq, rs.bi, rs.bj, rs.b2, a.ids, sCand);
  if (threadIdx.x == 0) {
    a.cand[blockIdx.x] = c;
    unsigned long long tot = 0;
    for (int w2 = 0; w2 < SCAN_WARPS; ++w2) tot += sRead[w2];
    p.cta_read[blockIdx.x] = tot;
    __threadfence();
    unsigned prev = atomicInc(a.counter, gridDim.x - 1);
    sLast = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!sLast) return;
  __threadfence();
  c = reduce_candidates(a.cand, (int)gridDim.x, sCand);
  {
    unsigned long long tot = 0;
    for (int x = threadIdx.x; x < (int)gridDim.x; x += SCAN_THREADS) tot += __ldcg(p.cta_read + x);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    if (lane == 0 && tot) atomicAdd(p.scanned, tot);
  }
  if (threadIdx.x == 0) {
    // every other CTA of this rank has finished (counter): clear the per-iteration state, commit the drift
    *p.gbest = enc_f64(INFINITY);
    p.wcount[2] = p.wcount[0];  // length of this iteration's work list, for the exact mode's resolver
    *p.wcount = 0u;
    prune_commit(p);
  }
  if (a.peers.ex[0] != nullptr) {
    if (threadIdx.x < a.G) {
      ShExHeader *h = ex_header(a.peers.ex[threadIdx.x]);
      h->cand[a.r] = c;
      __threadfence_system();
      post_flag(&h->flag1[a.r], a.tag);
    }
  } else if (threadIdx.x == 0) {
    *a.out = c;
  }
}

static size_t sh_sums_offset(int64_t n, int world, int dtype) {
  const size_t esz = dtype == CAS_F64 ? 8 : 4;
  const size_t chunk = (size_t)sh_chunk((int)(n > 0 ? n : 1), world > 0 ? world : 1);
  size_t per = 3 * chunk * esz;
  if (per < chunk * 8) per = chunk * 8;
  return align_up(sizeof(ShExHeader) + per * (size_t)(world > 0 ? world : 1), 256);
}

// ---- exact mode, strict tail (agglom_exact.cuh): once an ambiguous iteration involved more than XR_SH_HEAVY slots
// every rank re-accumulates the reference's sequential sums of ITS rows before each scan (rows are whole on their
// owner, so the chains are local), publishes them to every rank's `sums` area and the scatter kernel installs them
// as R.  Both kernels are launched for live sizes <= STRICT_TAIL_K and return at once until the flag is raised.
constexpr int XR_SH_HEAVY = 64;

struct ShStrictArgs {
  const double *Dloc;
  int64_t ld;
  int k, G, r, t;
  ShPeers peers;
  size_t sums_off;        // byte offset of the [n] doubles `sums` area inside an exchange buffer
  unsigned long long tag;
  const int32_t *order;   // live slots in id order (replicated)
  const unsigned *gate;
  unsigned *counter;
  double *R;
  unsigned long long *rmax_bits;
  int *epoch;
};

__global__ void __launch_bounds__(256) sh_rowsum_seq_kernel(ShStrictArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  if (!*a.gate) return;
  __shared__ double tile[XR_ROWS][XR_POS + 1];
  __shared__ bool sLast;
  const int nloc = sh_count(a.k, a.r, a.G);
  const int l0 = blockIdx.x * XR_ROWS;
  const int M = max(0, min(XR_ROWS, nloc - l0));
  const double *Dl = a.Dloc;
  const int64_t ld = a.ld;
  const int G = a.G, r = a.r;
  ShPeers peers = a.peers;
  const size_t off = a.sums_off;
  exact_rowsums_generic(
      a.k, a.order, M, tile, [&](int m) { return Dl + (int64_t)(l0 + m) * ld; },
      [&](int m, double sum) {
        const int slot = sh_slot(l0 + m, r, G);
        for (int pr = 0; pr < G; ++pr) reinterpret_cast<double *>(peers.ex[pr] + off)[slot] = sum;
      });
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev = atomicInc(a.counter, gridDim.x - 1);
    sLast = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (sLast && threadIdx.x < a.G) {
    __threadfence_system();
    post_flag(&ex_header(a.peers.ex[threadIdx.x])->flagS[a.r], a.tag);
  }
}

__global__ void __launch_bounds__(256) sh_strict_scatter_kernel(ShStrictArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  if (!*a.gate) return;
  ShExHeader *own = ex_header(a.peers.ex[a.r]);
  if (threadIdx.x < a.G) wait_flag(&own->flagS[threadIdx.x], a.tag, &own->error);
  __syncthreads();
  const double *sums = reinterpret_cast<const double *>(a.peers.ex[a.r] + a.sums_off);
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (v < a.k) {
    s = sums[v];
    a.R[v] = s;
  }
  warp_atomic_absmax(a.rmax_bits, s);
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.epoch = a.t;
}

struct ShColsArgs {
  const void *Dloc;
  int64_t ld;
  int k, G, r;
  int chunk;             // rows per rank chunk in the gather buffer
  const Cand *cand_all;  // [G] gathered candidates
  void *mine;            // this rank's chunk: [3][chunk] values of type T
  const int32_t *ids;
  const int32_t *sizes;
  Sel *sel;
  int32_t *joins;
  int t;
  ShPeers peers;
  unsigned long long tag;
  unsigned *counter;
  // pruned scan (or E == null): the bounds of the two rows the merge rewrites become unknown
  unsigned *E; int64_t e_stride; unsigned *beta;
  // exact NJ mode (P2P exchange only): see "exact mode, sharded" below
  int exact; const double *R; double tinv; PruneArgs p; ExactArgs ex; unsigned *xstate;
  const int *strict_epoch;  // exact mode: == t when this iteration runs on the reference's sequential sums
};

// The part of the cols kernel that needs the selected pair: marks the two rows the merge rewrites unknown,
// publishes this rank's pieces of columns lo / hi / last, the selection and the join.  All threads of all CTAs.
// `vl_pre` (have_vl): this thread's element of column `last`, loaded by the caller before it waited for the peers'
// candidates -- the column does not depend on the selection, and its scattered loads (one per 400 KB row) are the
// slowest of the three.
template <typename T>
__device__ __forceinline__ void sh_cols_body(const ShColsArgs &a, int si, int sj, bool have_vl = false, T vl_pre = (T)0) {
  __shared__ bool sLast;
  const bool p2p = a.peers.ex[0] != nullptr;
  const int lo = min(si, sj), hi = max(si, sj), last = a.k - 1;
  const T *__restrict__ D = static_cast<const T *>(a.Dloc);
  if (lo < 0) {  // no candidate anywhere (a map without a finite entry): flag it, publish nothing
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      Sel s;
      s.lo = -1; s.hi = -1; s.dij = 0.0; s.size_lo = 0; s.size_hi = 0; s.id_lo = -1; s.id_hi = -1;
      *a.sel = s;
      a.joins[2 * a.t] = -1;
      a.joins[2 * a.t + 1] = -1;
    }
  } else {
    if (a.E && blockIdx.x == 0) {
      const int nsk = (a.k + SW - 1) / SW;
      const bool own_lo = sh_owner(lo, a.G) == a.r, own_hi = sh_owner(hi, a.G) == a.r;
      for (int sgm = threadIdx.x; sgm < nsk; sgm += 256) {
        if (own_lo) a.E[(int64_t)sh_local(lo, a.G) * a.e_stride + sgm] = enc_f32(-INFINITY);
        if (own_hi) a.E[(int64_t)sh_local(hi, a.G) * a.e_stride + sgm] = enc_f32(-INFINITY);
      }
      if (threadIdx.x == 0) {
        if (own_lo) a.beta[sh_local(lo, a.G)] = enc_f32(-INFINITY);
        if (own_hi) a.beta[sh_local(hi, a.G)] = enc_f32(-INFINITY);
      }
    }
    const int lr = blockIdx.x * 256 + threadIdx.x;
    if (lr < sh_count(a.k, a.r, a.G)) {
      const T *row = D + (int64_t)lr * a.ld;
      const T va = row[lo], vb = row[hi], vl = have_vl ? vl_pre : row[last];
      if (p2p) {
        // write this rank's chunk straight into every rank's gather area over NVLink
        for (int pr = 0; pr < a.G; ++pr) {
          T *out = reinterpret_cast<T *>(ex_gath(a.peers.ex[pr])) + (size_t)a.r * 3 * a.chunk;
          out[lr] = va;
          out[a.chunk + lr] = vb;
          out[2 * a.chunk + lr] = vl;
        }
      } else {
        T *__restrict__ out = static_cast<T *>(a.mine);
        out[lr] = va;
        out[a.chunk + lr] = vb;
        out[2 * a.chunk + lr] = vl;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      Sel s;
      s.lo = lo; s.hi = hi;
      s.dij = 0.0;  // filled from the gathered row in the merge kernel
      s.size_lo = a.sizes ? a.sizes[lo] : 1;
      s.size_hi = a.sizes ? a.sizes[hi] : 1;
      s.id_lo = a.ids[lo];
      s.id_hi = a.ids[hi];
      *a.sel = s;
      a.joins[2 * a.t] = min(s.id_lo, s.id_hi);
      a.joins[2 * a.t + 1] = max(s.id_lo, s.id_hi);
    }
  }
  if (p2p) {
    // last block of this kernel signals "chunk landed" to every rank (also when nothing could be selected:
    // the peers' merge kernels must not wait for ever).  One system-scope fence per CTA, after the barrier that orders
    // every thread's peer stores before it (fences are cumulative), instead of one per thread.
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      unsigned prev = atomicInc(a.counter, gridDim.x - 1);
      sLast = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (sLast && threadIdx.x < a.G) {
      __threadfence_system();
      post_flag(&ex_header(a.peers.ex[threadIdx.x])->flag2[a.r], a.tag);
    }
  }
}

// ---- exact mode, sharded ----------------------------------------------------------------------------------
// Every rank reduces the same G candidates and holds the same replicated bounds, so all ranks agree whether an
// iteration is ambiguous (a second pair inside the error window, agglom_exact.cuh).  Then, instead of the cols body:
//   cols kernel (X1)   : all CTAs collect the pairs with q~ <= capq among this rank's listed rows and mark their
//                        slots; the last CTA publishes the rank's list of marked slots to every rank (flagA);
//   exact sums kernel  : one CTA per rank: sorted union of the G lists (identical on every rank); for the slots this
//                        rank owns, the reference's sequential row sum and the distances to all other marked slots,
//                        published to every rank (flagB);
//   second cols kernel : every CTA re-evaluates the marked pairs with the sequential sums (redundantly, identically
//                        on every rank), applies the (value, id_lo, id_hi) rule and runs the cols body.
// The two extra kernels are launched every iteration and return at once unless the iteration is ambiguous.
template <typename T>
__device__ __forceinline__ Cand sh_cols_reduce(const ShColsArgs &a, Cand *sCand) {
  const bool p2p = a.peers.ex[0] != nullptr;
  const Cand *cand_all = a.cand_all;
  if (p2p) {
    ShExHeader *own = ex_header(a.peers.ex[a.r]);
    if (threadIdx.x < a.G) wait_flag(&own->flag1[threadIdx.x], a.tag, &own->error);
    __syncthreads();
    cand_all = own->cand;
  }
  return reduce_candidates(cand_all, a.G, sCand);
}

template <typename T>
__global__ void __launch_bounds__(256) sh_cols_kernel(ShColsArgs a) {
  pdl_wait();  // no-ops unless launched with programmatic stream serialization (the pruned loop)
  pdl_launch_dependents();
  __shared__ Cand sCand[8];
  __shared__ bool sLastX;
  // independent of the peers' candidates: in flight during the flag wait
  const int lr_pre = blockIdx.x * 256 + threadIdx.x;
  const bool have_vl = lr_pre < sh_count(a.k, a.r, a.G);
  const T vl_pre = have_vl ? static_cast<const T *>(a.Dloc)[(int64_t)lr_pre * a.ld + (a.k - 1)] : (T)0;
  const int epoch_pre = (a.exact && a.strict_epoch) ? __ldcg(a.strict_epoch) : 0;
  const Cand c = sh_cols_reduce<T>(a, sCand);
  if constexpr (sizeof(T) == 8) {
    if (a.exact && epoch_pre != a.t) {
      const double W = exact_window(a.p.xw, a.k, a.tinv);
      if (c.si >= 0 && c.q2 - c.q <= W) {
        // X1: collect among this rank's listed rows (work-list entries, E and beta are indexed by LOCAL row)
        const double *Dl = static_cast<const double *>(a.Dloc);
        const int G = a.G, r = a.r;
        const int64_t ld = a.ld;
        exact_collect_listed(a.R, a.tinv, a.k, a.p, a.ex, c.q + W, a.t + 1, __ldcg(a.p.wcount + 2), 0u,
                             (int)(blockIdx.x * 8 + (threadIdx.x >> 5)), (int)(gridDim.x * 8),
                             [&](int lr, int &i, const double *&row) {
                               i = sh_slot(lr, r, G);
                               row = Dl + (int64_t)lr * ld;
                             });
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
          unsigned prev = atomicInc(a.xstate + 1, gridDim.x - 1);
          sLastX = (prev == gridDim.x - 1);
        }
        __syncthreads();
        if (!sLastX) return;
        __threadfence();
        ShExHeader *own = ex_header(a.peers.ex[a.r]);
        int M = (int)__ldcg(a.ex.mcount);
        if (M > XT) {
          if (threadIdx.x == 0) atomicExch(&own->error, 2);
          M = XT;
        }
        for (int pr = 0; pr < a.G; ++pr) {
          ShExHeader *h = ex_header(a.peers.ex[pr]);
          for (int e = threadIdx.x; e < M; e += 256) h->listA[a.r][e] = __ldcg(a.ex.mlist + e);
          if (threadIdx.x == 0) h->countA[a.r] = M;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < a.G) post_flag(&ex_header(a.peers.ex[threadIdx.x])->flagA[a.r], a.tag);
        if (threadIdx.x == 0) {
          *a.ex.mcount = 0u;
          a.xstate[0] = 1u;  // deferred: the exact sums kernel and the second cols kernel finish this iteration
        }
        return;
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) a.xstate[0] = 0u;
    } else if (a.exact) {
      if (blockIdx.x == 0 && threadIdx.x == 0) a.xstate[0] = 0u;  // strict tail: nothing to resolve
    }
  }
  sh_cols_body<T>(a, c.si, c.sj, have_vl, vl_pre);
}

struct ShSumsArgs {
  unsigned *strict_flag;  // raised when the union of marked slots exceeds XR_SH_HEAVY (identical on every rank)
  const double *Dloc;
  int64_t ld;
  int k, G, r;
  ShPeers peers;
  unsigned long long tag;
  const int32_t *order;   // live slots in id order (replicated)
  int32_t *ulist;         // out: sorted union of the marked slots
  unsigned *xstate;       // [0] deferred, [2] size of the union
};

__global__ void __launch_bounds__(256) sh_exact_sums_kernel(ShSumsArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  if (!a.xstate[0]) return;
  __shared__ double tile[XR_ROWS][XR_POS + 1];
  __shared__ int sU[XT];
  __shared__ int sMine[XT];
  __shared__ int sT, sNm;
  ShExHeader *own = ex_header(a.peers.ex[a.r]);
  if (threadIdx.x < a.G) wait_flag(&own->flagA[threadIdx.x], a.tag, &own->error);
  if (threadIdx.x == 0) { sT = 0; sNm = 0; }
  __syncthreads();
  // sorted union of the G lists: an entry is kept if no earlier entry (rank, position order) holds the same slot;
  // its rank is the number of kept entries with a smaller slot
  int cnt[SH_MAXG];
  int total = 0;
  for (int g = 0; g < a.G; ++g) { cnt[g] = min(max(own->countA[g], 0), XT); total += cnt[g]; }
  auto entry = [&](int e) {
    int g = 0;
    while (e >= cnt[g]) { e -= cnt[g]; ++g; }
    return own->listA[g][e];
  };
  for (int e = threadIdx.x; e < total; e += 256) {
    const int s = entry(e);
    bool first = true;
    for (int f = 0; f < e && first; ++f) first = entry(f) != s;
    if (!first) continue;
    int rank = 0;
    for (int f = 0; f < total; ++f) {
      const int sf = entry(f);
      if (sf < s) {
        bool ff = true;
        for (int g2 = 0; g2 < f && ff; ++g2) ff = entry(g2) != sf;
        rank += ff ? 1 : 0;
      }
    }
    if (rank < XT) sU[rank] = s;
    atomicAdd(&sT, 1);
  }
  __syncthreads();
  int T = sT;
  if (T > XT) {
    if (threadIdx.x == 0) atomicExch(&own->error, 2);
    T = XT;
  }
  for (int u = threadIdx.x; u < T; u += 256) {
    a.ulist[u] = sU[u];
    if (sh_owner(sU[u], a.G) == a.r) sMine[atomicAdd(&sNm, 1)] = u;
  }
  if (threadIdx.x == 0) {
    a.xstate[2] = (unsigned)T;
    if (T > XR_SH_HEAVY) *a.strict_flag = 1u;
  }
  __syncthreads();
  // the slots this rank owns: sequential sums (reference order) and distances to every marked slot
  const int nm = sNm;
  const double *Dl = a.Dloc;
  const int G = a.G;
  const int64_t ld = a.ld;
  ShPeers peers = a.peers;
  exact_rowsums_generic(
      a.k, a.order, nm, tile, [&](int m) { return Dl + (int64_t)sh_local(sU[sMine[m]], G) * ld; },
      [&](int m, double sum) {
        for (int pr = 0; pr < G; ++pr) ex_header(peers.ex[pr])->xsum[sMine[m]] = sum;
      });
  for (int m = 0; m < nm; ++m) {
    const int u = sMine[m];
    const double *row = Dl + (int64_t)sh_local(sU[u], G) * ld;
    for (int v = threadIdx.x; v < T; v += 256) {
      const double d = row[sU[v]];
      for (int pr = 0; pr < G; ++pr) ex_header(peers.ex[pr])->xd[u][v] = d;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < a.G) post_flag(&ex_header(a.peers.ex[threadIdx.x])->flagB[a.r], a.tag);
}

template <typename T>
__global__ void __launch_bounds__(256) sh_cols2_kernel(ShColsArgs a, const int32_t *ulist) {
  pdl_wait();
  pdl_launch_dependents();
  if (!a.xstate[0]) return;
  __shared__ Cand sCand[8];
  __shared__ double sQ[8];
  __shared__ unsigned long long sK[8];
  __shared__ int sI[8], sJ[8];
  const Cand c = sh_cols_reduce<T>(a, sCand);  // the same G candidates as in the first cols kernel
  ShExHeader *own = ex_header(a.peers.ex[a.r]);
  if (threadIdx.x < a.G) wait_flag(&own->flagB[threadIdx.x], a.tag, &own->error);
  __syncthreads();
  const double W = exact_window(a.p.xw, a.k, a.tinv);
  const double capq = c.q + W;
  const int Tn = (int)a.xstate[2];
  double bq = INFINITY;
  unsigned long long bk = ~0ull;
  int bi = -1, bj = -1;
  for (int e = threadIdx.x; e < Tn * Tn; e += 256) {
    const int u = e / Tn, v = e - u * Tn;
    if (v >= u) continue;  // ulist is sorted by slot: slot_u > slot_v
    const int i = ulist[u], j = ulist[v];
    const double d = own->xd[u][v];
    if (exact_q(d, a.tinv, a.R[i], a.R[j]) > capq) continue;
    const double q = exact_q(d, a.tinv, own->xsum[u], own->xsum[v]);
    const unsigned long long key = make_key(a.ids[i], a.ids[j]);
    if (cand_better(q, key, bq, bk)) { bq = q; bk = key; bi = i; bj = j; }
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
  for (int w = 1; w < 8; ++w)
    if (cand_better(sQ[w], sK[w], bq, bk)) { bq = sQ[w]; bk = sK[w]; bi = sI[w]; bj = sJ[w]; }
  if (bi < 0) { bi = c.si; bj = c.sj; }  // cannot happen: the arg-min of q~ is always among the marked pairs
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    atomicAdd(a.ex.xstats + 0, 1ull);
    atomicAdd(a.ex.xstats + 1, (unsigned long long)Tn);
    atomicMax(a.ex.xstats + 2, (unsigned long long)Tn);
    if (max(bi, bj) != max(c.si, c.sj) || min(bi, bj) != min(c.si, c.sj)) atomicAdd(a.ex.xstats + 3, 1ull);
  }
  __syncthreads();
  sh_cols_body<T>(a, bi, bj);
}

struct ShMergeArgs {
  void *Dloc;
  int64_t ld;
  int k, G, r, chunk;
  const void *gath;  // [G][3][chunk] values of type T
  const Sel *sel;
  double *R;
  int32_t *ids;
  int32_t *sizes;
  double *partial;
  unsigned *counter;
  unsigned long long *rmax_bits;
  int new_id;
  int maintain_R;
  ShPeers peers;
  unsigned long long tag;
  // pruned scan (prune != 0), see MergeArgs in agglom.cu
  int prune; float *uf; double tinv_next; unsigned *dmax;
  const Cand *cand; int ncand; const int32_t *sugg; unsigned long long *gbest; int32_t *pool; int rot;
  // exact NJ mode (replicated like R): bounds of the maintained sums, their maxima, id-ordered slot lists
  int exact; double *Eb; double *Ab; unsigned long long *xw; double *apartial;
  const int32_t *order_old; const int32_t *pos_old; int32_t *order_new; int32_t *pos_new;
  void *newcol; int64_t newcol_stride;  // contiguous copies of the two rewritten columns (PruneArgs), replicated
  double *pmax; unsigned *pufup;        // per-CTA maxima ([grid][4]) and largest uf increase ([grid])
};

// Per-rank warm start of the pruned scan's bound (prune_warm_start, agglom_pruned.cuh): only pairs whose row
// this rank owns can be evaluated here, and only those enter its pool.
// Called either by all MERGE_THREADS threads (UPGMA) or by warps 1.. of the last CTA while warp 0 sums (NJ with
// maintained sums): WARM_SH threads with indices 0 .. WARM_SH - 1.
template <typename T, bool NJ>
__device__ __forceinline__ void sh_warm_start(const ShMergeArgs &a, int lo, int hi, int last) {
  const T *D = static_cast<const T *>(a.Dloc);
  const int64_t ld = a.ld;
  const int G = a.G, r = a.r;
  auto addr = [=](int si, int sj) -> const T * {
    if (sh_owner(si, G) != r) return nullptr;
    return D + (int64_t)sh_local(si, G) * ld + sj;
  };
  if (a.maintain_R)
    prune_warm_start<T, NJ, MERGE_THREADS - 32>(a.pool, a.cand, a.ncand, a.sugg, a.gbest, a.R, a.tinv_next, lo, hi, last,
                                                a.rot, addr, (int)threadIdx.x - 32);
  else
    prune_warm_start<T, NJ, MERGE_THREADS>(a.pool, a.cand, a.ncand, a.sugg, a.gbest, a.R, a.tinv_next, lo, hi, last, a.rot,
                                           addr, (int)threadIdx.x);
}

// Latency structure as in merge_kernel (agglom.cu; profiles/r2_ncu_kernels.md has the per-line stall profile that
// motivated it: a chain of ~8 lazily issued, dependent loads): (1) the thread's replicated state is loaded before the
// NVLink flag wait, the gathered values right after it, all in one wave; (2) the selection and the handful of
// broadcast values that depend on it are loaded once per CTA into shared memory; (3) running maxima go through
// per-CTA slots that the last CTA reduces instead of per-warp load + atomic pairs; (4) in the last CTA warp 0 sums
// while the other warps warm-start gbest.  The arithmetic is unchanged (the merged row is computed redundantly and
// identically on every rank, and identically to the single-GPU loop).
template <typename T, bool NJ>
__global__ void __launch_bounds__(MERGE_THREADS) sh_merge_kernel(ShMergeArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ double sPart[MERGE_THREADS / 32];
  __shared__ double sPartA[MERGE_THREADS / 32];
  __shared__ double sMax[4][MERGE_THREADS / 32];
  __shared__ unsigned sUfUp[MERGE_THREADS / 32];
  __shared__ Sel sSel;
  __shared__ double sDij, sG0l, sG1l;
  __shared__ int sPlo, sPhi, sIdLast, sSzLast;
  __shared__ bool sLast;
  T *__restrict__ D = static_cast<T *>(a.Dloc);
  const T *g = static_cast<const T *>(a.gath);
  const int k = a.k, last = k - 1, G = a.G;
  const int64_t ld = a.ld;
  const int64_t cstride = 3 * (int64_t)a.chunk;
  const int v = blockIdx.x * MERGE_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool live = v < k;
  // replicated state of this slot: independent of the exchange, in flight during the flag wait
  const double Rv0 = (live && a.maintain_R) ? a.R[v] : 0.0;
  const double Eb0 = (live && a.exact) ? a.Eb[v] : 0.0, Ab0 = (live && a.exact) ? a.Ab[v] : 0.0;
  const float uf0 = (live && a.prune && a.maintain_R) ? a.uf[v] : 0.f;
  const int ord0 = (live && a.order_old) ? a.order_old[v] : 0;
  if (a.peers.ex[0] != nullptr) {
    ShExHeader *own = ex_header(a.peers.ex[a.r]);
    if (threadIdx.x < a.G) wait_flag(&own->flag2[threadIdx.x], a.tag, &own->error);
    __syncthreads();
    g = reinterpret_cast<const T *>(ex_gath(a.peers.ex[a.r]));
  }
  auto gathered = [&](int which, int slot) -> T {
    return g[(int64_t)sh_owner(slot, G) * cstride + (int64_t)which * a.chunk + sh_local(slot, G)];
  };
  // the slot's three gathered values and the selection: one wave
  const T gv0 = live ? gathered(0, v) : (T)0, gv1 = live ? gathered(1, v) : (T)0, gv2 = live ? gathered(2, v) : (T)0;
  if (threadIdx.x == 0) sSel = *a.sel;
  __syncthreads();
  const Sel s = sSel;
  if (s.lo < 0) return;  // no candidate (sh_cols_body)
  const int lo = s.lo, hi = s.hi;
  // broadcast values that depend on the selection: one thread each, then shared memory
  if (threadIdx.x == 0) sDij = (double)gathered(0, hi);  // d(hi, lo)
  if (threadIdx.x == 32) sG0l = (double)gathered(0, last);
  if (threadIdx.x == 64) sG1l = (double)gathered(1, last);
  if (threadIdx.x == 96) sPlo = a.order_old ? a.pos_old[lo] : 0;
  if (threadIdx.x == 128) sPhi = a.order_old ? a.pos_old[hi] : 0;
  if (threadIdx.x == 160) sIdLast = a.ids[last];
  if (threadIdx.x == 192) sSzLast = a.sizes ? a.sizes[last] : 0;
  __syncthreads();
  const double dij = sDij;
  auto merged = [&](double da, double db) -> double {
    if (NJ) return __dmul_rn(0.5, __dsub_rn(__dadd_rn(da, db), dij));
    const double zi = (double)s.size_lo, zj = (double)s.size_hi;
    return __ddiv_rn(__dadd_rn(__dmul_rn(zi, da), __dmul_rn(zj, db)), (double)(s.size_lo + s.size_hi));
  };
  const bool own_lo = sh_owner(lo, G) == a.r, own_hi = sh_owner(hi, G) == a.r;
  T *row_lo = own_lo ? D + (int64_t)sh_local(lo, G) * ld : nullptr;
  T *row_hi = own_hi ? D + (int64_t)sh_local(hi, G) * ld : nullptr;
  // merged distance of the node that sits in the last slot (needed at (lo,hi) / (hi,lo))
  T nv_last = (T)0;
  if (hi != last) nv_last = round_to<T>(merged(sG0l, sG1l));

  double contrib = 0.0, mine = 0.0;
  float uf_up = 0.f;
  double ebv = 0.0, abv = 0.0;
  if (live) {
    const bool own_v = sh_owner(v, G) == a.r;
    T *row_v = own_v ? D + (int64_t)sh_local(v, G) * ld : nullptr;
    if (v == lo) {
      if (own_lo) row_lo[lo] = (T)0;
      if (own_hi && hi != last) row_hi[lo] = nv_last;
      a.ids[lo] = a.new_id;
      if (a.sizes) a.sizes[lo] = s.size_lo + s.size_hi;
    } else if (v == hi) {
      if (own_hi) row_hi[hi] = (T)0;
      if (own_lo && hi != last) row_lo[hi] = nv_last;
    } else {
      const double da = (double)gv0, db = (double)gv1;
      const T nvs = round_to<T>(merged(da, db));
      contrib = (double)nvs;
      const bool moves = (v == last && hi != last);
      const int vdst = moves ? hi : v;
      double Rv = 0.0;
      if (a.maintain_R) {
        const double x1 = __dsub_rn(Rv0, da), x2 = __dsub_rn(x1, db);
        Rv = __dadd_rn(x2, contrib);
        mine = Rv;
        a.R[vdst] = Rv;
        if (a.exact) {  // as in merge_kernel (agglom.cu)
          ebv = __dadd_ru(Eb0, __dmul_ru(__dadd_ru(__dadd_ru(fabs(x1), fabs(x2)), fabs(Rv)), EXACT_U1));
          abv = fmax(__dsub_ru(__dsub_ru(__dadd_ru(Ab0, fabs(contrib)), fabs(da)), fabs(db)), 0.0);
          a.Eb[vdst] = ebv;
          a.Ab[vdst] = abv;
        }
      }
      if (hi != last && v != last) {
        const T l = gv2;
        if (own_hi) row_hi[v] = l;
        if (own_v) row_v[hi] = l;
        if (a.prune) static_cast<T *>(a.newcol)[a.newcol_stride + v] = l;
      }
      if (own_lo) row_lo[v] = nvs;
      if (own_v) row_v[lo] = nvs;
      if (a.prune) static_cast<T *>(a.newcol)[vdst] = nvs;
      if (a.prune && a.maintain_R) {
        const float un = __double2float_rn(__dmul_rn(a.tinv_next, Rv));
        if (!moves) uf_up = un - uf0;
        a.uf[vdst] = un;
      }
      if (moves) {
        a.ids[hi] = sIdLast;
        if (a.sizes) a.sizes[hi] = sSzLast;
      }
    }
    // id-ordered list of live slots (exact mode), as in merge_kernel
    if (a.order_old) {
      const int p = v;
      const int p_lo = sPlo, p_hi = sPhi;
      if (p != p_lo && p != p_hi) {
        int sl = ord0;
        if (sl == last) sl = hi;
        const int np = p - (p > p_lo ? 1 : 0) - (p > p_hi ? 1 : 0);
        a.order_new[np] = sl;
        a.pos_new[sl] = np;
      }
      if (p == 0) {
        a.order_new[k - 2] = lo;
        a.pos_new[lo] = k - 2;
      }
    }
  }
  if (a.prune && !a.maintain_R) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      unsigned prev = atomicInc(a.counter, gridDim.x - 1);
      sLast = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (sLast) sh_warm_start<T, NJ>(a, lo, hi, last);
    return;
  }
  if (!a.maintain_R) return;
  // per-warp reductions: the new row's sum (fixed shuffle tree: deterministic), |new row| and the maxima
  double acontrib = fabs(contrib);
  double m0 = fabs(mine), m1 = ebv, m2 = abv, m3 = acontrib;
  unsigned up = enc_f32(uf_up);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    contrib = __dadd_rn(contrib, __shfl_down_sync(0xffffffffu, contrib, off));
    m0 = fmax(m0, __shfl_down_sync(0xffffffffu, m0, off));
    if (a.exact) {
      acontrib = __dadd_rn(acontrib, __shfl_down_sync(0xffffffffu, acontrib, off));
      m1 = fmax(m1, __shfl_down_sync(0xffffffffu, m1, off));
      m2 = fmax(m2, __shfl_down_sync(0xffffffffu, m2, off));
      m3 = fmax(m3, __shfl_down_sync(0xffffffffu, m3, off));
    }
  }
  if (a.prune) up = __reduce_max_sync(0xffffffffu, up);
  if (lane == 0) {
    sPart[warp] = contrib; sPartA[warp] = acontrib;
    sMax[0][warp] = m0; sMax[1][warp] = m1; sMax[2][warp] = m2; sMax[3][warp] = m3;
    sUfUp[warp] = up;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0, asum = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
    unsigned u = enc_f32(0.f);
    for (int w = 0; w < MERGE_THREADS / 32; ++w) {
      sum = __dadd_rn(sum, sPart[w]);
      asum = __dadd_rn(asum, sPartA[w]);
      x0 = fmax(x0, sMax[0][w]); x1 = fmax(x1, sMax[1][w]); x2 = fmax(x2, sMax[2][w]); x3 = fmax(x3, sMax[3][w]);
      u = max(u, sUfUp[w]);
    }
    a.partial[blockIdx.x] = sum;
    a.apartial[blockIdx.x] = asum;
    double *pm = a.pmax + 4 * (size_t)blockIdx.x;
    pm[0] = x0; pm[1] = x1; pm[2] = x2; pm[3] = x3;
    a.pufup[blockIdx.x] = u;
    __threadfence();
    unsigned prev = atomicInc(a.counter, gridDim.x - 1);
    sLast = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!sLast) return;
  // Last CTA (everything it reads of the other CTAs' results is read through L2)
  if (warp == 0) {
    const int nb = (int)gridDim.x;
    double sum, asum;
    sum_partials_one_warp(a.partial, a.apartial, nb, a.exact != 0, sum, asum);  // as merge_kernel
    double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
    unsigned u = enc_f32(0.f);
    for (int b0 = 0; b0 < nb; b0 += 128) {  // four slots per lane in flight
      double2 pa[4], pb[4];
      unsigned pu[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int b = b0 + q * 32 + lane;
        const bool in = b < nb;
        const double2 *pm = reinterpret_cast<const double2 *>(a.pmax + 4 * (size_t)(in ? b : 0));
        pa[q] = in ? __ldcg(pm) : make_double2(0.0, 0.0);
        pb[q] = (in && a.exact) ? __ldcg(pm + 1) : make_double2(0.0, 0.0);
        pu[q] = (in && a.prune) ? __ldcg(a.pufup + b) : enc_f32(0.f);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        x0 = fmax(x0, pa[q].x); x1 = fmax(x1, pa[q].y); x2 = fmax(x2, pb[q].x); x3 = fmax(x3, pb[q].y);
        u = max(u, pu[q]);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      x0 = fmax(x0, __shfl_down_sync(0xffffffffu, x0, off));
      x1 = fmax(x1, __shfl_down_sync(0xffffffffu, x1, off));
      x2 = fmax(x2, __shfl_down_sync(0xffffffffu, x2, off));
      x3 = fmax(x3, __shfl_down_sync(0xffffffffu, x3, off));
    }
    u = __reduce_max_sync(0xffffffffu, u);
    if (lane == 0) {
      a.R[lo] = sum;
      x0 = fmax(x0, fabs(sum));
      atomicMax(a.rmax_bits, (unsigned long long)__double_as_longlong(x0));
      if (a.prune) {
        a.uf[lo] = __double2float_rn(__dmul_rn(a.tinv_next, sum));
        atomicMax(a.dmax, u);
      }
      if (a.exact) {
        const double ab = up30(asum);
        const double eb = up30(new_row_sum_depth(nb) * EXACT_U1 * ab);
        a.Ab[lo] = ab;
        a.Eb[lo] = eb;
        atomicMax(a.xw + XW_EBMAX, (unsigned long long)__double_as_longlong(fmax(x1, eb)));
        atomicMax(a.xw + XW_ABMAX, (unsigned long long)__double_as_longlong(fmax(x2, ab)));
        atomicMax(a.xw + XW_DABS, (unsigned long long)__double_as_longlong(x3));
      }
    }
  } else if (a.prune) {
    // entries touching slot lo are dropped, so the warm start needs neither R[lo] nor anything warp 0 writes
    sh_warm_start<T, NJ>(a, lo, hi, last);
  }
}

// initial row sums of the local rows (same lane-strided order as rowsum_init_kernel)
// nval = 1: the row sum.  nval = 3 (exact mode): also sum |d| and max |d| of the row, at [chunk + lr], [2 chunk + lr]
// of the rank's part of the gather area (rank stride nval * chunk).
template <typename T>
__global__ void __launch_bounds__(256) sh_rowsum_kernel(const T *__restrict__ D, int64_t ld, int k, int G,
                                                        int r, double *__restrict__ mine, int chunk,
                                                        ShPeers peers, unsigned long long tag,
                                                        unsigned *counter, int nval) {
  __shared__ bool sLast;
  const int lr = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const bool p2p = peers.ex[0] != nullptr;
  if (lr < sh_count(k, r, G)) {
    const T *__restrict__ row = D + (int64_t)lr * ld;
    double s = 0.0, as = 0.0, dm = 0.0;
    for (int j = lane; j < k; j += 32) {
      const double x = (double)row[j];
      s = __dadd_rn(s, x);
      as = __dadd_rn(as, fabs(x));
      dm = fmax(dm, fabs(x));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s = __dadd_rn(s, __shfl_down_sync(0xffffffffu, s, off));
      as = __dadd_rn(as, __shfl_down_sync(0xffffffffu, as, off));
      dm = fmax(dm, __shfl_down_sync(0xffffffffu, dm, off));
    }
    if (lane == 0) {
      if (p2p) {
        for (int pr = 0; pr < G; ++pr) {
          double *o = reinterpret_cast<double *>(ex_gath(peers.ex[pr])) + (size_t)r * nval * chunk;
          o[lr] = s;
          if (nval == 3) { o[chunk + lr] = as; o[2 * chunk + lr] = dm; }
        }
      } else {
        mine[lr] = s;
        if (nval == 3) { mine[chunk + lr] = as; mine[2 * chunk + lr] = dm; }
      }
    }
  }
  if (p2p) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned prev = atomicInc(counter, gridDim.x - 1);
      sLast = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (sLast && threadIdx.x < G) {
      __threadfence_system();
      post_flag(&ex_header(peers.ex[threadIdx.x])->flag2[r], tag);
    }
  }
}

__global__ void sh_rscatter_kernel(const double *gath, int chunk, int k, int G, int r,
                                   double *__restrict__ R, unsigned long long *rmax_bits, ShPeers peers,
                                   unsigned long long tag, float *uf, double tinv, int nval, double *Eb, double *Ab) {
  if (peers.ex[0] != nullptr) {
    ShExHeader *own = ex_header(peers.ex[r]);
    if (threadIdx.x < G) wait_flag(&own->flag2[threadIdx.x], tag, &own->error);
    __syncthreads();
    gath = reinterpret_cast<const double *>(ex_gath(peers.ex[r]));
  }
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0, eb = 0.0, ab = 0.0, dm = 0.0;
  if (v < k) {
    const double *o = gath + (int64_t)sh_owner(v, G) * nval * chunk;
    const int lr = sh_local(v, G);
    s = o[lr];
    R[v] = s;
    if (uf) uf[v] = __double2float_rn(__dmul_rn(tinv, s));
    if (nval == 3) {  // exact mode: bounds as in rowsum_init_kernel (agglom.cu)
      ab = up30(o[chunk + lr]);
      eb = up30((double)((k + 31) / 32 + 6) * EXACT_U1 * ab);
      dm = o[2 * chunk + lr];
      Ab[v] = ab;
      Eb[v] = eb;
    }
  }
  warp_atomic_absmax(rmax_bits, s);
  if (nval == 3) {
    warp_atomic_max_nonneg(rmax_bits + XW_EBMAX, eb);
    warp_atomic_max_nonneg(rmax_bits + XW_ABMAX, ab);
    warp_atomic_max_nonneg(rmax_bits + XW_DABS, dm);
  }
}

__global__ void sh_init_kernel(int n, int32_t *ids, int32_t *sizes, unsigned *counters,
                               unsigned long long *rmax_bits, int32_t *order, int32_t *pos, int32_t *mark,
                               unsigned long long *xstats, unsigned *xstate) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) {
    ids[v] = v;
    sizes[v] = 1;
    if (order) { order[v] = v; pos[v] = v; mark[v] = 0; }
  }
  if (v < 8) { counters[v] = 0; rmax_bits[v] = 0ull; xstate[v] = 0u; }
  if (v < 4) xstats[v] = 0ull;
}

__global__ void sh_prune_init_kernel(unsigned *E, int64_t count, unsigned *beta, int64_t nrows, float *hdr,
                                     unsigned *dmax, unsigned long long *gbest, unsigned *wcount,
                                     unsigned long long *scanned, int32_t *pool) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned ninf = enc_f32(-INFINITY);
  for (int64_t e = x; e < count; e += (int64_t)gridDim.x * blockDim.x) E[e] = ninf;
  for (int64_t e = x; e < nrows; e += (int64_t)gridDim.x * blockDim.x) beta[e] = ninf;
  if (x < 2 * POOL) pool[x] = -1;
  if (x == 0) {
    hdr[0] = 0.f;
    *dmax = enc_f32(0.f);
    *gbest = enc_f64(INFINITY);
    *wcount = 0u;
    *scanned = 0ull;
  }
}

__global__ void sh_last2_kernel(const int32_t *ids, int32_t *last2) {
  if (threadIdx.x == 0) {
    last2[0] = min(ids[0], ids[1]);
    last2[1] = max(ids[0], ids[1]);
  }
}

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen (so the library also loads on boxes without NCCL)

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};

struct UniqueId128 { char bytes[128]; };
typedef int (*CommInitRankFn)(void **, int, UniqueId128, int);

static NcclApi g_nccl;
static CommInitRankFn g_comm_init = nullptr;

static int nccl_load() {
  if (g_nccl.handle) return CAS_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *nm : names) {
    h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return CAS_ERR_NCCL;
  }
  g_nccl.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
  g_comm_init = (CommInitRankFn)dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
  g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_comm_init || !g_nccl.CommDestroy || !g_nccl.AllGather) {
    set_error("libnccl lacks a required symbol");
    return CAS_ERR_NCCL;
  }
  g_nccl.handle = h;
  return CAS_OK;
}

#define CAS_NCCL_TRY(expr)                                                                  \
  do {                                                                                      \
    int _r = (expr);                                                                        \
    if (_r != 0) {                                                                          \
      set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
      return CAS_ERR_NCCL;                                                                  \
    }                                                                                       \
  } while (0)

constexpr int NCCL_INT8 = 0;  // ncclInt8 / ncclChar

// ---------------------------------------------------------------------------------------------

struct ShWorkspace {
  double *R;
  int32_t *ids, *sizes;
  Cand *cand;
  Cand *cand_all;
  Sel *sel;
  double *partial;
  unsigned *counters;
  unsigned long long *rmax_bits;
  char *gath;  // max(G*3*chunk*esz, G*chunk*8) bytes
  size_t gath_bytes;
  // pruned scan
  unsigned *E; unsigned *beta; float *uf; float *hdr; unsigned *dmax; unsigned long long *gbest;
  unsigned long long *scanned; unsigned long long *cta_read; int32_t *pool; int32_t *worklist; unsigned *wcount;
  int32_t *sugg; int64_t lb_stride;
  // exact NJ mode
  double *Eb, *Ab, *apartial; int32_t *order[2], *pos[2], *mark, *mlist, *ulist; unsigned *mcount, *xstate;
  unsigned long long *xstats;
  double *newcol;
  unsigned *strict_flag; int *strict_epoch;
  double *pmax; unsigned *pufup;
  size_t bytes;
};

static ShWorkspace sh_carve(void *base, int64_t n, int G, size_t esz) {
  Carver c(base);
  ShWorkspace w;
  w.R = c.take<double>(n);
  w.ids = c.take<int32_t>(n);
  w.sizes = c.take<int32_t>(n);
  w.cand = c.take<Cand>(MAX_SCAN_CTAS);
  w.cand_all = c.take<Cand>(64);
  w.sel = c.take<Sel>(4);
  w.partial = c.take<double>((n + MERGE_THREADS - 1) / MERGE_THREADS + 1);
  w.counters = c.take<unsigned>(64);
  w.rmax_bits = c.take<unsigned long long>(8);
  const size_t chunk = (size_t)sh_chunk((int)n, G);
  size_t per = 3 * chunk * esz;
  if (per < chunk * 8) per = chunk * 8;
  w.gath_bytes = per * G;
  w.gath = c.take<char>(w.gath_bytes);
  w.lb_stride = ((n + SW - 1) / SW + 31) / 32 * 32;
  w.E = c.take<unsigned>((size_t)w.lb_stride * chunk);
  w.beta = c.take<unsigned>(chunk + 32);
  w.uf = c.take<float>((size_t)n + 32);
  w.hdr = c.take<float>(8);
  w.dmax = c.take<unsigned>(4);
  w.wcount = w.dmax + 1;
  w.gbest = c.take<unsigned long long>(2);
  w.scanned = w.gbest + 1;
  w.cta_read = c.take<unsigned long long>(MAX_SCAN_CTAS);
  w.pool = c.take<int32_t>(2 * POOL);
  w.worklist = c.take<int32_t>(chunk + 32);
  w.sugg = c.take<int32_t>(2 * MAX_SCAN_CTAS);
  w.Eb = c.take<double>(n);
  w.Ab = c.take<double>(n);
  w.apartial = c.take<double>((n + MERGE_THREADS - 1) / MERGE_THREADS + 1);
  for (int b = 0; b < 2; ++b) { w.order[b] = c.take<int32_t>(n); w.pos[b] = c.take<int32_t>(n); }
  w.mark = c.take<int32_t>(n);
  w.mlist = c.take<int32_t>(n);
  w.ulist = c.take<int32_t>(XT);
  w.mcount = c.take<unsigned>(4);
  w.xstate = c.take<unsigned>(8);
  w.xstats = c.take<unsigned long long>(4);
  w.newcol = c.take<double>(2 * (size_t)n);
  w.strict_flag = c.take<unsigned>(4);
  w.strict_epoch = reinterpret_cast<int *>(w.strict_flag + 1);
  w.pmax = c.take<double>(4 * ((n + MERGE_THREADS - 1) / MERGE_THREADS + 1));
  w.pufup = c.take<unsigned>((n + MERGE_THREADS - 1) / MERGE_THREADS + 1);
  w.bytes = c.used();
  return w;
}

struct ShRank {
  void *D;
  ShWorkspace w;
  int r;
};

// Runs the loop for the local virtual ranks `ranks` (one in real multi-process mode, G in the
// emulation).  Exchange modes:
//   peers != nullptr : P2P -- kernels store into every rank's exchange buffer and poll tagged flags
//   comm  != nullptr : NCCL all-gathers
//   neither          : single-process emulation with shared gather buffers (no exchange at all)
template <typename T, bool NJ>
static int sh_run(std::vector<ShRank> &ranks, int64_t n, int64_t ld, int G, int64_t max_iters, void *comm,
                  const ShPeers *peers_in, unsigned long long epoch, int32_t *joins, int32_t *last2,
                  cudaStream_t stream, int mode) {
  const int sms = sm_count();
  const int target_ctas = sms * 4;
  ShPeers peers;
  memset(&peers, 0, sizeof(peers));
  if (peers_in) peers = *peers_in;
  const bool p2p = peers.ex[0] != nullptr;
  const bool emu = (comm == nullptr) && !p2p;
  const unsigned long long tag0 = epoch << 32;
  Cand *cand_all = ranks[0].w.cand_all;
  char *gath = ranks[0].w.gath;
  // exact (adaptive strict) NJ: float64, P2P exchange (real or emulated), pruned scan down to the last iteration
  const bool exact = NJ && sizeof(T) == 8 && mode == CAS_MODE_EXACT;
  if (exact && !p2p) {
    set_error("the sharded exact mode needs the P2P exchange (CAS_SHARDED_EXCHANGE=p2p)");
    return CAS_ERR_UNSUPPORTED;
  }
  const bool prune = (exact || prune_enabled_sharded(NJ)) && n > 2;
  const int min_k = exact ? 0 : prune_min_k();  // exact pruned scan (agglom_pruned.cuh; policy in agglom.cu)
  for (auto &rk : ranks) {
    sh_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        (int)n, rk.w.ids, rk.w.sizes, rk.w.counters, rk.w.rmax_bits, exact ? rk.w.order[0] : nullptr,
        exact ? rk.w.pos[0] : nullptr, rk.w.mark, rk.w.xstats, rk.w.xstate);
    CAS_LAUNCH_CHECK();
    CAS_CUDA_TRY(cudaMemsetAsync(rk.w.mcount, 0, 4 * sizeof(unsigned), stream));
    {
      const int init[4] = {0, -1, 0, 0};  // strict tail: flag clear, no strict iteration yet
      CAS_CUDA_TRY(cudaMemcpyAsync(rk.w.strict_flag, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    }
    if (prune) {
      sh_prune_init_kernel<<<sms * 4, 256, 0, stream>>>(rk.w.E, rk.w.lb_stride * sh_chunk((int)n, G), rk.w.beta,
                                                        sh_chunk((int)n, G), rk.w.hdr, rk.w.dmax, rk.w.gbest,
                                                        rk.w.wcount, rk.w.scanned, rk.w.pool);
      CAS_LAUNCH_CHECK();
    }
  }
  if (NJ && n > 2) {
    const int chunk = sh_chunk((int)n, G);
    const int nval = exact ? 3 : 1;
    for (auto &rk : ranks) {
      const int nloc = sh_count((int)n, rk.r, G);
      double *mine = reinterpret_cast<double *>(emu ? gath : rk.w.gath) + (size_t)rk.r * nval * chunk;
      const unsigned blocks = (unsigned)((nloc + 7) / 8 > 0 ? (nloc + 7) / 8 : 1);
      sh_rowsum_kernel<T><<<blocks, 256, 0, stream>>>((const T *)rk.D, ld, (int)n, G, rk.r, mine, chunk, peers,
                                                      tag0 + 1, rk.w.counters + 2, nval);
      CAS_LAUNCH_CHECK();
    }
    if (!emu && !p2p) {
      ShRank &rk = ranks[0];
      double *buf = reinterpret_cast<double *>(rk.w.gath);
      CAS_NCCL_TRY(g_nccl.AllGather(buf + (size_t)rk.r * chunk, buf, (size_t)chunk * 8, NCCL_INT8, comm, stream));
    }
    for (auto &rk : ranks) {
      const double *buf = reinterpret_cast<const double *>(emu ? gath : rk.w.gath);
      sh_rscatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
          buf, chunk, (int)n, G, rk.r, rk.w.R, rk.w.rmax_bits, peers, tag0 + 1, prune ? rk.w.uf : nullptr,
          1.0 / (double)(n - 2), nval, rk.w.Eb, rk.w.Ab);
      CAS_LAUNCH_CHECK();
    }
  }
  int rk_grid[SH_MAXG] = {0};
  for (int64_t t = 0; t + 2 < n; ++t) {
    if (max_iters >= 0 && t >= max_iters) break;
    const int k = (int)(n - t);
    const int chunk = sh_chunk(k, G);
    const int ncb = (k + TC - 1) / TC;
    const int cur = (int)(t & 1);
    const unsigned long long tag = tag0 + 2 + (unsigned long long)t;
    // one-way switch to the dense scan once a rank's share of the map is small (cf. PRUNE_MIN_K in agglom.cu)
    const bool prune_it = prune && (double)k * (double)k > (double)min_k * (double)min_k * (double)G;
    // --- exact mode, strict tail: gated re-accumulation of every row sum (returns at once until the flag is raised)
    if (exact && k <= STRICT_TAIL_K) {
      const size_t sums_off = sh_sums_offset(n, G, sizeof(T) == 8 ? CAS_F64 : CAS_F32);
      std::vector<ShStrictArgs> sx(ranks.size());
      for (size_t x = 0; x < ranks.size(); ++x) {
        ShRank &rk = ranks[x];
        ShStrictArgs &st = sx[x];
        st.Dloc = (const double *)rk.D; st.ld = ld; st.k = k; st.G = G; st.r = rk.r; st.t = (int)t; st.peers = peers;
        st.sums_off = sums_off; st.tag = tag; st.order = rk.w.order[cur]; st.gate = rk.w.strict_flag;
        st.counter = rk.w.counters + 3; st.R = rk.w.R; st.rmax_bits = rk.w.rmax_bits; st.epoch = rk.w.strict_epoch;
        const int nloc = sh_count(k, rk.r, G);
        const unsigned g1 = (unsigned)((nloc + XR_ROWS - 1) / XR_ROWS > 0 ? (nloc + XR_ROWS - 1) / XR_ROWS : 1);
        CAS_CUDA_TRY(launch_pdl(sh_rowsum_seq_kernel, g1, 256u, stream, st));
        CAS_LAUNCH_CHECK();
      }
      for (size_t x = 0; x < ranks.size(); ++x) {
        CAS_CUDA_TRY(launch_pdl(sh_strict_scatter_kernel, (unsigned)((k + 255) / 256), 256u, stream, sx[x]));
        CAS_LAUNCH_CHECK();
      }
    }
    // --- scan
    for (auto &rk : ranks) {
      ShScanArgs sa;
      sa.Dloc = rk.D; sa.ld = ld; sa.k = k; sa.G = G; sa.r = rk.r;
      const int nloc = sh_count(k, rk.r, G);
      sa.nlrb = (nloc + SH_TR - 1) / SH_TR;
      sa.ncb = ncb;
      long long cnt = 0;
      for (int cb = 0; cb < ncb; ++cb) {
        int first = sh_count(cb * TC + 1, rk.r, G) / SH_TR;
        if (sa.nlrb > first) cnt += sa.nlrb - first;
      }
      sa.ntiles = (int)cnt;
      sa.R = rk.w.R; sa.ids = rk.w.ids; sa.tinv = 1.0 / (double)(k - 2); sa.rmax_bits = rk.w.rmax_bits;
      sa.cand = rk.w.cand; sa.counter = rk.w.counters + 0;
      sa.out = (emu ? cand_all : rk.w.cand_all) + rk.r;
      sa.peers = peers; sa.tag = tag;
      int grid = sa.ntiles < target_ctas ? sa.ntiles : target_ctas;
      if (grid > MAX_SCAN_CTAS) grid = MAX_SCAN_CTAS;
      if (grid < 1) grid = 1;
      if (prune_it) {
        PruneArgs pa{};  // split = 0: one work list, short units
        pa.E = rk.w.E; pa.stride = rk.w.lb_stride; pa.beta = rk.w.beta; pa.uf = rk.w.uf; pa.hdr = rk.w.hdr;
        pa.dmax = rk.w.dmax; pa.gbest = rk.w.gbest; pa.worklist = rk.w.worklist; pa.wcount = rk.w.wcount;
        pa.sugg = rk.w.sugg; pa.scanned = rk.w.scanned; pa.cta_read = rk.w.cta_read;
        pa.rmaxb = NJ ? rk.w.rmax_bits : nullptr;
        pa.xw = exact ? rk.w.rmax_bits : nullptr;
        pa.newcol = rk.w.newcol; pa.newcol_stride = n;
        pa.strict_epoch = exact ? rk.w.strict_epoch : nullptr; pa.t = (int)t;
        CAS_CUDA_TRY(launch_pdl(sh_rowtest_kernel<T, NJ>, (unsigned)((nloc + 255) / 256 > 0 ? (nloc + 255) / 256 : 1),
                                256u, stream, pa, (const T *)rk.D, ld, k, G, rk.r, (const double *)rk.w.R, sa.tinv,
                                (const Sel *)rk.w.sel, t > 0 ? 1 : 0));
        CAS_LAUNCH_CHECK();
        grid = sms * 2;
        CAS_CUDA_TRY(launch_pdl(sh_scan_pruned_kernel<T, NJ>, (unsigned)grid, (unsigned)SCAN_THREADS, stream, sa, pa));
      } else {
        sh_scan_kernel<T, NJ><<<grid, SCAN_THREADS, 0, stream>>>(sa);
      }
      CAS_LAUNCH_CHECK();
      rk_grid[rk.r % SH_MAXG] = grid;
    }
    // --- C1: candidates
    if (!emu && !p2p) {
      ShRank &rk = ranks[0];
      CAS_NCCL_TRY(g_nccl.AllGather(rk.w.cand_all + rk.r, rk.w.cand_all, sizeof(Cand), NCCL_INT8, comm, stream));
    }
    // --- cols
    std::vector<ShColsArgs> cas_(ranks.size());
    for (size_t x = 0; x < ranks.size(); ++x) {
      ShRank &rk = ranks[x];
      ShColsArgs &ca = cas_[x];
      ca.Dloc = rk.D; ca.ld = ld; ca.k = k; ca.G = G; ca.r = rk.r; ca.chunk = chunk;
      ca.cand_all = emu ? cand_all : rk.w.cand_all;
      ca.mine = (emu ? gath : rk.w.gath) + (size_t)rk.r * 3 * chunk * sizeof(T);
      ca.ids = rk.w.ids; ca.sizes = NJ ? nullptr : rk.w.sizes; ca.sel = rk.w.sel;
      ca.joins = joins; ca.t = (int)t;
      ca.peers = peers; ca.tag = tag; ca.counter = rk.w.counters + 2;
      ca.E = prune_it ? rk.w.E : nullptr; ca.e_stride = rk.w.lb_stride; ca.beta = rk.w.beta;
      ca.exact = exact ? 1 : 0; ca.R = rk.w.R; ca.tinv = 1.0 / (double)(k - 2);
      ca.p = PruneArgs{};
      ca.p.E = rk.w.E; ca.p.stride = rk.w.lb_stride; ca.p.beta = rk.w.beta; ca.p.uf = rk.w.uf; ca.p.hdr = rk.w.hdr;
      ca.p.dmax = rk.w.dmax; ca.p.gbest = rk.w.gbest; ca.p.worklist = rk.w.worklist; ca.p.wcount = rk.w.wcount;
      ca.p.rmaxb = NJ ? rk.w.rmax_bits : nullptr; ca.p.xw = exact ? rk.w.rmax_bits : nullptr;
      ca.ex = ExactArgs{};
      ca.ex.Eb = rk.w.Eb; ca.ex.Ab = rk.w.Ab; ca.ex.xw = rk.w.rmax_bits; ca.ex.mark = rk.w.mark;
      ca.ex.mlist = rk.w.mlist; ca.ex.mcount = rk.w.mcount; ca.ex.xstats = rk.w.xstats; ca.ex.order = rk.w.order[cur];
      ca.xstate = rk.w.xstate;
      ca.strict_epoch = rk.w.strict_epoch;
      const int nloc = sh_count(k, rk.r, G);
      const unsigned cgrid = (unsigned)((nloc + 255) / 256 > 0 ? (nloc + 255) / 256 : 1);
      if (prune_it) {
        CAS_CUDA_TRY(launch_pdl(sh_cols_kernel<T>, cgrid, 256u, stream, ca));
      } else {
        sh_cols_kernel<T><<<cgrid, 256, 0, stream>>>(ca);
      }
      CAS_LAUNCH_CHECK();
    }
    if (exact) {
      // the two resolver kernels of an ambiguous iteration (return at once otherwise)
      for (auto &rk : ranks) {
        ShSumsArgs xa;
        xa.strict_flag = rk.w.strict_flag;
        xa.Dloc = (const double *)rk.D; xa.ld = ld; xa.k = k; xa.G = G; xa.r = rk.r; xa.peers = peers; xa.tag = tag;
        xa.order = rk.w.order[cur]; xa.ulist = rk.w.ulist; xa.xstate = rk.w.xstate;
        CAS_CUDA_TRY(launch_pdl(sh_exact_sums_kernel, 1u, 256u, stream, xa));
        CAS_LAUNCH_CHECK();
      }
      for (size_t x = 0; x < ranks.size(); ++x) {
        const int nloc = sh_count(k, ranks[x].r, G);
        const unsigned cgrid = (unsigned)((nloc + 255) / 256 > 0 ? (nloc + 255) / 256 : 1);
        CAS_CUDA_TRY(launch_pdl(sh_cols2_kernel<T>, cgrid, 256u, stream, cas_[x], (const int32_t *)ranks[x].w.ulist));
        CAS_LAUNCH_CHECK();
      }
    }
    // --- C2: rows lo, hi, last
    if (!emu && !p2p) {
      ShRank &rk = ranks[0];
      const size_t per = (size_t)3 * chunk * sizeof(T);
      CAS_NCCL_TRY(g_nccl.AllGather(rk.w.gath + (size_t)rk.r * per, rk.w.gath, per, NCCL_INT8, comm, stream));
    }
    // --- merge
    for (auto &rk : ranks) {
      ShMergeArgs ma;
      ma.Dloc = rk.D; ma.ld = ld; ma.k = k; ma.G = G; ma.r = rk.r; ma.chunk = chunk;
      ma.gath = emu ? gath : rk.w.gath;
      ma.sel = rk.w.sel; ma.R = rk.w.R; ma.ids = rk.w.ids; ma.sizes = NJ ? nullptr : rk.w.sizes;
      ma.partial = rk.w.partial; ma.counter = rk.w.counters + 1; ma.rmax_bits = rk.w.rmax_bits;
      ma.new_id = (int)(n + t); ma.maintain_R = NJ ? 1 : 0;
      ma.peers = peers; ma.tag = tag;
      ma.prune = prune_it ? 1 : 0; ma.uf = rk.w.uf; ma.dmax = rk.w.dmax;
      ma.tinv_next = k - 3 > 0 ? 1.0 / (double)(k - 3) : 0.0;
      ma.cand = rk.w.cand; ma.ncand = rk_grid[rk.r % SH_MAXG]; ma.sugg = rk.w.sugg; ma.gbest = rk.w.gbest;
      ma.pool = rk.w.pool; ma.rot = (int)((t * 61) % POOL);
      ma.newcol = rk.w.newcol; ma.newcol_stride = n;
      ma.pmax = rk.w.pmax; ma.pufup = rk.w.pufup;
      ma.exact = exact ? 1 : 0; ma.Eb = rk.w.Eb; ma.Ab = rk.w.Ab; ma.xw = rk.w.rmax_bits; ma.apartial = rk.w.apartial;
      ma.order_old = exact ? rk.w.order[cur] : nullptr; ma.pos_old = exact ? rk.w.pos[cur] : nullptr;
      ma.order_new = exact ? rk.w.order[cur ^ 1] : nullptr; ma.pos_new = exact ? rk.w.pos[cur ^ 1] : nullptr;
      if (prune_it) {
        CAS_CUDA_TRY(launch_pdl(sh_merge_kernel<T, NJ>, (unsigned)((k + MERGE_THREADS - 1) / MERGE_THREADS),
                                (unsigned)MERGE_THREADS, stream, ma));
      } else {
        sh_merge_kernel<T, NJ><<<(unsigned)((k + MERGE_THREADS - 1) / MERGE_THREADS), MERGE_THREADS, 0, stream>>>(ma);
      }
      CAS_LAUNCH_CHECK();
    }
  }
  if (n >= 2) {
    sh_last2_kernel<<<1, 32, 0, stream>>>(ranks[0].w.ids, last2);
    CAS_LAUNCH_CHECK();
  }
  return CAS_OK;
}

static int sh_dispatch(std::vector<ShRank> &ranks, int64_t n, int64_t ld, int G, int dtype, int algo,
                       int64_t max_iters, void *comm, const ShPeers *peers, unsigned long long epoch,
                       int32_t *joins, int32_t *last2, cudaStream_t stream, int mode = CAS_MODE_FAST) {
  CAS_REQUIRE(mode == CAS_MODE_FAST || mode == CAS_MODE_EXACT, "the sharded loop runs in fast or exact mode");
  CAS_REQUIRE(!(mode == CAS_MODE_EXACT && dtype != CAS_F64), "the exact mode requires a float64 map");
  if (algo == CAS_ALGO_NJ) {
    if (dtype == CAS_F64)
      return sh_run<double, true>(ranks, n, ld, G, max_iters, comm, peers, epoch, joins, last2, stream, mode);
    return sh_run<float, true>(ranks, n, ld, G, max_iters, comm, peers, epoch, joins, last2, stream, mode);
  }
  if (dtype == CAS_F64)
    return sh_run<double, false>(ranks, n, ld, G, max_iters, comm, peers, epoch, joins, last2, stream, mode);
  return sh_run<float, false>(ranks, n, ld, G, max_iters, comm, peers, epoch, joins, last2, stream, mode);
}

}  // namespace cas

using namespace cas;

extern "C" {

int64_t cas_sharded_local_rows(int64_t n, int32_t rank, int32_t world) {
  if (world < 1 || rank < 0 || rank >= world || n < 0) return -1;
  return sh_count((int)n, rank, world);
}

int64_t cas_sharded_local_capacity(int64_t n, int32_t world) {
  if (world < 1 || n < 0) return -1;
  return sh_chunk((int)n, world);
}

int cas_sharded_scanned_elements(const void *workspace, int64_t n, int32_t world, int32_t dtype, void *stream,
                                 int64_t *elements) {
  CAS_REQUIRE(workspace && elements, "null pointer argument");
  ShWorkspace w = sh_carve(const_cast<void *>(workspace), n > 0 ? n : 1, world > 0 ? world : 1,
                           dtype == CAS_F64 ? 8 : 4);
  unsigned long long h = 0;
  CAS_CUDA_TRY(cudaMemcpyAsync(&h, w.scanned, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CAS_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  *elements = (int64_t)h;
  return CAS_OK;
}

int cas_sharded_exact_stats(const void *workspace, int64_t n, int32_t world, int32_t dtype, void *stream,
                            int64_t *out4) {
  CAS_REQUIRE(workspace && out4, "null pointer argument");
  ShWorkspace w = sh_carve(const_cast<void *>(workspace), n > 0 ? n : 1, world > 0 ? world : 1,
                           dtype == CAS_F64 ? 8 : 4);
  unsigned long long h[4] = {0, 0, 0, 0};
  CAS_CUDA_TRY(cudaMemcpyAsync(h, w.xstats, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CAS_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  for (int i = 0; i < 4; ++i) out4[i] = (int64_t)h[i];
  return CAS_OK;
}

size_t cas_sharded_workspace_bytes(int64_t n, int32_t world, int32_t dtype) {
  return sh_carve(nullptr, n > 0 ? n : 1, world > 0 ? world : 1, dtype == CAS_F64 ? 8 : 4).bytes;
}

int cas_nccl_unique_id(void *out128) {
  CAS_REQUIRE(out128, "null pointer argument");
  int rc = nccl_load();
  if (rc != CAS_OK) return rc;
  CAS_NCCL_TRY(g_nccl.GetUniqueId(out128));
  return CAS_OK;
}

int cas_nccl_comm_create(const void *id128, int32_t rank, int32_t world, void **comm_out) {
  CAS_REQUIRE(id128 && comm_out, "null pointer argument");
  int rc = nccl_load();
  if (rc != CAS_OK) return rc;
  UniqueId128 id;
  memcpy(id.bytes, id128, 128);
  void *comm = nullptr;
  CAS_NCCL_TRY(g_comm_init(&comm, world, id, rank));
  *comm_out = comm;
  return CAS_OK;
}

int cas_nccl_comm_destroy(void *comm) {
  if (!comm) return CAS_OK;
  int rc = nccl_load();
  if (rc != CAS_OK) return rc;
  CAS_NCCL_TRY(g_nccl.CommDestroy(comm));
  return CAS_OK;
}

static int sh_check(int64_t n, int64_t ld, int dtype, int algo, int world) {
  CAS_REQUIRE(n >= 2 && n < (int64_t)1 << 30, "n=%lld out of range", (long long)n);
  CAS_REQUIRE(ld >= n && ld % 32 == 0, "ld=%lld must be >= n and a multiple of 32", (long long)ld);
  CAS_REQUIRE(dtype == CAS_F32 || dtype == CAS_F64, "bad dtype %d", dtype);
  CAS_REQUIRE(algo == CAS_ALGO_NJ || algo == CAS_ALGO_UPGMA, "bad algo %d", algo);
  CAS_REQUIRE(world >= 1 && world <= 64, "world=%d out of range", world);
  return CAS_OK;
}

int cas_sharded_solve(void *D_local, int64_t n, int64_t ld, int32_t dtype, int32_t algo, int64_t max_iters,
                      int32_t rank, int32_t world, void *nccl_comm, int32_t *joins, int32_t *last2,
                      void *workspace, size_t workspace_bytes, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D_local && joins && last2 && workspace && nccl_comm, "null pointer argument");
  int rc = sh_check(n, ld, dtype, algo, world);
  if (rc != CAS_OK) return rc;
  CAS_REQUIRE(rank >= 0 && rank < world, "bad rank %d", rank);
  CAS_REQUIRE(workspace_bytes >= cas_sharded_workspace_bytes(n, world, dtype), "workspace too small");
  rc = nccl_load();
  if (rc != CAS_OK) return rc;
  std::vector<ShRank> ranks(1);
  ranks[0].D = D_local;
  ranks[0].w = sh_carve(workspace, n, world, dtype == CAS_F64 ? 8 : 4);
  ranks[0].r = rank;
  return sh_dispatch(ranks, n, ld, world, dtype, algo, max_iters, nccl_comm, nullptr, 0, joins, last2,
                     (cudaStream_t)stream);
}

int cas_sharded_solve_emulated(void *const *D_locals, int64_t n, int64_t ld, int32_t dtype, int32_t algo,
                               int64_t max_iters, int32_t world, int32_t *joins, int32_t *last2,
                               void *const *workspaces, size_t workspace_bytes, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D_locals && joins && last2 && workspaces, "null pointer argument");
  int rc = sh_check(n, ld, dtype, algo, world);
  if (rc != CAS_OK) return rc;
  CAS_REQUIRE(workspace_bytes >= cas_sharded_workspace_bytes(n, world, dtype), "workspace too small");
  std::vector<ShRank> ranks(world);
  for (int r = 0; r < world; ++r) {
    CAS_REQUIRE(D_locals[r] && workspaces[r], "null shard pointer for rank %d", r);
    ranks[r].D = D_locals[r];
    ranks[r].w = sh_carve(workspaces[r], n, world, dtype == CAS_F64 ? 8 : 4);
    ranks[r].r = r;
  }
  return sh_dispatch(ranks, n, ld, world, dtype, algo, max_iters, nullptr, nullptr, 0, joins, last2,
                     (cudaStream_t)stream);
}

// ---- P2P exchange buffers ------------------------------------------------------------------
size_t cas_exchange_bytes(int64_t n, int32_t world, int32_t dtype) {
  const size_t esz = dtype == CAS_F64 ? 8 : 4;
  const size_t chunk = (size_t)sh_chunk((int)(n > 0 ? n : 1), world > 0 ? world : 1);
  size_t per = 3 * chunk * esz;
  if (per < chunk * 8) per = chunk * 8;
  // header | gather area | [n] doubles: sequential sums of the exact mode's strict tail
  return sh_sums_offset(n, world, dtype) + align_up((size_t)(n > 0 ? n : 1) * 8, 256) + 256;
}

int cas_exchange_alloc(size_t bytes, void **ptr_out, void *handle64_out) {
  CAS_REQUIRE(ptr_out && handle64_out, "null pointer argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
  void *p = nullptr;
  CAS_CUDA_TRY(cudaMalloc(&p, bytes));
  CAS_CUDA_TRY(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    memset(&h, 0, sizeof(h));  // same-process use (emulation) still works without a handle
  }
  memcpy(handle64_out, &h, 64);
  *ptr_out = p;
  return CAS_OK;
}

int cas_exchange_open(const void *handle64, void **ptr_out) {
  CAS_REQUIRE(handle64 && ptr_out, "null pointer argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  CAS_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return CAS_OK;
}

int cas_exchange_close(void *ptr) {
  if (ptr) CAS_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return CAS_OK;
}

int cas_exchange_free(void *ptr) {
  if (ptr) CAS_CUDA_TRY(cudaFree(ptr));
  return CAS_OK;
}

// 0 = fine, 1 = a flag wait timed out in some kernel (synchronises the stream)
int cas_exchange_error(const void *own_exchange, void *stream, int32_t *error_out) {
  CAS_REQUIRE(own_exchange && error_out, "null pointer argument");
  int err = 0;
  CAS_CUDA_TRY(cudaMemcpyAsync(&err, (const char *)own_exchange + offsetof(ShExHeader, error), sizeof(int),
                               cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CAS_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  *error_out = err;
  return CAS_OK;
}

static int fill_peers(ShPeers &peers, void *const *exchange_ptrs, int world) {
  memset(&peers, 0, sizeof(peers));
  CAS_REQUIRE(world <= SH_MAXG, "P2P exchange supports at most %d ranks", SH_MAXG);
  for (int r = 0; r < world; ++r) {
    CAS_REQUIRE(exchange_ptrs[r] != nullptr, "null exchange buffer for rank %d", r);
    peers.ex[r] = static_cast<char *>(exchange_ptrs[r]);
  }
  return CAS_OK;
}

int cas_sharded_solve_p2p(void *D_local, int64_t n, int64_t ld, int32_t dtype, int32_t algo, int32_t mode,
                          int64_t max_iters, int32_t rank, int32_t world, void *const *exchange_ptrs, uint64_t epoch,
                          int32_t *joins, int32_t *last2, void *workspace, size_t workspace_bytes,
                          void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D_local && joins && last2 && workspace && exchange_ptrs, "null pointer argument");
  int rc = sh_check(n, ld, dtype, algo, world);
  if (rc != CAS_OK) return rc;
  CAS_REQUIRE(rank >= 0 && rank < world, "bad rank %d", rank);
  CAS_REQUIRE(workspace_bytes >= cas_sharded_workspace_bytes(n, world, dtype), "workspace too small");
  ShPeers peers;
  rc = fill_peers(peers, exchange_ptrs, world);
  if (rc != CAS_OK) return rc;
  std::vector<ShRank> ranks(1);
  ranks[0].D = D_local;
  ranks[0].w = sh_carve(workspace, n, world, dtype == CAS_F64 ? 8 : 4);
  ranks[0].r = rank;
  return sh_dispatch(ranks, n, ld, world, dtype, algo, max_iters, nullptr, &peers, epoch, joins, last2,
                     (cudaStream_t)stream, mode);
}

int cas_sharded_solve_emulated_p2p(void *const *D_locals, int64_t n, int64_t ld, int32_t dtype, int32_t algo,
                                   int32_t mode, int64_t max_iters, int32_t world, void *const *exchange_ptrs,
                                   uint64_t epoch, int32_t *joins, int32_t *last2, void *const *workspaces,
                                   size_t workspace_bytes, void *stream) {
  reset_launch_count();
  CAS_REQUIRE(D_locals && joins && last2 && workspaces && exchange_ptrs, "null pointer argument");
  int rc = sh_check(n, ld, dtype, algo, world);
  if (rc != CAS_OK) return rc;
  CAS_REQUIRE(workspace_bytes >= cas_sharded_workspace_bytes(n, world, dtype), "workspace too small");
  ShPeers peers;
  rc = fill_peers(peers, exchange_ptrs, world);
  if (rc != CAS_OK) return rc;
  std::vector<ShRank> ranks(world);
  for (int r = 0; r < world; ++r) {
    CAS_REQUIRE(D_locals[r] && workspaces[r], "null shard pointer for rank %d", r);
    ranks[r].D = D_locals[r];
    ranks[r].w = sh_carve(workspaces[r], n, world, dtype == CAS_F64 ? 8 : 4);
    ranks[r].r = r;
  }
  return sh_dispatch(ranks, n, ld, world, dtype, algo, max_iters, nullptr, &peers, epoch, joins, last2,
                     (cudaStream_t)stream, mode);
}

}  // extern "C"
