// Tensor-core dissimilarity map: the weighted Hamming distance recast as one-hot contractions
// and run on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA with the 128-byte swizzle).  Replaces the same reference code as
// whd_exact.cu (cassiopeia/solver/dissimilarity_functions.py:46-72 over all pairs).
//
// Unweighted ("plain Hamming", costs 0/1/2) -- kind::i8, int32 accumulation, EXACT:
//   per cell one row of A and one of B, K = m*(S+2) int8 columns
//     A = [ U | P | -2 X ]      B = [ P | U | X ]
//   P = present, U = present & cut, X = one-hot over (character, cut state).  Then
//     acc_ij = U.P + P.U - 2 X.X = the integer numerator; the shared-present count is
//   popcount(mask_i & mask_j) of the cells' presence bit masks in the epilogue; D = num / den
//   in IEEE arithmetic => bit-identical to the reference.
// Weighted (mutation priors) -- kind::f16 (bf16 x bf16 -> fp32), cancellation-free form
//     num_ij = sum_{c,s} Xw_i[c,s] * Y_j[c,s] + Y_i[c,s] * Xw_j[c,s],   Y = present & (state != s)
//   every term is >= 0.  The float64 weight is split into three bf16 planes (hi, mid, lo:
//   24 significant bits); each plane has its own fp32 accumulator in TMEM (a product with the
//   exact 0/1 operand is exact, and a plane's partial sums stay within fp32's 24 bits for
//   realistic weight ranges) and the planes are added in float64 in the epilogue.  Relative
//   error vs the exact kernel <= 1e-6 (tests/test_gpu_gemm.py).
//
// Weighted, fixed point (default for prior-weighted maps; `CAS_GEMM_WEIGHTED=bf16` selects the path above) --
//   kind::i8 again.  Every weight is quantised to a 28-bit integer W = round(w 2^F) (F from the largest weight, so
//   the absolute error per weight is 2^-(F+1) and that of a distance at most 2^-F: 6e-8 for weights below 16) and
//   split into four 7-bit digits; digit plane p has its own int32 accumulator in TMEM and
//     acc_p = U_p[I] . P[J] + P[I] . U_p[J] - 2 X_p[I] . X[J]
//   (U_p = digit of the cell's own weight per character, X_p = the digit at the one-hot position) is EXACT integer
//   arithmetic, so the subtraction form needs no cancellation-free doubling; the epilogue recombines
//   num = sum_p 128^p acc_p in int64.  The big one-hot operand of the column block (X[J], -2 folded in) is shared by
//   the four planes: five tile loads feed four MMA groups per K chunk (the bf16 path: eight loads for six groups of
//   half-rate MMAs over twice the bytes).
//
// Kernel: one CTA per 128x128 lower-triangle output tile (tiles rasterised in bands of 12 tile
// rows, column-major inside a band, so concurrently running CTAs share operand blocks in L2),
// 6 warps: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5
// = epilogue (TMEM -> registers -> D, mirrored).  K is walked in 128-byte chunks (4 MMAs each)
// through a full/empty mbarrier ring (3 stages and 2 CTAs per SM for int8, 5 stages for bf16).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace cas {

constexpr int GT = 128;            // tile edge (UMMA M = N = 128)
constexpr int CHUNK_BYTES = 128;   // K bytes per pipeline stage row (one swizzle atom)
constexpr int TILE_BYTES = GT * CHUNK_BYTES;  // 16 KB per operand tile
constexpr int GEMM_THREADS = 320;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int SLOTS = 9;            // operand-tile ring (16 KB each)
constexpr int MAXPW = 8;           // presence-mask words per cell (m <= 512)
constexpr int BAND = 12;           // tile rows per rasterisation band
constexpr int MAX_MAPS = 10;  // fixed-point weighted path: X_p (4), -2 X, U_p (4), P

struct GemmParams {
  CUtensorMap maps[MAX_MAPS];
  int chunks;    // 128-byte K chunks per operand row
  int ntile_rows;
  int pw;        // presence-mask words per cell
  const unsigned long long *pmask;  // [n, pw]
  int64_t n;
  void *D;
  int64_t ld;
  int head_chunks;       // fixed-point weighted path: 128-byte chunks of the [U_p | P] operands (ceil(m / 128))
  const double *scale;   // fixed-point weighted path: device scalar 2^-F
  // Shard mode (shard_world > 0): D holds this rank's rows of the block-cyclic layout (64-row blocks r, r + G, ...:
  // local row lr is global row ((lr / 64) * G + r) * 64 + lr % 64), every column of each: a 128-row tile is two
  // 64-row blocks that are G blocks apart in the operand arrays, loaded as two half boxes; no mirrored stores.
  CUtensorMap maps_half[MAX_MAPS];  // box 128 B x 64 rows
  int shard_world, shard_rank;
  int nloc_rows;   // rows of this rank's shard
  int nloc_tiles;  // ceil(nloc_rows / 128)
  int ncol_tiles;  // ceil(n / 128)
};

__host__ __device__ __forceinline__ int64_t shard_global_row(int64_t lr, int G, int r) {
  return ((lr >> 6) * G + r) * 64 + (lr & 63);
}

// ---- PTX helpers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}

// K-major operand tile [128 rows][128 bytes], 128-byte swizzle (8-row groups of 1024 bytes):
// start address >> 4, LBO unused (0), SBO = 1024 >> 4, descriptor version 1 (sm_100), SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <bool I8>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                     uint32_t accumulate) {
  if (I8) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Per-K-chunk program: which operand tiles are loaded (tensor map, row block I or column block J)
// and which MMAs consume them.  Unweighted: A[I] x B[J].  Weighted: the one-hot "differs" operand Y
// is loaded once per side and shared by the three weight planes -- 8 loads feed 6 MMA groups
// (12 loads if every MMA group fetched its own pair), which matters because this kernel is bound
// by L2->SM bandwidth, not by the tensor pipe.
template <bool I8> struct Prog;
template <> struct Prog<true> {
  static constexpr int NL = 2, NM = 1;
  __device__ static __forceinline__ int load_map(int l) { return l; }
  __device__ static __forceinline__ int load_side(int l) { return l; }  // 0 = rows I, 1 = columns J
  __device__ static __forceinline__ int mma_a(int) { return 0; }
  __device__ static __forceinline__ int mma_b(int) { return 1; }
  __device__ static __forceinline__ int mma_acc(int) { return 0; }
  __device__ static __forceinline__ unsigned release(int) { return 0x3u; }
};
template <> struct Prog<false> {
  // maps: 0..2 = Xw planes, 3 = Y.  loads: 0 Y[J] | 1..3 Xw_p[I] | 4 Y[I] | 5..7 Xw_p[J]
  static constexpr int NL = 8, NM = 6;
  __device__ static __forceinline__ int load_map(int l) { return l == 0 || l == 4 ? 3 : (l < 4 ? l - 1 : l - 5); }
  __device__ static __forceinline__ int load_side(int l) { return (l == 0 || l >= 5) ? 1 : 0; }
  __device__ static __forceinline__ int mma_a(int m) { return m < 3 ? m + 1 : 4; }
  __device__ static __forceinline__ int mma_b(int m) { return m < 3 ? 0 : m + 2; }
  __device__ static __forceinline__ int mma_acc(int m) { return m % 3; }
  // operand tiles whose last consumer is MMA group m (bit l = load l)
  __device__ static __forceinline__ unsigned release(int m) {
    return m == 0 ? 0x02u : m == 1 ? 0x04u : m == 2 ? 0x09u : m == 3 ? 0x20u : m == 4 ? 0x40u : 0x90u;
  }
};

template <typename TO>
struct __align__(1024) GemmSmem {
  static constexpr int HALF = sizeof(TO) == 4 ? 64 : 32;  // columns staged at a time per warp
  uint8_t tile[SLOTS][TILE_BYTES];
  TO stage[8][32][HALF + 1];         // per-epilogue-warp transpose tile for the direct stores
  unsigned long long pm[GT][MAXPW];  // presence masks of the tile's columns
  uint64_t full[SLOTS];
  uint64_t empty[SLOTS];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
};

// linear tile index -> lower-triangle tile (ti >= tj): bands of BAND tile rows, column-major in a band
__device__ __forceinline__ void tile_of(int64_t t, int T, int &ti, int &tj) {
  int b0 = 0;
  for (;;) {
    const int hb = min(BAND, T - b0);
    const int64_t cnt = (int64_t)b0 * hb + (int64_t)hb * (hb + 1) / 2;
    if (t < cnt) {
      const int64_t full = (int64_t)b0 * hb;
      if (t < full) {
        tj = (int)(t / hb);
        ti = b0 + (int)(t % hb);
      } else {
        int64_t rem = t - full;
        int c = 0;
        while (rem >= hb - c) { rem -= hb - c; ++c; }
        tj = b0 + c;
        ti = b0 + c + (int)rem;
      }
      return;
    }
    t -= cnt;
    b0 += hb;
  }
}

// Persistent kernel, one CTA per SM.  NACC accumulator buffers in TMEM: 2 for int8 (the epilogue of
// tile t overlaps the MMAs of tile t+1), 1 for bf16 (3 planes x 128 columns leave no room for a
// second set; the TMA producer still runs ahead into the next tile during the epilogue).
// KIND: 0 = bf16 x 3 planes (weighted), 1 = int8 (unweighted, exact), 2 = int8 x 4 digit planes (weighted, fixed point)
template <int KIND> struct ProgOf { using type = Prog<KIND == 1>; };
struct ProgW8 {
  // per K chunk: load 0 = the operand shared by the four planes, loads 1..4 = the plane operands; MMA group m
  // pairs plane m with the shared tile; chunk types: 0 U_p[I] x P[J], 1 P[I] x U_p[J], 2 X_p[I] x (-2 X)[J]
  static constexpr int NL = 5, NM = 4;
  __device__ static __forceinline__ unsigned release(int m) { return m == 3 ? 0x11u : (1u << (m + 1)); }
};
template <> struct ProgOf<2> { using type = ProgW8; };

template <int KIND, typename TO, int PW>
__global__ void __launch_bounds__(GEMM_THREADS, 1) whd_gemm_kernel(const __grid_constant__ GemmParams p) {
  constexpr bool I8 = KIND != 0;  // int8 operands, int32 accumulators
  using P = typename ProgOf<KIND>::type;
  constexpr int NPL = KIND == 2 ? 4 : (KIND == 0 ? 3 : 1);  // accumulator planes
  constexpr int NACC = KIND == 1 ? 2 : 1;
  constexpr int ACCW = NPL * GT;                       // TMEM columns per accumulator buffer
  constexpr uint32_t TMEM_COLS = KIND == 1 ? 256u : 512u;  // power of two >= NACC * ACCW
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GemmSmem<TO> &sm = *reinterpret_cast<GemmSmem<TO> *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool shard = p.shard_world > 0;
  const int64_t ntiles = shard ? (int64_t)p.nloc_tiles * p.ncol_tiles
                               : (int64_t)p.ntile_rows * (p.ntile_rows + 1) / 2;
  // shard mode: consecutive CTAs take different row tiles of the same column tile (the column operand is shared in L2)
  auto tile_pos = [&](int64_t tile, int &ti, int &tj) {
    if (shard) { tj = (int)(tile / p.nloc_tiles); ti = (int)(tile - (int64_t)tj * p.nloc_tiles); }
    else tile_of(tile, p.ntile_rows, ti, tj);
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.tfull[s], 1); mbar_init(&sm.tempty[s], 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < (KIND == 2 ? 10 : (KIND == 1 ? 2 : 4)); ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.maps[i]) : "memory");
      if (shard) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.maps_half[i]) : "memory");
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&sm.tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer: one operand tile per ring slot =====
    if (lane == 0) {
      uint32_t lc = 0;  // loads issued so far
      const int elems_per_chunk = I8 ? CHUNK_BYTES : CHUNK_BYTES / 2;
      // one operand tile: 128 rows starting at `row` -- or, for the row side of a shard, two 64-row half boxes
      auto load_tile = [&](uint32_t slot, int map, int kc, bool row_side, int ti, int col0) {
        if (shard && row_side) {
          const int r0 = (int)shard_global_row((int64_t)ti * GT, p.shard_world, p.shard_rank);
          const int r1 = (int)shard_global_row((int64_t)ti * GT + 64, p.shard_world, p.shard_rank);
          tma_load_2d(sm.tile[slot], &p.maps_half[map], &sm.full[slot], kc, r0);
          tma_load_2d(sm.tile[slot] + 64 * CHUNK_BYTES, &p.maps_half[map], &sm.full[slot], kc, r1);
        } else {
          tma_load_2d(sm.tile[slot], &p.maps[map], &sm.full[slot], kc, row_side ? ti * GT : col0);
        }
      };
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int ti, tj;
        tile_pos(tile, ti, tj);
        const int col0 = tj * GT;
        for (int c = 0; c < p.chunks; ++c) {
#pragma unroll
          for (int l = 0; l < P::NL; ++l, ++lc) {
            const uint32_t slot = lc % SLOTS, phase = (lc / SLOTS) & 1u;
            mbar_wait(&sm.empty[slot], phase ^ 1u);
            mbar_expect_tx(&sm.full[slot], TILE_BYTES);
            if constexpr (KIND == 2) {
              // maps: 0..3 X_p, 4 -2 X, 5..8 U_p, 9 P
              const int hc = p.head_chunks;
              const int type = c < hc ? 0 : (c < 2 * hc ? 1 : 2);
              const int kc = (type == 0 ? c : (type == 1 ? c - hc : c - 2 * hc)) * CHUNK_BYTES;
              int map, side;  // side 1 = column block J
              if (type == 0) { map = l == 0 ? 9 : 4 + l; side = l == 0 ? 1 : 0; }
              else if (type == 1) { map = l == 0 ? 9 : 4 + l; side = l == 0 ? 0 : 1; }
              else { map = l == 0 ? 4 : l - 1; side = l == 0 ? 1 : 0; }
              load_tile(slot, map, kc, side == 0, ti, col0);
            } else {
              load_tile(slot, P::load_map(l), c * elems_per_chunk, P::load_side(l) == 0, ti, col0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      // instruction descriptor: D format (F32 = 1 / S32 = 2) bits [4,6), A/B format bits [7,10)/[10,13)
      // (BF16 = 1, signed int8 = 1), K-major A and B, N >> 3 at bit 17, M >> 4 at bit 24
      const uint32_t idesc = ((I8 ? 2u : 1u) << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GT >> 3) << 17) |
                             ((uint32_t)(GT >> 4) << 24);
      uint32_t lc0 = 0;  // load counter of the current chunk's first load
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int as = (int)(it % NACC);
        const uint32_t aphase = (uint32_t)((it / NACC) & 1);
        mbar_wait(&sm.tempty[as], aphase ^ 1);  // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        for (int c = 0; c < p.chunks; ++c, lc0 += P::NL) {
#pragma unroll
          for (int m = 0; m < P::NM; ++m) {
            uint32_t la, lb;
            int acc_plane;
            if constexpr (KIND == 2) {
              const int hc = p.head_chunks;
              const bool shared_is_a = c >= hc && c < 2 * hc;  // chunk type 1: P[I] x U_p[J]
              la = lc0 + (shared_is_a ? 0 : 1 + m);
              lb = lc0 + (shared_is_a ? 1 + m : 0);
              acc_plane = m;
            } else {
              la = lc0 + P::mma_a(m);
              lb = lc0 + P::mma_b(m);
              acc_plane = P::mma_acc(m);
            }
            const uint32_t sa = la % SLOTS, sb = lb % SLOTS;
            mbar_wait(&sm.full[sa], (la / SLOTS) & 1u);
            mbar_wait(&sm.full[sb], (lb / SLOTS) & 1u);
            tcgen05_fence_after();
            const uint64_t adesc = make_smem_desc(smem_u32(sm.tile[sa]));
            const uint64_t bdesc = make_smem_desc(smem_u32(sm.tile[sb]));
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * ACCW + acc_plane * GT);
            const bool first = (c == 0) && (m < NPL);  // first MMA group into this accumulator
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              // advance 32 bytes of K inside the swizzle atom: +2 in the (>>4) start-address field
              umma<I8>(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                       (first && kk == 0) ? 0u : 1u);
            }
            // free the operand tiles whose last consumer was this MMA group
#pragma unroll
            for (int l = 0; l < P::NL; ++l)
              if ((P::release(m) >> l) & 1u) tcgen05_commit(&sm.empty[(lc0 + l) % SLOTS]);
          }
        }
        tcgen05_commit(&sm.tfull[as]);
      }
    }
  } else {
    // ===== epilogue: 8 warps; warp w owns TMEM lanes 32*(w%4).. and the column half (w-2)/4 =====
    constexpr int HALF = GemmSmem<TO>::HALF;
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;        // which 64 columns of the tile
    const int lrow = q * 32 + lane;        // row inside the tile == TMEM lane
    const int et = threadIdx.x - 64;       // 0..255
    TO *__restrict__ D = static_cast<TO *>(p.D);
    TO(*stage)[HALF + 1] = sm.stage[warp - 2];
    const double wscale = KIND == 2 ? *p.scale : 1.0;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      int ti, tj;
      tile_pos(tile, ti, tj);
      const int row0 = ti * GT, col0 = tj * GT;  // row0: first row of the tile in D (a LOCAL row in shard mode)
      // global row of this thread's accumulator lane
      const int64_t gi = shard ? ((int64_t)row0 + lrow < p.nloc_rows
                                      ? shard_global_row((int64_t)row0 + lrow, p.shard_world, p.shard_rank) : p.n)
                               : (int64_t)row0 + lrow;
      // rows of D this tile may write; global row of the tile's row rr (diagonal test)
      const int64_t row_limit = shard ? (int64_t)p.nloc_rows : p.n;
      auto global_row = [&](int rr) -> int64_t {
        return shard ? shard_global_row((int64_t)row0 + rr, p.shard_world, p.shard_rank) : (int64_t)row0 + rr;
      };
      // the diagonal crosses the tile: same tile indices (full map) / the tile's global rows overlap its columns (shard)
      const bool diag_tile = shard ? true : (ti == tj);
      const bool mirror = !shard && !diag_tile;  // the transposed tile is written too (full map, off-diagonal tiles)
      const bool whole = (int64_t)row0 + GT <= row_limit && (int64_t)col0 + GT <= p.n &&
                         (!shard || global_row(GT - 1) < p.n);
      // presence masks: this thread's row in registers, the tile's columns in shared memory
      unsigned long long mrow[PW];
      asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's readers are done with sm.pm
#pragma unroll
      for (int w = 0; w < PW; ++w) mrow[w] = gi < p.n ? p.pmask[gi * PW + w] : 0ull;
      if (et < GT) {
        const int64_t gj = (int64_t)col0 + et;
#pragma unroll
        for (int w = 0; w < PW; ++w) sm.pm[et][w] = gj < p.n ? p.pmask[gj * PW + w] : 0ull;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int as = (int)(it % NACC);
      mbar_wait(&sm.tfull[as], (uint32_t)((it / NACC) & 1));
      tcgen05_fence_after();
      // The accumulator row of a lane is written twice: mirrored (column gi of rows gj: the 32 lanes
      // hit 32 consecutive addresses) straight from registers, and direct (row gi) through the
      // warp's transpose tile so that those stores are coalesced too.
      if constexpr (KIND == 2 && sizeof(TO) == 4) {
        // Fixed-point weighted, float32 output.  The four digit planes fill TMEM (no second accumulator set), so
        // the numerators of the thread's 64 columns are first pulled into registers -- digits recombined with three
        // fused multiply-adds in float32, 2^-23 relative, inside the stated 1e-6 -- and the accumulator is handed
        // back at once: popcounts, divisions and stores then overlap the next tile's MMAs.
        const int cbase = ch * 64;
        float numf[64];
        const float sc = (float)wscale;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t v0[16], v1[16], v2[16], v3[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACCW + cbase + g * 16);
          tmem_ld16(taddr, v0);
          tmem_ld16(taddr + GT, v1);
          tmem_ld16(taddr + 2 * GT, v2);
          tmem_ld16(taddr + 3 * GT, v3);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            float x = fmaf((float)(int)v3[e], 128.f, (float)(int)v2[e]);
            x = fmaf(x, 128.f, (float)(int)v1[e]);
            x = fmaf(x, 128.f, (float)(int)v0[e]);
            numf[g * 16 + e] = x * sc;
          }
        }
        tcgen05_fence_before();
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.tempty[as])) : "memory");
        TO *mir0 = D + ((int64_t)col0 + cbase) * p.ld + gi;
#pragma unroll
        for (int e = 0; e < 64; ++e) {
          int den = 0;
#pragma unroll
          for (int w = 0; w < PW; ++w) den += __popcll(mrow[w] & sm.pm[cbase + e][w]);
          float d = den > 0 ? __fdividef(numf[e], (float)den) : 0.f;
          if (diag_tile && ((int64_t)col0 + cbase + e == gi)) d = 0.f;
          stage[lane][e] = (TO)d;
          if (mirror && (whole || (gi < p.n && (int64_t)col0 + cbase + e < p.n))) mir0[(int64_t)e * p.ld] = (TO)d;
        }
        __syncwarp();
        TO *dst = D + ((int64_t)row0 + q * 32) * p.ld + col0 + cbase;
        for (int rr = 0; rr < 32; ++rr, dst += p.ld) {
          if (!whole && ((int64_t)row0 + q * 32 + rr >= row_limit || global_row(q * 32 + rr) >= p.n)) break;
#pragma unroll
          for (int h = 0; h < HALF; h += 32)
            if (whole || (int64_t)col0 + cbase + h + lane < p.n) dst[h + lane] = stage[rr][h + lane];
        }
        __syncwarp();
      } else
      for (int part = 0; part < 64; part += HALF) {
        const int cbase = ch * 64 + part;
        for (int cb = cbase; cb < cbase + HALF; cb += 16) {
          uint32_t v0[16], v1[16], v2[16], v3[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACCW + cb);
          tmem_ld16(taddr, v0);
          if (KIND != 1) {
            tmem_ld16(taddr + GT, v1);
            tmem_ld16(taddr + 2 * GT, v2);
          }
          if (KIND == 2) tmem_ld16(taddr + 3 * GT, v3);
          tmem_ld_wait();
          TO *mir = D + ((int64_t)col0 + cb) * p.ld + gi;  // element (gj0, gi), next column = + ld
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            int den = 0;
#pragma unroll
            for (int w = 0; w < PW; ++w) den += __popcll(mrow[w] & sm.pm[cb + e][w]);
            double num;
            if (KIND == 1) {
              num = (double)(int)v0[e];
            } else if (KIND == 2) {
              // digits recombined in int64 (exact), then one scaling by the power of two 2^-F
              const long long fixed = (long long)(int)v0[e] + ((long long)(int)v1[e] << 7) +
                                      ((long long)(int)v2[e] << 14) + ((long long)(int)v3[e] << 21);
              num = (double)fixed * wscale;
            } else {
              num = ((double)__uint_as_float(v0[e]) + (double)__uint_as_float(v1[e])) + (double)__uint_as_float(v2[e]);
            }
            TO d = (TO)0;
            if (den > 0) {
              if (sizeof(TO) == 8) {
                d = (TO)__ddiv_rn(num, (double)den);
              } else if (KIND == 1) {
                // small integers: the correctly rounded float quotient equals fl32(fl64(num/den))
                d = (TO)__fdiv_rn((float)num, (float)den);
              } else {
                d = (TO)__double2float_rn(__ddiv_rn(num, (double)den));
              }
            }
            if (diag_tile && ((int64_t)col0 + cb + e == gi)) d = (TO)0;
            stage[lane][(cb - cbase) + e] = d;
            if (mirror && (whole || (gi < p.n && (int64_t)col0 + cb + e < p.n))) mir[(int64_t)e * p.ld] = d;
          }
        }
        if (part + HALF == 64) {
          // all TMEM reads of this tile are done: hand the accumulator buffer back to the MMA warp
          tcgen05_fence_before();
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&sm.tempty[as])) : "memory");
        }
        __syncwarp();
        // direct rows: lane l writes columns l (+32) of each of the warp's 32 rows
        TO *dst = D + ((int64_t)row0 + q * 32) * p.ld + col0 + cbase;
        for (int rr = 0; rr < 32; ++rr, dst += p.ld) {
          if (!whole && ((int64_t)row0 + q * 32 + rr >= row_limit || global_row(q * 32 + rr) >= p.n)) break;
#pragma unroll
          for (int h = 0; h < HALF; h += 32)
            if (whole || (int64_t)col0 + cbase + h + lane < p.n) dst[h + lane] = stage[rr][h + lane];
        }
        __syncwarp();
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// =============================================================================================
// 2-SM variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x N2 output tile
// (N2 = 256 for int8, 128 for bf16 whose three planes already fill TMEM).  Each CTA stages its own
// 128 rows of the row-block operand and HALF of the column-block operand; one thread of the leader
// CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and writes
// each CTA's 128 accumulator rows into that CTA's TMEM.  Per MMA cycle every SM fetches half as many
// operand bytes as the 1-SM kernel above, which is what bounds that kernel (L2->SM ~46 B/clk/SM).
// =============================================================================================
constexpr int BAND2 = 6;  // 256-row tile rows per rasterisation band

struct Gemm2Params {
  CUtensorMap maps[4];       // box 128 B x 128 rows
  CUtensorMap maps_half[4];  // box 128 B x 64 rows (bf16: half of the 128-column block)
  int chunks;
  int nrow_tiles;            // ceil(n / 256)
  int ncol_tiles;            // ceil(n / N2)
  int pw;
  const unsigned long long *pmask;
  int64_t n;
  void *D;
  int64_t ld;
  int mchars;                // characters (int8 float path: side of the quotient table), 0 = no table
};

// QTAB floats of shared memory hold the table of correctly rounded quotients num / den of the unweighted
// float32 epilogue (num <= 2m, den <= m): one LDS replaces two conversions and an IEEE division per pair.
template <typename TO, int PW>
struct __align__(1024) Gemm2Smem {
  static constexpr int NSLOT = sizeof(TO) == 4 ? 10 : 8;
  static constexpr int HALF = 32;
  static constexpr int QTAB = (sizeof(TO) == 4 && PW == 1) ? 7680 : 1;
  uint8_t tile[NSLOT][TILE_BYTES];
  TO stage[8][32][HALF + 1];
  float qtab[QTAB];
  unsigned long long pm[256][PW];
  uint64_t full[NSLOT];
  uint64_t empty[NSLOT];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
};

static_assert(sizeof(Gemm2Smem<float, 1>) + 1024 <= 227 * 1024 && sizeof(Gemm2Smem<double, 8>) + 1024 <= 227 * 1024 &&
                  sizeof(Gemm2Smem<float, 8>) + 1024 <= 227 * 1024,
              "SM-pair kernel: shared memory exceeds 227 KB");

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(addr), "r"(cta_rank));
  return out;
}
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *map, uint32_t leader_bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t *bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)0x3)
      : "memory");
}
template <bool I8>
__device__ __forceinline__ void umma_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  if (I8) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(
            tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(
            tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// work item -> (row tile of 256 rows, column tile of N2 columns), r = 256 / N2.  Column tile c is live
// for row tile ti when it starts at or left of the row tile's last row: c < min(ncol, r * (ti + 1));
// the cap only ever binds for the last row tile.  Bands of BAND2 row tiles, column-major in a band,
// so the clusters running at any moment share a handful of operand blocks in L2.
__host__ __device__ __forceinline__ int64_t tile2_prefix(int b0, int r) {
  return (int64_t)r * b0 * (b0 + 1) / 2;  // work items of row tiles [0, b0)
}
__device__ __forceinline__ void tile2_of(int64_t t, int T, int r, int ncol, int &ti, int &tj) {
  int b0 = (int)((sqrt(8.0 * (double)t / r + 1.0) - 1.0) * 0.5);
  b0 = min(b0, T - 1) / BAND2 * BAND2;
  while (b0 > 0 && tile2_prefix(b0, r) > t) b0 -= BAND2;
  while (b0 + BAND2 < T && tile2_prefix(b0 + BAND2, r) <= t) b0 += BAND2;
  t -= tile2_prefix(b0, r);
  const int hb = min(BAND2, T - b0);
  const int cfull = min(ncol, r * (b0 + 1));  // columns live for every row tile of the band
  if (t < (int64_t)cfull * hb) {
    tj = (int)(t / hb);
    ti = b0 + (int)(t % hb);
    return;
  }
  t -= (int64_t)cfull * hb;
  const int cend = min(ncol, r * (b0 + hb));
  for (int c = cfull; c < cend; ++c) {
    const int first = c / r;  // first row tile for which column c is live
    const int cnt = b0 + hb - first;
    if (t < cnt) { tj = c; ti = first + (int)t; return; }
    t -= cnt;
  }
  ti = T - 1; tj = 0;  // not reached for t < tile2_count
}
static int64_t tile2_count(int T, int r, int ncol) {
  int64_t cnt = 0;
  for (int ti = 0; ti < T; ++ti) {
    const int c = r * (ti + 1);
    cnt += c < ncol ? c : ncol;
  }
  return cnt;
}

template <bool I8, typename TO, int PW>
__global__ void __launch_bounds__(GEMM_THREADS, 1) whd_gemm2_kernel(const __grid_constant__ Gemm2Params p,
                                                                    const int64_t nwork) {
  using P = Prog<I8>;
  using SM = Gemm2Smem<TO, PW>;
  constexpr int NSLOT = SM::NSLOT;
  constexpr int N2 = I8 ? 256 : 128;
  constexpr int NACC = I8 ? 2 : 1;
  constexpr int ACCW = I8 ? 256 : 3 * 128;  // TMEM columns per accumulator buffer
  constexpr uint32_t TMEM_COLS = 512u;
  constexpr int BROWS = N2 / 2;             // rows of the column-block operand staged by one CTA
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SM &sm = *reinterpret_cast<SM *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_rank();
  const bool leader = crank == 0;
  const int64_t wstart = blockIdx.x >> 1, wstride = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.tfull[s], 1); mbar_init(&sm.tempty[s], 512); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < (I8 ? 2 : 4); ++i) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.maps[i]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.maps_half[i]) : "memory");
    }
  }
  constexpr bool CAN_TAB = I8 && sizeof(TO) == 4 && PW == 1;
  const int mp1 = p.mchars + 1;
  const bool use_tab = CAN_TAB && p.mchars > 0;
  if (use_tab) {
    const int total = (2 * p.mchars + 1) * mp1;
    for (int x = threadIdx.x; x < total; x += GEMM_THREADS) {
      const int num = x / mp1, den = x - num * mp1;
      sm.qtab[x] = den > 0 ? __fdiv_rn((float)num, (float)den) : 0.f;
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&sm.tmem_base)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own rows of the I operand, own half of the J operand =====
    if (lane == 0) {
      uint32_t lc = 0;
      const int elems_per_chunk = I8 ? CHUNK_BYTES : CHUNK_BYTES / 2;
      for (int64_t work = wstart; work < nwork; work += wstride) {
        int ti, tj;
        tile2_of(work, p.nrow_tiles, 256 / N2, p.ncol_tiles, ti, tj);
        const int row0 = ti * 256 + (int)crank * 128, col0 = tj * N2 + (int)crank * BROWS;
        for (int c = 0; c < p.chunks; ++c) {
#pragma unroll
          for (int l = 0; l < P::NL; ++l, ++lc) {
            const uint32_t slot = lc % NSLOT, phase = (lc / NSLOT) & 1u;
            mbar_wait(&sm.empty[slot], phase ^ 1u);
            const bool jside = P::load_side(l) != 0;
            const uint32_t my_bytes = jside ? (uint32_t)(BROWS * CHUNK_BYTES) : (uint32_t)TILE_BYTES;
            // all bytes of both CTAs are accounted on the leader's barrier
            if (leader) mbar_expect_tx(&sm.full[slot], 2 * my_bytes);
            const uint32_t leader_bar = mapa_u32(smem_u32(&sm.full[slot]), 0);
            const CUtensorMap *map = (jside && BROWS == 64) ? &p.maps_half[P::load_map(l)] : &p.maps[P::load_map(l)];
            tma_load_2d_2sm(sm.tile[slot], map, leader_bar, c * elems_per_chunk, jside ? col0 : row0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores =====
    if (leader && lane == 0) {
      const uint32_t idesc = ((I8 ? 2u : 1u) << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N2 >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      uint32_t lc0 = 0;
      int64_t it = 0;
      for (int64_t work = wstart; work < nwork; work += wstride, ++it) {
        const int as = (int)(it % NACC);
        const uint32_t aphase = (uint32_t)((it / NACC) & 1);
        mbar_wait(&sm.tempty[as], aphase ^ 1);  // both CTAs' epilogues have drained this buffer
        tcgen05_fence_after();
        for (int c = 0; c < p.chunks; ++c, lc0 += P::NL) {
#pragma unroll
          for (int m = 0; m < P::NM; ++m) {
            const uint32_t la = lc0 + P::mma_a(m), lb = lc0 + P::mma_b(m);
            const uint32_t sa = la % NSLOT, sb = lb % NSLOT;
            mbar_wait(&sm.full[sa], (la / NSLOT) & 1u);
            mbar_wait(&sm.full[sb], (lb / NSLOT) & 1u);
            tcgen05_fence_after();
            const uint64_t adesc = make_smem_desc(smem_u32(sm.tile[sa]));
            const uint64_t bdesc = make_smem_desc(smem_u32(sm.tile[sb]));
            const uint32_t tmem_d = tmem_base + (uint32_t)(as * ACCW + P::mma_acc(m) * 128);
            const bool first = (c == 0) && (m < (I8 ? 1 : 3));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_2sm<I8>(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                           (first && kk == 0) ? 0u : 1u);
#pragma unroll
            for (int l = 0; l < P::NL; ++l)
              if ((P::release(m) >> l) & 1u) tcgen05_commit_2sm(&sm.empty[(lc0 + l) % NSLOT]);
          }
        }
        tcgen05_commit_2sm(&sm.tfull[as]);
      }
    }
  } else {
    // ===== epilogue (both CTAs): this CTA's 128 rows x N2 columns =====
    constexpr int HALF = SM::HALF;
    constexpr int WCOLS = N2 / 2;          // columns per epilogue warp
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int lrow = q * 32 + lane;
    const int et = threadIdx.x - 64;       // 0..255
    TO *__restrict__ D = static_cast<TO *>(p.D);
    TO(*stage)[HALF + 1] = sm.stage[warp - 2];
    const uint32_t leader_tempty0 = mapa_u32(smem_u32(&sm.tempty[0]), 0);
    const uint32_t leader_tempty1 = mapa_u32(smem_u32(&sm.tempty[1]), 0);
    int64_t it = 0;
    for (int64_t work = wstart; work < nwork; work += wstride, ++it) {
      int ti, tj;
      tile2_of(work, p.nrow_tiles, 256 / N2, p.ncol_tiles, ti, tj);
      const int row0 = ti * 256 + (int)crank * 128, col0 = tj * N2;
      const int64_t gi = (int64_t)row0 + lrow;
      // strictly below the diagonal and fully inside the map: no per-element tests
      const bool whole = (int64_t)col0 + N2 <= row0 && (int64_t)row0 + 128 <= p.n;
      unsigned long long mrow[PW];
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int w = 0; w < PW; ++w) mrow[w] = gi < p.n ? p.pmask[gi * PW + w] : 0ull;
      if (et < N2) {
        const int64_t gj = (int64_t)col0 + et;
#pragma unroll
        for (int w = 0; w < PW; ++w) sm.pm[et][w] = gj < p.n ? p.pmask[gj * PW + w] : 0ull;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int as = (int)(it % NACC);
      mbar_wait(&sm.tfull[as], (uint32_t)((it / NACC) & 1));
      tcgen05_fence_after();
      for (int part = 0; part < WCOLS; part += HALF) {
        const int cbase = ch * WCOLS + part;
        for (int cb = cbase; cb < cbase + HALF; cb += 16) {
          uint32_t v0[16], v1[16], v2[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACCW + cb);
          tmem_ld16(taddr, v0);
          if (!I8) {
            tmem_ld16(taddr + 128, v1);
            tmem_ld16(taddr + 256, v2);
          }
          tmem_ld_wait();
          TO *mir = D + ((int64_t)col0 + cb) * p.ld + gi;
          // position of the diagonal inside this 16-column block (or outside [0, 16))
          const int64_t dd64 = gi - ((int64_t)col0 + cb);
          const int dd = (dd64 >= 0 && dd64 < 16) ? (int)dd64 : -1;
          const int below = gi < p.n ? (int)min((int64_t)16, max((int64_t)0, dd64)) : 0;  // columns e < below: gj < gi
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            int den = 0;
#pragma unroll
            for (int w = 0; w < PW; ++w) den += __popcll(mrow[w] & sm.pm[cb + e][w]);
            TO d = (TO)0;
            if (CAN_TAB && use_tab) {
              // correctly rounded num / den from the table (den == 0 -> 0); num <= 2 m by construction
              d = (TO)sm.qtab[(int)v0[e] * mp1 + den];
            } else if (den > 0) {
              if (I8) {
                if (sizeof(TO) == 8) d = (TO)__ddiv_rn((double)(int)v0[e], (double)den);
                else d = (TO)__fdiv_rn((float)(int)v0[e], (float)den);
              } else {
                const double num = ((double)__uint_as_float(v0[e]) + (double)__uint_as_float(v1[e])) +
                                   (double)__uint_as_float(v2[e]);
                if (sizeof(TO) == 8) d = (TO)__ddiv_rn(num, (double)den);
                else d = (TO)__double2float_rn(__ddiv_rn(num, (double)den));
              }
            }
            if (e == dd) d = (TO)0;
            stage[lane][(cb - cbase) + e] = d;
            // every unordered pair is produced once: only elements on or below the diagonal are used
            if (whole || e < below) mir[(int64_t)e * p.ld] = d;
          }
        }
        if (part + HALF == WCOLS) {
          tcgen05_fence_before();
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                           as == 0 ? leader_tempty0 : leader_tempty1)
                       : "memory");
        }
        __syncwarp();
        TO *dst = D + ((int64_t)row0 + q * 32) * p.ld + col0 + cbase;
        for (int rr = 0; rr < 32; ++rr, dst += p.ld) {
          const int64_t grow = (int64_t)row0 + q * 32 + rr;
          if (!whole && grow >= p.n) break;
          // direct rows: columns up to and including the diagonal
          if (whole || (int64_t)col0 + cbase + lane <= grow) dst[lane] = stage[rr][lane];
        }
        __syncwarp();
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// ---- operand encoding (K1) -----------------------------------------------------------------
// one thread per (cell, character); buffers are zero-filled beforehand
__global__ void encode_i8_kernel(const int32_t *__restrict__ cm, int64_t n, int m, int S, int32_t missing,
                                 int8_t *__restrict__ A, int8_t *__restrict__ B, int64_t Kp,
                                 int *__restrict__ err_flag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int64_t i = idx / m;
  const int c = (int)(idx - i * m);
  const int32_t s = cm[idx];
  if (s == missing) return;
  if (s < 0 || s > S) { atomicExch(err_flag, 1); return; }
  int8_t *a = A + i * Kp, *b = B + i * Kp;
  const int cut = s != 0;
  a[c] = (int8_t)cut;          b[c] = 1;              // U . P
  a[m + c] = 1;                b[m + c] = (int8_t)cut;  // P . U
  if (cut) {
    const int64_t k = 2 * (int64_t)m + (int64_t)c * S + (s - 1);
    a[k] = -2;
    b[k] = 1;                                          // -2 X . X
  }
}

// presence bit masks: one thread per (cell, 64-character word)
__global__ void encode_mask_kernel(const int32_t *__restrict__ cm, int64_t n, int m, int pw, int32_t missing,
                                   unsigned long long *__restrict__ pmask) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * pw) return;
  const int64_t i = idx / pw;
  const int w = (int)(idx - i * pw);
  unsigned long long bits = 0;
  for (int b = 0; b < 64; ++b) {
    const int c = w * 64 + b;
    if (c < m && cm[i * m + c] != missing) bits |= 1ull << b;
  }
  pmask[idx] = bits;
}

__global__ void encode_bf16_kernel(const int32_t *__restrict__ cm, int64_t n, int m, int S, int32_t missing,
                                   const double *__restrict__ w, __nv_bfloat16 *__restrict__ XW0,
                                   __nv_bfloat16 *__restrict__ XW1, __nv_bfloat16 *__restrict__ XW2,
                                   __nv_bfloat16 *__restrict__ Y, int64_t Kx, int *__restrict__ err_flag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int64_t i = idx / m;
  const int c = (int)(idx - i * m);
  const int32_t s = cm[idx];
  if (s == missing) return;
  if (s < 0 || s > S) { atomicExch(err_flag, 1); return; }
  __nv_bfloat16 *y = Y + i * Kx + (int64_t)c * S;
  const __nv_bfloat16 one = __float2bfloat16(1.0f);
  for (int t = 0; t < S; ++t)
    if (t != s - 1) y[t] = one;  // present and state != t+1
  if (s != 0) {
    const double wt = w[(int64_t)c * (S + 1) + s];
    const __nv_bfloat16 h0 = __double2bfloat16(wt);
    const double r1 = wt - (double)__bfloat162float(h0);
    const __nv_bfloat16 h1 = __double2bfloat16(r1);
    const double r2 = r1 - (double)__bfloat162float(h1);
    const __nv_bfloat16 h2 = __double2bfloat16(r2);
    const int64_t k = i * Kx + (int64_t)c * S + (s - 1);
    XW0[k] = h0;
    XW1[k] = h1;
    XW2[k] = h2;
  }
}

// ---- fixed-point weighted path: scale and operand encoding ----------------------------------
// scale[0] = 2^-F, scale[1] = 2^F with F the largest integer such that round(wmax 2^F) < 2^27 (one bit of slack
// below the 28 bits the four 7-bit digits hold).  One CTA.
__global__ void wscale_kernel(const double *__restrict__ w, int64_t count, double *__restrict__ scale,
                              int *__restrict__ err_flag) {
  __shared__ double sm[32];
  double mx = 0.0;
  for (int64_t e = threadIdx.x; e < count; e += blockDim.x) {
    const double x = w[e];
    if (!(x >= 0.0) || !isfinite(x)) atomicExch(err_flag, 2);  // negative / NaN / inf weight
    else mx = fmax(mx, x);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, sm[i]);
    int e = 0;
    frexp(mx > 0.0 ? mx : 1.0, &e);  // mx < 2^e
    const int F = 27 - e;
    scale[0] = ldexp(1.0, -F);
    scale[1] = ldexp(1.0, F);
  }
}

// one thread per (cell, character); buffers are zero-filled beforehand.  Xp / Up: four digit planes.
__global__ void encode_w8_kernel(const int32_t *__restrict__ cm, int64_t n, int m, int S, int32_t missing,
                                 const double *__restrict__ w, const double *__restrict__ scale,
                                 int8_t *__restrict__ Xp, int64_t plane_x, int8_t *__restrict__ Xm2,
                                 int8_t *__restrict__ Up, int64_t plane_u, int8_t *__restrict__ Pm, int64_t Kx,
                                 int64_t Km, int *__restrict__ err_flag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int64_t i = idx / m;
  const int c = (int)(idx - i * m);
  const int32_t s = cm[idx];
  if (s == missing) return;
  if (s < 0 || s > S) { atomicExch(err_flag, 1); return; }
  Pm[i * Km + c] = 1;
  if (s != 0) {
    const double wt = w[(int64_t)c * (S + 1) + s];
    const long long W = (wt >= 0.0 && isfinite(wt)) ? llrint(wt * scale[1]) : 0ll;
    const int64_t k = i * Kx + (int64_t)c * S + (s - 1);
    Xm2[k] = -2;
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) {
      const int8_t digit = (int8_t)((W >> (7 * pl)) & 127);
      Xp[pl * plane_x + k] = digit;
      Up[pl * plane_u + i * Km + c] = digit;
    }
  }
}

// ---- host side -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// [rows, kcols] row-major matrix of `esz`-byte elements, box = 128 bytes x 128 rows, 128B swizzle
static int make_map(CUtensorMap *map, void *base, int64_t rows, int64_t kcols, int esz, int box_rows = GT) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return CAS_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(kcols * esz)};
  cuuint32_t box[2] = {(cuuint32_t)(CHUNK_BYTES / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult rc = fn(map, esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base,
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with %d", (int)rc); return CAS_ERR_CUDA; }
  return CAS_OK;
}

static inline int64_t round_up64(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct GemmLayout {
  int64_t Kp;   // int8: padded K of A / B
  int64_t Kx;   // bf16: padded m*S
  int64_t Kd;   // bf16: padded m (denominator operand)
};

static GemmLayout layout_of(int m, int S) {
  GemmLayout L;
  L.Kp = round_up64((int64_t)m * (S + 2), CHUNK_BYTES);
  L.Kx = round_up64((int64_t)m * S, CHUNK_BYTES / 2);
  L.Kd = round_up64(m, CHUNK_BYTES / 2);
  return L;
}

// CAS_GEMM_CTA_GROUP=1 forces the single-SM 128 x 128 kernel, =2 the SM-pair kernel.  Default: the
// SM-pair kernel for int8 (operand-feed bound: 1.2-1.7x faster, 75 % tensor-pipe activity at K = 20k)
// and the single-SM kernel for bf16 (bound by operand latency, not bandwidth; measured 7 % faster).
// weighted maps: "i8" (default) = four 7-bit digit planes on the int8 path, "bf16" = three bf16 planes
static bool weighted_fixed_point() {
  const char *e = getenv("CAS_GEMM_WEIGHTED");
  return !(e && e[0] == 'b');
}

static int gemm_cta_group(bool weighted) {
  const char *e = getenv("CAS_GEMM_CTA_GROUP");
  if (e && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
  return weighted ? 1 : 2;
}

size_t whd_gemm_workspace_bytes(int64_t n, int32_t m, int32_t n_states, int32_t weighted) {
  const int S = n_states > 0 ? n_states : 1;
  GemmLayout L = layout_of(m, S);
  Carver c(nullptr);
  c.take<int>(64);
  c.take<unsigned long long>((size_t)n * 8);
  if (!weighted) {
    c.take<int8_t>((size_t)n * L.Kp);
    c.take<int8_t>((size_t)n * L.Kp);
  } else {
    // the larger of the two weighted layouts (the choice is an environment switch read at call time)
    const size_t bf16_bytes = 4 * align_up((size_t)n * L.Kx * 2, 256);
    const int64_t Kx8 = round_up64((int64_t)m * S, CHUNK_BYTES), Km = round_up64(m, CHUNK_BYTES);
    const size_t w8_bytes = 256 + 5 * align_up((size_t)n * Kx8, 256) + 5 * align_up((size_t)n * Km, 256);
    c.take<char>(bf16_bytes > w8_bytes ? bf16_bytes : w8_bytes);
  }
  return c.used();
}

// world > 1: D is rank `rank`'s shard of the block-cyclic layout (sharded.cu: 64-row blocks r, r + G, ...), filled
// by the single-SM kernel in shard mode (every column of the rank's rows, no mirroring); world = 1: the whole map.
int whd_gemm_map(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w, int32_t n_states,
                 void *D, int64_t ld, int32_t dtype, int64_t row0, int64_t row1, void *workspace,
                 size_t workspace_bytes, cudaStream_t stream, int world, int rank) {
  const bool weighted = (w != nullptr);
  const int S = n_states > 0 ? n_states : 1;
  const bool shard = world > 1;
  if (!(row0 == 0 && row1 == n)) {
    set_error("CAS_MAP_TENSOR computes the whole map or a block-cyclic shard (use CAS_MAP_EXACT for row blocks)");
    return CAS_ERR_UNSUPPORTED;
  }
  CAS_REQUIRE(!shard || (rank >= 0 && rank < world), "bad rank %d of %d", rank, world);
  CAS_REQUIRE(m <= 64 * MAXPW, "CAS_MAP_TENSOR supports at most %d characters (m=%d)", 64 * MAXPW, m);
  CAS_REQUIRE(workspace_bytes >= whd_gemm_workspace_bytes(n, m, n_states, weighted), "workspace too small");
  GemmLayout L = layout_of(m, S);
  Carver carve(workspace);
  int *err_flag = carve.take<int>(64);
  CAS_CUDA_TRY(cudaMemsetAsync(err_flag, 0, sizeof(int), stream));

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.D = D; p.ld = ld;
  void *map_bases[4] = {nullptr, nullptr, nullptr, nullptr};
  const int64_t nm = n * (int64_t)m;
  const unsigned eb = (unsigned)((nm + 255) / 256);
  int pw = (m + 63) / 64;
  pw = pw <= 2 ? pw : (pw <= 4 ? 4 : 8);  // kernel templates: 1, 2, 4 or 8 words per cell
  unsigned long long *pmask = carve.take<unsigned long long>((size_t)n * pw);
  encode_mask_kernel<<<(unsigned)((n * pw + 255) / 256), 256, 0, stream>>>(cm, n, m, pw, missing, pmask);
  CAS_LAUNCH_CHECK();
  p.pw = pw; p.pmask = pmask;
  p.ntile_rows = (int)((n + GT - 1) / GT);
  if (shard) {
    // rows of this rank: 64-row blocks r, r + G, ... below n (the last block may be partial)
    const int64_t nblocks = (n + 63) / 64;
    int64_t rows = 0;
    for (int64_t b = rank; b < nblocks; b += world) rows += (b * 64 + 64 <= n) ? 64 : (n - b * 64);
    p.shard_world = world; p.shard_rank = rank;
    p.nloc_rows = (int)rows;
    p.nloc_tiles = (int)((rows + GT - 1) / GT);
    p.ncol_tiles = (int)((n + GT - 1) / GT);
    if (rows == 0) return CAS_OK;  // fewer 64-row blocks than ranks: this rank holds nothing
  }
  // shard mode loads the row-side operand as two 64-row boxes
  auto half_map = [&](int i, void *base, int64_t kcols, int esz) -> int {
    return shard ? make_map(&p.maps_half[i], base, n, kcols, esz, 64) : CAS_OK;
  };
  if (!weighted) {
    int8_t *A = carve.take<int8_t>((size_t)n * L.Kp);
    int8_t *B = carve.take<int8_t>((size_t)n * L.Kp);
    CAS_CUDA_TRY(cudaMemsetAsync(A, 0, (size_t)n * L.Kp, stream));
    CAS_CUDA_TRY(cudaMemsetAsync(B, 0, (size_t)n * L.Kp, stream));
    encode_i8_kernel<<<eb, 256, 0, stream>>>(cm, n, m, S, missing, A, B, L.Kp, err_flag);
    CAS_LAUNCH_CHECK();
    int rc = make_map(&p.maps[0], A, n, L.Kp, 1);
    if (rc != CAS_OK) return rc;
    rc = make_map(&p.maps[1], B, n, L.Kp, 1);
    if (rc != CAS_OK) return rc;
    rc = half_map(0, A, L.Kp, 1);
    if (rc != CAS_OK) return rc;
    rc = half_map(1, B, L.Kp, 1);
    if (rc != CAS_OK) return rc;
    map_bases[0] = A; map_bases[1] = B;
    p.chunks = (int)(L.Kp / CHUNK_BYTES);
  } else if (weighted_fixed_point()) {
    const int64_t Kx8 = round_up64((int64_t)m * S, CHUNK_BYTES), Km = round_up64(m, CHUNK_BYTES);
    double *scale = carve.take<double>(32);
    int8_t *Xp = carve.take<int8_t>(4 * (size_t)n * Kx8);
    int8_t *Xm2 = carve.take<int8_t>((size_t)n * Kx8);
    int8_t *Up = carve.take<int8_t>(4 * (size_t)n * Km);
    int8_t *Pm = carve.take<int8_t>((size_t)n * Km);
    CAS_CUDA_TRY(cudaMemsetAsync(Xp, 0, 4 * (size_t)n * Kx8, stream));
    CAS_CUDA_TRY(cudaMemsetAsync(Xm2, 0, (size_t)n * Kx8, stream));
    CAS_CUDA_TRY(cudaMemsetAsync(Up, 0, 4 * (size_t)n * Km, stream));
    CAS_CUDA_TRY(cudaMemsetAsync(Pm, 0, (size_t)n * Km, stream));
    wscale_kernel<<<1, 1024, 0, stream>>>(w, (int64_t)m * (S + 1), scale, err_flag);
    CAS_LAUNCH_CHECK();
    encode_w8_kernel<<<eb, 256, 0, stream>>>(cm, n, m, S, missing, w, scale, Xp, (int64_t)n * Kx8, Xm2, Up,
                                             (int64_t)n * Km, Pm, Kx8, Km, err_flag);
    CAS_LAUNCH_CHECK();
    for (int i = 0; i < 4; ++i) {
      int rc = make_map(&p.maps[i], Xp + (size_t)i * n * Kx8, n, Kx8, 1);
      if (rc != CAS_OK) return rc;
      rc = make_map(&p.maps[5 + i], Up + (size_t)i * n * Km, n, Km, 1);
      if (rc != CAS_OK) return rc;
      rc = half_map(i, Xp + (size_t)i * n * Kx8, Kx8, 1);
      if (rc != CAS_OK) return rc;
      rc = half_map(5 + i, Up + (size_t)i * n * Km, Km, 1);
      if (rc != CAS_OK) return rc;
    }
    int rc = make_map(&p.maps[4], Xm2, n, Kx8, 1);
    if (rc != CAS_OK) return rc;
    rc = make_map(&p.maps[9], Pm, n, Km, 1);
    if (rc != CAS_OK) return rc;
    rc = half_map(4, Xm2, Kx8, 1);
    if (rc != CAS_OK) return rc;
    rc = half_map(9, Pm, Km, 1);
    if (rc != CAS_OK) return rc;
    p.head_chunks = (int)(Km / CHUNK_BYTES);
    p.chunks = 2 * p.head_chunks + (int)(Kx8 / CHUNK_BYTES);
    p.scale = scale;
    const int64_t tiles8 = (n + GT - 1) / GT;
    const int64_t ntiles8 = shard ? (int64_t)p.nloc_tiles * p.ncol_tiles : tiles8 * (tiles8 + 1) / 2;
    CAS_REQUIRE(ntiles8 < 2147483647LL, "too many tiles");
#define LAUNCH_W8(TO, PWV)                                                                         \
  do {                                                                                             \
    const size_t smem = sizeof(GemmSmem<TO>) + 1024;                                               \
    CAS_CUDA_TRY(cudaFuncSetAttribute(whd_gemm_kernel<2, TO, PWV>,                                 \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    const int64_t grid = ntiles8 < sm_count() ? ntiles8 : sm_count();                              \
    whd_gemm_kernel<2, TO, PWV><<<(unsigned)grid, GEMM_THREADS, smem, stream>>>(p);                \
  } while (0)
#define LAUNCH_W8_PW(TO)                                                                           \
  do {                                                                                             \
    if (pw == 1) LAUNCH_W8(TO, 1);                                                                 \
    else if (pw == 2) LAUNCH_W8(TO, 2);                                                            \
    else if (pw <= 4) LAUNCH_W8(TO, 4);                                                            \
    else LAUNCH_W8(TO, 8);                                                                         \
  } while (0)
    if (dtype == CAS_F64) LAUNCH_W8_PW(double); else LAUNCH_W8_PW(float);
#undef LAUNCH_W8_PW
#undef LAUNCH_W8
    CAS_LAUNCH_CHECK();
    return CAS_OK;
  } else {
    __nv_bfloat16 *XW[3];
    for (int i = 0; i < 3; ++i) XW[i] = carve.take<__nv_bfloat16>((size_t)n * L.Kx);
    __nv_bfloat16 *Y = carve.take<__nv_bfloat16>((size_t)n * L.Kx);
    for (int i = 0; i < 3; ++i) CAS_CUDA_TRY(cudaMemsetAsync(XW[i], 0, (size_t)n * L.Kx * 2, stream));
    CAS_CUDA_TRY(cudaMemsetAsync(Y, 0, (size_t)n * L.Kx * 2, stream));
    encode_bf16_kernel<<<eb, 256, 0, stream>>>(cm, n, m, S, missing, w, XW[0], XW[1], XW[2], Y, L.Kx, err_flag);
    CAS_LAUNCH_CHECK();
    for (int i = 0; i < 3; ++i) {
      int rc = make_map(&p.maps[i], XW[i], n, L.Kx, 2);
      if (rc != CAS_OK) return rc;
    }
    int rc = make_map(&p.maps[3], Y, n, L.Kx, 2);
    if (rc != CAS_OK) return rc;
    for (int i = 0; i < 3; ++i) {
      rc = half_map(i, XW[i], L.Kx, 2);
      if (rc != CAS_OK) return rc;
    }
    rc = half_map(3, Y, L.Kx, 2);
    if (rc != CAS_OK) return rc;
    for (int i = 0; i < 3; ++i) map_bases[i] = XW[i];
    map_bases[3] = Y;
    p.chunks = (int)(L.Kx / (CHUNK_BYTES / 2));
  }
  if (!shard && gemm_cta_group(weighted) == 2) {
    // 2-SM kernel: 256-row tiles, N2 = 256 (int8) / 128 (bf16) columns, clusters of two CTAs
    Gemm2Params q;
    memset(&q, 0, sizeof(q));
    q.n = n; q.D = D; q.ld = ld; q.pw = pw; q.pmask = pmask; q.chunks = p.chunks;
    // quotient table of the unweighted float32 epilogue, when it fits (m <= 60)
    q.mchars = (!weighted && dtype == CAS_F32 && pw == 1 && (2 * (int64_t)m + 1) * ((int64_t)m + 1) <= 7680) ? m : 0;
    const int nmaps = weighted ? 4 : 2;
    const int esz = weighted ? 2 : 1;
    const int64_t kcols = weighted ? L.Kx : L.Kp;
    const int N2 = weighted ? 128 : 256;
    for (int i = 0; i < nmaps; ++i) {
      q.maps[i] = p.maps[i];
      void *base = map_bases[i];
      int rc = make_map(&q.maps_half[i], base, n, kcols, esz, 64);
      if (rc != CAS_OK) return rc;
    }
    q.nrow_tiles = (int)((n + 255) / 256);
    q.ncol_tiles = (int)((n + N2 - 1) / N2);
    const int64_t nwork = tile2_count(q.nrow_tiles, 256 / N2, q.ncol_tiles);
#define LAUNCH_GEMM2(I8, TO, PWV)                                                                  \
  do {                                                                                             \
    const size_t smem = sizeof(Gemm2Smem<TO, PWV>) + 1024;                                         \
    CAS_CUDA_TRY(cudaFuncSetAttribute(whd_gemm2_kernel<I8, TO, PWV>,                               \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    cudaLaunchConfig_t cfg;                                                                        \
    memset(&cfg, 0, sizeof(cfg));                                                                  \
    cudaLaunchAttribute attr[1];                                                                   \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                              \
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;      \
    cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;           \
    cfg.attrs = attr; cfg.numAttrs = 1;                                                            \
    cfg.gridDim = dim3(2);                                                                         \
    int nclusters = 0;                                                                             \
    if (cudaOccupancyMaxActiveClusters(&nclusters, whd_gemm2_kernel<I8, TO, PWV>, &cfg) !=         \
            cudaSuccess || nclusters <= 0) {                                                       \
      (void)cudaGetLastError();                                                                    \
      nclusters = sm_count() / 2;                                                                  \
    }                                                                                              \
    const int64_t g = nwork < nclusters ? nwork : nclusters;                                       \
    cfg.gridDim = dim3((unsigned)(2 * g));                                                         \
    CAS_CUDA_TRY(cudaLaunchKernelEx(&cfg, whd_gemm2_kernel<I8, TO, PWV>, q, nwork));               \
  } while (0)
#define LAUNCH_PW2(I8, TO)                                                                         \
  do {                                                                                             \
    if (pw == 1) LAUNCH_GEMM2(I8, TO, 1);                                                          \
    else if (pw == 2) LAUNCH_GEMM2(I8, TO, 2);                                                     \
    else if (pw <= 4) LAUNCH_GEMM2(I8, TO, 4);                                                     \
    else LAUNCH_GEMM2(I8, TO, 8);                                                                  \
  } while (0)
    if (!weighted) {
      if (dtype == CAS_F64) LAUNCH_PW2(true, double); else LAUNCH_PW2(true, float);
    } else {
      if (dtype == CAS_F64) LAUNCH_PW2(false, double); else LAUNCH_PW2(false, float);
    }
#undef LAUNCH_PW2
#undef LAUNCH_GEMM2
    CAS_LAUNCH_CHECK();
    return CAS_OK;
  }
  const int64_t tiles = (n + GT - 1) / GT;
  const int64_t ntiles = shard ? (int64_t)p.nloc_tiles * p.ncol_tiles : tiles * (tiles + 1) / 2;
  CAS_REQUIRE(ntiles < 2147483647LL, "too many tiles");
#define LAUNCH_GEMM(I8, TO, PWV)                                                                   \
  do {                                                                                             \
    const size_t smem = sizeof(GemmSmem<TO>) + 1024;                                               \
    CAS_CUDA_TRY(cudaFuncSetAttribute(whd_gemm_kernel<I8, TO, PWV>,                                \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    const int64_t grid = ntiles < sm_count() ? ntiles : sm_count();                                \
    whd_gemm_kernel<I8, TO, PWV><<<(unsigned)grid, GEMM_THREADS, smem, stream>>>(p);               \
  } while (0)
#define LAUNCH_PW(I8, TO)                                                                          \
  do {                                                                                             \
    if (pw == 1) LAUNCH_GEMM(I8, TO, 1);                                                           \
    else if (pw == 2) LAUNCH_GEMM(I8, TO, 2);                                                      \
    else if (pw <= 4) LAUNCH_GEMM(I8, TO, 4);                                                      \
    else LAUNCH_GEMM(I8, TO, 8);                                                                   \
  } while (0)
  if (!weighted) {
    if (dtype == CAS_F64) LAUNCH_PW(true, double); else LAUNCH_PW(true, float);
  } else {
    if (dtype == CAS_F64) LAUNCH_PW(false, double); else LAUNCH_PW(false, float);
  }
#undef LAUNCH_PW
#undef LAUNCH_GEMM
  CAS_LAUNCH_CHECK();
  return CAS_OK;
}

// ---- measured peak of the int8 tensor pipe ---------------------------------------------------
// One CTA per SM issues `iters` x 4 back-to-back tcgen05.mma kind::i8 (M = 128, N = 256, K = 32) on operand tiles that
// already sit in shared memory (zero-filled: the pipe's rate does not depend on the data), accumulating in TMEM.  No
// loads, no epilogue: what is timed is the issue rate of the MMA pipe itself, the denominator the map kernels'
// tensor fractions are quoted against (scripts/map_sweep.py).  cuBLASLt's int8 GEMM reached through torch._int_mm
// is far below it on this image (636 TOP/s at 16384^3), so it is no roofline.
__global__ void __launch_bounds__(128) i8_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *base = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t *tileA = base;               // [128 rows][128 bytes]
  uint8_t *tileB = base + 128 * 128;   // [256 rows][128 bytes]
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (128 + 256) * 128 / 16; i += blockDim.x)
    reinterpret_cast<uint4 *>(base)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async-proxy (MMA) reads
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = make_smem_desc(smem_u32(tileA)), bdesc = make_smem_desc(smem_u32(tileB));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma<true>(tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (it | kk) ? 1u : 0u);
    }
    tcgen05_commit(&done_bar);
    mbar_wait(&done_bar, 0);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

int int8_mma_peak(int iters, cudaStream_t stream, double *ops_per_s) {
  const size_t smem = (128 + 256) * 128 + 1024;
  CAS_CUDA_TRY(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = sm_count();
  cudaEvent_t e0, e1;
  CAS_CUDA_TRY(cudaEventCreate(&e0));
  CAS_CUDA_TRY(cudaEventCreate(&e1));
  i8_peak_kernel<<<grid, 128, smem, stream>>>(iters / 8 + 1);  // warm-up
  CAS_CUDA_TRY(cudaEventRecord(e0, stream));
  i8_peak_kernel<<<grid, 128, smem, stream>>>(iters);
  CAS_CUDA_TRY(cudaEventRecord(e1, stream));
  CAS_LAUNCH_CHECK();
  CAS_CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0.f;
  CAS_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ops_per_s = (double)grid * iters * 4.0 * (2.0 * 128 * 256 * 32) / ((double)ms * 1e-3);
  return CAS_OK;
}

int whd_gemm_read_error_flag(const void *workspace, cudaStream_t stream, int *flag) {
  CAS_CUDA_TRY(cudaMemcpyAsync(flag, workspace, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CAS_CUDA_TRY(cudaStreamSynchronize(stream));
  return CAS_OK;
}

}  // namespace cas
