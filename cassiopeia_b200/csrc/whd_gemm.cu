// Tensor-core dissimilarity map: the weighted Hamming distance recast as one-hot contractions
// and run on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged by TMA with the 128-byte swizzle).  Replaces the same reference code as
// whd_exact.cu (cassiopeia/solver/dissimilarity_functions.py:46-72 over all pairs).
//
// Unweighted ("plain Hamming", costs 0/1/2) -- kind::i8, int32 accumulation, EXACT:
//   per cell one row of A and one of B, K = m*(S+3) int8 columns
//     A = [ U | P | -2 X | 64 P ]      B = [ P | U | X | 64 P ]
//   P = present, U = present & cut, X = one-hot over (character, cut state).  Then
//     acc_ij = (U.P + P.U - 2 X.X) + 4096 * (P.P) = numerator + 4096 * shared_present
//   (numerator <= 2m < 4096), so one accumulator yields both; D = num / den in IEEE
//   arithmetic => bit-identical to the reference.
// Weighted (mutation priors) -- kind::f16 (bf16 x bf16 -> fp32), cancellation-free form
//     num_ij = sum_{c,s} Xw_i[c,s] * Y_j[c,s] + Y_i[c,s] * Xw_j[c,s],   Y = present & (state != s)
//   every term is >= 0.  The float64 weight is split into three bf16 planes (hi, mid, lo:
//   24 significant bits); each plane has its own fp32 accumulator in TMEM (a product with the
//   exact 0/1 operand is exact, and a plane's partial sums stay within fp32's 24 bits for
//   realistic weight ranges), the planes are added in float64 in the epilogue, and the
//   denominator P.P is a fourth accumulator.  Relative error vs the exact kernel <= 1e-6
//   (tests/test_gpu_gemm.py).
//
// Kernel: one CTA per 128x128 lower-triangle output tile, 6 warps: warp 0 = TMA producer,
// warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 = epilogue (TMEM -> registers
// -> D, mirrored).  K is walked in 128-byte chunks (4 MMAs each) through a 4-stage
// full/empty mbarrier ring.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace cas {

constexpr int GT = 128;            // tile edge (UMMA M = N = 128)
constexpr int CHUNK_BYTES = 128;   // K bytes per pipeline stage row (one swizzle atom)
constexpr int TILE_BYTES = GT * CHUNK_BYTES;  // 16 KB per operand tile
constexpr int STAGES = 4;
constexpr int GEMM_THREADS = 192;
constexpr int MAX_SEGS = 8;
constexpr int MAX_MAPS = 5;

struct GemmSeg {
  int map_a, map_b;  // tensor-map indices of the A / B operand matrices
  int chunks;        // number of 128-byte K chunks
  int acc;           // accumulator index (TMEM column block of 128)
};

struct GemmParams {
  CUtensorMap maps[MAX_MAPS];
  GemmSeg segs[MAX_SEGS];
  int nsegs;
  int kind_i8;   // 1: kind::i8 (int32 acc)  0: kind::f16 bf16 (fp32 acc)
  int naccs;     // accumulators in use (1 or 4)
  int64_t n;
  void *D;
  int64_t ld;
  int out_f64;
};

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

struct __align__(1024) GemmSmem {
  uint8_t a[STAGES][TILE_BYTES];
  uint8_t b[STAGES][TILE_BYTES];
  uint64_t full[STAGES];
  uint64_t empty[STAGES];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

template <bool I8, typename TO>
__global__ void __launch_bounds__(GEMM_THREADS, 1) whd_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  GemmSmem &sm = *reinterpret_cast<GemmSmem *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // lower-triangle tile (ti >= tj)
  int64_t t = blockIdx.x;
  int64_t r = (int64_t)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  while (r * (r + 1) / 2 > t) --r;
  const int ti = (int)r, tj = (int)(t - r * (r + 1) / 2);
  const int row0 = ti * GT, col0 = tj * GT;
  const uint32_t tmem_cols = p.naccs * GT;  // 128 or 512 (power of two >= 32)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&sm.tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int sg = 0; sg < p.nsegs; ++sg) {
        const GemmSeg seg = p.segs[sg];
        const int elems_per_chunk = I8 ? CHUNK_BYTES : CHUNK_BYTES / 2;
        for (int c = 0; c < seg.chunks; ++c) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_expect_tx(&sm.full[stage], 2 * TILE_BYTES);
          tma_load_2d(sm.a[stage], &p.maps[seg.map_a], &sm.full[stage], c * elems_per_chunk, row0);
          tma_load_2d(sm.b[stage], &p.maps[seg.map_b], &sm.full[stage], c * elems_per_chunk, col0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
      int stage = 0;
      uint32_t phase = 0;
      uint32_t started = 0;  // bit a set once accumulator a holds data
      for (int sg = 0; sg < p.nsegs; ++sg) {
        const GemmSeg seg = p.segs[sg];
        const uint32_t tmem_d = tmem_base + (uint32_t)(seg.acc * GT);
        for (int c = 0; c < seg.chunks; ++c) {
          mbar_wait(&sm.full[stage], phase);
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(sm.a[stage]));
          const uint64_t bdesc = make_smem_desc(smem_u32(sm.b[stage]));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // advance 32 bytes of K inside the swizzle atom: +2 in the (>>4) start-address field
            umma<I8>(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                     ((started >> seg.acc) & 1u) | (kk > 0 ? 1u : 0u));
          }
          started |= 1u << seg.acc;
          tcgen05_commit(&sm.empty[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      tcgen05_commit(&sm.tmem_full);
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4) .. +32 =====
    mbar_wait(&sm.tmem_full, 0);
    tcgen05_fence_after();
    const int q = warp & 3;
    const int lrow = q * 32 + lane;            // row inside the tile == TMEM lane
    const int64_t gi = (int64_t)row0 + lrow;   // global row
    TO *__restrict__ D = static_cast<TO *>(p.D);
    const bool diag_tile = (ti == tj);
    for (int cb = 0; cb < GT; cb += 16) {
      uint32_t v0[16], v1[16], v2[16], v3[16];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb;
      tmem_ld16(taddr, v0);
      if (!I8) {
        tmem_ld16(taddr + GT, v1);
        tmem_ld16(taddr + 2 * GT, v2);
        tmem_ld16(taddr + 3 * GT, v3);
      }
      tmem_ld_wait();
      TO out[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        double num, den;
        if (I8) {
          const int acc = (int)v0[e];
          den = (double)(acc >> 12);
          num = (double)(acc & 4095);
        } else {
          num = ((double)__uint_as_float(v0[e]) + (double)__uint_as_float(v1[e])) + (double)__uint_as_float(v2[e]);
          den = (double)__uint_as_float(v3[e]);
        }
        double d = den > 0.0 ? __ddiv_rn(num, den) : 0.0;
        if (diag_tile && (cb + e == lrow)) d = 0.0;
        out[e] = (TO)d;
      }
      if (gi < p.n) {
        const int64_t gj0 = (int64_t)col0 + cb;
        TO *dst = D + gi * p.ld + gj0;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (gj0 + e < p.n) dst[e] = out[e];
        if (!diag_tile) {
          // mirrored element (gj, gi): the 32 lanes of a warp write 32 consecutive columns
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (gj0 + e < p.n) D[(gj0 + e) * p.ld + gi] = out[e];
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
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
  const int64_t kd = 2 * (int64_t)m + (int64_t)m * S + c;
  a[kd] = 64;
  b[kd] = 64;                                          // 4096 * P . P
}

__global__ void encode_bf16_kernel(const int32_t *__restrict__ cm, int64_t n, int m, int S, int32_t missing,
                                   const double *__restrict__ w, __nv_bfloat16 *__restrict__ XW0,
                                   __nv_bfloat16 *__restrict__ XW1, __nv_bfloat16 *__restrict__ XW2,
                                   __nv_bfloat16 *__restrict__ Y, __nv_bfloat16 *__restrict__ P, int64_t Kx,
                                   int64_t Kd, int *__restrict__ err_flag) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int64_t i = idx / m;
  const int c = (int)(idx - i * m);
  const int32_t s = cm[idx];
  if (s == missing) return;
  if (s < 0 || s > S) { atomicExch(err_flag, 1); return; }
  P[i * Kd + c] = __float2bfloat16(1.0f);
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
static int make_map(CUtensorMap *map, void *base, int64_t rows, int64_t kcols, int esz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return CAS_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(kcols * esz)};
  cuuint32_t box[2] = {(cuuint32_t)(CHUNK_BYTES / esz), (cuuint32_t)GT};
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
  L.Kp = round_up64((int64_t)m * (S + 3), CHUNK_BYTES);
  L.Kx = round_up64((int64_t)m * S, CHUNK_BYTES / 2);
  L.Kd = round_up64(m, CHUNK_BYTES / 2);
  return L;
}

size_t whd_gemm_workspace_bytes(int64_t n, int32_t m, int32_t n_states, int32_t weighted) {
  const int S = n_states > 0 ? n_states : 1;
  GemmLayout L = layout_of(m, S);
  Carver c(nullptr);
  c.take<int>(64);
  if (!weighted) {
    c.take<int8_t>((size_t)n * L.Kp);
    c.take<int8_t>((size_t)n * L.Kp);
  } else {
    for (int i = 0; i < 4; ++i) c.take<__nv_bfloat16>((size_t)n * L.Kx);
    c.take<__nv_bfloat16>((size_t)n * L.Kd);
  }
  return c.used();
}

int whd_gemm_map(const int32_t *cm, int64_t n, int32_t m, int32_t missing, const double *w, int32_t n_states,
                 void *D, int64_t ld, int32_t dtype, int64_t row0, int64_t row1, void *workspace,
                 size_t workspace_bytes, cudaStream_t stream) {
  const bool weighted = (w != nullptr);
  const int S = n_states > 0 ? n_states : 1;
  if (!(row0 == 0 && row1 == n)) {
    set_error("CAS_MAP_TENSOR computes the whole map only (use CAS_MAP_EXACT for row blocks)");
    return CAS_ERR_UNSUPPORTED;
  }
  CAS_REQUIRE(m < 2048, "CAS_MAP_TENSOR supports at most 2047 characters (m=%d)", m);
  CAS_REQUIRE(workspace_bytes >= whd_gemm_workspace_bytes(n, m, n_states, weighted), "workspace too small");
  GemmLayout L = layout_of(m, S);
  Carver carve(workspace);
  int *err_flag = carve.take<int>(64);
  CAS_CUDA_TRY(cudaMemsetAsync(err_flag, 0, sizeof(int), stream));

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.n = n; p.D = D; p.ld = ld; p.out_f64 = (dtype == CAS_F64);
  const int64_t nm = n * (int64_t)m;
  const unsigned eb = (unsigned)((nm + 255) / 256);
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
    p.nsegs = 1;
    p.segs[0] = GemmSeg{0, 1, (int)(L.Kp / CHUNK_BYTES), 0};
    p.kind_i8 = 1;
    p.naccs = 1;
  } else {
    __nv_bfloat16 *XW[3];
    for (int i = 0; i < 3; ++i) XW[i] = carve.take<__nv_bfloat16>((size_t)n * L.Kx);
    __nv_bfloat16 *Y = carve.take<__nv_bfloat16>((size_t)n * L.Kx);
    __nv_bfloat16 *P = carve.take<__nv_bfloat16>((size_t)n * L.Kd);
    for (int i = 0; i < 3; ++i) CAS_CUDA_TRY(cudaMemsetAsync(XW[i], 0, (size_t)n * L.Kx * 2, stream));
    CAS_CUDA_TRY(cudaMemsetAsync(Y, 0, (size_t)n * L.Kx * 2, stream));
    CAS_CUDA_TRY(cudaMemsetAsync(P, 0, (size_t)n * L.Kd * 2, stream));
    encode_bf16_kernel<<<eb, 256, 0, stream>>>(cm, n, m, S, missing, w, XW[0], XW[1], XW[2], Y, P, L.Kx, L.Kd,
                                               err_flag);
    CAS_LAUNCH_CHECK();
    for (int i = 0; i < 3; ++i) {
      int rc = make_map(&p.maps[i], XW[i], n, L.Kx, 2);
      if (rc != CAS_OK) return rc;
    }
    int rc = make_map(&p.maps[3], Y, n, L.Kx, 2);
    if (rc != CAS_OK) return rc;
    rc = make_map(&p.maps[4], P, n, L.Kd, 2);
    if (rc != CAS_OK) return rc;
    const int chunks = (int)(L.Kx / (CHUNK_BYTES / 2));
    int sgi = 0;
    for (int pl = 0; pl < 3; ++pl) {
      p.segs[sgi++] = GemmSeg{pl, 3, chunks, pl};  // Xw_p[I] . Y[J]
      p.segs[sgi++] = GemmSeg{3, pl, chunks, pl};  // Y[I] . Xw_p[J]
    }
    p.segs[sgi++] = GemmSeg{4, 4, (int)(L.Kd / (CHUNK_BYTES / 2)), 3};  // P[I] . P[J]
    p.nsegs = sgi;
    p.kind_i8 = 0;
    p.naccs = 4;
  }
  const int64_t tiles = (n + GT - 1) / GT;
  const int64_t ntiles = tiles * (tiles + 1) / 2;
  CAS_REQUIRE(ntiles < 2147483647LL, "too many tiles");
  const size_t smem = sizeof(GemmSmem) + 1024;
#define LAUNCH_GEMM(I8, TO)                                                                        \
  do {                                                                                             \
    CAS_CUDA_TRY(cudaFuncSetAttribute(whd_gemm_kernel<I8, TO>,                                     \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    whd_gemm_kernel<I8, TO><<<(unsigned)ntiles, GEMM_THREADS, smem, stream>>>(p);                  \
  } while (0)
  if (!weighted) {
    if (dtype == CAS_F64) LAUNCH_GEMM(true, double); else LAUNCH_GEMM(true, float);
  } else {
    if (dtype == CAS_F64) LAUNCH_GEMM(false, double); else LAUNCH_GEMM(false, float);
  }
#undef LAUNCH_GEMM
  CAS_LAUNCH_CHECK();
  return CAS_OK;
}

int whd_gemm_read_error_flag(const void *workspace, cudaStream_t stream, int *flag) {
  CAS_CUDA_TRY(cudaMemcpyAsync(flag, workspace, sizeof(int), cudaMemcpyDeviceToHost, stream));
  CAS_CUDA_TRY(cudaStreamSynchronize(stream));
  return CAS_OK;
}

}  // namespace cas
