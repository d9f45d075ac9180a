/*
 * cas_b200.h -- C ABI of the B200-native distance-solver hot path.
 *
 * Drop-in boundary for YosefLab/Cassiopeia's DistanceSolver path.  Every entry
 * point takes plain pointers and sizes (no torch / Python types) so it can be
 * bound from ctypes (see INTEGRATION.md for the stub a Cassiopeia maintainer
 * would add).  All `*_dev` pointers are device pointers owned by the caller
 * (the Python host keeps them alive as torch tensors); `stream` is a
 * `cudaStream_t` passed as `void*` (NULL = legacy default stream).  Functions
 * enqueue work asynchronously on `stream` unless they say otherwise, return 0
 * on success and a negative code on failure; `cas_last_error()` then holds a
 * human-readable message (thread-local).
 *
 * Reference interfaces replaced (paths relative to the Cassiopeia checkout):
 *   cas_whd_map        <- cassiopeia/data/utilities.py:189-295 `compute_dissimilarity_map`
 *                         with cassiopeia/solver/dissimilarity_functions.py:12-72
 *                         `weighted_hamming_distance` (called from
 *                         cassiopeia/data/CassiopeiaTree.py:1926)
 *   cas_nj_solve       <- the `while N > 2` loop of cassiopeia/solver/DistanceSolver.py:185-205
 *                         specialised by NeighborJoiningSolver.find_cherry / compute_q /
 *                         update_dissimilarity_map (cassiopeia/solver/NeighborJoiningSolver.py:123-260)
 *   cas_upgma_solve    <- same loop with UPGMASolver.find_cherry / update_dissimilarity_map
 *                         (cassiopeia/solver/UPGMASolver.py:118-238)
 *   cas_solve_host     <- DistanceSolver.solve's numeric part end to end
 *                         (cassiopeia/solver/DistanceSolver.py:173-205) from HOST buffers
 *   cas_sharded_*      <- no reference counterpart (the reference is single-host); row-block
 *                         sharded variant of the two loops for n^2 maps larger than one GPU.
 */
#ifndef CAS_B200_H_
#define CAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAS_ABI_VERSION 2

#if defined(__GNUC__)
#define CAS_API __attribute__((visibility("default")))
#else
#define CAS_API
#endif

/* element type of the dissimilarity map */
#define CAS_F32 0
#define CAS_F64 1
/* agglomeration rule */
#define CAS_ALGO_NJ 0
#define CAS_ALGO_UPGMA 1
/* arithmetic mode of the loop:
 *   STRICT: float64 map, NJ row sums re-accumulated every iteration sequentially in
 *           id order, no FMA -- reproduces the reference join for join.
 *   FAST  : float32 (or float64) map, NJ row sums maintained incrementally in float64. */
#define CAS_MODE_STRICT 0
#define CAS_MODE_FAST 1
/*   EXACT : float64 map; NJ row sums maintained incrementally together with rigorous error bounds;
 *           the reference's sequential sums are evaluated only for the rows of pairs that fall
 *           inside the error window of the minimum.  Same join list as STRICT (= the reference's)
 *           at the cost of the FAST loop.  For UPGMA (no sums) identical to STRICT. */
#define CAS_MODE_EXACT 2
/* dissimilarity-map kernel:
 *   EXACT : CUDA-core kernel, float64 accumulation in character order (bit-exact vs reference)
 *   TENSOR: one-hot contraction on tcgen05 tensor cores (int8 unweighted / split-bf16 weighted) */
#define CAS_MAP_EXACT 0
#define CAS_MAP_TENSOR 1

/* error codes */
#define CAS_OK 0
#define CAS_ERR_INVALID (-1)
#define CAS_ERR_CUDA (-2)
#define CAS_ERR_UNSUPPORTED (-3)
#define CAS_ERR_WORKSPACE (-4)
#define CAS_ERR_NCCL (-5)

CAS_API int cas_abi_version(void);
CAS_API const char *cas_last_error(void);

/* sm_count, compute capability and free/total device memory of the current device */
CAS_API int cas_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, size_t *free_bytes,
                    size_t *total_bytes);

/* ---- pairwise weighted-Hamming dissimilarity map -------------------------
 * cm_dev   : [n, m] int32 character matrix, row-major.  States are
 *            0 (uncut), 1..n_states (indels, compact codes) or `missing`.
 * w_dev    : [m, n_states+1] float64 weights (w[c][0] ignored) or NULL for the
 *            reference's unweighted costs (0/1/2).
 * D_dev    : output, element (i, j) at D[(i-row0)*ld + j] for i in [row0,row1),
 *            j in [0,n); dtype CAS_F32 / CAS_F64; diagonal 0.  When the block is
 *            the whole matrix only lower-triangle tiles are computed and mirrored.
 * workspace: cas_whd_workspace_bytes(...) bytes of device scratch.
 */
CAS_API size_t cas_whd_workspace_bytes(int64_t n, int32_t m, int32_t n_states, int32_t weighted,
                               int32_t method);
CAS_API int cas_whd_map(const int32_t *cm_dev, int64_t n, int32_t m, int32_t missing, const double *w_dev,
                int32_t n_states, void *D_dev, int64_t ld, int32_t dtype, int64_t row0,
                int64_t row1, int32_t method, void *workspace_dev, size_t workspace_bytes,
                void *stream);

/* ---- the other pairwise functions of cassiopeia/solver/dissimilarity_functions.py ----------
 * Same layout, workspace (cas_whd_workspace_bytes(..., CAS_MAP_EXACT)) and bit-exact float64
 * accumulation in character order as cas_whd_map(method = CAS_MAP_EXACT); `fn` selects the pair
 * function.  Replaces cassiopeia/data/utilities.py:189-295 `compute_dissimilarity_map` called with
 *   CAS_FN_WEIGHTED_HAMMING_DISTANCE                    dissimilarity_functions.py:12-72
 *   CAS_FN_HAMMING_SIMILARITY_WITHOUT_MISSING           :75-113  (SharedMutationJoiningSolver's default)
 *   CAS_FN_HAMMING_SIMILARITY_NORMALIZED_OVER_MISSING   :116-160
 *   CAS_FN_HAMMING_DISTANCE                             :163-194 (integer; w_dev ignored; flag
 *                                                       CAS_FN_FLAG_IGNORE_MISSING = ignore_missing_state)
 *   CAS_FN_WEIGHTED_HAMMING_SIMILARITY                  :197-240
 *   CAS_FN_EXPONENTIAL_NEGATIVE_HAMMING_DISTANCE        :243-300 (exp() within 1 ulp of the host libm)
 * The diagonal is 0 for every function (the reference fills only i < j, scipy squareform zeroes it).
 */
#define CAS_FN_WEIGHTED_HAMMING_DISTANCE 0
#define CAS_FN_HAMMING_SIMILARITY_WITHOUT_MISSING 1
#define CAS_FN_HAMMING_SIMILARITY_NORMALIZED_OVER_MISSING 2
#define CAS_FN_HAMMING_DISTANCE 3
#define CAS_FN_WEIGHTED_HAMMING_SIMILARITY 4
#define CAS_FN_EXPONENTIAL_NEGATIVE_HAMMING_DISTANCE 5
#define CAS_FN_FLAG_IGNORE_MISSING 1
CAS_API int cas_pairwise_map(int32_t fn, int32_t flags, const int32_t *cm_dev, int64_t n, int32_t m,
                     int32_t missing, const double *w_dev, int32_t n_states, void *D_dev, int64_t ld,
                     int32_t dtype, int64_t row0, int64_t row1, void *workspace_dev, size_t workspace_bytes,
                     void *stream);

/* ---- agglomeration loops ---------------------------------------------------
 * D_dev    : [n, ld] symmetric map (both triangles valid), destroyed.  ld % 32 == 0.
 * joins_dev: [max(n-2,0), 2] int32 out; join t merges ids (lo, hi) into id n+t
 *            (leaf at row p has id p).  Ties resolve to the lexicographically smallest
 *            (value, id_lo, id_hi) == the reference's first row-major arg-min.
 * last2_dev: [2] int32 out, ids of the two clusters left (ascending); only meaningful
 *            when the loop ran to the end.
 * max_iters: < 0 runs all n-2 joins; otherwise stops after that many (bounded samples,
 *            single-step hooks, profiling).
 */
CAS_API size_t cas_agglom_workspace_bytes(int64_t n, int32_t algo, int32_t mode);
CAS_API int cas_nj_solve(void *D_dev, int64_t n, int64_t ld, int32_t dtype, int32_t mode,
                 int64_t max_iters, int32_t *joins_dev, int32_t *last2_dev, void *workspace_dev,
                 size_t workspace_bytes, void *stream);
CAS_API int cas_upgma_solve(void *D_dev, int64_t n, int64_t ld, int32_t dtype, int32_t mode,
                    int64_t max_iters, int32_t *joins_dev, int32_t *last2_dev,
                    void *workspace_dev, size_t workspace_bytes, void *stream);
/* ---- Shared-Mutation-Joining loop ------------------------------------------------------------
 * Replaces the `while N > 1` loop of cassiopeia/solver/SharedMutationJoiningSolver.py:176-205:
 * find_cherry (:211-228, first row-major arg-max of the similarity map) and
 * update_similarity_map_and_character_matrix (:230-343: the new node gets the Camin-Sokal LCA of the two
 * joined character rows, its similarities are recomputed with the pair function `fn`).
 * cm_dev   : [n, m] int32 compact state codes (as for cas_whd_map), rows in map order
 * S_dev    : [n, ld] float64 similarity map of those rows (e.g. from cas_pairwise_map with the same fn),
 *            destroyed
 * joins_dev: [max(n-2,0), 2] ids as for cas_nj_solve; last2_dev: the two clusters left, which the
 *            reference joins last (its root).  Bit-identical to the reference's float64 arithmetic. */
CAS_API size_t cas_smj_workspace_bytes(int64_t n, int32_t m);
CAS_API int cas_smj_solve(const int32_t *cm_dev, int64_t n, int32_t m, int32_t missing, const double *w_dev,
                  int32_t n_states, int32_t fn, double *S_dev, int64_t ld, int64_t max_iters,
                  int32_t *joins_dev, int32_t *last2_dev, void *workspace_dev, size_t workspace_bytes,
                  void *stream);

/* ---- row-sharded loop (n^2 map larger than one GPU; no reference counterpart) ----------
 * Slots are dealt to ranks in 64-row blocks, block-cyclically; a rank stores its rows in local
 * order, [cas_sharded_local_capacity(n, world), ld].  cas_whd_map_cyclic (method = CAS_MAP_EXACT, or
 * CAS_MAP_TENSOR: the tcgen05 kernel in shard mode, workspace = cas_whd_workspace_bytes(..., CAS_MAP_TENSOR)) fills such a shard
 * (every rank needs only the replicated character matrix: no exchange).  cas_sharded_solve runs
 * the fast-mode loop with one NCCL all-gather of the ranks' candidates and one of the rows
 * (lo, hi, last) per iteration; every rank receives the full, identical join list.
 * cas_sharded_solve_emulated runs all `world` virtual ranks in one process on one GPU with no
 * collective (shared gather buffers) -- the same kernels, used to test the sharded path. */
CAS_API int64_t cas_sharded_local_rows(int64_t n, int32_t rank, int32_t world);
CAS_API int64_t cas_sharded_local_capacity(int64_t n, int32_t world);
/* map elements this rank's pruned scans read during the last sharded solve with this workspace
 * (0 for the dense scan); synchronises the stream */
CAS_API int cas_sharded_scanned_elements(const void *workspace_dev, int64_t n, int32_t world, int32_t dtype,
                                 void *stream, int64_t *elements);
CAS_API size_t cas_sharded_workspace_bytes(int64_t n, int32_t world, int32_t dtype);
CAS_API int cas_whd_map_cyclic(const int32_t *cm_dev, int64_t n, int32_t m, int32_t missing,
                       const double *w_dev, int32_t n_states, void *D_local_dev, int64_t ld,
                       int32_t dtype, int32_t rank, int32_t world, int32_t method,
                       void *workspace_dev, size_t workspace_bytes, void *stream);
CAS_API int cas_nccl_unique_id(void *out128);
CAS_API int cas_nccl_comm_create(const void *id128, int32_t rank, int32_t world, void **comm_out);
CAS_API int cas_nccl_comm_destroy(void *comm);
CAS_API int cas_sharded_solve(void *D_local_dev, int64_t n, int64_t ld, int32_t dtype, int32_t algo,
                      int64_t max_iters, int32_t rank, int32_t world, void *nccl_comm,
                      int32_t *joins_dev, int32_t *last2_dev, void *workspace_dev,
                      size_t workspace_bytes, void *stream);
CAS_API int cas_sharded_solve_emulated(void *const *D_locals_dev, int64_t n, int64_t ld, int32_t dtype,
                               int32_t algo, int64_t max_iters, int32_t world, int32_t *joins_dev,
                               int32_t *last2_dev, void *const *workspaces_dev,
                               size_t workspace_bytes, void *stream);

/* P2P variant: no collective library in the loop.  Every rank owns an exchange buffer
 * (cas_exchange_alloc: cudaMalloc + cudaIpc handle) that its peers map (cas_exchange_open) and
 * write with plain NVLink stores from inside the scan / cols kernels; iteration-tagged flags
 * (release/acquire at system scope) replace the two all-gathers.  exchange_ptrs[r] = rank r's buffer
 * as mapped in this process; `epoch` must increase with every solve on the same buffers. */
CAS_API size_t cas_exchange_bytes(int64_t n, int32_t world, int32_t dtype);
CAS_API int cas_exchange_alloc(size_t bytes, void **ptr_out, void *handle64_out);
CAS_API int cas_exchange_open(const void *handle64, void **ptr_out);
CAS_API int cas_exchange_close(void *ptr);
CAS_API int cas_exchange_free(void *ptr);
CAS_API int cas_exchange_error(const void *own_exchange_dev, void *stream, int32_t *error_out);
/* mode: CAS_MODE_FAST (float32 or float64 map, maintained sums) or CAS_MODE_EXACT (float64 map, NJ: the reference's
 * join list -- in an ambiguous iteration the ranks exchange their marked slots, the owners evaluate the reference's
 * sequential row sums of those slots and every rank decides identically; at most 1024 slots per iteration, else the
 * exchange error flag is set to 2).  cas_sharded_exact_stats: as cas_agglom_exact_stats, for a sharded workspace. */
CAS_API int cas_sharded_solve_p2p(void *D_local_dev, int64_t n, int64_t ld, int32_t dtype, int32_t algo, int32_t mode,
                          int64_t max_iters, int32_t rank, int32_t world, void *const *exchange_ptrs,
                          uint64_t epoch, int32_t *joins_dev, int32_t *last2_dev, void *workspace_dev,
                          size_t workspace_bytes, void *stream);
CAS_API int cas_sharded_solve_emulated_p2p(void *const *D_locals_dev, int64_t n, int64_t ld, int32_t dtype,
                                   int32_t algo, int32_t mode, int64_t max_iters, int32_t world,
                                   void *const *exchange_ptrs, uint64_t epoch, int32_t *joins_dev,
                                   int32_t *last2_dev, void *const *workspaces_dev,
                                   size_t workspace_bytes, void *stream);
CAS_API int cas_sharded_exact_stats(const void *workspace_dev, int64_t n, int32_t world, int32_t dtype, void *stream,
                            int64_t *out4);

/* ---- single-step hooks (float64, reference operation order) -----------------
 * cas_nj_q       : Q[n,n] (zero diagonal) of NeighborJoiningSolver.compute_q
 *                  (cassiopeia/solver/NeighborJoiningSolver.py:142-171); rowsum_scratch = n doubles.
 * cas_update_map : the (n-1)x(n-1) map after joining rows i and j -- the two rows/columns are
 *                  dropped and the merged node appended last
 *                  (NeighborJoiningSolver.py:218-260 / UPGMASolver.py:195-238). */
CAS_API int cas_nj_q(const double *D_dev, int64_t n, int64_t ld, double *Q_dev, double *rowsum_scratch_dev,
             void *stream);
CAS_API int cas_update_map(const double *D_dev, int64_t n, int64_t ld, int32_t algo, int64_t i, int64_t j,
                   int64_t size_i, int64_t size_j, double *out_dev, int64_t ld_out, void *stream);

/* Near-tie diagnostics of the last loop run with this workspace (synchronises the stream):
 * near_ties  = iterations whose runner-up criterion value lay within 1e-6 (relative; absolute
 *              below 1) of the winner's -- the joins a float32 map may decide differently from
 *              the reference; exact_ties = iterations decided by the (id_lo, id_hi) rule alone.
 * Counted by the multi-kernel single-GPU loop. */
CAS_API int cas_agglom_stats(const void *workspace_dev, int64_t n, void *stream, int64_t *near_ties,
                     int64_t *exact_ties);

/* CAS_MODE_EXACT diagnostics of the last NJ loop run with this workspace (synchronises the stream):
 * out4[0] iterations in which a second pair lay inside the error window (resolved with sequential sums),
 * out4[1] rows whose sequential sum was evaluated, out4[2] largest such row set in one iteration,
 * out4[3] iterations whose pair differs from the arg-min of the maintained criterion. */
CAS_API int cas_agglom_exact_stats(const void *workspace_dev, int64_t n, void *stream, int64_t *out4);
/* the first (at most 4096) ambiguous iterations of that loop, one row of four int32 each:
 * iteration index, live size k, rows whose sequential sum was evaluated, 0 */
CAS_API int cas_agglom_exact_log(const void *workspace_dev, int64_t n, void *stream, int32_t *out_host, int64_t cap_rows);

/* Map elements the scan kernels of the last loop run with this workspace actually read (synchronises
 * the stream).  The pruned scan (default for every mode but strict NJ; CAS_PRUNE=0 selects the dense
 * scan) skips row segments whose lower bound proves they hold neither the minimum nor a near-tie; the
 * dense scan reads k(k-1)/2 elements per iteration and reports 0 here. */
CAS_API int cas_agglom_scanned_elements(const void *workspace_dev, int64_t n, void *stream, int64_t *elements);

/* pruned-scan diagnostics of the last single-GPU loop run with this workspace (synchronises the stream): rows put
 * on the work list summed over all iterations, and the final cumulative drift C of the bounds */
CAS_API int cas_agglom_prune_diag(const void *workspace_dev, int64_t n, void *stream, int64_t *rows_listed,
                          double *drift);

/* Measurement aid (bench.py's latency model): enqueues `iters` iterations of the pruned loop's launch pattern at
 * live size k -- `kernels_per_iter` programmatically chained launches with the loop's grid sizes -- with kernels
 * that only pass one word along.  The device time of this sequence is the launch floor of the loop. */
CAS_API int cas_loop_launch_floor(int64_t k, int64_t iters, int32_t kernels_per_iter, void *workspace_dev,
                          size_t workspace_bytes, void *stream);

/* Measurement aid (scripts/map_sweep.py): issue rate of the int8 tensor pipe -- every SM issues iters x 4
 * tcgen05.mma kind::i8 (128 x 256 x 32) on operands resident in shared memory; *ops_per_s = 2 M N K ops / s over the
 * device.  The measured denominator of the tensor-core map's roofline fraction. */
CAS_API int cas_int8_mma_peak(int32_t iters, void *stream, double *ops_per_s);

/* number of kernels the last cas_*_solve / cas_whd_map / cas_solve_host call on this
 * thread launched (bench.py's gpu_launches) */
CAS_API int64_t cas_last_launch_count(void);

/* ---- host passes of the array-based solve tail (no device work) ------------
 * The reference finishes DistanceSolver.solve with per-node Python passes over the joined tree:
 * root_tree's depth-first orientation (cassiopeia/solver/NeighborJoiningSolver.py:100-121,
 * UPGMASolver.py:89-116), populate_tree's times (cassiopeia/data/CassiopeiaTree.py:141-203) and the
 * post-order splicing of collapse_unifurcations / collapse_mutationless_edges (:1669-1775).
 * cassiopeia_b200/tree_tail.py computes their result from integer arrays; these are its three
 * per-node loops.  Trees are given as child0 / child1 arrays over node ids [0, n_nodes), -1 = none
 * (a node has 0, 1 (child0 only) or 2 children). */

/* Depth-first preorder, first child first (= node insertion order of the reference's rooted DiGraph)
 * and the depth of every node reached.  Returns the number of nodes written to `order`, or < 0. */
CAS_API int64_t cas_tail_preorder(const int64_t *child0, const int64_t *child1, int64_t n_nodes, int64_t root,
                                  int64_t *order, int64_t *depth);

/* height[v] above the deepest leaf below v (leaves 0), for the nodes of `order` (a preorder). */
CAS_API int cas_tail_heights(const int64_t *order, int64_t n_order, const int64_t *child0, const int64_t *child1,
                             int64_t *height);

/* Successor lists after contracting the edges into the nodes flagged in `removed`, as linked lists
 * (head / tail per node, next per list element; -1 terminated; `next` must come in filled with -1):
 * a node's kept children first, then the lists of its removed children, in child order -- the order
 * the reference's post-order splicing (add_edge appends) leaves in the successor dict. */
CAS_API int cas_tail_successor_lists(const int64_t *order, int64_t n_order, const int64_t *child0,
                                     const int64_t *child1, const uint8_t *removed, int64_t *head, int64_t *tail,
                                     int64_t *next);

/* ---- end to end from HOST buffers (synchronous) ----------------------------
 * Copies the character matrix (and weights) to the device, builds the map, runs the
 * loop and copies the join list back.  timings_ms (optional, [4]) receives device
 * times: h2d, map, loop, d2h. */
CAS_API int cas_solve_host(const int32_t *cm_host, int64_t n, int32_t m, int32_t missing,
                   const double *w_host, int32_t n_states, int32_t algo, int32_t mode,
                   int32_t map_method, int32_t *joins_host, int32_t *last2_host,
                   double *timings_ms);
/* cas_solve_host keeps its device buffers (character matrix, map, workspaces, join list) between calls and re-uses
 * them when they are large enough; this frees them. */
CAS_API int cas_solve_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* CAS_B200_H_ */
