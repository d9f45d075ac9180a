"""Row-sharded loop on ONE GPU: all virtual ranks run in this process with shared gather buffers
(cas_sharded_solve_emulated) -- the same kernels the multi-process NCCL path launches.  Sharding
must not change the join list."""
import numpy as np
import pytest

from conftest import Golden

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["pruned", "pruned-upgma-too", "dense"], autouse=True)
def scan_policy(request, monkeypatch):
    """Every sharded test runs with the pruned scan kept down to the last iteration (the default policy
    switches to the dense scan below 2048 live nodes, which these fixtures never exceed), with pruning forced
    for UPGMA as well, and with the dense scan."""
    monkeypatch.setenv("CAS_PRUNE_MIN_K", "0")
    if request.param == "dense":
        monkeypatch.setenv("CAS_PRUNE", "0")
    elif request.param == "pruned-upgma-too":
        monkeypatch.setenv("CAS_PRUNE", "1")
    else:
        monkeypatch.delenv("CAS_PRUNE", raising=False)
    return request.param


def _inputs(g):
    w = g.weights
    if w is not None:
        w = np.where(np.isnan(w), 0.0, w)
        S = w.shape[1] - 1
    else:
        S = int(g.ordered_cm.max())
    return g.ordered_cm.astype(np.int32), w, S


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["sim256_s1_nj_priors", "sim1024_s0_nj_priors", "ties120_nj_plain"])
def test_sharded_nj_equals_single_gpu_fast(name, world):
    from cassiopeia_b200 import engine, sharded

    g = Golden(name)
    cm, w, S = _inputs(g)
    n = cm.shape[0]
    D = engine.whd_map(cm, g.missing, w, S, dtype="f32")
    single, last_single = engine.agglomerate(D, n, "nj", "fast")
    shard, last_shard = sharded.emulated_solve(cm, g.missing, w, S, "nj", world, dtype="f32")
    bad = np.nonzero((single != shard).any(axis=1))[0]
    assert bad.size == 0, f"world={world}: first differing join t={bad[0]} {single[bad[0]]} vs {shard[bad[0]]}"
    assert np.array_equal(last_single, last_shard)


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("name", ["sim1024_s0_nj_priors", "ties120_nj_plain"])
def test_sharded_p2p_kernels_equal_single_gpu(name, world):
    """The P2P exchange path (stores into every rank's exchange buffer + tagged flags), emulated in
    one process where the "peer" buffers are local."""
    from cassiopeia_b200 import engine, sharded

    g = Golden(name)
    cm, w, S = _inputs(g)
    n = cm.shape[0]
    D = engine.whd_map(cm, g.missing, w, S, dtype="f32")
    single, last_single = engine.agglomerate(D, n, "nj", "fast")
    shard, last_shard = sharded.emulated_solve(cm, g.missing, w, S, "nj", world, dtype="f32", p2p=True)
    assert np.array_equal(single, shard) and np.array_equal(last_single, last_shard)


def test_sharded_p2p_upgma_f64_reproduces_reference():
    from cassiopeia_b200 import sharded

    g = Golden("sim256_s1_upgma_priors")
    cm, w, S = _inputs(g)
    joins, _ = sharded.emulated_solve(cm, g.missing, w, S, "upgma", 4, dtype="f64", p2p=True)
    assert np.array_equal(joins, g.joins)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", ["sim1024_s0_upgma_plain", "sim256_s1_upgma_priors", "ties120_upgma_priors"])
def test_sharded_upgma_f64_reproduces_reference(name, world):
    """UPGMA has no reductions, so the float64 sharded loop matches the reference join for join."""
    from cassiopeia_b200 import sharded

    g = Golden(name)
    cm, w, S = _inputs(g)
    joins, last2 = sharded.emulated_solve(cm, g.missing, w, S, "upgma", world, dtype="f64")
    assert np.array_equal(joins, g.joins)


def test_cyclic_shards_tile_the_full_map():
    import torch
    from cassiopeia_b200 import engine, sharded

    g = Golden("sim256_s1_nj_priors")
    cm, w, S = _inputs(g)
    n = cm.shape[0]
    full = engine.whd_map(cm, g.missing, w, S, dtype="f64")[:, :n].cpu().numpy()
    cm_dev = torch.from_numpy(cm).cuda()
    for world in (2, 3):
        seen = np.zeros(n, dtype=bool)
        for r in range(world):
            shard = sharded.build_shard(cm_dev, g.missing, w, S, r, world, "f64")[:, :n].cpu().numpy()
            rows = sharded.shard_rows(n, r, world)
            assert np.array_equal(shard[: len(rows)], full[rows])
            seen[rows] = True
        assert seen.all()


@pytest.mark.parametrize("weighted_path", ["i8", "bf16"])
@pytest.mark.parametrize("n,m,S,world", [(5, 3, 2, 2), (130, 7, 3, 2), (257, 20, 9, 3), (700, 33, 17, 8),
                                         (3000, 40, 50, 4), (4100, 60, 100, 8)])
def test_tensor_core_shards_tile_the_full_tensor_map(n, m, S, world, weighted_path, monkeypatch):
    """cas_whd_map_cyclic with CAS_MAP_TENSOR (tcgen05 kernel in shard mode: two 64-row half boxes per row tile, every
    column of the rank's rows, no mirroring): unweighted shards are bit-identical to the rows of the exact map,
    prior-weighted shards bit-identical to the rows of the full tensor map (same products, same accumulation order)
    and within 1e-6 of the exact map."""
    import torch
    from cassiopeia_b200 import engine, sharded
    from cassiopeia_b200.synthetic import simulate_character_matrix

    monkeypatch.setenv("CAS_GEMM_WEIGHTED", weighted_path)
    monkeypatch.setenv("CAS_GEMM_CTA_GROUP", "1")  # the shard mode lives in the single-SM kernel
    cm, priors, table = simulate_character_matrix(n, m, S, seed=n, mutation_rate=0.6, missing_fraction=0.15)
    with np.errstate(invalid="ignore"):
        w = np.nan_to_num(-np.log(table))
    cm_dev = torch.from_numpy(np.ascontiguousarray(cm, dtype=np.int32)).cuda()
    for weights, dtype in ((None, "f64"), (None, "f32"), (w, "f64"), (w, "f32")):
        if weights is None and weighted_path == "bf16":
            continue  # the unweighted path does not depend on the switch
        exact = engine.whd_map(cm, -1, weights, S, dtype=dtype, method="exact")[:, :n]
        full = engine.whd_map(cm, -1, weights, S, dtype=dtype, method="tensor")[:, :n]
        seen = np.zeros(n, dtype=bool)
        for r in range(world):
            shard = sharded.build_shard(cm_dev, -1, weights, S, r, world, dtype, method="tensor")[:, :n]
            rows = sharded.shard_rows(n, r, world)
            got = shard[: len(rows)]
            if len(rows) == 0:
                assert float(shard.abs().max()) == 0.0  # fewer 64-row blocks than ranks: an empty shard
                continue
            idx = torch.from_numpy(np.asarray(rows, dtype=np.int64)).cuda()
            if weights is None:
                assert torch.equal(got, exact[idx]), (dtype, r)
            else:
                assert torch.equal(got, full[idx]), (dtype, r)
                rel = ((got.double() - exact[idx].double()).abs() / exact[idx].double().abs().clamp(min=1.0)).max().item()
                assert rel <= 1e-6
            assert float(shard[len(rows):].abs().max()) == 0.0 if shard.shape[0] > len(rows) else True
            seen[rows] = True
        assert seen.all()


def test_sharded_solve_on_tensor_core_shards_equals_single_gpu():
    """Unweighted data: shards built by the tensor kernel are the exact map, so the sharded exact-mode loop on them
    gives the single-GPU exact-mode join list."""
    import torch
    from cassiopeia_b200 import engine, sharded
    from cassiopeia_b200.synthetic import simulate_character_matrix

    n, m, S, world = 2500, 40, 50, 4
    cm, _, _ = simulate_character_matrix(n, m, S, seed=3, mutation_rate=0.3, missing_fraction=0.1)
    D = engine.whd_map(cm, -1, None, S, dtype="f64", method="tensor")
    want, want_last = engine.agglomerate(D, n, "nj", "exact")
    got, got_last = sharded.emulated_solve(cm, -1, None, S, "nj", world, dtype="f64", p2p=True, mode="exact",
                                           map_method="tensor")
    assert np.array_equal(got, want) and np.array_equal(got_last, want_last)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["sim64_s0_nj_priors", "sim256_s1_nj_plain", "sim1024_s0_nj_priors",
                                  "ties120_nj_plain", "ties120_nj_priors"])
def test_sharded_exact_mode_reproduces_reference_joins(name, world, scan_policy):
    """Exact (adaptive strict) mode, row-sharded: in an ambiguous iteration the ranks exchange their marked slots,
    the owners evaluate the reference's sequential sums and every rank decides identically -- the join list must be
    the reference's (golden fixtures generated by the reference itself) for every number of ranks."""
    from cassiopeia_b200 import sharded

    if scan_policy == "dense":
        pytest.skip("the sharded exact mode always uses the pruned scan")
    g = Golden(name)
    cm, w, S = _inputs(g)
    joins, last2 = sharded.emulated_solve(cm, g.missing, w, S, "nj", world, dtype="f64", p2p=True, mode="exact")
    bad = np.nonzero((joins != g.joins).any(axis=1))[0]
    assert bad.size == 0, f"world={world}: first differing join t={bad[0]} {joins[bad[0]]} vs {g.joins[bad[0]]}"
    st = sharded.last_exact_stats
    assert st["ambiguous_iterations"] >= 2 and st["rows_resolved"] >= 7  # at least the last two iterations


def test_sharded_exact_mode_equals_single_gpu_exact_mode_at_6000_cells():
    from cassiopeia_b200 import engine, sharded
    from cassiopeia_b200.synthetic import simulate_character_matrix

    n, m, S = 6000, 40, 50
    cm, priors, table = simulate_character_matrix(n, m, S, seed=5, mutation_rate=0.1, missing_fraction=0.1)
    cm = np.vstack([cm, np.zeros((1, m), dtype=cm.dtype)])
    with np.errstate(invalid="ignore"):
        w = np.nan_to_num(-np.log(table))
    D = engine.whd_map(cm, -1, w, S, dtype="f64")
    single, last_single = engine.agglomerate(D, n + 1, "nj", "exact")
    amb_single = engine.last_stats["ambiguous_iterations"]
    shard, last_shard = sharded.emulated_solve(cm, -1, w, S, "nj", 4, dtype="f64", p2p=True, mode="exact")
    bad = np.nonzero((single != shard).any(axis=1))[0]
    assert bad.size == 0, f"first differing join t={bad[0]} {single[bad[0]]} vs {shard[bad[0]]}"
    assert np.array_equal(last_single, last_shard)
    print("ambiguous iterations: single", amb_single, "sharded", sharded.last_exact_stats)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("weighted", [False, True])
def test_sharded_exact_mode_on_a_star_like_map_uses_the_strict_tail(world, weighted, scan_policy):
    """Every cell carries its own state at every character, so d_ij = a_i + a_j (all equal without priors): every q
    ties in exact arithmetic and the reference's choice hangs on the rounding of its sequential row sums in every
    iteration.  The sharded resolver meets more than 64 tied slots, raises the strict flag, and from then on the
    ranks re-accumulate their rows' sequential sums before each scan."""
    from cassiopeia_b200 import sharded
    from oracle import pyoracle as po

    if scan_policy == "dense":
        pytest.skip("the sharded exact mode always uses the pruned scan")
    n, m = 150, 6
    rng = np.random.default_rng(world)
    cm = np.tile(np.arange(1, n + 1, dtype=np.int32)[:, None], (1, m))
    cm[-1] = 0  # the implicit root row
    w = None
    if weighted:
        w = np.concatenate([np.zeros((m, 1)), rng.random((m, n)) + 0.5], axis=1)
    want, want_last = po.nj(po.whd_map(cm.astype(np.int64), -1, w))
    joins, last2 = sharded.emulated_solve(cm, -1, w, n, "nj", world, dtype="f64", p2p=True, mode="exact")
    bad = np.nonzero((joins != want).any(axis=1))[0]
    assert bad.size == 0, f"world={world}: first differing join t={bad[0]} {joins[bad[0]]} vs {want[bad[0]]}"
    assert np.array_equal(last2, want_last)
    st = sharded.last_exact_stats
    assert st["max_rows_resolved"] > 64 and st["ambiguous_iterations"] < n // 2
