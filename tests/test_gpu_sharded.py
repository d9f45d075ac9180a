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
    switches to the dense scan below 3072 live nodes, which these fixtures never exceed), with pruning forced
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
