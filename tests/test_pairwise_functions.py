"""The other pair functions of cassiopeia/solver/dissimilarity_functions.py.

CPU: the C oracle and the Python single-pair mirrors against the reference's unit-test answers
(test/solver_tests/dissimilarity_functions_test.py:150-305) and against golden maps produced by the
reference itself (oracle/make_golden_pairwise.py).  GPU: ``cas_pairwise_map`` against the same goldens
(bit-exact; exp() within 2 ulp) and against the oracle on larger seeded inputs."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import pyoracle as po

S1 = [0, 1, 0, -1, 1, 2]
S2 = [1, 1, 0, 0, 1, 3]
ALL_MISSING = [-1] * 6
PRIORS = {0: {1: 0.5, 2: 0.5}, 1: {1: 0.5, 2: 0.5}, 2: {1: 0.25, 2: 0.75}, 3: {1: 0.3, 2: 0.7},
          4: {1: 0.4, 2: 0.6}, 5: {1: 0.1, 2: 0.05, 3: 0.85}}
NLW = {c: {s: -np.log(p) for s, p in sp.items()} for c, sp in PRIORS.items()}
FUNCTIONS = ("weighted_hamming_distance", "hamming_similarity_without_missing",
             "hamming_similarity_normalized_over_missing", "weighted_hamming_similarity",
             "exponential_negative_hamming_distance")


def table_of(weights, m, S):
    tab = np.zeros((m, S + 1))
    for c, sw in weights.items():
        for s, w in sw.items():
            tab[c, s] = w
    return tab


NLTAB = table_of(NLW, 6, 3)


def both(name, a, b, missing=-1, weights=None, **kw):
    """(oracle value, Python mirror value) of one pair."""
    from cassiopeia_b200.solver import dissimilarity_functions as df

    o = po.pair_fn(name, a, b, missing, None if weights is None else NLTAB, **kw)
    if name == "hamming_distance":
        p = df.hamming_distance(a, b, kw.get("ignore_missing", False), missing)
    else:
        p = getattr(df, name)(a, b, missing, weights)
    return o, p


@pytest.mark.parametrize("name,a,b,weights,expected", [
    # dissimilarity_functions_test.py:150-180
    ("hamming_similarity_without_missing", S1, S1, None, 3),
    ("hamming_similarity_without_missing", S1, S2, None, 2),
    ("hamming_similarity_without_missing", S1, S2, NLW, -np.log(0.5) + -np.log(0.4)),
    ("hamming_similarity_without_missing", S1, ALL_MISSING, NLW, 0),
    # :182-220
    ("hamming_similarity_normalized_over_missing", S1, S1, None, 3 / 5),
    ("hamming_similarity_normalized_over_missing", S1, S2, None, 2 / 5),
    ("hamming_similarity_normalized_over_missing", S1, S2, NLW, (-np.log(0.5) + -np.log(0.4)) / 5),
    ("hamming_similarity_normalized_over_missing", S1, ALL_MISSING, NLW, 0),
    # :222-252
    ("weighted_hamming_similarity", S1, S1, None, 8 / 5),
    ("weighted_hamming_similarity", S1, S2, None, 1),
    ("weighted_hamming_similarity", S1, S2, NLW, (-np.log(0.5) * 2 + -np.log(0.4) * 2) / 5),
    ("weighted_hamming_similarity", S1, ALL_MISSING, NLW, 0),
    # exp(-whd): :94-99 through :293-300
    ("exponential_negative_hamming_distance", S1, S2, None, np.exp(-3 / 5)),
    ("exponential_negative_hamming_distance", S1, ALL_MISSING, None, 0),
])
def test_reference_known_answers(name, a, b, weights, expected):
    o, p = both(name, a, b, -1, weights)
    if name.startswith("exponential") and expected != 0:
        assert abs(o - expected) <= 2 * np.spacing(expected) and abs(p - expected) <= 2 * np.spacing(expected)
    else:
        assert o == expected and p == expected


def test_reference_known_answers_hamming_distance():
    # dissimilarity_functions_test.py:289-305
    assert both("hamming_distance", S1, S2) == (3, 3)
    assert both("hamming_distance", S1, S2, ignore_missing=True) == (2, 2)
    assert both("hamming_distance", S1, ALL_MISSING, ignore_missing=True) == (0, 0)


def pairwise_goldens():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "pairwise", "*.npz")))


def golden_inputs(path):
    z = np.load(path)
    cm = z["cm"].astype(np.int64)
    pri = z["priors"]
    with np.errstate(divide="ignore", invalid="ignore"):
        tab = np.nan_to_num(-np.log(pri), nan=0.0, posinf=0.0)
    return z, cm, int(z["missing"]), tab


def assert_map_equal(name, got, want):
    if name.startswith("exponential"):
        np.testing.assert_array_max_ulp(got, want, maxulp=2)
        assert np.array_equal(got == 0, want == 0)
    else:
        assert np.array_equal(got, want), name


@pytest.mark.parametrize("path", pairwise_goldens(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_maps(path):
    z, cm, missing, tab = golden_inputs(path)
    assert len(z.files) == 3 + 2 * len(FUNCTIONS) + 2
    for name in FUNCTIONS:
        assert_map_equal(name, po.pairwise_map(name, cm, missing, None), z[f"{name}_u"])
        assert_map_equal(name, po.pairwise_map(name, cm, missing, tab), z[f"{name}_w"])
    assert np.array_equal(po.pairwise_map("hamming_distance", cm, missing), z["hamming_distance_u"])
    assert np.array_equal(po.pairwise_map("hamming_distance", cm, missing, ignore_missing=True),
                          z["hamming_distance_ignore_u"])
    # fn 0 of the generic entry point is the map oracle pinned elsewhere
    assert np.array_equal(po.pairwise_map("weighted_hamming_distance", cm, missing, tab), po.whd_map(cm, missing, tab))


def test_kernel_name_resolution():
    from cassiopeia_b200.solver import dissimilarity_functions as df

    for name in FUNCTIONS:
        assert df.gpu_kernel_name(getattr(df, name)) == name
    assert df.gpu_kernel_name(df.hamming_distance) is None   # different signature, see docstring
    assert df.gpu_kernel_name(lambda a, b, m, w: 0.0) is None

    def weighted_hamming_distance(a, b, m, w):               # same name, foreign module: not ours
        return 0.0

    assert df.gpu_kernel_name(weighted_hamming_distance) is None


def test_match_lookup_key_errors():
    """A similarity function looks a state up when two samples share it (reference :104-110)."""
    from cassiopeia_b200 import host_prep

    cm = np.array([[1, 2], [1, 3], [0, 2]], dtype=np.int64)
    # state 1 of character 0 and state 2 of character 1 are shared; state 3 occurs once
    host_prep.compact_states(cm, -1, {0: {1: 0.5}, 1: {2: 0.7}}, lookup="match")
    with pytest.raises(KeyError):
        host_prep.compact_states(cm, -1, {0: {1: 0.5}, 1: {3: 0.7}}, lookup="match")
    with pytest.raises(KeyError):   # mismatch lookup needs state 3 too
        host_prep.compact_states(cm, -1, {0: {1: 0.5}, 1: {2: 0.7}}, lookup="mismatch")


# ------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", pairwise_goldens(), ids=lambda p: os.path.basename(p)[:-4])
def test_gpu_pairwise_maps_match_reference(path):
    from cassiopeia_b200 import engine

    z, cm, missing, tab = golden_inputs(path)
    n = cm.shape[0]
    n_states = tab.shape[1] - 1
    codes, missing_code = cm.astype(np.int32), missing
    for name in FUNCTIONS:
        for tag, w in (("u", None), ("w", tab)):
            D = engine.whd_map(codes, missing_code, w, n_states, dtype="f64", function=name)[:, :n].cpu().numpy()
            assert_map_equal(name, D, z[f"{name}_{tag}"])
            D32 = engine.whd_map(codes, missing_code, w, n_states, dtype="f32", function=name)[:, :n].cpu().numpy()
            if not name.startswith("exponential"):
                assert np.array_equal(D32, z[f"{name}_{tag}"].astype(np.float32))
    for tag, ignore in (("hamming_distance_u", False), ("hamming_distance_ignore_u", True)):
        D = engine.whd_map(codes, missing_code, None, n_states, dtype="f64", function="hamming_distance",
                           ignore_missing=ignore)[:, :n].cpu().numpy()
        assert np.array_equal(D, z[tag])


@pytest.mark.gpu
def test_gpu_pairwise_maps_match_oracle_at_size():
    from cassiopeia_b200 import engine
    from cassiopeia_b200.synthetic import simulate_character_matrix

    n, m, S = 1500, 47, 30
    cm, priors, table = simulate_character_matrix(n, m, S, seed=5, mutation_rate=0.7, missing_fraction=0.2)
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.nan_to_num(-np.log(table), nan=0.0, posinf=0.0)
    for name in FUNCTIONS:
        for ww in (None, w):
            D = engine.whd_map(cm, -1, ww, S, dtype="f64", function=name)[:, :n].cpu().numpy()
            assert_map_equal(name, D, po.pairwise_map(name, cm.astype(np.int64), -1, ww))
    for ignore in (False, True):
        D = engine.whd_map(cm, -1, None, S, dtype="f64", function="hamming_distance",
                           ignore_missing=ignore)[:, :n].cpu().numpy()
        assert np.array_equal(D, po.pairwise_map("hamming_distance", cm.astype(np.int64), -1, ignore_missing=ignore))
    # row blocks agree with the full map
    full = engine.whd_map(cm, -1, w, S, dtype="f64", function="weighted_hamming_similarity")[:, :n]
    blk = engine.whd_map(cm, -1, w, S, dtype="f64", function="weighted_hamming_similarity", rows=(100, 300))[:, :n]
    assert bool((full[100:300] == blk).all())


@pytest.mark.gpu
def test_gpu_compute_dissimilarity_map_with_similarity_function():
    """The reference-facing wrapper (data/utilities.py:189-295 contract: condensed vector, pair order)."""
    from cassiopeia_b200.data import utilities
    from cassiopeia_b200.mixins import DistanceSolverError
    from cassiopeia_b200.solver import dissimilarity_functions as df

    z, cm, missing, tab = golden_inputs(pairwise_goldens()[0])
    n = cm.shape[0]
    weights = {c: {s: float(tab[c, s]) for s in range(1, tab.shape[1])} for c in range(tab.shape[0])}
    vec = utilities.compute_dissimilarity_map(cm, n, df.hamming_similarity_without_missing, weights, missing)
    iu = np.triu_indices(n, k=1)
    assert np.array_equal(vec, z["hamming_similarity_without_missing_w"][iu])
    with pytest.raises(DistanceSolverError):
        utilities.compute_dissimilarity_map(cm, n, lambda a, b, m, w: 0.0, None, missing)
