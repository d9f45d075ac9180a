"""Tensor-core (tcgen05) dissimilarity map vs the exact kernel / the reference goldens."""
import numpy as np
import pytest

from conftest import Golden, golden_names

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6  # north-star tolerance for prior-weighted distances (float32 storage)


@pytest.fixture(params=["2", "1"], ids=["sm_pair", "single_sm"], autouse=True)
def cta_group(request, monkeypatch):
    """Both tensor-core kernels are tested: the SM-pair kernel (tcgen05 cta_group::2, the default) and
    the single-SM 128 x 128 kernel (CAS_GEMM_CTA_GROUP=1); the library reads the variable per call."""
    monkeypatch.setenv("CAS_GEMM_CTA_GROUP", request.param)
    return request.param


@pytest.fixture(params=["i8", "bf16"], ids=["fixed_point_int8", "bf16_planes"], autouse=True)
def weighted_path(request, monkeypatch):
    """Both weighted tensor paths: four 7-bit digit planes on the int8 MMAs (default) and three bf16 planes."""
    monkeypatch.setenv("CAS_GEMM_WEIGHTED", request.param)
    return request.param


def _inputs(g):
    w = g.weights
    if w is not None:
        w = np.where(np.isnan(w), 0.0, w)
        S = w.shape[1] - 1
    else:
        S = max(int(g.ordered_cm.max()), 1)
    return g.ordered_cm.astype(np.int32), w, S


@pytest.mark.parametrize("name", [n for n in golden_names() if n.endswith("plain")])
def test_int8_gemm_is_bit_exact(name):
    """Unweighted: integer numerator/denominator from the int8 MMA, IEEE division => bit-identical."""
    from cassiopeia_b200 import engine

    g = Golden(name)
    cm, w, S = _inputs(g)
    n = cm.shape[0]
    D = engine.whd_map(cm, g.missing, None, S, dtype="f64", method="tensor")[:, :n].cpu().numpy()
    ref = engine.whd_map(cm, g.missing, None, S, dtype="f64", method="exact")[:, :n].cpu().numpy()
    assert np.array_equal(D, ref)
    if g.D is not None:
        assert np.array_equal(D, g.D)
    D32 = engine.whd_map(cm, g.missing, None, S, dtype="f32", method="tensor")[:, :n].cpu().numpy()
    assert np.array_equal(D32, ref.astype(np.float32))


@pytest.mark.parametrize("name", [n for n in golden_names() if not n.endswith("plain")])
def test_weighted_gemm_within_tolerance(name):
    from cassiopeia_b200 import engine

    g = Golden(name)
    cm, w, S = _inputs(g)
    n = cm.shape[0]
    ref = engine.whd_map(cm, g.missing, w, S, dtype="f64", method="exact")[:, :n].cpu().numpy()
    D = engine.whd_map(cm, g.missing, w, S, dtype="f64", method="tensor")[:, :n].cpu().numpy()
    err = np.abs(D - ref) / np.maximum(np.abs(ref), 1.0)
    print(f"{name}: max rel err {err.max():.3e}")
    assert err.max() <= REL_TOL
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    # exact zeros stay exact zeros (duplicates, all-missing rows)
    assert np.array_equal(D == 0, ref == 0)


def test_gemm_sizes_not_multiple_of_tile():
    from cassiopeia_b200 import engine
    from cassiopeia_b200.synthetic import simulate_character_matrix

    for n, m, S in ((5, 3, 2), (129, 7, 3), (255, 20, 9), (257, 20, 9), (300, 33, 17), (1000, 40, 50), (1537, 70, 20),
                    (700, 150, 6)):
        cm, priors, table = simulate_character_matrix(n, m, S, seed=n, mutation_rate=0.7, missing_fraction=0.15)
        with np.errstate(invalid="ignore"):
            w = np.nan_to_num(-np.log(table))
        ref = engine.whd_map(cm, -1, None, S, dtype="f64", method="exact")[:, :n].cpu().numpy()
        D = engine.whd_map(cm, -1, None, S, dtype="f64", method="tensor")[:, :n].cpu().numpy()
        assert np.array_equal(D, ref), (n, m, S)
        refw = engine.whd_map(cm, -1, w, S, dtype="f64", method="exact")[:, :n].cpu().numpy()
        Dw = engine.whd_map(cm, -1, w, S, dtype="f64", method="tensor")[:, :n].cpu().numpy()
        assert (np.abs(Dw - refw) / np.maximum(np.abs(refw), 1.0)).max() <= REL_TOL, (n, m, S)
