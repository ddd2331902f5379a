"""CPU tests of the oracle: it must reproduce (a) the reference's own source compiled here
(oracle/_ref, when built), (b) the golden vectors that source produced, (c) the reference's own
known-answer / property tests for this path (tests/testthat/test-core.R:106-127,
test-diagnostics.R:5-20, test-edge-cases.R:5-108)."""
import numpy as np
import pytest

from conftest import load_golden, random_r_matrix, small_problem
from oracle import cpu_oracle, r_glue

GOLDENS = ["triangle", "small_thresholds", "small_sparse", "h3n2_ndim5", "hiv_ndim5"]


@pytest.mark.parametrize("name", GOLDENS)
def test_oracle_reproduces_reference_golden(name):
    z, args, seed = load_golden(name)
    res = cpu_oracle.optimize_layout_exact(*args, seed=seed)
    assert np.array_equal(res["positions"], z["positions"])       # bit for bit
    assert res["iterations"] == int(z["iterations"])
    assert res["converged"] == bool(z["converged"])
    assert res["final_mae"] == float(z["final_mae"])
    assert res["final_k"] == float(z["final_k"])


@pytest.mark.parametrize("name", GOLDENS[:3])
def test_sparse_lookup_and_explicit_order_are_the_same_loop(name):
    z, args, seed = load_golden(name)
    n = args[0].shape[0]
    sparse = cpu_oracle.optimize_layout_exact(*args, seed=seed, dense=False)
    assert np.array_equal(sparse["positions"], z["positions"])
    order = cpu_oracle.pair_orders(n, args[6], seed)
    explicit = cpu_oracle.optimize_layout_exact(*args, pair_order=order)
    assert np.array_equal(explicit["positions"], z["positions"])
    assert explicit["iterations"] == int(z["iterations"])


@pytest.mark.skipif(not cpu_oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [0, 5, 99])
def test_oracle_equals_compiled_reference_source(seed):
    args = small_problem(45, 3, 0.35, seed)
    a = cpu_oracle.optimize_layout_exact(*args, 150, 4.0, 0.02, 0.03, seed=seed)
    b = cpu_oracle.ref_optimize_layout_exact(*args, 150, 4.0, 0.02, 0.03, seed=seed)
    assert np.array_equal(a["positions"], b["positions"])
    for k in ("iterations", "converged", "final_mae", "final_k"):
        assert a[k] == b[k]


def test_guards_follow_the_reference():
    # src/optimization.cpp:131
    with pytest.raises(cpu_oracle.OracleError, match="Need at least 2 points"):
        cpu_oracle.optimize_layout_exact(np.zeros((1, 2)), [1], [], [], [], [], 5, 1.0, 0.01, 0.01, dense=False)
    # :359-361 - a huge repulsion on coincident points overflows within 10 iterations
    init = np.zeros((4, 2))
    with pytest.raises(cpu_oracle.OracleError, match="Numerical instability at iteration 10"):
        cpu_oracle.optimize_layout_exact(init, [2, 2, 1, 1], [0], [1], [1.0], [0], 20, 1e300, 0.01, 1e308,
                                         convergence_window=100, convergence_check_freq=50)


def test_triangle_relations_like_test_core():
    # tests/testthat/test-core.R:106-127
    m = np.array([[0, 1, 2], [1, 0, 1], [2, 1, 0]], dtype=float)
    r = r_glue.euclidean_embedding(m, 2, 10, 1.0, 0.01, 0.01, seed=3)
    p = r["positions"]
    o = r["order"] if r["order"] is not None else np.arange(3)
    pos = {int(o[i]): p[i] for i in range(3)}
    dist = lambda a, b: float(np.linalg.norm(pos[a] - pos[b]))
    assert dist(0, 2) > dist(0, 1)
    assert dist(0, 2) < dist(0, 1) + dist(1, 2)
    assert set(r) >= {"positions", "est_distances", "mae", "iter", "parameters", "convergence"}


def test_thresholds_and_na_like_test_core():
    # tests/testthat/test-core.R:90-104
    m = np.array([[0, ">2", None], [">2", 0, 4], [None, 4, 0]], dtype=object)
    r = r_glue.euclidean_embedding(m, 2, 10, 1.0, 0.01, 0.01, seed=1)
    e = r["est_distances"]
    assert np.isfinite(e[0, 2]) and e[0, 2] == e[2, 0]


@pytest.mark.parametrize("m", [
    np.zeros((3, 3)),                                                                 # all-zero
    np.array([["0", ">5", "<10"], [">5", "0", ">20"], ["<10", ">20", "0"]], dtype=object),  # thresholds only
])
def test_degenerate_matrices_like_test_edge_cases(m):
    # tests/testthat/test-edge-cases.R:5-63
    r = r_glue.euclidean_embedding(m, 2, 20, 1.0, 0.01, 0.01, seed=2)
    assert np.all(np.isfinite(r["positions"]))


def test_one_measurement_matrix_like_test_edge_cases():
    # tests/testthat/test-edge-cases.R:65-82
    m = np.full((4, 4), np.nan)
    m[0, 1] = m[1, 0] = 5
    np.fill_diagonal(m, 0)
    r = r_glue.euclidean_embedding(m, 2, 50, 1.0, 0.01, 0.1, seed=4)
    assert np.all(np.isfinite(r["positions"]))


def test_error_calculator_known_answers():
    # tests/testthat/test-diagnostics.R:5-20
    true = np.array([[0, 1, 2], [1, 0, 3], [2, 3, 0]], dtype=float)
    pred = true + 0.1
    inp = true.copy()
    inp[0, 2] = inp[2, 0] = np.nan
    e = r_glue.error_calculator_comparison(pred, true, inp)
    assert np.sum(~np.isnan(e["OutSampleError"])) == 2
    assert np.sum(~np.isnan(e["InSampleError"])) == 7
    # tests/testthat/test-edge-cases.R:84-108
    ident = np.full((3, 3), 5.0)
    np.fill_diagonal(ident, 0)
    assert r_glue.error_calculator_comparison(ident, ident)["Completeness"] == 1
    assert r_glue.error_calculator_comparison(np.full((3, 3), np.nan), true)["Completeness"] == 0
    inf_true = np.array([[0, 1, np.inf], [1, 0, 2], [np.inf, 2, 0]])
    r_glue.error_calculator_comparison(true + 0.1, inf_true)


def test_edge_error_masks():
    # src/optimization.cpp:68-78: exact always, '>' only if dist < target, '<' only if dist > target
    pos = np.array([[0.0, 0.0], [3.0, 4.0]])     # dist 5
    tot, cnt = cpu_oracle.edge_error(pos, [0, 0, 0, 0, 0], [1, 1, 1, 1, 1], [4.0, 6.0, 6.0, 4.0, 4.0], [0, 1, -1, 1, -1])
    assert cnt == 3 and tot == pytest.approx(1.0 + 1.0 + 1.0)


def test_likelihood_function_pooling():
    # R/adaptive_sampling.R:2710-2725
    m = random_r_matrix(14, 0.7, 3, thresholds=False)
    r = r_glue.likelihood_function(m, 20, 1e-3, 2, 1.0, 0.01, 0.01, folds=3, seed=5)
    rows = [f for f in r["folds"] if f["n_samples"] > 0]
    tot = sum(f["n_samples"] for f in rows)
    assert r["Holdout_MAE"] == pytest.approx(sum(f["sum_abs_errors"] for f in rows) / tot)
    assert r["NLL"] == pytest.approx(tot * (1 + np.log(2 * r["Holdout_MAE"])))
