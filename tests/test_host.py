"""CPU tests of the host side: C-ABI surface, schedule coverage, the Python mirror of the R glue
(against oracle/r_glue.py), fold construction, sharding over ranks (gloo, world size 2)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys
import warnings

import numpy as np
import pytest

from conftest import ROOT, has_cuda, random_r_matrix, small_problem
from oracle import r_glue
from topolow_b200 import _lib, core, cv, shard


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "topolow_b200.h")).read()
    declared = set(re.findall(r"TOPOLOW_API\s+[\w\s\*]+?\b(topolow_\w+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (topolow_\w+)", nm))
    assert declared <= exported
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name)
    assert b"sm_100a" in L.topolow_version()
    sizes = (C.c_int64 * 3)()
    L.topolow_abi_sizes(C.byref(sizes))       # the ctypes mirror of the three structs matches the header
    assert list(sizes) == [C.sizeof(_lib.Problem), C.sizeof(_lib.Params), C.sizeof(_lib.Result)]


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    args = small_problem(20, 2, 0.5, 0)
    with pytest.raises(_lib.TopolowError) as e:
        _lib.fit(*args, 5, 1.0, 0.01, 0.01)
    assert e.value.status == _lib.ERR_CUDA
    with pytest.raises(_lib.TopolowError):
        _lib.est_distances(np.zeros((4, 2)))


def test_too_few_points_status():
    L = _lib.lib()
    pa = _lib.ProblemArrays(np.zeros((1, 2)), [1], [], [], [], [])
    pr, _ = _lib.make_params(5, 1.0, 0.01, 0.01)
    out = np.zeros((1, 2))
    res = _lib.Result()
    res.positions = out.ctypes.data_as(_lib._dp)
    import ctypes as C
    rc = L.topolow_fit(C.byref(pa.struct), C.byref(pr), C.byref(res))
    assert rc == _lib.ERR_TOO_FEW_POINTS
    assert res.message.decode() == "Need at least 2 points for embedding"   # src/optimization.cpp:131


@pytest.mark.parametrize("n", [2, 3, 31, 33, 65, 97, 193, 335, 700, 1025, 1500])
@pytest.mark.parametrize("prec", [_lib.PREC_F32, _lib.PREC_F64_EXACT])
@pytest.mark.parametrize("tile_points", [32, 64, 96])
def test_schedule_visits_every_pair_exactly_once(n, prec, tile_points):
    for it, max_ctas in ((0, 0), (7, 3)):
        order, geo = _lib.schedule_enumerate(n, 5, it, precision=prec, seed=n + it, max_ctas=max_ctas,
                                             tile_points=tile_points)
        a = np.minimum(order[:, 0], order[:, 1]).astype(np.int64)
        b = np.maximum(order[:, 0], order[:, 1]).astype(np.int64)
        assert (a != b).all() and a.min() >= 0 and b.max() < n
        assert len(np.unique(a * n + b)) == n * (n - 1) // 2
        assert geo["super_blocks"] == 2 * geo["ctas"] * geo["tasks_per_cta"]
        assert geo["super_blocks"] * geo["warps_per_cta"] >= geo["tiles"]


def test_schedule_steps_are_matchings_and_orders_differ_between_iterations():
    n = 200
    o0, _ = _lib.schedule_enumerate(n, 3, 0, seed=1)
    o1, _ = _lib.schedule_enumerate(n, 3, 1, seed=1)
    assert not np.array_equal(o0, o1)
    # the enumerator emits one ring / xor step of one tile pair at a time: consecutive pairs sharing
    # a tile pair and step never repeat a point.  Check greedily: a run of pairs with all-distinct points
    # must on average be long (>= 8) - a sequential-only order would give runs of ~ sqrt(n).
    runs, seen, cur = [], set(), 0
    for i, j in o0:
        if i in seen or j in seen:
            runs.append(cur); seen, cur = set(), 0
        seen.update((int(i), int(j))); cur += 1
    assert np.mean(runs) >= 8


VALIDATION = [
    (dict(ndim=-1), "ndim must be a positive integer"),
    (dict(k0=-1), "k0 must be a positive number"),
    (dict(cooling_rate=1.5), "cooling_rate must be between 0 and 1"),
    (dict(c_repulsion=0), "c_repulsion must be a positive number"),
    (dict(relative_epsilon=-1), "relative_epsilon must be a positive number"),
    (dict(convergence_counter=0.5), "convergence_counter must be a positive integer"),
    (dict(mapping_max_iter=0), "mapping_max_iter must be a positive integer"),
    (dict(convergence_check_freq=0), "convergence_check_freq must be a positive integer"),
]


@pytest.mark.parametrize("override,msg", VALIDATION)
def test_validation_messages_match_the_reference(override, msg):
    # tests/testthat/test-core.R:22-65, R/core.R:202-264
    m = np.array(random_r_matrix(4, 1.0, 0, thresholds=False), dtype=float)
    kw = dict(ndim=2, mapping_max_iter=10, k0=1.0, cooling_rate=0.01, c_repulsion=0.01, relative_epsilon=1e-4,
              convergence_counter=5)
    kw.update(override)
    with pytest.raises(ValueError, match=re.escape(msg)):
        core.euclidean_embedding(m, **kw)


def test_matrix_validation_and_k0_warning():
    with pytest.raises(ValueError, match="dissimilarity_matrix must be a matrix"):
        core.euclidean_embedding("not a matrix", 2, 10, 1.0, 0.01, 0.01)
    with pytest.raises(ValueError, match="dissimilarity_matrix must be square"):
        core.euclidean_embedding(np.arange(6.0).reshape(2, 3), 2, 10, 1.0, 0.01, 0.01)
    m = np.array(random_r_matrix(4, 1.0, 0, thresholds=False), dtype=float)
    with pytest.raises(ValueError, match="initial_positions must have same number of rows"):
        core.euclidean_embedding(m, 2, 10, 1.0, 0.01, 0.01, initial_positions=np.zeros((5, 2)))
    with pytest.raises(ValueError, match="initial_positions must have ndim columns"):
        core.euclidean_embedding(m, 2, 10, 1.0, 0.01, 0.01, initial_positions=np.zeros((4, 3)))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        try:
            core.euclidean_embedding(m, 2, 10, 35.0, 0.01, 0.01)
        except _lib.TopolowError:
            pass  # no GPU here: the warning fires before the native call
        assert any("High k0 value" in str(x.message) for x in w)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("preserve", [True, False])
def test_build_problem_matches_the_r_glue_restatement(seed, preserve):
    m = random_r_matrix(23, 0.4, seed)
    got = core.build_problem(m, preserve)
    want = r_glue.prepare(m, preserve)
    assert (got["order"] is None) == (want["order"] is None)
    if got["order"] is not None:
        assert np.array_equal(got["order"], want["order"])
    for k in ("degrees", "edge_i", "edge_j", "edge_dist", "edge_thresh"):
        assert np.array_equal(got[k], want[k]), k


def test_parse_numeric_and_character_matrices_agree():
    m = random_r_matrix(12, 0.6, 9, thresholds=False)
    num = np.array([[np.nan if x is None else float(x) for x in row] for row in m])
    a, b = core.build_problem(m, True), core.build_problem(num, True)
    for k in ("degrees", "edge_i", "edge_j", "edge_dist", "edge_thresh"):
        assert np.array_equal(a[k], b[k])


def test_folds_follow_the_reference_scheme():
    # R/adaptive_sampling.R:2568-2598
    m = random_r_matrix(20, 0.6, 4)
    folds_a = cv.make_folds(m, 5, np.random.default_rng(7))
    folds_b = r_glue.make_folds(m, 5, np.random.default_rng(7))
    assert len(folds_a) == len(folds_b) == 5
    for a, b in zip(folds_a, folds_b):
        assert np.array_equal(a, b)
    non_na = sum(x is not None for x in m.ravel())
    assert all(len(f) == non_na // 10 for f in folds_a)
    seen = set()
    for f in folds_a:            # a cell and its mirror leave the pool together (a single draw may
        cells = {(int(i) % 20, int(i) // 20) for i in f}   # still hold both orientations of a pair)
        assert not (cells & seen)
        seen |= cells | {(c, r) for r, c in cells}


def test_error_calculator_matches_oracle_and_reference_kats():
    rng = np.random.default_rng(0)
    true = random_r_matrix(9, 0.8, 2)
    inp = true.copy()
    inp[1, 4] = inp[4, 1] = None
    inp[0, 0] = None
    pred = rng.uniform(0, 5, size=(9, 9))
    got = cv.error_calculator_comparison(pred, true, inp)
    want = r_glue.error_calculator_comparison(pred, true, inp)
    for k in ("InSampleError", "OutSampleError", "InSamplePercentageError", "OutSamplePercentageError"):
        np.testing.assert_array_equal(got["report_df"][k], want[k])
    assert got["Completeness"] == want["Completeness"]
    # tests/testthat/test-diagnostics.R:5-20
    t3 = np.array([[0, 1, 2], [1, 0, 3], [2, 3, 0]], dtype=float)
    i3 = t3.copy(); i3[0, 2] = i3[2, 0] = np.nan
    e = cv.error_calculator_comparison(t3 + 0.1, t3, i3)["report_df"]
    assert np.sum(~np.isnan(e["OutSampleError"])) == 2 and np.sum(~np.isnan(e["InSampleError"])) == 7
    with pytest.raises(ValueError, match="All matrices must have the same dimensions"):
        cv.error_calculator_comparison(np.zeros((2, 2)), t3)


def test_partition_is_longest_first_and_complete():
    costs = [5, 1, 9, 3, 3, 7, 2]
    parts = shard.partition(costs, 3)
    assert sorted(j for p in parts for j in p) == list(range(7))
    loads = [sum(costs[j] for j in p) for p in parts]
    assert max(loads) - min(loads) <= max(costs)
    assert shard.partition(costs, 1) == [[2, 5, 0, 3, 4, 6, 1]]


_GLOO_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch.distributed as dist
from topolow_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
jobs = [dict(degrees=np.ones(n, np.int32), n_iter=it, tag=i) for i, (n, it) in enumerate([(50, 10), (20, 5), (80, 3), (10, 100), (60, 7)])]
calls = []
def fake_batch(js, device=0):            # stands in for _lib.fit_batch: no GPU on this box
    calls.append(len(js))
    return [dict(status=0, tag=j["tag"], rank=dist.get_rank()) for j in js]
out = shard.run_sharded(jobs, fit_batch=fake_batch)
assert [r["tag"] for r in out] == [0, 1, 2, 3, 4], out
assert len(calls) == 1                      # one batch call per rank
assert {{r["rank"] for r in out}} == {{0, 1}}
parts = shard.partition([shard.job_cost(j) for j in jobs], 2)
assert all(out[j]["rank"] == r for r, p in enumerate(parts) for j in p)
dist.barrier(); dist.destroy_process_group()
print("ok")
"""


def test_sharding_across_two_ranks_with_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "gloo_shard.py"
    script.write_text(_GLOO_SCRIPT.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    for p in procs:
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0, err[-2000:]
        assert "ok" in out


def test_likelihood_batch_holdout_cells_follow_the_reordering(monkeypatch):
    """likelihood_batch hands every fit its hold-out cells in the row numbering of the fold's training
    problem (which euclidean_embedding-style reordering permutes unless preserve_order).  A stub "fit"
    that returns its initial positions makes the expected residuals computable by hand."""
    m = random_r_matrix(30, 0.6, 11)
    value, code, is_na = core.parse_dissimilarity(m)
    folds = cv.make_folds(m, 3, np.random.default_rng(2))
    G = np.random.default_rng(5).normal(size=(30, 2)) * 4          # "embedding" in the caller's row numbering

    def stub_fit_batch(jobs, device=0):
        out = []
        for j in jobs:
            pos = np.asarray(j["initial_positions"], dtype=float)
            hi, hj, ht = j["holdout"]
            ok = ~np.isnan(ht)
            d = np.linalg.norm(pos[hi] - pos[hj], axis=1)
            out.append(dict(status=_lib.OK, positions=pos, holdout_sum_abs=float(np.abs(ht[ok] - d[ok]).sum()),
                            holdout_count=int(ok.sum()), iterations=1, converged=True))
        return out

    monkeypatch.setattr(_lib, "fit_batch", stub_fit_batch)
    for preserve in (True, False):
        inits, want = [], []
        for h in folds:
            prob, ci, cj, tr = cv._fold_job(value, code, is_na, np.asarray(h), preserve)
            inits.append(G if prob["order"] is None else G[prob["order"]])
            want.append(np.abs(tr - np.linalg.norm(G[ci] - G[cj], axis=1)).mean())
        assert preserve or any(cv._fold_job(value, code, is_na, np.asarray(h), False)[0]["order"] is not None for h in folds)
        got = cv.likelihood_batch(m, [dict(N=2, k0=1.0, cooling_rate=0.01, c_repulsion=0.01)], 5, 1e-4, folds=3,
                                  preserve_order=preserve, fold_indices=folds, init_list=[inits])
        assert [f["Holdout_MAE"] for f in got[0]["folds"]] == pytest.approx(want, rel=1e-12)


# ------------------------------------------------------------------ the .Call shim ---------------
@pytest.fixture(scope="module")
def shim_cpu(tmp_path_factory):
    from conftest import build_r_shim_harness
    return build_r_shim_harness(str(tmp_path_factory.mktemp("shim") / "harness_cpu"), real_library=False)


def _harness(exe, *args):
    out = subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


def _clean(r):
    assert r["protect_depth"] == 0 and r["type_errors"] == 0 and r["rng_violations"] == 0 and r["rng_open"] == 0
    assert r["raw_interrupt_jumps"] == 0


def test_r_shim_registers_both_entries(shim_cpu):
    r = _harness(shim_cpu, "register")
    assert r["routines"] == {"_topolow_optimize_layout_b200": 16, "_topolow_fit_batch_b200": 3}    # src/RcppExports.cpp:41-49
    assert r["dynamic_symbols"] == 0


def test_r_shim_single_call_marshals_every_field(shim_cpu, tmp_path):
    from conftest import write_problem_bin
    init, deg, ei, ej, ed, et = small_problem(40, 3, 0.3, 1)
    write_problem_bin(tmp_path / "p.bin", init, deg, ei, ej, ed, et)
    r = _harness(shim_cpu, "single", tmp_path / "p.bin", tmp_path / "o.bin", 100, 5.0, 0.01, 0.02, 1e-4, 5, 3, "rowblock")
    _clean(r)
    assert r["names"] == ["positions", "converged", "iterations", "final_mae", "final_k"]      # src/optimization.cpp:375-381
    assert r["types"] == [14, 10, 13, 14, 14] and r["dim"] == [40, 3]                          # REALSXP, LGLSXP, INTSXP, REALSXP x 2
    out = np.fromfile(tmp_path / "o.bin")
    # the recording fake answers with functions of what it was handed (integration/r_stub/fake_topolow.c)
    assert out[0] == 1 and out[1] == 100 - 3
    assert out[2] == pytest.approx(float(np.sum(ed * (1 + et) + ei - ej)), rel=1e-12)
    assert out[3] == pytest.approx(5.0 * 0.99 + 0.02 + 1e-4, rel=1e-15)
    pos = out[4:].reshape(3, 40).T
    np.testing.assert_allclose(pos, 2 * init + deg[:, None], rtol=0, atol=0)
    # a second scenario in the default mode: seed drawn inside GetRNGstate / PutRNGstate (two draws of the fixed stream)
    r2 = _harness(shim_cpu, "single", tmp_path / "p.bin", tmp_path / "o2.bin", 50, 1.0, 0.5, 0.0, 0.0, 4, 1)
    _clean(r2)
    assert np.fromfile(tmp_path / "o2.bin")[0] == 0          # convergence_window reached the library as 4


def test_r_shim_errors_and_interrupts_leave_through_the_shim(shim_cpu, tmp_path):
    from conftest import write_problem_bin
    r = _harness(shim_cpu, "too_few")
    assert r["left_by"] == "Rf_error" and r["message"] == "Need at least 2 points for embedding"   # src/optimization.cpp:131
    _clean(r)
    write_problem_bin(tmp_path / "p.bin", *small_problem(30, 2, 0.3, 2))
    r = _harness(shim_cpu, "interrupt", tmp_path / "p.bin", 3)
    # the interrupt is caught inside R_ToplevelExec, reported to the library as a return value, and only after the
    # library has returned TOPOLOW_ERR_INTERRUPTED does the shim raise the R condition - never a jump through it
    assert r["left_by"] == "Rf_onintr" and r["onintr_calls"] == 1 and r["interrupt_checks"] == 3
    _clean(r)
    r = _harness(shim_cpu, "interrupt", tmp_path / "p.bin", 0)      # no interrupt: polled between chunks, returns normally
    assert r["left_by"] == "return" and r["interrupt_checks"] == 4
    _clean(r)


def test_r_shim_batch_shares_fold_arrays_and_reports_per_job_status(shim_cpu, tmp_path):
    from conftest import write_problem_bin
    init, deg, ei, ej, ed, et = small_problem(50, 4, 0.2, 3)
    hold = (np.array([0, 3, 7], np.int32), np.array([5, 9, 11], np.int32), np.array([1.5, 2.5, 0.25]))
    write_problem_bin(tmp_path / "p.bin", init, deg, ei, ej, ed, et, hold)
    r = _harness(shim_cpu, "batch", tmp_path / "p.bin", tmp_path / "o.bin", 5, 250, 4.0, 0.01, 0.02, 1e-4, 5, 3, 1)
    _clean(r)
    assert r["n"] == 5 and r["names"][:8] == ["converged", "iterations", "final_mae", "final_k", "holdout_sum_abs",
                                               "holdout_count", "status", "message"]
    assert "shared=4" in r["jobs"][0]["message"]             # the four other jobs carried the same edge vectors
    assert [j["status"] for j in r["jobs"]] == [0] * 5 and all(j["pos_len"] == 200 for j in r["jobs"])
    out = np.fromfile(tmp_path / "o.bin").reshape(5, 7 + 200)
    for j in range(5):
        assert out[j, 1] == 247 and out[j, 3] == pytest.approx(4.0 * (1 + 0.25 * j) * 0.99 + 0.02 + 1e-4, rel=1e-15)
        assert out[j, 4] == pytest.approx(hold[2].sum() + hold[0].sum() + 2 * hold[1].sum()) and out[j, 5] == 3
    r = _harness(shim_cpu, "batch", tmp_path / "p.bin", tmp_path / "o.bin", 2, 250, 4.0, 0.01, 0.02, 1e-4, 5, 3, 0)
    assert all(j["pos_len"] == 0 for j in r["jobs"])         # positions stay on the device side unless asked for
