"""Generate tests/golden/*.npz.  Runs in the BUILD container only (needs /root/reference).

Two kinds of fixtures:
  * fixtures/h3n2.npz, fixtures/hiv.npz - the edge-list form of BASELINE.json configs[0] and [1]:
    /root/reference/data-raw/Smith2004-data.csv through a restatement of process_antigenic_data
    (R/data_preprocessing.R:488-684: log2(titer/10), Smith distance = per-serum max - value, '<' titers
    become '>' distances, repeated (virusStrain, serumStrain) pairs averaged, V/ and S/ prefixes, points
    sorted by year) and /root/reference/data-raw/hiv_filtered_long_data.csv (`distance` column, the matrix
    the paper maps, inst/examples/methods-comparison-h3n2-hiv-denv.Rmd:472-479) through
    titers_list_to_matrix (R/data_preprocessing.R:743-844).
  * golden_*.npz - inputs + outputs of the REFERENCE'S OWN source file
    (/root/reference/src/optimization.cpp compiled against oracle/ref_shim with std::random_device forced
    to a seed, oracle/_ref/libtopolow_ref.so).  tests/test_oracle.py requires oracle/topolow_oracle.cpp to
    reproduce them bit for bit; the GPU replay test requires the same of the CUDA path.
"""
from __future__ import annotations

import csv
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _matrix_to_edges(names, cells):
    """cells: dict[(i, j)] -> (value, code) with i < j in sorted-name index space."""
    n = len(names)
    keys = sorted(cells, key=lambda ij: (ij[1], ij[0]))  # R's which(arr.ind): by column, then row
    ei = np.array([k[0] for k in keys], dtype=np.int32)
    ej = np.array([k[1] for k in keys], dtype=np.int32)
    ed = np.array([cells[k][0] for k in keys], dtype=np.float64)
    et = np.array([cells[k][1] for k in keys], dtype=np.int32)
    deg = (np.bincount(ei, minlength=n) + np.bincount(ej, minlength=n) + 1).astype(np.int32)  # + diagonal 0
    return dict(n=n, names=np.array(names), edge_i=ei, edge_j=ej, edge_dist=ed, edge_thresh=et, degrees=deg)


def h3n2_problem():
    rows = list(csv.DictReader(open(os.path.join(REF, "data-raw", "Smith2004-data.csv"), encoding="utf-8-sig")))
    rec = []
    for r in rows:
        t = r["titer"].strip()
        if not t or t[0] not in "0123456789<>":
            continue
        sign = t[0] if t[0] in "<>" else ""
        val = math.log2(float(t.lstrip("<>")) / 10.0)
        rec.append((r["virusStrain"], r["serumStrain"], sign, val, int(r["virusYear"]), int(r["serumYear"])))
    smax = {}
    for v, s, sign, val, vy, sy in rec:
        smax[s] = max(smax.get(s, -1e300), val)
    groups = {}
    for v, s, sign, val, vy, sy in rec:
        dist = smax[s] - val
        dsign = {"<": ">", ">": "<", "": ""}[sign]   # a '<' titer is a '>' distance
        groups.setdefault((v, s), []).append((dsign, dist, vy, sy))
    years = {}
    cells_by_name = {}
    for (v, s), lst in groups.items():
        signs = [x[0] for x in lst if x[0]]
        sign = ("<" if "<" in signs else ">") if signs else ""
        mean = float(np.mean([x[1] for x in lst]))
        vn, sn = "V/" + v, "S/" + s
        years[vn] = min(years.get(vn, 9999), lst[0][2])
        years[sn] = min(years.get(sn, 9999), lst[0][3])
        cells_by_name[(vn, sn)] = (mean, {">": 1, "<": -1, "": 0}[sign])
    names = sorted(years, key=lambda k: (years[k], k))   # order(ranks) over the sorted names: stable
    idx = {k: i for i, k in enumerate(names)}
    cells = {}
    for (a, b), vc in cells_by_name.items():
        i, j = sorted((idx[a], idx[b]))
        cells[(i, j)] = vc
    return _matrix_to_edges(names, cells)


def hiv_problem():
    rows = list(csv.DictReader(open(os.path.join(REF, "data-raw", "hiv_filtered_long_data.csv"))))
    names = sorted({"V/" + r["Virus"] for r in rows} | {"S/" + r["Antibody"] for r in rows})
    idx = {k: i for i, k in enumerate(names)}
    cells = {}
    for r in rows:
        d = r["distance"].strip()
        code = 1 if d.startswith(">") else (-1 if d.startswith("<") else 0)
        i, j = sorted((idx["V/" + r["Virus"]], idx["S/" + r["Antibody"]]))
        cells[(i, j)] = (float(d.lstrip("<>")), code)   # later rows overwrite, as the R loop does
    return _matrix_to_edges(names, cells)


def init_positions(n, ndim, max_d, seed):
    rng = np.random.default_rng(seed)
    step = max_d / n
    return np.vstack([np.zeros((1, ndim)), np.cumsum(rng.uniform(0, 2 * step, size=(n - 1, ndim)), axis=0)])


def golden_case(name, prob, ndim, n_iter, k0, cooling, c_rep, seed, eps=1e-4, window=5, freq=3):
    from oracle import cpu_oracle
    exact = prob["edge_dist"][prob["edge_thresh"] == 0]
    init = init_positions(prob["n"], ndim, float(exact.max() if len(exact) else prob["edge_dist"].max()), seed)
    res = cpu_oracle.ref_optimize_layout_exact(init, prob["degrees"], prob["edge_i"], prob["edge_j"], prob["edge_dist"],
                                               prob["edge_thresh"], n_iter, k0, cooling, c_rep, eps, window, freq, seed=seed)
    np.savez_compressed(
        os.path.join(OUT, f"golden_{name}.npz"), initial_positions=init, degrees=prob["degrees"], edge_i=prob["edge_i"],
        edge_j=prob["edge_j"], edge_dist=prob["edge_dist"], edge_thresh=prob["edge_thresh"],
        params=np.array([n_iter, k0, cooling, c_rep, eps, window, freq, seed], dtype=np.float64),
        positions=res["positions"], converged=res["converged"], iterations=res["iterations"],
        final_mae=res["final_mae"], final_k=res["final_k"])
    print(f"golden_{name}: n={prob['n']} E={len(prob['edge_i'])} iters={res['iterations']} conv={res['converged']} "
          f"mae={res['final_mae']:.6f}")


def small_problem(n, density, seed, thresholds=True):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, 3)) * 3
    iu = np.triu_indices(n, 1)
    keep = rng.random(len(iu[0])) < density
    keep[rng.integers(0, len(keep))] = True
    ei, ej = iu[0][keep], iu[1][keep]
    order = np.lexsort((ei, ej))
    ei, ej = ei[order].astype(np.int32), ej[order].astype(np.int32)
    ed = np.linalg.norm(X[ei] - X[ej], axis=1)
    et = (rng.choice([0, 0, 0, 1, -1], size=len(ei)) if thresholds else np.zeros(len(ei))).astype(np.int32)
    deg = (np.bincount(ei, minlength=n) + np.bincount(ej, minlength=n) + 1).astype(np.int32)
    return dict(n=n, edge_i=ei, edge_j=ej, edge_dist=ed, edge_thresh=et, degrees=deg)


def main():
    from oracle import cpu_oracle
    cpu_oracle.build(ref=True)
    assert cpu_oracle.have_ref(), "oracle/_ref is needed to generate golden vectors"
    os.makedirs(os.path.join(OUT, "fixtures"), exist_ok=True)
    h3 = h3n2_problem()
    hv = hiv_problem()
    for nm, p in (("h3n2", h3), ("hiv", hv)):
        np.savez_compressed(os.path.join(OUT, "fixtures", nm + ".npz"), **p)
        print(nm, "n =", p["n"], "E =", len(p["edge_i"]), "thresholds:", np.bincount(p["edge_thresh"] + 1))
    # the reference's own triangle test matrix (tests/testthat/test-core.R:108-109): (1, 2, 1)
    tri = dict(n=3, edge_i=np.array([0, 0, 1], np.int32), edge_j=np.array([1, 2, 2], np.int32),
               edge_dist=np.array([1.0, 2.0, 1.0]), edge_thresh=np.zeros(3, np.int32), degrees=np.array([3, 3, 3], np.int32))
    golden_case("triangle", tri, 2, 10, 1.0, 0.01, 0.01, seed=11)
    golden_case("small_thresholds", small_problem(60, 0.3, 1), 3, 200, 5.0, 0.01, 0.02, seed=7)
    golden_case("small_sparse", small_problem(97, 0.05, 2, thresholds=False), 2, 120, 2.0, 0.02, 0.05, seed=3)
    # configs[0] / configs[1] with the published hyper-parameters at the config's ndim = 5
    golden_case("h3n2_ndim5", h3, 5, 60, 14.76214, 0.03641074, 0.002943064, seed=1)
    golden_case("hiv_ndim5", hv, 5, 60, 3.550036, 0.04130713, 0.0007038619, seed=2)


if __name__ == "__main__":
    main()
