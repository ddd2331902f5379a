"""Golden numbers of the REFERENCE LOOP (oracle/topolow_oracle.cpp = src/optimization.cpp:108-382 restated) on
BASELINE.json configs[2]: synthetic 10 000 points, 95 % missing, ndim 10, thresholds on, 10 % of the exact cells
held out, 100 iterations, one run per shuffle seed.  Each run is ~8 minutes of one CPU core (5e9 pair visits on a
dense 10k x 10k lookup), which is why the numbers are committed instead of recomputed on the GPU box:

    python tools/make_golden_cfg3.py            # writes tests/golden/cfg3_reference_loop.json

tests/test_gpu_rowblock.py and tests/test_gpu_parity.py regenerate the same problem (tools/synth.py, seed 3) and
compare the GPU modes with these numbers."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N, D, MISSING, ITERS, SEEDS = 10_000, 10, 0.95, 100, (0, 1, 2)
HP = (5.0, 0.01, 0.02, 1e-4, ITERS + 1, 3)     # no early stop: every seed runs the same 100 iterations


def problem():
    from tools import synth
    prob = synth.make_problem(N, D, MISSING, seed=3)
    ei, ej, ed, et = prob["edge_i"], prob["edge_j"], prob["edge_dist"], prob["edge_thresh"]
    held = (np.random.default_rng(11).random(len(ei)) < 0.10) & (et == 0)
    tr = ~held
    deg = (np.bincount(ei[tr], minlength=N) + np.bincount(ej[tr], minlength=N) + 1).astype(np.int32)
    train = (prob["initial_positions"], deg, ei[tr], ej[tr], ed[tr], et[tr])
    return train, (ei[held], ej[held], ed[held]), (ei, ej, ed, et)


def run(seed):
    from oracle import cpu_oracle
    train, held, _all = problem()
    t0 = time.time()
    r = cpu_oracle.optimize_layout_exact(*train, ITERS, *HP, seed=seed, trace=True)
    dist = np.linalg.norm(r["positions"][held[0]] - r["positions"][held[1]], axis=1)
    tr = r["trace_mae"]
    return {"seed": seed, "final_mae": r["final_mae"], "best_iteration": r["iterations"],
            "heldout_mae": float(np.abs(held[2] - dist).mean()), "seconds": time.time() - t0,
            "mae_trace": [float(x) for x in tr[~np.isnan(tr)]]}


if __name__ == "__main__":
    with mp.get_context("fork").Pool(len(SEEDS)) as pool:
        out = pool.map(run, SEEDS)
    train, held, _ = problem()
    doc = {"what": __doc__.split("\n\n")[0], "n": N, "ndim": D, "missing": MISSING, "iterations": ITERS, "synth_seed": 3,
           "holdout_seed": 11, "train_edges": int(len(train[2])), "heldout_cells": int(len(held[0])),
           "hyper": {"k0": HP[0], "cooling_rate": HP[1], "c_repulsion": HP[2], "relative_epsilon": HP[3],
                     "convergence_counter": HP[4], "convergence_check_freq": HP[5]}, "runs": out}
    with open(os.path.join(ROOT, "tests", "golden", "cfg3_reference_loop.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc)[:600])
