"""cProfile of one cfg5 step (296 fits through _lib.fit_batch): where the host time outside the library goes."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from topolow_b200 import _lib
prob, jobs, meta = bench.cfg5_jobs(5000, 0.95, 74, 250, seed=0)
held = {}
for (_s, f, h) in meta:
    held.setdefault(f, (np.ascontiguousarray(prob["edge_i"][h]), np.ascontiguousarray(prob["edge_j"][h]), np.ascontiguousarray(prob["edge_dist"][h])))
jobs = [dict(j, holdout=held[f]) for j, (_s, f, _h) in zip(jobs, meta)]
_lib.fit_batch(jobs[:40])
t0 = time.perf_counter(); out = _lib.fit_batch(jobs); print("step wall", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); out = _lib.fit_batch(jobs); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
