"""CV-grid throughput probe: many independent fits of one synthetic matrix (BASELINE.json configs[4] shape)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from topolow_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_fits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
fixed_d = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.default_rng(0)
prob = synth.make_problem(n, 5, 0.95, seed=0)
jobs = []
for j in range(n_fits):
    d = fixed_d or int(rng.integers(2, 11))
    init = rng.normal(size=(n, d)) * 3
    jobs.append(dict(initial_positions=init, degrees=prob["degrees"], edge_i=prob["edge_i"], edge_j=prob["edge_j"],
                     edge_dist=prob["edge_dist"], edge_thresh=prob["edge_thresh"], n_iter=iters, k0=float(rng.uniform(1, 10)),
                     cooling_rate=float(rng.uniform(0.005, 0.05)), c_repulsion=float(rng.uniform(0.001, 0.05)),
                     convergence_window=10**6, seed=j))
_lib.fit_batch(jobs[:4])
t0 = time.perf_counter()
out = _lib.fit_batch(jobs)
dt = time.perf_counter() - t0
pu = sum(r["pair_updates"] for r in out)
print(f"n={n} fits={n_fits} iters={iters} wall={dt:.2f}s fits/min={n_fits/dt*60:.1f} pair-updates/s={pu/dt:.3e} "
      f"dev_ms min/med/max={min(r['device_ms'] for r in out):.1f}/{np.median([r['device_ms'] for r in out]):.1f}/{max(r['device_ms'] for r in out):.1f}")
