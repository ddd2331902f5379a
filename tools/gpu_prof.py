"""Profiling target: a few iterations of the cfg4-shaped workload (or a smaller one) and nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth
from topolow_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
miss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.99
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
max_ctas = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tile_points = int(sys.argv[6]) if len(sys.argv) > 6 else 0
prob = synth.make_problem(n, d, miss, seed=0)
plan = _lib.Plan(*synth.fit_args(prob), iters + 3, 5.0, 0.01, 0.02, convergence_window=10**6, max_ctas=max_ctas, tile_points=tile_points)
plan.run(3)
ms = plan.run(iters)
info = plan.info()
print(f"n={n} d={d} ms/iter={ms/iters:.3f} pair-updates/s={info['pairs_per_iter']*iters/(ms*1e-3):.3e} {info}")
plan.close()
