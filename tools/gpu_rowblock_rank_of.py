"""Timing of ONE rank of an n_ranks-way sharded map on a single GPU, its peers ignored (TOPOLOW_IGNORE_PEERS=1:
waits return at once, peer stores go to replicas that nobody advances).  The numbers are the rank's kernel times
in the shape it has on a real box; the positions are meaningless.
  TOPOLOW_IGNORE_PEERS=1 python tools/gpu_rowblock_rank_of.py [n_ranks] [n] [ndim] [missing] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from tools import synth
from topolow_b200 import rowblock

assert os.environ.get("TOPOLOW_IGNORE_PEERS"), "set TOPOLOW_IGNORE_PEERS=1"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 16
miss = float(sys.argv[4]) if len(sys.argv) > 4 else 0.99
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 30
cache = "/tmp/rowblock_quick_%d_%d_%g.npz" % (n, d, miss)
if os.path.exists(cache):
    z = np.load(cache); prob = {k: z[k] for k in z.files}
else:
    prob = synth.make_problem(n, d, miss, seed=0); np.savez(cache, **prob)
fa = synth.fit_args(prob)
ls = rowblock.LocalShards(*fa, 6 + 6 + iters, 5.0, 0.01, 0.02, 1e-4, 10 ** 6, 3, n_ranks=G, seed=0)
sh = ls.shards[0]
sh.run(6)
ms = sh.run(iters)
print("rank 0 of %d: ms/iter %.3f" % (G, ms / iters), flush=True)
print("kernels (in order, alone):", {k: round(v, 4) if isinstance(v, float) else v for k, v in sh.time_kernels(6).items()})
ls.close()
