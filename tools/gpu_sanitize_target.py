"""Small run of every kernel family for compute-sanitizer (tools/gpu_sanitize.sh): the coloured schedule with several CTAs
(warp-to-warp tile hand-off through shared-memory flags, per-super-block release / acquire counters, the grid barrier),
the many-fits kernel, replay, the row-block kernels with two shards in lock step (peer stores + epoch flags), the
components kernels.  Sizes are tiny: the tools slow kernels down by one to two orders of magnitude."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import small_problem
from topolow_b200 import _lib, rowblock

which = sys.argv[1] if len(sys.argv) > 1 else "all"
hp = (5.0, 0.02, 0.02, 1e-4, 5, 2)
if which in ("all", "coloured"):
    args = small_problem(700, 3, 0.05, 1)
    for tp in (32, 64, 96):
        r = _lib.fit(*args, 6, *hp, seed=1, tile_points=tp)
        print("coloured f32 tile", tp, r["final_mae"])
    r = _lib.fit(*args, 4, *hp, seed=1, precision=_lib.PREC_F64_EXACT, tile_points=32)
    print("coloured f64", r["final_mae"])
if which in ("all", "batch"):
    a = small_problem(200, 4, 0.1, 2)
    jobs = [dict(initial_positions=a[0], degrees=a[1], edge_i=a[2], edge_j=a[3], edge_dist=a[4], edge_thresh=a[5], n_iter=8, k0=3.0 + j,
                 cooling_rate=0.02, c_repulsion=0.01, seed=j, holdout=(a[2][:20], a[3][:20], a[4][:20])) for j in range(18)]
    print("batch", sum(r["final_mae"] for r in _lib.fit_batch(jobs)))
if which in ("all", "replay"):
    a = small_problem(60, 2, 0.3, 3)
    print("replay", _lib.fit(*a, 5, *hp, mode=_lib.MODE_REPLAY, seed=3)["final_mae"])
if which in ("all", "rowblock"):
    a = small_problem(600, 5, 0.08, 4)
    print("rowblock", _lib.fit(*a, 6, *hp, mode=_lib.MODE_ROWBLOCK, seed=4)["final_mae"])
    ls = rowblock.LocalShards(*a, 6, *hp, n_ranks=2, seed=4)
    ls.run(6)
    print("rowblock 2 shards", ls.result(rank=1)["final_mae"])
    ls.close()
if which in ("all", "graph"):
    a = small_problem(300, 2, 0.02, 5)
    print("components", _lib.components(300, a[2], a[3], np.random.default_rng(0).random((3, 300)) < 0.7))
