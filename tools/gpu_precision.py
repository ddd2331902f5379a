"""FP32 production arithmetic against the exact-FP64 run of the SAME schedule at larger N: does the FP32
position update lose the many tiny far-field repulsion kicks (each below half an ulp of a coordinate)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from topolow_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
miss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.99
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 60
prob = synth.make_problem(n, d, miss, seed=3)
fa = synth.fit_args(prob)
hp = (5.0, 0.01, 0.02, 1e-4, 10**6, 3)
a = _lib.fit(*fa, iters, *hp, precision=_lib.PREC_F64_EXACT, seed=1, tile_points=96, trace=True)
b = _lib.fit(*fa, iters, *hp, precision=_lib.PREC_F32, seed=1, tile_points=96, max_warps=2, trace=True)
ta, tb = a["trace_mae"], b["trace_mae"]
ok = ~np.isnan(ta)
print("n", n, "d", d, "iters", iters, "E", len(prob["edge_i"]), "f64 ms", a["device_ms"], "f32 ms", b["device_ms"])
print("edge MAE trace f64:", np.round(ta[ok][::4], 5))
print("edge MAE trace f32:", np.round(tb[ok][::4], 5))
print("final MAE f64 %.6f f32 %.6f rel diff %.2e" % (a["final_mae"], b["final_mae"], abs(a["final_mae"] - b["final_mae"]) / a["final_mae"]))
pa, pb = a["positions"], b["positions"]
print("coordinate scale %.2f  |f32 - f64| median %.3e  99%% %.3e  max %.3e" % (np.abs(pa).max(), np.median(np.abs(pa - pb)),
      np.quantile(np.abs(pa - pb), 0.99), np.abs(pa - pb).max()))
# mean pairwise distance on a sample (a lost outward repulsion would shrink the map)
rng = np.random.default_rng(0)
i, j = rng.integers(0, n, 200000), rng.integers(0, n, 200000)
da, db = np.linalg.norm(pa[i] - pa[j], axis=1), np.linalg.norm(pb[i] - pb[j], axis=1)
print("mean random-pair distance f64 %.5f f32 %.5f ratio %.6f" % (da.mean(), db.mean(), db.mean() / da.mean()))
