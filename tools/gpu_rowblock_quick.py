"""Quick timing of the row-block mode on one GPU: python tools/gpu_rowblock_quick.py [n] [ndim] [missing] [iters]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from tools import synth
from topolow_b200 import _lib, rowblock

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
miss = float(sys.argv[3]) if len(sys.argv) > 3 else 0.99
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 12
t0 = time.time()
cache = "/tmp/rowblock_quick_%d_%d_%g.npz" % (n, d, miss)
if os.path.exists(cache):
    z = np.load(cache)
    prob = {k: z[k] for k in z.files}
else:
    prob = synth.make_problem(n, d, miss, seed=0)
    np.savez(cache, **prob)
fa = synth.fit_args(prob)
print("problem %.1fs  E=%d" % (time.time() - t0, len(prob["edge_i"])), flush=True)
if os.environ.get("QUICK_TWICE"):      # a first shard pays for module loading and pool growth; time the second one
    rowblock.Shard(*fa, 3, 5.0, 0.01, 0.02, 1e-4, 10 ** 6, 3, seed=0).close()
    print("---- second create ----", flush=True)
t0 = time.time()
sh = rowblock.Shard(*fa, int(os.environ.get('QUICK_WARM', '3')) + 6 + iters, 5.0, 0.01, 0.02, 1e-4, 10 ** 6, 3, seed=0)
print("create %.2fs" % (time.time() - t0), sh.info(), flush=True)
sh.run(int(os.environ.get('QUICK_WARM', '3')))
tk = sh.time_kernels(6)
print("kernels ms:", {k: round(v, 4) if isinstance(v, float) else v for k, v in tk.items()}, flush=True)
ms = sh.run(iters)
res = sh.result(trace=True)
pairs = n * (n - 1) // 2
print("ms/iter %.3f  pair-updates/s %.3e  mae %.5f  iters %d" % (ms / iters, pairs * iters / (ms * 1e-3), res["final_mae"], res["iterations_run"]))
print("trace", np.round(res["trace_mae"][~np.isnan(res["trace_mae"])], 4))
flop = (7 * d + 8) * pairs
print("algorithmic TFLOP/s (SURVEY 8d: 7d+8 per pair): %.2f;  repulse kernel alone: %.2f" % (flop / (ms / iters * 1e-3) / 1e12, flop / (tk["repulse"] * 1e-3) / 1e12))
sh.close()
