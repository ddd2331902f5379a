"""BASELINE.json configs[0] / configs[1]: the bundled H3N2 / HIV matrices, ndim = 5, 1000 iterations:
replay mode (bit-exact) and coloured mode against the CPU oracle, with timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cpu_oracle
from topolow_b200 import _lib
HP = {"h3n2": (14.76214, 0.03641074, 0.002943064), "hiv": (3.550036, 0.04130713, 0.0007038619)}
for name in ("h3n2", "hiv"):
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "fixtures", name + ".npz"))
    n = int(z["n"])
    rng = np.random.default_rng(1)
    exact = z["edge_dist"][z["edge_thresh"] == 0]
    init = np.vstack([np.zeros((1, 5)), np.cumsum(rng.uniform(0, 2 * exact.max() / n, size=(n - 1, 5)), axis=0)])
    args = (init, z["degrees"], z["edge_i"], z["edge_j"], z["edge_dist"], z["edge_thresh"], 1000, *HP[name], 1e-4, 1001, 3)
    t0 = time.perf_counter(); cpu = cpu_oracle.optimize_layout_exact(*args, seed=7); t_cpu = time.perf_counter() - t0
    _lib.fit(*args[:6], 5, *args[7:], mode=_lib.MODE_REPLAY, seed=7)
    t0 = time.perf_counter(); rep = _lib.fit(*args, mode=_lib.MODE_REPLAY, seed=7); t_rep = time.perf_counter() - t0
    _lib.fit(*args[:6], 5, *args[7:], seed=7)
    t0 = time.perf_counter(); col = _lib.fit(*args, seed=7); t_col = time.perf_counter() - t0
    P = n * (n - 1) // 2
    print(f"{name}: n={n} E={len(z['edge_i'])} 1000 iterations: CPU oracle {t_cpu:.2f} s ({P*1000/t_cpu:.2e} pu/s) | "
          f"replay {t_rep:.2f} s wall, {rep['device_ms']/1e3:.2f} s device, bit-exact={np.array_equal(rep['positions'], cpu['positions'])} | "
          f"coloured FP32 {t_col:.3f} s wall, {col['device_ms']/1e3:.3f} s device ({P*1000/t_col:.2e} pu/s), "
          f"MAE cpu {cpu['final_mae']:.4f} replay {rep['final_mae']:.4f} coloured {col['final_mae']:.4f}")
