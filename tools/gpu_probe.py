"""First-contact GPU script: smoke parity, pipe-rate microbenchmarks, throughput at a few sizes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
from topolow_b200 import _lib

def synth(n, d, density, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, d)) * 3
    E = int(density * n * (n - 1) / 2)
    ei = rng.integers(0, n, size=int(E * 1.1)); ej = rng.integers(0, n, size=int(E * 1.1))
    ok = ei != ej
    a = np.minimum(ei[ok], ej[ok]); b = np.maximum(ei[ok], ej[ok])
    key = np.unique(a.astype(np.int64) * n + b)
    ei = (key // n).astype(np.int32); ej = (key % n).astype(np.int32)
    ed = np.linalg.norm(X[ei] - X[ej], axis=1)
    et = np.zeros(len(ei), np.int32)
    deg = (np.bincount(np.r_[ei, ej], minlength=n) + 1).astype(np.int32)
    init = rng.normal(size=(n, d)) * 3
    return init, deg, ei, ej, ed, et

out = {}
print(_lib.device_info())
try:
    ge.smoke(); out["smoke"] = "ok"
except Exception as e:
    import traceback; traceback.print_exc(); out["smoke"] = repr(e)
names = ["ffma_flops", "ffma2_flops", "dfma_flops", "shfl_warp_inst", "mufu_warp_inst", "copy_bytes"]
for w, nm in enumerate(names):
    try:
        out[nm] = _lib.microbench(w)
    except Exception as e:
        out[nm] = repr(e)
    print(nm, out[nm], flush=True)
for (n, d, dens, iters, prec) in [(335, 5, 0.06, 200, 0), (2000, 5, 0.05, 30, 0), (10000, 10, 0.05, 6, 0),
                                  (30000, 16, 0.01, 3, 0), (100000, 16, 0.01, 2, 0), (10000, 10, 0.05, 3, 1)]:
    try:
        t0 = time.time(); args = synth(n, d, dens); t1 = time.time()
        plan = _lib.Plan(*args, iters + 2, 5.0, 0.01, 0.02, convergence_window=10**6, precision=prec)
        t2 = time.time()
        plan.run(2)
        ms = plan.run(iters)
        info = plan.info(); r = plan.result(); plan.close()
        rate = info["pairs_per_iter"] * iters / (ms * 1e-3)
        print(f"n={n} d={d} prec={prec} E={len(args[2])} geo={info} ms/iter={ms/iters:.3f} pair-updates/s={rate:.3e} "
              f"mae={r['final_mae']:.4f} synth_s={t1-t0:.1f} plan_s={t2-t1:.1f}", flush=True)
        out[f"rate_n{n}_d{d}_p{prec}"] = rate
    except Exception as e:
        import traceback; traceback.print_exc(); out[f"rate_n{n}_d{d}_p{prec}"] = repr(e)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
