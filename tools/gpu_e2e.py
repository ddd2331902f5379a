"""End-to-end timing of one fit through topolow_fit on host buffers (cfg4 shape by default).
TOPOLOW_DEBUG=1 prints the phase times of the native side.  usage: gpu_e2e.py [n d missing iters] [--pinned] [--rowblock]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tools import synth
from topolow_b200 import _lib
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(argv[0]) if len(argv) > 0 else 100000
d = int(argv[1]) if len(argv) > 1 else 16
miss = float(argv[2]) if len(argv) > 2 else 0.99
iters = int(argv[3]) if len(argv) > 3 else 20
prob = synth.make_problem(n, d, miss, seed=0)
fa = list(synth.fit_args(prob))
if "--pinned" in sys.argv:
    import torch
    keep = []
    for i, a in enumerate(fa):
        a = np.asarray(a)
        t = torch.from_numpy(np.ascontiguousarray(a) if a.flags.c_contiguous else np.asfortranarray(a).T.copy()).pin_memory()
        keep.append(t)
        fa[i] = t.numpy() if a.flags.c_contiguous else t.numpy().T
mode = _lib.MODE_ROWBLOCK if "--rowblock" in sys.argv else _lib.MODE_COLOURED
_lib.fit(*fa, 2, 5.0, 0.01, 0.02, 1e-4, 3, 3, mode=mode)   # warm-up (context, module load)
t0 = time.perf_counter()
_pa = _lib.ProblemArrays(*fa)
print(f"ProblemArrays (host marshalling alone): {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
del _pa
for rep in range(2):
    t0 = time.perf_counter()
    r = _lib.fit(*fa, iters, 5.0, 0.01, 0.02, 1e-4, iters + 1, 3, mode=mode)
    wall = time.perf_counter() - t0
    pairs = n * (n - 1) // 2
    print(f"e2e wall {wall*1e3:.1f} ms, device {r['device_ms']:.1f} ms, overhead {wall*1e3 - r['device_ms']:.1f} ms, "
          f"{pairs * r['iterations_run'] / wall:.3e} pair-updates/s", flush=True)
