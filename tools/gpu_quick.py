"""Quick GPU check: smoke parity + throughput at cfg3/cfg4 sizes (synthetic generator of bench.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
from tools import synth
from topolow_b200 import _lib
ge.smoke()
sizes = [(335, 5, 0.93, 200), (10000, 10, 0.95, 12), (100000, 16, 0.99, 6)]
if len(sys.argv) > 1:
    sizes = [s for s in sizes if str(s[0]) in sys.argv[1:]]
for (n, d, miss, iters) in sizes:
    prob = synth.make_problem(n, d, miss, seed=0)
    plan = _lib.Plan(*synth.fit_args(prob), iters + 3, 5.0, 0.01, 0.02, convergence_window=10**6)
    plan.run(3)
    ms = plan.run(iters)
    info = plan.info(); r = plan.result(); plan.close()
    print(f"n={n} d={d} W={info['warps_per_cta']} G={info['ctas']} ms/iter={ms/iters:.3f} "
          f"pair-updates/s={info['pairs_per_iter']*iters/(ms*1e-3):.3e} mae={r['final_mae']:.4f}", flush=True)
