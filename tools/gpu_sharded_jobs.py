"""Per-job kernel times of the sharded schedule (emulated on one GPU): where does an iteration go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import synth
from topolow_b200.sharded import ShardedMap
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
prob = synth.make_problem(100000, 16, 0.99, seed=0)
sm = ShardedMap(*synth.fit_args(prob), 10, 5.0, 0.01, 0.02, 1e-4, 10**6, 3, world_size=world, seed=0, emulate=True, tile_points=tp)
def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
sm.step(1)
row = sm.jobs(1)[0]
print("mega-blocks", sm.M, "x", sm.Tm, "tiles; layout", sm.plan.layout())
print("bipartite job (rank 0 of a round): %.2f ms" % timed(lambda: sm.plan.run_job(*row[0])))
print("bipartite job again: %.2f ms" % timed(lambda: sm.plan.run_job(*row[0])))
print("diag job (one mega-block): %.2f ms" % timed(lambda: sm.plan.run_job(0, 0, sm.Tm)))
print("end phase: %.2f ms" % timed(lambda: sm.plan.end_iteration()))
