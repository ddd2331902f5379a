"""Multi-GPU check of ShardedMap (run under torchrun, one rank per GPU):
  1. a small problem: every rank's result must equal, bit for bit, the single-GPU emulation of the
     same sharded schedule (FP32 and exact FP64) - i.e. the NCCL exchange moves exactly the right blocks;
  2. cfg4-shaped timing: iterations of the N = 100k map across the ranks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from tools import synth
from topolow_b200 import _lib
from topolow_b200.sharded import ShardedMap

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hp = (5.0, 0.01, 0.02, 1e-4, 10**6, 3)

for prec in (_lib.PREC_F64_EXACT, _lib.PREC_F32):
    prob = synth.make_problem(3000, 6, 0.95, seed=1)
    fa = synth.fit_args(prob)
    sm = ShardedMap(*fa, 6, *hp, world_size=world, rank=rank, device=local, precision=prec, seed=3)
    sm.step(6)
    got = sm.result()
    sm.close()
    em = ShardedMap(*fa, 6, *hp, world_size=world, device=local, precision=prec, seed=3, emulate=True)
    em.step(6)
    want = em.result()
    em.close()
    same = np.array_equal(got["positions"], want["positions"]) and got["final_mae"] == want["final_mae"]
    flags = [None] * world
    dist.all_gather_object(flags, bool(same))
    if rank == 0:
        print(f"precision {prec}: NCCL-sharded == emulated on every rank: {all(flags)}  (mae {got['final_mae']:.6f})", flush=True)
    assert same

n, d, miss = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (100000, 16, 0.99)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
prob = synth.make_problem(n, d, miss, seed=0)
sm = ShardedMap(*synth.fit_args(prob), iters + 2, *hp, world_size=world, rank=rank, device=local, seed=0)
sm.step(2)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
sm.step(iters)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
r = sm.result()
if rank == 0:
    P = n * (n - 1) // 2
    print(f"n={n} d={d} ranks={world}: {dt/iters*1e3:.2f} ms/iter, {P*iters/dt:.3e} pair-updates/s, mae {r['final_mae']:.4f}, "
          f"mega-blocks {sm.M} x {sm.Tm} tiles", flush=True)
sm.close()
dist.destroy_process_group()
