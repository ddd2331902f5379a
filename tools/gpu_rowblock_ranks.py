"""torchrun helper: every rank builds its shard of ONE map (row-block mode), the ranks exchange CUDA-IPC
handles and step the map together; rank 0 writes the result for the caller to compare with the
single-rank run (tests/test_gpu_rowblock.py, bench.py use the same RowBlockMap class).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/gpu_rowblock_ranks.py --points 3000 --ndim 5 --iters 12 --out gpurun_out/ranks2.npz
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=3000)
    ap.add_argument("--ndim", type=int, default=5)
    ap.add_argument("--missing", type=float, default=0.9)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tools import synth
    from topolow_b200.rowblock import RowBlockMap

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    prob = synth.make_problem(a.points, a.ndim, a.missing, seed=0)
    fa = synth.fit_args(prob)
    m = RowBlockMap(*fa, a.iters, 5.0, 0.01, 0.02, 1e-4, 5, 3, rank=rank, world_size=world, device=local, seed=3)
    ms = m.step(a.iters)
    res = m.result(trace=True)
    info = m.info()
    m.close()
    if rank == 0:
        print("ranks", world, "ms/iter %.3f" % (ms / max(res["iterations_run"], 1)), "mae", res["final_mae"], "iters", res["iterations_run"],
              "peer bytes/iter", info["peer_store_bytes_per_iteration"], flush=True)
        if a.out:
            np.savez(a.out, positions=res["positions"], final_mae=res["final_mae"], iterations=res["iterations"],
                     iterations_run=res["iterations_run"], trace=res["trace_mae"])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
