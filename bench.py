#!/usr/bin/env python
"""bench.py - pair-updates/s of the force-directed embedding loop on B200.

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg4|cfg3|cfg2|small]

A step is ONE iteration of the reference's loop (src/optimization.cpp:193-366): all N(N-1)/2 pair
updates, cooling, and - every convergence_check_freq iterations - the edge MAE + controller.  The
default workload is BASELINE.json configs[3]: synthetic, 100 000 points, 99 % missing, ndim = 16
(tools/synth.py, seed 0), early stopping disabled so that exactly K iterations run
(convergence_counter = n_iter + 1).

  value      whole-job pair-updates/s with the problem resident in HBM (topolow_plan_run), CUDA events
             on the launching stream, max over ranks.
  e2e        the same metric through the reference-facing call with HOST buffers (topolow_fit: upload,
             device set-up, K iterations, download), wall clock around the call.
  roofline   the persistent tile kernel is FP32-FMA bound (no tensor-core work exists on this path):
             achieved = algorithmic flop per iteration / kernel time, peak = FFMA rate measured live by
             topolow_microbench on the same GPU.  `hbm` gives the edge stream against MEASURED_PEAKS.
  cpu_baseline / --impl reference: oracle/ (the CPU restatement of src/optimization.cpp) on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, ndim, missing fraction)
    "cfg4": (100_000, 16, 0.99),
    "cfg3": (10_000, 10, 0.95),
    "small": (2_000, 5, 0.90),
    "cfg5": (5_000, 5, 0.95),     # Euclidify grid: parameter samples x 4 folds, independent fits
}
HYPER = dict(k0=5.0, cooling_rate=0.01, c_repulsion=0.02, relative_epsilon=1e-4, convergence_check_freq=3)


def flop_per_iter(n, d, n_edges):
    """7d+8 per unmeasured pair update, 7d+12 per spring pair update."""
    pairs = n * (n - 1) // 2
    return (pairs - n_edges) * (7 * d + 8) + n_edges * (7 * d + 12)


def edge_bytes_per_iter(n, d, n_edges):
    """Edge stream: 16-byte record per measured pair + every FP32 position read and written once."""
    return n_edges * 16 + 2 * n * d * 4


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._idx = gpu_index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self._idx), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return None
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def cpu_sample(prob, n, d, sample_pairs, kind="O2", seed=0):
    """Time the oracle's pair loop on the same workload.  n <= 12000: one full iteration exactly as the
    reference runs it (dense lookup, std::shuffle of all pairs).  Larger n (the reference cannot run
    these: int overflow and dense n x n inputs, src/optimization.cpp:143-150): a bounded sample of
    uniformly random pairs in explicit order with the sparse lookup; set-up (hash build, final MAE) is
    measured by a zero-pair call and subtracted.  Returns (pair-updates/s, seconds, description)."""
    from oracle import cpu_oracle
    from tools import synth
    fa = synth.fit_args(prob)
    hp = (HYPER["k0"], HYPER["cooling_rate"], HYPER["c_repulsion"], HYPER["relative_epsilon"], 3, 7)
    if n <= 12000:
        t0 = time.perf_counter()
        cpu_oracle.optimize_layout_exact(*fa, 0, *hp, seed=seed, dense=True, kind=kind)
        setup = time.perf_counter() - t0          # dense matrices + all_pairs construction
        t0 = time.perf_counter()
        res = cpu_oracle.optimize_layout_exact(*fa, 2, *hp, seed=seed, dense=True, kind=kind)
        dt = max(time.perf_counter() - t0 - setup, 1e-9)
        return res["visited"] / dt, dt, ("two full iterations, dense lookup + std::shuffle of all pairs (the reference's own "
                                         "scheme); set-up time subtracted")
    rng = np.random.default_rng(seed)
    i = rng.integers(0, n, size=sample_pairs, dtype=np.int32)
    j = rng.integers(0, n - 1, size=sample_pairs, dtype=np.int32)
    j = np.where(j >= i, j + 1, j).astype(np.int32)
    order = np.stack([i, j], axis=1)[None]
    empty = np.full((1, 1, 2), -1, dtype=np.int32)
    t0 = time.perf_counter()
    cpu_oracle.optimize_layout_exact(*fa, 1, *hp, pair_order=empty, dense=False, kind=kind)
    setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = cpu_oracle.optimize_layout_exact(*fa, 1, *hp, pair_order=order, dense=False, kind=kind)
    dt = max(time.perf_counter() - t0 - setup, 1e-9)
    return res["visited"] / dt, dt, (f"{sample_pairs} uniformly random pair visits in explicit order, sparse (hash) lookup; "
                                     "set-up time subtracted")


def run_reference(args, n, d, missing):
    """--impl reference: the CPU implementation of the path on this box's host cores."""
    from tools import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = synth.make_problem(n, d, missing, seed=0)
    pairs = n * (n - 1) // 2
    sample = int(min(pairs, 4e7, 3e8 / max(args.steps, 1)))   # bounded: the whole run stays within a few minutes
    from oracle import cpu_oracle
    cpu_oracle.build(ref=False)
    for w in range(args.warmup):
        cpu_sample(prob, n, d, max(sample // 8, 1000), seed=100 + w)
    visited, busy, desc = 0.0, 0.0, ""
    for k in range(args.steps):
        rate, dt, desc = cpu_sample(prob, n, d, sample, seed=k)
        visited += rate * dt
        busy += dt
    wall = busy
    value = visited / wall
    sample = int(round(visited / args.steps))
    line = {
        "impl": "reference", "metric": "pair-updates/s", "value": value, "unit": "pair-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3 * (pairs / sample),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "n_points": n, "ndim": d, "missing": missing, "n_edges": int(len(prob["edge_i"])),
                   "pairs_per_iteration": pairs},
        "cpu_baseline": {"value": value, "unit": "pair-updates/s", "cores": 1, "kind": "port",
                         "sample": f"per step: {desc} (oracle = CPU restatement of src/optimization.cpp, g++ -O2, "
                                   "1 thread: a fit is single-threaded in the reference); ms_per_step is scaled to a full iteration"},
        "e2e": {"value": value, "unit": "pair-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



def cfg5_jobs(n, missing, samples, fit_iters, seed):
    """`samples` parameter draws x 4 CV folds on one synthetic matrix: the unit of work of
    initial_parameter_optimization (R/adaptive_sampling.R:419-425,516-528)."""
    from tools import synth
    rng = np.random.default_rng(seed)
    prob = synth.make_problem(n, 5, missing, seed=0, thresholds=False)
    ei, ej, ed = prob["edge_i"], prob["edge_j"], prob["edge_dist"]
    E = len(ei)
    folds = []
    perm = rng.permutation(E)
    hold = E // 8            # floor(num_non_NA / (2 * folds)) cells = E/8 pairs in each of the 4 folds... per fold
    for f in range(4):
        mask = np.ones(E, dtype=bool)
        mask[perm[f * hold:(f + 1) * hold]] = False
        deg = (np.bincount(ei[mask], minlength=n) + np.bincount(ej[mask], minlength=n) + 1).astype(np.int32)
        # one training edge list per fold, shared by every parameter sample (topolow_fit_batch then builds
        # its device records once per fold; cv.likelihood_batch hands over its folds the same way)
        train = dict(edge_i=np.ascontiguousarray(ei[mask], dtype=np.int32), edge_j=np.ascontiguousarray(ej[mask], dtype=np.int32),
                     edge_dist=np.ascontiguousarray(ed[mask], dtype=np.float64),
                     edge_thresh=np.ascontiguousarray(prob["edge_thresh"][mask], dtype=np.int32))
        folds.append((train, deg, perm[f * hold:(f + 1) * hold]))
    jobs, meta = [], []
    for s_ in range(samples):
        ndim = int(rng.integers(2, 11))
        k0, cr, c = float(rng.uniform(1, 15)), float(10 ** rng.uniform(-3, -1.3)), float(10 ** rng.uniform(-3, -1.3))
        for f, (train, deg, held) in enumerate(folds):
            init = np.vstack([np.zeros((1, ndim)), np.cumsum(rng.uniform(0, 2 * ed.max() / n, size=(n - 1, ndim)), axis=0)])
            jobs.append(dict(initial_positions=init, degrees=deg, **train, n_iter=fit_iters, k0=k0, cooling_rate=cr, c_repulsion=c,
                             relative_epsilon=1e-4, convergence_window=5, seed=1000 * s_ + f))
            meta.append((s_, f, held))
    return prob, jobs, meta


def main_cfg5(args, n, missing):
    import torch
    import torch.distributed as dist
    from topolow_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prob, jobs, meta = cfg5_jobs(n, missing, args.samples, args.fit_iters, seed=rank)
    warm = [dict(j, n_iter=2) for j in jobs[:: max(1, len(jobs) // 18)]]   # loads every ndim variant of the kernel
    _lib.fit_batch(warm, device=local)
    for _ in range(max(args.warmup - 1, 0)):
        _lib.fit_batch(warm, device=local)
    barrier()
    t0 = time.perf_counter()
    done, pair_updates, dev_ms, held_mae = 0, 0, 0.0, []
    held_cells = {}   # the hold-out cells of a fold (fixed for the whole grid), scored on the device inside the batch
    for (_s, f, held) in meta:
        if f not in held_cells:
            held_cells[f] = (np.ascontiguousarray(prob["edge_i"][held]), np.ascontiguousarray(prob["edge_j"][held]),
                             np.ascontiguousarray(prob["edge_dist"][held]))
    jobs = [dict(j, holdout=held_cells[f]) for j, (_s, f, _h) in zip(jobs, meta)]
    with ClockSampler(local) as clk:
        for _ in range(args.steps):
            # fits + hold-out residuals (R/error_metrics.R:95-114, R/adaptive_sampling.R:2642-2647) in one call
            out = _lib.fit_batch(jobs, device=local)
            held_mae.extend(r["holdout_sum_abs"] / max(r["holdout_count"], 1) for r in out)
            done += len(out)
            pair_updates += sum(r["pair_updates"] for r in out)
            dev_ms = max(dev_ms, max(r["device_ms"] for r in out))
        barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    tot = torch.tensor([done, pair_updates], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    wall = float(t[0])
    fits = float(tot[0])
    cpu = None
    if not args.no_cpu:
        from oracle import cpu_oracle
        from tools import synth
        cpu_oracle.build(ref=False)
        j = jobs[0]
        t1 = time.perf_counter()
        res = cpu_oracle.optimize_layout_exact(j["initial_positions"], j["degrees"], j["edge_i"], j["edge_j"], j["edge_dist"],
                                               j["edge_thresh"], 2, j["k0"], j["cooling_rate"], j["c_repulsion"], 1e-4, 5, 3,
                                               seed=0, dense=True)
        dt = time.perf_counter() - t1
        per_fit_s = dt / 2 * args.fit_iters
        cores = os.cpu_count() or 1
        cpu = {"value": 60.0 / per_fit_s * cores, "unit": "CV embeddings/min", "cores": cores, "kind": "port",
               "sample": f"2 iterations of one fold fit on 1 core ({dt:.1f} s incl. set-up), scaled to {args.fit_iters} iterations "
                         f"and to one fit per core on {cores} cores (mclapply, R/adaptive_sampling.R:645-672); no early stopping assumed"}
    line = {
        "metric": "CV embeddings/min", "value": fits / wall * 60.0, "unit": "CV embeddings/min", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg5: Euclidify grid, {args.samples} parameter samples x 4 folds per GPU per step, {n}-point synthetic "
                               f"matrix ({missing:.0%} missing), ndim sampled in 2..10, mapping_max_iter={args.fit_iters}, early stopping on",
                   "fits_per_step_per_gpu": len(jobs), "parallelism": "independent fits, one CTA per fit, no collective",
                   "l2": "each fit streams its own edge list (about 9 MB per iteration); 128 fits in flight exceed L2"},
        "pair_updates_per_s": float(tot[1]) / wall, "mean_holdout_mae": float(np.mean(held_mae)),
        "clocks": clk.summary(),
        "e2e": {"value": fits / wall * 60.0, "unit": "CV embeddings/min",
                "h2d_bytes_per_step": int(sum(j["initial_positions"].nbytes + len(j["edge_i"]) * 20 + n * 4 for j in jobs)),
                "d2h_bytes_per_step": int(sum(j["initial_positions"].nbytes for j in jobs)),
                "note": "the measured path IS end to end: topolow_fit_batch on host buffers (set-up, upload, fits, download) "
                        "including the hold-out residual kernel of every fit"},
        # per step: one tile_batch_kernel launch per ndim group (one CTA per fit) + one hold-out kernel per fit
        "gpu_launches": int(args.steps * (len({j["initial_positions"].shape[1] for j in jobs}) + len(jobs))),
        "roofline": None, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"],
                    help="cfg4 with --gpus > 1: one map sharded over the GPUs (exact, position all-gather between rounds) "
                         "or one independent replica per GPU")
    ap.add_argument("--samples", type=int, default=74, help="cfg5: parameter samples per GPU per step (x 4 folds)")
    ap.add_argument("--fit-iters", type=int, default=250, help="cfg5: mapping_max_iter of every fit (R/core.R:945)")
    args = ap.parse_args()
    n, d, missing = WORKLOADS[args.workload]
    if args.workload == "cfg5":
        return main_cfg5(args, n, missing)

    if args.impl == "reference":
        run_reference(args, n, d, missing)
        return

    import torch
    import torch.distributed as dist
    from tools import synth
    from topolow_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: topolow_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prec = {"f32": _lib.PREC_F32, "f64": _lib.PREC_F64_EXACT}[args.precision]
    prob = synth.make_problem(n, d, missing, seed=0)
    fa = synth.fit_args(prob)
    E = int(len(prob["edge_i"]))
    pairs = n * (n - 1) // 2
    total_iters = args.steps + (max(args.warmup, 3) if args.warmup else 0)   # every iteration run is inside n_iter
    nw = total_iters + 1  # convergence window > n_iter: no early stop

    sharded = world > 1 and args.multi == "sharded"
    hp = (HYPER["k0"], HYPER["cooling_rate"], HYPER["c_repulsion"], HYPER["relative_epsilon"])
    # ---- device-resident leg -------------------------------------------------------------------
    if sharded:
        from topolow_b200.sharded import ShardedMap
        sm = ShardedMap(*fa, total_iters, *hp, nw, HYPER["convergence_check_freq"], world_size=world, rank=rank,
                        device=local, precision=prec, seed=0)
        info0 = dict(sm.plan.info(), mega_blocks=sm.M, tiles_per_mega_block=sm.Tm)
        sm.step(max(args.warmup, 3) if args.warmup else 0)
        launches_before = sm.plan.info()["launches"]
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            t0 = time.perf_counter()
            ev0.record()                      # jobs and NCCL exchanges all run on torch's current stream
            sm.step(args.steps)
            ev1.record()
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
        ms = ev0.elapsed_time(ev1)
        info1 = sm.plan.info()
        res = sm.result()
        sm.close()
    else:
        plan = _lib.Plan(*fa, total_iters, *hp, nw, HYPER["convergence_check_freq"], precision=prec, seed=0, device=local)
        info0 = plan.info()
        plan.run(max(args.warmup, 3) if args.warmup else 0)
        launches_before = plan.info()["launches"]
        barrier()
        with ClockSampler(local) as clk:
            t0 = time.perf_counter()
            ms = plan.run(args.steps)
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
        info1 = plan.info()
        res = plan.result()
        plan.close()
    launches = info1["launches"] - launches_before
    t = torch.tensor([ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, wall_max = float(t[0]), float(t[1])
    # sharded: ONE map, every unordered pair once per iteration whatever the rank count (strong scaling);
    # replicas: `world` independent maps (weak scaling)
    value = (1 if sharded else world) * pairs * args.steps / (ms_max * 1e-3)

    # ---- end-to-end leg: host buffers through the public call -----------------------------------
    e2e = None
    if not args.no_e2e:
        # the step's inputs wait in pinned host memory (page-locked views handed to the C ABI as plain pointers)
        keep_pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in fa[1:]]
        fa = (fa[0],) + tuple(t.numpy() for t in keep_pinned)
        barrier()
        t0 = time.perf_counter()
        if sharded:
            sm = ShardedMap(*fa, args.steps, *hp, args.steps + 1, HYPER["convergence_check_freq"], world_size=world,
                            rank=rank, device=local, precision=prec, seed=0)
            sm.step(args.steps)
            r2 = sm.result()
            sm.close()
        else:
            r2 = _lib.fit(*fa, args.steps, *hp, args.steps + 1, HYPER["convergence_check_freq"], precision=prec, seed=0,
                          device=local)
        barrier()
        wall = time.perf_counter() - t0
        tt = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        h2d = n * d * 8 + n * 4 + E * (4 + 4 + 8 + 4)
        d2h = n * d * 8
        e2e = {"value": (1 if sharded else world) * pairs * r2["iterations_run"] / float(tt[0]), "unit": "pair-updates/s",
               "h2d_bytes_per_step": h2d // max(args.steps, 1), "d2h_bytes_per_step": d2h // max(args.steps, 1),
               "note": "one fit of K iterations from pinned host buffers (per rank when sharded): upload, bucket build, "
                       "K iterations, download; bytes are the call's totals divided by K",
               "wall_s": float(tt[0]), "kernel_ms": r2["device_ms"]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline ---------------------------------------------------------------------------------
    peaks, peak_src = measured_peaks()
    ffma_peak = _lib.microbench(0, local)      # flop/s, measured now on this GPU
    flop = flop_per_iter(n, d, E)
    kernel_s = ms_max * 1e-3 / args.steps        # device time of one iteration (all of it is tile_kernel launches)
    ach = flop / kernel_s / (world if sharded else 1)   # per GPU
    roofline = {"bound": "fp32", "kernel": "tile_kernel<D,%s>" % ("FastF32" if prec == 0 else "ExactF64"),
                "achieved": ach / 1e12, "peak": ffma_peak / 1e12, "unit": "TFLOP/s", "frac": ach / ffma_peak,
                "peak_source": "topolow_microbench FFMA, measured live on this GPU (not in MEASURED_PEAKS.json)",
                "flop_per_iteration": flop,
                # dram__bytes_read + dram__bytes_write of one launch (= one iteration) of this kernel on this
                # workload, from the ncu --set full capture summarised in profiles/r1_ncu_tile_kernel_cfg4.md
                "traffic": 878.6e6 if (args.workload == "cfg4" and prec == 0) else None,
                "traffic_unit": "bytes per launch (one iteration); algorithmic edge stream is bytes_per_iteration below",
                "hbm": {"achieved": edge_bytes_per_iter(n, d, E) / kernel_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": edge_bytes_per_iter(n, d, E) / kernel_s / 1e9 / peaks["hbm_gbs"], "peak_source": peak_src,
                        "bytes_per_iteration": edge_bytes_per_iter(n, d, E)}}

    cpu = None
    if not args.no_cpu:
        from oracle import cpu_oracle
        cpu_oracle.build(ref=False)
        sample = int(min(pairs, 2e8))
        rate, dt, desc = cpu_sample(prob, n, d, sample)
        cpu = {"value": rate, "unit": "pair-updates/s", "cores": 1, "kind": "port",
               "sample": f"{desc}; {dt:.1f} s of CPU work (oracle = CPU restatement of src/optimization.cpp, g++ -O2, 1 thread)"}

    line = {
        "metric": "pair-updates/s", "value": value, "unit": "pair-updates/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak" if (world > 1 and not sharded) else "strong", "vs_baseline": None,
        "dtype": "f32" if prec == 0 else "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: synthetic low-rank, {n} points, {missing:.0%} missing, ndim={d}",
                   "n_points": n, "ndim": d, "n_edges": E, "pairs_per_iteration": pairs, "schedule": info0,
                   "parallelism": "1 GPU" if world == 1 else (
                       f"one map row-sharded over {world} GPUs: tournament over {2 * world} mega-blocks, NCCL all-gather of the "
                       "changed position blocks after each of its rounds, exact sequential semantics" if sharded
                       else f"{world} independent replicas (no collective)"),
                   "l2": "edge stream (%.0f MB/iteration) exceeds L2; the %.1f MB position array is the resident working set"
                         % (E * 16 / 1e6, n * d * 4 / 1e6),
                   "early_stop": "disabled (convergence_counter = n_iter + 1)"},
        "wall_ms_per_step": wall_max / args.steps,
        "final_mae": res["final_mae"],
        "clocks": clk.summary(),
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
