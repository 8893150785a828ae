#!/usr/bin/env python3
"""BASELINE configs[4], second half: 64 independent planning scenarios sharded across the GPUs of one box.

    python tools/scenario_sweep.py [--scenarios 64] [--check K]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scenario_sweep.py

Scenario s (seed 100 + s) is a C3-style world -- the C2 map and ribbons, 50 Gaussian obstacles drawn from the seed, a
start state drawn from the seed -- planned by the product's BatchedAStarPlanner on a VIRTUAL clock (0.95 s budget,
--tick seconds per now() call), so every scenario is deterministic.  Scenario s runs on rank s mod N (replicas only: nothing is
exchanged but the 64 final plan costs, one all_gather).  --check K re-plans the first K scenarios of rank 0 with the
reference's CPU planner and requires identical words / counters and 1e-9 costs.
Rank 0 prints one JSON line.

Status (round 1): start-up, world generation, planning and the reference check run on one GPU, but the sweep was NOT
timed to completion this round -- with a 4 ms tick several scenarios double their sample set past 10^5 and the
adapter's host-side heap work dominates (minutes per scenario); the default tick is now 20 ms.  No number from this
tool is quoted anywhere."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def scenario_world(s):
    from path_planner_b200 import synth
    w = synth.world_c2()
    w.name = "scenario-%d" % s
    synth._add_obstacles(w, "gaussian", 50, 300.0, 700.0, 100 + s)
    rng = np.random.default_rng(1000 + s)
    start = np.array([rng.uniform(385, 595), rng.uniform(385, 615), rng.uniform(0, 2 * np.pi), 2.5, 1.0])
    return w, start


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--tick", type=float, default=2e-2, help="virtual seconds per now() call")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from path_planner_b200 import sharding
    from tests import common

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = common.load_harness()
    mine = sharding.scenario_assignment(args.scenarios, rank, world_size)
    costs = torch.full((args.scenarios,), float("nan"), dtype=torch.float64, device=dev)
    expanded = 0
    checked = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k, s in enumerate(mine):
        w, start = scenario_world(s)
        sid = w.upload_ref(lib)
        plan, st = common.run_plan(lib, "harness", sid, start, 0.95, 1000.0, args.tick, 100, device=local_rank)
        costs[s] = st["f"] if len(plan) else float("inf")
        expanded += int(st["expanded"])
        if rank == 0 and k < args.check:
            ref_plan, rs = common.run_plan(lib, "ref", sid, start, 0.95, 1000.0, args.tick, 100)
            assert ref_plan.shape == plan.shape and np.array_equal(ref_plan[:, 7], plan[:, 7]), "scenario %d: plans differ" % s
            assert rs["expanded"] == st["expanded"] and rs["generated"] == st["generated"], "scenario %d: search differs" % s
            assert np.isclose(rs["f"], st["f"], rtol=1e-9, atol=1e-9)
            checked += 1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = torch.tensor([dt, float(expanded)], dtype=torch.float64, device=dev)
    if world_size > 1:
        # each rank filled its own slots (NaN elsewhere): the gather of the 64 plan costs is a max over ranks
        costs = torch.nan_to_num(costs, nan=-float("inf"))
        dist.all_reduce(costs, op=dist.ReduceOp.MAX)
        tmax = stats[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats[1:], op=dist.ReduceOp.SUM)
        stats[0] = tmax[0]
    if rank == 0:
        c = costs.cpu().numpy()
        print(json.dumps({
            "metric": "planning_scenarios_per_sec", "value": args.scenarios / float(stats[0]), "unit": "scenarios/s",
            "n_gpus": world_size, "scenarios": args.scenarios, "wall_s": float(stats[0]), "expanded_total": int(stats[1]),
            "plans_found": int(np.isfinite(c).sum()), "f_mean": float(c[np.isfinite(c)].mean()), "f_checksum": float(np.nansum(np.where(np.isfinite(c), c, 0.0))),
            "checked_against_reference": checked,
            "config": {"workload": "independent C3-style scenarios (seeds 100..), virtual clock 0.95 s / %g s per now()" % args.tick,
                       "parallelism": "scenario s -> rank s mod N, replicas only"}}), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
