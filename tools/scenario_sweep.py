#!/usr/bin/env python3
"""BASELINE configs[4], second half: 64 independent planning scenarios sharded across the GPUs of one box.

    python tools/scenario_sweep.py [--scenarios 64] [--check K] [--tick T]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scenario_sweep.py

Scenario s (seed 100 + s) is a C3-style world -- the C2 map and ribbons, 50 Gaussian obstacles drawn from the seed, a
start state drawn from the seed -- planned by the product's standalone harness (path_planner_b200/libppe_harness.so:
BatchedAStarPlanner on this rank's GPU) on a VIRTUAL clock (0.95 s budget, --tick seconds per now() call), so every scenario
is deterministic.  Scenario s runs on rank s mod N (replicas only: nothing is exchanged but the final plan costs).
--check K re-plans the first K scenarios of rank 0 with the reference's own CPU planner (oracle/_ref/libref_planner.so) and
requires identical words / counters and 1e-9 costs.  The same sweep is part of bench.py's default line ("scenarios").
Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--tick", type=float, default=4e-3, help="virtual seconds per now() call")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    from path_planner_b200 import harness as ph
    from path_planner_b200 import sharding

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
    checked = 0
    if rank == 0 and args.check > 0:
        from tests import common
        ref = common.load_ref()
        h = ph.PlanningHarness(local_rank)
        for s in sharding.scenario_assignment(args.scenarios, 0, world_size)[: args.check]:
            w, start = bench.scenario_world(s)
            h.set_world(w)
            plan, st = h.plan(start, 0.95, clock0=1000.0, tick=args.tick, sample_tick=bench.SCENARIO_SAMPLE_TICK)
            sid = w.upload_ref(ref)
            ref_plan, rs = common.run_plan(ref, "ref", sid, start, 0.95, 1000.0, args.tick, 100, sample_tick=bench.SCENARIO_SAMPLE_TICK)
            assert len(ref_plan) == len(plan) and np.array_equal(ref_plan[:, 7], plan["type"]), "scenario %d: plans differ" % s
            assert rs["expanded"] == st["expanded"] and rs["generated"] == st["generated"] and rs["samples"] == st["samples"], \
                "scenario %d: search differs" % s
            assert np.isclose(rs["f"], st["plan_f"], rtol=1e-9, atol=1e-9), "scenario %d: plan cost differs" % s
            checked += 1
        h.close()
    if world_size > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall, expanded, mine = bench.scenario_sweep(args.scenarios, rank, world_size, local_rank, args.tick)
    costs = torch.full((args.scenarios,), float("-inf"), dtype=torch.float64, device=dev)
    for s_, f_ in mine:
        costs[s_] = f_
    stats = torch.tensor([wall, float(expanded)], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(costs, op=dist.ReduceOp.MAX)
        tmax = stats[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats[1:], op=dist.ReduceOp.SUM)
        stats[0] = tmax[0]
    if rank == 0:
        c = costs.cpu().numpy()
        fin = np.isfinite(c)
        print(json.dumps({
            "metric": "planning_scenarios_per_sec", "value": args.scenarios / float(stats[0]), "unit": "scenarios/s",
            "n_gpus": world_size, "scenarios": args.scenarios, "wall_s": float(stats[0]), "expanded_total": int(stats[1]),
            "plans_found": int(fin.sum()), "f_mean": float(c[fin].mean()) if fin.any() else None, "f_checksum": float(c[fin].sum()),
            "checked_against_reference": checked,
            "config": {"workload": "independent C3-style scenarios (seeds 100..), virtual clock 0.95 s / %g s per now()" % args.tick,
                       "parallelism": "scenario s -> rank s mod N, replicas only"}}), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
