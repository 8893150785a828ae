#!/usr/bin/env python3
"""bench.py -- Dubins edge true-cost evaluations per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c3b|c4|c1] [--edges E] [--scaling strong|weak]
    python bench.py --impl reference ...      # the reference's own CPU implementation of the path

One "step" = one pass of the hot path (Edge::computeTrueCost incl. the Dubins solve, Edge.cpp:68-206)
over one batch of E synthetic edges.  Default workload: BASELINE.json configs[4], the 1M-edge batch config --
2^20 edges on the 4096^2 map with 100 ribbons and 50 Gaussian obstacles (C5).  With N GPUs that ONE batch is
split into N contiguous shards (strong scaling, path_planner_b200.sharding.shard_range) and the per-shard best
records are all-gathered over NCCL; --scaling weak gives every GPU its own E-edge batch instead.  At N = 1 the
line also carries the other configs (C2 = configs[1], C3, C4) under "workloads", the 64-scenario sweep of
configs[4] under "scenarios" and the plan-cost comparison under "plan".

  value     edges/s, whole job, inputs already resident in HBM, timed with CUDA events on the
            launching stream (max over ranks).
  e2e       the same metric through the public host-buffer call (EdgeEngine.true_cost_batch ->
            ppe_true_cost_batch): pinned host inputs, H2D + kernels + D2H inside the timed region.
  roofline  the launch group of one step against the MEASURED fp64 FMA peak of this GPU (tensor cores are not
            used; HBM is not the bound -- its fraction is reported too); fp64_pipe_pct / traffic come from the
            committed ncu captures of this very command (profiles/r02_ncu_metrics.json).
  cpu_baseline  the compiled reference (oracle/_ref/libref_planner.so, kind "reference") or the C
            restatement (kind "port") single-threaded on this host, on a bounded sample.
  dubins    secondary line: K1 Dubins solves per second (HBM-resident), BASELINE metric part (ii).
  plan      BASELINE metric part (iii), "plan cost at 1 s budget": the reference's AStarPlanner (CPU,
            oracle/_ref/libref_planner.so) and the product's standalone harness (path_planner_b200/libppe_harness.so:
            BatchedAStarPlanner on this GPU) each get the same world, start state and a REAL 1.0 s wall-clock
            budget; f-value of the returned plan (lower is better), expansions and expansions/s of both.
"""
import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    "c1": "C1: single ribbon, empty map, no dynamic obstacles",
    "c2": "C2: 1 km^2 grid map (1000x1000 @ 1 m, 40 static rectangles), 10 survey ribbons, no dynamic obstacles",
    "c3": "C3: C2 + 50 Gaussian dynamic obstacles",
    "c3b": "C3b: C2 + 50 binary (10 m x 30 m) dynamic obstacles",
    "c4": "C4: 4096^2 occupancy map, 100 ribbons",
    "c5": "C5: 4096^2 occupancy map, 100 ribbons, 50 Gaussian dynamic obstacles",
}
F_OBS = {"none": 0, "binary": 18, "gaussian": 51}  # SURVEY.md section 8d flop weights


def algorithmic_flops(n_edges, sum_samples, sum_checkpoints, n_ribbons, n_obs, obs_kind):
    """F_edge = 1451 + 43 R + S (111 + N F_obs) + C R 180  (SURVEY.md section 8d), summed over a batch."""
    return (n_edges * (1451.0 + 43.0 * n_ribbons) + sum_samples * (111.0 + n_obs * F_OBS[obs_kind])
            + sum_checkpoints * n_ribbons * 180.0)


def algorithmic_bytes(n_edges, sum_samples):
    """208 B in + 120 B out per edge as the reference's structs hold them + S/8 bitmap bytes (8d)."""
    return n_edges * (208.0 + 120.0) + sum_samples / 8.0


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (5 ms period; the
    timed region of a default run is ~0.1 s, shorter than nvidia-smi's start-up), nvidia-smi -lms as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.thread = None
        self.samples = []
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((sm, reasons))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            import threading
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self.nvml
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": ["no samples"]}
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            reasons = sorted({n for n, bit in names for _, r in self.samples if r & bit})
            sm = [v for v, _ in self.samples]
            busy = [v for v in sm if v >= 0.5 * max(sm)] or sm
            return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.smax, "reasons": reasons, "samples": len(sm), "source": "nvml, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [v for v in sm if v >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def load_cpu_lib():
    """(world wrapper, kind): the compiled reference when present, else the C restatement."""
    from tests import common
    if common.have_ref():
        return common.load_ref(), "reference"
    return common.load_oracle("glibc"), "port"


def cpu_rate(world, edges, budget_s, threads):
    """edges/s of the CPU implementation on a bounded sample of `edges` sized for ~budget_s."""
    from tests import common
    w, kind = load_cpu_lib()
    if kind == "reference":
        world.upload_ref(w)
    else:
        world.upload(w)
    probe = edges[: min(len(edges), 256 if threads == 1 else 1024)]
    t0 = time.perf_counter()
    common.true_cost_mt(w, probe, threads)
    dt = time.perf_counter() - t0
    rate = len(probe) / max(dt, 1e-9)
    n = int(max(len(probe), min(len(edges), rate * budget_s)))
    sample = edges[:n]
    t0 = time.perf_counter()
    common.true_cost_mt(w, sample, threads)
    dt = time.perf_counter() - t0
    return n / dt, kind, n, dt


PLAN_SCENARIOS = (("c1", None), ("c2", None), ("c2", (470.0, 610.0, 3.0, 2.5, 1.0)), ("c3", (420.0, 395.0, 0.0, 2.5, 1.0)),
                  ("c3b", (420.0, 395.0, 0.0, 2.5, 1.0)), ("c4", None))


def plan_at_budget(budget_s, device, reps=3):
    """Reference AStarPlanner (CPU) vs the product harness (BatchedAStarPlanner on `device`) with a real wall-clock
    budget (tick = 0 -> real clock) on the BASELINE worlds (the start states of tests/plan_cases.py)."""
    from path_planner_b200 import harness as ph
    from path_planner_b200 import synth
    from tests import common
    if not ph.available():
        return {"unavailable": "path_planner_b200/libppe_harness.so not built (needs the reference sources at build time)"}
    ref = common.load_ref() if common.have_ref() else None
    h = ph.PlanningHarness(device)
    out = {"budget_s": budget_s, "repetitions": reps,
           "unit": "f = g + h of the returned plan, seconds (lower is better); every scenario is planned `repetitions` times by "
                   "each planner on the real clock, repetition r of both rebased to the same start instant so that the sampler "
                   "seed (the integer second of the deadline, AStarPlanner.cpp:33) is the same: f per repetition, median, best, "
                   "mean expansions and expansions per second of wall time",
           "scenarios": []}
    for wname, start in PLAN_SCENARIOS:
        world = synth.WORLDS[wname]()
        st0 = world.start if start is None else np.array(start, dtype=np.float64)
        rec = {"world": wname, "start": [float(v) for v in st0]}
        initial = 10000 if wname == "c4" else 100  # C4 runs with initialSamples = 10 000
        # the planner seeds its sampler with the integer second of its deadline (AStarPlanner.cpp:33): repetition r of both
        # planners runs on the real clock rebased to start at the same instant 1e9 + 10 r + 0.25, so they draw the same samples
        epochs = [1.0e9 + 10.0 * r + 0.25 for r in range(reps)]
        if ref is not None:
            sid = world.upload_ref(ref)
            fs, exp, smp, wall = [], [], [], []
            for c0 in epochs:
                t0 = time.perf_counter()
                plan, st = common.run_plan(ref, "ref", sid, st0, budget_s, c0, 0.0, initial)
                wall.append(time.perf_counter() - t0)
                fs.append(st["f"] if len(plan) else float("inf"))
                exp.append(st["expanded"])
                smp.append(st["samples"])
            rec["reference_cpu"] = {"f": fs, "f_median": statistics.median(fs), "f_best": min(fs),
                                    "plans_found": sum(1 for f in fs if f < float("inf")), "expanded_mean": sum(exp) / reps,
                                    "samples_mean": sum(smp) / reps, "expansions_per_s": sum(exp) / sum(wall),
                                    "wall_s_mean": round(sum(wall) / reps, 3)}
        h.set_world(world)
        fs, exp, smp, wall, hits, batches, exact = [], [], [], [], [], [], []
        where = {"engine_expand": 0.0, "replay": 0.0, "add_samples": 0.0, "exact": 0.0}
        for c0 in epochs:
            plan, st = h.plan(st0, budget_s, clock0=c0, initial_samples=initial)
            wall.append(st["wall_seconds"])
            fs.append(st["plan_f"] if len(plan) else float("inf"))
            exp.append(st["expanded"])
            smp.append(st["samples"])
            hits.append(st["frontier_hits"])
            batches.append(st["engine_batches"])
            exact.append((st["exact_expansions"], st["exact_for_ties"], st["exact_for_overflow"]))
            for k in where:
                where[k] += st["seconds_" + k]
        rec["engine"] = {"f": fs, "f_median": statistics.median(fs), "f_best": min(fs),
                         "plans_found": sum(1 for f in fs if f < float("inf")),
                         "expanded_mean": sum(exp) / reps, "samples_mean": sum(smp) / reps, "expansions_per_s": sum(exp) / sum(wall),
                         "frontier_hit_rate": sum(hits) / max(1, sum(exp)), "engine_batches_mean": sum(batches) / reps,
                         "exact_expansions_mean": sum(e[0] for e in exact) / reps, "exact_for_ties": sum(e[1] for e in exact),
                         "exact_for_overflow": sum(e[2] for e in exact), "wall_s_mean": round(sum(wall) / reps, 3),
                         "wall_share": {k: round(v / max(1e-9, sum(wall)), 3) for k, v in where.items()}}
        if ref is not None:
            rec["engine_f_not_worse"] = [bool(a <= b + 1e-9 * max(1.0, abs(b))) for a, b in zip(fs, rec["reference_cpu"]["f"])]
        if ref is not None and rec["reference_cpu"]["expansions_per_s"] > 0:
            rec["expansions_per_s_ratio"] = rec["engine"]["expansions_per_s"] / rec["reference_cpu"]["expansions_per_s"]
        out["scenarios"].append(rec)
    return out


SCENARIO_SAMPLE_TICK = 2e-7  # virtual seconds per generated sample: bounds the sample doubling as generation time does


def scenario_world(s):
    """Scenario s of BASELINE configs[4]: a C3-style world (C2 map and ribbons, 50 Gaussian obstacles from seed 100 + s)
    and a start state drawn from seed 1000 + s."""
    from path_planner_b200 import synth
    w = synth.world_c2()
    w.name = "scenario-%d" % s
    synth._add_obstacles(w, "gaussian", 50, 300.0, 700.0, 100 + s)
    rng = np.random.default_rng(1000 + s)
    start = np.array([rng.uniform(385, 595), rng.uniform(385, 615), rng.uniform(0, 2 * np.pi), 2.5, 1.0])
    return w, start


def scenario_sweep(n_scenarios, rank, world_size, device, tick):
    """Independent planning scenarios sharded over the ranks (scenario s -> rank s mod N, replicas only): each one is a
    whole Planner::plan call of the product harness on a virtual clock (0.95 s budget, `tick` s per now() call), so
    the work per scenario is deterministic.  Returns (wall seconds of this rank, expansions, [(s, f)])."""
    from path_planner_b200 import harness as ph
    from path_planner_b200 import sharding
    h = ph.PlanningHarness(device)
    out, expanded = [], 0
    t0 = time.perf_counter()
    for s in sharding.scenario_assignment(n_scenarios, rank, world_size):
        w, start = scenario_world(s)
        h.set_world(w)
        plan, st = h.plan(start, 0.95, clock0=1000.0, tick=tick, sample_tick=SCENARIO_SAMPLE_TICK)
        out.append((s, st["plan_f"] if len(plan) else float("inf")))
        expanded += st["expanded"]
    return time.perf_counter() - t0, expanded, out


def dubins_rate(eng, torch, dev, sh, n=1 << 22, steps=5):
    """K1: solves/s with inputs resident in HBM (same pose distribution as the edge sweep)."""
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    q0 = torch.rand((n, 3), dtype=torch.float64, device=dev, generator=g) * torch.tensor([200.0, 200.0, 2 * math.pi], dtype=torch.float64, device=dev)
    q1 = q0.clone()
    q1[:, :2] += (torch.rand((n, 2), dtype=torch.float64, device=dev, generator=g) - 0.5) * 150.0
    q1[:, 2] = torch.rand(n, dtype=torch.float64, device=dev, generator=g) * 2 * math.pi
    rho = torch.full((n,), 8.0, dtype=torch.float64, device=dev)
    rho[1::2] = 16.0
    typ = torch.empty(n, dtype=torch.int32, device=dev)
    err = torch.empty(n, dtype=torch.int32, device=dev)
    par = torch.empty((n, 3), dtype=torch.float64, device=dev)
    length = torch.empty(n, dtype=torch.float64, device=dev)
    args = (n, q0.data_ptr(), q1.data_ptr(), rho.data_ptr(), typ.data_ptr(), par.data_ptr(), length.data_ptr(), err.data_ptr(), sh)
    for _ in range(2):
        eng.dubins_batch_device(*args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        eng.dubins_batch_device(*args)
    b.record()
    torch.cuda.synchronize()
    return n * steps / (a.elapsed_time(b) * 1e-3)


def ncu_metrics(workload, n):
    """fp64 pipe utilisation (time-weighted over the kernels of one step) and DRAM bytes per step from the committed
    `ncu --set full` captures of this very command (profiles/r02_ncu_metrics.json, written by tools/ncu_summary.py), or
    {} when there is no capture for this workload / batch size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")) as f:
            t = json.load(f)
        for rec in t.get("captures", []):
            if rec.get("workload") == workload and rec.get("edges") == n:
                return rec
        for rec in t.get("captures", []):  # a shard of the captured batch (--gpus N): the pipe share carries over, bytes scale with n
            if rec.get("workload") == workload and rec.get("edges") and n < rec["edges"]:
                out = dict(rec)
                out["dram_bytes"] = rec["dram_bytes"] * n / rec["edges"] if rec.get("dram_bytes") is not None else None
                out["source"] = "%s; scaled from the %d-edge capture to this rank's %d edges" % (rec.get("source", ""), rec["edges"], n)
                return out
    except (OSError, ValueError):
        pass
    return {}


def run_reference(args, world, edges_fn):
    """--impl reference: the reference's CPU implementation, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import common
    w, kind = load_cpu_lib()
    if kind == "reference":
        world.upload_ref(w)
    else:
        world.upload(w)
    cores = os.cpu_count() or 1
    edges = edges_fn(0, 1)
    probe = edges[:1024]
    t0 = time.perf_counter()
    common.true_cost_mt(w, probe, 0)
    rate = len(probe) / (time.perf_counter() - t0)
    total_steps = args.steps + args.warmup
    n = int(max(1024, min(len(edges), rate * (100.0 / total_steps))))
    sample = edges[:n]
    for _ in range(args.warmup):
        common.true_cost_mt(w, sample, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        common.true_cost_mt(w, sample, 0)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    # north_star's stated baseline is the SINGLE-threaded planner: time that too, on a smaller sample
    rate1, _, n1, dt1 = cpu_rate(world, edges, 10.0, 1)
    line = {
        "impl": "reference", "metric": "dubins_edge_true_cost_evals_per_sec", "value": value, "unit": "edges/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "edges_total": args.edges if args.scaling == "strong" else args.edges * args.gpus,
                   "note": "CPU arm: each step evaluates a bounded sample of the same edge batch on all host threads"},
        "cpu_baseline": {"value": value, "unit": "edges/s", "cores": cores, "kind": kind,
                         "sample": "%d edges of the %d-edge batch per step, %d host threads" % (n, len(edges), cores)},
        "cpu_baseline_1thread": {"value": rate1, "unit": "edges/s", "cores": 1, "kind": kind,
                                 "sample": "first %d edges of the same batch, single thread, %.1f s" % (n1, dt1)},
        "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure(torch, eng, world, edges, steps, warmup, stream, after_step=None):
    """Device-timed steps of one resident batch through ppe_true_cost_batch_device.  Returns a dict of timings and
    the per-batch work counters read back from the result records."""
    from path_planner_b200 import abi
    n = len(edges)
    dev = torch.device("cuda", torch.cuda.current_device())
    h_edges = torch.from_numpy(edges.view(np.uint8).reshape(n, abi.EDGE_DTYPE.itemsize)).pin_memory()
    d_edges = h_edges.to(dev, non_blocking=True)
    d_results = torch.empty((n, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    sh = stream.cuda_stream
    torch.cuda.synchronize()

    def step():
        eng.true_cost_batch_device(n, d_edges.data_ptr(), d_results.data_ptr(), sh)
        if after_step is not None:
            after_step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    return step, h_edges, d_results


def work_counters(torch, d_results, n):
    from path_planner_b200 import abi
    r32 = d_results.view(torch.int32).reshape(n, abi.RESULT_DTYPE.itemsize // 4)
    return {
        "culled_samples": int((r32[:, 51] & 0xFFFFFF).to(torch.int64).sum().item()),   # `reserved` bits 0-23
        "thread_walked": int(((r32[:, 51] >> 24) & 1).to(torch.int64).sum().item()),   # bit 24: walked by a K2t thread
        "sum_samples": int(r32[:, 47].to(torch.int64).sum().item()),
        "sum_cp": int(r32[:, 48].to(torch.int64).sum().item()),
        "infeasible": int(r32[:, 45].to(torch.int64).sum().item()),
        "bad_status": int((r32[:, 46] != 0).sum().item()),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--edges", type=int, default=1 << 20, help="edges of the batch (strong: in total; weak: per GPU)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--near-ribbons", type=float, default=0.0)
    ap.add_argument("--no-plan", action="store_true", help="skip the plan-cost-at-1-s-budget comparison")
    ap.add_argument("--plan-budget", type=float, default=1.0)
    ap.add_argument("--no-extra", action="store_true", help="skip the other workloads, the scenario sweep and the Dubins line")
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--scenario-tick", type=float, default=4e-3, help="virtual seconds per now() call in the scenario sweep")
    args = ap.parse_args()

    from path_planner_b200 import abi, sharding, synth

    world = synth.WORLDS[args.workload]()

    def edges_fn(rank, world_size):
        """This rank's edges: a shard of ONE seeded batch (strong) or the rank's own batch (weak)."""
        if args.scaling == "strong" or world_size == 1:
            e = synth.make_edges(world, args.edges, seed=5, near_ribbons=args.near_ribbons)
            lo, hi = sharding.shard_range(args.edges, rank, world_size)
            return e[lo:hi]
        return synth.make_edges(world, args.edges, seed=5 + 1000 * rank, near_ribbons=args.near_ribbons)

    if args.impl == "reference":
        run_reference(args, world, edges_fn)
        return

    import torch
    import torch.distributed as dist
    from path_planner_b200 import EdgeEngine

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world_size > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on STDOUT when the communicator comes up; stdout carries exactly one JSON line,
        # so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    eng = EdgeEngine(local_rank)
    set_id = world.upload(eng)
    edges = edges_fn(rank, world_size)
    edges["ribbon_set"] = set_id
    n = len(edges)
    lo = sharding.shard_range(args.edges, rank, world_size)[0] if args.scaling == "strong" else 0
    n_total = args.edges if (args.scaling == "strong" or not distributed) else args.edges * world_size

    best_local = torch.zeros(2, dtype=torch.float64, device=dev)  # {f64 f, i64 idx} as 16 raw bytes
    best_all = torch.zeros(2 * world_size, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream

    def gather_best():
        # the path's only exchange: one (best f, GLOBAL edge index) record per GPU back to the planning rank
        eng.best_copy_device(best_local.data_ptr(), lo, sh)
        dist.all_gather_into_tensor(best_all, best_local)

    step, h_edges, d_results = measure(torch, eng, world, edges, args.steps, args.warmup, stream, gather_best if distributed else None)
    fp64_peak = eng.measure_fp64_peak(sh) if rank == 0 else 0.0
    torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for k in range(args.steps):
        step()
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    launches = eng.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    wc = work_counters(torch, d_results, n)

    gather_check = None
    if distributed:
        # the gathered record must be the minimum of the per-rank records, with GLOBAL edge indices
        f_loc, i_loc = eng.best_device(sh)
        mine = torch.tensor([f_loc if i_loc >= 0 else float("inf"), float(lo + i_loc if i_loc >= 0 else -1)], dtype=torch.float64, device=dev)
        allr = torch.zeros(2 * world_size, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        rec = best_all.cpu().numpy().reshape(world_size, 2)
        fs, idx = rec[:, 0].copy(), rec[:, 1].copy().view(np.int64)
        exp = allr.cpu().numpy().reshape(world_size, 2)
        gather_check = bool(np.array_equal(fs, exp[:, 0]) and np.array_equal(idx, exp[:, 1].astype(np.int64)))
        if not gather_check:
            raise SystemExit("bench.py: the NCCL-gathered best records differ from the per-rank ppe_best_device records")

    # ---- end to end through the public host-buffer API -----------------------------------------
    res_host = torch.empty((n, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    lib = eng._lib
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        rc = lib.ppe_true_cost_batch(eng._ctx, n, C.c_void_p(h_edges.data_ptr()), C.c_void_p(res_host.data_ptr()))
        if rc != 0:
            raise RuntimeError("ppe_true_cost_batch failed: %d" % rc)

    e2e_step()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s_max = float(te.item())

    # ---- secondary line: weak scaling (every GPU its own full batch), 3 steps -------------------
    weak = None
    if distributed and args.scaling == "strong" and not args.no_extra:
        e_w = synth.make_edges(world, args.edges, seed=5 + 1000 * rank, near_ribbons=args.near_ribbons)
        e_w["ribbon_set"] = set_id

        def gather_best_weak():
            eng.best_copy_device(best_local.data_ptr(), rank * args.edges, sh)
            dist.all_gather_into_tensor(best_all, best_local)

        step_w, h_w, r_w = measure(torch, eng, world, e_w, 3, 2, stream, gather_best_weak)
        dist.barrier()
        torch.cuda.synchronize()
        a_w, b_w = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_w.record(stream)
        for _ in range(3):
            step_w()
        b_w.record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t_w = torch.tensor([a_w.elapsed_time(b_w)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_w, op=dist.ReduceOp.MAX)
        ms_w = float(t_w.item()) / 3
        weak = {"scaling": "weak", "value": args.edges * world_size / (ms_w * 1e-3), "unit": "edges/s", "ms_per_step": ms_w, "steps": 3,
                "edges_per_gpu": args.edges, "note": "one %d-edge batch per GPU (seed 5 + 1000 rank), same kernels and 16-byte gather" % args.edges}
        del step_w, h_w, r_w

    # ---- BASELINE configs[4], second half: independent scenarios sharded over the ranks ----------
    scen = None
    if not args.no_extra and args.scenarios > 0:
        from path_planner_b200 import harness as ph
        if ph.available():
            if distributed:
                dist.barrier()
            wall, expanded, mine = scenario_sweep(args.scenarios, rank, world_size, local_rank, args.scenario_tick)
            costs = torch.full((args.scenarios,), float("-inf"), dtype=torch.float64, device=dev)
            for s_, f_ in mine:
                costs[s_] = f_
            st = torch.tensor([wall, float(expanded)], dtype=torch.float64, device=dev)
            if distributed:  # the gather of the 64 final plan costs: each rank filled its own slots
                dist.all_reduce(costs, op=dist.ReduceOp.MAX)
                tmax = st[:1].clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(st[1:], op=dist.ReduceOp.SUM)
                st[0] = tmax[0]
            c = costs.cpu().numpy()
            fin = np.isfinite(c)
            scen = {"metric": "planning_scenarios_per_sec", "value": args.scenarios / float(st[0]), "unit": "scenarios/s",
                    "scenarios": args.scenarios, "wall_s": float(st[0]), "expanded_total": int(st[1]), "plans_found": int(fin.sum()),
                    "f_checksum": float(c[fin].sum()), "scaling": "strong",
                    "config": "C3-style worlds (seeds 100..), whole Planner::plan per scenario, virtual clock 0.95 s / %g s per now() "
                              "+ %g s per generated sample; scenario s -> rank s mod N, replicas only" % (args.scenario_tick, SCENARIO_SAMPLE_TICK)}
        else:
            scen = {"unavailable": "path_planner_b200/libppe_harness.so not built"}

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = n_total * args.steps / (total_ms_max * 1e-3)
        n_rib = len(world.ribbons)
        n_obs = 0 if world.obstacle_kind == "none" else len(world.obstacles["x"])
        flops = algorithmic_flops(n, wc["sum_samples"], wc["sum_cp"], n_rib, n_obs, world.obstacle_kind)
        kernel_s = total_ms * 1e-3 / args.steps  # rank 0's own average launch-group duration
        achieved_tf = flops / kernel_s / 1e12
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bytes_alg = algorithmic_bytes(n, wc["sum_samples"])
        nm = ncu_metrics(args.workload, n)
        line = {
            "metric": "dubins_edge_true_cost_evals_per_sec", "value": value, "unit": "edges/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "edges_total": n_total, "edges_rank0": n, "edge_bytes": abi.EDGE_DTYPE.itemsize,
                       "result_bytes": abi.RESULT_DTYPE.itemsize, "l2": "inputs+outputs (%.0f MB on rank 0) larger than L2" %
                       (n * (abi.EDGE_DTYPE.itemsize + abi.RESULT_DTYPE.itemsize) / 1e6),
                       "mean_samples_per_edge": wc["sum_samples"] / n, "mean_checkpoints_per_edge": wc["sum_cp"] / n,
                       "infeasible_edges": wc["infeasible"], "edges_with_status": wc["bad_status"],
                       "culled_sample_fraction": wc["culled_samples"] / max(1, wc["sum_samples"]),
                       "thread_walked_edge_fraction": wc["thread_walked"] / n,
                       "parallelism": ("one %d-edge batch split into %d contiguous shards (sharding.shard_range), one process per GPU, "
                                       "16-byte best record all-gathered over NCCL" % (n_total, world_size)) if args.scaling == "strong" or not distributed
                       else "one %d-edge batch per GPU, %d GPUs" % (args.edges, world_size),
                       "best_gather_verified": gather_check},
            "e2e": {"value": n_total * e2e_steps / e2e_s_max, "unit": "edges/s",
                    "h2d_bytes_per_step": n * abi.EDGE_DTYPE.itemsize, "d2h_bytes_per_step": n * abi.RESULT_DTYPE.itemsize + 8,
                    "steps": e2e_steps, "note": "bytes are rank 0's share per step"},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {
                "bound": "fp64", "kernel": "k2a_prepare + k2t_thread_walk + k2_true_cost (the launch group of one step)",
                "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None,
                "peak_source": "measured live: ppe_measure_fp64_peak (DFMA chain, 2 flop/FMA) -- MEASURED_PEAKS.json has no fp64 entry",
                "flops_per_launch": flops,
                "fp64_pipe_pct": nm.get("fp64_pipe_pct"), "traffic": nm.get("dram_bytes"),
                "ncu_source": nm.get("source"),
                "note": "achieved = ALGORITHMIC flops (SURVEY 8d: every executed sample point of the reference loop) / launch time; "
                        "the kernels prove config.culled_sample_fraction of the sample points clean in 32-sample chunks and do "
                        "not evaluate them, so frac measures work done per second in the reference's units and can exceed 1; "
                        "fp64_pipe_pct is what the hardware's FP64 pipe was busy with (ncu, time-weighted over the step's kernels)",
                "hbm": {"achieved": bytes_alg / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_alg / kernel_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                        "bytes_moved_per_launch": n * (abi.EDGE_DTYPE.itemsize + abi.RESULT_DTYPE.itemsize)},
            },
        }
        if weak is not None:
            line["weak"] = weak
        if scen is not None:
            line["scenarios"] = scen
        if world_size == 1 and not args.no_extra:
            # the other BASELINE configs as secondary lines: same metric, fewer steps
            others = {}
            for wname in ("c2", "c3", "c4"):
                if wname == args.workload:
                    continue
                w2 = synth.WORLDS[wname]()
                e2 = synth.make_edges(w2, args.edges, seed=5, near_ribbons=args.near_ribbons)
                e2["ribbon_set"] = w2.upload(eng)
                step2, h2, r2 = measure(torch, eng, w2, e2, 3, 2, stream)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(3):
                    step2()
                b.record(stream)
                torch.cuda.synchronize()
                ms2 = a.elapsed_time(b) / 3
                wc2 = work_counters(torch, r2, len(e2))
                nobs2 = 0 if w2.obstacle_kind == "none" else len(w2.obstacles["x"])
                fl2 = algorithmic_flops(len(e2), wc2["sum_samples"], wc2["sum_cp"], len(w2.ribbons), nobs2, w2.obstacle_kind)
                nm2 = ncu_metrics(wname, len(e2))
                others[wname] = {"workload": WORKLOADS[wname], "value": len(e2) / (ms2 * 1e-3), "unit": "edges/s", "ms_per_step": ms2,
                                 "steps": 3, "roofline_frac": fl2 / (ms2 * 1e-3) / 1e12 / fp64_peak if fp64_peak > 0 else None,
                                 "culled_sample_fraction": wc2["culled_samples"] / max(1, wc2["sum_samples"]),
                                 "thread_walked_edge_fraction": wc2["thread_walked"] / len(e2),
                                 "fp64_pipe_pct": nm2.get("fp64_pipe_pct"), "traffic": nm2.get("dram_bytes")}
                del step2, h2, r2
            line["workloads"] = others
            world.upload(eng)
            line["dubins"] = {"metric": "dubins_solves_per_sec", "value": dubins_rate(eng, torch, dev, sh), "unit": "solves/s",
                              "n": 1 << 22, "note": "K1, one GPU, HBM-resident, correctly rounded transcendentals"}
        if world_size == 1 and not args.no_plan:
            try:
                line["plan"] = plan_at_budget(args.plan_budget, local_rank)
            except Exception as ex:  # the comparison is auxiliary: never lose the bench line over it
                line["plan"] = {"unavailable": "%s: %s" % (type(ex).__name__, ex)}
        if world_size == 1 and not args.no_cpu_baseline:
            rate, kind, ns, dt = cpu_rate(world, edges, args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "edges/s", "cores": 1, "kind": kind,
                                    "sample": "first %d edges of the same batch, single thread, %.1f s" % (ns, dt),
                                    "host_cores": os.cpu_count()}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
