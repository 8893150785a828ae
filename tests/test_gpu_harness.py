"""The product's standalone harness (path_planner_b200/libppe_harness.so + ppe_plan_main, include/ppe_harness.h) on the GPU:
its plans against the reference's own AStarPlanner (oracle/_ref/libref_planner.so, a separate library in the same
process) on a virtual clock, the Plan.msg export, the visualization stream and the command-line driver."""
import json
import os
import subprocess

import numpy as np
import pytest

from path_planner_b200 import harness as ph
from path_planner_b200 import synth
from tests import common, plan_cases

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ph.available() or not common.have_ref(),
                                                  reason="libppe_harness.so / libref_planner.so not built (need /root/reference)")]


def as12(plan):
    """pph_dubins_path records -> the 12-double rows tests/common.run_plan returns."""
    out = np.zeros((len(plan), 12))
    for i, k in enumerate(("initial_x", "initial_y", "initial_yaw", "length0", "length1", "length2", "rho", "type", "speed",
                           "start_time", "end_time")):
        out[:, i] = plan[k]
    return out


@pytest.fixture(scope="module")
def ref():
    return common.load_ref()


@pytest.mark.parametrize("i", [0, 2, 5, 6, 8])
def test_harness_plans_equal_the_reference(ref, i):
    _, wname, start, budget, tick, initial = plan_cases.CASES[i]
    world = plan_cases.make_world(wname)
    start = world.start if start is None else np.array(start, dtype=np.float64)
    sid = world.upload_ref(ref)
    want_plan, want = common.run_plan(ref, "ref", sid, start, budget, 1000.0, tick, initial)
    h = ph.PlanningHarness(0)
    h.set_world(world)
    plan, st = h.plan(start, budget, clock0=1000.0, tick=tick, initial_samples=initial)
    got = as12(plan)
    assert got.shape == want_plan.shape
    assert np.array_equal(got[:, 7], want_plan[:, 7]) and np.array_equal(got[:, 6], want_plan[:, 6])
    assert np.allclose(got, want_plan, rtol=common.RTOL, atol=common.ATOL)
    for a, b in (("samples", "samples"), ("generated", "generated"), ("expanded", "expanded"), ("iterations", "iterations"),
                 ("plan_depth", "depth"), ("now_calls", "now_calls")):
        assert st[a] == want[b], (a, st, want)
    assert np.isclose(st["plan_f"], want["f"], rtol=common.RTOL, atol=common.ATOL)
    assert st["true_cost_edges"] > 0 and st["frontier_vertices"] >= st["expanded"]


@pytest.mark.parametrize("i", [1, 6])
def test_plan_identity_over_sampler_seeds(ref, i):
    """The sampler's seed is the integer second of the deadline (AStarPlanner.cpp:33): other clock origins give other sample
    sets; the plans must stay the reference's for each of them.  The virtual clock charges time per generated sample so that
    the anytime loop's sample doubling stays bounded whatever the seed does to the length of the iterations."""
    _, wname, start, budget, tick, initial = plan_cases.CASES[i]
    world = plan_cases.make_world(wname)
    start = world.start if start is None else np.array(start, dtype=np.float64)
    sid = world.upload_ref(ref)
    h = ph.PlanningHarness(0)
    h.set_world(world)
    for clock0 in (1.0e9 + 1.25, 1.0e9 + 11.25, 77.5):
        want_plan, want = common.run_plan(ref, "ref", sid, start, budget, clock0, tick, initial, sample_tick=2e-6)
        plan, st = h.plan(start, budget, clock0=clock0, tick=tick, initial_samples=initial, sample_tick=2e-6)
        got = as12(plan)
        assert got.shape == want_plan.shape and st["expanded"] == want["expanded"] and st["samples"] == want["samples"], (clock0, st, want)
        assert np.array_equal(got[:, 7], want_plan[:, 7])
        assert np.allclose(got, want_plan, rtol=common.RTOL, atol=plan_cases.RETRY_FLIP_SLACK), clock0


def test_two_cycles_with_previous_plan(ref):
    """Receding-horizon use: cycle 2 starts one second along cycle 1's plan and gets it as previousPlan."""
    world = synth.world_c2()
    start = np.array([395.0, 390.0, 0.3, 2.5, 1.0])
    sid = world.upload_ref(ref)
    h = ph.PlanningHarness(0)
    h.set_world(world)
    plan1, st1 = h.plan(start, 0.95, clock0=1000.0, tick=2e-3)
    ref1, _ = common.run_plan(ref, "ref", sid, start, 0.95, 1000.0, 2e-3, 100)
    assert np.allclose(as12(plan1), ref1, rtol=common.RTOL, atol=common.ATOL)
    nxt = plan_cases.state_along(ref1, start[4] + 1.0)
    # the harness's own bookkeeping gives the same next start state (DubinsPlan::sample of the reference's classes)
    assert np.allclose(h.advance(start[4] + 1.0), nxt, rtol=1e-9, atol=1e-9)
    plan2, st2 = h.plan(nxt, 0.95, clock0=1007.0, tick=2e-3, previous=plan1)
    ref2, want2 = common.run_plan(ref, "ref", sid, nxt, 0.95, 1007.0, 2e-3, 100, previous=ref1)
    # NB the harness covered the ribbons between the two starts (executive.cpp:188); the reference context did not,
    # so only compare when nothing was covered on that first second
    if st2["expanded"] == want2["expanded"]:
        assert np.allclose(as12(plan2), ref2, rtol=common.RTOL, atol=common.ATOL)
    assert len(plan2) >= 1


def test_plan_msg_export_and_visualization_stream(tmp_path):
    world = synth.world_c1()
    h = ph.PlanningHarness(0)
    h.set_world(world)
    vis = str(tmp_path / "vis.txt")
    plan, st = h.plan(world.start, 0.95, clock0=1000.0, tick=5e-4, visualization_path=vis)
    assert len(plan) >= 1
    msg = str(tmp_path / "plan.yaml")
    h.write_plan_msg(plan, msg)
    text = open(msg).read().split("\n")
    # Plan.msg / DubinsPath.msg field order (path_planner_common/msg, NodeBase.h:201-220)
    assert text[0] == "paths:"
    keys = [ln.strip().lstrip("- ").split(":")[0] for ln in text[1:11]]
    assert keys == ["initial_x", "initial_y", "initial_yaw", "length0", "length1", "length2", "rho", "type", "speed", "start_time"]
    assert text[-2].startswith("endtime: ") and float(text[-2].split()[1]) == plan["end_time"][-1]
    # the stream visualizer.py reads: vertices, samples, trajectories, ribbons, the plan
    dump = open(vis).read()
    for needle in ("Trajectory:", " trajectory", " sample", "Expanded State: (", "Generated State: (", " vertex ", " start ",
                   "Ribbons: ", "End Ribbons", "Incumbent f-value", " plan"):
        assert needle in dump, needle
    line = next(ln for ln in dump.split("\n") if ln.endswith(" trajectory"))
    assert line.startswith("State: (") and ", f: " in line and ", g: " in line and ", h: " in line


def test_ppe_plan_main_runs_a_scenario_file(tmp_path):
    scenario = tmp_path / "scenario.txt"
    scenario.write_text("\n".join([
        "# single ribbon, empty map",
        "config ribbon_width 1.5", "config heuristic 0",
        "map none", "ribbon 0 10 0 30", "start 0 0 0 2.5 1", "budget 0.95", "virtual_clock 1000 0.0005", "cycles 2", "period 1.0"]))
    out = subprocess.run([ph.PLAN_MAIN_PATH, str(scenario), "--plan-out", str(tmp_path / "p.yaml")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(ln) for ln in out.stdout.strip().split("\n")]
    assert len(lines) == 2 and lines[0]["paths"] >= 1 and lines[0]["expanded"] == 191  # the c1 case of tests/plan_cases.py
    assert lines[1]["paths"] >= 1 and lines[1]["true_cost_edges"] > 0
    assert open(tmp_path / "p.yaml").read().startswith("paths:")
