"""The product harness's C ABI (include/ppe_harness.h: pph_*) without a GPU: oracle/_ref/libppe_harness_cpu.so is
path_planner_b200/harness/harness_capi.cpp + BatchedAStarPlanner.cpp linked against the CPU test double of the engine
(oracle/ppe_on_oracle.c).  World set-up, plan cycles with a previous plan, the Plan.msg export and the visualization stream
are host logic; with the bit-identical evaluator behind it the plans must be the compiled reference's bit for bit."""
import os

import numpy as np
import pytest

from path_planner_b200 import harness as ph
from path_planner_b200 import synth
from tests import common, plan_cases

CPU_SO = os.path.join(common.ORACLE_DIR, "_ref", "libppe_harness_cpu.so")
pytestmark = pytest.mark.skipif(not os.path.exists(CPU_SO) or not common.have_ref(),
                                reason="oracle/_ref/libppe_harness_cpu.so / libref_planner.so not built (need /root/reference)")


def as12(plan):
    out = np.zeros((len(plan), 12))
    for i, k in enumerate(("initial_x", "initial_y", "initial_yaw", "length0", "length1", "length2", "rho", "type", "speed",
                           "start_time", "end_time")):
        out[:, i] = plan[k]
    return out


@pytest.mark.parametrize("i", [0, 6, 8])
def test_pph_plan_equals_the_reference_bit_for_bit(i):
    _, wname, start, budget, tick, initial = plan_cases.CASES[i]
    world = plan_cases.make_world(wname)
    start = world.start if start is None else np.array(start, dtype=np.float64)
    ref = common.load_ref()
    sid = world.upload_ref(ref)
    want_plan, want = common.run_plan(ref, "ref", sid, start, budget, 1000.0, tick, initial)
    h = ph.PlanningHarness(0, lib_path=CPU_SO)
    h.set_world(world)
    plan, st = h.plan(start, budget, clock0=1000.0, tick=tick, initial_samples=initial)
    assert np.array_equal(as12(plan), want_plan)
    for a, b in (("samples", "samples"), ("generated", "generated"), ("expanded", "expanded"), ("iterations", "iterations"),
                 ("plan_depth", "depth"), ("now_calls", "now_calls")):
        assert st[a] == want[b], (a, st, want)
    assert st["plan_f"] == want["f"] and st["plan_h"] == want["h"]
    # second cycle from one second along the plan with the first plan as previousPlan (executive.cpp:146,189)
    nxt = plan_cases.state_along(want_plan, start[4] + 1.0)
    plan2, st2 = h.plan(nxt, budget, clock0=1007.0, tick=tick, initial_samples=initial, previous=plan)
    ref2, want2 = common.run_plan(ref, "ref", sid, nxt, budget, 1007.0, tick, initial, previous=want_plan)
    assert np.array_equal(as12(plan2), ref2) and st2["expanded"] == want2["expanded"]


def test_sample_cost_clock_bounds_the_sample_doubling():
    """A virtual clock that only ticks per now() call lets the anytime loop double its sample set without bound where
    iterations are short; charging virtual time per generated sample bounds it as generation time does on a real clock.
    The reference planner under the same clock gives the same plan."""
    world = synth.world_c1()
    ref = common.load_ref()
    sid = world.upload_ref(ref)
    h = ph.PlanningHarness(0, lib_path=CPU_SO)
    h.set_world(world)
    plan, st = h.plan(world.start, 0.95, clock0=1000.0, tick=5e-3, sample_tick=2e-6)
    want_plan, want = common.run_plan(ref, "ref", sid, world.start, 0.95, 1000.0, 5e-3, 100, sample_tick=2e-6)
    assert st["samples"] == want["samples"] and st["samples"] < 1e6
    assert np.array_equal(as12(plan), want_plan)


def test_plan_msg_export_and_visualization_stream(tmp_path):
    world = synth.world_c1()
    h = ph.PlanningHarness(0, lib_path=CPU_SO)
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
                   "Ribbons: ", "End Ribbons", "Incumbent f-value", " plan", " goal "):
        assert needle in dump, needle
    line = next(ln for ln in dump.split("\n") if ln.endswith(" trajectory"))
    assert line.startswith("State: (") and ", f: " in line and ", g: " in line and ", h: " in line
    # the same run with the visualization off gives the same plan (the dump is a side channel)
    plan2, st2 = h.plan(world.start, 0.95, clock0=1000.0, tick=5e-4)
    assert np.array_equal(as12(plan), as12(plan2)) and st2["expanded"] == st["expanded"]
