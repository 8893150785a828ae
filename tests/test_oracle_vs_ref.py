"""Pins the oracle: the plain-C restatement (oracle/ppe_oracle.c, glibc variant) must be
BIT-IDENTICAL to the reference's own sources compiled in place (oracle/_ref/libref_planner.so)
on seeded batches of every world, and to the committed golden vectors made from that reference."""
import glob
import math
import os

import numpy as np
import pytest

from path_planner_b200 import abi, synth
from tests import common

needs_ref = pytest.mark.skipif(not common.have_ref(), reason="oracle/_ref/libref_planner.so not built (needs /root/reference)")


def _ribbon_lists_equal(a, b, idx):
    for i in idx:
        ra, rb = a.ribbons_after(i), b.ribbons_after(i)
        if ra.shape != rb.shape or not np.array_equal(ra, rb):
            return i
    return None


@needs_ref
@pytest.mark.parametrize("name,near,n", [("c1", 0.5, 3000), ("c2", 0.0, 3000), ("c2", 0.6, 3000), ("c3", 0.3, 400),
                                          ("c3b", 0.3, 600), ("c4", 0.5, 1200), ("c5", 0.3, 300)])
def test_oracle_is_bit_identical_to_the_compiled_reference(name, near, n):
    ref = common.load_ref()
    ora = common.load_oracle("glibc")
    world = synth.WORLDS[name]()
    sid = world.upload_ref(ref)
    order = common.ref_obstacle_order(ref) if world.obstacle_kind != "none" else None
    assert world.upload(ora, order) == sid
    edges = synth.make_edges(world, n, seed=17, near_ribbons=near)
    edges["ribbon_set"] = sid
    want = ref.true_cost_batch(edges)
    got = ora.true_cost_batch(edges)
    bad = common.diff_results(got, want, exact=True, check_counts=False)
    assert not bad, common.describe(bad, got, want)
    assert _ribbon_lists_equal(ora, ref, np.flatnonzero(want["ribbons_changed"])[:400]) is None
    assert want["ribbons_changed"].sum() > 0 or near == 0.0
    assert (want["infeasible"] == 1).any() or name == "c1"


NEW_WORLDS = {
    "c2-epoch": (lambda: synth.with_time_offset(synth.world_c2(), 1.7e9), 0.5, 1200),
    "c3-epoch": (lambda: synth.with_time_offset(synth.WORLDS["c3"](), 1.7e9), 0.3, 250),
    "c3b-epoch": (lambda: synth.with_time_offset(synth.WORLDS["c3b"](), 1.7e9), 0.3, 300),
    "c2@0.5m": (lambda: synth.with_resolution(synth.world_c2(), 0.5), 0.3, 1200),
    "c2@2.5m": (lambda: synth.with_resolution(synth.world_c2(), 2.5), 0.3, 1200),
    "c2@0.3m": (lambda: synth.with_resolution(synth.world_c2(), 0.3), 0.3, 1200),
    "c4@2.5m": (lambda: synth.with_resolution(synth.world_c4(), 2.5), 0.3, 600),
    "c3+cov": (lambda: synth.with_covariances(synth.WORLDS["c3"]()), 0.2, 250),
}


@needs_ref
@pytest.mark.parametrize("key", sorted(NEW_WORLDS))
def test_oracle_is_bit_identical_on_epoch_times_resolutions_and_covariances(key):
    """The configurations tests/test_gpu_parity.py adds in round 2 (epoch-scale state times as the ROS node feeds them,
    map resolutions != 1 m incl. non powers of two, per-obstacle covariances): the oracle is pinned against the
    compiled reference on them before the GPU is compared with the oracle."""
    mk, near, n = NEW_WORLDS[key]
    ref = common.load_ref()
    ora = common.load_oracle("glibc")
    world = mk()
    sid = world.upload_ref(ref)
    order = common.ref_obstacle_order(ref) if world.obstacle_kind != "none" else None
    assert world.upload(ora, order) == sid
    edges = synth.make_edges(world, n, seed=12, near_ribbons=near)
    edges["ribbon_set"] = sid
    want = ref.true_cost_batch(edges)
    got = ora.true_cost_batch(edges)
    bad = common.diff_results(got, want, exact=True, check_counts=False)
    assert not bad, common.describe(bad, got, want)
    assert _ribbon_lists_equal(ora, ref, np.flatnonzero(want["ribbons_changed"])[:200]) is None
    assert want["ribbons_changed"].sum() > 0 and (want["infeasible"] == 1).any()


def tsp_world(heuristic, n_ribbons):
    """C2 with its first n ribbons and one of the point-robot TSP heuristics (Executive installs the K variant with K = 2,
    executive.cpp:391; more than five ribbons would force MaxDistance, RibbonManager.cpp:381-385)."""
    w = synth.world_c2()
    w.cfg.heuristic = heuristic
    w.ribbons = w.ribbons[:n_ribbons].copy()
    w.name = "C2-tsp%d-%dr" % (heuristic, n_ribbons)
    return w


@needs_ref
@pytest.mark.parametrize("heuristic", [abi.H_TSP_POINT_ROBOT_NO_SPLIT_K, abi.H_TSP_POINT_ROBOT_NO_SPLIT_ALL])
@pytest.mark.parametrize("n_ribbons", [1, 3, 5])
def test_oracle_tsp_heuristics_are_bit_identical_to_the_reference(heuristic, n_ribbons):
    """h of RibbonManager::tspPointRobotNoSplitKRibbons / ...AllRibbons (RibbonManager.cpp:53-95) on the ribbons-after of
    every edge, including lists that splits have grown to 7 ribbons."""
    ref = common.load_ref()
    ora = common.load_oracle("glibc")
    world = tsp_world(heuristic, n_ribbons)
    sid = world.upload_ref(ref)
    assert world.upload(ora) == sid
    edges = synth.make_edges(world, 1200, seed=12, near_ribbons=0.7)
    edges["ribbon_set"] = sid
    want = ref.true_cost_batch(edges)
    got = ora.true_cost_batch(edges)
    bad = common.diff_results(got, want, exact=True, check_counts=False)
    assert not bad, common.describe(bad, got, want)
    assert (got["h"] >= 0).all() and want["n_ribbons_after"].max() > n_ribbons


@needs_ref
def test_oracle_has_path_edges_and_dubins_match_the_reference():
    """Winner edges of expand() (pre-solved wrapper, speed change) and previous-plan style wrappers
    (earlier start time, truncated end time), AStarPlanner.cpp:46-59."""
    ref = common.load_ref()
    ora = common.load_oracle("glibc")
    world = synth.world_c2()
    sid = world.upload_ref(ref)
    world.upload(ora)
    n = 2000
    edges = synth.make_edges(world, n, seed=23, near_ribbons=0.5)
    edges["ribbon_set"] = sid
    yaw = lambda h: np.where(math.pi / 2 - h < 0, math.pi / 2 - h + 2 * math.pi, math.pi / 2 - h)
    q0 = np.column_stack([edges["src"][:, 0], edges["src"][:, 1], yaw(edges["src"][:, 2])])
    q1 = np.column_stack([edges["dst"][:, 0], edges["dst"][:, 1], yaw(edges["dst"][:, 2])])
    rho = np.where(edges["coverage_allowed"] == 1, world.cfg.coverage_turning_radius, world.cfg.turning_radius)
    rt = ref.dubins_batch(q0, q1, rho)
    ot = ora.dubins_batch(q0, q1, rho)
    for a, b in zip(rt, ot):
        assert np.array_equal(a, b)
    typ, par, length, err = rt
    edges["has_path"] = 1
    edges["path_qi"] = q0
    edges["path_param"] = par
    edges["path_rho"] = rho
    edges["path_type"] = typ
    edges["w_speed"] = edges["dst"][:, 3]
    edges["w_start_time"] = edges["src"][:, 4]
    edges["w_end_time"] = edges["src"][:, 4] + length / edges["dst"][:, 3]
    # a third of them: wrapper started earlier than the source vertex and was truncated before
    k = np.arange(n) % 3 == 0
    edges["w_start_time"][k] -= 0.5
    edges["w_end_time"][k] = edges["w_start_time"][k] + 0.7 * length[k] / edges["dst"][k, 3]
    # a few with a radius that matches neither config radius -> re-solve (Edge.cpp:78-80)
    edges["path_rho"][::50] = 11.0
    want = ref.true_cost_batch(edges)
    got = ora.true_cost_batch(edges)
    bad = common.diff_results(got, want, exact=True, check_counts=False)
    assert not bad, common.describe(bad, got, want)


@needs_ref
def test_error_edges_match_the_reference():
    """Edges on which the reference throws out of computeTrueCost must carry a non-zero status."""
    ref = common.load_ref()
    ora = common.load_oracle("glibc")
    world = synth.world_c1()
    sid = world.upload_ref(ref)
    world.upload(ora)
    e = np.zeros(3, dtype=abi.EDGE_DTYPE)
    e["ribbon_set"] = sid
    e[0]["src"] = [0, 0, 0, 2.5, 1]; e[0]["dst"] = [0, 0, 0, 2.5]          # co-located: unset wrapper
    e[1]["src"] = [0, 0, 0, 2.5, 40]; e[1]["dst"] = [0, 50, 0, 2.5]        # source beyond the horizon
    e[2]["src"] = [0, 0, 0, 2.5, 1]; e[2]["dst"] = [0, 20, 0, 2.5]         # fine
    want = ref.true_cost_batch(e)
    got = ora.true_cost_batch(e)
    assert (want["status"] != 0).tolist() == [True, True, False]
    assert (got["status"] != 0).tolist() == [True, True, False]
    assert got["infeasible"][1] == want["infeasible"][1] == 1


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "edges_*.npz"))))
def test_oracle_reproduces_the_committed_golden_vectors(path):
    """tests/golden/*.npz were produced by the compiled reference (tests/golden/make_golden.py);
    this test needs no /root/reference."""
    z = np.load(path)
    name = os.path.basename(path)[len("edges_"):-len(".npz")]
    world = synth.WORLDS[name]()
    ora = common.load_oracle("glibc")
    order = z["obstacle_order"] if len(z["obstacle_order"]) else None
    sid = world.upload(ora, order)
    edges = z["edges"].copy()
    edges["ribbon_set"] = sid
    want = z["results"]
    got = ora.true_cost_batch(edges)
    bad = common.diff_results(got, want, exact=True, check_counts=False)
    assert not bad, common.describe(bad, got, want)
    offs = np.concatenate([[0], np.cumsum(z["ribbons_count"])])
    for i in range(len(edges)):
        assert np.array_equal(ora.ribbons_after(i), z["ribbons_flat"][offs[i]:offs[i + 1]]), i
