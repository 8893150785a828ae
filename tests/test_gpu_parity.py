"""GPU parity: the CUDA engine (through the C ABI, host buffers) against the CPU oracle on the
same seeded inputs.  Bit-exact for word choice, flags, sample / check-point counts and ribbon
counts; 1e-9 relative for costs, end states and ribbon coordinates (BASELINE.json north_star)."""
import math

import numpy as np
import pytest

from path_planner_b200 import EdgeEngine, abi, synth
from tests import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["thread+warp", "warp-only"])
def engine(request):
    """Both evaluation paths must produce the oracle's bits: the default pipeline (K2t thread walker for simple edges,
    K2b warp walker for the heavy list) and the warp walker alone (PPE_THREAD_WALKER=0, read at ppe_create)."""
    import os
    old = os.environ.get("PPE_THREAD_WALKER")
    if request.param == "warp-only":
        os.environ["PPE_THREAD_WALKER"] = "0"
    else:
        os.environ.pop("PPE_THREAD_WALKER", None)
    try:
        return EdgeEngine(0)
    finally:
        if old is None:
            os.environ.pop("PPE_THREAD_WALKER", None)
        else:
            os.environ["PPE_THREAD_WALKER"] = old


@pytest.fixture(scope="module")
def oracle():
    # oracle-B: every edge must match (discrete fields exactly, continuous ones to 1e-9)
    return common.load_oracle("cr")


@pytest.fixture(scope="module")
def oracle_glibc():
    # oracle-A == the compiled reference bit for bit; glibc's non-correctly-rounded last bits make
    # a few edges per 100k take the other side of the end-sample retry (DubinsWrapper.cpp:39-42)
    return common.load_oracle("glibc")


# measured on the CPU (tests/test_oracle_variants.py): 4e-5 of edges; allow 5x head-room (and never less than one edge)
GLIBC_EDGE_BUDGET = 2e-4


def _check_batch(engine, oracle, world, edges, ribbon_lists=200):
    sid_o = world.upload(oracle)
    sid_e = world.upload(engine)
    assert sid_o == sid_e
    edges["ribbon_set"] = sid_e
    want = oracle.true_cost_batch(edges)
    got = engine.true_cost_batch(edges)
    bad = common.diff_results(got, want)
    assert not bad, common.describe(bad, got, want)
    changed = np.flatnonzero(want["ribbons_changed"])
    for i in changed[:ribbon_lists]:
        a = engine.ribbons_after(i)
        b = oracle.ribbons_after(i)
        assert a.shape == b.shape, (i, a.shape, b.shape)
        assert np.allclose(a, b, rtol=common.RTOL, atol=common.ATOL), (i, a, b)
    # unchanged edges hand back the parent's set
    unchanged = np.flatnonzero(want["ribbons_changed"] == 0)
    for i in unchanged[:5]:
        assert np.array_equal(engine.ribbons_after(i), oracle.ribbons_after(i))
    # K3: best feasible f
    f, idx = engine.best()
    ok = (want["infeasible"] == 0) & (want["status"] == 0)
    if ok.any():
        fo = (got["g"] + got["h"])[ok]
        assert idx >= 0 and ok[idx]
        assert f == fo.min()
        assert idx == np.flatnonzero(ok)[np.argmin(fo)]
    else:
        assert idx == -1 and math.isinf(f)
    return got, want


def _vs_glibc_reference(engine, oracle_glibc, world, edges, got, label):
    """GPU vs oracle-A (glibc libm == the compiled reference bit for bit).  The engine's Dubins transcendentals are
    correctly rounded, glibc's are faithful (< 1 ulp): on a few edges per 100 k the last bit differs and with it the
    side of the 1e-5 m end-sample retry (DubinsWrapper.cpp:39-42).  Costs, flags, status, sample counts and end
    times must match on EVERY edge; every other field -- all discrete ones included -- on all but a measured handful,
    and the measured counts are printed per config."""
    n = len(edges)
    world.upload(oracle_glibc)
    ref = oracle_glibc.true_cost_batch(edges)
    bad = common.diff_results(got, ref)
    for name_ in ("true_cost", "collision_penalty", "g", "infeasible", "status", "n_samples", "w_end_time"):
        assert name_ not in bad, common.describe(bad, got, ref)
    counts = {k: int(len(v)) for k, v in bad.items()}
    print("[glibc-parity] %s n=%d differing edges per field: %s" % (label, n, counts or "none"))
    budget = max(1, int(GLIBC_EDGE_BUDGET * n))
    for name_ in common.DISCRETE:  # word, check-point count, ribbon count, changed flag: a handful at most
        assert counts.get(name_, 0) <= budget, (name_, counts, common.describe(bad, got, ref))
    idx = set()
    for v in bad.values():
        idx |= set(v.tolist())
    assert len(idx) <= max(1, int(GLIBC_EDGE_BUDGET * n * 5)), common.describe(bad, got, ref)
    return counts


@pytest.mark.parametrize("name,near,n", [("c1", 0.5, 3000), ("c2", 0.0, 4000), ("c2", 0.6, 4000),
                                          ("c3", 0.3, 1000), ("c3b", 0.3, 1500), ("c4", 0.5, 2000), ("c5", 0.2, 600)])
def test_true_cost_matches_oracle(engine, oracle, oracle_glibc, name, near, n):
    world = synth.WORLDS[name]()
    edges = synth.make_edges(world, n, seed=11, near_ribbons=near)
    got, want = _check_batch(engine, oracle, world, edges)
    assert (want["status"] == 0).all()
    _vs_glibc_reference(engine, oracle_glibc, world, edges, got, "%s near=%.1f" % (name, near))


# ---- configurations the first round never ran through the CUDA path ("What's weak" #2) ----------------
@pytest.mark.parametrize("name,near,n", [("c2", 0.5, 3000), ("c3", 0.3, 800), ("c3b", 0.3, 1000)])
def test_epoch_scale_state_times(engine, oracle, oracle_glibc, name, near, n):
    """State times ~1.7e9 s, what the ROS node feeds: one binade, ulp(t) = 2.4e-7 s, so `t += dt` (Edge.cpp:173)
    rounds every step and the sample count / int(ribbonsDoneTime) depend on it."""
    world = synth.with_time_offset(synth.WORLDS[name](), 1.7e9)
    edges = synth.make_edges(world, n, seed=12, near_ribbons=near)
    assert edges["src"][:, 4].min() > 1.6e9
    got, want = _check_batch(engine, oracle, world, edges)
    assert (want["status"] == 0).all() and want["ribbons_changed"].sum() > 0
    _vs_glibc_reference(engine, oracle_glibc, world, edges, got, world.name)


@pytest.mark.parametrize("name,res,near,n", [("c2", 0.5, 0.3, 3000), ("c2", 2.5, 0.3, 3000), ("c2", 0.3, 0.3, 2000),
                                              ("c4", 0.5, 0.3, 1500), ("c4", 2.5, 0.3, 1500), ("c5", 2.5, 0.2, 500)])
def test_map_resolutions(engine, oracle, oracle_glibc, name, res, near, n):
    """GridWorldMap::isBlocked divides by the resolution (GridWorldMap.cpp:84-93): 0.5 m is the exact-reciprocal path,
    2.5 m and 0.3 m the division path; the dilation radius of the chunk-culling safe map depends on it too."""
    world = synth.with_resolution(synth.WORLDS[name](), res)
    edges = synth.make_edges(world, n, seed=13, near_ribbons=near)
    got, want = _check_batch(engine, oracle, world, edges)
    assert (want["status"] == 0).all()
    assert (want["infeasible"] == 1).any() and (want["infeasible"] == 0).any()
    _vs_glibc_reference(engine, oracle_glibc, world, edges, got, world.name)


@pytest.mark.parametrize("heuristic", [abi.H_TSP_POINT_ROBOT_NO_SPLIT_K, abi.H_TSP_POINT_ROBOT_NO_SPLIT_ALL])
@pytest.mark.parametrize("n_ribbons", [1, 3, 5])
def test_tsp_heuristics_on_device(engine, oracle, oracle_glibc, heuristic, n_ribbons):
    """Executive's default heuristic (TspPointRobotNoSplitKRibbons, K = 2, executive.cpp:391) and the all-ribbons variant
    are evaluated by the kernels on the ribbons-after of every edge (RibbonManager.cpp:53-95): h within 1e-9, and the
    prune record of ppe_best is available for them."""
    from tests.test_oracle_vs_ref import tsp_world
    world = tsp_world(heuristic, n_ribbons)
    edges = synth.make_edges(world, 2500, seed=15, near_ribbons=0.7)
    got, want = _check_batch(engine, oracle, world, edges)
    assert (got["h"] >= 0).all() and want["n_ribbons_after"].max() > n_ribbons
    _vs_glibc_reference(engine, oracle_glibc, world, edges, got, world.name)


def test_non_default_covariances(engine, oracle, oracle_glibc):
    """Per-obstacle SPD covariances instead of the manager default (GaussianDynamicObstaclesManager.h:24-25,39-44)."""
    for base in ("c3", "c5"):
        world = synth.with_covariances(synth.WORLDS[base]())
        edges = synth.make_edges(world, 800 if base == "c3" else 500, seed=14, near_ribbons=0.2)
        got, want = _check_batch(engine, oracle, world, edges)
        assert (want["collision_penalty"] > 0).any()
        _vs_glibc_reference(engine, oracle_glibc, world, edges, got, world.name)


def test_previous_plan_style_wrappers(engine, oracle, oracle_glibc):
    """has_path edges as AStarPlanner.cpp:46-59 builds them from a previous plan: wrapper started earlier than the
    source vertex, end time truncated, and a foreign radius that forces the re-solve of Edge.cpp:78-80
    (mirrors tests/test_oracle_vs_ref.py::test_oracle_has_path_edges_and_dubins_match_the_reference)."""
    world = synth.world_c2()
    n = 3000
    edges = synth.make_edges(world, n, seed=23, near_ribbons=0.5)
    cfg = world.cfg
    q0 = np.column_stack([edges["src"][:, 0], edges["src"][:, 1], _yaw(edges["src"][:, 2])])
    q1 = np.column_stack([edges["dst"][:, 0], edges["dst"][:, 1], _yaw(edges["dst"][:, 2])])
    rho = np.where(edges["coverage_allowed"] == 1, cfg.coverage_turning_radius, cfg.turning_radius)
    world.upload(engine)
    typ, par, length, err = engine.dubins_batch(q0, q1, rho)
    edges["has_path"] = 1
    edges["path_qi"] = q0
    edges["path_param"] = par
    edges["path_rho"] = rho
    edges["path_type"] = typ
    edges["w_speed"] = edges["dst"][:, 3]
    edges["w_start_time"] = edges["src"][:, 4]
    edges["w_end_time"] = edges["src"][:, 4] + length / edges["dst"][:, 3]
    k = np.arange(n) % 3 == 0
    edges["w_start_time"][k] -= 0.5
    edges["w_end_time"][k] = edges["w_start_time"][k] + 0.7 * length[k] / edges["dst"][k, 3]
    edges["path_rho"][::50] = 11.0
    got, want = _check_batch(engine, oracle, world, edges)
    _vs_glibc_reference(engine, oracle_glibc, world, edges, got, "c2 previous-plan wrappers")


def test_has_path_edges_match_oracle(engine, oracle):
    """Winner edges of expand(): pre-solved wrapper + speed change (SamplingBasedPlanner.cpp:134-149)."""
    world = synth.world_c2()
    edges = synth.make_edges(world, 3000, seed=3, near_ribbons=0.4)
    cfg = world.cfg
    q0 = np.column_stack([edges["src"][:, 0], edges["src"][:, 1], _yaw(edges["src"][:, 2])])
    q1 = np.column_stack([edges["dst"][:, 0], edges["dst"][:, 1], _yaw(edges["dst"][:, 2])])
    rho = np.where(edges["coverage_allowed"] == 1, cfg.coverage_turning_radius, cfg.turning_radius)
    world.upload(engine)
    typ, par, length, err = engine.dubins_batch(q0, q1, rho)
    assert (err == 0).all()
    edges["has_path"] = 1
    edges["path_qi"] = q0
    edges["path_param"] = par
    edges["path_rho"] = rho
    edges["path_type"] = typ
    edges["w_speed"] = edges["dst"][:, 3]
    edges["w_start_time"] = edges["src"][:, 4]
    edges["w_end_time"] = edges["src"][:, 4] + length / edges["dst"][:, 3]
    _check_batch(engine, oracle, world, edges)


def _yaw(h):
    y = math.pi / 2 - h
    return np.where(y < 0, y + 2 * math.pi, y)


def test_dubins_batch_matches_oracle(engine, oracle):
    n = 1 << 20
    rng = np.random.default_rng(0)
    q0 = np.column_stack([rng.uniform(-100, 100, n), rng.uniform(-100, 100, n), rng.uniform(0, 2 * math.pi, n)])
    q1 = np.column_stack([q0[:, 0] + rng.uniform(-75, 75, n), q0[:, 1] + rng.uniform(-75, 75, n), rng.uniform(0, 2 * math.pi, n)])
    q1[: n // 4, 0] = q0[: n // 4, 0] + rng.uniform(-10, 10, n // 4)
    q1[: n // 4, 1] = q0[: n // 4, 1] + rng.uniform(-10, 10, n // 4)
    rho = np.where(np.arange(n) % 2 == 0, 8.0, 16.0)
    engine.set_config(abi.PpeConfig())
    oracle.set_config(abi.PpeConfig())
    gt, gp, gl, ge = engine.dubins_batch(q0, q1, rho)
    wt, wp, wl, we = oracle.dubins_batch(q0, q1, rho)
    assert np.array_equal(ge, we)
    mism = np.flatnonzero(gt != wt)
    assert mism.size == 0, "word mismatches: %s" % mism[:10]
    assert np.allclose(gp, wp, rtol=1e-9, atol=1e-9)
    assert np.allclose(gl, wl, rtol=1e-9, atol=1e-9)
    assert np.bincount(gt, minlength=6).min() > 1000  # every word exercised


def test_degenerate_dubins(engine, oracle):
    """Co-located, collinear, d ~ 0 and tie cases: straight lines tie LSL/RSR/LSR/RSL -> LSL wins."""
    q0 = [[0, 0, math.pi / 2], [0, 0, 0], [0, 0, 0], [5, 5, 1.0], [0, 0, math.pi / 2], [0, 0, 0.0], [0, 0, 0.0]]
    q1 = [[0, 5, math.pi / 2], [25, 0, 0], [0, 0, 0], [5, 5, 2.0], [0, 16, 3 * math.pi / 2], [0, 1e-9, 0.0], [-10, 0, 0.0]]
    rho = [8, 8, 8, 8, 8, 8, 8]
    engine.set_config(abi.PpeConfig())
    oracle.set_config(abi.PpeConfig())
    gt, gp, gl, ge = engine.dubins_batch(q0, q1, rho)
    wt, wp, wl, we = oracle.dubins_batch(q0, q1, rho)
    assert np.array_equal(gt, wt) and np.array_equal(ge, we)
    assert np.allclose(gp, wp, rtol=1e-9, atol=1e-12)
    assert gt[0] == abi.LSL and gt[1] == abi.LSL
    assert gl[0] == 5 and gl[1] == 25


def test_deep_thread_walker_matches_oracle(oracle):
    """K2c (PPE_DEEP_WALKER=1, off by default): the edges K2t catches covering a ribbon are walked one thread per edge with a
    two-entry override table over the parent's ribbon list; what changes the list's structure goes on to the warp walker.
    Same bits as the oracle, and some edges must actually have taken that path (bit 25 of the instrumentation word)."""
    import os
    os.environ["PPE_DEEP_WALKER"] = "1"
    try:
        eng = EdgeEngine(0)
    finally:
        os.environ.pop("PPE_DEEP_WALKER", None)
    deep = 0
    for name, near, n in (("c2", 0.6, 4000), ("c1", 0.5, 2000), ("c5", 0.3, 600)):
        world = synth.WORLDS[name]()
        edges = synth.make_edges(world, n, seed=21, near_ribbons=near)
        got, want = _check_batch(eng, oracle, world, edges)
        deep += int(((got["reserved"] >> 25) & 1).sum())
    assert deep > 0


@pytest.mark.parametrize("name,near,n", [("c2", 0.6, 9000), ("c5", 0.3, 5000)])
def test_host_buffer_pipeline_matches_oracle(oracle, name, near, n):
    """Large host-buffer batches run K2a + K2t slice by slice under the copies and K2b once at the end, whose records reach the
    caller's array by a scatter kernel (pinned, mapped memory) or a packed copy + host scatter (pageable).  Forced here with a
    small slice (PPE_LATE_SLICE, read at ppe_create): both variants must give the oracle's records, the ribbons-after pool
    and the K3 record, and must agree with the one-launch-group path byte for byte."""
    import ctypes as C
    import os
    import torch
    os.environ["PPE_LATE_SLICE"] = "1024"
    try:
        eng = EdgeEngine(0)
    finally:
        os.environ.pop("PPE_LATE_SLICE", None)
    world = synth.WORLDS[name]()
    edges = synth.make_edges(world, n, seed=33, near_ribbons=near)
    launches0 = eng.launch_count()
    got, want = _check_batch(eng, oracle, world, edges)  # pageable numpy buffers
    slices = (n + 1023) // 1024
    assert eng.launch_count() - launches0 >= 2 * slices + 3, "the sliced pipeline did not run"
    assert int(((got["reserved"] >> 24) & 1).sum()) < n, "no edge went to K2b: the scatter was not exercised"
    # pinned buffers: K2b's records are written into the caller's array by the scatter kernel
    h_edges = torch.from_numpy(edges.view(np.uint8).reshape(n, -1).copy()).pin_memory()
    h_res = torch.zeros((n, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    rc = eng._lib.ppe_true_cost_batch(eng._ctx, n, C.c_void_p(h_edges.data_ptr()), C.c_void_p(h_res.data_ptr()))
    assert rc == 0
    pinned = h_res.numpy().view(abi.RESULT_DTYPE).reshape(n)
    # the ribbons-after pool is filled in whatever order the warps finish: offsets differ run to run, nothing else may
    a, b = pinned.copy(), got.copy()
    a["ribbons_offset"] = 0
    b["ribbons_offset"] = 0
    assert a.tobytes() == b.tobytes()
    # and the one-launch-group path (default slice: a batch this small is not pipelined)
    ref = EdgeEngine(0)
    sid = world.upload(ref)
    assert sid == edges["ribbon_set"][0]
    c = ref.true_cost_batch(edges)
    c["ribbons_offset"] = 0
    assert c.tobytes() == b.tobytes()


@pytest.mark.parametrize("pipeline", ["late-k2b", "sliced-whole"])
def test_ribbons_after_pool_overflow_is_retried(oracle, pipeline):
    """The ribbons-after pool is bump-allocated; a batch that outgrows it is run again with a larger pool (both host-buffer
    pipelines, forced here with a 64-ribbon pool and small slices).  The caller sees the oracle's records and lists."""
    import os
    env = {"PPE_RIBBON_POOL": "64", "PPE_LATE_SLICE": "1024", "PPE_SLICE_EDGES": "1024"}
    if pipeline == "sliced-whole":
        env["PPE_LATE_K2B"] = "0"
    os.environ.update(env)
    try:
        eng = EdgeEngine(0)
        world = synth.WORLDS["c2"]()
        edges = synth.make_edges(world, 6000, seed=44, near_ribbons=0.6)
        got, want = _check_batch(eng, oracle, world, edges)
    finally:
        for k in env:
            os.environ.pop(k, None)
    assert int(want["ribbons_changed"].sum()) * 5 > 64, "the batch does not overflow a 64-ribbon pool"
