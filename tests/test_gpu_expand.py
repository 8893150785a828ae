"""GPU parity of the frontier expansion: ppe_add_samples / ppe_expand_batch (path_planner_b200/csrc/ppe_expand.cu)
against the CPU restatement of SamplingBasedPlanner::addSamples / ::expand (oracle/ppe_oracle_expand.c) on the same
seeded vertices and samples.  Which samples win the k-nearest selection, in which heap order, which Dubins word, how many
samples the loop popped and every flag are bit-exact; costs and end states within 1e-9 (BASELINE.json north_star)."""
import math
import os

import numpy as np
import pytest

from path_planner_b200 import EdgeEngine, abi, synth
from tests import common

pytestmark = pytest.mark.gpu

CHILD_DISCRETE = ("sample_index", "path_type", "infeasible", "status", "coverage_allowed", "n_ribbons_after", "ribbons_changed")
CHILD_CONT = ("true_cost", "collision_penalty", "approx_cost", "end", "g", "h", "coverage_completed_time", "path_param", "w_end_time")


@pytest.fixture(scope="module")
def engine():
    return EdgeEngine(0)


@pytest.fixture(scope="module")
def oracle():
    return common.load_oracle("cr")


def make_samples(world, n, seed, centre, reach=75.0):
    """What StateGenerator draws (StateGenerator.cpp:15-31): uniform in the box around the start, 1 % on a ribbon."""
    rng = np.random.default_rng(seed)
    x = centre[0] + rng.uniform(-reach, reach, n)
    y = centre[1] + rng.uniform(-reach, reach, n)
    h = rng.uniform(0, 2 * math.pi, n)
    on = np.flatnonzero(rng.uniform(size=n) < 0.01)
    if len(world.ribbons) and on.size:
        rb = world.ribbons[rng.integers(0, len(world.ribbons), on.size)]
        u = rng.uniform(0, 1, on.size)
        x[on] = rb[:, 0] + u * (rb[:, 2] - rb[:, 0])
        y[on] = rb[:, 1] + u * (rb[:, 3] - rb[:, 1])
        h[on] = np.mod(math.pi / 2 - np.arctan2(rb[:, 3] - rb[:, 1], rb[:, 2] - rb[:, 0]), 2 * math.pi)
    return x, y, h


def make_vertices(world, n, seed, centre, set_id, reach=60.0):
    rng = np.random.default_rng(seed)
    v = np.zeros(n, dtype=abi.VERTEX_DTYPE)
    todo = np.arange(n)
    while todo.size:
        x = centre[0] + rng.uniform(-reach, reach, todo.size)
        y = centre[1] + rng.uniform(-reach, reach, todo.size)
        ok = ~world.is_blocked(x, y)
        v["state"][todo[ok], 0] = x[ok]
        v["state"][todo[ok], 1] = y[ok]
        todo = todo[~ok]
    v["state"][:, 2] = rng.uniform(0, 2 * math.pi, n)
    v["state"][:, 3] = world.cfg.max_speed
    v["state"][:, 4] = world.cfg.start_state_time + rng.uniform(0, 15, n)
    v["g"] = rng.uniform(0, 20, n)
    v["ribbon_set"] = set_id
    # nearest ribbon end point pulled inside, both headings (a stand-in for getNearestEndpointAsState: any state works)
    he = rng.uniform(size=n) < 0.85
    rb = world.ribbons[rng.integers(0, len(world.ribbons), n)]
    v["has_endpoint"] = he.astype(np.int32)
    v["endpoint"][:, 0] = rb[:, 0]
    v["endpoint"][:, 1] = rb[:, 1] + 2.0
    v["endpoint"][:, 2] = np.mod(math.pi / 2 - np.arctan2(rb[:, 3] - rb[:, 1], rb[:, 2] - rb[:, 0]), 2 * math.pi)
    return v


def _compare(engine, oracle, world, n_samples, n_vertices, seed, duplicate=False):
    centre = world.start[:2]
    sid_e = world.upload(engine)
    sid_o = world.upload(oracle)
    assert sid_e == sid_o
    x, y, h = make_samples(world, n_samples, seed, centre)
    if duplicate:  # two samples at exactly the same place: an exact distance tie for every vertex that pops them
        x[1::7] = x[0::7][: len(x[1::7])]
        y[1::7] = y[0::7][: len(y[1::7])]
    engine.clear_samples()
    oracle.clear_samples()
    # added in two calls, as the anytime loop does (initial samples, then doubling)
    half = n_samples // 2
    ke = np.concatenate([engine.add_samples(x[:half], y[:half], h[:half]), engine.add_samples(x[half:], y[half:], h[half:])])
    ko = np.concatenate([oracle.add_samples(x[:half], y[:half], h[:half]), oracle.add_samples(x[half:], y[half:], h[half:])])
    assert np.array_equal(ke, ko)
    assert np.array_equal(ke, ~world.is_blocked(x, y))
    assert engine.sample_count() == oracle.sample_count() == int(ke.sum())
    verts = make_vertices(world, n_vertices, seed + 1, centre, sid_e)
    gn, gc, gf, gp, gpool = engine.expand_batch(verts)
    wn, wc, wf, wp, wpool = oracle.expand_batch(verts)
    assert np.array_equal(gn, wn), (gn, wn)
    assert np.array_equal(gf, wf), (gf, wf)
    assert np.array_equal(gp, wp), (gp, wp)
    for v in range(n_vertices):
        a, b = gc[v, : gn[v]], wc[v, : wn[v]]
        for name in CHILD_DISCRETE:
            assert np.array_equal(a[name], b[name]), (v, name, a[name], b[name])
        for name in CHILD_CONT:
            assert np.allclose(a[name], b[name], rtol=common.RTOL, atol=common.ATOL, equal_nan=True), (v, name, a[name], b[name])
        assert (gc[v, gn[v]:]["status"] == abi.EDGE_SKIPPED).all()
        for c in np.flatnonzero(b["ribbons_changed"] == 1):
            ra = gpool[a["ribbons_offset"][c]: a["ribbons_offset"][c] + a["n_ribbons_after"][c]]
            rb_ = wpool[b["ribbons_offset"][c]: b["ribbons_offset"][c] + b["n_ribbons_after"][c]]
            assert ra.shape == rb_.shape and np.allclose(ra, rb_, rtol=common.RTOL, atol=common.ATOL), (v, c)
    return gn, gc, gf, gp


@pytest.mark.parametrize("name,n_samples,n_vertices", [("c1", 1000, 40), ("c2", 100, 30), ("c2", 12800, 64), ("c3", 3200, 48),
                                                       ("c3b", 6400, 32), ("c4", 10000, 96), ("c5", 40000, 64)])
def test_expand_batch_matches_oracle(engine, oracle, name, n_samples, n_vertices):
    world = synth.WORLDS[name]()
    gn, gc, gf, gp = _compare(engine, oracle, world, n_samples, n_vertices, seed=31)
    assert (gf == 0).all()
    assert gn.max() == 4 + 4 * world.cfg.branching_factor  # 2 x 2 end-point edges + 2 radii x k winners x 2 speeds
    assert (gp > 2 * world.cfg.branching_factor).all()


def test_expand_batch_large_sample_set(engine, oracle):
    """The deep end of the anytime loop: 6e5 samples (the doubling of AStarPlanner.cpp:101-102 after ~13 iterations);
    thousands of candidates per vertex go through the sort and the replay."""
    world = synth.world_c2()
    gn, gc, gf, gp = _compare(engine, oracle, world, 600000, 12, seed=33)
    assert (gf == 0).all() and gp.min() > 500


def test_expand_batch_flags_exact_distance_ties(engine, oracle):
    world = synth.world_c2()
    gn, gc, gf, gp = _compare(engine, oracle, world, 4000, 40, seed=35, duplicate=True)
    assert (gf & abi.EXPAND_TIE).any()


def test_expand_batch_search_radius_retries(oracle):
    """The candidate collection starts from a density-based search disc and widens / shrinks it on its own; a far too small
    and a far too large first guess (PPE_EXPAND_R2, read per call) must give the same children."""
    world = synth.world_c2()
    for r2 in ("0.01", "1e9"):
        os.environ["PPE_EXPAND_R2"] = r2
        try:
            eng = EdgeEngine(0)
            _compare(eng, oracle, world, 30000, 16, seed=37)
        finally:
            os.environ.pop("PPE_EXPAND_R2", None)


def test_branching_factors_and_equal_radii(engine, oracle):
    """k = 3 and k = 16 (the device limit); coverage radius == turning radius and slow speed == max speed drop the second
    configuration (SamplingBasedPlanner.cpp:58-63)."""
    for k, same in ((3, False), (16, False), (9, True)):
        world = synth.world_c2()
        world.cfg.branching_factor = k
        if same:
            world.cfg.coverage_turning_radius = world.cfg.turning_radius
            world.cfg.slow_speed = world.cfg.max_speed
        gn, gc, gf, gp = _compare(engine, oracle, world, 5000, 24, seed=39)
        per = (1 if same else 2) * (1 if same else 2)
        assert gn.max() == per + per * k


def test_map_tile_staging_gives_identical_results(oracle):
    """N2 A/B variant (PPE_MAP_TILE=1): the thread walker reads the occupancy / safe bitmaps of the batch's bounding box
    from a shared-memory tile staged with TMA bulk copies; every result must equal the global-memory path bit for bit --
    for a frontier batch (window derived from the vertices) and for an edge sweep with an explicit window, including edges
    that leave the window (they fall back to the L2 path)."""
    world = synth.world_c4()
    os.environ["PPE_MAP_TILE"] = "1"
    try:
        tiled = EdgeEngine(0)
    finally:
        os.environ.pop("PPE_MAP_TILE", None)
    plain = EdgeEngine(0)
    _compare(tiled, oracle, world, 10000, 48, seed=41)
    cx, cy = world.start[0], world.start[1]
    edges = synth.make_edges(world, 20000, seed=43, near_ribbons=0.0, extent=None)
    # pull the sources into a 300 m box around the start so that the window matters
    edges["src"][:, 0] = cx + (edges["src"][:, 0] % 300.0) - 150.0
    edges["src"][:, 1] = cy + (edges["src"][:, 1] % 300.0) - 150.0
    edges["dst"][:, 0] = edges["src"][:, 0] + (edges["dst"][:, 0] % 120.0) - 60.0
    edges["dst"][:, 1] = edges["src"][:, 1] + (edges["dst"][:, 1] % 120.0) - 60.0
    for eng in (tiled, plain):
        edges["ribbon_set"] = world.upload(eng)
    tiled._lib.ppe_set_map_window.argtypes = [__import__("ctypes").c_void_p] + [__import__("ctypes").c_double] * 4
    tiled._lib.ppe_set_map_window(tiled._ctx, cx - 180.0, cy - 180.0, cx + 180.0, cy + 180.0)  # some edges reach outside
    a = tiled.true_cost_batch(edges)
    b = plain.true_cost_batch(edges)
    for name in abi.RESULT_DTYPE.names:
        if name not in ("ribbons_offset",):
            assert np.array_equal(a[name], b[name]), name
    assert (a["infeasible"] == 1).any() and (a["infeasible"] == 0).any()
