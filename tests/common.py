"""Shared helpers of the test-suite: loading the oracle libraries and diffing result batches.

oracle/_ref/libppe_oracle.so   plain-C restatement of the path (oracle/ppe_oracle.c)
oracle/_ref/libref_planner.so  the reference's own sources compiled in place (oracle/ref_shim.cpp)
Both are test infrastructure; the product (path_planner_b200) never loads them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from path_planner_b200 import abi
from path_planner_b200._capi import CApiWorld

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_ref", "libppe_oracle.so")
ORACLE_CR_SO = os.path.join(ORACLE_DIR, "_ref", "libppe_oracle_cr.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_planner.so")

DISCRETE = ("path_type", "infeasible", "status", "n_samples", "n_checkpoints", "n_ribbons_after", "ribbons_changed")
CONTINUOUS = ("true_cost", "collision_penalty", "approx_cost", "end", "g", "h", "coverage_completed_time",
              "path_qi", "path_param", "path_rho", "w_speed", "w_start_time", "w_end_time")
RTOL = 1e-9  # BASELINE.json north_star: "Edge costs must agree within 1e-9 relative"
ATOL = 1e-9  # absolute floor for quantities that are legitimately ~0 (e.g. x of a due-north path)


def build_oracles():
    """Build what can be built here: the C restatement always, the reference when its sources exist."""
    if not os.path.exists(ORACLE_SO) or not os.path.exists(ORACLE_CR_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    if not os.path.exists(REF_SO) and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)


def _load(path, prefix):
    lib = C.CDLL(path)
    create = getattr(lib, prefix + "create")
    create.argtypes = [C.POINTER(C.c_void_p)]
    ctx = C.c_void_p()
    assert create(C.byref(ctx)) == 0
    w = CApiWorld(lib, ctx, prefix)
    w.lib = lib
    w.ctx = ctx
    return w


def load_oracle(variant="glibc"):
    """variant "glibc": oracle-A, bit-identical to the compiled reference.
    variant "cr": oracle-B, the same restatement with the Dubins transcendentals correctly rounded
    (oracle/crmath_redirect.h) -- bit-identical to the engine's per-edge arithmetic."""
    build_oracles()
    w = _load(ORACLE_CR_SO if variant == "cr" else ORACLE_SO, "oracle_")
    w.lib.oracle_true_cost_batch_mt.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return w


def have_ref():
    build_oracles()
    return os.path.exists(REF_SO)


def load_ref():
    build_oracles()
    w = _load(REF_SO, "ref_")
    w.lib.ref_get_obstacles.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    w.lib.ref_true_cost_batch_mt.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    w.lib.ref_get_ribbon_set.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int]
    D = C.POINTER(C.c_double)
    w.lib.ref_plan.argtypes = [C.c_void_p, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, D, C.c_int, D]
    w.lib.ref_plan2.argtypes = [C.c_void_p, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, D, C.c_int, D,
                                C.c_int, D]
    w.lib.ref_plan3.argtypes = [C.c_void_p, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, D,
                                C.c_int, D, C.c_int, D]
    return w


def ref_obstacle_order(ref):
    buf = np.zeros((1024, 9))
    n = ref.lib.ref_get_obstacles(ref.ctx, abi.dptr(buf), 1024)
    return buf[:n].copy()


def true_cost_mt(w, edges, threads=0):
    """Multi-threaded CPU batch (oracle or reference) without keeping ribbons-after."""
    edges = np.ascontiguousarray(edges, dtype=abi.EDGE_DTYPE)
    res = np.zeros(edges.shape[0], dtype=abi.RESULT_DTYPE)
    fn = getattr(w.lib, w.prefix + "true_cost_batch_mt")
    rc = fn(w.ctx, edges.shape[0], abi.vptr(edges), abi.vptr(res), threads, 0)
    assert rc == 0
    return res


def diff_results(got, want, exact=False, check_counts=True):
    """Returns a dict field -> indices of mismatching edges.  `want` is the checker (oracle/ref).
    Edges whose status is non-zero on both sides are compared on status/infeasible only."""
    n = len(want)
    bad = {}
    ok_both = (got["status"] == 0) & (want["status"] == 0)
    for name in DISCRETE:
        if not check_counts and name in ("n_samples", "n_checkpoints"):
            continue
        a, b = got[name], want[name]
        m = a != b
        if name not in ("status", "infeasible"):
            m &= ok_both
        if m.any():
            bad[name] = np.flatnonzero(m)
    for name in CONTINUOUS:
        a = got[name].reshape(n, -1)
        b = want[name].reshape(n, -1)
        if exact:
            m = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        else:
            m = ~np.isclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True)
        m = m.any(axis=1) & ok_both
        if m.any():
            bad[name] = np.flatnonzero(m)
    return bad


def describe(bad, got, want, limit=3):
    lines = []
    for name, idx in bad.items():
        lines.append("%s: %d edges, first %s" % (name, len(idx), idx[:limit].tolist()))
        for i in idx[:limit]:
            lines.append("   [%d] got %s want %s" % (i, got[name][i], want[name][i]))
    return "\n".join(lines)


# ---- whole-plan runs: the reference's AStarPlanner and the product's BatchedAStarPlanner --------------
HARNESS_SO = os.path.join(ORACLE_DIR, "_ref", "libplan_compare.so")
PLAN_STATS = ("samples", "generated", "expanded", "iterations", "f", "collision_penalty", "time_penalty", "h", "depth",
              "now_calls", "true_cost_edges", "dubins_solves", "batches", "frontier_vertices", "frontier_hits", "exact_expansions")


def have_harness():
    return os.path.exists(HARNESS_SO)


def load_harness(path=None):
    """oracle/_ref/libplan_compare.so = the compiled reference objects + ref_shim + a TEST BUILD of the product's C++ host
    adapter (path_planner_b200/harness/BatchedAStarPlanner.cpp) linked against libppe.so.  One ref_ctx holds the
    world both planners read, so `ref_plan` and `harness_plan` see identical inputs.  (The product's own build of the
    harness is path_planner_b200/libppe_harness.so, bound in path_planner_b200/harness.py.)"""
    w = _load(path or HARNESS_SO, "ref_")
    D = C.POINTER(C.c_double)
    w.lib.ref_get_obstacles.argtypes = [C.c_void_p, D, C.c_int]
    w.lib.ref_plan.argtypes = [C.c_void_p, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, D, C.c_int, D]
    w.lib.harness_plan.argtypes = [C.c_void_p, C.c_int, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_int, D, C.c_int, D]
    w.lib.ref_plan2.argtypes = [C.c_void_p, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, D, C.c_int, D,
                                C.c_int, D]
    w.lib.ref_plan3.argtypes = [C.c_void_p, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, D,
                                C.c_int, D, C.c_int, D]
    w.lib.harness_plan2.argtypes = [C.c_void_p, C.c_int, C.c_int, D, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                    C.c_int, C.c_int, D, C.c_int, D, C.c_int, D]
    return w


def run_plan(w, which, ribbon_set, start, time_remaining, clock0, tick, initial_samples=100, brown=0, knn_chunk=128,
             device=0, cap=256, frontier=-1, previous=None, sample_tick=0.0):
    """which = "ref" (AStarPlanner) or "harness" (BatchedAStarPlanner on `device`).  tick > 0 selects the
    virtual clock (now() = clock0 + calls * tick).  `frontier`: vertices per ppe_expand_batch (-1 default, 0 = exact host
    replay).  `previous`: a plan [n, 12] handed in as previousPlan.  Returns (plan [n,12], stats dict)."""
    start = np.ascontiguousarray(start, dtype=np.float64)
    plan = np.zeros((cap, 12))
    stats = np.zeros(16)
    prev = np.ascontiguousarray(previous if previous is not None else np.zeros((0, 12)), dtype=np.float64).reshape(-1, 12)
    pp = abi.dptr(prev) if len(prev) else None
    if which == "ref":
        n = w.lib.ref_plan3(w.ctx, ribbon_set, abi.dptr(start), time_remaining, clock0, tick, sample_tick, initial_samples, brown,
                            pp, len(prev), abi.dptr(plan), cap, abi.dptr(stats))
    else:
        n = w.lib.harness_plan2(w.ctx, device, ribbon_set, abi.dptr(start), time_remaining, clock0, tick, initial_samples,
                                brown, knn_chunk, frontier, pp, len(prev), abi.dptr(plan), cap, abi.dptr(stats))
    if n < 0:
        raise RuntimeError("%s plan failed (%d): %s" % (which, n, (w.lib.ref_last_error(w.ctx) or b"").decode()))
    return plan[:min(n, cap)].copy(), dict(zip(PLAN_STATS, stats.tolist()))
