"""Python binding of the standalone planning harness (include/ppe_harness.h, path_planner_b200/libppe_harness.so):
the reference's Planner / Edge / Vertex / State classes driven by BatchedAStarPlanner over the B200 edge engine.
Built by path_planner_b200/harness/Makefile where the reference sources exist; the .so travels with the repo.
No arithmetic lives here."""
import ctypes as C
import os

import numpy as np

from . import abi
from ._capi import PpeError

_HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS_PATH = os.path.join(_HERE, "libppe_harness.so")
PLAN_MAIN_PATH = os.path.join(_HERE, "ppe_plan_main")

# pph_dubins_path: DubinsPath.msg field order + end time
PATH_DTYPE = np.dtype([("initial_x", "<f8"), ("initial_y", "<f8"), ("initial_yaw", "<f8"), ("length0", "<f8"), ("length1", "<f8"),
                       ("length2", "<f8"), ("rho", "<f8"), ("type", "<i4"), ("pad", "<i4"), ("speed", "<f8"), ("start_time", "<f8"),
                       ("end_time", "<f8")], align=True)
assert PATH_DTYPE.itemsize == 88


class PlanOptions(C.Structure):
    _fields_ = [("time_remaining", C.c_double), ("clock0", C.c_double), ("tick", C.c_double), ("sample_tick", C.c_double),
                ("initial_samples", C.c_int32),
                ("use_brown_paths", C.c_int32), ("frontier", C.c_int32), ("knn_chunk", C.c_int32), ("visualize", C.c_int32),
                ("reserved", C.c_int32), ("visualization_path", C.c_char_p)]


class PlanStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("samples", "generated", "expanded", "iterations", "plan_depth")] + \
               [(n, C.c_double) for n in ("plan_f", "plan_collision_penalty", "plan_time_penalty", "plan_h", "plan_endtime")] + \
               [(n, C.c_uint64) for n in ("now_calls", "true_cost_edges", "dubins_solves", "engine_batches", "frontier_vertices",
                                          "frontier_hits", "exact_expansions")] + \
               [(n, C.c_double) for n in ("wall_seconds", "seconds_engine_expand", "seconds_replay", "seconds_add_samples", "seconds_exact")] + \
               [(n, C.c_uint64) for n in ("exact_for_ties", "exact_for_overflow")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def available():
    return os.path.exists(HARNESS_PATH)


_libs = {}


def load_library(path=None):
    """`path`: another build of the same C ABI (tests bind the CPU test double of the engine through it)."""
    path = path or HARNESS_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise PpeError("%s is missing: build it with `make -C path_planner_b200/harness` (needs the reference sources)" % path)
    lib = C.CDLL(path)
    D = C.POINTER(C.c_double)
    lib.pph_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.pph_destroy.argtypes = [C.c_void_p]
    lib.pph_destroy.restype = None
    lib.pph_last_error.argtypes = [C.c_void_p]
    lib.pph_last_error.restype = C.c_char_p
    lib.pph_set_config.argtypes = [C.c_void_p, C.POINTER(abi.PpeConfig)]
    lib.pph_set_map_none.argtypes = [C.c_void_p]
    lib.pph_set_map_bitmap.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
    lib.pph_load_gridworld_map.argtypes = [C.c_void_p, C.c_char_p]
    lib.pph_set_obstacles_none.argtypes = [C.c_void_p]
    lib.pph_set_obstacles_binary.argtypes = [C.c_void_p, C.c_int, D, D, D, D, D, D, D]
    lib.pph_set_obstacles_gaussian.argtypes = [C.c_void_p, C.c_int, D, D, D, D, D, D]
    lib.pph_set_ribbons.argtypes = [C.c_void_p, C.c_int, D]
    lib.pph_plan.argtypes = [C.c_void_p, D, C.c_void_p, C.c_int, C.POINTER(PlanOptions), C.c_void_p, C.c_int, C.POINTER(PlanStats)]
    lib.pph_advance.argtypes = [C.c_void_p, C.c_double, D]
    lib.pph_write_plan_msg.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    _libs[path] = lib
    return lib


class PlanningHarness:
    """One planning thread: a pph_ctx (its own ppe_ctx on `device`)."""

    def __init__(self, device=0, lib_path=None):
        self._lib = load_library(lib_path)
        self._ctx = C.c_void_p()
        rc = self._lib.pph_create(int(device), C.byref(self._ctx))
        if rc != abi.PPE_OK:
            raise PpeError("pph_create(device=%d) failed (%d): no usable sm_100 CUDA device; the engine has no CPU path" % (device, rc))

    def close(self):
        if self._ctx:
            self._lib.pph_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise PpeError("pph_%s failed (%d): %s" % (what, rc, (self._lib.pph_last_error(self._ctx) or b"").decode()))
        return rc

    def set_world(self, world):
        """`world`: path_planner_b200.synth.World (config, map, obstacles in the managers' heading convention, ribbons)."""
        self._check(self._lib.pph_set_config(self._ctx, C.byref(world.cfg)), "set_config")
        if world.map_bits is None:
            self._check(self._lib.pph_set_map_none(self._ctx), "set_map_none")
        else:
            bits = np.ascontiguousarray(world.map_bits, dtype=np.uint8)
            self._check(self._lib.pph_set_map_bitmap(self._ctx, abi.vptr(bits), world.rows, world.cols, bits.shape[1], float(world.resolution)),
                        "set_map_bitmap")
        o = world.obstacles
        if world.obstacle_kind == "none":
            self._check(self._lib.pph_set_obstacles_none(self._ctx), "set_obstacles_none")
        else:
            a = [np.ascontiguousarray(o[k], dtype=np.float64) for k in ("x", "y", "heading", "speed", "time")]
            if world.obstacle_kind == "binary":
                a += [np.ascontiguousarray(o[k], dtype=np.float64) for k in ("width", "length")]
                self._check(self._lib.pph_set_obstacles_binary(self._ctx, len(a[0]), *[abi.dptr(v) for v in a]), "set_obstacles_binary")
            else:
                cov = o.get("cov")
                covp = abi.dptr(np.ascontiguousarray(cov, dtype=np.float64).reshape(-1, 4)) if cov is not None else None
                self._check(self._lib.pph_set_obstacles_gaussian(self._ctx, len(a[0]), *[abi.dptr(v) for v in a], covp), "set_obstacles_gaussian")
        rib = np.ascontiguousarray(world.ribbons, dtype=np.float64).reshape(-1, 4)
        self._check(self._lib.pph_set_ribbons(self._ctx, rib.shape[0], abi.dptr(rib)), "set_ribbons")

    def plan(self, start, time_remaining, clock0=0.0, tick=0.0, initial_samples=100, brown=0, frontier=-1, knn_chunk=0, previous=None,
             visualization_path=None, cap=256, sample_tick=0.0):
        """Planner::plan.  Returns (plan records [n] of PATH_DTYPE, stats dict)."""
        start = np.ascontiguousarray(start, dtype=np.float64)
        opt = PlanOptions(time_remaining, clock0, tick, sample_tick, initial_samples, brown, frontier, knn_chunk, 1 if visualization_path else 0, 0,
                          visualization_path.encode() if visualization_path else None)
        prev = np.ascontiguousarray(previous if previous is not None else np.zeros(0, dtype=PATH_DTYPE), dtype=PATH_DTYPE)
        out = np.zeros(cap, dtype=PATH_DTYPE)
        st = PlanStats()
        n = self._check(self._lib.pph_plan(self._ctx, abi.dptr(start), abi.vptr(prev) if len(prev) else None, len(prev), C.byref(opt),
                                           abi.vptr(out), cap, C.byref(st)), "plan")
        return out[:min(n, cap)].copy(), st.as_dict()

    def advance(self, time):
        s = np.zeros(5)
        self._check(self._lib.pph_advance(self._ctx, float(time), abi.dptr(s)), "advance")
        return s

    def write_plan_msg(self, plan, path):
        plan = np.ascontiguousarray(plan, dtype=PATH_DTYPE)
        self._check(self._lib.pph_write_plan_msg(abi.vptr(plan), len(plan), path.encode()), "write_plan_msg")
