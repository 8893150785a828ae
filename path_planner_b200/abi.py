"""Binary layout of include/ppe.h for Python callers (ctypes + numpy).

Every struct here mirrors a struct in include/ppe.h field for field; `tests/test_abi.py` checks
the sizes against the compiled library (`ppe_abi_sizeof_*`).  The engine itself is the C-ABI
shared library built from path_planner_b200/csrc; nothing in this file computes anything.
"""
import ctypes as C

import numpy as np

PPE_ABI_VERSION = 3

# ppe_status
PPE_OK = 0
PPE_ERR_NO_DEVICE = -1
PPE_ERR_CUDA = -2
PPE_ERR_INVALID = -3
PPE_ERR_STATE = -4
PPE_ERR_CAPACITY = -5

# Dubins words (DubinsPath.msg:17)
LSL, LSR, RSL, RSR, RLR, LRL = range(6)
WORD_NAMES = ("LSL", "LSR", "RSL", "RSR", "RLR", "LRL")

# dubins.h error codes
EDUBOK, EDUBCOCONFIGS, EDUBPARAM, EDUBBADRHO, EDUBNOPATH = range(5)

# RibbonManager::Heuristic (RibbonManager.h:19-25)
H_MAX_DISTANCE = 0
H_TSP_POINT_ROBOT_NO_SPLIT_ALL = 1
H_TSP_POINT_ROBOT_NO_SPLIT_K = 2
H_TSP_DUBINS_NO_SPLIT_ALL = 3
H_TSP_DUBINS_NO_SPLIT_K = 4

# per-edge status
EDGE_OK = 0
EDGE_ERR_END_SAMPLE = 1
EDGE_ERR_NO_PATH = 2
EDGE_ERR_RIBBON_CAPACITY = 3


class PpeConfig(C.Structure):
    """ppe_config: PlannerConfig scalars (PlannerConfig.h:179-207), Ribbon::RibbonWidth
    (Ribbon.cpp:4) and the Edge penalty factors (Edge.h:151-152), with the reference defaults."""

    _fields_ = [
        ("max_speed", C.c_double),
        ("slow_speed", C.c_double),
        ("turning_radius", C.c_double),
        ("coverage_turning_radius", C.c_double),
        ("time_horizon", C.c_double),
        ("time_minimum", C.c_double),
        ("collision_checking_increment", C.c_double),
        ("start_state_time", C.c_double),
        ("ribbon_width", C.c_double),
        ("collision_penalty_factor", C.c_double),
        ("time_penalty_factor", C.c_double),
        ("heuristic", C.c_int32),
        ("branching_factor", C.c_int32),
        ("tsp_k", C.c_int32),
        ("reserved0", C.c_int32),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.max_speed = 2.5
        self.slow_speed = 0.5
        self.turning_radius = 8.0
        self.coverage_turning_radius = 16.0
        self.time_horizon = 30.0
        self.time_minimum = 5.0
        self.collision_checking_increment = 0.05
        self.start_state_time = 1.0
        self.ribbon_width = 1.5
        self.collision_penalty_factor = 600.0
        self.time_penalty_factor = 1.0
        self.heuristic = H_MAX_DISTANCE
        self.branching_factor = 9
        self.tsp_k = 2
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# ppe_edge (176 bytes)
EDGE_DTYPE = np.dtype(
    [
        ("src", "<f8", (5,)),
        ("src_g", "<f8"),
        ("dst", "<f8", (4,)),
        ("path_qi", "<f8", (3,)),
        ("path_param", "<f8", (3,)),
        ("path_rho", "<f8"),
        ("w_speed", "<f8"),
        ("w_start_time", "<f8"),
        ("w_end_time", "<f8"),
        ("path_type", "<i4"),
        ("has_path", "<i4"),
        ("coverage_allowed", "<i4"),
        ("ribbon_set", "<i4"),
    ],
    align=True,
)

# ppe_edge_result (208 bytes)
RESULT_DTYPE = np.dtype(
    [
        ("true_cost", "<f8"),
        ("collision_penalty", "<f8"),
        ("approx_cost", "<f8"),
        ("end", "<f8", (5,)),
        ("g", "<f8"),
        ("h", "<f8"),
        ("coverage_completed_time", "<f8"),
        ("path_qi", "<f8", (3,)),
        ("path_param", "<f8", (3,)),
        ("path_rho", "<f8"),
        ("w_speed", "<f8"),
        ("w_start_time", "<f8"),
        ("w_end_time", "<f8"),
        ("ribbons_offset", "<i8"),
        ("path_type", "<i4"),
        ("infeasible", "<i4"),
        ("status", "<i4"),
        ("n_samples", "<i4"),
        ("n_checkpoints", "<i4"),
        ("n_ribbons_after", "<i4"),
        ("ribbons_changed", "<i4"),
        ("reserved", "<i4"),
    ],
    align=True,
)

# ppe_vertex (80 bytes) / ppe_child (160 bytes): frontier expansion
VERTEX_DTYPE = np.dtype(
    [("state", "<f8", (5,)), ("g", "<f8"), ("endpoint", "<f8", (3,)), ("ribbon_set", "<i4"), ("has_endpoint", "<i4")],
    align=True,
)
CHILD_DTYPE = np.dtype(
    [
        ("true_cost", "<f8"),
        ("collision_penalty", "<f8"),
        ("approx_cost", "<f8"),
        ("end", "<f8", (5,)),
        ("g", "<f8"),
        ("h", "<f8"),
        ("coverage_completed_time", "<f8"),
        ("path_param", "<f8", (3,)),
        ("w_end_time", "<f8"),
        ("ribbons_offset", "<i8"),
        ("sample_index", "<i4"),
        ("path_type", "<i4"),
        ("infeasible", "<i4"),
        ("status", "<i4"),
        ("coverage_allowed", "<i4"),
        ("n_ribbons_after", "<i4"),
        ("ribbons_changed", "<i4"),
        ("reserved", "<i4"),
    ],
    align=True,
)
EXPAND_TIE = 1
EXPAND_OVERFLOW = 2
EDGE_SKIPPED = 4

assert EDGE_DTYPE.itemsize == 176, EDGE_DTYPE.itemsize
assert RESULT_DTYPE.itemsize == 208, RESULT_DTYPE.itemsize
assert VERTEX_DTYPE.itemsize == 80, VERTEX_DTYPE.itemsize
assert CHILD_DTYPE.itemsize == 160, CHILD_DTYPE.itemsize


def dptr(a):
    """double* of a C-contiguous float64 array (or None)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def vptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def declare_world_api(lib, prefix, ctx_t=C.c_void_p):
    """Attach argtypes/restype for the world-state + batch entry points that the engine
    (`ppe_`), the C oracle (`oracle_`) and the compiled reference (`ref_`) all share."""
    D = C.POINTER(C.c_double)
    I = C.POINTER(C.c_int32)

    def f(name, argtypes, restype=C.c_int):
        fn = getattr(lib, prefix + name)
        fn.argtypes = argtypes
        fn.restype = restype
        return fn

    f("destroy", [ctx_t], None)
    f("last_error", [ctx_t], C.c_char_p)
    f("set_config", [ctx_t, C.POINTER(PpeConfig)])
    f("set_map_none", [ctx_t])
    f("set_map_bitmap", [ctx_t, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double])
    f("set_obstacles_none", [ctx_t])
    f("set_obstacles_binary", [ctx_t, C.c_int, D, D, D, D, D, D, D])
    f("set_obstacles_gaussian", [ctx_t, C.c_int, D, D, D, D, D, D])
    f("put_ribbon_set", [ctx_t, C.c_int, D, C.c_double, I])
    f("clear_ribbon_sets", [ctx_t])
    f("dubins_batch", [ctx_t, C.c_int64, D, D, D, I, D, D, I])
    f("true_cost_batch", [ctx_t, C.c_int64, C.c_void_p, C.c_void_p])
    f("get_ribbons_after", [ctx_t, C.c_int64, D, C.c_int])
    if hasattr(lib, prefix + "expand_batch"):  # the engine and the C oracle; the compiled reference expands through its planner
        f("clear_samples", [ctx_t])
        f("add_samples", [ctx_t, C.c_int64, D, D, D, C.c_void_p], C.c_int64)
        f("sample_count", [ctx_t], C.c_int64)
        f("expand_stride", [ctx_t])
        f("expand_batch", [ctx_t, C.c_int, C.c_void_p, I, C.c_void_p, I, I])
        f("ribbon_pool", [ctx_t, C.POINTER(C.c_int64)], D)
        f("expand_solve_count", [ctx_t], C.c_int64)
