"""Python host side of the B200 edge engine: a thin mirror of include/ppe.h over ctypes.

`EdgeEngine` binds path_planner_b200/libppe.so (hand-written sm_100a CUDA behind a C ABI).  It is
the batched drop-in for Edge::computeApproxCost / Edge::computeTrueCost
(path_planner/src/planner/search/Edge.cpp:11-20,68-206).  There is no CPU fallback: a missing
library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from ._capi import CApiWorld, PpeError

_HERE = os.path.dirname(os.path.abspath(__file__))
# PPE_LIB_PATH: build-variant experiments only (profiles/); the shipped library is libppe.so
LIB_PATH = os.environ.get("PPE_LIB_PATH") or os.path.join(_HERE, "libppe.so")

_lib = None


def load_library():
    """dlopen libppe.so and declare the ABI.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PpeError(
            "CUDA extension %s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    D = C.POINTER(C.c_double)
    I = C.POINTER(C.c_int32)
    lib.ppe_abi_version.restype = C.c_int
    for name in ("ppe_abi_sizeof_config", "ppe_abi_sizeof_edge", "ppe_abi_sizeof_edge_result"):
        getattr(lib, name).restype = C.c_int
    lib.ppe_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.ppe_create.restype = C.c_int
    abi.declare_world_api(lib, "ppe_")
    lib.ppe_best.argtypes = [C.c_void_p, D, C.POINTER(C.c_int64)]
    lib.ppe_best.restype = C.c_int
    lib.ppe_dubins_batch_device.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_void_p]
    lib.ppe_dubins_batch_device.restype = C.c_int
    lib.ppe_true_cost_batch_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ppe_true_cost_batch_device.restype = C.c_int
    lib.ppe_best_device.argtypes = [C.c_void_p, D, C.POINTER(C.c_int64), C.c_void_p]
    lib.ppe_best_device.restype = C.c_int
    lib.ppe_best_copy_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.ppe_best_copy_device.restype = C.c_int
    lib.ppe_launch_count.argtypes = [C.c_void_p]
    lib.ppe_launch_count.restype = C.c_int64
    lib.ppe_map_generation.argtypes = [C.c_void_p]
    lib.ppe_map_generation.restype = C.c_uint64
    lib.ppe_measure_fp64_peak.argtypes = [C.c_void_p, D, C.c_void_p]
    lib.ppe_measure_fp64_peak.restype = C.c_int
    if lib.ppe_abi_version() != abi.PPE_ABI_VERSION:
        raise PpeError("libppe.so ABI version %d != %d" % (lib.ppe_abi_version(), abi.PPE_ABI_VERSION))
    if lib.ppe_abi_sizeof_edge() != abi.EDGE_DTYPE.itemsize or lib.ppe_abi_sizeof_edge_result() != abi.RESULT_DTYPE.itemsize \
            or lib.ppe_abi_sizeof_config() != C.sizeof(abi.PpeConfig):
        raise PpeError("libppe.so struct layout differs from path_planner_b200/abi.py")
    _lib = lib
    return lib


class EdgeEngine(CApiWorld):
    """One engine context = one GPU (one process per GPU).  Methods mirror include/ppe.h."""

    def __init__(self, device=0):
        lib = load_library()
        ctx = C.c_void_p()
        rc = lib.ppe_create(int(device), C.byref(ctx))
        if rc != abi.PPE_OK:
            raise PpeError(
                "ppe_create(device=%d) failed (%d): no usable sm_100 CUDA device; the engine has no CPU path" % (device, rc)
            )
        super().__init__(lib, ctx, "ppe_")
        self.device = device

    # ---- K3 --------------------------------------------------------------------------------
    def best(self):
        f = C.c_double()
        idx = C.c_int64()
        self._check(self._lib.ppe_best(self._ctx, C.byref(f), C.byref(idx)), "best")
        return f.value, idx.value

    # ---- device-resident variants (raw device pointers + cudaStream_t as integers) -----------
    def dubins_batch_device(self, n, q0, q1, rho, typ, param, length, err, stream=0):
        self._check(
            self._lib.ppe_dubins_batch_device(self._ctx, int(n), q0, q1, rho, typ, param, length, err, C.c_void_p(stream)),
            "dubins_batch_device",
        )

    def true_cost_batch_device(self, n, d_edges, d_results, stream=0):
        self._check(
            self._lib.ppe_true_cost_batch_device(self._ctx, int(n), C.c_void_p(d_edges), C.c_void_p(d_results), C.c_void_p(stream)),
            "true_cost_batch_device",
        )

    def best_device(self, stream=0):
        f = C.c_double()
        idx = C.c_int64()
        self._check(self._lib.ppe_best_device(self._ctx, C.byref(f), C.byref(idx), C.c_void_p(stream)), "best_device")
        return f.value, idx.value

    def best_copy_device(self, d_dst16, index_base=0, stream=0):
        """{f64 f, i64 edge_index + index_base} of the last device batch -> 16 bytes at d_dst16 (no sync)."""
        self._check(self._lib.ppe_best_copy_device(self._ctx, C.c_void_p(d_dst16), int(index_base), C.c_void_p(stream)), "best_copy_device")

    def map_generation(self):
        return int(self._lib.ppe_map_generation(self._ctx))

    # ---- instrumentation ---------------------------------------------------------------------
    def launch_count(self):
        return int(self._lib.ppe_launch_count(self._ctx))

    def measure_fp64_peak(self, stream=0):
        t = C.c_double()
        self._check(self._lib.ppe_measure_fp64_peak(self._ctx, C.byref(t), C.c_void_p(stream)), "measure_fp64_peak")
        return t.value
