"""ctypes glue shared by every library that speaks the include/ppe.h batch structs.

`CApiWorld` wraps a context of a library exporting `<prefix>set_config`, `<prefix>set_map_*`,
`<prefix>set_obstacles_*`, `<prefix>put_ribbon_set`, `<prefix>dubins_batch`,
`<prefix>true_cost_batch`, `<prefix>get_ribbons_after`.  The engine (`ppe_`, the product) is
bound in path_planner_b200/engine.py; tests bind the CPU oracle (`oracle_`) and the compiled
reference (`ref_`) through the same class so that the parity tests read identically on both
sides.  This module contains no arithmetic.
"""
import ctypes as C

import numpy as np

from . import abi


class PpeError(RuntimeError):
    pass


class CApiWorld:
    prefix = "ppe_"

    def __init__(self, lib, ctx, prefix):
        self._lib = lib
        self._ctx = ctx
        self.prefix = prefix
        abi.declare_world_api(lib, prefix)
        self._keep = []
        self.config = None

    # -- helpers ---------------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self.prefix + name)

    def _check(self, rc, what):
        if rc < 0:
            msg = self._fn("last_error")(self._ctx)
            raise PpeError("%s%s failed (%d): %s" % (self.prefix, what, rc, (msg or b"").decode()))
        return rc

    def close(self):
        if self._ctx is not None:
            self._fn("destroy")(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- world state -------------------------------------------------------------------------
    def set_config(self, cfg):
        self.config = cfg
        self._check(self._fn("set_config")(self._ctx, C.byref(cfg)), "set_config")

    def set_map_none(self):
        self._check(self._fn("set_map_none")(self._ctx), "set_map_none")

    def set_map_bitmap(self, bits, rows, cols, resolution):
        """bits: uint8 array [rows, stride]; bit (r, c) = bits[r, c // 8] >> (c % 8) & 1; row 0 = y 0."""
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        assert bits.ndim == 2 and bits.shape[0] == rows and bits.shape[1] * 8 >= cols
        self._check(
            self._fn("set_map_bitmap")(self._ctx, abi.vptr(bits), rows, cols, bits.shape[1], float(resolution)),
            "set_map_bitmap",
        )

    def set_obstacles_none(self):
        self._check(self._fn("set_obstacles_none")(self._ctx), "set_obstacles_none")

    def set_obstacles_binary(self, x, y, ang, speed, time, width, length):
        """`ang` is yaw for the engine/oracle and heading for the compiled reference (`ref_`)."""
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, ang, speed, time, width, length)]
        self._check(self._fn("set_obstacles_binary")(self._ctx, len(a[0]), *[abi.dptr(v) for v in a]), "set_obstacles_binary")

    def set_obstacles_gaussian(self, x, y, ang, speed, time, cov=None):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, ang, speed, time)]
        covp = None
        if cov is not None:
            cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(len(a[0]), 4)
            covp = abi.dptr(cov)
        self._check(
            self._fn("set_obstacles_gaussian")(self._ctx, len(a[0]), *[abi.dptr(v) for v in a], covp),
            "set_obstacles_gaussian",
        )

    def put_ribbon_set(self, xyxy, coverage_completed_time=-1.0):
        xyxy = np.ascontiguousarray(xyxy, dtype=np.float64).reshape(-1, 4)
        sid = np.zeros(1, dtype=np.int32)
        self._check(
            self._fn("put_ribbon_set")(self._ctx, xyxy.shape[0], abi.dptr(xyxy), float(coverage_completed_time), abi.iptr(sid)),
            "put_ribbon_set",
        )
        return int(sid[0])

    def clear_ribbon_sets(self):
        self._check(self._fn("clear_ribbon_sets")(self._ctx), "clear_ribbon_sets")

    # -- batches -----------------------------------------------------------------------------
    def dubins_batch(self, q0, q1, rho):
        """DubinsWrapper::set / Edge::computeApproxCost for n (q0, q1, rho) triples.
        Returns (type[n] int32, param[n,3], length[n], err[n] int32)."""
        q0 = np.ascontiguousarray(q0, dtype=np.float64).reshape(-1, 3)
        q1 = np.ascontiguousarray(q1, dtype=np.float64).reshape(-1, 3)
        n = q0.shape[0]
        rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (n,)))
        typ = np.zeros(n, dtype=np.int32)
        par = np.zeros((n, 3), dtype=np.float64)
        length = np.zeros(n, dtype=np.float64)
        err = np.zeros(n, dtype=np.int32)
        self._check(
            self._fn("dubins_batch")(self._ctx, n, abi.dptr(q0), abi.dptr(q1), abi.dptr(rho), abi.iptr(typ),
                                     abi.dptr(par), abi.dptr(length), abi.iptr(err)),
            "dubins_batch",
        )
        return typ, par, length, err

    def true_cost_batch(self, edges):
        """Edge::computeTrueCost for a batch of `abi.EDGE_DTYPE` records -> `abi.RESULT_DTYPE`."""
        edges = np.ascontiguousarray(edges, dtype=abi.EDGE_DTYPE)
        res = np.zeros(edges.shape[0], dtype=abi.RESULT_DTYPE)
        self._check(
            self._fn("true_cost_batch")(self._ctx, edges.shape[0], abi.vptr(edges), abi.vptr(res)), "true_cost_batch"
        )
        return res

    def ribbons_after(self, edge_index, cap=1024):
        buf = np.zeros((cap, 4), dtype=np.float64)
        n = self._check(self._fn("get_ribbons_after")(self._ctx, int(edge_index), abi.dptr(buf), cap), "get_ribbons_after")
        if n > cap:
            return self.ribbons_after(edge_index, cap=n)
        return buf[:n].copy()

    # -- frontier expansion (SamplingBasedPlanner::addSamples / ::expand for many vertices) ---------------
    def clear_samples(self):
        self._check(self._fn("clear_samples")(self._ctx), "clear_samples")

    def add_samples(self, x, y, heading):
        """Appends the states the map does not block to the resident sample set; returns the keep mask."""
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, y, heading)]
        keep = np.zeros(len(a[0]), dtype=np.uint8)
        kept = self._fn("add_samples")(self._ctx, len(a[0]), abi.dptr(a[0]), abi.dptr(a[1]), abi.dptr(a[2]), abi.vptr(keep))
        self._check(int(min(kept, 0)), "add_samples")
        assert kept == int(keep.sum())
        return keep.astype(bool)

    def sample_count(self):
        return int(self._fn("sample_count")(self._ctx))

    def expand_batch(self, vertices):
        """`abi.VERTEX_DTYPE` records -> (n_children[n], children[n, stride] of `abi.CHILD_DTYPE`, flags[n], n_popped[n],
        ribbon pool [m, 4])."""
        vertices = np.ascontiguousarray(vertices, dtype=abi.VERTEX_DTYPE)
        n = vertices.shape[0]
        stride = int(self._fn("expand_stride")(self._ctx))
        nch = np.zeros(n, dtype=np.int32)
        flags = np.zeros(n, dtype=np.int32)
        pops = np.zeros(n, dtype=np.int32)
        children = np.zeros((n, stride), dtype=abi.CHILD_DTYPE)
        self._check(self._fn("expand_batch")(self._ctx, n, abi.vptr(vertices), abi.iptr(nch), abi.vptr(children), abi.iptr(flags),
                                             abi.iptr(pops)), "expand_batch")
        m = C.c_int64()
        p = self._fn("ribbon_pool")(self._ctx, C.byref(m))
        pool = np.ctypeslib.as_array(p, shape=(m.value * 4,)).reshape(-1, 4).copy() if m.value > 0 else np.zeros((0, 4))
        return nch, children, flags, pops, pool
