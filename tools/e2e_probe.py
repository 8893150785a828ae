"""A/B of the host-buffer pipeline of ppe_true_cost_batch on one GPU: the legacy slicing of the whole kernel sequence
(PPE_LATE_K2B=0) against K2a + K2t per slice / K2b once (PPE_LATE_SLICE = 64 k, 128 k, 256 k), with the pinned-memory copy
rates of the box beside them.  python tools/e2e_probe.py [c5 c2 ...]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from path_planner_b200 import EdgeEngine, abi, synth  # noqa: E402


def copy_rates(nbytes):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        out[name + "_GBs"] = 5 * nbytes / (time.perf_counter() - t0) / 1e9
    return out


def run(workload, n, env, steps=6, near=0.0):
    for k, v in env.items():
        os.environ[k] = v
    try:
        eng = EdgeEngine(0)
    finally:
        for k in env:
            os.environ.pop(k, None)
    world = synth.WORLDS[workload]()
    sid = world.upload(eng)
    edges = synth.make_edges(world, n, seed=5, near_ribbons=near)
    edges["ribbon_set"] = sid
    h_edges = torch.from_numpy(edges.view(np.uint8).reshape(n, -1).copy()).pin_memory()
    h_res = torch.empty((n, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8).pin_memory()

    def step():
        rc = eng._lib.ppe_true_cost_batch(eng._ctx, n, C.c_void_p(h_edges.data_ptr()), C.c_void_p(h_res.data_ptr()))
        assert rc == 0, rc

    step()
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    res = h_res.numpy().view(abi.RESULT_DTYPE).reshape(n).copy()
    res["ribbons_offset"] = 0
    return dt, res


def main():
    workloads = sys.argv[1:] or ["c5", "c2"]
    n = 1 << 20
    print(json.dumps({"copy": copy_rates(200 << 20)}))
    for wl in workloads:
        base = None
        for label, env in (("legacy-256k", {"PPE_LATE_K2B": "0"}), ("late-64k", {"PPE_LATE_SLICE": "65536"}),
                           ("late-128k", {}), ("late-256k", {"PPE_LATE_SLICE": "262144"})):
            dt, res = run(wl, n, env)
            same = None if base is None else bool(res.tobytes() == base.tobytes())
            if base is None:
                base = res
            print(json.dumps({"workload": wl, "pipeline": label, "ms_per_call": dt * 1e3, "edges_per_s": n / dt, "same_bytes_as_legacy": same}))


if __name__ == "__main__":
    main()
