"""Where the warp walker K2b spends its cycles (GPU, development build `make -C path_planner_b200/csrc prof`):
per workload the share of probe passes / lane poses + map / fast and general check-points / obstacle penalties / tail."""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
os.environ["PPE_LIB_PATH"] = os.path.join("/root/repo", "path_planner_b200", "libppe_prof.so")
from path_planner_b200 import EdgeEngine, synth

eng = EdgeEngine(0)
lib = eng._lib
lib.ppe_debug_k2b_profile.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
names = ["edges", "total", "probe", "poses+map", "cp fast", "cp general", "obstacles", "tail", "n fast", "n general"]
for name, near, n in [("c2", 0.0, 1 << 18), ("c3", 0.0, 1 << 18), ("c5", 0.0, 1 << 18), ("c2", 0.6, 1 << 16)]:
    world = synth.WORLDS[name]()
    edges = synth.make_edges(world, n, seed=5, near_ribbons=near)
    edges["ribbon_set"] = world.upload(eng)
    buf = (C.c_ulonglong * 16)()
    lib.ppe_debug_k2b_profile(eng._ctx, buf)
    r = eng.true_cost_batch(edges)
    k = lib.ppe_debug_k2b_profile(eng._ctx, buf)
    v = np.array(list(buf)[:10], dtype=np.float64)
    tot = max(v[1], 1)
    who = (r["reserved"] >> 24) & 1
    print("%s near %.1f: %d edges, %d on the warp walker (%.1f%%), mean %.0f kcycles/edge, max check-points %d" % (
        name, near, n, int(v[0]), 100 * v[0] / n, tot / max(v[0], 1) / 1e3, int(r["n_checkpoints"].max())))
    print("   " + "  ".join("%s %.1f%%" % (names[i], 100 * v[i] / tot) for i in range(2, 8)) +
          "  | check-points fast %d general %d (%.0f / %.0f cycles each)" % (v[8], v[9], v[4] / max(v[8], 1), v[5] / max(v[9], 1)))
    hv = r[who == 0]
    print("   warp-walker edges: mean samples %.0f check-points %.1f changed %.2f infeasible %.2f" % (
        hv["n_samples"].mean(), hv["n_checkpoints"].mean(), (hv["ribbons_changed"] != 0).mean(), hv["infeasible"].mean()))
