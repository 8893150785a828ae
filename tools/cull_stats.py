"""Instrumentation read-out (GPU): how many 32-sample chunks the probe pass culls and which walker evaluated the edges."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo")
from path_planner_b200 import EdgeEngine, synth
eng = EdgeEngine(0)
for name, near in [("c1", 0.0), ("c2", 0.0), ("c2", 0.6), ("c3", 0.0), ("c3b", 0.0), ("c4", 0.0), ("c5", 0.0)]:
    world = synth.WORLDS[name]()
    edges = synth.make_edges(world, 20000, seed=5, near_ribbons=near)
    edges["ribbon_set"] = world.upload(eng)
    r = eng.true_cost_batch(edges)
    who = (r["reserved"] >> 24) & 1   # 1: K2t thread walker, 2: K2h heavy thread walker, 0: K2b warp walker
    culled = r["reserved"] & 0xFFFFFF
    ch = np.ceil(r["n_samples"] / 32)
    hv = who != 1
    print("%-4s near %.1f: chunks/edge %.1f culled %.3f  check-points/edge %.2f  walked by K2t %.3f K2h %.3f K2b %.4f  (not K2t: changed %.3f, mean cps %.1f)" % (
        name, near, ch.mean(), culled.sum() / ch.sum(), r["n_checkpoints"].mean(), (who == 1).mean(), (who == 2).mean(), (who == 0).mean(),
        (r["ribbons_changed"][hv] != 0).mean() if hv.any() else 0, r["n_checkpoints"][hv].mean() if hv.any() else 0))
