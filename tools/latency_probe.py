"""Latency of the calls one BatchedAStarPlanner::expand makes (GPU): 40-edge true-cost batches and 256-solve Dubins
batches through the host-buffer C ABI, on the C1 / C2 worlds, for edges far from and along the survey lines."""
import sys
import time
import numpy as np
sys.path.insert(0, "/root/repo")
from path_planner_b200 import EdgeEngine, synth

eng = EdgeEngine(0)
for name, near in [("c1", 0.0), ("c1", 1.0), ("c2", 0.0), ("c2", 1.0), ("c3", 0.3)]:
    world = synth.WORLDS[name]()
    edges = synth.make_edges(world, 40 * 50, seed=9, near_ribbons=near)
    edges["ribbon_set"] = world.upload(eng)
    eng.true_cost_batch(edges[:40])
    ts = []
    cps = []
    for k in range(50):
        b = edges[40 * k:40 * k + 40]
        t0 = time.perf_counter()
        r = eng.true_cost_batch(b)
        ts.append(time.perf_counter() - t0)
        cps.append(r["n_checkpoints"].max())
    ts = np.array(ts) * 1e3
    print("%-3s near %.1f: 40-edge batch median %.3f ms  p90 %.3f ms  max %.3f ms   (max check-points in a batch: median %d)" % (
        name, near, np.median(ts), np.percentile(ts, 90), ts.max(), int(np.median(cps))))
q0 = np.random.default_rng(0).uniform(0, 100, (256, 3))
q1 = np.random.default_rng(1).uniform(0, 100, (256, 3))
eng.dubins_batch(q0, q1, 8.0)
t0 = time.perf_counter()
for _ in range(100):
    eng.dubins_batch(q0, q1, 8.0)
print("256-solve Dubins batch: %.3f ms" % ((time.perf_counter() - t0) * 10))
rib = np.array([[0.0, 10.0, 0.0, 30.0]])
t0 = time.perf_counter()
for _ in range(100):
    eng.put_ribbon_set(rib, -1.0)
print("put_ribbon_set: %.3f ms" % ((time.perf_counter() - t0) * 10))
