"""The two host-buffer pipelines of ppe_true_cost_batch must return the same bytes for a batch size that the adaptive slice
size pipelines (n / 8 per slice): python tools/pipeline_equivalence.py [n] [workload]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from path_planner_b200 import EdgeEngine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
wl = sys.argv[2] if len(sys.argv) > 2 else "c5"
world = synth.WORLDS[wl]()
edges = synth.make_edges(world, n, seed=77, near_ribbons=0.2)
out = []
for env in ({}, {"PPE_LATE_K2B": "0"}):
    os.environ.update(env)
    try:
        eng = EdgeEngine(0)
    finally:
        for k in env:
            os.environ.pop(k, None)
    edges["ribbon_set"] = world.upload(eng)
    l0 = eng.launch_count()
    r = eng.true_cost_batch(edges)
    launches = eng.launch_count() - l0
    r["ribbons_offset"] = 0
    out.append((r.tobytes(), launches, int(((r["reserved"] >> 24) & 1).sum())))
print("n %d %s: launches late-K2b %d, sliced-whole %d; thread-walked %d; same bytes: %s" % (n, wl, out[0][1], out[1][1], out[0][2], out[0][0] == out[1][0]))
sys.exit(0 if out[0][0] == out[1][0] else 1)
