"""One-off stress run (GPU): large seeded batches of every world against the CPU oracle (oracle-B, all host threads),
beyond the sizes the test-suite uses.  Prints one line per batch; exits non-zero on any mismatch."""
import sys
import time
import numpy as np
sys.path.insert(0, "/root/repo")
from path_planner_b200 import EdgeEngine, synth
from tests import common

eng = EdgeEngine(0)
ora = common.load_oracle("cr")
bad_total = 0
for name, near, n, seed in [("c1", 0.3, 60000, 101), ("c2", 0.0, 150000, 102), ("c2", 0.5, 40000, 103), ("c3", 0.1, 40000, 104),
                            ("c3b", 0.1, 60000, 105), ("c4", 0.2, 60000, 106), ("c5", 0.1, 30000, 107)]:
    world = synth.WORLDS[name]()
    edges = synth.make_edges(world, n, seed=seed, near_ribbons=near)
    sid = world.upload(ora)
    assert world.upload(eng) == sid
    edges["ribbon_set"] = sid
    t0 = time.time()
    want = common.true_cost_mt(ora, edges, 0)
    t1 = time.time()
    got = eng.true_cost_batch(edges)
    t2 = time.time()
    bad = common.diff_results(got, want)
    nbad = len(set().union(*[set(v.tolist()) for v in bad.values()])) if bad else 0
    bad_total += nbad
    who = (got["reserved"] >> 24) & 1
    print("%-4s near %.1f n %6d: mismatching edges %d %s | oracle %.1f s, engine %.3f s | K2t %.3f, infeasible %.3f, changed %.3f" % (
        name, near, n, nbad, sorted(bad.keys()) if bad else "", t1 - t0, t2 - t1, who.mean(), got["infeasible"].mean(), got["ribbons_changed"].mean()), flush=True)
    if bad:
        print(common.describe(bad, got, want))
sys.exit(1 if bad_total else 0)
