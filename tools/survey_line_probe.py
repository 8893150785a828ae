"""Latency of ONE edge that runs along a survey line for its whole 30 s (GPU): the longest strictly sequential item an
edge batch can hold -- a check-point on every 5 cm sample (Edge.cpp:153-172).  C2-style (10 ribbons) and C5-style (100
ribbons) parent sets, coverage-allowed and not, single-edge batches through ppe_true_cost_batch, timed on the host
(median of 20) and checked against the oracle.  VERDICT r1 target: <= 0.5 ms."""
import sys
import time
import numpy as np
sys.path.insert(0, "/root/repo")
from path_planner_b200 import EdgeEngine, abi, synth
from tests import common

eng = EdgeEngine(0)
ora = common.load_oracle("cr")
for name in ("c2", "c5"):
    world = synth.WORLDS[name]()
    sid = world.upload(eng)
    world.upload(ora)
    rb = world.ribbons[3]
    for cov, rev in ((1, 0), (0, 0), (1, 1)):
        e = np.zeros(1, dtype=abi.EDGE_DTYPE)
        # on the ribbon's line, heading along it for 80 m: from its start towards its end, or (rev) the other way round
        if not rev:
            e["src"][0] = [rb[0], rb[1] + 20.0, 0.0, 2.5, 1.0]
            e["dst"][0] = [rb[0], rb[1] + 100.0, 0.0, 2.5]
        else:
            e["src"][0] = [rb[0], rb[3] - 20.0, np.pi, 2.5, 1.0]
            e["dst"][0] = [rb[0], rb[3] - 100.0, np.pi, 2.5]
        e["coverage_allowed"] = cov
        e["ribbon_set"] = sid
        want = ora.true_cost_batch(e)
        got = eng.true_cost_batch(e)
        bad = common.diff_results(got, want)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            eng.true_cost_batch(e)
            ts.append(time.perf_counter() - t0)
        # a trivial edge for the fixed cost of a one-edge batch
        f = e.copy()
        f["src"][0] = [world.start[0], world.start[1], 0.0, 2.5, 1.0]
        f["dst"][0] = [world.start[0] + 1.0, world.start[1] + 5.0, 0.0, 2.5]
        tf = []
        for _ in range(20):
            t0 = time.perf_counter()
            eng.true_cost_batch(f)
            tf.append(time.perf_counter() - t0)
        print("%s cov=%d rev=%d: survey-line edge %.3f ms (fixed cost of a 1-edge batch %.3f ms) check-points %d samples %d changed %d parity %s" % (
            name, cov, rev, np.median(ts) * 1e3, np.median(tf) * 1e3, int(got["n_checkpoints"][0]), int(got["n_samples"][0]),
            int(got["ribbons_changed"][0]), "ok" if not bad else bad))
