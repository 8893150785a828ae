#!/usr/bin/env python3
"""Generates the committed golden vectors in tests/golden/*.npz from the reference itself.

Run in the build container (needs /root/reference): the reference's own Edge.cpp / Vertex.cpp /
Ribbon*.cpp / *ObstaclesManager.cpp / GridWorldMap.cpp, compiled in place into
oracle/_ref/libref_planner.so (oracle/Makefile), evaluates seeded synthetic edge batches; inputs,
outputs and ribbons-after are stored so that boxes without /root/reference (the GPU box) can still
pin the oracle and the engine against real reference output.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from path_planner_b200 import synth  # noqa: E402
from tests import common  # noqa: E402

CASES = [("c1", 0.5, 256, 101), ("c2", 0.5, 384, 102), ("c3", 0.4, 192, 103), ("c3b", 0.4, 192, 104), ("c4", 0.5, 256, 105)]


def main():
    ref = common.load_ref()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, near, n, seed in CASES:
        world = synth.WORLDS[name]()
        sid = world.upload_ref(ref)
        edges = synth.make_edges(world, n, seed=seed, near_ribbons=near)
        edges["ribbon_set"] = sid
        res = ref.true_cost_batch(edges)
        order = common.ref_obstacle_order(ref) if world.obstacle_kind != "none" else np.zeros((0, 9))
        lists = [ref.ribbons_after(i) for i in range(n)]
        counts = np.array([len(a) for a in lists], dtype=np.int32)
        flat = np.concatenate(lists) if counts.sum() else np.zeros((0, 4))
        # the reference does not expose its loop counters
        res["n_samples"] = -1
        res["n_checkpoints"] = -1
        np.savez_compressed(os.path.join(out_dir, "edges_%s.npz" % name), edges=edges, results=res,
                            obstacle_order=order, ribbons_flat=flat, ribbons_count=counts,
                            seed=seed, near=near)
        print(name, n, "edges; infeasible", int(res["infeasible"].sum()), "changed", int(res["ribbons_changed"].sum()),
              "penalised", int((res["collision_penalty"] > 0).sum()))


if __name__ == "__main__":
    main()
