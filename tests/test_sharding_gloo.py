"""The N > 1 host logic on CPU: world_size-2 `gloo` process group, each rank evaluates its contiguous
shard of one edge batch (here with the CPU oracle standing in for its GPU) and the ranks exchange ONLY
the 16-byte best-f record; the gathered winner must equal the single-process answer, including the
smaller-index tie-break.  Also the round-robin scenario assignment of the 64-scenario config."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from path_planner_b200 import sharding, synth
from tests import common


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_best(res):
    ok = (res["infeasible"] == 0) & (res["status"] == 0)
    if not ok.any():
        return np.inf, -1
    f = res["g"] + res["h"]
    f = np.where(ok, f, np.inf)
    i = int(np.argmin(f))  # first minimum = smaller index on ties
    return float(f[i]), i


def _worker(rank, world_size, port, n, out_q, force_tie):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        world = synth.world_c2()
        edges = synth.make_edges(world, n, seed=41, near_ribbons=0.3)
        if force_tie:  # the same edge at both ends of the batch -> equal f on two ranks
            edges[-1] = edges[0]
        o = common.load_oracle("glibc")
        edges["ribbon_set"] = world.upload(o)
        lo, hi = sharding.shard_range(n, rank, world_size)
        res = o.true_cost_batch(edges[lo:hi])
        f, i = _local_best(res)
        rec = torch.from_numpy(sharding.pack_best(f, i, lo))
        best = sharding.gather_best(rec)
        out_q.put((rank, lo, hi, best))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("force_tie", [False, True])
def test_two_rank_gather_equals_single_process(force_tie):
    n, world_size = 601, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world_size, port, n, q, force_tie)) for r in range(world_size)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    world = synth.world_c2()
    edges = synth.make_edges(world, n, seed=41, near_ribbons=0.3)
    if force_tie:
        edges[-1] = edges[0]
    o = common.load_oracle("glibc")
    edges["ribbon_set"] = world.upload(o)
    f, i = _local_best(o.true_cost_batch(edges))
    covered = sorted((lo, hi) for _, lo, hi, _ in got)
    assert covered == [(0, 301), (301, 601)]
    for rank, lo, hi, best in got:
        assert best[0] == f and best[1] == i, (rank, best, f, i)
        assert best[2] == (0 if i < 301 else 1)


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 1 << 20, (1 << 20) + 3):
        for ws in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(ws - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_scenarios_round_robin():
    for ws in (1, 2, 4, 8):
        seen = sorted(s for k in range(ws) for s in sharding.scenario_assignment(64, k, ws))
        assert seen == list(range(64))
        assert all(len(sharding.scenario_assignment(64, k, ws)) == 64 // ws for k in range(ws))
