"""BASELINE-size batches (2^20 edges, the bench configuration) on the GPU: the oracle cannot evaluate a million edges
in test time, so the full batch is checked through size-independent properties and a random subset against the
oracle:
  * the pipelined host-buffer call (slices over two kernel lanes and two copy streams) returns bit for bit what
    the single-launch device-resident call returns, including the ribbons-after handles' contents;
  * K3's best record equals the minimum recomputed on the host over all result records;
  * a shuffled batch returns the same per-edge records (no edge's result depends on its neighbours or on which
    walker -- thread, warp -- picked it up);
  * 3 000 randomly chosen edges of the batch agree with the oracle (discrete fields exactly, continuous to 1e-9).
"""
import numpy as np
import pytest
import torch

from path_planner_b200 import EdgeEngine, abi, synth
from tests import common

pytestmark = pytest.mark.gpu
N = 1 << 20


@pytest.fixture(scope="module")
def engine():
    return EdgeEngine(0)


def _device_batch(engine, edges):
    n = len(edges)
    dev = torch.device("cuda", 0)
    d_edges = torch.from_numpy(edges.view(np.uint8).reshape(n, abi.EDGE_DTYPE.itemsize)).to(dev)
    d_res = torch.zeros((n, abi.RESULT_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    engine.true_cost_batch_device(n, d_edges.data_ptr(), d_res.data_ptr(), s)
    f, idx = engine.best_device(s)
    torch.cuda.synchronize()
    return d_res.cpu().numpy().view(abi.RESULT_DTYPE).reshape(n), f, idx


FIELDS = [f for f in abi.RESULT_DTYPE.names if f not in ("ribbons_offset", "reserved")]


@pytest.mark.parametrize("name,near", [("c2", 0.05), ("c5", 0.02)])
def test_full_size_batch(engine, name, near):
    world = synth.WORLDS[name]()
    edges = synth.make_edges(world, N, seed=77, near_ribbons=near)
    edges["ribbon_set"] = world.upload(engine)

    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(N, 3000, replace=False))
    host = engine.true_cost_batch(edges)                 # pipelined: 4 slices of 262 144 edges
    f_host, i_host = engine.best()
    changed_in_pick = [int(i) for i in pick if host["ribbons_changed"][i]][:40]
    ribbons_host = {i: engine.ribbons_after(i) for i in changed_in_pick}
    dev, f_dev, i_dev = _device_batch(engine, edges)     # one launch group over the whole batch
    for fld in FIELDS:
        a, b = host[fld], dev[fld]
        assert np.array_equal(a, b) or np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), fld
    assert (f_host, i_host) == (f_dev, i_dev)
    assert (host["status"] == 0).all()

    ok = (host["infeasible"] == 0) & (host["status"] == 0) & (host["h"] >= 0)
    fo = np.where(ok, host["g"] + host["h"], np.inf)
    assert f_host == fo.min() and i_host == int(np.argmin(fo))

    # structural invariants of Edge::computeTrueCost that hold for every edge
    assert (host["n_samples"] <= 1502).all() and (host["n_checkpoints"] <= host["n_samples"] + 1).all()
    assert (host["end"][:, 4] <= world.cfg.start_state_time + world.cfg.time_horizon + 1e-9).all()
    assert np.array_equal(host["g"], edges["src_g"] + host["true_cost"])
    assert (host["true_cost"] >= host["collision_penalty"]).all()

    # permutation invariance
    perm = rng.permutation(N)
    shuffled = engine.true_cost_batch(edges[perm])
    for fld in FIELDS:
        a, b = shuffled[fld], host[fld][perm]
        assert np.array_equal(a, b) or np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)]), fld

    # a random subset against the oracle
    oracle = common.load_oracle("cr")
    world.upload(oracle)
    want = oracle.true_cost_batch(edges[pick])
    bad = common.diff_results(host[pick], want)
    assert not bad, common.describe(bad, host[pick], want)
    for j, i in enumerate(pick):
        if int(i) in ribbons_host:
            assert np.allclose(ribbons_host[int(i)], oracle.ribbons_after(j), rtol=common.RTOL, atol=common.ATOL)
