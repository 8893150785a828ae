"""`.map` ingest: path_planner_b200/mapio.py must read a GridWorld map file exactly as
the reference's GridWorldMap does (GridWorldMap.cpp:10-64) -- ragged rows cut to the shortest, last line = y 0,
'#' blocked -- checked cell by cell and on probe points against the reference's own parser and isBlocked."""
import ctypes as C
import os

import numpy as np
import pytest

from path_planner_b200 import abi, mapio, synth
from tests import common


def _write(tmp_path, text, name="m.map"):
    p = os.path.join(str(tmp_path), name)
    with open(p, "w", newline="") as f:
        f.write(text)
    return p


RAGGED = "0.5  # resolution, metres\n" + "\n".join(["__##____#_", "_#________xx", "##########", "____#_____", "_________#ab", "#_#_#_#_#_"]) + "\n"


def test_parser_follows_the_reference_format(tmp_path):
    bits, rows, cols, res = mapio.load_gridworld_map(_write(tmp_path, RAGGED))
    assert (rows, cols, res) == (6, 10, 0.5)
    cell = lambda r, c: (bits[r, c // 8] >> (c % 8)) & 1
    assert [cell(0, c) for c in range(10)] == [1, 0, 1, 0, 1, 0, 1, 0, 1, 0]   # last line of the file = row 0
    assert [cell(5, c) for c in range(10)] == [0, 0, 1, 1, 0, 0, 0, 0, 1, 0]   # first row line = top row
    assert cell(1, 9) == 1 and cell(3, 0) == 1 and cell(4, 1) == 1


def test_round_trip_of_a_synthetic_world(tmp_path):
    world = synth.world_c2()
    blocked = np.unpackbits(world.map_bits, axis=1, bitorder="little")[:, :world.cols].astype(bool)
    p = os.path.join(str(tmp_path), "c2.map")
    mapio.save_gridworld_map(p, blocked, world.resolution)
    bits, rows, cols, res = mapio.load_gridworld_map(p)
    assert (rows, cols, res) == (world.rows, world.cols, world.resolution)
    assert np.array_equal(bits, world.map_bits[:, :bits.shape[1]])


@pytest.mark.skipif(not common.have_ref(), reason="oracle/_ref/libref_planner.so not built (needs /root/reference)")
@pytest.mark.parametrize("text", [RAGGED, "1\n" + "\n".join("".join("#" if (3 * r + 5 * c) % 7 == 0 else "_" for c in range(37)) for r in range(23)) + "\n",
                                  "2.5\n__#__\r\n_#___\r\n"], ids=["ragged", "37x23", "crlf"])
def test_blocked_lookups_agree_with_the_reference_parser(tmp_path, text):
    path = _write(tmp_path, text)
    ref = common.load_ref()
    ref.lib.ref_set_map_file.argtypes = [C.c_void_p, C.c_char_p]
    ref.lib.ref_is_blocked.argtypes = [C.c_void_p, C.c_double, C.c_double]
    ref.lib.ref_map_resolution.argtypes = [C.c_void_p]
    ref.lib.ref_map_resolution.restype = C.c_double
    assert ref.lib.ref_set_map_file(ref.ctx, path.encode()) == 0
    ora = common.load_oracle("glibc")
    ora.lib.oracle_is_blocked.argtypes = [C.c_void_p, C.c_double, C.c_double]
    ora.set_config(abi.PpeConfig())
    rows, cols, res = mapio.set_map_from_file(ora, path)
    assert ref.lib.ref_map_resolution(ref.ctx) == res
    rng = np.random.default_rng(3)
    xs = np.concatenate([rng.uniform(-2 * res, (cols + 2) * res, 4000), np.arange(-1, cols + 2) * res, (np.arange(cols) + 0.5) * res])
    ys = np.concatenate([rng.uniform(-2 * res, (rows + 2) * res, 4000), np.arange(-1, rows + 2) * res, (np.arange(rows) + 0.5) * res])
    for x in xs[::7]:
        for y in ys[::11]:
            assert ora.lib.oracle_is_blocked(ora.ctx, x, y) == ref.lib.ref_is_blocked(ref.ctx, x, y), (x, y)
    for x, y in zip(xs, rng.permutation(ys)[:len(xs)] if len(ys) >= len(xs) else np.resize(ys, len(xs))):
        assert ora.lib.oracle_is_blocked(ora.ctx, x, y) == ref.lib.ref_is_blocked(ref.ctx, x, y), (x, y)


@pytest.mark.gpu
def test_engine_evaluates_edges_on_a_map_loaded_from_file(tmp_path):
    """The engine on a map that went through the file format == the engine on the bitmap it came from."""
    from path_planner_b200 import EdgeEngine
    world = synth.world_c2()
    blocked = np.unpackbits(world.map_bits, axis=1, bitorder="little")[:, :world.cols].astype(bool)
    p = os.path.join(str(tmp_path), "c2.map")
    mapio.save_gridworld_map(p, blocked, world.resolution)
    eng = EdgeEngine(0)
    edges = synth.make_edges(world, 4000, seed=21)
    edges["ribbon_set"] = world.upload(eng)
    want = eng.true_cost_batch(edges)
    mapio.set_map_from_file(eng, p)
    got = eng.true_cost_batch(edges)
    for f in ("infeasible", "n_samples", "true_cost", "g", "h", "status"):
        assert np.array_equal(got[f], want[f]), f
    assert want["infeasible"].sum() > 50
