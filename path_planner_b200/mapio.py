"""GridWorld `.map` ingest: the text format the reference's GridWorldMap reads
(path_planner/src/common/map/GridWorldMap.cpp:10-64) -> the bit-packed occupancy map of ppe_set_map_bitmap.

Format, as the reference parses it: line 1 holds the resolution (metres per cell, read with `>>`, so leading blanks
and trailing text are tolerated); every following line is one row of cells, '#' = blocked, anything else = free; the
number of columns is the length of the SHORTEST row (longer rows are cut); the LAST line of the file is row 0 (y = 0).
A '\\r' left by CRLF files counts as a (free) cell, exactly as std::getline leaves it for the reference.
"""
import re

import numpy as np

_FLOAT_PREFIX = re.compile(r"\s*([+-]?(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?))")


def parse_gridworld_map(text):
    """Returns (bits uint8 [rows, stride], rows, cols, resolution); bit (r, c) = bits[r, c // 8] >> (c % 8) & 1."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()  # std::getline does not produce a line after a trailing newline
    if not lines:
        raise ValueError("empty map file")
    m = _FLOAT_PREFIX.match(lines[0])
    if not m:
        raise ValueError("first line of a GridWorld map must start with the resolution")
    resolution = float(m.group(1))
    rows_txt = lines[1:]
    if not rows_txt:
        raise ValueError("GridWorld map without rows")
    cols = min(len(l) for l in rows_txt)
    rows = len(rows_txt)
    if cols == 0:
        raise ValueError("GridWorld map with an empty row (the reference would index an empty vector)")
    rows_txt = rows_txt[::-1]  # last line = y 0 (GridWorldMap.cpp:25)
    blocked = np.zeros((rows, cols), dtype=bool)
    for r, line in enumerate(rows_txt):
        blocked[r] = np.frombuffer(line[:cols].encode("latin-1"), dtype=np.uint8) == ord("#")
    return pack_bits(blocked), rows, cols, resolution


def load_gridworld_map(path):
    with open(path, "r", newline="", encoding="latin-1") as f:
        return parse_gridworld_map(f.read())


def pack_bits(blocked):
    """bool [rows, cols] -> uint8 [rows, ceil(cols / 8)], bit c % 8 of byte c // 8 (little-endian bit order)."""
    return np.packbits(np.asarray(blocked, dtype=bool), axis=1, bitorder="little")


def save_gridworld_map(path, blocked, resolution):
    """Writes bool [rows, cols] (row 0 = y 0) in the reference's format."""
    blocked = np.asarray(blocked, dtype=bool)
    with open(path, "w", newline="") as f:
        f.write("%.17g\n" % resolution)
        for r in range(blocked.shape[0] - 1, -1, -1):
            f.write("".join("#" if b else "_" for b in blocked[r]) + "\n")


def set_map_from_file(world_api, path):
    """Loads a `.map` file into anything that speaks the C-ABI world calls (EdgeEngine, the oracle wrappers)."""
    bits, rows, cols, resolution = load_gridworld_map(path)
    world_api.set_map_bitmap(bits, rows, cols, resolution)
    return rows, cols, resolution
