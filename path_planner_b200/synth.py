"""Synthetic worlds and edge batches of the shapes BASELINE.json names.

Host-side data generation only (numpy, seeded); used by bench.py, the harness and the tests.

  C1  single ribbon, empty map, no dynamic obstacles (mirrors RHRSAStarTest1Ribbons,
      path_planner/test/planner/test_planner.cpp:1276-1286)
  C2  1000 x 1000 cells @ 1 m, 40 blocked rectangles, 10 ribbons
  C3  C2 + 50 dynamic obstacles (Gaussian default covariance, or Binary 10 m x 30 m)
  C4  4096 x 4096 cells @ 1 m, 600 rectangles, 100 ribbons
  C5  1M-edge sweep on the C4 world + 50 Gaussian obstacles
"""
import math

import numpy as np

from . import abi


class World:
    """Everything an edge batch is evaluated against (the read-only state replicated per GPU)."""

    def __init__(self, name, cfg, start):
        self.name = name
        self.cfg = cfg
        self.start = np.asarray(start, dtype=np.float64)  # x, y, heading, speed, time
        self.map_bits = None  # uint8 [rows, stride] bit-packed, row 0 = y 0
        self.rows = self.cols = 0
        self.resolution = 1.0
        self.ribbons = np.zeros((0, 4))
        self.obstacle_kind = "none"  # none | binary | gaussian
        self.obstacles = None  # dict of arrays (heading convention, as the managers' update() takes)

    # ---- map helpers -----------------------------------------------------------------------
    def set_grid(self, blocked, resolution=1.0):
        """blocked: bool [rows, cols], row 0 = y 0."""
        blocked = np.asarray(blocked, dtype=bool)
        self.rows, self.cols = blocked.shape
        self.resolution = float(resolution)
        self.map_bits = np.packbits(blocked, axis=1, bitorder="little")
        self._blocked = blocked

    def blocked_cells(self):
        return self._blocked

    def is_blocked(self, x, y):
        """Vectorised GridWorldMap::isBlocked (GridWorldMap.cpp:84-93) for generator use."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if self.map_bits is None:
            return np.zeros(x.shape, dtype=bool)
        cx = x / self.resolution
        cy = y / self.resolution
        oob = (x < 0) | (cx >= self.cols) | (y < 0) | (cy >= self.rows)
        ix = np.clip(cx, 0, self.cols - 1).astype(np.int64)
        iy = np.clip(cy, 0, self.rows - 1).astype(np.int64)
        return oob | self._blocked[iy, ix]

    # ---- upload into any library speaking the ppe.h world API --------------------------------
    def upload(self, w, obstacle_order=None):
        """w: CApiWorld-like.  `obstacle_order`: n x 9 array from the compiled reference
        (container iteration order, stored yaws); default = generation order with
        yaw = pi/2 - heading as the managers' constructors compute it."""
        w.set_config(self.cfg)
        if self.map_bits is None:
            w.set_map_none()
        else:
            w.set_map_bitmap(self.map_bits, self.rows, self.cols, self.resolution)
        if self.obstacle_kind == "none":
            w.set_obstacles_none()
        else:
            o = obstacle_order if obstacle_order is not None else self.obstacles_yaw_order()
            if self.obstacle_kind == "binary":
                w.set_obstacles_binary(o[:, 0], o[:, 1], o[:, 2], o[:, 3], o[:, 4], o[:, 5], o[:, 6])
            else:
                w.set_obstacles_gaussian(o[:, 0], o[:, 1], o[:, 2], o[:, 3], o[:, 4], o[:, 5:9])
        w.clear_ribbon_sets()
        return w.put_ribbon_set(self.ribbons, -1.0)

    def upload_ref(self, w):
        """Same for the compiled reference (`ref_` prefix): obstacles go in through the managers'
        update(mmsi, x, y, heading, ...) API."""
        w.set_config(self.cfg)
        if self.map_bits is None:
            w.set_map_none()
        else:
            w.set_map_bitmap(self.map_bits, self.rows, self.cols, self.resolution)
        o = self.obstacles
        if self.obstacle_kind == "none":
            w.set_obstacles_none()
        elif self.obstacle_kind == "binary":
            w.set_obstacles_binary(o["x"], o["y"], o["heading"], o["speed"], o["time"], o["width"], o["length"])
        else:
            w.set_obstacles_gaussian(o["x"], o["y"], o["heading"], o["speed"], o["time"], o.get("cov"))
        w.clear_ribbon_sets()
        return w.put_ribbon_set(self.ribbons, -1.0)

    def obstacles_yaw_order(self):
        """n x 9: X Y Yaw Speed Time + (Width Length 0 0 | cov) in generation order."""
        o = self.obstacles
        n = len(o["x"])
        out = np.zeros((n, 9))
        out[:, 0] = o["x"]
        out[:, 1] = o["y"]
        out[:, 2] = math.pi / 2 - np.asarray(o["heading"])  # Obstacle ctor: Yaw(M_PI_2 - heading)
        out[:, 3] = o["speed"]
        out[:, 4] = o["time"]
        if self.obstacle_kind == "binary":
            out[:, 5] = o["width"]
            out[:, 6] = o["length"]
        else:
            cov = o.get("cov")
            out[:, 5:9] = np.asarray(cov).reshape(n, 4) if cov is not None else np.array([30.0, 10.0, 10.0, 30.0])
        return out


def _rects(world_size, n, lo, hi, rng, keep_out_pts, keep_out_r, ribbons, ribbon_margin):
    """n axis-aligned blocked rectangles with sides U[lo,hi], rejected near keep-out points / ribbons."""
    blocked = np.zeros((world_size, world_size), dtype=bool)
    placed = 0
    guard = 0
    while placed < n and guard < 100 * n:
        guard += 1
        w, h = rng.uniform(lo, hi, 2)
        x0 = rng.uniform(0, world_size - w)
        y0 = rng.uniform(0, world_size - h)
        x1, y1 = x0 + w, y0 + h
        bad = False
        for (px, py) in keep_out_pts:
            dx = max(x0 - px, 0, px - x1)
            dy = max(y0 - py, 0, py - y1)
            if math.hypot(dx, dy) < keep_out_r:
                bad = True
                break
        if not bad:
            for (ax, ay, bx, by) in ribbons:
                # ribbon bounding box inflated by the margin
                rx0, rx1 = min(ax, bx) - ribbon_margin, max(ax, bx) + ribbon_margin
                ry0, ry1 = min(ay, by) - ribbon_margin, max(ay, by) + ribbon_margin
                if x0 < rx1 and x1 > rx0 and y0 < ry1 and y1 > ry0:
                    bad = True
                    break
        if bad:
            continue
        blocked[int(y0):int(math.ceil(y1)), int(x0):int(math.ceil(x1))] = True
        placed += 1
    return blocked


def world_c1():
    cfg = abi.PpeConfig(ribbon_width=1.5, start_state_time=1.0)
    w = World("C1", cfg, [0, 0, 0, 2.5, 1])
    w.ribbons = np.array([[0.0, 10.0, 0.0, 30.0]])
    return w


def world_c2(size=1000, n_rects=40):
    cfg = abi.PpeConfig(ribbon_width=2.0, start_state_time=1.0)
    s = size / 1000.0
    start = [380 * s, 380 * s, 0, 2.5, 1]
    w = World("C2", cfg, start)
    w.ribbons = np.array([[(400 + 20 * i) * s, 400 * s, (400 + 20 * i) * s, 600 * s] for i in range(10)], dtype=np.float64)
    rng = np.random.default_rng(2)
    blocked = _rects(size, n_rects, 10 * s, 60 * s, rng, [(start[0], start[1])], 50 * s, w.ribbons, 5.0)
    w.set_grid(blocked, 1.0)
    return w


def _add_obstacles(w, kind, n=50, lo=300.0, hi=700.0, seed=3):
    rng = np.random.default_rng(seed)
    o = {
        "x": rng.uniform(lo, hi, n),
        "y": rng.uniform(lo, hi, n),
        "heading": rng.uniform(0, 2 * math.pi, n),
        "speed": rng.uniform(0, 3, n),
        "time": np.full(n, 1.0),
    }
    if kind == "binary":
        o["width"] = np.full(n, 10.0)   # path_planner_node.cpp:163-164 defaults
        o["length"] = np.full(n, 30.0)
    w.obstacle_kind = kind
    w.obstacles = o
    return w


def world_c3(kind="gaussian", size=1000):
    w = world_c2(size=size)
    w.name = "C3-" + kind
    s = size / 1000.0
    return _add_obstacles(w, kind, 50, 300 * s, 700 * s, 3)


def world_c4(size=4096, n_rects=600):
    cfg = abi.PpeConfig(ribbon_width=2.0, start_state_time=1.0)
    s = size / 4096.0
    start = [2048 * s, 2048 * s, 0, 2.5, 1]
    w = World("C4", cfg, start)
    ribbons = []
    rng = np.random.default_rng(4)
    # 10 blocks x 10 parallel 200 m lines spaced 20 m
    for b in range(10):
        bx = (300 + (b % 5) * 750) * s
        by = (800 + (b // 5) * 2000) * s
        for i in range(10):
            ribbons.append([bx + 20 * i * s, by, bx + 20 * i * s, by + 200 * s])
    w.ribbons = np.array(ribbons, dtype=np.float64)
    blocked = _rects(size, n_rects, 10 * s, 80 * s, rng, [(start[0], start[1])], 50 * s, w.ribbons, 5.0)
    w.set_grid(blocked, 1.0)
    return w


def world_c5(size=4096):
    w = world_c4(size=size)
    w.name = "C5"
    s = size / 4096.0
    # 50 Gaussian obstacles spread over the sampled region (edge sources are uniform over the map)
    return _add_obstacles(w, "gaussian", 50, 0.05 * size, 0.95 * size, 5)


def with_resolution(world, resolution):
    """The same world with its occupancy grid resampled to `resolution` metres per cell over the same extent
    (nearest cell).  Exercises x / res by division (non power-of-two resolutions, GridWorldMap.cpp:84-93) and the
    resolution-dependent dilation radius of the engine's chunk culling."""
    old = world.blocked_cells()
    ext_y, ext_x = old.shape[0] * world.resolution, old.shape[1] * world.resolution
    rows, cols = int(round(ext_y / resolution)), int(round(ext_x / resolution))
    ry = np.clip(((np.arange(rows) + 0.5) * resolution / world.resolution).astype(np.int64), 0, old.shape[0] - 1)
    rx = np.clip(((np.arange(cols) + 0.5) * resolution / world.resolution).astype(np.int64), 0, old.shape[1] - 1)
    world.set_grid(old[np.ix_(ry, rx)], resolution)
    world.name += "@%gm" % resolution
    return world


def with_time_offset(world, t0):
    """Epoch-scale clock: the ROS node feeds state times ~1.7e9 s (seconds since 1970).  Start time, obstacle
    observation times and (through make_edges) every edge's source time move by t0."""
    world.cfg.start_state_time += t0
    world.start = world.start.copy()
    world.start[4] += t0
    if world.obstacles is not None:
        world.obstacles["time"] = world.obstacles["time"] + t0
    world.name += "+%g s" % t0
    return world


def with_covariances(world, seed=9):
    """Per-obstacle random symmetric positive-definite covariances instead of the manager's default
    [[30, 10], [10, 30]] (GaussianDynamicObstaclesManager.h:24-25)."""
    assert world.obstacle_kind == "gaussian"
    rng = np.random.default_rng(seed)
    n = len(world.obstacles["x"])
    a = rng.uniform(5, 60, n)
    d = rng.uniform(5, 60, n)
    b = rng.uniform(-0.8, 0.8, n) * np.sqrt(a * d)
    world.obstacles["cov"] = np.column_stack([a, b, b, d])
    world.name += "+cov"
    return world


WORLDS = {
    "c1": world_c1,
    "c2": world_c2,
    "c3": lambda: world_c3("gaussian"),
    "c3b": lambda: world_c3("binary"),
    "c4": world_c4,
    "c5": world_c5,
}


def make_edges(world, n, seed=5, ribbon_set=0, near_ribbons=0.0, reach=75.0, extent=None):
    """Edge sweep of SURVEY.md section 8d (C5): source uniform over free cells (or, with probability
    `near_ribbons`, within a few metres of a ribbon), heading U[0,2pi), source time = start + U[0,20],
    g U[0,20]; destination = source + U[-reach,reach]^2 (resampled while blocked / outside),
    heading U[0,2pi); rho alternates turning / coverage radius with coverageAllowed = (rho == coverage
    radius); speed = max (75 %) or slow (25 %).  Edges carry has_path = 0: the engine constructs the
    shortest Dubins path itself (Edge.cpp:78-80), as for every (vertex, sample) pair in expand()."""
    rng = np.random.default_rng(seed)
    cfg = world.cfg
    e = np.zeros(n, dtype=abi.EDGE_DTYPE)
    if world.map_bits is not None:
        xmax, ymax = world.cols * world.resolution, world.rows * world.resolution
        xmin = ymin = 0.0
    elif extent is not None:
        xmin, xmax, ymin, ymax = extent
    else:
        xmin, xmax = world.start[0] - 100, world.start[0] + 100
        ymin, ymax = world.start[1] - 100, world.start[1] + 100

    def sample_free(count, gen):
        xs = np.empty(count)
        ys = np.empty(count)
        todo = np.arange(count)
        while todo.size:
            x, y = gen(todo.size)
            ok = ~world.is_blocked(x, y)
            xs[todo[ok]] = x[ok]
            ys[todo[ok]] = y[ok]
            todo = todo[~ok]
        return xs, ys

    sx, sy = sample_free(n, lambda k: (rng.uniform(xmin, xmax, k), rng.uniform(ymin, ymax, k)))
    if near_ribbons > 0 and len(world.ribbons):
        pick = rng.uniform(size=n) < near_ribbons
        idx = np.flatnonzero(pick)
        rb = world.ribbons[rng.integers(0, len(world.ribbons), idx.size)]
        u = rng.uniform(-0.1, 1.1, idx.size)
        px = rb[:, 0] + u * (rb[:, 2] - rb[:, 0]) + rng.normal(0, 3.0, idx.size)
        py = rb[:, 1] + u * (rb[:, 3] - rb[:, 1]) + rng.normal(0, 3.0, idx.size)
        ok = ~world.is_blocked(px, py)
        sx[idx[ok]] = px[ok]
        sy[idx[ok]] = py[ok]
    e["src"][:, 0] = sx
    e["src"][:, 1] = sy
    e["src"][:, 2] = rng.uniform(0, 2 * math.pi, n)
    e["src"][:, 3] = cfg.max_speed
    e["src"][:, 4] = cfg.start_state_time + rng.uniform(0, 20, n)
    e["src_g"] = rng.uniform(0, 20, n)

    dx = np.empty(n)
    dy = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = sx[todo] + rng.uniform(-reach, reach, todo.size)
        y = sy[todo] + rng.uniform(-reach, reach, todo.size)
        ok = ~world.is_blocked(x, y)
        if world.map_bits is None:
            ok[:] = True
        dx[todo[ok]] = x[ok]
        dy[todo[ok]] = y[ok]
        todo = todo[~ok]
    e["dst"][:, 0] = dx
    e["dst"][:, 1] = dy
    e["dst"][:, 2] = rng.uniform(0, 2 * math.pi, n)
    if near_ribbons > 0 and len(world.ribbons):
        # a share of the near-ribbon edges runs along the ribbon direction so that coverage happens
        idx = np.flatnonzero(pick)
        along = idx[rng.uniform(size=idx.size) < 0.6]
        if along.size:
            rb = world.ribbons[rng.integers(0, len(world.ribbons), along.size)]
            hd = np.pi / 2 - np.arctan2(rb[:, 3] - rb[:, 1], rb[:, 2] - rb[:, 0])
            hd = np.where(rng.uniform(size=along.size) < 0.5, hd, hd + np.pi)
            hd = np.mod(hd, 2 * np.pi)
            e["src"][along, 2] = hd
            e["dst"][along, 2] = hd
            L = rng.uniform(5, reach, along.size)
            e["dst"][along, 0] = e["src"][along, 0] + L * np.sin(hd)
            e["dst"][along, 1] = e["src"][along, 1] + L * np.cos(hd)
    slow = rng.uniform(size=n) < 0.25
    e["dst"][:, 3] = np.where(slow, cfg.slow_speed, cfg.max_speed)
    cov = (np.arange(n) % 2) == 1
    e["coverage_allowed"] = cov.astype(np.int32)
    e["has_path"] = 0
    e["ribbon_set"] = ribbon_set
    return e
