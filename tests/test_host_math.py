"""Host build (g++ / glibc) of the engine's scalar building blocks (path_planner_b200/csrc/
ppe_math.cuh, ppe_crmath.cuh) pinned against the oracle: same formulas, same operation order."""
import ctypes as C
import math

import numpy as np
import pytest

from path_planner_b200 import abi
from tests import common

D = C.POINTER(C.c_double)


def _ulps(a, b):
    return np.abs(a.view(np.int64) - b.view(np.int64))


def _pairs(n, seed=0):
    rng = np.random.default_rng(seed)
    q0 = np.column_stack([rng.uniform(-100, 100, n), rng.uniform(-100, 100, n), rng.uniform(0, 2 * math.pi, n)])
    q1 = np.column_stack([q0[:, 0] + rng.uniform(-75, 75, n), q0[:, 1] + rng.uniform(-75, 75, n), rng.uniform(0, 2 * math.pi, n)])
    q1[: n // 4, 0] = q0[: n // 4, 0] + rng.uniform(-10, 10, n // 4)
    q1[: n // 4, 1] = q0[: n // 4, 1] + rng.uniform(-10, 10, n // 4)
    rho = np.where(np.arange(n) % 2 == 0, 8.0, 16.0)
    return np.ascontiguousarray(q0), np.ascontiguousarray(q1), rho


def _hh_dubins(hh, q0, q1, rho):
    n = len(rho)
    t = np.zeros(n, np.int32); p = np.zeros((n, 3)); l = np.zeros(n); e = np.zeros(n, np.int32)
    hh.hh_dubins_batch(C.c_int64(n), abi.dptr(q0), abi.dptr(q1), abi.dptr(rho), abi.iptr(t), abi.dptr(p), abi.dptr(l), abi.iptr(e))
    return t, p, l, e


def test_dubins_solver_is_bit_identical_to_the_cr_oracle(host_helpers):
    q0, q1, rho = _pairs(300_000)
    got = _hh_dubins(host_helpers, q0, q1, rho)
    want = common.load_oracle("cr").dubins_batch(q0, q1, rho)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert np.bincount(got[0], minlength=6).min() > 1000


def test_dubins_solver_vs_glibc_oracle(host_helpers):
    """Against oracle-A (== reference arithmetic): same word everywhere, parameters equal except
    where glibc's last bit is not the correctly rounded one."""
    q0, q1, rho = _pairs(300_000, seed=3)
    gt, gp, gl, ge = _hh_dubins(host_helpers, q0, q1, rho)
    wt, wp, wl, we = common.load_oracle("glibc").dubins_batch(q0, q1, rho)
    assert np.array_equal(gt, wt) and np.array_equal(ge, we)
    assert np.allclose(gp, wp, rtol=1e-12, atol=1e-12)
    frac = np.mean((gp != wp).any(axis=1))
    assert frac < 0.02, frac


def test_time_walker_replays_repeated_addition(host_helpers):
    hh = host_helpers
    hh.hh_time_walk.argtypes = [C.c_double, C.c_double, C.c_int, D]
    hh.hh_time_walk_lane.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, D]
    rng = np.random.default_rng(5)
    starts = list(rng.uniform(0.03, 40, 200)) + [1.0, 1.02, 2.0, 4.0, 1.6e9 + 0.3, 0.0625, 0.05, 0.031, 1e-3, 123456.789, 0.0]
    for t0 in starts:
        for dt in (0.02, 0.05 / 2.5, 0.1 / 3.0, 0.05, 0.25, 1e-3):
            m = 2200
            seq = np.zeros(m)
            t = t0
            for i in range(m):
                seq[i] = t
                t = t + dt
            out = np.zeros(m)
            hh.hh_time_walk(t0, dt, m, abi.dptr(out))
            assert np.array_equal(out, seq), (t0, dt)
            for lane in (0, 7, 31):
                k = (m - lane + 31) // 32
                lo = np.zeros(k)
                hh.hh_time_walk_lane(t0, dt, lane, k, abi.dptr(lo))
                assert np.array_equal(lo, seq[lane::32][:k]), (t0, dt, lane)


def test_skip_count_replays_repeated_subtraction(host_helpers):
    hh = host_helpers
    hh.hh_skip_count.argtypes = [C.c_double, C.c_double, C.c_int]
    hh.hh_skip_count.restype = C.c_int
    rng = np.random.default_rng(6)
    xs = list(rng.uniform(0, 300, 2000)) + list(rng.uniform(0, 1, 500)) + [0.0, 0.05, 0.1, 0.15000000000000002, 0.25, 0.125, 0.2, 64.0, 128.0, 100.0]
    for x in xs:
        for c in (0.05, 0.1, 0.03):
            v, k = x, 0
            while v > c:
                v -= c
                k += 1
            assert hh.hh_skip_count(x, c, 1 << 28) == k, (x, c)
    assert hh.hh_skip_count(1e9, 0.05, 1000) == 1000  # capped
    # adversarial: distances within a few ulps of an integer multiple of the increment, where the closed form
    # (ceil(x / c) - 1) must hand over to the exact replay
    for c in (0.05, 0.1, 0.03, 0.01):
        for m in list(rng.integers(1, 6000, 300)) + [1, 2, 3, 1500, 1501]:
            x0 = float(m) * c
            for d in (-3, -2, -1, 0, 1, 2, 3):
                x = x0
                for _ in range(abs(d)):
                    x = float(np.nextafter(x, np.inf if d > 0 else -np.inf))
                v, k = x, 0
                while v > c:
                    v -= c
                    k += 1
                assert hh.hh_skip_count(x, c, 1 << 28) == k, (x, c, m, d)


def test_hoisted_sampler_matches_dubins_path_sample(host_helpers):
    hh = host_helpers
    ora = common.load_oracle("cr")
    q0, q1, rho = _pairs(400, seed=9)
    typ, par, length, err = ora.dubins_batch(q0, q1, rho)
    rng = np.random.default_rng(1)
    I = C.POINTER(C.c_int32)
    ora.lib.oracle_wrapper_sample.argtypes = [D, D, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, D, D, D, D, I]
    hh.hh_sample.argtypes = [D, D, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, D, D, D, D, I]
    for i in range(len(rho)):
        speed = 2.5 if i % 3 else 0.5
        t_end = 1.0 + length[i] / speed
        times = np.concatenate([np.linspace(1.0, t_end, 40), [t_end, t_end + 1e-9, 1.0 + rng.uniform(0, 1) * (t_end - 1.0)]])
        m = len(times)
        res = []
        for fn in (hh.hh_sample, ora.lib.oracle_wrapper_sample):
            x = np.zeros(m); y = np.zeros(m); h = np.zeros(m); ok = np.zeros(m, np.int32)
            qi = np.ascontiguousarray(q0[i]); pp = np.ascontiguousarray(par[i])
            fn(abi.dptr(qi), abi.dptr(pp), float(rho[i]), int(typ[i]), 1.0, speed, m, abi.dptr(times), abi.dptr(x), abi.dptr(y), abi.dptr(h), abi.iptr(ok))
            res.append((x, y, h, ok))
        inside = times <= t_end  # the oracle refuses times beyond the wrapper's end (containsTime)
        assert np.array_equal(res[0][3][inside], res[1][3][inside])
        for a, b in zip(res[0][:3], res[1][:3]):
            assert np.array_equal(a[inside], b[inside]), i


def test_cr_functions_are_correctly_rounded_and_agree_with_glibc(host_helpers):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 300
    hh = host_helpers
    rng = np.random.default_rng(2)
    n = 400_000
    x = np.concatenate([rng.uniform(-4 * math.pi, 4 * math.pi, n), [k * math.pi / 2 for k in range(-8, 9)], [0.0, 1e-300, -1e-12]])
    s = np.zeros_like(x); c = np.zeros_like(x); gs = np.zeros_like(x); gc = np.zeros_like(x)
    hh.hh_cr_sincos(C.c_int64(len(x)), abi.dptr(x), abi.dptr(s), abi.dptr(c))
    hh.hh_libm_sincos(C.c_int64(len(x)), abi.dptr(x), abi.dptr(gs), abi.dptr(gc))
    assert _ulps(s, gs).max() <= 1 and _ulps(c, gc).max() <= 1
    assert np.mean(s != gs) < 5e-3 and np.mean(c != gc) < 5e-3
    y = np.concatenate([rng.uniform(-3, 3, n), [0.0, -0.0, 0.0, -0.0, 2.0, -2.0, 5.0]])
    xx = np.concatenate([rng.uniform(-200, 200, n), [1.0, 1.0, -1.0, -1.0, 0.0, 0.0, -0.0]])
    o = np.zeros_like(y); g = np.zeros_like(y)
    hh.hh_cr_atan2(C.c_int64(len(y)), abi.dptr(y), abi.dptr(xx), abi.dptr(o))
    hh.hh_libm_atan2(C.c_int64(len(y)), abi.dptr(y), abi.dptr(xx), abi.dptr(g))
    assert _ulps(o, g).max() <= 1 and np.mean(o != g) < 5e-3
    assert np.array_equal(np.signbit(o[-7:]), np.signbit(g[-7:])) and np.array_equal(o[-7:], g[-7:])
    a = np.concatenate([rng.uniform(-1, 1, n), [1.0, -1.0, 0.0, 0.5, -0.5]])
    oa = np.zeros_like(a); ga = np.zeros_like(a)
    hh.hh_cr_acos(C.c_int64(len(a)), abi.dptr(a), abi.dptr(oa))
    hh.hh_libm_acos(C.c_int64(len(a)), abi.dptr(a), abi.dptr(ga))
    assert _ulps(oa, ga).max() <= 1 and np.mean(oa != ga) < 5e-3

    def err(v, true):
        u = abs(float(np.spacing(v)))
        return float(abs(mp.mpf(float(v)) - true) / u) if u > 0 else 0.0

    # correct rounding against 300-bit mpmath, including every place where glibc disagrees
    idx = np.concatenate([rng.integers(0, n, 300), np.flatnonzero(s != gs)[:50], np.flatnonzero(c != gc)[:50]])
    for i in idx:
        xm = mp.mpf(float(x[i]))
        assert err(s[i], mp.sin(xm)) <= 0.5000001 and err(c[i], mp.cos(xm)) <= 0.5000001
    for i in np.concatenate([rng.integers(0, n, 300), np.flatnonzero(o != g)[:50]]):
        assert err(o[i], mp.atan2(mp.mpf(float(y[i])), mp.mpf(float(xx[i])))) <= 0.5000001
    for i in np.concatenate([rng.integers(0, n, 300), np.flatnonzero(oa != ga)[:50]]):
        assert err(oa[i], mp.acos(mp.mpf(float(a[i])))) <= 0.5000001


def test_zero_aware_division_is_the_ieee_quotient(host_helpers):
    """`div_zero_aware` (the guard that keeps exact-zero numerators away from CUDA's out-of-line division routine) returns
    the IEEE quotient's bits for every operand class, signs of zero included; `ribbon_projection` through it equals
    Ribbon::getProjection (Ribbon.cpp:72-78) written with plain divisions, axis-aligned ribbons included."""
    rng = np.random.default_rng(9)
    special = np.array([0.0, -0.0, 1.0, -1.0, 5e-324, -5e-324, 1e-310, 1e308, -1e308, np.inf, -np.inf, np.nan, 3.5, -2.25e-200])
    num = np.concatenate([np.repeat(special, len(special)), rng.normal(0, 1e3, 4000), np.zeros(500), -np.zeros(500)])
    den = np.concatenate([np.tile(special, len(special)), rng.normal(0, 1e3, 4000), rng.uniform(1e-3, 1e6, 500), rng.uniform(1e-3, 1e6, 500)])
    out = np.zeros(len(num))
    host_helpers.hh_div_zero_aware(C.c_int64(len(num)), abi.dptr(num), abi.dptr(den), abi.dptr(out))
    with np.errstate(all="ignore"):
        want = num / den
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(out), np.isnan(want))
    assert np.array_equal(out[ok].view(np.int64), want[ok].view(np.int64))
    n = 6000
    rib = rng.uniform(-500, 500, (n, 4))
    rib[:2000, 2] = rib[:2000, 0]  # vertical ribbons: (ex - sx) * dot is an exact zero
    rib[2000:4000, 3] = rib[2000:4000, 1]  # horizontal ones
    rib = np.ascontiguousarray(rib)
    x, y = rng.uniform(-500, 500, n), rng.uniform(-500, 500, n)
    px, py = np.zeros(n), np.zeros(n)
    host_helpers.hh_ribbon_projection(C.c_int64(n), abi.dptr(rib), abi.dptr(x), abi.dptr(y), abi.dptr(px), abi.dptr(py))
    sx, sy, ex, ey = rib.T
    sq = (ex - sx) * (ex - sx) + (ey - sy) * (ey - sy)
    dot = (x - sx) * (ex - sx) + (y - sy) * (ey - sy)
    wx = (ex - sx) * dot / sq + sx
    wy = (ey - sy) * dot / sq + sy
    assert np.array_equal(px.view(np.int64), wx.view(np.int64)) and np.array_equal(py.view(np.int64), wy.view(np.int64))
