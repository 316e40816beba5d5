"""Synthetic obstacle worlds and spaces shared by the tests, the smoke test and bench.py.

The generators are the record of "what exactly was measured" for SURVEY §8(d) configs 1-4.
"""
from __future__ import annotations

import math

import numpy as np

# Franka-Panda-like joint limits (config 4)
PANDA_LO = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
PANDA_HI = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])


def capsule_world(seed: int, lo, hi, n_capsules: int = 256, max_len: float = 3.0,
                  r_lo: float = 1.0, r_hi: float = 1.6) -> np.ndarray:
    """n capsules in R^D: endpoint a uniform in the box, b = a + u * len (u uniform direction,
    len uniform in [0, max_len]), radius uniform in [r_lo, r_hi].  Rows: [a..., b..., r].
    Defaults are the config-4 world: with the Panda-like box they leave ~82% of uniform samples
    valid (radii of 0.1-0.3 would leave the 7-D box 100% free — measured with the oracle)."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    d = lo.shape[0]
    a = lo + (hi - lo) * rng.random((n_capsules, d))
    u = rng.normal(size=(n_capsules, d))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    length = max_len * rng.random((n_capsules, 1))
    b = a + u * length
    r = r_lo + (r_hi - r_lo) * rng.random((n_capsules, 1))
    return np.ascontiguousarray(np.hstack([a, b, r]))


def sphere_world(seed: int, lo, hi, n: int, r_lo: float, r_hi: float, keep_clear=()) -> np.ndarray:
    """n spheres, centres uniform in the box, radius uniform in [r_lo, r_hi]; spheres containing
    any of the `keep_clear` points are re-drawn.  Rows: [c..., r]."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    d = lo.shape[0]
    out = np.empty((n, d + 1))
    i = 0
    while i < n:
        c = lo + (hi - lo) * rng.random(d)
        r = r_lo + (r_hi - r_lo) * rng.random()
        if any(np.linalg.norm(np.asarray(p) - c) <= r * 1.5 for p in keep_clear):
            continue
        out[i, :d] = c
        out[i, d] = r
        i += 1
    return out


def aabb_world(seed: int, lo, hi, n: int, side_lo: float, side_hi: float, keep_clear=()) -> np.ndarray:
    """n axis-aligned boxes with side lengths uniform in [side_lo, side_hi]; boxes containing a
    `keep_clear` point (with a margin) are re-drawn.  Rows: [lo..., hi...]."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    d = lo.shape[0]
    out = np.empty((n, 2 * d))
    i = 0
    while i < n:
        side = side_lo + (side_hi - side_lo) * rng.random(d)
        c = lo + (hi - lo) * rng.random(d)
        blo, bhi = c - side / 2, c + side / 2
        if any(np.all(np.asarray(p) >= blo - 1.0) and np.all(np.asarray(p) <= bhi + 1.0)
               for p in keep_clear):
            continue
        out[i, :d] = blo
        out[i, d:] = bhi
        i += 1
    return out


def sample_stream(seed: int, n: int, lo, hi, goal_bias: float, goal_point) -> np.ndarray:
    """Pre-generated planner sample stream: with probability goal_bias the goal
    point, else uniform in the box."""
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)
    s = lo + (hi - lo) * rng.random((n, lo.shape[0]))
    if goal_bias > 0:
        coin = rng.random(n) < goal_bias
        s[coin] = np.asarray(goal_point, float)
    return s


# config 1: README RRT world (README.md:97-150, oxmpl-js/examples/simple_2d_planning.js:8-40)
CONFIG1 = dict(
    lo=[-10.0, -10.0], hi=[10.0, 10.0],
    spheres=np.array([[0.0, 0.0, 2.0]]),
    start=[-5.0, -5.0], goal=[5.0, 5.0], goal_radius=0.5,
    max_distance=0.5, goal_bias=0.05,
)

# wall fixture: oxmpl/tests/rrt_rvss_tests.rs:111-143
WALL = dict(
    lo=[0.0, 0.0], hi=[10.0, 10.0],
    aabbs=np.array([[5.0 - 0.5 / 2.0, 2.0, 5.0 + 0.5 / 2.0, 8.0]]),
    start=[1.0, 5.0], goal=[9.0, 5.0], goal_radius=0.5,
)

# forbidden-angle fixture: oxmpl/tests/rrt_so2ss_tests.rs:17-27,99-125
SO2_FORBIDDEN = dict(interval=(0, -0.5, 0.5), start=[-math.pi / 2], goal=[math.pi / 2],
                     goal_radius=0.1, max_distance=0.2, goal_bias=0.05)


def calibrated_obstacles(seed: int, dim: int, n_spheres: int, n_aabbs: int, n_capsules: int,
                         p_each: float = 0.012):
    """Obstacles in [-1,1]^dim sized so that EACH primitive covers about `p_each` of the box
    whatever the dimension (radii from the empirical distance quantile, box sides from
    p^(1/dim)) — small fixed radii cover nothing in 6+ dimensions.  Returns (spheres, aabbs,
    capsules) in the C-ABI row formats."""
    rng = np.random.default_rng(seed)
    lo, hi = -np.ones(dim), np.ones(dim)
    a = rng.uniform(-1, 1, (4000, dim))
    b = rng.uniform(-1, 1, (4000, dim))
    r_q = float(np.quantile(np.linalg.norm(a - b, axis=1), p_each))
    spheres = sphere_world(seed + 1, lo, hi, n_spheres, 0.85 * r_q, 1.1 * r_q)
    side = 2.0 * p_each ** (1.0 / dim)
    aabbs = aabb_world(seed + 2, lo, hi, n_aabbs, 0.85 * side, 1.1 * side)
    capsules = capsule_world(seed + 3, lo, hi, n_capsules, 0.8, 0.7 * r_q, 0.95 * r_q)
    return spheres, aabbs, capsules
