"""numpy-facing wrappers over the C ABI (include/oxg.h).

These are the calls a user (or the parity tests) makes; each one is a single liboxg entry
point with host buffers.  Names follow the reference's domain: a *store* holds the planner's
tree / roadmap states (``Vec<Node<S>>`` in rrt.rs:61, prm.rs:58), a *world* is a declarative
``StateValidityChecker`` (oxmpl/src/base/validity.rs:39-48).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import PrmStats, SpaceDesc, check, lib

INVALID_INDEX = _lib.OXG_INVALID_INDEX


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a, shape_last=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None and (a.ndim != 2 or a.shape[1] != shape_last):
        raise ValueError(f"expected an array of shape (n, {shape_last}), got {a.shape}")
    return a


@dataclass(frozen=True)
class Space:
    """Mirror of ``oxg_space_desc``: a RealVector block of ``n_real`` coordinates followed by
    ``n_so2`` angles.  ``Space.real_vector(d)`` is RealVectorStateSpace
    (real_vector_state_space.rs), ``Space.so2()`` is SO2StateSpace (so2_state_space.rs),
    ``Space.compound(d, 1)`` is the SE2-style composition (not in the reference)."""

    n_real: int
    n_so2: int = 0
    w_real: float = 1.0
    w_so2: float = 1.0

    @staticmethod
    def real_vector(dim: int) -> "Space":
        return Space(dim, 0)

    @staticmethod
    def so2() -> "Space":
        return Space(0, 1)

    @staticmethod
    def compound(n_real: int, n_so2: int, w_real: float = 1.0, w_so2: float = 0.5) -> "Space":
        return Space(n_real, n_so2, w_real, w_so2)

    @property
    def dim(self) -> int:
        return self.n_real + self.n_so2

    def desc(self) -> SpaceDesc:
        return SpaceDesc(self.n_real, self.n_so2, self.w_real, self.w_so2)


def real_vector_max_extent(bounds) -> float:
    """RealVectorStateSpace::get_maximum_extent (real_vector_state_space.rs:103-118).
    Python floats are IEEE binary64 and ``math.sqrt`` is correctly rounded, so this is
    bit-identical to the Rust."""
    if any((not math.isfinite(lo)) or (not math.isfinite(hi)) for lo, hi in bounds):
        return 1.0
    s = 0.0
    for lo, hi in bounds:
        d = hi - lo
        s = s + d * d
    return math.sqrt(s)


def motion_resolution(extent: float, fraction: float = 0.05) -> float:
    """``space.get_longest_valid_segment_length() * 0.1`` (real_vector_state_space.rs:251-253,
    rrt.rs:97); fraction defaults to 0.05 (real_vector_state_space.rs:98)."""
    return (extent * fraction) * 0.1


def device_count() -> int:
    n = C.c_int(0)
    code = lib.oxg_device_count(C.byref(n))
    return n.value if code == 0 else 0


def fp64_peak():
    """(unfused FP64 add|mul operations per second, kernel ms) measured on the current GPU."""
    ops = C.c_double(0.0)
    ms = C.c_double(0.0)
    check(lib.oxg_fp64_peak(C.byref(ops), C.byref(ms)))
    return ops.value, ms.value


def fp32_peak():
    """(FP32 lane-instructions per second of the pre-filter's FADD2+FFMA2 mix, kernel ms)."""
    ops = C.c_double(0.0)
    ms = C.c_double(0.0)
    check(lib.oxg_fp32_peak(C.byref(ops), C.byref(ms)))
    return ops.value, ms.value


def l2_peak():
    """(bytes per second all SMs read from an L2-resident buffer, kernel ms) on the current GPU."""
    bps = C.c_double(0.0)
    ms = C.c_double(0.0)
    check(lib.oxg_l2_peak(C.byref(bps), C.byref(ms)))
    return bps.value, ms.value


def kernel_launches(reset: bool = False) -> int:
    return int(lib.oxg_kernel_launches(1 if reset else 0))


def tuning_set(key: str, value: int) -> None:
    """A/B switch for measurements (``oxg_tuning_set``): e.g. ``tuning_set("knn_tree", 1)`` forces
    the bounding-box hierarchy for every eligible k-NN, 0 the all-pairs scan, -1 automatic."""
    check(lib.oxg_tuning_set(key.encode(), int(value)))


class StateStore:
    """GPU-resident SoA f64 state store (replaces ``Vec<Node<S>>``)."""

    def __init__(self, space: Space, capacity: int = 0):
        self.space = space
        self._desc = space.desc()
        self._h = C.c_void_p()
        check(lib.oxg_store_create(C.byref(self._desc), capacity, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.oxg_store_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def clear(self):
        check(lib.oxg_store_clear(self._h))

    def reserve(self, capacity: int):
        check(lib.oxg_store_reserve(self._h, capacity))

    def append(self, rows) -> int:
        rows = _f64(np.atleast_2d(rows), self.space.dim)
        first = C.c_uint64(0)
        check(lib.oxg_store_append(self._h, _ptr(rows), rows.shape[0], C.byref(first)))
        return first.value

    def append_ptr(self, host_ptr: int, n: int) -> int:
        """Append n row-major states from a raw host address (e.g. pinned memory)."""
        first = C.c_uint64(0)
        check(lib.oxg_store_append(self._h, C.c_void_p(host_ptr), n, C.byref(first)))
        return first.value

    def __len__(self) -> int:
        n = C.c_uint64(0)
        check(lib.oxg_store_size(self._h, C.byref(n)))
        return n.value

    def read(self, first: int = 0, n: int | None = None) -> np.ndarray:
        if n is None:
            n = len(self) - first
        out = np.empty((n, self.space.dim), dtype=np.float64)
        check(lib.oxg_store_read(self._h, first, n, _ptr(out)))
        return out

    def sync(self):
        check(lib.oxg_store_sync(self._h))

    def last_scan_ms(self) -> float:
        """Device time of the kernels of the most recent nn1 / knn call (no copies)."""
        ms = C.c_double(0.0)
        check(lib.oxg_store_last_scan_ms(self._h, C.byref(ms)))
        return ms.value

    def tree_stats(self) -> dict:
        """Counters of the last k-NN answered through the bounding-box hierarchy (needs
        ``tuning_set("tree_stats", 1)`` before the call)."""
        out = (C.c_double * 8)()
        check(lib.oxg_store_tree_stats(self._h, out))
        names = ("build_ms", "query_kernel_ms", "leaves_visited", "box_groups_tested", "candidates",
                 "merges", "exact_leaves", "queries")
        return dict(zip(names, [float(v) for v in out]))

    def stream(self) -> int:
        s = C.c_void_p()
        check(lib.oxg_store_stream(self._h, C.byref(s)))
        return s.value or 0

    # -- nearest neighbour ---------------------------------------------------------
    def nn1(self, queries):
        """1-NN scan of rrt.rs:187-196 for each query row -> (idx u32[Q], dist f64[Q])."""
        q = _f64(np.atleast_2d(queries), self.space.dim)
        idx = np.empty(q.shape[0], dtype=np.uint32)
        dist = np.empty(q.shape[0], dtype=np.float64)
        check(lib.oxg_nn1(self._h, _ptr(q), q.shape[0], _ptr(idx), _ptr(dist)))
        return idx, dist

    def knn(self, queries, k: int, exclude=None):
        q = _f64(np.atleast_2d(queries), self.space.dim)
        idx = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float64)
        ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint32)
        check(lib.oxg_knn(self._h, _ptr(q), q.shape[0], k, _ptr(ex), _ptr(idx), _ptr(dist)))
        return idx, dist

    def radius(self, queries, r: float, idx_limit=None, want_dist: bool = False):
        """Radius scan of rrt_star.rs:121-131 / prm.rs:131-138 -> CSR (offsets, idx[, dist])."""
        q = _f64(np.atleast_2d(queries), self.space.dim)
        Q = q.shape[0]
        lim = None if idx_limit is None else np.ascontiguousarray(idx_limit, dtype=np.uint32)
        offsets = np.zeros(Q + 1, dtype=np.uint64)
        cap = max(1024, 64 * Q)
        while True:
            idx = np.empty(cap, dtype=np.uint32)
            dist = np.empty(cap, dtype=np.float64) if want_dist else None
            code = lib.oxg_radius(
                self._h, _ptr(q), Q, float(r), _ptr(lim), _ptr(offsets), _ptr(idx), cap, _ptr(dist)
            )
            if code == _lib.OXG_E_CAPACITY:
                cap = int(offsets[Q])
                continue
            check(code)
            break
        n = int(offsets[Q])
        if want_dist:
            return offsets, idx[:n].copy(), dist[:n].copy()
        return offsets, idx[:n].copy()


    # -- SO(3): stores of unit quaternions (x, y, z, w), include/oxg_so3.h --------------------
    def so3_nn1(self, queries):
        """1-NN under ``SO3StateSpace::distance`` (so3_state_space.rs:102-112): the reference's
        winner (lowest index at the minimum distance) and its libm distance."""
        q = _f64(np.atleast_2d(queries), 4)
        idx = np.empty(q.shape[0], dtype=np.uint32)
        dist = np.empty(q.shape[0], dtype=np.float64)
        check(lib.oxg_so3_nn1(self._h, _ptr(q), q.shape[0], _ptr(idx), _ptr(dist)))
        return idx, dist

    def so3_knn(self, queries, k: int, exclude=None):
        """k nearest stored quaternions by (SO3 distance, index) ascending (``oxg_so3_knn``)."""
        q = _f64(np.atleast_2d(queries), 4)
        ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint32)
        if ex is not None and ex.shape != (q.shape[0],):
            raise ValueError("exclude needs one index per query")
        idx = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float64)
        check(lib.oxg_so3_knn(self._h, _ptr(q), q.shape[0], int(k), _ptr(ex), _ptr(idx), _ptr(dist)))
        return idx, dist

    def so3_radius(self, queries, r: float, want_dist: bool = False):
        """Radius scan ``distance(query, stored) < r`` over SO3 -> CSR (offsets, idx[, dist])."""
        q = _f64(np.atleast_2d(queries), 4)
        Q = q.shape[0]
        offsets = np.zeros(Q + 1, dtype=np.uint64)
        cap = max(1024, 64 * Q)
        while True:
            idx = np.empty(cap, dtype=np.uint32)
            dist = np.empty(cap, dtype=np.float64) if want_dist else None
            code = lib.oxg_so3_radius(self._h, _ptr(q), Q, float(r), _ptr(offsets), _ptr(idx), cap,
                                      _ptr(dist))
            if code == _lib.OXG_E_CAPACITY:
                cap = int(offsets[Q])
                continue
            check(code)
            break
        n = int(offsets[Q])
        if want_dist:
            return offsets, idx[:n].copy(), dist[:n].copy()
        return offsets, idx[:n].copy()


class So3Cones:
    """Forbidden cones over SO(3): a state is valid iff ``distance(centre, state) > angle`` for
    every cone (the reference fixture ``ForbiddenConeChecker``, oxmpl/tests/rrt_so3ss_tests.rs:46-56).
    ``cones``: (n, 5) rows ``[x, y, z, w, angle]``."""

    def __init__(self, cones):
        c = _f64(np.reshape(np.asarray(cones, dtype=np.float64), (-1, 5)), 5)
        self._h = C.c_void_p()
        check(lib.oxg_so3_cones_create(_ptr(c), c.shape[0], C.byref(self._h)))

    def close(self):
        if self._h:
            lib.oxg_so3_cones_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def is_valid(self, states) -> np.ndarray:
        st = _f64(np.atleast_2d(states), 4)
        out = np.empty(st.shape[0], dtype=np.uint8)
        check(lib.oxg_so3_is_valid(self._h, _ptr(st), st.shape[0], _ptr(out)))
        return out

    def check_motion(self, frm, to, resolution: float, want_first_invalid=False):
        """``check_motion(from, to)`` (rrt.rs:90-116) with SO3 interpolation."""
        a = _f64(np.atleast_2d(frm), 4)
        b = _f64(np.atleast_2d(to), 4)
        if a.shape != b.shape:
            raise ValueError("from/to shape mismatch")
        valid = np.empty(a.shape[0], dtype=np.uint8)
        first = np.empty(a.shape[0], dtype=np.uint32) if want_first_invalid else None
        check(lib.oxg_so3_check_motion(self._h, _ptr(a), _ptr(b), a.shape[0], float(resolution),
                                       _ptr(valid), _ptr(first)))
        return (valid, first) if want_first_invalid else valid


class World:
    """Declarative obstacle set implementing ``StateValidityChecker::is_valid``.

    spheres: (n, dim+1) rows ``[centre..., r]``; aabbs: (n, 2*dim) rows ``[lo..., hi...]``;
    capsules: (n, 2*dim+1) rows ``[a..., b..., r]``."""

    def __init__(self, dim: int, spheres=None, aabbs=None, capsules=None, so2_intervals=()):
        self.dim = dim
        sp = _f64(spheres, dim + 1) if spheres is not None and len(spheres) else None
        bx = _f64(aabbs, 2 * dim) if aabbs is not None and len(aabbs) else None
        cp = _f64(capsules, 2 * dim + 1) if capsules is not None and len(capsules) else None
        self._h = C.c_void_p()
        check(
            lib.oxg_world_create(
                dim,
                _ptr(sp), 0 if sp is None else sp.shape[0],
                _ptr(bx), 0 if bx is None else bx.shape[0],
                _ptr(cp), 0 if cp is None else cp.shape[0],
                C.byref(self._h),
            )
        )
        for comp, lo, hi in so2_intervals:
            check(lib.oxg_world_add_so2_interval(self._h, comp, float(lo), float(hi)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib.oxg_world_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def is_valid(self, space: Space, states) -> np.ndarray:
        st = _f64(np.atleast_2d(states), space.dim)
        out = np.empty(st.shape[0], dtype=np.uint8)
        d = space.desc()
        check(lib.oxg_is_valid(self._h, C.byref(d), _ptr(st), st.shape[0], _ptr(out)))
        return out

    def check_motion(self, space: Space, frm, to, resolution: float, want_first_invalid=False):
        """Batched ``check_motion(from, to)`` (rrt.rs:90-116)."""
        a = _f64(np.atleast_2d(frm), space.dim)
        b = _f64(np.atleast_2d(to), space.dim)
        if a.shape != b.shape:
            raise ValueError("from/to shape mismatch")
        E = a.shape[0]
        valid = np.empty(E, dtype=np.uint8)
        first = np.empty(E, dtype=np.uint32) if want_first_invalid else None
        d = space.desc()
        check(
            lib.oxg_check_motion(
                self._h, C.byref(d), _ptr(a), _ptr(b), E, float(resolution), _ptr(valid), _ptr(first)
            )
        )
        return (valid, first) if want_first_invalid else valid

    def check_edges(self, store: StateStore, from_idx, to_idx, resolution: float) -> np.ndarray:
        fi = np.ascontiguousarray(from_idx, dtype=np.uint32)
        ti = np.ascontiguousarray(to_idx, dtype=np.uint32)
        valid = np.empty(fi.shape[0], dtype=np.uint8)
        check(
            lib.oxg_check_edges(
                self._h, store.handle, _ptr(fi), _ptr(ti), fi.shape[0], float(resolution), _ptr(valid)
            )
        )
        return valid


def _stats_dict(st: PrmStats) -> dict:
    return {name: getattr(st, name) for name, _ in PrmStats._fields_}


def prm_build_knn(world: World, store: StateStore, k: int, resolution: float, row_begin: int = 0,
                  row_end: int | None = None):
    """k-NN PRM rows -> (idx[rows,k], dist[rows,k], valid[rows,k], stats)."""
    if row_end is None:
        row_end = len(store)
    rows = row_end - row_begin
    idx = np.empty((rows, k), dtype=np.uint32)
    dist = np.empty((rows, k), dtype=np.float64)
    valid = np.empty((rows, k), dtype=np.uint8)
    st = PrmStats()
    check(
        lib.oxg_prm_build_knn(
            world.handle, store.handle, k, float(resolution), row_begin, row_end, _ptr(idx),
            _ptr(dist), _ptr(valid), C.byref(st),
        )
    )
    return idx, dist, valid, _stats_dict(st)


def prm_solve(world: World, store: StateStore, lower_offsets, lower_idx, connection_radius: float,
              resolution: float, start, goal_target, goal_radius: float):
    """PRM::solve (prm.rs:227-307) on the device -> (status, path node indices, result dict)."""
    off = np.ascontiguousarray(lower_offsets, dtype=np.uint64)
    idx = np.ascontiguousarray(lower_idx, dtype=np.uint32)
    st, tg = _f64(start), _f64(goal_target)
    path = np.empty(max(len(store), 1), dtype=np.uint32)
    r = _lib.PrmQueryResult()
    check(lib.oxg_prm_solve(world.handle, store.handle, _ptr(off), _ptr(idx), float(connection_radius),
                            float(resolution), _ptr(st), _ptr(tg), float(goal_radius), _ptr(path),
                            path.shape[0], C.byref(r)))
    return int(r.status), path[: r.path_len].copy(), {n: getattr(r, n) for n, _ in r._fields_}


def prm_solve_knn(world: World, store: StateStore, idx, valid, connection_radius: float,
                  resolution: float, start, goal_target, goal_radius: float):
    """The same query on a k-NN roadmap (idx / valid [N][k] of prm_build_knn)."""
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    valid = np.ascontiguousarray(valid, dtype=np.uint8)
    st, tg = _f64(start), _f64(goal_target)
    path = np.empty(max(len(store), 1), dtype=np.uint32)
    r = _lib.PrmQueryResult()
    check(lib.oxg_prm_solve_knn(world.handle, store.handle, _ptr(idx), _ptr(valid), idx.shape[1],
                                float(connection_radius), float(resolution), _ptr(st), _ptr(tg),
                                float(goal_radius), _ptr(path), path.shape[0], C.byref(r)))
    return int(r.status), path[: r.path_len].copy(), {n: getattr(r, n) for n, _ in r._fields_}


def prm_build_radius(world: World, store: StateStore, connection_radius: float, resolution: float,
                     row_begin: int = 0, row_end: int | None = None):
    """Reference-faithful radius PRM rows -> CSR (offsets, idx) of ``new_node.edges``
    (prm.rs:131-138), plus stats."""
    if row_end is None:
        row_end = len(store)
    rows = row_end - row_begin
    offsets = np.zeros(rows + 1, dtype=np.uint64)
    cap = max(1024, 32 * rows)
    st = PrmStats()
    while True:
        idx = np.empty(cap, dtype=np.uint32)
        code = lib.oxg_prm_build_radius(
            world.handle, store.handle, float(connection_radius), float(resolution), row_begin,
            row_end, _ptr(offsets), _ptr(idx), cap, C.byref(st),
        )
        if code == _lib.OXG_E_CAPACITY:
            cap = int(offsets[rows])
            continue
        check(code)
        break
    return offsets, idx[: int(offsets[rows])].copy(), _stats_dict(st)


# -- fused sequential planners ---------------------------------------------------------------
PLAN_STATUS = {0: "solved", 1: "samples_exhausted", 2: "tree_full", 3: "timeout"}


def _plan_params(max_distance, resolution, max_nodes, goal_radius, search_radius=0.0,
                 timeout_s=0.0, stop_at_goal=True):
    return _lib.PlanParams(float(max_distance), float(search_radius), float(resolution),
                           float(goal_radius), float(timeout_s), int(max_nodes),
                           1 if stop_at_goal else 0, 0)


def _plan_result(res) -> dict:
    return {
        "n_nodes": (int(res.n_nodes[0]), int(res.n_nodes[1])),
        "iters": int(res.iters),
        "motion_checks": int(res.motion_checks),
        "motion_checks_executed": int(res.motion_checks_executed),
        "goal_node": (int(res.goal_node[0]), int(res.goal_node[1])),
        "solved": bool(res.solved),
        "status": PLAN_STATUS.get(int(res.status), int(res.status)),
        "device_ms": float(res.device_ms),
        "ctas": int(res.ctas),
        "phase_ms": [float(x) for x in res.phase_ms],
    }


def rrt_solve(world: World, tree: StateStore, samples, max_distance: float, resolution: float,
              max_nodes: int, goal_target=None, goal_radius: float = -1.0, timeout_s: float = 0.0):
    """RRT::solve (rrt.rs:158-227) over a pre-drawn sample stream, as one persistent kernel.
    `tree` must hold the start state.  Returns (parents i64[n], result dict)."""
    s = _f64(np.atleast_2d(samples), tree.space.dim)
    g = None if goal_target is None else _f64(goal_target)
    prm = _plan_params(max_distance, resolution, max_nodes, goal_radius, timeout_s=timeout_s)
    parents = np.full(max_nodes, -1, dtype=np.int64)
    res = _lib.PlanResult()
    check(lib.oxg_rrt_solve(world.handle, tree.handle, C.byref(prm), _ptr(g), _ptr(s), s.shape[0],
                            _ptr(parents), C.byref(res)))
    return parents[: int(res.n_nodes[0])].copy(), _plan_result(res)


def so3_rrt_solve(cones: "So3Cones", tree: StateStore, samples, max_distance: float, resolution: float,
                  max_nodes: int, goal_target=None, goal_radius: float = -1.0, timeout_s: float = 0.0,
                  parents=None):
    """RRT::solve over SO3 (``oxg_so3_rrt_solve``): speculative chunks of samples per device call,
    results those of the one-sample-at-a-time loop.  ``parents`` continues an earlier call's tree.
    Returns (parents i64[n], result dict); ``result["ctas"]`` is the number of chunks."""
    s = _f64(np.atleast_2d(samples), 4)
    g = None if goal_target is None else _f64(goal_target)
    prm = _plan_params(max_distance, resolution, max_nodes, goal_radius, timeout_s=timeout_s)
    par = np.full(max_nodes, -1, dtype=np.int64)
    if parents is not None:
        par[: len(parents)] = parents
    res = _lib.PlanResult()
    check(lib.oxg_so3_rrt_solve(tree.handle, cones._h, C.byref(prm), _ptr(g), _ptr(s), s.shape[0],
                                _ptr(par), C.byref(res)))
    return par[: int(res.n_nodes[0])].copy(), _plan_result(res)


def so3_rrt_connect_solve(cones: "So3Cones", start_tree: StateStore, goal_tree: StateStore, samples,
                          max_distance: float, resolution: float, max_nodes: int, goal_target=None,
                          goal_radius: float = -1.0, timeout_s: float = 0.0, start_parents=None,
                          goal_parents=None):
    """RRTConnect::solve over SO3 (``oxg_so3_rrt_connect_solve``), speculative chunks.
    Returns (start_parents, goal_parents, result dict)."""
    s = _f64(np.atleast_2d(samples), 4)
    g = None if goal_target is None else _f64(goal_target)
    prm = _plan_params(max_distance, resolution, max_nodes, goal_radius, timeout_s=timeout_s)
    pa = np.full(max_nodes, -1, dtype=np.int64)
    pb = np.full(max_nodes, -1, dtype=np.int64)
    if start_parents is not None:
        pa[: len(start_parents)] = start_parents
    if goal_parents is not None:
        pb[: len(goal_parents)] = goal_parents
    res = _lib.PlanResult()
    check(lib.oxg_so3_rrt_connect_solve(start_tree.handle, goal_tree.handle, cones._h, C.byref(prm),
                                        _ptr(g), _ptr(s), s.shape[0], _ptr(pa), _ptr(pb),
                                        C.byref(res)))
    return pa[: int(res.n_nodes[0])].copy(), pb[: int(res.n_nodes[1])].copy(), _plan_result(res)


def rrt_connect_solve(world: World, start_tree: StateStore, goal_tree: StateStore, samples,
                      max_distance: float, resolution: float, max_nodes: int, goal_target=None,
                      goal_radius: float = -1.0, timeout_s: float = 0.0):
    """RRTConnect::solve (rrt_connect.rs:227-309).  Returns (start_parents, goal_parents, result)."""
    s = _f64(np.atleast_2d(samples), start_tree.space.dim)
    g = None if goal_target is None else _f64(goal_target)
    prm = _plan_params(max_distance, resolution, max_nodes, goal_radius, timeout_s=timeout_s)
    pa = np.full(max_nodes, -1, dtype=np.int64)
    pb = np.full(max_nodes, -1, dtype=np.int64)
    res = _lib.PlanResult()
    check(lib.oxg_rrt_connect_solve(world.handle, start_tree.handle, goal_tree.handle, C.byref(prm),
                                    _ptr(g), _ptr(s), s.shape[0], _ptr(pa), _ptr(pb), C.byref(res)))
    return pa[: int(res.n_nodes[0])].copy(), pb[: int(res.n_nodes[1])].copy(), _plan_result(res)


def rrt_star_solve(world: World, tree: StateStore, samples, max_distance: float,
                   search_radius: float, resolution: float, max_nodes: int, goal_target=None,
                   goal_radius: float = -1.0, stop_at_goal: bool = True, timeout_s: float = 0.0):
    """RRTStar::solve (rrt_star.rs:171-290).  Returns (parents, costs, result)."""
    s = _f64(np.atleast_2d(samples), tree.space.dim)
    g = None if goal_target is None else _f64(goal_target)
    prm = _plan_params(max_distance, resolution, max_nodes, goal_radius, search_radius, timeout_s,
                       stop_at_goal)
    parents = np.full(max_nodes, -1, dtype=np.int64)
    costs = np.zeros(max_nodes, dtype=np.float64)
    res = _lib.PlanResult()
    check(lib.oxg_rrt_star_solve(world.handle, tree.handle, C.byref(prm), _ptr(g), _ptr(s),
                                 s.shape[0], _ptr(parents), _ptr(costs), C.byref(res)))
    n = int(res.n_nodes[0])
    return parents[:n].copy(), costs[:n].copy(), _plan_result(res)
