"""ctypes binding of the CPU oracle (oracle/oxg_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, by ``__graft_entry__.smoke()`` and by
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs — never by the product package
``oxmpl_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboxg_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oxg_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, env={**os.environ, "CXX": ""})
    return LIB_PATH


class OrcSpace(C.Structure):
    _fields_ = [("n_real", C.c_uint32), ("n_so2", C.c_uint32), ("w_real", C.c_double),
                ("w_so2", C.c_double)]


class OrcWorld(C.Structure):
    _fields_ = [
        ("dim", C.c_uint32),
        ("n_spheres", C.c_uint64), ("spheres", C.c_void_p),
        ("n_aabbs", C.c_uint64), ("aabbs", C.c_void_p),
        ("n_capsules", C.c_uint64), ("capsules", C.c_void_p),
        ("n_so2", C.c_uint64), ("so2_comp", C.c_void_p), ("so2_lohi", C.c_void_p),
    ]


class OrcCounters(C.Structure):
    _fields_ = [("states", C.c_uint64), ("prim_tests", C.c_uint64), ("flops", C.c_double)]


class OrcGoal(C.Structure):
    _fields_ = [("target", C.c_void_p), ("radius", C.c_double)]


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self):
        build()
        self.lib = lib = C.CDLL(LIB_PATH)
        d, vp, u64, u32 = C.c_double, C.c_void_p, C.c_uint64, C.c_uint32
        sp, w = C.POINTER(OrcSpace), C.POINTER(OrcWorld)
        cnt = C.POINTER(OrcCounters)
        sig = {
            "orc_so2_normalise": (d, [d]),
            "orc_distance": (d, [sp, vp, vp]),
            "orc_interpolate": (None, [sp, vp, vp, d, vp]),
            "orc_rv_max_extent": (d, [vp, vp, u32]),
            "orc_longest_valid_segment_length": (d, [d, d]),
            "orc_so2_max_extent": (d, []),
            "orc_motion_resolution": (d, [d]),
            "orc_clamp_fraction": (d, [d]),
            "orc_is_valid": (C.c_int, [w, sp, vp, cnt]),
            "orc_is_valid_batch": (None, [w, sp, vp, u64, vp]),
            "orc_check_motion": (C.c_int, [w, sp, vp, vp, d, C.POINTER(u32), cnt]),
            "orc_check_motion_batch": (None, [w, sp, vp, vp, u64, d, vp, vp, cnt]),
            "orc_nn1": (None, [sp, vp, u64, vp, C.POINTER(u32), C.POINTER(d)]),
            "orc_nn1_batch": (None, [sp, vp, u64, vp, u64, vp, vp]),
            "orc_radius": (u64, [sp, vp, u64, vp, d, u64, vp, vp, u64]),
            "orc_knn": (None, [sp, vp, u64, vp, u32, u32, vp, vp]),
            "orc_knn_batch": (None, [sp, vp, u64, vp, u64, u32, vp, vp, vp]),
            "orc_rrt": (u64, [w, sp, C.POINTER(OrcGoal), d, d, vp, u64, u64, vp, vp,
                              C.POINTER(C.c_int), C.POINTER(u64)]),
            "orc_rrt_connect": (C.c_int, [w, sp, C.POINTER(OrcGoal), d, d, vp, u64, u64, vp, vp,
                                          C.POINTER(u64), vp, vp, C.POINTER(u64),
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                          C.POINTER(u64)]),
            "orc_rrt_star": (u64, [w, sp, C.POINTER(OrcGoal), d, d, d, vp, u64, u64, C.c_int, vp,
                                   vp, vp, C.POINTER(C.c_int), C.POINTER(u64), C.POINTER(u64)]),
            "orc_prm_radius": (u64, [w, sp, d, d, vp, u64, vp, vp, vp, u64, C.POINTER(u64),
                                     C.POINTER(u64)]),
            "orc_prm_knn": (None, [w, sp, vp, u64, u32, d, u64, u64, vp, vp, vp, cnt]),
            "orc_prm_solve": (C.c_int, [w, sp, vp, u64, vp, vp, d, d, vp, C.POINTER(OrcGoal), vp, u32,
                                        C.POINTER(u32)]),
            "orc_so3_distance": (d, [vp, vp]),
            "orc_so3_interpolate": (None, [vp, vp, d, vp]),
            "orc_so3_nn1_batch": (None, [vp, u64, vp, u64, vp, vp]),
            "orc_so3_radius": (u64, [vp, u64, vp, d, vp, vp, u64]),
            "orc_so3_knn_batch": (None, [vp, u64, vp, u64, C.c_uint32, vp, vp, vp]),
            "orc_so3_is_valid": (C.c_int, [vp, u64, vp]),
            "orc_so3_rrt_connect": (C.c_int, [vp, u64, vp, d, d, d, vp, u64, u64, vp, vp,
                                              C.POINTER(u64), vp, vp, C.POINTER(u64),
                                              C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                              C.POINTER(u64)]),
            "orc_so3_rrt": (u64, [vp, u64, vp, d, d, d, vp, u64, u64, vp, vp, C.POINTER(C.c_int),
                                  C.POINTER(u64)]),
            "orc_so3_check_motion": (C.c_int, [vp, u64, vp, vp, d, C.POINTER(u32)]),
            "orc_build_info": (C.c_int, []),
            "orc_probe_mul_add": (d, [d, d, d]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args

    # -- helpers -------------------------------------------------------------------
    @staticmethod
    def space(n_real, n_so2=0, w_real=1.0, w_so2=1.0) -> OrcSpace:
        return OrcSpace(n_real, n_so2, w_real, w_so2)

    @staticmethod
    def world(dim, spheres=None, aabbs=None, capsules=None, so2_intervals=()):
        """Returns (OrcWorld, keepalive) — keep the second value referenced."""
        keep = []

        def arr(a):
            if a is None or len(a) == 0:
                return None, 0
            a = _f64(a)
            keep.append(a)
            return a, a.shape[0]

        s, ns = arr(spheres)
        b, nb = arr(aabbs)
        c, nc = arr(capsules)
        comp = np.ascontiguousarray([t[0] for t in so2_intervals], dtype=np.uint32)
        lohi = _f64([[t[1], t[2]] for t in so2_intervals]) if so2_intervals else _f64([])
        keep += [comp, lohi]
        w = OrcWorld(dim, ns, _p(s), nb, _p(b), nc, _p(c), len(so2_intervals), _p(comp), _p(lohi))
        return w, keep

    # -- scalar functions ----------------------------------------------------------
    def so2_normalise(self, v):
        return self.lib.orc_so2_normalise(v)

    def distance(self, sp, a, b):
        a, b = _f64(a), _f64(b)
        return self.lib.orc_distance(C.byref(sp), _p(a), _p(b))

    def interpolate(self, sp, a, b, t):
        a, b = _f64(a), _f64(b)
        out = np.empty_like(a)
        self.lib.orc_interpolate(C.byref(sp), _p(a), _p(b), t, _p(out))
        return out

    def rv_max_extent(self, lo, hi):
        lo, hi = _f64(lo), _f64(hi)
        return self.lib.orc_rv_max_extent(_p(lo), _p(hi), lo.shape[0])

    def resolution(self, extent, fraction=0.05):
        return self.lib.orc_motion_resolution(
            self.lib.orc_longest_valid_segment_length(extent, fraction))

    # -- batches -------------------------------------------------------------------
    def is_valid(self, w, sp, states):
        st = _f64(np.atleast_2d(states))
        out = np.empty(st.shape[0], dtype=np.uint8)
        self.lib.orc_is_valid_batch(C.byref(w), C.byref(sp), _p(st), st.shape[0], _p(out))
        return out

    def check_motion(self, w, sp, frm, to, resolution, counters=None):
        a, b = _f64(np.atleast_2d(frm)), _f64(np.atleast_2d(to))
        E = a.shape[0]
        valid = np.empty(E, dtype=np.uint8)
        first = np.empty(E, dtype=np.uint32)
        self.lib.orc_check_motion_batch(C.byref(w), C.byref(sp), _p(a), _p(b), E, resolution,
                                        _p(valid), _p(first),
                                        C.byref(counters) if counters is not None else None)
        return valid, first

    def nn1(self, sp, rows, queries):
        rows, q = _f64(rows), _f64(np.atleast_2d(queries))
        idx = np.empty(q.shape[0], dtype=np.uint32)
        dist = np.empty(q.shape[0], dtype=np.float64)
        self.lib.orc_nn1_batch(C.byref(sp), _p(rows), rows.shape[0], _p(q), q.shape[0], _p(idx),
                               _p(dist))
        return idx, dist

    def knn(self, sp, rows, queries, k, exclude=None):
        rows, q = _f64(rows), _f64(np.atleast_2d(queries))
        idx = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float64)
        ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint32)
        self.lib.orc_knn_batch(C.byref(sp), _p(rows), rows.shape[0], _p(q), q.shape[0], k, _p(ex),
                               _p(idx), _p(dist))
        return idx, dist

    def radius(self, sp, rows, queries, r, idx_limit=None, want_dist=False):
        rows, q = _f64(rows), _f64(np.atleast_2d(queries))
        N = rows.shape[0]
        offsets = np.zeros(q.shape[0] + 1, dtype=np.uint64)
        idxs, dists = [], []
        buf_i = np.empty(max(N, 1), dtype=np.uint32)
        buf_d = np.empty(max(N, 1), dtype=np.float64)
        for i in range(q.shape[0]):
            lim = N if idx_limit is None else int(idx_limit[i])
            qi = np.ascontiguousarray(q[i])
            n = self.lib.orc_radius(C.byref(sp), _p(rows), N, _p(qi), r, lim, _p(buf_i), _p(buf_d),
                                    buf_i.shape[0])
            idxs.append(buf_i[:n].copy())
            dists.append(buf_d[:n].copy())
            offsets[i + 1] = offsets[i] + n
        idx = np.concatenate(idxs) if idxs else np.empty(0, np.uint32)
        dist = np.concatenate(dists) if dists else np.empty(0, np.float64)
        return (offsets, idx, dist) if want_dist else (offsets, idx)

    # -- SO(3) ----------------------------------------------------------------------
    def so3_distance(self, a, b):
        a, b = _f64(a), _f64(b)
        return self.lib.orc_so3_distance(_p(a), _p(b))

    def so3_interpolate(self, a, b, t):
        a, b = _f64(a), _f64(b)
        out = np.empty(4, dtype=np.float64)
        self.lib.orc_so3_interpolate(_p(a), _p(b), float(t), _p(out))
        return out

    def so3_nn1(self, rows, queries):
        rows, q = _f64(rows), _f64(np.atleast_2d(queries))
        idx = np.empty(q.shape[0], dtype=np.uint32)
        dist = np.empty(q.shape[0], dtype=np.float64)
        self.lib.orc_so3_nn1_batch(_p(rows), rows.shape[0], _p(q), q.shape[0], _p(idx), _p(dist))
        return idx, dist

    def so3_knn(self, rows, queries, k, exclude=None):
        rows, q = _f64(rows), _f64(np.atleast_2d(queries))
        idx = np.empty((q.shape[0], k), dtype=np.uint32)
        dist = np.empty((q.shape[0], k), dtype=np.float64)
        ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.uint32)
        self.lib.orc_so3_knn_batch(_p(rows), rows.shape[0], _p(q), q.shape[0], k, _p(ex), _p(idx),
                                   _p(dist))
        return idx, dist

    def so3_radius(self, rows, queries, r):
        rows, q = _f64(rows), _f64(np.atleast_2d(queries))
        N = rows.shape[0]
        offsets = np.zeros(q.shape[0] + 1, dtype=np.uint64)
        idxs, dists = [], []
        buf_i = np.empty(max(N, 1), dtype=np.uint32)
        buf_d = np.empty(max(N, 1), dtype=np.float64)
        for i in range(q.shape[0]):
            qi = np.ascontiguousarray(q[i])
            n = self.lib.orc_so3_radius(_p(rows), N, _p(qi), float(r), _p(buf_i), _p(buf_d), N)
            idxs.append(buf_i[:n].copy())
            dists.append(buf_d[:n].copy())
            offsets[i + 1] = offsets[i] + n
        idx = np.concatenate(idxs) if idxs else np.empty(0, np.uint32)
        dist = np.concatenate(dists) if dists else np.empty(0, np.float64)
        return offsets, idx, dist

    def so3_rrt(self, cones, start, samples, max_distance, resolution, max_nodes, goal_target=None,
                goal_radius=-1.0):
        """RRT::solve over SO3 -> (tree rows, parents, solved, iters)."""
        cones = _f64(np.reshape(cones, (-1, 5)))
        s = _f64(np.atleast_2d(samples))
        tree = np.zeros((max_nodes, 4), dtype=np.float64)
        tree[0] = start
        parents = np.full(max_nodes, -1, dtype=np.int64)
        g = None if goal_target is None else _f64(goal_target)
        solved, iters = C.c_int(0), C.c_uint64(0)
        n = self.lib.orc_so3_rrt(_p(cones), cones.shape[0], _p(g), float(goal_radius),
                                 float(max_distance), float(resolution), _p(s), s.shape[0],
                                 int(max_nodes), _p(tree), _p(parents), C.byref(solved),
                                 C.byref(iters))
        return tree[:n].copy(), parents[:n].copy(), bool(solved.value), int(iters.value)

    def so3_rrt_connect(self, cones, start, goal_root, samples, max_distance, resolution, max_nodes,
                        goal_target=None, goal_radius=-1.0):
        """RRTConnect::solve over SO3 -> dict(start_tree, start_parents, goal_tree, goal_parents,
        solved, conn, iters)."""
        cones = _f64(np.reshape(cones, (-1, 5)))
        s = _f64(np.atleast_2d(samples))
        ta = np.zeros((max_nodes, 4), dtype=np.float64)
        tb = np.zeros((max_nodes, 4), dtype=np.float64)
        ta[0], tb[0] = start, goal_root
        pa = np.full(max_nodes, -1, dtype=np.int64)
        pb = np.full(max_nodes, -1, dtype=np.int64)
        g = None if goal_target is None else _f64(goal_target)
        na, nb, iters = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        cs, cg = C.c_int64(-1), C.c_int64(-1)
        solved = self.lib.orc_so3_rrt_connect(_p(cones), cones.shape[0], _p(g), float(goal_radius),
                                              float(max_distance), float(resolution), _p(s),
                                              s.shape[0], int(max_nodes), _p(ta), _p(pa),
                                              C.byref(na), _p(tb), _p(pb), C.byref(nb),
                                              C.byref(cs), C.byref(cg), C.byref(iters))
        return {"start_tree": ta[:na.value].copy(), "start_parents": pa[:na.value].copy(),
                "goal_tree": tb[:nb.value].copy(), "goal_parents": pb[:nb.value].copy(),
                "solved": bool(solved), "conn": (cs.value, cg.value), "iters": int(iters.value)}

    def so3_is_valid(self, cones, states):
        cones, st = _f64(np.reshape(cones, (-1, 5))), _f64(np.atleast_2d(states))
        return np.array([self.lib.orc_so3_is_valid(_p(cones), cones.shape[0],
                                                   _p(np.ascontiguousarray(st[i])))
                         for i in range(st.shape[0])], dtype=np.uint8)

    def so3_check_motion(self, cones, frm, to, resolution):
        cones = _f64(np.reshape(cones, (-1, 5)))
        frm, to = _f64(np.atleast_2d(frm)), _f64(np.atleast_2d(to))
        valid = np.empty(frm.shape[0], dtype=np.uint8)
        first = np.empty(frm.shape[0], dtype=np.uint32)
        for e in range(frm.shape[0]):
            fi = C.c_uint32(0)
            v = self.lib.orc_so3_check_motion(_p(cones), cones.shape[0],
                                              _p(np.ascontiguousarray(frm[e])),
                                              _p(np.ascontiguousarray(to[e])), float(resolution),
                                              C.byref(fi))
            valid[e] = v == 1
            first[e] = fi.value
        return valid, first

    def prm_knn(self, w, sp, rows, k, resolution, row_begin=0, row_end=None, counters=None):
        rows = _f64(rows)
        N = rows.shape[0]
        if row_end is None:
            row_end = N
        n = row_end - row_begin
        idx = np.empty((n, k), dtype=np.uint32)
        dist = np.empty((n, k), dtype=np.float64)
        valid = np.empty((n, k), dtype=np.uint8)
        self.lib.orc_prm_knn(C.byref(w), C.byref(sp), _p(rows), N, k, resolution, row_begin, row_end,
                             _p(idx), _p(dist), _p(valid),
                             C.byref(counters) if counters is not None else None)
        return idx, dist, valid

    def prm_radius(self, w, sp, radius, resolution, samples):
        s = _f64(np.atleast_2d(samples))
        n_s, dim = s.shape
        nodes = np.empty((n_s, dim), dtype=np.float64)
        offsets = np.zeros(n_s + 1, dtype=np.uint64)
        cap = max(1024, 64 * n_s)
        pairs, mcs = C.c_uint64(0), C.c_uint64(0)
        while True:
            idx = np.empty(cap, dtype=np.uint32)
            n = self.lib.orc_prm_radius(C.byref(w), C.byref(sp), radius, resolution, _p(s), n_s,
                                        _p(nodes), _p(offsets), _p(idx), cap, C.byref(pairs),
                                        C.byref(mcs))
            if int(offsets[n]) > cap:
                cap = int(offsets[n])
                continue
            break
        return (nodes[:n].copy(), offsets[: n + 1].copy(), idx[: int(offsets[n])].copy(),
                pairs.value, mcs.value)

    def prm_solve(self, w, sp, rows, lower_off, lower_idx, connection_radius, resolution, start,
                  goal_target, goal_radius):
        """PRM::solve (prm.rs:227-307) -> (status, path node indices)."""
        rows = _f64(rows)
        off = np.ascontiguousarray(lower_off, dtype=np.uint64)
        idx = np.ascontiguousarray(lower_idx, dtype=np.uint32)
        st, tg = _f64(start), _f64(goal_target)
        g = OrcGoal(C.c_void_p(tg.ctypes.data), float(goal_radius))
        path = np.empty(max(rows.shape[0], 1), dtype=np.uint32)
        plen = C.c_uint32(0)
        status = self.lib.orc_prm_solve(C.byref(w), C.byref(sp), _p(rows), rows.shape[0], _p(off),
                                        _p(idx), connection_radius, resolution, _p(st), C.byref(g),
                                        _p(path), path.shape[0], C.byref(plen))
        return int(status), path[: plen.value].copy()

    def rrt(self, w, sp, start, goal_target, goal_radius, max_distance, resolution, samples,
            max_nodes):
        s = _f64(np.atleast_2d(samples))
        dim = s.shape[1]
        tree = np.zeros((max_nodes, dim), dtype=np.float64)
        tree[0] = start
        parents = np.full(max_nodes, -1, dtype=np.int64)
        tgt = _f64(goal_target) if goal_target is not None else None
        goal = OrcGoal(_p(tgt), goal_radius) if tgt is not None else None
        solved, iters = C.c_int(0), C.c_uint64(0)
        n = self.lib.orc_rrt(C.byref(w), C.byref(sp), C.byref(goal) if goal else None, max_distance,
                             resolution, _p(s), s.shape[0], max_nodes, _p(tree), _p(parents),
                             C.byref(solved), C.byref(iters))
        return tree[:n].copy(), parents[:n].copy(), bool(solved.value), iters.value

    def rrt_connect(self, w, sp, start, goal_state, goal_target, goal_radius, max_distance,
                    resolution, samples, max_nodes):
        s = _f64(np.atleast_2d(samples))
        dim = s.shape[1]
        ta = np.zeros((max_nodes, dim)); ta[0] = start
        tb = np.zeros((max_nodes, dim)); tb[0] = goal_state
        pa = np.full(max_nodes, -1, dtype=np.int64)
        pb = np.full(max_nodes, -1, dtype=np.int64)
        tgt = _f64(goal_target)
        goal = OrcGoal(_p(tgt), goal_radius)
        na, nb, it = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        cs, cg = C.c_int64(-1), C.c_int64(-1)
        solved = self.lib.orc_rrt_connect(C.byref(w), C.byref(sp), C.byref(goal), max_distance,
                                          resolution, _p(s), s.shape[0], max_nodes, _p(ta), _p(pa),
                                          C.byref(na), _p(tb), _p(pb), C.byref(nb), C.byref(cs),
                                          C.byref(cg), C.byref(it))
        return dict(start_tree=ta[: na.value].copy(), start_parents=pa[: na.value].copy(),
                    goal_tree=tb[: nb.value].copy(), goal_parents=pb[: nb.value].copy(),
                    solved=bool(solved), conn_start=cs.value, conn_goal=cg.value, iters=it.value)

    def rrt_star(self, w, sp, start, goal_target, goal_radius, max_distance, search_radius,
                 resolution, samples, max_nodes, stop_at_goal=True):
        s = _f64(np.atleast_2d(samples))
        dim = s.shape[1]
        tree = np.zeros((max_nodes, dim)); tree[0] = start
        parents = np.full(max_nodes, -1, dtype=np.int64)
        costs = np.zeros(max_nodes)
        tgt = _f64(goal_target) if goal_target is not None else None
        goal = OrcGoal(_p(tgt), goal_radius) if tgt is not None else None
        solved, iters, mcs = C.c_int(0), C.c_uint64(0), C.c_uint64(0)
        n = self.lib.orc_rrt_star(C.byref(w), C.byref(sp), C.byref(goal) if goal else None,
                                  max_distance, search_radius, resolution, _p(s), s.shape[0],
                                  max_nodes, 1 if stop_at_goal else 0, _p(tree), _p(parents),
                                  _p(costs), C.byref(solved), C.byref(iters), C.byref(mcs))
        return dict(tree=tree[:n].copy(), parents=parents[:n].copy(), costs=costs[:n].copy(),
                    solved=bool(solved.value), iters=iters.value, motion_checks=mcs.value)


_ORACLE = None


def get() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
