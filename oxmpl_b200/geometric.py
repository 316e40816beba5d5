"""Mirror of ``oxmpl_py.geometric`` on the GPU path: ``RRT``, ``RRTConnect``, ``RRTStar``,
``PRM`` with the constructors and methods of oxmpl-py/src/geometric/*.rs
(``RRT(max_distance, goal_bias, problem_definition)``, ``setup(validity)``,
``solve(timeout_secs)`` -> ``Path``; ``PRM(timeout, connection_radius, problem_definition)``,
``construct_roadmap()``).  ``setup`` takes a :class:`oxmpl_b200.base.GpuValidityChecker`
instead of a Python callback — the point of the GPU path is that validity never re-enters the
interpreter; a plain callable is rejected loudly (there is no CPU fallback)."""
from __future__ import annotations

import random
import time
from collections import deque

import numpy as np

from . import api
from .base import (GpuConeChecker, GpuIntervalChecker, GpuValidityChecker, Path, PlanningError,
                   ProblemDefinition, RealVectorState, SO2State, SO2StateSpace, SO3State, SO3StateSpace)


def _checker(validity, space=None):
    want = GpuIntervalChecker if isinstance(space, SO2StateSpace) else GpuValidityChecker
    if not isinstance(validity, want):
        raise TypeError(f"the GPU planners take a declarative {want.__name__}, not a Python "
                        "callback (oxmpl_b200 has no CPU fallback)")
    return validity


def _state_maker(space):
    """Row of the store -> the space's state type."""
    if isinstance(space, SO2StateSpace):
        return lambda r: SO2State(r[0])
    if isinstance(space, SO3StateSpace):
        return lambda r: SO3State(*r)
    return RealVectorState


class _TreePlanner:
    chunk = 1 << 14
    max_nodes = 1 << 20

    def __init__(self, max_distance: float, goal_bias: float, problem_definition: ProblemDefinition,
                 seed: int = 0):
        self.max_distance = float(max_distance)
        self.goal_bias = float(goal_bias)
        self.pd = problem_definition
        self.rng = random.Random(seed)
        self.vc = None
        self.tree = None

    def setup(self, validity):
        self.vc = _checker(validity, self.pd.space)
        self.tree = api.StateStore(self.pd.space._space(), capacity=1024)
        self.tree.append(np.asarray([self.pd.start_state.values]))   # rrt.rs:147-155

    def _stream(self, n):
        # the reference's per-iteration draw (rrt.rs:177-184), n iterations ahead
        pd, out = self.pd, np.empty((n, self.pd.space.dimension))
        for i in range(n):
            q = pd.goal.sample_goal() if self.rng.random() < self.goal_bias \
                else pd.space.sample_uniform(self.rng)
            out[i] = q.values
        return out

    def _resolution(self):
        return self.pd.space.get_longest_valid_segment_length() * 0.1   # rrt.rs:97

    def _trace(self, store, parents, node):
        rows, out, mk = store.read(), [], _state_maker(self.pd.space)
        while node >= 0:
            out.append(mk(rows[node]))
            node = int(parents[node])
        return out[::-1]


class _So3Tree:
    """A tree of SO3 states: device store (for the scans) + host copies (for steering / paths)."""

    def __init__(self, space, root: SO3State):
        self.store = api.StateStore(space._space(), capacity=1024)
        self.store.append(np.asarray([root.values]))
        self.nodes, self.parents, self.costs = [root], [-1], [0.0]

    def nearest(self, q: SO3State):
        idx, dist = self.store.so3_nn1(np.asarray([q.values]))
        return int(idx[0]), float(dist[0])

    def push(self, q: SO3State, parent: int, cost: float = 0.0) -> int:
        self.store.append(np.asarray([q.values]))
        self.nodes.append(q)
        self.parents.append(parent)
        self.costs.append(cost)
        return len(self.nodes) - 1

    def path_to(self, node: int):
        out = []
        while node >= 0:
            out.append(self.nodes[node])
            node = self.parents[node]
        return out[::-1]


def _so3_checker(validity) -> GpuConeChecker:
    if not isinstance(validity, GpuConeChecker):
        raise TypeError("SO3 planning takes a declarative GpuConeChecker, not a Python "
                        "callback (oxmpl_b200 has no CPU fallback)")
    return validity


def _so3_steer(space, q_near: SO3State, q_rand: SO3State, min_dist: float, max_distance: float):
    """rrt.rs:200-208 — returns (q_new, reached)."""
    if min_dist > max_distance:
        return space.interpolate(q_near, q_rand, max_distance / min_dist), False
    return q_rand, True


def _so3_motion(vc: GpuConeChecker, frm, to, res) -> np.ndarray:
    a = np.asarray([s.values for s in frm]).reshape(-1, 4)
    b = np.asarray([s.values for s in to]).reshape(-1, 4)
    return vc.cones.check_motion(a, b, res).astype(bool) if len(a) else np.empty(0, bool)


class RRT(_TreePlanner):
    """oxmpl_py.geometric.RRT (oxmpl-py/src/geometric/rrt.rs) over oxg_rrt_solve.  Over an
    :class:`SO3StateSpace` it runs the reference loop unfused: ``oxg_so3_nn1`` -> steer on the
    host (SLERP needs libm) -> ``oxg_so3_check_motion`` -> append (include/oxg_so3.h)."""

    def setup(self, validity):
        if not isinstance(self.pd.space, SO3StateSpace):
            return super().setup(validity)
        if not isinstance(validity, GpuConeChecker):
            raise TypeError("SO3 planning takes a declarative GpuConeChecker, not a Python "
                            "callback (oxmpl_b200 has no CPU fallback)")
        self.vc = validity
        self.tree = api.StateStore(self.pd.space._space(), capacity=1024)
        self.tree.append(np.asarray([self.pd.start_state.values]))
        self.nodes = [self.pd.start_state]
        self.parents = [-1]

    so3_chunk = 1024

    def _solve_so3_chunked(self, timeout_secs: float) -> Path:
        """Ball goal: ``oxg_so3_rrt_solve`` consumes a pre-drawn stream in speculative chunks (one
        1-NN launch + one cone launch per chunk instead of per iteration, include/oxg_so3.h)."""
        pd, res = self.pd, self._resolution()
        par = np.asarray(self.parents, dtype=np.int64)
        t0 = time.perf_counter()
        while True:
            left = timeout_secs - (time.perf_counter() - t0)
            if left <= 0:
                raise PlanningError(PlanningError.TIMEOUT)
            stream = np.empty((self.so3_chunk, 4))
            for i in range(self.so3_chunk):                        # rrt.rs:177-184, ahead of time
                q = pd.goal.sample_goal() if self.rng.random() < self.goal_bias \
                    else pd.space.sample_uniform(self.rng)
                stream[i] = q.values
            par, r = api.so3_rrt_solve(self.vc.cones, self.tree, stream, self.max_distance, res,
                                       self.max_nodes, goal_target=pd.goal.target.values,
                                       goal_radius=pd.goal.radius, timeout_s=left, parents=par)
            for row in self.tree.read(len(self.nodes)):
                self.nodes.append(SO3State(*row))
            self.parents = [int(v) for v in par]
            self.last_result = r
            if r["solved"]:
                out, node = [], r["goal_node"][0]
                while node >= 0:
                    out.append(self.nodes[node])
                    node = self.parents[node]
                return Path(out[::-1])
            if r["status"] in ("timeout", "tree_full"):
                raise PlanningError(PlanningError.TIMEOUT)

    def _solve_so3(self, timeout_secs: float) -> Path:
        pd, space = self.pd, self.pd.space
        if hasattr(pd.goal, "target") and hasattr(pd.goal, "radius"):
            return self._solve_so3_chunked(timeout_secs)
        res = self._resolution()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 <= timeout_secs:          # rrt.rs:170-225
            q_rand = pd.goal.sample_goal() if self.rng.random() < self.goal_bias \
                else space.sample_uniform(self.rng)
            idx, dist = self.tree.so3_nn1(np.asarray([q_rand.values]))
            near, min_dist = int(idx[0]), float(dist[0])
            q_near = self.nodes[near]
            if min_dist > self.max_distance:
                q_new = space.interpolate(q_near, q_rand, self.max_distance / min_dist)
            else:
                q_new = q_rand
            if self.vc.cones.check_motion(np.asarray([q_near.values]), np.asarray([q_new.values]),
                                          res)[0]:
                self.tree.append(np.asarray([q_new.values]))
                self.nodes.append(q_new)
                self.parents.append(near)
                if pd.goal.is_satisfied(q_new):
                    out, node = [], len(self.nodes) - 1
                    while node >= 0:
                        out.append(self.nodes[node])
                        node = self.parents[node]
                    return Path(out[::-1])
        raise PlanningError(PlanningError.TIMEOUT)

    def _solve_unfused(self, timeout_secs: float) -> Path:
        """rrt.rs:170-225 one iteration at a time, for goals the device cannot test (anything but a
        ball): oxg_nn1 -> steer on the host (rrt.rs:200-208) -> oxg_check_motion -> append -> the
        user's ``is_satisfied``."""
        pd, space = self.pd, self.pd.space
        sp, res = space._space(), self._resolution()
        nodes, parents = [pd.start_state], [-1]
        t0 = time.perf_counter()
        while time.perf_counter() - t0 <= timeout_secs:
            q_rand = pd.goal.sample_goal() if self.rng.random() < self.goal_bias \
                else space.sample_uniform(self.rng)
            idx, dist = self.tree.nn1(np.asarray([q_rand.values]))
            near, min_dist = int(idx[0]), float(dist[0])
            q_near = nodes[near]
            q_new = space.interpolate(q_near, q_rand, self.max_distance / min_dist) \
                if min_dist > self.max_distance else q_rand
            if self.vc.world.check_motion(sp, np.asarray([q_near.values]), np.asarray([q_new.values]),
                                          res)[0]:
                self.tree.append(np.asarray([q_new.values]))
                nodes.append(q_new)
                parents.append(near)
                if pd.goal.is_satisfied(q_new):
                    out, node = [], len(nodes) - 1
                    while node >= 0:
                        out.append(nodes[node])
                        node = parents[node]
                    return Path(out[::-1])
        raise PlanningError(PlanningError.TIMEOUT)

    def solve(self, timeout_secs: float) -> Path:
        if self.vc is None:
            raise PlanningError(PlanningError.UNINITIALISED)
        if isinstance(self.pd.space, SO3StateSpace):
            return self._solve_so3(timeout_secs)
        if not (hasattr(self.pd.goal, "target") and hasattr(self.pd.goal, "radius")):
            return self._solve_unfused(timeout_secs)
        t0 = time.perf_counter()
        while True:
            left = timeout_secs - (time.perf_counter() - t0)
            if left <= 0:
                raise PlanningError(PlanningError.TIMEOUT)
            parents, r = api.rrt_solve(self.vc.world, self.tree, self._stream(self.chunk),
                                       self.max_distance, self._resolution(), self.max_nodes,
                                       goal_target=self.pd.goal.target.values,
                                       goal_radius=self.pd.goal.radius, timeout_s=left)
            if r["solved"]:
                return Path(self._trace(self.tree, parents, r["goal_node"][0]))
            if r["status"] in ("timeout", "tree_full"):
                raise PlanningError(PlanningError.TIMEOUT)


class RRTStar(_TreePlanner):
    """oxmpl_py.geometric.RRTStar(max_distance, goal_bias, search_radius, problem_definition).
    Over an :class:`SO3StateSpace` the reference loop (rrt_star.rs:182-289) runs unfused on
    ``oxg_so3_nn1`` / ``oxg_so3_radius`` / ``oxg_so3_check_motion``: the choose-parent and rewire
    candidates are motion-checked in one batch each and the strict-`<` selection is replayed in
    neighbour order on the host."""

    def __init__(self, max_distance, goal_bias, search_radius, problem_definition, seed: int = 0):
        super().__init__(max_distance, goal_bias, problem_definition, seed)
        self.search_radius = float(search_radius)

    def setup(self, validity):
        if not isinstance(self.pd.space, SO3StateSpace):
            return super().setup(validity)
        self.vc = _so3_checker(validity)
        self.t = _So3Tree(self.pd.space, self.pd.start_state)

    def _solve_so3(self, timeout_secs: float) -> Path:
        pd, space, t, res = self.pd, self.pd.space, self.t, self._resolution()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 <= timeout_secs:
            q_rand = pd.goal.sample_goal() if self.rng.random() < self.goal_bias \
                else space.sample_uniform(self.rng)
            near, min_dist = t.nearest(q_rand)
            q_new, _ = _so3_steer(space, t.nodes[near], q_rand, min_dist, self.max_distance)
            if not _so3_motion(self.vc, [t.nodes[near]], [q_new], res)[0]:
                continue
            q = np.asarray([q_new.values])
            _, nbrs, nd = t.store.so3_radius(q, self.search_radius, want_dist=True)  # find_neighbours
            nbrs = [int(j) for j in nbrs]
            # choose parent (rrt_star.rs:233-249); cost(new, nbr) = nbr.cost + distance(new, nbr)
            best = near
            min_cost = t.costs[near] + space.distance(q_new, t.nodes[near])
            via = [t.costs[j] + float(d) for j, d in zip(nbrs, nd)]
            cand = [u for u, c in enumerate(via) if c < min_cost]
            ok = _so3_motion(self.vc, [t.nodes[nbrs[u]] for u in cand], [q_new] * len(cand), res)
            for u, good in zip(cand, ok):
                if via[u] < min_cost and good:
                    min_cost, best = via[u], nbrs[u]
            new = t.push(q_new, best, min_cost)
            # rewire (rrt_star.rs:261-282); cost(nbr, new) = new.cost + distance(nbr, new)
            thru = [min_cost + float(d) for d in nd]
            cand = [u for u, j in enumerate(nbrs) if j != best and thru[u] < t.costs[j]]
            ok = _so3_motion(self.vc, [q_new] * len(cand), [t.nodes[nbrs[u]] for u in cand], res)
            for u, good in zip(cand, ok):
                if good:
                    t.parents[nbrs[u]], t.costs[nbrs[u]] = new, thru[u]
            if pd.goal.is_satisfied(q_new):
                return Path(t.path_to(new))
        raise PlanningError(PlanningError.TIMEOUT)

    def solve(self, timeout_secs: float) -> Path:
        if self.vc is None:
            raise PlanningError(PlanningError.UNINITIALISED)
        if isinstance(self.pd.space, SO3StateSpace):
            return self._solve_so3(timeout_secs)
        t0 = time.perf_counter()
        while True:
            left = timeout_secs - (time.perf_counter() - t0)
            if left <= 0:
                raise PlanningError(PlanningError.TIMEOUT)
            if not (hasattr(self.pd.goal, "target") and hasattr(self.pd.goal, "radius")):
                raise TypeError("the fused RRT* loop tests a ball goal (target + radius); use RRT for "
                                "other goal objects")
            try:
                parents, costs, r = api.rrt_star_solve(
                    self.vc.world, self.tree, self._stream(self.chunk), self.max_distance,
                    self.search_radius, self._resolution(), self.max_nodes,
                    goal_target=self.pd.goal.target.values, goal_radius=self.pd.goal.radius,
                    stop_at_goal=True, timeout_s=left)
            except api._lib.OxgError as e:
                if e.code == api._lib.OXG_E_CAPACITY:   # OXG_RRT_STAR_MAX_NEIGHBOURS (include/oxg.h)
                    raise PlanningError(PlanningError.NO_SOLUTION + " " + str(e)) from e
                raise
            if r["solved"]:
                return Path(self._trace(self.tree, parents, r["goal_node"][0]))
            if r["status"] in ("timeout", "tree_full"):
                raise PlanningError(PlanningError.TIMEOUT)


class RRTConnect(_TreePlanner):
    """oxmpl_py.geometric.RRTConnect over oxg_rrt_connect_solve; over an :class:`SO3StateSpace` the
    reference loop (rrt_connect.rs:121-159,241-308) runs unfused on the SO3 entry points."""

    def _extend_so3(self, t: "_So3Tree", target: SO3State):
        """RRTConnect::extend — (reached, new index) or None when the motion is invalid."""
        near, min_dist = t.nearest(target)
        q_new, reached = _so3_steer(self.pd.space, t.nodes[near], target, min_dist, self.max_distance)
        if not _so3_motion(self.vc, [t.nodes[near]], [q_new], self._resolution())[0]:
            return None
        return reached, t.push(q_new, near)

    so3_chunk = 1024

    def _solve_so3_chunked(self, timeout_secs: float) -> Path:
        """Ball goal: ``oxg_so3_rrt_connect_solve`` on a pre-drawn stream (speculative chunks, six
        launches per chunk instead of four synchronous calls per iteration)."""
        pd, res = self.pd, self._resolution()
        pa = np.asarray(self.ta.parents, dtype=np.int64)
        pb = np.asarray(self.tb.parents, dtype=np.int64)
        t0 = time.perf_counter()
        while True:
            left = timeout_secs - (time.perf_counter() - t0)
            if left <= 0:
                raise PlanningError(PlanningError.TIMEOUT)
            stream = np.empty((self.so3_chunk, 4))
            for i in range(self.so3_chunk):                        # rrt_connect.rs:256-262
                q = pd.goal.sample_goal() if self.rng.random() < self.goal_bias \
                    else pd.space.sample_uniform(self.rng)
                stream[i] = q.values
            pa, pb, r = api.so3_rrt_connect_solve(
                self.vc.cones, self.ta.store, self.tb.store, stream, self.max_distance, res,
                self.max_nodes, goal_target=pd.goal.target.values, goal_radius=pd.goal.radius,
                timeout_s=left, start_parents=pa, goal_parents=pb)
            for t, par in ((self.ta, pa), (self.tb, pb)):
                for row in t.store.read(len(t.nodes)):
                    t.nodes.append(SO3State(*row))
                    t.costs.append(0.0)
                t.parents = [int(v) for v in par]
            self.last_result = r
            if r["solved"]:
                s_idx, g_idx = r["goal_node"]
                if g_idx < 0:
                    return Path(self.ta.path_to(s_idx))
                return Path(self.ta.path_to(s_idx) + self.tb.path_to(g_idx)[::-1][1:])
            if r["status"] in ("timeout", "tree_full"):
                raise PlanningError(PlanningError.TIMEOUT)

    def _solve_so3(self, timeout_secs: float) -> Path:
        pd, space = self.pd, self.pd.space
        if hasattr(pd.goal, "target") and hasattr(pd.goal, "radius"):
            return self._solve_so3_chunked(timeout_secs)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 <= timeout_secs:
            grow_start = len(self.ta.nodes) <= len(self.tb.nodes)          # rrt_connect.rs:249-254
            a, b = (self.ta, self.tb) if grow_start else (self.tb, self.ta)
            q_rand = pd.goal.sample_goal() if self.rng.random() < self.goal_bias \
                else space.sample_uniform(self.rng)
            ext = self._extend_so3(a, q_rand)
            if ext is None:
                continue
            q_new = a.nodes[ext[1]]
            if grow_start and pd.goal.is_satisfied(q_new):
                return Path(self.ta.path_to(ext[1]))
            con = self._extend_so3(b, q_new)
            if con is not None and con[0]:
                s_idx, g_idx = (ext[1], con[1]) if grow_start else (con[1], ext[1])
                return Path(self.ta.path_to(s_idx) + self.tb.path_to(g_idx)[::-1][1:])
        raise PlanningError(PlanningError.TIMEOUT)

    def setup(self, validity):
        if isinstance(self.pd.space, SO3StateSpace):
            self.vc = _so3_checker(validity)
            self.ta = _So3Tree(self.pd.space, self.pd.start_state)
            self.tb = _So3Tree(self.pd.space, self.pd.goal.sample_goal())   # rrt_connect.rs:207-214
            return
        super().setup(validity)
        self.goal_tree = api.StateStore(self.pd.space._space(), capacity=1024)
        self.goal_tree.append(np.asarray([self.pd.goal.sample_goal().values]))  # rrt_connect.rs:207-214

    def solve(self, timeout_secs: float) -> Path:
        if self.vc is None:
            raise PlanningError(PlanningError.UNINITIALISED)
        if isinstance(self.pd.space, SO3StateSpace):
            return self._solve_so3(timeout_secs)
        t0 = time.perf_counter()
        while True:
            left = timeout_secs - (time.perf_counter() - t0)
            if left <= 0:
                raise PlanningError(PlanningError.TIMEOUT)
            pa, pb, r = api.rrt_connect_solve(
                self.vc.world, self.tree, self.goal_tree, self._stream(self.chunk),
                self.max_distance, self._resolution(), self.max_nodes,
                goal_target=self.pd.goal.target.values, goal_radius=self.pd.goal.radius,
                timeout_s=left)
            if r["solved"]:
                states = self._trace(self.tree, pa, r["goal_node"][0])
                if r["goal_node"][1] >= 0:     # merge (rrt_connect.rs:288-304)
                    states += self._trace(self.goal_tree, pb, r["goal_node"][1])[::-1][1:]
                return Path(states)
            if r["status"] in ("timeout", "tree_full"):
                raise PlanningError(PlanningError.TIMEOUT)


class PRM:
    """oxmpl_py.geometric.PRM(timeout, connection_radius, problem_definition)
    (oxmpl-py/src/geometric/prm.rs): construct_roadmap() via oxg_prm_build_radius in batches
    (identical to the incremental build: node i only looks at nodes j < i), solve() = start
    connection (radius scan + check_motion) + the reference's BFS (prm.rs:227-307)."""

    batch = 1 << 13

    def __init__(self, timeout: float, connection_radius: float, problem_definition: ProblemDefinition,
                 seed: int = 0):
        self.timeout = float(timeout)
        self.connection_radius = float(connection_radius)
        self.pd = problem_definition
        self.rng = random.Random(seed)
        self.vc = None
        self.edges = []

    def setup(self, validity):
        self.so3 = isinstance(self.pd.space, SO3StateSpace)
        if self.so3:
            if not isinstance(validity, GpuConeChecker):
                raise TypeError("SO3 planning takes a declarative GpuConeChecker, not a Python "
                                "callback (oxmpl_b200 has no CPU fallback)")
            self.vc = validity
        else:
            self.vc = _checker(validity, self.pd.space)
        self.store = api.StateStore(self.pd.space._space(), capacity=1024)
        self.edges = []                                       # prm.rs:224

    def get_roadmap(self):
        return self.edges

    def construct_roadmap(self, max_samples: int | None = None):
        if self.vc is None:
            raise PlanningError(PlanningError.UNINITIALISED)
        if self.edges:
            return                                            # prm.rs:106-113
        if self.so3:
            return self._construct_so3(max_samples)
        sp, res = self.pd.space, self.pd.space.get_longest_valid_segment_length() * 0.1
        t0, drawn = time.perf_counter(), 0
        while time.perf_counter() - t0 <= self.timeout and (max_samples is None or drawn < max_samples):
            nb = self.batch if max_samples is None else min(self.batch, max_samples - drawn)
            rows = np.array([sp.sample_uniform(self.rng).values for _ in range(nb)])
            drawn += nb
            keep = rows[self.vc.is_valid_batch(rows).astype(bool)]          # prm.rs:123
            if not len(keep):
                continue
            first = self.store.append(keep)
            off, idx, _ = api.prm_build_radius(self.vc.world, self.store, self.connection_radius,
                                               res, first, first + len(keep))
            self.edges.extend([] for _ in range(len(keep)))
            for i in range(len(keep)):
                for j in idx[int(off[i]):int(off[i + 1])]:
                    self.edges[first + i].append(int(j))                    # prm.rs:135
                    self.edges[int(j)].append(first + i)                    # prm.rs:143-145

    def _construct_so3(self, max_samples):
        """prm.rs:117-146 over SO3: node i joins every earlier node j with distance(q_i, q_j) <
        connection_radius whose check_motion(q_i, q_j) holds — batched: one radius scan and one
        motion-check call per batch of samples (include/oxg_so3.h)."""
        sp, res = self.pd.space, self.pd.space.get_longest_valid_segment_length() * 0.1
        t0, drawn = time.perf_counter(), 0
        self.rows = np.empty((0, 4))
        while time.perf_counter() - t0 <= self.timeout and (max_samples is None or drawn < max_samples):
            nb = min(self.batch, 2048) if max_samples is None else min(self.batch, 2048, max_samples - drawn)
            rows = np.array([sp.sample_uniform(self.rng).values for _ in range(nb)])
            drawn += nb
            keep = rows[self.vc.is_valid_batch(rows).astype(bool)]          # prm.rs:123
            if not len(keep):
                continue
            first = self.store.append(keep)
            self.rows = np.vstack([self.rows, keep])
            off, idx = self.store.so3_radius(keep, self.connection_radius)
            src = np.repeat(np.arange(first, first + len(keep)), np.diff(off).astype(np.int64))
            lower = idx < src                                               # only j < i (prm.rs:131)
            src, dst = src[lower], idx[lower].astype(np.int64)
            ok = self.vc.cones.check_motion(self.rows[src], self.rows[dst], res).astype(bool) \
                if len(src) else np.empty(0, bool)
            self.edges.extend([] for _ in range(len(keep)))
            for i, j in zip(src[ok], dst[ok]):
                self.edges[int(i)].append(int(j))                           # prm.rs:135
                self.edges[int(j)].append(int(i))                           # prm.rs:143-145

    def solve(self, timeout_secs: float) -> Path:
        if self.vc is None:
            raise PlanningError(PlanningError.UNINITIALISED)
        if not self.edges:
            raise PlanningError(PlanningError.UNSAMPLED)
        start = self.pd.start_state
        if not self.vc(start):
            raise PlanningError(PlanningError.INVALID_START)
        sp, res = self.pd.space, self.pd.space.get_longest_valid_segment_length() * 0.1
        rows = self.store.read()
        q = np.asarray([start.values])
        mk = _state_maker(self.pd.space)
        if self.so3:
            off, nb = self.store.so3_radius(q, self.connection_radius)      # prm.rs:249-256
            ok = self.vc.cones.check_motion(np.repeat(q, len(nb), axis=0), rows[nb], res) \
                if len(nb) else np.empty(0, np.uint8)
            return self._search(start, rows, nb, ok, mk)
        goal = self.pd.goal
        if hasattr(goal, "target") and hasattr(goal, "radius"):
            # ball goal: the whole query runs in the library (oxg_prm_solve: start connection, goal
            # scan, symmetrised adjacency and the reference-order BFS on the device)
            lower = [[j for j in e if j < i] for i, e in enumerate(self.edges)]
            off = np.zeros(len(lower) + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(e) for e in lower])
            idx = np.fromiter((j for e in lower for j in e), dtype=np.uint32, count=int(off[-1]))
            status, path, _ = api.prm_solve(self.vc.world, self.store, off, idx, self.connection_radius,
                                            res, start.values, goal.target.values, float(goal.radius))
            if status == 2:
                raise PlanningError(PlanningError.INVALID_START)
            if status != 0:
                raise PlanningError(PlanningError.NO_SOLUTION)
            return Path([start] + [mk(rows[i]) for i in path])
        off, nb = self.store.radius(q, self.connection_radius)              # prm.rs:249-256
        ok = self.vc.world.check_motion(sp._space(), np.repeat(q, len(nb), axis=0), rows[nb], res) \
            if len(nb) else np.empty(0, np.uint8)
        return self._search(start, rows, nb, ok, mk)

    def _search(self, start, rows, nb, ok, mk) -> Path:
        start_conn = [int(j) for j, v in zip(nb, ok) if v]
        goal = [self.pd.goal.is_satisfied(mk(r)) for r in rows]             # prm.rs:259-264
        if not start_conn or not any(goal):
            raise PlanningError(PlanningError.NO_SOLUTION)
        parent = {j: None for j in start_conn}
        queue = deque(start_conn)
        while queue:                                                          # prm.rs:271-301
            cur = queue.popleft()
            if goal[cur]:
                out = []
                while cur is not None:
                    out.append(mk(rows[cur]))
                    cur = parent[cur]
                return Path([start] + out[::-1])
            for j in self.edges[cur]:
                if j not in parent:
                    parent[j] = cur
                    queue.append(j)
        raise PlanningError(PlanningError.NO_SOLUTION)
