"""GPU: the reference's Python integration tests (oxmpl-py/tests/test_{rrt,rrt_connect,
rrt_star,prm}_rvss.py) against the `oxmpl_b200.base` / `oxmpl_b200.geometric` mirror: same
world, same planner parameters, same assertions; the Python `is_state_valid` callback is
replaced by the declarative GpuValidityChecker and re-used here as the independent check."""
import math
import random

import numpy as np
import pytest

from oxmpl_b200.base import (GpuValidityChecker, PlanningError, ProblemDefinition, RealVectorState,
                             RealVectorStateSpace)
from oxmpl_b200.geometric import PRM, RRT, RRTConnect, RRTStar

pytestmark = pytest.mark.gpu


class CircularGoal:                      # oxmpl-py/tests/test_rrt_rvss.py:9-26
    def __init__(self, space, x, y, radius):
        self.space = space
        self.target = RealVectorState([x, y])
        self.radius = radius
        self.rng = random.Random(123)

    def is_satisfied(self, state):
        return self.space.distance(self.target, state) <= self.radius

    def sample_goal(self):
        angle = self.rng.uniform(0, 2 * math.pi)
        radius = self.radius * math.sqrt(self.rng.uniform(0, 1))
        return RealVectorState([self.target.values[0] + radius * math.cos(angle),
                                self.target.values[1] + radius * math.sin(angle)])


def is_state_valid(state):               # oxmpl-py/tests/test_rrt_rvss.py:29-45
    x, y = state.values
    in_wall = (5.0 - 0.25 <= x <= 5.0 + 0.25) and (2.0 <= y <= 8.0)
    return not in_wall


def _problem():
    space = RealVectorStateSpace(dimension=2, bounds=[(0.0, 10.0), (0.0, 10.0)])
    start = RealVectorState([1.0, 5.0])
    goal = CircularGoal(space, x=9.0, y=5.0, radius=0.5)
    return space, start, goal, ProblemDefinition.from_real_vector(space, start, goal)


def _check(path, space, start, goal):
    assert len(path.states) > 1
    assert space.distance(path.states[0], start) < 1e-9
    assert goal.is_satisfied(path.states[-1])
    for s in path.states:
        assert is_state_valid(s), s.values


WALL = [[4.75, 2.0, 5.25, 8.0]]


def test_rrt_finds_path_in_rvss():
    space, start, goal, pd = _problem()
    planner = RRT(max_distance=0.5, goal_bias=0.05, problem_definition=pd)
    planner.setup(GpuValidityChecker(2, aabbs=WALL))
    _check(planner.solve(timeout_secs=5.0), space, start, goal)


def test_rrt_connect_finds_path_in_rvss():
    space, start, goal, pd = _problem()
    planner = RRTConnect(max_distance=0.5, goal_bias=0.05, problem_definition=pd)
    planner.setup(GpuValidityChecker(2, aabbs=WALL))
    _check(planner.solve(timeout_secs=5.0), space, start, goal)


def test_rrt_star_finds_path_in_rvss():
    space, start, goal, pd = _problem()
    planner = RRTStar(max_distance=0.5, goal_bias=0.05, search_radius=0.25, problem_definition=pd)
    planner.setup(GpuValidityChecker(2, aabbs=WALL))
    _check(planner.solve(timeout_secs=5.0), space, start, goal)


def test_prm_finds_path_in_rvss():
    space, start, goal, pd = _problem()
    planner = PRM(timeout=5.0, connection_radius=0.5, problem_definition=pd)
    planner.setup(GpuValidityChecker(2, aabbs=WALL))
    with pytest.raises(PlanningError, match="not sampled"):
        planner.solve(timeout_secs=5.0)
    planner.construct_roadmap(max_samples=12000)
    assert len(planner.get_roadmap()) > 5000
    _check(planner.solve(timeout_secs=5.0), space, start, goal)


def test_python_callback_is_rejected_and_checker_matches_callback():
    space, start, goal, pd = _problem()
    planner = RRT(max_distance=0.5, goal_bias=0.05, problem_definition=pd)
    with pytest.raises(TypeError):
        planner.setup(is_state_valid)
    vc = GpuValidityChecker(2, aabbs=WALL)
    rng = random.Random(1)
    for _ in range(200):
        s = space.sample_uniform(rng)
        assert vc(s) == is_state_valid(s)
    for xy in ([4.75, 2.0], [5.25, 8.0], [5.0, 5.0], [5.2500001, 8.0], [4.75, 1.9999999]):
        assert vc(RealVectorState(xy)) == is_state_valid(RealVectorState(xy))
    with pytest.raises(PlanningError, match="uninitialised"):
        RRT(0.5, 0.05, pd).solve(1.0)


# ---------------------------------------------------------------------------------------------
# SO(3): oxmpl-py/tests/test_rrt_so3ss.py against the mirror (unfused loop over include/oxg_so3.h)
# ---------------------------------------------------------------------------------------------
def _quat_axis_angle(axis, angle):
    from oxmpl_b200.base import SO3State
    s, c = math.sin(angle * 0.5), math.cos(angle * 0.5)
    return SO3State(x=axis[0] * s, y=axis[1] * s, z=axis[2] * s, w=c)


class _QuaternionGoalRegion:
    """oxmpl-py/tests/test_rrt_so3ss.py:14-33.  `valid_only` additionally rejects samples inside the
    forbidden cone: the upstream goal region (10 deg around a rotation 45 deg from the identity)
    lies mostly inside the 44.9 deg cone, so most draws are invalid states — harmless as RRT targets,
    but RRTConnect::setup roots its goal tree at one draw (rrt_connect.rs:207-214) and an invalid
    root can never be extended."""

    def __init__(self, space, target, radius, valid_only=False):
        self.space, self.target, self.radius = space, target, radius
        self.rng = random.Random(123)
        self.valid_only = valid_only

    def is_satisfied(self, state):
        return self.space.distance(self.target, state) <= self.radius

    def sample_goal(self):
        from oxmpl_b200.base import SO3State
        while True:
            x, y, z, w = [self.rng.uniform(-1.0, 1.0) for _ in range(4)]
            norm_sq = x * x + y * y + z * z + w * w
            if 1e-9 < norm_sq < 1.0:
                norm = math.sqrt(norm_sq)
                q = SO3State(x / norm, y / norm, z / norm, w / norm)
                if self.is_satisfied(q) and not (
                        self.valid_only
                        and self.space.distance(SO3State.identity(), q) <= math.radians(44.9)):
                    return q


def _so3_problem(valid_goal=False):
    from oxmpl_b200.base import GpuConeChecker, ProblemDefinition, SO3State, SO3StateSpace
    space = SO3StateSpace()
    start = _quat_axis_angle((0.0, 1.0, 0.0), math.pi / 2.0)
    goal = _QuaternionGoalRegion(space, _quat_axis_angle((0.0, 1.0, 0.0), -math.pi / 2.0),
                                 math.radians(10.0), valid_only=valid_goal)
    checker = GpuConeChecker([(SO3State.identity(), math.radians(44.9))])
    return space, start, goal, checker, ProblemDefinition.from_so3(space, start, goal)


def test_rrt_finds_path_in_so3ss(oxg):
    from oxmpl_b200.geometric import RRT
    space, start, goal, checker, pd = _so3_problem()
    assert checker(start) and checker(goal.target)
    planner = RRT(max_distance=0.5, goal_bias=0.05, problem_definition=pd)
    with pytest.raises(TypeError):
        planner.setup(lambda s: True)
    planner.setup(checker)
    path = planner.solve(timeout_secs=20.0)
    assert len(path.states) > 1
    assert space.distance(path.states[0], start) < 1e-9
    assert goal.is_satisfied(path.states[-1])
    forbidden = math.radians(44.9)
    ident = type(start).identity()
    for a, b in zip(path.states[:-1], path.states[1:]):        # is_path_valid, rrt_so3ss_tests.rs:106-143
        assert space.distance(ident, a) > forbidden and space.distance(ident, b) > forbidden
        n = math.ceil(space.distance(a, b) / space.get_longest_valid_segment_length())
        for j in range(1, n + 1 if n > 1 else 0):
            assert space.distance(ident, space.interpolate(a, b, j / n)) > forbidden


def test_so3_rrt_tree_identical_to_oracle_loop(oxg, orc):
    """The unfused GPU loop and the same loop driven by the oracle grow the same tree, bit for bit."""
    from oxmpl_b200.geometric import RRT
    space, start, goal, checker, pd = _so3_problem()
    planner = RRT(max_distance=0.5, goal_bias=0.05, problem_definition=pd, seed=7)
    planner.setup(checker)
    path = planner.solve(timeout_secs=20.0)
    got = np.asarray([n.values for n in planner.nodes])
    # replay with the oracle: same draws (seed 7 for the planner, 123 inside the goal)
    space2, start2, goal2, _, _ = _so3_problem()
    rng = random.Random(7)
    cones = np.array([[0.0, 0.0, 0.0, 1.0, math.radians(44.9)]])
    res = space2.get_longest_valid_segment_length() * 0.1
    rows, parents = [np.asarray(start2.values)], [-1]
    while True:
        q = goal2.sample_goal() if rng.random() < 0.05 else space2.sample_uniform(rng)
        q = np.asarray(q.values)
        idx, dist = orc.so3_nn1(np.asarray(rows), q)
        near, d = int(idx[0]), float(dist[0])
        q_new = orc.so3_interpolate(rows[near], q, 0.5 / d) if d > 0.5 else q
        valid, _ = orc.so3_check_motion(cones, rows[near], q_new, res)
        if valid[0]:
            rows.append(q_new)
            parents.append(near)
            if orc.so3_distance(goal2.target.values, q_new) <= goal2.radius:
                break
        assert len(rows) <= len(got)
    want = np.asarray(rows)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got.view(np.uint64), want.view(np.uint64))
    assert planner.parents == parents
    assert len(path.states) >= 2


def test_prm_finds_path_in_so3ss_and_roadmap_matches_oracle(oxg, orc):
    """oxmpl-py/tests/test_prm_so3ss.py: PRM(timeout=5, connection_radius=0.5) around the forbidden
    cone; the roadmap (prm.rs:117-146: lower-index neighbours inside the radius whose motion is
    valid) is also rebuilt with the oracle and compared edge list by edge list."""
    from oxmpl_b200.geometric import PRM
    space, start, goal, checker, pd = _so3_problem()
    planner = PRM(timeout=5.0, connection_radius=0.5, problem_definition=pd, seed=3)
    with pytest.raises(PlanningError):
        planner.solve(1.0)                                    # PlannerUninitialised
    planner.setup(checker)
    with pytest.raises(PlanningError):
        planner.solve(1.0)                                    # UnsampledStateSpace
    planner.construct_roadmap(max_samples=2500)
    path = planner.solve(timeout_secs=5.0)
    assert len(path.states) > 1
    assert space.distance(path.states[0], start) < 1e-9
    assert goal.is_satisfied(path.states[-1])
    forbidden = math.radians(44.9)
    ident = type(start).identity()
    for s in path.states:
        assert space.distance(ident, s) > forbidden
    # oracle rebuild of the same roadmap from the same valid samples
    rows = planner.rows
    cones = np.array([[0.0, 0.0, 0.0, 1.0, forbidden]])
    res = space.get_longest_valid_segment_length() * 0.1
    assert len(rows) > 800 and orc.so3_is_valid(cones, rows).all()
    edges = [[] for _ in range(len(rows))]
    for i in range(1, len(rows)):
        _, nb, _ = orc.so3_radius(rows[:i], rows[i], 0.5)
        if len(nb):
            ok, _ = orc.so3_check_motion(cones, np.repeat(rows[i:i + 1], len(nb), axis=0), rows[nb], res)
            for j in nb[ok.astype(bool)]:
                edges[i].append(int(j))
                edges[int(j)].append(i)
    assert planner.get_roadmap() == edges
    assert sum(len(e) for e in edges) > 2 * len(rows)


def _check_so3_path(space, start, goal, path):
    forbidden = math.radians(44.9)
    ident = type(start).identity()
    assert len(path.states) > 1
    assert space.distance(path.states[0], start) < 1e-9
    assert goal.is_satisfied(path.states[-1])
    for a, b in zip(path.states[:-1], path.states[1:]):
        assert space.distance(ident, a) > forbidden and space.distance(ident, b) > forbidden
        n = math.ceil(space.distance(a, b) / space.get_longest_valid_segment_length())
        for j in range(1, n + 1 if n > 1 else 0):
            assert space.distance(ident, space.interpolate(a, b, j / n)) > forbidden


@pytest.mark.parametrize("max_distance", [1.0, 0.05])
def test_rrt_connect_so3ss_and_trees_match_oracle_loop(oxg, orc, max_distance):
    """oxmpl-py/tests/test_rrt_connect_so3ss.py (RRTConnect(1.0, 0.05); 0.05 grows longer trees) +
    both trees rebuilt by the same loop on the oracle (rrt_connect.rs:121-159,241-308)."""
    from oxmpl_b200.geometric import RRTConnect
    space, start, goal, checker, pd = _so3_problem(valid_goal=True)
    planner = RRTConnect(max_distance=max_distance, goal_bias=0.05, problem_definition=pd, seed=11)
    planner.setup(checker)
    path = planner.solve(timeout_secs=20.0)
    _check_so3_path(space, start, goal, path)
    # oracle replay: same draws (planner seed 11; the goal's own generator, first used at setup)
    space2, start2, goal2, _, _ = _so3_problem(valid_goal=True)
    rng = random.Random(11)
    cones = np.array([[0.0, 0.0, 0.0, 1.0, math.radians(44.9)]])
    res = space2.get_longest_valid_segment_length() * 0.1
    trees = [[np.asarray(start2.values)], [np.asarray(goal2.sample_goal().values)]]
    parents = [[-1], [-1]]

    def extend(t, target):
        idx, dist = orc.so3_nn1(np.asarray(trees[t]), target)
        near, d = int(idx[0]), float(dist[0])
        reached = not d > max_distance
        q_new = target if reached else orc.so3_interpolate(trees[t][near], target, max_distance / d)
        if not orc.so3_check_motion(cones, trees[t][near], q_new, res)[0][0]:
            return None
        trees[t].append(q_new)
        parents[t].append(near)
        return reached, len(trees[t]) - 1

    for _ in range(100000):
        a = 0 if len(trees[0]) <= len(trees[1]) else 1
        q = goal2.sample_goal() if rng.random() < 0.05 else space2.sample_uniform(rng)
        ext = extend(a, np.asarray(q.values))
        if ext is None:
            continue
        q_new = trees[a][ext[1]]
        if a == 0 and orc.so3_distance(goal2.target.values, q_new) <= goal2.radius:
            break
        con = extend(a ^ 1, q_new)
        if con is not None and con[0]:
            break
    for t, tree in ((0, planner.ta), (1, planner.tb)):
        got = np.asarray([n.values for n in tree.nodes])
        want = np.asarray(trees[t])
        assert got.shape == want.shape
        np.testing.assert_array_equal(got.view(np.uint64), want.view(np.uint64))
        assert tree.parents == parents[t]


def test_rrt_star_so3ss_and_tree_matches_oracle_loop(oxg, orc):
    """oxmpl-py/tests/test_rrt_star_so3ss.py (RRTStar(0.5, 0.05, 0.25)) + tree, parents and costs
    rebuilt by the sequential reference loop on the oracle (rrt_star.rs:182-289)."""
    from oxmpl_b200.geometric import RRTStar
    space, start, goal, checker, pd = _so3_problem()
    planner = RRTStar(max_distance=0.5, goal_bias=0.05, search_radius=0.25, problem_definition=pd,
                      seed=5)
    planner.setup(checker)
    path = planner.solve(timeout_secs=30.0)
    _check_so3_path(space, start, goal, path)
    space2, start2, goal2, _, _ = _so3_problem()
    rng = random.Random(5)
    cones = np.array([[0.0, 0.0, 0.0, 1.0, math.radians(44.9)]])
    res = space2.get_longest_valid_segment_length() * 0.1
    rows, parents, costs = [np.asarray(start2.values)], [-1], [0.0]

    def motion(a, b):
        return bool(orc.so3_check_motion(cones, a, b, res)[0][0])

    for _ in range(100000):
        q = goal2.sample_goal() if rng.random() < 0.05 else space2.sample_uniform(rng)
        q = np.asarray(q.values)
        idx, dist = orc.so3_nn1(np.asarray(rows), q)
        near, d = int(idx[0]), float(dist[0])
        q_new = orc.so3_interpolate(rows[near], q, 0.5 / d) if d > 0.5 else q
        if not motion(rows[near], q_new):
            continue
        _, nbrs, _ = orc.so3_radius(np.asarray(rows), q_new, 0.25)
        best, min_cost = near, costs[near] + orc.so3_distance(q_new, rows[near])
        for j in nbrs:                                        # sequential, lazy: the reference loop
            c = costs[j] + orc.so3_distance(q_new, rows[j])
            if c < min_cost and motion(rows[j], q_new):
                min_cost, best = c, int(j)
        rows.append(q_new)
        parents.append(best)
        costs.append(min_cost)
        new = len(rows) - 1
        for j in nbrs:
            if parents[new] == j:
                continue
            c = costs[new] + orc.so3_distance(rows[j], q_new)
            if c < costs[j] and motion(q_new, rows[j]):
                parents[j], costs[j] = new, c
        if orc.so3_distance(goal2.target.values, q_new) <= goal2.radius:
            break
    got = np.asarray([n.values for n in planner.t.nodes])
    want = np.asarray(rows)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got.view(np.uint64), want.view(np.uint64))
    assert planner.t.parents == parents
    np.testing.assert_array_equal(np.asarray(planner.t.costs).view(np.uint64),
                                  np.asarray(costs).view(np.uint64))


# ---------------------------------------------------------------------------------------------
# SO2: oxmpl-py/tests/test_{rrt,rrt_connect,rrt_star,prm}_so2ss.py — forbidden zone [-0.5, 0.5],
# start -pi/2, goal pi/2 +- 0.1; the callback is_angle_valid is kept as the independent check.
# ---------------------------------------------------------------------------------------------
class AngleGoalRegion:                   # oxmpl-py/tests/test_rrt_so2ss.py:9-21
    def __init__(self, space, target_angle, radius):
        from oxmpl_b200.base import SO2State
        self.space = space
        self.target = SO2State(target_angle)
        self.radius = radius
        self.rng = random.Random(456)

    def is_satisfied(self, state):
        return self.space.distance(self.target, state) <= self.radius

    def sample_goal(self):
        from oxmpl_b200.base import SO2State
        return SO2State(self.target.value + self.rng.uniform(-self.radius, self.radius))


def is_angle_valid(state):               # oxmpl-py/tests/test_rrt_so2ss.py:24-34
    return not (-0.5 <= state.value <= 0.5)


def _so2_problem():
    from oxmpl_b200.base import GpuIntervalChecker, SO2State, SO2StateSpace
    space = SO2StateSpace()
    start = SO2State(-math.pi / 2.0)
    goal = AngleGoalRegion(space, target_angle=math.pi / 2.0, radius=0.1)
    return space, start, goal, ProblemDefinition.from_so2(space, start, goal), GpuIntervalChecker([(-0.5, 0.5)])


def _check_so2(path, space, start, goal):
    assert len(path.states) > 1
    assert space.distance(path.states[0], start) < 1e-9
    assert goal.is_satisfied(path.states[-1])
    for s in path.states:
        assert is_angle_valid(s), s.value


def test_so2_state_and_space_follow_the_reference(orc):
    from oxmpl_b200.base import SO2State, SO2StateSpace
    assert SO2State(3.0 * math.pi / 2.0).value == -math.pi / 2.0       # so2_state.rs:80-87, exactly
    assert abs(SO2State(5.0 * math.pi).value + math.pi) < 1e-9          # so2_state.rs:28-32
    space = SO2StateSpace()
    assert space.get_maximum_extent() == math.pi
    rng = random.Random(3)
    for _ in range(300):
        a, b, t = rng.uniform(-7, 7), rng.uniform(-7, 7), rng.random()
        sa, sb = SO2State(a), SO2State(b)
        sp_o = orc.space(0, 1)
        assert space.distance(sa, sb) == orc.distance(sp_o, [sa.value], [sb.value])
        assert space.interpolate(sa, sb, t).value == orc.interpolate(sp_o, [sa.value], [sb.value], t)[0]


@pytest.mark.parametrize("planner_name", ["RRT", "RRTConnect", "RRTStar"])
def test_tree_planners_find_path_in_so2ss(planner_name):
    from oxmpl_b200 import geometric
    space, start, goal, pd, checker = _so2_problem()
    kw = dict(max_distance=0.5, goal_bias=0.05, problem_definition=pd)
    if planner_name == "RRTStar":
        kw["search_radius"] = 0.25
    planner = getattr(geometric, planner_name)(**kw)
    with pytest.raises(TypeError):
        planner.setup(is_angle_valid)                # a Python callback is not a GPU checker
    planner.setup(checker)
    _check_so2(planner.solve(timeout_secs=5.0), space, start, goal)


def test_prm_finds_path_in_so2ss():
    space, start, goal, pd, checker = _so2_problem()
    planner = PRM(timeout=5.0, connection_radius=0.5, problem_definition=pd)
    planner.setup(checker)
    planner.construct_roadmap(max_samples=4000)
    _check_so2(planner.solve(timeout_secs=5.0), space, start, goal)


def test_rrt_star_dense_neighbourhood_is_a_planning_error():
    """More than OXG_RRT_STAR_MAX_NEIGHBOURS tree nodes inside search_radius (include/oxg.h): the
    mirror reports the reference's NoSolutionFound with the library's explanation."""
    space, start, goal, pd = _problem()
    planner = RRTStar(max_distance=0.5, goal_bias=0.0, search_radius=1e9, problem_definition=pd)
    planner.chunk = 6000
    planner.setup(GpuValidityChecker(2, aabbs=[[20.0, 20.0, 21.0, 21.0]]))
    planner.pd.goal.radius = 1e-12                   # never satisfied: the tree keeps growing
    with pytest.raises(PlanningError, match="No solution found.*search_radius"):
        planner.solve(timeout_secs=20.0)


def test_rrt_accepts_a_goal_that_is_not_a_ball():
    """A reference-style user Goal with an arbitrary is_satisfied (here: a half-plane) has no device
    form; RRT then runs the unfused per-iteration path through the same C-ABI entry points."""
    class RightEdgeGoal:
        def __init__(self):
            self.rng = random.Random(9)

        def is_satisfied(self, state):
            return state.values[0] >= 9.0

        def sample_goal(self):
            return RealVectorState([self.rng.uniform(9.0, 10.0), self.rng.uniform(0.0, 10.0)])

    space = RealVectorStateSpace(dimension=2, bounds=[(0.0, 10.0), (0.0, 10.0)])
    start = RealVectorState([1.0, 5.0])
    goal = RightEdgeGoal()
    pd = ProblemDefinition.from_real_vector(space, start, goal)
    planner = RRT(max_distance=0.5, goal_bias=0.05, problem_definition=pd)
    planner.setup(GpuValidityChecker(2, aabbs=WALL))
    path = planner.solve(timeout_secs=20.0)
    assert len(path.states) > 1 and space.distance(path.states[0], start) < 1e-9
    assert goal.is_satisfied(path.states[-1])
    for s in path.states:
        assert is_state_valid(s), s.values
    for a, b in zip(path.states[:-1], path.states[1:]):
        assert space.distance(a, b) <= 0.5 + 1e-9
