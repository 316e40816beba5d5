"""GPU: the reference's Python integration tests (oxmpl-py/tests/test_{rrt,rrt_connect,
rrt_star,prm}_rvss.py) against the `oxmpl_b200.base` / `oxmpl_b200.geometric` mirror: same
world, same planner parameters, same assertions; the Python `is_state_valid` callback is
replaced by the declarative GpuValidityChecker and re-used here as the independent check."""
import math
import random

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
