"""Mirror of ``oxmpl_py.base`` for the GPU path.

Same names and call shapes as the reference's Python binding (oxmpl-py/src/base/*.rs):
``RealVectorState(values)``, ``RealVectorStateSpace(dimension, bounds)``,
``ProblemDefinition.from_real_vector(space, start_state, goal)``, ``Path.states``.  The one
addition is :class:`GpuValidityChecker`: a declarative obstacle set that replaces the Python
``is_state_valid`` callback, so the hot loop never re-enters the interpreter (the reference
calls back into Python once per interpolated state,
oxmpl-py/src/base/state_validity_checker.rs:30-48).
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field

import numpy as np

from . import api


class RealVectorState:
    """oxmpl_py.base.RealVectorState (oxmpl-py/src/base/real_vector_state.rs)."""

    def __init__(self, values):
        self.values = [float(v) for v in values]

    def __repr__(self):
        return f"RealVectorState({self.values})"


class RealVectorStateSpace:
    """oxmpl_py.base.RealVectorStateSpace; host arithmetic is the reference's
    (real_vector_state_space.rs:103-155,251-253) in Python floats (IEEE f64, no FMA)."""

    def __init__(self, dimension: int, bounds=None):
        self.dimension = int(dimension)
        if bounds is None:
            bounds = [(-math.inf, math.inf)] * self.dimension
        if len(bounds) != self.dimension:
            raise ValueError("DimensionMismatch: bounds do not match the dimension")
        for lo, hi in bounds:
            if lo >= hi:
                raise ValueError("InvalidBound: lower bound must be less than upper bound")
        self.bounds = [(float(lo), float(hi)) for lo, hi in bounds]
        self._fraction = 0.05

    def distance(self, state1: RealVectorState, state2: RealVectorState) -> float:
        s = 0.0
        for a, b in zip(state1.values, state2.values):
            d = a - b
            s = s + d * d
        return math.sqrt(s)

    def interpolate(self, frm: RealVectorState, to: RealVectorState, t: float) -> RealVectorState:
        # real_vector_state_space.rs:161-186: from + (to - from) * t, each operation rounded
        return RealVectorState([f + (g - f) * t for f, g in zip(frm.values, to.values)])

    def get_maximum_extent(self) -> float:
        return api.real_vector_max_extent(self.bounds)

    def set_longest_valid_segment_fraction(self, fraction: float):
        self._fraction = fraction if 0.0 < fraction <= 1.0 else (0.0 if fraction <= 0.0 else 1.0)

    def get_longest_valid_segment_length(self) -> float:
        return self.get_maximum_extent() * self._fraction

    def sample_uniform(self, rng: random.Random) -> RealVectorState:
        return RealVectorState([rng.uniform(lo, hi) for lo, hi in self.bounds])

    def _space(self) -> api.Space:
        return api.Space.real_vector(self.dimension)


class GpuValidityChecker:
    """Declarative ``StateValidityChecker`` (oxmpl/src/base/validity.rs:39-48): spheres
    ``[c..., r]`` (valid iff distance > r), boxes ``[lo..., hi...]`` (invalid inside, faces
    inclusive), capsules ``[a..., b..., r]``.  Callable like the reference's callback."""

    def __init__(self, dimension: int, spheres=None, aabbs=None, capsules=None):
        self.dimension = dimension
        self.world = api.World(dimension, spheres=spheres, aabbs=aabbs, capsules=capsules)
        self._space = api.Space.real_vector(dimension)

    def __call__(self, state: RealVectorState) -> bool:
        return bool(self.world.is_valid(self._space, np.asarray([state.values]))[0])

    def is_valid_batch(self, rows) -> np.ndarray:
        return self.world.is_valid(self._space, rows)


def _so2_normalise(v: float) -> float:
    """SO2State::normalise (oxmpl/src/base/states/so2_state.rs:55-59):
    ``(value + PI).rem_euclid(2.0 * PI) - PI`` — ``%`` on f64 is C fmod (exact), rem_euclid adds
    ``|rhs|`` when the remainder is negative."""
    r = math.fmod(v + math.pi, 2.0 * math.pi)
    if r < 0.0:
        r = r + abs(2.0 * math.pi)
    return r - math.pi


class SO2State:
    """oxmpl_py.base.SO2State(value) (oxmpl-py/src/base/so2_state.rs): an angle, normalised to
    [-pi, pi) on construction as ``SO2State::new`` does (so2_state.rs:34-38)."""

    __slots__ = ("value",)

    def __init__(self, value: float):
        self.value = _so2_normalise(float(value))

    @property
    def values(self):
        return [self.value]

    def __repr__(self):
        return f"SO2State(value={self.value})"


class SO2StateSpace:
    """oxmpl_py.base.SO2StateSpace(bounds=None); host arithmetic restates
    oxmpl/src/base/spaces/so2_state_space.rs:57-79,97-122,161-170 in Python floats."""

    dimension = 1

    def __init__(self, bounds=None):
        if bounds is None:
            bounds = (-math.pi, math.pi)
        lo, hi = float(bounds[0]), float(bounds[1])
        if lo >= hi:
            raise ValueError("InvalidBound: lower bound must be less than upper bound")
        self.bounds = (lo, hi)
        self._fraction = 0.05

    def distance(self, state1: SO2State, state2: SO2State) -> float:
        diff = state1.value - state2.value                      # so2_state_space.rs:97-101
        return abs(_so2_normalise(diff))

    def interpolate(self, frm: SO2State, to: SO2State, t: float) -> SO2State:
        diff = _so2_normalise(to.value) - _so2_normalise(frm.value)   # so2_state_space.rs:107-122
        if diff > math.pi:
            diff -= 2.0 * math.pi
        elif diff < -math.pi:
            diff += 2.0 * math.pi
        return SO2State(frm.value + diff * t)

    def get_maximum_extent(self) -> float:
        return math.pi

    def set_longest_valid_segment_fraction(self, fraction: float):
        self._fraction = fraction if 0.0 < fraction <= 1.0 else (0.0 if fraction <= 0.0 else 1.0)

    def get_longest_valid_segment_length(self) -> float:
        return self.get_maximum_extent() * self._fraction

    def satisfies_bounds(self, state: SO2State) -> bool:
        v = _so2_normalise(state.value)
        return self.bounds[0] <= v <= self.bounds[1]

    def sample_uniform(self, rng: random.Random) -> SO2State:
        return SO2State(rng.uniform(self.bounds[0], self.bounds[1]))

    def _space(self) -> api.Space:
        return api.Space.so2()


class GpuIntervalChecker:
    """Declarative ``StateValidityChecker<SO2State>``: forbidden angle intervals ``(lo, hi)``; a
    state is invalid iff ``lo <= normalise(angle) <= hi`` for one of them (the reference's SO2
    fixture, oxmpl/tests/rrt_so2ss_tests.rs:17-27 / oxmpl-py/tests/test_rrt_so2ss.py:24-34)."""

    dimension = 0

    def __init__(self, intervals):
        self.world = api.World(0, so2_intervals=[(0, float(lo), float(hi)) for lo, hi in intervals])
        self._space = api.Space.so2()

    def __call__(self, state: SO2State) -> bool:
        return bool(self.world.is_valid(self._space, np.asarray([state.values]))[0])

    def is_valid_batch(self, rows) -> np.ndarray:
        return self.world.is_valid(self._space, rows)


class SO3State:
    """oxmpl_py.base.SO3State(x, y, z, w) (oxmpl-py/src/base/so3_state.rs): a unit quaternion."""

    __slots__ = ("x", "y", "z", "w")

    def __init__(self, x: float, y: float, z: float, w: float):
        self.x, self.y, self.z, self.w = float(x), float(y), float(z), float(w)

    @staticmethod
    def identity() -> "SO3State":
        return SO3State(0.0, 0.0, 0.0, 1.0)

    @property
    def values(self):
        return [self.x, self.y, self.z, self.w]

    def __repr__(self):
        return f"SO3State(x={self.x}, y={self.y}, z={self.z}, w={self.w})"


class SO3StateSpace:
    """oxmpl_py.base.SO3StateSpace(bounds=None); host arithmetic restates
    oxmpl/src/base/spaces/so3_state_space.rs:26-60,99-159,196-232 in Python floats (IEEE f64, no
    FMA; math.acos / math.sin are the libm calls Rust makes)."""

    def __init__(self, bounds=None):
        if bounds is None:
            self.bounds = (SO3State.identity(), math.pi)
        else:
            centre, max_angle = bounds
            if max_angle < 0.0:
                raise ValueError("InvalidAngularDistance: the maximum angle must be non-negative")
            self.bounds = (centre, min(float(max_angle), math.pi))
        self._fraction = 0.05

    def distance(self, state1: SO3State, state2: SO3State) -> float:
        abs_dot = abs(state1.x * state2.x + state1.y * state2.y + state1.z * state2.z
                      + state1.w * state2.w)
        return 0.0 if abs_dot > 1.0 - 1e-9 else math.acos(abs_dot)

    def interpolate(self, frm: SO3State, to: SO3State, t: float) -> SO3State:
        dot = frm.x * to.x + frm.y * to.y + frm.z * to.z + frm.w * to.w
        sign = -1.0 if dot < 0.0 else 1.0
        dot *= sign
        f, g = frm.values, to.values
        if dot > 0.9995:
            v = [f[c] + t * (g[c] * sign - f[c]) for c in range(4)]
            norm = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3])
            return SO3State(*[c / norm for c in v])
        theta = math.acos(dot)
        sin_theta = math.sin(theta)
        s0 = math.sin((1.0 - t) * theta) / sin_theta
        s1 = math.sin(t * theta) / sin_theta * sign
        return SO3State(*[f[c] * s0 + g[c] * s1 for c in range(4)])

    def get_maximum_extent(self) -> float:
        return 0.5 * math.pi

    def set_longest_valid_segment_fraction(self, fraction: float):
        self._fraction = fraction if 0.0 < fraction <= 1.0 else (0.0 if fraction <= 0.0 else 1.0)

    def get_longest_valid_segment_length(self) -> float:
        return self.get_maximum_extent() * self._fraction

    def satisfies_bounds(self, state: SO3State) -> bool:
        return self.distance(self.bounds[0], state) <= self.bounds[1]

    def sample_uniform(self, rng: random.Random) -> SO3State:
        centre, max_angle = self.bounds
        if max_angle < 1e-9:
            return SO3State(*centre.values)
        while True:     # rejection sampling inside the unit 4-ball, so3_state_space.rs:205-228
            x, y, z, w = (rng.uniform(-1.0, 1.0) for _ in range(4))
            norm_sq = x * x + y * y + z * z + w * w
            if 1e-9 < norm_sq < 1.0:
                norm = math.sqrt(norm_sq)
                q = SO3State(x / norm, y / norm, z / norm, w / norm)
                if self.distance(centre, q) <= max_angle:
                    return q

    def _space(self) -> api.Space:
        return api.Space.real_vector(4)


class GpuConeChecker:
    """Declarative ``StateValidityChecker<SO3State>``: forbidden cones ``(centre, angle)``; a state
    is valid iff ``distance(centre, state) > angle`` for all of them (the reference's SO3 fixture,
    oxmpl/tests/rrt_so3ss_tests.rs:46-56 / oxmpl-py/tests/test_rrt_so3ss.py:38-44)."""

    def __init__(self, cones):
        rows = [list(c.values) + [float(a)] for c, a in cones]
        self.cones = api.So3Cones(np.asarray(rows, dtype=np.float64).reshape(-1, 5))

    def __call__(self, state: SO3State) -> bool:
        return bool(self.cones.is_valid(np.asarray([state.values]))[0])

    def is_valid_batch(self, rows) -> np.ndarray:
        return self.cones.is_valid(rows)


@dataclass
class ProblemDefinition:
    """oxmpl_py.base.ProblemDefinition (oxmpl-py/src/base/problem_definition.rs:51-77)."""

    space: RealVectorStateSpace
    start_state: RealVectorState
    goal: object   # .is_satisfied(), .sample_goal(); ball goals also expose .target and .radius

    @staticmethod
    def from_real_vector(space, start_state, goal) -> "ProblemDefinition":
        """oxmpl-py/src/base/problem_definition.rs:51-77.  A ball goal (``target`` + ``radius``,
        the CircularGoalRegion of the reference's tests) runs on the fused device loops; any other
        goal object with ``is_satisfied(state)`` and ``sample_goal()`` is accepted too and makes
        ``RRT`` take the unfused per-iteration path (``oxg_nn1`` -> steer -> ``oxg_check_motion``
        -> append), one round trip per iteration."""
        for attr in ("is_satisfied", "sample_goal"):
            if not hasattr(goal, attr):
                raise TypeError(f"goal must provide `{attr}`")
        return ProblemDefinition(space, start_state, goal)

    @staticmethod
    def from_so2(space, start_state, goal) -> "ProblemDefinition":
        """oxmpl-py/src/base/problem_definition.rs: an SO2 problem; the fused kernels test the goal
        as the ball ``distance(state, target) <= radius``, so the goal must expose both."""
        for attr in ("target", "radius", "is_satisfied", "sample_goal"):
            if not hasattr(goal, attr):
                raise TypeError(f"goal must provide `{attr}` (a ball goal: target + radius)")
        return ProblemDefinition(space, start_state, goal)

    @staticmethod
    def from_so3(space, start_state, goal) -> "ProblemDefinition":
        """oxmpl-py/src/base/problem_definition.rs:108-133: any goal object with
        ``is_satisfied(state)`` and ``sample_goal()``."""
        for attr in ("is_satisfied", "sample_goal"):
            if not hasattr(goal, attr):
                raise TypeError(f"goal must provide `{attr}`")
        return ProblemDefinition(space, start_state, goal)


@dataclass
class Path:
    """oxmpl_py.base.Path: ``.states`` is the list of states (oxmpl-py/src/base/path.rs:48-82)."""

    states: list = field(default_factory=list)

    def __len__(self):
        return len(self.states)


class PlanningError(RuntimeError):
    """oxmpl/src/base/error.rs:97-108 (the reference raises these messages as Python exceptions)."""

    TIMEOUT = "No solution found within timeout."
    NO_SOLUTION = "No solution found."
    UNINITIALISED = "<Planner>.setup() was not called, thus Planner is uninitialised."
    INVALID_START = "Start state is not valid in the current StateSpace."
    UNSAMPLED = "StateSpace is not sampled. Either Tree or Roadmap is empty."
