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


@dataclass
class ProblemDefinition:
    """oxmpl_py.base.ProblemDefinition (oxmpl-py/src/base/problem_definition.rs:51-77)."""

    space: RealVectorStateSpace
    start_state: RealVectorState
    goal: object   # needs .target (RealVectorState), .radius, .is_satisfied(), .sample_goal()

    @staticmethod
    def from_real_vector(space, start_state, goal) -> "ProblemDefinition":
        for attr in ("target", "radius", "is_satisfied", "sample_goal"):
            if not hasattr(goal, attr):
                raise TypeError(f"goal must provide `{attr}` (a ball goal: target + radius)")
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
