"""CPU: host-side logic of the Python mirror (oxmpl_b200.base / geometric) that needs no GPU —
the reference's space arithmetic, error strings, problem-definition checks — against the oracle
and the reference's own constants."""
import math
import random

import numpy as np
import pytest

from oxmpl_b200 import sharding
from oxmpl_b200.base import (PlanningError, ProblemDefinition, RealVectorState,
                             RealVectorStateSpace)


def test_real_vector_space_matches_oracle(orc):
    rng = np.random.default_rng(3)
    for d in (1, 2, 7, 14):
        space = RealVectorStateSpace(d, [(-2.0, 3.0)] * d)
        sp_o = orc.space(d)
        for _ in range(50):
            a, b = rng.normal(size=d) * 3, rng.normal(size=d) * 3
            got = space.distance(RealVectorState(a), RealVectorState(b))
            want = orc.distance(sp_o, a, b)
            assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64)
        assert space.get_maximum_extent() == orc.rv_max_extent([-2.0] * d, [3.0] * d)
        lvsl = space.get_longest_valid_segment_length()
        assert lvsl * 0.1 == orc.resolution(space.get_maximum_extent())


def test_reference_constants_and_clamps():
    space = RealVectorStateSpace(2, [(-10.0, 10.0), (-10.0, 10.0)])      # README world
    assert space.get_maximum_extent() == 28.284271247461902
    assert space.get_longest_valid_segment_length() * 0.1 == 0.14142135623730953
    space.set_longest_valid_segment_fraction(2.0)                           # clamp (rs:121-129)
    assert space.get_longest_valid_segment_length() == space.get_maximum_extent()
    space.set_longest_valid_segment_fraction(-1.0)
    assert space.get_longest_valid_segment_length() == 0.0
    assert RealVectorStateSpace(3).get_maximum_extent() == 1.0             # unbounded -> 1.0 (rs:103-118)
    with pytest.raises(ValueError):
        RealVectorStateSpace(2, [(0.0, 1.0)])
    with pytest.raises(ValueError):
        RealVectorStateSpace(1, [(1.0, 1.0)])
    s = space.sample_uniform(random.Random(0))
    assert all(-10.0 <= v <= 10.0 for v in s.values)


def test_problem_definition_and_errors():
    space = RealVectorStateSpace(2, [(0.0, 1.0), (0.0, 1.0)])

    class Goal:
        target = RealVectorState([0.5, 0.5])
        radius = 0.1

        def is_satisfied(self, s):
            return space.distance(self.target, s) <= self.radius

        def sample_goal(self):
            return self.target

    pd = ProblemDefinition.from_real_vector(space, RealVectorState([0.0, 0.0]), Goal())
    assert pd.goal.is_satisfied(RealVectorState([0.55, 0.5]))
    with pytest.raises(TypeError):
        ProblemDefinition.from_real_vector(space, RealVectorState([0.0, 0.0]), object())
    # messages of oxmpl/src/base/error.rs:110-140
    assert PlanningError.TIMEOUT == "No solution found within timeout."
    assert "uninitialised" in PlanningError.UNINITIALISED
    assert math.isclose(space.distance(RealVectorState([0, 0]), RealVectorState([3, 4])), 5.0)


def test_shard_ownership_is_a_partition():
    """Every candidate slot of a k-NN table is owned by exactly one of the two mutual slots."""
    rng = np.random.default_rng(5)
    n, k = 400, 6
    idx = np.empty((n, k), dtype=np.uint32)
    for i in range(n):
        idx[i] = rng.choice([j for j in range(n) if j != i], size=k, replace=False)
    idx[7, 3] = 0xFFFFFFFF                                            # a padded slot
    own_all = sharding.owned_slots(idx, 0, n)
    parts = [sharding.owned_slots(idx, *sharding.row_range(n, r, 3)) for r in range(3)]
    np.testing.assert_array_equal(np.vstack(parts), own_all)
    pairs = {}
    for i in range(n):
        for s in range(k):
            j = int(idx[i, s])
            if j == 0xFFFFFFFF:
                assert not own_all[i, s]
                continue
            pairs.setdefault((min(i, j), max(i, j)), []).append(bool(own_all[i, s]))
    assert all(sum(v) == 1 for v in pairs.values())
