"""CPU: the oracle's SO(3) restatement (oxmpl/src/base/spaces/so3_state_space.rs:99-159) against an
independent pure-Python restatement — math.acos / math.sin are the same libm calls — and the few
closed-form facts the metric has.  The reference holds no golden value for this path."""
import math

import numpy as np


def _unit(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _py_distance(a, b):
    dot = abs(float(a[0]) * float(b[0]) + float(a[1]) * float(b[1]) + float(a[2]) * float(b[2])
              + float(a[3]) * float(b[3]))
    return 0.0 if dot > 1.0 - 1e-9 else math.acos(dot)


def _py_interpolate(f, g, t):
    f = [float(v) for v in f]
    g = [float(v) for v in g]
    dot = f[0] * g[0] + f[1] * g[1] + f[2] * g[2] + f[3] * g[3]
    sign = -1.0 if dot < 0.0 else 1.0
    dot *= sign
    if dot > 0.9995:
        v = [f[c] + t * (g[c] * sign - f[c]) for c in range(4)]
        norm = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3])
        return [x / norm for x in v]
    theta = math.acos(dot)
    st = math.sin(theta)
    s0 = math.sin((1.0 - t) * theta) / st
    s1 = math.sin(t * theta) / st * sign
    return [f[c] * s0 + g[c] * s1 for c in range(4)]


def test_so3_distance_matches_python_restatement(orc):
    rng = np.random.default_rng(11)
    a, b = _unit(rng, 500), _unit(rng, 500)
    for i in range(500):
        got = orc.so3_distance(a[i], b[i])
        assert np.float64(got).view(np.uint64) == np.float64(_py_distance(a[i], b[i])).view(np.uint64)


def test_so3_distance_closed_forms(orc):
    ident = [0.0, 0.0, 0.0, 1.0]
    assert orc.so3_distance(ident, ident) == 0.0
    q = _unit(np.random.default_rng(1), 1)[0]
    assert orc.so3_distance(q, -q) == 0.0                       # double cover: q and -q coincide
    assert orc.so3_distance(q, q * (1.0 - 1e-10)) == 0.0        # the 1 - 1e-9 clamp
    # rotation by theta about z: dot = cos(theta/2), so the metric is half the rotation angle
    for theta in (0.1, 1.0, math.pi / 2, 3.0):
        r = [0.0, 0.0, math.sin(theta / 2), math.cos(theta / 2)]
        assert abs(orc.so3_distance(ident, r) - theta / 2) < 1e-15 * 4
    assert orc.so3_distance([1.0, 0, 0, 0], [0, 1.0, 0, 0]) == math.acos(0.0)   # maximum extent pi/2


def test_so3_interpolate_matches_python_restatement_and_stays_unit(orc):
    rng = np.random.default_rng(12)
    a, b = _unit(rng, 200), _unit(rng, 200)
    # half of the pairs are close (LERP branch), a quarter have a negative dot (sign flip)
    b[:100] = a[:100] + rng.normal(size=(100, 4)) * 0.01
    b[:100] /= np.linalg.norm(b[:100], axis=1, keepdims=True)
    b[100:150] = np.where((np.sum(a[100:150] * b[100:150], axis=1) > 0)[:, None], -b[100:150], b[100:150])
    for i in range(200):
        for t in (0.0, 0.25, 1.0 / 3.0, 1.0):
            got = orc.so3_interpolate(a[i], b[i], t)
            want = np.array(_py_interpolate(a[i], b[i], t))
            np.testing.assert_array_equal(got.view(np.uint64), want.view(np.uint64))
            assert abs(np.linalg.norm(got) - 1.0) < 1e-12


def test_so3_check_motion_steps(orc):
    res = 0.5 * math.pi * 0.05 * 0.1     # SO3 get_maximum_extent (= pi/2) * 0.05 * 0.1, rrt.rs:97
    ident = np.array([0.0, 0.0, 0.0, 1.0])
    theta = 1.0
    to = np.array([0.0, 0.0, math.sin(theta / 2), math.cos(theta / 2)])
    # a cone around the rotation by 0.5 about z, of metric radius 0.05, blocks the z-sweep
    mid = [0.0, 0.0, math.sin(0.25), math.cos(0.25), 0.05]
    valid, first = orc.so3_check_motion([mid], ident, to, res)
    n = math.ceil(0.5 / res)
    assert valid[0] == 0 and 1 < first[0] < n
    # the first blocked step is the first whose half-angle i/n * 0.5 is within 0.05 of 0.25
    want = next(i for i in range(1, n + 1) if abs(i / n * 0.5 - 0.25) <= 0.05 - 1e-12)
    assert abs(int(first[0]) - want) <= 1
    valid, first = orc.so3_check_motion([[1.0, 0.0, 0.0, 0.0, 0.05]], ident, to, res)
    assert valid[0] == 1 and first[0] == 0


def test_so3_knn_is_the_sorted_scan(orc):
    """orc.so3_knn = the k smallest (distance, index) pairs of the pure-Python metric; its first
    column is the 1-NN scan's winner."""
    rng = np.random.default_rng(5)
    rows, qs = _unit(rng, 400), _unit(rng, 7)
    rows[10] = rows[3]                                  # duplicate state: index 3 comes first
    qs[0] = rows[3]
    ex = np.array([3, 0xFFFFFFFF, 1, 2, 0xFFFFFFFF, 5, 399], np.uint32)
    for k, exclude in ((1, None), (6, None), (6, ex)):
        idx, dist = orc.so3_knn(rows, qs, k, exclude=exclude)
        for q in range(len(qs)):
            pairs = sorted((_py_distance(qs[q], rows[i]), i) for i in range(len(rows))
                           if exclude is None or i != exclude[q])[:k]
            assert idx[q].tolist() == [i for _, i in pairs]
            assert dist[q].tolist() == [d for d, _ in pairs]
    i1, d1 = orc.so3_nn1(rows, qs)
    idx, dist = orc.so3_knn(rows, qs, 3)
    assert (idx[:, 0] == i1).all() and (dist[:, 0] == d1).all()
    idx, dist = orc.so3_knn(rows[:2], qs[:1], 4)        # padding
    assert idx[0, 2:].tolist() == [0xFFFFFFFF] * 2 and np.isinf(dist[0, 2:]).all()
