"""Pins the CPU oracle against every exact known answer the reference's own tests hold for
this path and against an independent pure-Python restatement of the cited Rust
(Python floats are IEEE binary64; math.sqrt / math.fmod / '/' are correctly rounded)."""
import math

import numpy as np
import pytest

PI = math.pi


# ---- independent pure-Python restatements (operation order from the cited Rust) -----------
def py_rem_euclid(v, rhs):
    r = math.fmod(v, rhs)
    return r + abs(rhs) if r < 0.0 else r


def py_normalise(v):  # so2_state.rs:55-59
    return py_rem_euclid(v + PI, 2.0 * PI) - PI


def py_rv_distance(a, b):  # real_vector_state_space.rs:148-154
    s = 0.0
    for x, y in zip(a, b):
        d = x - y
        s = s + d * d
    return math.sqrt(s)


def py_so2_distance(a, b):  # so2_state_space.rs:97-101
    diff = a - b
    diff = py_rem_euclid(diff + PI, 2.0 * PI) - PI
    return abs(diff)


def py_so2_interpolate(f, t_, t):  # so2_state_space.rs:107-122
    diff = py_normalise(t_) - py_normalise(f)
    if diff > PI:
        diff -= 2.0 * PI
    elif diff < -PI:
        diff += 2.0 * PI
    return py_normalise(f + diff * t)


def py_check_motion(dist_fn, interp_fn, valid_fn, frm, to, res):  # rrt.rs:90-116
    dist = dist_fn(frm, to)
    x = dist / res
    n = 0 if math.isnan(x) else max(0, math.ceil(x)) if math.isfinite(x) else 2**64 - 1
    if n <= 1:
        return (1, 0) if valid_fn(to) else (0, 1)
    for i in range(1, n + 1):
        t = float(i) / float(n)
        if not valid_fn(interp_fn(frm, to, t)):
            return (0, i)
    return (1, 0)


# ---- reference known answers -------------------------------------------------------------
def test_so2_normalise_known_answers(orc):
    # test_so2_state_normalise: normalise(3*PI/2) == -PI/2 exactly (so2_state.rs:80-87)
    assert orc.so2_normalise(3.0 * PI / 2.0) == -PI / 2.0
    # doctests so2_state.rs:28-32
    assert abs(orc.so2_normalise(3.0 * PI / 2.0) - (-PI / 2.0)) < 1e-9
    assert abs(orc.so2_normalise(5.0 * PI) + PI) < 1e-9
    for v in (0.0, 1.0, -1.0, 3.0, -3.0):
        assert abs(orc.so2_normalise(v) - v) < 1e-15


def test_config1_resolution_constants(orc):
    # real_vector_state_space.rs:103-118,251-253 and rrt.rs:96-97
    ext = orc.rv_max_extent([-10.0, -10.0], [10.0, 10.0])
    assert ext == 28.284271247461902
    assert orc.lib.orc_longest_valid_segment_length(ext, 0.05) == 1.4142135623730951
    res = orc.resolution(ext)
    assert res == 0.14142135623730953
    assert math.ceil(0.5 / res) == 4
    assert orc.rv_max_extent([0.0, 0.0], [math.inf, 1.0]) == 1.0     # non-finite bound -> 1.0
    assert orc.lib.orc_so2_max_extent() == PI                         # so2_state_space.rs:77-79
    # set_longest_valid_segment_fraction clamp (real_vector_state_space.rs:121-129)
    assert orc.lib.orc_clamp_fraction(0.3) == 0.3
    assert orc.lib.orc_clamp_fraction(-1.0) == 0.0
    assert orc.lib.orc_clamp_fraction(7.0) == 1.0


def test_oracle_is_not_contracted(orc):
    a = 1.0 + 2.0 ** -27
    assert orc.lib.orc_probe_mul_add(a, a, -1.0) == 2.0 ** -26       # an FMA would add 2^-54
    assert orc.lib.orc_build_info() & 1


# ---- oracle vs the independent restatement ---------------------------------------------------
@pytest.mark.parametrize("d", [1, 2, 3, 6, 7, 14])
def test_rv_distance_and_interpolate_bit_exact(orc, d):
    rng = np.random.default_rng(d)
    sp = orc.space(d)
    for _ in range(300):
        a = rng.normal(scale=5.0, size=d)
        b = rng.normal(scale=5.0, size=d)
        assert orc.distance(sp, a, b) == py_rv_distance(a.tolist(), b.tolist())
        assert orc.distance(sp, a, b) == orc.distance(sp, b, a)       # RealVector is symmetric
        t = float(rng.random())
        got = orc.interpolate(sp, a, b, t)
        exp = [x + (y - x) * t for x, y in zip(a.tolist(), b.tolist())]
        assert got.tolist() == exp


def test_so2_distance_interpolate_bit_exact_and_asymmetric(orc):
    rng = np.random.default_rng(0)
    sp = orc.space(0, 1)
    asym = 0
    for _ in range(3000):
        a = float(rng.uniform(-4 * PI, 4 * PI))
        b = float(rng.uniform(-4 * PI, 4 * PI))
        d_ab = orc.distance(sp, [a], [b])
        assert d_ab == py_so2_distance(a, b)
        asym += d_ab != orc.distance(sp, [b], [a])
        t = float(rng.random())
        assert orc.interpolate(sp, [a], [b], t)[0] == py_so2_interpolate(a, b, t)
        assert 0.0 <= d_ab <= PI + 1e-12
    assert asym > 0        # argument order is observable (SURVEY hard part 3)


def test_compound_distance_definition(orc):
    rng = np.random.default_rng(1)
    sp = orc.space(2, 1, 1.0, 0.5)
    for _ in range(500):
        a = np.r_[rng.uniform(0, 100, 2), rng.uniform(-PI, PI)]
        b = np.r_[rng.uniform(0, 100, 2), rng.uniform(-PI, PI)]
        exp = 1.0 * py_rv_distance(a[:2].tolist(), b[:2].tolist()) + 0.5 * py_so2_distance(a[2], b[2])
        assert orc.distance(sp, a, b) == exp


def test_fixture_semantics(orc):
    sp = orc.space(2)
    # README circle: valid iff sqrt(x^2+y^2) > 2.0 — strict (README.md:147-150)
    w, keep = orc.world(2, spheres=[[0.0, 0.0, 2.0]])
    pts = [[2.0, 0.0], [math.nextafter(2.0, 3.0), 0.0], [1.2, 1.6], [3.0, 4.0]]
    assert orc.is_valid(w, sp, pts).tolist() == [0, 1, 0, 1]
    # wall: inclusive faces (oxmpl/tests/rrt_rvss_tests.rs:24-36)
    w, keep = orc.world(2, aabbs=[[4.75, 2.0, 5.25, 8.0]])
    pts = [[4.75, 2.0], [5.25, 8.0], [math.nextafter(4.75, 0.0), 5.0], [5.0, math.nextafter(8.0, 9.0)]]
    assert orc.is_valid(w, sp, pts).tolist() == [0, 0, 1, 1]
    # forbidden angle: normalise, invalid iff -0.5 <= v <= 0.5 (oxmpl/tests/rrt_so2ss_tests.rs:17-27)
    sp1 = orc.space(0, 1)
    w, keep = orc.world(0, so2_intervals=[(0, -0.5, 0.5)])
    pts = [[0.5], [-0.5], [0.5001], [2 * PI + 0.1], [PI]]
    assert orc.is_valid(w, sp1, pts).tolist() == [0, 0, 1, 0, 1]


def test_capsule_definition(orc):
    """The capsule predicate is defined by the oracle (parity unpinned): check it against the
    documented formula in pure Python."""
    rng = np.random.default_rng(2)
    d = 3
    caps = np.hstack([rng.uniform(-1, 1, (20, d)), rng.uniform(-1, 1, (20, d)), rng.uniform(0.1, 0.5, (20, 1))])
    caps[0, d:2 * d] = caps[0, :d]                        # degenerate segment: den == 0 -> t = 0
    sp = orc.space(d)
    w, keep = orc.world(d, capsules=caps)
    pts = rng.uniform(-1.2, 1.2, (2000, d))

    def valid(x):
        for c in caps.tolist():
            a, b, r = c[:d], c[d:2 * d], c[2 * d]
            den = num = None
            for i in range(d):
                ab = b[i] - a[i]
                ap = x[i] - a[i]
                den = ab * ab if den is None else den + ab * ab
                num = ap * ab if num is None else num + ap * ab
            t = num / den if den > 0.0 else 0.0
            t = 0.0 if t < 0.0 else t
            t = 1.0 if t > 1.0 else t
            d2 = None
            for i in range(d):
                e = x[i] - (a[i] + (b[i] - a[i]) * t)
                d2 = e * e if d2 is None else d2 + e * e
            if not math.sqrt(d2) > r:
                return False
        return True

    got = orc.is_valid(w, sp, pts)
    assert got.tolist() == [int(valid(p)) for p in pts.tolist()]
    assert 0 < got.sum() < len(got)


def test_check_motion_matches_reference_loop(orc):
    rng = np.random.default_rng(3)
    sp = orc.space(2)
    spheres = np.hstack([rng.uniform(-8, 8, (30, 2)), rng.uniform(0.3, 1.5, (30, 1))])
    w, keep = orc.world(2, spheres=spheres)
    res = orc.resolution(orc.rv_max_extent([-10, -10], [10, 10]))

    def valid(x):
        return all(py_rv_distance(x, c[:2]) > c[2] for c in spheres.tolist())

    def interp(f, t_, t):
        return [a + (b - a) * t for a, b in zip(f, t_)]

    frm = rng.uniform(-10, 10, (400, 2))
    to = frm + rng.normal(scale=1.0, size=(400, 2))
    to[:5] = frm[:5]
    v, first = orc.check_motion(w, sp, frm, to, res)
    for e in range(400):
        exp = py_check_motion(py_rv_distance, interp, valid, frm[e].tolist(), to[e].tolist(), res)
        assert (int(v[e]), int(first[e])) == exp
    assert 0 < v.sum() < 400


def test_nn_scan_semantics(orc):
    sp = orc.space(2)
    rows = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [3.0, 0.0], [0.0, 0.0]])
    # 1-NN: strict '<' -> the lowest index wins exact ties (rrt.rs:191-195)
    idx, dist = orc.nn1(sp, rows[:4], [[0.0, 0.0]])
    assert idx[0] == 0 and dist[0] == 1.0
    # radius: strict '<' (rrt_star.rs:125), ascending index, idx_limit = "only earlier nodes"
    off, ridx = orc.radius(sp, rows, [[0.0, 0.0]], 1.0)
    assert ridx.tolist() == [4]
    off, ridx = orc.radius(sp, rows, [[0.0, 0.0]], math.nextafter(1.0, 2.0))
    assert ridx.tolist() == [0, 1, 2, 4]
    off, ridx = orc.radius(sp, rows, [[0.0, 0.0]], 5.0, idx_limit=[3])
    assert ridx.tolist() == [0, 1, 2]
    # k-NN (oracle-defined): ascending (dist, idx), exclusion by index, padding
    idx, dist = orc.knn(sp, rows, [[0.0, 0.0]], 4, exclude=[4])
    assert idx[0].tolist() == [0, 1, 2, 3] and dist[0].tolist() == [1.0, 1.0, 1.0, 3.0]
    idx, dist = orc.knn(sp, rows[:2], [[0.0, 0.0]], 4)
    assert idx[0].tolist() == [0, 1, 0xFFFFFFFF, 0xFFFFFFFF] and math.isinf(dist[0][3])


@pytest.mark.parametrize("d", [2, 7])
def test_knn_matches_numpy_lexsort(orc, d):
    rng = np.random.default_rng(d)
    rows = rng.integers(0, 5, (400, d)).astype(float)         # many exact ties
    qs = rng.integers(0, 5, (50, d)).astype(float)
    sp = orc.space(d)
    idx, dist = orc.knn(sp, rows, qs, 9)
    for q in range(50):
        dd = np.array([py_rv_distance(qs[q].tolist(), r.tolist()) for r in rows])
        order = np.lexsort((np.arange(400), dd))[:9]
        assert idx[q].tolist() == order.tolist()
        assert dist[q].tolist() == dd[order].tolist()
