"""GPU parity of the SO(3) scans / cone validity (include/oxg_so3.h) against the oracle: indices and
f64 bit patterns must be identical.  The adversarial sets use the query (1, 0, 0, 0), for which
a = |q . s| is exactly |s.x|, so the stored x column dictates `a` to the last bit."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RES = 0.5 * math.pi * 0.05 * 0.1
E1 = np.array([[1.0, 0.0, 0.0, 0.0]])


def _unit(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _store(oxg, rows):
    st = oxg.StateStore(oxg.Space.real_vector(4))
    st.append(rows)
    return st


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _x_rows(xs):
    rows = np.zeros((len(xs), 4))
    rows[:, 0] = xs
    return rows


def _plateau(x0, width=4):
    """`width` consecutive doubles ending at the top of an acos plateau near x0 that has at least
    two members (several adjacent x share one acos value when ulp(acos) > ulp(x))."""
    x = x0
    for _ in range(10000):
        lo = np.nextafter(x, 0.0)
        if math.acos(lo) == math.acos(x) and math.acos(np.nextafter(x, 1.0)) != math.acos(x):
            out = [x]
            for _ in range(width - 1):
                out.append(np.nextafter(out[-1], 0.0))
            return out[::-1]            # ascending, last = top of the plateau
        x = np.nextafter(x, 1.0)
    raise AssertionError("no acos plateau found")


@pytest.mark.parametrize("n,q", [(1, 1), (255, 3), (20000, 50), (5003, 3001), (300000, 9)])
def test_so3_nn1_random(oxg, orc, n, q):
    rng = np.random.default_rng(n + q)
    rows, qs = _unit(rng, n), _unit(rng, q)
    qs[::7] = rows[rng.integers(0, n, size=len(qs[::7]))]        # exact members (distance 0)
    with _store(oxg, rows) as st:
        idx, dist = st.so3_nn1(qs)
    oidx, odist = orc.so3_nn1(rows, qs)
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_array_equal(_bits(dist), _bits(odist))


def test_so3_nn1_plateaus_clamp_and_nan(oxg, orc):
    plat = _plateau(0.3)
    assert math.acos(plat[-1]) == math.acos(plat[-2])
    cases = []
    # the largest a sits at the highest index; a lower index one ulp below shares its acos
    cases.append([0.1, plat[-2], 0.2, plat[-1]])
    cases.append([0.1, plat[0], plat[-2], 0.2, plat[-1], 0.25])
    # next to a = 0 everything rounds to acos = pi/2: index 0 wins although its a is the smallest
    cases.append([0.0, 1e-30, 1e-20, 1e-18, 5e-324])
    cases.append([1e-300, 0.0, 1e-17])
    # clamp: every a > 1 - 1e-9 has distance 0, lowest index wins; +inf and values above 1 too
    cases.append([0.5, 1.0 - 1e-10, 1.0, 0.9999999995, np.inf, 7.0])
    cases.append([0.5, 1.0 - 1e-9, np.nextafter(1.0 - 1e-9, 2.0), 1.0])
    # NaN at index 0 wins (nothing is < NaN); NaN elsewhere never does
    cases.append([np.nan, 0.9, 1.0])
    cases.append([0.3, np.nan, 0.9, np.nan])
    cases.append([-0.7, 0.6, -0.6])                                # |.|: -0.7 is the closest
    for xs in cases:
        rows = _x_rows(xs)
        with _store(oxg, rows) as st:
            idx, dist = st.so3_nn1(E1)
        oidx, odist = orc.so3_nn1(rows, E1)
        assert idx[0] == oidx[0], (xs, idx, oidx)
        np.testing.assert_array_equal(_bits(dist), _bits(odist))
    # the same plateau buried in a large store that spans many slices
    rng = np.random.default_rng(9)
    xs = rng.uniform(-0.29, 0.29, size=70001)
    xs[60000] = plat[-1]
    xs[123] = -plat[-2]
    xs[40000] = plat[0]
    rows = _x_rows(xs)
    rows[:, 1:] = rng.normal(size=(len(xs), 3))                    # ignored by the query e1
    with _store(oxg, rows) as st:
        idx, dist = st.so3_nn1(np.vstack([E1, -E1, _unit(rng, 5)]))
    oidx, odist = orc.so3_nn1(rows, np.vstack([E1, -E1]))
    expected = 123 if math.acos(plat[-2]) == math.acos(plat[-1]) else 60000
    assert idx[0] == oidx[0] == expected and idx[1] == oidx[1]


@pytest.mark.parametrize("r", [0.3, 1.0, math.pi / 2, 2.0, 0.0, -1.0, float("nan"), 4.5e-5, 1e-300])
def test_so3_radius_random(oxg, orc, r):
    rng = np.random.default_rng(21)
    rows, qs = _unit(rng, 9001), _unit(rng, 21)
    qs[3] = rows[77]
    rows[100] = rows[77] * (1.0 - 1e-12)
    with _store(oxg, rows) as st:
        off, idx, dist = st.so3_radius(qs, r, want_dist=True)
        off2, idx2 = st.so3_radius(qs, r)
    ooff, oidx, odist = orc.so3_radius(rows, qs, r)
    np.testing.assert_array_equal(off, ooff)
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_array_equal(_bits(dist), _bits(odist))
    np.testing.assert_array_equal(off2, ooff)
    np.testing.assert_array_equal(idx2, oidx)


def test_so3_radius_threshold_shell_and_many_queries(oxg, orc):
    # consecutive doubles on both sides of cos(r): the strict `<` must flip at the oracle's ulp
    for r in (0.3, 1.2661036727794992, 1.5):
        c = math.cos(r)
        xs = [c]
        for _ in range(12):
            xs.append(np.nextafter(xs[-1], 1.0))
        lo = c
        for _ in range(12):
            lo = np.nextafter(lo, 0.0)
            xs.append(lo)
        xs += [0.0, 1.0, -c, np.nan, np.inf]
        rows = _x_rows(xs)
        with _store(oxg, rows) as st:
            off, idx, dist = st.so3_radius(E1, r, want_dist=True)
        ooff, oidx, odist = orc.so3_radius(rows, E1, r)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(idx, oidx)
        np.testing.assert_array_equal(_bits(dist), _bits(odist))
        assert 0 < len(idx) < len(xs)
    rng = np.random.default_rng(5)
    rows, qs = _unit(rng, 3000), _unit(rng, 2501)
    with _store(oxg, rows) as st:
        off, idx, dist = st.so3_radius(qs, 0.35, want_dist=True)
    ooff, oidx, odist = orc.so3_radius(rows, qs, 0.35)
    np.testing.assert_array_equal(off, ooff)
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_array_equal(_bits(dist), _bits(odist))


def test_so3_store_shape_and_errors(oxg):
    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        st.append(np.zeros((2, 3)))
        with pytest.raises(oxg.OxgError) as e:
            st.so3_nn1(np.zeros((1, 4)))
        assert e.value.code == -2                                  # OXG_E_DIM
    with oxg.StateStore(oxg.Space.real_vector(4)) as st:
        with pytest.raises(oxg.OxgError) as e:
            st.so3_nn1(E1)
        assert e.value.code == -7                                  # OXG_E_EMPTY
        off, idx = st.so3_radius(E1, 1.0)
        assert off.tolist() == [0, 0] and len(idx) == 0


def test_so3_cones_validity_and_boundaries(oxg, orc):
    rng = np.random.default_rng(31)
    cones = np.hstack([_unit(rng, 40), rng.uniform(0.02, 0.25, size=(40, 1))])
    states = _unit(rng, 20000)
    states[:40] = cones[:, :4]                                     # the centres themselves
    c = oxg.So3Cones(cones)
    got = c.is_valid(states)
    want = orc.so3_is_valid(cones, states)
    np.testing.assert_array_equal(got, want)
    assert 0 < int(got.sum()) < len(got)
    # exact boundary: states whose a walks through cos(angle) one ulp at a time (centre e1)
    for angle in (0.25, 1.0, 0.0, -0.5, float("nan"), math.pi / 2, 3.0):
        cone = np.array([[1.0, 0.0, 0.0, 0.0, angle]])
        base = math.cos(angle) if angle == angle else 0.5
        xs = [base]
        for _ in range(10):
            xs.append(np.nextafter(xs[-1], 2.0))
        lo = base
        for _ in range(10):
            lo = np.nextafter(lo, -2.0)
            xs.append(lo)
        xs += [0.0, 1.0, np.nan, 1.0 - 1e-9, 1.0 - 1e-10]
        st = _x_rows(xs)
        got = oxg.So3Cones(cone).is_valid(st)
        want = orc.so3_is_valid(cone, st)
        np.testing.assert_array_equal(got, want, err_msg=f"angle {angle}")
    assert oxg.So3Cones(np.zeros((0, 5))).is_valid(states[:5]).tolist() == [1] * 5


def test_so3_check_motion_matches_oracle(oxg, orc):
    rng = np.random.default_rng(41)
    cones = np.hstack([_unit(rng, 12), rng.uniform(0.1, 0.35, size=(12, 1))])
    frm, to = _unit(rng, 600), _unit(rng, 600)
    to[:150] = frm[:150] + rng.normal(size=(150, 4)) * 0.01       # LERP branch (dot > 0.9995)
    to[:150] /= np.linalg.norm(to[:150], axis=1, keepdims=True)
    to[150:160] = frm[150:160]                                     # zero-length edges (n <= 1)
    to[160:170] = -frm[160:170]
    c = oxg.So3Cones(cones)
    valid, first = c.check_motion(frm, to, RES, want_first_invalid=True)
    ovalid, ofirst = orc.so3_check_motion(cones, frm, to, RES)
    np.testing.assert_array_equal(valid, ovalid)
    np.testing.assert_array_equal(first, ofirst)
    assert 0 < int(valid.sum()) < len(valid)
    with pytest.raises(oxg.OxgError) as e:
        c.check_motion(frm[:1], to[:1], 0.0)
    assert e.value.code == -5                                      # OXG_E_RESOLUTION
