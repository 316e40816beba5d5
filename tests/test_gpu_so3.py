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


@pytest.mark.parametrize("n,q,k", [(1, 1, 1), (5, 3, 8), (255, 3, 10), (20000, 50, 10), (5003, 3001, 4),
                                   (300000, 9, 32), (70000, 1, 10), (33, 40, 32)])
def test_so3_knn_random(oxg, orc, n, q, k):
    rng = np.random.default_rng(n + q + k)
    rows, qs = _unit(rng, n), _unit(rng, q)
    qs[::7] = rows[rng.integers(0, n, size=len(qs[::7]))]        # exact members (distance 0)
    with _store(oxg, rows) as st:
        idx, dist = st.so3_knn(qs, k)
        ex = rng.integers(0, n, size=q).astype(np.uint32)
        ex[::3] = 0xFFFFFFFF
        idx_e, dist_e = st.so3_knn(qs, k, exclude=ex)
        i1, d1 = st.so3_nn1(qs)
    oidx, odist = orc.so3_knn(rows, qs, k)
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_array_equal(_bits(dist), _bits(odist))
    oidx, odist = orc.so3_knn(rows, qs, k, exclude=ex)
    np.testing.assert_array_equal(idx_e, oidx)
    np.testing.assert_array_equal(_bits(dist_e), _bits(odist))
    np.testing.assert_array_equal(idx[:, 0], i1)                   # the list starts at the 1-NN
    np.testing.assert_array_equal(_bits(dist[:, 0]), _bits(d1))


def test_so3_knn_plateaus_clamp_nan_and_duplicates(oxg, orc):
    plat = _plateau(0.3, width=5)
    cases = []
    # members of one acos plateau order by index, not by a; the k-th sits inside the plateau
    cases.append([plat[4], 0.1, plat[0], plat[3], 0.31, plat[1], plat[2], 0.29])
    cases.append([0.1, plat[-2], 0.2, plat[-1]])
    # next to a = 0 everything rounds to acos = pi/2
    cases.append([0.0, 1e-30, 1e-20, 1e-18, 5e-324, 0.5, 1e-200])
    # clamp region: distance 0 for every a > 1 - 1e-9 (+inf and values above 1 too)
    cases.append([0.5, 1.0 - 1e-10, 1.0, 0.9999999995, np.inf, 7.0, 1.0 - 1e-9,
                  np.nextafter(1.0 - 1e-9, 2.0)])
    # NaN never enters, wherever it sits
    cases.append([np.nan, 0.9, 1.0, np.nan, 0.3])
    cases.append([np.nan, np.nan])
    # duplicates of the k-th value
    cases.append([0.7, 0.2, 0.7, 0.7, 0.9, 0.7, 0.1, 0.7])
    cases.append([-0.7, 0.6, -0.6, 0.7])
    for xs in cases:
        rows = _x_rows(xs)
        for k in (1, 2, 3, 4, len(xs), len(xs) + 3):
            if k > 32:
                continue
            with _store(oxg, rows) as st:
                idx, dist = st.so3_knn(E1, k)
                idx_e, dist_e = st.so3_knn(E1, k, exclude=np.array([2], np.uint32))
            oidx, odist = orc.so3_knn(rows, E1, k)
            np.testing.assert_array_equal(idx, oidx, err_msg=f"{xs} k={k}")
            np.testing.assert_array_equal(_bits(dist), _bits(odist), err_msg=f"{xs} k={k}")
            oidx, odist = orc.so3_knn(rows, E1, k, exclude=np.array([2], np.uint32))
            np.testing.assert_array_equal(idx_e, oidx, err_msg=f"{xs} k={k} excl")
            np.testing.assert_array_equal(_bits(dist_e), _bits(odist), err_msg=f"{xs} k={k} excl")
    # a plateau spread over a large store that spans many slices and warps
    rng = np.random.default_rng(11)
    xs = rng.uniform(-0.29, 0.29, size=90001)
    for pos, v in zip((60000, 123, 40000, 89999, 7), (plat[4], -plat[3], plat[0], plat[2], plat[1])):
        xs[pos] = v
    rows = _x_rows(xs)
    rows[:, 1:] = rng.normal(size=(len(xs), 3))
    qs = np.vstack([E1, -E1])
    with _store(oxg, rows) as st:
        for k in (1, 3, 5, 6, 10):
            idx, dist = st.so3_knn(qs, k)
            oidx, odist = orc.so3_knn(rows, qs, k)
            np.testing.assert_array_equal(idx, oidx)
            np.testing.assert_array_equal(_bits(dist), _bits(odist))


def test_so3_knn_errors_and_empty(oxg):
    with oxg.StateStore(oxg.Space.real_vector(4)) as st:
        idx, dist = st.so3_knn(E1, 3)                               # empty store: padded rows
        assert (idx == 0xFFFFFFFF).all() and np.isinf(dist).all()
        st.append(_x_rows([0.5, 0.4]))
        with pytest.raises(oxg.OxgError):
            st.so3_knn(E1, 33)
        idx, dist = st.so3_knn(E1, 3)
        assert idx.tolist() == [[0, 1, 0xFFFFFFFF]] and np.isinf(dist[0, 2])
    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        st.append(np.zeros((2, 3)))
        with pytest.raises(oxg.OxgError):
            st.so3_knn(E1, 1)


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


def _so3_problem(rng, n_cones, n_samples, goal_bias=0.05):
    cones = np.hstack([_unit(rng, n_cones), rng.uniform(0.2, 0.5, size=(n_cones, 1))])
    start = np.array([0.0, 0.0, 0.0, 1.0])
    goal = _unit(rng, 1)[0]
    samples = _unit(rng, n_samples)
    samples[rng.random(n_samples) < goal_bias] = goal
    return cones, start, goal, samples


@pytest.mark.parametrize("seed,max_distance,n_cones", [(1, 0.5, 1), (2, 0.2, 3), (3, 0.05, 2), (4, 2.0, 6)])
def test_so3_rrt_chunked_loop_is_the_sequential_loop(oxg, orc, seed, max_distance, n_cones):
    """oxg_so3_rrt_solve runs speculative chunks; tree, parents, iteration count and outcome must
    be the one-sample-at-a-time loop's (orc_so3_rrt), bit for bit."""
    rng = np.random.default_rng(seed)
    cones, start, goal, samples = _so3_problem(rng, n_cones, 3000, goal_bias=0.0)
    while not orc.so3_is_valid(cones, start[None])[0]:
        cones, start, goal, samples = _so3_problem(rng, n_cones, 3000, goal_bias=0.0)
    c = oxg.So3Cones(cones)
    otree, opar, osolved, oiters = orc.so3_rrt(cones, start, samples, max_distance, RES, 4000)
    with _store(oxg, start[None]) as st:
        par, r = oxg.so3_rrt_solve(c, st, samples, max_distance, RES, 4000)
        rows = st.read(0, len(st))
    assert r["iters"] == oiters == 3000 and not r["solved"] and r["status"] == "samples_exhausted"
    np.testing.assert_array_equal(par, opar)
    np.testing.assert_array_equal(_bits(rows), _bits(otree))
    assert len(par) > 100
    assert r["ctas"] < r["iters"] / 4, r                            # it really ran in chunks


def test_so3_rrt_goal_continuation_and_capacity(oxg, orc):
    rng = np.random.default_rng(8)
    cones, start, goal, samples = _so3_problem(rng, 2, 4000, goal_bias=0.05)
    while not orc.so3_is_valid(cones, goal[None])[0] or not orc.so3_is_valid(cones, start[None])[0]:
        cones, start, goal, samples = _so3_problem(rng, 2, 4000, goal_bias=0.05)
    c = oxg.So3Cones(cones)
    otree, opar, osolved, oiters = orc.so3_rrt(cones, start, samples, 0.3, RES, 4000, goal, 0.1)
    assert osolved
    with _store(oxg, start[None]) as st:
        par, r = oxg.so3_rrt_solve(c, st, samples, 0.3, RES, 4000, goal_target=goal, goal_radius=0.1)
        rows = st.read(0, len(st))
    assert r["solved"] and r["iters"] == oiters and r["goal_node"][0] == len(opar) - 1
    np.testing.assert_array_equal(par, opar)
    np.testing.assert_array_equal(_bits(rows), _bits(otree))
    # the same stream in two calls (the second continues the first call's tree)
    cut = oiters // 2
    with _store(oxg, start[None]) as st:
        p1, r1 = oxg.so3_rrt_solve(c, st, samples[:cut], 0.3, RES, 4000, goal_target=goal,
                                   goal_radius=0.1)
        assert not r1["solved"] and r1["iters"] == cut
        p2, r2 = oxg.so3_rrt_solve(c, st, samples[cut:], 0.3, RES, 4000, goal_target=goal,
                                   goal_radius=0.1, parents=p1)
        rows = st.read(0, len(st))
    assert r2["solved"] and r2["iters"] == oiters - cut
    np.testing.assert_array_equal(p2, opar)
    np.testing.assert_array_equal(_bits(rows), _bits(otree))
    # capacity: the loop stops when the tree holds max_nodes states
    otree, opar, osolved, oiters = orc.so3_rrt(cones, start, samples, 0.3, RES, 50)
    with _store(oxg, start[None]) as st:
        par, r = oxg.so3_rrt_solve(c, st, samples, 0.3, RES, 50)
    assert r["status"] == "tree_full" and r["iters"] == oiters and len(par) == 50
    np.testing.assert_array_equal(par, opar)
    with oxg.StateStore(oxg.Space.real_vector(4)) as st:
        with pytest.raises(oxg.OxgError):
            oxg.so3_rrt_solve(c, st, samples, 0.3, RES, 50)         # no start state


@pytest.mark.parametrize("seed,max_distance,n_cones,goal_radius",
                         [(1, 0.5, 1, -1.0), (2, 0.2, 3, -1.0), (3, 0.05, 2, -1.0), (4, 1.0, 2, 0.05),
                          (5, 0.3, 4, 0.1)])
def test_so3_rrt_connect_chunked_loop_is_the_sequential_loop(oxg, orc, seed, max_distance, n_cones,
                                                             goal_radius):
    """Both trees, their parents, the joined nodes and the iteration count of the chunked
    RRT-Connect equal the sequential loop's (orc_so3_rrt_connect)."""
    rng = np.random.default_rng(100 + seed)
    while True:
        cones, start, goal, samples = _so3_problem(rng, n_cones, 2500, goal_bias=0.05)
        if orc.so3_is_valid(cones, np.vstack([start, goal])).all():
            break
    c = oxg.So3Cones(cones)
    gt = goal if goal_radius >= 0 else None
    # an unreachable "connection" keeps the loops running when no goal is set: forbid `reached`
    # by never letting the trees meet only if max_distance is small; otherwise compare as is
    o = orc.so3_rrt_connect(cones, start, goal, samples, max_distance, RES, 3000, gt, goal_radius)
    with _store(oxg, start[None]) as sa, _store(oxg, goal[None]) as sb:
        pa, pb, r = oxg.so3_rrt_connect_solve(c, sa, sb, samples, max_distance, RES, 3000,
                                              goal_target=gt, goal_radius=goal_radius)
        ra, rb = sa.read(), sb.read()
    assert r["solved"] == o["solved"] and r["iters"] == o["iters"]
    np.testing.assert_array_equal(pa, o["start_parents"])
    np.testing.assert_array_equal(pb, o["goal_parents"])
    np.testing.assert_array_equal(_bits(ra), _bits(o["start_tree"]))
    np.testing.assert_array_equal(_bits(rb), _bits(o["goal_tree"]))
    if o["solved"]:
        assert r["goal_node"] == o["conn"]
    assert r["iters"] <= 8 or r["ctas"] < r["iters"]
    # the same stream in two calls continues the same trees
    cut = max(1, o["iters"] // 2)
    with _store(oxg, start[None]) as sa, _store(oxg, goal[None]) as sb:
        pa1, pb1, r1 = oxg.so3_rrt_connect_solve(c, sa, sb, samples[:cut], max_distance, RES, 3000,
                                                 goal_target=gt, goal_radius=goal_radius)
        if not r1["solved"]:
            pa2, pb2, r2 = oxg.so3_rrt_connect_solve(c, sa, sb, samples[cut:], max_distance, RES, 3000,
                                                     goal_target=gt, goal_radius=goal_radius,
                                                     start_parents=pa1, goal_parents=pb1)
            assert r1["iters"] + r2["iters"] == o["iters"] and r2["solved"] == o["solved"]
            np.testing.assert_array_equal(pa2, o["start_parents"])
            np.testing.assert_array_equal(pb2, o["goal_parents"])
            np.testing.assert_array_equal(_bits(sb.read()), _bits(o["goal_tree"]))


def test_so3_rrt_connect_small_steps(oxg, orc):
    """Small steps: hundreds of iterations in which the node added last is usually the nearest one
    of the next scan, so chunks are cut early and often."""
    rng = np.random.default_rng(77)
    while True:
        cones, start, goal, samples = _so3_problem(rng, 3, 6000, goal_bias=0.0)
        if orc.so3_is_valid(cones, np.vstack([start, goal])).all():
            break
    c = oxg.So3Cones(cones)
    o = orc.so3_rrt_connect(cones, start, goal, samples, 0.02, RES, 1500)
    with _store(oxg, start[None]) as sa, _store(oxg, goal[None]) as sb:
        pa, pb, r = oxg.so3_rrt_connect_solve(c, sa, sb, samples, 0.02, RES, 1500)
        ra, rb = sa.read(), sb.read()
    assert r["solved"] == o["solved"] and r["iters"] == o["iters"]
    np.testing.assert_array_equal(pa, o["start_parents"])
    np.testing.assert_array_equal(pb, o["goal_parents"])
    np.testing.assert_array_equal(_bits(ra), _bits(o["start_tree"]))
    np.testing.assert_array_equal(_bits(rb), _bits(o["goal_tree"]))
    if not o["solved"]:
        assert r["status"] in ("tree_full", "samples_exhausted")
    assert r["ctas"] <= r["iters"], r      # tiny steps: the newest node is usually the nearest one
