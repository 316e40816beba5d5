"""GPU parity: nn1 / knn / radius through the C ABI vs the CPU oracle (bit-exact)."""
import math

import numpy as np
import pytest

from conftest import uniform_states

pytestmark = pytest.mark.gpu


def _adversarial(seed, n, d):
    """Duplicates, an integer lattice (mass exact ties) and 1-ulp near-ties."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=(n, d)).astype(np.float64)      # lattice: many exact ties
    dup = base[rng.integers(0, n, size=n // 4)]                      # exact duplicates
    near = base[: n // 4] + 0.0
    near[:, 0] = np.nextafter(near[:, 0] + 0.5, np.inf)              # 1-ulp perturbations
    near2 = base[: n // 4] + 0.0
    near2[:, 0] = near2[:, 0] + 0.5
    pts = np.vstack([base, dup, near, near2])
    rng.shuffle(pts)
    return np.ascontiguousarray(pts)


@pytest.mark.parametrize("d", [1, 2, 3, 6, 7, 14, 16])
@pytest.mark.parametrize("n", [1, 5, 127, 128, 129, 1000, 20011])
def test_nn1_knn_radius_uniform(oxg, orc, d, n):
    rows = uniform_states(100 + d, n, [0.0] * d, [1.0] * d)
    queries = uniform_states(200 + d, 67, [0.0] * d, [1.0] * d)
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        assert st.append(rows) == 0
        assert len(st) == n
        np.testing.assert_array_equal(st.read(), rows)
        idx, dist = st.nn1(queries)
        oi, od = orc.nn1(sp_o, rows, queries)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        for k in (1, 3, 10):
            idx, dist = st.knn(queries, k)
            oi, od = orc.knn(sp_o, rows, queries, k)
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        r = 0.3 * np.sqrt(d)
        off, ridx, rdist = st.radius(queries, r, want_dist=True)
        ooff, oidx, odist = orc.radius(sp_o, rows, queries, r, want_dist=True)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(ridx, oidx)
        np.testing.assert_array_equal(rdist.view(np.uint64), odist.view(np.uint64))


@pytest.mark.parametrize("d", [2, 3, 7])
def test_adversarial_ties(oxg, orc, d):
    rows = _adversarial(7 + d, 2000, d)
    queries = np.vstack([rows[:200], _adversarial(99, 100, d)])
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(rows)
        idx, dist = st.nn1(queries)
        oi, od = orc.nn1(sp_o, rows, queries)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        idx, dist = st.knn(queries, 10)
        oi, od = orc.knn(sp_o, rows, queries, 10)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        # self-exclusion by index, not by zero distance (duplicates stay)
        ex = np.arange(200, dtype=np.uint32)
        idx, dist = st.knn(rows[:200], 5, exclude=ex)
        oi, od = orc.knn(sp_o, rows, rows[:200], 5, exclude=ex)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        for r in (0.5, 1.0, np.nextafter(1.0, 2.0), 1.5):
            off, ridx = st.radius(queries, r)
            ooff, oidx = orc.radius(sp_o, rows, queries, r)
            np.testing.assert_array_equal(off, ooff)
            np.testing.assert_array_equal(ridx, oidx)


def test_sqrt_collapse_tie(oxg, orc):
    """Two stored states whose squared distances differ by an ulp but whose sqrt is equal:
    the reference compares after sqrt, so the LOWER index must win (SURVEY hard part 1)."""
    rng = np.random.default_rng(3)
    pairs = []
    while len(pairs) < 64:
        a = rng.random(2) + 0.5
        b = a.copy()
        b[0] = np.nextafter(b[0], np.inf)
        da = a[0] * a[0] + a[1] * a[1]
        db = b[0] * b[0] + b[1] * b[1]
        if da != db and np.sqrt(da) == np.sqrt(db):
            pairs.append((a, b))
    sp_o = orc.space(2)
    q = np.zeros((1, 2))
    for a, b in pairs:
        rows = np.array([b, a, [5.0, 5.0]])      # the farther-in-d2 state has the LOWER index
        with oxg.StateStore(oxg.Space.real_vector(2)) as st:
            st.append(rows)
            idx, dist = st.nn1(q)
            oi, od = orc.nn1(sp_o, rows, q)
            assert oi[0] == 0                     # reference: tie after sqrt -> first index
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
            idx, dist = st.knn(q, 2)
            oi, od = orc.knn(sp_o, rows, q, 2)
            np.testing.assert_array_equal(idx, oi)
            off, ridx = st.radius(q, float(np.sqrt(b[0] * b[0] + b[1] * b[1])))
            ooff, oidx = orc.radius(sp_o, rows, q, float(np.sqrt(b[0] * b[0] + b[1] * b[1])))
            np.testing.assert_array_equal(off, ooff)
            np.testing.assert_array_equal(ridx, oidx)


def test_radius_idx_limit_and_empty(oxg, orc):
    d = 3
    rows = uniform_states(5, 3000, [0.0] * d, [1.0] * d)
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(rows)
        lim = np.arange(0, 3000, 7, dtype=np.uint32)
        q = rows[lim]
        off, ridx = st.radius(q, 0.2, idx_limit=lim)
        ooff, oidx = orc.radius(sp_o, rows, q, 0.2, idx_limit=lim)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(ridx, oidx)
        off, ridx = st.radius(q, 0.0)            # strict '<': nothing, not even self
        assert off[-1] == 0 and ridx.size == 0
        off, ridx = st.radius(q[:3], np.inf)
        assert off[-1] == 3 * 3000


def test_incremental_append_matches_tree_growth(oxg, orc):
    """append one state at a time (tree.push) and query in between, as RRT does."""
    d = 2
    rows = uniform_states(11, 300, [-10.0] * d, [10.0] * d)
    qs = uniform_states(12, 300, [-10.0] * d, [10.0] * d)
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d), capacity=4) as st:
        for i in range(300):
            assert st.append(rows[i]) == i
            if i % 17 == 0 or i < 5:
                idx, dist = st.nn1(qs[i])
                oi, od = orc.nn1(sp_o, rows[: i + 1], qs[i])
                assert idx[0] == oi[0] and dist[0] == od[0]
        st.clear()
        assert len(st) == 0
        st.append(rows[:10])
        idx, _ = st.nn1(qs[:5])
        oi, _ = orc.nn1(sp_o, rows[:10], qs[:5])
        np.testing.assert_array_equal(idx, oi)


@pytest.mark.parametrize("space_kind", ["so2", "se2", "rv20"])
def test_generic_spaces(oxg, orc, space_kind):
    rng = np.random.default_rng(21)
    if space_kind == "so2":
        sp, sp_o = oxg.Space.so2(), orc.space(0, 1)
        rows = rng.uniform(-4 * np.pi, 4 * np.pi, size=(777, 1))
        qs = rng.uniform(-4 * np.pi, 4 * np.pi, size=(50, 1))
        r = 0.3
    elif space_kind == "se2":
        sp, sp_o = oxg.Space.compound(2, 1, 1.0, 0.5), orc.space(2, 1, 1.0, 0.5)
        rows = np.hstack([rng.uniform(0, 100, (1500, 2)), rng.uniform(-np.pi, np.pi, (1500, 1))])
        qs = np.hstack([rng.uniform(0, 100, (60, 2)), rng.uniform(-np.pi, np.pi, (60, 1))])
        r = 8.0
    else:
        sp, sp_o = oxg.Space.real_vector(20), orc.space(20)
        rows = rng.random((900, 20))
        qs = rng.random((40, 20))
        r = 1.5
    with oxg.StateStore(sp) as st:
        st.append(rows)
        idx, dist = st.nn1(qs)
        oi, od = orc.nn1(sp_o, rows, qs)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        idx, dist = st.knn(qs, 7)
        oi, od = orc.knn(sp_o, rows, qs, 7)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        off, ridx, rdist = st.radius(qs, r, want_dist=True)
        ooff, oidx, odist = orc.radius(sp_o, rows, qs, r, want_dist=True)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(ridx, oidx)
        np.testing.assert_array_equal(rdist.view(np.uint64), odist.view(np.uint64))


def test_large_batch_two_queries_per_thread(oxg, orc):
    """Q large enough to take the 2-queries-per-thread kernel; checked on a sample."""
    d, n, Q = 7, 30000, 160000
    rows = uniform_states(31, n, [0.0] * d, [1.0] * d)
    qs = uniform_states(32, Q, [0.0] * d, [1.0] * d)
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(rows)
        idx, dist = st.knn(qs, 10)
        sel = np.r_[0:300, Q - 300:Q, 75000:75300]
        oi, od = orc.knn(sp_o, rows, qs[sel], 10)
        np.testing.assert_array_equal(idx[sel], oi)
        np.testing.assert_array_equal(dist[sel].view(np.uint64), od.view(np.uint64))
        # properties over the full batch: sorted by (dist, idx), no duplicates per row
        assert np.all(np.diff(dist, axis=1) >= 0)
        tie = np.diff(dist, axis=1) == 0
        assert np.all(np.diff(idx.astype(np.int64), axis=1)[tie] > 0)


def test_errors(oxg):
    from oxmpl_b200 import OxgError

    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        with pytest.raises(OxgError) as e:
            st.nn1(np.zeros((1, 3)))
        assert e.value.code == -7  # OXG_E_EMPTY
        st.append(np.zeros((4, 3)))
        with pytest.raises(OxgError):
            st.knn(np.zeros((1, 3)), 0)
        with pytest.raises(OxgError):
            st.knn(np.zeros((1, 3)), 1000)
    with pytest.raises(OxgError) as e:
        oxg.StateStore(oxg.Space(0, 0))
    assert e.value.code == -2  # OXG_E_DIM


def _scaled_sets(d, nq=64):
    """Inputs that stress the certified FP32 pre-filter: it must never change a result, only
    skip certain rejections (csrc/oxg_nn.cu, filter_threshold_dev)."""
    rng = np.random.default_rng(77 + d)
    u = rng.random((3000, d))
    q = rng.random((nq, d))
    mixed = np.array([10.0 ** rng.integers(-6, 7) for _ in range(d)])
    return {
        "offset_1e7": (1.0e7 + u, 1.0e7 + q, 0.4),          # all f32 images collide
        "offset_1e15": (1.0e15 + 64.0 * u, 1.0e15 + 64.0 * q, 30.0),
        "tiny_1e-30": (1e-30 * u, 1e-30 * q, 0.3e-30),
        "f32_subnormal_1e-42": (1e-42 * u, 1e-42 * q, 0.3e-42),
        "huge_1e200": (1e200 * u, 1e200 * q, 0.3e200),        # f32 overflow -> inf / NaN
        "f64_near_overflow_1e153": (1e153 * u, 1e153 * q, 0.3e153),
        "mixed_scales": (u * mixed, q * mixed, 0.3 * float(np.linalg.norm(mixed))),
        "negative_box": (-5.0 + 10.0 * u, -5.0 + 10.0 * q, 2.5),
    }


@pytest.mark.parametrize("nq", [64, 1100], ids=["diff_form", "dot_form"])
@pytest.mark.parametrize("d", [2, 7])
def test_fp32_prefilter_never_changes_results(oxg, orc, d, nq):
    """64 queries take the difference-form kernel, >= 1024 the dot-form kernel (whose bound is
    hopeless on the offset / mixed-scale sets: its warps must fall back, results unchanged)."""
    sp_o = orc.space(d)
    for name, (rows, queries, r) in _scaled_sets(d, nq).items():
        rows = np.ascontiguousarray(rows)
        queries = np.ascontiguousarray(queries)
        with oxg.StateStore(oxg.Space.real_vector(d)) as st:
            st.append(rows)
            idx, dist = st.nn1(queries)
            oi, od = orc.nn1(sp_o, rows, queries)
            np.testing.assert_array_equal(idx, oi, err_msg=name)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64), err_msg=name)
            idx, dist = st.knn(queries, 5)
            oi, od = orc.knn(sp_o, rows, queries, 5)
            np.testing.assert_array_equal(idx, oi, err_msg=name)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64), err_msg=name)
            off, ridx, rdist = st.radius(queries, r, want_dist=True)
            ooff, oidx, odist = orc.radius(sp_o, rows, queries, r, want_dist=True)
            np.testing.assert_array_equal(off, ooff, err_msg=name)
            np.testing.assert_array_equal(ridx, oidx, err_msg=name)
            np.testing.assert_array_equal(rdist.view(np.uint64), odist.view(np.uint64),
                                          err_msg=name)


def test_store_clear_resets_filter_scale(oxg, orc):
    """max |coordinate| is tracked per store for the filter bound; clear() must reset it and a
    later append of huge values must widen it again."""
    d = 3
    sp_o = orc.space(d)
    small = uniform_states(1, 500, [0.0] * d, [1.0] * d)
    huge = 1e12 + 1e3 * uniform_states(2, 500, [0.0] * d, [1.0] * d)
    q = uniform_states(3, 40, [0.0] * d, [1.0] * d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(huge)
        st.clear()
        st.append(small)
        idx, dist = st.knn(q, 3)
        oi, od = orc.knn(sp_o, small, q, 3)
        np.testing.assert_array_equal(idx, oi)
        st.append(huge)
        both = np.vstack([small, huge])
        idx, dist = st.knn(np.vstack([q, huge[:10]]), 3)
        oi, od = orc.knn(sp_o, both, np.vstack([q, huge[:10]]), 3)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))


@pytest.mark.parametrize("q_count", [1, 2, 3, 5, 8, 9, 27])
def test_few_query_stream_scan(oxg, orc, q_count):
    """Q <= 8 takes the state-per-thread streaming kernel (the planners' per-iteration query)."""
    for d, n in ((2, 1), (3, 300), (7, 70001), (14, 4097)):
        rows = uniform_states(300 + d, n, [0.0] * d, [1.0] * d)
        queries = uniform_states(400 + d, q_count, [0.0] * d, [1.0] * d)
        with oxg.StateStore(oxg.Space.real_vector(d)) as st:
            st.append(rows)
            idx, dist = st.nn1(queries)
            oi, od = orc.nn1(orc.space(d), rows, queries)
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
            assert st.last_scan_ms() > 0.0
    rows = _adversarial(5, 3000, 3)
    queries = rows[:q_count] + 0.0
    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        st.append(rows)
        idx, dist = st.nn1(queries)
        oi, od = orc.nn1(orc.space(3), rows, queries)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))


@pytest.mark.parametrize("q_count", [1, 2, 5, 8])
def test_few_query_stream_radius(oxg, orc, q_count):
    """Q <= 8 radius scans (find_neighbours / PRM start connection) take the state-per-thread
    two-pass kernel: ascending indices, idx_limit, distances, empty results."""
    for d, n in ((2, 1), (3, 777), (6, 70001), (14, 4097)):
        rows = uniform_states(500 + d, n, [0.0] * d, [1.0] * d)
        queries = uniform_states(600 + d, q_count, [0.0] * d, [1.0] * d)
        sp_o = orc.space(d)
        with oxg.StateStore(oxg.Space.real_vector(d)) as st:
            st.append(rows)
            for r in (0.0, 0.35 * np.sqrt(d), 10.0):
                off, idx, dist = st.radius(queries, r, want_dist=True)
                ooff, oidx, odist = orc.radius(sp_o, rows, queries, r, want_dist=True)
                np.testing.assert_array_equal(off, ooff)
                np.testing.assert_array_equal(idx, oidx)
                np.testing.assert_array_equal(dist.view(np.uint64), odist.view(np.uint64))
            lim = np.arange(q_count, dtype=np.uint32) * (n // 8)
            off, idx = st.radius(queries, 0.4 * np.sqrt(d), idx_limit=lim)
            ooff, oidx = orc.radius(sp_o, rows, queries, 0.4 * np.sqrt(d), idx_limit=lim)
            np.testing.assert_array_equal(off, ooff)
            np.testing.assert_array_equal(idx, oidx)
    rows = _adversarial(6, 3000, 2)
    with oxg.StateStore(oxg.Space.real_vector(2)) as st:
        st.append(rows)
        for r in (0.5, 1.0, np.nextafter(1.0, 2.0)):
            off, idx = st.radius(rows[:q_count], r)
            ooff, oidx = orc.radius(orc.space(2), rows, rows[:q_count], r)
            np.testing.assert_array_equal(off, ooff)
            np.testing.assert_array_equal(idx, oidx)


def test_randomised_shapes_against_oracle(oxg, orc):
    """Seeded random (dim, N, Q, k, radius) combinations, including k > N, k = 64, Q = 0,
    several store slices (small Q over a large N) and query batches that straddle the
    kernel-selection thresholds (8 / 1024 queries)."""
    rng = np.random.default_rng(2026)
    for trial in range(28):
        d = int(rng.integers(1, 17))
        n = int(rng.choice([1, 2, 31, 257, 1500, 40_000]))
        q = int(rng.choice([0, 1, 8, 9, 130, 1023, 1024, 2100]))
        k = int(rng.choice([1, 2, 10, 33, 64]))
        if q * n > 3e7:
            q = max(1, int(3e7 // n))
        lo = float(rng.choice([-3.0, 0.0, 100.0]))
        rows = lo + rng.random((n, d)) * float(rng.choice([1.0, 7.5]))
        queries = lo + rng.random((q, d)) * 7.5
        sp_o = orc.space(d)
        with oxg.StateStore(oxg.Space.real_vector(d)) as st:
            st.append(rows)
            msg = f"trial {trial}: d={d} n={n} q={q} k={k}"
            if q == 0:
                i0, d0 = st.knn(np.empty((0, d)), k)
                assert i0.shape == (0, k) and d0.shape == (0, k)
                off, idx = st.radius(np.empty((0, d)), 1.0)
                assert off.tolist() == [0] and len(idx) == 0
                continue
            idx, dist = st.knn(queries, k)
            oi, od = orc.knn(sp_o, rows, queries, k)
            np.testing.assert_array_equal(idx, oi, err_msg=msg)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64), err_msg=msg)
            idx, dist = st.nn1(queries)
            oi, od = orc.nn1(sp_o, rows, queries)
            np.testing.assert_array_equal(idx, oi, err_msg=msg)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64), err_msg=msg)
            r = float(np.median(od)) * float(rng.choice([0.5, 1.0, 2.0]))
            off, ridx, rdist = st.radius(queries, r, want_dist=True)
            ooff, oidx, odist = orc.radius(sp_o, rows, queries, r, want_dist=True)
            np.testing.assert_array_equal(off, ooff, err_msg=msg)
            np.testing.assert_array_equal(ridx, oidx, err_msg=msg)
            np.testing.assert_array_equal(rdist.view(np.uint64), odist.view(np.uint64), err_msg=msg)


@pytest.mark.parametrize("q_count,k", [(1, 10), (3, 2), (8, 16), (5, 7), (19, 10), (64, 3)])
def test_few_query_stream_knn(oxg, orc, q_count, k):
    """Q <= 8, 2 <= k <= 16 over >= 4096 states: state-per-thread k-NN with per-CTA partial lists
    and the parallel (dist, idx) selection; ties, self-exclusion and k > (states per slice)."""
    for d, n in ((2, 4096), (7, 70001), (14, 9000)):
        rows = uniform_states(700 + d, n, [0.0] * d, [1.0] * d)
        queries = np.vstack([uniform_states(800 + d, q_count, [0.0] * d, [1.0] * d)])
        sp_o = orc.space(d)
        with oxg.StateStore(oxg.Space.real_vector(d)) as st:
            st.append(rows)
            idx, dist = st.knn(queries, k)
            oi, od = orc.knn(sp_o, rows, queries, k)
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
            ex = np.arange(q_count, dtype=np.uint32) * 5
            idx, dist = st.knn(rows[ex], k, exclude=ex)
            oi, od = orc.knn(sp_o, rows, rows[ex], k, exclude=ex)
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
    rows = np.vstack([_adversarial(8, 3000, 3), _adversarial(9, 3000, 3)])   # mass exact ties
    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        st.append(rows)
        idx, dist = st.knn(rows[:q_count], k)
        oi, od = orc.knn(orc.space(3), rows, rows[:q_count], k)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))


@pytest.mark.parametrize("offset,delta", [(0.0, 1e-9), (10.0, 1e-7), (1000.0, 1e-6), (3.0, 1e-15)])
@pytest.mark.parametrize("nq", [40, 1100], ids=["diff_form", "dot_form"])
def test_prefilter_at_the_threshold(oxg, orc, offset, delta, nq):
    """Every query sits at the centre of a thin shell of stored states whose distances differ by
    relative `delta` (down to a few ulps), and k is chosen so that the k-th neighbour lies in
    the middle of its shell: the admission threshold U cuts through a cloud of near-equal d2
    values, which is where a filter that rejected too eagerly (or a compare-before-sqrt) would
    pick different indices.  Offsets make the FP32 images of states and queries collide."""
    d, per = 7, 48
    rng = np.random.default_rng(int(offset) + nq)
    centres = offset + rng.random((nq, d)) * 4.0
    u = rng.normal(size=(nq, per, d))
    u /= np.linalg.norm(u, axis=2, keepdims=True)
    radii = 0.25 * (1.0 + (np.arange(per) - per // 2) * delta)
    rows = (centres[:, None, :] + u * radii[None, :, None]).reshape(-1, d)
    perm = rng.permutation(len(rows))
    rows = np.ascontiguousarray(rows[perm])
    sp_o = orc.space(d)
    k = per // 2
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(rows)
        idx, dist = st.knn(centres, k)
        oi, od = orc.knn(sp_o, rows, centres, k)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        i1, d1 = st.nn1(centres)
        o1, od1 = orc.nn1(sp_o, rows, centres)
        np.testing.assert_array_equal(i1, o1)
        np.testing.assert_array_equal(d1.view(np.uint64), od1.view(np.uint64))
        off, ridx = st.radius(centres, 0.25)
        ooff, oidx = orc.radius(sp_o, rows, centres, 0.25)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(ridx, oidx)


@pytest.mark.parametrize("q_count,k", [(70, 1), (130, 1), (130, 4), (300, 10)])
def test_medium_batches_on_a_large_store(oxg, orc, q_count, k):
    """On stores of >= 4e6 states, batches of 65..1023 queries take the dot-form kernel (one query
    per thread, the store cut into slices) instead of the difference-form one."""
    d, n = 2, 4_200_000
    rows = uniform_states(900, n, [0.0] * d, [1.0] * d)
    rows[1_000_000] = rows[3]                                    # a far-apart duplicate pair
    queries = uniform_states(901 + q_count, q_count, [0.0] * d, [1.0] * d)
    queries[0] = rows[3]
    with oxg.StateStore(oxg.Space.real_vector(d), capacity=n) as st:
        st.append(rows)
        idx, dist = st.knn(queries, k)
        oi, od = orc.knn(orc.space(d), rows, queries, k)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        if k == 1:
            idx1, dist1 = st.nn1(queries)
            o1, od1 = orc.nn1(orc.space(d), rows, queries)
            np.testing.assert_array_equal(idx1, o1)
            np.testing.assert_array_equal(dist1.view(np.uint64), od1.view(np.uint64))


def test_so2_remainder_is_exact_on_adversarial_angles(oxg, orc):
    """The device computes `fmod` with one FMA from a corrected quotient (oxg_math.cuh
    fmod_exact_dev) instead of libdevice's division loop: every pairwise SO2 distance — all k = N
    neighbours of every query — must carry the oracle's bit pattern, for angles at and next to the
    multiples of pi and 2*pi, huge, tiny, signed zeros and non-finite values."""
    pi = math.pi
    base = [0.0, -0.0, pi, -pi, 2 * pi, -2 * pi, 3 * pi, -3 * pi, 4 * pi, pi / 2, -pi / 2, 1e-300,
            -1e-300, 5e-324, 1.0, -1.0, 100.0, -100.0, 1e6, -1e6, 1e15, -1e15, 2.0 ** 52 * 2 * pi,
            2.0 ** 53 * 2 * pi, 1e300, -1e300, 710.0 * pi, 12345.678]
    vals = []
    for v in base:
        vals.append(v)
        up = down = v
        for _ in range(3):
            up, down = np.nextafter(up, np.inf), np.nextafter(down, -np.inf)
            vals += [up, down]
    rng = np.random.default_rng(3)
    vals += list(rng.uniform(-50, 50, size=60)) + list(rng.uniform(-1e9, 1e9, size=40))
    vals = np.array(vals, dtype=np.float64)
    sp, sp_o = oxg.Space.so2(), orc.space(0, 1)
    qs = np.concatenate([vals, [np.inf, -np.inf, np.nan]])[:, None]
    for lo in range(0, len(vals), 64):
        rows = vals[lo:lo + 64, None]
        with oxg.StateStore(sp) as st:
            st.append(rows)
            k = len(rows)
            idx, dist = st.knn(qs, k)
            oi, od = orc.knn(sp_o, rows, qs, k)
            np.testing.assert_array_equal(idx, oi)
            np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
            i1, d1 = st.nn1(qs)
            o1, od1 = orc.nn1(sp_o, rows, qs)
            np.testing.assert_array_equal(i1, o1)
            nan = np.isnan(od1)                      # NaN payloads are not part of the contract
            np.testing.assert_array_equal(np.isnan(d1), nan)
            np.testing.assert_array_equal(d1[~nan].view(np.uint64), od1[~nan].view(np.uint64))
    # interpolation / normalise through check_motion over SO2 with an interval obstacle
    w_o, keep = orc.world(0, so2_intervals=[(0, 1.0, 1.3)])
    frm = rng.uniform(-20, 20, size=(400, 1))
    to = rng.uniform(-20, 20, size=(400, 1))
    frm[:20, 0] = vals[:20]
    to[:20, 0] = vals[20:40]
    res = math.pi * 0.05 * 0.1
    with oxg.World(0, so2_intervals=[(0, 1.0, 1.3)]) as gw:
        valid, first = gw.check_motion(sp, frm, to, res, want_first_invalid=True)
    ovalid, ofirst = orc.check_motion(w_o, sp_o, frm, to, res)
    np.testing.assert_array_equal(valid, ovalid)
    np.testing.assert_array_equal(first, ofirst)


def test_nn1_nan_first_state_keeps_index_zero(oxg, orc):
    """rrt.rs:187-196 starts from min = distance(tree[0], q): when that is NaN nothing is `<` it and
    the reference answers (0, NaN) even if other states are finite; a NaN elsewhere never wins."""
    for d in (1, 3, 7, 20):
        sp, sp_o = oxg.Space.real_vector(d), orc.space(d)
        rng = np.random.default_rng(d)
        rows = rng.random((5000, d))
        qs = rng.random((40, d))
        rows[0, d - 1] = np.nan
        rows[77, 0] = np.nan
        with oxg.StateStore(sp) as st:
            st.append(rows)
            for q in (qs[:1], qs[:9], qs):
                idx, dist = st.nn1(q)
                oi, od = orc.nn1(sp_o, rows, q)
                assert (oi == 0).all() and np.isnan(od).all()
                np.testing.assert_array_equal(idx, oi)
                assert np.isnan(dist).all()
        rows[0] = rng.random(d)                       # only a later state is NaN now
        qs[3, 0] = np.inf                              # inf - finite = inf: not NaN, a plain loser
        qs[5, 0] = np.nan
        with oxg.StateStore(sp) as st:
            st.append(rows)
            idx, dist = st.nn1(qs)
            oi, od = orc.nn1(sp_o, rows, qs)
            np.testing.assert_array_equal(idx, oi)
            nan = np.isnan(od)
            assert nan[5] and not nan[0]
            np.testing.assert_array_equal(np.isnan(dist), nan)
            np.testing.assert_array_equal(dist[~nan].view(np.uint64), od[~nan].view(np.uint64))
