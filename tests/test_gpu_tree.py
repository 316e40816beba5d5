"""GPU parity of the bounding-box-hierarchy k-NN (csrc/oxg_tree.cu) against the CPU oracle and
against the all-pairs scan kernels: same lists bit for bit, whatever the visiting order."""
import numpy as np
import pytest

from conftest import uniform_states

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["key-order", "partition-order"])
def tree_on(oxg, request):
    """Forces the hierarchy for every eligible k-NN, in both state orders: the space-filling key
    order and the balanced partition order (the 4-4-2 splits of csrc/oxg_tree.cu)."""
    oxg.tuning_set("knn_tree", 1)
    oxg.tuning_set("tree_kind", request.param)
    yield
    oxg.tuning_set("knn_tree", -1)
    oxg.tuning_set("tree_kind", -1)


def _same(a, b):
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1].view(np.uint64), b[1].view(np.uint64))


@pytest.mark.parametrize("d", [1, 2, 3, 6, 7, 8])
@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 1000, 1024, 1025, 33000])
def test_tree_knn_uniform(oxg, orc, tree_on, d, n):
    rows = uniform_states(300 + d, n, [-1.0] * d, [2.0] * d)
    queries = np.vstack([uniform_states(400 + d, 61, [-1.5] * d, [2.5] * d), rows[: min(n, 20)]])
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(rows)
        for k in (1, 3, 10, 32):
            _same(st.knn(queries, k), orc.knn(sp_o, rows, queries, k))
        _same(st.nn1(queries), orc.nn1(sp_o, rows, queries))
        ex = np.arange(min(n, 20), dtype=np.uint32)
        _same(st.knn(rows[: min(n, 20)], 4, exclude=ex), orc.knn(sp_o, rows, rows[: min(n, 20)], 4, exclude=ex))


@pytest.mark.parametrize("d", [2, 3, 7])
def test_tree_adversarial_ties(oxg, orc, tree_on, d):
    """Integer lattice + duplicates: every list is full of exact ties, the lowest index must win
    although the hierarchy visits the states in key order."""
    rng = np.random.default_rng(11 + d)
    base = rng.integers(0, 4, size=(6000, d)).astype(np.float64)
    near = base[:1500] + 0.0
    near[:, 0] = np.nextafter(near[:, 0] + 0.5, np.inf)
    rows = np.ascontiguousarray(np.vstack([base, base[rng.integers(0, 6000, 1500)], near]))
    rng.shuffle(rows)
    queries = np.vstack([rows[:300], rng.integers(0, 4, size=(100, d)).astype(np.float64) + 0.5])
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(rows)
        for k in (1, 10, 32):
            _same(st.knn(queries, k), orc.knn(sp_o, rows, queries, k))
        ex = np.arange(300, dtype=np.uint32)
        _same(st.knn(rows[:300], 7, exclude=ex), orc.knn(sp_o, rows, rows[:300], 7, exclude=ex))


def test_tree_sqrt_collapse_and_special_values(oxg, orc, tree_on):
    rng = np.random.default_rng(5)
    rows = rng.random((5000, 3))
    rows[17] = [np.nan, 0.5, 0.5]          # NaN states never enter a list
    rows[40] = [np.inf, 0.5, 0.5]          # inf is an ordinary (last) distance
    rows[41] = [-np.inf, 0.1, 0.2]
    rows[100:132] = rows[99]               # a whole leaf of duplicates
    queries = np.vstack([rng.random((50, 3)), [[np.nan, 0.1, 0.1]], [[np.inf, 0.0, 0.0]], rows[99:100]])
    sp_o = orc.space(3)
    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        st.append(rows)
        for k in (1, 5, 32):
            _same(st.knn(queries, k), orc.knn(sp_o, rows, queries, k))
    # fewer than k admissible states: padded rows
    few = rng.random((7, 2))
    with oxg.StateStore(oxg.Space.real_vector(2)) as st:
        st.append(few)
        q9 = np.random.default_rng(6).random((9, 2))
        _same(st.knn(q9, 10), orc.knn(orc.space(2), few, q9, 10))


@pytest.mark.parametrize("scale,offset", [(1.0, 1e7), (1.0, 1e15), (1e-30, 0.0), (1e-160, 0.0),
                                          (1e153, 0.0), (1e200, 0.0)])
def test_tree_extreme_scales(oxg, orc, tree_on, scale, offset):
    """The box bound must stay a lower bound when squares overflow, underflow or cancel."""
    rng = np.random.default_rng(21)
    rows = (rng.random((4000, 3)) + offset) * scale
    queries = np.vstack([(rng.random((40, 3)) + offset) * scale, rows[:40]])
    sp_o = orc.space(3)
    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        st.append(rows)
        for k in (1, 10):
            _same(st.knn(queries, k), orc.knn(sp_o, rows, queries, k))


def test_tree_clustered_and_degenerate_boxes(oxg, orc, tree_on):
    """Clusters, a constant coordinate (zero extent) and a line: skewed keys, same lists."""
    rng = np.random.default_rng(31)
    centres = rng.random((12, 4)) * 100.0
    rows = np.vstack([c + rng.normal(size=(1500, 4)) * 0.01 for c in centres])
    rows[:, 2] = 7.25                              # zero-extent dimension
    line = np.zeros((2000, 4))
    line[:, 0] = np.linspace(0.0, 1.0, 2000)
    rows = np.ascontiguousarray(np.vstack([rows, line]))
    rng.shuffle(rows)
    queries = np.vstack([rows[:200], rng.random((100, 4)) * 100.0])
    sp_o = orc.space(4)
    with oxg.StateStore(oxg.Space.real_vector(4)) as st:
        st.append(rows)
        _same(st.knn(queries, 10), orc.knn(sp_o, rows, queries, 10))


def test_tree_follows_the_store(oxg, orc, tree_on):
    """append / clear invalidate the hierarchy (it is rebuilt for the next batch)."""
    rng = np.random.default_rng(41)
    a, b = rng.random((3000, 5)), rng.random((2000, 5)) * 0.5
    q = rng.random((64, 5))
    sp_o = orc.space(5)
    with oxg.StateStore(oxg.Space.real_vector(5)) as st:
        st.append(a)
        _same(st.knn(q, 6), orc.knn(sp_o, a, q, 6))
        st.append(b)
        ab = np.vstack([a, b])
        _same(st.knn(q, 6), orc.knn(sp_o, ab, q, 6))
        st.clear()
        st.append(b)
        _same(st.knn(q, 6), orc.knn(sp_o, b, q, 6))


def test_tree_equals_all_pairs_scan_on_a_prm_table(oxg, orc):
    """200 000 x 200 000, D = 7, k = 10: the hierarchy and the all-pairs scan kernels give the
    same table; a sample of rows is compared with the oracle."""
    from oxmpl_b200 import worlds

    n = 200_000
    rows = uniform_states(4, n, worlds.PANDA_LO, worlds.PANDA_HI)
    with oxg.StateStore(oxg.Space.real_vector(7)) as st:
        st.append(rows)
        oxg.tuning_set("knn_tree", 1)
        try:
            oxg.tuning_set("tree_kind", 0)
            ki, kd = st.knn(rows, 10, exclude=np.arange(n, dtype=np.uint32))
            oxg.tuning_set("tree_kind", 1)
            ti, td = st.knn(rows, 10, exclude=np.arange(n, dtype=np.uint32))
        finally:
            oxg.tuning_set("tree_kind", -1)
            oxg.tuning_set("knn_tree", 0)
        try:
            bi, bd = st.knn(rows, 10, exclude=np.arange(n, dtype=np.uint32))
        finally:
            oxg.tuning_set("knn_tree", -1)
    np.testing.assert_array_equal(ti, bi)
    np.testing.assert_array_equal(td.view(np.uint64), bd.view(np.uint64))
    np.testing.assert_array_equal(ki, bi)
    np.testing.assert_array_equal(kd.view(np.uint64), bd.view(np.uint64))
    pick = np.random.default_rng(1).choice(n, 512, replace=False)
    oi, od = orc.knn(orc.space(7), rows, rows[pick], 10, exclude=pick.astype(np.uint32))
    np.testing.assert_array_equal(ti[pick], oi)
    np.testing.assert_array_equal(td[pick].view(np.uint64), od.view(np.uint64))


@pytest.mark.parametrize("d", [1, 2, 3, 7, 8])
@pytest.mark.parametrize("n", [20000, 33000])
def test_tree_radius_uniform(oxg, orc, tree_on, d, n):
    """The radius scans (rrt_star.rs:121-131, prm.rs:131-138) through the hierarchy: strict `<`,
    ascending indices, optional per-query index limits, distances bit for bit."""
    rows = uniform_states(500 + d, n, [-1.0] * d, [2.0] * d)
    queries = np.vstack([uniform_states(600 + d, 80, [-1.5] * d, [2.5] * d), rows[:40]])
    sp_o = orc.space(d)
    r = 3.0 * (200.0 / n) ** (1.0 / d)
    with oxg.StateStore(oxg.Space.real_vector(d)) as st:
        st.append(rows)
        off, idx, dist = st.radius(queries, r, want_dist=True)
        ooff, oidx, odist = orc.radius(sp_o, rows, queries, r, want_dist=True)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(idx, oidx)
        np.testing.assert_array_equal(dist.view(np.uint64), odist.view(np.uint64))
        assert off[-1] > 0
        lim = np.random.default_rng(d).integers(0, n, size=len(queries)).astype(np.uint32)
        off, idx = st.radius(queries, r, idx_limit=lim)
        ooff, oidx = orc.radius(sp_o, rows, queries, r, idx_limit=lim)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(idx, oidx)
        for bad in (0.0, -1.0, float("nan")):
            off, idx = st.radius(queries, bad)
            assert off[-1] == 0 and len(idx) == 0


def test_tree_radius_boundary_and_ties(oxg, orc, tree_on):
    """States exactly at distance r are excluded (strict `<`), one ulp inside are included; the
    lattice makes whole shells tie."""
    rng = np.random.default_rng(77)
    rows = rng.integers(0, 30, size=(40000, 3)).astype(np.float64)
    queries = np.vstack([rows[:100], rng.integers(0, 30, size=(60, 3)).astype(np.float64)])
    sp_o = orc.space(3)
    with oxg.StateStore(oxg.Space.real_vector(3)) as st:
        st.append(rows)
        for r in (1.0, np.nextafter(1.0, 2.0), float(np.sqrt(2.0)), float(np.nextafter(np.sqrt(2.0), 2.0)), 3.0):
            off, idx = st.radius(queries, r)
            ooff, oidx = orc.radius(sp_o, rows, queries, r)
            np.testing.assert_array_equal(off, ooff)
            np.testing.assert_array_equal(idx, oidx)


def test_radius_prm_through_the_hierarchy_equals_the_oracle(oxg, orc, tree_on):
    """Reference-faithful PRM rows (node i joins the earlier nodes within the radius whose motion is
    valid, prm.rs:131-138) with the neighbour phase answered by the hierarchy."""
    from oxmpl_b200 import worlds

    n, d = 24000, 7
    lo, hi = worlds.PANDA_LO, worlds.PANDA_HI
    caps = worlds.capsule_world(4, lo, hi, 64)
    sp, sp_o = oxg.Space.real_vector(d), orc.space(d)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))
    samples = lo + (hi - lo) * np.random.default_rng(9).random((n, d))
    ow, keep = orc.world(d, capsules=caps)
    o_rows, o_off, o_idx, _, _ = orc.prm_radius(ow, sp_o, 1.6, res, samples)
    with oxg.World(d, capsules=caps) as gw, oxg.StateStore(sp) as st:
        st.append(o_rows)
        off, idx, stats = oxg.prm_build_radius(gw, st, 1.6, res)
    np.testing.assert_array_equal(off, o_off)
    np.testing.assert_array_equal(idx, o_idx)
    assert off[-1] > 10000
