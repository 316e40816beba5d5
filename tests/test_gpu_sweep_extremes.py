"""BASELINE config 5's extreme points as driver-run GPU tests (round 1 ran them only through
tests/probes/sweep.py): the largest store (1e7 x 14-D, 1.12 GB, HBM-resident) with 1 and 100
queries, and the largest batch (1e6 x 1e6, D = 2) — nn1 / knn / radius against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _same(a, b):
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1].view(np.uint64), b[1].view(np.uint64))


def test_largest_store_1e7_x_14d_few_queries(oxg, orc):
    n, d = 10_000_000, 14
    rng = np.random.default_rng(5)
    rows = rng.random((n, d))
    rows[123_456] = rows[7]                      # an exact duplicate far apart in index: tie -> lower index
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d), capacity=n) as st:
        for a in range(0, n, 2_000_000):
            st.append(rows[a:a + 2_000_000])
        assert len(st) == n
        for nq in (1, 100):
            q = np.vstack([rng.random((nq - 1, d)), rows[7:8]]) if nq > 1 else rows[7:8] + 0.0
            _same(st.nn1(q), orc.nn1(sp_o, rows, q))
            _same(st.knn(q, 10), orc.knn(sp_o, rows, q, 10))
            r = 0.62
            off, idx, dist = st.radius(q, r, want_dist=True)
            ooff, oidx, odist = orc.radius(sp_o, rows, q, r, want_dist=True)
            np.testing.assert_array_equal(off, ooff)
            np.testing.assert_array_equal(idx, oidx)
            np.testing.assert_array_equal(dist.view(np.uint64), odist.view(np.uint64))
            assert off[-1] > 0


def test_largest_batch_1e6_x_1e6_2d(oxg, orc):
    """1e12 pairs: the whole 1e6 x 10 table through size-independent properties, a sample of rows
    against the oracle, and the 1-NN of 1e6 external queries likewise."""
    n, d, k = 1_000_000, 2, 10
    rng = np.random.default_rng(55)
    rows = rng.random((n, d))
    rows[500_000:500_064] = rows[10:74]          # duplicates: ties by index inside the lists
    sp_o = orc.space(d)
    with oxg.StateStore(oxg.Space.real_vector(d), capacity=n) as st:
        st.append(rows)
        idx, dist = st.knn(rows, k, exclude=np.arange(n, dtype=np.uint32))
        q = rng.random((n, d))
        i1, d1 = st.nn1(q)
    own = np.arange(n, dtype=np.uint32)[:, None]
    assert (idx != own).all() and (idx < n).all()
    a = rows[np.repeat(np.arange(n), k)]
    b = rows[idx.reshape(-1)]
    acc = np.zeros(n * k)
    for c in range(d):
        df = a[:, c] - b[:, c]
        acc = acc + df * df
    assert np.array_equal(dist.view(np.uint64), np.sqrt(acc).reshape(n, k).view(np.uint64))
    dd, ii = np.diff(dist, axis=1), np.diff(idx.astype(np.int64), axis=1)
    assert ((dd > 0) | ((dd == 0) & (ii > 0))).all()
    pick = np.sort(rng.choice(n, size=2048, replace=False)).astype(np.uint32)
    pick[:64] = np.arange(500_000, 500_064)
    _same((idx[pick], dist[pick]), orc.knn(sp_o, rows, rows[pick], k, exclude=pick))
    # 1-NN of external queries: distance recomputed in reference order, no closer state in a sample
    acc = np.zeros(n)
    for c in range(d):
        df = rows[i1, c] - q[:, c]               # distance(stored, query) — rrt.rs:191
        acc = acc + df * df
    assert np.array_equal(d1.view(np.uint64), np.sqrt(acc).view(np.uint64))
    p2 = rng.choice(n, size=2048, replace=False)
    _same((i1[p2], d1[p2]), orc.nn1(sp_o, rows, q[p2]))
