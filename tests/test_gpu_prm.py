"""GPU parity: batched PRM construction (radius = reference-faithful, k-NN = config 4)."""
import numpy as np
import pytest

from conftest import uniform_states
from oxmpl_b200 import worlds

pytestmark = pytest.mark.gpu


def _panda_world(oxg, orc, n_caps=64):
    caps = worlds.capsule_world(4, worlds.PANDA_LO, worlds.PANDA_HI, n_caps)
    gw = oxg.World(7, capsules=caps)
    ow, keep = orc.world(7, capsules=caps)
    ext = oxg.real_vector_max_extent(list(zip(worlds.PANDA_LO, worlds.PANDA_HI)))
    return gw, ow, keep, oxg.motion_resolution(ext)


def test_prm_radius_matches_incremental_reference(oxg, orc):
    """The batched build must equal PRM::construct_roadmap (prm.rs:117-146) run sample by sample."""
    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    gw, ow, keep, res = _panda_world(oxg, orc)
    samples = uniform_states(4, 2500, worlds.PANDA_LO, worlds.PANDA_HI)
    radius = 1.6
    nodes, ooff, oidx, pairs, mcs = orc.prm_radius(ow, sp_o, radius, res, samples)
    ok = gw.is_valid(sp, samples).astype(bool)                 # prm.rs:123
    np.testing.assert_array_equal(samples[ok], nodes)
    with oxg.StateStore(sp) as st:
        st.append(samples[ok])
        off, idx, stats = oxg.prm_build_radius(gw, st, radius, res)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(idx, oidx)
        assert stats["n_candidates"] == mcs and stats["n_edges_valid"] == len(oidx)
        assert 0 < len(oidx) < mcs                              # some edges blocked, some free
        # sharded rows reproduce the same CSR pieces (multi-GPU unit of work)
        n = len(nodes)
        cut = n // 3
        off_a, idx_a, _ = oxg.prm_build_radius(gw, st, radius, res, 0, cut)
        off_b, idx_b, _ = oxg.prm_build_radius(gw, st, radius, res, cut, n)
        np.testing.assert_array_equal(np.concatenate([idx_a, idx_b]), oidx)
        np.testing.assert_array_equal(off_b + off_a[-1], ooff[cut:])
    gw.close()


@pytest.mark.parametrize("n_caps,scale,offset", [(256, 1.0, 0.0), (64, 1.0, 0.0), (300, 1e-12, 0.0),
                                                 (256, 1.0, 1.0e5), (500, 1.0, 0.0)])
def test_row_cull_never_changes_the_validity_table(oxg, orc, n_caps, scale, offset):
    """The edge phase rules obstacles out once per row (ball around the row's state, radius = its
    longest candidate edge) before the per-edge cull.  The validity table must be the one the
    per-edge-only path gives, and the oracle's, for thin, dense, tiny-scale and far-offset worlds
    (at 6000 states the rows' balls are wide: with 256+ capsules most rows overflow
    their lists and take the full pass, with 64 none does)."""
    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    lo, hi = np.asarray(worlds.PANDA_LO), np.asarray(worlds.PANDA_HI)
    caps = worlds.capsule_world(9, worlds.PANDA_LO, worlds.PANDA_HI, n_caps)
    caps[:, :14] = caps[:, :14] * scale + offset
    caps[:, 14] *= scale
    gw = oxg.World(7, capsules=caps)
    ow, keep = orc.world(7, capsules=caps)
    rows = uniform_states(45, 6000, worlds.PANDA_LO, worlds.PANDA_HI) * scale + offset
    rows[17] = rows[16]                                   # duplicate state: zero-length edge
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo * scale, hi * scale))))
    with oxg.StateStore(sp) as st:
        st.append(rows)
        oxg.tuning_set("row_cull", 0)
        try:
            idx0, dist0, valid0, stats0 = oxg.prm_build_knn(gw, st, 10, res)
        finally:
            oxg.tuning_set("row_cull", 1)
        idx1, dist1, valid1, stats1 = oxg.prm_build_knn(gw, st, 10, res)
    np.testing.assert_array_equal(idx0, idx1)
    np.testing.assert_array_equal(valid0, valid1)
    assert stats0["n_edges_valid"] == stats1["n_edges_valid"]
    # the radius build takes the same lists with the connection radius as every row's ball
    radius = float(np.sort(dist1[:, 3])[len(dist1) // 2])       # ~4 earlier neighbours per node
    with oxg.StateStore(sp) as st:
        st.append(rows)
        oxg.tuning_set("row_cull", 0)
        try:
            off0, nbr0, rs0 = oxg.prm_build_radius(gw, st, radius, res)
        finally:
            oxg.tuning_set("row_cull", 1)
        off1, nbr1, rs1 = oxg.prm_build_radius(gw, st, radius, res)
    np.testing.assert_array_equal(off0, off1)
    np.testing.assert_array_equal(nbr0, nbr1)
    assert rs1["n_candidates"] > 148 * 64 and 0 < rs1["n_edges_valid"] <= rs1["n_candidates"]
    oi, od, ov = orc.prm_knn(ow, sp_o, rows[:1500], 10, res)  # oracle on a prefix store ...
    with oxg.StateStore(sp) as st:
        st.append(rows[:1500])
        idx2, dist2, valid2, _ = oxg.prm_build_knn(gw, st, 10, res)
    np.testing.assert_array_equal(idx2, oi)
    np.testing.assert_array_equal(valid2, ov)
    assert 0 < int((valid1 == 1).sum()) < valid1.size
    gw.close()


@pytest.mark.parametrize("n,k", [(3000, 10), (9, 10), (513, 4)])
def test_prm_knn_matches_oracle(oxg, orc, n, k):
    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    gw, ow, keep, res = _panda_world(oxg, orc)
    rows = uniform_states(44, n, worlds.PANDA_LO, worlds.PANDA_HI)
    oi, od, ov = orc.prm_knn(ow, sp_o, rows, k, res)
    with oxg.StateStore(sp) as st:
        st.append(rows)
        idx, dist, valid, stats = oxg.prm_build_knn(gw, st, k, res)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        np.testing.assert_array_equal(valid, ov)
        assert stats["n_candidates"] == int((oi != 0xFFFFFFFF).sum())
        assert stats["n_edges_checked"] <= stats["n_candidates"]
        # row-range call (shard): same rows, mutual-pair dedup limited to the range
        a, b = n // 4, n // 2
        idx2, dist2, valid2, _ = oxg.prm_build_knn(gw, st, k, res, a, b)
        np.testing.assert_array_equal(idx2, oi[a:b])
        np.testing.assert_array_equal(valid2, ov[a:b])
    gw.close()


@pytest.mark.parametrize("n,k,shards", [(3001, 10, 2), (2000, 6, 4)])
def test_device_pointer_pipeline_sharded_on_one_gpu(oxg, orc, n, k, shards):
    """The multi-GPU PRM path (oxg_knn_rows_dev -> all-gather -> oxg_prm_edges_dev -> all-gather
    -> oxg_prm_resolve_dev, as bench.py drives it with NCCL) emulated on one GPU: every shard's
    rows are computed by separate calls into shared [n, k] device buffers (what the all-gathers
    produce), the result must equal the unsharded oracle roadmap, and every candidate edge must
    be motion-checked exactly once across the shards."""
    import ctypes as C

    import torch

    from oxmpl_b200 import sharding
    from oxmpl_b200._lib import check, lib

    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    gw, ow, keep, res = _panda_world(oxg, orc)
    rows = uniform_states(45, n, worlds.PANDA_LO, worlds.PANDA_HI)
    oi, od, ov = orc.prm_knn(ow, sp_o, rows, k, res)
    dev = torch.device("cuda", 0)
    with oxg.StateStore(sp) as st:
        st.append(rows)
        ext = torch.cuda.ExternalStream(st.stream(), device=dev)
        with torch.cuda.stream(ext):
            idx_all = torch.empty((n, k), dtype=torch.int32, device=dev)
            dist_all = torch.empty((n, k), dtype=torch.float64, device=dev)
            valid_all = torch.full((n, k), 7, dtype=torch.uint8, device=dev)
            counters = torch.zeros(2, dtype=torch.int64, device=dev)
        ext.synchronize()
        for r in range(shards):
            rb, re = sharding.row_range(n, r, shards)
            check(lib.oxg_knn_rows_dev(st.handle, rb, re, k, C.c_void_p(idx_all[rb:re].data_ptr()),
                                       C.c_void_p(dist_all[rb:re].data_ptr())))
        checked = 0
        for r in range(shards):
            rb, re = sharding.row_range(n, r, shards)
            check(lib.oxg_prm_edges_dev(gw.handle, st.handle, C.c_void_p(idx_all.data_ptr()), k,
                                        float(res), rb, re, C.c_void_p(valid_all.data_ptr()),
                                        C.c_void_p(counters.data_ptr())))
            st.sync()
            checked += int(counters.cpu()[0])
            own = sharding.owned_slots(oi, rb, re)                 # host mirror of the ownership rule
            assert int(counters.cpu()[0]) == int(own.sum())
        check(lib.oxg_prm_resolve_dev(st.handle, C.c_void_p(idx_all.data_ptr()), k,
                                      C.c_void_p(valid_all.data_ptr())))
        st.sync()
        np.testing.assert_array_equal(idx_all.cpu().numpy().view(np.uint32), oi)
        np.testing.assert_array_equal(dist_all.cpu().numpy().view(np.uint64), od.view(np.uint64))
        np.testing.assert_array_equal(valid_all.cpu().numpy(), ov)
        pairs = {(min(i, int(j)), max(i, int(j))) for i in range(n) for j in oi[i] if j != 0xFFFFFFFF}
        assert checked == len(pairs)
    gw.close()


def test_config4_full_size_roadmap_properties(oxg, orc):
    """BASELINE config 4 at FULL size (1 000 000 valid 7-D samples, k = 10, 256 capsules — the
    bench.py workload).  The oracle cannot rebuild 1e12 pairs, so the whole table is checked through
    size-independent properties, and a random sample of rows / edges against the oracle:
      * every listed distance equals the reference-order f64 distance recomputed from the rows;
      * rows are sorted by (distance, index), exclude themselves, hold k distinct neighbours;
      * mutual consistency: if j is listed for i at distance d and (d, i) beats j's k-th entry,
        then i is listed for j — with the same distance bits (the metric is bitwise symmetric);
      * an undirected edge gets the same motion verdict from both of its slots;
      * idempotence: a second build returns the same bits."""
    n, k, d = 1_000_000, 10, 7
    lo, hi = worlds.PANDA_LO, worlds.PANDA_HI
    caps = worlds.capsule_world(4, lo, hi, 256)
    sp, sp_o = oxg.Space.real_vector(d), orc.space(d)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))
    rng = np.random.default_rng(4)
    with oxg.World(d, capsules=caps) as gw, oxg.StateStore(sp, capacity=n) as st:
        rows = np.empty((0, d))
        while len(rows) < n:
            s = lo + (hi - lo) * rng.random((n // 2, d))
            rows = np.vstack([rows, s[gw.is_valid(sp, s).astype(bool)]])
        rows = np.ascontiguousarray(rows[:n])
        st.append(rows)
        idx, dist, valid, stats = oxg.prm_build_knn(gw, st, k, res)
        idx2, dist2, valid2, _ = oxg.prm_build_knn(gw, st, k, res)
        # the same table from the other neighbour algorithm: all 1e12 pairs scanned instead of
        # the bounding-box hierarchy (two independent kernels, one answer)
        oxg.tuning_set("knn_tree", 0)
        try:
            idx3, dist3, valid3, _ = oxg.prm_build_knn(gw, st, k, res)
        finally:
            oxg.tuning_set("knn_tree", -1)
    assert np.array_equal(idx, idx2) and np.array_equal(dist.view(np.uint64), dist2.view(np.uint64))
    assert np.array_equal(valid, valid2)
    assert np.array_equal(idx, idx3) and np.array_equal(dist.view(np.uint64), dist3.view(np.uint64))
    assert np.array_equal(valid, valid3)
    own = np.arange(n, dtype=np.uint32)[:, None]
    assert (idx != own).all() and (idx < n).all()
    # distances: reference order s = 0; s = s + (a_d - b_d)^2 ...; sqrt(s)   (no reassociation)
    a = rows[np.repeat(np.arange(n), k)]
    b = rows[idx.reshape(-1)]
    acc = np.zeros(n * k)
    for c in range(d):
        df = a[:, c] - b[:, c]
        acc = acc + df * df
    want = np.sqrt(acc).reshape(n, k)
    del a, b, acc
    assert np.array_equal(dist.view(np.uint64), want.view(np.uint64))
    # sorted by (distance, index); k distinct neighbours
    dd, ii = np.diff(dist, axis=1), np.diff(idx.astype(np.int64), axis=1)
    assert ((dd > 0) | ((dd == 0) & (ii > 0))).all()
    # mutual consistency against the neighbour's own k-th entry
    src = np.repeat(np.arange(n, dtype=np.int64), k)
    dst = idx.reshape(-1).astype(np.int64)
    dflat = dist.reshape(-1)
    kth_d, kth_i = dist[dst, k - 1], idx[dst, k - 1].astype(np.int64)
    must = (dflat < kth_d) | ((dflat == kth_d) & (src <= kth_i))
    listed = (idx[dst] == src[:, None].astype(np.uint32))
    assert (listed.any(axis=1) >= must).all()
    slot = listed.argmax(axis=1)
    both = listed.any(axis=1)
    assert np.array_equal(dist[dst[both], slot[both]].view(np.uint64), dflat[both].view(np.uint64))
    # the two slots of a mutual pair describe the same undirected edge: same verdict
    vflat = valid.reshape(-1)
    assert np.array_equal(valid[dst[both], slot[both]], vflat[both])
    assert 0.5 < vflat.mean() < 1.0 and stats["n_edges_checked"] > 5_000_000
    # no missed neighbour, mutual or not: for random pairs (i, j), whenever (d(i,j), j) sorts before
    # i's k-th entry, j must be listed for i
    pr = np.random.default_rng(3)
    pi = pr.integers(0, n, size=400_000)
    near = idx[pi, pr.integers(0, k, size=pi.size)].astype(np.int64)       # neighbours of neighbours:
    pj = idx[near, pr.integers(0, k, size=pi.size)].astype(np.int64)       # pairs that are close
    pj = np.where(pr.random(pi.size) < 0.25, pr.integers(0, n, size=pi.size), pj)
    keep_p = pi != pj
    pi, pj = pi[keep_p], pj[keep_p]
    acc = np.zeros(pi.size)
    for c in range(d):
        df = rows[pi, c] - rows[pj, c]
        acc = acc + df * df
    dij = np.sqrt(acc)
    beats = (dij < dist[pi, k - 1]) | ((dij == dist[pi, k - 1]) & (pj <= idx[pi, k - 1].astype(np.int64)))
    assert beats.sum() > 50_000
    assert (idx[pi[beats]] == pj[beats, None].astype(np.uint32)).any(axis=1).all()
    # a sample of rows and of edges against the oracle
    pick = np.sort(np.random.default_rng(1).choice(n, size=4096, replace=False)).astype(np.uint32)
    oi, od = orc.knn(sp_o, rows, rows[pick], k, exclude=pick)
    assert np.array_equal(idx[pick], oi)
    assert np.array_equal(dist[pick].view(np.uint64), od.view(np.uint64))
    ow, keep = orc.world(d, capsules=caps)
    e = np.random.default_rng(2).choice(n * k, size=50_000, replace=False)
    i_e, j_e = src[e], dst[e]
    ov, _ = orc.check_motion(ow, sp_o, rows[np.maximum(i_e, j_e)], rows[np.minimum(i_e, j_e)], res)
    assert np.array_equal(vflat[e], ov)


@pytest.mark.parametrize("concurrent", ["1", "0"])
def test_row_shards_of_the_full_build_equal_the_full_table(oxg, orc, concurrent, monkeypatch):
    """The multi-GPU build gives each rank a contiguous block of query rows of the 1M-node store
    (500 000 / 250 000 / 125 000 rows at 2 / 4 / 8 ranks).  Such a block is a few whole waves of CTAs
    plus a tail; the tail runs with the store cut into slices, concurrently with the whole waves on a
    second stream.  Every shard must reproduce the rows of the one-piece build bit for bit."""
    monkeypatch.setenv("OXG_KNN_TAIL_CONCURRENT", concurrent)
    n, k, d = 1_000_000, 10, 7
    lo, hi = worlds.PANDA_LO, worlds.PANDA_HI
    caps = worlds.capsule_world(4, lo, hi, 256)
    sp = oxg.Space.real_vector(d)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))
    rows = lo + (hi - lo) * np.random.default_rng(14).random((n, d))
    with oxg.World(d, capsules=caps) as gw, oxg.StateStore(sp, capacity=n) as st:
        st.append(rows)
        idx, dist, valid, _ = oxg.prm_build_knn(gw, st, k, res)
        for ranks in (2, 4, 8):
            for r in ((0, ranks - 1) if ranks > 2 else (0, 1)):
                rb, re = r * n // ranks, (r + 1) * n // ranks
                for rep in range(2):
                    i2, d2, v2, _ = oxg.prm_build_knn(gw, st, k, res, rb, re)
                    assert np.array_equal(i2, idx[rb:re]), (ranks, r, rep)
                    assert np.array_equal(d2.view(np.uint64), dist[rb:re].view(np.uint64))
                    assert np.array_equal(v2, valid[rb:re])
        # odd shard sizes around the wave boundaries (444 CTAs x 256 queries = 113 664 rows)
        for rows_q in (113_664, 113_665, 113_664 + 256 * 221, 113_664 + 256 * 222 + 1, 227_328 + 17):
            i2, d2, v2, _ = oxg.prm_build_knn(gw, st, k, res, 1000, 1000 + rows_q)
            assert np.array_equal(i2, idx[1000:1000 + rows_q]), rows_q
            assert np.array_equal(d2.view(np.uint64), dist[1000:1000 + rows_q].view(np.uint64))


def test_radius_prm_full_size_two_algorithms_one_roadmap(oxg, orc):
    """The reference's own PRM (radius mode, prm.rs:117-146) at 1 000 000 nodes: the hierarchy and the
    all-pairs radius kernels must emit the same CSR, and sampled rows must equal the oracle's
    `new_node.edges` (radius scan over the earlier nodes + check_motion(q_i, q_j))."""
    n, d, radius = 1_000_000, 7, 0.75
    lo, hi = worlds.PANDA_LO, worlds.PANDA_HI
    caps = worlds.capsule_world(4, lo, hi, 256)
    sp, sp_o = oxg.Space.real_vector(d), orc.space(d)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))
    rng = np.random.default_rng(4)
    with oxg.World(d, capsules=caps) as gw, oxg.StateStore(sp, capacity=n) as st:
        s = lo + (hi - lo) * rng.random((n + n // 3, d))
        rows = np.ascontiguousarray(s[gw.is_valid(sp, s).astype(bool)][:n])
        assert len(rows) == n
        st.append(rows)
        off, idx, stats = oxg.prm_build_radius(gw, st, radius, res)
        oxg.tuning_set("knn_tree", 0)
        try:
            off2, idx2, _ = oxg.prm_build_radius(gw, st, radius, res)
        finally:
            oxg.tuning_set("knn_tree", -1)
    np.testing.assert_array_equal(off, off2)
    np.testing.assert_array_equal(idx, idx2)
    assert stats["n_candidates"] > 4_000_000 and off[-1] > 0.95 * stats["n_candidates"]
    # rows are ascending and strictly below their own index
    starts = off[:-1].astype(np.int64)
    row_of = np.repeat(np.arange(n, dtype=np.int64), np.diff(off).astype(np.int64))
    assert (idx.astype(np.int64) < row_of).all()
    inner = np.ones(len(idx), dtype=bool)
    inner[starts[starts < len(idx)]] = False
    assert (np.diff(idx.astype(np.int64))[inner[1:]] > 0).all()
    # sampled rows against the oracle
    ow, keep = orc.world(d, capsules=caps)
    pick = np.sort(np.random.default_rng(5).choice(n, size=256, replace=False))
    ooff, oidx = orc.radius(sp_o, rows, rows[pick], radius, idx_limit=pick.astype(np.uint32))
    for t, i in enumerate(pick):
        cand = oidx[int(ooff[t]):int(ooff[t + 1])]
        if len(cand):
            ok, _ = orc.check_motion(ow, sp_o, np.repeat(rows[i:i + 1], len(cand), axis=0), rows[cand], res)
            cand = cand[ok.astype(bool)]
        np.testing.assert_array_equal(idx[int(off[i]):int(off[i + 1])], cand)
