"""GPU parity of the PRM query phase (oxg_prm_solve / oxg_prm_solve_knn, csrc/oxg_query.cu) with the
oracle's restatement of PRM::solve (prm.rs:227-307): same status, same path — the breadth-first
search must assign the reference's parents, not just find some shortest path."""
import numpy as np
import pytest

from oxmpl_b200 import worlds

pytestmark = pytest.mark.gpu

WALL = np.array([[4.75, 2.0, 5.25, 8.0]])        # oxmpl/tests/prm_rvss_tests.rs: x = 5 +- 0.25, y in [2, 8]


def _wall_roadmap(oxg, orc, n_samples, seed):
    lo, hi = [0.0, 0.0], [10.0, 10.0]
    sp, sp_o = oxg.Space.real_vector(2), orc.space(2)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))
    gw = oxg.World(2, aabbs=WALL)
    ow, keep = orc.world(2, aabbs=WALL)
    rng = np.random.default_rng(seed)
    s = rng.random((n_samples, 2)) * 10.0
    rows = np.ascontiguousarray(s[gw.is_valid(sp, s).astype(bool)])
    st = oxg.StateStore(sp)
    st.append(rows)
    off, idx, _ = oxg.prm_build_radius(gw, st, 0.5, res)
    return gw, ow, keep, sp_o, st, rows, off, idx, res


def test_wall_world_paths_equal_the_reference_bfs(oxg, orc):
    """The reference's own PRM fixture (PRM::new(5.0, 0.5), start (1,5), goal (9,5) r 0.5) and a
    batch of other start / goal pairs on the same roadmap (multi-query PRM, prm.rs:88-90)."""
    gw, ow, keep, sp_o, st, rows, off, idx, res = _wall_roadmap(oxg, orc, 9000, 7)
    rng = np.random.default_rng(8)
    pairs = [((1.0, 5.0), (9.0, 5.0), 0.5)]
    pairs += [(tuple(rng.random(2) * 10.0), tuple(rng.random(2) * 10.0), 0.4) for _ in range(24)]
    solved = 0
    for start, goal, gr in pairs:
        status, path, info = oxg.prm_solve(gw, st, off, idx, 0.5, res, start, goal, gr)
        o_status, o_path = orc.prm_solve(ow, sp_o, rows, off, idx, 0.5, res, start, goal, gr)
        assert status == o_status, (start, goal)
        np.testing.assert_array_equal(path, o_path)
        if status == 0:
            solved += 1
            assert info["goal_node"] == path[-1] and info["path_len"] == len(path)
            assert np.linalg.norm(rows[path[-1]] - np.asarray(goal)) <= gr
            assert np.linalg.norm(rows[path[0]] - np.asarray(start)) < 0.5
            hops = np.linalg.norm(np.diff(rows[path], axis=0), axis=1)
            assert (hops < 0.5).all()                      # consecutive nodes are roadmap edges
    assert solved >= 10 and pairs[0] is not None
    # the wall fixture itself must be solved around the wall
    status, path, _ = oxg.prm_solve(gw, st, off, idx, 0.5, res, (1.0, 5.0), (9.0, 5.0), 0.5)
    assert status == 0 and len(path) > 10
    st.close()
    gw.close()


def test_query_error_statuses(oxg, orc):
    gw, ow, keep, sp_o, st, rows, off, idx, res = _wall_roadmap(oxg, orc, 3000, 9)
    cases = [((5.0, 5.0), (9.0, 5.0), 0.5, 2),       # start inside the wall: InvalidStartState
             ((1.0, 5.0), (9.0, 5.0), 1e-9, 1),      # no roadmap node satisfies the goal
             ((1.0, 5.0), (5.0, 5.0), 0.2, 1)]       # goal region inside the wall: no goal nodes
    for start, goal, gr, want in cases:
        status, path, _ = oxg.prm_solve(gw, st, off, idx, 0.5, res, start, goal, gr)
        o_status, o_path = orc.prm_solve(ow, sp_o, rows, off, idx, 0.5, res, start, goal, gr)
        assert status == o_status == want and len(path) == 0 == len(o_path)
    # a start with nothing within the connection radius: NoSolutionFound
    status, path, info = oxg.prm_solve(gw, st, off, idx, 1e-6, res, (1.0, 5.0), (9.0, 5.0), 0.5)
    assert status == 1 and info["n_start_connections"] == 0
    # an empty roadmap: UnsampledStateSpace (prm.rs:239-241)
    with oxg.StateStore(oxg.Space.real_vector(2)) as empty:
        status, path, _ = oxg.prm_solve(gw, empty, np.zeros(1, np.uint64), np.zeros(0, np.uint32), 0.5, res,
                                        (1.0, 5.0), (9.0, 5.0), 0.5)
        assert status == 3
    st.close()
    gw.close()


def test_knn_roadmap_query_7d(oxg, orc):
    """Config-4-shaped roadmap (7-D, capsules, k = 10): the k-NN table's valid edges, symmetrised,
    searched on the device; the oracle gets the same roadmap as lower-index lists."""
    n, k, d = 30_000, 10, 7
    lo, hi = worlds.PANDA_LO, worlds.PANDA_HI
    caps = worlds.capsule_world(4, lo, hi, 64)
    sp, sp_o = oxg.Space.real_vector(d), orc.space(d)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))
    rng = np.random.default_rng(4)
    with oxg.World(d, capsules=caps) as gw, oxg.StateStore(sp) as st:
        s = lo + (hi - lo) * rng.random((2 * n, d))
        rows = np.ascontiguousarray(s[gw.is_valid(sp, s).astype(bool)][:n])
        st.append(rows)
        idx, dist, valid, _ = oxg.prm_build_knn(gw, st, k, res)
        ii = np.repeat(np.arange(n), k)
        jj = idx.reshape(-1).astype(np.int64)
        ok = valid.reshape(-1) == 1
        a, b = np.maximum(ii[ok], jj[ok]), np.minimum(ii[ok], jj[ok])
        und = np.unique(np.stack([a, b], axis=1), axis=0)          # sorted by (higher, lower)
        off = np.zeros(n + 1, dtype=np.uint64)
        np.add.at(off, und[:, 0] + 1, 1)
        off = np.cumsum(off).astype(np.uint64)
        lower = und[:, 1].astype(np.uint32)
        ow, keep = orc.world(d, capsules=caps)
        solved = 0
        for t in range(12):
            start = lo + (hi - lo) * rng.random(d)
            goal = rows[rng.integers(0, n)] + 0.01
            status, path, info = oxg.prm_solve_knn(gw, st, idx, valid, 1.2, res, start, goal, 0.6)
            o_status, o_path = orc.prm_solve(ow, sp_o, rows, off, lower, 1.2, res, start, goal, 0.6)
            assert status == o_status
            np.testing.assert_array_equal(path, o_path)
            solved += status == 0
            if status == 0:
                assert info["n_adjacency"] == 2 * len(und)
        assert solved >= 6
