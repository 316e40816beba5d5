"""GPU parity: the fused planner kernels (oxg_rrt_solve / oxg_rrt_connect_solve /
oxg_rrt_star_solve) against the oracle's sample-stream-driven restatements of RRT::solve
(rrt.rs:158-227), RRTConnect::solve (rrt_connect.rs:227-309) and RRTStar::solve
(rrt_star.rs:171-290): identical trees (state bit patterns), parents, costs and outcomes."""
import math
import os

import numpy as np
import pytest

from oxmpl_b200 import worlds

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _res(oxg, orc, lo, hi):
    return oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))


@pytest.fixture(params=[0, 1, 7, 148], ids=lambda c: f"ctas{c or 'auto'}")
def ctas(request, monkeypatch):
    """The grid size must not change any result: 1 CTA (no grid barrier), a few, all SMs."""
    if request.param:
        monkeypatch.setenv("OXG_PLAN_CTAS", str(request.param))
    else:
        monkeypatch.delenv("OXG_PLAN_CTAS", raising=False)
    return request.param


def test_config1_readme_rrt_identical_tree(oxg, orc, ctas):
    """BASELINE config 1: README RRT world (README.md:97-150), fixed sample stream."""
    c = worlds.CONFIG1
    sp, sp_o = oxg.Space.real_vector(2), orc.space(2)
    res = _res(oxg, orc, c["lo"], c["hi"])
    assert res == 0.14142135623730953
    samples = worlds.sample_stream(1, 20000, c["lo"], c["hi"], c["goal_bias"], c["goal"])
    ow, keep = orc.world(2, spheres=c["spheres"])
    max_nodes = 20000
    t_o, p_o, solved_o, it_o = orc.rrt(ow, sp_o, c["start"], c["goal"], c["goal_radius"],
                                       c["max_distance"], res, samples, max_nodes)
    assert solved_o
    with oxg.World(2, spheres=c["spheres"]) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([c["start"]]))
        parents, r = oxg.rrt_solve(gw, st, samples, c["max_distance"], res, max_nodes,
                                   goal_target=c["goal"], goal_radius=c["goal_radius"])
        assert r["solved"] and r["status"] == "solved"
        assert r["iters"] == it_o and r["n_nodes"][0] == len(t_o)
        np.testing.assert_array_equal(_bits(st.read()), _bits(t_o))
        np.testing.assert_array_equal(parents, p_o)
        assert r["goal_node"][0] == len(t_o) - 1
        # the path the reference would reconstruct (rrt.rs:118-128) is valid and reaches the goal
        path, i = [], len(t_o) - 1
        while i >= 0:
            path.append(t_o[i])
            i = parents[i]
        path = np.array(path[::-1])
        assert np.array_equal(path[0], c["start"])
        assert math.dist(path[-1], c["goal"]) <= c["goal_radius"]
        assert gw.is_valid(sp, path).all()
        # the grown tree is a normal store: later NN queries see every appended state
        q = samples[:50]
        i1, d1 = st.nn1(q)
        oi, od = orc.nn1(sp_o, t_o, q)
        np.testing.assert_array_equal(i1, oi)
        np.testing.assert_array_equal(_bits(d1), _bits(od))


def test_rrt_wall_fixture_and_resume(oxg, orc, ctas):
    """oxmpl/tests/rrt_rvss_tests.rs world (wall AABB, inclusive faces), RRT(0.5, 0.0); the
    stream is fed in two calls and must give the tree of one call."""
    c = worlds.WALL
    sp, sp_o = oxg.Space.real_vector(2), orc.space(2)
    res = _res(oxg, orc, c["lo"], c["hi"])
    samples = worlds.sample_stream(11, 6000, c["lo"], c["hi"], 0.0, c["goal"])
    ow, keep = orc.world(2, aabbs=c["aabbs"])
    t_o, p_o, solved_o, it_o = orc.rrt(ow, sp_o, c["start"], c["goal"], c["goal_radius"], 0.5, res,
                                       samples, 6001)
    with oxg.World(2, aabbs=c["aabbs"]) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([c["start"]]))
        cut = min(150, it_o - 1)
        _, r1 = oxg.rrt_solve(gw, st, samples[:cut], 0.5, res, 6001, goal_target=c["goal"],
                              goal_radius=c["goal_radius"])
        assert r1["status"] == "samples_exhausted" and not r1["solved"]
        parents, r2 = oxg.rrt_solve(gw, st, samples[cut:], 0.5, res, 6001, goal_target=c["goal"],
                                    goal_radius=c["goal_radius"])
        assert r2["solved"] == solved_o and r1["iters"] + r2["iters"] == it_o
        np.testing.assert_array_equal(_bits(st.read()), _bits(t_o))
        np.testing.assert_array_equal(parents, p_o)


def test_rrt_so2_forbidden_angle(oxg, orc):
    """oxmpl/tests/rrt_so2ss_tests.rs: SO2 space, forbidden interval [-0.5, 0.5], RRT(0.2, 0.05)."""
    f = worlds.SO2_FORBIDDEN
    sp, sp_o = oxg.Space.so2(), orc.space(0, 1)
    res = (math.pi * 0.05) * 0.1           # so2_state_space.rs:77-79,168-170; rrt.rs:97
    rng = np.random.default_rng(5)
    samples = rng.uniform(-math.pi, math.pi, (4000, 1))
    samples[rng.random(4000) < f["goal_bias"]] = f["goal"][0]
    ow, keep = orc.world(0, so2_intervals=[f["interval"]])
    t_o, p_o, solved_o, it_o = orc.rrt(ow, sp_o, f["start"], f["goal"], f["goal_radius"],
                                       f["max_distance"], res, samples, 4001)
    assert solved_o
    with oxg.World(0, so2_intervals=[f["interval"]]) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([f["start"]]))
        parents, r = oxg.rrt_solve(gw, st, samples, f["max_distance"], res, 4001,
                                   goal_target=f["goal"], goal_radius=f["goal_radius"])
        assert r["solved"] and r["iters"] == it_o
        np.testing.assert_array_equal(_bits(st.read()), _bits(t_o))
        np.testing.assert_array_equal(parents, p_o)


def _config2(n_boxes=1000):
    lo, hi = [0.0, 0.0], [100.0, 100.0]
    start, goal = [5.0, 5.0, 0.0], [95.0, 95.0, 1.0]
    boxes = worlds.aabb_world(2, lo, hi, n_boxes, 1.0, 3.5, keep_clear=[start[:2], goal[:2]])
    rng = np.random.default_rng(2)
    samples = np.column_stack([rng.uniform(0, 100, 30000), rng.uniform(0, 100, 30000),
                               rng.uniform(-math.pi, math.pi, 30000)])
    return lo, hi, start, goal, boxes, samples


def test_config2_rrt_connect_compound_identical_trees(oxg, orc, ctas):
    """BASELINE config 2: RRT-Connect on R^2 x SO2 (weights 1.0 / 0.5) through 1000 AABBs."""
    lo, hi, start, goal, boxes, samples = _config2()
    sp = oxg.Space.compound(2, 1, 1.0, 0.5)
    sp_o = orc.space(2, 1, 1.0, 0.5)
    # resolution of the composed space: builder-defined as the RealVector block's
    res = _res(oxg, orc, lo, hi)
    ow, keep = orc.world(2, aabbs=boxes)
    max_nodes = 30000
    o = orc.rrt_connect(ow, sp_o, start, goal, goal, 1.0, 2.0, res, samples, max_nodes)
    assert o["solved"]
    with oxg.World(2, aabbs=boxes) as gw, oxg.StateStore(sp) as ta, oxg.StateStore(sp) as tb:
        ta.append(np.array([start]))
        tb.append(np.array([goal]))
        pa, pb, r = oxg.rrt_connect_solve(gw, ta, tb, samples, 2.0, res, max_nodes,
                                          goal_target=goal, goal_radius=1.0)
        assert r["solved"] and r["iters"] == o["iters"]
        np.testing.assert_array_equal(_bits(ta.read()), _bits(o["start_tree"]))
        np.testing.assert_array_equal(_bits(tb.read()), _bits(o["goal_tree"]))
        np.testing.assert_array_equal(pa, o["start_parents"])
        np.testing.assert_array_equal(pb, o["goal_parents"])
        assert r["goal_node"] == (o["conn_start"], o["conn_goal"])


def _config3(n_spheres, dim=6):
    lo, hi = [-1.0] * dim, [1.0] * dim
    start = [-0.9] * dim
    goal = [0.9] * dim
    # sphere radii scaled so a useful fraction of a 6-D box is blocked (0.05-0.12 blocks nothing)
    sph = worlds.sphere_world(3, lo, hi, n_spheres, 0.25, 0.45, keep_clear=[start, goal])
    return lo, hi, start, goal, sph


@pytest.mark.parametrize("n_spheres,n_samples,search_radius", [(300, 2500, 0.55), (2000, 1200, 0.7)])
def test_config3_rrt_star_identical_tree_costs(oxg, orc, ctas, n_spheres, n_samples, search_radius):
    """BASELINE config 3 shape: RRT* in 6-D among spheres; count-bounded, no goal stop.  Parents
    AND costs after every rewire must equal the sequential reference."""
    lo, hi, start, goal, sph = _config3(n_spheres)
    sp, sp_o = oxg.Space.real_vector(6), orc.space(6)
    res = _res(oxg, orc, lo, hi)
    samples = worlds.sample_stream(33, n_samples, lo, hi, 0.05, goal)
    ow, keep = orc.world(6, spheres=sph)
    max_nodes = n_samples + 1
    o = orc.rrt_star(ow, sp_o, start, goal, 0.3, 0.2, search_radius, res, samples, max_nodes,
                     stop_at_goal=False)
    assert len(o["tree"]) > n_samples // 4
    with oxg.World(6, spheres=sph) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([start]))
        parents, costs, r = oxg.rrt_star_solve(gw, st, samples, 0.2, search_radius, res, max_nodes,
                                               goal_target=goal, goal_radius=0.3,
                                               stop_at_goal=False)
        assert r["iters"] == o["iters"] and r["n_nodes"][0] == len(o["tree"])
        np.testing.assert_array_equal(_bits(st.read()), _bits(o["tree"]))
        np.testing.assert_array_equal(parents, o["parents"])
        np.testing.assert_array_equal(_bits(costs), _bits(o["costs"]))
        assert r["motion_checks"] == o["motion_checks"]
        assert r["motion_checks_executed"] >= r["motion_checks"]
        assert (parents[1:] != np.arange(1, len(parents))).all()


def test_rrt_star_rvss_fixture_stops_at_goal(oxg, orc):
    """oxmpl/tests/rrt_star_rvss_tests.rs: wall world, RRTStar(0.5, 0.0, 0.25)."""
    c = worlds.WALL
    sp, sp_o = oxg.Space.real_vector(2), orc.space(2)
    res = _res(oxg, orc, c["lo"], c["hi"])
    samples = worlds.sample_stream(12, 8000, c["lo"], c["hi"], 0.0, c["goal"])
    ow, keep = orc.world(2, aabbs=c["aabbs"])
    o = orc.rrt_star(ow, sp_o, c["start"], c["goal"], c["goal_radius"], 0.5, 0.25, res, samples, 8001)
    assert o["solved"]
    with oxg.World(2, aabbs=c["aabbs"]) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([c["start"]]))
        parents, costs, r = oxg.rrt_star_solve(gw, st, samples, 0.5, 0.25, res, 8001,
                                               goal_target=c["goal"], goal_radius=c["goal_radius"])
        assert r["solved"] and r["iters"] == o["iters"]
        np.testing.assert_array_equal(_bits(st.read()), _bits(o["tree"]))
        np.testing.assert_array_equal(parents, o["parents"])
        np.testing.assert_array_equal(_bits(costs), _bits(o["costs"]))
        assert r["goal_node"][0] == len(o["tree"]) - 1


def test_planner_bounds_and_errors(oxg, orc):
    c = worlds.CONFIG1
    sp = oxg.Space.real_vector(2)
    res = _res(oxg, orc, c["lo"], c["hi"])
    samples = worlds.sample_stream(1, 500, c["lo"], c["hi"], 0.0, c["goal"])
    with oxg.World(2, spheres=c["spheres"]) as gw, oxg.StateStore(sp) as st:
        with pytest.raises(oxg.OxgError) as e:                       # empty tree: setup not done
            oxg.rrt_solve(gw, st, samples, 0.5, res, 100)
        assert e.value.code == -7
        st.append(np.array([c["start"]]))
        with pytest.raises(oxg.OxgError) as e:
            oxg.rrt_solve(gw, st, samples, 0.5, 0.0, 100)
        assert e.value.code == -5
        parents, r = oxg.rrt_solve(gw, st, samples, 0.5, res, 40)    # capacity bound
        assert r["status"] == "tree_full" and r["n_nodes"][0] == 40 and len(st) == 40
        st.append(worlds.sample_stream(9, 3000, [3.0, 3.0], [9.0, 9.0], 0.0, c["goal"]))
        with pytest.raises(oxg.OxgError) as e:                       # > 2048 RRT* neighbours
            oxg.rrt_star_solve(gw, st, samples, 0.5, 1e9, res, 5000)
        assert e.value.code == -3
    with oxg.World(2, spheres=c["spheres"]) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([c["start"]]))
        _, r = oxg.rrt_solve(gw, st, np.tile(samples, (400, 1)), 1e-4, res, 10**6, timeout_s=0.05)
        assert r["status"] in ("timeout", "samples_exhausted")


def test_grid_grows_with_the_tree(oxg, orc, monkeypatch):
    """Without OXG_PLAN_CTAS a run starts on one CTA and is relaunched on a larger grid when a
    tree passes 4096 nodes (the kernels resume from the stores): sample consumption, trees,
    parents, costs and counters must be those of one uninterrupted run."""
    monkeypatch.delenv("OXG_PLAN_CTAS", raising=False)
    c = worlds.CONFIG1
    sp, sp_o = oxg.Space.real_vector(2), orc.space(2)
    res = _res(oxg, orc, c["lo"], c["hi"])
    # RRT, no goal: 9000 samples grow ~8000 nodes; max_nodes = 40000 asks for 3 CTAs later on
    samples = worlds.sample_stream(71, 9000, c["lo"], c["hi"], 0.0, c["goal"])
    ow, keep = orc.world(2, spheres=c["spheres"])
    t_o, p_o, solved_o, it_o = orc.rrt(ow, sp_o, c["start"], c["goal"], -1.0, c["max_distance"], res,
                                       samples, 40000)
    assert len(t_o) > 6000 and not solved_o
    with oxg.World(2, spheres=c["spheres"]) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([c["start"]]))
        parents, r = oxg.rrt_solve(gw, st, samples, c["max_distance"], res, 40000)
        assert r["status"] == "samples_exhausted" and r["iters"] == it_o == 9000
        assert r["ctas"] > 1                                           # the last stage's grid
        np.testing.assert_array_equal(_bits(st.read()), _bits(t_o))
        np.testing.assert_array_equal(parents, p_o)
    # RRT*: choose-parent / rewire state (costs) must survive the relaunch
    samples = worlds.sample_stream(72, 5200, c["lo"], c["hi"], 0.0, c["goal"])
    o = orc.rrt_star(ow, sp_o, c["start"], c["goal"], -1.0, 0.5, 0.6, res, samples, 40000,
                     stop_at_goal=False)
    assert len(o["tree"]) > 4300
    with oxg.World(2, spheres=c["spheres"]) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([c["start"]]))
        parents, costs, r = oxg.rrt_star_solve(gw, st, samples, 0.5, 0.6, res, 40000,
                                               stop_at_goal=False)
        assert r["iters"] == o["iters"] and r["n_nodes"][0] == len(o["tree"])
        np.testing.assert_array_equal(_bits(st.read()), _bits(o["tree"]))
        np.testing.assert_array_equal(parents, o["parents"])
        np.testing.assert_array_equal(_bits(costs), _bits(o["costs"]))
        assert r["motion_checks"] == o["motion_checks"]
    # RRT-Connect: both trees cross the threshold
    lo, hi = [0.0, 0.0], [100.0, 100.0]
    start, goal = [5.0, 5.0], [95.0, 95.0]
    samples = np.random.default_rng(73).uniform(0.0, 100.0, size=(18000, 2))
    res2 = _res(oxg, orc, lo, hi)
    wall = np.array([[49.0, -1.0, 51.0, 101.0]])                        # keeps the trees apart
    ow2, keep2 = orc.world(2, aabbs=wall)
    o = orc.rrt_connect(ow2, sp_o, start, goal, goal, 0.5, 0.7, res2, samples, 40000)
    assert min(len(o["start_tree"]), len(o["goal_tree"])) > 4300
    with oxg.World(2, aabbs=wall) as gw, oxg.StateStore(sp) as ta, \
            oxg.StateStore(sp) as tb:
        ta.append(np.array([start]))
        tb.append(np.array([goal]))
        pa, pb, r = oxg.rrt_connect_solve(gw, ta, tb, samples, 0.7, res2, 40000,
                                          goal_target=goal, goal_radius=0.5)
        assert r["iters"] == o["iters"] and r["solved"] == o["solved"]
        np.testing.assert_array_equal(_bits(ta.read()), _bits(o["start_tree"]))
        np.testing.assert_array_equal(_bits(tb.read()), _bits(o["goal_tree"]))
        np.testing.assert_array_equal(pa, o["start_parents"])
        np.testing.assert_array_equal(pb, o["goal_parents"])


def test_config3_full_size_tree_properties(oxg, orc, monkeypatch):
    """BASELINE config 3 at FULL size: RRT* in 6-D among 10 000 spheres, run to 200 000 nodes.  The
    oracle needs minutes for that, so the full tree is checked through properties — it is a tree
    (rewiring may point a node to a later one, so: acyclic with a single root), every node is valid, costs are finite and positive, a sample of tree edges passes the oracle's
    check_motion in the direction the planner checked them (parent -> child) — and the first
    3 000 iterations are compared with the oracle bit for bit."""
    monkeypatch.delenv("OXG_PLAN_CTAS", raising=False)
    dim, n_nodes = 6, 200_000
    lo, hi = [-1.0] * dim, [1.0] * dim
    start, goal = [-0.9] * dim, [0.9] * dim
    sph = worlds.sphere_world(3, lo, hi, 10000, 0.22, 0.32, keep_clear=[start, goal])
    sp, sp_o = oxg.Space.real_vector(dim), orc.space(dim)
    res = _res(oxg, orc, lo, hi)
    samples = worlds.sample_stream(33, int(n_nodes * 1.6), lo, hi, 0.0, goal)
    search_radius, max_distance = 0.41, 0.2
    ow, keep = orc.world(dim, spheres=sph)
    with oxg.World(dim, spheres=sph) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([start]))
        parents, costs, r = oxg.rrt_star_solve(gw, st, samples, max_distance, search_radius, res,
                                               n_nodes, stop_at_goal=False)
        tree = st.read()
        assert r["status"] == "tree_full" and len(tree) == n_nodes
        assert gw.is_valid(sp, tree).all()
    assert parents[0] == -1 and (parents[1:] >= 0).all() and (parents < n_nodes).all()
    assert (parents[1:] != np.arange(1, n_nodes)).all()
    # acyclic: pointer-jumping reaches the root from every node within log2(n) rounds of doubling
    hop = np.where(parents < 0, 0, parents).astype(np.int64)
    for _ in range(18):
        hop = hop[hop]
    assert (hop == 0).all()
    assert costs[0] == 0.0 and np.isfinite(costs).all() and (costs[1:] > 0.0).all()
    # every tree edge was accepted by check_motion(parent -> child): re-check a sample on the oracle
    pick = np.random.default_rng(5).choice(np.arange(1, n_nodes), size=3000, replace=False)
    ov, _ = orc.check_motion(ow, sp_o, tree[parents[pick]], tree[pick], res)
    assert ov.all()
    # edges never exceed the steering / rewiring reach
    seg = np.sqrt(((tree[1:] - tree[parents[1:]]) ** 2).sum(axis=1))
    assert (seg <= max(max_distance, search_radius) * (1 + 1e-12)).all()
    # prefix parity with the sequential reference
    o = orc.rrt_star(ow, sp_o, start, None, -1.0, max_distance, search_radius, res, samples[:3000],
                     3001, stop_at_goal=False)
    with oxg.World(dim, spheres=sph) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([start]))
        p2, c2, r2 = oxg.rrt_star_solve(gw, st, samples[:3000], max_distance, search_radius, res,
                                        3001, stop_at_goal=False)
        np.testing.assert_array_equal(_bits(st.read()), _bits(o["tree"]))
        np.testing.assert_array_equal(p2, o["parents"])
        np.testing.assert_array_equal(_bits(c2), _bits(o["costs"]))
