"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): the CPU test
pins the oracle against them, the GPU test runs liboxg through the C ABI on the stored inputs
and compares with the stored outputs bit for bit."""
import os

import numpy as np
import pytest

from oxmpl_b200 import worlds

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def _load(name):
    return np.load(os.path.join(G, name))


# ------------------------------------------------------------------------------ CPU: oracle
def test_oracle_matches_golden_scans(orc):
    g = _load("scans.npz")
    for d in (3, 7):
        sp = orc.space(d)
        rows, q = g[f"rows{d}"], g[f"queries{d}"]
        i1, d1 = orc.nn1(sp, rows, q)
        np.testing.assert_array_equal(i1, g[f"nn1_idx{d}"])
        np.testing.assert_array_equal(_bits(d1), _bits(g[f"nn1_dist{d}"]))
        ik, dk = orc.knn(sp, rows, q, 6)
        np.testing.assert_array_equal(ik, g[f"knn_idx{d}"])
        np.testing.assert_array_equal(_bits(dk), _bits(g[f"knn_dist{d}"]))
        off, ir, dr = orc.radius(sp, rows, q, 0.9, want_dist=True)
        np.testing.assert_array_equal(off, g[f"rad_off{d}"])
        np.testing.assert_array_equal(ir, g[f"rad_idx{d}"])
        np.testing.assert_array_equal(_bits(dr), _bits(g[f"rad_dist{d}"]))
    spc = orc.space(2, 1, 1.0, 0.5)
    ic, dc = orc.nn1(spc, g["rows_c"], g["queries_c"])
    np.testing.assert_array_equal(ic, g["nn1_idx_c"])
    np.testing.assert_array_equal(_bits(dc), _bits(g["nn1_dist_c"]))


def test_oracle_matches_golden_validity_and_planners(orc):
    g = _load("validity.npz")
    ow, keep = orc.world(3, spheres=g["spheres"], aabbs=g["aabbs"], capsules=g["capsules"])
    sp = orc.space(3)
    np.testing.assert_array_equal(orc.is_valid(ow, sp, g["a"]), g["is_valid"])
    mv, first = orc.check_motion(ow, sp, g["a"], g["b"], float(g["resolution"]))
    np.testing.assert_array_equal(mv, g["motion"])
    np.testing.assert_array_equal(first, g["first_invalid"])
    p = _load("planners.npz")
    c = worlds.CONFIG1
    ow1, keep1 = orc.world(2, spheres=c["spheres"])
    tree, parents, solved, iters = orc.rrt(ow1, orc.space(2), c["start"], c["goal"], c["goal_radius"],
                                           c["max_distance"], float(p["rrt_resolution"]),
                                           p["rrt_samples"], 2001)
    np.testing.assert_array_equal(_bits(tree), _bits(p["rrt_tree"]))
    np.testing.assert_array_equal(parents, p["rrt_parents"])
    assert solved == bool(p["rrt_solved"]) and iters == int(p["rrt_iters"])


# ------------------------------------------------------------------------------ GPU: liboxg
@pytest.mark.gpu
def test_gpu_matches_golden_scans(oxg):
    g = _load("scans.npz")
    for d in (3, 7):
        with oxg.StateStore(oxg.Space.real_vector(d)) as st:
            st.append(g[f"rows{d}"])
            q = g[f"queries{d}"]
            i1, d1 = st.nn1(q)
            np.testing.assert_array_equal(i1, g[f"nn1_idx{d}"])
            np.testing.assert_array_equal(_bits(d1), _bits(g[f"nn1_dist{d}"]))
            ik, dk = st.knn(q, 6)
            np.testing.assert_array_equal(ik, g[f"knn_idx{d}"])
            np.testing.assert_array_equal(_bits(dk), _bits(g[f"knn_dist{d}"]))
            off, ir, dr = st.radius(q, 0.9, want_dist=True)
            np.testing.assert_array_equal(off, g[f"rad_off{d}"])
            np.testing.assert_array_equal(ir, g[f"rad_idx{d}"])
            np.testing.assert_array_equal(_bits(dr), _bits(g[f"rad_dist{d}"]))
            i1, d1 = st.nn1(q[:3])                        # few-query streaming kernels
            np.testing.assert_array_equal(i1, g[f"nn1_idx{d}"][:3])
    with oxg.StateStore(oxg.Space.compound(2, 1, 1.0, 0.5)) as st:
        st.append(g["rows_c"])
        ic, dc = st.nn1(g["queries_c"])
        np.testing.assert_array_equal(ic, g["nn1_idx_c"])
        np.testing.assert_array_equal(_bits(dc), _bits(g["nn1_dist_c"]))
        ik, dk = st.knn(g["queries_c"], 4)
        np.testing.assert_array_equal(ik, g["knn_idx_c"])
        np.testing.assert_array_equal(_bits(dk), _bits(g["knn_dist_c"]))


@pytest.mark.gpu
def test_gpu_matches_golden_validity_and_planners(oxg):
    g = _load("validity.npz")
    sp = oxg.Space.real_vector(3)
    with oxg.World(3, spheres=g["spheres"], aabbs=g["aabbs"], capsules=g["capsules"]) as gw:
        np.testing.assert_array_equal(gw.is_valid(sp, g["a"]), g["is_valid"])
        mv, first = gw.check_motion(sp, g["a"], g["b"], float(g["resolution"]), want_first_invalid=True)
        np.testing.assert_array_equal(mv, g["motion"])
        np.testing.assert_array_equal(first, g["first_invalid"])
    p = _load("planners.npz")
    c = worlds.CONFIG1
    sp2 = oxg.Space.real_vector(2)
    with oxg.World(2, spheres=c["spheres"]) as gw, oxg.StateStore(sp2) as st:
        st.append(np.array([c["start"]]))
        parents, r = oxg.rrt_solve(gw, st, p["rrt_samples"], c["max_distance"],
                                   float(p["rrt_resolution"]), 2001, goal_target=c["goal"],
                                   goal_radius=c["goal_radius"])
        np.testing.assert_array_equal(_bits(st.read()), _bits(p["rrt_tree"]))
        np.testing.assert_array_equal(parents, p["rrt_parents"])
        assert r["solved"] == bool(p["rrt_solved"]) and r["iters"] == int(p["rrt_iters"])
    w = worlds.WALL
    with oxg.World(2, aabbs=w["aabbs"]) as gw, oxg.StateStore(sp2) as st:
        st.append(np.array([w["start"]]))
        parents, costs, r = oxg.rrt_star_solve(gw, st, p["star_samples"], 0.5, 0.6,
                                               float(p["star_resolution"]), 701, stop_at_goal=False)
        np.testing.assert_array_equal(_bits(st.read()), _bits(p["star_tree"]))
        np.testing.assert_array_equal(parents, p["star_parents"])
        np.testing.assert_array_equal(_bits(costs), _bits(p["star_costs"]))
        assert r["motion_checks"] == int(p["star_motion_checks"])


# ------------------------------------------------------------------------------ SO(3)
def _check_so3_against_golden(nn1, radius, is_valid, check_motion):
    """Shared by the CPU (oracle) and GPU (liboxg) tests: callables with the oracle's shapes."""
    g = _load("so3.npz")
    idx, dist = nn1(g["rows"], g["queries"])
    np.testing.assert_array_equal(idx, g["nn1_idx"])
    np.testing.assert_array_equal(_bits(dist), _bits(g["nn1_dist"]))
    off, ridx, rdist = radius(g["rows"], g["queries"], float(g["radius"]))
    np.testing.assert_array_equal(off, g["rad_off"])
    np.testing.assert_array_equal(ridx, g["rad_idx"])
    np.testing.assert_array_equal(_bits(rdist), _bits(g["rad_dist"]))
    np.testing.assert_array_equal(is_valid(g["cones"], g["states"]), g["valid"])
    mv, mf = check_motion(g["cones"], g["frm"], g["to"], float(g["resolution"]))
    np.testing.assert_array_equal(mv, g["motion_valid"])
    np.testing.assert_array_equal(mf, g["motion_first"])
    assert 0 < int(g["valid"].sum()) < len(g["valid"]) and 0 < int(g["motion_valid"].sum()) < 120
    assert g["nn1_idx"][0] == 3 and g["nn1_idx"][2] == 3 and g["nn1_dist"][1] == 0.0


def test_oracle_matches_golden_so3(orc):
    _check_so3_against_golden(orc.so3_nn1, orc.so3_radius, orc.so3_is_valid, orc.so3_check_motion)


@pytest.mark.gpu
def test_gpu_matches_golden_so3(oxg):
    def nn1(rows, queries):
        with oxg.StateStore(oxg.Space.real_vector(4)) as st:
            st.append(rows)
            return st.so3_nn1(queries)

    def radius(rows, queries, r):
        with oxg.StateStore(oxg.Space.real_vector(4)) as st:
            st.append(rows)
            return st.so3_radius(queries, r, want_dist=True)

    def is_valid(cones, states):
        return oxg.So3Cones(cones).is_valid(states)

    def check_motion(cones, frm, to, res):
        return oxg.So3Cones(cones).check_motion(frm, to, res, want_first_invalid=True)

    _check_so3_against_golden(nn1, radius, is_valid, check_motion)
