"""GPU parity: is_valid / check_motion / check_edges through the C ABI vs the CPU oracle."""
import math

import numpy as np
import pytest

from conftest import uniform_states
from oxmpl_b200 import worlds

pytestmark = pytest.mark.gpu


def _world_pair(oxg, orc, dim, spheres=None, aabbs=None, capsules=None, so2=()):
    gw = oxg.World(dim, spheres=spheres, aabbs=aabbs, capsules=capsules, so2_intervals=so2)
    ow, keep = orc.world(dim, spheres=spheres, aabbs=aabbs, capsules=capsules, so2_intervals=so2)
    return gw, ow, keep


def test_readme_circle_boundary_and_wall_fixture(oxg, orc):
    """Operator strictness of the reference's own fixtures."""
    sp, sp_o = oxg.Space.real_vector(2), orc.space(2)
    # README circle: valid iff sqrt(x^2 + y^2) > 2.0 (README.md:147-150) — boundary invalid
    gw, ow, keep = _world_pair(oxg, orc, 2, spheres=worlds.CONFIG1["spheres"])
    pts = np.array([[2.0, 0.0], [np.nextafter(2.0, 3.0), 0.0], [np.nextafter(2.0, 0.0), 0.0],
                    [0.0, -2.0], [1.2, 1.6], [3.0, 4.0], [0.0, 0.0], [-5.0, -5.0]])
    got = gw.is_valid(sp, pts)
    np.testing.assert_array_equal(got, orc.is_valid(ow, sp_o, pts))
    np.testing.assert_array_equal(got, [0, 1, 0, 0, 0, 1, 0, 1])
    gw.close()
    # wall: invalid iff lo <= x <= hi on all faces, inclusive (oxmpl/tests/rrt_rvss_tests.rs:24-36)
    gw, ow, keep = _world_pair(oxg, orc, 2, aabbs=worlds.WALL["aabbs"])
    pts = np.array([[4.75, 2.0], [5.25, 8.0], [np.nextafter(4.75, 0.0), 5.0], [5.0, 5.0],
                    [5.0, np.nextafter(8.0, 9.0)], [1.0, 5.0]])
    got = gw.is_valid(sp, pts)
    np.testing.assert_array_equal(got, orc.is_valid(ow, sp_o, pts))
    np.testing.assert_array_equal(got, [0, 0, 1, 0, 1, 1])
    gw.close()


@pytest.mark.parametrize("dim", [2, 3, 6, 7, 14])
def test_is_valid_and_motion_random_worlds(oxg, orc, dim):
    lo, hi = [-1.0] * dim, [1.0] * dim
    spheres, aabbs, capsules = worlds.calibrated_obstacles(dim, dim, 40, 25, 33)
    sp, sp_o = oxg.Space.real_vector(dim), orc.space(dim)
    gw, ow, keep = _world_pair(oxg, orc, dim, spheres, aabbs, capsules)
    states = uniform_states(50 + dim, 20000, lo, hi)
    np.testing.assert_array_equal(gw.is_valid(sp, states), orc.is_valid(ow, sp_o, states))
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))
    assert res == orc.resolution(orc.rv_max_extent(lo, hi))
    frm = uniform_states(60 + dim, 6000, lo, hi)
    rng = np.random.default_rng(70 + dim)
    to = frm + rng.normal(scale=0.25, size=frm.shape) * rng.random((6000, 1))
    to[:50] = frm[:50]                                  # zero-length: num_steps = 0 -> is_valid(to)
    to[50:100] = frm[50:100] + 1e-4                     # one step
    v, first = gw.check_motion(sp, frm, to, res, want_first_invalid=True)
    ov, ofirst = orc.check_motion(ow, sp_o, frm, to, res)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(first, ofirst)
    assert 0 < v.sum() < v.size                          # the world is neither empty nor solid
    # the few-edges launch shape (a warp per edge) must agree too
    v2, f2 = gw.check_motion(sp, frm[:97], to[:97], res, want_first_invalid=True)
    np.testing.assert_array_equal(v2, ov[:97])
    np.testing.assert_array_equal(f2, ofirst[:97])
    # index-pair form against a store
    with oxg.StateStore(sp) as st:
        st.append(np.vstack([frm[:3000], to[:3000]]))
        fi = np.arange(3000, dtype=np.uint32)
        ti = fi + 3000
        np.testing.assert_array_equal(gw.check_edges(st, fi, ti, res), ov[:3000])
    gw.close()


def test_motion_direction_and_so2(oxg, orc):
    """SO2 / compound spaces: argument order matters (SURVEY hard part 3)."""
    rng = np.random.default_rng(5)
    # pure SO2 with the forbidden-angle interval (oxmpl/tests/rrt_so2ss_tests.rs:17-27)
    sp, sp_o = oxg.Space.so2(), orc.space(0, 1)
    gw, ow, keep = _world_pair(oxg, orc, 0, so2=[worlds.SO2_FORBIDDEN["interval"]])
    a = rng.uniform(-3 * math.pi, 3 * math.pi, (4000, 1))
    b = rng.uniform(-3 * math.pi, 3 * math.pi, (4000, 1))
    np.testing.assert_array_equal(gw.is_valid(sp, a), orc.is_valid(ow, sp_o, a))
    res = (math.pi * 0.05) * 0.1                    # so2_state_space.rs:77-79,168-170; rrt.rs:97
    for x, y in ((a, b), (b, a)):
        v, f = gw.check_motion(sp, x, y, res, want_first_invalid=True)
        ov, of = orc.check_motion(ow, sp_o, x, y, res)
        np.testing.assert_array_equal(v, ov)
        np.testing.assert_array_equal(f, of)
    gw.close()
    # SE2-style compound: AABBs over (x, y), angle free
    sp, sp_o = oxg.Space.compound(2, 1, 1.0, 0.5), orc.space(2, 1, 1.0, 0.5)
    boxes = worlds.aabb_world(9, [0, 0], [100, 100], 300, 1.0, 5.0)
    gw, ow, keep = _world_pair(oxg, orc, 2, aabbs=boxes)
    a = np.hstack([rng.uniform(0, 100, (5000, 2)), rng.uniform(-math.pi, math.pi, (5000, 1))])
    b = a + np.hstack([rng.normal(scale=2.0, size=(5000, 2)), rng.normal(scale=1.0, size=(5000, 1))])
    np.testing.assert_array_equal(gw.is_valid(sp, a), orc.is_valid(ow, sp_o, a))
    for x, y in ((a, b), (b, a)):
        v, f = gw.check_motion(sp, x, y, 0.7, want_first_invalid=True)
        ov, of = orc.check_motion(ow, sp_o, x, y, 0.7)
        np.testing.assert_array_equal(v, ov)
        np.testing.assert_array_equal(f, of)
    gw.close()


def test_resolution_rejected(oxg):
    from oxmpl_b200 import OxgError

    sp = oxg.Space.real_vector(2)
    with oxg.World(2, spheres=worlds.CONFIG1["spheres"]) as gw:
        for bad in (0.0, -1.0, float("nan")):
            with pytest.raises(OxgError) as e:
                gw.check_motion(sp, [[0.0, 0.0]], [[1.0, 1.0]], bad)
            assert e.value.code == -5  # OXG_E_RESOLUTION


def test_config1_step_count(oxg, orc):
    """Config-1 arithmetic: extent 28.28..., resolution 0.1414..., a 0.5-long edge has 4 steps
    (SURVEY §8a row A6): an obstacle that only the 3rd interpolated state touches must be found
    at step 3."""
    sp, sp_o = oxg.Space.real_vector(2), orc.space(2)
    res = oxg.motion_resolution(oxg.real_vector_max_extent([(-10.0, 10.0), (-10.0, 10.0)]))
    assert res == 0.14142135623730953
    frm, to = np.array([[3.0, 3.0]]), np.array([[3.5, 3.0]])
    sph = np.array([[3.375, 3.0, 0.01]])
    gw, ow, keep = _world_pair(oxg, orc, 2, spheres=sph)
    v, f = gw.check_motion(sp, frm, to, res, want_first_invalid=True)
    ov, of = orc.check_motion(ow, sp_o, frm, to, res)
    assert (v[0], f[0]) == (0, 3) == (ov[0], of[0])
    gw.close()


@pytest.mark.parametrize("dim", [2, 7])
@pytest.mark.parametrize("name,scale,offset", [
    ("offset_1e6", 1.0, 1.0e6),          # f32 images of obstacles and states nearly collide
    ("offset_1e12", 1.0, 1.0e12),
    ("tiny_1e-20", 1e-20, 0.0),
    ("huge_1e150", 1e150, 0.0),           # f32 overflow everywhere, f64 squares still finite
    ("unit", 1.0, 0.0),
])
def test_validity_prefilter_never_changes_results(oxg, orc, dim, name, scale, offset):
    """The FP32 pre-filter of the validity kernels (oxg_math.cuh, *_maybe_hit32) may only skip
    exact tests whose outcome is certain: scaled / shifted worlds and states that sit exactly on
    obstacle boundaries must give the oracle's booleans."""
    spheres, aabbs, capsules = worlds.calibrated_obstacles(dim, dim, 24, 24, 24, p_each=0.02)

    def tr(a, radius_col=None):
        a = a.copy()
        if radius_col is None:
            return a * scale + offset
        a[:, :radius_col] = a[:, :radius_col] * scale + offset
        a[:, radius_col] = a[:, radius_col] * scale
        return a

    S, B, C = tr(spheres, dim), tr(aabbs), tr(capsules, 2 * dim)
    sp, sp_o = oxg.Space.real_vector(dim), orc.space(dim)
    gw, ow, keep = _world_pair(oxg, orc, dim, S, B, C)
    states = uniform_states(90 + dim, 6000, [-1.0] * dim, [1.0] * dim) * scale + offset
    # states exactly on boundaries: sphere surface along +e0, box faces and corners, capsule caps
    on_sphere = S[:, :dim].copy()
    on_sphere[:, 0] = on_sphere[:, 0] + S[:, dim]
    on_box = np.vstack([B[:, :dim], B[:, dim:], 0.5 * (B[:, :dim] + B[:, dim:])])
    on_cap = C[:, dim:2 * dim].copy()
    on_cap[:, -1] = on_cap[:, -1] + C[:, 2 * dim]
    states = np.ascontiguousarray(np.vstack([states, on_sphere, on_box, on_cap,
                                             np.nextafter(on_sphere, np.inf),
                                             np.nextafter(on_box, -np.inf)]))
    got, want = gw.is_valid(sp, states), orc.is_valid(ow, sp_o, states)
    np.testing.assert_array_equal(got, want, err_msg=name)
    frm = states[:3000]
    to = states[3000:6000]
    res = 0.05 * scale
    v, first = gw.check_motion(sp, frm, to, res, want_first_invalid=True)
    ov, ofirst = orc.check_motion(ow, sp_o, frm, to, res)
    np.testing.assert_array_equal(v, ov, err_msg=name)
    np.testing.assert_array_equal(first, ofirst, err_msg=name)
    if name in ("unit", "tiny_1e-20", "huge_1e150"):
        assert 0 < want.sum() < want.size
    gw.close()


@pytest.mark.parametrize("dim", [2, 7])
def test_few_states_large_world(oxg, orc, dim):
    """<= 592 states / edges against >= 64 obstacles take the CTA-per-edge kernel (the trait
    path `is_valid(&state)` and an unfused planner iteration): booleans and first invalid step
    must equal the sequential loop's."""
    spheres, aabbs, capsules = worlds.calibrated_obstacles(dim + 20, dim, 120, 90, 110, p_each=0.004)
    sp, sp_o = oxg.Space.real_vector(dim), orc.space(dim)
    gw, ow, keep = _world_pair(oxg, orc, dim, spheres, aabbs, capsules)
    for n in (1, 7, 300):
        states = uniform_states(70 + n, n, [-1.0] * dim, [1.0] * dim)
        np.testing.assert_array_equal(gw.is_valid(sp, states), orc.is_valid(ow, sp_o, states))
        to = states + np.random.default_rng(n).normal(scale=0.3, size=states.shape)
        to[0] = states[0]
        for res in (0.02, 0.5, 100.0):
            v, first = gw.check_motion(sp, states, to, res, want_first_invalid=True)
            ov, ofirst = orc.check_motion(ow, sp_o, states, to, res)
            np.testing.assert_array_equal(v, ov)
            np.testing.assert_array_equal(first, ofirst)
    gw.close()
    # compound space with SO2 intervals through the generic (runtime-dimension) instantiation
    spc, spc_o = oxg.Space.compound(2, 1, 1.0, 0.5), orc.space(2, 1, 1.0, 0.5)
    boxes = worlds.aabb_world(3, [0.0, 0.0], [10.0, 10.0], 80, 0.3, 1.0)
    gw2 = oxg.World(2, aabbs=boxes, so2_intervals=[(0, -0.4, 0.2)])
    ow2, keep2 = orc.world(2, aabbs=boxes, so2_intervals=[(0, -0.4, 0.2)])
    rng = np.random.default_rng(8)
    a = np.column_stack([rng.uniform(0, 10, (200, 2)), rng.uniform(-3.1, 3.1, 200)])
    b = np.column_stack([rng.uniform(0, 10, (200, 2)), rng.uniform(-3.1, 3.1, 200)])
    np.testing.assert_array_equal(gw2.is_valid(spc, a), orc.is_valid(ow2, spc_o, a))
    v, first = gw2.check_motion(spc, a, b, 0.3, want_first_invalid=True)
    ov, ofirst = orc.check_motion(ow2, spc_o, a, b, 0.3)
    np.testing.assert_array_equal(v, ov)
    np.testing.assert_array_equal(first, ofirst)
    gw2.close()


@pytest.mark.parametrize("dim,scale,offset", [(2, 1.0, 0.0), (3, 1.0, 0.0), (7, 1.0, 0.0), (7, 1.0, 1e6),
                                              (6, 1e-20, 0.0), (7, 1e150, 0.0), (14, 1.0, 0.0)])
def test_whole_edge_cull_never_changes_results(oxg, orc, dim, scale, offset):
    """Many edges take motion_cull_kernel: an obstacle is dropped for a whole edge only when the
    ball around the midpoint provably clears it.  Validity AND the first invalid step must be the
    sequential loop's (rrt.rs:90-116), for short, long, grazing, zero-length and NaN edges."""
    spheres, aabbs, capsules = worlds.calibrated_obstacles(3 * dim, dim, 30, 20, 40, p_each=0.01)

    def tr(a, radius_col=None):
        a = a.copy()
        if radius_col is None:
            return a * scale + offset
        a[:, :radius_col] = a[:, :radius_col] * scale + offset
        a[:, radius_col] = a[:, radius_col] * scale
        return a

    S, B, C = tr(spheres, dim), tr(aabbs), tr(capsules, 2 * dim)
    sp, sp_o = oxg.Space.real_vector(dim), orc.space(dim)
    gw, ow, keep = _world_pair(oxg, orc, dim, S, B, C)
    rng = np.random.default_rng(700 + dim)
    n = 16000
    frm = rng.uniform(-1.0, 1.0, size=(n, dim))
    to = frm + rng.normal(scale=0.2, size=frm.shape) * rng.random((n, 1))
    to[:2000] = rng.uniform(-1.0, 1.0, size=(2000, dim))            # long edges across the world
    to[2000:2100] = frm[2000:2100]                                  # zero length
    # grazing: start on a sphere's surface and leave tangentially / radially
    k = min(len(spheres), 200)
    c, r = spheres[:k, :dim], spheres[:k, dim]
    u = rng.normal(size=(k, dim))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    frm[3000:3000 + k] = c + u * r[:, None]
    to[3000:3000 + k] = frm[3000:3000 + k] + u * 0.05
    v = rng.normal(size=(k, dim))
    v -= (v * u).sum(1, keepdims=True) * u
    frm[4000:4000 + k] = c + u * (r[:, None] * (1 + 1e-12))
    to[4000:4000 + k] = frm[4000:4000 + k] + v * 0.1
    frm, to = frm * scale + offset, to * scale + offset
    frm[5000, 0] = np.nan
    to[5001, dim - 1] = np.inf
    res = 0.02 * scale
    got_v, got_f = gw.check_motion(sp, frm, to, res, want_first_invalid=True)
    want_v, want_f = orc.check_motion(ow, sp_o, frm, to, res)
    np.testing.assert_array_equal(got_v, want_v)
    # an infinitely long edge: the reference loop would never end; the library reports
    # "invalid, step 0xFFFFFFFF" (include/oxg.h), the oracle wrapper "invalid, step 0"
    assert got_v[5001] == 0 and got_f[5001] == 0xFFFFFFFF
    got_f[5001] = want_f[5001]
    np.testing.assert_array_equal(got_f, want_f)
    if offset == 0.0:
        assert 0 < want_v.sum() < want_v.size
    with oxg.StateStore(sp) as st:
        st.append(np.vstack([frm, to]))
        fi = np.arange(n, dtype=np.uint32)
        np.testing.assert_array_equal(gw.check_edges(st, fi, fi + n, res), want_v)
    gw.close()
