#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the CPU oracle (oracle/oxg_oracle.cpp).

The reference (Rust) cannot be built in this image and its own tests hold no golden vectors for
this path (they assert path properties with an unseeded RNG — SURVEY §4), so the fixtures are
oracle outputs on seeded inputs: they pin the oracle against regressions (CPU test) and give
the GPU path a second, file-based target (GPU test).  Inputs are stored with the outputs, so
the files are self-contained.  Re-run:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402
from oxmpl_b200 import worlds  # noqa: E402
from oxmpl_b200.api import motion_resolution, real_vector_max_extent  # noqa: E402


def lattice_with_ties(seed, n, d):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 4, size=(n, d)).astype(np.float64)
    near = base[: n // 4].copy()
    near[:, 0] = np.nextafter(near[:, 0] + 0.5, np.inf)
    pts = np.vstack([base, base[rng.integers(0, n, n // 4)], near])
    rng.shuffle(pts)
    return np.ascontiguousarray(pts)


def main():
    orc = oracle.get()
    # ---- scans: uniform + lattice (mass exact ties) in 3-D and 7-D
    out = {}
    for d in (3, 7):
        rng = np.random.default_rng(10 + d)
        rows = np.vstack([rng.random((700, d)) * 3.0, lattice_with_ties(20 + d, 400, d)])
        queries = np.vstack([rng.random((40, d)) * 3.0, rows[690:720]])
        sp = orc.space(d)
        i1, d1 = orc.nn1(sp, rows, queries)
        ik, dk = orc.knn(sp, rows, queries, 6)
        off, ir, dr = orc.radius(sp, rows, queries, 0.9, want_dist=True)
        out.update({f"rows{d}": rows, f"queries{d}": queries, f"nn1_idx{d}": i1, f"nn1_dist{d}": d1,
                    f"knn_idx{d}": ik, f"knn_dist{d}": dk, f"rad_off{d}": off, f"rad_idx{d}": ir,
                    f"rad_dist{d}": dr})
    # SO2 and compound spaces (argument order matters)
    rng = np.random.default_rng(31)
    rows = np.column_stack([rng.random((300, 2)) * 10, rng.uniform(-3.5, 3.5, 300)])
    queries = np.column_stack([rng.random((25, 2)) * 10, rng.uniform(-3.5, 3.5, 25)])
    spc = orc.space(2, 1, 1.0, 0.5)
    ic, dc = orc.nn1(spc, rows, queries)
    ikc, dkc = orc.knn(spc, rows, queries, 4)
    out.update({"rows_c": rows, "queries_c": queries, "nn1_idx_c": ic, "nn1_dist_c": dc,
                "knn_idx_c": ikc, "knn_dist_c": dkc})
    np.savez_compressed(os.path.join(HERE, "scans.npz"), **out)

    # ---- validity / motion on a mixed 3-D world
    spheres, aabbs, capsules = worlds.calibrated_obstacles(5, 3, 12, 9, 10, p_each=0.02)
    ow, keep = orc.world(3, spheres=spheres, aabbs=aabbs, capsules=capsules)
    sp = orc.space(3)
    rng = np.random.default_rng(41)
    a = rng.uniform(-1, 1, (400, 3))
    b = a + rng.normal(scale=0.3, size=a.shape)
    b[:5] = a[:5]
    res = motion_resolution(real_vector_max_extent([(-1.0, 1.0)] * 3))
    valid = orc.is_valid(ow, sp, a)
    mv, first = orc.check_motion(ow, sp, a, b, res)
    np.savez_compressed(os.path.join(HERE, "validity.npz"), spheres=spheres, aabbs=aabbs,
                        capsules=capsules, a=a, b=b, resolution=res, is_valid=valid, motion=mv,
                        first_invalid=first)

    # ---- BASELINE config 1: README RRT, fixed sample stream -> tree, parents
    c = worlds.CONFIG1
    res1 = motion_resolution(real_vector_max_extent(list(zip(c["lo"], c["hi"]))))
    samples = worlds.sample_stream(1, 2000, c["lo"], c["hi"], c["goal_bias"], c["goal"])
    ow1, keep1 = orc.world(2, spheres=c["spheres"])
    tree, parents, solved, iters = orc.rrt(ow1, orc.space(2), c["start"], c["goal"], c["goal_radius"],
                                           c["max_distance"], res1, samples, 2001)
    samples = samples[:iters]
    # RRT*: wall world, count-bounded
    w = worlds.WALL
    res2 = motion_resolution(real_vector_max_extent(list(zip(w["lo"], w["hi"]))))
    s2 = worlds.sample_stream(12, 700, w["lo"], w["hi"], 0.0, w["goal"])
    ow2, keep2 = orc.world(2, aabbs=w["aabbs"])
    st = orc.rrt_star(ow2, orc.space(2), w["start"], w["goal"], w["goal_radius"], 0.5, 0.6, res2, s2,
                      701, stop_at_goal=False)
    np.savez_compressed(os.path.join(HERE, "planners.npz"), rrt_samples=samples, rrt_tree=tree,
                        rrt_parents=parents, rrt_solved=solved, rrt_iters=iters, rrt_resolution=res1,
                        star_samples=s2, star_tree=st["tree"], star_parents=st["parents"],
                        star_costs=st["costs"], star_resolution=res2,
                        star_motion_checks=st["motion_checks"])
    assert res1 == 0.14142135623730953 and math.isclose(res2, 0.07071067811865477)
    for f in ("scans.npz", "validity.npz", "planners.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


def main_so3():
    """SO(3): 1-NN / radius over unit quaternions, cone validity, check_motion (so3.npz).  acos / sin
    come from this machine's libm (glibc), the one Rust's f64 methods call on linux-gnu."""
    orc = oracle.get()
    rng = np.random.default_rng(77)

    def unit(n):
        q = rng.normal(size=(n, 4))
        return q / np.linalg.norm(q, axis=1, keepdims=True)

    rows, queries = unit(900), unit(30)
    rows[400] = rows[3]                                   # duplicate: lowest index wins
    rows[401] = -rows[5]                                  # antipodal: the same rotation
    queries[:4] = rows[[3, 5, 400, 17]]
    i1, d1 = orc.so3_nn1(rows, queries)
    off, ir, dr = orc.so3_radius(rows, queries, 0.4)
    cones = np.hstack([unit(6), rng.uniform(0.1, 0.4, size=(6, 1))])
    states = unit(500)
    ok = orc.so3_is_valid(cones, states)
    frm, to = unit(120), unit(120)
    to[:40] = frm[:40] + rng.normal(size=(40, 4)) * 0.01
    to[:40] /= np.linalg.norm(to[:40], axis=1, keepdims=True)
    res = 0.5 * math.pi * 0.05 * 0.1
    mv, mf = orc.so3_check_motion(cones, frm, to, res)
    np.savez_compressed(os.path.join(HERE, "so3.npz"), rows=rows, queries=queries, nn1_idx=i1,
                        nn1_dist=d1, rad_off=off, rad_idx=ir, rad_dist=dr, radius=0.4, cones=cones,
                        states=states, valid=ok, frm=frm, to=to, resolution=res, motion_valid=mv,
                        motion_first=mf)
    print("so3.npz", os.path.getsize(os.path.join(HERE, "so3.npz")), "bytes")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "so3":
        main_so3()
    else:
        main()
        main_so3()
