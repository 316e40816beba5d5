#!/usr/bin/env python
"""Small end-to-end pass over every kernel family, meant to be run under
`compute-sanitizer --tool memcheck` (sizes chosen so the run stays short)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oxmpl_b200 as oxg  # noqa: E402
from oxmpl_b200 import worlds  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    for d in (2, 7, 16):
        sp = oxg.Space.real_vector(d)
        with oxg.StateStore(sp) as st:
            rows = rng.random((2501, d))
            st.append(rows[:1000])
            st.append(rows[1000:])          # growth + second append
            for q in (1, 5, 67, 1100):
                qs = rng.random((q, d))
                st.nn1(qs)
                st.knn(qs, 10)
                st.radius(qs, 0.3 * np.sqrt(d), want_dist=True)
            st.knn(rng.random((3, d)), 64)
    sp = oxg.Space.compound(2, 1)
    with oxg.StateStore(sp) as st:
        st.append(rng.random((777, 3)))
        st.nn1(rng.random((9, 3)))
        st.knn(rng.random((9, 3)), 4)
        st.radius(rng.random((9, 3)), 0.4)
    spheres, aabbs, capsules = worlds.calibrated_obstacles(3, 7, 40, 30, 50)
    sp = oxg.Space.real_vector(7)
    with oxg.World(7, spheres=spheres, aabbs=aabbs, capsules=capsules) as gw, oxg.StateStore(sp) as st:
        a = rng.uniform(-1, 1, (3000, 7))
        b = a + rng.normal(scale=0.2, size=a.shape)
        gw.is_valid(sp, a)
        gw.is_valid(sp, a[:5])
        gw.check_motion(sp, a, b, 0.02, want_first_invalid=True)
        gw.check_motion(sp, a[:9], b[:9], 0.02, want_first_invalid=True)
        ok = gw.is_valid(sp, a).astype(bool)
        st.append(a[ok])
        oxg.prm_build_knn(gw, st, 10, 0.03)
        oxg.prm_build_radius(gw, st, 0.9, 0.03)
    c = worlds.CONFIG1
    sp2 = oxg.Space.real_vector(2)
    samples = worlds.sample_stream(1, 3000, c["lo"], c["hi"], 0.05, c["goal"])
    for ctas in ("1", "7"):
        os.environ["OXG_PLAN_CTAS"] = ctas
        with oxg.World(2, spheres=c["spheres"]) as gw, oxg.StateStore(sp2) as t0, oxg.StateStore(sp2) as t1:
            t0.append(np.array([c["start"]]))
            oxg.rrt_solve(gw, t0, samples, 0.5, 0.14, 3001, goal_target=c["goal"], goal_radius=0.5)
            t0.clear()
            t0.append(np.array([c["start"]]))
            oxg.rrt_star_solve(gw, t0, samples[:800], 0.5, 1.0, 0.14, 801, stop_at_goal=False)
            t0.clear()
            t0.append(np.array([c["start"]]))
            t1.append(np.array([c["goal"]]))
            oxg.rrt_connect_solve(gw, t0, t1, samples, 0.5, 0.14, 3001, goal_target=c["goal"],
                                  goal_radius=0.5)
    print("sanitize probe done")


if __name__ == "__main__":
    main()
