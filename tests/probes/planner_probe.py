#!/usr/bin/env python
"""GPU probe for the fused planner kernels at BASELINE configs 2 and 3 scale: iterations/s on
the GPU for the full run, and (cpu=True / cpu_prefix > 0) GPU vs the CPU oracle on the same
stream, parity-checked.  Prints JSON lines.  Test infrastructure: the oracle is imported only when
the CPU leg is requested (bench.py's extras run the GPU part; its cpu_baseline leg the CPU part)."""
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oxmpl_b200 as oxg  # noqa: E402
from oxmpl_b200 import worlds  # noqa: E402


def res_of(lo, hi):
    return oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(lo, hi))))


def config3(n_nodes, n_spheres, cpu_prefix):
    dim = 6
    lo, hi = [-1.0] * dim, [1.0] * dim
    start, goal = [-0.9] * dim, [0.9] * dim
    sph = worlds.sphere_world(3, lo, hi, n_spheres, 0.22, 0.32, keep_clear=[start, goal])
    sp = oxg.Space.real_vector(dim)
    res = res_of(lo, hi)
    n_samples = int(n_nodes * 1.6)
    samples = worlds.sample_stream(33, n_samples, lo, hi, 0.0, goal)
    max_distance = 0.2
    # search radius s.t. the mean neighbour count at n_nodes is ~30 (unit-ball volume in 6-D = pi^3/6)
    free = float(oxg.World(dim, spheres=sph).is_valid(sp, samples[:20000]).mean())
    vol = (2.0 ** dim) * free
    search_radius = (30.0 * vol / n_nodes / (math.pi ** 3 / 6.0)) ** (1.0 / dim)
    out = {"config": "3: RRT* 6-D", "spheres": n_spheres, "free_fraction": free,
           "search_radius": search_radius, "max_distance": max_distance, "resolution": res}
    with oxg.World(dim, spheres=sph) as gw, oxg.StateStore(sp) as st:
        st.append(np.array([start]))
        t0 = time.perf_counter()
        parents, costs, r = oxg.rrt_star_solve(gw, st, samples, max_distance, search_radius, res,
                                               n_nodes, stop_at_goal=False)
        wall = time.perf_counter() - t0
        out.update({"gpu_nodes": r["n_nodes"][0], "gpu_iters": r["iters"], "gpu_ms": r["device_ms"],
                    "gpu_wall_ms": wall * 1e3, "gpu_iters_per_s": r["iters"] / (r["device_ms"] * 1e-3),
                    "gpu_us_per_iter": r["device_ms"] * 1e3 / r["iters"], "status": r["status"],
                    "motion_checks_ref": r["motion_checks"],
                    "motion_checks_executed": r["motion_checks_executed"],
                    "motion_checks_per_s": r["motion_checks"] / (r["device_ms"] * 1e-3),
                    "ctas": r["ctas"], "phase_ms": r["phase_ms"]})
    if cpu_prefix:
        from oracle import oracle

        orc = oracle.get()
        sp_o = orc.space(dim)
        ow, keep = orc.world(dim, spheres=sph)
        t0 = time.perf_counter()
        o = orc.rrt_star(ow, sp_o, start, None, -1.0, max_distance, search_radius, res,
                         samples[:cpu_prefix], cpu_prefix + 1, stop_at_goal=False)
        cpu_s = time.perf_counter() - t0
        with oxg.World(dim, spheres=sph) as gw, oxg.StateStore(sp) as st:
            st.append(np.array([start]))
            parents, costs, r = oxg.rrt_star_solve(gw, st, samples[:cpu_prefix], max_distance,
                                                   search_radius, res, cpu_prefix + 1,
                                                   stop_at_goal=False)
            same = (np.array_equal(st.read().view(np.uint64), o["tree"].view(np.uint64))
                    and np.array_equal(parents, o["parents"])
                    and np.array_equal(costs.view(np.uint64), o["costs"].view(np.uint64)))
        out.update({"prefix_iters": cpu_prefix, "prefix_cpu_s": cpu_s,
                    "prefix_cpu_iters_per_s": cpu_prefix / cpu_s,
                    "prefix_gpu_ms": r["device_ms"],
                    "prefix_gpu_iters_per_s": cpu_prefix / (r["device_ms"] * 1e-3),
                    "prefix_bit_identical": bool(same)})
    return out


def config2(n_samples, cpu=True):
    lo, hi = [0.0, 0.0], [100.0, 100.0]
    start, goal = [5.0, 5.0, 0.0], [95.0, 95.0, 1.0]
    boxes = worlds.aabb_world(2, lo, hi, 1000, 1.0, 3.5, keep_clear=[start[:2], goal[:2]])
    rng = np.random.default_rng(2)
    samples = np.column_stack([rng.uniform(0, 100, n_samples), rng.uniform(0, 100, n_samples),
                               rng.uniform(-math.pi, math.pi, n_samples)])
    sp = oxg.Space.compound(2, 1, 1.0, 0.5)
    res = res_of(lo, hi)
    with oxg.World(2, aabbs=boxes) as gw, oxg.StateStore(sp) as ta, oxg.StateStore(sp) as tb:
        ta.append(np.array([start]))
        tb.append(np.array([goal]))
        pa, pb, r = oxg.rrt_connect_solve(gw, ta, tb, samples, 2.0, res, n_samples,
                                          goal_target=goal, goal_radius=1.0)
        tree_a, tree_b = ta.read(), tb.read()
    out = {"config": "2: RRT-Connect R2xSO2, 1000 AABBs", "iters": r["iters"],
           "solved": r["solved"], "nodes": r["n_nodes"], "gpu_ms": r["device_ms"],
           "gpu_us_per_iter": r["device_ms"] * 1e3 / max(r["iters"], 1), "ctas": r["ctas"],
           "phase_us_per_iter": [round(v * 1e3 / max(r["iters"], 1), 3) for v in r["phase_ms"]]}
    if cpu:
        from oracle import oracle

        orc = oracle.get()
        sp_o = orc.space(2, 1, 1.0, 0.5)
        ow, keep = orc.world(2, aabbs=boxes)
        t0 = time.perf_counter()
        o = orc.rrt_connect(ow, sp_o, start, goal, goal, 1.0, 2.0, res, samples, n_samples)
        cpu_s = time.perf_counter() - t0
        same = (np.array_equal(tree_a.view(np.uint64), o["start_tree"].view(np.uint64))
                and np.array_equal(tree_b.view(np.uint64), o["goal_tree"].view(np.uint64))
                and np.array_equal(pa, o["start_parents"]) and np.array_equal(pb, o["goal_parents"]))
        out.update({"cpu_s": cpu_s, "cpu_us_per_iter": cpu_s * 1e6 / max(o["iters"], 1),
                    "bit_identical": bool(same)})
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "3"
    if which == "2":
        print(json.dumps(config2(int(sys.argv[2]) if len(sys.argv) > 2 else 30000)), flush=True)
    else:
        print(json.dumps(config3(int(sys.argv[2]) if len(sys.argv) > 2 else 200000,
                                 int(sys.argv[3]) if len(sys.argv) > 3 else 10000,
                                 int(sys.argv[4]) if len(sys.argv) > 4 else 3000)), flush=True)
