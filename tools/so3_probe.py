"""Times the SO(3) scans on the device (oxg_store_last_scan_ms brackets the kernels of the last call).
usage: python tools/so3_probe.py  -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oxmpl_b200 as oxg  # noqa: E402


def unit(rng, n):
    q = rng.normal(size=(n, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def main():
    rng = np.random.default_rng(0)
    out = {}
    for n, q in ((10_000_000, 1), (10_000_000, 8), (1_000_000, 1000), (100_000, 100_000)):
        rows, qs = unit(rng, n), unit(rng, q)
        with oxg.StateStore(oxg.Space.real_vector(4)) as st:
            st.append(rows)
            best = None
            for _ in range(4):
                st.so3_nn1(qs)
                ms = st.last_scan_ms()
                best = ms if best is None else min(best, ms)
            rec = {"nn1_ms": round(best, 4), "store_GBps": round(n * 32 / best / 1e6, 1),
                   "pairs_per_s": round(n * q / best * 1e3, 0)}
            best = None
            for _ in range(3):
                off, idx = st.so3_radius(qs[: min(q, 20000)], 0.2)
                ms = st.last_scan_ms()
                best = ms if best is None else min(best, ms)
            rec["radius_ms"] = round(best, 4)
            rec["radius_hits"] = int(off[-1])
            if q >= 8:
                for _ in range(2):                                   # second call: buffers exist
                    t0 = time.perf_counter()
                    st.so3_knn(qs[: min(q, 20000)], 10)
                    rec["knn10_wall_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        out[f"n{n}_q{q}"] = rec
    out["rrt"] = rrt_loop(rng)
    print(json.dumps(out))


def rrt_loop(rng):
    """RRT over SO3, 20 000 iterations, no goal: one C-ABI call per step of every iteration
    (round 1's loop) against oxg_so3_rrt_solve's speculative chunks; same tree."""
    import math
    res = 0.5 * math.pi * 0.05 * 0.1
    cones = np.array([[0.0, 0.0, 0.0, 1.0, math.radians(44.9)]])
    c = oxg.So3Cones(cones)
    start = np.array([[1.0, 0.0, 0.0, 0.0]])
    samples = unit(rng, 20000)
    rec = {}
    with oxg.StateStore(oxg.Space.real_vector(4)) as st:
        st.append(start)
        oxg.so3_rrt_solve(c, st, samples[:64], 0.3, res, 30000)      # warm-up
    with oxg.StateStore(oxg.Space.real_vector(4)) as st:
        st.append(start)
        t0 = time.perf_counter()
        par, r = oxg.so3_rrt_solve(c, st, samples, 0.3, res, 30000)
        dt = time.perf_counter() - t0
        rec["chunked"] = {"iters_per_s": round(r["iters"] / dt), "nodes": len(par), "chunks": r["ctas"],
                          "motion_checks_launched": r["motion_checks_executed"]}
        rows_chunked = st.read()
    n_loop = 3000
    with oxg.StateStore(oxg.Space.real_vector(4)) as st:
        st.append(start)
        nodes = [start[0]]
        t0 = time.perf_counter()
        for q in samples[:n_loop]:
            idx, dist = st.so3_nn1(q[None])
            near, d = int(idx[0]), float(dist[0])
            q_new = slerp(nodes[near], q, 0.3 / d) if d > 0.3 else q
            if c.check_motion(nodes[near][None], q_new[None], res)[0]:
                st.append(q_new[None])
                nodes.append(q_new)
        dt = time.perf_counter() - t0
        rec["per_iteration_calls"] = {"iters_per_s": round(n_loop / dt), "nodes": len(nodes)}
        rec["same_tree_prefix"] = bool((st.read().view(np.uint64) ==
                                        rows_chunked[: len(nodes)].view(np.uint64)).all())
    # RRT-Connect, no goal ball, trees rooted at antipodal-ish states; 0.3 rad steps
    goal_root = np.array([[0.0, 1.0, 0.0, 0.0]])
    for md in (0.3, 0.05):
        with oxg.StateStore(oxg.Space.real_vector(4)) as sa, oxg.StateStore(oxg.Space.real_vector(4)) as sb:
            sa.append(start)
            sb.append(goal_root)
            t0 = time.perf_counter()
            pa, pb, r = oxg.so3_rrt_connect_solve(c, sa, sb, samples, md, res, 30000)
            dt = time.perf_counter() - t0
            rec[f"connect_chunked_step{md}"] = {
                "iters": r["iters"], "iters_per_s": round(r["iters"] / dt), "chunks": r["ctas"],
                "solved": r["solved"], "nodes": [len(pa), len(pb)],
                "motion_checks": r["motion_checks"], "launched": r["motion_checks_executed"]}
    return rec


def slerp(f, g, t):
    from oxmpl_b200.base import SO3State, SO3StateSpace
    return np.asarray(SO3StateSpace().interpolate(SO3State(*f), SO3State(*g), t).values)


if __name__ == "__main__":
    main()
