#!/usr/bin/env python
"""BASELINE config 5: batched NN / validity sweep — N stored states x D dims x Q queries per
batch, ops nn1 / knn(k=10) / radius / is_valid / check_motion — GPU (liboxg through the C ABI,
host buffers) next to the CPU oracle on a bounded sample of the same batch, with the sampled
rows compared bit for bit.  Writes one JSON document (default profiles/r01_sweep.json).

The oracle is only the checker / CPU baseline here (test infrastructure)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oxmpl_b200 as oxg  # noqa: E402
from oracle import oracle  # noqa: E402
from oxmpl_b200 import worlds  # noqa: E402

PAIR_CAP = 2.0e12        # Q*N cap per batch
CPU_PAIR_BUDGET = 6.0e7  # oracle pairs per measured point (~0.3 s single thread)


def best_of(fn, reps=3):
    ts = []
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_sweep.json")
    quick = os.environ.get("SWEEP_QUICK") == "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    orc = oracle.get()
    try:  # the oracle's batch entry points are OpenMP loops: pin them to ONE thread (the
        # reference is single-threaded), as bench.py's cpu_baseline does
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(1)
    except OSError:
        pass
    Ns = [10_000, 100_000, 1_000_000, 10_000_000]
    Ds = [2, 3, 6, 7, 14]
    Qs = [1, 100, 10_000, 1_000_000]
    if quick:
        Ns, Ds, Qs = [10_000, 1_000_000], [2, 7], [1, 10_000]
    rows_out = []
    rng = np.random.default_rng(5)
    for D in Ds:
        sp, sp_o = oxg.Space.real_vector(D), orc.space(D)
        for N in Ns:
            store_rows = rng.random((N, D))
            with oxg.StateStore(sp, capacity=N) as st:
                for a in range(0, N, 1_000_000):
                    st.append(store_rows[a:a + 1_000_000])
                # radius giving ~10 neighbours per query under uniform density
                from math import gamma, pi
                ball = pi ** (D / 2) / gamma(D / 2 + 1)
                r = (10.0 / N / ball) ** (1.0 / D)
                for Q in Qs:
                    if Q * N > PAIR_CAP:
                        continue
                    queries = rng.random((Q, D))
                    q_cpu = int(max(1, min(Q, CPU_PAIR_BUDGET // N)))
                    sel = np.sort(rng.choice(Q, size=q_cpu, replace=False))
                    rec = {"N": N, "D": D, "Q": Q, "cpu_sample_queries": q_cpu}
                    # ---- nn1
                    t, (gi, gd) = best_of(lambda: st.nn1(queries))
                    tc, (ci, cd) = best_of(lambda: orc.nn1(sp_o, store_rows, queries[sel]), 1)
                    rec["nn1"] = {"gpu_s": t, "gpu_device_ms": st.last_scan_ms(),
                                  "gpu_queries_per_s": Q / t, "cpu_queries_per_s": q_cpu / tc,
                                  "bit_exact": bool(np.array_equal(gi[sel], ci) and
                                                    np.array_equal(gd[sel].view(np.uint64),
                                                                   cd.view(np.uint64)))}
                    # ---- knn k=10
                    k = min(10, N)
                    t, (gi, gd) = best_of(lambda: st.knn(queries, k))
                    tc, (ci, cd) = best_of(lambda: orc.knn(sp_o, store_rows, queries[sel], k), 1)
                    rec["knn10"] = {"gpu_s": t, "gpu_device_ms": st.last_scan_ms(),
                                    "gpu_queries_per_s": Q / t, "cpu_queries_per_s": q_cpu / tc,
                                    "bit_exact": bool(np.array_equal(gi[sel], ci) and
                                                      np.array_equal(gd[sel].view(np.uint64),
                                                                     cd.view(np.uint64)))}
                    # ---- radius
                    t, (go, gx) = best_of(lambda: st.radius(queries, r))
                    tc, (co, cx) = best_of(lambda: orc.radius(sp_o, store_rows, queries[sel], r), 1)
                    same = all(np.array_equal(gx[int(go[s]):int(go[s + 1])],
                                              cx[int(co[i]):int(co[i + 1])])
                               for i, s in enumerate(sel))
                    rec["radius"] = {"gpu_s": t, "gpu_queries_per_s": Q / t,
                                     "cpu_queries_per_s": q_cpu / tc, "mean_neighbours":
                                     float(go[-1]) / Q, "bit_exact": bool(same)}
                    rows_out.append(rec)
                    print(json.dumps(rec), flush=True)
    # ---- validity / motion: D x obstacle count x batch size
    val_out = []
    for D in ([2, 7] if quick else [2, 6, 7, 14]):
        sp, sp_o = oxg.Space.real_vector(D), orc.space(D)
        for n_obs in ([256] if quick else [16, 256, 4096]):
            third = n_obs // 3 + 1
            spheres, aabbs, capsules = worlds.calibrated_obstacles(D, D, third, third, third,
                                                                   p_each=0.3 / n_obs)
            ow, keep = orc.world(D, spheres=spheres, aabbs=aabbs, capsules=capsules)
            with oxg.World(D, spheres=spheres, aabbs=aabbs, capsules=capsules) as gw:
                for B in ([10_000] if quick else [1, 100, 10_000, 1_000_000]):
                    states = rng.uniform(-1, 1, (B, D))
                    to = states + rng.normal(scale=0.2, size=states.shape)
                    res = oxg.motion_resolution(oxg.real_vector_max_extent([(-1.0, 1.0)] * D))
                    b_cpu = int(max(1, min(B, 3.0e7 // (n_obs * D))))
                    t, gv = best_of(lambda: gw.is_valid(sp, states))
                    tc, cv = best_of(lambda: orc.is_valid(ow, sp_o, states[:b_cpu]), 1)
                    t2, gm = best_of(lambda: gw.check_motion(sp, states, to, res))
                    b_cpu2 = max(1, b_cpu // 8)
                    tc2, (cm, _) = best_of(lambda: orc.check_motion(ow, sp_o, states[:b_cpu2],
                                                                    to[:b_cpu2], res), 1)
                    rec = {"D": D, "obstacles": 3 * third, "batch": B,
                           "is_valid": {"gpu_states_per_s": B / t, "cpu_states_per_s": b_cpu / tc,
                                        "bit_exact": bool(np.array_equal(gv[:b_cpu], cv))},
                           "check_motion": {"gpu_edges_per_s": B / t2,
                                            "cpu_edges_per_s": b_cpu2 / tc2,
                                            "bit_exact": bool(np.array_equal(gm[:b_cpu2], cm))}}
                    val_out.append(rec)
                    print(json.dumps(rec), flush=True)
    doc = {"what": "BASELINE config 5 sweep; GPU = host-buffer C-ABI calls (H2D + kernels + D2H, "
                   "best of 3); CPU = oracle, single thread, bounded sample of the same batch",
           "host_cores": os.cpu_count(), "cpu_threads_used": 1, "scans": rows_out, "validity": val_out,
           "all_bit_exact": all(r[k]["bit_exact"] for r in rows_out for k in ("nn1", "knn10", "radius"))
           and all(r[k]["bit_exact"] for r in val_out for k in ("is_valid", "check_motion"))}
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1)
    print("all_bit_exact", doc["all_bit_exact"], "->", out_path)


if __name__ == "__main__":
    main()
