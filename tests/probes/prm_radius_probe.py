#!/usr/bin/env python
"""SURVEY §8(d) config 4, second half: the reference-faithful radius PRM
(PRM::construct_roadmap, prm.rs:117-146 — node i connects to every earlier node within
connection_radius whose motion is valid) at N = 1e5 in the 7-D capsule world: GPU batched build
(oxg_prm_build_radius) vs the oracle's sample-by-sample loop, CSR compared bit for bit."""
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


def main():
    n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 122_000   # ~1e5 valid
    radius = float(sys.argv[2]) if len(sys.argv) > 2 else 1.05
    orc = oracle.get()
    caps = worlds.capsule_world(4, worlds.PANDA_LO, worlds.PANDA_HI, 256)
    rng = np.random.default_rng(4)
    samples = worlds.PANDA_LO + (worlds.PANDA_HI - worlds.PANDA_LO) * rng.random((n_samples, 7))
    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(worlds.PANDA_LO, worlds.PANDA_HI))))
    out = {"config": "4b: radius PRM (reference-faithful), 7-D, 256 capsules", "samples": n_samples,
           "connection_radius": radius, "resolution": res}
    with oxg.World(7, capsules=caps) as gw, oxg.StateStore(sp) as st:
        t0 = time.perf_counter()
        ok = gw.is_valid(sp, samples).astype(bool)          # prm.rs:123
        st.append(samples[ok])
        off, idx, stats = oxg.prm_build_radius(gw, st, radius, res)
        gpu_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        off, idx, stats = oxg.prm_build_radius(gw, st, radius, res)
        gpu_s2 = time.perf_counter() - t0
        out.update({"nodes": int(ok.sum()), "gpu_wall_s_incl_filter_upload": gpu_s,
                    "gpu_build_s": gpu_s2, "scan_ms": stats["knn_ms"], "edge_ms": stats["edge_ms"],
                    "candidates": stats["n_candidates"], "edges_valid": stats["n_edges_valid"],
                    "mean_degree": 2.0 * stats["n_edges_valid"] / max(int(ok.sum()), 1)})
    if os.environ.get("SKIP_CPU") != "1":
        ow, keep = orc.world(7, capsules=caps)
        t0 = time.perf_counter()
        nodes, ooff, oidx, pairs, mcs = orc.prm_radius(ow, sp_o, radius, res, samples)
        cpu_s = time.perf_counter() - t0
        out.update({"cpu_s_single_thread": cpu_s, "cpu_pairs": pairs, "cpu_motion_checks": mcs,
                    "bit_identical": bool(np.array_equal(off, ooff) and np.array_equal(idx, oidx)
                                          and np.array_equal(nodes, samples[ok])),
                    "speedup_build": cpu_s / gpu_s2})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
