"""Times the SO(3) scans on the device (oxg_store_last_scan_ms brackets the kernels of the last call).
usage: python tools/so3_probe.py  -> one JSON line"""
import json
import os
import sys

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
        out[f"n{n}_q{q}"] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()
