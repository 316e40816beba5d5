#!/usr/bin/env python
"""GPU probe: device time of the batched k-NN rows kernel for a given N / D / k, under the
current OXG_KNN_* environment switches.  Prints one JSON line.  (Measurement aid.)"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import oxmpl_b200 as oxg  # noqa: E402
from oxmpl_b200._lib import check, lib  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    rows_q = int(sys.argv[5]) if len(sys.argv) > 5 else n   # query rows [0, rows_q) (a shard)
    rng = np.random.default_rng(0)
    rows = rng.random((n, d)) * 5.0 - 2.5
    dev = torch.device("cuda", 0)
    st = oxg.StateStore(oxg.Space.real_vector(d), capacity=n)
    st.append(rows)
    ext = torch.cuda.ExternalStream(st.stream(), device=dev)
    idx = torch.empty((n, k), dtype=torch.int32, device=dev)
    dist = torch.empty((n, k), dtype=torch.float64, device=dev)
    times = []
    for it in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record(ext)
            check(lib.oxg_knn_rows_dev(st.handle, 0, rows_q, k, C.c_void_p(idx.data_ptr()),
                                       C.c_void_p(dist.data_ptr())))
            e1.record(ext)
        ext.synchronize()
        if it > 0:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    chk = int(idx[:rows_q].to(torch.int64).sum().item())
    print(json.dumps({"n": n, "d": d, "k": k, "rows": rows_q, "ms": ms,
                      "pairs_per_s": n * rows_q / (ms * 1e-3),
                      "filter": os.environ.get("OXG_KNN_FILTER", "1"),
                      "qpt": os.environ.get("OXG_KNN_QPT", "auto"), "idx_checksum": chk}))


if __name__ == "__main__":
    main()
