"""Device time of k-NN batches of a few hundred to a few thousand queries (the regime between the
few-query streaming kernels and the machine-filling batched kernels).
usage: python tools/medq_probe.py [n ...]      env OXG_KNN_QPT=1|2 forces queries per thread"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import oxmpl_b200 as oxg  # noqa: E402

rng = np.random.default_rng(0)
d = 7
out = {"qpt_env": os.environ.get("OXG_KNN_QPT", "")}
for n in [int(float(a)) for a in sys.argv[1:]] or [1_000_000]:
    st = oxg.StateStore(oxg.Space.real_vector(d), capacity=n)
    st.append(rng.random((n, d)))
    for Q in (100, 256, 512, 1000, 1024, 2048, 4096, 10000, 40000):
        q = rng.random((Q, d))
        for k in (1, 10):
            best = None
            for _ in range(3):
                st.knn(q, k)
                ms = st.last_scan_ms()
                best = ms if best is None else min(best, ms)
            out[f"n{n}_q{Q}_k{k}"] = [round(best, 3), float("%.3g" % (n * Q / best * 1e3))]
    st.close()
print(json.dumps(out))
