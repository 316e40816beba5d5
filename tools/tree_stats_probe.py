"""GPU probe: counters of the hierarchy k-NN on the config-4 store (leaves / box groups / candidates /
merges per query, kernel ms) for both state orders.  usage: python tools/tree_stats_probe.py [n] [k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oxmpl_b200 as oxg
from oxmpl_b200 import worlds


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rng = np.random.default_rng(4)
    lo, hi = np.asarray(worlds.PANDA_LO), np.asarray(worlds.PANDA_HI)
    rows = lo + (hi - lo) * rng.random((n, 7))
    oxg.tuning_set("knn_tree", 1)
    oxg.tuning_set("tree_stats", 2)
    with oxg.StateStore(oxg.Space.real_vector(7)) as st:
        st.append(rows)
        for kind in (0, 1):
            oxg.tuning_set("tree_kind", kind)
            oxg.tuning_set("tree_cache", 0)
            for rep in range(2):
                st.knn(rows, k, exclude=np.arange(n, dtype=np.uint32))
            s = st.tree_stats()
            q = s["queries"]
            print(f"kind={kind} build_ms={s['build_ms']:.3f} kernel_ms={s['query_kernel_ms']:.3f} "
                  f"leaves/q={s['leaves_visited'] / q:.1f} groups/q={s['box_groups_tested'] / q:.1f} "
                  f"cand/q={s['candidates'] / q:.1f} merges/q={s['merges'] / q:.2f} exact leaves/q={s['exact_leaves'] / q:.1f}", flush=True)


if __name__ == "__main__":
    main()
