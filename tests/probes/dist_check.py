"""Multi-rank GPU check of the library's sharded PRM builds (run under torchrun, one rank per
GPU): oxg_prm_build_knn_dist / oxg_prm_build_radius_dist must return, on every rank, exactly what
the single-GPU entry points return.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/probes/dist_check.py [n]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    import torch
    import torch.distributed as dist

    import oxmpl_b200 as oxg
    from oxmpl_b200 import sharding, worlds
    from oxmpl_b200._lib import check, lib

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    check(lib.oxg_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_003
    k = 10
    rng = np.random.default_rng(4)
    rows = worlds.PANDA_LO + (worlds.PANDA_HI - worlds.PANDA_LO) * rng.random((n, 7))
    caps = worlds.capsule_world(4, worlds.PANDA_LO, worlds.PANDA_HI, 256)
    res = oxg.motion_resolution(oxg.real_vector_max_extent(list(zip(worlds.PANDA_LO, worlds.PANDA_HI))))
    sp = oxg.Space.real_vector(7)
    ok = True
    with oxg.World(7, capsules=caps) as gw, oxg.StateStore(sp) as st:
        st.append(rows)
        d = sharding.Dist.from_torch(dist, rank, world)
        for nn in (n, n - (n % world)):          # uneven and even shards
            if nn != n:
                st.clear()
                st.append(rows[:nn])
            ri, rd, rv, _ = oxg.prm_build_knn(gw, st, k, res)
            gi, gd, gv, stats = d.prm_build_knn(gw, st, k, res)
            same = (np.array_equal(ri, gi) and np.array_equal(rd.view(np.uint64), gd.view(np.uint64))
                    and np.array_equal(rv, gv))
            checked = torch.tensor([stats["n_edges_checked"]], dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(checked)
            pairs = {(min(i, int(j)), max(i, int(j))) for i in range(0, nn, 97) for j in ri[i]}
            print(f"rank {rank}/{world} n={nn}: knn build identical={same} ms={stats['knn_ms']:.2f} "
                  f"edges checked (all ranks)={int(checked)}", flush=True)
            ok = ok and same
        # collective append: every rank uploads only its slice, the rest arrives over NVLink
        st.clear()
        rb, re = d.row_range(n)
        d.append(st, rows[rb:re])
        same = len(st) == n and np.array_equal(st.read().view(np.uint64), rows.view(np.uint64))
        gi2, _, _, _ = d.prm_build_knn(gw, st, k, res)
        st.clear()
        st.append(rows)
        ri2, _, _, _ = oxg.prm_build_knn(gw, st, k, res)
        same = same and np.array_equal(gi2, ri2)
        print(f"rank {rank}/{world}: sharded append identical={same}", flush=True)
        ok = ok and same
        # radius mode on a smaller store (brute-force scans)
        st.clear()
        m = 20_001
        st.append(rows[:m])
        ro, rx, _ = oxg.prm_build_radius(gw, st, 1.1, res)
        go, gx = d.prm_build_radius(gw, st, 1.1, res)
        same = np.array_equal(ro, go) and np.array_equal(rx, gx)
        print(f"rank {rank}/{world}: radius build identical={same} entries={len(gx)} "
              f"rows={d.row_range(m, True)}", flush=True)
        ok = ok and same
        d.close()
    if world > 1:
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(int(t))
        dist.destroy_process_group()
    print(f"rank {rank}: {'DIST_CHECK_OK' if ok else 'DIST_CHECK_FAILED'}", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
