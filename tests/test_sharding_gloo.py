"""CPU, world_size 2 (gloo): the host-side sharding of the PRM build — row partition, uneven
all-gather, single ownership of every candidate edge — reproduces the unsharded result.  The
per-shard compute is stood in for by the CPU oracle (test infrastructure); on the GPU box the
same host logic drives liboxg (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from oracle import oracle
        from oxmpl_b200 import sharding, worlds
        from oxmpl_b200.api import motion_resolution, real_vector_max_extent

        orc = oracle.get()
        rng = np.random.default_rng(4)
        rows = worlds.PANDA_LO + (worlds.PANDA_HI - worlds.PANDA_LO) * rng.random((n, 7))
        caps = worlds.capsule_world(4, worlds.PANDA_LO, worlds.PANDA_HI, 48)
        res = motion_resolution(real_vector_max_extent(list(zip(worlds.PANDA_LO, worlds.PANDA_HI))))
        sp = orc.space(7)
        ow, keep = orc.world(7, capsules=caps)

        rb, re = sharding.row_range(n, rank, world_size)
        idx_all = torch.zeros((n, k), dtype=torch.int32)
        dist_all = torch.zeros((n, k), dtype=torch.float64)
        valid_all = torch.full((n, k), 2, dtype=torch.uint8)
        rmax = sharding.rows_max(n, world_size)
        pads = [torch.zeros((world_size, rmax, k), dtype=t) for t in
                (torch.int32, torch.float64, torch.uint8)]
        # phase 1: this rank's k-NN rows
        qi = np.arange(rb, re, dtype=np.uint32)
        i_loc, d_loc = orc.knn(sp, rows, rows[rb:re], k, exclude=qi)
        idx_all[rb:re] = torch.from_numpy(i_loc.view(np.int32))
        dist_all[rb:re] = torch.from_numpy(d_loc)
        sharding.all_gather_rows(dist, idx_all, rank, world_size, pads[0])
        sharding.all_gather_rows(dist, dist_all, rank, world_size, pads[1])
        # phase 2: motion-check the candidate edges this shard owns
        idx_np = idx_all.numpy().view(np.uint32)
        own = sharding.owned_slots(idx_np, rb, re)
        ii, ss = np.nonzero(own)
        gi = ii + rb
        gj = idx_np[gi, ss].astype(np.int64)
        frm, to = rows[np.maximum(gi, gj)], rows[np.minimum(gi, gj)]
        v, _ = orc.check_motion(ow, sp, frm, to, res)
        loc = np.full((re - rb, k), 2, dtype=np.uint8)
        loc[idx_np[rb:re] == 0xFFFFFFFF] = 0
        loc[ii, ss] = v
        valid_all[rb:re] = torch.from_numpy(loc)
        sharding.all_gather_rows(dist, valid_all, rank, world_size, pads[2])
        # phase 3: resolve mirror-owned slots from the owner's value
        val = valid_all.numpy().copy()
        for i, s in zip(*np.nonzero(val == 2)):
            j = int(idx_np[i, s])
            (t,) = np.nonzero(idx_np[j] == i)
            val[i, s] = val[j, t[0]]
        n_checked = torch.tensor([len(ii)], dtype=torch.int64)
        dist.all_reduce(n_checked)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), idx=idx_np, dist=dist_all.numpy(),
                 valid=val, n_checked=n_checked.numpy(), own_rows=np.array([rb, re]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(600, 10), (601, 6)])
def test_sharded_prm_equals_unsharded(tmp_path, n, k):
    from oracle import oracle
    from oxmpl_b200 import sharding, worlds
    from oxmpl_b200.api import motion_resolution, real_vector_max_extent

    world_size = 2
    port = _free_port()
    mp.spawn(_worker, args=(world_size, port, n, k, str(tmp_path)), nprocs=world_size, join=True)

    orc = oracle.get()
    rng = np.random.default_rng(4)
    rows = worlds.PANDA_LO + (worlds.PANDA_HI - worlds.PANDA_LO) * rng.random((n, 7))
    caps = worlds.capsule_world(4, worlds.PANDA_LO, worlds.PANDA_HI, 48)
    res = motion_resolution(real_vector_max_extent(list(zip(worlds.PANDA_LO, worlds.PANDA_HI))))
    ow, keep = orc.world(7, capsules=caps)
    oi, od, ov = orc.prm_knn(ow, orc.space(7), rows, k, res)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world_size)]
    for o in outs:                       # every rank ends with the complete, identical roadmap
        np.testing.assert_array_equal(o["idx"], oi)
        np.testing.assert_array_equal(o["dist"].view(np.uint64), od.view(np.uint64))
        np.testing.assert_array_equal(o["valid"], ov)
    # every undirected candidate pair was motion-checked exactly once across the ranks
    pairs = {(min(i, int(j)), max(i, int(j))) for i in range(n) for j in oi[i] if j != 0xFFFFFFFF}
    assert int(outs[0]["n_checked"][0]) == len(pairs)
    # and the ownership rule balances the shards
    own = [int(sharding.owned_slots(oi, *sharding.row_range(n, r, world_size)).sum())
           for r in range(world_size)]
    assert abs(own[0] - own[1]) < 0.1 * sum(own)


def test_row_ranges_cover_and_owner_rule():
    from oxmpl_b200 import sharding

    for n in (1, 7, 1000, 1_000_003):
        for w in (1, 2, 4, 8):
            r = [sharding.row_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert sharding.rows_max(n, w) == max(b - a for a, b in r)
    i = np.array([3, 4, 10, 11]); j = np.array([8, 9, 2, 5])
    o = sharding.edge_owner(i, j)
    np.testing.assert_array_equal(o, sharding.edge_owner(j, i))     # symmetric
    np.testing.assert_array_equal(o, [3, 4, 10, 11])  # 3^8 odd->lo 3; 4^9 odd->lo 4; 10^2 even->hi; 11^5 even->hi


def _range_worker(rank, world_size, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        import ctypes as C

        from oxmpl_b200 import sharding
        from oxmpl_b200._lib import lib

        got = []
        for radius_mode in (0, 1):
            rb, re = C.c_uint64(0), C.c_uint64(0)
            assert lib.oxg_dist_row_range(n, rank, world_size, radius_mode, C.byref(rb), C.byref(re)) == 0
            got += [rb.value, re.value]
        assert tuple(got[:2]) == sharding.row_range(n, rank, world_size)
        assert tuple(got[2:]) == sharding.radius_row_range(n, rank, world_size)
        t = torch.tensor(got, dtype=torch.int64)
        allr = [torch.zeros(4, dtype=torch.int64) for _ in range(world_size)]
        dist.all_gather(allr, t)
        np.save(os.path.join(out_dir, f"ranges{rank}.npy"), torch.stack(allr).numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [601, 1_000_000])
def test_library_row_ranges_partition_the_rows(tmp_path, n):
    """The library's own partition (oxg_dist_row_range, used by oxg_prm_build_knn_dist /
    oxg_prm_build_radius_dist) seen from two ranks: both modes tile [0, n) without gaps, and the
    radius-mode cut balances the triangular work (node i scans i earlier nodes)."""
    world_size = 2
    mp.spawn(_range_worker, args=(world_size, _free_port(), n, str(tmp_path)), nprocs=world_size, join=True)
    a = np.load(tmp_path / "ranges0.npy")
    np.testing.assert_array_equal(a, np.load(tmp_path / "ranges1.npy"))
    for c in (0, 2):
        assert a[0, c] == 0 and a[-1, c + 1] == n and a[0, c + 1] == a[1, c]
    work = [(a[r, 3] ** 2 - a[r, 2] ** 2) / 2 for r in range(world_size)]
    assert abs(work[0] - work[1]) <= 0.01 * sum(work) + 2 * n


def test_dist_entry_points_fail_loudly_without_a_device():
    """No CPU fallback: without a GPU the communicator cannot be created; bad arguments are
    rejected before anything else."""
    import ctypes as C

    import oxmpl_b200 as oxg
    from oxmpl_b200._lib import OXG_E_ARG, OXG_E_CUDA, OXG_E_NULL, lib

    h = C.c_void_p()
    assert lib.oxg_dist_create(None, 2, 2, C.byref(h)) == OXG_E_ARG
    assert lib.oxg_dist_create(None, 0, 2, C.byref(h)) == OXG_E_NULL       # world > 1 needs the id
    assert lib.oxg_dist_row_range(10, 0, 1, 0, None, None) == OXG_E_NULL
    if oxg.device_count() == 0:
        assert lib.oxg_dist_create(None, 0, 1, C.byref(h)) == OXG_E_CUDA
