"""GPU parity: batched PRM construction (radius = reference-faithful, k-NN = config 4)."""
import numpy as np
import pytest

from conftest import uniform_states
from oxmpl_b200 import worlds

pytestmark = pytest.mark.gpu


def _panda_world(oxg, orc, n_caps=64):
    caps = worlds.capsule_world(4, worlds.PANDA_LO, worlds.PANDA_HI, n_caps)
    gw = oxg.World(7, capsules=caps)
    ow, keep = orc.world(7, capsules=caps)
    ext = oxg.real_vector_max_extent(list(zip(worlds.PANDA_LO, worlds.PANDA_HI)))
    return gw, ow, keep, oxg.motion_resolution(ext)


def test_prm_radius_matches_incremental_reference(oxg, orc):
    """The batched build must equal PRM::construct_roadmap (prm.rs:117-146) run sample by sample."""
    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    gw, ow, keep, res = _panda_world(oxg, orc)
    samples = uniform_states(4, 2500, worlds.PANDA_LO, worlds.PANDA_HI)
    radius = 1.6
    nodes, ooff, oidx, pairs, mcs = orc.prm_radius(ow, sp_o, radius, res, samples)
    ok = gw.is_valid(sp, samples).astype(bool)                 # prm.rs:123
    np.testing.assert_array_equal(samples[ok], nodes)
    with oxg.StateStore(sp) as st:
        st.append(samples[ok])
        off, idx, stats = oxg.prm_build_radius(gw, st, radius, res)
        np.testing.assert_array_equal(off, ooff)
        np.testing.assert_array_equal(idx, oidx)
        assert stats["n_candidates"] == mcs and stats["n_edges_valid"] == len(oidx)
        assert 0 < len(oidx) < mcs                              # some edges blocked, some free
        # sharded rows reproduce the same CSR pieces (multi-GPU unit of work)
        n = len(nodes)
        cut = n // 3
        off_a, idx_a, _ = oxg.prm_build_radius(gw, st, radius, res, 0, cut)
        off_b, idx_b, _ = oxg.prm_build_radius(gw, st, radius, res, cut, n)
        np.testing.assert_array_equal(np.concatenate([idx_a, idx_b]), oidx)
        np.testing.assert_array_equal(off_b + off_a[-1], ooff[cut:])
    gw.close()


@pytest.mark.parametrize("n,k", [(3000, 10), (9, 10), (513, 4)])
def test_prm_knn_matches_oracle(oxg, orc, n, k):
    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    gw, ow, keep, res = _panda_world(oxg, orc)
    rows = uniform_states(44, n, worlds.PANDA_LO, worlds.PANDA_HI)
    oi, od, ov = orc.prm_knn(ow, sp_o, rows, k, res)
    with oxg.StateStore(sp) as st:
        st.append(rows)
        idx, dist, valid, stats = oxg.prm_build_knn(gw, st, k, res)
        np.testing.assert_array_equal(idx, oi)
        np.testing.assert_array_equal(dist.view(np.uint64), od.view(np.uint64))
        np.testing.assert_array_equal(valid, ov)
        assert stats["n_candidates"] == int((oi != 0xFFFFFFFF).sum())
        assert stats["n_edges_checked"] <= stats["n_candidates"]
        # row-range call (shard): same rows, mutual-pair dedup limited to the range
        a, b = n // 4, n // 2
        idx2, dist2, valid2, _ = oxg.prm_build_knn(gw, st, k, res, a, b)
        np.testing.assert_array_equal(idx2, oi[a:b])
        np.testing.assert_array_equal(valid2, ov[a:b])
    gw.close()


@pytest.mark.parametrize("n,k,shards", [(3001, 10, 2), (2000, 6, 4)])
def test_device_pointer_pipeline_sharded_on_one_gpu(oxg, orc, n, k, shards):
    """The multi-GPU PRM path (oxg_knn_rows_dev -> all-gather -> oxg_prm_edges_dev -> all-gather
    -> oxg_prm_resolve_dev, as bench.py drives it with NCCL) emulated on one GPU: every shard's
    rows are computed by separate calls into shared [n, k] device buffers (what the all-gathers
    produce), the result must equal the unsharded oracle roadmap, and every candidate edge must
    be motion-checked exactly once across the shards."""
    import ctypes as C

    import torch

    from oxmpl_b200 import sharding
    from oxmpl_b200._lib import check, lib

    sp, sp_o = oxg.Space.real_vector(7), orc.space(7)
    gw, ow, keep, res = _panda_world(oxg, orc)
    rows = uniform_states(45, n, worlds.PANDA_LO, worlds.PANDA_HI)
    oi, od, ov = orc.prm_knn(ow, sp_o, rows, k, res)
    dev = torch.device("cuda", 0)
    with oxg.StateStore(sp) as st:
        st.append(rows)
        ext = torch.cuda.ExternalStream(st.stream(), device=dev)
        with torch.cuda.stream(ext):
            idx_all = torch.empty((n, k), dtype=torch.int32, device=dev)
            dist_all = torch.empty((n, k), dtype=torch.float64, device=dev)
            valid_all = torch.full((n, k), 7, dtype=torch.uint8, device=dev)
            counters = torch.zeros(2, dtype=torch.int64, device=dev)
        ext.synchronize()
        for r in range(shards):
            rb, re = sharding.row_range(n, r, shards)
            check(lib.oxg_knn_rows_dev(st.handle, rb, re, k, C.c_void_p(idx_all[rb:re].data_ptr()),
                                       C.c_void_p(dist_all[rb:re].data_ptr())))
        checked = 0
        for r in range(shards):
            rb, re = sharding.row_range(n, r, shards)
            check(lib.oxg_prm_edges_dev(gw.handle, st.handle, C.c_void_p(idx_all.data_ptr()), k,
                                        float(res), rb, re, C.c_void_p(valid_all.data_ptr()),
                                        C.c_void_p(counters.data_ptr())))
            st.sync()
            checked += int(counters.cpu()[0])
            own = sharding.owned_slots(oi, rb, re)                 # host mirror of the ownership rule
            assert int(counters.cpu()[0]) == int(own.sum())
        check(lib.oxg_prm_resolve_dev(st.handle, C.c_void_p(idx_all.data_ptr()), k,
                                      C.c_void_p(valid_all.data_ptr())))
        st.sync()
        np.testing.assert_array_equal(idx_all.cpu().numpy().view(np.uint32), oi)
        np.testing.assert_array_equal(dist_all.cpu().numpy().view(np.uint64), od.view(np.uint64))
        np.testing.assert_array_equal(valid_all.cpu().numpy(), ov)
        pairs = {(min(i, int(j)), max(i, int(j))) for i in range(n) for j in oi[i] if j != 0xFFFFFFFF}
        assert checked == len(pairs)
    gw.close()
