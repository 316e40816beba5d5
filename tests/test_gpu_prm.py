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
