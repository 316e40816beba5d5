"""Sharding of the PRM build over ranks: one process per GPU, the store and the
world replicated, query rows partitioned, results exchanged with all-gathers.

``Dist`` is the product path: the library's own multi-GPU entry points (include/oxg.h,
``oxg_prm_build_knn_dist`` / ``oxg_prm_build_radius_dist``), NCCL inside liboxg, the host only
carries the 128-byte NCCL id between the ranks (``Dist.from_torch`` uses torch.distributed for
that rendezvous and nothing else).

The numpy helpers below restate the partitioning rules (row ranges, edge ownership) for the CPU
tests (tests/test_sharding_gloo.py, world_size-2 gloo) and ``all_gather_rows`` is the
torch.distributed form of the exchange they exercise."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np


def radius_row_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Rows of `rank` in radius mode: node i scans the i earlier nodes (prm.rs:131), so the cuts
    sit at n*sqrt(g/G) (equal areas of the triangle).  Mirrors oxg_dist_row_range(radius_mode=1)."""
    def cut(g):
        if g <= 0:
            return 0
        if g >= world_size:
            return n
        return min(n, int(math.floor(float(n) * math.sqrt(float(g) / float(world_size)))))
    return cut(rank), cut(rank + 1)


class Dist:
    """``oxg_dist``: this rank's handle on the library's NCCL communicator."""

    def __init__(self, rank: int, world_size: int, unique_id: bytes | None = None):
        from ._lib import check, lib

        self.rank, self.world_size = rank, world_size
        self._h = C.c_void_p()
        buf = None
        if world_size > 1:
            if unique_id is None or len(unique_id) != 128:
                raise ValueError("a 128-byte NCCL unique id is required for world_size > 1")
            buf = C.create_string_buffer(bytes(unique_id), 128)
        check(lib.oxg_dist_create(buf, rank, world_size, C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        from ._lib import check, lib

        buf = C.create_string_buffer(128)
        check(lib.oxg_dist_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch(cls, dist, rank: int, world_size: int) -> "Dist":
        """Rendezvous through an initialised torch.distributed group: rank 0 draws the NCCL id
        and broadcasts the 128 bytes; everything after that is the library's."""
        if world_size == 1:
            return cls(0, 1)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(rank, world_size, box[0])

    @property
    def handle(self):
        return self._h

    def row_range(self, n: int, radius_mode: bool = False) -> tuple[int, int]:
        from ._lib import check, lib

        rb, re = C.c_uint64(0), C.c_uint64(0)
        check(lib.oxg_dist_row_range(n, self.rank, self.world_size, 1 if radius_mode else 0,
                                     C.byref(rb), C.byref(re)))
        return rb.value, re.value

    def append(self, store, rows_local) -> int:
        """Collective append (``oxg_store_append_dist``): every rank passes its own slice; all
        stores end up holding all rows in rank order.  Returns the first new index."""
        from ._lib import check, lib

        rows = np.ascontiguousarray(rows_local, dtype=np.float64)
        first = C.c_uint64(0)
        check(lib.oxg_store_append_dist(self._h, store.handle, C.c_void_p(rows.ctypes.data),
                                        rows.shape[0], C.byref(first)))
        return first.value

    def prm_build_knn(self, world, store, k: int, resolution: float, want_host: bool = True):
        """The sharded k-NN build -> the complete (idx[N,k], dist[N,k], valid[N,k], stats)."""
        from ._lib import PrmStats, check, lib

        n = len(store)
        idx = np.empty((n, k), dtype=np.uint32) if want_host else None
        dst = np.empty((n, k), dtype=np.float64) if want_host else None
        valid = np.empty((n, k), dtype=np.uint8) if want_host else None
        st = PrmStats()
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)  # noqa: E731
        check(lib.oxg_prm_build_knn_dist(self._h, world.handle, store.handle, k, float(resolution),
                                         p(idx), p(dst), p(valid), C.byref(st)))
        return idx, dst, valid, {name: getattr(st, name) for name, _ in PrmStats._fields_}

    def prm_build_radius(self, world, store, connection_radius: float, resolution: float):
        """The sharded radius build -> CSR (offsets[N+1], idx) of every node's lower-index edges."""
        from . import _lib
        from ._lib import PrmStats, check, lib

        n = len(store)
        offsets = np.zeros(n + 1, dtype=np.uint64)
        cap = max(1024, 32 * n)
        st = PrmStats()
        while True:
            idx = np.empty(cap, dtype=np.uint32)
            code = lib.oxg_prm_build_radius_dist(self._h, world.handle, store.handle,
                                                 float(connection_radius), float(resolution),
                                                 C.c_void_p(offsets.ctypes.data),
                                                 C.c_void_p(idx.ctypes.data), cap, C.byref(st))
            if code == _lib.OXG_E_CAPACITY:
                cap = int(offsets[n])
                continue
            check(code)
            break
        return offsets, idx[: int(offsets[n])].copy()

    def close(self):
        from ._lib import lib

        if getattr(self, "_h", None) is not None and self._h:
            lib.oxg_dist_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def row_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous query rows of `rank`: k-NN-all work is uniform per row."""
    return (n * rank) // world_size, (n * (rank + 1)) // world_size


def rows_max(n: int, world_size: int) -> int:
    return max(row_range(n, r, world_size)[1] - row_range(n, r, world_size)[0]
               for r in range(world_size))


def edge_owner(i, j):
    """Which endpoint's row owns the motion check of the candidate pair {i, j} when both rows
    list each other (mutual neighbours): the parity of i^j picks the lower or the higher index,
    so owners spread evenly over rows — and therefore over the contiguous row shards.  Mirrors
    prm_mark_edges_kernel (csrc/oxg_prm.cu)."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    return np.where(((i ^ j) & 1) == 1, lo, hi)


def owned_slots(idx_all: np.ndarray, row_begin: int, row_end: int) -> np.ndarray:
    """Boolean [rows, k]: the candidate slots of rows [row_begin, row_end) whose motion check
    this shard executes.  A slot (i, j) is owned unless the pair is mutual and the owner is j."""
    n, k = idx_all.shape
    rows = np.arange(row_begin, row_end, dtype=np.int64)
    sub = idx_all[row_begin:row_end].astype(np.int64)
    valid = sub != 0xFFFFFFFF
    jj = np.where(valid, sub, 0)
    mutual = (idx_all[jj].astype(np.int64) == rows[:, None, None]).any(axis=2) & valid
    owner = edge_owner(np.broadcast_to(rows[:, None], sub.shape), jj)
    return valid & (~mutual | (owner == rows[:, None]))


def all_gather_rows(dist, all_t, rank: int, world_size: int, pad_t=None):
    """In-place all-gather of the row shards of `all_t` ([n, k] tensor, every rank holds its own
    rows [row_range)).  Uneven shards go through `pad_t` ([world, rows_max, k])."""
    if world_size == 1:
        return
    n = all_t.shape[0]
    rb, re = row_range(n, rank, world_size)
    if n % world_size == 0:
        dist.all_gather_into_tensor(all_t, all_t[rb:re].clone())
        return
    pad_t[rank, : re - rb].copy_(all_t[rb:re])
    dist.all_gather_into_tensor(pad_t.view(-1), pad_t[rank].reshape(-1).clone())
    for r in range(world_size):
        a, b = row_range(n, r, world_size)
        all_t[a:b].copy_(pad_t[r, : b - a])
