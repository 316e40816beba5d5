"""Host-side sharding of the PRM build over ranks: one process per GPU, the store
and the world replicated, query rows partitioned, results exchanged with all-gathers.

Works on any torch.distributed backend: NCCL on the GPU box (bench.py), gloo in the CPU tests
(tests/test_sharding_gloo.py).  No data-path collective other than the all-gathers of the
fixed-width per-node rows exists."""
from __future__ import annotations

import numpy as np


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
