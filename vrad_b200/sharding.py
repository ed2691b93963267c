"""Host-side sharding rules for the multi-GPU path (one process per GPU).

Rays / luxels are independent: each rank takes a contiguous range.  Patch rows of the transfer matrix
are partitioned in contiguous blocks that tile [0, N) in rank order.  `row_partition` is the default rule (equal
blocks of ceil(N / world) rows rounded up to a multiple of 4, what vrad_build_transfers uses without a communicator and what callers of
vrad_transfers_upload typically pass); with a communicator vrad_build_transfers balances the blocks by estimated
transfers instead.  Radiance buffers are indexed by global patch number, so any such tiling works.  The only
data-path exchange is the per-bounce hand-over of the new radiance rows.
"""
from __future__ import annotations

import numpy as np

from .environment import row_partition  # noqa: F401  (re-exported)


def rows_per_rank(n_rows: int, world: int) -> int:
    return ((n_rows + world - 1) // world + 3) & ~3          # blocks start on multiples of 4 (environment.row_partition)


def range_partition(n: int, world: int):
    """Contiguous ranges for independent work items (rays, luxels)."""
    return [(r * n // world, (r + 1) * n // world) for r in range(world)]


def slice_csr(rowptr, col, w, row0: int, row1: int):
    """Rows [row0,row1) of a CSR matrix, rowptr rebased to 0 (what vrad_transfers_upload takes)."""
    rowptr = np.asarray(rowptr, np.int64)
    s, e = int(rowptr[row0]), int(rowptr[row1])
    return rowptr[row0:row1 + 1] - s, np.asarray(col[s:e], np.int32), np.asarray(w[s:e], np.float32)


def padded_gather_buffer(n_rows: int, world: int, width: int = 3, dtype=np.float32):
    """The all-gather buffer: world * rows_per_rank rows (tail rows of the last rank are padding)."""
    return np.zeros((rows_per_rank(n_rows, world) * world, width), dtype)


def collect_rows(parent, child1, child2, area):
    """CollectLight for interior patches, flattened the way the library does it (vrad_patches_set_hierarchy): for every
    interior patch the leaves of its subtree with weight = product of the area fractions area_child / (area_c1 + area_c2)
    along the path (fp32, top down).  Returns (interior ids, ptr, leaf, weight)."""
    parent = np.asarray(parent, np.int32); child1 = np.asarray(child1, np.int32); child2 = np.asarray(child2, np.int32)
    area = np.asarray(area, np.float32)
    ids, ptr, leaf, wt = [], [0], [], []
    for p in np.nonzero(child1 != -1)[0]:
        ids.append(int(p))
        work = [(int(p), np.float32(1.0))]
        while work:
            q, wq = work.pop()
            if child1[q] == -1:
                leaf.append(q); wt.append(wq)
                continue
            a1, a2 = area[child1[q]], area[child2[q]]
            work.append((int(child2[q]), np.float32(wq * np.float32(a2 / np.float32(a1 + a2)))))
            work.append((int(child1[q]), np.float32(wq * np.float32(a1 / np.float32(a1 + a2)))))
        ptr.append(len(leaf))
    return np.asarray(ids, np.int32), np.asarray(ptr, np.int64), np.asarray(leaf, np.int32), np.asarray(wt, np.float32)


def apply_collect(values, ids, ptr, leaf, wt):
    """values[interior] = sum of wt * values[leaf] over the subtree (in place; values: [N, 3])."""
    for k, p in enumerate(ids):
        s, e = int(ptr[k]), int(ptr[k + 1])
        values[p] = (values[leaf[s:e]] * wt[s:e, None]).sum(axis=0, dtype=np.float32)
    return values
