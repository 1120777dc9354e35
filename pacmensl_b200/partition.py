"""Host-side statement of the multi-GPU layout rules (pure numpy; no compute path uses this module):

* block_layout      == StateSetBase::update_layout (pacmensl_b200/host/StateSetBase.cpp): contiguous equal-count BLOCK
                       split of the global ordering, ranks < N % P own one extra state (Zoltan LB_METHOD=BLOCK with unit
                       weights, reference src/Partitioner/StatePartitionerBase.cpp:81-83,144); sinks on the last rank.
* ghost_plan        == fspmat_build_ghosts + FspMatrixBase::SetupGhosts_: columns outside the own block become ghost
                       slots (sorted unique global ids, grouped by owner), col is re-encoded (>= 0 local, -1 none,
                       <= -2 ghost slot -(col+2)), and every rank learns which of its entries each peer needs.

tests/test_partition_gloo.py runs these rules with world_size 2 on the gloo backend against the CPU oracle; the CUDA/NCCL
implementation is checked against the same oracle by tests/multirank_check.py on real GPUs.
"""
import numpy as np


def block_layout(n_global, world):
    base, rem = divmod(int(n_global), int(world))
    starts = np.zeros(world + 1, dtype=np.int64)
    for r in range(world):
        starts[r + 1] = starts[r] + base + (1 if r < rem else 0)
    return starts


def ghost_plan(col_global, starts, rank):
    """col_global: int array [P, n_local] of GLOBAL column indices (-1 = none) for this rank's rows.
    Returns (col_local, ghost_gids, recv_counts) with ghost_gids sorted ascending (=> grouped by owner)."""
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    col = np.asarray(col_global).copy()
    outside = (col >= 0) & ((col < lo) | (col >= hi))
    ghost_gids = np.unique(col[outside])
    local = (col >= lo) & (col < hi)
    col_local = col.copy()
    col_local[local] = col[local] - lo
    col_local[outside] = -(np.searchsorted(ghost_gids, col[outside]) + 2)
    owners = np.searchsorted(starts, ghost_gids, side="right") - 1
    recv_counts = np.bincount(owners, minlength=len(starts) - 1).astype(np.int64)
    return col_local, ghost_gids, recv_counts


def fetch_x(x_local, ghost, c):
    """value of x referenced by the encoded column c (vectorised)"""
    c = np.asarray(c)
    out = np.zeros(c.shape, dtype=np.float64)
    loc = c >= 0
    out[loc] = x_local[c[loc]]
    gh = c <= -2
    out[gh] = ghost[-(c[gh] + 2)]
    return out
