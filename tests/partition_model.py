"""Host-side statement of the multi-GPU layout rules (pure numpy; no compute path uses this module):

* block_layout      == StateSetBase::update_layout (pacmensl_b200/host/StateSetBase.cpp): contiguous equal-count BLOCK
                       split of the global ordering, ranks < N % P own one extra state (Zoltan LB_METHOD=BLOCK with unit
                       weights, reference src/Partitioner/StatePartitionerBase.cpp:81-83,144); sinks on the last rank.
* ghost_plan        == fspmat_build_ghosts + FspMatrixBase::SetupGhosts_: columns outside the own block become ghost
                       slots (sorted unique global ids, grouped by owner), col is re-encoded (>= 0 local, -1 none,
                       <= -2 ghost slot -(col+2)), and every rank learns which of its entries each peer needs.

* window_offsets / cta_issue_order / EpochProtocol == the peer-memory path (csrc/fspcomm.cu fsphalo_*, csrc/fspmat.cu
                       fsp_action_p2p_kernel): where a sender's segment starts in a receiver's ghost window, the CTA
                       issue order (ghost-free CTAs first), and the flag/parity protocol that makes two ghost buffers
                       enough.

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


def window_offsets(recv_counts):
    """Start of every source rank's segment in this rank's ghost window (ghost slots are laid out per source in rank
    order): the prefix sum fsphalo_create computes before it tells every peer its own start (remote_off)."""
    off = np.zeros(len(recv_counts) + 1, dtype=np.int64)
    off[1:] = np.cumsum(recv_counts)
    return off


def cta_issue_order(n_local, col_local, threads=256):
    """CTA issue order of the single-kernel peer-memory action: CTAs whose rows reference no ghost slot first (in
    ascending order), then the others.  Returns (order, n_interior)."""
    n_ctas = (n_local + threads - 1) // threads
    has_ghost = np.zeros(n_ctas, dtype=bool)
    rows = np.nonzero((np.asarray(col_local) <= -2).any(axis=0))[0]
    has_ghost[rows // threads] = True
    interior = np.nonzero(~has_ghost)[0]
    boundary = np.nonzero(has_ghost)[0]
    return np.concatenate([interior, boundary]), len(interior)


class EpochProtocol:
    """Model of the flag/parity protocol of the peer-memory halo (fspcomm.cu): every rank, every epoch e = 1, 2, ...
    (1) push: stores its boundary values into buffer[e & 1] of every peer, then sets flag[e & 1][me] = e on the peer;
    (2) consume: waits until flag[e & 1][p] >= e for every peer p, then reads buffer[e & 1].
    A rank runs push(e) only after its own consume(e - 1).  step(rank) advances one rank by one micro-step if it can;
    any interleaving of step() calls is a legal execution.  check() is asserted inside: a consumer must read exactly
    the values of its epoch (i.e. no peer may have overwritten the buffer with epoch e + 2 data before it was read)."""

    def __init__(self, world):
        self.world = world
        self.buf = [[[None] * world for _ in range(2)] for _ in range(world)]   # buf[dst][parity][src] = epoch of data
        self.flag = [[[0] * world for _ in range(2)] for _ in range(world)]
        self.epoch = [1] * world
        self.phase = [0] * world   # 0: about to push, 1: about to consume
        self.consumed = [0] * world

    def step(self, r):
        e = self.epoch[r]
        par = e & 1
        if self.phase[r] == 0:
            for p in range(self.world):
                self.buf[p][par][r] = e
                self.flag[p][par][r] = e
            self.phase[r] = 1
            return True
        if all(self.flag[r][par][p] >= e for p in range(self.world)):
            for p in range(self.world):
                assert self.buf[r][par][p] == e, "rank %d epoch %d read data of epoch %s from %d" % (r, e, self.buf[r][par][p], p)
            self.consumed[r] = e
            self.epoch[r] = e + 1
            self.phase[r] = 0
            return True
        return False
