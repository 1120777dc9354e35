"""CPU: the host arithmetic of the sharded state set's re-balance (fspset_rebalance_plan in pacmensl_b200/csrc/fspset.cu,
what the reference does with Zoltan_LB_Partition + Zoltan_Migrate, src/Partitioner/StatePartitionerBase.cpp:136-239, for
LB_METHOD=BLOCK): the rank-concatenated listing of the states is re-cut into contiguous equal-count blocks and every
rank pulls its new block out of the old ones.  Checked: every state arrives exactly once, the order of the listing is
kept, the layout is the BLOCK split -- single process over many random configurations, and with two gloo ranks that
actually move their blocks.  The device side of the same function (peer-to-peer copies between the GPUs' windows) is
checked on real GPUs by tests/multirank_sharded_check.py."""
import ctypes as C
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def plan(lib, counts, rank):
    n = len(counts)
    cnt = (C.c_long * n)(*[int(c) for c in counts])
    starts = (C.c_long * (n + 1))()
    src, off, dst, ln = [(C.c_long * n)() for _ in range(4)]
    n_seg, same = C.c_int(), C.c_int()
    rc = lib.fspset_rebalance_plan(n, cnt, rank, starts, src, off, dst, ln, C.byref(n_seg), C.byref(same))
    assert rc == 0
    segs = [(src[k], off[k], dst[k], ln[k]) for k in range(n_seg.value)]
    return list(starts), segs, bool(same.value)


def block_layout(n, world):
    base, rem = divmod(n, world)
    st = [0]
    for r in range(world):
        st.append(st[-1] + base + (1 if r < rem else 0))
    return st


def test_plan_moves_every_state_once_and_keeps_the_order():
    from pacmensl_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(7)
    for trial in range(300):
        world = int(rng.integers(1, 17))
        kind = trial % 4
        if kind == 0:      # everything on the last rank (a BFS frontier after a BLOCK split)
            counts = [0] * (world - 1) + [int(rng.integers(0, 5000))]
        elif kind == 1:    # already balanced
            n = int(rng.integers(0, 5000))
            st = block_layout(n, world)
            counts = [st[r + 1] - st[r] for r in range(world)]
        elif kind == 2:    # fewer states than ranks
            counts = [int(v) for v in rng.integers(0, 2, size=world)]
        else:
            counts = [int(v) for v in rng.integers(0, 3000, size=world)]
        old = [[(r, i) for i in range(counts[r])] for r in range(world)]
        listing = [x for blk in old for x in blk]
        n = len(listing)
        new = []
        for r in range(world):
            starts, segs, same = plan(lib, counts, r)
            assert starts == block_layout(n, world)
            assert same == (counts == [starts[q + 1] - starts[q] for q in range(world)])
            assert len(segs) <= world
            blk = [None] * (starts[r + 1] - starts[r])
            last_dst = -1
            for (q, off, dst, ln) in segs:
                assert ln > 0 and dst > last_dst and off + ln <= counts[q]
                last_dst = dst
                blk[dst: dst + ln] = old[q][off: off + ln]
            assert all(b is not None for b in blk)
            new.append(blk)
        assert [x for blk in new for x in blk] == listing


def test_plan_rejects_bad_arguments():
    from pacmensl_b200 import _capi
    lib = _capi.lib()
    cnt = (C.c_long * 2)(3, -1)
    out = [(C.c_long * 3)() for _ in range(5)]
    a, b = C.c_int(), C.c_int()
    assert lib.fspset_rebalance_plan(2, cnt, 0, *out, C.byref(a), C.byref(b)) != 0
    cnt = (C.c_long * 2)(3, 1)
    assert lib.fspset_rebalance_plan(2, cnt, 2, *out, C.byref(a), C.byref(b)) != 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pacmensl_b200 import _capi
    lib = _capi.lib()
    ok = True
    rng = np.random.default_rng(11)  # same stream on both ranks: both know every rank's count, as after the gather
    for trial in range(20):
        counts = [int(v) for v in rng.integers(0, 400, size=world)]
        if trial == 0:
            counts = [1] + [0] * (world - 1)  # the initial state of an FSP solve
        S = 3
        mine = np.stack([np.full(counts[rank], rank), np.arange(counts[rank]), np.arange(counts[rank]) ** 2 % 7], axis=1).astype(np.int32).reshape(-1, S)
        windows = [None] * world  # "peer windows": every rank can read every block
        dist.all_gather_object(windows, mine)
        starts, segs, same = plan(lib, counts, rank)
        new = np.zeros((starts[rank + 1] - starts[rank], S), dtype=np.int32)
        for (q, off, dst, ln) in segs:
            new[dst: dst + ln] = windows[q][off: off + ln]
        gathered = [None] * world
        dist.all_gather_object(gathered, new)
        listing_old = np.concatenate([w.reshape(-1, S) for w in windows])
        listing_new = np.concatenate([g.reshape(-1, S) for g in gathered])
        ok &= bool((listing_old == listing_new).all())
        ok &= [len(g) for g in gathered] == [starts[r + 1] - starts[r] for r in range(world)]
    ret[rank] = ok
    dist.destroy_process_group()


def test_rebalance_between_two_gloo_ranks():
    import torch.multiprocessing as mp
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


# ---- the whole sharded construction, as a host model on two gloo ranks, against the oracle's state set --------------
# Rules restated from pacmensl_b200/csrc/fspset.cu (sharded mode): a state's directory shard is hash(state) mod N; every
# rank generates the children of the frontier states IT owns; a key claimed by several candidates goes to the smallest
# (is_candidate, rank, position) -- stored states beat candidates, then the lowest rank, then the lowest position;
# winners are appended on the discovering rank with status 1; a frontier state ends with status 0 (all children inside)
# or -1; after the BFS the blocks are re-cut by the product's own fspset_rebalance_plan.
TOGGLE_SM = [[1, 1, -1, 0, 0, 0], [0, 0, 0, 1, 1, -1]]  # the toggle-switch fixture (pacmensl_b200/fixtures/fsp_models.h)


def _shard(key, world):
    return (key[0] * 1000003 + key[1] * 10007) % world   # any function of the key that all ranks agree on


def _sharded_bfs_worker(rank, world, port, name, bounds, ret):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from pacmensl_b200 import _capi
    lib = _capi.lib()
    info = O.fixture_info(name)
    SM = np.asarray(TOGGLE_SM, dtype=np.int64)           # S x R
    x0 = tuple(int(v) for v in info["x0"])
    S, R = SM.shape
    b = np.asarray(bounds, dtype=np.int64)

    def valid(x):
        return all(v >= 0 for v in x) and all(x[s] <= b[s] for s in range(len(b)))

    states, status = [], []                               # this rank's block
    shard = {}                                            # my directory shard: key -> (rank, position)
    if rank == 0:
        states.append(x0); status.append(1)
    if _shard(x0, world) == rank:
        shard[x0] = (0, 0)
    while True:
        frontier = [i for i, s in enumerate(status) if s == 1]
        tot = [None] * world
        dist.all_gather_object(tot, len(frontier))
        if sum(tot) == 0:
            break
        cand, fstat = [], {i: 0 for i in frontier}
        for j in range(R):                                # reaction-major inside the batch
            for i in frontier:
                c = tuple(int(states[i][s] + SM[s, j]) for s in range(S))
                if valid(c):
                    cand.append(c)
                else:
                    fstat[i] = -1
        # claims: every candidate goes to the shard of its key, tagged (rank, position in my candidate list)
        out = [[] for _ in range(world)]
        for pos, c in enumerate(cand):
            out[_shard(c, world)].append((c, rank, pos))
        allout = [None] * world
        dist.all_gather_object(allout, out)
        claims = [m for r in range(world) for m in allout[r][rank]]
        winners = {}
        for (c, r, pos) in claims:                        # atomicMin over (rank, position); stored states always stay
            if c in shard:
                continue
            if c not in winners or (r, pos) < winners[c]:
                winners[c] = (r, pos)
        allwin = [None] * world
        dist.all_gather_object(allwin, winners)
        mine = sorted(pos for w in allwin for (c, (r, pos)) in w.items() if r == rank)
        new_pos = {}
        for pos in mine:                                  # winners are compacted in candidate order
            new_pos[pos] = len(states)
            states.append(cand[pos]); status.append(1)
        # the shard owners store (rank, final position)
        final = {cand[pos]: (rank, p) for pos, p in new_pos.items()}
        allfinal = [None] * world
        dist.all_gather_object(allfinal, final)
        for f in allfinal:
            for c, v in f.items():
                if _shard(c, world) == rank:
                    shard[c] = v
        for i in frontier:
            status[i] = fstat[i]
    # re-balance with the product's plan
    counts = [None] * world
    dist.all_gather_object(counts, len(states))
    windows = [None] * world
    dist.all_gather_object(windows, states)
    starts, segs, same = plan(lib, counts, rank)
    new = [None] * (starts[rank + 1] - starts[rank])
    for (q, off, dst, ln) in segs:
        new[dst: dst + ln] = windows[q][off: off + ln]
    blocks = [None] * world
    dist.all_gather_object(blocks, new)
    union = [x for blk in blocks for x in blk]
    ok = len(union) == len(set(union))                    # no state twice
    if rank == 0:
        so = O.StateSet(fixture=name, bounds=list(bounds))
        so.expand()
        ok &= set(union) == set(map(tuple, so.states().tolist()))
        ok &= [len(blk) for blk in blocks] == [starts[r + 1] - starts[r] for r in range(world)]
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_construction_model_on_two_gloo_ranks_equals_the_oracle_set():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_bfs_worker, args=(world, _free_port(), "toggle", [14, 11], ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
