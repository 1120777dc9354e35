"""CPU, world_size 2, gloo: the multi-GPU layout rules (BLOCK partition, ghost lists, halo exchange, sink reduction,
all-reduced inner products) reproduce the single-rank oracle Action.  This covers the N > 1 host logic without GPUs;
the CUDA/NCCL implementation of the same rules is checked on real GPUs by tests/multirank_check.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, bounds, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import planes_from_oracle
    from oracle import oracle as O
    from partition_model import block_layout, fetch_x, ghost_plan

    # every rank builds the same (replicated) state directory, like the product does
    st = O.StateSet(fixture=name, bounds=bounds)
    st.expand()
    A = O.FspMatrix(constrained=True)
    A.generate_fixture(st, name)
    d = planes_from_oracle(A)
    N, K, P = d["n"], d["K"], d["col"].shape[0]
    starts = block_layout(N, world)
    lo, hi = int(starts[rank]), int(starts[rank + 1])
    n_loc = hi - lo
    t = 3.0
    coef = O.fixture_tcoef(name, t, st.R)[1]
    order = list(d["tv"]) + list(d["ti"])
    c = np.array([coef[r] if p < len(d["tv"]) else 1.0 for p, r in enumerate(order)])

    rng = np.random.default_rng(5)
    xg = rng.random(N + K)
    x_loc = xg[lo:hi].copy()

    col_local, ghost_gids, recv_counts = ghost_plan(d["col"][:, lo:hi], starts, rank)
    # tell each peer which entries we need (counts, then ids), receive what they need from us
    send_counts = torch.zeros(world, dtype=torch.int64)
    dist.all_to_all_single(send_counts, torch.from_numpy(recv_counts))
    want = torch.from_numpy(ghost_gids.astype(np.int64))
    need_from_me = torch.zeros(int(send_counts.sum()), dtype=torch.int64)
    dist.all_to_all_single(need_from_me, want, output_split_sizes=send_counts.tolist(), input_split_sizes=recv_counts.tolist())
    send_idx = need_from_me.numpy() - lo
    assert ((send_idx >= 0) & (send_idx < n_loc)).all()
    # halo exchange of x
    ghost = torch.zeros(len(ghost_gids), dtype=torch.float64)
    dist.all_to_all_single(ghost, torch.from_numpy(x_loc[send_idx]), output_split_sizes=recv_counts.tolist(),
                           input_split_sizes=send_counts.tolist())
    ghost = ghost.numpy()
    assert np.array_equal(ghost, xg[ghost_gids])

    # peer-memory form of the same exchange (fsphalo_*): every sender learns where its segment starts in the receiver's
    # ghost window (remote_off) and "stores" its values there, two parities, three epochs with changing x
    from partition_model import cta_issue_order, window_offsets
    recv_off = window_offsets(recv_counts)
    remote_off = torch.zeros(world, dtype=torch.int64)
    dist.all_to_all_single(remote_off, torch.from_numpy(recv_off[:world].copy()))
    send_off = window_offsets(send_counts.numpy())
    window = np.full((2, len(ghost_gids)), np.nan)
    for epoch in (1, 2, 3):
        par = epoch & 1
        xe = x_loc * epoch
        # what lands in peer p's window: (offset, values); delivered here by an all-to-all of offsets and values
        offs_in = torch.zeros(world, dtype=torch.int64)
        dist.all_to_all_single(offs_in, remote_off.clone())
        vals_in = torch.zeros(len(ghost_gids), dtype=torch.float64)
        dist.all_to_all_single(vals_in, torch.from_numpy(xe[send_idx]), output_split_sizes=recv_counts.tolist(),
                               input_split_sizes=send_counts.tolist())
        pos = 0
        for p in range(world):
            cnt = int(recv_counts[p])
            window[par, int(offs_in[p]): int(offs_in[p]) + cnt] = vals_in[pos: pos + cnt].numpy()
            pos += cnt
        assert np.array_equal(window[par], xg[ghost_gids] * epoch)
    assert send_off[-1] == len(send_idx)
    order, n_int = cta_issue_order(n_loc, col_local, threads=8)
    assert sorted(order.tolist()) == list(range((n_loc + 7) // 8))
    ghost_rows = np.nonzero((col_local <= -2).any(axis=0))[0]
    assert not set((ghost_rows // 8).tolist()) & set(order[:n_int].tolist())
    assert set((ghost_rows // 8).tolist()) == set(order[n_int:].tolist())

    # local rows of the fused operator
    y_loc = np.zeros(n_loc)
    for p in range(P):
        y_loc += c[p] * (d["off"][p, lo:hi] * fetch_x(x_loc, ghost, col_local[p]) - d["diag"][p, lo:hi] * x_loc)
    # sink partial sums over local states, reduced to the last rank
    sink = np.zeros(K)
    for p in range(P):
        for k in range(K):
            b, e = d["sink_ptr"][p * K + k], d["sink_ptr"][p * K + k + 1]
            idx = d["sink_idx"][b:e]
            m = (idx >= lo) & (idx < hi)
            sink[k] += c[p] * np.dot(d["sink_val"][b:e][m], xg[idx[m]])
    sink_t = torch.from_numpy(sink)
    dist.all_reduce(sink_t)
    # all-reduced inner product (what VecDot does multi-rank)
    dot = torch.tensor([float(y_loc @ x_loc)], dtype=torch.float64)
    dist.all_reduce(dot)

    ierr, y_ref = A.action(t, xg)
    err = np.abs(y_loc - y_ref[lo:hi]).max() / np.abs(y_ref).max()
    ok = err <= 1e-12
    if rank == world - 1:
        ok &= np.abs(sink_t.numpy() - y_ref[N:]).max() <= 1e-12 * np.abs(y_ref).max()
    ok &= abs(float(dot) - float(y_ref[:N] @ xg[:N])) <= 1e-12 * np.abs(y_ref[:N]).dot(np.abs(xg[:N]))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,bounds", [("birth_death_3d_tv", [9, 7, 5]), ("toggle_custom", [12, 9, 40]), ("hog1p", [3, 4, 4, 3, 3])])
def test_two_rank_partitioned_action_matches_oracle(name, bounds):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    O.lib()  # build once before forking
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, bounds, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert all(ret.get(r, False) for r in range(world))


def test_block_layout_rule():
    sys.path.insert(0, ROOT)
    from partition_model import block_layout
    assert block_layout(10, 4).tolist() == [0, 3, 6, 8, 10]
    assert block_layout(13, 2).tolist() == [0, 7, 13]
    assert block_layout(3, 8).tolist() == [0, 1, 2, 3, 3, 3, 3, 3, 3]


@pytest.mark.parametrize("world", [2, 3, 8])
def test_epoch_protocol_two_parities_suffice_under_any_interleaving(world):
    """The peer-memory halo keeps only TWO ghost buffers per rank.  Randomised interleavings of the ranks' push/consume
    steps (including ranks racing ahead as far as the protocol lets them) never read overwritten data."""
    sys.path.insert(0, ROOT)
    from partition_model import EpochProtocol
    rng = np.random.default_rng(world)
    for trial in range(20):
        prot = EpochProtocol(world)
        bias = rng.random(world) + 0.05   # some ranks are much "faster" than others
        bias /= bias.sum()
        for _ in range(4000):
            prot.step(int(rng.choice(world, p=bias)))
        assert min(prot.consumed) >= 5
        assert max(prot.consumed) - min(prot.consumed) <= 1   # nobody can run more than one epoch ahead
