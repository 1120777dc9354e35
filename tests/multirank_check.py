"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multirank_check.py

1. Action on the block-partitioned birth-death lattice (halo exchange + sink all-reduce) == CPU oracle (1e-12).
2. Adaptive FSP solves (Krylov and BDF) on N ranks == analytic Poisson pmf (KAT-F4/F5 bounds), which exercises the
   multi-rank ExpandVec redistribution and the all-reduced inner products.
Rank 0 prints "MULTIRANK OK" on success.
"""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def gather_blocks(local, sizes, device):
    """all-gather variable-size float64 blocks (padded)"""
    world = dist.get_world_size()
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.float64, device=device)
    buf[: local.numel()] = local
    outs = [torch.zeros(pad, dtype=torch.float64, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([o[: sizes[r]].cpu().numpy() for r, o in enumerate(outs)])


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pacmensl_b200 import api
    from pacmensl_b200.lattice import Lattice
    api.init(local_rank, dist)
    dev = torch.device("cuda", local_rank)
    ok = True
    if rank == 0:
        print("peer-memory fast path: %s" % ("on" if api.p2p_enabled() else "off (NCCL)"))

    # ---- 1. partitioned Action vs oracle ----
    from helpers import rel_err
    from oracle import oracle as O
    for tv, name in ((False, "birth_death_3d"), (True, "birth_death_3d_tv")):
        upper = [21, 17, 13]
        lat = Lattice(upper, tv=tv)
        N = lat.n_global
        K = 3
        rng = np.random.default_rng(99)
        xg = rng.random(N + K)
        sizes_rows = [0] * world
        t_sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        t_sizes[rank] = lat.n_rows
        dist.all_reduce(t_sizes)
        sizes_rows = [int(v) for v in t_sizes.tolist()]
        xl = np.concatenate([xg[lat.start: lat.start + lat.n_local], xg[N:] if rank == world - 1 else np.zeros(0)])
        assert len(xl) == lat.n_rows
        for t in (0.0, 4.0):
            xd = torch.from_numpy(xl).to(dev)
            yd = torch.empty_like(xd)
            for rep in range(5):  # repeated calls walk through both parities of the double-buffered ghost windows
                yd.fill_(float("nan"))
                lat.action(t, xd if rep == 4 else xd * (rep + 2.0), yd)
            torch.cuda.synchronize()
            yg = gather_blocks(yd, sizes_rows, dev)   # states of rank 0.., then the K sinks of the last rank
            if rank == 0:
                st = O.StateSet(fixture=name, bounds=upper)
                st.expand()
                A = O.FspMatrix(constrained=True)
                A.generate_fixture(st, name)
                Ls = [u + 1 for u in upper]
                idx = np.arange(N)
                X = np.stack([idx % Ls[0], (idx // Ls[0]) % Ls[1], idx // (Ls[0] * Ls[1])], axis=1).astype(np.int32)
                perm = st.state2index(X)
                x_or = np.zeros(N + K)
                x_or[perm] = xg[:N]
                x_or[N:] = xg[N:]
                ierr, y_or = A.action(t, x_or)
                e1 = rel_err(yg[:N], y_or[perm], scale=np.abs(y_or).max())
                e2 = rel_err(yg[N:], y_or[N:], scale=np.abs(y_or).max())
                print("action parity tv=%d t=%g: rel_err states %.2e sinks %.2e" % (tv, t, e1, e2))
                ok &= e1 <= 1e-12 and e2 <= 1e-12
        del lat

    # ---- 2. adaptive FSP solve on N ranks vs Poisson ----
    for ode, label in ((api.KRYLOV, "krylov"), (api.CVODE, "cvode")):
        s, m = api.fixture_solver("pure_birth", ode)
        states, p = s.solve(10.0, 1e-6)
        lam_t = 20.0
        pdf = np.array([math.exp(-lam_t) * lam_t ** int(n) / math.gamma(int(n) + 1) for n in states[:, 0]])
        err = torch.tensor([np.abs(p - pdf).sum()], dtype=torch.float64, device=dev)
        dist.all_reduce(err)
        stt = s.stats()
        if rank == 0:
            print("poisson %s on %d ranks: %d states, %d expansions, L1 error %.3e" % (label, world, stt["n_states"], stt["expansions"], float(err)))
            ok &= float(err) <= 1e-6
        s.clear()

    # toggle with custom constraints (host lhs callbacks) on N ranks: mass conservation
    s, m = api.fixture_solver("toggle_custom", api.KRYLOV)
    s.set_initial_bounds([10, 10, 30])
    states, p = s.solve(100.0, 1e-6)
    tot = torch.tensor([p.sum()], dtype=torch.float64, device=dev)
    dist.all_reduce(tot)
    if rank == 0:
        print("toggle_custom krylov on %d ranks: %d states, 1 - sum(p) = %.3e" % (world, s.stats()["n_states"], 1.0 - float(tot)))
        ok &= (1.0 - float(tot)) <= 1e-6 + 1e-9 and float(tot) <= 1.0 + 1e-8
    s.clear()

    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    if rank == 0:
        print("MULTIRANK OK" if ok else "MULTIRANK FAILED")
    dist.barrier()
    api.finalize()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
