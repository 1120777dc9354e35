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

    # ---- 1. partitioned Action vs oracle, by state key (lattice TI/TV, transcr_reg_6d, hog1p, and pure_birth where
    #         rank 0 references no ghost column at all -- the case that broke round 1's single-kernel path) ----
    import parity_leg
    par = parity_leg.run(api, dist, dev, verbose=True)
    ok &= par["ok"] if rank == 0 else True

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
