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

    # ---- 3. the personalised all-to-all of the set-up paths (ghost id lists, routed ExpandVec) vs torch.distributed ----
    import ctypes as C
    from pacmensl_b200 import _capi
    L = _capi.lib()
    comm = api.world_comm()
    rng = np.random.default_rng(100 + rank)
    a2a_ok = True
    for trial, (dtype, esz) in enumerate(((torch.int32, 4), (torch.float64, 8), (torch.int32, 4), (torch.float64, 8))):
        hi = 3 if trial < 2 else 200000   # tiny segments (some empty), then segments that make the window grow
        sc = torch.from_numpy(rng.integers(0, hi, size=world)).to(torch.int64)
        rc = torch.zeros(world, dtype=torch.int64)
        sc_d, rc_d = sc.to(dev), rc.to(dev)
        dist.all_to_all_single(rc_d, sc_d)
        rc = rc_d.cpu()
        send = (torch.rand(int(sc.sum()), device=dev, dtype=torch.float64) * 1e6).to(dtype)
        want = torch.empty(int(rc.sum()), device=dev, dtype=dtype)
        dist.all_to_all_single(want, send, output_split_sizes=rc.tolist(), input_split_sizes=sc.tolist())
        got = torch.full((max(int(rc.sum()), 1),), -7, device=dev, dtype=dtype)
        scc = (C.c_long * world)(*sc.tolist())
        rcc = (C.c_long * world)(*rc.tolist())
        ierr = L.fspcomm_alltoallv(comm, C.c_void_p(send.data_ptr()), scc, C.c_void_p(got.data_ptr()), rcc, esz, None)
        a2a_ok &= ierr == 0 and bool((got[: int(rc.sum())] == want).all())
    a2a = torch.tensor([1.0 if a2a_ok else 0.0], device=dev)
    dist.all_reduce(a2a, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("fspcomm_alltoallv (peer windows) == torch all_to_all_single on %d ranks: %s" % (world, a2a.item() == 1.0))
        ok &= a2a.item() == 1.0

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
