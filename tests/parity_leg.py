"""Partitioned-Action parity against the CPU oracle, by state key (shared by bench.py's parity leg, which runs at every
N before the timed region, and by tests/multirank_check.py).

For every case the state set and the operator are built through the host C ABI on all ranks (BLOCK partition,
src/StateSet/StateSetBase.cpp:286-301 semantics: rank r owns a contiguous range of global indices, the K sink rows
live on the last rank), Action(t, x, y) is applied to a seeded global vector, and rank 0 compares the gathered result
with the oracle's reference-shaped Action on the same set: states are matched by their integer key (State2Index of the
oracle), never by raw position.  Bound: 1e-12 relative to max|y| (BASELINE north_star).
"""
import numpy as np

TOL = 1e-12

# (fixture, bounds or None = the fixture's own, times)
DEFAULT_CASES = [
    ("birth_death_3d", [21, 17, 13], (0.0,)),
    ("birth_death_3d_tv", [21, 17, 13], (0.0, 4.0)),
    ("transcr_reg_6d", [10, 6, 1, 2, 1, 1], (0.0, 40.0)),
    ("hog1p", [3, 6, 6, 5, 5], (25.0,)),
    ("pure_birth", [5], (0.0,)),  # 6 states: with >= 2 ranks the first rank references no ghost column, with > 6 some own no state
]


def _gather(local, dist, dev, dtype):
    """all-gather variable-length 1-D blocks; returns the concatenation in rank order (numpy)."""
    import torch
    if dist is None:
        return local.cpu().numpy()
    world = dist.get_world_size()
    cnt = torch.zeros(world, dtype=torch.int64, device=dev)
    cnt[dist.get_rank()] = local.numel()
    dist.all_reduce(cnt)
    sizes = [int(v) for v in cnt.tolist()]
    pad = max(max(sizes), 1)
    buf = torch.zeros(pad, dtype=dtype, device=dev)
    buf[: local.numel()] = local
    outs = [torch.zeros(pad, dtype=dtype, device=dev) for _ in range(world)]
    dist.all_gather(outs, buf)
    return np.concatenate([o[: sizes[r]].cpu().numpy() for r, o in enumerate(outs)])


def run_case(api, dist, dev, name, bounds, times, reps=3, seed=99, sharded=None):
    """Returns (states_rel_err, sinks_rel_err, n_states) on rank 0 (zeros elsewhere)."""
    import torch
    rank = dist.get_rank() if dist is not None else 0
    world = dist.get_world_size() if dist is not None else 1
    if name.startswith("birth_death_3d"):
        from pacmensl_b200.lattice import Lattice
        lat = Lattice(bounds, tv=name.endswith("_tv"), sharded=sharded)
        st, mat = lat.set, lat.mat
    else:
        st, mat = api.fixture_set_and_matrix(name, bounds=np.asarray(bounds, dtype=np.int32), sharded=sharded)
    n_local, N, start = st.sizes()
    n_rows = mat.n_rows
    K = 0
    if world == 1:
        K = n_rows - n_local
    else:
        k_t = torch.tensor([float(n_rows - n_local)], dtype=torch.float64, device=dev)
        dist.all_reduce(k_t, op=dist.ReduceOp.MAX)
        K = int(k_t.item())
    S = st.S
    states_g = _gather(torch.from_numpy(st.states().reshape(-1)).to(dev), dist, dev, torch.int32).reshape(-1, S)
    xg = np.random.default_rng(seed).random(N + K)
    owns_sinks = n_rows > n_local
    xl = np.concatenate([xg[start: start + n_local], xg[N:] if owns_sinks else np.zeros(0)])
    assert len(xl) == n_rows
    e_states = e_sinks = 0.0
    so = A = perm = None
    if rank == 0:
        # a failure here must not raise on rank 0 alone (the other ranks would hang in the next collective): it is
        # reported as an infinite error instead
        from oracle import oracle as O
        so = O.StateSet(fixture=name, bounds=list(bounds))
        A = O.FspMatrix(constrained=True)
        if so.expand() != 0 or A.generate_fixture(so, name) != 0 or so.n != N:
            e_states = float("inf")
        else:
            perm = so.state2index(states_g.astype(np.int32))
            if not ((perm >= 0).all() and len(np.unique(perm)) == N):  # the two state sets differ
                e_states, perm = float("inf"), None
    if dist is not None:
        # rank 0 may have spent seconds in the oracle: the other ranks wait HERE (no time limit), not inside the Action's
        # device-side flag waits, which give up after FSP_SPIN_TIMEOUT_MS (20 s) and poison the result
        dist.barrier()
    for t in times:
        xd = torch.from_numpy(xl).to(dev)
        yd = torch.empty_like(xd)
        for rep in range(reps):  # repeated calls walk through both parities of the double-buffered ghost windows
            yd.fill_(float("nan"))
            if mat.action(t, xd if rep == reps - 1 else xd * (rep + 2.0), yd) != 0:
                yd.fill_(float("nan"))  # reported as a non-finite result below
        torch.cuda.synchronize()
        # rank r contributes its n_local states (+ K sinks on the last rank): global order = states, then sinks
        yg = _gather(yd, dist, dev, torch.float64)
        if rank == 0 and perm is not None:
            x_or = np.zeros(N + K)
            x_or[perm] = xg[:N]
            x_or[N:] = xg[N:]
            ierr, y_or = A.action(t, x_or)
            scale = float(np.abs(y_or).max()) or 1.0
            e_states = max(e_states, float(np.abs(yg[:N] - y_or[perm]).max() / scale))
            if K:
                e_sinks = max(e_sinks, float(np.abs(yg[N:] - y_or[N:]).max() / scale))
            if ierr != 0 or not np.isfinite(yg).all():
                e_states = float("inf")
    del mat, st
    return e_states, e_sinks, N


def run(api, dist, dev, cases=None, verbose=False, sharded=None):
    """Runs all cases; returns {"max_rel_err", "sinks_rel_err", "cases": {...}, "ok"} (meaningful on rank 0)."""
    out = {"max_rel_err": 0.0, "sinks_rel_err": 0.0, "tol": TOL, "cases": {}}
    rank = dist.get_rank() if dist is not None else 0
    for name, bounds, times in (cases or DEFAULT_CASES):
        es, ek, n = run_case(api, dist, dev, name, bounds, times, sharded=sharded)
        out["cases"][name] = {"states": n, "rel_err": es, "sinks_rel_err": ek}
        out["max_rel_err"] = max(out["max_rel_err"], es)
        out["sinks_rel_err"] = max(out["sinks_rel_err"], ek)
        if verbose and rank == 0:
            print("action parity %-18s N=%-7d rel_err states %.2e sinks %.2e%s" % (name, n, es, ek, "  (sharded set)" if sharded else ("  (replicated set)" if sharded is False else "")))
    out["ok"] = bool(out["max_rel_err"] <= TOL and out["sinks_rel_err"] <= TOL)
    return out
