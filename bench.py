#!/usr/bin/env python
"""bench.py -- FSP Action() bandwidth on the synthetic 3-D birth-death lattice (BASELINE.json config 4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--lattice 465] [--tv]

A "step" is one Action(t, x, y) over the whole lattice vector.  Metric = algorithmic GB moved per
second, bytes/row = 16 + 12 R + 8 (R_tv + [R_ti>0]) (+ sinks), SURVEY.md section 8(d).
One JSON line is printed by rank 0.  For N > 1 launch with torchrun (one rank per GPU): the lattice is
split into N contiguous row blocks (strong scaling, halo exchange over NCCL).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lattice", type=int, default=465, help="states per species axis (L+1); N = lattice^3")
    ap.add_argument("--lattice-dims", default=None, help="a,b,c states per axis (overrides --lattice; diagnostics)")
    ap.add_argument("--tv", action="store_true", help="time-varying births (R_tv = 3)")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (0 = default lean kernel; 1,2,3,4 = alternatives, see fspmat.cu)")
    ap.add_argument("--cpu-lattice", type=int, default=0, help="lattice edge of the CPU-baseline arm (0 = the same lattice as the GPU arm)")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-parity", action="store_true", help="skip the partitioned-Action parity leg (oracle check before the timed region)")
    ap.add_argument("--trace-steps", action="store_true", help="print per-rank per-step event times of the timed region to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-expand", action="store_true", help="skip the Expand() closure check of the lattice set")
    ap.add_argument("--no-solve", action="store_true", help="skip the solve-to-t_f side measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra single-GPU rooflines (TV lattice, BFS-ordered example sets)")
    ap.add_argument("--solve-lattice", type=int, default=0, help="lattice edge of the GPU solve-to-t_f measurement (0 = --lattice)")
    ap.add_argument("--cpu-solve-lattice", type=int, default=128, help="lattice edge of the bounded CPU solve sample")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def lattice_bytes_per_row(tv):
    R, ntv = 6, (3 if tv else 0)
    return 16 + 12 * R + 8 * (ntv + (1 if R - ntv > 0 else 0))


def workload_name(dims, tv):
    n = dims[0] * dims[1] * dims[2]
    return ("synthetic 3-D birth-death lattice %s = %d states, S=3 R=6 K=3 sinks, %s, FspMatrixConstrained::Action(t,x,y)"
            % ("x".join(str(d) for d in dims), n, "R_tv=3" if tv else "time-invariant"))


def lattice_dims(args):
    return [args.lattice] * 3 if not args.lattice_dims else [int(v) for v in args.lattice_dims.split(",")]


def cpu_baseline(args, steps=None, warmup=2):
    """Reference-shaped multi-pass Action (oracle port: one CSR SpMV into a work vector + AXPY per matrix,
    src/Matrix/FspMatrixBase.cpp:36-62) on the SAME lattice as the GPU arm, all host cores."""
    steps = steps or args.cpu_steps
    import numpy as np
    from oracle import oracle as O
    cores = O.use_all_cores()  # torchrun exports OMP_NUM_THREADS=1; the baseline must use the box's cores
    dims = lattice_dims(args)
    if args.cpu_lattice:
        dims = [args.cpu_lattice] * 3
    n = dims[0] * dims[1] * dims[2]
    # ~115 B/state of host memory (CSR + vectors); fall back to a bounded sample if the box cannot hold it
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 1 << 40
    if n * 130 > avail:
        edge = int((avail / 130.0) ** (1.0 / 3.0)) // 5 * 5
        dims = [edge] * 3
        n = edge ** 3
    A = O.FspMatrix(constrained=True)
    t0 = time.perf_counter()
    assert A.generate_lattice(dims, tv=args.tv) == 0
    t_gen = time.perf_counter() - t0
    rng = np.random.default_rng(12345)
    x = rng.random(A.nrows)
    x[n:] = 0.0
    x /= x.sum()
    y = np.empty_like(x)
    for _ in range(max(warmup, 1)):
        A.action_into(0.3, x, y)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        A.action_into(0.3, x, y)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    nbytes = n * lattice_bytes_per_row(args.tv) + 12 * (dims[0] * dims[1] + dims[1] * dims[2] + dims[0] * dims[2]) + 8 * 3
    return {"value": nbytes / med / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
            "sample": "%s lattice = %d states (%s), median of %d reference-shaped (CSR SpMV + AXPY per matrix) OpenMP Actions; "
                      "oracle/fsp_oracle.c (the PETSc reference cannot be built in this image)"
                      % ("x".join(str(d) for d in dims), n, "the GPU arm's workload" if dims == lattice_dims(args) else "bounded sample", steps),
            "ms_per_step": med * 1e3, "dims": dims, "same_config": dims == lattice_dims(args), "build_seconds": round(t_gen, 2),
            "sum_y": float(y.sum())}


EXTRA_WORKLOADS = [
    # (key, fixture, bounds): time-varying lattice + the BFS-ordered example sets at the sizes the example solves end with
    ("lattice_tv", "birth_death_3d_tv", None),
    ("hog1p_bfs_23.9M", "hog1p", [3, 36, 73, 36, 58]),
    ("hog1p_bfs_858k", "hog1p", [3, 17, 36, 13, 22]),
    ("repressilator_bfs_3.95M", "repressilator", [151, 146, 176]),
    ("transcr_reg_6d_bfs_1.64M", "transcr_reg_6d", [68, 119, 2, 2, 2, 32]),
]


def extra_workloads(args, steps=20, warmup=3, only=None):
    """Action() on the other operators of BASELINE's configs, one GPU: the time-varying lattice (R_tv = 3, 120 B/row)
    and the BFS-ordered example state sets, where the x[col] gathers are irregular.  Same metric, same roofline
    arithmetic as the headline (algorithmic bytes = fspmat_action_bytes / CUDA-event time)."""
    import numpy as np
    import torch
    from pacmensl_b200 import api
    from pacmensl_b200.lattice import Lattice
    peak, _ = measured_peak()
    out = {}
    for key, name, bounds in EXTRA_WORKLOADS:
        if only and key not in only:
            continue
        try:
            t0 = time.perf_counter()
            if name.startswith("birth_death_3d"):
                lat = Lattice([d - 1 for d in lattice_dims(args)], tv=name.endswith("_tv"), expand=False)
                st, mat = lat.set, lat.mat
            else:
                st, mat = api.fixture_set_and_matrix(name, bounds=np.asarray(bounds, dtype=np.int32))
            torch.cuda.synchronize()
            t_build = time.perf_counter() - t0
            n_rows, flops, nbytes = mat.info()
            x = torch.rand(n_rows, dtype=torch.float64, device="cuda")
            x /= x.sum()
            y = torch.empty_like(x)
            for _ in range(warmup):
                assert mat.action(25.0, x, y) == 0
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(steps):
                mat.action(25.0, x, y)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / steps
            P = st.R
            out[key] = {"states": st.n_global, "reactions": P, "bytes_per_row": round(nbytes / max(st.n_global, 1), 1),
                        "ms_per_action": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
                        "frac": nbytes / (ms * 1e-3) / 1e9 / peak, "build_seconds": round(t_build, 2), "sum_y": float(y.sum())}
            del mat, st, x, y
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            out[key] = {"unavailable": repr(e)[:200]}
    return out


def solve_to_tf(args, rank, world, local_rank, port_base):
    """Second half of BASELINE's metric: wall time of KrylovFsp / CvodeFsp solves to t_f = 1 on the fixed lattice
    (p0 = delta(0), analytic check: product of three Poisson pmfs), on N GPUs.  Every rank of this job runs
    examples/lattice_solve (the host C++ classes on the CUDA library) as ITS rank of an N-rank solve, after the Action
    workload has been released; rank 0 reports.  CPU beside it (N = 1 only): the oracle's restatement of KrylovFsp
    (oracle/krylov_oracle.py) on a bounded sample."""
    exe = os.path.join(ROOT, "build", "examples", "lattice_solve")
    edge = args.solve_lattice or args.lattice
    out = {"t_final": 1.0, "lattice_edge": edge, "ranks": world}

    def gpu(solver, e, port):
        env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(local_rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        r = subprocess.run([exe, "--edge", str(e), "--solver", solver, "--repeat", "2"], capture_output=True, text=True,
                           timeout=420, env=env)
        if rank != 0:
            return None
        j = json.loads(r.stdout.strip().splitlines()[-1])
        return {"states": j["states"], "wall_s": j["wall_s"], "action_calls": j["action_calls"], "status": j["status"],
                "l1_err_vs_poisson": j["l1_err_vs_poisson"], "sum_p": j["sum_p"]}

    for k, (solver, label) in enumerate((("krylov", "KrylovFsp (defaults: IOP q=2, m in [25,60], atol 1e-14)"),
                                         ("cvode", "CvodeFsp (BDF, rtol 1e-6, atol 1e-14)"))):
        try:
            res = gpu(solver, edge, port_base + 40 * k)
            if rank == 0:
                res["solver"] = label
                out[solver] = res
        except Exception as e:  # noqa: BLE001
            out[solver] = {"unavailable": repr(e)[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            import numpy as np
            L = args.cpu_solve_lattice
            out["krylov_at_cpu_sample_size"] = gpu("krylov", L, port_base + 80)
            from oracle import oracle as O
            from oracle.krylov_oracle import KrylovOracle
            cores = O.use_all_cores()
            A = O.FspMatrix(constrained=True)
            assert A.generate_lattice([L] * 3) == 0
            p0 = np.zeros(A.nrows)
            p0[0] = 1.0  # lexicographic order: state (0,0,0) has index 0
            kry = KrylovOracle(A)
            t0 = time.perf_counter()
            p = kry.solve(p0, 1.0)
            out["cpu_krylov"] = {"states": L ** 3, "wall_s": time.perf_counter() - t0, "action_calls": kry.num_rhs, "cores": cores,
                                 "kind": "port", "sum_p": float(p.sum())}
        except Exception as e:  # noqa: BLE001
            out["cpu_krylov"] = {"unavailable": repr(e)[:200]}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # K steps / W warm-ups as asked, each step one reference-shaped Action on the GPU arm's lattice (~0.1 s per step)
    steps, warmup = max(1, min(args.steps, 200)), max(args.warmup, 3)
    cb = cpu_baseline(args, steps=steps, warmup=warmup)
    line = {
        "metric": "FSP Action() GB/s", "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": workload_name(cb["dims"], args.tv), "states": cb["dims"][0] * cb["dims"][1] * cb["dims"][2],
                   "bytes_per_row": lattice_bytes_per_row(args.tv), "same_config": cb["same_config"],
                   "build_seconds": cb["build_seconds"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from pacmensl_b200 import _capi, api
    from pacmensl_b200.lattice import Lattice
    L = _capi.lib()
    api.init(local_rank, dist)

    # ---- parity leg: the partitioned Action at THIS N against the CPU oracle, by state key (before anything is timed)
    parity = None
    if not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import parity_leg
        parity = parity_leg.run(api, dist, dev)
        flag = torch.tensor([1.0 if (rank != 0 or parity["ok"]) else 0.0], device=dev)
        if dist is not None:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() != 1.0:
            if rank == 0:
                print(json.dumps({"metric": "FSP Action() GB/s", "n_gpus": world, "parity": parity,
                                  "error": "partitioned Action differs from the oracle"}))
            raise SystemExit(3)

    # ---- build: state set (device hash directory) + operator through the host C++ classes -------------------
    t_build0 = time.perf_counter()
    dims = lattice_dims(args)
    lat = Lattice([d - 1 for d in dims], tv=args.tv, expand=not args.no_expand)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build0
    lat.mat.set_variant(args.variant)
    N = lat.n_global
    n_rows = lat.n_rows
    bytes_local = lat.bytes

    g = torch.Generator(device="cuda").manual_seed(12345 + rank)
    x = torch.rand(n_rows, generator=g, dtype=torch.float64, device="cuda")
    if lat.n_local < n_rows:
        x[lat.n_local:] = 0.0
    s = x.sum()
    if dist is not None:
        dist.all_reduce(s)
    x /= s
    y = torch.empty_like(x)
    t_eval = 0.3

    def step():
        lat.action(t_eval, x, y)  # FspMatrixConstrained::Action: ONE launch (N > 1: push + sinks + rows + finish)

    # the clock sampler (a fork + exec of nvidia-smi, tens of ms) starts BEFORE the warm-up and the barrier, so that
    # no rank enters the timed region later than the others because of it
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    tick = torch.zeros(1, device=dev)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    # ---- timed region: exactly K steps, CUDA events on the launching stream ----------------------------------
    launches0 = L.fsp_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)] if args.trace_steps else None
    if dist is not None:
        # device-side barrier on the launching stream: every GPU leaves this all-reduce at the same moment, and its
        # timed region starts right there (host-side barrier exits are tens of microseconds apart)
        dist.all_reduce(tick)
    ev0.record()
    for k in range(args.steps):
        step()
        if step_ev:
            step_ev[k].record()
    ev1.record()
    torch.cuda.synchronize()
    launches = L.fsp_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    if step_ev:
        ts = [ev0.elapsed_time(e) for e in step_ev]
        sys.stderr.write("[trace rank %d] total %.4f ms; per-step end times (ms): %s\n"
                         % (rank, ms, " ".join("%.3f" % v for v in ts)))
    if dist is not None:
        dist.barrier()
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
        tot = torch.tensor([bytes_local, float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tot)
        bytes_total, launches_total = float(tot[0].item()), int(tot[1].item())
    else:
        bytes_total, launches_total = bytes_local, launches
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / args.steps
    value = bytes_total / (ms_per_step * 1e-3) / 1e9
    # A step IS one launch of the Action kernel on every GPU (N = 1: fsp_action_lean; N > 1: fsp_action_halo_kernel),
    # so the per-launch duration of the dominant kernel is the event-timed step.
    kms = ms_per_step

    # ---- N > 1: the halo exchange alone (push CTAs + finishing CTA, no rows): what crosses NVLink per Action ----------
    halo = None
    if dist is not None and api.p2p_enabled():
        for _ in range(5):
            sent = lat.mat.halo_only(x, y)
        torch.cuda.synchronize()
        dist.all_reduce(tick)
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(50):
            lat.mat.halo_only(x, y)
        h1.record()
        torch.cuda.synchronize()
        th = torch.tensor([h0.elapsed_time(h1) / 50.0, float(sent)], dtype=torch.float64, device="cuda")
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
        us = float(th[0].item()) * 1e3
        halo = {"us_per_exchange": us, "bytes_pushed_per_rank_max": int(th[1].item()),
                "GBps_per_direction": float(th[1].item()) / (us * 1e-6) / 1e9 if us > 0 else None,
                "what": "pack + store into the peers' windows over NVLink + epoch flag + wait for every peer's flag, one launch, "
                        "back to back (latency-bound: 2 planes of 465^2 doubles per interior rank); nvidia-smi nvlink counters "
                        "report N/A on this pool"}

    # ---- e2e: the same Action through the C ABI with HOST buffers (H2D x, D2H y inside the timed region) -----
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(n_rows, dtype=torch.float64).pin_memory()
        yh = torch.empty(n_rows, dtype=torch.float64).pin_memory()
        xh.copy_(x)
        k2 = args.steps

        def e2e_step():
            ierr = lat.mat.action_host(t_eval, xh, yh)  # pfsp_mat_action_host: H2D(x) + Action + D2H(y)
            if ierr:
                raise SystemExit("action_host failed")

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        e2e_ok = bool(torch.equal(yh.to(dev), y))  # bit-identical to the device-vector Action of the same x
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k2):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k2
        if dist is not None:
            tt = torch.tensor([dt, 0.0 if e2e_ok else 1.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt, e2e_ok = float(tt[0].item()), tt[1].item() == 0.0
        # what the host <-> device links of THIS box carry when every rank moves the same bytes both ways at once and
        # nothing is computed (two streams, the same pinned buffers): the floor of any e2e step on this machine
        dx = torch.empty_like(x)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
        link = []
        for rep in range(4):
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                dx.copy_(xh, non_blocking=True)
            with torch.cuda.stream(s_dn):
                yh.copy_(y, non_blocking=True)
            torch.cuda.synchronize()
            link.append(time.perf_counter() - t0)
        dl = min(link[1:])
        if dist is not None:
            tt = torch.tensor([dl], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dl = float(tt.item())
        del dx
        e2e = {"value": bytes_total / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(n_rows * 8),
               "d2h_bytes_per_step": int(n_rows * 8), "ms_per_step": dt * 1e3, "steps": k2,
               "bit_identical_to_device_action": e2e_ok,
               "copies_only_ms_per_step": dl * 1e3, "frac_of_link_floor": dl / dt,
               "link_note": "copies_only = the same H2D + D2H bytes on all ranks at once with no compute (max over ranks): "
                            "the PCIe / host-memory floor of this box; e2e cannot scale beyond the box's aggregate host links",
               "call": "pfsp_mat_action_host (include/pacmensl_b200_host.h) with pinned host x, y"}
        del xh, yh

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        ach = bytes_local / (kms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "action_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("lattice") == args.lattice and bool(tj.get("tv")) == bool(args.tv) and world == 1:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                pass
        p2p = api.p2p_enabled()
        kernel = ("fsp_action_lean<6,0>" if world == 1 else
                  ("fsp_action_halo_kernel<6> (push + sink + row + finishing CTAs in one launch)" if p2p
                   else "fsp_action_lean<6,2> + boundary rows + NCCL halo (rank 0 share)"))
        line = {
            "metric": "FSP Action() GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic (x ~ U(0,1) normalised, torch Philox seed 12345)",
            "config": {"workload": workload_name(dims, args.tv),
                       "states": N, "bytes_per_row": lattice_bytes_per_row(args.tv),
                       "l2": "inputs (%.2f GB per Action) exceed the 126 MB L2; no flush needed" % (bytes_total / 1e9),
                       "partition": "1 block" if world == 1 else (
                           "%d contiguous row blocks (BLOCK); " % world + (
                               "peer-memory path: ONE kernel per Action and GPU -- leading CTAs pack the boundary entries of x and store "
                               "them into the peers' ghost windows (CUDA IPC over NVLink) + epoch flag, row CTAs wait in device code "
                               "only where a warp meets a ghost column, sink slots summed by the owner; no NCCL call per Action"
                               if p2p else "NCCL halo exchange + K-double sink all-reduce per Action")),
                       "state_set": "one block" if world == 1 else (
                           "sharded construction: each rank built its block, directory striped over the GPUs' HBM (peer windows)"
                           if lat.set.is_sharded() else "replicated directory (every rank built the whole set)"),
                       "kernel_variant": args.variant, "build_seconds": round(t_build, 2),
                       "timing": "device barrier (all-reduce) on the launching stream, then CUDA events around exactly K steps; max over ranks"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": kernel + " (pacmensl_b200/csrc/fspmat.cu)",
                         "kernel_ms": kms, "algorithmic_bytes_per_launch": bytes_local},
            "e2e": e2e, "gpu_launches": int(launches_total), "clocks": clocks, "parity": parity, "halo_exchange": halo,
        }
        if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only; `--impl reference` is the CPU arm at every N
            cb = cpu_baseline(args)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "same_config")}
    if dist is not None:
        dist.barrier()
    del lat, x, y
    torch.cuda.empty_cache()
    if world == 1 and not args.no_extra:
        line["roofline_other_operators"] = extra_workloads(args, steps=min(args.steps, 50))
    api.finalize()
    port_base = int(os.environ.get("MASTER_PORT", "29500")) + 101
    if dist is not None:
        dist.destroy_process_group()
    torch.cuda.empty_cache()
    if not args.no_solve:
        # side measurement after the device memory of the Action workload has been released; never fatal
        try:
            sol = solve_to_tf(args, rank, world, local_rank, port_base)
        except Exception as e:  # noqa: BLE001
            sol = {"unavailable": repr(e)[:200]}
        if rank == 0:
            line["solve_to_tf"] = sol
    if rank == 0:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
