"""ncu target (1 GPU): a few Action() launches on every operator of bench.py's EXTRA_WORKLOADS plus the headline lattice,
so that `ncu --set full -k regex:fsp_action_lean` captures the kernels that ship, on the inputs that are hard (BFS order).
    python tools/profile_actions.py [--lattice 465] [--only key,key]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lattice", type=int, default=465)
    ap.add_argument("--only", default="")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch
    from pacmensl_b200 import api
    from pacmensl_b200.lattice import Lattice
    torch.cuda.set_device(0)
    api.init(0, None)
    args = argparse.Namespace(lattice=a.lattice, lattice_dims=None, tv=False)
    only = [k for k in a.only.split(",") if k] or None
    out = {}
    if not only or "lattice_ti" in only:
        lat = Lattice([a.lattice - 1] * 3, tv=False, expand=False)
        x = torch.rand(lat.n_rows, dtype=torch.float64, device="cuda")
        y = torch.empty_like(x)
        for _ in range(a.warmup + a.steps):
            lat.action(0.3, x, y)
        torch.cuda.synchronize()
        out["lattice_ti"] = {"states": lat.n_global, "bytes": lat.bytes}
        del lat, x, y
        torch.cuda.empty_cache()
    out.update(bench.extra_workloads(args, steps=a.steps, warmup=a.warmup, only=only))
    api.finalize()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
