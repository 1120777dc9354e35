"""What can the host <-> device links of this box carry when N GPUs copy at once?  (context for bench.py's e2e number)
Run under torchrun: every rank copies `mb` MB up and `mb` MB down concurrently (two streams, pinned memory), `reps` times;
prints per-rank GB/s per direction, the aggregate, the GPU <-> CPU topology and each rank's CPU affinity.
  --bind : first bind the process to the CPUs local to its GPU (sysfs local_cpulist), so pinned memory is NUMA-local."""
import argparse
import os
import subprocess
import time

import torch
import torch.distributed as dist


def local_cpus(index):
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip().lower()
        bus = bus[4:] if len(bus) > 12 else bus  # 00000000:1B:00.0 -> 0000:1b:00.0
        path = "/sys/bus/pci/devices/%s/" % bus
        node = open(path + "numa_node").read().strip()
        cpus = open(path + "local_cpulist").read().strip()
        return node, cpus
    except Exception as e:  # noqa: BLE001
        return "?", "? (%r)" % (e,)


def parse_cpulist(s):
    out = set()
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        elif part.strip().isdigit():
            out.add(int(part))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=200)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--bind", action="store_true")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    node, cpus = local_cpus(lr)
    if a.bind:
        want = parse_cpulist(cpus) & os.sched_getaffinity(0)
        if want:
            os.sched_setaffinity(0, want)
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n = a.mb * 1000 * 1000 // 8
    hx = torch.empty(n, dtype=torch.float64).pin_memory()
    hy = torch.empty(n, dtype=torch.float64).pin_memory()
    hx.fill_(1.0)
    dx = torch.empty(n, dtype=torch.float64, device="cuda")
    dy = torch.ones(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for it in range(2):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(a.reps):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        dx.copy_(hx, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        hy.copy_(dy, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        res[mode] = a.mb * 1e6 * a.reps / dt / 1e9
    line = "rank %d gpu %d numa %s cpus %s affinity %d cores | per direction GB/s: h2d alone %.1f, d2h alone %.1f, both at once %.1f" % (
        rank, lr, node, cpus, len(os.sched_getaffinity(0)), res["h2d"], res["d2h"], res["both"])
    if world > 1:
        outs = [None] * world
        dist.all_gather_object(outs, line)
        tot = torch.tensor([res["both"]], device="cuda")
        dist.all_reduce(tot)
        if rank == 0:
            print("\n".join(outs))
            print("aggregate with %d GPUs copying both ways at once: %.1f GB/s per direction (bind=%s)" % (world, tot.item(), a.bind))
            print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:3000])
        dist.destroy_process_group()
    else:
        print(line)


if __name__ == "__main__":
    main()
