"""A peer that never arrives must surface as an error, not as a wrong result (run under torchrun with 2 ranks and
FSP_SPIN_TIMEOUT_MS=300): rank 1 skips one Action; rank 0's CTAs that need the halo time out, poison their rows of y
with NaN and raise the communicator's error flag, which pfsp_check() reports as a non-zero return code."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    from pacmensl_b200 import api
    from pacmensl_b200.lattice import Lattice
    api.init(local_rank, dist)
    ok = api.p2p_enabled()
    # replicated set: the collective (sharded) construction has device-side barriers of its own, which the 300 ms limit
    # of this test would cut short whenever the ranks' first kernels load a few hundred ms apart
    lat = Lattice([21, 17, 13], sharded=False)
    x = torch.rand(lat.n_rows, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    lat.action(0.0, x, y)            # a healthy Action on every rank
    ok &= api.health_check() == 0 and bool(torch.isfinite(y).all())
    dist.barrier()
    if rank == 0:
        lat.action(0.0, x, y)        # rank 1 never launches this one
        rc = api.health_check()
        nan_rows = int(torch.isnan(y).sum())
        print("after the missing peer: pfsp_check rc=%d, %d NaN rows in y" % (rc, nan_rows))
        ok &= rc != 0 and nan_rows > 0
    dist.barrier()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("TIMEOUT CHECK OK" if flag.item() == 1.0 else "TIMEOUT CHECK FAILED")
    del lat
    api.finalize()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
