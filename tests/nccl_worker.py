"""Worker of tests/test_gpu_nccl.py: one rank of a 2-GPU job.  A small stepped channel is cut into one part per rank
and marched over the real transports -- the library's own NCCL communicator (scg_md_march_dist) and the
torch.distributed one -- and compared on rank 0, bit for bit, with the undecomposed device march."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from solvcon_b200 import _lib, bench_md
    from solvcon_b200.decomp import comm_init
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.init(local)
    comm_init(rank, world, dist=dist)
    ok = True
    for transport in ("lib", "torch"):
        res, good = bench_md.nccl_parity_3d(rank, world, local, dist, "rcm", ncell_target=60_000, steps=4, transport=transport)
        ok = ok and good
        if rank == 0:
            print(transport, res, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
