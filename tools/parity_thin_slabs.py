"""N-vs-1 bitwise check of the slab-decomposed propagator on THIN slabs (the persistent slab kernel's range) for a few
grid sizes -- widths that exercise every CTA width the kernel picks.  Run under torchrun on >= 2 GPUs:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/parity_thin_slabs.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as B  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, device = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(device)
    dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    bad = 0
    for n in (700, 1200, 2128, 2500):
        v = B.parity_n_vs_1(rank, world, device, n=n, nt=60)
        if rank == 0:
            print(json.dumps(v), flush=True)
            bad += v["result"] != "bitwise"
    dist.destroy_process_group()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
