"""Multi-GPU check, run under torchrun on a GPU box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_slab_gpu.py

Every rank propagates its slab of a 4096 x 4096 grid for 40 levels with the
NCCL halo exchange; rank 0 also runs the whole grid alone and the gathered
slabs must equal it bit for bit (decomposition does not change the
per-point arithmetic)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402
from parallel_finite_difference_computation_b200 import distributed as D  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    n, nb, nt = 4096, 40, 40
    nx = nz = n - 2 * nb
    rng = np.random.default_rng(7)
    ve = np.empty((n, n), np.float32)
    ve[:, : n // 3] = 2000.0
    ve[:, n // 3:] = 3500.0
    v2 = ve * ve
    a = rng.standard_normal((n, n), dtype=np.float32)
    b = rng.standard_normal((n, n), dtype=np.float32)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 25.0, fdw.FAMILY_GPU)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, rank=rank, world=world, device=lrank, order=8,
                          fac=0.75, family=fdw.FAMILY_GPU, taper=fdw.TAPER_TOP, nt=nt)
    sp.set_stream(stream.cuda_stream)
    x0, x1 = sp.slab
    sp.set_v2_local(v2[x0:x1])
    sp.set_wavelet(srce)
    sp.set_source(D.slab_rows(n, world, 0)[1] - 2, nb)  # the same global point on every rank, next to a slab cut
    na, nb_ = a[x0:x1].copy(), b[x0:x1].copy()  # copies: propagate_local works in place
    sp.propagate_local(na, nb_, 0, nt)
    torch.cuda.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, (na, nb_))
    ok = True
    if rank == 0:
        with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU,
                        taper=fdw.TAPER_TOP, device=lrank, nt=nt) as w:
            w.set_v2(v2)
            w.set_wavelet(srce)
            w.set_source(D.slab_rows(n, world, 0)[1] - 2, nb)
            w.propagate(a, b, 0, nt)
        newest = np.concatenate([p[0] for p in parts])
        older = np.concatenate([p[1] for p in parts])
        ok = np.array_equal(newest.view(np.uint32), a.view(np.uint32)) and \
            np.array_equal(older.view(np.uint32), b.view(np.uint32))
        print("slab x%d vs single domain bitwise: %s" % (world, "OK" if ok else "MISMATCH"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
