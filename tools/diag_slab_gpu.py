"""Diagnostic (2-GPU box, torchrun): locate slab-vs-single-domain mismatches level by level."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import parallel_finite_difference_computation_b200 as fdw
from parallel_finite_difference_computation_b200 import distributed as D, _lib
import ctypes as C

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
n, nb = 1024, 40
nx = nz = n - 2 * nb
rng = np.random.default_rng(7)
v2 = np.full((n, n), np.float32(2500.0) ** 2, np.float32)
a0 = rng.standard_normal((n, n), dtype=np.float32); b0 = rng.standard_normal((n, n), dtype=np.float32)
srce = fdw.host.ricker_wavelet(64, 0.001, 25.0, fdw.FAMILY_GPU)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for taper in (fdw.TAPER_NONE, fdw.TAPER_TOP):
  for nt in (0, 1, 2, 5):
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, rank=rank, world=world, device=lrank, order=8,
                          fac=0.75, family=fdw.FAMILY_GPU, taper=taper, nt=64)
    sp.set_stream(stream.cuda_stream)
    x0, x1 = sp.slab
    sp.set_v2_local(v2[x0:x1]); sp.set_wavelet(srce); sp.set_source(n // 2 - 2, nb)
    na, nb_ = a0[x0:x1].copy(), b0[x0:x1].copy()
    if nt == 0:
        # aliasing check: after upload+refresh, the ghost rows must hold the neighbour's rows
        sp.upload_local(na, nb_)
        hl = _lib.Halo(); _lib.check(sp.L, sp.L.fdw_halo_get(sp.h, 0, C.byref(hl)))
        t_lo = sp._tensor(hl.recv_lo, hl.count).cpu().numpy().reshape(4, -1)[:, :n]
        t_hi = sp._tensor(hl.recv_hi, hl.count).cpu().numpy().reshape(4, -1)[:, :n]
        if rank == 0:
            print("rank0 ghost-above == rank1 rows 0..3:", np.array_equal(t_hi, a0[x1:x1 + 4]), flush=True)
        else:
            print("rank1 ghost-below == rank0 last rows:", np.array_equal(t_lo, a0[x0 - 4:x0]), flush=True)
        sp.close(); continue
    sp.propagate_local(na, nb_, 0, nt)
    torch.cuda.synchronize()
    parts = [None] * world
    dist.all_gather_object(parts, (na, nb_))
    if rank == 0:
        a, b = a0.copy(), b0.copy()
        with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU, taper=taper,
                        device=lrank, nt=64) as w:
            w.set_v2(v2); w.set_wavelet(srce); w.set_source(n // 2 - 2, nb); w.propagate(a, b, 0, nt)
        newest = np.concatenate([p[0] for p in parts]); older = np.concatenate([p[1] for p in parts])
        for nm, x, y in (("newest", newest, a), ("older", older, b)):
            bad = np.argwhere(x.view(np.uint32) != y.view(np.uint32))
            print("taper %d nt %d %s: %d bad" % (taper, nt, nm, len(bad)) + ("" if not len(bad) else
                  " rows %d..%d cols %d..%d" % (bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max())), flush=True)
    sp.close()
dist.barrier(); dist.destroy_process_group()
