"""CPU-only, world_size 2 and 3 over gloo: the slab-decomposed driver
(parallel_finite_difference_computation_b200.distributed) must reproduce the
single-domain result bit for bit.  Runs the library's host logic + kernel
bodies through tests/emu (no GPU here); the transport is real torch.distributed."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, family, taper, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch.distributed as dist

    import parity_cases as PC
    from emu_loader import load as load_emu
    from oracle import oracle as O
    from parallel_finite_difference_computation_b200 import FAMILY_GPU, SRC_POINT, distributed as D

    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = load_emu()
    rng = np.random.default_rng(42)  # same on every rank
    nx, nz, nxb, nzb, nt = 53, 37, 9, 8, 14
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    v2 = PC.layered_v2(nx, nz, nxb, nzb, rng)
    a = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    b = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    fam_o = O.FAM_G if family == FAMILY_GPU else O.FAM_C
    srce = O.ricker_wavelet(nt, 0.001, 30.0, fam_o)
    fac = 0.6 if family == FAMILY_GPU else 0.11
    sp = D.SlabPropagator(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, rank=rank, world=world, lib=emu, on_gpu=False,
                          order=8, fac=fac, family=family, taper=taper, nt=nt)
    x0, x1 = sp.slab
    sp.set_v2_local(v2[x0:x1])
    sp.set_wavelet(srce)
    sp.set_source(nxb + nx // 2, nzb + 1, SRC_POINT)  # the source sits in one slab, near a cut for world=3
    na, nb_ = np.ascontiguousarray(a[x0:x1]), np.ascontiguousarray(b[x0:x1])
    sp.propagate_local(na, nb_, 0, nt)
    np.save(os.path.join(outdir, "newest_%d.npy" % rank), na)
    np.save(os.path.join(outdir, "older_%d.npy" % rank), nb_)
    img = D.reduce_image(np.full((4, 5), np.float32(rank + 1)), "ordered")
    assert np.all(img == sum(range(1, world + 1)))
    img = D.reduce_image(np.full((4, 5), np.float32(rank + 1)), "allreduce")
    assert np.all(img == sum(range(1, world + 1)))
    sp.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("family,taper", [(0, 1), (1, 2)])
def test_slab_matches_single_domain_bitwise(tmp_path, world, family, taper):
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_cases as PC
    from emu_loader import load as load_emu
    from oracle import oracle as O
    from parallel_finite_difference_computation_b200 import FAMILY_GPU, SRC_POINT, Wave2D, distributed as D

    emu = load_emu()  # build once before the workers race for it
    mp.spawn(_worker, args=(world, _free_port(), family, taper, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    nx, nz, nxb, nzb, nt = 53, 37, 9, 8, 14
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    v2 = PC.layered_v2(nx, nz, nxb, nzb, rng)
    a = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    b = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    fam_o = O.FAM_G if family == FAMILY_GPU else O.FAM_C
    srce = O.ricker_wavelet(nt, 0.001, 30.0, fam_o)
    fac = 0.6 if family == FAMILY_GPU else 0.11
    with Wave2D(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, order=8, fac=fac, family=family, taper=taper, nt=nt,
                lib=emu) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.set_source(nxb + nx // 2, nzb + 1, SRC_POINT)
        w.propagate(a, b, 0, nt)
    newest = np.concatenate([np.load(tmp_path / ("newest_%d.npy" % r)) for r in range(world)])
    older = np.concatenate([np.load(tmp_path / ("older_%d.npy" % r)) for r in range(world)])
    PC.assert_bit_equal(newest, a, "slab newest world=%d" % world)
    PC.assert_bit_equal(older, b, "slab older world=%d" % world)


def test_partitions():
    from parallel_finite_difference_computation_b200 import distributed as D
    for nxe in (71, 16384, 131072):
        for world in (1, 2, 3, 8):
            rows = [D.slab_rows(nxe, world, r) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == nxe
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1
    assert sorted(sum((D.shot_partition(64, 8, r) for r in range(8)), [])) == list(range(64))
    assert sorted(sum((D.shot_partition(7, 3, r) for r in range(3)), [])) == list(range(7))
