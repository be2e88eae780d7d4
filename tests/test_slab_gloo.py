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


def _worker(rank, world, port, family, taper, outdir, halo=None):
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
    if halo == "p2p":  # the peer mapping fails on rank 1: every slab must fall back to send/recv together
        os.environ["FDW_TEST_FAIL_PEER_ATTACH"] = "1"
    sp = D.SlabPropagator(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, rank=rank, world=world, lib=emu, on_gpu=False,
                          order=8, fac=fac, family=family, taper=taper, nt=nt, halo=halo)
    if halo == "p2p":
        assert sp.halo == "nccl" and not sp.p2p
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


def test_peer_mapping_failure_falls_back_to_send_recv_on_every_rank(tmp_path):
    """halo='p2p' requested, one rank cannot map its neighbours: the slabs agree (all_reduce) to use
    send/recv instead; the result is still the single-domain one bit for bit"""
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_cases as PC
    from emu_loader import load as load_emu
    from oracle import oracle as O
    from parallel_finite_difference_computation_b200 import FAMILY_GPU, SRC_POINT, TAPER_TOP, Wave2D

    emu = load_emu()
    world = 3
    mp.spawn(_worker, args=(world, _free_port(), FAMILY_GPU, TAPER_TOP, str(tmp_path), "p2p"), nprocs=world, join=True)
    rng = np.random.default_rng(42)
    nx, nz, nxb, nzb, nt = 53, 37, 9, 8, 14
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    v2 = PC.layered_v2(nx, nz, nxb, nzb, rng)
    a = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    b = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_G)
    with Wave2D(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, order=8, fac=0.6, family=FAMILY_GPU, taper=TAPER_TOP, nt=nt,
                lib=emu) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.set_source(nxb + nx // 2, nzb + 1, SRC_POINT)
        w.propagate(a, b, 0, nt)
    newest = np.concatenate([np.load(tmp_path / ("newest_%d.npy" % r)) for r in range(world)])
    older = np.concatenate([np.load(tmp_path / ("older_%d.npy" % r)) for r in range(world)])
    PC.assert_bit_equal(newest, a, "fallback newest")
    PC.assert_bit_equal(older, b, "fallback older")


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


# ------------------------------------------------------------------ shot parallelism
def _shot_case():
    import parity_cases as PC
    from oracle import oracle as O
    rng = np.random.default_rng(17)
    nx, nz, nb, nt, ns = 37, 29, 8, 30, 5
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    v2s = [PC.layered_v2(nx, nz, nb, nb, np.random.default_rng(100 + k), random_border=True) for k in range(ns)]
    dobs = rng.uniform(-1, 1, (ns, nx, nt)).astype(np.float32)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_G)
    return nx, nz, nb, nt, ns, v2s, dobs, srce


def _shot_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch.distributed as dist
    from emu_loader import load as load_emu
    from parallel_finite_difference_computation_b200 import FAMILY_GPU, TAPER_TOP, Wave2D, distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = load_emu()
    nx, nz, nb, nt, ns, v2s, dobs, srce = _shot_case()
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=FAMILY_GPU, taper=TAPER_TOP,
                compat_extents=True, nt=nt, lib=emu) as w:
        w.set_wavelet(srce)
        for mode, contiguous in (("chain", True), ("allreduce", False)):
            shots = D.shot_partition(ns, world, rank, contiguous=contiguous)
            img = D.migrate_shots_gpu_family(w, shots, lambda k: v2s[k], lambda k: dobs[k],
                                             lambda k: nb + 3 + 6 * k, nb, nb, stack=mode)
            if rank == 0:
                np.save(os.path.join(outdir, "img_%s.npy" % mode), img)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shot_parallel_image_stack(tmp_path, world):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_cases as PC
    from emu_loader import load as load_emu
    from oracle import oracle as O
    load_emu()
    mp.spawn(_shot_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    nx, nz, nb, nt, ns, v2s, dobs, srce = _shot_case()
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    want = np.zeros((nx, nz), np.float32)
    for k in range(ns):  # the reference's sequential loop, with the oracle
        cfg = O.GpuCfg(8, nxe, nze, nb, nb, nt, 10.0, 10.0, 0.001, 0.75, 1)
        P, PP = O.gpu_forward(cfg, v2s[k], srce, nb + 3 + 6 * k, nb)
        want += O.gpu_back(cfg, P, PP, v2s[k], dobs[k], nb)
    PC.assert_bit_equal(np.load(tmp_path / "img_chain.npy"), want, "chained stack vs sequential oracle")
    got = np.load(tmp_path / "img_allreduce.npy")
    assert PC.rel_l2(got, want) < 1e-6


# ------------------------------------------------------------------ config 5: domain-divided mod_main + rtm_main
def _c5_case():
    import parity_cases as PC
    from oracle import oracle as O
    rng = np.random.default_rng(23)
    nx, nz, nb, nt = 47, 31, 7, 36
    v2 = PC.layered_v2(nx, nz, nb, nb, rng)
    srce = O.ricker_wavelet(nt, 0.001, 35.0, O.FAM_C)
    return nx, nz, nb, nt, v2, srce


def _c5_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch.distributed as dist
    from emu_loader import load as load_emu
    from parallel_finite_difference_computation_b200 import FAMILY_CPU, TAPER_FOUR, TAPER_TOP, distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = load_emu()
    nx, nz, nb, nt, v2, srce = _c5_case()
    sx, sz, gz = nb + nx // 2, nb, nb
    # mod_main, slab-decomposed
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, rank=rank, world=world, lib=emu, on_gpu=False,
                          order=8, fac=0.05, family=FAMILY_CPU, taper=TAPER_FOUR, nt=nt)
    x0, x1 = sp.slab
    sp.set_v2_local(v2[x0:x1])
    sp.set_wavelet(srce)
    data = sp.gather_rows(sp.model_shot(sx, sz, gz))
    sp.close()
    # rtm_main, slab-decomposed (history sharded with the slabs)
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, rank=rank, world=world, lib=emu, on_gpu=False,
                          order=8, fac=0.05, family=FAMILY_CPU, taper=TAPER_TOP, nt=nt, history=True)
    sp.set_v2_local(v2[x0:x1])
    sp.set_wavelet(srce)
    img = sp.gather_rows(sp.rtm_shot_cpu(sx, sz, gz, data[None], 0))
    sp.close()
    if rank == 0:
        np.save(os.path.join(outdir, "data.npy"), data)
        np.save(os.path.join(outdir, "img.npy"), img)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_domain_divided_mod_main_rtm_main(tmp_path, world):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_cases as PC
    from emu_loader import load as load_emu
    from oracle import oracle as O
    load_emu()
    mp.spawn(_c5_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    nx, nz, nb, nt, v2, srce = _c5_case()
    cfg = O.CpuCfg(8, nx, nz, nb, nb, nt, 10.0, 10.0, 0.001, 0.05)
    want = O.mod_shot(cfg, v2, srce, nb + nx // 2, nb, nb)
    PC.assert_bit_equal(np.load(tmp_path / "data.npy"), want, "domain-divided mod_main seismogram")
    wimg = O.rtm_shot(cfg, v2, srce, nb + nx // 2, nb, nb, want[None], 0)
    assert np.abs(wimg).max() > 0
    PC.assert_bit_equal(np.load(tmp_path / "img.npy"), wimg, "domain-divided rtm_main image")
