"""CPU-only: the peer-memory halo exchange (fdw_peer_*: boundary rows stored straight into the
neighbour's ghost rows by the step kernel, flag release/acquire instead of a communication
library) must reproduce the single-domain result bit for bit.

Runs on the host build (tests/emu): there "CUDA IPC" hands over plain pointers, so the slab
contexts of ONE process can map each other, and kernels execute synchronously -- the slabs are
therefore stepped in lock-step, one level at a time, and an acquire wait that finds its flag unmet
is reported as a protocol error instead of spinning.  The real thing (one process per GPU, NVLink)
is checked by tests/run_slab_gpu.py on a multi-GPU box."""
import ctypes as C

import numpy as np
import pytest

import parity_cases as PC
from emu_loader import load as load_emu
from oracle import oracle as O
from parallel_finite_difference_computation_b200 import (FAMILY_CPU, FAMILY_GPU, SRC_GAUSS7, SRC_POINT, TAPER_FOUR,
                                                         TAPER_TOP, FdwError, Wave2D, _lib)
from parallel_finite_difference_computation_b200.distributed import slab_rows


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def _attach_all(emu, slabs):
    infos = []
    for w in slabs:
        pi = _lib.PeerInfo()
        _lib.check(emu, emu.fdw_peer_export(w.h, C.byref(pi)))
        infos.append(pi)
    for r, w in enumerate(slabs):
        lo = C.byref(infos[r - 1]) if r > 0 else None
        hi = C.byref(infos[r + 1]) if r < len(slabs) - 1 else None
        _lib.check(emu, emu.fdw_peer_attach(w.h, lo, hi))


def _lockstep(emu, slabs, it0, n):
    for it in range(it0, it0 + n):
        for w in slabs:
            _lib.check(emu, emu.fdw_peer_levels(w.h, it, 1))
    for w in slabs:
        _lib.check(emu, emu.fdw_peer_fence(w.h))
        w.sync()  # reports a failed acquire


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("family,taper", [(FAMILY_GPU, TAPER_TOP), (FAMILY_CPU, TAPER_FOUR)])
def test_peer_push_matches_single_domain_bitwise(emu, world, family, taper):
    rng = np.random.default_rng(7)
    nx, nz, nxb, nzb, nt = 53, 37, 9, 8, 14
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    v2 = PC.layered_v2(nx, nz, nxb, nzb, rng)
    a = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    b = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_G if family == FAMILY_GPU else O.FAM_C)
    fac = 0.6 if family == FAMILY_GPU else 0.11
    kw = dict(order=8, fac=fac, family=family, taper=taper, nt=nt, lib=emu)
    # single domain
    with Wave2D(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, **kw) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.set_source(nxb + nx // 2, nzb + 1, SRC_POINT)
        ra, rb = a.copy(), b.copy()
        w.propagate(ra, rb, 0, nt)
    # slabs, peer exchange
    slabs = [Wave2D(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, slab=slab_rows(nxe, world, r), **kw) for r in range(world)]
    try:
        _attach_all(emu, slabs)
        for r, w in enumerate(slabs):
            x0, x1 = slab_rows(nxe, world, r)
            _lib.check(emu, emu.fdw_set_v2_local(w.h, np.ascontiguousarray(v2[x0:x1])))
            w.set_wavelet(srce)
            w.set_source(nxb + nx // 2, nzb + 1, SRC_POINT)
            _lib.check(emu, emu.fdw_fields_upload_local(w.h, 0, np.ascontiguousarray(a[x0:x1]),
                                                        np.ascontiguousarray(b[x0:x1])))
        for w in slabs:
            _lib.check(emu, emu.fdw_peer_refresh(w.h))
        _lockstep(emu, slabs, 0, nt)
        for r, w in enumerate(slabs):
            x0, x1 = slab_rows(nxe, world, r)
            n = np.zeros((x1 - x0, nze), np.float32)
            o = np.zeros((x1 - x0, nze), np.float32)
            _lib.check(emu, emu.fdw_fields_download_local(w.h, 0, n.ctypes.data_as(C.c_void_p),
                                                          o.ctypes.data_as(C.c_void_p)))
            PC.assert_bit_equal(n, ra[x0:x1], "newest, slab %d/%d" % (r, world))
            PC.assert_bit_equal(o, rb[x0:x1], "older, slab %d/%d" % (r, world))
    finally:
        for w in slabs:
            w.close()


def test_peer_modelling_shot_records_like_single_domain(emu):
    """mod_main shot (7x7 source, four-sided sponge, seismogram epilogue) through the peer path"""
    rng = np.random.default_rng(11)
    nx, nz, nxb, nzb, nt, world = 41, 29, 8, 8, 12, 2
    nxe = nx + 2 * nxb
    v2 = PC.layered_v2(nx, nz, nxb, nzb, rng)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_C)
    kw = dict(order=8, fac=0.11, family=FAMILY_CPU, recipe=_lib.RECIPE_C, taper=TAPER_FOUR, nt=nt, lib=emu)
    sx, sz, gz = nxb + nx // 2, nzb + 2, nzb + 1
    with Wave2D(nx, nz, nxb, nzb, 10.0, 10.0, 0.001, **kw) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        want = w.model_shot(sx, sz, gz)
    slabs = [Wave2D(nx, nz, nxb, nzb, 10.0, 10.0, 0.001, slab=slab_rows(nxe, world, r), **kw) for r in range(world)]
    try:
        _attach_all(emu, slabs)
        for r, w in enumerate(slabs):
            x0, x1 = slab_rows(nxe, world, r)
            _lib.check(emu, emu.fdw_set_v2_local(w.h, np.ascontiguousarray(v2[x0:x1])))
            w.set_wavelet(srce)
            _lib.check(emu, emu.fdw_shot_begin(w.h, _lib.PHASE_MODEL, sx, sz, gz, None, 1, 0))
        for w in slabs:
            _lib.check(emu, emu.fdw_peer_refresh(w.h))
        _lockstep(emu, slabs, 0, nt)
        rows = []
        for w in slabs:
            d = w.devinfo()
            out = np.zeros((max(d.nli, 1), nt), np.float32)
            _lib.check(emu, emu.fdw_shot_end(w.h, out))
            rows.append(out[: d.nli])
        PC.assert_bit_equal(np.concatenate(rows), want, "seismograms")
    finally:
        for w in slabs:
            w.close()


def test_peer_protocol_errors(emu):
    nx, nz, nb = 24, 16, 4
    nxe = nx + 2 * nb
    kw = dict(order=8, fac=0.5, family=FAMILY_GPU, taper=TAPER_TOP, nt=4, lib=emu)
    a = Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, slab=(0, 16), **kw)
    b = Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, slab=(16, nxe), **kw)
    c = Wave2D(nx, nz + 32, nb, nb, 10.0, 10.0, 0.001, slab=(16, nxe), **kw)  # different pitch
    try:
        with pytest.raises(FdwError) as e:  # stepping without neighbours
            _lib.check(emu, emu.fdw_peer_levels(a.h, 0, 1))
        assert e.value.code == -5
        ia, ib, ic = _lib.PeerInfo(), _lib.PeerInfo(), _lib.PeerInfo()
        for w, i in ((a, ia), (b, ib), (c, ic)):
            _lib.check(emu, emu.fdw_peer_export(w.h, C.byref(i)))
        with pytest.raises(FdwError) as e:  # wrong side: b is above a, not below
            _lib.check(emu, emu.fdw_peer_attach(a.h, C.byref(ib), None))
        assert e.value.code == -1 and "adjoin" in str(e.value)
        with pytest.raises(FdwError):  # pitch mismatch
            _lib.check(emu, emu.fdw_peer_attach(a.h, None, C.byref(ic)))
        _lib.check(emu, emu.fdw_peer_attach(a.h, None, C.byref(ib)))
        _lib.check(emu, emu.fdw_peer_attach(b.h, C.byref(ia), None))
        for w in (a, b):
            w.set_v2(np.full((nxe, nz + 2 * nb), 4e6, np.float32))
            w.set_wavelet(np.ones(4, np.float32))
            w.set_source(nb + 3, nb + 3, SRC_POINT)
        # a runs two levels while b has not moved: the second acquire finds b's flag unmet
        _lib.check(emu, emu.fdw_peer_levels(a.h, 0, 1))
        _lib.check(emu, emu.fdw_peer_levels(a.h, 1, 1))
        with pytest.raises(FdwError) as e:
            a.sync()
        assert "never arrived" in str(e.value)
    finally:
        for w in (a, b, c):
            w.close()


def test_async_local_round_trip_equals_synchronous(emu):
    """propagate_local_async (upload, levels and download enqueued without a final synchronisation; the
    building block of the double-buffered end-to-end loop in bench.py) gives what propagate_local gives"""
    from parallel_finite_difference_computation_b200 import distributed as D
    rng = np.random.default_rng(5)
    nx, nz, nb, nt = 33, 27, 8, 9
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    v2 = PC.layered_v2(nx, nz, nb, nb, rng)
    a = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    b = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_G)
    outs = []
    for use_async in (False, True):
        sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, lib=emu, on_gpu=False, order=8, fac=0.6,
                              family=FAMILY_GPU, taper=TAPER_TOP, nt=nt)
        sp.set_v2_local(v2)
        sp.set_wavelet(srce)
        sp.set_source(nb + 11, nb + 2, SRC_POINT)
        x, y = a.copy(), b.copy()
        if use_async:
            sp.propagate_local_async(x, y, 0, nt)
            sp.sync()
        else:
            sp.propagate_local(x, y, 0, nt)
        outs.append((x, y))
        sp.close()
    PC.assert_bit_equal(outs[1][0], outs[0][0], "async newest")
    PC.assert_bit_equal(outs[1][1], outs[0][1], "async older")


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("fuse,glevels", [("1", 8), ("0", 8), ("1", 2), ("1", 4)])
def test_peer_level_loop_graph_replay_threads(emu, world, fuse, glevels, monkeypatch):
    """The whole level loop in ONE fdw_peer_levels call per slab (as on the GPUs): the first two levels are issued
    directly, every further group of FDW_GRAPH_LEVELS levels (default 8) is recorded, turned into a graph and
    replayed with refreshed node arguments, a remainder shorter than a group is issued directly
    (the host stand-in keeps the kernel nodes and runs them in creation order).  Each slab runs in its own host
    thread and the acquire really waits (FDW_EMU_SPIN), so the slabs are coupled only through the flags --
    with the acquire/release inside the boundary kernels (fuse=1) and as stand-alone kernels (fuse=0)."""
    import threading
    monkeypatch.setenv("FDW_EMU_SPIN", "1")
    monkeypatch.setenv("FDW_FUSE_FLAGS", fuse)
    monkeypatch.setenv("FDW_GRAPH", "1")
    monkeypatch.setenv("FDW_GRAPH_LEVELS", str(glevels))
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    rng = np.random.default_rng(23)
    nx, nz, nxb, nzb, nt = 47, 33, 8, 8, 21  # 2 direct levels + 19: 2 groups of 8 (+3 direct) / 4 of 4 (+3) / 9 pairs (+1)
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    v2 = PC.layered_v2(nx, nz, nxb, nzb, rng)
    a = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    b = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_G)
    kw = dict(order=8, fac=0.6, family=FAMILY_GPU, taper=TAPER_TOP, nt=nt, lib=emu)
    with Wave2D(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, **kw) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.set_source(nxb + nx // 2, nzb + 1, SRC_POINT)
        ra, rb = a.copy(), b.copy()
        w.propagate(ra, rb, 0, nt)
    launches0 = C.c_longlong.in_dll(emu, "emu_graph_launches").value
    slabs = [Wave2D(nx, nz, nxb, nzb, 10.0, 12.5, 0.001, slab=slab_rows(nxe, world, r), **kw) for r in range(world)]
    try:
        _attach_all(emu, slabs)
        for r, w in enumerate(slabs):
            x0, x1 = slab_rows(nxe, world, r)
            _lib.check(emu, emu.fdw_set_v2_local(w.h, np.ascontiguousarray(v2[x0:x1])))
            w.set_wavelet(srce)
            w.set_source(nxb + nx // 2, nzb + 1, SRC_POINT)
            _lib.check(emu, emu.fdw_fields_upload_local(w.h, 0, np.ascontiguousarray(a[x0:x1]),
                                                        np.ascontiguousarray(b[x0:x1])))
        for w in slabs:
            _lib.check(emu, emu.fdw_peer_refresh(w.h))
        errors = []

        def run(w):
            try:
                _lib.check(emu, emu.fdw_peer_levels(w.h, 0, nt))
                _lib.check(emu, emu.fdw_peer_fence(w.h))
                w.sync()
            except Exception as e:  # noqa: BLE001 -- reported by the main thread
                errors.append(e)

        threads = [threading.Thread(target=run, args=(w,)) for w in slabs]
        for t in threads:
            t.start()
        for t in threads:
            t.join(120)
        assert not errors, errors
        assert not any(t.is_alive() for t in threads)
        assert C.c_longlong.in_dll(emu, "emu_graph_launches").value - launches0 == world * ((nt - 2) // glevels)
        for r, w in enumerate(slabs):
            x0, x1 = slab_rows(nxe, world, r)
            n = np.zeros((x1 - x0, nze), np.float32)
            o = np.zeros((x1 - x0, nze), np.float32)
            _lib.check(emu, emu.fdw_fields_download_local(w.h, 0, n.ctypes.data_as(C.c_void_p),
                                                          o.ctypes.data_as(C.c_void_p)))
            PC.assert_bit_equal(n, ra[x0:x1], "newest, slab %d/%d" % (r, world))
            PC.assert_bit_equal(o, rb[x0:x1], "older, slab %d/%d" % (r, world))
    finally:
        for w in slabs:
            w.close()


def test_peer_rtm_main_shot_domain_divided_threads(emu, monkeypatch):
    """config 5 in miniature: rtm_main's shot (forward with the history sharded over the slabs, backward with
    back-injection and imaging) through the peer path, graph replay on, one host thread per slab"""
    import threading
    monkeypatch.setenv("FDW_EMU_SPIN", "1")
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    rng = np.random.default_rng(31)
    nx, nz, nb, nt, world = 45, 31, 8, 12, 3
    nxe = nx + 2 * nb
    v2 = PC.layered_v2(nx, nz, nb, nb, rng)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_C)
    dobs = rng.uniform(-1, 1, (1, nx, nt)).astype(np.float32)
    sx, sz, gz = nb + nx // 2 - 2, nb + 1, nb
    kw = dict(order=8, fac=0.11, family=FAMILY_CPU, recipe=_lib.RECIPE_C, taper=TAPER_TOP, nt=nt, history=True, lib=emu)
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, **kw) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        want = w.rtm_shot_cpu(sx, sz, gz, dobs, 0)
    assert np.abs(want).max() > 0
    slabs = [Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, slab=slab_rows(nxe, world, r), **kw) for r in range(world)]
    try:
        _attach_all(emu, slabs)
        for r, w in enumerate(slabs):
            x0, x1 = slab_rows(nxe, world, r)
            _lib.check(emu, emu.fdw_set_v2_local(w.h, np.ascontiguousarray(v2[x0:x1])))
            w.set_wavelet(srce)
        rows, errors = [None] * world, []

        def run(r, w):
            try:
                for phase, traces in ((_lib.PHASE_RTM_FWD, None), (_lib.PHASE_RTM_BWD, dobs)):
                    ptr = traces.ctypes.data_as(C.c_void_p) if traces is not None else None
                    _lib.check(emu, emu.fdw_shot_begin(w.h, phase, sx, sz, gz, ptr, 1, 0))
                    _lib.check(emu, emu.fdw_peer_refresh(w.h))  # the phase zeroed the fields
                    _lib.check(emu, emu.fdw_peer_fence(w.h))
                    _lib.check(emu, emu.fdw_peer_levels(w.h, 0, nt))
                    _lib.check(emu, emu.fdw_peer_fence(w.h))
                d = w.devinfo()
                out = np.zeros((max(d.nli, 1), nz), np.float32)
                _lib.check(emu, emu.fdw_shot_end(w.h, out))
                rows[r] = out[: d.nli]
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        threads = [threading.Thread(target=run, args=(r, w)) for r, w in enumerate(slabs)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(120)
        assert not errors, errors
        PC.assert_bit_equal(np.concatenate(rows), want, "domain-divided image")
    finally:
        for w in slabs:
            w.close()
