"""Parity cases shared by the CPU-only emulation tests (host logic + kernel
bodies compiled for the host) and the GPU tests (the real CUDA library through
the C ABI).  Every case compares with the oracle on the same seeded inputs and
returns nothing; it asserts.  `lib` is a bound CDLL (emu or libfdwave.so)."""
import numpy as np

from oracle import oracle as O
from parallel_finite_difference_computation_b200 import (FAMILY_CPU, FAMILY_GPU, RECIPE_C, RECIPE_FAST, RECIPE_G,
                                                         SRC_GAUSS7, SRC_POINT, TAPER_FOUR, TAPER_NONE, TAPER_TOP,
                                                         Wave2D, stencil)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        raise AssertionError("%s: %d/%d words differ, first at %s, max abs diff %g (|ref|max %g)" % (
            what, len(bad), a.size, tuple(bad[0]), np.nanmax(np.abs(a - b)), np.abs(b).max()))


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-300))


def layered_v2(nx, nz, nxb, nzb, rng, random_border=False):
    """3-layer velocity (2000/3000/4000 m/s) + small lateral perturbation, extended."""
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    vp = np.empty((nx, nz), np.float32)
    vp[:, : nz // 3] = 2000.0
    vp[:, nz // 3: 2 * nz // 3] = 3000.0
    vp[:, 2 * nz // 3:] = 4000.0
    vp += rng.uniform(-50, 50, (nx, nz)).astype(np.float32)
    ve = np.zeros((nxe, nze), np.float32)
    ve[nxb:nxb + nx, nzb:nzb + nz] = vp
    if random_border:
        ve = O.extendvel_linear(nx, nz, nxb, nzb, ve, seed=7)
    else:
        ve = O.extendvel(nx, nz, nxb, nzb, ve)
    return (ve * ve).astype(np.float32)


# ------------------------------------------------------------------ config 1
def case_stencil(lib, order=8, shape=(61, 47), seed=1):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-1, 1, shape).astype(np.float32)
    got = stencil(p, order=order, dx=10.0, dz=12.5, lib=lib)
    assert_bit_equal(got, O.stencil(order, 10.0, 12.5, p), "stencil order %d %s" % (order, shape))


def case_stencil_golden(lib, golden_dir):
    import os
    p = np.fromfile(os.path.join(golden_dir, "stencil/input.bin"), np.float32).reshape(415, 295)
    g = np.fromfile(os.path.join(golden_dir, "stencil/output_teste.bin"), np.float32).reshape(415, 295)
    assert_bit_equal(stencil(p, order=8, dx=10.0, dz=10.0, lib=lib), g, "output_teste.bin")


# ------------------------------------------------------------------ plain propagation
def case_advance(lib, family, recipe, taper, order=8, nx=37, nz=29, nxb=9, nzb=8, nt=25, compat=False,
                 src_kind=SRC_POINT, seed=3, tol=None, random_init=False):
    """nt steps of the forward loop vs the oracle blocks composed in the reference's order."""
    rng = np.random.default_rng(seed)
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    dx, dz, dt, fac = 10.0, 12.5, 0.001, (0.6 if family == FAMILY_GPU else 0.11)
    v2 = layered_v2(nx, nz, nxb, nzb, rng)
    fam_o = O.FAM_G if family == FAMILY_GPU else O.FAM_C
    srce = O.ricker_wavelet(nt, dt, 30.0, fam_o)
    sx, sz = nxb + nx // 2, nzb + 1
    # ---- oracle
    tx, tz = O.taper_table(nxb, fac, fam_o), O.taper_table(nzb, fac, fam_o)
    a = np.zeros((nxe, nze), np.float32)  # newest
    b = np.zeros((nxe, nze), np.float32)  # older
    if random_init:
        a = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
        b = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    lap = np.zeros_like(a)
    h = order // 2
    if compat:
        ux, uz, uzb = (nxe // 8) * 8, (nze // 8) * 8, (nzb // 8) * 8
        ilim, jlim = h + ux, h + uz
        # The reference never updates rows >= ux / columns >= uz (quirk Q1) and only ever
        # starts from zero fields, so that region is identically zero in every reference
        # run; the library supports exactly that ("compat extents").
        a[ux:] = 0; a[:, uz:] = 0; b[ux:] = 0; b[:, uz:] = 0
    else:
        ux, uz, uzb, ilim, jlim = nxe, nze, nzb, nxe, nze
    a0, b0 = a.copy(), b.copy()
    cx, cz = O.premult_coefs(order, dx, dz)
    coefs = O.calc_coefs(order, fam_o)
    dx2inv, dz2inv, dt2 = O.scalars(dx, dz, dt)

    def sponge(f):
        if taper == TAPER_TOP:
            O.taper_top(f, nxb, tx, tz, ux, uzb)
        elif taper == TAPER_FOUR:
            O.taper_4(f, nx, nz, nxb, nzb, tx, tz)

    oracle_recipe = RECIPE_G if recipe == RECIPE_FAST else recipe
    for it in range(nt):
        if family == FAMILY_GPU:
            sponge(a)
            sponge(b)
        if oracle_recipe == RECIPE_G:
            O.lap_G(order, a, cx, cz, ilim, jlim, out=lap)
        else:
            O.lap_C(order, a, coefs, dx2inv, dz2inv, out=lap)
        O.time_update(a, b, v2, lap, dt2, ux, uz)
        if src_kind == SRC_POINT:
            b[sx, sz] += srce[it]
        else:
            O.ptsrc(sx, sz, srce[it], b)
        if family == FAMILY_CPU:
            sponge(b)
            sponge(a)
        a, b = b, a
    # ---- candidate
    with Wave2D(nx, nz, nxb, nzb, dx, dz, dt, order=order, fac=fac, family=family, recipe=recipe, taper=taper,
                compat_extents=compat, nt=nt, lib=lib) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.set_source(sx, sz, src_kind)
        if random_init:
            w.upload(a0, b0)
        else:
            w.zero()
        w.advance(0, nt)
        newest, older = w.download()
    what = "advance fam=%d recipe=%d taper=%d order=%d compat=%d" % (family, recipe, taper, order, compat)
    if tol is None:
        assert_bit_equal(newest, a, what + " newest")
        assert_bit_equal(older, b, what + " older")
    else:
        assert rel_l2(newest, a) < tol and rel_l2(older, b) < tol, (what, rel_l2(newest, a), rel_l2(older, b))


# ------------------------------------------------------------------ GPU-family RTM shot
def case_gpu_rtm(lib, nx=41, nz=33, nxb=8, nzb=8, nt=60, compat=True, seed=5, host_roundtrip=False, order=8):
    rng = np.random.default_rng(seed)
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    dx = dz = 10.0
    dt, fac, fpeak = 0.001, 0.75, 30.0
    v2 = layered_v2(nx, nz, nxb, nzb, rng, random_border=True)
    srce = O.ricker_wavelet(nt, dt, fpeak, O.FAM_G)
    sx, sz, gz = nxb + nx // 3, nzb, nzb
    dobs = rng.uniform(-1, 1, (nx, nt)).astype(np.float32)
    cfg = O.GpuCfg(order, nxe, nze, nxb, nzb, nt, dx, dz, dt, fac, int(compat))
    P, PP = O.gpu_forward(cfg, v2, srce, sx, sz)
    im = O.gpu_back(cfg, P, PP, v2, dobs, gz)
    with Wave2D(nx, nz, nxb, nzb, dx, dz, dt, order=order, fac=fac, family=FAMILY_GPU, taper=TAPER_TOP,
                compat_extents=compat, nt=nt, lib=lib) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        if host_roundtrip:
            gP, gPP = w.forward(sx, sz)
            assert_bit_equal(gP, P, "fd_forward P")
            assert_bit_equal(gPP, PP, "fd_forward PP")
            gim = w.backward(dobs, gz, gP, gPP)
        else:
            w.forward(sx, sz, download=False)
            gim = w.backward(dobs, gz)
    assert np.abs(im).max() > 0
    assert_bit_equal(gim, im, "fd_back image compat=%d roundtrip=%d" % (compat, host_roundtrip))


# ------------------------------------------------------------------ CPU family programs
def case_mod_shot(lib, nx=33, nz=27, nxb=7, nzb=6, nt=50, seed=9, order=8, gz=None):
    rng = np.random.default_rng(seed)
    dx, dz, dt, fac, fpeak = 10.0, 10.0, 0.001, 0.05, 35.0
    v2 = layered_v2(nx, nz, nxb, nzb, rng)
    srce = O.ricker_wavelet(nt, dt, fpeak, O.FAM_C)
    sx, sz, gz = nxb + 2, nzb, (nzb if gz is None else gz)
    cfg = O.CpuCfg(order, nx, nz, nxb, nzb, nt, dx, dz, dt, fac)
    want = O.mod_shot(cfg, v2, srce, sx, sz, gz)
    with Wave2D(nx, nz, nxb, nzb, dx, dz, dt, order=order, fac=fac, family=FAMILY_CPU, taper=TAPER_FOUR, nt=nt,
                lib=lib) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        got = w.model_shot(sx, sz, gz)
    assert np.abs(want).max() > 0
    assert_bit_equal(got, want, "mod_main seismogram")


def case_rtm_shot_cpu(lib, nx=31, nz=31, nxb=7, nzb=7, nt=40, ns=2, is_=0, seed=11, order=8):
    rng = np.random.default_rng(seed)
    dx, dz, dt, fac, fpeak = 10.0, 10.0, 0.001, 0.05, 35.0
    v2 = layered_v2(nx, nz, nxb, nzb, rng)
    srce = O.ricker_wavelet(nt, dt, fpeak, O.FAM_C)
    sx, sz, gz = nxb + 5, nzb, nzb
    dobs = rng.uniform(-1, 1, (ns, nx, nt)).astype(np.float32)
    cfg = O.CpuCfg(order, nx, nz, nxb, nzb, nt, dx, dz, dt, fac)
    want = O.rtm_shot(cfg, v2, srce, sx, sz, gz, dobs, is_)
    with Wave2D(nx, nz, nxb, nzb, dx, dz, dt, order=order, fac=fac, family=FAMILY_CPU, taper=TAPER_TOP, nt=nt,
                history=True, lib=lib) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        got = w.rtm_shot_cpu(sx, sz, gz, dobs, is_)
    assert np.abs(want).max() > 0
    assert_bit_equal(got, want, "rtm_main image is=%d" % is_)
