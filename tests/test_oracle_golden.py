"""The oracle pinned against the reference's own golden files and against
vectors produced by the reference's own objects (tests/golden/make_golden.py).
CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle as O


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b):
    assert a.shape == b.shape
    assert np.array_equal(bits(a), bits(b)), "max abs diff %g" % np.abs(a - b).max()


def test_config1_stencil_bit_exact(golden_dir):
    """dpct_migrated_stencil_computation/output_teste.bin."""
    p = np.fromfile(os.path.join(golden_dir, "stencil/input.bin"), np.float32).reshape(415, 295)
    g = np.fromfile(os.path.join(golden_dir, "stencil/output_teste.bin"), np.float32).reshape(415, 295)
    assert_bit_equal(O.stencil(8, 10.0, 10.0, p), g)
    assert not g[:4].any() and not g[:, :4].any() and not g[-4:].any() and not g[:, -4:].any()


@pytest.fixture(scope="module")
def lay3(golden_dir):
    d = os.path.join(golden_dir, "3lay_mod")
    nx = nz = 151
    nxb = nzb = 40
    nt = 1001
    vp = np.fromfile(os.path.join(d, "3layer_151x151.bin"), np.float32).reshape(nx, nz)
    v2 = np.zeros((nx + 2 * nxb, nz + 2 * nzb), np.float32)
    v2[nxb:nxb + nx, nzb:nzb + nz] = vp * vp
    v2 = O.extendvel(nx, nz, nxb, nzb, v2)
    cfg = O.CpuCfg(8, nx, nz, nxb, nzb, nt, 10.0, 10.0, 0.001, 0.010)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_C)
    dobs = np.fromfile(os.path.join(d, "dobs.bin"), np.float32).reshape(1, nx, nt)
    return d, cfg, v2, srce, dobs


def test_mod_main_3lay_dobs_bit_exact(lay3):
    d, cfg, v2, srce, dobs = lay3
    data = O.mod_shot(cfg, v2, srce, 0 + cfg.nxb, 0 + cfg.nzb, 0 + cfg.nzb)
    assert_bit_equal(data, dobs[0])


def test_rtm_main_3lay_image_bit_exact(lay3):
    d, cfg, v2, srce, dobs = lay3
    im = O.rtm_shot(cfg, v2, srce, cfg.nxb, cfg.nzb, cfg.nzb, dobs, 0)
    assert_bit_equal(im, np.fromfile(os.path.join(d, "dir.img"), np.float32).reshape(151, 151))
    assert_bit_equal(im, np.fromfile(os.path.join(d, "dir.image"), np.float32).reshape(151, 151))


def test_gpu_family_forward_matches_shipped_snapshot(golden_dir):
    """cuda_reference_stencil_computation/input.bin is the reference's own
    forward wavefield (new_mod, shot 5, after 1700 steps).  It was
    produced by a different build of the same arithmetic, so the pin is a
    tolerance (recipe-noise floor ~6e-6), and it only holds with the truncated
    launch extents (quirk Q1) reproduced."""
    nx, nz, nxb, nzb, nt = 315, 195, 50, 50, 1700
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    ve = np.fromfile(os.path.join(golden_dir, "new_mod/vel_ext_rnd.shot5.bin"), np.float32).reshape(nxe, nze)
    g = np.fromfile(os.path.join(golden_dir, "stencil/input.bin"), np.float32).reshape(nxe, nze)
    srce = O.ricker_wavelet(nt, 0.001, 20.0, O.FAM_G)
    O.set_threads(min(8, O.max_threads()))
    try:
        cfg = O.GpuCfg(8, nxe, nze, nxb, nzb, nt, 10.0, 10.0, 0.001, 0.75, 1)
        P, PP = O.gpu_forward(cfg, (ve * ve).astype(np.float32), srce, 7 + 5 * 60 + nxb, nzb)
    finally:
        O.set_threads(1)
    rel = np.linalg.norm(P - g) / np.linalg.norm(g)
    assert rel < 2e-5, rel
    assert np.abs(P - g).max() < 1e-5
    assert not P[408:].any() and not P[:, 288:].any()  # Q1 evidence, same as the shipped file
    assert not g[408:].any() and not g[:, 288:].any()


# ------------------------------------------------------------------ tables
def test_tables_vs_reference_objects(refvec):
    for order in (2, 4, 6, 8, 10, 12, 16):
        assert_bit_equal(O.calc_coefs(order, O.FAM_C), refvec["coefs_cpu_%d" % order])
        assert_bit_equal(O.calc_coefs(order, O.FAM_G), refvec["coefs_gpu_%d" % order])
    for k, (nt, dt, fp) in enumerate(refvec["ricker_cases"]):
        assert_bit_equal(O.ricker_wavelet(int(nt), float(dt), float(fp), O.FAM_C), refvec["ricker_cpu_%d" % k])
        assert_bit_equal(O.ricker_wavelet(int(nt), float(dt), float(fp), O.FAM_G), refvec["ricker_gpu_%d" % k])
    for k, (nb, fac) in enumerate(refvec["taper_cases"]):
        assert_bit_equal(O.taper_table(int(nb), float(fac), O.FAM_G), refvec["taper_gpu_%d" % k])
        assert_bit_equal(O.taper_table(int(nb), float(fac), O.FAM_C), refvec["taper_cpu_%d" % k])


def test_velocity_extension_vs_reference_objects(refvec):
    nx, nz, nxb, nzb = (int(v) for v in refvec["ext_dims"])
    assert_bit_equal(O.extendvel(nx, nz, nxb, nzb, refvec["ext_in"]), refvec["extendvel_out"])
    assert_bit_equal(O.extendvel_linear(nx, nz, nxb, nzb, refvec["ext_in"], seed=1),
                     refvec["extendvel_linear_seed1_out"])


def test_cpu_family_functions_vs_reference_objects(refvec):
    nx, nz, nxb, nzb, _ = (int(v) for v in refvec["step_dims"])
    dx, dz, dt, fac = (float(v) for v in refvec["step_scal"])
    p, pp, v2 = refvec["step_p"], refvec["step_pp"], refvec["step_v2"]
    for order in (2, 4, 6, 8):
        a, b, lap = p.copy(), pp.copy(), np.zeros_like(p)
        for _ in range(3):
            O.fd_step(order, a, b, v2, lap, dx, dz, dt)
            a, b = b, a
        assert_bit_equal(a, refvec["fd_step3_o%d_p" % order])
        assert_bit_equal(b, refvec["fd_step3_o%d_pp" % order])
    tx, tz = O.taper_table(nxb, fac, O.FAM_C), O.taper_table(nzb, fac, O.FAM_C)
    a = p.copy(); O.taper_4(a, nx, nz, nxb, nzb, tx, tz)
    assert_bit_equal(a, refvec["taper_apply_out"])
    a = p.copy(); O.taper_top(a, nxb, tx, tz)
    assert_bit_equal(a, refvec["taper_apply2_out"])
    for k, (xs, zs) in enumerate(refvec["ptsrc_pos"]):
        a = p.copy(); O.ptsrc(int(xs), int(zs), np.float32(0.731), a)
        assert_bit_equal(a, refvec["ptsrc_out_%d" % k])


def test_gpu_launch_extents(refvec):
    assert tuple(refvec["gpu_launch_extents_415x295_nb50"]) == (408, 288, 48)


def test_image_laplacian_matches_the_fortran_expression():
    """oracle of the image post-filter (laplace.f90:24-28) against an independent float32 numpy restatement of the
    same left-to-right expression; ring of width 1 is zero.  (Parity unpinned: no gfortran here, no shipped output.)"""
    rng = np.random.default_rng(4)
    for nx, nz, dx, dz in ((151, 151, 10.0, 10.0), (37, 90, 25.0, 8.0), (3, 3, 1.0, 1.0), (2, 5, 10.0, 10.0)):
        img = rng.standard_normal((nx, nz)).astype(np.float32)
        got = O.image_laplacian(img, dx, dz)
        want = np.zeros_like(img)
        f = np.float32
        c = img[1:-1, 1:-1]
        tz = ((img[1:-1, 2:] - f(2.0) * c) + img[1:-1, :-2]) / (f(dz) * f(dz))
        tx = ((img[2:, 1:-1] - f(2.0) * c) + img[:-2, 1:-1]) / (f(dx) * f(dx))
        want[1:-1, 1:-1] = tz + tx
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
