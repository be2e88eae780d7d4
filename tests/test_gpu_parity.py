"""GPU parity tests (run with -m gpu on the B200 box): the real CUDA library
through the C ABI vs the oracle, the committed golden files, and -- at
BASELINE sizes the oracle cannot reach -- size-independent properties."""
import os

import numpy as np
import pytest

import parity_cases as PC
import parallel_finite_difference_computation_b200 as fdw
from oracle import oracle as O
from parallel_finite_difference_computation_b200 import (FAMILY_CPU, FAMILY_GPU, RECIPE_C, RECIPE_FAST, RECIPE_G,
                                                         SRC_GAUSS7, SRC_POINT, TAPER_FOUR, TAPER_NONE, TAPER_TOP,
                                                         Wave2D)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    L = fdw.load()
    assert L.fdw_device_count() >= 1, "no CUDA device: the GPU tests must run the CUDA library"
    return L


@pytest.mark.parametrize("order", [2, 4, 6, 8, 10, 12, 14, 16])
@pytest.mark.parametrize("shape", [(61, 47), (40, 64), (9, 9), (300, 130), (1030, 2052)])
def test_stencil(lib, order, shape):
    PC.case_stencil(lib, order, shape)


def test_stencil_golden_output_teste(lib, golden_dir):
    PC.case_stencil_golden(lib, golden_dir)


@pytest.fixture(params=["tile", "persistent", "one-launch", "rectangles+fork", "level-graph", "inplace-sponge"])
def launch_mode(request, monkeypatch):
    # tile: the default for small grids (shared-memory tiles, one cooperative launch per phase);
    # persistent: round 1's L2-resident kernel; the others: one (or several) launches per level, issued directly or
    # (level-graph: what mid-size grids get) as a replayed CUDA graph of two levels
    if request.param == "persistent":
        monkeypatch.setenv("FDW_TILE", "0")
    if request.param not in ("persistent", "tile"):
        monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")  # one launch per level
    if request.param not in ("one-launch", "persistent", "tile"):
        monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
        monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_LEVEL_GRAPH", "1" if request.param == "level-graph" else "0")
    # inplace-sponge: what mid-size whole grids get -- pending sponge passes applied in place, then ONE plain launch
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "1" if request.param == "inplace-sponge" else "0")
    return request.param


@pytest.mark.parametrize("order", [2, 4, 6, 8, 10, 12, 14, 16])
@pytest.mark.parametrize("family,recipe,taper,src", [
    (FAMILY_GPU, RECIPE_G, TAPER_TOP, SRC_POINT),
    (FAMILY_GPU, RECIPE_G, TAPER_NONE, SRC_POINT),
    (FAMILY_CPU, RECIPE_C, TAPER_TOP, SRC_POINT),
    (FAMILY_CPU, RECIPE_C, TAPER_FOUR, SRC_GAUSS7),
])
def test_advance_bit_exact(lib, launch_mode, order, family, recipe, taper, src):
    PC.case_advance(lib, family, recipe, taper, order=order, src_kind=src)


@pytest.mark.parametrize("family,recipe,taper", [(FAMILY_GPU, RECIPE_G, TAPER_TOP), (FAMILY_CPU, RECIPE_C, TAPER_FOUR),
                                                 (FAMILY_CPU, RECIPE_C, TAPER_TOP)])
def test_advance_nonzero_initial_fields(lib, family, recipe, taper):
    PC.case_advance(lib, family, recipe, taper, random_init=True, nt=7)


@pytest.mark.parametrize("dims", [(37, 29, 9, 8), (50, 43, 16, 16), (41, 35, 11, 13)])
def test_advance_compat_extents(lib, dims):
    nx, nz, nxb, nzb = dims
    PC.case_advance(lib, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=nx, nz=nz, nxb=nxb, nzb=nzb, compat=True)
    PC.case_advance(lib, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=nx, nz=nz, nxb=nxb, nzb=nzb, compat=True,
                    random_init=True, nt=5)


def test_advance_medium_grid_many_ctas(lib):
    # 1100 x 2100 extended grid: several z blocks, many x chunks, ragged edges
    PC.case_advance(lib, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=1060, nz=2060, nxb=20, nzb=20, nt=12)
    PC.case_advance(lib, FAMILY_CPU, RECIPE_C, TAPER_FOUR, nx=1060, nz=2061, nxb=20, nzb=20, nt=6,
                    src_kind=SRC_GAUSS7)


@pytest.mark.parametrize("taper", [TAPER_TOP, TAPER_NONE])
def test_advance_ragged_bulk_tail(lib, launch_mode, taper, monkeypatch):
    # > one 256-thread CTA column plus a short remainder: the bulk's ragged tail becomes its own rectangle
    monkeypatch.setenv("FDW_THREADS", "256")  # the tail is only cut for wide CTAs (default: 64 threads)
    PC.case_advance(lib, FAMILY_GPU, RECIPE_G, taper, nx=40, nz=1300, nxb=10, nzb=12, nt=5)
    PC.case_advance(lib, FAMILY_GPU, RECIPE_G, taper, nx=1200, nz=4136, nxb=20, nzb=20, nt=4)  # the C4 width


def test_fast_recipe_within_tolerance(lib):
    # FAST = symmetric pairs + FMA + float update; tolerance from SURVEY 8d: rel-L2 <= 5e-5
    PC.case_advance(lib, FAMILY_GPU, RECIPE_FAST, TAPER_TOP, nt=200, nx=120, nz=100, nxb=20, nzb=20, tol=5e-5)


@pytest.mark.parametrize("compat", [True, False])
@pytest.mark.parametrize("roundtrip", [True, False])
def test_gpu_family_rtm_shot(lib, launch_mode, compat, roundtrip):
    PC.case_gpu_rtm(lib, compat=compat, host_roundtrip=roundtrip)
    PC.case_gpu_rtm(lib, nx=150, nz=130, nxb=24, nzb=24, nt=300, compat=compat, host_roundtrip=roundtrip)


def test_mod_main_shot(lib, launch_mode):
    PC.case_mod_shot(lib)
    PC.case_mod_shot(lib, order=4, nx=30, nz=41, nxb=5, nzb=9)


@pytest.mark.parametrize("is_", [0, 1])
def test_rtm_main_shot(lib, launch_mode, is_):
    PC.case_rtm_shot_cpu(lib, is_=is_)


def test_level_loop_is_a_replayed_graph(lib, monkeypatch):
    """a grid above the tile / persistent range: pairs of levels go out as one replayed CUDA graph (arguments
    refreshed per pair), the odd last level directly; bit-exact like every other path (launch_mode level-graph)"""
    from parallel_finite_difference_computation_b200 import Wave2D
    monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
    monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_LEVEL_GRAPH", "1")
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "0")
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=lib) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        w.advance(0, 9)
        assert w.graph_replays() == 4
    monkeypatch.setenv("FDW_LEVEL_GRAPH", "0")
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=lib) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        w.advance(0, 9)
        assert w.graph_replays() == 0


@pytest.mark.parametrize("multirect", ["1", "0"])
def test_sponge_strips_share_one_launch(lib, multirect, monkeypatch):
    """a four-sided sponge on a grid that is not "small": bulk + 4 strips per level; the strips go out as ONE
    multi-rectangle launch (2 launches per level) unless FDW_MULTIRECT=0 (5 per level) -- same bits either way"""
    from parallel_finite_difference_computation_b200 import Wave2D
    monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
    monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_MULTIRECT", multirect)
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "0")
    PC.case_advance(lib, FAMILY_CPU, RECIPE_C, TAPER_FOUR, nx=70, nz=300, nxb=12, nzb=10, nt=6, src_kind=SRC_GAUSS7)
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=lib) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        n0 = w.launch_count()
        w.advance(0, 4)
        # level 0 has no sponge pass pending yet (CPU family: update, then sponge): one launch
        assert w.launch_count() - n0 == (1 + 3 * 2 if multirect == "1" else 1 + 3 * 5)


def test_sponge_in_place_then_one_plain_launch(lib, monkeypatch):
    """mid-size whole grids: the pending sponge passes are applied in place on the sponge regions (one element-wise
    launch for both levels), then ONE plain launch covers the grid -- 2 launches per level; same bits (the
    inplace-sponge launch mode runs every advance / shot case this way); receivers inside the sponge keep the strips"""
    from parallel_finite_difference_computation_b200 import Wave2D
    monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
    monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "1")
    monkeypatch.setenv("FDW_TILE", "0")
    PC.case_advance(lib, FAMILY_CPU, RECIPE_C, TAPER_FOUR, nx=70, nz=300, nxb=12, nzb=10, nt=6, src_kind=SRC_GAUSS7)
    PC.case_advance(lib, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=70, nz=300, nxb=12, nzb=10, nt=6, compat=True)
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=lib) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        n0 = w.launch_count()
        w.advance(0, 4)
        assert w.launch_count() - n0 == 1 + 3 * 2
    PC.case_mod_shot(lib, nx=40, nz=60, nxb=8, nzb=9)  # gz = nzb: receivers just outside the sponge -> in place
    PC.case_mod_shot(lib, nx=40, nz=60, nxb=8, nzb=9, gz=4)  # receivers inside the sponge -> strips, sample after the pass


@pytest.mark.parametrize("order", [10, 12, 16])
def test_shots_above_order_8(lib, order):
    """the reference's windowed-sinc weights (functions.c:119-157 makeo2) drive the same kernels at
    orders 10..16 -- whole shots of both families (forward + backward + imaging, modelling, RTM with history)"""
    PC.case_gpu_rtm(lib, compat=False, order=order)
    PC.case_gpu_rtm(lib, compat=True, host_roundtrip=True, order=order)
    PC.case_mod_shot(lib, order=order)
    PC.case_rtm_shot_cpu(lib, order=order)


def test_tile_kernel_runs_whole_phases(lib):
    """small grids: every phase is ONE launch of the shared-memory tile kernel (forward, the GPU family's
    backward pass with both field pairs, mod_main and rtm_main shots), bit-exact like every other path"""
    nx, nz, nb, nt = 150, 130, 24, 60
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    rng = np.random.default_rng(3)
    v2 = np.full((nxe, nze), np.float32(2500.0) ** 2, np.float32)
    dobs = rng.standard_normal((nx, nt)).astype(np.float32)
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=FAMILY_GPU, taper=TAPER_TOP,
                compat_extents=True, nt=nt) as w:
        w.set_v2(v2)
        w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 20.0, FAMILY_GPU))
        l0 = w.launch_count()
        w.forward(nb + 10, nb, download=False)
        assert w.tile_launches() == 1
        w.backward(dobs, nb)
        assert w.tile_launches() == 2
        assert w.launch_count() - l0 <= 6  # two phases + the materialisation of the two saved levels
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=FAMILY_CPU, taper=TAPER_TOP, nt=nt,
                history=True) as w:
        w.set_v2(v2)
        w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 20.0, FAMILY_CPU))
        w.rtm_shot_cpu(nb + 10, nb, nb, dobs[None], 0)
        assert w.tile_launches() == 2


@pytest.mark.parametrize("small", [True, False])
def test_shot_pipeline_device_stack_equals_sequential_loop(lib, small, monkeypatch):
    """fdw_v2_stage / commit, fdw_backward_device, fdw_stack_*: the shot loop without host round trips gives the
    reference's sequential img += imloc bit for bit (tile kernel and per-level launches)"""
    from parallel_finite_difference_computation_b200 import distributed as D
    if not small:
        monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    nx, nz, nb, nt, ns = 90, 70, 16, 80, 3
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    rng = np.random.default_rng(11)
    v2s = [np.full((nxe, nze), np.float32(2000.0 + 300.0 * k) ** 2, np.float32) for k in range(ns)]
    dobs = rng.standard_normal((ns, nx, nt)).astype(np.float32)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 25.0, FAMILY_GPU)
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=FAMILY_GPU, taper=TAPER_TOP, nt=nt) as w:
        w.set_wavelet(srce)
        seq = np.zeros((nx, nz), np.float32)
        for k in range(ns):
            w.set_v2(v2s[k])
            w.forward(nb + 5 + 20 * k, nb, download=False)
            seq += w.backward(dobs[k], nb)
    pipe = D.ShotPipeline(nx, nz, nb, nb, 10.0, 10.0, 0.001, nt=nt, order=8, fac=0.75)
    pipe.set_wavelet(srce)
    pinned = [pipe.pinned((nxe, nze)) for _ in range(ns)]
    for k in range(ns):
        pinned[k][:] = v2s[k]
    got = pipe.run_shots(list(range(ns)), lambda k: pinned[k], lambda k: dobs[k], lambda k: nb + 5 + 20 * k, nb, nb)
    pipe.close()
    PC.assert_bit_equal(got, seq, "shot pipeline image stack")
    assert np.abs(seq).max() > 0


@pytest.mark.parametrize("shape", [(151, 151), (37, 90), (3, 3), (2, 5), (1030, 517)])
def test_image_laplacian_bit_exact(lib, shape):
    """the image post-filter (laplace.f90:24-28) on the GPU vs the oracle"""
    rng = np.random.default_rng(9)
    img = rng.standard_normal(shape).astype(np.float32)
    PC.assert_bit_equal(fdw.image_laplacian(img, dx=25.0, dz=8.0), O.image_laplacian(img, 25.0, 8.0), "image laplacian")


# ------------------------------------------------------------------ reference golden files
def test_golden_3lay_mod_seismogram_and_image(lib, golden_dir):
    """mod_main + rtm_main on build/3lay_mod reproduce the reference's shipped
    dobs.bin / dir.img / dir.image bit for bit."""
    d = os.path.join(golden_dir, "3lay_mod")
    nx = nz = 151
    nb, nt = 40, 1001
    vp = np.fromfile(os.path.join(d, "3layer_151x151.bin"), np.float32).reshape(nx, nz)
    v2 = np.zeros((nx + 2 * nb, nz + 2 * nb), np.float32)
    v2[nb:nb + nx, nb:nb + nz] = vp * vp
    v2 = fdw.host.extendvel(nx, nz, nb, nb, v2)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 30.0, FAMILY_CPU)
    gold = np.fromfile(os.path.join(d, "dobs.bin"), np.float32).reshape(1, nx, nt)
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.010, family=FAMILY_CPU, taper=TAPER_FOUR,
                nt=nt) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        data = w.model_shot(nb, nb, nb)
    PC.assert_bit_equal(data, gold[0], "3lay_mod dobs.bin")
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.010, family=FAMILY_CPU, taper=TAPER_TOP, nt=nt,
                history=True) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        im = w.rtm_shot_cpu(nb, nb, nb, gold, 0)
    PC.assert_bit_equal(im, np.fromfile(os.path.join(d, "dir.img"), np.float32).reshape(nx, nz), "3lay_mod dir.img")
    PC.assert_bit_equal(im, np.fromfile(os.path.join(d, "dir.image"), np.float32).reshape(nx, nz), "dir.image")


def test_golden_new_mod_forward_snapshot(lib, golden_dir):
    """The reference's shipped input.bin is its own forward wavefield (new_mod,
    shot 5, 1700 steps).  Shipped file: tolerance (other build of the same
    arithmetic); oracle: bit-exact."""
    nx, nz, nb, nt = 315, 195, 50, 1700
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    ve = np.fromfile(os.path.join(golden_dir, "new_mod/vel_ext_rnd.shot5.bin"), np.float32).reshape(nxe, nze)
    g = np.fromfile(os.path.join(golden_dir, "stencil/input.bin"), np.float32).reshape(nxe, nze)
    v2 = (ve * ve).astype(np.float32)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 20.0, FAMILY_GPU)
    sx, sz = 7 + 5 * 60 + nb, nb
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=FAMILY_GPU, taper=TAPER_TOP,
                compat_extents=True, nt=nt) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        P, PP = w.forward(sx, sz)
    assert PC.rel_l2(P, g) < 2e-5 and np.abs(P - g).max() < 1e-5
    O.set_threads(min(8, O.max_threads()))
    try:
        oP, oPP = O.gpu_forward(O.GpuCfg(8, nxe, nze, nb, nb, nt, 10.0, 10.0, 0.001, 0.75, 1), v2, srce, sx, sz)
    finally:
        O.set_threads(1)
    PC.assert_bit_equal(P, oP, "new_mod forward P")
    PC.assert_bit_equal(PP, oPP, "new_mod forward PP")


# ------------------------------------------------------------------ BASELINE-size properties
def _big_setup(n, seed=0):
    rng = np.random.default_rng(seed)
    nb = 40
    nx = nz = n - 2 * nb
    ve = np.empty((n, n), np.float32)
    ve[:, : n // 3] = 2000.0
    ve[:, n // 3: 2 * n // 3] = 3000.0
    ve[:, 2 * n // 3:] = 4000.0
    v2 = ve * ve
    a = rng.standard_normal((n, n), dtype=np.float32)
    b = rng.standard_normal((n, n), dtype=np.float32)
    return nx, nz, nb, v2, a, b


@pytest.mark.parametrize("n", [4096, 16384])
def test_full_size_window_vs_oracle(lib, n):
    """16384^2 (config 3 size): k steps of the fused kernel; cut-outs of the
    result are checked bit for bit against the oracle run on the cut-out plus
    its domain of dependence (4 points per step) -- in the top sponge corner,
    at the source, at the bottom-right edge and in the bulk."""
    k = 6
    nx, nz, nb, v2, a, b = _big_setup(n)
    srce = fdw.host.ricker_wavelet(k, 0.001, 20.0, FAMILY_GPU) * np.float32(1e3)
    sx, sz = n // 2, nb
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=FAMILY_GPU, taper=TAPER_TOP,
                nt=k) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.set_source(sx, sz)
        newest, older = a.copy(), b.copy()
        w.propagate(newest, older, 0, k)
    m = 4 * k
    tx, tz = O.taper_table(nb, 0.75, O.FAM_G), O.taper_table(nb, 0.75, O.FAM_G)
    cx, cz = O.premult_coefs(8, 10.0, 10.0)
    _, _, dt2 = O.scalars(10.0, 10.0, 0.001)
    size = 96
    for (i0, j0) in [(0, 0), (sx - size // 2, 0), (n - size, n - size), (n // 3, n // 2), (n - size, 0)]:
        # window [i0,i0+size) x [j0,j0+size) needs data on the window grown by m (clipped at the grid edge)
        ia, ib = max(0, i0 - m), min(n, i0 + size + m)
        ja, jb = max(0, j0 - m), min(n, j0 + size + m)
        wa, wb = a[ia:ib, ja:jb].copy(), b[ia:ib, ja:jb].copy()
        wv = np.ascontiguousarray(v2[ia:ib, ja:jb])
        lap = np.zeros_like(wa)
        for it in range(k):
            for f in (wa, wb):  # sponge on load: global-position factors
                for jj in range(ja, min(jb, nb)):
                    f[:, jj - ja] *= tz[jj]
                    for ii in range(ia, ib):
                        if ii < nb:
                            f[ii - ia, jj - ja] *= tx[ii]
                        elif ii >= n - nb:
                            f[ii - ia, jj - ja] *= tx[n - 1 - ii]
            O.lap_G(8, wa, cx, cz, out=lap)
            # the global ring (first/last 4 rows/cols of the full grid) has lap = 0
            if ia == 0: lap[:4] = 0
            if ja == 0: lap[:, :4] = 0
            if ib == n: lap[-4:] = 0
            if jb == n: lap[:, -4:] = 0
            O.time_update(wa, wb, wv, lap, dt2)
            if ia <= sx < ib and ja <= sz < jb:
                wb[sx - ia, sz - ja] += srce[it]
            wa, wb = wb, wa
        sl = (slice(i0 - ia, i0 - ia + size), slice(j0 - ja, j0 - ja + size))
        PC.assert_bit_equal(newest[i0:i0 + size, j0:j0 + size], wa[sl], "window %d,%d newest" % (i0, j0))
        PC.assert_bit_equal(older[i0:i0 + size, j0:j0 + size], wb[sl], "window %d,%d older" % (i0, j0))


def test_full_size_translation_invariance(lib):
    """In a homogeneous medium, far from the edges, shifting the source by
    (dx,dz) grid points shifts the wavefield bit for bit."""
    n, nb, k = 8192, 40, 40
    nx = nz = n - 2 * nb
    v2 = np.full((n, n), np.float32(2500.0) ** 2, np.float32)
    srce = fdw.host.ricker_wavelet(k, 0.001, 25.0, FAMILY_GPU)
    out = []
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, family=FAMILY_GPU, taper=TAPER_NONE, nt=k) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        for (sx, sz) in [(3000, 3100), (3000 + 1237, 3100 + 2051)]:
            w.zero()
            w.set_source(sx, sz)
            w.advance(0, k)
            newest, _ = w.download()
            r = 4 * k + 8
            out.append(newest[sx - r:sx + r, sz - r:sz + r].copy())
            assert np.abs(out[-1]).max() > 0
            total = np.count_nonzero(newest)
            assert total == np.count_nonzero(out[-1])  # nothing outside the cone of influence
    PC.assert_bit_equal(out[0], out[1], "translated source")


def test_launch_geometry_invariance(lib, monkeypatch):
    """The result must not depend on how x is cut into CTA chunks or on the CTA width."""
    res = []
    for rpc, thr in [(None, None), ("8", "32"), ("37", "128"), ("512", "256")]:
        if rpc is None:
            monkeypatch.delenv("FDW_ROWS_PER_CTA", raising=False)
            monkeypatch.delenv("FDW_THREADS", raising=False)
        else:
            monkeypatch.setenv("FDW_ROWS_PER_CTA", rpc)
            monkeypatch.setenv("FDW_THREADS", thr)
        rng = np.random.default_rng(5)
        nx, nz, nb, nt = 700, 1500, 30, 9
        v2 = PC.layered_v2(nx, nz, nb, nb, rng)
        a = rng.standard_normal((nx + 2 * nb, nz + 2 * nb), dtype=np.float32)
        b = rng.standard_normal((nx + 2 * nb, nz + 2 * nb), dtype=np.float32)
        with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=FAMILY_GPU, taper=TAPER_TOP,
                    nt=nt) as w:
            w.set_v2(v2)
            w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 25.0, FAMILY_GPU))
            w.set_source(nb + 100, nb)
            w.propagate(a, b, 0, nt)
        res.append((a, b))
    for a, b in res[1:]:
        PC.assert_bit_equal(a, res[0][0], "geometry newest")
        PC.assert_bit_equal(b, res[0][1], "geometry older")
