"""CPU-only: libfdwave's host logic and kernel bodies (host build against the
fake CUDA runtime in tests/emu/) vs the oracle.  The GPU tests run the same
cases through the real library (test_gpu_parity.py)."""
import numpy as np
import pytest

import parity_cases as PC
from emu_loader import load as load_emu
from parallel_finite_difference_computation_b200 import (FAMILY_CPU, FAMILY_GPU, RECIPE_C, RECIPE_FAST, RECIPE_G,
                                                         SRC_GAUSS7, SRC_POINT, TAPER_FOUR, TAPER_NONE, TAPER_TOP)


@pytest.fixture(scope="module")
def emu():
    return load_emu()


@pytest.fixture(params=["persistent", "one-launch", "rectangles", "rectangles+fork", "level-graph", "inplace-sponge"])
def launch_mode(request, monkeypatch):
    """the library picks one sponge-kernel launch for tiny grids and plain/sponge rectangles
    (optionally forked to a side stream) for large ones, issued directly or -- mid-size grids -- as a replayed CUDA
    graph of two levels; force each path on the small test grids"""
    if request.param != "persistent":
        monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")  # one launch per level
    if request.param not in ("one-launch", "persistent"):
        monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
        monkeypatch.setenv("FDW_FORK_LIMIT", "0" if request.param != "rectangles" else str(1 << 40))
    monkeypatch.setenv("FDW_LEVEL_GRAPH", "1" if request.param == "level-graph" else "0")
    # inplace-sponge: what mid-size whole grids get -- pending sponge passes applied in place, then ONE plain launch
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "1" if request.param == "inplace-sponge" else "0")
    return request.param


@pytest.mark.parametrize("order", [2, 4, 6, 8, 10, 12, 14, 16])
@pytest.mark.parametrize("shape", [(61, 47), (40, 64), (9, 9), (300, 130)])
def test_stencil(emu, order, shape):
    PC.case_stencil(emu, order, shape)


def test_stencil_golden(emu, golden_dir):
    PC.case_stencil_golden(emu, golden_dir)


@pytest.mark.parametrize("order", [2, 4, 6, 8, 10, 12, 14, 16])
@pytest.mark.parametrize("family,recipe,taper,src", [
    (FAMILY_GPU, RECIPE_G, TAPER_TOP, SRC_POINT),
    (FAMILY_GPU, RECIPE_G, TAPER_NONE, SRC_POINT),
    (FAMILY_CPU, RECIPE_C, TAPER_TOP, SRC_POINT),
    (FAMILY_CPU, RECIPE_C, TAPER_FOUR, SRC_GAUSS7),
])
def test_advance_bit_exact(emu, launch_mode, order, family, recipe, taper, src):
    PC.case_advance(emu, family, recipe, taper, order=order, src_kind=src)


@pytest.mark.parametrize("family,recipe,taper", [(FAMILY_GPU, RECIPE_G, TAPER_TOP), (FAMILY_CPU, RECIPE_C, TAPER_FOUR),
                                                 (FAMILY_CPU, RECIPE_C, TAPER_TOP)])
def test_advance_nonzero_initial_fields(emu, family, recipe, taper):
    PC.case_advance(emu, family, recipe, taper, random_init=True, nt=7)


@pytest.mark.parametrize("dims", [(37, 29, 9, 8), (50, 43, 16, 16), (41, 35, 11, 13)])
def test_advance_compat_extents(emu, launch_mode, dims):
    nx, nz, nxb, nzb = dims
    PC.case_advance(emu, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=nx, nz=nz, nxb=nxb, nzb=nzb, compat=True)
    PC.case_advance(emu, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=nx, nz=nz, nxb=nxb, nzb=nzb, compat=True,
                    random_init=True, nt=5)


def test_advance_wide_grid_many_chunks(emu, launch_mode):
    PC.case_advance(emu, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=90, nz=1100, nxb=10, nzb=12, nt=6)


@pytest.mark.parametrize("taper", [TAPER_TOP, TAPER_NONE])
def test_advance_ragged_bulk_tail(emu, launch_mode, taper, monkeypatch):
    """more than one 256-thread CTA column plus a short remainder: the ragged tail of the bulk
    rectangle is launched as its own narrow rectangle (side stream when forking)"""
    monkeypatch.setenv("FDW_THREADS", "256")  # the tail is only cut for wide CTAs (default: 64 threads)
    PC.case_advance(emu, FAMILY_GPU, RECIPE_G, taper, nx=40, nz=1300, nxb=10, nzb=12, nt=5)


def test_fast_recipe_within_tolerance(emu):
    PC.case_advance(emu, FAMILY_GPU, RECIPE_FAST, TAPER_TOP, nt=60, tol=5e-5)


@pytest.mark.parametrize("compat", [True, False])
@pytest.mark.parametrize("roundtrip", [True, False])
def test_gpu_family_rtm_shot(emu, launch_mode, compat, roundtrip):
    PC.case_gpu_rtm(emu, compat=compat, host_roundtrip=roundtrip)


def test_mod_main_shot(emu, launch_mode):
    PC.case_mod_shot(emu)
    PC.case_mod_shot(emu, order=4, nx=30, nz=41, nxb=5, nzb=9)


@pytest.mark.parametrize("is_", [0, 1])
def test_rtm_main_shot(emu, launch_mode, is_):
    PC.case_rtm_shot_cpu(emu, is_=is_)


@pytest.mark.parametrize("threads", ["32", "64", "96", "256"])
@pytest.mark.parametrize("multirect", ["1", "0"])
def test_strip_folding_at_every_cta_width(emu, threads, multirect, monkeypatch):
    """the sponge strips' multi-rectangle launch folds narrow strips to the lane group that pads them least, for
    whatever CTA width is in force (FDW_THREADS; 96 = a width that is not a power of two), orders 8 and 12"""
    monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
    monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "0")
    monkeypatch.setenv("FDW_THREADS", threads)
    monkeypatch.setenv("FDW_MULTIRECT", multirect)
    PC.case_advance(emu, FAMILY_CPU, RECIPE_C, TAPER_FOUR, nx=70, nz=300, nxb=12, nzb=10, nt=5, src_kind=SRC_GAUSS7)
    PC.case_advance(emu, FAMILY_CPU, RECIPE_C, TAPER_FOUR, order=12, nx=50, nz=200, nxb=12, nzb=10, nt=4)


def test_level_loop_is_a_replayed_graph(emu, monkeypatch):
    """a grid above the tile / persistent range: pairs of levels go out as one replayed CUDA graph (arguments
    refreshed per pair), the odd last level directly; bit-exact like every other path (launch_mode level-graph)"""
    from parallel_finite_difference_computation_b200 import Wave2D
    monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
    monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_LEVEL_GRAPH", "1")
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "0")
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=emu) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        w.advance(0, 9)
        assert w.graph_replays() == 4
    monkeypatch.setenv("FDW_LEVEL_GRAPH", "0")
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=emu) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        w.advance(0, 9)
        assert w.graph_replays() == 0


@pytest.mark.parametrize("multirect", ["1", "0"])
def test_sponge_strips_share_one_launch(emu, multirect, monkeypatch):
    """a four-sided sponge on a grid that is not "small": bulk + 4 strips per level; the strips go out as ONE
    multi-rectangle launch (2 launches per level) unless FDW_MULTIRECT=0 (5 per level) -- same bits either way"""
    from parallel_finite_difference_computation_b200 import Wave2D
    monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
    monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_MULTIRECT", multirect)
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "0")
    PC.case_advance(emu, FAMILY_CPU, RECIPE_C, TAPER_FOUR, nx=70, nz=300, nxb=12, nzb=10, nt=6, src_kind=SRC_GAUSS7)
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=emu) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        n0 = w.launch_count()
        w.advance(0, 4)
        # level 0 has no sponge pass pending yet (CPU family: update, then sponge): one launch
        assert w.launch_count() - n0 == (1 + 3 * 2 if multirect == "1" else 1 + 3 * 5)


def test_sponge_in_place_then_one_plain_launch(emu, monkeypatch):
    """mid-size whole grids: the pending sponge passes are applied in place on the sponge regions (one element-wise
    launch for both levels), then ONE plain launch covers the grid -- 2 launches per level; same bits (the
    inplace-sponge launch mode runs every advance / shot case this way); receivers inside the sponge keep the strips"""
    from parallel_finite_difference_computation_b200 import Wave2D
    monkeypatch.setenv("FDW_PERSIST_LIMIT", "0")
    monkeypatch.setenv("FDW_SMALL_GRID_LIMIT", "0")
    monkeypatch.setenv("FDW_FORK_LIMIT", "0")
    monkeypatch.setenv("FDW_SPONGE_INPLACE", "1")
    monkeypatch.setenv("FDW_TILE", "0")
    PC.case_advance(emu, FAMILY_CPU, RECIPE_C, TAPER_FOUR, nx=70, nz=300, nxb=12, nzb=10, nt=6, src_kind=SRC_GAUSS7)
    PC.case_advance(emu, FAMILY_GPU, RECIPE_G, TAPER_TOP, nx=70, nz=300, nxb=12, nzb=10, nt=6, compat=True)
    with Wave2D(70, 300, 12, 10, 10.0, 10.0, 0.001, family=FAMILY_CPU, recipe=RECIPE_C, taper=TAPER_FOUR, lib=emu) as w:
        w.set_v2(np.full((94, 320), 4.0e6, np.float32))
        w.zero()
        n0 = w.launch_count()
        w.advance(0, 4)
        assert w.launch_count() - n0 == 1 + 3 * 2
    PC.case_mod_shot(emu, nx=40, nz=60, nxb=8, nzb=9)  # gz = nzb: receivers just outside the sponge -> in place
    PC.case_mod_shot(emu, nx=40, nz=60, nxb=8, nzb=9, gz=4)  # receivers inside the sponge -> strips, sample after the pass


@pytest.mark.parametrize("order", [10, 12, 16])
def test_shots_above_order_8(emu, order):
    """the reference's windowed-sinc weights (functions.c:119-157 makeo2) drive the same kernels at
    orders 10..16 -- whole shots of both families (forward + backward + imaging, modelling, RTM with history)"""
    PC.case_gpu_rtm(emu, compat=False, order=order)
    PC.case_gpu_rtm(emu, compat=True, host_roundtrip=True, order=order)
    PC.case_mod_shot(emu, order=order)
    PC.case_rtm_shot_cpu(emu, order=order)
