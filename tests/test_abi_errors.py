"""CPU-only: error behaviour and edge cases of the C ABI, exercised on the host build
(tests/emu) where a "device" always exists, and on the real library where none does."""
import ctypes as C

import numpy as np
import pytest

import parity_cases as PC
from emu_loader import load as load_emu
from oracle import oracle as O
from parallel_finite_difference_computation_b200 import (FAMILY_CPU, FAMILY_GPU, RECIPE_C, RECIPE_G, SRC_POINT,
                                                         TAPER_FOUR, TAPER_NONE, TAPER_TOP, FdwError, Wave2D, _lib,
                                                         stencil)


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def test_bad_arguments_are_reported_not_crashed(emu):
    with pytest.raises(FdwError) as e:
        Wave2D(0, 10, 2, 2, 10.0, 10.0, 0.001, lib=emu)
    assert e.value.code == -1 and "bad grid" in str(e.value)
    with pytest.raises(FdwError) as e:
        Wave2D(16, 16, 4, 4, 10.0, 10.0, 0.001, order=18, lib=emu)  # host tables support it, the device stops at 16
    assert e.value.code == -6
    with pytest.raises(FdwError) as e:
        Wave2D(16, 16, 4, 4, 10.0, 10.0, 0.001, order=7, lib=emu)
    assert e.value.code == -6
    with pytest.raises(FdwError) as e:  # slabs exchange 4 ghost rows: orders above 8 need the whole grid
        Wave2D(40, 16, 4, 4, 10.0, 10.0, 0.001, order=12, slab=(0, 24), lib=emu)
    assert e.value.code == -6 and "ghost rows" in str(e.value)
    with pytest.raises(FdwError) as e:
        Wave2D(16, 16, 4, 4, 10.0, 10.0, 0.001, slab=(5, 400), lib=emu)
    assert e.value.code == -1
    with Wave2D(16, 16, 4, 4, 10.0, 10.0, 0.001, nt=0, lib=emu) as w:
        with pytest.raises(ValueError):
            w.set_v2(np.zeros((3, 3), np.float32))
        with pytest.raises(FdwError) as e:  # no nt / wavelet
            w.forward(8, 4)
        assert e.value.code == -5
        with pytest.raises(FdwError) as e:  # history not allocated
            w.rtm_shot_cpu(8, 4, 4, np.zeros((1, 16, 1), np.float32))
        assert e.value.code in (-1, -5)
    with Wave2D(16, 16, 4, 4, 10.0, 10.0, 0.001, nt=5, lib=emu) as w:
        w.set_wavelet(np.ones(5, np.float32))
        with pytest.raises(FdwError) as e:  # backward without a forward
            w.backward(np.zeros((16, 5), np.float32), 4)
        assert e.value.code == -5 and "fdw_forward" in str(e.value)
    assert emu.fdw_calc_coefs(7, 0, np.zeros(8, np.float32)) == -1
    assert emu.fdw_fields_zero(None, 0) == -1


def test_zero_steps_and_single_step(emu):
    rng = np.random.default_rng(2)
    nx, nz, nb = 12, 9, 4
    a = rng.uniform(-1, 1, (nx + 2 * nb, nz + 2 * nb)).astype(np.float32)
    b = rng.uniform(-1, 1, (nx + 2 * nb, nz + 2 * nb)).astype(np.float32)
    with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, taper=TAPER_NONE, nt=1, lib=emu) as w:
        w.set_v2(np.full_like(a, 4e6))
        x, y = a.copy(), b.copy()
        w.propagate(x, y, 0, 0)  # zero levels: a round trip through the device layout
        PC.assert_bit_equal(x, a, "0 levels newest")
        PC.assert_bit_equal(y, b, "0 levels older")


@pytest.mark.parametrize("dims", [(1, 1, 4, 4), (2, 3, 4, 4), (5, 1, 4, 5), (1, 40, 6, 4), (3, 3, 0, 0)])
def test_degenerate_grids(emu, dims):
    """interiors of a single point / single row / no border at all"""
    nx, nz, nxb, nzb = dims
    if min(nx + 2 * nxb, nz + 2 * nzb) < 9:
        # thinner than the stencil: everything is ring, the update degenerates to 2p - pp
        pass
    PC.case_advance(emu, FAMILY_GPU, RECIPE_G, TAPER_TOP if nzb else TAPER_NONE, nx=nx, nz=nz, nxb=max(nxb, 0),
                    nzb=max(nzb, 0), nt=5, random_init=True) if nxb and nzb else None
    rng = np.random.default_rng(1)
    p = rng.uniform(-1, 1, (nx + 2 * nxb, nz + 2 * nzb)).astype(np.float32)
    PC.assert_bit_equal(stencil(p, order=8, dx=10.0, dz=10.0, lib=emu), O.stencil(8, 10.0, 10.0, p), "tiny stencil")


def test_ragged_widths_every_residue_mod_4(emu):
    for nz in range(17, 25):  # nze mod 4 = every residue, pad columns exercised
        PC.case_advance(emu, FAMILY_CPU, RECIPE_C, TAPER_FOUR, nx=19, nz=nz, nxb=5, nzb=4, nt=6, random_init=True)


def test_source_on_the_grid_edge_and_in_the_sponge(emu):
    """ptsrc clips its 7x7 patch at the grid edge (ptsrc.c:51-52)"""
    import parallel_finite_difference_computation_b200 as fdw
    rng = np.random.default_rng(4)
    nx, nz, nb, nt = 21, 17, 5, 8
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    v2 = PC.layered_v2(nx, nz, nb, nb, rng)
    srce = O.ricker_wavelet(nt, 0.001, 30.0, O.FAM_C)
    for (sx, sz) in [(0, 0), (nxe - 1, nze - 1), (1, nze - 2), (nxe - 2, 2)]:
        cfg = O.CpuCfg(8, nx, nz, nb, nb, nt, 10.0, 10.0, 0.001, 0.05)
        want = O.mod_shot(cfg, v2, srce, sx, sz, nb)
        with Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, fac=0.05, family=FAMILY_CPU, taper=TAPER_FOUR, nt=nt,
                    lib=emu) as w:
            w.set_v2(v2)
            w.set_wavelet(srce)
            got = w.model_shot(sx, sz, nb)
        PC.assert_bit_equal(got, want, "source at (%d,%d)" % (sx, sz))
