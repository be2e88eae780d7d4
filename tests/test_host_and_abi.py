"""CPU-only: the product's host tables/readers vs the reference's own outputs
(tests/golden/ref_vectors.npz), and the C-ABI surface of libfdwave.so."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import parallel_finite_difference_computation_b200 as fdw
from parallel_finite_difference_computation_b200 import _lib, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bit_eq(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_library_loads_and_exports_every_declared_symbol():
    lib = fdw.load()  # raises if libfdwave.so is missing: there is no fallback
    hdr = open(os.path.join(ROOT, "include", "fdwave.h")).read()
    declared = set(re.findall(r"\b(fdw_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback_without_device():
    lib = fdw.load()
    if lib.fdw_device_count() > 0:
        pytest.skip("a CUDA device is visible here")
    with pytest.raises(fdw.FdwError) as e:
        fdw.Wave2D(16, 16, 4, 4, 10.0, 10.0, 0.001)
    assert e.value.code == -2  # FDW_ERR_CUDA
    with pytest.raises(fdw.FdwError):
        fdw.stencil(np.zeros((16, 16), np.float32))


def test_tables_bit_exact_vs_reference(refvec):
    for order in (2, 4, 6, 8, 10, 12, 16):
        assert bit_eq(host.calc_coefs(order, fdw.FAMILY_CPU), refvec["coefs_cpu_%d" % order])
        assert bit_eq(host.calc_coefs(order, fdw.FAMILY_GPU), refvec["coefs_gpu_%d" % order])
    for k, (nt, dt, fp) in enumerate(refvec["ricker_cases"]):
        assert bit_eq(host.ricker_wavelet(int(nt), float(dt), float(fp), fdw.FAMILY_CPU), refvec["ricker_cpu_%d" % k])
        assert bit_eq(host.ricker_wavelet(int(nt), float(dt), float(fp), fdw.FAMILY_GPU), refvec["ricker_gpu_%d" % k])
    for k, (nb, fac) in enumerate(refvec["taper_cases"]):
        assert bit_eq(host.taper_table(int(nb), float(fac), fdw.FAMILY_GPU), refvec["taper_gpu_%d" % k])
        assert bit_eq(host.taper_table(int(nb), float(fac), fdw.FAMILY_CPU), refvec["taper_cpu_%d" % k])


def test_velocity_extension_bit_exact_vs_reference(refvec):
    nx, nz, nxb, nzb = (int(v) for v in refvec["ext_dims"])
    assert bit_eq(host.extendvel(nx, nz, nxb, nzb, refvec["ext_in"]), refvec["extendvel_out"])
    assert bit_eq(host.extendvel_linear(nx, nz, nxb, nzb, refvec["ext_in"], seed=1),
                  refvec["extendvel_linear_seed1_out"])


def test_ptsrc_weights_match_reference_ptsrc(refvec):
    # ptsrc adds ts*w to a field; recover w from the reference's output around (20,15)
    p, out = refvec["step_p"], refvec["ptsrc_out_0"]
    w = host.ptsrc_weights()
    ts = np.float32(0.731)
    want = p.copy()
    want[17:24, 12:19] += ts * w
    assert bit_eq(want, out)


GPU_DAT = """tmpdir=./output
vpfile=./models/new_mod/vel-koslov.1
datfile=./models/new_mod/dobs.6
vel_ext_file=./models/new_mod/vel_ext_rnd.6
nz=195
nx=315
nt=1700
dz=10
dx=10
dt=0.001
fpeak=20.
ns=6
iss=0
sz=0
fsx=7
ds=60
gz=0
nxb=50
nzb=50
rnd=1
fac=0.75
order=8
"""


def test_gpu_dialect_reader(tmp_path):
    f = tmp_path / "input.dat"
    f.write_text(GPU_DAT)
    d = host.read_input_gpu(str(f))
    assert (d["nz"], d["nx"], d["nt"], d["ns"], d["fsx"], d["ds"], d["nxb"], d["nzb"], d["order"]) == \
        (195, 315, 1700, 6, 7, 60, 50, 50, 8)
    assert d["tmpdir"] == "./output" and d["vel_ext_file"].endswith("vel_ext_rnd.6") and d["has_vel_ext_file"] == 1
    assert abs(d["fac"] - 0.75) < 1e-7 and abs(d["fpeak"] - 20.0) < 1e-6 and abs(d["dt"] - 0.001) < 1e-9
    # substring matching of the reference parser (functions.c:19): "rnd" first matches the
    # vel_ext_file line, whose value atoi()s to 0
    assert d["rnd"] == 0
    # order sensitivity: nzb before nz makes "nz" read the nzb line
    f.write_text("nzb=50\nnz=195\nnx=10\n")
    d = host.read_input_gpu(str(f))
    assert d["nz"] == 50 and d["nzb"] == 50
    # defaults (fd-code.cu:367-377)
    assert (d["ns"], d["sz"], d["fsx"], d["ds"], d["gz"], d["order"], d["nxb"]) == (1, 0, 0, 1, 0, 8, 40)
    assert abs(d["fac"] - 0.7) < 1e-7
    d = host.read_input_gpu(str(f), apply_defaults=False)
    assert d["ns"] == -1 and d["fac"] == -1.0 and d["has_datfile"] == 0
    with pytest.raises(fdw.FdwError):
        host.read_input_gpu(str(tmp_path / "missing.dat"))


def test_stencil_dialect_reader(tmp_path):
    f = tmp_path / "input.dat"
    f.write_text("tmpdir=./input.bin\nnz=195\nnx=315\ndz=10\ndx=10\nnxb=50\nnzb=50\norder=8\n")
    d = host.read_input_stencil(str(f))
    assert d["tmpdir"] == "./input.bin"
    assert (d["nz"], d["nx"], d["nxb"], d["nzb"], d["order"]) == (195, 315, 50, 50, 8)
    assert d["dz"] == 10.0 and d["dx"] == 10.0


def test_cpu_dialect_reader(tmp_path):
    f = tmp_path / "par.dat"
    f.write_text("tmpdir=./ vpfile=3layer_151x151.bin\ndatfile=dobs.bin\nnz=151\nnx=151 nt=1001\n"
                 "dz=10\ndx=10\ndt=0.001\nfpeak=30.\nns=1\nnz=152\nfac=0.010\n")
    d = host.read_input_cpu(str(f))
    assert d["nz"] == 152  # last occurrence wins (getpars.c:447-453)
    assert (d["nx"], d["nt"], d["ns"], d["order"], d["nxb"], d["nzb"]) == (151, 1001, 1, 8, 40, 40)
    assert d["vpfile"] == "3layer_151x151.bin" and d["datfile"] == "dobs.bin" and d["has_datfile"] == 1
    assert abs(d["fac"] - 0.010) < 1e-9


def test_float_files_and_image_num(tmp_path):
    """the raw float32 readers / writers and the image.num text dump of the drop-in surface live in the library
    (SURVEY 8f.2): short reads leave the rest of the buffer alone, append mode, the reference's text format
    (fd-code.cu:521-528: "======== is ========" then " %f " per point, iz outer / ix inner, img += imloc)"""
    lib = fdw.load()
    p = str(tmp_path / "a.bin").encode()
    a = np.arange(7, dtype=np.float32)
    assert lib.fdw_write_floats(p, a.ctypes.data, 7, 0) == 0
    assert lib.fdw_write_floats(p, a.ctypes.data, 3, 1) == 0
    buf = np.full(12, -1.0, np.float32)
    assert lib.fdw_read_floats(p, buf.ctypes.data, 12) == 10
    assert np.array_equal(buf, np.concatenate([a, a[:3], [-1.0, -1.0]]).astype(np.float32))
    assert lib.fdw_read_floats(str(tmp_path / "missing.bin").encode(), buf.ctypes.data, 12) == -1
    assert lib.fdw_write_floats(str(tmp_path / "no_dir" / "x.bin").encode(), a.ctypes.data, 7, 0) == -4
    nx, nz = 3, 2
    img = np.zeros((nx, nz), np.float32)
    num = str(tmp_path / "image.num").encode()
    want = ""
    acc = np.zeros((nx, nz), np.float32)
    for is_ in range(2):
        imloc = (np.arange(nx * nz, dtype=np.float32).reshape(nx, nz) + 0.5) * (is_ + 1)
        assert lib.fdw_image_stack_shot(num, is_, nx, nz, img.ctypes.data, imloc.ctypes.data) == 0
        acc += imloc
        want += "======== %i ========\n" % is_
        for iz in range(nz):
            for ix in range(nx):
                want += " %f \n" % acc[ix, iz]
    assert np.array_equal(img, acc)
    assert open(num.decode()).read() == want
