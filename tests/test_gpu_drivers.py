"""GPU: the C drivers (apps/ -> bin/) behave like the reference programs --
same input.dat files in, same binary files out -- checked against the
reference's shipped golden files and against the reference's OWN CUDA programs
(oracle/_ref/rtm_code_ref, stencil_code_ref: built from /root/reference for
sm_100 with the reference flags) run side by side on the same B200."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import parity_cases as PC
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def run(cmd, cwd, **kw):
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=600, **kw)
    assert p.returncode == 0, "%s failed:\n%s\n%s" % (cmd, p.stdout[-2000:], p.stderr[-2000:])
    return p


@pytest.fixture(scope="module", autouse=True)
def built():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps")])


def test_stencil_code_golden(tmp_path, golden_dir):
    shutil.copy(os.path.join(golden_dir, "stencil/input.bin"), tmp_path / "input.bin")
    (tmp_path / "input.dat").write_text("tmpdir=./input.bin\nnz=195\nnx=315\ndz=10\ndx=10\nnxb=50\nnzb=50\norder=8\n")
    run([os.path.join(BIN, "stencil_code"), "./input.dat", "out.bin"], tmp_path)
    got = np.fromfile(tmp_path / "out.bin", np.float32).reshape(415, 295)
    gold = np.fromfile(os.path.join(golden_dir, "stencil/output_teste.bin"), np.float32).reshape(415, 295)
    PC.assert_bit_equal(got, gold, "stencil_code vs output_teste.bin")
    if R.path("stencil_code_ref"):  # the reference CUDA program, same box
        os.makedirs(tmp_path / "bin", exist_ok=True)
        os.makedirs(tmp_path / "run", exist_ok=True)
        shutil.copy(tmp_path / "input.bin", tmp_path / "run" / "input.bin")
        shutil.copy(tmp_path / "input.dat", tmp_path / "run" / "input.dat")
        run([R.path("stencil_code_ref"), "./input.dat"], tmp_path / "run")
        ref = np.fromfile(tmp_path / "bin" / "output_cuda.bin", np.float32).reshape(415, 295)
        # cudaMalloc'ed laplace ring is uninitialised in the reference (quirk Q2): compare the interior
        PC.assert_bit_equal(got[4:-4, 4:-4], ref[4:-4, 4:-4], "stencil_code vs reference CUDA program")


def test_mod_main_rtm_main_golden_3lay(tmp_path, golden_dir):
    d = os.path.join(golden_dir, "3lay_mod")
    shutil.copy(os.path.join(d, "3layer_151x151.bin"), tmp_path / "3layer_151x151.bin")
    par = ("tmpdir=./\nvpfile=3layer_151x151.bin\ndatfile=dobs.bin\nnz=151\nnx=151\nnt=1001\ndz=10\ndx=10\n"
           "dt=0.001\nfpeak=30.\nns=1\nsz=0\nfsx=0\nds=10\ngz=0\nnxb=40\nnzb=40\nfac=0.010\norder=8\n")
    (tmp_path / "input.dat").write_text(par)
    run([os.path.join(BIN, "mod_main"), "par=input.dat"], tmp_path)
    got = np.fromfile(tmp_path / "dobs.bin", np.float32)
    PC.assert_bit_equal(got, np.fromfile(os.path.join(d, "dobs.bin"), np.float32), "mod_main dobs.bin")
    run([os.path.join(BIN, "rtm_main"), "par=input.dat"], tmp_path)
    for f in ("dir.img", "dir.image"):
        PC.assert_bit_equal(np.fromfile(tmp_path / f, np.float32), np.fromfile(os.path.join(d, f), np.float32),
                            "rtm_main " + f)


def _write_rtm_case(dirpath, nx, nz, nb, nt, ns, seed):
    rng = np.random.default_rng(seed)
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    vp = np.empty((nx, nz), np.float32)
    vp[:, : nz // 2] = 2200.0
    vp[:, nz // 2:] = 3400.0
    vp.tofile(os.path.join(dirpath, "vp.bin"))
    ext = []
    for is_ in range(ns):
        ve = np.zeros((nxe, nze), np.float32)
        ve[nb:nb + nx, nb:nb + nz] = vp
        ext.append(O.extendvel_linear(nx, nz, nb, nb, ve, seed=100 + is_))
    np.stack(ext).tofile(os.path.join(dirpath, "vel_ext.bin"))
    dobs = (rng.standard_normal((ns, nx, nt)) * 0.1).astype(np.float32)
    dobs.tofile(os.path.join(dirpath, "dobs.bin"))
    os.makedirs(os.path.join(dirpath, "out"), exist_ok=True)
    # key order matters to the reference's substring parser (nz/nx before nzb/nxb)
    with open(os.path.join(dirpath, "input.dat"), "w") as f:
        f.write("tmpdir=./out\nvpfile=./vp.bin\ndatfile=./dobs.bin\nvel_ext_file=./vel_ext.bin\n"
                "nz=%d\nnx=%d\nnt=%d\ndz=10\ndx=10\ndt=0.001\nfpeak=25.\nns=%d\niss=0\nsz=0\nfsx=%d\nds=%d\ngz=0\n"
                "nxb=%d\nnzb=%d\nfac=0.75\norder=8\n" % (nz, nx, nt, ns, nx // 4, nx // 3, nb, nb))


@pytest.mark.parametrize("dims", [(101, 83, 24, 400, 2), (151, 151, 40, 300, 1)])
def test_rtm_code_vs_reference_cuda_program(tmp_path, dims):
    """our rtm_code and the reference's own rtm_code (rebuilt for sm_100 with
    --fmad=false) on identical inputs, same GPU.

    The reference's backward pass is NOT reproducible on a B200: its racy
    kernels (kernel_sism: 8 replicas of a non-atomic += per address, 2 warps;
    fd-code.cu:124-131) make two runs of the reference itself differ by
    rel-L2 ~1e-5..1e-4 (measured: tools/diag_ref_back.py).  Its
    forward pass is reproducible and equals the oracle bit for bit
    (tools/diag_ref_race.py).  So the image is compared with a tolerance of
    rel-L2 <= 1e-3 (north_star: "within a stated float32 relative-L2 and
    max-abs tolerance"), and we also require that our distance to
    the reference is of the order of the reference's distance to itself."""
    if not R.path("rtm_code_ref"):
        pytest.skip("oracle/_ref/rtm_code_ref not built")
    nx, nz, nb, nt, ns = dims
    ours, theirs = tmp_path / "ours", tmp_path / "ref"
    refs = [tmp_path / ("ref%d" % k) for k in range(3)]
    for d in [ours] + refs:
        os.makedirs(d)
        _write_rtm_case(str(d), nx, nz, nb, nt, ns, seed=3)
    run([os.path.join(BIN, "rtm_code"), "./input.dat"], ours)
    for d in refs:
        run([R.path("rtm_code_ref"), "./input.dat"], d)
    theirs = refs[0]
    a = np.fromfile(ours / "out" / "dir.image", np.float32).reshape(nx, nz)
    bs = [np.fromfile(d / "out" / "dir.image", np.float32).reshape(nx, nz) for d in refs]
    b = bs[0]
    assert np.abs(b).max() > 0, "reference image is empty"
    self_spread = max(PC.rel_l2(bs[i], bs[j]) for i in range(3) for j in range(i))
    ours_vs_ref = min(PC.rel_l2(a, x) for x in bs)
    print("reference vs itself (3 runs) rel-L2 %.3g | ours vs reference rel-L2 %.3g max-abs %.3g (|img|max %.3g)"
          % (self_spread, ours_vs_ref, min(np.abs(a - x).max() for x in bs), np.abs(b).max()))
    # stated tolerance: rel-L2 <= 5e-3 and max-abs <= 5e-3*|img|max, and never further from the
    # reference than 10x the reference is from itself (floor 1e-5)
    assert ours_vs_ref <= 5e-3
    assert min(np.abs(a - x).max() for x in bs) <= 5e-3 * np.abs(b).max()
    assert ours_vs_ref <= max(10 * self_spread, 1e-5)
    la, lb = open(ours / "image.num").read().splitlines(), open(theirs / "image.num").read().splitlines()
    assert len(la) == len(lb) == ns * (nx * nz + 1)
    assert [l for l in la if l.startswith("=")] == [l for l in lb if l.startswith("=")]
    assert os.path.getsize(ours / "out" / "dir.image_lap") == nx * nz * 4
    for f in ("dir.snaps", "dir.snaps_rec", "dir.snapr"):
        assert os.path.getsize(ours / "out" / f) == 0


def test_reference_forward_equals_oracle_and_library():
    """function-level: the reference's own CUDA fd_forward (libref_gpufam.so) on this GPU vs the
    oracle vs our library, from zero fields as in every reference call path.  When the reference
    reproduces itself (two runs bit-identical -- the usual case; its corner sponge is formally
    racy, quirk Q4) the three must agree bit for bit; otherwise within 10x its own spread."""
    if not R.available("libref_gpufam.so"):
        pytest.skip("oracle/_ref/libref_gpufam.so not built")
    import parallel_finite_difference_computation_b200 as fdw
    nx, nz, nb, nt = 101, 83, 24, 400
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    rng = np.random.default_rng(3)
    v2 = PC.layered_v2(nx, nz, nb, nb, rng, random_border=True)
    srce = O.ricker_wavelet(nt, 0.001, 25.0, O.FAM_G)
    sx, sz = nx // 4 + nb, nb
    g = R.GpuFam()
    g.fd_init(8, nxe, nze, nb, nb, nt, 1, 0.75, 10.0, 10.0, 0.001)
    runs = []
    for _ in range(2):
        rP, rPP = np.zeros((nxe, nze), np.float32), np.zeros((nxe, nze), np.float32)
        g.fd_forward(8, rP, rPP, v2, nt, 0, sz, [sx], srce)
        runs.append((rP, rPP))
    (rP, rPP), (rP2, rPP2) = runs
    assert np.abs(rPP).max() > 0
    oP, oPP = O.gpu_forward(O.GpuCfg(8, nxe, nze, nb, nb, nt, 10.0, 10.0, 0.001, 0.75, 1), v2, srce, sx, sz)
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU,
                    taper=fdw.TAPER_TOP, compat_extents=True, nt=nt) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        P, PP = w.forward(sx, sz)
    PC.assert_bit_equal(P, oP, "libfdwave vs oracle P")
    PC.assert_bit_equal(PP, oPP, "libfdwave vs oracle PP")
    if np.array_equal(rP, rP2) and np.array_equal(rPP, rPP2):
        PC.assert_bit_equal(oP, rP, "oracle vs reference CUDA fd_forward P")
        PC.assert_bit_equal(oPP, rPP, "oracle vs reference CUDA fd_forward PP")
    else:
        spread = PC.rel_l2(rPP2, rPP)
        print("reference fd_forward not reproducible in this run: self rel-L2 %.3g" % spread)
        assert PC.rel_l2(oPP, rPP) <= max(10 * spread, 1e-6)


def test_reference_mains_relinked_against_our_shim(tmp_path, golden_dir):
    """The truest drop-in test: the reference's UNMODIFIED mod_main.cpp and
    rtm_main.cpp, linked against libfdwave_cpufam.so instead of the reference's fd.c /
    taper.c / ptsrc.c (oracle/Makefile target `relinked`), must reproduce the shipped
    3lay_mod golden files bit for bit with fd_step running on the GPU."""
    if not R.path("mod_main_relinked"):
        pytest.skip("oracle/_ref/mod_main_relinked not built")
    d = os.path.join(golden_dir, "3lay_mod")
    shutil.copy(os.path.join(d, "3layer_151x151.bin"), tmp_path / "3layer_151x151.bin")
    (tmp_path / "input.dat").write_text(
        "tmpdir=./\nvpfile=3layer_151x151.bin\ndatfile=dobs.bin\nnz=151\nnx=151\nnt=1001\ndz=10\ndx=10\n"
        "dt=0.001\nfpeak=30.\nns=1\nsz=0\nfsx=0\nds=10\ngz=0\nnxb=40\nnzb=40\nfac=0.010\norder=8\n")
    run([R.path("mod_main_relinked"), "par=input.dat"], tmp_path)
    PC.assert_bit_equal(np.fromfile(tmp_path / "dobs.bin", np.float32),
                        np.fromfile(os.path.join(d, "dobs.bin"), np.float32), "relinked mod_main dobs.bin")
    run([R.path("rtm_main_relinked"), "par=input.dat"], tmp_path)
    for f in ("dir.img", "dir.image"):
        PC.assert_bit_equal(np.fromfile(tmp_path / f, np.float32), np.fromfile(os.path.join(d, f), np.float32),
                            "relinked rtm_main " + f)


@pytest.mark.parametrize("dims", [(101, 83, 24, 400, 2), (151, 151, 40, 300, 1)])
def test_reference_rtm_main_relinked_against_gpufam_shim(tmp_path, dims):
    """Function-level boundary of the GPU family: the reference's OWN main() of fd-code.cu (cut out of
    the file at build time, oracle/Makefile target rtm_code_relinked) linked against libfdwave_gpufam.so -- i.e.
    the reference program with its kernels, fd_init, fd_forward and fd_back replaced by this library -- must write
    the same dir.image as our own driver bin/rtm_code bit for bit, and match the reference's CUDA program within
    the tolerance of test_rtm_code_vs_reference_cuda_program."""
    if not R.path("rtm_code_relinked"):
        pytest.skip("oracle/_ref/rtm_code_relinked not built")
    nx, nz, nb, nt, ns = dims
    ours, relinked, ref = tmp_path / "ours", tmp_path / "relinked", tmp_path / "ref"
    for d in (ours, relinked, ref):
        os.makedirs(d)
        _write_rtm_case(str(d), nx, nz, nb, nt, ns, seed=5)
    run([os.path.join(BIN, "rtm_code"), "./input.dat"], ours)
    run([R.path("rtm_code_relinked"), "./input.dat"], relinked)
    a = np.fromfile(ours / "out" / "dir.image", np.float32).reshape(nx, nz)
    b = np.fromfile(relinked / "out" / "dir.image", np.float32).reshape(nx, nz)
    assert np.abs(a).max() > 0
    PC.assert_bit_equal(b, a, "reference main() on libfdwave_gpufam vs bin/rtm_code dir.image")
    assert open(ours / "image.num").read() == open(relinked / "image.num").read()
    if R.path("rtm_code_ref"):
        run([R.path("rtm_code_ref"), "./input.dat"], ref)
        c = np.fromfile(ref / "out" / "dir.image", np.float32).reshape(nx, nz)
        print("relinked vs reference CUDA program: rel-L2 %.3g" % PC.rel_l2(b, c))
        assert PC.rel_l2(b, c) <= 5e-3


def test_rtm_code_image_lap_opt_in(tmp_path):
    """dir.image_lap: zeros by default like the reference (fd-code.cu:477,542); FDW_IMAGE_LAP=1 fills it with the
    laplace.f90 filter of dir.image"""
    nx, nz, nb, nt, ns = 61, 47, 16, 120, 1
    d = tmp_path / "case"
    os.makedirs(d)
    _write_rtm_case(str(d), nx, nz, nb, nt, ns, seed=8)
    run([os.path.join(BIN, "rtm_code"), "./input.dat"], d)
    assert not np.fromfile(d / "out" / "dir.image_lap", np.float32).any()
    run([os.path.join(BIN, "rtm_code"), "./input.dat"], d, env=dict(os.environ, FDW_IMAGE_LAP="1"))
    img = np.fromfile(d / "out" / "dir.image", np.float32).reshape(nx, nz)
    lap = np.fromfile(d / "out" / "dir.image_lap", np.float32).reshape(nx, nz)
    assert np.abs(img).max() > 0
    PC.assert_bit_equal(lap, O.image_laplacian(img, 10.0, 10.0), "dir.image_lap")
