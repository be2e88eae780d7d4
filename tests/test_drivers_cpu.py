"""CPU-only: the C drivers exist, link against the C ABI only, and refuse to
run without a GPU (no CPU fallback)."""
import os
import subprocess

import parallel_finite_difference_computation_b200 as fdw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_drivers_build_and_fail_loudly_without_gpu(tmp_path):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "apps")])
    for prog in ("stencil_code", "rtm_code", "mod_main", "rtm_main"):
        assert os.access(os.path.join(ROOT, "bin", prog), os.X_OK)
    nm = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(ROOT, "bin", "rtm_code")],
                        capture_output=True, text=True).stdout
    assert "fdw_forward" in nm and "fdw_backward" in nm and "cuda" not in nm.lower()
    if fdw.load().fdw_device_count() > 0:
        return
    (tmp_path / "input.bin").write_bytes(b"\0" * (24 * 24 * 4))
    (tmp_path / "input.dat").write_text("tmpdir=./input.bin\nnz=16\nnx=16\ndz=10\ndx=10\nnxb=4\nnzb=4\norder=8\n")
    p = subprocess.run([os.path.join(ROOT, "bin", "stencil_code"), "./input.dat", "o.bin"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU path" in p.stderr


def test_shim_libraries_export_the_reference_symbols():
    """libfdwave_cpufam.so / libfdwave_gpufam.so export exactly the (mangled) names the
    reference's callers bind (include/fdwave_cpufam.h, include/fdwave_gpufam.h)."""
    pkg = os.path.join(ROOT, "parallel_finite_difference_computation_b200")

    def exported(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        return {l.split()[-1] for l in out.splitlines() if " T " in l}

    cpu = exported(os.path.join(pkg, "libfdwave_cpufam.so"))
    want_cpu = {"_Z7fd_initiiifff", "_Z7fd_stepiPPfS0_S0_ii", "_Z10fd_destroyv", "_Z10calc_coefsi",
                "_Z9extendveliiiiPf", "_Z10taper_initiif", "_Z11taper_applyPPfiiii", "_Z12taper_apply2PPfiiii",
                "_Z13taper_destroyv", "_Z5ptsrciiiifPPf", "_Z14ricker_waveletiffPf", "_Z6rickerff"}
    assert want_cpu <= cpu, want_cpu - cpu
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_cpufam.so")
    if os.path.exists(ref):  # the reference's own objects export the same names
        assert want_cpu <= exported(ref)
    gpu = exported(os.path.join(pkg, "libfdwave_gpufam.so"))
    want_gpu = {"fd_init", "fd_init_cuda", "_Z13write_buffersPPfS0_S0_S_S_S0_S0_ii",
                "_Z10fd_forwardiPPfS0_S0_iiiiiPiS_i", "_Z7fd_backiPPfS0_S0_S0_S0_iiiiiiPS0_S0_S0_"}
    assert want_gpu <= gpu, want_gpu - gpu
