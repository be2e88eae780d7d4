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
