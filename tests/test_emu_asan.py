"""CPU-only: the kernel bodies under AddressSanitizer.  The host stand-in allocates every "device" buffer with
malloc, so an access of a kernel body outside its allocation (the guard rows and pad columns that absorb the
stencil's edge loads are part of the allocation) is an ASan error.  compute-sanitizer is not available on the GPU
pool; this is the bounds check the new launch paths get (orders up to 16 with two z-neighbour blocks per side,
folded multi-rectangle strips, the in-place sponge pass, graph replay, the peer halo protocol)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libasan():
    try:
        p = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    except OSError:
        return None
    return p if p and os.path.isabs(p) and os.path.exists(p) else None


def test_kernel_bodies_run_clean_under_address_sanitizer():
    lib = _libasan()
    if lib is None:
        pytest.skip("no libasan.so next to gcc")
    env = dict(os.environ, FDW_EMU_ASAN="1", LD_PRELOAD=lib, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
    sel = "stencil or sponge or above_order or level_loop or compat or wide_grid or rtm_main_shot or graph_replay"
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(ROOT, "tests", "test_emu_parity.py"), os.path.join(ROOT, "tests", "test_peer_halo_emu.py"),
                        "-k", sel], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "AddressSanitizer" not in tail, tail
    assert " passed" in r.stdout
