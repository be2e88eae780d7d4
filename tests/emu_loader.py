"""TEST INFRASTRUCTURE ONLY: builds and loads tests/emu/libfdwave_emu.so, the
host build of libfdwave's sources against a fake CUDA runtime.  Used by the
CPU-only tests to exercise the library's host logic and kernel bodies; never
imported by the product package."""
import ctypes as C
import os
import subprocess

from parallel_finite_difference_computation_b200 import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMU = None


def load():
    global _EMU
    if _EMU is None:
        d = os.path.join(_HERE, "emu")
        # FDW_EMU_ASAN=1 (run pytest with LD_PRELOAD=$(gcc -print-file-name=libasan.so)): the same host build with
        # -fsanitize=address in its own object directory -- every emulated device allocation is a malloc, so the kernel
        # bodies' out-of-bounds accesses (absorbed by guard rows / pad columns by design) are checked against the
        # allocation bounds; compute-sanitizer is not available on the GPU pool
        if os.environ.get("FDW_EMU_ASAN") == "1":
            subprocess.check_call(["make", "-s", "-C", d, "-j8", "ASAN=1"], stdout=subprocess.DEVNULL)
            _EMU = _lib.bind(C.CDLL(os.path.join(d, "asan", "libfdwave_emu.so")))
            return _EMU
        subprocess.check_call(["make", "-s", "-C", d, "-j8"], stdout=subprocess.DEVNULL)
        _EMU = _lib.bind(C.CDLL(os.path.join(d, "libfdwave_emu.so")))
    return _EMU
