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
        subprocess.check_call(["make", "-s", "-C", d, "-j8"], stdout=subprocess.DEVNULL)
        _EMU = _lib.bind(C.CDLL(os.path.join(d, "libfdwave_emu.so")))
    return _EMU
