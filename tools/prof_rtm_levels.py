"""Profiling helper: a few levels of the GPU-family RTM (forward, then backward = reconstruction + receiver step with
back-injection and imaging) on the C4 grid (8192 x 4096 + 40 border), for an `ncu` launch list / full capture."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402

nx, nz, nb, nt = 8192, 4096, 40, int(os.environ.get("NT", "8"))
nxe, nze = nx + 2 * nb, nz + 2 * nb
ve = np.full((nxe, nze), 3000.0, np.float32)
with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU, taper=fdw.TAPER_TOP,
                nt=nt) as w:
    w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_GPU))
    w.set_v2(ve * ve)
    w.forward(64, nb, download=False)
    w.sync()
    w.mark_begin()
    w.forward(64, nb, download=False)
    f_ms = w.mark_end()
    img = w.backward(np.zeros((nx, nt), np.float32), nb)
    print("forward %.3f ms/level" % (f_ms / nt))
