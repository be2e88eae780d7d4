"""Profiling helper: one forward phase + one backward pass of the shared-memory tile kernel on the new_mod size
(415 x 295 extended grid), for `ncu -k regex:k_tile`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402
nx, nz, nb, nt = 315, 195, 50, int(os.environ.get("NT", "400"))
v2 = np.full((nx + 2 * nb, nz + 2 * nb), np.float32(2500.0) ** 2, np.float32)
with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU, taper=fdw.TAPER_TOP,
                compat_extents=True, nt=nt) as w:
    w.set_v2(v2); w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 20.0, fdw.FAMILY_GPU))
    w.forward(nb + 10, nb, download=False)
    w.backward(np.zeros((nx, nt), np.float32), nb)
    print("tile launches", w.tile_launches())
