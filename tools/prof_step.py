"""Profiling helper: a few levels of the headline propagator (16384 x 16384, recipe G, top sponge, point source)
for `ncu -k regex:k_step` (launch list, --set full captures, DRAM byte counts)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402

n, nb = int(os.environ.get("N", "16384")), 40
recipe = {"G": fdw.RECIPE_G, "C": fdw.RECIPE_C}[os.environ.get("RECIPE", "G")]
family = fdw.FAMILY_GPU if recipe == fdw.RECIPE_G else fdw.FAMILY_CPU
with fdw.Wave2D(n - 2 * nb, n - 2 * nb, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=family, recipe=recipe,
                taper=fdw.TAPER_TOP) as w:
    v2 = np.full((n, n), np.float32(3000.0) ** 2, np.float32)
    w.set_v2(v2)
    w.set_wavelet(fdw.host.ricker_wavelet(64, 0.001, 20.0, family))
    w.set_source(n // 2, nb)
    w.zero()
    w.advance(0, int(os.environ.get("LEVELS", "6")))
    w.sync()
    print("launches", w.launch_count())
