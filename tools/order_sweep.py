"""Throughput of the plain propagation at every supported order (GPU-family recipe G, top sponge, point source) on an
N x N extended grid: orders 2..8 run the tuned 64-register kernels, orders 10..16 the one-kernel-per-epilogue
128-register instantiation."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402

n, nb, nt = int(os.environ.get("N", "8192")), 40, int(os.environ.get("NT", "60"))
v2 = np.full((n, n), np.float32(3000.0) ** 2, np.float32)
for order in (2, 4, 6, 8, 10, 12, 14, 16):
    with fdw.Wave2D(n - 2 * nb, n - 2 * nb, nb, nb, 10.0, 10.0, 0.001, order=order, fac=0.75, family=fdw.FAMILY_GPU,
                    taper=fdw.TAPER_TOP) as w:
        w.set_v2(v2)
        w.set_wavelet(fdw.host.ricker_wavelet(2 * nt, 0.001, 20.0, fdw.FAMILY_GPU))
        w.set_source(n // 2, nb)
        w.zero()
        w.advance(0, nt)
        w.sync()
        w.mark_begin()
        w.advance(nt, nt)
        ms = w.mark_end()
        g = float(n) * n * nt / ms / 1e6
        print("order %2d  %8.3f ms/level  %7.1f Gpts/s  %6.0f GB/s at 16 B/pt" % (order, ms / nt, g, g * 16))
