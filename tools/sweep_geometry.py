"""In-process A/B of the launch geometry of the bulk kernel on the headline grid (box-to-box spread is +-3 %, so
variants must be compared inside one process, interleaved).  FDW_ROWS_PER_CTA / FDW_THREADS are read at
fdw_create, so one context per variant."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402

n, nb, levels = 16384, 40, 60
nx = nz = n - 2 * nb
ve = np.empty((n, n), np.float32)
ve[:, : n // 3] = 2000.0
ve[:, n // 3: 2 * n // 3] = 3000.0
ve[:, 2 * n // 3:] = 4000.0
v2 = ve * ve
srce = fdw.host.ricker_wavelet(10000, 0.001, 20.0, fdw.FAMILY_GPU)
variants = [(7, 64), (7, 32), (6, 32), (8, 32), (5, 64), (6, 64), (9, 64), (10, 64), (14, 64), (7, 128), (3, 64), (4, 64)]
ctxs = {}
for rpc, thr in variants:
    os.environ["FDW_ROWS_PER_CTA"], os.environ["FDW_THREADS"] = str(rpc), str(thr)
    w = fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU, taper=fdw.TAPER_TOP)
    w.set_v2(v2)
    w.set_wavelet(srce)
    w.set_source(n // 2, nb)
    w.zero()
    w.advance(0, 20)
    w.sync()
    ctxs[(rpc, thr)] = w
res = {v: [] for v in variants}
for rnd in range(4):
    for v in variants:
        w = ctxs[v]
        w.mark_begin()
        w.advance(100 + rnd * levels, levels)
        ms = w.mark_end()
        res[v].append(n * n * levels / (ms * 1e-3) / 1e9)
for v in variants:
    print("rows_per_cta=%3d threads=%3d  Gpts/s per round: %s  median %.1f" % (
        v[0], v[1], " ".join("%.1f" % x for x in res[v]), float(np.median(res[v]))))
