"""Where does a time level of a mid-size grid go?  C5 grid (8192 x 2048 + 40 border), CPU-family recipe C, one GPU:
us per level of the modelling phase for each sponge kind (NONE = one bulk launch per level, TOP = bulk + 1 strip,
FOUR = bulk + 4 strips), CUDA events around the level loop.  Run it under different FDW_* knobs (FDW_FORK_LIMIT,
FDW_LEVEL_GRAPH, ...) to separate kernel time from launch / dependency overhead."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402

nx, nz, nb = int(os.environ.get("NX", "8192")), int(os.environ.get("NZ", "2048")), 40
nt = int(os.environ.get("NT", "300"))
nxe, nze = nx + 2 * nb, nz + 2 * nb
v2 = np.full((nxe, nze), np.float32(3000.0) ** 2, np.float32)
srce = fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_CPU)
pts = float(nxe) * nze
out = []
for name, taper in (("NONE", fdw.TAPER_NONE), ("TOP", fdw.TAPER_TOP), ("FOUR", fdw.TAPER_FOUR)):
    for phase, pname, hist in ((fdw.PHASE_MODEL, "model", False), (fdw.PHASE_RTM_FWD, "fwd+hist", True)):
        with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU, taper=taper, nt=nt,
                        history=hist) as w:
            w.set_v2(v2)
            w.set_wavelet(srce)
            w.shot_phase_device(phase, nb + nx // 2, nb, nb)
            w.sync()
            l0 = w.launch_count()
            w.mark_begin()
            w.shot_phase_device(phase, nb + nx // 2, nb, nb)
            ms = w.mark_end()
            out.append("%-4s %-8s %7.2f us/level %6.1f Gpts/s  %.1f launches/level" %
                       (name, pname, ms / nt * 1e3, pts * nt / ms / 1e6, (w.launch_count() - l0) / nt))
print(os.environ.get("TAG", ""), " | ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("FDW_")))
print("\n".join(out))
