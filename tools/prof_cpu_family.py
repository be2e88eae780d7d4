"""Profiling helper for the epilogue kernels: a few levels of each propagation phase on a grid well
above L2, for `ncu -k regex:k_step --set full` and for launch lists:
  CPU family (recipe C):  mod_main  -> k_step<8,1,1>  (record)
                          rtm fwd   -> k_step<8,1,4>  (history store)
                          rtm bwd   -> k_step<8,1,10> (back-injection + imaging against the history)
                          plain     -> k_step<8,1,0>
  GPU family (recipe G):  fd_back   -> k_step<8,0,0> (reconstruction) + k_step<8,0,18> (injection + imaging)
Grid: NX x NZ interior + 40 border (default 8192 x 8192: 1.1 GB per field)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402

nx, nz, nb = int(os.environ.get("NX", "8192")), int(os.environ.get("NZ", "8192")), 40
nt = int(os.environ.get("NT", "3"))
nxe, nze = nx + 2 * nb, nz + 2 * nb
v2 = np.full((nxe, nze), np.float32(3000.0) ** 2, np.float32)
sx, sz, gz = nb + nx // 2, nb, nb
which = os.environ.get("WHICH", "plain,model,rtm,gpufam").split(",")

if "plain" in which:
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU, taper=fdw.TAPER_TOP) as w:
        w.set_v2(v2)
        w.set_wavelet(fdw.host.ricker_wavelet(64, 0.001, 15.0, fdw.FAMILY_CPU))
        w.set_source(sx, sz)
        w.zero()
        w.advance(0, nt)
        w.sync()
if "model" in which:
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU, taper=fdw.TAPER_TOP,
                    nt=nt) as w:
        w.set_v2(v2)
        w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_CPU))
        w.shot_phase_device(fdw.PHASE_MODEL, sx, sz, gz)
        w.sync()
if "rtm" in which:
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU, taper=fdw.TAPER_TOP,
                    nt=nt, history=True) as w:
        w.set_v2(v2)
        w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_CPU))
        w.shot_phase_device(fdw.PHASE_RTM_FWD, sx, sz, gz)
        w.shot_phase_device(fdw.PHASE_RTM_BWD, sx, sz, gz, np.zeros((1, nx, nt), np.float32), 0)
        w.sync()
if "gpufam" in which:
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU, taper=fdw.TAPER_TOP,
                    nt=nt) as w:
        w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_GPU))
        w.set_v2(v2)
        w.forward(sx, sz, download=False)
        w.backward_device(np.zeros((nx, nt), np.float32), gz)
        w.sync()
print("ok")
