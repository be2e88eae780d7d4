"""Diagnostic (GPU box): how reproducible is the reference's own CUDA fd_forward, and which
reading of its racy corner sponge (kernel_tapper, fd-code.cu:94-117) does the hardware produce?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from oracle import ref as R
import parallel_finite_difference_computation_b200 as fdw

nx, nz, nb, nt = 101, 83, 24, 400
nxe, nze = nx + 2 * nb, nz + 2 * nb
rng = np.random.default_rng(3)
vp = np.empty((nx, nz), np.float32); vp[:, : nz // 2] = 2200.0; vp[:, nz // 2:] = 3400.0
ve = np.zeros((nxe, nze), np.float32); ve[nb:nb + nx, nb:nb + nz] = vp
ve = O.extendvel_linear(nx, nz, nb, nb, ve, seed=100)
v2 = (ve * ve).astype(np.float32)
srce = O.ricker_wavelet(nt, 0.001, 25.0, O.FAM_G)
sx, sz = nx // 4 + nb, nb
g = R.GpuFam()
g.fd_init(8, nxe, nze, nb, nb, nt, 1, 0.75, 10.0, 10.0, 0.001)
runs = []
for k in range(3):
    P = np.zeros((nxe, nze), np.float32); PP = np.zeros((nxe, nze), np.float32)
    g.fd_forward(8, P, PP, v2, nt, 0, sz, [sx], srce)
    runs.append((P.copy(), PP.copy()))
for k in (1, 2):
    print("reference run0 vs run%d: P equal %s  PP equal %s  maxdiff %g" % (
        k, np.array_equal(runs[0][0], runs[k][0]), np.array_equal(runs[0][1], runs[k][1]),
        np.abs(runs[0][1] - runs[k][1]).max()))
cfg = O.GpuCfg(8, nxe, nze, nb, nb, nt, 10.0, 10.0, 0.001, 0.75, 1)
oP, oPP = O.gpu_forward(cfg, v2, srce, sx, sz)
rP, rPP = runs[0]
d = np.abs(oPP - rPP)
print("oracle(both factors) vs reference: PP equal %s maxdiff %g relL2 %g ; |PP|max %g" % (
    np.array_equal(oPP, rPP), d.max(), np.linalg.norm(oPP - rPP) / np.linalg.norm(rPP), np.abs(rPP).max()))
bad = np.argwhere(oPP != rPP)
if len(bad):
    print("differing rows range", bad[:, 0].min(), bad[:, 0].max(), "cols", bad[:, 1].min(), bad[:, 1].max(), "count", len(bad))
# short runs: where do the first differences appear?
for n_short in (1, 2, 3, 5, 10, 30):
    g.fd_init(8, nxe, nze, nb, nb, n_short, 1, 0.75, 10.0, 10.0, 0.001)
    P = rng.standard_normal((nxe, nze)).astype(np.float32); PP = rng.standard_normal((nxe, nze)).astype(np.float32)
    ux, uz = (nxe // 8) * 8, (nze // 8) * 8
    P[ux:] = 0; P[:, uz:] = 0; PP[ux:] = 0; PP[:, uz:] = 0
    P0, PP0 = P.copy(), PP.copy()
    g.fd_forward(8, P, PP, v2, n_short, 0, sz, [sx], srce)
    cfg2 = O.GpuCfg(8, nxe, nze, nb, nb, n_short, 10.0, 10.0, 0.001, 0.75, 1)
    oP, oPP = O.gpu_forward(cfg2, v2, srce, sx, sz, P0, PP0)
    bad = np.argwhere((oPP != PP) | (oP != P))
    msg = "nt=%d random init: differing %d" % (n_short, len(bad))
    if len(bad):
        msg += " rows %d..%d cols %d..%d" % (bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max())
        i, j = bad[0]
        msg += " first (%d,%d): oracle %r ref %r ratio %r" % (i, j, oPP[i, j], PP[i, j], PP[i, j] / oPP[i, j] if oPP[i, j] else None)
    print(msg)
