"""Diagnostic (GPU box): reproducibility of the reference's own CUDA fd_back (racy kernel_sism,
fd-code.cu:124-131) and its distance to the +1x oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from oracle import ref as R

nx, nz, nb, nt = 101, 83, 24, 400
nxe, nze = nx + 2 * nb, nz + 2 * nb
rng = np.random.default_rng(3)
vp = np.empty((nx, nz), np.float32); vp[:, : nz // 2] = 2200.0; vp[:, nz // 2:] = 3400.0
ve = np.zeros((nxe, nze), np.float32); ve[nb:nb + nx, nb:nb + nz] = vp
ve = O.extendvel_linear(nx, nz, nb, nb, ve, seed=100)
v2 = (ve * ve).astype(np.float32)
srce = O.ricker_wavelet(nt, 0.001, 25.0, O.FAM_G)
sx, sz, gz = nx // 4 + nb, nb, nb
dobs = (rng.standard_normal((1, nx, nt)) * 0.1).astype(np.float32)
g = R.GpuFam()
g.fd_init(8, nxe, nze, nb, nb, nt, 1, 0.75, 10.0, 10.0, 0.001)
P = np.zeros((nxe, nze), np.float32); PP = np.zeros((nxe, nze), np.float32)
g.fd_forward(8, P, PP, v2, nt, 0, sz, [sx], srce)
snaps = np.stack([P, PP]).copy()
imgs = []
for k in range(4):
    z = lambda: np.zeros((nxe, nze), np.float32)
    im = np.zeros((nx, nz), np.float32)
    g.fd_back(8, z(), z(), z(), z(), v2, nt, 0, sz, gz, snaps, im, dobs.reshape(1, nx * nt).copy())
    imgs.append(im.copy())
print()
for k in range(1, 4):
    print("reference fd_back run0 vs run%d: equal %s maxdiff %g relL2 %g" % (
        k, np.array_equal(imgs[0], imgs[k]), np.abs(imgs[0] - imgs[k]).max(),
        np.linalg.norm(imgs[0] - imgs[k]) / np.linalg.norm(imgs[0])))
cfg = O.GpuCfg(8, nxe, nze, nb, nb, nt, 10.0, 10.0, 0.001, 0.75, 1)
oim = O.gpu_back(cfg, snaps[0], snaps[1], v2, dobs[0], gz)
for k in range(4):
    print("oracle(+1x) vs reference run%d: equal %s maxdiff %g relL2 %g |img|max %g" % (
        k, np.array_equal(oim, imgs[k]), np.abs(oim - imgs[k]).max(),
        np.linalg.norm(oim - imgs[k]) / np.linalg.norm(imgs[k]), np.abs(imgs[k]).max()))
# one backward step with a huge single trace sample: multiplicity of the injection
for trial in range(3):
    g.fd_init(8, nxe, nze, nb, nb, 1, 1, 0.75, 10.0, 10.0, 0.001)
    d1 = np.zeros((1, nx * 1), np.float32); d1[0, :] = np.arange(1, nx + 1, dtype=np.float32)
    z = lambda: np.zeros((nxe, nze), np.float32)
    im = np.zeros((nx, nz), np.float32)
    s2 = np.ones((2, nxe, nze), np.float32)
    g.fd_back(8, z(), z(), z(), z(), v2, 1, 0, sz, gz, s2, im, d1)
    # img = p(=ones) * ppr_new ; ppr_new at (i+nb, gz) = m * d1[i]
    mult = im[:, gz - nb] / d1[0]
    print("trial %d injection multiplicity: min %g max %g, values != 1: %d" % (trial, mult.min(), mult.max(), int((mult != 1).sum())))
