"""Regenerate tests/golden/ from the reference checkout.

Run in the authoring container (needs /root/reference and `make -C oracle`):
    python tests/golden/make_golden.py

Two kinds of fixtures are produced:
 (1) the reference's OWN shipped inputs / golden outputs for the hot path,
     copied byte-for-byte (data files, not sources):
       stencil/input.bin, stencil/output_teste.bin      (config 1 pin)
       3lay_mod/{3layer_151x151.bin,dobs.bin,dir.img,dir.image}  (CPU family pins)
       new_mod/vel_ext_rnd.shot5.bin                     (slice [5] of vel_ext_rnd.6)
 (2) ref_vectors.npz: outputs of the reference's own functions, executed here
     from oracle/_ref/libref_*.so (built from the reference sources in place),
     on small seeded inputs -- host tables and one call of each hot-path
     function.  tests/ compare the oracle and the CUDA path against these, so
     the checks still bind on the GPU box where /root/reference is absent.
"""
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("FDW_REFERENCE", "/root/reference")

from oracle import ref as R  # noqa: E402


def copy(src, dst):
    dst = os.path.join(HERE, dst)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(os.path.join(REF, src), dst)
    os.chmod(dst, 0o644)


def main():
    copy("cuda_reference_stencil_computation/input.bin", "stencil/input.bin")
    copy("dpct_migrated_stencil_computation/output_teste.bin", "stencil/output_teste.bin")
    for f in ("3layer_151x151.bin", "dobs.bin", "dir.img", "dir.image"):
        copy("dpct_gpu_rtm_domain_division/build/3lay_mod/" + f, "3lay_mod/" + f)
    nxe, nze = 415, 295
    ve = np.fromfile(os.path.join(REF, "cuda_reference_RTM/models/new_mod/vel_ext_rnd.6"), np.float32)
    ve = ve.reshape(6, nxe, nze)
    os.makedirs(os.path.join(HERE, "new_mod"), exist_ok=True)
    ve[5].tofile(os.path.join(HERE, "new_mod/vel_ext_rnd.shot5.bin"))

    cpu, gh, gf = R.CpuFam(), R.GpuHost(), R.GpuFam()
    out = {}
    # ---- host tables
    for order in (2, 4, 6, 8, 10, 12, 16):
        out["coefs_cpu_%d" % order] = cpu.calc_coefs(order)
        out["coefs_gpu_%d" % order] = gh.calc_coefs(order)
    ricker_cases = [(1001, 0.001, 30.0), (1700, 0.001, 20.0), (3004, 0.001, 6.5), (2000, 0.0015, 7.0),
                    (401, 0.001, 40.0)]
    out["ricker_cases"] = np.array(ricker_cases, np.float64)
    for k, (nt, dt, fp) in enumerate(ricker_cases):
        out["ricker_cpu_%d" % k] = cpu.ricker_wavelet(nt, dt, fp)
        out["ricker_gpu_%d" % k] = gh.ricker_wavelet(nt, dt, fp)
    taper_cases = [(50, 0.75), (40, 0.010), (40, 0.7), (37, 0.33), (40, 0.75)]
    out["taper_cases"] = np.array(taper_cases, np.float64)
    for k, (nb, fac) in enumerate(taper_cases):
        tx, tz = gf.taper_tables(415, 295, nb, nb, fac)
        out["taper_gpu_%d" % k] = tx
        # CPU family table: recover it by tapering a field of ones
        cpu.taper_init(nb, nb, fac)
        ones = np.ones((2 * nb + 3, 2 * nb + 5), np.float32)
        cpu.taper_apply2(ones, 3, 5, nb, nb)
        out["taper_cpu_%d" % k] = ones[nb + 1, :nb].copy()
        cpu.taper_destroy()
    out["gpu_launch_extents_415x295_nb50"] = np.array(
        (gf.taper_tables(415, 295, 50, 50, 0.75), gf.launch_extents())[1], np.int32)
    # ---- velocity extension
    rng = np.random.default_rng(20261018)
    nx, nz, nxb, nzb = 23, 17, 6, 5
    v = np.zeros((nx + 2 * nxb, nz + 2 * nzb), np.float32)
    v[nxb:nxb + nx, nzb:nzb + nz] = rng.uniform(1500, 4500, (nx, nz)).astype(np.float32)
    out["ext_in"] = v
    out["ext_dims"] = np.array([nx, nz, nxb, nzb], np.int32)
    out["extendvel_out"] = cpu.extendvel(nx, nz, nxb, nzb, v)
    out["extendvel_linear_seed1_out"] = gh.extendvel_linear(nx, nz, nxb, nzb, v, seed=1)
    # ---- one call of each CPU-family hot-path function
    nx, nz, nxb, nzb, order = 41, 29, 7, 6, 8
    nxe, nze = nx + 2 * nxb, nz + 2 * nzb
    p = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    pp = rng.uniform(-1, 1, (nxe, nze)).astype(np.float32)
    v2 = (rng.uniform(1500, 4500, (nxe, nze)).astype(np.float32)) ** 2
    out["step_dims"] = np.array([nx, nz, nxb, nzb, order], np.int32)
    out["step_scal"] = np.array([10.0, 12.5, 0.001, 0.35], np.float32)  # dx dz dt fac
    out["step_p"], out["step_pp"], out["step_v2"] = p, pp, v2
    for order_k in (2, 4, 6, 8):
        cpu.fd_init(order_k, nxe, nze, 10.0, 12.5, 0.001)
        a, b = p.copy(), pp.copy()
        for _ in range(3):
            cpu.fd_step(order_k, a, b, v2)
            a, b = b, a
        cpu.fd_destroy()
        out["fd_step3_o%d_p" % order_k], out["fd_step3_o%d_pp" % order_k] = a, b
    cpu.taper_init(nxb, nzb, 0.35)
    a = p.copy(); cpu.taper_apply(a, nx, nz, nxb, nzb); out["taper_apply_out"] = a
    a = p.copy(); cpu.taper_apply2(a, nx, nz, nxb, nzb); out["taper_apply2_out"] = a
    cpu.taper_destroy()
    for k, (xs, zs) in enumerate(((20, 15), (1, 2), (nxe - 2, nze - 1))):
        a = p.copy(); cpu.ptsrc(xs, zs, np.float32(0.731), a); out["ptsrc_out_%d" % k] = a
    out["ptsrc_pos"] = np.array(((20, 15), (1, 2), (nxe - 2, nze - 1)), np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "vectors")


if __name__ == "__main__":
    main()
