"""ctypes front-end of the CPU oracle (oracle/fdwave_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(parallel_finite_difference_computation_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class GpuCfg(C.Structure):
    _fields_ = [("order", C.c_int), ("nxe", C.c_int), ("nze", C.c_int), ("nxb", C.c_int),
                ("nzb", C.c_int), ("nt", C.c_int), ("dx", C.c_float), ("dz", C.c_float),
                ("dt", C.c_float), ("fac", C.c_float), ("compat", C.c_int)]


class CpuCfg(C.Structure):
    _fields_ = [("order", C.c_int), ("nx", C.c_int), ("nz", C.c_int), ("nxb", C.c_int),
                ("nzb", C.c_int), ("nt", C.c_int), ("dx", C.c_float), ("dz", C.c_float),
                ("dt", C.c_float), ("fac", C.c_float)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "fdwave_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_get_max_threads.restype = C.c_int
        L.orc_calc_coefs.argtypes = [C.c_int, C.c_int, f32p]
        L.orc_scalars.argtypes = [C.c_float] * 3 + [C.POINTER(C.c_float)] * 3
        L.orc_premult_coefs.argtypes = [C.c_int, C.c_float, C.c_float, f32p, f32p]
        L.orc_ricker_wavelet.argtypes = [C.c_int, C.c_float, C.c_float, C.c_int, f32p]
        L.orc_taper_table.argtypes = [C.c_int, C.c_float, C.c_int, f32p]
        L.orc_extendvel.argtypes = [C.c_int] * 4 + [f32p]
        L.orc_extendvel_linear.argtypes = [C.c_int] * 4 + [f32p]
        L.orc_srand.argtypes = [C.c_uint]
        L.orc_lap_G.argtypes = [C.c_int] * 5 + [f32p, f32p, f32p, f32p]
        L.orc_lap_C.argtypes = [C.c_int] * 3 + [f32p, f32p, f32p, C.c_float, C.c_float]
        L.orc_time.argtypes = [C.c_int] * 4 + [f32p, f32p, f32p, f32p, C.c_float]
        L.orc_taper_top.argtypes = [f32p] + [C.c_int] * 5 + [f32p, f32p]
        L.orc_taper_4.argtypes = [f32p] + [C.c_int] * 4 + [f32p, f32p]
        L.orc_ptsrc.argtypes = [C.c_int] * 4 + [C.c_float, f32p]
        L.orc_stencil.argtypes = [C.c_int] * 3 + [C.c_float, C.c_float, f32p, f32p]
        L.orc_gpu_forward.argtypes = [C.POINTER(GpuCfg), f32p, f32p, f32p, f32p, C.c_int, C.c_int]
        L.orc_gpu_back.argtypes = [C.POINTER(GpuCfg), f32p, f32p, f32p, f32p, C.c_int, f32p]
        L.orc_image_laplacian.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p]
        L.orc_fd_step.argtypes = [C.c_int] * 3 + [f32p, f32p, f32p, f32p] + [C.c_float] * 3
        L.orc_mod_shot.argtypes = [C.POINTER(CpuCfg), f32p, f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.orc_rtm_shot.argtypes = [C.POINTER(CpuCfg), f32p, f32p, C.c_int, C.c_int, C.c_int, f32p,
                                   C.c_int, C.c_int, f32p, C.c_void_p]
        _LIB = L
    return _LIB


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return lib().orc_get_max_threads()


# ---------------------------------------------------------------- tables
FAM_G, FAM_C, FAM_GHOST = 0, 1, 2


def calc_coefs(order, family=FAM_G):
    out = np.zeros(order + 1, np.float32)
    lib().orc_calc_coefs(order, family, out)
    return out


def scalars(dx, dz, dt):
    a, b, c = C.c_float(), C.c_float(), C.c_float()
    lib().orc_scalars(dx, dz, dt, C.byref(a), C.byref(b), C.byref(c))
    return np.float32(a.value), np.float32(b.value), np.float32(c.value)


def premult_coefs(order, dx, dz):
    cx = np.zeros(order + 1, np.float32)
    cz = np.zeros(order + 1, np.float32)
    lib().orc_premult_coefs(order, dx, dz, cx, cz)
    return cx, cz


def ricker_wavelet(nt, dt, fpeak, family):
    s = np.zeros(nt, np.float32)
    lib().orc_ricker_wavelet(nt, dt, fpeak, family, s)
    return s


def taper_table(nb, fac, family):
    t = np.zeros(nb, np.float32)
    lib().orc_taper_table(nb, fac, family, t)
    return t


def extendvel(nx, nz, nxb, nzb, vel):
    vel = np.ascontiguousarray(vel, np.float32).copy()
    lib().orc_extendvel(nx, nz, nxb, nzb, vel)
    return vel


def extendvel_linear(nx, nz, nxb, nzb, vel, seed=None):
    vel = np.ascontiguousarray(vel, np.float32).copy()
    if seed is not None:
        lib().orc_srand(seed)
    lib().orc_extendvel_linear(nx, nz, nxb, nzb, vel)
    return vel


# ---------------------------------------------------------------- kernels
def lap_G(order, p, cx, cz, ilim=None, jlim=None, out=None):
    nxe, nze = p.shape
    lap = np.zeros_like(p) if out is None else out
    lib().orc_lap_G(order, nxe, nze, nxe if ilim is None else ilim, nze if jlim is None else jlim,
                    p, lap, cx, cz)
    return lap


def lap_C(order, p, coefs, dx2inv, dz2inv, out=None):
    nxe, nze = p.shape
    lap = np.zeros_like(p) if out is None else out
    lib().orc_lap_C(order, nxe, nze, p, lap, coefs, dx2inv, dz2inv)
    return lap


def time_update(p, pp, v2, lap, dt2, ux=None, uz=None):
    nxe, nze = p.shape
    lib().orc_time(nxe, nze, nxe if ux is None else ux, nze if uz is None else uz, p, pp, v2, lap, dt2)


def taper_top(a, nxb, taperx, taperz, ux=None, uzb=None):
    nxe, nze = a.shape
    lib().orc_taper_top(a, nxe, nze, nxb, nxe if ux is None else ux, len(taperz) if uzb is None else uzb,
                        taperx, taperz)


def taper_4(a, nx, nz, nxb, nzb, taperx, taperz):
    lib().orc_taper_4(a, nx, nz, nxb, nzb, taperx, taperz)


def ptsrc(xs, zs, ts, s):
    nxe, nze = s.shape
    lib().orc_ptsrc(xs, zs, nxe, nze, ts, s)


def stencil(order, dx, dz, p):
    nxe, nze = p.shape
    out = np.empty_like(p)
    lib().orc_stencil(order, nxe, nze, dx, dz, p, out)
    return out


def fd_step(order, p, pp, v2, lap, dx, dz, dt):
    nxe, nze = p.shape
    lib().orc_fd_step(order, nxe, nze, p, pp, v2, lap, dx, dz, dt)


# ---------------------------------------------------------------- pipelines
def gpu_forward(cfg, v2, srce, sx, sz, p=None, pp=None):
    """fd_forward of the GPU family. Returns (P, PP) = (older, newest)."""
    shp = (cfg.nxe, cfg.nze)
    p = np.zeros(shp, np.float32) if p is None else p.copy()
    pp = np.zeros(shp, np.float32) if pp is None else pp.copy()
    lib().orc_gpu_forward(C.byref(cfg), p, pp, v2, srce, sx, sz)
    return p, pp


def gpu_back(cfg, snap0, snap1, v2, dobs, gz):
    nx, nz = cfg.nxe - 2 * cfg.nxb, cfg.nze - 2 * cfg.nzb
    imloc = np.zeros((nx, nz), np.float32)
    lib().orc_gpu_back(C.byref(cfg), snap0, snap1, v2, np.ascontiguousarray(dobs, np.float32), gz, imloc)
    return imloc


def mod_shot(cfg, v2, srce, sx, sz, gz):
    data = np.zeros((cfg.nx, cfg.nt), np.float32)
    lib().orc_mod_shot(C.byref(cfg), v2, srce, sx, sz, gz, data)
    return data


def rtm_shot(cfg, v2, srce, sx, sz, gz, dobs_all, is_=0, want_swf=False):
    dobs_all = np.ascontiguousarray(dobs_all, np.float32)
    ns = dobs_all.shape[0]
    imloc = np.zeros((cfg.nx, cfg.nz), np.float32)
    swf = np.zeros((cfg.nt, cfg.nx, cfg.nz), np.float32) if want_swf else None
    lib().orc_rtm_shot(C.byref(cfg), v2, srce, sx, sz, gz, dobs_all, ns, is_, imloc,
                       swf.ctypes.data if want_swf else None)
    return (imloc, swf) if want_swf else imloc


def image_laplacian(img, dx, dz):
    """laplace.f90:24-28 on an [nx][nz] image (parity unpinned: no gfortran, no shipped output)"""
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty_like(img)
    lib().orc_image_laplacian(img.shape[0], img.shape[1], dx, dz, img, out)
    return out
