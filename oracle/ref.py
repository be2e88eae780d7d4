"""ctypes front-end of the reference's OWN objects built into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  oracle/_ref/ is produced by oracle/Makefile from the
sources under /root/reference (never copied into the repo) and is git-ignored;
it travels to the GPU box as prebuilt files.  Every accessor returns None /
raises RefUnavailable when the objects are missing so that callers can skip.

Libraries:
  libref_cpufam.so  fd.c + ptsrc.c + taper.c of dpct_gpu_rtm_domain_division
                    (built as C++ by g++ like the reference => mangled names)
  libref_gpuhost.so cuda_reference_RTM/lib/src/functions.c (C, gcc)
  libref_gpufam.so  cuda_reference_RTM/src/fd-code.cu for sm_100 (needs a GPU
                    for anything beyond the host tables)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


class RefUnavailable(RuntimeError):
    pass


def path(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def _load(name):
    p = path(name)
    if p is None:
        raise RefUnavailable("oracle/_ref/%s not built (run `make -C oracle`)" % name)
    return C.CDLL(p)


def available(name="libref_cpufam.so"):
    return path(name) is not None


def rows(a):
    """float** row-pointer table over a contiguous 2-D float32 array (the
    alloc2float contract, functions.c:168-182)."""
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] and a.ndim == 2
    n2, n1 = a.shape
    tab = (C.POINTER(C.c_float) * n2)()
    base = a.ctypes.data
    for i in range(n2):
        tab[i] = C.cast(base + 4 * n1 * i, C.POINTER(C.c_float))
    return tab


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class CpuFam:
    """Reference CPU family, function level (include/timestep/fd.h:4-7,
    include/boundary/taper.h:4-8, include/source/ptsrc.h:4-6)."""

    def __init__(self):
        L = self.L = _load("libref_cpufam.so")
        self._fd_init = L._Z7fd_initiiifff
        self._fd_init.argtypes = [C.c_int] * 3 + [C.c_float] * 3
        self._fd_step = L._Z7fd_stepiPPfS0_S0_ii
        self._fd_step.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        self._fd_destroy = L._Z10fd_destroyv
        self._calc_coefs = L._Z10calc_coefsi
        self._calc_coefs.argtypes = [C.c_int]
        self._calc_coefs.restype = C.POINTER(C.c_float)
        self._taper_init = L._Z10taper_initiif
        self._taper_init.argtypes = [C.c_int, C.c_int, C.c_float]
        self._taper_apply = L._Z11taper_applyPPfiiii
        self._taper_apply.argtypes = [C.c_void_p] + [C.c_int] * 4
        self._taper_apply2 = L._Z12taper_apply2PPfiiii
        self._taper_apply2.argtypes = [C.c_void_p] + [C.c_int] * 4
        self._taper_destroy = L._Z13taper_destroyv
        self._extendvel = L._Z9extendveliiiiPf
        self._extendvel.argtypes = [C.c_int] * 4 + [_f32p]
        self._ptsrc = L._Z5ptsrciiiifPPf
        self._ptsrc.argtypes = [C.c_int] * 4 + [C.c_float, C.c_void_p]
        self._ricker_wavelet = L._Z14ricker_waveletiffPf
        self._ricker_wavelet.argtypes = [C.c_int, C.c_float, C.c_float, _f32p]

    def calc_coefs(self, order):
        p = self._calc_coefs(order)
        return np.array([p[i] for i in range(order + 1)], np.float32)

    def ricker_wavelet(self, nt, dt, fpeak):
        s = np.zeros(nt, np.float32)
        self._ricker_wavelet(nt, dt, fpeak, s)
        return s

    def extendvel(self, nx, nz, nxb, nzb, vel):
        vel = np.ascontiguousarray(vel, np.float32).copy()
        self._extendvel(nx, nz, nxb, nzb, vel)
        return vel

    def fd_init(self, order, nxe, nze, dx, dz, dt):
        self._fd_init(order, nxe, nze, dx, dz, dt)

    def fd_step(self, order, p, pp, v2):
        nxe, nze = p.shape
        self._fd_step(order, rows(p), rows(pp), rows(v2), nze, nxe)

    def fd_destroy(self):
        self._fd_destroy()

    def taper_init(self, nxb, nzb, fac):
        self._taper_init(nxb, nzb, fac)

    def taper_apply(self, a, nx, nz, nxb, nzb):
        self._taper_apply(rows(a), nx, nz, nxb, nzb)

    def taper_apply2(self, a, nx, nz, nxb, nzb):
        self._taper_apply2(rows(a), nx, nz, nxb, nzb)

    def taper_destroy(self):
        self._taper_destroy()

    def ptsrc(self, xs, zs, ts, s):
        nxe, nze = s.shape
        self._ptsrc(xs, zs, nxe, nze, ts, rows(s))


class GpuHost:
    """Reference GPU-family host helpers (lib/include/functions.h:11-35)."""

    def __init__(self):
        L = self.L = _load("libref_gpuhost.so")
        L.calc_coefs.argtypes = [C.c_int]
        L.calc_coefs.restype = C.POINTER(C.c_float)
        L.ricker_wavelet.argtypes = [C.c_int, C.c_float, C.c_float, _f32p]
        L.extendvel_linear.argtypes = [C.c_int] * 4 + [C.c_void_p]
        self.libc = C.CDLL(None)

    def calc_coefs(self, order):
        p = self.L.calc_coefs(order)
        return np.array([p[i] for i in range(order + 1)], np.float32)

    def ricker_wavelet(self, nt, dt, fpeak):
        s = np.zeros(nt, np.float32)
        self.L.ricker_wavelet(nt, dt, fpeak, s)
        return s

    def extendvel_linear(self, nx, nz, nxb, nzb, vel, seed=None):
        vel = np.ascontiguousarray(vel, np.float32).copy()
        if seed is not None:
            self.libc.srand(C.c_uint(seed))
        self.L.extendvel_linear(nx, nz, nxb, nzb, rows(vel))
        return vel


class GpuFam:
    """Reference CUDA RTM (fd-code.cu) as a shared object.  taper_tables()
    works without a GPU (cudaMalloc just fails); forward()/back() need one."""

    def __init__(self):
        L = self.L = _load("libref_gpufam.so")
        L.fd_init_cuda.argtypes = [C.c_int] * 7 + [C.c_float]
        L.fd_init.argtypes = [C.c_int] * 7 + [C.c_float] * 4
        self._fwd = L._Z10fd_forwardiPPfS0_S0_iiiiiPiS_i
        self._fwd.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_void_p, _f32p, C.c_int]
        self._back = L._Z7fd_backiPPfS0_S0_S0_S0_iiiiiiPS0_S0_S0_
        self._back.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 6 + [C.c_void_p] * 3

    def taper_tables(self, nxe, nze, nxb, nzb, fac):
        self.L.fd_init_cuda(8, nxe, nze, nxb, nzb, 1, 1, fac)
        tx = C.POINTER(C.c_float).in_dll(self.L, "taper_x")
        tz = C.POINTER(C.c_float).in_dll(self.L, "taper_z")
        return (np.array([tx[i] for i in range(nxb)], np.float32),
                np.array([tz[i] for i in range(nzb)], np.float32))

    def launch_extents(self):
        g = lambda n: C.c_int.in_dll(self.L, n).value
        return g("gridx") * 8, g("gridz") * 8, g("gridBorder_z") * 8

    def fd_init(self, order, nxe, nze, nxb, nzb, nt, ns, fac, dx, dz, dt):
        self.L.fd_init(order, nxe, nze, nxb, nzb, nt, ns, fac, dx, dz, dt)

    def fd_forward(self, order, p, pp, v2, nt, is_, sz, sx, srce):
        """fd-code.cu:247; callers pass nz=nze, nx=nxe (fd-code.cu:499)."""
        nxe, nze = p.shape
        sxa = (C.c_int * len(sx))(*sx)
        self._fwd(order, rows(p), rows(pp), rows(v2), nze, nxe, nt, is_, sz, sxa,
                  np.ascontiguousarray(srce, np.float32), is_)

    def fd_back(self, order, p, pp, pr, ppr, v2, nt, is_, sz, gz, snaps, imloc, dobs_aux):
        """fd-code.cu:290.  snaps: (2,nxe,nze); dobs_aux: (ns, nx*nt)."""
        nxe, nze = p.shape
        s0, s1 = rows(snaps[0]), rows(snaps[1])
        snap_tab = (C.c_void_p * 2)(C.cast(s0, C.c_void_p), C.cast(s1, C.c_void_p))
        self._keep = (s0, s1)
        self._back(order, rows(p), rows(pp), rows(pr), rows(ppr), rows(v2), nze, nxe, nt, is_, sz, gz,
                   snap_tab, rows(imloc), rows(dobs_aux))
