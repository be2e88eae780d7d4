"""Host-side tables and input.dat readers (no GPU needed).  Thin wrappers over
the C functions of csrc/fdw_host.c; each cites the reference function it
reproduces bit for bit."""
import ctypes as C

import numpy as np

from . import _lib


def _L(lib=None):
    return lib if lib is not None else _lib.load()


def calc_coefs(order, family=_lib.FAMILY_GPU, lib=None):
    """calc_coefs: functions.c:78-123 / fd.c:54-97."""
    L = _L(lib)
    out = np.zeros(order + 1, np.float32)
    _lib.check(L, L.fdw_calc_coefs(order, family, out))
    return out


def ricker_wavelet(nt, dt, fpeak, family=_lib.FAMILY_GPU, lib=None):
    """ricker_wavelet: functions.c:293-299 / ptsrc.c:88-99."""
    L = _L(lib)
    s = np.zeros(nt, np.float32)
    _lib.check(L, L.fdw_ricker_wavelet(nt, dt, fpeak, family, s))
    return s


def taper_table(nb, fac, family=_lib.FAMILY_GPU, lib=None):
    """sponge table: fd-code.cu:159-166 / taper.c:33-42."""
    L = _L(lib)
    t = np.zeros(nb, np.float32)
    _lib.check(L, L.fdw_taper_table(nb, fac, family, t))
    return t


def extendvel(nx, nz, nxb, nzb, vel, lib=None):
    """extendvel: taper.c:7-23 (returns a copy)."""
    L = _L(lib)
    vel = np.ascontiguousarray(vel, np.float32).copy()
    _lib.check(L, L.fdw_extendvel(nx, nz, nxb, nzb, vel))
    return vel


def extendvel_linear(nx, nz, nxb, nzb, vel, seed=None, lib=None):
    """extendvel_linear: functions.c:301-359 (libc rand(); seed=None keeps the stream)."""
    L = _L(lib)
    vel = np.ascontiguousarray(vel, np.float32).copy()
    if seed is not None:
        C.CDLL(None).srand(C.c_uint(seed))
    _lib.check(L, L.fdw_extendvel_linear(nx, nz, nxb, nzb, vel))
    return vel


def ptsrc_weights(lib=None):
    L = _L(lib)
    w = np.zeros(49, np.float32)
    _lib.check(L, L.fdw_ptsrc_weights(w))
    return w.reshape(7, 7)


def _input_dict(inp):
    d = {}
    for name, _ in _lib.Input._fields_:
        v = getattr(inp, name)
        d[name] = v.decode() if isinstance(v, bytes) else v
    return d


def read_input_gpu(path, apply_defaults=True, lib=None):
    """GPU-family input.dat dialect (functions.c:10-75, defaults fd-code.cu:367-377)."""
    L = _L(lib)
    inp = _lib.Input()
    _lib.check(L, L.fdw_read_input_gpu(path.encode(), int(apply_defaults), C.byref(inp)))
    return _input_dict(inp)


def read_input_stencil(path, lib=None):
    """stencil program dialect (fd-source-code.cu:34-108)."""
    L = _L(lib)
    inp = _lib.Input()
    _lib.check(L, L.fdw_read_input_stencil(path.encode(), C.byref(inp)))
    return _input_dict(inp)


def read_input_cpu(path, apply_defaults=True, lib=None):
    """CPU-family par= file (CWP getpar semantics, mod_main.cpp:58-85)."""
    L = _L(lib)
    inp = _lib.Input()
    _lib.check(L, L.fdw_read_input_cpu(path.encode(), int(apply_defaults), C.byref(inp)))
    return _input_dict(inp)
