"""Wave2D: one libfdwave context (= one GPU, one slab) behind a small Python
class.  All numerics happen in the CUDA library; this file only marshals
numpy host buffers across the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (FAMILY_CPU, FAMILY_GPU, RECIPE_C, RECIPE_FAST, RECIPE_G, SRC_GAUSS7, SRC_POINT,  # noqa: F401
                   TAPER_FOUR, TAPER_NONE, TAPER_TOP)


def _opt(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Wave2D:
    def __init__(self, nx, nz, nxb, nzb, dx, dz, dt, order=8, fac=0.7, family=FAMILY_GPU, recipe=None,
                 taper=TAPER_TOP, compat_extents=False, device=0, slab=None, history=False, nt=0, lib=None):
        self.L = lib if lib is not None else _lib.load()
        if recipe is None:
            recipe = RECIPE_G if family == FAMILY_GPU else RECIPE_C
        x0, x1 = slab if slab is not None else (0, 0)
        self.prm = _lib.Params(nx, nz, nxb, nzb, order, dx, dz, dt, fac, family, recipe, taper,
                               int(compat_extents), device, x0, x1, int(history), nt)
        self.nxe, self.nze = nx + 2 * nxb, nz + 2 * nzb
        self.nx, self.nz, self.nxb, self.nzb, self.nt = nx, nz, nxb, nzb, nt
        h = C.c_void_p()
        _lib.check(self.L, self.L.fdw_create(C.byref(self.prm), C.byref(h)))
        self.h = h

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.L.fdw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        _lib.check(self.L, rc)

    def _grid(self, a, name):
        a = np.ascontiguousarray(a, np.float32)
        if a.shape != (self.nxe, self.nze):
            raise ValueError("%s must have shape (%d,%d), got %s" % (name, self.nxe, self.nze, a.shape))
        return a

    # -- set-up
    def set_stream(self, cuda_stream):
        self._ck(self.L.fdw_set_stream(self.h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self.L.fdw_sync(self.h))

    def set_v2(self, v2):
        self._ck(self.L.fdw_set_v2(self.h, self._grid(v2, "v2")))

    def set_wavelet(self, srce):
        srce = np.ascontiguousarray(srce, np.float32)
        self._ck(self.L.fdw_set_wavelet(self.h, srce, len(srce)))

    def set_source(self, sx, sz, kind=SRC_POINT):
        self._ck(self.L.fdw_set_source(self.h, sx, sz, kind))

    # -- fields
    def zero(self, pair=0):
        self._ck(self.L.fdw_fields_zero(self.h, pair))

    def upload(self, newest, older, pair=0):
        self._ck(self.L.fdw_fields_upload(self.h, pair, self._grid(newest, "newest"), self._grid(older, "older")))

    def download(self, pair=0, out_newest=None, out_older=None):
        n = np.zeros((self.nxe, self.nze), np.float32) if out_newest is None else out_newest
        o = np.zeros((self.nxe, self.nze), np.float32) if out_older is None else out_older
        self._ck(self.L.fdw_fields_download(self.h, pair, _opt(n), _opt(o)))
        return n, o

    def advance(self, it0, nsteps):
        self._ck(self.L.fdw_advance(self.h, it0, nsteps))

    def propagate(self, newest, older, it0, nsteps):
        """host fields in, nsteps levels, host fields out (in place)."""
        assert newest.dtype == np.float32 and older.dtype == np.float32
        assert newest.flags["C_CONTIGUOUS"] and older.flags["C_CONTIGUOUS"]
        assert newest.shape == (self.nxe, self.nze) and older.shape == (self.nxe, self.nze)
        self._ck(self.L.fdw_propagate(self.h, newest, older, it0, nsteps))

    # -- pipelines
    def forward(self, sx, sz, download=True):
        """fd_forward (fd-code.cu:247-288) -> (P, PP) = (older, newest)."""
        if not download:
            self._ck(self.L.fdw_forward(self.h, sx, sz, None, None))
            return None
        P = np.zeros((self.nxe, self.nze), np.float32)
        PP = np.zeros_like(P)
        self._ck(self.L.fdw_forward(self.h, sx, sz, _opt(P), _opt(PP)))
        return P, PP

    def backward(self, dobs, gz, P=None, PP=None):
        """fd_back (fd-code.cu:290-341) -> imloc [nx][nz]."""
        dobs = np.ascontiguousarray(dobs, np.float32)
        assert dobs.size == self.nx * self.nt
        im = np.zeros((self.nx, self.nz), np.float32)
        if P is not None:
            P, PP = self._grid(P, "P"), self._grid(PP, "PP")
        self._ck(self.L.fdw_backward(self.h, _opt(P), _opt(PP), dobs, gz, im))
        return im

    def model_shot(self, sx, sz, gz):
        """one shot of mod_main (mod_main.cpp:141-169) -> data [nx][nt]."""
        data = np.zeros((self.nx, self.nt), np.float32)
        self._ck(self.L.fdw_model_shot(self.h, sx, sz, gz, data))
        return data

    def rtm_shot_cpu(self, sx, sz, gz, dobs_all, is_=0):
        """one shot of rtm_main (rtm_main.cpp:158-240) -> imloc [nx][nz]."""
        dobs_all = np.ascontiguousarray(dobs_all, np.float32)
        ns = dobs_all.size // (self.nx * self.nt) if self.nt > 0 else 1
        im = np.zeros((self.nx, self.nz), np.float32)
        self._ck(self.L.fdw_rtm_shot_cpu(self.h, sx, sz, gz, dobs_all.reshape(-1), ns, is_, im))
        return im

    def shot_end(self, out):
        """host result of the phase opened by shot_phase_device: traces [nx][nt] or image [nx][nz]"""
        self._ck(self.L.fdw_shot_end(self.h, out))
        return out

    # -- shot loop without host round trips (fdw_v2_stage .. fdw_stack_download)
    def v2_stage(self, v2):
        """upload + premultiply the NEXT shot's squared velocity into the second buffer on the copy stream
        (asynchronous; pass pinned memory and keep it alive until v2_commit)"""
        self._ck(self.L.fdw_v2_stage(self.h, self._grid(v2, "v2")))

    def v2_commit(self):
        self._ck(self.L.fdw_v2_commit(self.h))

    def backward_device(self, dobs, gz):
        """fd_back on the levels forward(download=False) left on the device; the image stays there"""
        dobs = np.ascontiguousarray(dobs, np.float32)
        assert dobs.size == self.nx * self.nt
        self._ck(self.L.fdw_backward_device(self.h, dobs.reshape(-1), gz))

    def stack_zero(self):
        self._ck(self.L.fdw_stack_zero(self.h))

    def stack_add(self):
        self._ck(self.L.fdw_stack_add(self.h))

    def stack_download(self):
        im = np.zeros((self.nx, self.nz), np.float32)
        self._ck(self.L.fdw_stack_download(self.h, im))
        return im

    def stack_devptr(self):
        ptr, pitch, rows = C.c_void_p(), C.c_longlong(), C.c_int()
        self._ck(self.L.fdw_stack_devptr(self.h, C.byref(ptr), C.byref(pitch), C.byref(rows)))
        return ptr.value, pitch.value, rows.value

    # -- benchmarks / plumbing
    def devinfo(self):
        d = _lib.DevInfo()
        self._ck(self.L.fdw_devinfo_get(self.h, C.byref(d)))
        return d

    def mark_begin(self):
        self._ck(self.L.fdw_mark_begin(self.h))

    def mark_end(self):
        ms = C.c_float()
        self._ck(self.L.fdw_mark_end(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return self.L.fdw_launch_count(self.h)

    def graph_replays(self):
        return self.L.fdw_counter(self.h, _lib.COUNTER_GRAPH_REPLAYS)

    def persist_launches(self):
        return self.L.fdw_counter(self.h, _lib.COUNTER_PERSIST_LAUNCHES)

    def tile_launches(self):
        return self.L.fdw_counter(self.h, _lib.COUNTER_TILE_LAUNCHES)

    def pslab_launches(self):
        return self.L.fdw_counter(self.h, _lib.COUNTER_PSLAB_LAUNCHES)

    def shot_phase_device(self, phase, sx, sz, gz, dobs_all=None, is_=0):
        """one phase of a CPU-family shot left on the device (fdw_shot_begin + fdw_shot_run, asynchronous):
        bracket with mark_begin()/mark_end() for the device time of the phase's level loop"""
        ns, ptr = 1, None
        if dobs_all is not None:
            dobs_all = np.ascontiguousarray(dobs_all, np.float32)
            ns = dobs_all.size // (self.nx * self.nt)
            ptr = dobs_all.ctypes.data_as(C.c_void_p)
        self._ck(self.L.fdw_shot_begin(self.h, phase, sx, sz, gz, ptr, ns, is_))
        self._ck(self.L.fdw_shot_run(self.h))

    def laplacian_device(self):
        self._ck(self.L.fdw_laplacian_device(self.h))


def stencil(p, order=8, dx=10.0, dz=10.0, device=0, lib=None):
    """The stencil program's kernel (fd-source-code.cu:325): Laplacian of an
    [nxe][nze] snapshot, ring of width order/2 zero."""
    L = lib if lib is not None else _lib.load()
    p = np.ascontiguousarray(p, np.float32)
    out = np.empty_like(p)
    _lib.check(L, L.fdw_stencil(order, p.shape[0], p.shape[1], dx, dz, p, out, device))
    return out


def image_laplacian(img, dx=10.0, dz=10.0, device=0, lib=None):
    """the image post-filter of cuda_reference_RTM/models/3lay_mod/laplace.f90:24-28 on the GPU: 2nd-order Laplacian
    of an [nx][nz] image, zero on the outermost ring"""
    L = lib if lib is not None else _lib.load()
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty_like(img)
    _lib.check(L, L.fdw_image_laplacian(img.shape[0], img.shape[1], dx, dz, img, out, device))
    return out
