"""ctypes binding of libfdwave.so (C ABI: include/fdwave.h).

The library is built in-tree by csrc/Makefile (`python __graft_entry__.py` or
`make -C parallel_finite_difference_computation_b200/csrc`).  There is no CPU
path: load() raises when the shared object is missing, and every compute
entry point raises FdwError when no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDW_LIBFDWAVE") or os.path.join(_HERE, "libfdwave.so")  # (override: A/B of two builds in tools/)

FAMILY_GPU, FAMILY_CPU = 0, 1
RECIPE_G, RECIPE_C, RECIPE_FAST = 0, 1, 2
TAPER_NONE, TAPER_TOP, TAPER_FOUR = 0, 1, 2
SRC_POINT, SRC_GAUSS7 = 0, 1
PHASE_PLAIN, PHASE_MODEL, PHASE_RTM_FWD, PHASE_RTM_BWD = 0, 1, 2, 3
COUNTER_LAUNCHES, COUNTER_GRAPH_REPLAYS, COUNTER_PERSIST_LAUNCHES, COUNTER_TILE_LAUNCHES, COUNTER_PSLAB_LAUNCHES = 0, 1, 2, 3, 4

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class FdwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libfdwave error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("nx", C.c_int), ("nz", C.c_int), ("nxb", C.c_int), ("nzb", C.c_int), ("order", C.c_int),
                ("dx", C.c_float), ("dz", C.c_float), ("dt", C.c_float), ("fac", C.c_float),
                ("family", C.c_int), ("recipe", C.c_int), ("taper", C.c_int), ("compat_extents", C.c_int),
                ("device", C.c_int), ("slab_x0", C.c_int), ("slab_x1", C.c_int), ("history", C.c_int),
                ("nt", C.c_int)]


class Input(C.Structure):
    _fields_ = [("tmpdir", C.c_char * 512), ("vpfile", C.c_char * 512), ("datfile", C.c_char * 512),
                ("vel_ext_file", C.c_char * 512),
                ("nz", C.c_int), ("nx", C.c_int), ("nt", C.c_int), ("ns", C.c_int), ("sz", C.c_int),
                ("fsx", C.c_int), ("ds", C.c_int), ("gz", C.c_int), ("order", C.c_int), ("nzb", C.c_int),
                ("nxb", C.c_int), ("iss", C.c_int), ("rnd", C.c_int),
                ("dz", C.c_float), ("dx", C.c_float), ("dt", C.c_float), ("fpeak", C.c_float), ("fac", C.c_float),
                ("has_datfile", C.c_int), ("has_vel_ext_file", C.c_int)]


class Halo(C.Structure):
    _fields_ = [("send_lo", C.c_void_p), ("send_hi", C.c_void_p), ("recv_lo", C.c_void_p), ("recv_hi", C.c_void_p),
                ("count", C.c_longlong)]


class DevInfo(C.Structure):
    _fields_ = [("newest", C.c_void_p), ("older", C.c_void_p), ("vdt", C.c_void_p), ("pitch", C.c_longlong),
                ("nloc", C.c_int), ("gx0", C.c_int), ("nxe", C.c_int), ("nze", C.c_int), ("guard", C.c_int),
                ("li0", C.c_int), ("nli", C.c_int)]


class PeerInfo(C.Structure):
    _fields_ = [("field", (C.c_ubyte * 64) * 4), ("flags", C.c_ubyte * 64), ("pitch", C.c_longlong),
                ("nloc", C.c_int), ("gx0", C.c_int), ("device", C.c_int), ("reserved", C.c_int)]


# every symbol include/fdwave.h declares: (restype, argtypes)
_optf32 = C.c_void_p  # optional float* (may be NULL)
SIGNATURES = {
    "fdw_calc_coefs": (C.c_int, [C.c_int, C.c_int, f32p]),
    "fdw_ricker_wavelet": (C.c_int, [C.c_int, C.c_float, C.c_float, C.c_int, f32p]),
    "fdw_taper_table": (C.c_int, [C.c_int, C.c_float, C.c_int, f32p]),
    "fdw_extendvel": (C.c_int, [C.c_int] * 4 + [f32p]),
    "fdw_extendvel_linear": (C.c_int, [C.c_int] * 4 + [f32p]),
    "fdw_ptsrc_weights": (C.c_int, [f32p]),
    "fdw_read_input_gpu": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(Input)]),
    "fdw_read_input_stencil": (C.c_int, [C.c_char_p, C.POINTER(Input)]),
    "fdw_read_input_cpu": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(Input)]),
    "fdw_read_floats": (C.c_longlong, [C.c_char_p, C.c_void_p, C.c_longlong]),
    "fdw_write_floats": (C.c_int, [C.c_char_p, C.c_void_p, C.c_longlong, C.c_int]),
    "fdw_image_stack_shot": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fdw_last_error": (C.c_char_p, []),
    "fdw_device_count": (C.c_int, []),
    "fdw_create": (C.c_int, [C.POINTER(Params), C.POINTER(C.c_void_p)]),
    "fdw_destroy": (None, [C.c_void_p]),
    "fdw_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fdw_sync": (C.c_int, [C.c_void_p]),
    "fdw_set_v2": (C.c_int, [C.c_void_p, f32p]),
    "fdw_set_wavelet": (C.c_int, [C.c_void_p, f32p, C.c_int]),
    "fdw_set_source": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "fdw_fields_zero": (C.c_int, [C.c_void_p, C.c_int]),
    "fdw_fields_upload": (C.c_int, [C.c_void_p, C.c_int, f32p, f32p]),
    "fdw_fields_download": (C.c_int, [C.c_void_p, C.c_int, _optf32, _optf32]),
    "fdw_advance": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "fdw_propagate": (C.c_int, [C.c_void_p, f32p, f32p, C.c_int, C.c_int]),
    "fdw_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _optf32, _optf32]),
    "fdw_backward": (C.c_int, [C.c_void_p, _optf32, _optf32, f32p, C.c_int, f32p]),
    "fdw_model_shot": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p]),
    "fdw_rtm_shot_cpu": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_int, f32p]),
    "fdw_stencil": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p, C.c_int]),
    "fdw_image_laplacian": (C.c_int, [C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p, C.c_int]),
    "fdw_shot_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _optf32, C.c_int, C.c_int]),
    "fdw_shot_end": (C.c_int, [C.c_void_p, f32p]),
    "fdw_step_begin": (C.c_int, [C.c_void_p, C.c_int]),
    "fdw_step_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "fdw_step_end": (C.c_int, [C.c_void_p]),
    "fdw_halo_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Halo)]),
    "fdw_set_v2_local": (C.c_int, [C.c_void_p, f32p]),
    "fdw_fields_upload_local": (C.c_int, [C.c_void_p, C.c_int, f32p, f32p]),
    "fdw_fields_download_local": (C.c_int, [C.c_void_p, C.c_int, _optf32, _optf32]),
    "fdw_fields_download_local_async": (C.c_int, [C.c_void_p, C.c_int, _optf32, _optf32]),
    "fdw_peer_export": (C.c_int, [C.c_void_p, C.POINTER(PeerInfo)]),
    "fdw_peer_attach": (C.c_int, [C.c_void_p, C.POINTER(PeerInfo), C.POINTER(PeerInfo)]),
    "fdw_peer_detach": (C.c_int, [C.c_void_p]),
    "fdw_peer_refresh": (C.c_int, [C.c_void_p]),
    "fdw_peer_levels": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "fdw_peer_fence": (C.c_int, [C.c_void_p]),
    "fdw_devinfo_get": (C.c_int, [C.c_void_p, C.POINTER(DevInfo)]),
    "fdw_mark_begin": (C.c_int, [C.c_void_p]),
    "fdw_mark_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "fdw_launch_count": (C.c_longlong, [C.c_void_p]),
    "fdw_laplacian_device": (C.c_int, [C.c_void_p]),
    "fdw_counter": (C.c_longlong, [C.c_void_p, C.c_int]),
    "fdw_v2_stage": (C.c_int, [C.c_void_p, f32p]),
    "fdw_v2_commit": (C.c_int, [C.c_void_p]),
    "fdw_backward_device": (C.c_int, [C.c_void_p, f32p, C.c_int]),
    "fdw_stack_zero": (C.c_int, [C.c_void_p]),
    "fdw_stack_add": (C.c_int, [C.c_void_p]),
    "fdw_stack_download": (C.c_int, [C.c_void_p, f32p]),
    "fdw_stack_devptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "fdw_shot_run": (C.c_int, [C.c_void_p]),
}


def bind(cdll):
    """attach argtypes/restypes for every declared symbol; raises if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)  # AttributeError = missing export
        fn.restype = res
        fn.argtypes = args
    return cdll


_LIB = None


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C parallel_finite_difference_computation_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
        _LIB = bind(C.CDLL(LIB_PATH))
    return _LIB


def check(lib, rc):
    if rc != 0:
        raise FdwError(rc, lib.fdw_last_error().decode(errors="replace"))
