"""B200-native 2-D acoustic finite-difference propagation (hand-written sm_100a
CUDA behind a C ABI).  Drop-in for the hot path of
FernandoSchett/parallel_finite_difference_computation."""
from ._lib import (FAMILY_CPU, FAMILY_GPU, PHASE_MODEL, PHASE_PLAIN, PHASE_RTM_BWD, PHASE_RTM_FWD, RECIPE_C, RECIPE_FAST,
                   RECIPE_G, SRC_GAUSS7, SRC_POINT, TAPER_FOUR, TAPER_NONE, TAPER_TOP, FdwError, load)
from . import host
from .propagator import Wave2D, image_laplacian, stencil

__all__ = ["Wave2D", "stencil", "image_laplacian", "host", "load", "FdwError", "FAMILY_CPU", "FAMILY_GPU", "RECIPE_C", "RECIPE_FAST",
           "RECIPE_G", "SRC_GAUSS7", "SRC_POINT", "TAPER_FOUR", "TAPER_NONE", "TAPER_TOP", "PHASE_PLAIN", "PHASE_MODEL",
           "PHASE_RTM_FWD", "PHASE_RTM_BWD"]
