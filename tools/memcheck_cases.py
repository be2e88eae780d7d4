"""A handful of small parity cases for `compute-sanitizer --tool memcheck` (out-of-bounds reads do not show in
results: guard rows / pad columns absorb them by design, so check that nothing reaches past the allocations):
orders 8 and 16, strips on load (forked / multi-rectangle, folded CTAs), in-place sponge pass, stencil, one shot of
each family.  Run:  compute-sanitizer --tool memcheck python tools/memcheck_cases.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_cases as PC  # noqa: E402
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402

lib = fdw.load()
for inplace in ("1", "0"):
    os.environ.update(FDW_PERSIST_LIMIT="0", FDW_SMALL_GRID_LIMIT="0", FDW_FORK_LIMIT="0", FDW_TILE="0",
                      FDW_SPONGE_INPLACE=inplace)
    for order in (8, 16):
        PC.case_advance(lib, fdw.FAMILY_CPU, fdw.RECIPE_C, fdw.TAPER_FOUR, order=order, nx=70, nz=300, nxb=12, nzb=10,
                        nt=5, src_kind=fdw.SRC_GAUSS7)
        PC.case_advance(lib, fdw.FAMILY_GPU, fdw.RECIPE_G, fdw.TAPER_TOP, order=order, nx=37, nz=29, nxb=9, nzb=8, nt=5,
                        compat=(order == 8))
        PC.case_stencil(lib, order, (61, 47))
    print("sponge in place = %s: ok" % inplace, flush=True)
for k in ("FDW_PERSIST_LIMIT", "FDW_SMALL_GRID_LIMIT", "FDW_FORK_LIMIT", "FDW_TILE", "FDW_SPONGE_INPLACE"):
    os.environ.pop(k, None)
PC.case_gpu_rtm(lib, nt=20)          # tile kernels
PC.case_mod_shot(lib, nt=20)
PC.case_rtm_shot_cpu(lib, nt=16)
PC.case_gpu_rtm(lib, nt=12, order=12, compat=False)
print("shots: ok")
