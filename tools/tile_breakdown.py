"""GPU box: where does a level of the shared-memory tile kernel go?  Times fd_forward on the shipped-model sizes
with parts of the level switched off (FDW_TILE_DBG: 1 no barrier, 2 no update, 4 no ring load -- results are then
wrong, this is a timing experiment), one subprocess per setting."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, numpy as np
sys.path.insert(0, %r)
import parallel_finite_difference_computation_b200 as fdw
nx, nz, nb, nt = [int(x) for x in sys.argv[1:5]]
nxe, nze = nx + 2 * nb, nz + 2 * nb
v2 = np.full((nxe, nze), np.float32(2500.0) ** 2, np.float32)
dobs = np.zeros((nx, nt), np.float32)
with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU, taper=fdw.TAPER_TOP,
                compat_extents=True, nt=nt) as w:
    w.set_v2(v2); w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 20.0, fdw.FAMILY_GPU))
    for rep in range(3):
        w.mark_begin(); w.forward(nb + 10, nb, download=False); f = w.mark_end()
    t0 = time.perf_counter(); w.backward(dobs, nb); b = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); w.forward(nb + 10, nb, download=False); w.backward(dobs, nb); b = (time.perf_counter() - t0) * 1e3 - f
    print("RESULT %%.3f %%.3f %%d" %% (f / nt * 1e3, b / nt * 1e3, w.tile_launches()))
''' % ROOT
cases = {"new_mod": (315, 195, 50, 1700), "marmousi": (369, 375, 40, 3004)}
for name, dims in cases.items():
    for dbg in (0, 8, 1, 2, 4, 3, 5, 6, 7):
        env = dict(os.environ, FDW_TILE_DBG=str(dbg), FDW_TILE_VERBOSE="1" if dbg == 0 else "")
        if dbg == 0:
            env.pop("FDW_TILE_DBG")
        if not env["FDW_TILE_VERBOSE"]:
            env.pop("FDW_TILE_VERBOSE")
        r = subprocess.run([sys.executable, "-c", CHILD] + [str(x) for x in dims], capture_output=True, text=True, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        plan = [l for l in r.stderr.splitlines() if "tile plan" in l][:1]
        print(name, "dbg", dbg, "(no:%s%s%s)" % (" barrier" if dbg & 1 else "", " update" if dbg & 2 else "", " ring" if dbg & 4 else ""),
              ("(counter barrier) " if dbg & 8 else "") + "fwd us/level, bwd us/level, tile launches:", line[0][7:] if line else r.stderr[-300:], plan[0] if plan else "", flush=True)
