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
cases = {"new_mod": (315, 195, 50, 1700), "3lay_mod": (151, 151, 40, 1001), "marmousi": (369, 375, 40, 3004)}
settings = [("flag-in-data halo (default)", {}),
            ("neighbour flags + ring loads (FDW_TILE_LL=0)", {"FDW_TILE_LL": "0"}),
            ("counter barrier + ring loads (FDW_TILE_LL=0, dbg 8)", {"FDW_TILE_LL": "0", "FDW_TILE_DBG": "8"}),
            ("flags, no barrier (dbg 1)", {"FDW_TILE_LL": "0", "FDW_TILE_DBG": "1"}),
            ("flags, no update (dbg 2)", {"FDW_TILE_LL": "0", "FDW_TILE_DBG": "2"}),
            ("flags, no ring load (dbg 4)", {"FDW_TILE_LL": "0", "FDW_TILE_DBG": "4"}),
            ("L2-resident persistent kernel of round 1 (FDW_TILE=0)", {"FDW_TILE": "0"})]
for name, dims in cases.items():
    for label, extra in settings:
        env = dict(os.environ, **extra)
        if not extra:
            env["FDW_TILE_VERBOSE"] = "1"
        r = subprocess.run([sys.executable, "-c", CHILD] + [str(x) for x in dims], capture_output=True, text=True, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
        plan = [l for l in r.stderr.splitlines() if "tile plan" in l][:1]
        print("%-9s %-52s fwd us/level, bwd us/level, tile launches: %s %s" % (name, label, line[0][7:] if line else r.stderr[-300:],
                                                                            plan[0] if plan else ""), flush=True)
