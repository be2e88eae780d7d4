"""GPU box: BASELINE config 2 -- single-shot RTM on shipped-model sizes, ours (bin/rtm_code and the
API) vs the reference's own CUDA program (oracle/_ref/rtm_code_ref, sm_100 rebuild), same GPU."""
import os, sys, time, subprocess, tempfile, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_drivers as T
from oracle import ref as R
import parallel_finite_difference_computation_b200 as fdw

cases = {"new_mod": (315, 195, 50, 1700), "3lay_mod": (151, 151, 40, 1001), "marmousi": (369, 375, 40, 3004)}
out = {}
for name, (nx, nz, nb, nt) in cases.items():
    base = tempfile.mkdtemp()
    res = {}
    for tag, prog in (("ours", os.path.join(ROOT, "bin", "rtm_code")), ("reference_cuda", R.path("rtm_code_ref"))):
        if prog is None:
            continue
        d = os.path.join(base, tag); os.makedirs(d)
        T._write_rtm_case(d, nx, nz, nb, nt, 1, seed=3)
        best = 1e9
        for rep in range(2):
            t0 = time.perf_counter()
            pr = subprocess.run([prog, "./input.dat"], cwd=d, capture_output=True, check=True, text=True)
            best = min(best, time.perf_counter() - t0)
        if tag == "ours":
            res["ours_breakdown"] = pr.stderr.strip().splitlines()[-1]
        res[tag + "_wall_s"] = best
    # API-level timing of the device work only (forward + backward, levels resident in HBM)
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    rng = np.random.default_rng(1)
    v2 = np.full((nxe, nze), np.float32(2500.0) ** 2, np.float32)
    dobs = rng.standard_normal((nx, nt)).astype(np.float32)
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU,
                    taper=fdw.TAPER_TOP, compat_extents=True, nt=nt) as w:
        w.set_v2(v2); w.set_wavelet(fdw.host.ricker_wavelet(nt, 0.001, 20.0, fdw.FAMILY_GPU))
        for rep in range(2):
            l0 = w.launch_count(); t0 = time.perf_counter()
            w.forward(nb + 10, nb, download=False); w.sync(); t1 = time.perf_counter()
            w.backward(dobs, nb); t2 = time.perf_counter()
        res.update(api_forward_s=t1 - t0, api_backward_s=t2 - t1, launches=w.launch_count() - l0,
                   us_per_forward_level=(t1 - t0) / nt * 1e6, us_per_backward_level=(t2 - t1) / nt * 1e6,
                   gpts_forward=nxe * nze * nt / (t1 - t0) / 1e9, gpts_backward=2 * nxe * nze * nt / (t2 - t1) / 1e9)
    # the reference's own fd_forward + fd_back, in process (libref_gpufam.so): like ours, no context start-up
    if R.available("libref_gpufam.so"):
        g = R.GpuFam()
        g.fd_init(8, nxe, nze, nb, nb, nt, 1, 0.75, 10.0, 10.0, 0.001)
        srce = fdw.host.ricker_wavelet(nt, 0.001, 20.0, fdw.FAMILY_GPU)
        for rep in range(2):
            z = lambda: np.zeros((nxe, nze), np.float32)
            P, PP = z(), z()
            t0 = time.perf_counter()
            g.fd_forward(8, P, PP, v2, nt, 0, nb, [nb + 10], srce)
            t1 = time.perf_counter()
            im = np.zeros((nx, nz), np.float32)
            g.fd_back(8, z(), z(), z(), z(), v2, nt, 0, nb, nb, np.stack([P, PP]), im, dobs.reshape(1, -1).copy())
            t2 = time.perf_counter()
        res.update(ref_forward_s=t1 - t0, ref_backward_s=t2 - t1,
                   speedup_forward=(t1 - t0) / res["api_forward_s"], speedup_backward=(t2 - t1) / res["api_backward_s"])
    out[name] = res
    print(name, json.dumps(res), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "rtm_small.json"), "w"), indent=1)
