#!/usr/bin/env python
"""bench_configs.py -- the BASELINE.json configs beside the headline one, as device-timed sub-records.

bench.py imports this module and puts the records into the `configs` object of its JSON line (so that the
driver-run bench sees every config); run stand-alone it prints one JSON object per config:

    python bench_configs.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... bench_configs.py   # N GPUs (C4 shot-parallel, C5 domain-divided)

  C1  stand-alone Laplacian (the stencil program's kernel) at 16384^2: 8 B/point
  C2  single-shot RTM (GPU-family algorithm) on the shipped-model sizes: us per level forward / backward,
      beside the reference's own CUDA fd_forward / fd_back (oracle/_ref, same process) when present
  C4  multi-shot RTM on the 8192 x 4096 model, nt = 4000: 8 shots per GPU (64 shots on 8 GPUs), shot-parallel,
      image stack reduced across ranks
  C5  mod_main + rtm_main algorithm (recipe C, forward history resident in HBM) on the 8192 x 2048 model:
      one GPU with nt cut to what 180 GB holds, domain-divided (slab decomposition, peer-memory halo) with the
      stated nt = 3000 on N > 1
  gpu_reference  the reference's CUDA kernels (fd_forward of cuda_reference_RTM rebuilt for sm_100) at 8192^2
  parity_n_vs_1  N > 1: the slab-decomposed propagator against one GPU, bit for bit

All times are CUDA events on the library's stream (fdw_mark_begin / fdw_mark_end) unless the record says wall.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import parallel_finite_difference_computation_b200 as fdw  # noqa: E402
from parallel_finite_difference_computation_b200 import distributed as D  # noqa: E402


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def layered(nxe, nze):
    ve = np.empty((nxe, nze), np.float32)
    ve[:, : nze // 3] = 2000.0
    ve[:, nze // 3: 2 * nze // 3] = 3000.0
    ve[:, 2 * nze // 3:] = 4000.0
    return ve * ve


class quiet_stdout:
    """the reference's CUDA library printf()s its progress ("* it = 100 / 1700") on the C stdout: send file
    descriptor 1 to /dev/null while its code runs so that bench.py's stdout stays ONE JSON line"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)
        return self

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def _roof(gpts, bytes_per_point):
    gbs = gpts * bytes_per_point
    return {"bytes_per_point": bytes_per_point, "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / hbm_peak()}


# ------------------------------------------------------------------------------------------------- C1
def c1_laplacian(wave, reps=20):
    """`wave`: any resident Wave2D of the size wanted (bench.py passes its 16384^2 propagator: the sweep reads
    the newest level and overwrites the older one, which the caller no longer needs)."""
    n_pts = float(wave.nxe) * wave.nze
    for _ in range(3):
        wave.laplacian_device()
    wave.sync()
    wave.mark_begin()
    for _ in range(reps):
        wave.laplacian_device()
    ms = wave.mark_end() / reps
    gpts = n_pts / (ms * 1e-3) / 1e9
    rec = {"config": "C1 stand-alone Laplacian, %d x %d (kernel of the stencil program, recipe G exact, ring 0)"
                     % (wave.nxe, wave.nze), "ms_per_sweep": ms, "value": gpts, "unit": "Gpts/s", "timing": "cuda events"}
    rec.update(_roof(gpts, 8))
    return rec


# ------------------------------------------------------------------------------------------------- C2
SHIPPED = {"new_mod": (315, 195, 50, 1700, 20.0), "3lay_mod": (151, 151, 40, 1001, 30.0), "marmousi": (369, 375, 40, 3004, 6.5)}


def c2_small_models(device, with_reference=True):
    """one RTM shot (fd_forward + fd_back of fd-code.cu:247-341) on the extended grids of the shipped models"""
    out = {"config": "C2 single-shot RTM (GPU-family algorithm, compat extents) at the shipped-model sizes; "
                     "us per time level, device time", "models": {}}
    ref = None
    if with_reference:
        try:
            from oracle import ref as R
            if R.available("libref_gpufam.so"):
                ref = R.GpuFam()
        except Exception:
            ref = None
    for name, (nx, nz, nb, nt, fpeak) in SHIPPED.items():
        nxe, nze = nx + 2 * nb, nz + 2 * nb
        rng = np.random.default_rng(1)
        v2 = np.full((nxe, nze), np.float32(2500.0) ** 2, np.float32)
        dobs = rng.standard_normal((nx, nt)).astype(np.float32)
        srce = fdw.host.ricker_wavelet(nt, 0.001, fpeak, fdw.FAMILY_GPU)
        with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU,
                        taper=fdw.TAPER_TOP, compat_extents=True, nt=nt, device=device) as w:
            w.set_v2(v2)
            w.set_wavelet(srce)
            best_f = best_b = 1e30
            for _ in range(3):
                l0 = w.launch_count()
                w.mark_begin()
                w.forward(nb + 10, nb, download=False)
                best_f = min(best_f, w.mark_end())
                w.mark_begin()
                w.backward(dobs, nb)  # incl. the trace upload and the image download of this small model
                best_b = min(best_b, w.mark_end())
                launches = w.launch_count() - l0
            tiles = w.tile_launches()
        r = {"grid": [nxe, nze], "nt": nt, "forward_us_per_level": best_f / nt * 1e3,
             "backward_us_per_level": best_b / nt * 1e3, "launches_per_shot": launches,
             "kernel": "shared-memory tile kernel (one launch per phase)" if tiles else "per-level launches",
             "forward_gpts": nxe * nze * nt / (best_f * 1e-3) / 1e9, "backward_gpts": 2.0 * nxe * nze * nt / (best_b * 1e-3) / 1e9}
        if ref is not None:
            try:
                with quiet_stdout():
                    ref.fd_init(8, nxe, nze, nb, nb, nt, 1, 0.75, 10.0, 10.0, 0.001)
                    tf = tb = 1e30
                    for _ in range(2):
                        z = lambda: np.zeros((nxe, nze), np.float32)
                        P, PP = z(), z()
                        t0 = time.perf_counter()
                        ref.fd_forward(8, P, PP, v2, nt, 0, nb, [nb + 10], srce)
                        t1 = time.perf_counter()
                        im = np.zeros((nx, nz), np.float32)
                        ref.fd_back(8, z(), z(), z(), z(), v2, nt, 0, nb, nb, np.stack([P, PP]), im, dobs.reshape(1, -1).copy())
                        t2 = time.perf_counter()
                        tf, tb = min(tf, t1 - t0), min(tb, t2 - t1)
                r.update(reference_cuda_forward_us_per_level=tf / nt * 1e6, reference_cuda_backward_us_per_level=tb / nt * 1e6,
                         speedup_forward=tf * 1e3 / best_f, speedup_backward=tb * 1e3 / best_b,
                         reference_note="fd_forward / fd_back of cuda_reference_RTM/src/fd-code.cu rebuilt for sm_100 "
                                        "(oracle/_ref), host wall clock around the call like ours incl. its own transfers")
            except Exception as e:  # the reference is optional here
                r["reference_error"] = str(e)[:200]
        out["models"][name] = r
    return out


# ------------------------------------------------------------------------------------------------- C4
def c4_rtm_shots(device, rank, world, shots_per_gpu=None, nt=None):
    nx, nz, nb = 8192, 4096, 40
    nt = int(os.environ.get("FDW_C4_NT", nt or 4000))
    spg = int(os.environ.get("FDW_C4_SHOTS_PER_GPU", shots_per_gpu or 8))
    ns = spg * world
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    v2 = layered(nxe, nze)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_GPU)
    shots = D.shot_partition(ns, world, rank, contiguous=True)
    pipe = D.ShotPipeline(nx, nz, nb, nb, 10.0, 10.0, 0.001, nt=nt, device=device, order=8, fac=0.75)
    pipe.set_wavelet(srce)
    dobs = pipe.pinned((nx, nt))  # zeros: pure-throughput run
    v2p = pipe.pinned((nxe, nze))
    v2p[:] = v2
    # one warm-up shot, then per-phase device times of one shot
    pipe.run_shots([0], lambda k: v2p, lambda k: dobs, lambda k: 64, nb, nb)
    times = pipe.time_one_shot(v2p, dobs, 64, nb, nb)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.perf_counter()
    img = pipe.run_shots(shots, lambda k: v2p, lambda k: dobs, lambda k: 64 + 126 * k, nb, nb,
                         stack="allreduce" if world > 1 else "device")
    dt = time.perf_counter() - t0
    pipe.close()
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    upd = 3.0 * nt * nxe * nze * ns
    kern = times["forward_ms"] + times["backward_ms"]
    rec = {"config": "C4 multi-shot RTM, 8192 x 4096 model (+40 border), GPU-family algorithm, %d shots x nt=%d on %d GPU(s), "
                     "shot-parallel (%d per GPU); velocity staged through pinned memory into a second buffer while the "
                     "previous shot runs, shot images stacked on the device, one image download (+ one all-reduce)"
                     % (ns, nt, world, spg),
           "value": upd / dt / 1e9, "unit": "Gpts/s", "seconds": dt, "updates": upd, "timing": "wall clock around the whole job, max over ranks",
           "one_shot_device_ms": times, "shot_time_over_kernel_time": (dt / max(len(shots), 1) * 1e3) / kern if kern > 0 else None,
           "forward_gpts": nt * nxe * nze / (times["forward_ms"] * 1e-3) / 1e9,
           "backward_gpts": 2.0 * nt * nxe * nze / (times["backward_ms"] * 1e-3) / 1e9,
           "image_checksum": float(np.abs(img).sum())}
    rec["forward_roofline"] = _roof(rec["forward_gpts"], 16)
    rec["backward_roofline"] = _roof(rec["backward_gpts"] / 2.0, 40)  # 40 B per point and level for the two updates + imaging
    return rec


# ------------------------------------------------------------------------------------------------- C5
def _c5_nt_that_fits(nx, pitch, want, device):
    import torch
    free, _ = torch.cuda.mem_get_info(device)
    per_level = nx * pitch * 4
    return int(max(50, min(want, 0.55 * free // per_level)))


def c5_cpu_family(device):
    """one GPU: the whole 8192 x 2048 model; nt as large as the forward history fits into HBM"""
    nx, nz, nb = 8192, 2048, 40
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    pitch = (nze + 4 + 31) // 32 * 32
    nt = int(os.environ.get("FDW_C5_NT", _c5_nt_that_fits(nx, pitch, 3000, device)))
    v2 = layered(nxe, nze)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_CPU)
    sx, sz, gz = nb + nx // 2, nb, nb
    out = {"config": "C5 mod_main + rtm_main algorithm (recipe C bit-exact), %d x %d model (+40 border), 1 GPU, nt=%d "
                     "(the stated nt=3000 needs 210 GB of forward history: it runs domain-divided on N>1)" % (nx, nz, nt),
           "nt": nt, "timing": "cuda events around the level loop of each phase"}
    pts = float(nxe) * nze
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU,
                    taper=fdw.TAPER_FOUR, device=device, nt=nt) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.shot_phase_device(fdw.PHASE_MODEL, sx, sz, gz)
        w.sync()
        w.mark_begin()
        w.shot_phase_device(fdw.PHASE_MODEL, sx, sz, gz)
        ms = w.mark_end()
        data = np.zeros((nx, nt), np.float32)
        w.shot_end(data)
        g = pts * nt / (ms * 1e-3) / 1e9
        out["mod_main"] = dict(value=g, unit="Gpts/s", us_per_level=ms / nt * 1e3, **_roof(g, 16))
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU,
                    taper=fdw.TAPER_TOP, device=device, nt=nt, history=True) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.mark_begin()
        w.shot_phase_device(fdw.PHASE_RTM_FWD, sx, sz, gz)
        msf = w.mark_end()
        w.mark_begin()
        w.shot_phase_device(fdw.PHASE_RTM_BWD, sx, sz, gz, data[None], 0)
        msb = w.mark_end()
        img = np.zeros((nx, nz), np.float32)
        w.shot_end(img)
        gf, gb = pts * nt / (msf * 1e-3) / 1e9, pts * nt / (msb * 1e-3) / 1e9
        frac_int = float(nx) * nz / pts
        out["rtm_main_forward_with_history"] = dict(value=gf, unit="Gpts/s", us_per_level=msf / nt * 1e3,
                                                    **_roof(gf, 16 + 4 * frac_int))
        out["rtm_main_backward_with_imaging"] = dict(value=gb, unit="Gpts/s", us_per_level=msb / nt * 1e3,
                                                     **_roof(gb, 16 + 12 * frac_int))
        out["history_GB"] = nt * nx * pitch * 4 / 1e9
        out["image_checksum"] = float(np.abs(img).sum())
    return out


def c5_domain_divided(device, rank, world):
    """BASELINE configs[4]: mod_main + rtm_main on the large synthetic model, domain-divided (slab decomposition
    along x, forward history sharded with the slabs, peer-memory halo exchange)"""
    import torch
    import torch.distributed as dist
    nx, nz, nb = int(os.environ.get("FDW_C5_NX", "8192")), int(os.environ.get("FDW_C5_NZ", "2048")), 40
    nt = int(os.environ.get("FDW_C5_NT", "3000"))
    halo = os.environ.get("FDW_BENCH_HALO", "p2p")
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    srce = fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_CPU)
    sx, sz, gz = nb + nx // 2, nb, nb
    kw = dict(order=8, fac=0.01, family=fdw.FAMILY_CPU, nt=nt, rank=rank, world=world, device=device, halo=halo)
    out = {"config": "C5 mod_main + rtm_main algorithm (recipe C bit-exact), %d x %d model (+40 border), nt=%d, "
                     "domain-divided over %d GPU(s), halo=%s" % (nx, nz, nt, world, halo),
           "timing": "wall clock between barriers around each shot phase incl. the result download, max over ranks"}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def timed(fn):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(); dist.barrier()
        return r, time.perf_counter() - t0

    pts = float(nxe) * nze
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, taper=fdw.TAPER_FOUR, **kw)
    sp.set_stream(stream.cuda_stream)
    x0, x1 = sp.slab
    v2 = layered(x1 - x0, nze)
    sp.set_v2_local(v2); sp.set_wavelet(srce)
    sp.model_shot(sx, sz, gz)
    rows, dt = timed(lambda: sp.model_shot(sx, sz, gz))
    g = pts * nt / dt / 1e9
    out["mod_main"] = dict(value=g, unit="Gpts/s", us_per_level=dt / nt * 1e6, per_gpu=_roof(g / world, 16),
                           level_loop_us_per_level_device_rank0=sp.levels_ms.get(fdw.PHASE_MODEL, 0.0) / nt * 1e3)
    data = sp.gather_rows(rows)
    sp.close()
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, taper=fdw.TAPER_TOP, history=True, **kw)
    sp.set_stream(stream.cuda_stream)
    sp.set_v2_local(v2); sp.set_wavelet(srce)
    sp.rtm_shot_cpu(sx, sz, gz, data[None], 0)
    _, dt = timed(lambda: sp.rtm_shot_cpu(sx, sz, gz, data[None], 0))
    g = 2 * pts * nt / dt / 1e9
    out["rtm_main"] = dict(value=g, unit="Gpts/s", us_per_level=dt / (2 * nt) * 1e6, per_gpu=_roof(g / world, 22),
                           forward_level_loop_us_per_level_device_rank0=sp.levels_ms.get(fdw.PHASE_RTM_FWD, 0.0) / nt * 1e3,
                           backward_level_loop_us_per_level_device_rank0=sp.levels_ms.get(fdw.PHASE_RTM_BWD, 0.0) / nt * 1e3)
    out["persistent_slab_launches_rank0"] = int(sp.w.pslab_launches())
    out["history_GB_per_gpu"] = nt * max(sp.owned_interior()[1], 0) * ((nze + 4 + 31) // 32 * 32) * 4 / 1e9
    out["halo"] = "p2p" if sp.p2p else "nccl"
    sp.close()
    return out


# ------------------------------------------------------------------------------------------------- reference CUDA kernels
def gpu_reference(n=8192, device=0):
    """the reference's own CUDA kernels (kernel_tapper, kernel_lap, kernel_time, kernel_src of fd-code.cu:53-122,
    rebuilt for sm_100 with the reference flags) at a headline-class size: time per level = difference of two
    fd_forward calls of different nt (their host transfers cancel)"""
    try:
        from oracle import ref as R
        if not R.available("libref_gpufam.so"):
            return {"unavailable": "oracle/_ref/libref_gpufam.so not present"}
        g = R.GpuFam()
        nb, nt_a, nt_b = 40, 4, 24
        with quiet_stdout():
            g.fd_init(8, n, n, nb, nb, nt_b, 1, 0.75, 10.0, 10.0, 0.001)
        v2 = layered(n, n)
        srce = fdw.host.ricker_wavelet(nt_b, 0.001, 20.0, fdw.FAMILY_GPU)
        P, PP = np.zeros((n, n), np.float32), np.zeros((n, n), np.float32)
        ts = {}
        with quiet_stdout():
            g.fd_forward(8, P, PP, v2, nt_a, 0, nb, [n // 2], srce)  # warm-up
            for nt in (nt_a, nt_b, nt_a, nt_b):
                t0 = time.perf_counter()
                g.fd_forward(8, P, PP, v2, nt, 0, nb, [n // 2], srce)
                ts[nt] = min(ts.get(nt, 1e30), time.perf_counter() - t0)
        per_level = (ts[nt_b] - ts[nt_a]) / (nt_b - nt_a)
        gpts = float(n) * n / per_level / 1e9
        return {"what": "reference CUDA kernels (cuda_reference_RTM fd_forward, sm_100 rebuild), %d x %d, per level = "
                        "(t[nt=%d] - t[nt=%d]) / %d" % (n, n, nt_b, nt_a, nt_b - nt_a),
                "value": gpts, "unit": "Gpts/s", "ms_per_level": per_level * 1e3, "reference_bytes_per_point_min": 28}
    except Exception as e:
        return {"unavailable": str(e)[:200]}


# ------------------------------------------------------------------------------------------------- N vs 1 parity
def parity_n_vs_1(rank, world, device, halo="p2p", n=4096, nt=40):
    """the slab-decomposed propagator (N ranks, peer-memory halo, CUDA-graph level loop) against rank 0 running the
    same job alone: "bitwise" or "MISMATCH" (decomposition does not change the per-point arithmetic)"""
    import torch
    import torch.distributed as dist
    nb = 40
    nx = nz = n - 2 * nb
    rng = np.random.default_rng(7)
    ve = np.empty((n, n), np.float32)
    ve[:, : n // 3] = 2000.0
    ve[:, n // 3:] = 3500.0
    v2 = ve * ve
    a = rng.standard_normal((n, n), dtype=np.float32)
    b = rng.standard_normal((n, n), dtype=np.float32)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 25.0, fdw.FAMILY_GPU)
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, rank=rank, world=world, device=device, order=8,
                          fac=0.75, family=fdw.FAMILY_GPU, taper=fdw.TAPER_TOP, nt=nt, halo=halo)
    sp.set_stream(torch.cuda.current_stream().cuda_stream)
    x0, x1 = sp.slab
    sp.set_v2_local(v2[x0:x1])
    sp.set_wavelet(srce)
    sx = D.slab_rows(n, world, 0)[1] - 2  # next to a slab cut
    sp.set_source(sx, nb)
    na, nb_ = a[x0:x1].copy(), b[x0:x1].copy()
    sp.propagate_local(na, nb_, 0, nt)
    torch.cuda.synchronize()
    used = "p2p" if sp.p2p else "nccl"
    replays, pslabs = sp.w.graph_replays(), sp.w.pslab_launches()
    sp.close()
    parts = [None] * world
    dist.all_gather_object(parts, (na, nb_))
    verdict = None
    if rank == 0:
        with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU,
                        taper=fdw.TAPER_TOP, device=device, nt=nt) as w:
            w.set_v2(v2)
            w.set_wavelet(srce)
            w.set_source(sx, nb)
            w.propagate(a, b, 0, nt)
        newest = np.concatenate([p[0] for p in parts])
        older = np.concatenate([p[1] for p in parts])
        ok = np.array_equal(newest.view(np.uint32), a.view(np.uint32)) and np.array_equal(older.view(np.uint32), b.view(np.uint32))
        verdict = {"result": "bitwise" if ok else "MISMATCH", "grid": [n, n], "levels": nt, "slabs": world, "halo": used,
                   "graph_replays_rank0": int(replays), "persistent_slab_launches_rank0": int(pslabs)}
    return verdict


# ------------------------------------------------------------------------------------------------- DRAM traffic
def measure_traffic(timeout=240):
    """DRAM bytes of the headline step kernel measured now, on this GPU: a short run of tools/prof_step.py under
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (a byte count, not a timing).  None when ncu cannot
    profile here."""
    ncu = "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not installed"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
           "-k", "regex:k_step", "-c", "12", sys.executable, os.path.join(ROOT, "tools", "prof_step.py")]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    except Exception as e:
        return None, "ncu failed: %s" % str(e)[:100]
    import csv
    rows = [row for row in csv.reader(r.stdout.splitlines()) if len(row) > 5]
    if not rows:
        return None, "ncu produced no rows (%s)" % (r.stderr.strip().splitlines()[-1][:120] if r.stderr.strip() else "no output")
    hdr = rows[0]
    try:
        i_name, i_metric, i_unit, i_val, i_id = (hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"),
                                                 hdr.index("Metric Value"), hdr.index("ID"))
    except ValueError:
        return None, "unexpected ncu csv header"
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    per_launch = {}
    for row in rows[1:]:
        if "k_step<" not in row[i_name] or "sponge" in row[i_name]:
            continue
        v = float(row[i_val].replace(",", "")) * scale.get(row[i_unit], 1.0)
        per_launch[row[i_id]] = per_launch.get(row[i_id], 0.0) + v
    big = [v for v in per_launch.values() if v > 1e9]  # the bulk launches (the strips move megabytes)
    if not big:
        return None, "no bulk launch captured"
    return float(np.median(big)), "ncu dram__bytes_read.sum + dram__bytes_write.sum, median of %d bulk launches, this run" % len(big)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    which = os.environ.get("FDW_CONFIGS", "c1,c2,c4,c5,ref").split(",")
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    res = []
    if world == 1:
        if "c1" in which:
            with fdw.Wave2D(16384, 16384, 0, 0, 10.0, 10.0, 0.001, order=8, taper=fdw.TAPER_NONE, device=device) as w:
                w.zero()
                res.append(c1_laplacian(w))
        if "c2" in which:
            res.append(c2_small_models(device))
        if "c5" in which:
            res.append(c5_cpu_family(device))
        if "ref" in which:
            res.append(gpu_reference(device=device))
    else:
        if "c5" in which:
            res.append(c5_domain_divided(device, rank, world))
        res.append(parity_n_vs_1(rank, world, device))
    if "c4" in which:
        res.append(c4_rtm_shots(device, rank, world))
    if rank == 0:
        for r in res:
            print(json.dumps(r))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
