#!/usr/bin/env python
"""bench_configs.py -- the other BASELINE.json configs beside the headline one (bench.py).

    python bench_configs.py                       # C1, C4 (1 GPU), C5 on one GPU
    torchrun --nproc-per-node N bench_configs.py  # C4 shot-parallel on N GPUs

Prints one JSON object per config (rank 0).  Device time only unless stated;
CUDA events / synchronised wall clock around resident work.
  C1  stand-alone Laplacian (stencil program kernel) at 16384^2: GB/s at 8 B/point
  C4  multi-shot RTM (GPU-family algorithm) on an 8192 x 4096 model, shot-parallel,
      reduced nt per shot; updates = 3 * nt * nxe * nze per shot
  C5  mod_main / rtm_main algorithm (CPU-family recipe C, history resident in HBM)
"""
import json
import os
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


def c1_laplacian(device):
    n = 16384
    with fdw.Wave2D(n, n, 0, 0, 10.0, 10.0, 0.001, order=8, taper=fdw.TAPER_NONE, device=device) as w:
        rng = np.random.default_rng(20261018)
        a = rng.uniform(-1, 1, (n, n)).astype(np.float32)
        w.upload(a, a)
        for _ in range(3):
            w.laplacian_device()
        w.sync()
        w.mark_begin()
        reps = 20
        for _ in range(reps):
            w.laplacian_device()
        ms = w.mark_end() / reps
    gbs = 8.0 * n * n / (ms * 1e-3) / 1e9
    return {"config": "C1 stand-alone Laplacian 16384x16384 (stencil program kernel, recipe G exact)",
            "ms_per_sweep": ms, "gpts_per_s": n * n / (ms * 1e-3) / 1e9, "hbm_gbs_at_8B_per_point": gbs,
            "frac_of_measured_peak": gbs / hbm_peak()}


def c4_rtm_shots(device, rank, world):
    import torch  # noqa: F401  (imported here so that its start-up cost stays outside the timed region)
    nx, nz, nb, nt = 8192, 4096, 40, int(os.environ.get("FDW_C4_NT", "400"))
    ns = int(os.environ.get("FDW_C4_SHOTS", str(2 * world)))
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    v2 = layered(nxe, nze)
    dobs = np.zeros((nx, nt), np.float32)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_GPU)
    shots = D.shot_partition(ns, world, rank, contiguous=True)
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.75, family=fdw.FAMILY_GPU,
                    taper=fdw.TAPER_TOP, device=device, nt=nt) as w:
        w.set_wavelet(srce)
        w.set_v2(v2)
        w.forward(64, nb, download=False)  # warm-up
        w.backward(dobs, nb)
        # breakdown of one shot (device work vs host transfers through pageable memory)
        t0 = time.perf_counter(); w.set_v2(v2); t1 = time.perf_counter()
        w.forward(64, nb, download=False); w.sync(); t2 = time.perf_counter()
        w.backward(dobs, nb); t3 = time.perf_counter()
        breakdown = {"set_v2_s": t1 - t0, "forward_s": t2 - t1, "backward_incl_traces_up_image_down_s": t3 - t2,
                     "forward_gpts_per_s": nt * nxe * nze / (t2 - t1) / 1e9,
                     "backward_gpts_per_s": 2.0 * nt * nxe * nze / (t3 - t2) / 1e9}
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        t0 = time.perf_counter()
        img = D.migrate_shots_gpu_family(w, shots, lambda k: v2, lambda k: dobs, lambda k: 64 + 126 * k, nb, nb,
                                         stack="allreduce" if world > 1 else "chain")
        dt = time.perf_counter() - t0
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    upd = 3.0 * nt * nxe * nze * ns
    return {"config": "C4 multi-shot RTM 8192x4096 (+40 border), GPU-family algorithm, %d shots x %d steps on %d GPU(s), "
                      "shot-parallel; per shot: v2 upload, forward, backward+imaging, image download; final image "
                      "stack" % (ns, nt, world),
            "seconds": dt, "gpts_per_s": upd / dt / 1e9, "updates": upd, "one_shot_breakdown": breakdown}


def c5_cpu_family(device):
    nx, nz, nb, nt = 4096, 2048, 40, int(os.environ.get("FDW_C5_NT", "100"))
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    v2 = layered(nxe, nze)
    srce = fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_CPU)
    out = {"config": "C5 mod_main + rtm_main algorithm (recipe C bit-exact, history in HBM), %dx%d model, %d steps, 1 GPU"
                     % (nx, nz, nt)}
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU,
                    taper=fdw.TAPER_FOUR, device=device, nt=nt) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.model_shot(nb + 100, nb, nb)
        t0 = time.perf_counter()
        data = w.model_shot(nb + 100, nb, nb)
        dt = time.perf_counter() - t0
        out["mod_main_gpts_per_s"] = nt * nxe * nze / dt / 1e9
    with fdw.Wave2D(nx, nz, nb, nb, 10.0, 10.0, 0.001, order=8, fac=0.01, family=fdw.FAMILY_CPU,
                    taper=fdw.TAPER_TOP, device=device, nt=nt, history=True) as w:
        w.set_v2(v2)
        w.set_wavelet(srce)
        w.rtm_shot_cpu(nb + 100, nb, nb, data[None], 0)
        t0 = time.perf_counter()
        w.rtm_shot_cpu(nb + 100, nb, nb, data[None], 0)
        dt = time.perf_counter() - t0
        out["rtm_main_gpts_per_s"] = 2 * nt * nxe * nze / dt / 1e9
        out["history_GB"] = nt * nx * ((nze + 4 + 31) // 32 * 32) * 4 / 1e9
    return out


def c5_domain_divided(device, rank, world):
    """BASELINE configs[4]: mod_main + rtm_main on a large synthetic model, domain-divided (slab
    decomposition along x, forward history sharded with the slabs, halo exchange FDW_BENCH_HALO)"""
    import torch
    import torch.distributed as dist
    nx, nz, nb = int(os.environ.get("FDW_C5_NX", "8192")), int(os.environ.get("FDW_C5_NZ", "2048")), 40
    nt = int(os.environ.get("FDW_C5_NT", "300"))
    halo = os.environ.get("FDW_BENCH_HALO", "p2p")
    nxe, nze = nx + 2 * nb, nz + 2 * nb
    srce = fdw.host.ricker_wavelet(nt, 0.001, 15.0, fdw.FAMILY_CPU)
    sx, sz, gz = nb + nx // 2, nb, nb
    kw = dict(order=8, fac=0.01, family=fdw.FAMILY_CPU, nt=nt, rank=rank, world=world, device=device, halo=halo)
    out = {"config": "C5 mod_main + rtm_main algorithm (recipe C bit-exact), %dx%d model (+40 border), %d steps, "
                     "domain-divided over %d GPU(s), halo=%s" % (nx, nz, nt, world, halo)}
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def timed(fn):
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(); dist.barrier()
        return r, time.perf_counter() - t0

    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, taper=fdw.TAPER_FOUR, **kw)
    sp.set_stream(stream.cuda_stream)
    x0, x1 = sp.slab
    v2 = layered(x1 - x0, nze)
    sp.set_v2_local(v2); sp.set_wavelet(srce)
    sp.model_shot(sx, sz, gz)
    rows, dt = timed(lambda: sp.model_shot(sx, sz, gz))
    out["mod_main_gpts_per_s"] = nt * nxe * nze / dt / 1e9
    data = sp.gather_rows(rows)
    sp.close()
    sp = D.SlabPropagator(nx, nz, nb, nb, 10.0, 10.0, 0.001, taper=fdw.TAPER_TOP, history=True, **kw)
    sp.set_stream(stream.cuda_stream)
    sp.set_v2_local(v2); sp.set_wavelet(srce)
    sp.rtm_shot_cpu(sx, sz, gz, data[None], 0)
    _, dt = timed(lambda: sp.rtm_shot_cpu(sx, sz, gz, data[None], 0))
    out["rtm_main_gpts_per_s"] = 2 * nt * nxe * nze / dt / 1e9
    out["history_GB_per_gpu"] = nt * max(sp.owned_interior()[1], 0) * ((nze + 4 + 31) // 32 * 32) * 4 / 1e9
    sp.close()
    return out


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    device = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    res = []
    if world == 1:
        res.append(c1_laplacian(device))
        res.append(c5_cpu_family(device))
    if world > 1 and os.environ.get("FDW_CONFIGS", "c4,c5") .find("c5") >= 0:
        res.append(c5_domain_divided(device, rank, world))
    if os.environ.get("FDW_CONFIGS", "c4,c5").find("c4") >= 0:
        res.append(c4_rtm_shots(device, rank, world))
    if rank == 0:
        for r in res:
            print(json.dumps(r))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
