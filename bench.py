#!/usr/bin/env python
"""bench.py -- headline benchmark of the propagation hot path.

Metric (BASELINE.json): grid-point updates/s (Gpts/s) of the fused
Laplacian+leapfrog propagator, plus achieved HBM GB/s against the measured
B200 roofline.

Workload (BASELINE.json configs[2]): synthetic 2-D
propagator on a 16384 x 16384 extended grid (40-point sponge inside it),
order 8, dx=dz=10 m, dt=1 ms, 3-layer velocity 2000/3000/4000 m/s, point
source, top sponge, GPU-family arithmetic (recipe G, bit-exact with the
reference's kernel_lap + kernel_time).  One bench "step" = LEVELS time levels
of that grid (the full job of 10k time levels = 10000/LEVELS steps).  With N
GPUs the grid is slab-decomposed along x with a per-level halo exchange and
the per-GPU slab stays 16384 x 16384 (weak scaling).

    python bench.py --gpus 1 --steps K --warmup W          # our arm
    python bench.py --impl reference --steps K --warmup W  # reference CPU path

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NGRID = int(os.environ.get("FDW_BENCH_N", "16384"))
LEVELS = int(os.environ.get("FDW_BENCH_LEVELS", "250"))
NB = 40
DX = DZ = 10.0
DT = 0.001
FPEAK = 20.0
FAC = 0.75
STRONG = os.environ.get("FDW_BENCH_STRONG", "0") == "1"  # 1: the x extent stays NGRID for every N (strong scaling)
HALO = os.environ.get("FDW_BENCH_HALO", "p2p")  # slab halo exchange at N>1: "p2p" (peer stores over NVLink) or "nccl"
RECIPE = os.environ.get("FDW_BENCH_RECIPE", "G")  # "G" = bit-exact reference arithmetic (headline); "FAST" = FMA recipe
BYTES_PER_POINT = 16  # read p, pp, v2*dt2 + write pp
METRIC = "grid-point updates/sec"
UNIT = "Gpts/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def layered_v2(nxe, nze):
    ve = np.empty((nxe, nze), np.float32)
    ve[:, : nze // 3] = 2000.0
    ve[:, nze // 3: 2 * nze // 3] = 3000.0
    ve[:, 2 * nze // 3:] = 4000.0
    return ve * ve


def bind_to_gpu_numa_node(index):
    """pin this process to the CPU cores of the GPU's NUMA node before any pinned host buffer is
    allocated (first-touch puts the pages on that node), so that the end-to-end leg's PCIe copies of N
    ranks do not cross the socket interconnect.  Best effort: silently does nothing off Linux/NVML."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------- reference arm
def reference_steps(nsteps, warmup, sample_n, sample_levels, force_port=False):
    """The reference's own CPU implementation of the path (oracle/_ref: fd_step +
    point source + taper_apply2, i.e. rtm_main's forward loop, rtm_main.cpp:166-176),
    serial like the reference; falls back to the oracle port with all cores
    (force_port: time the all-core port even when the reference objects exist)."""
    from oracle import oracle as O
    from oracle import ref as R
    nb = NB
    n = sample_n
    nx = nz = n - 2 * nb
    v2 = layered_v2(n, n)
    srce = O.ricker_wavelet(10000, DT, FPEAK, O.FAM_C)
    rng = np.random.default_rng(1)
    p = rng.standard_normal((n, n), dtype=np.float32)
    pp = rng.standard_normal((n, n), dtype=np.float32)
    sx, sz = n // 2, nb
    times = []
    if R.available("libref_cpufam.so") and not force_port:
        kind, cores = "reference", 1
        cpu = R.CpuFam()
        cpu.fd_init(8, n, n, DX, DZ, DT)
        cpu.taper_init(nb, nb, FAC)

        def level(it, p, pp):
            cpu.fd_step(8, p, pp, v2)
            pp[sx, sz] += srce[it]
            cpu.taper_apply2(pp, nx, nz, nb, nb)
            cpu.taper_apply2(p, nx, nz, nb, nb)
    else:
        kind, cores = "port", O.max_threads()
        O.set_threads(cores)
        lap = np.zeros_like(p)
        tx, tz = O.taper_table(nb, FAC, O.FAM_C), O.taper_table(nb, FAC, O.FAM_C)

        def level(it, p, pp):
            O.fd_step(8, p, pp, v2, lap, DX, DZ, DT)
            pp[sx, sz] += srce[it]
            O.taper_top(pp, nb, tx, tz)
            O.taper_top(p, nb, tx, tz)
    it = 0
    for s in range(warmup + nsteps):
        t0 = time.perf_counter()
        for _ in range(sample_levels):
            level(it, p, pp)
            p, pp = pp, p
            it += 1
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    pts = float(n) * n * sample_levels
    return pts * len(times) / sum(times) / 1e9, sum(times) / len(times), kind, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, lv = 2048, 2
    value, t_step, kind, cores = reference_steps(args.steps, args.warmup, n, lv)
    pv, _, _, pcores = reference_steps(max(1, args.steps), 1, n, lv, force_port=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d levels of a %dx%d sub-grid of the workload per step "
                                   "(fd_step + point source + taper_apply2, serial as shipped)" % (lv, n, n),
                         "port_all_cores": {"value": pv, "unit": UNIT, "cores": pcores, "kind": "port",
                                            "note": "the oracle's C restatement with OpenMP over x on every host core "
                                                    "(the reference itself has no threading)"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


REFERENCE_ARM = {
    "what": "the reference's own CPU implementation of this path (fd_step + point source + taper_apply2 = the time "
            "loop of rtm_main.cpp:166-176, from oracle/_ref/libref_cpufam.so), serial as shipped",
    "sample": "each step = 2 time levels of a 2048 x 2048 sub-grid of the workload (bounded so that the run ends "
              "in minutes); throughput in the same unit",
    "recipe": "C (fd.c:24-46: the only arithmetic the reference's CPU path has); the GPU arm runs recipe G "
              "(kernel_lap + kernel_time, fd-code.cu:53-92) -- the same 2-D order-8 update, different rounding",
}


def workload_config(ngpus):
    cfg = _workload_config(ngpus)
    cfg["reference_arm"] = REFERENCE_ARM
    return cfg


def _workload_config(ngpus):
    if STRONG:
        return {"workload": "synthetic 2D stencil propagator, %dx%d extended grid in total, cut into %d slab%s along x, "
                            "order 8, recipe %s, top sponge, point source; %d time levels per step"
                            % (NGRID, NGRID, ngpus, "s" if ngpus > 1 else "", RECIPE, LEVELS),
                "grid": [NGRID, NGRID], "levels_per_step": LEVELS, "order": 8, "recipe": RECIPE,
                "partition": "slab-x%d" % ngpus if ngpus > 1 else "single", "halo_exchange": HALO if ngpus > 1 else None,
                "l2": "working set 3 GiB in total"}
    return {"workload": "synthetic 2D stencil propagator, %dx%d extended grid per GPU (x extent %d over %d slab%s), "
                        "order 8, recipe %s, top sponge, point source; %d time levels per step"
                        % (NGRID, NGRID, NGRID * ngpus, ngpus, "s" if ngpus > 1 else "",
                           "G (bit-exact)" if RECIPE != "FAST" else "FAST (FMA, tolerance-checked)", LEVELS),
            "grid": [NGRID * ngpus, NGRID], "levels_per_step": LEVELS, "order": 8, "recipe": RECIPE,
            "partition": "slab-x%d" % ngpus if ngpus > 1 else "single",
            "halo_exchange": ("peer stores into the neighbour's ghost rows from the boundary kernel (CUDA IPC, "
                              "NVLink), device-side flags" if HALO == "p2p" else "NCCL send/recv") if ngpus > 1 else None,
            "l2": "working set 3 GiB per GPU > 126 MB L2, no flush needed"}


# ---------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch

    import parallel_finite_difference_computation_b200 as fdw
    from parallel_finite_difference_computation_b200 import distributed as fdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    numa_node = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nxe_g, nze = (NGRID if STRONG else NGRID * world), NGRID
    nb = NB
    nx, nz = nxe_g - 2 * nb, nze - 2 * nb
    srce = fdw.host.ricker_wavelet(10000, DT, FPEAK, fdw.FAMILY_GPU)
    prop = fdist.SlabPropagator(nx, nz, nb, nb, DX, DZ, DT, order=8, fac=FAC, family=fdw.FAMILY_GPU,
                                recipe=fdw.RECIPE_FAST if RECIPE == "FAST" else fdw.RECIPE_G,
                                taper=fdw.TAPER_TOP, device=local_rank, rank=rank, world=world, halo=HALO)
    if world > 1 and HALO == "p2p" and not prop.p2p:
        globals()["HALO"] = "nccl"  # the slabs agreed to fall back (no peer access); report what actually ran
    # a non-default torch stream: the library launches on it, and the torch events below time it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    prop.set_stream(stream.cuda_stream)
    x0, x1 = prop.slab
    v2_local = layered_v2(x1 - x0, nze)
    prop.set_v2_local(v2_local)
    prop.set_wavelet(srce)
    prop.set_source(nxe_g // 2, nb)
    prop.zero()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps of LEVELS levels
    it = 0
    for _ in range(args.warmup):
        prop.advance(it, LEVELS)
        it += LEVELS
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = prop.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        prop.advance(it, LEVELS)
        it += LEVELS
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.result()
    launches = prop.launch_count() - l0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pts_per_step = float(nxe_g) * nze * LEVELS
    value = pts_per_step * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end: host (pinned) fields in, LEVELS levels, host fields out, through the C ABI.
    # Every step uploads its own two time levels and downloads its two result levels.  Steps are independent
    # jobs, so two propagators on two streams keep two of them in flight: one job's PCIe transfers overlap the
    # other job's levels (double buffering, as a production caller would); the timed region still contains
    # every step's H2D and D2H copies.
    nloc = x1 - x0
    pipelined = world == 1 or prop.p2p
    props, bufs = [prop], []
    if pipelined:
        prop2 = fdist.SlabPropagator(nx, nz, nb, nb, DX, DZ, DT, order=8, fac=FAC, family=fdw.FAMILY_GPU,
                                     recipe=fdw.RECIPE_FAST if RECIPE == "FAST" else fdw.RECIPE_G,
                                     taper=fdw.TAPER_TOP, device=local_rank, rank=rank, world=world, halo=HALO)
        stream2 = torch.cuda.Stream()
        prop2.set_stream(stream2.cuda_stream)
        prop2.set_v2_local(v2_local)
        prop2.set_wavelet(srce)
        prop2.set_source(nxe_g // 2, nb)
        if world == 1 or prop2.p2p:
            props.append(prop2)
        else:  # (cannot happen if the first one attached; be safe) one step at a time
            prop2.close()
            pipelined = False
    for _ in props:
        hn = torch.zeros((nloc, nze), dtype=torch.float32).pin_memory()
        ho = torch.zeros((nloc, nze), dtype=torch.float32).pin_memory()
        bufs.append((hn, ho, hn.numpy(), ho.numpy()))
    e2e_steps = max(4, min(args.steps, 40))  # K steps like the device-timed leg (the whole 10 000-level job is 40)

    e2e_streams = [stream, stream2] if pipelined and len(props) == 2 else None
    lv_done = [torch.cuda.Event(), torch.cuda.Event()]
    serial_levels = os.environ.get("FDW_BENCH_E2E_SERIAL_LEVELS", "1") != "0"

    def e2e_job(k, s):
        if pipelined:
            props[k].sync()  # the job issued two steps ago on this propagator has delivered its results
            if e2e_streams is not None and serial_levels:
                # the other job's transfers overlap these levels, its level loop does not (see propagate_local_async)
                props[k].propagate_local_async(bufs[k][2], bufs[k][3], s * LEVELS, LEVELS, stream=e2e_streams[k],
                                               levels_after=lv_done[1 - k], levels_done=lv_done[k])
            else:
                props[k].propagate_local_async(bufs[k][2], bufs[k][3], s * LEVELS, LEVELS)
        else:
            props[k].propagate_local(bufs[k][2], bufs[k][3], s * LEVELS, LEVELS)

    for k in range(len(props)):  # warm-up
        e2e_job(k, 0)
    for pr in props:
        pr.sync()
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        e2e_job(s % len(props), s)
    for pr in props:
        pr.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = pts_per_step * e2e_steps / e2e_s / 1e9
    field_bytes = nloc * nze * 4

    # ---- the other BASELINE configs, device-timed, in the same line (outside the headline's timed regions)
    configs, parity = {}, None
    for pr in props[1:]:
        pr.close()
    if not args.no_configs:
        import bench_configs as BC

        def guarded(name, fn):
            try:
                configs[name] = fn()
            except Exception as e:  # a failing side config must not take the headline with it
                configs[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

        if world == 1:
            guarded("C1", lambda: BC.c1_laplacian(prop.w))
        prop.close()
        if world == 1:
            guarded("C2", lambda: BC.c2_small_models(local_rank))
        guarded("C4", lambda: BC.c4_rtm_shots(local_rank, rank, world))
        if world == 1:
            guarded("C5", lambda: BC.c5_cpu_family(local_rank))
            guarded("gpu_reference", lambda: BC.gpu_reference(device=local_rank))
        else:
            guarded("C5", lambda: BC.c5_domain_divided(local_rank, rank, world))
            try:
                parity = BC.parity_n_vs_1(rank, world, local_rank, HALO)
            except Exception as e:
                parity = {"result": "error", "error": str(e)[:300]}
    else:
        prop.close()

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (one launch per level per GPU)
    peak, peak_src = peaks()
    per_launch_ms = ms / (args.steps * LEVELS)
    launch_bytes = float(nloc) * nze * BYTES_PER_POINT  # rank 0's slab (= NGRID rows in the weak-scaling default)
    achieved = launch_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, "not measured (--no-traffic)"
    if not args.no_traffic and NGRID == 16384:
        import bench_configs as BC
        traffic, traffic_src = BC.measure_traffic()

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, _, kind, cores = reference_steps(6, 1, 2048, 1)
        pv, _, _, pcores = reference_steps(6, 1, 2048, 1, force_port=True)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "6 time levels of a 2048x2048 sub-grid of the workload "
                         "(reference fd_step + point source + taper_apply2, serial as shipped)",
               "port_all_cores": {"value": pv, "unit": UNIT, "cores": pcores, "kind": "port",
                                  "note": "the oracle's C restatement with OpenMP over x on every host core "
                                          "(the reference itself has no threading)"}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if STRONG else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * field_bytes * world,
                "d2h_bytes_per_step": 2 * field_bytes * world, "steps": e2e_steps, "host_numa_node_rank0": numa_node,
                "note": "per step: both time levels H2D from pinned host memory, %d levels, both levels D2H; %s" % (
                    LEVELS, "two independent steps in flight on two contexts/streams (double-buffered: one step's "
                    "transfers overlap the other's levels)" if pipelined else "one step at a time")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": "k_step<8,%s> fused Laplacian+leapfrog+sponge+source" % RECIPE,
                     "algorithmic_bytes_per_launch": launch_bytes, "avg_launch_ms": per_launch_ms},
        "cpu_baseline": cpu,
        "configs": configs,
    }
    if world > 1:
        line["parity_n_vs_1"] = parity["result"] if parity else None
        line["parity_n_vs_1_detail"] = parity
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only (skip the C1/C2/C4/C5 sub-records)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu DRAM-byte measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
