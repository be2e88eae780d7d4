"""Multi-GPU partitioning of the propagation path (one process per GPU).

The reference has no multi-GPU code; both partitionings are
new, and neither changes the per-point arithmetic:

* slab decomposition (SlabPropagator): the extended grid is cut along x (the
  slow axis, so a halo is GUARD contiguous rows) into `world` slabs.  Two
  exchange mechanisms, same results bit for bit (and equal to one domain):
    halo="p2p"  (default on GPUs) neighbouring processes map each other's field
                buffers through CUDA IPC once; per level the boundary-strip
                kernel stores its rows locally AND into the neighbour's ghost
                rows over NVLink, a release flag follows, the interior update
                overlaps the transfer, the next level's boundary launch waits
                on the device for the neighbours' flags (fdw_peer_levels: the
                whole level loop runs inside the C library, no per-level host
                communication call);
    halo="nccl" boundary strips -> batch_isend_irecv on a communication stream
                while the interior updates -> join (also the gloo path of the
                CPU unit tests).
* shot parallelism (shot_partition / reduce_image): shots are independent
  units; the only collective is the final image sum.

`torch.distributed` (NCCL on GPUs, gloo in the CPU unit tests) is plumbing:
rendezvous, the halo send/recv and the image reduction.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .propagator import Wave2D

GUARD = 4


def slab_rows(nxe, world, rank):
    """contiguous, near-equal row blocks of the extended grid"""
    x0 = (nxe * rank) // world
    x1 = (nxe * (rank + 1)) // world
    return x0, x1


def shot_partition(ns, world, rank, contiguous=False):
    """shots handled by `rank`: round-robin (even load when ns % world != 0), or contiguous
    blocks (needed by the bit-reproducible chained image stack, stack_images_chain)"""
    if contiguous:
        return list(range((ns * rank) // world, (ns * (rank + 1)) // world))
    return list(range(rank, ns, world))


class _CudaView:
    """minimal __cuda_array_interface__ carrier so torch can alias library-owned device memory"""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class SlabPropagator:
    """Wave2D over one slab of a slab-decomposed grid (world == 1: the whole grid)."""

    def __init__(self, nx, nz, nxb, nzb, dx, dz, dt, rank=0, world=1, device=0, lib=None, on_gpu=True, halo=None,
                 **kw):
        self.rank, self.world, self.on_gpu = rank, world, on_gpu
        self.halo = halo or ("p2p" if on_gpu else "nccl")
        self._attached = False
        self.nxe, self.nze = nx + 2 * nxb, nz + 2 * nzb
        self.slab = slab_rows(self.nxe, world, rank)
        self.nloc = self.slab[1] - self.slab[0]
        if world > 1 and self.nloc < 2 * GUARD:
            raise ValueError("slab of %d rows is thinner than two halos" % self.nloc)
        self.w = Wave2D(nx, nz, nxb, nzb, dx, dz, dt, device=device, slab=self.slab if world > 1 else None,
                        lib=lib, **kw)
        self.L, self.h = self.w.L, self.w.h
        self._views = {}
        self.levels_ms = {}
        self._compute_stream = None
        self._comm_stream = None
        self._event = None
        if world > 1 and on_gpu:
            import torch
            self._comm_stream = torch.cuda.Stream(priority=-1)
            self._event = torch.cuda.Event()
        if world > 1 and self.halo == "p2p":
            self.attach_peers()  # collective; settles the exchange mechanism before first use

    # -- plumbing
    def close(self):
        self.w.close()

    def attach_peers(self):
        """halo="p2p": exchange the CUDA IPC handles of the slab buffers (all_gather over the
        process group -- plumbing, once) and map the two neighbours' buffers.  Collective.  If ANY
        rank cannot map its neighbours (GPUs without peer access, IPC unavailable), every rank
        drops back to halo="nccl" together -- a change of transport, same results -- and says so."""
        if self.world == 1 or self._attached or self.halo != "p2p":
            return
        import sys

        import torch
        import torch.distributed as dist
        if not self.on_gpu and not os.environ.get("FDW_TEST_FAIL_PEER_ATTACH"):
            raise ValueError("halo='p2p' maps GPU buffers across processes; use halo='nccl' on the CPU")
        ok, why, mine = 1, "", _lib.PeerInfo()
        try:
            if os.environ.get("FDW_TEST_FAIL_PEER_ATTACH") == str(self.rank):
                raise _lib.FdwError(-2, "peer attach failure injected by FDW_TEST_FAIL_PEER_ATTACH")
            _lib.check(self.L, self.L.fdw_peer_export(self.h, C.byref(mine)))
        except _lib.FdwError as e:
            ok, why = 0, str(e)
        infos = [None] * self.world
        dist.all_gather_object(infos, bytes(mine) if ok else None)
        if ok and all(i is not None for i in infos) and not os.environ.get("FDW_TEST_FAIL_PEER_ATTACH"):
            lo = _lib.PeerInfo.from_buffer_copy(infos[self.rank - 1]) if self.rank > 0 else None
            hi = _lib.PeerInfo.from_buffer_copy(infos[self.rank + 1]) if self.rank < self.world - 1 else None
            try:
                _lib.check(self.L, self.L.fdw_peer_attach(self.h, C.byref(lo) if lo else None,
                                                          C.byref(hi) if hi else None))
            except _lib.FdwError as e:
                ok, why = 0, str(e)
        else:
            ok = 0
        agree = torch.tensor([ok], dtype=torch.int32,
                             device="cuda" if self.on_gpu and dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)  # also the barrier: nobody pushes before every slab has zeroed its flags
        if int(agree.item()) == 0:
            self.L.fdw_peer_detach(self.h)
            self.halo = "nccl"
            if why or self.rank == 0:
                print("fdwave: peer-memory halo exchange unavailable (%s); all slabs use NCCL send/recv"
                      % (why or "a neighbour could not map its peers"), file=sys.stderr)
            return
        self._attached = True

    def _peer_refresh(self):
        """after this slab's buffers were zeroed or uploaded: push the newest level's boundary rows
        and raise the flag; the neighbours' first boundary launch of the next level waits for it
        (so it can never push into a buffer that is still to be zeroed)"""
        _lib.check(self.L, self.L.fdw_peer_refresh(self.h))
        _lib.check(self.L, self.L.fdw_peer_fence(self.h))

    @property
    def p2p(self):
        return self.world > 1 and self.halo == "p2p" and self._attached

    def set_stream(self, cuda_stream):
        self.w.set_stream(cuda_stream)
        self._compute_stream = cuda_stream

    def set_wavelet(self, s):
        self.w.set_wavelet(s)

    def set_source(self, sx, sz, kind=_lib.SRC_POINT):
        self.w.set_source(sx, sz, kind)

    def zero(self):
        self.w.zero()
        if self.p2p:
            self._peer_refresh()

    def launch_count(self):
        return self.w.launch_count()

    def set_v2_local(self, v2_rows):
        v2_rows = np.ascontiguousarray(v2_rows, np.float32)
        assert v2_rows.shape == (self.nloc, self.nze)
        _lib.check(self.L, self.L.fdw_set_v2_local(self.h, v2_rows))

    def upload_local(self, newest, older):
        _lib.check(self.L, self.L.fdw_fields_upload_local(self.h, 0, newest, older))
        self.refresh_halos()

    def download_local(self, newest=None, older=None):
        n = np.zeros((self.nloc, self.nze), np.float32) if newest is None else newest
        o = np.zeros((self.nloc, self.nze), np.float32) if older is None else older
        _lib.check(self.L, self.L.fdw_fields_download_local(
            self.h, 0, n.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p)))
        return n, o

    # -- halo exchange
    def _tensor(self, ptr, count):
        key = (ptr, count)
        t = self._views.get(key)
        if t is None:
            import torch
            if self.on_gpu:
                t = torch.as_tensor(_CudaView(ptr, count), device="cuda")
            else:
                buf = (C.c_float * count).from_address(ptr)
                t = torch.from_numpy(np.ctypeslib.as_array(buf))
            self._views[key] = t
        return t

    def _exchange_ops(self, level):
        import torch.distributed as dist
        hl = _lib.Halo()
        _lib.check(self.L, self.L.fdw_halo_get(self.h, level, C.byref(hl)))
        n = hl.count
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, self._tensor(hl.send_lo, n), self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, self._tensor(hl.recv_lo, n), self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, self._tensor(hl.send_hi, n), self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, self._tensor(hl.recv_hi, n), self.rank + 1))
        return ops

    def refresh_halos(self):
        """exchange the boundary rows of the newest level (after an upload)"""
        if self.world == 1:
            return
        import torch.distributed as dist
        if self.p2p:
            self._peer_refresh()
            return
        if self.on_gpu:
            import torch
            self.w.sync()  # the upload ran on the library stream
            torch.cuda.current_stream().synchronize()
        for r in dist.batch_isend_irecv(self._exchange_ops(0)):
            r.wait()
        if self.on_gpu:
            torch.cuda.current_stream().synchronize()

    # -- slab-decomposed shots of the CPU family (config 5: mod_main + rtm_main, domain-divided)
    def owned_interior(self):
        """global interior x rows owned by this slab: (first, count)"""
        d = self.w.devinfo()
        return d.li0 - self.w.nxb, d.nli

    def _shot(self, phase, sx, sz, gz, dobs_all=None, is_=0):
        L, h = self.L, self.h
        ns = 1
        ptr = None
        if dobs_all is not None:
            dobs_all = np.ascontiguousarray(dobs_all, np.float32)
            ns = dobs_all.size // (self.w.nx * self.w.nt)
            ptr = dobs_all.ctypes.data_as(C.c_void_p)
        _lib.check(L, L.fdw_shot_begin(h, phase, sx, sz, gz, ptr, ns, is_))
        if self.p2p:
            self._peer_refresh()  # the phase zeroed the fields
        if self.on_gpu:
            self.w.mark_begin()
        self._levels(0, self.w.nt)
        if self.on_gpu:  # device time of the level loop of this phase (CUDA events on the library's stream)
            self.levels_ms[phase] = self.w.mark_end()

    def model_shot(self, sx, sz, gz):
        """one shot of mod_main (mod_main.cpp:141-169) on the slab-decomposed grid; returns this
        slab's traces [nli][nt] (rows owned_interior())"""
        self._shot(_lib.PHASE_MODEL, sx, sz, gz)
        out = np.zeros((max(self.owned_interior()[1], 0), self.w.nt), np.float32)
        _lib.check(self.L, self.L.fdw_shot_end(self.h, out if out.size else np.zeros(1, np.float32)))
        return out

    def rtm_shot_cpu(self, sx, sz, gz, dobs_all, is_=0):
        """one shot of rtm_main (rtm_main.cpp:158-240) on the slab-decomposed grid: the forward
        history is sharded with the slabs (each GPU keeps its own rows in HBM); returns this
        slab's image rows [nli][nz]"""
        self._shot(_lib.PHASE_RTM_FWD, sx, sz, gz)
        self._shot(_lib.PHASE_RTM_BWD, sx, sz, gz, dobs_all, is_)
        out = np.zeros((max(self.owned_interior()[1], 0), self.w.nz), np.float32)
        _lib.check(self.L, self.L.fdw_shot_end(self.h, out if out.size else np.zeros(1, np.float32)))
        return out

    def gather_rows(self, local_rows):
        """concatenate the per-slab row blocks on every rank (seismograms, images)"""
        if self.world == 1:
            return local_rows
        import torch.distributed as dist
        parts = [None] * self.world
        dist.all_gather_object(parts, local_rows)
        return np.concatenate([p for p in parts if p.shape[0] > 0])

    # -- time stepping
    def advance(self, it0, nsteps):
        if self.world == 1:
            self.w.advance(it0, nsteps)
            return
        _lib.check(self.L, self.L.fdw_shot_begin(self.h, _lib.PHASE_PLAIN, 0, 0, 0, None, 1, 0))
        self._levels(it0, nsteps)

    def _levels(self, it0, nsteps):
        """nsteps time levels of the current phase with the overlapped halo exchange"""
        L, h = self.L, self.h
        if self.world == 1:
            for it in range(it0, it0 + nsteps):
                _lib.check(L, L.fdw_step_begin(h, it))
                _lib.check(L, L.fdw_step_rows(h, 0, self.nloc, None))
                _lib.check(L, L.fdw_step_end(h))
            return
        if self.p2p:
            _lib.check(L, L.fdw_peer_levels(h, it0, nsteps))
            _lib.check(L, L.fdw_peer_fence(h))
            return
        import torch.distributed as dist
        lo_nb, hi_nb = self.rank > 0, self.rank < self.world - 1
        ilo = GUARD if lo_nb else 0
        ihi = self.nloc - GUARD if hi_nb else self.nloc
        if self.on_gpu:
            import torch
            # the levels, the library's own zero / upload / download work and the halo send/recv must be
            # ordered on ONE stream: adopt torch's current stream as the library stream (fdw_set_stream
            # drains the previous one first), then launch the rows on it
            compute = torch.cuda.current_stream()
            if self._compute_stream != compute.cuda_stream:
                self.set_stream(compute.cuda_stream)
            cs = C.c_void_p(compute.cuda_stream)
        else:
            cs = None
        for it in range(it0, it0 + nsteps):
            _lib.check(L, L.fdw_step_begin(h, it))
            if lo_nb:
                _lib.check(L, L.fdw_step_rows(h, 0, GUARD, cs))
            if hi_nb:
                _lib.check(L, L.fdw_step_rows(h, self.nloc - GUARD, self.nloc, cs))
            ops = self._exchange_ops(1)
            if self.on_gpu:
                self._event.record(compute)
                self._comm_stream.wait_event(self._event)
                with torch.cuda.stream(self._comm_stream):
                    reqs = dist.batch_isend_irecv(ops)
                _lib.check(L, L.fdw_step_rows(h, ilo, ihi, cs))  # interior overlaps the exchange
                with torch.cuda.stream(self._comm_stream):
                    for r in reqs:
                        r.wait()
                    self._event.record(self._comm_stream)
                compute.wait_event(self._event)
            else:
                reqs = dist.batch_isend_irecv(ops)
                _lib.check(L, L.fdw_step_rows(h, ilo, ihi, cs))
                for r in reqs:
                    r.wait()
            _lib.check(L, L.fdw_step_end(h))

    def propagate_local(self, newest, older, it0, nsteps):
        """slab-local host rows in, nsteps levels, host rows out (in place)"""
        if self.world == 1:
            self.w.propagate(newest, older, it0, nsteps)
            return
        self.upload_local(newest, older)
        self.advance(it0, nsteps)
        self.download_local(newest, older)


    def propagate_local_async(self, newest, older, it0, nsteps, stream=None, levels_after=None, levels_done=None):
        """propagate_local without the final synchronisation: everything (uploads from the -- pinned --
        host rows, halo refresh, nsteps levels, downloads into the same rows) is enqueued on this
        propagator's stream; call sync() before reading the arrays.  Two propagators on two streams
        pipeline independent jobs: one's PCIe transfers overlap the other's levels.

        stream (the torch stream given to set_stream) + levels_after / levels_done (torch CUDA events): the levels
        start only after `levels_after` and `levels_done` is recorded behind them, so that two pipelined jobs overlap
        their TRANSFERS with each other's levels but not their level loops (two grids' step kernels interleaved on one
        GPU evict each other's halo rows from L2)."""
        if self.world > 1 and not self.p2p:
            raise NotImplementedError("propagate_local_async needs the peer-memory halo exchange (halo='p2p')")
        _lib.check(self.L, self.L.fdw_fields_upload_local(self.h, 0, newest, older))
        if self.world > 1:
            self._peer_refresh()
        if stream is not None and levels_after is not None:
            stream.wait_event(levels_after)
        self.advance(it0, nsteps)
        if stream is not None and levels_done is not None:
            levels_done.record(stream)
        _lib.check(self.L, self.L.fdw_fields_download_local_async(
            self.h, 0, newest.ctypes.data_as(C.c_void_p), older.ctypes.data_as(C.c_void_p)))

    def sync(self):
        self.w.sync()


def reduce_image(img, op="ordered"):
    """stack per-rank partial images (img += imloc of fd-code.cu:525 / rtm_main.cpp:237).
    'ordered' gathers to every rank and sums in rank order (bit-reproducible);
    'allreduce' is a single NCCL/gloo all-reduce (fast path, order unspecified)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return img
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(img, np.float32)).to(dev)
    if op == "allreduce":
        dist.all_reduce(t)
        return t.cpu().numpy()
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    acc = parts[0].clone()
    for q in parts[1:]:
        acc += q
    return acc.cpu().numpy()


def stack_images_chain(shot_images, shape):
    """Bit-reproducible image stack for shot parallelism with contiguous shot blocks.

    The reference stacks sequentially, img += imloc in shot order (fd-code.cu:525,
    rtm_main.cpp:237).  Float addition is not associative, so per-rank partial sums followed by
    a reduction give a different rounding.  Here the running sum travels down the ranks: rank r
    receives the stack of all earlier shots, adds its own shot images in order and passes it
    on; the last rank broadcasts the result.  Same additions in the same order as one process."""
    import torch
    import torch.distributed as dist
    acc = np.zeros(shape, np.float32)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        for im in shot_images:
            acc += im
        return acc
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(shape, dtype=torch.float32, device=dev)
    if rank > 0:
        dist.recv(t, rank - 1)
    for im in shot_images:
        t += torch.from_numpy(np.ascontiguousarray(im, np.float32)).to(dev)
    if rank < world - 1:
        dist.send(t, rank + 1)
    dist.broadcast(t, world - 1)
    return t.cpu().numpy()


class ShotPipeline:
    """Shot-parallel RTM of the GPU family (main() loop of fd-code.cu:480-529) without host round trips: this rank
    migrates its shots on its own GPU; the next shot's velocity is staged (pinned memory -> second device buffer,
    premultiplied on a copy stream) while the current shot runs, the shot images are stacked on the device in shot
    order, the cross-rank stack is reduced straight from device memory, and the image is downloaded once.

    stack = "device"    single rank: sequential stack, bit-identical to the reference's img += imloc
            "chain"     the running stack travels down the ranks in shot order (contiguous shot blocks): still
                        bit-identical to the sequential reference order
            "allreduce" per-rank partial stacks + one all-reduce (fast path; summation order unspecified)"""

    def __init__(self, nx, nz, nxb, nzb, dx, dz, dt, nt, device=0, order=8, fac=0.7, compat_extents=False, lib=None):
        self.w = Wave2D(nx, nz, nxb, nzb, dx, dz, dt, order=order, fac=fac, family=_lib.FAMILY_GPU,
                        taper=_lib.TAPER_TOP, compat_extents=compat_extents, device=device, nt=nt, lib=lib)
        self.nx, self.nz, self.nt = nx, nz, nt
        self._pinned = []

    def close(self):
        self.w.close()

    def set_wavelet(self, s):
        self.w.set_wavelet(s)

    def pinned(self, shape):
        """page-locked float32 host array (torch owns the allocation; kept alive by the pipeline)"""
        import torch
        t = torch.zeros(shape, dtype=torch.float32).pin_memory()
        self._pinned.append(t)
        return t.numpy()

    def _stack_tensor(self):
        import torch
        ptr, pitch, rows = self.w.stack_devptr()
        return torch.as_tensor(_CudaView(ptr, pitch * rows), device="cuda")

    def run_shots(self, shots, v2_of_shot, dobs_of_shot, sx_of_shot, sz, gz, stack="device"):
        import torch.distributed as dist
        w = self.w
        multi = stack in ("chain", "allreduce") and dist.is_initialized() and dist.get_world_size() > 1
        w.stack_zero()
        if shots:
            w.v2_stage(v2_of_shot(shots[0]))
        nccl = multi and dist.get_backend() == "nccl"  # gloo (tests on one device): bounce through the host
        if multi and stack == "chain" and dist.get_rank() > 0:
            w.sync()
            t = self._stack_tensor()  # the stack of all earlier shots arrives here
            if nccl:
                dist.recv(t, dist.get_rank() - 1)
            else:
                h = t.cpu()
                dist.recv(h, dist.get_rank() - 1)
                t.copy_(h)
            import torch
            torch.cuda.current_stream().synchronize()
        for i, is_ in enumerate(shots):
            w.v2_commit()
            if i + 1 < len(shots):
                w.v2_stage(v2_of_shot(shots[i + 1]))  # overlaps this shot
            w.forward(sx_of_shot(is_), sz, download=False)
            w.backward_device(dobs_of_shot(is_), gz)
            w.stack_add()
        if multi:
            w.sync()
            import torch
            t = self._stack_tensor()
            x = t if nccl else t.cpu()
            if stack == "chain":
                rank, world = dist.get_rank(), dist.get_world_size()
                if rank < world - 1:
                    dist.send(x, rank + 1)
                dist.broadcast(x, world - 1)
            else:
                dist.all_reduce(x)
            if not nccl:
                t.copy_(x)
            torch.cuda.current_stream().synchronize()
        return w.stack_download()

    def time_one_shot(self, v2, dobs, sx, sz, gz):
        """device times (CUDA events) of the pieces of one shot"""
        w = self.w
        w.sync()
        t0 = _now()
        w.v2_stage(v2)
        w.v2_commit()
        w.sync()
        stage_ms = (_now() - t0) * 1e3
        w.mark_begin()
        w.forward(sx, sz, download=False)
        f_ms = w.mark_end()
        w.mark_begin()
        w.backward_device(dobs, gz)
        b_ms = w.mark_end()
        return {"v2_stage_ms_wall_unoverlapped": stage_ms, "forward_ms": f_ms, "backward_ms": b_ms}


def _now():
    import time
    return time.perf_counter()


def migrate_shots_gpu_family(wave, shots, v2_of_shot, dobs_of_shot, sx_of_shot, sz, gz, stack="chain"):
    """Shot-parallel RTM with the GPU family's algorithm (main() loop of fd-code.cu:480-529):
    this rank migrates `shots` (from shot_partition) on its own GPU -- forward with the two
    saved levels left in HBM, time-reversed reconstruction + receiver back-propagation +
    imaging -- and the per-shot images are stacked across ranks.

    stack = "chain": bit-identical to the sequential reference order (needs contiguous blocks);
            "allreduce": per-rank partial sums + one all-reduce (fast path)."""
    images = []
    for is_ in shots:
        wave.set_v2(v2_of_shot(is_))
        wave.forward(sx_of_shot(is_), sz, download=False)
        images.append(wave.backward(dobs_of_shot(is_), gz))
    shape = (wave.nx, wave.nz)
    if stack == "chain":
        return stack_images_chain(images, shape)
    part = np.zeros(shape, np.float32)
    for im in images:
        part += im
    return reduce_image(part, "allreduce")
