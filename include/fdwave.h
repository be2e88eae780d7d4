/*
 * fdwave.h -- C ABI of the B200-native finite-difference acoustic propagation
 * library (libfdwave.so).  Plain C: pointers, sizes and ints only.
 *
 * The reference (FernandoSchett/parallel_finite_difference_computation) has no
 * plugin or FFI boundary: its hot path is a handful of C functions called by
 * each program's main().  Every entry point below names the reference
 * function(s) it replaces.  Shims with the reference's exact signatures live
 * in fdwave_cpufam.h / fdwave_gpufam.h (two libraries, because both families
 * export an `fd_init` with different argument lists).
 *
 * Conventions
 *   - all arrays are float32, logical shape [nxe][nze], z fastest, exactly the
 *     reference's alloc2float(nze,nxe) block (functions.c:168-182);
 *     nxe = nx + 2*nxb, nze = nz + 2*nzb (fd-code.cu:410-411);
 *   - "newest"/"older" name the two time levels explicitly (the reference's
 *     P/PP naming differs between its two families);
 *   - every function returns FDW_OK (0) or a negative error code;
 *     fdw_last_error() gives the message.  No function falls back to the CPU:
 *     without a CUDA device the compute entry points fail with FDW_ERR_CUDA.
 *   - a context is bound to one device and one stream; calls on it are
 *     asynchronous up to the next download/sync.
 */
#ifndef FDWAVE_H
#define FDWAVE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDW_OK 0
#define FDW_ERR_ARG (-1)
#define FDW_ERR_CUDA (-2)
#define FDW_ERR_NOMEM (-3)
#define FDW_ERR_IO (-4)
#define FDW_ERR_STATE (-5)
#define FDW_ERR_UNSUPPORTED (-6)

/* table family: which of the reference's two code families a host table or a
 * context reproduces */
#define FDW_FAMILY_GPU 0 /* cuda_reference_RTM / dpct_migrated_* */
#define FDW_FAMILY_CPU 1 /* dpct_gpu_rtm_domain_division */

/* arithmetic recipe of the Laplacian + update */
#define FDW_RECIPE_G 0    /* kernel_lap + kernel_time, fd-code.cu:53-92 (bit-exact) */
#define FDW_RECIPE_C 1    /* fd_step, fd.c:24-46 (bit-exact) */
#define FDW_RECIPE_FAST 2 /* symmetric pairs + FMA, float update (tolerance-checked) */

/* sponge geometry */
#define FDW_TAPER_NONE 0
#define FDW_TAPER_TOP 1  /* kernel_tapper fd-code.cu:94-117 == taper_apply2 taper.c:69-84 */
#define FDW_TAPER_FOUR 2 /* taper_apply taper.c:47-67 */

/* source kind */
#define FDW_SRC_POINT 0  /* kernel_src fd-code.cu:119-122, rtm_main.cpp:171 */
#define FDW_SRC_GAUSS7 1 /* ptsrc ptsrc.c:12-58 */

typedef struct fdw_ctx fdw_ctx;

typedef struct fdw_params {
    int nx, nz;   /* interior grid */
    int nxb, nzb; /* sponge / border widths */
    int order;    /* even, 2..16; 2-8 are the tabulated weights, 10-16 the makeo2 windowed-sinc ones (functions.c:119-157).
                   * Orders above 8 run whole-grid contexts only (no slab decomposition: slabs exchange 4 ghost rows) */
    float dx, dz, dt;
    float fac;          /* sponge factor (meaning differs per family, see fdw_taper_table) */
    int family;         /* FDW_FAMILY_*: tables + step ordering */
    int recipe;         /* FDW_RECIPE_* */
    int taper;          /* FDW_TAPER_* */
    int compat_extents; /* 1: reproduce the reference GPU family's truncated launch
                           extents floor(n/8)*8 (fd-code.cu:185-195) */
    int device;         /* CUDA device ordinal */
    int slab_x0, slab_x1; /* extended-grid rows owned by this context; 0,0 = all rows */
    int history;        /* 1: allocate the forward history needed by fdw_rtm_shot_cpu */
    int nt;             /* time steps per shot (sizes wavelet, traces, history) */
} fdw_params;

/* ---------------------------------------------------------------- host tables
 * Pure host code, callable without a GPU. */

/* calc_coefs: functions.c:78-123 (family GPU), fd.c:54-97 (family CPU). order+1 floats. */
int fdw_calc_coefs(int order, int family, float *coefs);
/* ricker_wavelet: functions.c:293-299 (GPU), ptsrc.c:88-99 (CPU: zero after 2/fpeak). */
int fdw_ricker_wavelet(int nt, float dt, float fpeak, int family, float *s);
/* sponge table: fd-code.cu:159-166 (GPU: fac -> dfrac), taper.c:33-42 (CPU: fac used directly). */
int fdw_taper_table(int nb, float fac, int family, float *tab);
/* extendvel, taper.c:7-23: constant extension of an [nxe][nze] array in place. */
int fdw_extendvel(int nx, int nz, int nxb, int nzb, float *vel);
/* extendvel_linear, functions.c:301-359: random "hybrid" border, libc rand(). */
int fdw_extendvel_linear(int nx, int nz, int nxb, int nzb, float *vel);
/* 7x7 Gaussian weights of ptsrc (ptsrc.c:51-56), row = x offset -3..3. */
int fdw_ptsrc_weights(float *w49);

/* input.dat, both dialects.  All reference keys; absent ints are -1, absent
 * floats -1.0, absent strings empty (then defaults as fd-code.cu:367-377 /
 * mod_main.cpp:76-85 are applied when apply_defaults != 0). */
typedef struct fdw_input {
    char tmpdir[512], vpfile[512], datfile[512], vel_ext_file[512];
    int nz, nx, nt, ns, sz, fsx, ds, gz, order, nzb, nxb, iss, rnd;
    float dz, dx, dt, fpeak, fac;
    int has_datfile, has_vel_ext_file;
} fdw_input;
/* GPU-family dialect: first line containing the key as a substring wins
 * (functions.c:10-75). */
int fdw_read_input_gpu(const char *path, int apply_defaults, fdw_input *out);
/* stencil program dialect (fd-source-code.cu:34-108): tmpdir is the input file. */
int fdw_read_input_stencil(const char *path, fdw_input *out);
/* CPU-family dialect: CWP getpar semantics on a par= file -- whitespace
 * separated name=value tokens, exact name match, last occurrence wins
 * (lib/cwp/src/par/lib/getpars.c:447-453). */
int fdw_read_input_cpu(const char *path, int apply_defaults, fdw_input *out);

/* raw float32 files of the drop-in surface (vpfile, datfile, vel_ext_file, input.bin, dir.image ...).
 * fdw_read_floats: like the reference's fread calls a short file leaves the rest of dst untouched; returns the
 * number of floats read, -1 when the file cannot be opened.  fdw_write_floats: append != 0 appends (the shot loop of
 * mod_main.cpp:163-166 writes datfile shot by shot); FDW_ERR_IO on failure. */
long long fdw_read_floats(const char *path, float *dst, long long n);
int fdw_write_floats(const char *path, const float *src, long long n, int append);
/* one shot of the image stack as rtm_code keeps it (fd-code.cu:521-528): img[ix][iz] += imloc[ix][iz] in the
 * reference's loop order and, when path is not NULL, the text block "======== is ========" followed by one
 * " %f " line per point appended to image.num (is == 0 truncates the file first). */
int fdw_image_stack_shot(const char *path, int is, int nx, int nz, float *img, const float *imloc);

/* ---------------------------------------------------------------- context */
const char *fdw_last_error(void);
int fdw_device_count(void);
int fdw_create(const fdw_params *prm, fdw_ctx **out);
void fdw_destroy(fdw_ctx *ctx);
/* run on an existing cudaStream_t (e.g. torch's current stream); NULL = own stream */
int fdw_set_stream(fdw_ctx *ctx, void *cuda_stream);
int fdw_sync(fdw_ctx *ctx);
/* squared velocity of the whole extended grid, host [nxe][nze] (vel2 of
 * fd-code.cu:490-494 / mod_main.cpp:120-126).  The slab's rows are taken. */
int fdw_set_v2(fdw_ctx *ctx, const float *v2);
/* source wavelet srce[nt] (fd-code.cu:402-403) */
int fdw_set_wavelet(fdw_ctx *ctx, const float *srce, int nt);
/* source position in extended-grid indices (sx[is], sz after "+= nzb") */
int fdw_set_source(fdw_ctx *ctx, int sx, int sz, int kind);
/* the two time levels of pair 0 (source field) or 1 (receiver field) */
int fdw_fields_zero(fdw_ctx *ctx, int pair);
int fdw_fields_upload(fdw_ctx *ctx, int pair, const float *newest, const float *older);
int fdw_fields_download(fdw_ctx *ctx, int pair, float *newest, float *older);
/* nsteps time levels of pair 0 starting at step index it0, with the context's
 * sponge and source: the loop body of fd_forward (fd-code.cu:259-284) for
 * family GPU, of rtm_main's forward loop (rtm_main.cpp:166-188) for family
 * CPU with FDW_SRC_POINT, of mod_main's (mod_main.cpp:146-168) with
 * FDW_SRC_GAUSS7 + FDW_TAPER_FOUR.  Asynchronous. */
int fdw_advance(fdw_ctx *ctx, int it0, int nsteps);

/* upload the two levels from host memory, run nsteps levels from step index
 * it0, download both levels back into the same host arrays: the body of
 * fd_forward (fd-code.cu:257-286) for arbitrary initial fields.  Synchronous. */
int fdw_propagate(fdw_ctx *ctx, float *newest, float *older, int it0, int nsteps);

/* ---------------------------------------------------------------- pipelines */
/* fd_forward (fd-code.cu:247-288): zero fields, nt steps, returns the two last
 * levels (P = older, PP = newest; either may be NULL to leave them on the
 * device for fdw_backward). */
int fdw_forward(fdw_ctx *ctx, int sx, int sz, float *P, float *PP);
/* fd_back (fd-code.cu:290-341): P/PP = the two saved levels (NULL,NULL = use
 * the ones fdw_forward left on the device), dobs = this shot's traces
 * [nx][nt], imloc [nx][nz] receives the shot image. */
int fdw_backward(fdw_ctx *ctx, const float *P, const float *PP, const float *dobs, int gz, float *imloc);
/* one shot of mod_main (mod_main.cpp:141-169): data [nx][nt]. */
int fdw_model_shot(fdw_ctx *ctx, int sx, int sz, int gz, float *data);
/* one shot of rtm_main (rtm_main.cpp:158-240): dobs_all = the whole
 * [ns][nx][nt] block (the reference's index quirk Q5 reaches into the next
 * trace), imloc [nx][nz].  Needs params.history = 1. */
int fdw_rtm_shot_cpu(fdw_ctx *ctx, int sx, int sz, int gz, const float *dobs_all, int ns, int is,
                     float *imloc);
/* ---- shot loop without host round trips (the reference's main() loop, fd-code.cu:480-529, re-staged for a GPU
 * that should never wait for the host): the next shot's velocity is uploaded and premultiplied on a copy stream
 * into a second buffer while the current shot runs; the shot image stays on the device and is stacked there in
 * call order (img += imloc, fd-code.cu:522-528 -- the same single float add per point, so the stack equals the
 * reference's sequential one bit for bit); one download at the end.  All calls are asynchronous; host arrays
 * should be pinned (pageable memory works but serialises). */
/* v2: whole extended grid [nxe][nze].  The copy is asynchronous: the array must stay valid and unchanged until it has
 * completed -- i.e. until a fdw_sync / fdw_stack_download that follows the matching fdw_v2_commit (the compute stream
 * waits for the copy there); a caller that refills host buffers per shot alternates two of them. */
int fdw_v2_stage(fdw_ctx *ctx, const float *v2);
int fdw_v2_commit(fdw_ctx *ctx);                  /* the staged velocity becomes the current one (stream-ordered) */
/* fd_back (fd-code.cu:290-341) on the levels fdw_forward left on the device; the shot image stays on the device */
int fdw_backward_device(fdw_ctx *ctx, const float *dobs, int gz);
int fdw_stack_zero(fdw_ctx *ctx);
int fdw_stack_add(fdw_ctx *ctx);                  /* stack += image of the last fdw_backward_device */
int fdw_stack_download(fdw_ctx *ctx, float *img); /* [nx][nz]; synchronous */
/* device address / row pitch (floats) / rows of the stack, for reductions straight from device memory (NCCL) */
int fdw_stack_devptr(fdw_ctx *ctx, void **ptr, long long *pitch, int *rows);

/* the stencil program (fd-source-code.cu:277-352): one Laplacian sweep of an
 * [nxe][nze] host array, ring of width order/2 = 0. */
int fdw_stencil(int order, int nxe, int nze, float dx, float dz, const float *in, float *out, int device);

/* image post-filter: the 2nd-order Laplacian of cuda_reference_RTM/models/3lay_mod/laplace.f90:24-28
 * applied to an [nx][nz] image on the GPU (one sweep; zero on the outermost ring).  rtm_code itself writes zeros
 * into dir.image_lap (fd-code.cu:477,542) -- bin/rtm_code does the same unless FDW_IMAGE_LAP=1 asks for the filter. */
int fdw_image_laplacian(int nx, int nz, float dx, float dz, const float *img, float *out, int device);

/* ---------------------------------------------------------------- slab decomposition
 * A context created with slab_x0 < slab_x1 owns those rows of the extended
 * grid plus GUARD ghost rows on each side.  One time level is then driven in
 * three phases so that the caller can overlap the halo exchange with the
 * interior update (the reference has no multi-GPU path; per-point arithmetic
 * is unchanged, so N slabs reproduce the 1-GPU result bit for bit):
 *   fdw_step_begin(it)            bookkeeping, builds the launch arguments
 *   fdw_step_rows(r0, r1, stream) update local rows [r0,r1) (any number of calls)
 *   fdw_step_end()                swaps the levels
 * fdw_halo_get returns, for the level being written (level=1, between begin and
 * end) or the newest level (level=0), the device addresses of the 4 boundary
 * row blocks (GUARD rows x pitch floats each, contiguous). */
typedef struct fdw_halo {
    void *send_lo, *send_hi; /* first / last GUARD owned rows */
    void *recv_lo, *recv_hi; /* ghost rows below row 0 / above row nloc-1 */
    long long count;         /* floats per block */
} fdw_halo;
/* shot phases for the split-phase API: fdw_shot_begin prepares a phase (zero fields, record /
 * history / image / trace buffers), the caller then drives params.nt levels with
 * fdw_step_begin/rows/end, and fdw_shot_end downloads THIS SLAB's part of the result:
 *   FDW_PHASE_MODEL    mod_main shot        -> traces of the owned interior x rows, [nli][nt]
 *   FDW_PHASE_RTM_FWD  rtm_main forward     -> nothing (history stays in HBM; no fdw_shot_end)
 *   FDW_PHASE_RTM_BWD  rtm_main backward    -> image rows of the owned interior x rows, [nli][nz]
 * (owned interior rows: fdw_devinfo.li0 / nli).  dobs_all is the whole [ns][nx][nt] block. */
#define FDW_PHASE_PLAIN 0
#define FDW_PHASE_MODEL 1
#define FDW_PHASE_RTM_FWD 2
#define FDW_PHASE_RTM_BWD 3
int fdw_shot_begin(fdw_ctx *ctx, int phase, int sx, int sz, int gz, const float *dobs_all, int ns, int is);
int fdw_shot_end(fdw_ctx *ctx, float *out);
/* all params.nt levels of the phase opened by fdw_shot_begin on a context that owns the whole grid
 * (the time loops of mod_main.cpp:146-168 / rtm_main.cpp:166-188,196-220), asynchronous: together with
 * fdw_mark_begin/end this gives the device time of a phase without its host transfers */
int fdw_shot_run(fdw_ctx *ctx);
int fdw_step_begin(fdw_ctx *ctx, int it);
int fdw_step_rows(fdw_ctx *ctx, int row0, int row1, void *cuda_stream);
int fdw_step_end(fdw_ctx *ctx);
int fdw_halo_get(fdw_ctx *ctx, int level, fdw_halo *out);
/* slab-local host I/O: arrays hold only the owned rows, [nloc][nze] */
int fdw_set_v2_local(fdw_ctx *ctx, const float *v2_rows);
int fdw_fields_upload_local(fdw_ctx *ctx, int pair, const float *newest, const float *older);
int fdw_fields_download_local(fdw_ctx *ctx, int pair, float *newest, float *older);
/* the same without the final synchronisation (pinned host arrays; fdw_sync before reading them):
 * lets a caller keep two contexts in flight so that one job's transfers overlap the other's levels */
int fdw_fields_download_local_async(fdw_ctx *ctx, int pair, float *newest, float *older);

/* ---------------------------------------------------------------- peer-memory halo exchange
 * (one process per GPU, NVLink / NVSwitch P2P; the reference has no multi-GPU path)
 * Instead of handing the boundary rows to a communication library, neighbouring slab contexts
 * map each other's buffers through CUDA IPC.  fdw_peer_levels then drives whole time levels:
 * the boundary-strip launch writes its rows locally and into the neighbour's ghost rows in the
 * same kernel, a release flag follows, the interior update overlaps the transfer and the next
 * level's boundary launch waits (on the device) for the neighbours' flags.  Results are bit for
 * bit those of fdw_step_begin/rows/end with any other exchange, and of the single domain.
 *   fdw_peer_export    this slab's handles (send to both neighbours, e.g. all_gather)
 *   fdw_peer_attach    map the lower / upper neighbour's buffers (NULL = grid edge)
 *   fdw_peer_refresh   push the newest level's boundary rows (after an upload); all slabs call it
 *   fdw_peer_levels    nsteps levels of the current phase (fdw_shot_begin), asynchronous
 *   fdw_peer_fence     stream-side wait until every push addressed to this slab has landed;
 *                      call before zeroing / uploading / downloading after fdw_peer_levels
 * Contract of a (re)initialisation: zeroing or uploading the fields of a slab-decomposed run is COLLECTIVE --
 * every slab initialises its buffers the same way (all zero, or all upload) and then calls fdw_peer_refresh,
 * which re-publishes its boundary rows into the neighbours' ghost rows and raises "my buffers are ready".
 * fdw_fields_zero leaves the ghost rows of an attached side alone (they belong to that neighbour's refresh,
 * which may land before or after the memset), so a refresh is never wiped; a slab that skips the refresh
 * leaves its neighbours with the previous run's ghost rows.
 * Device-side waits (acquire of the neighbours' flags, grid barriers of the persistent / tile kernels) give up
 * after FDW_TIMEOUT_MS of wall time (default 4000; %globaltimer, so independent of SM clocks and time slicing):
 * the strip is then neither updated nor pushed on, and the next synchronising call returns FDW_ERR_CUDA.  A context
 * whose level failed keeps consistent sponge counts / flag sequence; fdw_peer_attach clears the error state. */
#define FDW_IPC_HANDLE_BYTES 64
typedef struct fdw_peer_info {
    unsigned char field[4][FDW_IPC_HANDLE_BYTES];
    unsigned char flags[FDW_IPC_HANDLE_BYTES];
    long long pitch;
    int nloc, gx0, device, reserved;
} fdw_peer_info;
int fdw_peer_export(fdw_ctx *ctx, fdw_peer_info *out);
int fdw_peer_attach(fdw_ctx *ctx, const fdw_peer_info *lower, const fdw_peer_info *upper);
int fdw_peer_detach(fdw_ctx *ctx);
int fdw_peer_refresh(fdw_ctx *ctx);
int fdw_peer_levels(fdw_ctx *ctx, int it0, int nsteps);
int fdw_peer_fence(fdw_ctx *ctx);

/* ---------------------------------------------------------------- device-resident access
 * (benchmarks / multi-GPU plumbing; pointers are CUDA device pointers) */
typedef struct fdw_devinfo {
    void *newest, *older; /* local row 0, column 0 of pair 0 */
    void *vdt;
    long long pitch; /* floats */
    int nloc, gx0, nxe, nze, guard;
    int li0, nli; /* interior x rows owned by this context: global rows [li0, li0+nli) */
} fdw_devinfo;
int fdw_devinfo_get(fdw_ctx *ctx, fdw_devinfo *out);
/* device time of the work enqueued between mark_begin and mark_end, in ms
 * (CUDA events on the context's stream) */
int fdw_mark_begin(fdw_ctx *ctx);
int fdw_mark_end(fdw_ctx *ctx, float *ms);
/* number of kernels this context has launched so far */
long long fdw_launch_count(fdw_ctx *ctx);
/* other counters of the same kind (what actually ran: tests and benchmarks assert on them) */
#define FDW_COUNTER_LAUNCHES 0
#define FDW_COUNTER_GRAPH_REPLAYS 1   /* CUDA-graph replays of the peer-memory level loop */
#define FDW_COUNTER_PERSIST_LAUNCHES 2 /* phases run by the persistent (L2-resident) kernel */
#define FDW_COUNTER_TILE_LAUNCHES 3    /* phases run by the shared-memory tile kernel */
#define FDW_COUNTER_PSLAB_LAUNCHES 4   /* level runs of a thin slab done by ONE launch of the persistent slab kernel */
long long fdw_counter(fdw_ctx *ctx, int which);
/* standalone Laplacian on device-resident data of pair 0 (newest -> older), for benchmarks */
int fdw_laplacian_device(fdw_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FDWAVE_H */
