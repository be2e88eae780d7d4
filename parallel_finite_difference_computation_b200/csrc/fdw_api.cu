/*
 * fdw_api.cu -- context management and the C ABI of libfdwave (include/fdwave.h).
 *
 * Data layout in HBM (per context = per GPU = per slab):
 *   every wavefield level, the premultiplied velocity fl32(v2*dt2) and the
 *   image share one pitched layout: row = one x index, `pitch` floats per row
 *   (multiple of 32 => every row starts on a 128-byte line, float4 columns are
 *   aligned), nze valid columns followed by >= order/2 zero pad columns.
 *   AGUARD+1 (9) zero guard rows precede and follow the slab's rows; the
 *   innermost GUARD (4) of them are the ghost rows a slab edge exchanges with
 *   its neighbour (zeros on a physical grid edge).
 *   Levels are stored RAW; the sponge multiplications still pending on a level
 *   are counted in Field::pend and either applied on load by the sponge
 *   instantiation of the step kernel (fdw_step_core.h: small grids, slabs) or
 *   applied in place to the sponge regions right before the level's plain
 *   launch (sponge_inplace: whole grids), and before an export.
 */
#ifdef FDW_EMU
#include "emu_cuda.h" /* tests/emu: host stand-in for the CUDA runtime, unit tests only */
#else
#include <cuda_runtime.h>
#endif

#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fdwave.h"
#include "fdw_internal.h"
#include "fdw_step_core.h"

using fdw::GUARD;
using fdw::AGUARD;
using fdw::StepArgs;
using fdw::PersistArgs;
using fdw::TileArgs;
using fdw::PSlabArgs;

/* ------------------------------------------------------------------ errors */
static thread_local char g_err[512] = "";

extern "C" void fdw_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

extern "C" const char *fdw_last_error(void) { return g_err; }

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            fdw_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return FDW_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define CHECK(x)               \
    do {                       \
        int r_ = (x);          \
        if (r_ != FDW_OK) return r_; \
    } while (0)

/* ------------------------------------------------------------------ context */
struct Field {
    float *base = nullptr; /* allocation */
    float *r0 = nullptr;   /* local row 0, column 0 */
    int pend = 0;          /* sponge multiplications not yet applied to the stored values */
};

/* One kernel launch of a time level, as recorded for the CUDA-graph replay of the level loop
 * (fdw_peer_levels): what would have been launched, on which of the two streams, with which
 * arguments.  kind 0 = step kernel (StepArgs by value), 1 = peer acquire-wait, 2 = peer release. */
struct RecLaunch {
    const void *kern = nullptr;
    dim3 grid, block;
    int side = 0, kind = 0, level = 0; /* side: 0 main stream, 1 boundary chain, 2 sponge strips (graph lanes) */
    int par = 0;                       /* 1: independent of the previous launch of its lane (runs beside it) */
    StepArgs a;
    const unsigned *w_mine = nullptr; unsigned v = 0; int need_lo = 0, need_hi = 0; int *err = nullptr;
    unsigned long long tmo = 0;
    unsigned *s_lo = nullptr, *s_hi = nullptr;
};

struct fdw_ctx {
    fdw_params prm;
    int nxe, nze, gx0, nloc, H;
    long long pitch;
    size_t rows_alloc, field_elems;
    Field f[4];
    int newest[2], older[2];
    float *vdt_base = nullptr, *vdt = nullptr;
    float *vdt2_base = nullptr;          /* staging copy of the velocity (fdw_v2_stage / fdw_v2_commit) */
    float *stack = nullptr;              /* device-resident image stack, layout of img */
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_staged = nullptr, ev_shot_done = nullptr;
    bool staged = false, shot_done_valid = false;
    float *tz_base = nullptr, *tz = nullptr, *tx_base = nullptr, *tx = nullptr;
    float cz[fdw::MAX_ORDER + 1], cx[fdw::MAX_ORDER + 1], dz2inv, dx2inv, dt2;
    int tx_jlim, tz_ilim, tap_jlo, tap_jhi, tap_ilo, tap_ihi;
    int lap_i0, lap_i1, lap_j0, lap_j1, upd_i1, upd_j1, ncol4;
    std::vector<float> wavelet;
    int sx = 0, sz = 0, src_kind = FDW_SRC_POINT;
    float src_w[49];
    float *hist = nullptr, *img = nullptr, *dobs_d = nullptr, *rec_d = nullptr;
    size_t dobs_cap = 0, rec_cap = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr, ev_pfork = nullptr, ev_pjoin = nullptr;
    cudaStream_t side = nullptr;
    long long launches = 0;
    int nsm = 0;
    bool saved_valid = false;
    int rows_per_cta_override = 0, threads_override = 0;
    long long small_grid_limit = 1LL << 18, fork_limit = 1LL << 20; /* in float4 columns x rows */
    float *wavelet_d = nullptr; /* device copy of the wavelet (persistent kernel) */
    unsigned *barrier_d = nullptr;
    unsigned long long *ll_d = nullptr;  /* {value, tag} planes of the tile kernels' flag-in-data halo exchange (4 planes) */
    int use_ll = 1;                      /* FDW_TILE_LL=0: neighbour flags + ring loads instead */
    unsigned *tileflags_d = nullptr; /* neighbour flags of the tile kernels: FDW_TILE_FLAG_WORDS words */
    int *errflag_d = nullptr;
    int coop = 0;                      /* device supports cooperative launches */
    long long persist_limit = 1LL << 18; /* float4 columns x rows below which phases run persistently */
    int use_pslab = 1;                   /* persistent slab kernel for thin slabs (FDW_PSLAB=0: graph-replayed launches) */
    long long pslab_limit = 1LL << 20;   /* float4 columns x rows below which a slab's levels run in one launch */
    long long pslab_launches = 0;
    int pslab_threads = 0;               /* CTA width of the persistent slab kernel; 0 = chosen per slab (FDW_PSLAB_THREADS: 32..128, multiple of 32) */
    int use_tile = 1;                    /* shared-memory tile kernel for those grids (FDW_TILE=0: L2-resident persistent kernel) */
    int smem_optin = 0;                  /* largest dynamic shared memory per CTA */
    long long persist_launches = 0;
    long long tile_launches = 0;
    int li0 = 0, nli = 0; /* interior x rows owned by this slab: global rows [li0, li0+nli) */
    /* split-phase shot state (slab decomposition) */
    int phase = 0, shot_gz = 0, shot_is = 0, shot_ns = 1;
    /* peer-memory halo exchange (slab decomposition over NVLink P2P) */
    struct PeerSide {
        bool on = false;
        float *fbase[4] = {nullptr, nullptr, nullptr, nullptr}; /* the neighbour's four field allocations, mapped here */
        unsigned *flags = nullptr;                              /* the neighbour's flag block, mapped here */
        int nloc = 0;
    } peer[2];                 /* [0] lower neighbour (rows below gx0), [1] upper neighbour */
    unsigned *flags_d = nullptr; /* written by the neighbours: [0] by the lower one, [1] by the upper one */
    unsigned *pcount_d = nullptr; /* CTAs of the current level's boundary launches that have finished */
    int fuse_flags = 1;          /* acquire/release inside the boundary kernels instead of separate launches */
    unsigned long long timeout_ns = 4000000000ull; /* device-side waits give up after this long (FDW_TIMEOUT_MS) */
    unsigned peer_seq = 0;       /* boundary-row pushes issued so far (lock-step on all slabs) */
    long long peer_waits = 0;
    /* CUDA-graph replay of the level loop: when `rec` is set, launches are recorded instead of issued */
    std::vector<RecLaunch> *rec = nullptr;
    int rec_level = 0, rec_lane = -1, rec_par = 0;
    cudaGraph_t lgraph = nullptr;
    cudaGraphExec_t lexec = nullptr;
    std::vector<cudaGraphNode_t> lnodes;   /* kernel nodes, in RecLaunch order */
    std::vector<RecLaunch> lshape;         /* the launch list the graph was built from */
    int use_graph = 1;
    long long graph_replays = 0;
    /* The replayed graph pays off where the HOST is the limit: ~12 driver calls per slab level from each of N
     * processes.  Measured: 8 B200s, 16384^2 per GPU, per-level cost of the slab machinery over one GPU: +21 us
     * replayed (round 1), +35 us with direct launches (profiles/r02t_bench_line_n8_direct_launches.json); the
     * 16384^2 grid cut over 8 GPUs: 121 us per level replayed, 126-136 direct.  On 2 GPUs with 4136 x 2128 slabs the
     * two are within the run-to-run noise of each other (profiles/r02q_, r02r_slab2_sweep*.log).  So the slab loop is
     * always replayed; graph_limit (float4 columns x rows, FDW_GRAPH_LIMIT) can restrict that to thin slabs.
     * graph_levels: levels per graph launch (FDW_GRAPH_LEVELS, even; 8 or 16 levels per launch measured no better
     * than 2 -- refreshing 40+ node arguments then takes the host as long as the levels run;
     * profiles/r02r_slab2_sweep_graph_levels.log) */
    long long graph_limit = 1LL << 62;
    int graph_levels = 2;
    int use_multirect = 1;       /* the sponge strips of a level in ONE launch (FDW_MULTIRECT=0: one launch per strip) */
    int use_pdl = 1;             /* programmatic dependent launch of the step kernels (FDW_PDL=0: plain launches) */
    bool pdl_now = false;        /* set around launches that follow their predecessor kernel directly in ONE stream (no
                                  * fork / join events, no slab boundary chain): only there is the attribute used */
    int inplace_sponge = 1;      /* mid-size whole grids: pending sponge passes applied in place before a plain launch */
    long long inplace_limit = 1LL << 62; /* float4 columns x rows below which that is done (FDW_SPONGE_INPLACE_LIMIT) */
    /* single-GPU level loop as a replayed CUDA graph (pairs of levels, node arguments refreshed), for grids whose
     * level is tens of microseconds.  OFF by default (FDW_LEVEL_GRAPH=1): measured on the 8272 x 2128 and 8272 x 4176
     * grids it gains 1 us per level with no sponge strips and LOSES 1.5-6 us with them (the stream-ordered launches
     * already keep the GPU fed at 60-130 us per level; profiles/r02n_level_graph_and_chunking.log) */
    int level_graph = 0;
    long long level_graph_limit = 1LL << 25; /* float4 columns x rows */
    /* split-phase step (slab decomposition) */
    bool step_open = false;
    StepArgs step_args;
    const void *step_kern = nullptr;
};

static int bind(fdw_ctx *c)
{
    CU(cudaSetDevice(c->prm.device));
    return FDW_OK;
}

/* ------------------------------------------------------------------ small kernels */
__global__ void k_scale_rows(float *v, long long n, float s)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = __fmul_rn(v[i], s);
}

/* apply the sponge cnt times in place (kernel_tapper / taper_apply as a
 * stand-alone pass; used only to materialise pending factors before export) */
__global__ void k_materialize(float *r0, long long pitch, int nze, int row_lo, int row_hi, int grow0,
                              const float *tz, const float *tx, int tx_jlim, int tz_ilim, int cnt)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int lr = row_lo + blockIdx.y;
    if (j >= nze || lr >= row_hi) return;
    float v = r0[(long long)lr * pitch + j];
    const float zf = (grow0 + lr < tz_ilim) ? tz[j] : 1.0f;
    const float xf = (j < tx_jlim) ? tx[lr] : 1.0f;
    for (int c = 0; c < cnt; c++) v = __fmul_rn(__fmul_rn(v, zf), xf);
    r0[(long long)lr * pitch + j] = v;
}

/* The sponge as an in-place pass over the regions where a factor differs from 1 (up to four rectangles), both
 * time levels in one launch: what kernel_tapper / taper_apply do, restricted to where it matters.  Used before the
 * update of a mid-size whole-grid level, so that ONE plain launch covers the whole grid (see sponge_inplace). */
struct SpongeRects {
    float *f[2];
    int cnt[2];
    long long pitch;
    int grow0, tx_jlim, tz_ilim, n;
    const float *tz, *tx;
    struct { int c0, c1, r0, r1; unsigned e0; } r[4];
    unsigned total; /* < 2^31 elements (checked by the caller) */
};
enum { SPONGE_INPLACE_PER_THREAD = 4 }; /* elements per thread, grid-stride: all loads in flight before the first store */
__global__ void k_sponge_inplace(const __grid_constant__ SpongeRects a)
{
    fdw::pdl_trigger();
    fdw::pdl_wait();
    const unsigned stride = gridDim.x * blockDim.x;
    float v[SPONGE_INPLACE_PER_THREAD][2], zf[SPONGE_INPLACE_PER_THREAD], xf[SPONGE_INPLACE_PER_THREAD];
    long long off[SPONGE_INPLACE_PER_THREAD];
#pragma unroll
    for (int q = 0; q < SPONGE_INPLACE_PER_THREAD; q++) {
        unsigned e = blockIdx.x * blockDim.x + threadIdx.x + (unsigned)q * stride;
        off[q] = -1;
        if (e >= a.total) continue;
        int k = 0;
        while (k + 1 < a.n && e >= a.r[k + 1].e0) k++;
        e -= a.r[k].e0;
        const unsigned w = (unsigned)(a.r[k].c1 - a.r[k].c0);
        const int lr = a.r[k].r0 + (int)(e / w), j = a.r[k].c0 + (int)(e % w);
        zf[q] = (a.grow0 + lr < a.tz_ilim) ? a.tz[j] : 1.0f;
        xf[q] = (j < a.tx_jlim) ? a.tx[lr] : 1.0f;
        off[q] = (long long)lr * a.pitch + j;
        for (int s = 0; s < 2; s++)
            if (a.f[s] && a.cnt[s] > 0) v[q][s] = a.f[s][off[q]];
    }
#pragma unroll
    for (int q = 0; q < SPONGE_INPLACE_PER_THREAD; q++) {
        if (off[q] < 0) continue;
        for (int s = 0; s < 2; s++) {
            if (!a.f[s] || a.cnt[s] <= 0) continue;
            float x = v[q][s];
            for (int c = 0; c < a.cnt[s]; c++) x = __fmul_rn(__fmul_rn(x, zf[q]), xf[q]);
            a.f[s][off[q]] = x;
        }
    }
}

/* peer-memory halo exchange: release / acquire of the "boundary rows delivered" counters.  The
 * signal kernel runs after the boundary-strip launch in stream order (its peer stores are
 * complete at the kernel boundary); the system-scope fence orders them before the flag. */
__global__ void k_peer_signal(unsigned *flag_lo, unsigned *flag_hi, unsigned v)
{
#ifndef FDW_EMU
    __threadfence_system();
#else
    __sync_synchronize();
#endif
    if (flag_lo) *(volatile unsigned *)flag_lo = v;
    if (flag_hi) *(volatile unsigned *)flag_hi = v;
#ifndef FDW_EMU
    __threadfence_system();
#endif
}

__global__ void k_peer_wait(const unsigned *mine, unsigned v, int need_lo, int need_hi, int *error_flag,
                            unsigned long long timeout_ns)
{
#ifdef FDW_EMU
    /* the host stand-in runs everything synchronously: an unmet flag is a protocol error */
    if ((need_lo && !emu_wait_flag(mine + 0, v)) || (need_hi && !emu_wait_flag(mine + 1, v))) *error_flag = 2;
#else
    const unsigned long long t0 = fdw::wall_ns();
    for (int s = 0; s < 2; s++) {
        if (!(s == 0 ? need_lo : need_hi)) continue;
        unsigned spins = 0;
        while (*((volatile const unsigned *)mine + s) < v) {
            if ((++spins & 255u) == 0 && fdw::wall_ns() - t0 > timeout_ns) { /* never hang the GPU on a lost neighbour */
                atomicExch(error_flag, 2);
                return;
            }
        }
    }
    __threadfence_system();
#endif
}

#ifdef FDW_EMU
static void thunk_peer_signal(void **a) { k_peer_signal(*(unsigned **)a[0], *(unsigned **)a[1], *(unsigned *)a[2]); }
static void thunk_peer_wait(void **a)
{
    k_peer_wait(*(const unsigned **)a[0], *(unsigned *)a[1], *(int *)a[2], *(int *)a[3], *(int **)a[4], *(unsigned long long *)a[5]);
}
static void thunk_scale_rows(void **a) { k_scale_rows(*(float **)a[0], *(long long *)a[1], *(float *)a[2]); }
static void thunk_sponge_inplace(void **a) { k_sponge_inplace(*(const SpongeRects *)a[0]); }
static void thunk_materialize(void **a)
{
    k_materialize(*(float **)a[0], *(long long *)a[1], *(int *)a[2], *(int *)a[3], *(int *)a[4], *(int *)a[5],
                  *(const float **)a[6], *(const float **)a[7], *(int *)a[8], *(int *)a[9], *(int *)a[10]);
}
#define FDW_KPTR(kernel, thunk) ((const void *)&thunk)
#else
#define FDW_KPTR(kernel, thunk) ((const void *)kernel)
#endif

/* ------------------------------------------------------------------ helpers */
/* launch of a kernel that begins with pdl_wait() (the step kernels, the in-place sponge pass): with programmatic
 * dependent launch its CTAs may be scheduled while the previous kernel of the stream drains (FDW_PDL=0: off) */
static cudaError_t launch_pdl(const fdw_ctx *c, const void *k, dim3 grid, dim3 block, void **params, cudaStream_t st);

static const void *step_kernel(int order, int recipe, int epi, int sponge)
{
    switch (order) {
    case 2: return fdw_step_kernel_o2(recipe, epi, sponge);
    case 4: return fdw_step_kernel_o4(recipe, epi, sponge);
    case 6: return fdw_step_kernel_o6(recipe, epi, sponge);
    case 8: return fdw_step_kernel_o8(recipe, epi, sponge);
    case 10: return fdw_step_kernel_o10(recipe, epi, sponge);
    case 12: return fdw_step_kernel_o12(recipe, epi, sponge);
    case 14: return fdw_step_kernel_o14(recipe, epi, sponge);
    case 16: return fdw_step_kernel_o16(recipe, epi, sponge);
    }
    return nullptr;
}

static const void *persist_kernel(int order, int recipe, int epi)
{
    switch (order) {
    case 2: return fdw_persist_kernel_o2(recipe, epi);
    case 4: return fdw_persist_kernel_o4(recipe, epi);
    case 6: return fdw_persist_kernel_o6(recipe, epi);
    case 8: return fdw_persist_kernel_o8(recipe, epi);
    case 10: return fdw_persist_kernel_o10(recipe, epi);
    case 12: return fdw_persist_kernel_o12(recipe, epi);
    case 14: return fdw_persist_kernel_o14(recipe, epi);
    case 16: return fdw_persist_kernel_o16(recipe, epi);
    }
    return nullptr;
}

static const void *tile_kernel(int order, int recipe, int epi)
{
    switch (order) {
    case 2: return fdw_tile_kernel_o2(recipe, epi);
    case 4: return fdw_tile_kernel_o4(recipe, epi);
    case 6: return fdw_tile_kernel_o6(recipe, epi);
    case 8: return fdw_tile_kernel_o8(recipe, epi);
    case 10: return fdw_tile_kernel_o10(recipe, epi);
    case 12: return fdw_tile_kernel_o12(recipe, epi);
    case 14: return fdw_tile_kernel_o14(recipe, epi);
    case 16: return fdw_tile_kernel_o16(recipe, epi);
    }
    return nullptr;
}

static const void *pslab_kernel(int order, int recipe, int epi)
{
    switch (order) {
    case 2: return fdw_pslab_kernel_o2(recipe, epi);
    case 4: return fdw_pslab_kernel_o4(recipe, epi);
    case 6: return fdw_pslab_kernel_o6(recipe, epi);
    case 8: return fdw_pslab_kernel_o8(recipe, epi);
    case 10: return fdw_pslab_kernel_o10(recipe, epi);
    case 12: return fdw_pslab_kernel_o12(recipe, epi);
    case 14: return fdw_pslab_kernel_o14(recipe, epi);
    case 16: return fdw_pslab_kernel_o16(recipe, epi);
    }
    return nullptr;
}

static const void *lap_kernel(int order)
{
    switch (order) {
    case 2: return fdw_lap_kernel_o2();
    case 4: return fdw_lap_kernel_o4();
    case 6: return fdw_lap_kernel_o6();
    case 8: return fdw_lap_kernel_o8();
    case 10: return fdw_lap_kernel_o10();
    case 12: return fdw_lap_kernel_o12();
    case 14: return fdw_lap_kernel_o14();
    case 16: return fdw_lap_kernel_o16();
    }
    return nullptr;
}

/* row pitch: >= order/2 zero pad columns after the valid ones (they absorb the z-neighbour loads of the edge threads
 * on both sides: the left neighbours of column 0 are the previous row's pad), whole 128-byte lines */
static long long pitch_for(int nze, int order) { return ((long long)nze + (order > 8 ? 8 : 4) + 31) / 32 * 32; }

/* launch geometry: one thread per float4 column; CTAs tile z, x is cut into chunks of rows_per_cta
 * rows.  Measured on B200, interleaved in one process (tools/sweep_geometry.py, profiles/
 * r01_sweep_geometry.log): SMALL work units win -- one-warp CTAs x 7 rows: 360 Gpts/s, 64 x 7: 354,
 * 128 x 8: 339, 256 x 32: 317 on the same GPU(s).  The 8 halo rows a chunk re-reads are L1/L2 hits, and with
 * ~300 k short CTAs per level neighbouring chunks run close together in time, so the L2 hit rate of those
 * rows rises (ncu: 42 % vs 33 %) and DRAM traffic falls from 1.05x to 0.99x of the algorithmic 16 B/point.
 * A warp is the natural unit of this kernel anyway: threads never exchange data or meet at a barrier. */
enum { FDW_CTA_THREADS = 32, FDW_CTA_ROWS = 7 };
static int cached_occupancy(const void *kern, int nthreads)
{
    /* the occupancy query costs microseconds per call: remember it per (kernel, block size) */
    static thread_local struct { const void *k; int nt, occ; } cache[64];
    static thread_local int ncache = 0;
    for (int i = 0; i < ncache; i++)
        if (cache[i].k == kern && cache[i].nt == nthreads) return cache[i].occ;
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthreads, 0) != cudaSuccess || occ < 1) occ = 1;
    if (ncache < 64) { cache[ncache].k = kern; cache[ncache].nt = nthreads; cache[ncache].occ = occ; ncache++; }
    return occ;
}

/* Rows per CTA of a launch that does not fill the machine many times over.  Such a launch is a latency chain: a
 * CTA loads its 8-row x window, then takes one dependent memory round trip per row, and the launch lasts as many
 * of those chains as it has waves of resident CTAs -- cost(rpc) = max(1, CTAs(rpc) / resident CTAs) x (8 + rpc),
 * minimised over rpc = 2..7.  Large grids get 7 (the measured optimum of the HBM-bound regime), launches below
 * one wave get 2 (shortest chain), and the sponge strips of a mid-size grid -- a few thousand CTAs -- get the
 * chunking that keeps them to about one wave (measured on the 8272 x 2128 grid: strips at 2 rows per CTA take 20 us
 * each in 1.5 waves; profiles/r02l_level_sweep.log).  nrect rectangles of one launch are chunked together. */
static int g_rpc_rule = 1; /* FDW_RPC_RULE=0: the earlier rule (halve 7 -> 4 -> 2 until the grid fills the machine twice) */
/* fold[i] (optional): row chunks one CTA of rectangle i holds side by side (StepArgs::rect[].lw) */
static int pick_rows_per_cta(long long cap, const int *gx, const int *rows, int nrect, const int *fold = nullptr)
{
    auto ctas_of = [&](int i, int rpc) {
        const int chunks = (rows[i] + rpc - 1) / rpc, f = fold ? fold[i] : 1;
        return (long long)gx[i] * ((chunks + f - 1) / f);
    };
    if (!g_rpc_rule) {
        int rpc = FDW_CTA_ROWS;
        long long ctas;
        do {
            ctas = 0;
            for (int i = 0; i < nrect; i++) ctas += ctas_of(i, rpc);
            if (rpc <= 2 || ctas >= 2 * cap) break;
            rpc = (rpc + 1) / 2;
        } while (true);
        return rpc;
    }
    /* a few waves (mid-size grids, e.g. 8272 x 2128 = 4.2 waves of 7-row CTAs): a nearly empty last wave costs a whole
     * CTA lifetime at low occupancy -- among 5 / 6 / 7 rows take the chunking whose last wave is fullest, weighed
     * against the halo rows a shorter chunk re-reads (measured on that grid: 7 rows 63.6 us per level, 6 rows
     * 61.7, 5 rows 61.7; profiles/r02m_c5_grid_level_sweep_folded_strips.log) */
    {
        long long c7 = 0;
        for (int i = 0; i < nrect; i++) c7 += ctas_of(i, FDW_CTA_ROWS);
        const double w7 = (double)c7 / (double)cap;
        if (w7 > 1.5 && w7 < 12.0) {
            int best = FDW_CTA_ROWS;
            double best_eff = -1.0;
            for (int rpc = FDW_CTA_ROWS; rpc >= 5; rpc--) {
                long long ctas = 0;
                for (int i = 0; i < nrect; i++) ctas += ctas_of(i, rpc);
                const double w = (double)ctas / (double)cap;
                const double eff = w / (double)(long long)(w + 0.999999) * rpc / (rpc + 1.0);
                if (eff > best_eff + 0.01) { best_eff = eff; best = rpc; }
            }
            return best;
        }
        if (w7 >= 12.0) return FDW_CTA_ROWS;
    }
    int best = FDW_CTA_ROWS;
    double best_cost = -1.0;
    for (int rpc = FDW_CTA_ROWS; rpc >= 2; rpc--) {
        long long ctas = 0;
        for (int i = 0; i < nrect; i++) ctas += ctas_of(i, rpc);
        const double waves = (double)ctas / (double)cap;
        const double cost = (waves > 1.0 ? waves : 1.0) * (8 + rpc);
        if (best_cost < 0.0 || cost < best_cost) { best_cost = cost; best = rpc; }
    }
    return best;
}

static void launch_geometry(const void *kern, int nsm, int ncols, int rows, int thr_override, int rpc_override,
                            dim3 *grid, dim3 *block, int *rows_per_cta)
{
    int nthreads = ncols >= FDW_CTA_THREADS ? FDW_CTA_THREADS : ((ncols + 31) / 32) * 32;
    if (thr_override > 0) nthreads = thr_override;
    if (nthreads < 32) nthreads = 32;
    int gx = (ncols + nthreads - 1) / nthreads;
    int rpc = FDW_CTA_ROWS;
    if (rpc_override > 0) rpc = rpc_override;
    else rpc = pick_rows_per_cta((long long)nsm * cached_occupancy(kern, nthreads), &gx, &rows, 1);
    while ((rows + rpc - 1) / rpc > 65535) rpc *= 2; /* gridDim.y limit */
    int gy = (rows + rpc - 1) / rpc;
    *grid = dim3(gx, gy < 1 ? 1 : gy, 1);
    *block = dim3(nthreads, 1, 1);
    *rows_per_cta = rpc;
}

static void base_args(const fdw_ctx *c, int pair, StepArgs *a)
{
    memset(a, 0, sizeof(*a));
    const Field &n = c->f[c->newest[pair]], &o = c->f[c->older[pair]];
    a->p = n.r0;
    a->pp = o.r0;
    a->vdt = c->vdt;
    a->pitch = c->pitch;
    a->apitch = c->pitch;
    a->col4_0 = 0;
    a->ncol4 = c->ncol4;
    a->grow0 = c->gx0;
    a->lap_i0 = c->lap_i0; a->lap_i1 = c->lap_i1; a->lap_j0 = c->lap_j0; a->lap_j1 = c->lap_j1;
    a->nze = c->nze;
    memcpy(a->cz, c->cz, sizeof a->cz);
    memcpy(a->cx, c->cx, sizeof a->cx);
    a->dz2inv = c->dz2inv;
    a->dx2inv = c->dx2inv;
    a->one = 1.0f;
    a->taper_on = 0;
    a->np = n.pend;
    a->no = o.pend;
    a->tz = c->tz;
    a->tx = c->tx;
    a->tx_jlim = c->tx_jlim; a->tz_ilim = c->tz_ilim;
    a->tap_jlo = c->tap_jlo; a->tap_jhi = c->tap_jhi; a->tap_ilo = c->tap_ilo; a->tap_ihi = c->tap_ihi;
    memcpy(a->src_w, c->src_w, sizeof a->src_w);
    a->row0 = 0;
    int hi = c->upd_i1 - c->gx0;
    a->row1 = hi < c->nloc ? (hi < 0 ? 0 : hi) : c->nloc;
}

/* One time level = one launch of the plain kernel on the sponge-free
 * rectangle plus launches of the sponge kernel on the thin strips (columns)
 * and row bands where some factor differs from 1.  The strips run on a side
 * stream, forked from and joined to `st` with events, so they overlap the
 * bulk launch instead of serialising behind it. */
struct Rect { int c0, c1, r0, r1, sponge; };

static cudaError_t launch_pdl(const fdw_ctx *c, const void *k, dim3 grid, dim3 block, void **params, cudaStream_t st)
{
#ifndef FDW_EMU
    if (c->use_pdl && c->pdl_now) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        memset(at, 0, sizeof at);
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelExC(&cfg, k, params);
    }
#endif
    return cudaLaunchKernel(k, grid, block, params, 0, st);
}

static int launch_rect(fdw_ctx *c, const StepArgs &base, int recipe, int epi, const Rect &rc, cudaStream_t st)
{
    if (rc.c1 <= rc.c0 || rc.r1 <= rc.r0) return FDW_OK;
    const void *k = step_kernel(c->prm.order, recipe, epi, rc.sponge);
    if (!k) {
        fdw_set_error("no kernel for order %d recipe %d epilogue %d", c->prm.order, recipe, epi);
        return FDW_ERR_UNSUPPORTED;
    }
    StepArgs a = base;
    a.col4_0 = rc.c0; a.ncol4 = rc.c1; a.row0 = rc.r0; a.row1 = rc.r1;
    dim3 grid, block;
    launch_geometry(k, c->nsm, rc.c1 - rc.c0, rc.r1 - rc.r0, c->threads_override, c->rows_per_cta_override, &grid,
                    &block, &a.rows_per_cta);
    c->launches++;
    if (c->rec) {
        RecLaunch r;
        r.kern = k; r.grid = grid; r.block = block; r.kind = 0;
        r.side = c->rec_lane >= 0 ? c->rec_lane : (st == c->side && st != c->stream);
        r.par = c->rec_par;
        r.level = c->rec_level; r.a = a;
        c->rec->push_back(r);
        return FDW_OK;
    }
    void *params[] = {&a};
    if (epi & fdw::EPI_PUSH) CU(cudaLaunchKernel(k, grid, block, params, 0, st)); /* boundary chain: acquire / release inside */
    else CU(launch_pdl(c, k, grid, block, params, st));
    return FDW_OK;
}

/* several disjoint rectangles of one level in ONE launch of the sponge kernel (StepArgs::nrect): the strips are
 * latency-bound short kernels (a few thousand one-warp CTAs, ~10-20 us each on an 8272 x 2128 grid) -- one after the
 * other on the side stream they outlast the bulk launch they are meant to hide behind (measured: 4 strips
 * 20+20+10+10 us against a 53 us bulk, profiles/r02k_c5_launches.csv); side by side they take as long as the
 * slowest.  Returns FDW_ERR_UNSUPPORTED (nothing launched) when the rectangles cannot share a launch. */
static int launch_rects(fdw_ctx *c, const StepArgs &base, int recipe, int epi, const Rect *rc, int n, cudaStream_t st)
{
    const void *k = step_kernel(c->prm.order, recipe, epi, 1);
    if (!k || n > (int)fdw::MAX_RECTS) return FDW_ERR_UNSUPPORTED;
    StepArgs a = base;
    a.nrect = 0;
    unsigned threads = 0;
    int gxs[fdw::MAX_RECTS], rws[fdw::MAX_RECTS], folds[fdw::MAX_RECTS];
    for (int i = 0; i < n; i++) {
        if (rc[i].c1 <= rc[i].c0 || rc[i].r1 <= rc[i].r0) continue;
        if (!rc[i].sponge) return FDW_ERR_UNSUPPORTED;
        dim3 grid, block;
        int rpc;
        launch_geometry(k, c->nsm, rc[i].c1 - rc[i].c0, rc[i].r1 - rc[i].r0, c->threads_override, c->rows_per_cta_override,
                        &grid, &block, &rpc);
        if (threads && block.x != threads) return FDW_ERR_UNSUPPORTED;
        threads = block.x;
        /* lanes side by side in z: the fold that pads the rectangle's width least (ties: the wider) */
        const int ncols = rc[i].c1 - rc[i].c0;
        int lw = (int)threads;
        for (int w = (int)threads / 2; w >= 8 && threads % w == 0; w /= 2)
            if ((ncols + w - 1) / w * w < (ncols + lw - 1) / lw * lw) lw = w;
        StepArgs::RectGeom &g = a.rect[a.nrect];
        g.c0 = rc[i].c0; g.c1 = rc[i].c1; g.r0 = rc[i].r0; g.r1 = rc[i].r1; g.rpc = rpc; g.lw = lw;
        g.nbx = (ncols + lw - 1) / lw;
        gxs[a.nrect] = g.nbx; rws[a.nrect] = rc[i].r1 - rc[i].r0; folds[a.nrect] = (int)threads / lw;
        a.nrect++;
    }
    if (a.nrect < 1) return FDW_ERR_UNSUPPORTED;
    /* the rectangles share the machine: one chunking for all of them */
    const int rpc_all = c->rows_per_cta_override > 0
                            ? c->rows_per_cta_override
                            : pick_rows_per_cta((long long)c->nsm * cached_occupancy(k, (int)threads), gxs, rws, a.nrect, folds);
    long long total = 0;
    for (int i = 0; i < a.nrect; i++) {
        StepArgs::RectGeom &g = a.rect[i];
        const int fold = (int)threads / g.lw, chunks = (g.r1 - g.r0 + rpc_all - 1) / rpc_all;
        g.rpc = rpc_all;
        g.cta0 = (int)total;
        total += (long long)g.nbx * ((chunks + fold - 1) / fold);
    }
    if (total > 0x7fffffffLL) return FDW_ERR_UNSUPPORTED;
    a.col4_0 = a.rect[0].c0; a.ncol4 = a.rect[0].c1; a.row0 = a.rect[0].r0; a.row1 = a.rect[0].r1; a.rows_per_cta = a.rect[0].rpc;
    const dim3 grid((unsigned)total, 1, 1), block(threads, 1, 1);
    c->launches++;
    if (c->rec) {
        RecLaunch r;
        r.kern = k; r.grid = grid; r.block = block; r.kind = 0;
        r.side = c->rec_lane >= 0 ? c->rec_lane : (st == c->side && st != c->stream);
        r.par = c->rec_par;
        r.level = c->rec_level; r.a = a;
        c->rec->push_back(r);
        return FDW_OK;
    }
    void *params[] = {&a};
    CU(launch_pdl(c, k, grid, block, params, st));
    return FDW_OK;
}

/* number of CTAs launch_rect will use for this rectangle (same geometry code) */
static long long rect_ctas(fdw_ctx *c, int recipe, int epi, const Rect &rc)
{
    if (rc.c1 <= rc.c0 || rc.r1 <= rc.r0) return 0;
    const void *k = step_kernel(c->prm.order, recipe, epi, rc.sponge);
    if (!k) return 0;
    dim3 grid, block;
    int rpc;
    launch_geometry(k, c->nsm, rc.c1 - rc.c0, rc.r1 - rc.r0, c->threads_override, c->rows_per_cta_override, &grid,
                    &block, &rpc);
    return (long long)grid.x * grid.y;
}

/* A bulk rectangle whose width is not a multiple of a WIDE CTA leaves the last CTA column nearly empty
 * while it still occupies a residency slot for the whole launch (measured with 256-thread CTAs on the
 * 8192 x 4096 RTM grid, 1044 float4 columns: 5 CTA columns instead of 4.08 -> +24 % launch time).  Cut
 * the ragged tail off into its own narrow rectangle; it runs on the side stream with the strips.  With
 * the default 64-thread CTAs the tail is at most one idle warp and nothing is cut. */
static bool split_ragged_tail(const fdw_ctx *c, Rect *bulk, Rect *tail)
{
    const int w = c->threads_override > 0 ? c->threads_override : FDW_CTA_THREADS;
    const int n = bulk->c1 - bulk->c0, r = n % w;
    if (w < 256 || n <= w || r == 0 || r > w / 2) return false;
    *tail = *bulk;
    tail->c0 = bulk->c1 - r;
    bulk->c1 -= r;
    return true;
}

static int launch_level(fdw_ctx *c, const StepArgs &base, int recipe, int epi, int row0, int row1, cudaStream_t st)
{
    if (row0 < base.row0) row0 = base.row0;
    if (row1 > base.row1) row1 = base.row1;
    if (row1 <= row0 || base.ncol4 <= 0) return FDW_OK;
    const int nc = base.ncol4, H = c->H;
    Rect side[5];
    int nside = 0;
    Rect bulk = {0, nc, row0, row1, 0};
    if (base.taper_on) {
        if ((long long)nc * (row1 - row0) < c->small_grid_limit) {
            /* launch-bound regime (shipped models are ~0.1 Mpoint): one launch of the sponge kernel over
             * everything -- its factors are exactly 1.0 outside the sponge, so the result is unchanged */
            Rect all = {0, nc, row0, row1, 1};
            return launch_rect(c, base, recipe, epi, all, st);
        }
        /* columns whose own samples or z neighbours (whole float4 blocks: +-4, +-8 above order 8) sit in a z
         * sponge, in whole warps */
        const int zr = H > 4 ? 8 : 4;
        int cs = 0, cb = nc;
        /* (strip launches fold their CTAs down to 8 lanes in z, see StepArgs::rect; the bulk's warps then start on a
         * 128-byte line all the same) */
        const int gran = c->use_multirect ? 8 : 32;
        if (c->tap_jlo > INT_MIN) cs = (((c->tap_jlo + zr + 3) / 4) + gran - 1) / gran * gran;
        if (c->tap_jhi < INT_MAX) cb = ((c->tap_jhi - 3 - zr > 0 ? c->tap_jhi - 3 - zr : 0) / 4) / gran * gran;
        if (cs > nc) cs = nc;
        if (cb < cs) cb = cs;
        /* rows whose x window touches an x sponge that applies to every column */
        int rl = row0, rh = row1;
        if (c->tap_ilo > INT_MIN) { int v = c->tap_ilo + H - c->gx0; rl = v < row0 ? row0 : (v > row1 ? row1 : v); }
        if (c->tap_ihi < INT_MAX) { int v = c->tap_ihi - H - c->gx0; rh = v < rl ? rl : (v > row1 ? row1 : v); }
        const Rect sponge[4] = {{0, cs, row0, row1, 1}, {cb, nc, row0, row1, 1}, {cs, cb, row0, rl, 1}, {cs, cb, rh, row1, 1}};
        for (int k = 0; k < 4; k++) side[nside++] = sponge[k];
        bulk = Rect{cs, cb, rl, rh, 0};
    }
    /* fork only when the bulk launch is long enough to be worth two event operations */
    const bool fork = bulk.c1 > bulk.c0 && bulk.r1 > bulk.r0 &&
                      (long long)(bulk.c1 - bulk.c0) * (bulk.r1 - bulk.r0) >= c->fork_limit;
    if (fork && split_ragged_tail(c, &bulk, &side[nside])) nside++;
    bool any_side = false;
    for (int k = 0; k < nside; k++) any_side = any_side || (side[k].c1 > side[k].c0 && side[k].r1 > side[k].r0);
    const bool do_fork = fork && any_side;
    c->pdl_now = !do_fork && !c->step_open && !c->rec;
    cudaStream_t ss = do_fork ? c->side : st;
    const bool ev = do_fork && !c->rec; /* a recorded level gets its fork/join as graph edges instead */
    if (ev) {
        CU(cudaEventRecord(c->ev_fork, st));
        CU(cudaStreamWaitEvent(ss, c->ev_fork, 0));
    }
    if (c->rec && do_fork) c->rec_lane = 2;
    if (c->use_multirect && nside > 0) {
        const int rc = launch_rects(c, base, recipe, epi, side, nside, ss);
        if (rc == FDW_OK) nside = 0; /* all strips are in flight */
        else if (rc != FDW_ERR_UNSUPPORTED) { c->rec_lane = -1; c->pdl_now = false; return rc; }
    }
    for (int k = 0; k < nside; k++) {
        if (c->rec && do_fork) c->rec_par = 1; /* disjoint rectangles of one level */
        int rc = launch_rect(c, base, recipe, epi, side[k], ss);
        c->rec_par = 0;
        if (rc != FDW_OK) { c->rec_lane = -1; c->pdl_now = false; return rc; }
    }
    c->rec_lane = -1;
    if (ev) CU(cudaEventRecord(c->ev_join, ss));
    const int rcb = launch_rect(c, base, recipe, epi, bulk, st);
    c->pdl_now = false;
    CHECK(rcb);
    if (ev) CU(cudaStreamWaitEvent(st, c->ev_join, 0));
    return FDW_OK;
}

static int launch_step(fdw_ctx *c, StepArgs *a, int recipe, int epi)
{
    return launch_level(c, *a, recipe, epi, a->row0, a->row1, c->stream);
}

static void set_source_args(const fdw_ctx *c, StepArgs *a, int it)
{
    a->src_on = 1;
    a->src_gi = c->sx;
    a->src_j = c->sz;
    a->src_rad = c->src_kind == FDW_SRC_GAUSS7 ? 3 : 0;
    a->src_amp = (it >= 0 && it < (int)c->wavelet.size()) ? c->wavelet[it] : 0.0f;
}

static int sponge_inplace(fdw_ctx *c, Field &n, Field &o, int epi);


/* one propagation step of `pair` with the context's sponge and step ordering.
 * fill(a) lets the caller add epilogue arguments. */
template <class Fill>
static int step_pair(fdw_ctx *c, int pair, int recipe, int epi, bool sponge, bool source, int it, Fill fill)
{
    Field &n = c->f[c->newest[pair]], &o = c->f[c->older[pair]];
    const bool tap = sponge && c->prm.taper != FDW_TAPER_NONE;
    if (tap && c->prm.family == FDW_FAMILY_GPU) { /* sponge first: fd-code.cu:264 */
        n.pend++;
        o.pend++;
    }
    if (n.pend || o.pend) CHECK(sponge_inplace(c, n, o, epi)); /* mid-size whole grids: pending passes applied in place */
    StepArgs a;
    base_args(c, pair, &a);
    a.taper_on = (n.pend || o.pend) ? 1 : 0;
    if (source) set_source_args(c, &a, it);
    fill(a);
    CHECK(launch_step(c, &a, recipe, epi));
    o.pend = 0; /* freshly written level */
    if (tap && c->prm.family == FDW_FAMILY_CPU) { /* sponge last: mod_main.cpp:155-156 */
        o.pend = 1;
        n.pend++;
    }
    int t = c->newest[pair];
    c->newest[pair] = c->older[pair];
    c->older[pair] = t;
    return FDW_OK;
}

/* host copy of the per-level bookkeeping step_pair does, for n levels at once */
static void replay_bookkeeping(fdw_ctx *c, int n, bool tap, int pair = 0)
{
    for (int l = 0; l < n; l++) {
        Field &nw = c->f[c->newest[pair]], &ol = c->f[c->older[pair]];
        if (tap && c->prm.family == FDW_FAMILY_GPU) { nw.pend++; ol.pend++; }
        ol.pend = 0;
        if (tap && c->prm.family == FDW_FAMILY_CPU) { ol.pend = 1; nw.pend++; }
        int t = c->newest[pair];
        c->newest[pair] = c->older[pair];
        c->older[pair] = t;
    }
}

/* ---- shared-memory tile kernels (small grids): tile plan + launch.
 * The grid (nc float4 columns x rows) is cut into at most one tile per SM; among the cuts that fit, the
 * one with the least work per CTA -- own points plus the halo ring it re-reads every level -- wins. */
enum { FDW_TILE_FLAG_WORDS = 32 * 1024 }; /* 32 words (128 B) per tile */
struct TilePlan { int tc4, tr, ntx, nty, ch, threads, sp; size_t smem; };

static bool tile_plan(const fdw_ctx *c, int nc, int rows, int nbuf, TilePlan *best)
{
    long long best_cost = -1;
    for (int ntx = 1; ntx <= nc && ntx <= c->nsm; ntx++) {
        TilePlan t;
        t.tc4 = (nc + ntx - 1) / ntx;
        t.ntx = (nc + t.tc4 - 1) / t.tc4;
        int nty = c->nsm / t.ntx;
        if (nty < 1) break;
        if (nty > rows) nty = rows;
        t.tr = (rows + nty - 1) / nty;
        if (t.tr < GUARD) t.tr = GUARD < rows ? GUARD : rows; /* a tile at least as tall as the halo it feeds */
        t.nty = (rows + t.tr - 1) / t.tr;
        if ((long long)t.ntx * t.nty > c->nsm) continue;
        const int npts = 4 * t.tc4 * t.tr; /* one point per thread and round */
        t.ch = (npts + 1023) / 1024;       /* rounds */
        t.threads = (((npts + t.ch - 1) / t.ch + 31) / 32) * 32;
        if (t.threads > 1024) t.threads = 1024;
        /* row pitch: room for one float4 of halo on each side and = tile width modulo 32, so that a warp whose
         * lanes run on from one tile row into the next still hits 32 different banks */
        t.sp = 4 * t.tc4 + 32;
        t.smem = ((size_t)nbuf * (t.tr + 2 * GUARD) * t.sp                    /* field / velocity tiles */
                  + t.sp + 8 + t.tr + 2 * GUARD + 8) * sizeof(float);         /* sponge tables */
        if (t.smem > (size_t)c->smem_optin) continue;
        const long long cost = (long long)t.tc4 * t.tr + 2LL * GUARD * t.tc4 + 2LL * t.tr;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; *best = t; }
    }
    return best_cost >= 0;
}

static int tile_launch(fdw_ctx *c, const void *k, TileArgs &ta, const TilePlan &tp)
{
    ta.tc4 = tp.tc4; ta.tr = tp.tr; ta.ntx = tp.ntx; ta.nty = tp.nty; ta.ch = tp.ch; ta.sp = tp.sp;
    if (c->use_ll) {
        const size_t plane = c->field_elems, bytes = 4 * plane * sizeof(unsigned long long);
        if (!c->ll_d && cudaMalloc(&c->ll_d, bytes) != cudaSuccess) { (void)cudaGetLastError(); c->ll_d = nullptr; }
        if (c->ll_d) {
            if (cudaMemsetAsync(c->ll_d, 0, bytes, c->stream) != cudaSuccess) return 0;
            ta.ll = c->ll_d + (size_t)(AGUARD + 1) * c->pitch; /* local row 0, column 0 of plane 0 */
            ta.ll_plane = (long long)plane;
        }
    }
    if (const char *e = getenv("FDW_TILE_DBG")) ta.dbg = atoi(e);
    if (getenv("FDW_TILE_VERBOSE"))
        fprintf(stderr, "fdwave tile plan: %d x %d tiles of %d float4 columns x %d rows, %d point(s)/thread, %d threads, %zu B smem\n",
                tp.ntx, tp.nty, tp.tc4, tp.tr, tp.ch, tp.threads, tp.smem);
    ta.pa.base.apitch = c->pitch;
    ta.pa.base.pitch = tp.sp;
    if (tp.smem > 48 * 1024 &&
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tp.smem) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    if (tp.ntx * tp.nty * 32 > FDW_TILE_FLAG_WORDS) return 0;
    if (cudaMemsetAsync(c->barrier_d, 0, sizeof(unsigned), c->stream) != cudaSuccess) return 0;
    if (cudaMemsetAsync(c->tileflags_d, 0, (size_t)tp.ntx * tp.nty * 32 * sizeof(unsigned), c->stream) != cudaSuccess) return 0;
    ta.flags = c->tileflags_d;
    void *params[] = {&ta};
    cudaError_t e = cudaLaunchCooperativeKernel(k, dim3(tp.ntx * tp.nty), dim3(tp.threads), params, tp.smem, c->stream);
    if (e != cudaSuccess) {
        (void)cudaGetLastError(); /* not co-resident: the caller falls back */
        return 0;
    }
    c->launches++;
    c->tile_launches++;
    c->persist_launches++; /* same device-side error flag protocol */
    return 1;
}

/* fill the PersistArgs part shared by the persistent and the tile kernels */
template <class Fill>
static void persist_args(fdw_ctx *c, int pair, StepArgs &base, bool tap, bool source, int it0, int n, int tidx_cpu,
                         Fill fill, PersistArgs *pa)
{
    memset(pa, 0, sizeof *pa);
    base.taper_on = 1; /* the sponge instantiation serves every CTA; counts of 0 make it a no-op */
    if (source) { set_source_args(c, &base, it0); base.src_on = 1; }
    fill(base);
    pa->base = base;
    pa->bufN = c->f[c->newest[pair]].r0; pa->bufO = c->f[c->older[pair]].r0;
    pa->pendN = c->f[c->newest[pair]].pend; pa->pendO = c->f[c->older[pair]].pend;
    pa->nlevels = n; pa->it0 = it0; pa->nt = c->prm.nt;
    pa->sponge = tap ? 1 : 0;
    pa->sponge_first = c->prm.family == FDW_FAMILY_GPU ? 1 : 0;
    pa->source = source ? 1 : 0;
    pa->tidx_cpu = tidx_cpu;
    pa->wavelet = c->wavelet_d;
    pa->hist = c->hist;
    pa->hist_slice = (long long)(c->nli > 0 ? c->nli : 1) * c->pitch;
    pa->barrier = c->barrier_d;
    pa->error_flag = c->errflag_d;
    pa->timeout_ns = c->timeout_ns;
}

static bool small_whole_grid(const fdw_ctx *c, const StepArgs &base)
{
    const int rows = base.row1 - base.row0, nc = base.ncol4;
    return c->coop && c->gx0 == 0 && c->nloc == c->nxe && !c->step_open && rows > 0 && nc > 0 &&
           (long long)nc * rows < c->persist_limit;
}

/* n consecutive levels of pair 0 by the tile kernel; 1 = done, 0 = fall back */
template <class Fill>
static int try_tile(fdw_ctx *c, int recipe, int epi, bool sponge, bool source, int it0, int n, int tidx_cpu, Fill fill)
{
    if (!c->use_tile || n < 2) return 0;
    const void *k = tile_kernel(c->prm.order, recipe, epi);
    if (!k) return 0;
    StepArgs base;
    base_args(c, 0, &base);
    if (!small_whole_grid(c, base)) return 0;
    if (source && (it0 < 0 || (size_t)(it0 + n) > c->wavelet.size() || !c->wavelet_d)) return 0;
    TilePlan tp;
    if (!tile_plan(c, base.ncol4, base.row1 - base.row0, 3, &tp)) return 0;
    const bool tap = sponge && c->prm.taper != FDW_TAPER_NONE;
    TileArgs ta;
    memset(&ta, 0, sizeof ta);
    persist_args(c, 0, base, tap, source, it0, n, tidx_cpu, fill, &ta.pa);
    if (!tile_launch(c, k, ta, tp)) return 0;
    replay_bookkeeping(c, n, tap);
    return 1;
}

/* Try to run n consecutive levels of pair 0 in ONE cooperative launch (small, launch-bound
 * grids).  Returns 1 when it did, 0 when the caller must fall back to per-level launches. */
template <class Fill>
static int try_persistent(fdw_ctx *c, int recipe, int epi, bool sponge, bool source, int it0, int n, int tidx_cpu,
                          Fill fill, int *rc)
{
    *rc = FDW_OK;
    if (!c->coop || n < 2 || c->gx0 != 0 || c->nloc != c->nxe || c->step_open) return 0;
    const void *k = persist_kernel(c->prm.order, recipe, epi);
    if (!k) return 0;
    StepArgs base;
    base_args(c, 0, &base);
    const int rows = base.row1 - base.row0, nc = base.ncol4;
    if (rows <= 0 || nc <= 0 || (long long)nc * rows >= c->persist_limit) return 0;
    if (source && (it0 < 0 || (size_t)(it0 + n) > c->wavelet.size() || !c->wavelet_d)) return 0;
    int nthreads = nc >= 256 ? 256 : ((nc + 31) / 32) * 32;
    if (c->threads_override > 0) nthreads = c->threads_override;
    const int gx = (nc + nthreads - 1) / nthreads;
    const long long cap = (long long)c->nsm * cached_occupancy(k, nthreads);
    int rpc = 2;
    while ((long long)gx * ((rows + rpc - 1) / rpc) > cap && rpc < rows) rpc *= 2;
    if ((long long)gx * ((rows + rpc - 1) / rpc) > cap) return 0;
    const bool tap = sponge && c->prm.taper != FDW_TAPER_NONE;
    PersistArgs pa;
    memset(&pa, 0, sizeof pa);
    base.taper_on = 1; /* the sponge instantiation serves every CTA; counts of 0 make it a no-op */
    base.rows_per_cta = rpc;
    if (source) { set_source_args(c, &base, it0); base.src_on = 1; }
    fill(base);
    pa.base = base;
    pa.bufN = c->f[c->newest[0]].r0; pa.bufO = c->f[c->older[0]].r0;
    pa.pendN = c->f[c->newest[0]].pend; pa.pendO = c->f[c->older[0]].pend;
    pa.nlevels = n; pa.it0 = it0; pa.nt = c->prm.nt;
    pa.sponge = tap ? 1 : 0;
    pa.sponge_first = c->prm.family == FDW_FAMILY_GPU ? 1 : 0;
    pa.source = source ? 1 : 0;
    pa.tidx_cpu = tidx_cpu;
    pa.wavelet = c->wavelet_d;
    pa.hist = c->hist;
    pa.hist_slice = (long long)(c->nli > 0 ? c->nli : 1) * c->pitch;
    pa.barrier = c->barrier_d;
    pa.error_flag = c->errflag_d;
    pa.timeout_ns = c->timeout_ns;
    if (cudaMemsetAsync(c->barrier_d, 0, sizeof(unsigned), c->stream) != cudaSuccess) return 0;
    dim3 grid(gx, (rows + rpc - 1) / rpc, 1), block(nthreads, 1, 1);
    void *params[] = {&pa};
#ifdef FDW_EMU
    emu_levels = n;
#endif
    cudaError_t e = cudaLaunchCooperativeKernel(k, grid, block, params, 0, c->stream);
    if (e != cudaSuccess) {
        (void)cudaGetLastError(); /* e.g. grid too large to be co-resident: fall back */
        return 0;
    }
    c->launches++;
    c->persist_launches++;
    replay_bookkeeping(c, n, tap);
    return 1;
}

static int graph_submit(fdw_ctx *c, std::vector<RecLaunch> &rec);

/* n levels of pair 0: one persistent launch when the grid is small, a replayed graph of two levels when it is
 * mid-size, else one launch (plus strips) per level */
template <class FillStatic, class FillLevel>
static int run_levels(fdw_ctx *c, int recipe, int epi, bool sponge, bool source, int it0, int n, int tidx_cpu,
                      FillStatic fill_static, FillLevel fill_level)
{
    int rc;
    if (try_tile(c, recipe, epi, sponge, source, it0, n, tidx_cpu, fill_static)) return FDW_OK;
    if (try_persistent(c, recipe, epi, sponge, source, it0, n, tidx_cpu, fill_static, &rc)) return rc;
    int it = it0;
    if (c->use_graph && c->level_graph && !c->step_open && !c->rec && n >= 4 &&
        (long long)c->ncol4 * c->nloc < c->level_graph_limit) {
        for (; it + 1 < it0 + n; it += 2) {
            std::vector<RecLaunch> rec;
            c->rec = &rec;
            rc = FDW_OK;
            for (int l = 0; l < 2 && rc == FDW_OK; l++) {
                c->rec_level = l;
                rc = step_pair(c, 0, recipe, epi, sponge, source, it + l, [&](StepArgs &a) { fill_level(a, it + l); });
            }
            c->rec = nullptr;
            CHECK(rc);
            CHECK(graph_submit(c, rec));
        }
    }
    for (; it < it0 + n; it++)
        CHECK(step_pair(c, 0, recipe, epi, sponge, source, it, [&](StepArgs &a) { fill_level(a, it); }));
    return FDW_OK;
}

static int check_device_flag(fdw_ctx *c)
{
    if (c->persist_launches == 0 && c->peer_waits == 0) return FDW_OK;
    int flag = 0;
    CU(cudaMemcpyAsync(&flag, c->errflag_d, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (flag == 2) { fdw_set_error("peer halo exchange: a neighbour's boundary rows never arrived"); return FDW_ERR_CUDA; }
    if (flag) { fdw_set_error("persistent kernel: grid barrier timed out"); return FDW_ERR_CUDA; }
    return FDW_OK;
}

static int materialize(fdw_ctx *c, Field &f)
{
    if (f.pend == 0) return FDW_OK;
    int lo = -GUARD, hi = c->nloc + GUARD;
    if (c->gx0 + lo < 0) lo = -c->gx0;
    if (c->gx0 + hi > c->nxe) hi = c->nxe - c->gx0;
    float *r0 = f.r0;
    long long pitch = c->pitch;
    const float *tz = c->tz, *tx = c->tx;
    int cnt = f.pend;
    for (int b0 = lo; b0 < hi; b0 += 65535) { /* gridDim.y limit: bands of at most 65535 rows */
        int b1 = b0 + 65535 < hi ? b0 + 65535 : hi;
        dim3 block(128), grid((c->nze + 127) / 128, b1 - b0);
        void *params[] = {&r0, &pitch, &c->nze, &b0, &b1, &c->gx0, &tz, &tx, &c->tx_jlim, &c->tz_ilim, &cnt};
        CU(cudaLaunchKernel(FDW_KPTR(k_materialize, thunk_materialize), grid, block, params, 0, c->stream));
        c->launches++;
    }
    f.pend = 0;
    return FDW_OK;
}

/* Whole-grid levels above the small-grid range: the sponge-on-load strips cost more than they hide -- 6-10 % of a
 * mid-size grid's points on the 110-register instantiation, resident from the start of the level, take the bulk
 * launch's occupancy away (measured on the 8272 x 2128 grid: four-sided sponge +15 us on a 61 us level even as one
 * folded multi-rectangle launch).  Instead the pending sponge passes are applied IN PLACE to the sponge regions of
 * both levels first (one launch of an element-wise kernel over a few per cent of the grid: ~12 MB of traffic against
 * 280 MB for that level) and ONE plain launch then covers the whole grid; the two follow each other in one stream,
 * so programmatic dependent launch applies (launch_pdl).  Same multiplications in the same order as on load (and
 * as the reference's own in-place passes), so the bits do not change.  Measured: mod_main level 78 -> 72 (-> 68 with
 * PDL) us; 16384^2 with a top sponge, where the one strip hides behind a 0.68 ms bulk launch: still +0.75 %
 * (profiles/r02s_, r02u_, r02v_).  Not used where a level goes out in pieces (slabs), on small grids (one
 * sponge-kernel launch), nor when a recorded seismogram sample would need a non-unit factor of the pass that is
 * still to come (receivers inside the sponge). */
static int sponge_inplace(fdw_ctx *c, Field &n, Field &o, int epi)
{
    const long long work = (long long)c->ncol4 * c->nloc;
    if (!c->inplace_sponge || c->rec || c->step_open || c->gx0 != 0 || c->nloc != c->nxe ||
        work < c->small_grid_limit || work >= c->inplace_limit)
        return FDW_OK;
    if ((epi & fdw::EPI_RECORD) && !(c->shot_gz >= c->tap_jlo && c->shot_gz < c->tap_jhi)) return FDW_OK;
    if (epi & fdw::EPI_PUSH) return FDW_OK;
    SpongeRects a;
    memset(&a, 0, sizeof a);
    a.f[0] = n.pend ? n.r0 : nullptr; a.cnt[0] = n.pend;
    a.f[1] = o.pend ? o.r0 : nullptr; a.cnt[1] = o.pend;
    a.pitch = c->pitch; a.grow0 = c->gx0; a.tx_jlim = c->tx_jlim; a.tz_ilim = c->tz_ilim;
    a.tz = c->tz; a.tx = c->tx;
    const int jlo = c->tap_jlo > INT_MIN ? (c->tap_jlo < c->nze ? c->tap_jlo : c->nze) : 0;
    const int jhi = c->tap_jhi < INT_MAX ? (c->tap_jhi > jlo ? c->tap_jhi : jlo) : c->nze;
    const int ilo = c->tap_ilo > INT_MIN ? (c->tap_ilo < c->nloc ? c->tap_ilo : c->nloc) : 0;
    const int ihi = c->tap_ihi < INT_MAX ? (c->tap_ihi > ilo ? c->tap_ihi : ilo) : c->nloc;
    const int rects[4][4] = {{0, jlo, 0, c->nloc}, {jhi, c->nze, 0, c->nloc}, {jlo, jhi, 0, ilo}, {jlo, jhi, ihi, c->nloc}};
    long long total = 0;
    for (int k = 0; k < 4; k++) {
        if (rects[k][1] <= rects[k][0] || rects[k][3] <= rects[k][2]) continue;
        a.r[a.n].c0 = rects[k][0]; a.r[a.n].c1 = rects[k][1]; a.r[a.n].r0 = rects[k][2]; a.r[a.n].r1 = rects[k][3];
        a.r[a.n].e0 = (unsigned)total;
        total += (long long)(rects[k][1] - rects[k][0]) * (rects[k][3] - rects[k][2]);
        a.n++;
    }
    if (total >= 0x7fffffffLL) return FDW_OK; /* (the strips do it) */
    a.total = (unsigned)total;
    if (total > 0) {
        const int threads = 256;
        const long long blocks = (total + (long long)threads * SPONGE_INPLACE_PER_THREAD - 1) / ((long long)threads * SPONGE_INPLACE_PER_THREAD);
        void *params[] = {&a};
        c->pdl_now = true; /* whole-grid level in one stream: in-place pass, plain launch, in-place pass, ... */
        const cudaError_t le = launch_pdl(c, FDW_KPTR(k_sponge_inplace, thunk_sponge_inplace), dim3((unsigned)blocks),
                                          dim3(threads), params, c->stream);
        c->pdl_now = false;
        CU(le);
        c->launches++;
    }
    n.pend = 0;
    o.pend = 0;
    return FDW_OK;
}

static int field_alloc(fdw_ctx *c, Field *f)
{
    CU(cudaMalloc(&f->base, c->field_elems * sizeof(float)));
    CU(cudaMemsetAsync(f->base, 0, c->field_elems * sizeof(float), c->stream));
    f->r0 = f->base + (size_t)(AGUARD + 1) * c->pitch;
    f->pend = 0;
    return FDW_OK;
}

static int field_zero(fdw_ctx *c, Field *f)
{
    /* With a neighbour attached, the ghost rows on that side belong to the neighbour's fdw_peer_refresh, which
     * may land before or after this memset: they are left alone (the refresh that must follow a zero on every
     * slab overwrites them), only the owned rows and the ghost rows of a physical grid edge are cleared. */
    const size_t ghost = (size_t)(AGUARD + 1) * c->pitch;
    float *lo = f->base + (c->peer[0].on ? ghost : 0);
    float *hi = f->base + c->field_elems - (c->peer[1].on ? ghost : 0);
    CU(cudaMemsetAsync(lo, 0, (size_t)(hi - lo) * sizeof(float), c->stream));
    f->pend = 0;
    return FDW_OK;
}

/* host [nxe][nze] rows gx0.. -> device local rows (optionally with ghost rows) */
static int field_h2d(fdw_ctx *c, float *r0, const float *host, int ghost)
{
    int lo = ghost ? -GUARD : 0, hi = c->nloc + (ghost ? GUARD : 0);
    if (c->gx0 + lo < 0) lo = -c->gx0;
    if (c->gx0 + hi > c->nxe) hi = c->nxe - c->gx0;
    CU(cudaMemcpy2DAsync(r0 + (long long)lo * c->pitch, c->pitch * sizeof(float),
                         host + (size_t)(c->gx0 + lo) * c->nze, (size_t)c->nze * sizeof(float),
                         (size_t)c->nze * sizeof(float), hi - lo, cudaMemcpyHostToDevice, c->stream));
    return FDW_OK;
}

static int field_d2h(fdw_ctx *c, const float *r0, float *host)
{
    CU(cudaMemcpy2DAsync(host + (size_t)c->gx0 * c->nze, (size_t)c->nze * sizeof(float), r0,
                         c->pitch * sizeof(float), (size_t)c->nze * sizeof(float), c->nloc,
                         cudaMemcpyDeviceToHost, c->stream));
    return FDW_OK;
}

/* ------------------------------------------------------------------ create / destroy */
extern "C" int fdw_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static void build_sponge_tables(fdw_ctx *c, std::vector<float> &tz, std::vector<float> &tx)
{
    const fdw_params &p = c->prm;
    std::vector<float> tabx(p.nxb > 0 ? p.nxb : 1), tabz(p.nzb > 0 ? p.nzb : 1);
    fdw_taper_table(p.nxb, p.fac, p.family, tabx.data());
    fdw_taper_table(p.nzb, p.fac, p.family, tabz.data());
    tz.assign((size_t)c->pitch + 2 * AGUARD, 1.0f); /* index AGUARD + j */
    tx.assign((size_t)c->nloc + 2 * AGUARD, 1.0f); /* index AGUARD + local row */
    c->tx_jlim = INT_MAX; c->tz_ilim = INT_MAX;
    c->tap_jlo = INT_MIN; c->tap_jhi = INT_MAX; c->tap_ilo = INT_MIN; c->tap_ihi = INT_MAX;
    if (p.taper == FDW_TAPER_NONE) return;
    int nzb_eff = p.nzb;
    if (p.taper == FDW_TAPER_TOP && p.compat_extents) {
        nzb_eff = (p.nzb / 8) * 8;            /* gridBorder_z, fd-code.cu:192-195 */
        c->tz_ilim = (c->nxe / 8) * 8;        /* z factor only on launched rows */
    }
    for (int j = 0; j < nzb_eff; j++) tz[AGUARD + j] = tabz[j];
    if (p.taper == FDW_TAPER_FOUR)
        for (int j = p.nz + p.nzb; j < c->nze; j++) tz[AGUARD + j] = tabz[c->nze - 1 - j];
    for (int lr = -AGUARD; lr < c->nloc + AGUARD; lr++) {
        int gi = c->gx0 + lr;
        if (gi < 0 || gi >= c->nxe) continue;
        if (gi < p.nxb) tx[AGUARD + lr] = tabx[gi];
        else if (gi >= c->nxe - p.nxb) tx[AGUARD + lr] = tabx[c->nxe - 1 - gi];
    }
    if (p.taper == FDW_TAPER_TOP) {
        c->tx_jlim = nzb_eff; /* x factor only on the two top corners */
        c->tap_jlo = nzb_eff;
    } else {
        c->tap_jlo = p.nzb; c->tap_jhi = p.nz + p.nzb;
        c->tap_ilo = p.nxb; c->tap_ihi = p.nx + p.nxb;
    }
}

extern "C" int fdw_create(const fdw_params *prm, fdw_ctx **out)
{
    if (!prm || !out) { fdw_set_error("fdw_create: null argument"); return FDW_ERR_ARG; }
    *out = nullptr;
    if (prm->nx < 1 || prm->nz < 1 || prm->nxb < 0 || prm->nzb < 0) {
        fdw_set_error("fdw_create: bad grid %dx%d border %d,%d", prm->nx, prm->nz, prm->nxb, prm->nzb);
        return FDW_ERR_ARG;
    }
    if (prm->order < 2 || prm->order > fdw::MAX_ORDER || (prm->order & 1)) {
        fdw_set_error("fdw_create: order %d not supported on the device (even, 2..%d)", prm->order, (int)fdw::MAX_ORDER);
        return FDW_ERR_UNSUPPORTED;
    }
    if (prm->order > 2 * GUARD && prm->slab_x1 > prm->slab_x0 &&
        (prm->slab_x0 > 0 || prm->slab_x1 < prm->nx + 2 * prm->nxb)) {
        fdw_set_error("fdw_create: slab decomposition exchanges %d ghost rows: order %d needs the whole grid on one GPU",
                      (int)GUARD, prm->order);
        return FDW_ERR_UNSUPPORTED;
    }
    int ndev = fdw_device_count();
    if (ndev < 1 || prm->device < 0 || prm->device >= ndev) {
        fdw_set_error("fdw_create: CUDA device %d not available (%d visible); libfdwave has no CPU path",
                      prm->device, ndev);
        return FDW_ERR_CUDA;
    }
    fdw_ctx *c = new fdw_ctx();
    c->prm = *prm;
    c->nxe = prm->nx + 2 * prm->nxb;
    c->nze = prm->nz + 2 * prm->nzb;
    c->H = prm->order / 2;
    c->gx0 = 0;
    c->nloc = c->nxe;
    if (prm->slab_x1 > prm->slab_x0) {
        if (prm->slab_x0 < 0 || prm->slab_x1 > c->nxe) {
            fdw_set_error("fdw_create: slab [%d,%d) outside [0,%d)", prm->slab_x0, prm->slab_x1, c->nxe);
            delete c;
            return FDW_ERR_ARG;
        }
        c->gx0 = prm->slab_x0;
        c->nloc = prm->slab_x1 - prm->slab_x0;
    }
    c->li0 = prm->nxb > c->gx0 ? prm->nxb : c->gx0;
    {
        int hi = prm->nxb + prm->nx < c->gx0 + c->nloc ? prm->nxb + prm->nx : c->gx0 + c->nloc;
        c->nli = hi > c->li0 ? hi - c->li0 : 0;
    }
    c->pitch = pitch_for(c->nze, prm->order);
    c->rows_alloc = (size_t)c->nloc + 2 * AGUARD + 2;
    c->field_elems = c->rows_alloc * (size_t)c->pitch;
    int rc = bind(c);
    if (rc != FDW_OK) { delete c; return rc; }
    cudaDeviceProp dp;
    if (cudaGetDeviceProperties(&dp, prm->device) != cudaSuccess) { delete c; fdw_set_error("cudaGetDeviceProperties failed"); return FDW_ERR_CUDA; }
    c->nsm = dp.multiProcessorCount;
    if (const char *e = getenv("FDW_ROWS_PER_CTA")) c->rows_per_cta_override = atoi(e);
    if (const char *e = getenv("FDW_THREADS")) c->threads_override = atoi(e);
    if (const char *e = getenv("FDW_SMALL_GRID_LIMIT")) c->small_grid_limit = atoll(e);
    if (const char *e = getenv("FDW_PERSIST_LIMIT")) c->persist_limit = atoll(e);
    cudaDeviceGetAttribute(&c->coop, cudaDevAttrCooperativeLaunch, prm->device);
    if (const char *e = getenv("FDW_TILE")) c->use_tile = atoi(e);
    if (const char *e = getenv("FDW_TILE_LL")) c->use_ll = atoi(e);
    if (const char *e = getenv("FDW_PSLAB")) c->use_pslab = atoi(e);
    if (const char *e = getenv("FDW_PSLAB_LIMIT")) c->pslab_limit = atoll(e);
    if (const char *e = getenv("FDW_PSLAB_THREADS")) {
        const int t = atoi(e);
        if (t >= 32 && t <= 128 && t % 32 == 0) c->pslab_threads = t;
    }
    cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, prm->device);
    if (const char *e = getenv("FDW_FORK_LIMIT")) c->fork_limit = atoll(e);
    if (const char *e = getenv("FDW_RPC_RULE")) g_rpc_rule = atoi(e);
    if (const char *e = getenv("FDW_MULTIRECT")) c->use_multirect = atoi(e);
    if (const char *e = getenv("FDW_PDL")) c->use_pdl = atoi(e);
    if (const char *e = getenv("FDW_SPONGE_INPLACE")) c->inplace_sponge = atoi(e);
    if (const char *e = getenv("FDW_SPONGE_INPLACE_LIMIT")) c->inplace_limit = atoll(e);
    if (const char *e = getenv("FDW_LEVEL_GRAPH")) c->level_graph = atoi(e);
    if (const char *e = getenv("FDW_LEVEL_GRAPH_LIMIT")) c->level_graph_limit = atoll(e);
    if (const char *e = getenv("FDW_GRAPH")) c->use_graph = atoi(e);
    if (const char *e = getenv("FDW_GRAPH_LIMIT")) c->graph_limit = atoll(e);
    if (const char *e = getenv("FDW_GRAPH_LEVELS")) { c->graph_levels = atoi(e) & ~1; if (c->graph_levels < 2) c->graph_levels = 2; }
    if (const char *e = getenv("FDW_FUSE_FLAGS")) c->fuse_flags = atoi(e);
    if (const char *e = getenv("FDW_TIMEOUT_MS")) c->timeout_ns = 1000000ull * (unsigned long long)atoll(e);

    /* coefficients and scalars: fd-code.cu:203-217 / fd.c:12-16 */
    float coefs[fdw::MAX_ORDER + 1];
    fdw_calc_coefs(prm->order, prm->family, coefs);
    c->dx2inv = (1. / prm->dx) * (1. / prm->dx);
    c->dz2inv = (1. / prm->dz) * (1. / prm->dz);
    c->dt2 = prm->dt * prm->dt;
    memset(c->cz, 0, sizeof c->cz);
    memset(c->cx, 0, sizeof c->cx);
    for (int k = 0; k <= prm->order; k++) {
        if (prm->recipe == FDW_RECIPE_C) {
            c->cz[k] = c->cx[k] = coefs[k];
        } else {
            c->cz[k] = c->dz2inv * coefs[k];
            c->cx[k] = c->dx2inv * coefs[k];
        }
    }
    fdw_ptsrc_weights(c->src_w);

    /* extents (quirk Q1 when asked for) */
    const int h = c->H;
    if (prm->compat_extents) {
        const int ux = (c->nxe / 8) * 8, uz = (c->nze / 8) * 8;
        c->upd_i1 = ux; c->upd_j1 = uz;
        c->lap_i1 = h + ux < c->nxe - h ? h + ux : c->nxe - h;
        c->lap_j1 = h + uz < c->nze - h ? h + uz : c->nze - h;
    } else {
        c->upd_i1 = c->nxe; c->upd_j1 = c->nze;
        c->lap_i1 = c->nxe - h; c->lap_j1 = c->nze - h;
    }
    c->lap_i0 = c->lap_j0 = h;
    c->ncol4 = (c->upd_j1 + 3) / 4;

#define TRY(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            fdw_set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));  \
            fdw_destroy(c);                                                                    \
            return e_ == cudaErrorMemoryAllocation ? FDW_ERR_NOMEM : FDW_ERR_CUDA;             \
        }                                                                                      \
    } while (0)
    TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    TRY(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    TRY(cudaEventCreate(&c->ev0));
    TRY(cudaEventCreate(&c->ev1));
    TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&c->ev_pfork, cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&c->ev_pjoin, cudaEventDisableTiming));
    for (int k = 0; k < 4; k++) {
        rc = field_alloc(c, &c->f[k]);
        if (rc != FDW_OK) { fdw_destroy(c); return rc; }
    }
    c->newest[0] = 0; c->older[0] = 1; c->newest[1] = 2; c->older[1] = 3;
    TRY(cudaMalloc(&c->vdt_base, c->field_elems * sizeof(float)));
    TRY(cudaMemsetAsync(c->vdt_base, 0, c->field_elems * sizeof(float), c->stream));
    c->vdt = c->vdt_base + (size_t)(AGUARD + 1) * c->pitch;
    std::vector<float> tz, tx;
    build_sponge_tables(c, tz, tx);
    TRY(cudaMalloc(&c->tz_base, tz.size() * sizeof(float)));
    TRY(cudaMalloc(&c->tx_base, tx.size() * sizeof(float)));
    TRY(cudaMemcpyAsync(c->tz_base, tz.data(), tz.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    TRY(cudaMemcpyAsync(c->tx_base, tx.data(), tx.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    c->tz = c->tz_base + AGUARD;
    c->tx = c->tx_base + AGUARD;
    const size_t img_elems = (size_t)(c->nli > 0 ? c->nli : 1) * c->pitch; /* image/history rows = owned interior rows */
    TRY(cudaMalloc(&c->img, img_elems * sizeof(float)));
    TRY(cudaMemsetAsync(c->img, 0, img_elems * sizeof(float), c->stream));
    if (prm->history) {
        if (prm->nt < 1) { fdw_destroy(c); fdw_set_error("fdw_create: history needs nt"); return FDW_ERR_ARG; }
        TRY(cudaMalloc(&c->hist, (size_t)prm->nt * img_elems * sizeof(float)));
    }
    TRY(cudaMalloc(&c->barrier_d, sizeof(unsigned)));
    TRY(cudaMalloc(&c->tileflags_d, FDW_TILE_FLAG_WORDS * sizeof(unsigned)));
    TRY(cudaMalloc(&c->flags_d, 2 * sizeof(unsigned)));
    TRY(cudaMemsetAsync(c->flags_d, 0, 2 * sizeof(unsigned), c->stream));
    TRY(cudaMalloc(&c->pcount_d, sizeof(unsigned)));
    TRY(cudaMemsetAsync(c->pcount_d, 0, sizeof(unsigned), c->stream));
    TRY(cudaMalloc(&c->errflag_d, sizeof(int)));
    TRY(cudaMemsetAsync(c->errflag_d, 0, sizeof(int), c->stream));
    TRY(cudaStreamSynchronize(c->stream)); /* the staging vectors go out of scope */
#undef TRY
    *out = c;
    return FDW_OK;
}

extern "C" void fdw_destroy(fdw_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->prm.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    fdw_peer_detach(c);
    if (c->lexec) cudaGraphExecDestroy(c->lexec);
    if (c->lgraph) cudaGraphDestroy(c->lgraph);
    cudaFree(c->flags_d);
    cudaFree(c->pcount_d);
    for (int k = 0; k < 4; k++) cudaFree(c->f[k].base);
    cudaFree(c->vdt_base); cudaFree(c->tz_base); cudaFree(c->tx_base);
    cudaFree(c->vdt2_base); cudaFree(c->stack);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_staged) cudaEventDestroy(c->ev_staged);
    if (c->ev_shot_done) cudaEventDestroy(c->ev_shot_done);
    cudaFree(c->hist); cudaFree(c->img); cudaFree(c->dobs_d); cudaFree(c->rec_d);
    cudaFree(c->wavelet_d); cudaFree(c->barrier_d); cudaFree(c->errflag_d); cudaFree(c->tileflags_d); cudaFree(c->ll_d);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_pfork) cudaEventDestroy(c->ev_pfork);
    if (c->ev_pjoin) cudaEventDestroy(c->ev_pjoin);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int fdw_set_stream(fdw_ctx *c, void *s)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    CU(cudaStreamSynchronize(c->stream));
    if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
    if (s) {
        c->stream = (cudaStream_t)s;
    } else {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return FDW_OK;
}

extern "C" int fdw_sync(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    CU(cudaStreamSynchronize(c->stream));
    return check_device_flag(c);
}

extern "C" int fdw_set_v2(fdw_ctx *c, const float *v2)
{
    if (!c || !v2) return FDW_ERR_ARG;
    CHECK(bind(c));
    CHECK(field_h2d(c, c->vdt, v2, 1));
    /* fl32(v2*dt2): the first of the two float multiplies of v2*dt2*lap (fd-code.cu:89) */
    long long n = (long long)c->field_elems;
    void *params[] = {&c->vdt_base, &n, &c->dt2};
    CU(cudaLaunchKernel(FDW_KPTR(k_scale_rows, thunk_scale_rows), dim3((unsigned)((n + 255) / 256)), dim3(256), params,
                        0, c->stream));
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
    return FDW_OK;
}

extern "C" int fdw_set_wavelet(fdw_ctx *c, const float *s, int nt)
{
    if (!c || !s || nt < 0) return FDW_ERR_ARG;
    c->wavelet.assign(s, s + nt);
    CHECK(bind(c));
    if (c->wavelet_d) { cudaFree(c->wavelet_d); c->wavelet_d = nullptr; }
    if (nt > 0) {
        CU(cudaMalloc(&c->wavelet_d, (size_t)nt * sizeof(float)));
        CU(cudaMemcpyAsync(c->wavelet_d, c->wavelet.data(), (size_t)nt * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    return FDW_OK;
}

extern "C" int fdw_set_source(fdw_ctx *c, int sx, int sz, int kind)
{
    if (!c || (kind != FDW_SRC_POINT && kind != FDW_SRC_GAUSS7)) return FDW_ERR_ARG;
    c->sx = sx; c->sz = sz; c->src_kind = kind;
    return FDW_OK;
}

extern "C" int fdw_fields_zero(fdw_ctx *c, int pair)
{
    if (!c || pair < 0 || pair > 1) return FDW_ERR_ARG;
    CHECK(bind(c));
    CHECK(field_zero(c, &c->f[c->newest[pair]]));
    CHECK(field_zero(c, &c->f[c->older[pair]]));
    if (pair == 0) c->saved_valid = false;
    return FDW_OK;
}

extern "C" int fdw_fields_upload(fdw_ctx *c, int pair, const float *newest, const float *older)
{
    if (!c || pair < 0 || pair > 1 || !newest || !older) return FDW_ERR_ARG;
    CHECK(bind(c));
    Field &n = c->f[c->newest[pair]], &o = c->f[c->older[pair]];
    CHECK(field_h2d(c, n.r0, newest, 1));
    CHECK(field_h2d(c, o.r0, older, 1));
    n.pend = o.pend = 0;
    return FDW_OK;
}

extern "C" int fdw_fields_download(fdw_ctx *c, int pair, float *newest, float *older)
{
    if (!c || pair < 0 || pair > 1) return FDW_ERR_ARG;
    CHECK(bind(c));
    Field &n = c->f[c->newest[pair]], &o = c->f[c->older[pair]];
    if (newest) { CHECK(materialize(c, n)); CHECK(field_d2h(c, n.r0, newest)); }
    if (older) { CHECK(materialize(c, o)); CHECK(field_d2h(c, o.r0, older)); }
    CU(cudaStreamSynchronize(c->stream));
    return check_device_flag(c);
}

extern "C" int fdw_advance(fdw_ctx *c, int it0, int nsteps)
{
    if (!c || nsteps < 0) return FDW_ERR_ARG;
    CHECK(bind(c));
    return run_levels(c, c->prm.recipe, 0, true, !c->wavelet.empty(), it0, nsteps, 0, [](StepArgs &) {},
                      [](StepArgs &, int) {});
}

extern "C" int fdw_propagate(fdw_ctx *c, float *newest, float *older, int it0, int nsteps)
{
    if (!c || !newest || !older) return FDW_ERR_ARG;
    CHECK(fdw_fields_upload(c, 0, newest, older));
    CHECK(fdw_advance(c, it0, nsteps));
    return fdw_fields_download(c, 0, newest, older);
}

/* ------------------------------------------------------------------ pipelines */
extern "C" int fdw_forward(fdw_ctx *c, int sx, int sz, float *P, float *PP)
{
    if (!c) return FDW_ERR_ARG;
    if (c->wavelet.size() < (size_t)c->prm.nt || c->prm.nt < 1) {
        fdw_set_error("fdw_forward: set params.nt and a wavelet of nt samples first");
        return FDW_ERR_STATE;
    }
    CHECK(fdw_fields_zero(c, 0));
    CHECK(fdw_set_source(c, sx, sz, FDW_SRC_POINT));
    CHECK(fdw_advance(c, 0, c->prm.nt));
    /* d_p (older, sponged once in the last iteration) and d_pp (newest), fd-code.cu:285-286 */
    CHECK(materialize(c, c->f[c->newest[0]]));
    CHECK(materialize(c, c->f[c->older[0]]));
    c->saved_valid = true;
    if (P || PP) CHECK(fdw_fields_download(c, 0, PP, P));
    return FDW_OK;
}

static int ensure_buffer(float **buf, size_t *cap, size_t need)
{
    if (*cap >= need) return FDW_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    CU(cudaMalloc(buf, need * sizeof(float)));
    *cap = need;
    return FDW_OK;
}

static int check_device_flag(fdw_ctx *c);

static int image_download(fdw_ctx *c, float *imloc)
{
    const fdw_params &p = c->prm;
    if (c->nli > 0)
        CU(cudaMemcpy2DAsync(imloc, (size_t)p.nz * sizeof(float), c->img + p.nzb, c->pitch * sizeof(float),
                             (size_t)p.nz * sizeof(float), c->nli, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return check_device_flag(c);
}

static int backward_core(fdw_ctx *c, const float *P, const float *PP, const float *dobs, int gz, float *imloc);
static void graph_drop_if_any(fdw_ctx *c); /* recorded launches hold the old velocity pointer */

extern "C" int fdw_backward(fdw_ctx *c, const float *P, const float *PP, const float *dobs, int gz, float *imloc)
{
    if (!c || !dobs || !imloc || ((P == nullptr) != (PP == nullptr))) return FDW_ERR_ARG;
    return backward_core(c, P, PP, dobs, gz, imloc);
}

extern "C" int fdw_backward_device(fdw_ctx *c, const float *dobs, int gz)
{
    if (!c || !dobs) return FDW_ERR_ARG;
    return backward_core(c, nullptr, nullptr, dobs, gz, nullptr);
}

/* imloc == nullptr: leave the shot image on the device (no download, no synchronisation) */
static int backward_core(fdw_ctx *c, const float *P, const float *PP, const float *dobs, int gz, float *imloc)
{
    const fdw_params &p = c->prm;
    const int nt = p.nt;
    if (nt < 1) { fdw_set_error("fdw_backward: params.nt not set"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    if (P) {
        CHECK(fdw_fields_upload(c, 0, PP, P)); /* newest = PP = u(T), older = P = u(T-1) */
    } else if (!c->saved_valid) {
        fdw_set_error("fdw_backward: no saved levels on the device (run fdw_forward first)");
        return FDW_ERR_STATE;
    }
    /* time-reversed roles: S_0 = PP, S_1 = P (fd-code.cu:304-314) */
    int s_pp = c->newest[0], s_p = c->older[0];
    CHECK(fdw_fields_zero(c, 1));
    if (c->nli != p.nx) { fdw_set_error("fdw_backward: not available on a slab context"); return FDW_ERR_UNSUPPORTED; }
    CU(cudaMemsetAsync(c->img, 0, (size_t)p.nx * c->pitch * sizeof(float), c->stream));
    const size_t ntr = (size_t)p.nx * nt;
    CHECK(ensure_buffer(&c->dobs_d, &c->dobs_cap, ntr));
    CU(cudaMemcpyAsync(c->dobs_d, dobs, ntr * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    const int nrec = p.compat_extents ? (p.nx < c->upd_i1 ? p.nx : c->upd_i1) : p.nx;
    auto fill_back = [&](StepArgs &a) {
        a.dobs = c->dobs_d; a.dobs_base = 0; a.dobs_len = (long long)ntr;
        a.inj_gi0 = p.nxb; a.inj_n = nrec; a.inj_j = gz; a.inj_nt = nt;
        a.img = c->img; a.img_gi0 = p.nxb; a.img_n = nrec;
    };
    /* small grids: the whole backward pass -- both field pairs, injection, imaging -- in ONE launch of the
     * shared-memory tile kernel */
    if (c->use_tile && nt >= 2) {
        const void *k = tile_kernel(p.order, p.recipe, -1);
        StepArgs base;
        base_args(c, 1, &base);
        TilePlan tp;
        if (k && small_whole_grid(c, base) && tile_plan(c, base.ncol4, base.row1 - base.row0, 5, &tp)) {
            const bool tap = p.taper != FDW_TAPER_NONE;
            TileArgs ta;
            memset(&ta, 0, sizeof ta);
            persist_args(c, 1, base, tap, false, 0, nt, 0, fill_back, &ta.pa);
            ta.sav0 = c->f[s_pp].r0;
            ta.sav1 = c->f[s_p].r0;
            if (tile_launch(c, k, ta, tp)) {
                replay_bookkeeping(c, nt, tap, 1);
                c->saved_valid = false;
                return imloc ? image_download(c, imloc) : FDW_OK;
            }
        }
    }
    int cur = s_pp, prev1 = -1, prev2 = -1;
    for (int it = 0; it < nt; it++) {
        if (it == 0) {
            cur = s_pp;
        } else if (it == 1) {
            prev1 = s_pp;
            cur = s_p;
        } else {
            /* same update run backwards in time: no sponge, no source (fd-code.cu:317-318) */
            prev2 = prev1;
            prev1 = cur;
            c->newest[0] = prev1;
            c->older[0] = prev2;
            CHECK(step_pair(c, 0, p.recipe, 0, false, false, it, [](StepArgs &) {}));
            cur = prev2;
        }
        const float *srcfield = c->f[cur].r0;
        CHECK(step_pair(c, 1, p.recipe, fdw::EPI_INJECT | fdw::EPI_IMG_FIELD, true, false, it, [&](StepArgs &a) {
            a.dobs = c->dobs_d; a.dobs_base = 0; a.dobs_len = (long long)ntr;
            a.inj_gi0 = p.nxb; a.inj_n = nrec; a.inj_j = gz; a.inj_nt = nt; a.inj_tidx = nt - 1 - it;
            a.img = c->img; a.img_gi0 = p.nxb; a.img_n = nrec; a.img_field = srcfield;
        }));
    }
    c->saved_valid = false;
    return imloc ? image_download(c, imloc) : FDW_OK;
}

/* ------------------------------------------------------------------ shot loop without host round trips */
__global__ void k_stack_add(float *stack, const float *img, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stack[i] = __fadd_rn(stack[i], img[i]); /* img += imloc, fd-code.cu:525 */
}
#ifdef FDW_EMU
static void thunk_stack_add(void **a) { k_stack_add(*(float **)a[0], *(const float **)a[1], *(long long *)a[2]); }
#endif

static int pipeline_init(fdw_ctx *c)
{
    if (c->copy_stream) return FDW_OK;
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_shot_done, cudaEventDisableTiming));
    return FDW_OK;
}

extern "C" int fdw_v2_stage(fdw_ctx *c, const float *v2)
{
    if (!c || !v2) return FDW_ERR_ARG;
    CHECK(bind(c));
    CHECK(pipeline_init(c));
    if (!c->vdt2_base) {
        CU(cudaMalloc(&c->vdt2_base, c->field_elems * sizeof(float)));
        CU(cudaMemsetAsync(c->vdt2_base, 0, c->field_elems * sizeof(float), c->copy_stream));
    }
    /* the staging buffer was the current velocity until the last commit: wait for the work enqueued before it */
    if (c->shot_done_valid) CU(cudaStreamWaitEvent(c->copy_stream, c->ev_shot_done, 0));
    float *r0 = c->vdt2_base + (size_t)(AGUARD + 1) * c->pitch;
    int lo = -GUARD, hi = c->nloc + GUARD;
    if (c->gx0 + lo < 0) lo = -c->gx0;
    if (c->gx0 + hi > c->nxe) hi = c->nxe - c->gx0;
    CU(cudaMemcpy2DAsync(r0 + (long long)lo * c->pitch, c->pitch * sizeof(float), v2 + (size_t)(c->gx0 + lo) * c->nze,
                         (size_t)c->nze * sizeof(float), (size_t)c->nze * sizeof(float), hi - lo, cudaMemcpyHostToDevice,
                         c->copy_stream));
    long long n = (long long)c->field_elems;
    void *params[] = {&c->vdt2_base, &n, &c->dt2};
    CU(cudaLaunchKernel(FDW_KPTR(k_scale_rows, thunk_scale_rows), dim3((unsigned)((n + 255) / 256)), dim3(256), params, 0,
                        c->copy_stream));
    c->launches++;
    CU(cudaEventRecord(c->ev_staged, c->copy_stream));
    c->staged = true;
    return FDW_OK;
}

extern "C" int fdw_v2_commit(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    if (!c->staged) { fdw_set_error("fdw_v2_commit: nothing staged"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    /* everything enqueued so far used the old velocity: the next stage may overwrite it only after that */
    CU(cudaEventRecord(c->ev_shot_done, c->stream));
    c->shot_done_valid = true;
    CU(cudaStreamWaitEvent(c->stream, c->ev_staged, 0));
    float *t = c->vdt_base;
    c->vdt_base = c->vdt2_base;
    c->vdt2_base = t;
    c->vdt = c->vdt_base + (size_t)(AGUARD + 1) * c->pitch;
    c->staged = false;
    graph_drop_if_any(c);
    return FDW_OK;
}

static size_t stack_elems(const fdw_ctx *c) { return (size_t)(c->nli > 0 ? c->nli : 1) * c->pitch; }

extern "C" int fdw_stack_zero(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    if (!c->stack) CU(cudaMalloc(&c->stack, stack_elems(c) * sizeof(float)));
    CU(cudaMemsetAsync(c->stack, 0, stack_elems(c) * sizeof(float), c->stream));
    return FDW_OK;
}

extern "C" int fdw_stack_add(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    if (!c->stack) { fdw_set_error("fdw_stack_add: call fdw_stack_zero first"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    long long n = (long long)stack_elems(c);
    const float *img = c->img;
    void *params[] = {&c->stack, &img, &n};
    CU(cudaLaunchKernel(FDW_KPTR(k_stack_add, thunk_stack_add), dim3((unsigned)((n + 255) / 256)), dim3(256), params, 0, c->stream));
    c->launches++;
    return FDW_OK;
}

extern "C" int fdw_stack_download(fdw_ctx *c, float *img)
{
    if (!c || !img) return FDW_ERR_ARG;
    if (!c->stack) { fdw_set_error("fdw_stack_download: no stack"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    const fdw_params &p = c->prm;
    if (c->nli > 0)
        CU(cudaMemcpy2DAsync(img, (size_t)p.nz * sizeof(float), c->stack + p.nzb, c->pitch * sizeof(float),
                             (size_t)p.nz * sizeof(float), c->nli, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return check_device_flag(c);
}

extern "C" int fdw_stack_devptr(fdw_ctx *c, void **ptr, long long *pitch, int *rows)
{
    if (!c || !ptr || !pitch || !rows) return FDW_ERR_ARG;
    if (!c->stack) { fdw_set_error("fdw_stack_devptr: no stack"); return FDW_ERR_STATE; }
    *ptr = c->stack; *pitch = c->pitch; *rows = c->nli > 0 ? c->nli : 1;
    return FDW_OK;
}

/* ---- the three CPU-family shot phases, shared by the one-call pipelines below and by the
 * split-phase (slab) API: which epilogue runs and what it is given at time level `it` */
enum { PHASE_PLAIN = FDW_PHASE_PLAIN, PHASE_MODEL = FDW_PHASE_MODEL, PHASE_RTM_FWD = FDW_PHASE_RTM_FWD,
       PHASE_RTM_BWD = FDW_PHASE_RTM_BWD };

static int phase_epi(int phase)
{
    switch (phase) {
    case PHASE_MODEL: return fdw::EPI_RECORD;
    case PHASE_RTM_FWD: return fdw::EPI_HSTORE;
    case PHASE_RTM_BWD: return fdw::EPI_INJECT | fdw::EPI_IMG_HIST;
    default: return 0;
    }
}

static void phase_fill(const fdw_ctx *c, int phase, int it, StepArgs &a)
{
    const fdw_params &p = c->prm;
    const int nt = p.nt;
    const size_t slice = (size_t)(c->nli > 0 ? c->nli : 1) * c->pitch;
    if (phase == PHASE_MODEL) { /* mod_main.cpp:159-161 */
        a.rec = c->rec_d; a.rec_gi0 = c->li0; a.rec_n = c->nli; a.rec_j = c->shot_gz; a.rec_nt = nt; a.rec_it = it;
    } else if (phase == PHASE_RTM_FWD) { /* rtm_main.cpp:177-181 */
        a.hist_w = c->hist + (size_t)it * slice; a.hist_gi0 = c->li0; a.hist_n = c->nli;
    } else if (phase == PHASE_RTM_BWD) { /* rtm_main.cpp:201-203, 223-229 */
        a.dobs = c->dobs_d;
        a.dobs_base = (long long)c->shot_is * p.nx * nt;
        a.dobs_len = (long long)c->shot_ns * p.nx * nt;
        a.inj_gi0 = p.nzb; /* quirk Q5: rtm_main.cpp:202 offsets x by nzb */
        a.inj_n = p.nx; a.inj_j = c->shot_gz; a.inj_nt = nt; a.inj_tidx = nt - it; /* Q5: nt-it */
        a.hist_r = c->hist + (size_t)(nt - 1 - it) * slice; a.hist_gi0 = c->li0; a.hist_n = c->nli;
        a.img = c->img; a.img_gi0 = c->li0; a.img_n = c->nli;
    }
}

/* prepare buffers for one phase of a shot (fields zeroed; record / image / traces set up) */
static int phase_begin(fdw_ctx *c, int phase, int sx, int sz, int gz, const float *dobs_all, int ns, int is)
{
    const fdw_params &p = c->prm;
    const int nt = p.nt;
    if (nt < 1 || c->wavelet.size() < (size_t)nt) { fdw_set_error("shot: params.nt / wavelet not set"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    CHECK(fdw_fields_zero(c, 0));
    c->phase = phase; c->shot_gz = gz; c->shot_is = is; c->shot_ns = ns;
    const size_t slice = (size_t)(c->nli > 0 ? c->nli : 1) * c->pitch;
    if (phase == PHASE_MODEL) {
        CHECK(fdw_set_source(c, sx, sz, FDW_SRC_GAUSS7));
        const size_t ntr = (size_t)(c->nli > 0 ? c->nli : 1) * nt;
        CHECK(ensure_buffer(&c->rec_d, &c->rec_cap, ntr));
        CU(cudaMemsetAsync(c->rec_d, 0, ntr * sizeof(float), c->stream));
    } else if (phase == PHASE_RTM_FWD) {
        if (!c->hist) { fdw_set_error("shot: context created without params.history"); return FDW_ERR_STATE; }
        CHECK(fdw_set_source(c, sx, sz, FDW_SRC_POINT));
    } else if (phase == PHASE_RTM_BWD) {
        if (!c->hist || !dobs_all) { fdw_set_error("shot: history / traces missing"); return FDW_ERR_STATE; }
        CU(cudaMemsetAsync(c->img, 0, slice * sizeof(float), c->stream));
        const size_t ntot = (size_t)ns * p.nx * nt;
        CHECK(ensure_buffer(&c->dobs_d, &c->dobs_cap, ntot));
        CU(cudaMemcpyAsync(c->dobs_d, dobs_all, ntot * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    return FDW_OK;
}

static int phase_run(fdw_ctx *c, int phase)
{
    const int nt = c->prm.nt;
    const bool source = phase != PHASE_RTM_BWD;
    return run_levels(c, c->prm.recipe, phase_epi(phase), true, source, 0, nt, phase == PHASE_RTM_BWD ? 1 : 0,
                      [&](StepArgs &a) { phase_fill(c, phase, 0, a); },
                      [&](StepArgs &a, int it) { phase_fill(c, phase, it, a); });
}

extern "C" int fdw_model_shot(fdw_ctx *c, int sx, int sz, int gz, float *data)
{
    if (!c || !data) return FDW_ERR_ARG;
    CHECK(phase_begin(c, PHASE_MODEL, sx, sz, gz, nullptr, 1, 0));
    CHECK(phase_run(c, PHASE_MODEL));
    return fdw_shot_end(c, data);
}

extern "C" int fdw_rtm_shot_cpu(fdw_ctx *c, int sx, int sz, int gz, const float *dobs_all, int ns, int is,
                                float *imloc)
{
    if (!c || !dobs_all || !imloc || ns < 1 || is < 0 || is >= ns) return FDW_ERR_ARG;
    /* forward with history (rtm_main.cpp:166-188) */
    CHECK(phase_begin(c, PHASE_RTM_FWD, sx, sz, gz, nullptr, ns, is));
    CHECK(phase_run(c, PHASE_RTM_FWD));
    /* backward + imaging on the fly (rtm_main.cpp:196-229): imloc += swf[nt-1-it]*rwf[it], it ascending --
     * the same float summation order as the reference's third loop, so rwf is never stored */
    CHECK(phase_begin(c, PHASE_RTM_BWD, sx, sz, gz, dobs_all, ns, is));
    CHECK(phase_run(c, PHASE_RTM_BWD));
    return fdw_shot_end(c, imloc);
}

/* split-phase versions (slab decomposition): the caller drives the levels with
 * fdw_step_begin / fdw_step_rows / fdw_step_end and exchanges halos in between */
extern "C" int fdw_shot_begin(fdw_ctx *c, int phase, int sx, int sz, int gz, const float *dobs_all, int ns, int is)
{
    if (!c || phase < PHASE_PLAIN || phase > PHASE_RTM_BWD) return FDW_ERR_ARG;
    if (phase == PHASE_PLAIN) { c->phase = phase; return FDW_OK; }
    return phase_begin(c, phase, sx, sz, gz, dobs_all, ns, is);
}

extern "C" int fdw_shot_end(fdw_ctx *c, float *out)
{
    if (!c || !out) return FDW_ERR_ARG;
    CHECK(bind(c));
    const int phase = c->phase;
    c->phase = PHASE_PLAIN;
    if (phase == PHASE_MODEL) { /* this slab's traces, [nli][nt] */
        const size_t ntr = (size_t)c->nli * c->prm.nt;
        if (ntr) CU(cudaMemcpyAsync(out, c->rec_d, ntr * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        return FDW_OK;
    }
    if (phase == PHASE_RTM_BWD) return image_download(c, out); /* this slab's image rows, [nli][nz] */
    fdw_set_error("fdw_shot_end: phase %d has no host output", phase);
    return FDW_ERR_STATE;
}

extern "C" int fdw_shot_run(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    if (c->phase < PHASE_MODEL || c->phase > PHASE_RTM_BWD) { fdw_set_error("fdw_shot_run: no shot phase open"); return FDW_ERR_STATE; }
    if (c->gx0 != 0 || c->nloc != c->nxe) { fdw_set_error("fdw_shot_run: slab contexts drive their levels with fdw_step_* / fdw_peer_levels"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    return phase_run(c, c->phase);
}

/* ------------------------------------------------------------------ stencil program */
static int launch_lap(fdw_ctx *c, const float *src, float *dst)
{
    const void *k = lap_kernel(c->prm.order);
    StepArgs a;
    base_args(c, 0, &a);
    a.p = src;
    a.col4_0 = 0;
    a.ncol4 = (c->nze + 3) / 4;
    a.row0 = 0;
    a.row1 = c->nloc;
    a.lap_i1 = c->nxe - c->H; a.lap_j1 = c->nze - c->H; /* the stencil program rounds its grid up */
    dim3 grid, block;
    launch_geometry(k, c->nsm, a.ncol4, a.row1, c->threads_override, c->rows_per_cta_override, &grid, &block,
                    &a.rows_per_cta);
    void *params[] = {&a, &dst};
    CU(cudaLaunchKernel(k, grid, block, params, 0, c->stream));
    c->launches++;
    return FDW_OK;
}

extern "C" int fdw_laplacian_device(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    return launch_lap(c, c->f[c->newest[0]].r0, c->f[c->older[0]].r0);
}

extern "C" int fdw_stencil(int order, int nxe, int nze, float dx, float dz, const float *in, float *out, int device)
{
    if (!in || !out || nxe < 1 || nze < 1) return FDW_ERR_ARG;
    fdw_params p;
    memset(&p, 0, sizeof p);
    p.nx = nxe; p.nz = nze; p.order = order; p.dx = dx; p.dz = dz; p.dt = 1.0f; p.fac = 0.5f;
    p.family = FDW_FAMILY_GPU; p.recipe = FDW_RECIPE_G; p.taper = FDW_TAPER_NONE; p.device = device;
    fdw_ctx *c = nullptr;
    CHECK(fdw_create(&p, &c));
    int rc = field_h2d(c, c->f[0].r0, in, 0);
    if (rc == FDW_OK) rc = launch_lap(c, c->f[0].r0, c->f[1].r0);
    if (rc == FDW_OK) rc = field_d2h(c, c->f[1].r0, out);
    if (rc == FDW_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) {
        fdw_set_error("fdw_stencil: %s", cudaGetErrorString(cudaGetLastError()));
        rc = FDW_ERR_CUDA;
    }
    fdw_destroy(c);
    return rc;
}

/* ------------------------------------------------------------------ image post-filter (laplace.f90:24-28) */
__global__ void k_image_lap(const float *img, float *out, int nx, int nz, float dx2, float dz2)
{
    const int iz = blockIdx.x * blockDim.x + threadIdx.x, ix = blockIdx.y;
    if (iz >= nz || ix >= nx) return;
    float o = 0.0f;
    if (ix >= 1 && ix < nx - 1 && iz >= 1 && iz < nz - 1) {
        const float *p = img + (size_t)ix * nz + iz;
        const float c = p[0];
        const float tz = __fdiv_rn(__fadd_rn(__fsub_rn(p[1], __fmul_rn(2.0f, c)), p[-1]), dz2);
        const float tx = __fdiv_rn(__fadd_rn(__fsub_rn(p[nz], __fmul_rn(2.0f, c)), p[-nz]), dx2);
        o = __fadd_rn(tz, tx);
    }
    out[(size_t)ix * nz + iz] = o;
}
#ifdef FDW_EMU
static void thunk_image_lap(void **a)
{
    k_image_lap(*(const float **)a[0], *(float **)a[1], *(int *)a[2], *(int *)a[3], *(float *)a[4], *(float *)a[5]);
}
#endif

extern "C" int fdw_image_laplacian(int nx, int nz, float dx, float dz, const float *img, float *out, int device)
{
    if (!img || !out || nx < 1 || nz < 1) return FDW_ERR_ARG;
    int ndev = fdw_device_count();
    if (ndev < 1 || device < 0 || device >= ndev) {
        fdw_set_error("fdw_image_laplacian: CUDA device %d not available (%d visible); libfdwave has no CPU path", device, ndev);
        return FDW_ERR_CUDA;
    }
    CU(cudaSetDevice(device));
    const size_t bytes = (size_t)nx * nz * sizeof(float);
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc(&d_in, bytes));
    if (cudaMalloc(&d_out, bytes) != cudaSuccess) { cudaFree(d_in); fdw_set_error("fdw_image_laplacian: out of device memory"); return FDW_ERR_NOMEM; }
    int rc = FDW_OK;
    float dx2 = dx * dx, dz2 = dz * dz;
    const float *cin = d_in;
    void *params[] = {&cin, &d_out, &nx, &nz, &dx2, &dz2};
    if (nx > 65535) { cudaFree(d_in); cudaFree(d_out); fdw_set_error("fdw_image_laplacian: nx > 65535 not supported"); return FDW_ERR_UNSUPPORTED; }
    if (cudaMemcpy(d_in, img, bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaLaunchKernel(FDW_KPTR(k_image_lap, thunk_image_lap), dim3((nz + 127) / 128, nx), dim3(128), params, 0, 0) != cudaSuccess ||
        cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
        fdw_set_error("fdw_image_laplacian: %s", cudaGetErrorString(cudaGetLastError()));
        rc = FDW_ERR_CUDA;
    }
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

/* ------------------------------------------------------------------ slab decomposition */
extern "C" int fdw_step_begin(fdw_ctx *c, int it)
{
    if (!c) return FDW_ERR_ARG;
    if (c->step_open) { fdw_set_error("fdw_step_begin: previous step not ended"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    Field &n = c->f[c->newest[0]], &o = c->f[c->older[0]];
    const bool tap = c->prm.taper != FDW_TAPER_NONE;
    if (tap && c->prm.family == FDW_FAMILY_GPU) { n.pend++; o.pend++; }
    base_args(c, 0, &c->step_args);
    c->step_args.taper_on = (n.pend || o.pend) ? 1 : 0;
    if (!c->wavelet.empty() && c->phase != PHASE_RTM_BWD) set_source_args(c, &c->step_args, it);
    phase_fill(c, c->phase, it, c->step_args);
    c->step_open = true;
    return FDW_OK;
}

extern "C" int fdw_step_rows(fdw_ctx *c, int row0, int row1, void *stream)
{
    if (!c) return FDW_ERR_ARG;
    if (!c->step_open) { fdw_set_error("fdw_step_rows: no open step"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    return launch_level(c, c->step_args, c->prm.recipe, phase_epi(c->phase), row0, row1,
                        stream ? (cudaStream_t)stream : c->stream);
}

extern "C" int fdw_step_end(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    if (!c->step_open) { fdw_set_error("fdw_step_end: no open step"); return FDW_ERR_STATE; }
    Field &n = c->f[c->newest[0]], &o = c->f[c->older[0]];
    o.pend = 0;
    if (c->prm.taper != FDW_TAPER_NONE && c->prm.family == FDW_FAMILY_CPU) { o.pend = 1; n.pend++; }
    int t = c->newest[0];
    c->newest[0] = c->older[0];
    c->older[0] = t;
    c->step_open = false;
    return FDW_OK;
}

extern "C" int fdw_halo_get(fdw_ctx *c, int level, fdw_halo *h)
{
    if (!c || !h || level < 0 || level > 1) return FDW_ERR_ARG;
    if (level == 1 && !c->step_open) { fdw_set_error("fdw_halo_get(level=1) outside a step"); return FDW_ERR_STATE; }
    /* level 1: the level being written = the "older" buffer of the open step */
    float *r0 = c->f[level == 1 ? c->older[0] : c->newest[0]].r0;
    h->send_lo = r0;
    h->send_hi = r0 + (long long)(c->nloc - GUARD) * c->pitch;
    h->recv_lo = r0 - (long long)GUARD * c->pitch;
    h->recv_hi = r0 + (long long)c->nloc * c->pitch;
    h->count = (long long)GUARD * c->pitch;
    return FDW_OK;
}

static int rows_h2d(fdw_ctx *c, float *r0, const float *host)
{
    CU(cudaMemcpy2DAsync(r0, c->pitch * sizeof(float), host, (size_t)c->nze * sizeof(float),
                         (size_t)c->nze * sizeof(float), c->nloc, cudaMemcpyHostToDevice, c->stream));
    return FDW_OK;
}

static int rows_d2h(fdw_ctx *c, const float *r0, float *host)
{
    CU(cudaMemcpy2DAsync(host, (size_t)c->nze * sizeof(float), r0, c->pitch * sizeof(float),
                         (size_t)c->nze * sizeof(float), c->nloc, cudaMemcpyDeviceToHost, c->stream));
    return FDW_OK;
}

extern "C" int fdw_set_v2_local(fdw_ctx *c, const float *v2)
{
    if (!c || !v2) return FDW_ERR_ARG;
    CHECK(bind(c));
    CU(cudaMemsetAsync(c->vdt_base, 0, c->field_elems * sizeof(float), c->stream));
    CHECK(rows_h2d(c, c->vdt, v2));
    long long n = (long long)c->field_elems;
    void *params[] = {&c->vdt_base, &n, &c->dt2};
    CU(cudaLaunchKernel(FDW_KPTR(k_scale_rows, thunk_scale_rows), dim3((unsigned)((n + 255) / 256)), dim3(256), params,
                        0, c->stream));
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
    return FDW_OK;
}

extern "C" int fdw_fields_upload_local(fdw_ctx *c, int pair, const float *newest, const float *older)
{
    if (!c || pair < 0 || pair > 1 || !newest || !older) return FDW_ERR_ARG;
    CHECK(bind(c));
    Field &n = c->f[c->newest[pair]], &o = c->f[c->older[pair]];
    CHECK(rows_h2d(c, n.r0, newest));
    CHECK(rows_h2d(c, o.r0, older));
    n.pend = o.pend = 0;
    return FDW_OK;
}

extern "C" int fdw_fields_download_local(fdw_ctx *c, int pair, float *newest, float *older)
{
    if (!c || pair < 0 || pair > 1) return FDW_ERR_ARG;
    CHECK(bind(c));
    Field &n = c->f[c->newest[pair]], &o = c->f[c->older[pair]];
    if (newest) { CHECK(materialize(c, n)); CHECK(rows_d2h(c, n.r0, newest)); }
    if (older) { CHECK(materialize(c, o)); CHECK(rows_d2h(c, o.r0, older)); }
    CU(cudaStreamSynchronize(c->stream));
    return FDW_OK;
}

extern "C" int fdw_fields_download_local_async(fdw_ctx *c, int pair, float *newest, float *older)
{
    if (!c || pair < 0 || pair > 1) return FDW_ERR_ARG;
    CHECK(bind(c));
    Field &n = c->f[c->newest[pair]], &o = c->f[c->older[pair]];
    if (newest) { CHECK(materialize(c, n)); CHECK(rows_d2h(c, n.r0, newest)); }
    if (older) { CHECK(materialize(c, o)); CHECK(rows_d2h(c, o.r0, older)); }
    return FDW_OK; /* no synchronisation: fdw_sync() before reading the host arrays */
}

/* ------------------------------------------------------------------ peer-memory halo exchange
 * One process per GPU: every slab exports CUDA IPC handles of its field buffers and flag block,
 * the neighbours map them, and from then on a time level needs no host-side communication call:
 * the boundary-strip launch stores its rows locally AND straight into the neighbour's ghost rows
 * over NVLink (EPI_PUSH), a release flag follows it in stream order, the interior launch overlaps
 * the transfer, and the next level's boundary launch is gated by an acquire wait on the flags the
 * neighbours wrote here.  Safety of the unsynchronised ghost-row writes: a slab writes level n+1
 * into the neighbour's copy of buffer Y only after it has seen the neighbour's flag n, which the
 * neighbour raises after its own last reader of Y's ghost rows (its boundary launch of level n)
 * has completed; fdw_peer_fence closes a run so that no push is still in flight when the caller
 * goes on to zero, upload or download the buffers. */
static_assert(sizeof(cudaIpcMemHandle_t) <= FDW_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int fdw_peer_export(fdw_ctx *c, fdw_peer_info *out)
{
    if (!c || !out) return FDW_ERR_ARG;
    CHECK(bind(c));
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    for (int k = 0; k < 4; k++) {
        CU(cudaIpcGetMemHandle(&h, c->f[k].base));
        memcpy(out->field[k], &h, sizeof h);
    }
    CU(cudaIpcGetMemHandle(&h, c->flags_d));
    memcpy(out->flags, &h, sizeof h);
    out->nloc = c->nloc; out->gx0 = c->gx0; out->device = c->prm.device; out->pitch = c->pitch;
    return FDW_OK;
}

extern "C" int fdw_peer_detach(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    for (int s = 0; s < 2; s++) {
        fdw_ctx::PeerSide &p = c->peer[s];
        if (!p.on) continue;
        for (int k = 0; k < 4; k++)
            if (p.fbase[k]) cudaIpcCloseMemHandle(p.fbase[k]);
        if (p.flags) cudaIpcCloseMemHandle(p.flags);
        p = fdw_ctx::PeerSide();
    }
    return FDW_OK;
}

extern "C" int fdw_peer_attach(fdw_ctx *c, const fdw_peer_info *lo, const fdw_peer_info *hi)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    CU(cudaStreamSynchronize(c->stream));
    fdw_peer_detach(c);
    const fdw_peer_info *info[2] = {lo, hi};
    for (int s = 0; s < 2; s++) {
        if (!info[s]) continue;
        const fdw_peer_info &pi = *info[s];
        const bool adjacent = s == 0 ? pi.gx0 + pi.nloc == c->gx0 : c->gx0 + c->nloc == pi.gx0;
        if (pi.pitch != c->pitch || !adjacent || pi.nloc < 2 * GUARD || c->nloc < 2 * GUARD) {
            fdw_set_error("fdw_peer_attach: neighbour %d (rows %d+%d, pitch %lld) does not adjoin this slab (rows %d+%d, pitch %lld)",
                          s, pi.gx0, pi.nloc, pi.pitch, c->gx0, c->nloc, c->pitch);
            fdw_peer_detach(c);
            return FDW_ERR_ARG;
        }
        fdw_ctx::PeerSide &p = c->peer[s];
        p.on = true; /* from here on fdw_peer_detach closes whatever has been mapped */
        p.nloc = pi.nloc;
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < 4 && e == cudaSuccess; k++) {
            memcpy(&h, pi.field[k], sizeof h);
            e = cudaIpcOpenMemHandle((void **)&p.fbase[k], h, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e == cudaSuccess) {
            memcpy(&h, pi.flags, sizeof h);
            e = cudaIpcOpenMemHandle((void **)&p.flags, h, cudaIpcMemLazyEnablePeerAccess);
        }
        if (e != cudaSuccess) {
            fdw_set_error("fdw_peer_attach: cudaIpcOpenMemHandle (neighbour %d, device %d): %s -- the slabs must live in "
                          "different processes on P2P-capable GPUs", s, pi.device, cudaGetErrorString(e));
            fdw_peer_detach(c);
            return FDW_ERR_CUDA;
        }
    }
    c->peer_seq = 0;
    CU(cudaMemsetAsync(c->errflag_d, 0, sizeof(int), c->stream)); /* a fresh attachment starts without a stale time-out */
    CU(cudaMemsetAsync(c->pcount_d, 0, sizeof(unsigned), c->stream));
    CU(cudaMemsetAsync(c->flags_d, 0, 2 * sizeof(unsigned), c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FDW_OK;
}

/* peer-mapped address that corresponds to this slab's local row 0, column 0 in the neighbour's
 * field buffer k: the lower neighbour's upper ghost rows start at its row nloc, the upper
 * neighbour's lower ghost rows end at its row 0 */
static float *peer_image(const fdw_ctx *c, int side, int k)
{
    const fdw_ctx::PeerSide &p = c->peer[side];
    if (!p.on) return nullptr;
    float *r0 = p.fbase[k] + (size_t)(AGUARD + 1) * c->pitch;
    return side == 0 ? r0 + (long long)p.nloc * c->pitch : r0 - (long long)c->nloc * c->pitch;
}

static int peer_wait(fdw_ctx *c, cudaStream_t st)
{
    int need_lo = c->peer[0].on, need_hi = c->peer[1].on;
    if (!need_lo && !need_hi) return FDW_OK;
    const unsigned *mine = c->flags_d;
    unsigned v = c->peer_seq;
    c->launches++;
    c->peer_waits++;
    if (c->rec) {
        RecLaunch r;
        r.kern = FDW_KPTR(k_peer_wait, thunk_peer_wait); r.grid = dim3(1); r.block = dim3(1);
        r.side = st == c->side && st != c->stream; r.kind = 1; r.level = c->rec_level;
        r.w_mine = mine; r.v = v; r.need_lo = need_lo; r.need_hi = need_hi; r.err = c->errflag_d; r.tmo = c->timeout_ns;
        c->rec->push_back(r);
        return FDW_OK;
    }
    void *params[] = {&mine, &v, &need_lo, &need_hi, &c->errflag_d, &c->timeout_ns};
    CU(cudaLaunchKernel(FDW_KPTR(k_peer_wait, thunk_peer_wait), dim3(1), dim3(1), params, 0, st));
    return FDW_OK;
}

static int peer_signal(fdw_ctx *c, cudaStream_t st)
{
    c->peer_seq++;
    /* this slab is the lower neighbour's upper neighbour (slot 1 there) and vice versa */
    unsigned *flo = c->peer[0].on ? c->peer[0].flags + 1 : nullptr;
    unsigned *fhi = c->peer[1].on ? c->peer[1].flags + 0 : nullptr;
    unsigned v = c->peer_seq;
    c->launches++;
    if (c->rec) {
        RecLaunch r;
        r.kern = FDW_KPTR(k_peer_signal, thunk_peer_signal); r.grid = dim3(1); r.block = dim3(1);
        r.side = st == c->side && st != c->stream; r.kind = 2; r.level = c->rec_level;
        r.s_lo = flo; r.s_hi = fhi; r.v = v;
        c->rec->push_back(r);
        return FDW_OK;
    }
    void *params[] = {&flo, &fhi, &v};
    CU(cudaLaunchKernel(FDW_KPTR(k_peer_signal, thunk_peer_signal), dim3(1), dim3(1), params, 0, st));
    return FDW_OK;
}

/* one time level of the peer-memory slab loop: issued directly, or recorded (c->rec) */
static int peer_level(fdw_ctx *c, int it)
{
    const int nc = c->ncol4;
    /* a failed level must leave this slab's sponge counts and flag sequence where its neighbours' are */
    const int pend_n0 = c->f[c->newest[0]].pend, pend_o0 = c->f[c->older[0]].pend;
    const unsigned seq0 = c->peer_seq;
    CHECK(fdw_step_begin(c, it));
    StepArgs a = c->step_args;
    const int epi = phase_epi(c->phase);
    const int wbuf = c->older[0]; /* the buffer this level is written to -- the same index on every slab */
    a.push_lo = peer_image(c, 0, wbuf);
    a.push_hi = peer_image(c, 1, wbuf);
    a.push_nloc = c->nloc;
    const int ilo = c->peer[0].on ? GUARD : 0, ihi = c->peer[1].on ? c->nloc - GUARD : c->nloc;
    /* the boundary chain (acquire -> boundary strips with the push -> release) runs on the side
     * stream, concurrently with the interior launch: a few microseconds of small kernels and the
     * NVLink transfer hide behind the interior update.  Both join before the next level. */
    const bool ev = !c->rec;
    int rc = FDW_OK;
    if (ev && (cudaEventRecord(c->ev_pfork, c->stream) != cudaSuccess ||
               cudaStreamWaitEvent(c->side, c->ev_pfork, 0) != cudaSuccess))
        rc = FDW_ERR_CUDA;
    Rect strips[2];
    int nstrip = 0;
    long long nctas = 0;
    for (int s = 0; s < 2; s++) {
        if (!c->peer[s].on) continue;
        int r0 = s == 0 ? 0 : c->nloc - GUARD, r1 = s == 0 ? GUARD : c->nloc;
        if (r0 < a.row0) r0 = a.row0;
        if (r1 > a.row1) r1 = a.row1;
        strips[nstrip] = Rect{0, nc, r0, r1, 1};
        nctas += rect_ctas(c, c->prm.recipe, epi | fdw::EPI_PUSH, strips[nstrip]);
        nstrip++;
    }
    /* acquire and release ride inside the boundary launches (two launches fewer on the critical path of a
     * thin slab's level); with nothing to launch they fall back to the stand-alone kernels */
    const bool fused = c->fuse_flags && nctas > 0;
    if (fused) {
        a.pw_flags = c->flags_d; a.pw_v = c->peer_seq; a.pw_lo = c->peer[0].on; a.pw_hi = c->peer[1].on;
        a.pw_err = c->errflag_d;
        a.pw_timeout_ns = c->timeout_ns;
        c->peer_seq++;
        a.ps_lo = c->peer[0].on ? c->peer[0].flags + 1 : nullptr; /* this slab is the lower neighbour's upper one */
        a.ps_hi = c->peer[1].on ? c->peer[1].flags + 0 : nullptr;
        a.ps_count = c->pcount_d; a.ps_v = c->peer_seq; a.ps_total = (unsigned)nctas;
        c->peer_waits++;
    } else if (rc == FDW_OK) {
        rc = peer_wait(c, c->side); /* the neighbours' rows of the newest level have landed */
    }
    for (int s = 0; s < nstrip && rc == FDW_OK; s++) {
        if (c->rec && s == 1) c->rec_par = 1; /* the two boundary strips are independent */
        rc = launch_rect(c, a, c->prm.recipe, epi | fdw::EPI_PUSH, strips[s], c->side);
        c->rec_par = 0;
    }
    if (rc == FDW_OK && !fused) rc = peer_signal(c, c->side);
    if (rc == FDW_OK && ev && cudaEventRecord(c->ev_pjoin, c->side) != cudaSuccess) rc = FDW_ERR_CUDA;
    if (rc == FDW_OK) rc = launch_level(c, c->step_args, c->prm.recipe, epi, ilo, ihi, c->stream);
    if (rc == FDW_OK && ev && cudaStreamWaitEvent(c->stream, c->ev_pjoin, 0) != cudaSuccess) rc = FDW_ERR_CUDA;
    if (rc != FDW_OK) {
        c->step_open = false;
        c->f[c->newest[0]].pend = pend_n0;
        c->f[c->older[0]].pend = pend_o0;
        c->peer_seq = seq0;
        return rc;
    }
    return fdw_step_end(c);
}

/* CUDA-graph replay of the level loop.  A thin slab's level is ~100 us of GPU work but 13 driver calls
 * (5 kernels, 4 event records, 4 stream waits): the host cannot enqueue them fast enough, the GPU starves
 * (measured: 16384^2 over 8 GPUs, 136 us per level for 94 us of arithmetic).  Two consecutive levels are
 * recorded (same bookkeeping, nothing launched), turned into one graph -- kernel nodes, the two streams'
 * orders and the fork/join points as edges -- and every further pair of levels only refreshes the nodes'
 * arguments (cudaGraphExecKernelNodeSetParams) and launches the graph once. */
static void graph_drop(fdw_ctx *c)
{
    if (c->lexec) cudaGraphExecDestroy(c->lexec);
    if (c->lgraph) cudaGraphDestroy(c->lgraph);
    c->lexec = nullptr; c->lgraph = nullptr;
    c->lnodes.clear(); c->lshape.clear();
}

static void graph_drop_if_any(fdw_ctx *c) { if (c->lexec) graph_drop(c); }

static void node_params(RecLaunch &r, cudaKernelNodeParams *kp, void **slots)
{
    *kp = cudaKernelNodeParams();
    kp->func = const_cast<void *>(r.kern);
    kp->gridDim = r.grid; kp->blockDim = r.block; kp->sharedMemBytes = 0; kp->extra = nullptr;
    if (r.kind == 0) { slots[0] = &r.a; }
    else if (r.kind == 1) { slots[0] = &r.w_mine; slots[1] = &r.v; slots[2] = &r.need_lo; slots[3] = &r.need_hi; slots[4] = &r.err; slots[5] = &r.tmo; }
    else { slots[0] = &r.s_lo; slots[1] = &r.s_hi; slots[2] = &r.v; }
    kp->kernelParams = slots;
#ifdef FDW_EMU
    /* the host stand-in has to copy the argument values; it is told their sizes through `extra` */
    static size_t sz_step[] = {sizeof(StepArgs), 0};
    static size_t sz_wait[] = {sizeof(const unsigned *), sizeof(unsigned), sizeof(int), sizeof(int), sizeof(int *), sizeof(unsigned long long), 0};
    static size_t sz_signal[] = {sizeof(unsigned *), sizeof(unsigned *), sizeof(unsigned), 0};
    kp->extra = (void **)(r.kind == 0 ? sz_step : r.kind == 1 ? sz_wait : sz_signal);
#endif
}

static bool same_shape(const std::vector<RecLaunch> &x, const std::vector<RecLaunch> &y)
{
    if (x.size() != y.size()) return false;
    for (size_t i = 0; i < x.size(); i++)
        if (x[i].kern != y[i].kern || x[i].side != y[i].side || x[i].par != y[i].par || x[i].level != y[i].level ||
            x[i].grid.x != y[i].grid.x || x[i].grid.y != y[i].grid.y || x[i].block.x != y[i].block.x)
            return false;
    return true;
}

static int graph_build(fdw_ctx *c, std::vector<RecLaunch> &rec)
{
    graph_drop(c);
    CU(cudaGraphCreate(&c->lgraph, 0));
    enum { NLANE = 3 };
    cudaGraphNode_t start = nullptr;          /* join of the previous level (empty node); none before the first */
    std::vector<cudaGraphNode_t> tails[NLANE]; /* frontier of each lane: what its next launch must wait for */
    std::vector<cudaGraphNode_t> prev_deps[NLANE];
    int level = -1;
    auto close_level = [&]() -> int {
        std::vector<cudaGraphNode_t> deps;
        for (int s = 0; s < NLANE; s++) { deps.insert(deps.end(), tails[s].begin(), tails[s].end()); tails[s].clear(); prev_deps[s].clear(); }
        if (deps.empty()) return FDW_OK;
        CU(cudaGraphAddEmptyNode(&start, c->lgraph, deps.data(), deps.size()));
        return FDW_OK;
    };
    for (size_t i = 0; i < rec.size(); i++) {
        RecLaunch &r = rec[i];
        if (r.level != level) { CHECK(close_level()); level = r.level; }
        cudaKernelNodeParams kp; void *slots[6];
        node_params(r, &kp, slots);
        const int ln = r.side;
        /* lane order within the level; a lane's first node hangs off the previous level's join; a `par`
         * node shares the dependencies of its predecessor and joins the lane's frontier beside it */
        const bool beside = r.par && !tails[ln].empty();
        std::vector<cudaGraphNode_t> deps;
        if (beside) deps = prev_deps[ln];
        else if (!tails[ln].empty()) deps = tails[ln];
        else if (start) deps.push_back(start);
        cudaGraphNode_t node;
        CU(cudaGraphAddKernelNode(&node, c->lgraph, deps.empty() ? nullptr : deps.data(), deps.size(), &kp));
        if (beside) tails[ln].push_back(node);
        else { prev_deps[ln] = deps; tails[ln].assign(1, node); }
        c->lnodes.push_back(node);
    }
    CHECK(close_level());
    CU(cudaGraphInstantiate(&c->lexec, c->lgraph, 0));
    c->lshape = rec;
    return FDW_OK;
}

/* nlev levels (an even number: the buffers alternate) from `it`: record, then build or refresh the graph and launch it */
static int peer_levels_graph(fdw_ctx *c, int it, int nlev)
{
    std::vector<RecLaunch> rec;
    c->rec = &rec;
    int rc = FDW_OK;
    for (int l = 0; l < nlev && rc == FDW_OK; l++) { c->rec_level = l; rc = peer_level(c, it + l); }
    c->rec = nullptr;
    CHECK(rc);
    return graph_submit(c, rec);
}

/* build the graph of a recorded launch list, or -- same kernels, lanes and grids as the last one -- refresh the
 * node arguments of the instantiated graph; launch it */
static int graph_submit(fdw_ctx *c, std::vector<RecLaunch> &rec)
{
    if (rec.empty()) return FDW_OK;
    if (!c->lexec || !same_shape(rec, c->lshape)) {
        CHECK(graph_build(c, rec));
    } else {
        for (size_t i = 0; i < rec.size(); i++) {
            cudaKernelNodeParams kp; void *slots[6];
            node_params(rec[i], &kp, slots);
            CU(cudaGraphExecKernelNodeSetParams(c->lexec, c->lnodes[i], &kp));
        }
    }
    CU(cudaGraphLaunch(c->lexec, c->stream));
    c->graph_replays++;
    return FDW_OK;
}

/* thin slabs: all nsteps levels in ONE cooperative launch of the persistent slab kernel; 1 = done */
static int try_pslab(fdw_ctx *c, int it0, int n)
{
    if (!c->use_pslab || !c->coop || n < 2 || c->step_open) return 0;
    const int epi = phase_epi(c->phase);
    const void *k = pslab_kernel(c->prm.order, c->prm.recipe, epi);
    if (!k) return 0;
    StepArgs base;
    base_args(c, 0, &base);
    const int rows = base.row1 - base.row0, nc = base.ncol4;
    if (rows <= 0 || nc <= 0 || (long long)nc * rows >= c->pslab_limit) return 0;
    const bool source = !c->wavelet.empty() && c->phase != PHASE_RTM_BWD;
    if (source && (it0 < 0 || (size_t)(it0 + n) > c->wavelet.size() || !c->wavelet_d)) return 0;
    const bool tap = c->prm.taper != FDW_TAPER_NONE;
    /* CTA width: the one that keeps the most useful lanes resident -- (fill of the last item column) x (resident
     * threads at this kernel's register count).  532 float4 columns: 128 threads -> 5 item columns, the last 16 % full,
     * 512 resident threads per SM; 96 threads -> 6 columns, 92 % full, 576 resident: measured 25.5-26.6 -> 23.9-24.4 us
     * per mod_main level, 28.1 -> 26.2 backward (profiles/r03b_pslab_threads.log) */
    int threads = c->pslab_threads;
    if (threads <= 0) {
        double best = -1.0;
        const int cand[2] = {128, 96}; /* (64: more resident lanes still, but twice the CTAs at the level barrier -- measured no gain) */
        for (int i = 0; i < 2; i++) {
            const int t = cand[i], cols = (nc + t - 1) / t;
            const double eff = (double)nc / ((double)cols * t) * (double)(cached_occupancy(k, t) * t);
            if (getenv("FDW_PSLAB_VERBOSE"))
                fprintf(stderr, "fdwave pslab: %d threads -> %d item columns, %d CTAs/SM, %.0f useful lanes/SM\n", t, cols,
                        cached_occupancy(k, t), eff);
            if (eff > best * 1.02) { best = eff; threads = t; }
        }
    }
    PSlabArgs sa;
    memset(&sa, 0, sizeof sa);
    persist_args(c, 0, base, tap, source, it0, n, c->phase == PHASE_RTM_BWD ? 1 : 0,
                 [&](StepArgs &a) { phase_fill(c, c->phase, it0, a); a.push_nloc = c->nloc; }, &sa.pa);
    sa.nbx = (nc + threads - 1) / threads;
    /* boundary rows only where a neighbour exists and inside the updated rows */
    sa.rows_lo = c->peer[0].on ? GUARD : 0;
    sa.rows_hi = c->peer[1].on ? GUARD : 0;
    if (rows < sa.rows_lo + sa.rows_hi || base.row0 != 0 || base.row1 != c->nloc) return 0;
    const int cap = c->nsm * cached_occupancy(k, threads);
    const int nbound = sa.nbx * ((sa.rows_lo ? 1 : 0) + (sa.rows_hi ? 1 : 0));
    const int mid_rows = rows - sa.rows_lo - sa.rows_hi;
    int chunks = (cap - nbound) / sa.nbx; /* interior row chunks so that every CTA has at most one item per level */
    if (chunks < 1) return 0;
    sa.rpc = mid_rows > 0 ? (mid_rows + chunks - 1) / chunks : 1;
    if (sa.rpc < 2) sa.rpc = 2;
    sa.nmid = mid_rows > 0 ? (mid_rows + sa.rpc - 1) / sa.rpc : 0;
    const int nitems = nbound + sa.nbx * sa.nmid;
    const int grid = nitems < cap ? nitems : cap;
    /* level l writes the buffer that is "older" at that level: older at entry for even l, newest at entry for odd l */
    const int wb[2] = {c->older[0], c->newest[0]};
    for (int par = 0; par < 2; par++) {
        sa.push_lo[par] = peer_image(c, 0, wb[par]);
        sa.push_hi[par] = peer_image(c, 1, wb[par]);
    }
    sa.pw_flags = c->flags_d;
    sa.ps_lo = c->peer[0].on ? c->peer[0].flags + 1 : nullptr; /* this slab is the lower neighbour's upper one */
    sa.ps_hi = c->peer[1].on ? c->peer[1].flags + 0 : nullptr;
    sa.ps_count = c->pcount_d;
    sa.pw_err = c->errflag_d;
    sa.seq0 = c->peer_seq;
    if (cudaMemsetAsync(c->barrier_d, 0, sizeof(unsigned), c->stream) != cudaSuccess) return 0;
    void *params[] = {&sa};
    cudaError_t e = cudaLaunchCooperativeKernel(k, dim3(grid), dim3(threads), params, 0, c->stream);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    c->launches++;
    c->pslab_launches++;
    c->persist_launches++;
    c->peer_waits++;
    c->peer_seq += (unsigned)n;
    replay_bookkeeping(c, n, tap);
    return 1;
}

extern "C" int fdw_peer_levels(fdw_ctx *c, int it0, int nsteps)
{
    if (!c || nsteps < 0) return FDW_ERR_ARG;
    if (!c->peer[0].on && !c->peer[1].on) { fdw_set_error("fdw_peer_levels: no neighbour attached"); return FDW_ERR_STATE; }
    CHECK(bind(c));
    if (try_pslab(c, it0, nsteps)) return FDW_OK;
    int it = it0;
    const int end = it0 + nsteps;
    if (c->use_graph && nsteps >= 6 && (long long)c->ncol4 * c->nloc < c->graph_limit) {
        /* the first two levels directly (their sponge bookkeeping differs from the steady state) */
        for (int k = 0; k < 2; k++) CHECK(peer_level(c, it++));
        const int g = end - it >= c->graph_levels ? c->graph_levels : 2;
        while (end - it >= g) { CHECK(peer_levels_graph(c, it, g)); it += g; }
    }
    while (it < end) CHECK(peer_level(c, it++));
    return FDW_OK;
}

extern "C" int fdw_peer_refresh(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    const int k = c->newest[0];
    const size_t bytes = (size_t)GUARD * c->pitch * sizeof(float);
    const float *r0 = c->f[k].r0;
    if (c->peer[0].on)
        CU(cudaMemcpyAsync(peer_image(c, 0, k), r0, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (c->peer[1].on)
        CU(cudaMemcpyAsync(peer_image(c, 1, k) + (long long)(c->nloc - GUARD) * c->pitch,
                           r0 + (long long)(c->nloc - GUARD) * c->pitch, bytes, cudaMemcpyDeviceToDevice, c->stream));
    CHECK(peer_signal(c, c->stream));
    return FDW_OK;
}

extern "C" int fdw_peer_fence(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    return peer_wait(c, c->stream);
}

/* ------------------------------------------------------------------ device-resident access */
extern "C" int fdw_devinfo_get(fdw_ctx *c, fdw_devinfo *o)
{
    if (!c || !o) return FDW_ERR_ARG;
    o->newest = c->f[c->newest[0]].r0;
    o->older = c->f[c->older[0]].r0;
    o->vdt = c->vdt;
    o->pitch = c->pitch;
    o->nloc = c->nloc; o->gx0 = c->gx0; o->nxe = c->nxe; o->nze = c->nze; o->guard = GUARD;
    o->li0 = c->li0; o->nli = c->nli;
    return FDW_OK;
}

extern "C" int fdw_mark_begin(fdw_ctx *c)
{
    if (!c) return FDW_ERR_ARG;
    CHECK(bind(c));
    CU(cudaEventRecord(c->ev0, c->stream));
    return FDW_OK;
}

extern "C" int fdw_mark_end(fdw_ctx *c, float *ms)
{
    if (!c || !ms) return FDW_ERR_ARG;
    CHECK(bind(c));
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return FDW_OK;
}

extern "C" long long fdw_launch_count(fdw_ctx *c) { return c ? c->launches : 0; }

extern "C" long long fdw_counter(fdw_ctx *c, int which)
{
    if (!c) return 0;
    switch (which) {
    case FDW_COUNTER_LAUNCHES: return c->launches;
    case FDW_COUNTER_GRAPH_REPLAYS: return c->graph_replays;
    case FDW_COUNTER_PERSIST_LAUNCHES: return c->persist_launches;
    case FDW_COUNTER_TILE_LAUNCHES: return c->tile_launches;
    case FDW_COUNTER_PSLAB_LAUNCHES: return c->pslab_launches;
    }
    return 0;
}
