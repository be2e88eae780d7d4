/*
 * fdw_tile_core.h -- shared-memory-resident propagation for small grids.
 *
 * The shipped models are 50-200 k points: one time level is well under a microsecond of arithmetic,
 * so a level's cost is whatever it takes the SMs to agree that the previous level is complete.
 * Round 1's persistent kernel kept the fields in L2 and re-read every operand (x window included)
 * after each device-wide barrier: 6.4 us per level.  Here the grid is cut into ntx x nty tiles, one
 * CTA per tile, all CTAs co-resident (cooperative launch), and each CTA keeps ITS tile of both time
 * levels and of v2*dt2 in shared memory for the whole phase:
 *
 *   per level   update own points from shared memory, ONE POINT PER THREAD (tile_point): with a few hundred
 *               points per SM the level is a latency chain, not a throughput problem -- the float4-column
 *               body of the big-grid kernels (one warp = ~700 dependent-ish instructions per level, measured
 *               2.7 us) is replaced by ~100 instructions per thread over 4x as many threads; the per-point
 *               operation sequence (recipe, sponge on load, source, epilogues) is the same, hence the same
 *               bits.  Each result goes into the tile AND into the global copy of the level (lv.mirror:
 *               fire-and-forget, L2)
 *               -> release/acquire counter barrier (measured on B200: 1.3 us for 148 CTAs,
 *                  profiles/r02a_kbench_sync_latencies.log)
 *               -> fetch only the halo ring (4 rows above/below, one float4 column left/right) of the
 *                  new level from L2 (ld.global.cg: other SMs wrote it) into the tile.
 *
 * The global copies stay complete at every level, so everything outside the kernel (exports, history,
 * images, the host bookkeeping of pending sponge counts) is unchanged.  The GPU family's backward pass
 * (fd_back, fd-code.cu:290-341) runs as ONE launch too: the time-reversed reconstruction of the source
 * field and the receiver field with back-injection and imaging advance together, one barrier per level.
 *
 * Device-only (shared memory, barriers): not part of the host-emulation build.
 */
#ifndef FDW_TILE_CORE_H
#define FDW_TILE_CORE_H

#include "fdw_step_core.h"

namespace fdw {

__device__ __forceinline__ void tile_red_release(unsigned *p)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned tile_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long tile_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

/* device-wide barrier of the co-resident grid: everything this CTA stored before it is visible to every
 * CTA after it.  Wall-clock (globaltimer) bail-out after pa.timeout_ns so that a lost CTA can never hang the GPU. */
/* `lost` is a shared-memory word of the CTA (zeroed at kernel start): a time-out is reported to the host
 * through error_flag and to the CTA's own threads through `lost`, so that the common path reads no global memory */
__device__ __forceinline__ bool tile_grid_barrier(unsigned *ctr, unsigned target, int *error_flag, volatile int *lost,
                                                  unsigned long long timeout_ns)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        tile_red_release(ctr);
        if (tile_ld_acquire(ctr) < target) {
            const unsigned long long t0 = tile_globaltimer();
            unsigned spins = 0;
            while (tile_ld_acquire(ctr) < target) {
                if ((++spins & 1023u) == 0 && tile_globaltimer() - t0 > timeout_ns) {
                    atomicExch(error_flag, 1);
                    *lost = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    return *lost == 0;
}

__device__ __forceinline__ void tile_st_release(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/* Neighbour synchronisation instead of a device-wide barrier: a tile's next level depends only on its own
 * points and on the halo ring its four edge neighbours own (cross stencil: no corners).  Every CTA publishes
 * "my level l is in the global copy" (bar.sync, then one release store: cumulative over the CTA's stores) and
 * waits for the same word of its neighbours only.  Measured on B200 (profiles/r02a_kbench_sync_latencies.log):
 * 1.1 us against 1.4 us for the counter barrier, and a slow tile only delays its neighbours.  Write-after-read
 * on the two alternating global copies is safe for the same reason: a tile starts writing level l+1 into the
 * copy that holds level l-1 only after its neighbours have published level l, i.e. after they finished reading
 * level l-1. */
__device__ __forceinline__ bool tile_neighbour_sync(unsigned *flags, const TileArgs &ta, unsigned level, int *error_flag,
                                                    volatile int *lost, unsigned long long timeout_ns)
{
    const int me = blockIdx.x, tx = me % ta.ntx, ty = me / ta.ntx;
    __syncthreads();
    if (threadIdx.x == 0) tile_st_release(flags + 32 * me, level);
    /* up to four pollers, in different warps when the CTA has them */
    const int stride = blockDim.x >= 160 ? 32 : 1;
    const int k = threadIdx.x / stride - 1; /* 0..3 for threads stride, 2*stride, .. */
    if (threadIdx.x % stride == 0 && k >= 0 && k < 4) {
        int nb = -1;
        if (k == 0 && tx > 0) nb = me - 1;
        if (k == 1 && tx + 1 < ta.ntx) nb = me + 1;
        if (k == 2 && ty > 0) nb = me - ta.ntx;
        if (k == 3 && ty + 1 < ta.nty) nb = me + ta.ntx;
        if (nb >= 0 && tile_ld_acquire(flags + 32 * nb) < level) {
            const unsigned long long t0 = tile_globaltimer();
            unsigned spins = 0;
            while (tile_ld_acquire(flags + 32 * nb) < level) {
                if ((++spins & 1023u) == 0 && tile_globaltimer() - t0 > timeout_ns) {
                    atomicExch(error_flag, 1);
                    *lost = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    return *lost == 0;
}

__device__ __forceinline__ bool tile_sync(const TileArgs &ta, unsigned level, unsigned nblk, volatile int *lost)
{
    if (ta.dbg & 8) return tile_grid_barrier(ta.pa.barrier, level * nblk, ta.pa.error_flag, lost, ta.pa.timeout_ns);
    return tile_neighbour_sync(ta.flags, ta, level, ta.pa.error_flag, lost, ta.pa.timeout_ns);
}

struct TileGeom {
    int c0, c1, r0, r1; /* owned float4 columns [c0,c1) and rows [r0,r1) */
    int sp;             /* shared-memory pitch in floats = 4 * (tc4 + 2) */
    int srows;          /* rows per shared buffer = tr + 2 * GUARD */
};

__device__ __forceinline__ TileGeom tile_geom(const TileArgs &ta)
{
    const StepArgs &a = ta.pa.base;
    TileGeom g;
    const int tx = blockIdx.x % ta.ntx, ty = blockIdx.x / ta.ntx;
    g.c0 = a.col4_0 + tx * ta.tc4;
    g.c1 = g.c0 + ta.tc4 < a.ncol4 ? g.c0 + ta.tc4 : a.ncol4;
    g.r0 = a.row0 + ty * ta.tr;
    g.r1 = g.r0 + ta.tr < a.row1 ? g.r0 + ta.tr : a.row1;
    g.sp = 4 * (ta.tc4 + 2);
    g.srows = ta.tr + 2 * GUARD;
    return g;
}

/* address of (row r, float column j) of a tile buffer, as an "origin" such that org + r*sp + j is it */
__device__ __forceinline__ float *tile_origin(float *buf, const TileGeom &g)
{
    return buf + (long long)(GUARD - g.r0) * g.sp + (4 - 4 * g.c0);
}
/* whole tile incl. halo ring <- global level (rows r0-GUARD .. r1+GUARD, float4 columns c0-1 .. c1) */
__device__ __forceinline__ void tile_load_all(float *buf, const float *glob, long long gpitch, const TileGeom &g)
{
    const int w4 = g.c1 - g.c0 + 2, nr = g.r1 - g.r0 + 2 * GUARD;
    float *org = tile_origin(buf, g);
    for (int e = threadIdx.x; e < w4 * nr; e += blockDim.x) {
        const int r = g.r0 - GUARD + e / w4, q = g.c0 - 1 + e % w4;
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(glob + (long long)r * gpitch + 4 * q));
        st4(org + (long long)r * g.sp + 4 * q, v);
    }
}

__device__ __forceinline__ void tile_zero(float *buf, const TileGeom &g)
{
    const int n4 = g.srows * g.sp / 4;
    for (int e = threadIdx.x; e < n4; e += blockDim.x) st4(buf + 4 * e, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
}

/* halo ring only: GUARD rows above and below (owned columns), one float4 column left and right (owned rows) */
__device__ __forceinline__ void tile_load_ring(float *buf, const float *glob, long long gpitch, const TileGeom &g)
{
    const int w4 = g.c1 - g.c0, nr = g.r1 - g.r0;
    const int nrow_part = 2 * GUARD * w4, ncol_part = 2 * nr;
    float *org = tile_origin(buf, g);
    for (int e = threadIdx.x; e < nrow_part + ncol_part; e += blockDim.x) {
        int r, q;
        if (e < nrow_part) {
            const int k = e / w4;
            r = k < GUARD ? g.r0 - GUARD + k : g.r1 + (k - GUARD);
            q = g.c0 + e % w4;
        } else {
            const int k = e - nrow_part;
            r = g.r0 + (k >> 1);
            q = (k & 1) ? g.c1 : g.c0 - 1;
        }
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(glob + (long long)r * gpitch + 4 * q));
        st4(org + (long long)r * g.sp + 4 * q, v);
    }
}

/* the sponge tables of this tile's columns / rows, staged once (the acquire of every level's barrier
 * invalidates L1, so reading them from global memory would cost an L2 round trip per dependent use) */
__device__ __forceinline__ void tile_load_tables(float *stz, float *stx, const StepArgs &a, const TileGeom &g)
{
    const int nz = 4 * (g.c1 - g.c0) + 12, nx = g.r1 - g.r0 + 2 * GUARD;
    for (int e = threadIdx.x; e < nz; e += blockDim.x) stz[e] = a.tz[4 * g.c0 - 4 + e];
    for (int e = threadIdx.x; e < nx; e += blockDim.x) stx[e] = a.tx[g.r0 - GUARD + e];
}

/* ---- one point.  Tile buffers are addressed through "origins" in the shared-memory window (32-bit byte
 * addresses, LDS / STS; arithmetic wraps modulo 2^32): org + 4 * (r * sp + j) is (global row r, column j). */
struct TilePt {
    unsigned n;   /* newer level (stencil input) */
    unsigned o;   /* older level in, new level out */
    unsigned v;   /* fl32(v2*dt2) */
    unsigned f;   /* EPI_IMG_FIELD: reconstructed source level */
    unsigned tz, tx; /* sponge tables: tz + 4*j, tx + 4*row */
    int sp;
};

__device__ __forceinline__ float lds(unsigned addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts(unsigned addr, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned tile_origin_s(const float *buf, const TileGeom &g)
{
    return (unsigned)__cvta_generic_to_shared(buf) + 4u * (unsigned)((GUARD - g.r0) * g.sp + (4 - 4 * g.c0));
}

/* sponge applied cnt times to one value at (row gi, column j): (v*Z)*X per application (kernel_tapper
 * fd-code.cu:94-117, taper_apply taper.c:47-67), Z only on rows the reference's launch covered (tz_ilim),
 * X only on the columns it applies to (tx_jlim) */
/* (a real call on purpose: the rare path must not bloat the instruction footprint of the level loop) */
static __device__ __noinline__ float tile_tap(const StepArgs &a, const TilePt &t, float v, int gi, int j, int cnt)
{
    const float zf = gi < a.tz_ilim ? lds(t.tz + 4u * (unsigned)j) : 1.0f;
    const float xf = j < a.tx_jlim ? lds(t.tx + 4u * (unsigned)gi) : 1.0f;
    for (int c = 0; c < cnt; c++) v = fmul(fmul(v, zf), xf);
    return v;
}

template <int ORDER, int RECIPE, bool TAPER, int EPI>
__device__ __forceinline__ void tile_point(const StepArgs &a, const Level &lv, const TilePt &t, const int gi, const int j)
{
    constexpr int H = ORDER / 2;
    const unsigned off = 4u * (unsigned)(gi * t.sp + j), rowb = 4u * (unsigned)t.sp;
    const unsigned ctr = t.n + off;
    float z[2 * H + 1], x[2 * H + 1];
#pragma unroll
    for (int k = 0; k <= 2 * H; k++) z[k] = lds(ctr + 4u * (unsigned)(k - H));
#pragma unroll
    for (int k = 0; k <= 2 * H; k++)
        if (k != H) x[k] = lds(ctr + (unsigned)(k - H) * rowb);
    float o = lds(t.o + off);
    const float v = lds(t.v + off);
    /* sponge on load; all factors are exactly 1 away from the sponge (StepArgs::tap_*), where nothing is done */
    if (TAPER && (lv.np | lv.no) &&
        (j - H < a.tap_jlo || j + H >= a.tap_jhi || gi - H < a.tap_ilo || gi + H >= a.tap_ihi)) {
#pragma unroll
        for (int k = 0; k <= 2 * H; k++) z[k] = tile_tap(a, t, z[k], gi, j + k - H, lv.np);
#pragma unroll
        for (int k = 0; k <= 2 * H; k++)
            if (k != H) x[k] = tile_tap(a, t, x[k], gi + k - H, j, lv.np);
        o = tile_tap(a, t, o, gi, j, lv.no);
    }
    x[H] = z[H];
    const float cc = z[H];
    float lap;
    if (RECIPE == RECIPE_G) { /* fd-code.cu:66-72 (the leading "0 +" dropped as in row_update) */
        float az = fmul(z[0], a.cz[0]), ax = fmul(x[0], a.cx[0]);
#pragma unroll
        for (int io = 1; io <= ORDER; io++) {
            az = fadd(az, fmul(z[io], a.cz[io]));
            ax = fadd(ax, fmul(x[io], a.cx[io]));
        }
        lap = fadd(az, ax);
    } else if (RECIPE == RECIPE_C) { /* fd.c:30-33 */
        float acm = fmul(fmul(z[0], a.cz[0]), a.dz2inv);
        acm = fadd(acm, fmul(fmul(x[0], a.cx[0]), a.dx2inv));
#pragma unroll
        for (int io = 1; io <= ORDER; io++) {
            acm = fadd(acm, fmul(fmul(z[io], a.cz[io]), a.dz2inv));
            acm = fadd(acm, fmul(fmul(x[io], a.cx[io]), a.dx2inv));
        }
        lap = acm;
    } else {
        float sm = fmul(cc, a.cz[H] + a.cx[H]);
#pragma unroll
        for (int d = 1; d <= H; d++) {
            sm = ffma(a.cz[H + d], z[H - d] + z[H + d], sm);
            sm = ffma(a.cx[H + d], x[H - d] + x[H + d], sm);
        }
        lap = sm;
    }
    if (gi < a.lap_i0 || gi >= a.lap_i1 || j < a.lap_j0 || j >= a.lap_j1) lap = 0.0f; /* ring (quirk Q2 read as 0) */
    float res = RECIPE == RECIPE_FAST ? ffma(v, lap, 2.0f * cc - o) : leap(cc, o, fmul(v, lap));
    /* ---- source */
    if (lv.src_on) {
        const int di = gi - a.src_gi, dj = j - a.src_j;
        if (di >= -a.src_rad && di <= a.src_rad && dj >= -a.src_rad && dj <= a.src_rad && j < a.nze)
            res = fadd(res, a.src_rad ? fmul(lv.src_amp, a.src_w[(di + 3) * 7 + (dj + 3)]) : lv.src_amp);
    }
    /* ---- receiver back-injection */
    if ((EPI & EPI_INJECT) && j == a.inj_j && gi >= a.inj_gi0 && gi < a.inj_gi0 + a.inj_n) {
        const long long idx = a.dobs_base + (long long)(gi - a.inj_gi0) * a.inj_nt + lv.inj_tidx;
        res = fadd(res, (idx >= 0 && idx < a.dobs_len) ? a.dobs[idx] : 0.0f);
    }
    sts(t.o + off, res);
    lv.mirror[(long long)gi * a.apitch + j] = res;
    /* ---- side outputs (global memory, pitch a.apitch) */
    if ((EPI & EPI_RECORD) && j == a.rec_j && gi >= a.rec_gi0 && gi < a.rec_gi0 + a.rec_n)
        a.rec[(long long)(gi - a.rec_gi0) * a.rec_nt + lv.rec_it] = TAPER ? tile_tap(a, t, cc, gi, j, 1) : cc;
    if ((EPI & EPI_HSTORE) && gi >= a.hist_gi0 && gi < a.hist_gi0 + a.hist_n)
        lv.hist_w[(long long)(gi - a.hist_gi0) * a.apitch + j] = cc;
    if ((EPI & EPI_IMG_HIST) && gi >= a.img_gi0 && gi < a.img_gi0 + a.img_n) {
        float *ip = a.img + (long long)(gi - a.img_gi0) * a.apitch + j;
        *ip = fadd(*ip, fmul(lv.hist_r[(long long)(gi - a.hist_gi0) * a.apitch + j], cc));
    }
    if ((EPI & EPI_IMG_FIELD) && gi >= a.img_gi0 && gi < a.img_gi0 + a.img_n) {
        float *ip = a.img + (long long)(gi - a.img_gi0) * a.apitch + j;
        *ip = fadd(*ip, fmul(lds(t.f + off), res));
    }
}

/* the points of the tile this thread owns: one per round (point index = thread + round * blockDim); the first
 * rounds' coordinates are worked out once, outside the level loop (an integer division per point otherwise) */
struct TileWork {
    enum { KEEP = 2 };
    int npts, w, r0, j00;
    int gi[KEEP], j[KEEP];
};

__device__ __forceinline__ TileWork tile_work(const TileGeom &g)
{
    TileWork k;
    k.w = 4 * (g.c1 - g.c0); k.npts = k.w * (g.r1 - g.r0); k.r0 = g.r0; k.j00 = 4 * g.c0;
#pragma unroll
    for (int r = 0; r < TileWork::KEEP; r++) {
        const int p = threadIdx.x + r * blockDim.x;
        k.gi[r] = g.r0 + p / k.w;
        k.j[r] = 4 * g.c0 + p % k.w;
    }
    return k;
}

template <int ORDER, int RECIPE, bool TAPER, int EPI>
__device__ __forceinline__ void tile_update(const StepArgs &a, const Level &lv, const TilePt &t, const TileWork &k)
{
#pragma unroll
    for (int r = 0; r < TileWork::KEEP; r++)
        if ((int)(threadIdx.x + r * blockDim.x) < k.npts) tile_point<ORDER, RECIPE, TAPER, EPI>(a, lv, t, k.gi[r], k.j[r]);
    for (int p = threadIdx.x + TileWork::KEEP * blockDim.x; p < k.npts; p += blockDim.x)
        tile_point<ORDER, RECIPE, TAPER, EPI>(a, lv, t, k.r0 + p / k.w, k.j00 + p % k.w);
}

/* forward phases: nlevels levels of pair 0 (plain, modelling, rtm forward with history, rtm backward) */
template <int ORDER, int RECIPE, int EPI>
__device__ __forceinline__ void tile_forward(const TileArgs &ta, float *smem)
{
    const PersistArgs &pa = ta.pa;
    const StepArgs &a = pa.base;
    const TileGeom g = tile_geom(ta);
    __shared__ int lost;
    if (threadIdx.x == 0) lost = 0;
    float *sbuf[2] = {smem, smem + (long long)g.srows * g.sp};
    float *sv = smem + 2LL * g.srows * g.sp;
    float *stz = smem + 3LL * g.srows * g.sp, *stx = stz + g.sp + 8;
    tile_load_tables(stz, stx, a, g);
    tile_load_all(sbuf[0], pa.bufN, a.apitch, g);
    tile_load_all(sbuf[1], pa.bufO, a.apitch, g);
    tile_load_all(sv, a.vdt, a.apitch, g);
    __syncthreads();
    const unsigned nblk = gridDim.x;
    const TileWork work = tile_work(g);
    const unsigned org[2] = {tile_origin_s(sbuf[0], g), tile_origin_s(sbuf[1], g)};
    TilePt t;
    t.v = tile_origin_s(sv, g); t.f = 0; t.sp = g.sp;
    t.tz = (unsigned)__cvta_generic_to_shared(stz) - 4u * (unsigned)(4 * g.c0 - 4);
    t.tx = (unsigned)__cvta_generic_to_shared(stx) - 4u * (unsigned)(g.r0 - GUARD);
    for (int l = 0; l < pa.nlevels; l++) {
        Level lv = persist_level_of(pa, l);
        float *gnew = lv.pp; /* global copy of the level being written */
        lv.mirror = gnew;
        t.n = org[l & 1];
        t.o = org[(l & 1) ^ 1];
        if (!(ta.dbg & 2)) tile_update<ORDER, RECIPE, true, EPI>(a, lv, t, work);
        if (l + 1 == pa.nlevels) break; /* the kernel boundary orders the last level */
        if (!(ta.dbg & 1) && !tile_sync(ta, (unsigned)(l + 1), nblk, &lost)) return;
        if (!(ta.dbg & 4)) tile_load_ring(sbuf[(l & 1) ^ 1], gnew, a.apitch, g);
        __syncthreads();
    }
}

/* GPU-family backward pass (fd_back, fd-code.cu:290-341) in one launch: per level the time-reversed
 * reconstruction of the source field (levels 0 and 1 are the two saved levels themselves, :304-314;
 * from level 2 on the plain update without sponge or source, :317-318) and the receiver field with the
 * sponge, the back-injected traces and the imaging condition against the reconstructed level. */
template <int ORDER, int RECIPE>
__device__ __forceinline__ void tile_backward(const TileArgs &ta, float *smem)
{
    const PersistArgs &pa = ta.pa; /* describes the RECEIVER pair (bufN / bufO / pend / sponge) */
    const StepArgs &a = pa.base;
    const TileGeom g = tile_geom(ta);
    __shared__ int lost;
    if (threadIdx.x == 0) lost = 0;
    const long long bsz = (long long)g.srows * g.sp;
    float *ss[2] = {smem, smem + bsz};           /* source field: [0] = u(T), [1] = u(T-1) at entry */
    float *sr[2] = {smem + 2 * bsz, smem + 3 * bsz}; /* receiver field */
    float *sv = smem + 4 * bsz;
    float *stz = smem + 5 * bsz, *stx = stz + g.sp + 8;
    tile_load_tables(stz, stx, a, g);
    float *gs[2] = {ta.sav0, ta.sav1};
    tile_load_all(ss[0], gs[0], a.apitch, g);
    tile_load_all(ss[1], gs[1], a.apitch, g);
    tile_load_all(sr[0], pa.bufN, a.apitch, g);
    tile_load_all(sr[1], pa.bufO, a.apitch, g);
    tile_load_all(sv, a.vdt, a.apitch, g);
    __syncthreads();
    const unsigned nblk = gridDim.x;
    const TileWork work = tile_work(g);
    const unsigned os[2] = {tile_origin_s(ss[0], g), tile_origin_s(ss[1], g)};
    const unsigned orr[2] = {tile_origin_s(sr[0], g), tile_origin_s(sr[1], g)};
    TilePt t;
    t.v = tile_origin_s(sv, g); t.f = 0; t.sp = g.sp;
    t.tz = (unsigned)__cvta_generic_to_shared(stz) - 4u * (unsigned)(4 * g.c0 - 4);
    t.tx = (unsigned)__cvta_generic_to_shared(stx) - 4u * (unsigned)(g.r0 - GUARD);
    for (int it = 0; it < pa.nlevels; it++) {
        /* ---- source field: cur = index of the level the imaging condition uses at this step */
        int cur = it & 1; /* it = 0: u(T) in ss[0]; it = 1: u(T-1) in ss[1]; then the update writes ss[it & 1] */
        if (it >= 2) {
            Level ls;
            ls.p = nullptr; ls.pp = nullptr; ls.vdt = nullptr; ls.tz = nullptr; ls.tx = nullptr;
            ls.mirror = gs[it & 1];
            ls.np = ls.no = 0; ls.src_on = 0; ls.src_amp = 0.0f; ls.rec_it = 0; ls.inj_tidx = 0;
            ls.hist_w = nullptr; ls.hist_r = nullptr; ls.img_field = nullptr;
            t.n = os[(it & 1) ^ 1]; /* prev1 */
            t.o = os[it & 1];       /* prev2, overwritten */
            tile_update<ORDER, RECIPE, false, 0>(a, ls, t, work);
            /* no barrier: the imaging below reads this level only at the thread's own points (same point-to-thread
             * mapping in both updates) */
        }
        /* ---- receiver field */
        Level lv = persist_level_of(pa, it);
        float *gnew = lv.pp;
        lv.mirror = gnew;
        lv.src_on = 0;
        t.n = orr[it & 1];
        t.o = orr[(it & 1) ^ 1];
        t.f = os[cur];
        tile_update<ORDER, RECIPE, true, EPI_INJECT | EPI_IMG_FIELD>(a, lv, t, work);
        if (it + 1 == pa.nlevels) break;
        if (!tile_sync(ta, (unsigned)(it + 1), nblk, &lost)) return;
        if (it >= 2) tile_load_ring(ss[it & 1], gs[it & 1], a.apitch, g);
        tile_load_ring(sr[(it & 1) ^ 1], gnew, a.apitch, g);
        __syncthreads();
    }
}

} /* namespace fdw */
#endif
