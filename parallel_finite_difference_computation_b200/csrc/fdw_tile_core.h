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
 *               fire-and-forget, L2).  The sponge is applied where a value ENTERS a tile buffer (own result,
 *               halo ring, initial load) instead of at each of its 17 uses: the tile is private to its CTA,
 *               so "sponge on load" degenerates to "sponge once" -- same multiplications in the same order
 *               (position-only factors), and the tiles that touch the sponge no longer set the pace
 *               (measured: 3.9 us of a 5.8 us level were the sponge tiles' 19 taps per point)
 *               -> FLAG-IN-DATA halo exchange: every result is also stored as one 8-byte {value, level tag}
 *                  pair into a global plane (aligned 8-byte stores are single transactions); a tile's threads
 *                  poll the pairs of their halo ring (4 points deep, no corners: cross stencil) until the tag
 *                  is the level they need and copy the value into the tile.  No flag, no fence, no barrier
 *                  between SMs: the data is its own "ready" signal, one store->L2->load trip per level
 *                  instead of store, release flag, poll, ring load (measured: 1.9 us -> see DESIGN.md).
 *                  Two planes per field alternate with the level parity: a tile can be at most one level
 *                  ahead of its neighbour (it needs the neighbour's previous level to advance), so a pair is
 *                  never overwritten before it was read.
 *                  (ta.ll == null keeps the earlier protocol: neighbour flags, measured 1.1 us against 1.4 us
 *                  for a counter barrier -- profiles/r02a_kbench_sync_latencies.log -- then ring loads.)
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
    int sp;             /* shared-memory pitch in floats: >= 4 * (tc4 + 2), = tile width mod 32 (conflict-free) */
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
    g.sp = ta.sp;
    g.srows = ta.tr + 2 * GUARD;
    return g;
}

/* address of (row r, float column j) of a tile buffer, as an "origin" such that org + r*sp + j is it */
__device__ __forceinline__ float *tile_origin(float *buf, const TileGeom &g)
{
    return buf + (long long)(GUARD - g.r0) * g.sp + (4 - 4 * g.c0);
}
/* ---- sponge, applied where a value enters a tile buffer */
struct TileSponge {
    const float *tz, *tx; /* shared-memory copies of the tables, origins: tz[j], tx[global row] */
};

/* some factor can differ from 1 at (row gi, column j) (StepArgs::tap_*) */
__device__ __forceinline__ bool tile_in_sponge(const StepArgs &a, int gi, int j)
{
    return j < a.tap_jlo || j >= a.tap_jhi || gi < a.tap_ilo || gi >= a.tap_ihi;
}

/* cnt applications of (v*Z)*X at (gi, j): kernel_tapper fd-code.cu:94-117, taper_apply taper.c:47-67; Z only on
 * rows the reference's launch covered (tz_ilim), X only on the columns it applies to (tx_jlim) */
__device__ __forceinline__ float tile_tap(const StepArgs &a, const TileSponge &sg, float v, int gi, int j, int cnt)
{
    if (cnt == 0 || !tile_in_sponge(a, gi, j)) return v;
    const float zf = gi < a.tz_ilim ? sg.tz[j] : 1.0f;
    const float xf = j < a.tx_jlim ? sg.tx[gi] : 1.0f;
    for (int c = 0; c < cnt; c++) v = fmul(fmul(v, zf), xf);
    return v;
}

__device__ __forceinline__ float4 tile_tap4(const StepArgs &a, const TileSponge &sg, float4 v, int gi, int j0, int cnt)
{
    if (cnt == 0) return v;
    v.x = tile_tap(a, sg, v.x, gi, j0, cnt);
    v.y = tile_tap(a, sg, v.y, gi, j0 + 1, cnt);
    v.z = tile_tap(a, sg, v.z, gi, j0 + 2, cnt);
    v.w = tile_tap(a, sg, v.w, gi, j0 + 3, cnt);
    return v;
}

/* whole tile incl. halo ring <- global level (rows r0-GUARD .. r1+GUARD, float4 columns c0-1 .. c1), the sponge
 * applied cnt times on the way in */
__device__ __forceinline__ void tile_load_all(float *buf, const float *glob, long long gpitch, const TileGeom &g,
                                              const StepArgs &a, const TileSponge &sg, int cnt)
{
    const int w4 = g.c1 - g.c0 + 2, nr = g.r1 - g.r0 + 2 * GUARD;
    float *org = tile_origin(buf, g);
    for (int e = threadIdx.x; e < w4 * nr; e += blockDim.x) {
        const int r = g.r0 - GUARD + e / w4, q = g.c0 - 1 + e % w4;
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(glob + (long long)r * gpitch + 4 * q));
        st4(org + (long long)r * g.sp + 4 * q, tile_tap4(a, sg, v, r, 4 * q, cnt));
    }
}

/* halo ring only: GUARD rows above and below (owned columns), one float4 column left and right (owned rows) */
__device__ __forceinline__ void tile_load_ring(float *buf, const float *glob, long long gpitch, const TileGeom &g,
                                               const StepArgs &a, const TileSponge &sg, int cnt)
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
        st4(org + (long long)r * g.sp + 4 * q, tile_tap4(a, sg, v, r, 4 * q, cnt));
    }
}

/* ---- flag-in-data halo exchange */
__device__ __forceinline__ void ll_store(unsigned long long *slot, float v, unsigned tag)
{
    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(slot), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_load(const unsigned long long *slot, unsigned &v, unsigned &tag)
{
    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(v), "=r"(tag) : "l"(slot) : "memory");
}

/* halo ring <- the neighbours' {value, tag} pairs of level `tag`: GUARD rows above / below the owned columns,
 * GUARD columns left / right of the owned rows; points outside the updated region never change and keep what
 * the initial load put there.  false = a neighbour never delivered (wall-clock time-out). */
__device__ __forceinline__ bool tile_poll_ring(float *buf, const unsigned long long *plane, long long gpitch, const TileGeom &g,
                                               const StepArgs &a, const TileSponge &sg, int cnt, unsigned tag, int *error_flag,
                                               unsigned long long timeout_ns)
{
    const int w = 4 * (g.c1 - g.c0), h = g.r1 - g.r0;
    const int nrow_part = 2 * GUARD * w, ncol_part = 2 * GUARD * h;
    float *org = tile_origin(buf, g);
    const int jlo = 4 * a.col4_0, jhi = 4 * a.ncol4;
    bool ok = true;
    for (int e = threadIdx.x; e < nrow_part + ncol_part; e += blockDim.x) {
        int r, j;
        if (e < nrow_part) {
            const int k = e / w;
            r = k < GUARD ? g.r0 - GUARD + k : g.r1 + (k - GUARD);
            j = 4 * g.c0 + e % w;
        } else {
            const int k = e - nrow_part, c = k % (2 * GUARD);
            r = g.r0 + k / (2 * GUARD);
            j = c < GUARD ? 4 * g.c0 - GUARD + c : 4 * g.c1 + (c - GUARD);
        }
        if (r < a.row0 || r >= a.row1 || j < jlo || j >= jhi) continue; /* nobody updates this point */
        const unsigned long long *slot = plane + (long long)r * gpitch + j;
        unsigned v, t;
        ll_load(slot, v, t);
        if (t != tag) {
            const unsigned long long t0 = wall_ns();
            unsigned spins = 0;
            do {
                ll_load(slot, v, t);
                if ((++spins & 1023u) == 0 && wall_ns() - t0 > timeout_ns) { atomicExch(error_flag, 1); ok = false; break; }
            } while (t != tag);
        }
        org[(long long)r * g.sp + j] = tile_tap(a, sg, __uint_as_float(v), r, j, cnt);
    }
    return ok;
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
    int sp;
    int tap_o;    /* sponge applications still missing on the older level's stored values at this level */
    int tap_new;  /* applications the new level needs as the NEXT level's stencil input (applied as it is stored) */
    unsigned long long *ll; /* {value, tag} plane of the level being written (origin: global row 0, column 0); null = off */
    unsigned tag;
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

template <int ORDER, int RECIPE, bool TAPER, int EPI>
__device__ __forceinline__ void tile_point(const StepArgs &a, const Level &lv, const TilePt &t, const TileSponge &sg,
                                           const int gi, const int j)
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
    /* the newer level sits in the tile with its sponge applications done (they were applied as the values
     * entered the buffer); the older level's own point may miss some */
    if (TAPER) o = tile_tap(a, sg, o, gi, j, t.tap_o);
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
    sts(t.o + off, TAPER ? tile_tap(a, sg, res, gi, j, t.tap_new) : res);
    lv.mirror[(long long)gi * a.apitch + j] = res;
    if (t.ll) ll_store(t.ll + (long long)gi * a.apitch + j, res, t.tag);
    /* ---- side outputs (global memory, pitch a.apitch) */
    if ((EPI & EPI_RECORD) && j == a.rec_j && gi >= a.rec_gi0 && gi < a.rec_gi0 + a.rec_n)
        a.rec[(long long)(gi - a.rec_gi0) * a.rec_nt + lv.rec_it] = TAPER ? tile_tap(a, sg, cc, gi, j, 1) : cc;
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
__device__ __forceinline__ void tile_update(const StepArgs &a, const Level &lv, const TilePt &t, const TileSponge &sg,
                                            const TileWork &k)
{
#pragma unroll
    for (int r = 0; r < TileWork::KEEP; r++)
        if ((int)(threadIdx.x + r * blockDim.x) < k.npts) tile_point<ORDER, RECIPE, TAPER, EPI>(a, lv, t, sg, k.gi[r], k.j[r]);
    for (int p = threadIdx.x + TileWork::KEEP * blockDim.x; p < k.npts; p += blockDim.x)
        tile_point<ORDER, RECIPE, TAPER, EPI>(a, lv, t, sg, k.r0 + p / k.w, k.j00 + p % k.w);
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
    float *const sbuf0 = smem, *const sbuf1 = smem + (long long)g.srows * g.sp; /* (no arrays indexed by the level parity:
                                                                                    they would live in local memory) */
    float *sv = smem + 2LL * g.srows * g.sp;
    float *stz = smem + 3LL * g.srows * g.sp, *stx = stz + g.sp + 8;
    tile_load_tables(stz, stx, a, g);
    __syncthreads();
    TileSponge sg;
    sg.tz = stz - (4 * g.c0 - 4);
    sg.tx = stx - (g.r0 - GUARD);
    /* the newer level enters its buffer with the sponge applications of level 0 done, the older one raw */
    tile_load_all(sbuf0, pa.bufN, a.apitch, g, a, sg, persist_np(pa, 0));
    tile_load_all(sbuf1, pa.bufO, a.apitch, g, a, sg, 0);
    tile_load_all(sv, a.vdt, a.apitch, g, a, sg, 0);
    __syncthreads();
    const unsigned nblk = gridDim.x;
    const TileWork work = tile_work(g);
    const unsigned org0 = tile_origin_s(sbuf0, g), org1 = tile_origin_s(sbuf1, g);
    TilePt t;
    t.v = tile_origin_s(sv, g); t.f = 0; t.sp = g.sp;
    for (int l = 0; l < pa.nlevels; l++) {
        Level lv = persist_level_of(pa, l);
        float *gnew = lv.pp; /* global copy of the level being written (raw values, like every other kernel's) */
        lv.mirror = gnew;
        t.n = (l & 1) ? org1 : org0;
        t.o = (l & 1) ? org0 : org1;
        t.tap_o = lv.no - (l ? persist_np(pa, l - 1) : 0); /* the older level was the newer one of level l-1 */
        t.tap_new = persist_np(pa, l + 1);
        t.ll = ta.ll ? ta.ll + (l & 1) * ta.ll_plane : nullptr;
        t.tag = (unsigned)(l + 1);
        if (!(ta.dbg & 2)) tile_update<ORDER, RECIPE, true, EPI>(a, lv, t, sg, work);
        if (l + 1 == pa.nlevels) break; /* the kernel boundary orders the last level */
        if (ta.ll) {
            if (!tile_poll_ring((l & 1) ? sbuf0 : sbuf1, t.ll, a.apitch, g, a, sg, t.tap_new, t.tag, pa.error_flag, pa.timeout_ns))
                lost = 1;
            __syncthreads();
            if (lost) return;
        } else {
            if (!(ta.dbg & 1) && !tile_sync(ta, (unsigned)(l + 1), nblk, &lost)) return;
            if (!(ta.dbg & 4)) tile_load_ring((l & 1) ? sbuf0 : sbuf1, gnew, a.apitch, g, a, sg, t.tap_new);
            __syncthreads();
        }
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
    float *const ss0 = smem, *const ss1 = smem + bsz;               /* source field: u(T), u(T-1) at entry */
    float *const sr0 = smem + 2 * bsz, *const sr1 = smem + 3 * bsz; /* receiver field */
    float *sv = smem + 4 * bsz;
    float *stz = smem + 5 * bsz, *stx = stz + g.sp + 8;
    tile_load_tables(stz, stx, a, g);
    __syncthreads();
    TileSponge sg;
    sg.tz = stz - (4 * g.c0 - 4);
    sg.tx = stx - (g.r0 - GUARD);
    float *const gs0 = ta.sav0, *const gs1 = ta.sav1;
    tile_load_all(ss0, gs0, a.apitch, g, a, sg, 0);
    tile_load_all(ss1, gs1, a.apitch, g, a, sg, 0);
    tile_load_all(sr0, pa.bufN, a.apitch, g, a, sg, persist_np(pa, 0));
    tile_load_all(sr1, pa.bufO, a.apitch, g, a, sg, 0);
    tile_load_all(sv, a.vdt, a.apitch, g, a, sg, 0);
    __syncthreads();
    const unsigned nblk = gridDim.x;
    const TileWork work = tile_work(g);
    const unsigned os0 = tile_origin_s(ss0, g), os1 = tile_origin_s(ss1, g);
    const unsigned or0 = tile_origin_s(sr0, g), or1 = tile_origin_s(sr1, g);
    TilePt t;
    t.v = tile_origin_s(sv, g); t.f = 0; t.sp = g.sp;
    for (int it = 0; it < pa.nlevels; it++) {
        /* ---- source field: cur = index of the level the imaging condition uses at this step */
        int cur = it & 1; /* it = 0: u(T) in ss[0]; it = 1: u(T-1) in ss[1]; then the update writes ss[it & 1] */
        if (it >= 2) {
            Level ls;
            ls.p = nullptr; ls.pp = nullptr; ls.vdt = nullptr; ls.tz = nullptr; ls.tx = nullptr;
            ls.mirror = (it & 1) ? gs1 : gs0;
            ls.np = ls.no = 0; ls.src_on = 0; ls.src_amp = 0.0f; ls.rec_it = 0; ls.inj_tidx = 0;
            ls.hist_w = nullptr; ls.hist_r = nullptr; ls.img_field = nullptr;
            t.n = (it & 1) ? os0 : os1; /* prev1 */
            t.o = (it & 1) ? os1 : os0; /* prev2, overwritten */
            t.tap_o = t.tap_new = 0;
            t.ll = ta.ll ? ta.ll + (2 + (it & 1)) * ta.ll_plane : nullptr; /* planes 2, 3: the source field */
            t.tag = (unsigned)(it + 1);
            tile_update<ORDER, RECIPE, false, 0>(a, ls, t, sg, work);
            /* no barrier: the imaging below reads this level only at the thread's own points (same point-to-thread
             * mapping in both updates) */
        }
        /* ---- receiver field */
        Level lv = persist_level_of(pa, it);
        float *gnew = lv.pp;
        lv.mirror = gnew;
        lv.src_on = 0;
        t.n = (it & 1) ? or1 : or0;
        t.o = (it & 1) ? or0 : or1;
        t.f = cur ? os1 : os0;
        t.tap_o = lv.no - (it ? persist_np(pa, it - 1) : 0);
        t.tap_new = persist_np(pa, it + 1);
        t.ll = ta.ll ? ta.ll + (it & 1) * ta.ll_plane : nullptr; /* planes 0, 1: the receiver field */
        t.tag = (unsigned)(it + 1);
        tile_update<ORDER, RECIPE, true, EPI_INJECT | EPI_IMG_FIELD>(a, lv, t, sg, work);
        if (it + 1 == pa.nlevels) break;
        if (ta.ll) {
            bool ok = true;
            if (it >= 2)
                ok = tile_poll_ring((it & 1) ? ss1 : ss0, ta.ll + (2 + (it & 1)) * ta.ll_plane, a.apitch, g, a, sg, 0, t.tag,
                                    pa.error_flag, pa.timeout_ns);
            ok = tile_poll_ring((it & 1) ? sr0 : sr1, t.ll, a.apitch, g, a, sg, t.tap_new, t.tag, pa.error_flag, pa.timeout_ns) && ok;
            if (!ok) lost = 1;
            __syncthreads();
            if (lost) return;
        } else {
            if (!tile_sync(ta, (unsigned)(it + 1), nblk, &lost)) return;
            if (it >= 2) tile_load_ring((it & 1) ? ss1 : ss0, (it & 1) ? gs1 : gs0, a.apitch, g, a, sg, 0);
            tile_load_ring((it & 1) ? sr0 : sr1, gnew, a.apitch, g, a, sg, t.tap_new);
            __syncthreads();
        }
    }
}

} /* namespace fdw */
#endif
