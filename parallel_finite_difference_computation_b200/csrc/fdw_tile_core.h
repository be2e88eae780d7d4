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
 *   per level   update own points from shared memory (the same step_rows body as every other kernel:
 *               identical arithmetic, sponge-on-load, source, epilogues), storing each result into the
 *               tile AND into the global copy of the level (lv.mirror: fire-and-forget, L2)
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
/* the same as a byte address in the shared-memory window (wraps modulo 2^32 like the kernel's 32-bit cursor) */
__device__ __forceinline__ unsigned tile_origin_s(float *buf, const TileGeom &g)
{
    return (unsigned)__cvta_generic_to_shared(buf) + 4u * (unsigned)((GUARD - g.r0) * g.sp + (4 - 4 * g.c0));
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

/* this thread's float4 column and rows inside the tile; false = idle thread */
__device__ __forceinline__ bool tile_my_rows(const TileArgs &ta, const TileGeom &g, int *j0, int *rb, int *re)
{
    const int col = threadIdx.x % ta.tc4, chunk = threadIdx.x / ta.tc4;
    const int q = g.c0 + col;
    *j0 = 4 * q;
    *rb = g.r0 + chunk * ta.ch;
    *re = *rb + ta.ch < g.r1 ? *rb + ta.ch : g.r1;
    return q < g.c1 && *rb < g.r1;
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
    int j0, rb, re;
    const bool active = tile_my_rows(ta, g, &j0, &rb, &re);
    const unsigned nblk = gridDim.x;
    for (int l = 0; l < pa.nlevels; l++) {
        Level lv = persist_level_of(pa, l);
        float *gnew = lv.pp; /* global copy of the level being written */
        lv.s_p = tile_origin_s(sbuf[l & 1], g);
        lv.s_pp = tile_origin_s(sbuf[(l & 1) ^ 1], g);
        lv.s_vdt = tile_origin_s(sv, g);
        lv.mirror = gnew;
        lv.tz = stz - (4 * g.c0 - 4);
        lv.tx = stx - (g.r0 - GUARD);
        if (active && !(ta.dbg & 2)) step_rows<ORDER, RECIPE, true, EPI, true, true>(a, lv, j0, rb, re);
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
    int j0, rb, re;
    const bool active = tile_my_rows(ta, g, &j0, &rb, &re);
    const unsigned nblk = gridDim.x;
    for (int it = 0; it < pa.nlevels; it++) {
        /* ---- source field: cur = index of the level the imaging condition uses at this step */
        int cur = it & 1; /* it = 0: u(T) in ss[0]; it = 1: u(T-1) in ss[1]; then the update writes ss[it & 1] */
        if (it >= 2) {
            Level ls;
            ls.p = nullptr; ls.pp = nullptr; ls.vdt = nullptr;
            ls.s_p = tile_origin_s(ss[(it & 1) ^ 1], g); /* prev1 */
            ls.s_pp = tile_origin_s(ss[it & 1], g);      /* prev2, overwritten */
            ls.s_vdt = tile_origin_s(sv, g);
            ls.s_field = 0;
            ls.tz = stz - (4 * g.c0 - 4);
            ls.tx = stx - (g.r0 - GUARD);
            ls.mirror = gs[it & 1];
            ls.np = ls.no = 0; ls.src_on = 0; ls.src_amp = 0.0f; ls.rec_it = 0; ls.inj_tidx = 0;
            ls.hist_w = nullptr; ls.hist_r = nullptr; ls.img_field = nullptr;
            if (active) step_rows<ORDER, RECIPE, false, 0, true, true>(a, ls, j0, rb, re);
        }
        /* ---- receiver field */
        Level lv = persist_level_of(pa, it);
        float *gnew = lv.pp;
        lv.s_p = tile_origin_s(sr[it & 1], g);
        lv.s_pp = tile_origin_s(sr[(it & 1) ^ 1], g);
        lv.s_vdt = tile_origin_s(sv, g);
        lv.mirror = gnew;
        lv.tz = stz - (4 * g.c0 - 4);
        lv.tx = stx - (g.r0 - GUARD);
        lv.src_on = 0;
        lv.s_field = tile_origin_s(ss[cur], g);
        if (active) step_rows<ORDER, RECIPE, true, EPI_INJECT | EPI_IMG_FIELD, true, true>(a, lv, j0, rb, re);
        if (it + 1 == pa.nlevels) break;
        if (!tile_sync(ta, (unsigned)(it + 1), nblk, &lost)) return;
        if (it >= 2) tile_load_ring(ss[it & 1], gs[it & 1], a.apitch, g);
        tile_load_ring(sr[(it & 1) ^ 1], gnew, a.apitch, g);
        __syncthreads();
    }
}

} /* namespace fdw */
#endif
