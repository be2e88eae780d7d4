/*
 * fdw_step_core.h -- per-thread body of the fused propagation kernel.
 *
 * One thread owns 4 consecutive z samples (one float4 column) and streams
 * along x (the slow axis) holding the (order+1)-row x window of the newer
 * time level in registers.  The z neighbours come from the two adjacent
 * aligned float4 of the centre row (L1 hits: the neighbouring threads
 * streamed them order/2 rows earlier).  Threads never exchange data and the
 * older level is read and overwritten by its owning thread only, so the
 * kernel is race-free without shared memory or barriers.
 *
 * What is fused (reference file:line):
 *   Laplacian      kernel_lap   cuda_reference_RTM/src/fd-code.cu:53-78
 *                  fd_step      dpct_gpu_rtm_domain_division/src/timestep/fd.c:28-37
 *   leapfrog       kernel_time  fd-code.cu:80-92, fd.c:39-43
 *   sponge         kernel_tapper fd-code.cu:94-117, taper_apply/2 taper.c:47-84
 *                  (applied ON LOAD to raw stored levels: position-only
 *                  factors, same sequence of float multiplies)
 *   source         kernel_src fd-code.cu:119-122, ptsrc ptsrc.c:12-58
 *   record         mod_main.cpp:159-161
 *   back-injection kernel_sism fd-code.cu:124-131, rtm_main.cpp:201-203
 *   imaging        kernel_img fd-code.cu:133-144, rtm_main.cpp:223-229
 *   history        rtm_main.cpp:177-181 (store), :223-229 (read)
 *
 * The header compiles for the device (nvcc) and, for the CPU-only unit tests
 * of the index/sponge/epilogue logic, for the host (tests/emu/).  The host
 * build is test infrastructure; the product only ever runs the device build.
 */
#ifndef FDW_STEP_CORE_H
#define FDW_STEP_CORE_H

#ifdef __CUDACC__
#define FDW_HD __host__ __device__ __forceinline__
#define FDW_HDM __host__ __device__ __forceinline__ /* for static member functions */
#define FDW_UNROLL _Pragma("unroll")
#else
#include <math.h>
#define FDW_HD static inline
#define FDW_HDM inline
#define FDW_UNROLL
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 v = {x, y, z, w}; return v; }
#endif

namespace fdw {

enum { RECIPE_G = 0, RECIPE_C = 1, RECIPE_FAST = 2 };
enum { EPI_RECORD = 1, EPI_INJECT = 2, EPI_HSTORE = 4, EPI_IMG_HIST = 8, EPI_IMG_FIELD = 16, EPI_PUSH = 32 };
enum { GUARD = 4 };  /* ghost rows exchanged between slabs (orders <= 8; slab decomposition needs order/2 <= GUARD) */
enum { AGUARD = 8 }; /* zero guard rows allocated above and below every slab; >= order/2 for every supported order */
enum { MAX_ORDER = 16 };
enum { MAX_RECTS = 5 };
/* float4 blocks of z neighbours a thread loads on each side of its own float4 column */
constexpr int zblocks(int order) { return (order / 2 + 3) / 4; }

struct StepArgs {
    const float *p;  /* newer level, local row 0 / column 0; stencil input, never written */
    float *pp;       /* older level in, next level out (in place, own column only) */
    const float *vdt; /* fl32(v2*dt2), same layout */
    long long pitch; /* floats per row, multiple of 32, >= nze+4 */
    long long apitch; /* row pitch of the history / image / halo-push targets (= pitch except in the tile kernel) */
    int col4_0;      /* first float4 column of this launch */
    int ncol4;       /* one past the last float4 column of this launch */
    int row0, row1;  /* local rows [row0,row1) handled by this launch */
    int rows_per_cta;
    int grow0;       /* global x index of local row 0 */
    int lap_i0, lap_i1, lap_j0, lap_j1; /* Laplacian is non-zero only inside (global) */
    int nze;         /* valid columns per row */
    float cz[MAX_ORDER + 1], cx[MAX_ORDER + 1]; /* G/FAST: premultiplied weights; C: both hold the raw weights */
    float dz2inv, dx2inv;
    float one;        /* 1.0f, opaque to the assembler: the packed IEEE add is fma(acc, one, prod) */
    /* sponge */
    int taper_on;
    int np, no;       /* multiplications pending on the newer / older level (0..2) */
    const float *tz;  /* z factor per column, AGUARD valid entries before [0] and up to pitch+AGUARD */
    const float *tx;  /* x factor per local row, valid on [-AGUARD, nloc+AGUARD) */
    int tx_jlim;      /* x factor applies to columns j < tx_jlim */
    int tz_ilim;      /* z factor applies to global rows < tz_ilim */
    int tap_jlo, tap_jhi; /* some factor != 1 only if j < tap_jlo or j >= tap_jhi ... */
    int tap_ilo, tap_ihi; /* ... or global row < tap_ilo or >= tap_ihi */
    /* source */
    int src_on, src_gi, src_j, src_rad;
    float src_amp;
    float src_w[49];
    /* receiver recording: rec[(gi-rec_gi0)*rec_nt + rec_it] = sample at column rec_j */
    float *rec;
    int rec_gi0, rec_n, rec_j, rec_nt, rec_it;
    /* back-injection: pp_new[gi][inj_j] += dobs[dobs_base + (gi-inj_gi0)*inj_nt + inj_tidx] */
    const float *dobs;
    long long dobs_base, dobs_len;
    int inj_gi0, inj_n, inj_j, inj_nt, inj_tidx;
    /* forward history (rows rec. as full pitched rows of the interior x range) */
    float *hist_w;
    const float *hist_r;
    int hist_gi0, hist_n;
    /* image, pitched like the field, row 0 = global row img_gi0 */
    float *img;
    int img_gi0, img_n;
    const float *img_field; /* EPI_IMG_FIELD: reconstructed source level, field layout */
    /* EPI_PUSH (slab decomposition over peer memory): the first / last GUARD owned rows are also
     * stored straight into the neighbour's ghost rows.  push_lo / push_hi are the peer-mapped
     * addresses that correspond to THIS slab's local row 0, column 0 in the lower / upper
     * neighbour's copy of the level being written (same pitch); null = no neighbour. */
    float *push_lo, *push_hi;
    int push_nloc;
    /* acquire / release folded into the boundary launches (optional; null = done by separate kernels):
     * every CTA first waits until the neighbours' "rows delivered" counters here reach pw_v, and the
     * last CTA of the level's boundary launches (ps_total CTAs in all) raises this slab's counter in
     * the neighbours' memory to ps_v after a system-scope fence. */
    const unsigned *pw_flags;
    unsigned pw_v;
    int pw_lo, pw_hi;
    int *pw_err;
    unsigned *ps_lo, *ps_hi, *ps_count;
    unsigned ps_v, ps_total;
    unsigned long long pw_timeout_ns; /* the acquire gives up (error flag 2, no update, no push) after this wall time */
    /* multi-rectangle launch (the sponge strips of one level side by side in ONE launch instead of one short
     * launch after the other): nrect > 0 -> a 1-D grid, CTA b belongs to the last rectangle with cta0 <= b and
     * replaces col4_0 / ncol4 / row0 / row1 / rows_per_cta by that rectangle's.  A rectangle narrower than a
     * warp folds the CTA: lw = 8 / 16 / 32 lanes side by side in z, blockDim/lw row chunks per CTA (a z sponge of
     * 40 columns then costs a 64-column strip of the slow sponge instantiation instead of a 128-column one). */
    int nrect;
    struct RectGeom { int c0, c1, r0, r1, rpc, nbx, cta0, lw; } rect[MAX_RECTS];
};

/* the few quantities that change from one time level to the next; the ordinary kernels copy
 * them out of StepArgs, the persistent multi-level kernel computes them per level */
struct Level {
    const float *p;
    float *pp;
    const float *vdt; /* fl32(v2*dt2) in the layout of p / pp */
    float *mirror;    /* tile kernel: the global copy of the level being written (null otherwise) */
    const float *tz, *tx; /* sponge tables (StepArgs::tz / tx, or the tile kernel's shared-memory copies) */
    float *push_lo, *push_hi; /* EPI_PUSH targets of this level (StepArgs::push_*; per level in the persistent slab kernel) */
    int np, no;
    int src_on;
    float src_amp;
    int rec_it, inj_tidx;
    float *hist_w;
    const float *hist_r;
    const float *img_field;
};

FDW_HD Level level_of(const StepArgs &a)
{
    Level lv;
    lv.p = a.p; lv.pp = a.pp; lv.vdt = a.vdt; lv.mirror = nullptr; lv.tz = a.tz; lv.tx = a.tx; lv.push_lo = a.push_lo; lv.push_hi = a.push_hi; lv.np = a.np; lv.no = a.no; lv.src_on = a.src_on; lv.src_amp = a.src_amp;
    lv.rec_it = a.rec_it; lv.inj_tidx = a.inj_tidx; lv.hist_w = a.hist_w; lv.hist_r = a.hist_r;
    lv.img_field = a.img_field;
    return lv;
}

/* arguments of the persistent kernel: `nlevels` consecutive time levels of ONE propagation in one
 * cooperative launch, a grid-wide barrier between levels (launch-bound small grids) */
struct PersistArgs {
    StepArgs base;
    float *bufN, *bufO; /* newest / older level at entry */
    int pendN, pendO;   /* their pending sponge counts at entry */
    int nlevels, it0, nt;
    int sponge;         /* 1: the context's sponge is active in this phase */
    int sponge_first;   /* 1: GPU-family order (sponge, update); 0: CPU-family order (update, sponge) */
    int source;         /* 1: add wavelet[it] at the source */
    int tidx_cpu;       /* back-injection sample: 0 -> nt-1-it (fd-code.cu:129), 1 -> nt-it (rtm_main.cpp:202) */
    const float *wavelet;
    float *hist;        /* forward history base, slices of hist_slice floats */
    long long hist_slice;
    unsigned *barrier;  /* zeroed before the launch */
    int *error_flag;
    unsigned long long timeout_ns; /* device-wide waits give up after this wall time */
};

/* sponge multiplications pending on the newer / older level at launch-relative level l, in closed form
 * (the host bookkeeping of step_pair, unrolled) */
FDW_HD int persist_np(const PersistArgs &pa, int l)
{
    if (!pa.sponge) return l == 0 ? pa.pendN : 0;
    if (pa.sponge_first) return l == 0 ? pa.pendN + 1 : 1;
    return l == 0 ? pa.pendN : 1;
}
FDW_HD int persist_no(const PersistArgs &pa, int l)
{
    if (!pa.sponge) return l == 0 ? pa.pendO : (l == 1 ? pa.pendN : 0);
    if (pa.sponge_first) return l == 0 ? pa.pendO + 1 : (l == 1 ? pa.pendN + 2 : 2);
    return l == 0 ? pa.pendO : (l == 1 ? pa.pendN + 1 : 2);
}

/* the Level of launch-relative level l, in closed form (see the host bookkeeping in step_pair) */
FDW_HD Level persist_level_of(const PersistArgs &pa, int l)
{
    Level lv;
    const int it = pa.it0 + l;
    lv.p = (l & 1) ? pa.bufO : pa.bufN;
    lv.pp = (l & 1) ? pa.bufN : pa.bufO;
    lv.vdt = pa.base.vdt;
    lv.mirror = nullptr;
   
    lv.tz = pa.base.tz; lv.tx = pa.base.tx;
    lv.push_lo = lv.push_hi = nullptr;
    lv.np = persist_np(pa, l);
    lv.no = persist_no(pa, l);
    lv.src_on = pa.source;
    lv.src_amp = pa.source ? pa.wavelet[it] : 0.0f;
    lv.rec_it = it;
    lv.inj_tidx = pa.tidx_cpu ? pa.nt - it : pa.nt - 1 - it;
    lv.hist_w = pa.hist ? pa.hist + (long long)it * pa.hist_slice : nullptr;
    lv.hist_r = pa.hist ? pa.hist + (long long)(pa.nt - 1 - it) * pa.hist_slice : nullptr;
    lv.img_field = nullptr;
    return lv;
}

FDW_HD float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
FDW_HD void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

#ifdef __CUDA_ARCH__
FDW_HD float4 ld4_stream(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
FDW_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
FDW_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
FDW_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
/* fl32( fl64( 2.*p - pp ) + t ): 2.*p is exact, so the DFMA rounds once exactly
 * like the reference's separate DMUL/DADD pair (fd-code.cu:89, fd.c:41). */
FDW_HD float leap(float p, float pp, float t)
{
    double d = __fma_rn(2.0, (double)p, -(double)pp);
    return __double2float_rn(__dadd_rn(d, (double)t));
}
#else
FDW_HD float4 ld4_stream(const float *p) { return ld4(p); }
FDW_HD float fmul(float a, float b) { return a * b; }
FDW_HD float fadd(float a, float b) { return a + b; }
FDW_HD float ffma(float a, float b, float c) { return fmaf(a, b, c); }
FDW_HD float leap(float p, float pp, float t)
{
    double d = 2.0 * (double)p - (double)pp;
    return (float)(d + (double)t);
}
#endif

/* wall-clock nanoseconds (time-outs of device-side waits must not depend on the SM clock or on time slicing) */
#ifdef __CUDA_ARCH__
FDW_HD unsigned long long wall_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#else
FDW_HD unsigned long long wall_ns() { return 0; }
#endif

/* ---- programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization
 * attribute may have its CTAs scheduled while the previous kernel of the stream drains; pdl_wait() -- before the
 * first global access -- blocks until that kernel has completed and its writes are visible; pdl_trigger() tells the
 * scheduler that the NEXT kernel's CTAs may be scheduled once every CTA of this one has started.  Both are no-ops in
 * a launch without the attribute.  On grids whose level is tens of microseconds this hides the launch gap and the
 * ramp-up behind the previous launch's tail. */
#ifdef __CUDA_ARCH__
FDW_HD void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
FDW_HD void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
FDW_HD void pdl_wait() {}
FDW_HD void pdl_trigger() {}
#endif

/* ---- packed FP32x2 arithmetic (sm_100a FMUL2 / FFMA2: two IEEE round-to-nearest results per
 * instruction, each lane rounded exactly like the scalar FMUL / FADD, so the reference's --fmad=false
 * bit pattern is kept while the FP32 issue slots halve).  ptxas contracts mul.rn.f32x2 + add.rn.f32x2
 * into one FFMA2 even under --fmad=false (checked in SASS), which would change the rounding; the
 * accumulate step is therefore written as fma(acc, one, prod) with `one` = 1.0f passed as a kernel
 * argument the assembler cannot see through: acc*1+prod rounds once, i.e. it IS the IEEE add, and an
 * FFMA2 cannot absorb a second multiply. */
#ifdef __CUDA_ARCH__
typedef unsigned long long f2;
FDW_HD f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
FDW_HD float lo2(f2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
FDW_HD float hi2(f2 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
FDW_HD f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
FDW_HD f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#else
struct f2 { float a, b; };
FDW_HD f2 pk(float a, float b) { f2 r = {a, b}; return r; }
FDW_HD float lo2(f2 v) { return v.a; }
FDW_HD float hi2(f2 v) { return v.b; }
FDW_HD f2 mul2(f2 a, f2 b) { return pk(a.a * b.a, a.b * b.b); }
FDW_HD f2 fma2(f2 a, f2 b, f2 c) { return pk(fmaf(a.a, b.a, c.a), fmaf(a.b, b.b, c.b)); }
#endif
FDW_HD f2 bc2(float c) { return pk(c, c); }
/* product of the z-row pair (za[i], za[i+1]) with one weight.  Even i: the pair is an aligned register
 * pair of a float4 load, one FMUL2.  Odd i: the two floats sit in different register pairs, so two scalar
 * FMULs write straight into a fresh aligned pair (no MOVs, no extra live registers); same value per lane. */
FDW_HD f2 zprod(const float *za, int i, float c)
{
    return (i & 1) ? pk(fmul(za[i], c), fmul(za[i + 1], c)) : mul2(pk(za[i], za[i + 1]), bc2(c));
}
/* IEEE add of two packed pairs: a*1 + b, one rounding (see above) */
FDW_HD f2 add2(f2 a, f2 one, f2 b) { return fma2(a, one, b); }

template <int K> FDW_HD float get(const float4 &v) { return K == 0 ? v.x : K == 1 ? v.y : K == 2 ? v.z : v.w; }
FDW_HD float getk(const float4 &v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

/* sponge applied cnt times to one float4 whose 4 columns have z factors zf[0..3];
 * the x factor xf applies to the columns flagged in xon (bit k).  Each
 * application is (v*Z)*X, the order of kernel_tapper / taper_apply. */
FDW_HD float4 tap4(float4 v, const float *zf, unsigned xon, float xf, bool zon, int cnt)
{
    float e[4] = {v.x, v.y, v.z, v.w};
    for (int c = 0; c < cnt; c++) {
        FDW_UNROLL
        for (int k = 0; k < 4; k++) {
            e[k] = fmul(e[k], zon ? zf[k] : 1.0f);
            e[k] = fmul(e[k], ((xon >> k) & 1u) ? xf : 1.0f);
        }
    }
    return make_float4(e[0], e[1], e[2], e[3]);
}

/* source patch on the 4 samples of one thread; deliberately not inlined so that the rare path
 * stays a real branch instead of being if-converted into every row of every thread */
#ifdef __CUDACC__
static __host__ __device__ __noinline__
#else
static
#endif
float4 add_source(const StepArgs &a, float src_amp, float4 r4, int gi, int j0)
{
    float res[4] = {r4.x, r4.y, r4.z, r4.w};
    const int di = gi - a.src_gi;
    for (int k = 0; k < 4; k++) {
        const int dj = j0 + k - a.src_j;
        if (dj >= -a.src_rad && dj <= a.src_rad && j0 + k < a.nze) {
            const float amp = a.src_rad ? fmul(src_amp, a.src_w[(di + 3) * 7 + (dj + 3)]) : src_amp;
            res[k] = fadd(res[k], amp);
        }
    }
    return make_float4(res[0], res[1], res[2], res[3]);
}

/* the z row of one thread as a flat array: zl[NB-1] .. zl[0], c4, zr[0] .. zr[NB-1] (zl[0] / zr[0] are the
 * aligned float4 next to the centre one); the centre sample k sits at index 4*NB + k */
template <int NB>
FDW_HD void z_row(const float4 (&zl)[NB], const float4 c4, const float4 (&zr)[NB], float (&za)[4 * (2 * NB + 1)])
{
    FDW_UNROLL
    for (int b = 0; b < NB; b++) {
        const float4 l = zl[NB - 1 - b], r = zr[b];
        za[4 * b + 0] = l.x; za[4 * b + 1] = l.y; za[4 * b + 2] = l.z; za[4 * b + 3] = l.w;
        za[4 * (NB + 1 + b) + 0] = r.x; za[4 * (NB + 1 + b) + 1] = r.y; za[4 * (NB + 1 + b) + 2] = r.z; za[4 * (NB + 1 + b) + 3] = r.w;
    }
    za[4 * NB + 0] = c4.x; za[4 * NB + 1] = c4.y; za[4 * NB + 2] = c4.z; za[4 * NB + 3] = c4.w;
}

/* Everything of ONE ROW of one thread's float4 column once its operands are in registers: Laplacian in the
 * selected recipe, leapfrog update, source patch, trace back-injection.  w[] is the rotating x window of the
 * newer level (w[(u+io)%W] = row lr-H+io), zl / zr the aligned float4 blocks left and right of the centre row
 * (one each up to order 8, two for orders 10..16), o4 the older level, v4 = fl32(v2*dt2). */
template <int ORDER, int RECIPE, int EPI, bool PACKED>
FDW_HD float4 row_update(const StepArgs &a, const Level &lv, const float4 (&w)[ORDER + 1], const int u,
                         const float4 (&zl)[zblocks(ORDER)], const float4 (&zr)[zblocks(ORDER)], const float4 o4,
                         const float4 v4, const int gi, const int j0, const bool ring, const bool near_src)
{
    constexpr int H = ORDER / 2, W = ORDER + 1, NB = zblocks(ORDER), C0 = 4 * NB;
    const float4 c4 = w[(u + H) % W];
    float za[4 * (2 * NB + 1)];
    z_row<NB>(zl, c4, zr, za);
    const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
    const float oo[4] = {o4.x, o4.y, o4.z, o4.w};
    const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
    float lap[4], res[4];
        if (!PACKED) {
            /* scalar arithmetic (round 1): kept for A/B measurements and as the readable statement
             * of the per-point operation sequence; the packed path below is bit-identical */
            FDW_UNROLL
            for (int k = 0; k < 4; k++) {
                if (RECIPE == RECIPE_G) {
                    float az = fmul(za[C0 + k - H], a.cz[0]);
                    float ax = fmul(getk(w[(u + 0) % W], k), a.cx[0]);
                    FDW_UNROLL
                    for (int io = 1; io <= ORDER; io++) {
                        az = fadd(az, fmul(za[C0 + k - H + io], a.cz[io]));
                        ax = fadd(ax, fmul(getk(w[(u + io) % W], k), a.cx[io]));
                    }
                    lap[k] = fadd(az, ax);
                } else if (RECIPE == RECIPE_C) {
                    float acm = fmul(fmul(za[C0 + k - H], a.cz[0]), a.dz2inv);
                    acm = fadd(acm, fmul(fmul(getk(w[(u + 0) % W], k), a.cx[0]), a.dx2inv));
                    FDW_UNROLL
                    for (int io = 1; io <= ORDER; io++) {
                        acm = fadd(acm, fmul(fmul(za[C0 + k - H + io], a.cz[io]), a.dz2inv));
                        acm = fadd(acm, fmul(fmul(getk(w[(u + io) % W], k), a.cx[io]), a.dx2inv));
                    }
                    lap[k] = acm;
                } else {
                    float sm = fmul(cc[k], a.cz[H] + a.cx[H]);
                    FDW_UNROLL
                    for (int d = 1; d <= H; d++) {
                        sm = ffma(a.cz[H + d], za[C0 + k - d] + za[C0 + k + d], sm);
                        sm = ffma(a.cx[H + d], getk(w[(u + H - d) % W], k) + getk(w[(u + H + d) % W], k), sm);
                    }
                    lap[k] = sm;
                }
            }
        } else {
            /* two packed pairs per float4: points (0,1) and (2,3).  Per lane the operation
             * sequence is the scalar one of the reference, only two lanes share an instruction. */
            const f2 one = bc2(a.one);
            constexpr int B = C0 - H; /* za index of the first z tap of point 0 */
            f2 l01, l23;
            if (RECIPE == RECIPE_G) {
                /* two accumulators, ascending io, summed last (fd-code.cu:66-72).
                 * The reference's "0 +" first add is dropped: it can only change
                 * the sign of an all-zero sum, which provably never reaches pp. */
                f2 az0 = zprod(za, B, a.cz[0]);
                f2 az1 = zprod(za, B + 2, a.cz[0]);
                FDW_UNROLL
                for (int io = 1; io <= ORDER; io++) { /* z first: l4 / r4 die here */
                    az0 = add2(az0, one, zprod(za, B + io, a.cz[io]));
                    az1 = add2(az1, one, zprod(za, B + io + 2, a.cz[io]));
                }
                f2 ax0 = mul2(pk(w[u % W].x, w[u % W].y), bc2(a.cx[0]));
                f2 ax1 = mul2(pk(w[u % W].z, w[u % W].w), bc2(a.cx[0]));
                FDW_UNROLL
                for (int io = 1; io <= ORDER; io++) {
                    const float4 &wx = w[(u + io) % W];
                    ax0 = add2(ax0, one, mul2(pk(wx.x, wx.y), bc2(a.cx[io])));
                    ax1 = add2(ax1, one, mul2(pk(wx.z, wx.w), bc2(a.cx[io])));
                }
                l01 = add2(az0, one, ax0);
                l23 = add2(az1, one, ax1);
            } else if (RECIPE == RECIPE_C) {
                /* one accumulator, z tap then x tap, (p*c)*d2inv (fd.c:30-33) */
                const f2 dz = bc2(a.dz2inv), dx = bc2(a.dx2inv);
                f2 a0 = mul2(zprod(za, B, a.cz[0]), dz);
                f2 a1 = mul2(zprod(za, B + 2, a.cz[0]), dz);
                a0 = add2(a0, one, mul2(mul2(pk(w[u % W].x, w[u % W].y), bc2(a.cx[0])), dx));
                a1 = add2(a1, one, mul2(mul2(pk(w[u % W].z, w[u % W].w), bc2(a.cx[0])), dx));
                FDW_UNROLL
                for (int io = 1; io <= ORDER; io++) {
                    const float4 &wx = w[(u + io) % W];
                    a0 = add2(a0, one, mul2(zprod(za, B + io, a.cz[io]), dz));
                    a1 = add2(a1, one, mul2(zprod(za, B + io + 2, a.cz[io]), dz));
                    a0 = add2(a0, one, mul2(mul2(pk(wx.x, wx.y), bc2(a.cx[io])), dx));
                    a1 = add2(a1, one, mul2(mul2(pk(wx.z, wx.w), bc2(a.cx[io])), dx));
                }
                l01 = a0;
                l23 = a1;
            } else {
                /* FAST: symmetric pairs + FMA; tolerance-checked, not bit-checked */
                f2 s0 = mul2(pk(cc[0], cc[1]), bc2(a.cz[H] + a.cx[H]));
                f2 s1 = mul2(pk(cc[2], cc[3]), bc2(a.cz[H] + a.cx[H]));
                FDW_UNROLL
                for (int d = 1; d <= H; d++) {
                    const float4 &wa = w[(u + H - d) % W], &wb = w[(u + H + d) % W];
                    s0 = fma2(bc2(a.cz[H + d]), add2(pk(za[C0 - d], za[C0 + 1 - d]), one, pk(za[C0 + d], za[C0 + 1 + d])), s0);
                    s1 = fma2(bc2(a.cz[H + d]), add2(pk(za[C0 + 2 - d], za[C0 + 3 - d]), one, pk(za[C0 + 2 + d], za[C0 + 3 + d])), s1);
                    s0 = fma2(bc2(a.cx[H + d]), add2(pk(wa.x, wa.y), one, pk(wb.x, wb.y)), s0);
                    s1 = fma2(bc2(a.cx[H + d]), add2(pk(wa.z, wa.w), one, pk(wb.z, wb.w)), s1);
                }
                l01 = s0;
                l23 = s1;
            }
            lap[0] = lo2(l01); lap[1] = hi2(l01); lap[2] = lo2(l23); lap[3] = hi2(l23);
        }
        if (ring) { /* rare: the ring of width order/2 keeps lap = 0 (quirk Q2 read as 0) */
            const bool row_in = gi >= a.lap_i0 && gi < a.lap_i1;
            FDW_UNROLL
            for (int k = 0; k < 4; k++)
                if (!row_in || j0 + k < a.lap_j0 || j0 + k >= a.lap_j1) lap[k] = 0.0f;
        }
        if (RECIPE == RECIPE_FAST) {
            FDW_UNROLL
            for (int k = 0; k < 4; k++) res[k] = ffma(vv[k], lap[k], 2.0f * cc[k] - oo[k]);
        } else if (!PACKED) {
            FDW_UNROLL
            for (int k = 0; k < 4; k++) res[k] = leap(cc[k], oo[k], fmul(vv[k], lap[k]));
        } else {
            const f2 t01 = mul2(pk(vv[0], vv[1]), pk(lap[0], lap[1]));
            const f2 t23 = mul2(pk(vv[2], vv[3]), pk(lap[2], lap[3]));
            res[0] = leap(cc[0], oo[0], lo2(t01));
            res[1] = leap(cc[1], oo[1], hi2(t01));
            res[2] = leap(cc[2], oo[2], lo2(t23));
            res[3] = leap(cc[3], oo[3], hi2(t23));
        }

    /* ---- source (after the update, before the sponge: both families) */
    if (near_src && gi >= a.src_gi - a.src_rad && gi <= a.src_gi + a.src_rad) {
        const float4 s4 = add_source(a, lv.src_amp, make_float4(res[0], res[1], res[2], res[3]), gi, j0);
        res[0] = s4.x; res[1] = s4.y; res[2] = s4.z; res[3] = s4.w;
    }
    /* ---- receiver back-injection */
    if ((EPI & EPI_INJECT) && gi >= a.inj_gi0 && gi < a.inj_gi0 + a.inj_n && a.inj_j >= j0 && a.inj_j < j0 + 4) {
        long long idx = a.dobs_base + (long long)(gi - a.inj_gi0) * a.inj_nt + lv.inj_tidx;
        float s = (idx >= 0 && idx < a.dobs_len) ? a.dobs[idx] : 0.0f;
        FDW_UNROLL
        for (int k = 0; k < 4; k++)
            if (j0 + k == a.inj_j) res[k] = fadd(res[k], s);
    }
    return make_float4(res[0], res[1], res[2], res[3]);
}

/* Side outputs of one row, all to global memory (aux pitch a.apitch): halo push, seismogram sample, forward
 * history, imaging.  c4 = the newer level's centre row (sponge applied as loaded), res = the new values,
 * f4 = the imaging operand of EPI_IMG_FIELD (the reconstructed source level at this point); zfc / xonc = the z
 * factors and the x-factor mask of this thread's own 4 columns. */
template <int EPI, bool TAPER>
FDW_HD void row_outputs(const StepArgs &a, const Level &lv, const float4 c4, const float4 res, const float4 f4, const int lr,
                        const int gi, const int j0, const float *zfc, const unsigned xonc)
{
    const long long ap = a.apitch;
    /* ---- halo push: boundary rows go to the neighbour's ghost rows over NVLink */
    if (EPI & EPI_PUSH) {
        if (lv.push_lo && lr < GUARD) st4(lv.push_lo + (long long)lr * ap + j0, res);
        if (lv.push_hi && lr >= a.push_nloc - GUARD) st4(lv.push_hi + (long long)lr * ap + j0, res);
    }
    /* ---- seismogram sample: the newer level after one more sponge pass */
    if ((EPI & EPI_RECORD) && gi >= a.rec_gi0 && gi < a.rec_gi0 + a.rec_n && a.rec_j >= j0 && a.rec_j < j0 + 4) {
        float4 s4 = c4;
        if (TAPER) s4 = tap4(s4, zfc, xonc, lv.tx[lr], gi < a.tz_ilim, 1);
        a.rec[(long long)(gi - a.rec_gi0) * a.rec_nt + lv.rec_it] = getk(s4, a.rec_j - j0);
    }
    /* ---- forward history: interior rows of the newer level (sponge factor there is 1) */
    if ((EPI & EPI_HSTORE) && gi >= a.hist_gi0 && gi < a.hist_gi0 + a.hist_n)
        st4(lv.hist_w + (long long)(gi - a.hist_gi0) * ap + j0, c4);
    /* ---- imaging condition */
    if ((EPI & EPI_IMG_HIST) && gi >= a.img_gi0 && gi < a.img_gi0 + a.img_n) {
        float *ip = a.img + (long long)(gi - a.img_gi0) * ap + j0;
        const float4 s4 = ld4_stream(lv.hist_r + (long long)(gi - a.hist_gi0) * ap + j0);
        float4 im = ld4(ip);
        im.x = fadd(im.x, fmul(s4.x, c4.x));
        im.y = fadd(im.y, fmul(s4.y, c4.y));
        im.z = fadd(im.z, fmul(s4.z, c4.z));
        im.w = fadd(im.w, fmul(s4.w, c4.w));
        st4(ip, im);
    }
    if ((EPI & EPI_IMG_FIELD) && gi >= a.img_gi0 && gi < a.img_gi0 + a.img_n) {
        float *ip = a.img + (long long)(gi - a.img_gi0) * ap + j0;
        float4 im = ld4(ip);
        im.x = fadd(im.x, fmul(f4.x, res.x));
        im.y = fadd(im.y, fmul(f4.y, res.y));
        im.z = fadd(im.z, fmul(f4.z, res.z));
        im.w = fadd(im.w, fmul(f4.w, res.w));
        st4(ip, im);
    }
}

/* per-thread invariants of the sponge: z factors of the NZ = 4*(2*NB+1) columns j0-4*NB .. j0+4*NB+3 (the thread's
 * own float4 and NB float4 blocks on each side), x-factor column mask */
template <int NZ>
FDW_HD bool sponge_setup(const StepArgs &a, const float *tz, int j0, float (&zf)[NZ], unsigned &xon)
{
    constexpr int C0 = (NZ - 4) / 2;
    xon = 0;
    bool zany = false; /* some z factor of these columns differs from 1 */
    FDW_UNROLL
    for (int m = 0; m < NZ; m++) {
        zf[m] = tz[j0 - C0 + m];
        zany = zany || zf[m] != 1.0f;
        if (j0 - C0 + m < a.tx_jlim) xon |= 1u << m;
    }
    return zany;
}

/* addresses are handled as 64-bit integers (byte addresses of global memory): the cursor arithmetic of
 * step_rows then compiles to plain integer adds from uniform registers */
template <bool TILE> struct Space;
template <> struct Space<false> {
    typedef unsigned long long addr;
    typedef long long diff;
#if defined(__CUDA_ARCH__) && defined(FDW_LDG_ASM)
    /* trial (tools/kbench5 -DFDW_LDG_ASM): explicit ld.global / st.global instead of the generic LD / ST the
     * integer-address formulation compiles to */
    static FDW_HDM float4 ld(addr x)
    {
        float4 v;
        asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(x));
        return v;
    }
    static FDW_HDM void st(addr x, float4 v)
    {
        asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(x), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
#else
    static FDW_HDM float4 ld(addr x) { return ld4((const float *)x); }
    static FDW_HDM void st(addr x, float4 v) { st4((float *)x, v); }
#endif
    static FDW_HDM float4 ld_stream(addr x) { return ld4_stream((const float *)x); }
};

/* rows [rb, re) of the float4 column at j0 (TILE is always false: the small-grid tile kernel has its own
 * point-per-thread body, fdw_tile_core.h) */
template <int ORDER, int RECIPE, bool TAPER, int EPI, bool PACKED, bool TILE>
FDW_HD void step_rows(const StepArgs &a, const Level &lv, const int j0, const int rb, const int re)
{
    constexpr int H = ORDER / 2, W = ORDER + 1, NB = zblocks(ORDER), C0 = 4 * NB;
    typedef Space<TILE> S;
    typedef typename S::addr addr;
    typedef typename S::diff diff;

    /* Per-thread invariants are folded into two rarely-true flags so that the streaming loop of
     * the ~99.9 % of threads that sit neither on the grid ring nor next to the source carries no
     * mask arithmetic (kept lean on purpose: at 64 registers ptxas rematerialises anything else
     * every row).  ring: some of this thread's 4 columns, or some of this CTA's rows, lie where
     * the Laplacian is defined as 0.  near_src: the source patch overlaps these 4 columns. */
    const bool ring = j0 < a.lap_j0 || j0 + 4 > a.lap_j1 || a.grow0 + rb < a.lap_i0 || a.grow0 + re > a.lap_i1;
    const bool near_src = lv.src_on && j0 + 3 >= a.src_j - a.src_rad && j0 <= a.src_j + a.src_rad;

    /* sponge on load.  A factor of exactly 1.0f changes nothing, so a thread whose 12 columns have no z
     * factor skips the multiplications of every row whose x factor is 1 as well: away from the sponge the
     * sponge instantiation then costs what the plain one does (it serves whole small grids). */
    float zf[4 * (2 * NB + 1)];
    unsigned xon = 0;
    bool zany = false;
    if (TAPER) zany = sponge_setup<4 * (2 * NB + 1)>(a, lv.tz, j0, zf, xon);

    /* ONE row cursor per thread (the centre row of the newer level); every other address of a row is the
     * cursor plus a launch-uniform byte delta, which costs an integer add from a uniform register instead
     * of a second and third loop-carried pointer pair (registers are what limits the resident warps). */
    const addr org_p = (addr)(size_t)lv.p, org_pp = (addr)(size_t)lv.pp, org_v = (addr)(size_t)lv.vdt,
               org_f = (addr)(size_t)lv.img_field;
    const diff rowb = (diff)(a.pitch * (long long)sizeof(float));
    const diff d_in = (diff)H * rowb;           /* centre row -> incoming row */
    const diff d_pp = (diff)(org_pp - org_p);   /* -> older level, same point */
    const diff d_v = (diff)(org_v - org_p);     /* -> v2*dt2 */
    const diff d_f = (diff)(org_f - org_p);     /* -> imaging operand (EPI_IMG_FIELD) */
    addr pc = org_p + (addr)((diff)(j0 * (int)sizeof(float)) + (diff)(rb - H) * rowb); /* prologue: row being loaded */

    float4 w[W];
    FDW_UNROLL
    for (int s = 0; s < 2 * H; s++) {
        w[s] = S::ld(pc);
        if (TAPER && lv.np) {
            const int lr = rb - H + s;
            const float xf = lv.tx[lr];
            const bool zon = a.grow0 + lr < a.tz_ilim;
            if ((zany && zon) || (xon && xf != 1.0f)) w[s] = tap4(w[s], zf + C0, xon >> C0, xf, zon, lv.np);
        }
        pc += (addr)rowb;
    }
    pc -= (addr)d_in; /* now the centre row of the first updated row */

    /* count-down loop: the only loop-carried integers are `left` and the row cursor */
    for (int left = re - rb; left > 0; left -= W) {
        FDW_UNROLL
        for (int u = 0; u < W; u++) {
            if (u < left) {
                const int lr = re - left + u; /* local row being updated (only rare paths use it) */
                const int gi = a.grow0 + lr;
                const addr ppc = pc + (addr)d_pp;
                float4 wn = S::ld(pc + (addr)d_in); /* row lr+H */
                float4 zl[NB], zr[NB];
                FDW_UNROLL
                for (int b = 0; b < NB; b++) {
                    zl[b] = S::ld(pc - (addr)(16 * (b + 1)));
                    zr[b] = S::ld(pc + (addr)(16 * (b + 1)));
                }
                float4 o4 = S::ld(ppc);
                const float4 v4 = S::ld_stream(pc + (addr)d_v);
                if (TAPER) {
                    const float xf = lv.tx[lr], xfn = lv.tx[lr + H];
                    const bool zon = gi < a.tz_ilim, zonn = gi + H < a.tz_ilim;
                    if ((zany && zonn) || (xon && xfn != 1.0f)) wn = tap4(wn, zf + C0, xon >> C0, xfn, zonn, lv.np);
                    if ((zany && zon) || (xon && xf != 1.0f)) {
                        FDW_UNROLL
                        for (int b = 0; b < NB; b++) {
                            zl[b] = tap4(zl[b], zf + C0 - 4 * (b + 1), xon >> (C0 - 4 * (b + 1)), xf, zon, lv.np);
                            zr[b] = tap4(zr[b], zf + C0 + 4 * (b + 1), xon >> (C0 + 4 * (b + 1)), xf, zon, lv.np);
                        }
                        o4 = tap4(o4, zf + C0, xon >> C0, xf, zon, lv.no);
                    }
                }
                w[(u + 2 * H) % W] = wn;
                const float4 res = row_update<ORDER, RECIPE, EPI, PACKED>(a, lv, w, u, zl, zr, o4, v4, gi, j0, ring, near_src);
                S::st(ppc, res);
                if (EPI) {
                    float4 f4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if ((EPI & EPI_IMG_FIELD) && gi >= a.img_gi0 && gi < a.img_gi0 + a.img_n)
                        f4 = S::ld_stream(pc + (addr)d_f);
                    row_outputs<EPI, TAPER>(a, lv, w[(u + H) % W], res, f4, lr, gi, j0, zf + C0, xon >> C0);
                }
                pc += (addr)rowb;
            }
        }
    }
}

template <int ORDER, int RECIPE, bool TAPER, int EPI, bool PACKED = true>
FDW_HD void step_thread(const StepArgs &a, const Level &lv, int bx, int by, int tid, int bdim)
{
    const int q = a.col4_0 + bx * bdim + tid;
    if (q >= a.ncol4) return;
    const int rb = a.row0 + by * a.rows_per_cta;
    const int re = rb + a.rows_per_cta < a.row1 ? rb + a.rows_per_cta : a.row1;
    if (rb >= re) return;
    step_rows<ORDER, RECIPE, TAPER, EPI, PACKED, false>(a, lv, q * 4, rb, re);
}

/* which float4 column and rows this thread owns: an ordinary 2-D launch over one rectangle, or CTA blockIdx.x
 * of a multi-rectangle launch (StepArgs::nrect > 0); false = nothing to do */
FDW_HD bool thread_work(const StepArgs &a, int bx, int by, int tid, int bdim, int *j0, int *rb, int *re)
{
    int c0 = a.col4_0, c1 = a.ncol4, r0 = a.row0, r1 = a.row1, rpc = a.rows_per_cta;
    if (a.nrect > 0) {
        int r = 0;
        while (r + 1 < a.nrect && bx >= a.rect[r + 1].cta0) r++;
        c0 = a.rect[r].c0; c1 = a.rect[r].c1; r0 = a.rect[r].r0; r1 = a.rect[r].r1; rpc = a.rect[r].rpc;
        const int b = bx - a.rect[r].cta0, nbx = a.rect[r].nbx, lw = a.rect[r].lw;
        bx = b % nbx;
        by = (b / nbx) * (bdim / lw) + tid / lw;
        tid %= lw;
        bdim = lw;
    }
    const int q = c0 + bx * bdim + tid;
    if (q >= c1) return false;
    *j0 = q * 4;
    *rb = r0 + by * rpc;
    *re = *rb + rpc < r1 ? *rb + rpc : r1;
    return *rb < *re;
}

/* step_thread for the kernels that may be launched over several rectangles at once (the sponge kernel) */
template <int ORDER, int RECIPE, bool TAPER, int EPI, bool PACKED = true>
FDW_HD void step_thread_rects(const StepArgs &a, const Level &lv, int bx, int by, int tid, int bdim)
{
    int j0, rb, re;
    if (thread_work(a, bx, by, tid, bdim, &j0, &rb, &re)) step_rows<ORDER, RECIPE, TAPER, EPI, PACKED, false>(a, lv, j0, rb, re);
}

/* arguments of the shared-memory tile kernels (small grids): the grid is cut into
 * ntx x nty tiles of tc4 float4 columns x tr rows, one CTA per tile, every tile resident in its
 * SM's shared memory for a whole propagation phase (see fdw_tile_core.h). */
struct TileArgs {
    PersistArgs pa;   /* pa.base.pitch = the tile's shared-memory pitch, pa.base.apitch = the global one */
    int tc4, tr, ntx, nty;
    int ch;           /* rows per thread */
    /* GPU-family backward (k_tile_back): the two saved levels of the source field, u(T) and u(T-1) */
    float *sav0, *sav1;
    unsigned *flags; /* one "levels completed" word per tile, 128 B apart, zeroed before the launch */
    /* flag-in-data halo exchange (fdw_tile_core.h): planes of {value, level tag} pairs in the global field layout,
     * two per field (level parity), zeroed before the launch; null = neighbour flags + plain ring loads */
    unsigned long long *ll;
    long long ll_plane; /* pairs per plane */
    int sp;             /* shared-memory row pitch of the tiles in floats */
    int dbg; /* 8: device-wide counter barrier instead of neighbour flags.  Timing experiments only (FDW_TILE_DBG; results are wrong when set): 1 no barrier, 2 no update, 4 no ring load */
};

/* arguments of the persistent slab kernel (thin slabs of a slab-decomposed run: a level is tens of microseconds,
 * so the per-level launches -- even replayed as a CUDA graph -- cost as much as the arithmetic): ONE cooperative
 * launch runs nlevels levels of the slab; per level the CTAs that own boundary rows first wait for the
 * neighbours' flags, update their rows and store them into the neighbours' ghost rows over NVLink (EPI_PUSH), the
 * last of them raises this slab's flag there; the other CTAs update the interior meanwhile; a device-wide barrier
 * closes the level. */
struct PSlabArgs {
    PersistArgs pa;          /* pa.base: the whole slab (col4_0..ncol4, row0..row1), EPI_PUSH bookkeeping fields unused */
    float *push_lo[2], *push_hi[2]; /* the neighbours' copies of the buffer written at even / odd levels (null: grid edge) */
    const unsigned *pw_flags; /* this slab's flag block (written by the neighbours) */
    unsigned *ps_lo, *ps_hi, *ps_count;
    int *pw_err;
    unsigned seq0;           /* flag value the neighbours reach before level 0 of this launch */
    int nbx;                 /* item columns (blockDim float4 columns each) */
    int rows_lo, rows_hi;    /* boundary rows with a neighbour below / above (0 or GUARD) */
    int rpc;                 /* rows per interior item */
    int nmid;                /* interior row chunks */
};

/* stand-alone Laplacian (config 1; kernel_lap fd-source-code.cu:110-135): the
 * exact reference sequence including the leading "0 +" adds; ring written 0.
 * Same streaming structure and the same lean loop as step_thread. */
template <int ORDER, bool PACKED = true>
FDW_HD void lap_thread(const StepArgs &a, float *lap, int bx, int by, int tid, int bdim)
{
    constexpr int H = ORDER / 2, W = ORDER + 1, NB = zblocks(ORDER), C0 = 4 * NB;
    const int q = a.col4_0 + bx * bdim + tid;
    if (q >= a.ncol4) return;
    const int j0 = q * 4;
    const int rb = a.row0 + by * a.rows_per_cta;
    const int re = rb + a.rows_per_cta < a.row1 ? rb + a.rows_per_cta : a.row1;
    if (rb >= re) return;
    const long long pitch = a.pitch;
    const bool ring = j0 < a.lap_j0 || j0 + 4 > a.lap_j1 || a.grow0 + rb < a.lap_i0 || a.grow0 + re > a.lap_i1;
    const float *__restrict__ pc = a.p + j0 + (long long)(rb - H) * pitch;
    float *__restrict__ out = lap + j0 + (long long)rb * pitch;
    float4 w[W];
    FDW_UNROLL
    for (int s = 0; s < 2 * H; s++) {
        w[s] = ld4(pc);
        pc += pitch;
    }
    for (int left = re - rb; left > 0; left -= W) {
        FDW_UNROLL
        for (int u = 0; u < W; u++) {
            if (u < left) {
                w[(u + 2 * H) % W] = ld4(pc);
                const float *ctr = pc - (long long)H * pitch;
                float4 zl[NB], zr[NB];
                FDW_UNROLL
                for (int b = 0; b < NB; b++) {
                    zl[b] = ld4(ctr - 4 * (b + 1));
                    zr[b] = ld4(ctr + 4 * (b + 1));
                }
                float za[4 * (2 * NB + 1)];
                z_row<NB>(zl, w[(u + H) % W], zr, za);
                float res[4];
                if (!PACKED) {
                    FDW_UNROLL
                    for (int k = 0; k < 4; k++) {
                        float az = 0.0f, ax = 0.0f;
                        FDW_UNROLL
                        for (int io = 0; io <= ORDER; io++) {
                            az = fadd(az, fmul(za[C0 + k - H + io], a.cz[io]));
                            ax = fadd(ax, fmul(getk(w[(u + io) % W], k), a.cx[io]));
                        }
                        res[k] = fadd(az, ax);
                    }
                } else {
                    /* packed pairs (0,1), (2,3); every lane runs the reference's scalar sequence,
                     * including the leading "0 +" adds (fd-source-code.cu:121-131) */
                    const f2 one = bc2(a.one);
                    constexpr int B = C0 - H;
                    f2 az0 = bc2(0.0f), az1 = bc2(0.0f), ax0 = bc2(0.0f), ax1 = bc2(0.0f);
                    FDW_UNROLL
                    for (int io = 0; io <= ORDER; io++) {
                        const float4 &wx = w[(u + io) % W];
                        az0 = add2(az0, one, zprod(za, B + io, a.cz[io]));
                        az1 = add2(az1, one, zprod(za, B + io + 2, a.cz[io]));
                        ax0 = add2(ax0, one, mul2(pk(wx.x, wx.y), bc2(a.cx[io])));
                        ax1 = add2(ax1, one, mul2(pk(wx.z, wx.w), bc2(a.cx[io])));
                    }
                    const f2 r01 = add2(az0, one, ax0), r23 = add2(az1, one, ax1);
                    res[0] = lo2(r01); res[1] = hi2(r01); res[2] = lo2(r23); res[3] = hi2(r23);
                }
                if (ring) {
                    const int gi = a.grow0 + re - left + u;
                    const bool row_in = gi >= a.lap_i0 && gi < a.lap_i1;
                    FDW_UNROLL
                    for (int k = 0; k < 4; k++)
                        if (!row_in || j0 + k < a.lap_j0 || j0 + k >= a.lap_j1) res[k] = 0.0f;
                }
                st4(out, make_float4(res[0], res[1], res[2], res[3]));
                pc += pitch;
                out += pitch;
            }
        }
    }
}

} /* namespace fdw */
#endif
