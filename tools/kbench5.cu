// tools/kbench5.cu -- round-2 A/B of the shipped step kernel body (csrc/fdw_step_core.h) on a large grid:
// scalar FP32 arithmetic (round 1) vs packed FP32x2 (FMUL2/FFMA2), at register budgets 64 / 72 / 80,
// for the plain kernels and the epilogue variants that spilled in round 1.  Every variant is checked
// bitwise against the scalar 64-register kernel, timed with CUDA events over REPS ping-pong launches.
// Build (tools/Makefile): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false \
//        -I../parallel_finite_difference_computation_b200/csrc -I../include -o kbench5 kbench5.cu
// Run:   ./kbench5 [nx nz reps]      (defaults 16384 16384 20)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fdw_step_core.h"

using namespace fdw;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int RECIPE, int EPI, bool PACKED, int MAXR>
__global__ void __maxnreg__(MAXR) k5(const __grid_constant__ StepArgs a)
{
    step_thread<8, RECIPE, false, EPI, PACKED>(a, level_of(a), blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}

template <bool PACKED, int MAXR>
__global__ void __maxnreg__(MAXR) k5lap(const __grid_constant__ StepArgs a, float *lap)
{
    lap_thread<8, PACKED>(a, lap, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
}

/* ---- round-2 trial: operands prefetched K rows ahead with cp.async (LDGSTS) into a per-warp shared-memory ring.
 * Bytes in flight no longer occupy registers: at 32 warps/SM the register kernel keeps ~49 KB/SM in flight, the
 * 3R+1W stream ceiling (7.2 TB/s, below) needs more.  One warp = one independent unit: lane t copies ITS float4
 * column of the incoming p row, of pp and of v2*dt2 for row r+K while row r is computed; the centre row's left /
 * right neighbours are read from the neighbour lanes' ring slots (lanes 0 / 31: from global, L2).  The x window
 * stays in registers.  Arithmetic = row_update of the shipped header, i.e. bit-identical. */
__device__ __forceinline__ void cpa_ca(const float4 *s, const float *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cpa_cg(const float4 *s, const float *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(s)), "l"(g) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int RECIPE, bool PACKED, int K, int MAXR>
__global__ void __maxnreg__(MAXR) k5pf(const __grid_constant__ StepArgs a)
{
    constexpr int H = 4, W = 9, RP = H + K;
    extern __shared__ float4 sm4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *sp = sm4 + warp * (RP + 2 * K) * 32; /* [slot][lane] */
    float4 *so = sp + RP * 32, *sv = so + K * 32;
    const int q = a.col4_0 + blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = q < a.ncol4;
    const int j0 = 4 * q;
    const int rb = a.row0 + blockIdx.y * a.rows_per_cta;
    const int re = rb + a.rows_per_cta < a.row1 ? rb + a.rows_per_cta : a.row1;
    if (rb >= re) return;
    const long long pitch = a.pitch;
    const Level lv = level_of(a);
    const bool ring = j0 < a.lap_j0 || j0 + 4 > a.lap_j1 || a.grow0 + rb < a.lap_i0 || a.grow0 + re > a.lap_i1;
    const bool near_src = lv.src_on && j0 + 3 >= a.src_j - a.src_rad && j0 <= a.src_j + a.src_rad;
    const float *gp = a.p + j0;
    float *gpp = a.pp + j0;
    const float *gv = a.vdt + j0;

    float4 w[W];
    int sc = 0;  /* ring slot of the centre row r */
    int so_i = 0; /* pp / vdt slot of row r */
    /* prologue: rows rb-H..rb-1 straight into registers; rows rb..rb+H-1 through the ring (their neighbours need them) */
    if (act) {
#pragma unroll
        for (int s2 = 0; s2 < H; s2++) w[s2] = ld4(gp + (long long)(rb - H + s2) * pitch);
#pragma unroll
        for (int s2 = 0; s2 < H; s2++) cpa_ca(&sp[s2 * 32 + lane], gp + (long long)(rb + s2) * pitch);
    }
    cpa_commit();
#pragma unroll
    for (int i = 0; i < K; i++) {
        if (act && rb + i < re) {
            cpa_ca(&sp[(H + i) * 32 + lane], gp + (long long)(rb + i + H) * pitch);
            cpa_cg(&so[i * 32 + lane], gpp + (long long)(rb + i) * pitch);
            cpa_cg(&sv[i * 32 + lane], gv + (long long)(rb + i) * pitch);
        }
        cpa_commit();
    }
    cpa_wait<K>();
    __syncwarp();
    if (act) {
#pragma unroll
        for (int s2 = 0; s2 < H; s2++) w[H + s2] = sp[s2 * 32 + lane];
    }
    for (int left = re - rb; left > 0; left -= W) {
#pragma unroll
        for (int u = 0; u < W; u++) {
            if (u < left) {
                const int r = re - left + u;
                const int gi = a.grow0 + r;
                cpa_wait<K - 1>();
                __syncwarp();
                int sin = sc + H; /* slot of the incoming row r+H */
                if (sin >= RP) sin -= RP;
                if (act) {
                    const float4 wn = sp[sin * 32 + lane];
                    const float4 o4 = so[so_i * 32 + lane], v4 = sv[so_i * 32 + lane];
                    const float *ctr = gp + (long long)r * pitch;
                    const float4 l4 = lane > 0 ? sp[sc * 32 + lane - 1] : ld4(ctr - 4);
                    const float4 r4 = lane < 31 ? sp[sc * 32 + lane + 1] : ld4(ctr + 4);
                    w[(u + 2 * H) % W] = wn;
                    const float4 res = row_update<8, RECIPE, 0, PACKED>(a, lv, w, u, l4, r4, o4, v4, gi, j0, ring, near_src);
                    st4(gpp + (long long)r * pitch, res);
                }
                __syncwarp(); /* every lane has read the centre row's slot before it is recycled */
                if (act && r + K < re) {
                    cpa_ca(&sp[sc * 32 + lane], gp + (long long)(r + K + H) * pitch);
                    cpa_cg(&so[so_i * 32 + lane], gpp + (long long)(r + K) * pitch);
                    cpa_cg(&sv[so_i * 32 + lane], gv + (long long)(r + K) * pitch);
                }
                cpa_commit();
                if (++sc == RP) sc = 0;
                if (++so_i == K) so_i = 0;
            }
        }
    }
}

/* access-pattern ceilings: the step kernel's streams without its arithmetic, halo or x window.
 * MODE 0: 3 reads + 1 write per float4 (p, pp, vdt -> pp), MODE 1: 3 reads only, MODE 2: 1 read + 1 write (copy) */
template <int MODE, int ILP>
__global__ void __maxnreg__(64) k5stream(const __grid_constant__ StepArgs a)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.ncol4) return;
    const int rb = a.row0 + blockIdx.y * a.rows_per_cta;
    const int re = rb + a.rows_per_cta < a.row1 ? rb + a.rows_per_cta : a.row1;
    const float *pc = a.p + 4 * q + (long long)rb * a.pitch;
    float *ppc = a.pp + 4 * q + (long long)rb * a.pitch;
    const float *vc = a.vdt + 4 * q + (long long)rb * a.pitch;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int r = rb; r < re; r += ILP) {
        float4 x[ILP], y[ILP], z[ILP];
#pragma unroll
        for (int i = 0; i < ILP; i++)
            if (r + i < re) {
                x[i] = ld4(pc + i * a.pitch);
                if (MODE != 2) { y[i] = ld4(ppc + i * a.pitch); z[i] = ld4_stream(vc + i * a.pitch); }
            }
#pragma unroll
        for (int i = 0; i < ILP; i++)
            if (r + i < re) {
                float4 o = x[i];
                if (MODE != 2) { o.x = fmaf(y[i].x, z[i].x, o.x); o.y = fmaf(y[i].y, z[i].y, o.y); o.z = fmaf(y[i].z, z[i].z, o.z); o.w = fmaf(y[i].w, z[i].w, o.w); }
                if (MODE == 1) { acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
                else st4(ppc + i * a.pitch, o);
            }
        pc += ILP * a.pitch; ppc += ILP * a.pitch; vc += ILP * a.pitch;
    }
    if (MODE == 1 && acc.x + acc.y + acc.z + acc.w == 1.2345e30f) st4(ppc, acc);
}

struct Variant { const char *name; const void *fn; int recipe, epi, packed, maxr; };

#define V(R, E, P, M) {#R "/" #E "/" #P "/" #M, (const void *)k5<R, E, P, M>, R, E, P, M}
static Variant variants[] = {
    V(0, 0, false, 64), V(0, 0, true, 64), V(0, 0, true, 72), V(0, 0, true, 80), V(0, 0, false, 72), V(0, 0, false, 80),
    V(1, 0, false, 64), V(1, 0, true, 64), V(1, 0, true, 72), V(1, 0, true, 80),
    V(1, 4, false, 64), V(1, 4, true, 64), V(1, 4, true, 72), V(1, 4, true, 80),
    V(1, 10, false, 64), V(1, 10, true, 64), V(1, 10, true, 72), V(1, 10, true, 80),
    V(1, 1, false, 64), V(1, 1, true, 64), V(1, 1, true, 72), V(1, 1, true, 80),
    V(0, 18, false, 64), V(0, 18, true, 64), V(0, 18, true, 72), V(0, 18, true, 80), V(0, 18, false, 72),
};

int main(int argc, char **argv)
{
    const int nxe = argc > 1 ? atoi(argv[1]) : 16384, nze = argc > 2 ? atoi(argv[2]) : 16384;
    const int reps = argc > 3 ? atoi(argv[3]) : 20;
    const int nthreads = argc > 4 ? atoi(argv[4]) : 32, rpc = argc > 5 ? atoi(argv[5]) : 7;
    const int nb = 40, nt = 64;
    const long long pitch = ((long long)nze + 4 + 31) / 32 * 32;
    const size_t rows = (size_t)nxe + 2 * GUARD + 2, elems = rows * pitch;
    const int nx = nxe - 2 * nb;
    printf("# grid %d x %d, pitch %lld, %d reps, CTA %d threads x %d rows\n", nxe, nze, pitch, reps, nthreads, rpc);

    float *f[2], *vdt, *ref[2], *img, *imgref, *hist, *rec, *dobs;
    for (int k = 0; k < 2; k++) { CK(cudaMalloc(&f[k], elems * 4)); CK(cudaMalloc(&ref[k], elems * 4)); }
    CK(cudaMalloc(&vdt, elems * 4));
    CK(cudaMalloc(&img, (size_t)nx * pitch * 4));
    CK(cudaMalloc(&imgref, (size_t)nx * pitch * 4));
    const int nhist = 4; /* history slices cycled through (each nx x pitch) */
    CK(cudaMalloc(&hist, (size_t)nhist * nx * pitch * 4));
    CK(cudaMalloc(&rec, (size_t)nx * nt * 4));
    CK(cudaMalloc(&dobs, (size_t)nx * nt * 4));
    std::vector<float> h(elems, 0.0f), hv(elems, 0.0f);
    std::vector<float> init[2];
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (int k = 0; k < 2; k++) {
        init[k].assign(elems, 0.0f);
        for (int i = 0; i < nxe; i++)
            for (int j = 0; j < nze; j++) init[k][(size_t)(i + GUARD + 1) * pitch + j] = rnd();
    }
    for (int i = 0; i < nxe; i++)
        for (int j = 0; j < nze; j++) {
            float v = j < nze / 3 ? 2000.f : (j < 2 * nze / 3 ? 3000.f : 4000.f);
            hv[(size_t)(i + GUARD + 1) * pitch + j] = (v * v) * (0.001f * 0.001f);
        }
    CK(cudaMemcpy(vdt, hv.data(), elems * 4, cudaMemcpyHostToDevice));
    {
        std::vector<float> hh((size_t)nhist * nx * pitch), hd((size_t)nx * nt);
        for (auto &x : hh) x = rnd();
        for (auto &x : hd) x = rnd();
        CK(cudaMemcpy(hist, hh.data(), hh.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dobs, hd.data(), hd.size() * 4, cudaMemcpyHostToDevice));
    }

    StepArgs a;
    memset(&a, 0, sizeof a);
    a.pitch = pitch; a.apitch = pitch; a.col4_0 = 0; a.ncol4 = (nze + 3) / 4; a.row0 = 0; a.row1 = nxe; a.rows_per_cta = rpc;
    a.grow0 = 0; a.lap_i0 = 4; a.lap_i1 = nxe - 4; a.lap_j0 = 4; a.lap_j1 = nze - 4; a.nze = nze;
    const double c8[9] = {-1. / 560, 8. / 315, -1. / 5, 8. / 5, -205. / 72, 8. / 5, -1. / 5, 8. / 315, -1. / 560};
    const float d2 = (float)((1. / 10.) * (1. / 10.));
    a.one = 1.0f;
    a.src_on = 1; a.src_gi = nxe / 2; a.src_j = nb; a.src_rad = 0; a.src_amp = 0.5f;
    a.rec = rec; a.rec_gi0 = nb; a.rec_n = nx; a.rec_j = nb; a.rec_nt = nt;
    a.dobs = dobs; a.dobs_base = 0; a.dobs_len = (long long)nx * nt; a.inj_gi0 = nb; a.inj_n = nx; a.inj_j = nb; a.inj_nt = nt;
    a.hist_gi0 = nb; a.hist_n = nx; a.img_gi0 = nb; a.img_n = nx;

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dim3 block(nthreads), grid((a.ncol4 + nthreads - 1) / nthreads, (nxe + rpc - 1) / rpc);
    double base_ms[32] = {0};
    for (const Variant &v : variants) {
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, v.fn));
        for (int k = 0; k <= 8; k++) {
            float c = (float)c8[k];
            a.cz[k] = v.recipe == 1 ? c : d2 * c;
            a.cx[k] = v.recipe == 1 ? c : d2 * c;
        }
        a.dz2inv = a.dx2inv = d2;
        const bool is_ref = !v.packed && v.maxr == 64;
        float *g[2] = {is_ref ? ref[0] : f[0], is_ref ? ref[1] : f[1]};
        float *im = is_ref ? imgref : img;
        float best = 1e30f;
        for (int trial = 0; trial < 3; trial++) {
            for (int k = 0; k < 2; k++) CK(cudaMemcpy(g[k], init[k].data(), elems * 4, cudaMemcpyHostToDevice));
            CK(cudaMemset(im, 0, (size_t)nx * pitch * 4));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int it = 0; it < reps; it++) {
                a.p = g[it & 1] + (size_t)(GUARD + 1) * pitch;
                a.pp = g[(it + 1) & 1] + (size_t)(GUARD + 1) * pitch;
                a.vdt = vdt + (size_t)(GUARD + 1) * pitch;
                a.rec_it = it % nt; a.inj_tidx = nt - 1 - (it % nt);
                a.hist_w = hist + (size_t)(it % nhist) * nx * pitch;
                a.hist_r = hist + (size_t)((it + 1) % nhist) * nx * pitch;
                a.img = im; a.img_field = vdt + (size_t)(GUARD + 1) * pitch;
                void *params[] = {&a};
                CK(cudaLaunchKernel(v.fn, grid, block, params, 0, 0));
            }
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        const double per = best / reps;
        const double gpts = (double)nxe * nze / (per * 1e-3) / 1e9;
        /* bitwise check against the scalar 64-register kernel of the same recipe/epilogue */
        long long bad = -1;
        if (!is_ref) {
            std::vector<float> x(elems), y(elems);
            bad = 0;
            for (int k = 0; k < 2; k++) {
                CK(cudaMemcpy(x.data(), g[k], elems * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(y.data(), ref[k], elems * 4, cudaMemcpyDeviceToHost));
                bad += memcmp(x.data(), y.data(), elems * 4) != 0;
            }
            if (v.epi & (EPI_IMG_HIST | EPI_IMG_FIELD)) {
                std::vector<float> xi((size_t)nx * pitch), yi((size_t)nx * pitch);
                CK(cudaMemcpy(xi.data(), img, xi.size() * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(yi.data(), imgref, yi.size() * 4, cudaMemcpyDeviceToHost));
                bad += memcmp(xi.data(), yi.data(), xi.size() * 4) != 0;
            }
        }
        printf("recipe %d epi %2d %-6s maxr %d: regs %3d spill-local %4zu B  %8.4f ms/level  %7.1f Gpts/s  %s\n", v.recipe, v.epi,
               v.packed ? "packed" : "scalar", v.maxr, fa.numRegs, fa.localSizeBytes, per, gpts,
               is_ref ? "(reference)" : (bad == 0 ? "bitwise OK" : "MISMATCH"));
        fflush(stdout);
    }
    /* ---- cp.async prefetch trial (recipe G plain and recipe C plain), vs ref[] of the matching recipe */
    {
        struct PF { const char *n; const void *fn; int recipe, k; } pf[] = {
            {"pf G scalar K2 r64", (const void *)k5pf<0, false, 2, 64>, 0, 2}, {"pf G scalar K3 r64", (const void *)k5pf<0, false, 3, 64>, 0, 3},
            {"pf G packed K3 r64", (const void *)k5pf<0, true, 3, 64>, 0, 3}, {"pf G scalar K3 r72", (const void *)k5pf<0, false, 3, 72>, 0, 3},
            {"pf G scalar K4 r80", (const void *)k5pf<0, false, 4, 80>, 0, 4}, {"pf G packed K4 r80", (const void *)k5pf<0, true, 4, 80>, 0, 4},
            {"pf C packed K3 r64", (const void *)k5pf<1, true, 3, 64>, 1, 3}, {"pf C packed K4 r80", (const void *)k5pf<1, true, 4, 80>, 1, 4}};
        const int geo[][2] = {{32, 7}, {32, 14}, {32, 28}, {32, 64}, {64, 28}, {128, 28}};
        std::vector<float> x(elems), y(elems);
        int last_recipe = -1;
        for (auto &v : pf) {
            for (int k = 0; k <= 8; k++) { float c = (float)c8[k]; a.cz[k] = v.recipe == 1 ? c : d2 * c; a.cx[k] = a.cz[k]; }
            if (v.recipe != last_recipe) { /* reference result of this recipe: scalar 64-register kernel, shipped geometry */
                const void *rf = v.recipe == 0 ? (const void *)k5<0, 0, false, 64> : (const void *)k5<1, 0, false, 64>;
                for (int k = 0; k < 2; k++) CK(cudaMemcpy(ref[k], init[k].data(), elems * 4, cudaMemcpyHostToDevice));
                a.rows_per_cta = rpc;
                for (int it = 0; it < reps; it++) {
                    a.p = ref[it & 1] + (size_t)(GUARD + 1) * pitch; a.pp = ref[(it + 1) & 1] + (size_t)(GUARD + 1) * pitch;
                    a.vdt = vdt + (size_t)(GUARD + 1) * pitch;
                    void *params[] = {&a};
                    CK(cudaLaunchKernel(rf, grid, block, params, 0, 0));
                }
                CK(cudaDeviceSynchronize());
                last_recipe = v.recipe;
            }
            CK(cudaFuncSetAttribute(v.fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            cudaFuncAttributes fa;
            CK(cudaFuncGetAttributes(&fa, v.fn));
            for (auto &gg : geo) {
                a.rows_per_cta = gg[1];
                dim3 b2(gg[0]), g2((a.ncol4 + gg[0] - 1) / gg[0], (nxe + gg[1] - 1) / gg[1]);
                const size_t smem = (size_t)(gg[0] / 32) * (4 + 3 * v.k) * 32 * 16;
                int occ = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, gg[0], smem));
                float best = 1e30f;
                for (int trial = 0; trial < 3; trial++) {
                    for (int k = 0; k < 2; k++) CK(cudaMemcpy(f[k], init[k].data(), elems * 4, cudaMemcpyHostToDevice));
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    for (int it = 0; it < reps; it++) {
                        a.p = f[it & 1] + (size_t)(GUARD + 1) * pitch; a.pp = f[(it + 1) & 1] + (size_t)(GUARD + 1) * pitch;
                        a.vdt = vdt + (size_t)(GUARD + 1) * pitch;
                        void *params[] = {&a};
                        CK(cudaLaunchKernel(v.fn, g2, b2, params, smem, 0));
                    }
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                int bad = 0;
                for (int k = 0; k < 2; k++) {
                    CK(cudaMemcpy(x.data(), f[k], elems * 4, cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(y.data(), ref[k], elems * 4, cudaMemcpyDeviceToHost));
                    bad += memcmp(x.data(), y.data(), elems * 4) != 0;
                }
                const double per = best / reps;
                printf("%-20s CTA %3d x %2d rows: regs %3d spill %3zu B, %2d warps/SM  %8.4f ms/level  %7.1f Gpts/s  %s\n", v.n, gg[0], gg[1],
                       fa.numRegs, fa.localSizeBytes, occ * gg[0] / 32, per, (double)nxe * nze / (per * 1e-3) / 1e9, bad ? "MISMATCH" : "bitwise OK");
                fflush(stdout);
            }
        }
        a.rows_per_cta = rpc;
    }
    /* ---- stand-alone Laplacian (config 1, 8 B/point): scalar vs packed */
    {
        struct { const char *n; const void *fn; } lv[] = {{"lap scalar 64", (const void *)k5lap<false, 64>}, {"lap packed 64", (const void *)k5lap<true, 64>},
                                                          {"lap packed 72", (const void *)k5lap<true, 72>}, {"lap scalar 72", (const void *)k5lap<false, 72>}};
        for (int k = 0; k <= 8; k++) { a.cz[k] = d2 * (float)c8[k]; a.cx[k] = d2 * (float)c8[k]; }
        std::vector<float> x(elems), y(elems);
        for (int vi = 0; vi < 4; vi++) {
            CK(cudaMemcpy(f[0], init[0].data(), elems * 4, cudaMemcpyHostToDevice));
            a.p = f[0] + (size_t)(GUARD + 1) * pitch;
            float *out = (vi == 0 ? ref[1] : f[1]) + (size_t)(GUARD + 1) * pitch;
            void *params[] = {&a, &out};
            float best = 1e30f;
            for (int trial = 0; trial < 3; trial++) {
                CK(cudaEventRecord(e0));
                for (int it = 0; it < reps; it++) CK(cudaLaunchKernel(lv[vi].fn, grid, block, params, 0, 0));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            int bad = 0;
            if (vi) {
                CK(cudaMemcpy(x.data(), f[1], elems * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(y.data(), ref[1], elems * 4, cudaMemcpyDeviceToHost));
                bad = memcmp(x.data(), y.data(), elems * 4) != 0;
            }
            const double per = best / reps;
            printf("%s: %8.4f ms/sweep  %7.1f Gpts/s  %6.1f GB/s at 8 B/pt  %s\n", lv[vi].n, per, (double)nxe * nze / (per * 1e-3) / 1e9,
                   8.0 * nxe * nze / (per * 1e-3) / 1e9, vi == 0 ? "(reference)" : (bad ? "MISMATCH" : "bitwise OK"));
        }
    }
    /* ---- access-pattern ceilings */
    {
        struct { const char *n; const void *fn; double bytes; } sv[] = {
            {"stream 3R+1W ilp1", (const void *)k5stream<0, 1>, 16}, {"stream 3R+1W ilp2", (const void *)k5stream<0, 2>, 16},
            {"stream 3R+1W ilp4", (const void *)k5stream<0, 4>, 16}, {"stream 3R ilp2", (const void *)k5stream<1, 2>, 12},
            {"stream 3R ilp4", (const void *)k5stream<1, 4>, 12}, {"copy 1R+1W ilp2", (const void *)k5stream<2, 2>, 8},
            {"copy 1R+1W ilp4", (const void *)k5stream<2, 4>, 8}};
        const int geo[][2] = {{32, 7}, {32, 16}, {128, 8}, {256, 32}};
        a.p = f[0] + (size_t)(GUARD + 1) * pitch; a.pp = f[1] + (size_t)(GUARD + 1) * pitch; a.vdt = vdt + (size_t)(GUARD + 1) * pitch;
        for (auto &sk : sv)
            for (auto &gg : geo) {
                a.rows_per_cta = gg[1];
                dim3 b2(gg[0]), g2((a.ncol4 + gg[0] - 1) / gg[0], (nxe + gg[1] - 1) / gg[1]);
                void *params[] = {&a};
                float best = 1e30f;
                for (int trial = 0; trial < 3; trial++) {
                    CK(cudaEventRecord(e0));
                    for (int it = 0; it < reps; it++) CK(cudaLaunchKernel(sk.fn, g2, b2, params, 0, 0));
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                const double per = best / reps;
                printf("%-18s CTA %3d x %2d rows: %8.4f ms  %7.1f GB/s\n", sk.n, gg[0], gg[1], per, sk.bytes * nxe * nze / (per * 1e-3) / 1e9);
            }
    }
    return 0;
}
