// tools/kbench_sync.cu -- what does one device-wide synchronisation cost on a B200?  Input to the design of
// the small-grid tile kernel and the persistent slab kernel: per-iteration latency of
//   A  the round-1 counter barrier (fence + atomicAdd + volatile spin + fence, k_persist)
//   B  cooperative_groups grid.sync()
//   C  release/acquire counter barrier (red.release.gpu / ld.acquire.gpu, no stand-alone fences)
//   D  neighbour flags (each CTA publishes its level, waits for CTA-1 and CTA+1 only)
//   E  cluster barrier (8 CTAs), for scale
// for several grid sizes (all CTAs co-resident, cooperative launch).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void kA(unsigned *ctr, int iters, float *sink)
{
    const unsigned nblk = gridDim.x;
    float acc = 0;
    for (int l = 0; l < iters; l++) {
        acc += sink[(blockIdx.x * 32 + (threadIdx.x & 31) + l) & 1023];
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(ctr, 1u);
            while (*((volatile unsigned *)ctr) < (unsigned)(l + 1) * nblk) { }
            __threadfence();
        }
        __syncthreads();
    }
    if (acc == 12345.f) sink[0] = acc;
}

__global__ void kB(int iters, float *sink)
{
    cg::grid_group g = cg::this_grid();
    float acc = 0;
    for (int l = 0; l < iters; l++) {
        acc += sink[(blockIdx.x * 32 + (threadIdx.x & 31) + l) & 1023];
        g.sync();
    }
    if (acc == 12345.f) sink[0] = acc;
}

__device__ __forceinline__ void red_release(unsigned *p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void kC(unsigned *ctr, int iters, float *sink)
{
    const unsigned nblk = gridDim.x;
    float acc = 0;
    for (int l = 0; l < iters; l++) {
        acc += sink[(blockIdx.x * 32 + (threadIdx.x & 31) + l) & 1023];
        __syncthreads();
        if (threadIdx.x == 0) {
            red_release(ctr);
            while (ld_acquire(ctr) < (unsigned)(l + 1) * nblk) { }
        }
        __syncthreads();
    }
    if (acc == 12345.f) sink[0] = acc;
}

/* flags 128 bytes apart; CTA b waits for b-1 and b+1 (a 1-D chain like x slabs / tile rows) */
__global__ void kD(unsigned *flags, int iters, float *sink)
{
    const int b = blockIdx.x, n = gridDim.x;
    float acc = 0;
    for (int l = 0; l < iters; l++) {
        acc += sink[(blockIdx.x * 32 + (threadIdx.x & 31) + l) & 1023];
        __syncthreads();
        if (threadIdx.x == 0) st_release(flags + b * 32, (unsigned)(l + 1));
        if (threadIdx.x == 1 && b > 0) while (ld_acquire(flags + (b - 1) * 32) < (unsigned)(l + 1)) { }
        if (threadIdx.x == 2 && b < n - 1) while (ld_acquire(flags + (b + 1) * 32) < (unsigned)(l + 1)) { }
        __syncthreads();
    }
    if (acc == 12345.f) sink[0] = acc;
}

/* 2-D neighbour flags: CTA (bx,by) of a gx x gy tile grid waits for its 4 edge neighbours */
__global__ void kD2(unsigned *flags, int iters, int gx, float *sink)
{
    const int b = blockIdx.x, n = gridDim.x, bx = b % gx, by = b / gx, gy = n / gx;
    float acc = 0;
    const int t = threadIdx.x;
    int nb = -1;
    if (t == 1 && bx > 0) nb = b - 1;
    if (t == 2 && bx < gx - 1) nb = b + 1;
    if (t == 3 && by > 0) nb = b - gx;
    if (t == 4 && by < gy - 1) nb = b + gx;
    for (int l = 0; l < iters; l++) {
        acc += sink[(blockIdx.x * 32 + (threadIdx.x & 31) + l) & 1023];
        __syncthreads();
        if (t == 0) st_release(flags + b * 32, (unsigned)(l + 1));
        if (nb >= 0) while (ld_acquire(flags + nb * 32) < (unsigned)(l + 1)) { }
        __syncthreads();
    }
    if (acc == 12345.f) sink[0] = acc;
}

__global__ void __cluster_dims__(8, 1, 1) kE(int iters, float *sink)
{
    cg::cluster_group c = cg::this_cluster();
    float acc = 0;
    for (int l = 0; l < iters; l++) {
        acc += sink[(blockIdx.x * 32 + (threadIdx.x & 31) + l) & 1023];
        c.sync();
    }
    if (acc == 12345.f) sink[0] = acc;
}

int main()
{
    int iters = 4000;
    unsigned *ctr, *flags;
    float *sink;
    CK(cudaMalloc(&ctr, 4));
    CK(cudaMalloc(&flags, 1024 * 128));
    CK(cudaMalloc(&sink, 4096));
    CK(cudaMemset(sink, 0, 4096));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grids[] = {37, 74, 148, 296, 592};
    const int threads[] = {128, 256, 512};
    for (int th : threads)
        for (int g : grids) {
            if ((long long)g * th > 148LL * 2048) continue;
            float ms[6] = {0, 0, 0, 0, 0, 0};
            for (int which = 0; which < 6; which++) {
                for (int rep = 0; rep < 2; rep++) {
                    CK(cudaMemset(ctr, 0, 4));
                    CK(cudaMemset(flags, 0, 1024 * 128));
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    void *pa[] = {&ctr, &iters, &sink};
                    void *pb[] = {&iters, &sink};
                    void *pd[] = {&flags, &iters, &sink};
                    int gx = g >= 148 ? 4 : 2;
                    void *pd2[] = {&flags, &iters, &gx, &sink};
                    cudaError_t e = cudaSuccess;
                    if (which == 0) e = cudaLaunchCooperativeKernel((void *)kA, dim3(g), dim3(th), pa, 0, 0);
                    if (which == 1) e = cudaLaunchCooperativeKernel((void *)kB, dim3(g), dim3(th), pb, 0, 0);
                    if (which == 2) e = cudaLaunchCooperativeKernel((void *)kC, dim3(g), dim3(th), pa, 0, 0);
                    if (which == 3) e = cudaLaunchCooperativeKernel((void *)kD, dim3(g), dim3(th), pd, 0, 0);
                    if (which == 4) e = cudaLaunchCooperativeKernel((void *)kD2, dim3(g / gx * gx), dim3(th), pd2, 0, 0);
                    if (which == 5) { kE<<<g / 8 * 8, th>>>(iters, sink); e = cudaGetLastError(); }
                    if (e != cudaSuccess) { ms[which] = -1; (void)cudaGetLastError(); break; }
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    CK(cudaEventElapsedTime(&ms[which], e0, e1));
                }
            }
            printf("grid %4d x %3d thr: A counter+fences %7.0f ns | B cg grid.sync %7.0f ns | C release/acquire %7.0f ns | "
                   "D 1-D neighbour flags %7.0f ns | D2 2-D neighbour flags %7.0f ns | E cluster(8).sync %7.0f ns\n",
                   g, th, ms[0] * 1e6 / iters, ms[1] * 1e6 / iters, ms[2] * 1e6 / iters, ms[3] * 1e6 / iters,
                   ms[4] * 1e6 / iters, ms[5] * 1e6 / iters);
            fflush(stdout);
        }
    return 0;
}
