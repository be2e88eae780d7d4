/* emu_cuda.cpp -- TEST INFRASTRUCTURE ONLY, see emu_cuda.h. */
#include "emu_cuda.h"

#include <stdint.h>
#include <time.h>

thread_local uint3_emu blockIdx, threadIdx;
thread_local dim3 blockDim, gridDim;
int emu_level = 0, emu_levels = 0;

cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    const char *e = getenv("FDW_EMU_NSM");
    p->multiProcessorCount = e ? atoi(e) : 3; /* tiny: forces several x chunks even on small grids */
    return cudaSuccess;
}
const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xFF, n); /* NaN poison */
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < h; r++) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)(uintptr_t)0x1; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = calloc(1, sizeof(double)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    *(double *)e = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(*(double *)b - *(double *)a); return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t) { *n = 2; return cudaSuccess; }

cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t, cudaStream_t)
{
    void (*thunk)(void **) = (void (*)(void **))func;
    const long long nblk = (long long)grid.x * grid.y * grid.z;
#pragma omp parallel for schedule(dynamic)
    for (long long b = 0; b < nblk; b++) {
        gridDim = grid;
        blockDim = block;
        blockIdx.x = (unsigned)(b % grid.x);
        blockIdx.y = (unsigned)((b / grid.x) % grid.y);
        blockIdx.z = (unsigned)(b / ((long long)grid.x * grid.y));
        for (unsigned tz = 0; tz < block.z; tz++)
            for (unsigned ty = 0; ty < block.y; ty++)
                for (unsigned tx = 0; tx < block.x; tx++) {
                    threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
                    thunk(args);
                }
    }
    return cudaSuccess;
}

cudaError_t cudaLaunchCooperativeKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t st)
{
    for (int l = 0; l < emu_levels; l++) {
        emu_level = l;
        cudaLaunchKernel(func, grid, block, args, smem, st);
    }
    emu_level = 0;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *value, int, int) { *value = 1; return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
