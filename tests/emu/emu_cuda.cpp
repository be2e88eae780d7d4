/* emu_cuda.cpp -- TEST INFRASTRUCTURE ONLY, see emu_cuda.h. */
#include "emu_cuda.h"

#include <stdint.h>
#include <vector>
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

/* ---- CUDA graphs (see emu_cuda.h) */
struct emu_node {
    bool kernel = false;
    void *func = nullptr;
    dim3 grid, block;
    std::vector<std::vector<unsigned char>> args;
    void set(const cudaKernelNodeParams *p)
    {
        func = p->func; grid = p->gridDim; block = p->blockDim;
        const size_t *sz = (const size_t *)p->extra;
        args.clear();
        for (int i = 0; sz && sz[i]; i++) {
            const unsigned char *src = (const unsigned char *)p->kernelParams[i];
            args.emplace_back(src, src + sz[i]);
        }
    }
};
struct emu_graph { std::vector<emu_node *> nodes; bool exec = false; };
long long emu_graph_launches = 0;

cudaError_t cudaGraphCreate(cudaGraph_t *g, unsigned) { *g = new emu_graph(); return cudaSuccess; }
cudaError_t cudaGraphDestroy(cudaGraph_t g)
{
    if (g) { for (emu_node *n : g->nodes) delete n; delete g; }
    return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; } /* shares the graph's nodes */
cudaError_t cudaGraphAddEmptyNode(cudaGraphNode_t *n, cudaGraph_t g, const cudaGraphNode_t *, size_t)
{
    *n = new emu_node();
    g->nodes.push_back(*n);
    return cudaSuccess;
}
cudaError_t cudaGraphAddKernelNode(cudaGraphNode_t *n, cudaGraph_t g, const cudaGraphNode_t *deps, size_t nd,
                                   const cudaKernelNodeParams *p)
{
    for (size_t i = 0; i < nd; i++) { /* every dependency must already exist: creation order is a topological order */
        bool found = false;
        for (emu_node *m : g->nodes) found = found || m == deps[i];
        if (!found) return 1;
    }
    *n = new emu_node();
    (*n)->kernel = true;
    (*n)->set(p);
    g->nodes.push_back(*n);
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, unsigned long long)
{
    *e = new emu_graph();
    (*e)->nodes = g->nodes; /* same node objects: SetParams on the exec updates what Launch runs */
    (*e)->exec = true;
    return cudaSuccess;
}
cudaError_t cudaGraphExecKernelNodeSetParams(cudaGraphExec_t, cudaGraphNode_t n, const cudaKernelNodeParams *p)
{
    if (!n->kernel || n->func != p->func || n->grid.x != p->gridDim.x || n->grid.y != p->gridDim.y ||
        n->block.x != p->blockDim.x)
        return 1; /* the real API rejects a change of kernel; be at least as strict */
    n->set(p);
    return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t st)
{
    for (emu_node *n : e->nodes) {
        if (!n->kernel) continue;
        std::vector<void *> ptrs;
        for (auto &a : n->args) ptrs.push_back(a.data());
        cudaLaunchKernel(n->func, n->grid, n->block, ptrs.data(), 0, st);
    }
    emu_graph_launches++;
    return cudaSuccess;
}
