/*
 * emu_cuda.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A host stand-in for the handful of CUDA runtime calls libfdwave makes, so
 * that the library's HOST logic (context set-up, sponge bookkeeping, launch
 * geometry, epilogue arguments, pipelines) and the per-thread kernel bodies
 * (fdw_step_core.h compiles for the host) can be unit-tested on a machine
 * without a GPU (`pytest -m "not gpu"`).  It builds tests/emu/libfdwave_emu.so;
 * the product package never loads that file -- parallel_finite_difference_
 * computation_b200/_lib.py loads libfdwave.so only and fails loudly when a
 * CUDA device is missing.  Device allocations are poisoned with NaNs to catch
 * reliance on zero-initialised memory.
 */
#ifndef EMU_CUDA_H
#define EMU_CUDA_H
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };
extern thread_local uint3_emu blockIdx, threadIdx;
extern thread_local dim3 blockDim, gridDim;
extern int emu_level;  /* level index seen by the persistent kernel's host stand-in */
extern int emu_levels; /* set by the caller before cudaLaunchCooperativeKernel */
struct cudaDeviceProp { int multiProcessorCount; };

#define __global__
#define __device__
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#define __grid_constant__
#define __launch_bounds__(...)

static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }

cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *, int);
const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError(void);
cudaError_t cudaMalloc(void **, size_t);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
cudaError_t cudaFree(void *);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t);
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k) { return cudaMemcpyAsync(d, s, n, k, (cudaStream_t)0); }
cudaError_t cudaMemcpy2DAsync(void *, size_t, const void *, size_t, size_t, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *, const void *, int, size_t);
/* `func` is a thunk void(*)(void **args) executed once per emulated thread */
cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t);
/* runs the thunk emu_levels times, level by level (the grid barrier of the real kernel) */
cudaError_t cudaLaunchCooperativeKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t);
enum { cudaDevAttrCooperativeLaunch = 95, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline cudaError_t cudaFuncSetAttribute(const void *, int, int) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *value, int attr, int device);
/* Acquire side of a flag in the host stand-in.  Default: kernels run synchronously in lock-step tests, so an
 * unmet flag is a protocol error (returns false at once).  With FDW_EMU_SPIN=1 the slabs run in separate host
 * threads, like GPUs, and the wait really waits (bounded: ~10 s). */
#include <sched.h>
#include <time.h>
static inline bool emu_wait_flag(const volatile unsigned *f, unsigned v)
{
    if (*f >= v) { __sync_synchronize(); return true; }
    const char *e = getenv("FDW_EMU_SPIN");
    if (!e || e[0] != '1') return false;
    struct timespec t0, t;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    while (*f < v) {
        sched_yield();
        clock_gettime(CLOCK_MONOTONIC, &t);
        if (t.tv_sec - t0.tv_sec > 10) return false;
    }
    __sync_synchronize();
    return true;
}

/* CUDA graphs: kernel nodes are kept in creation order (the library creates them in a valid topological
 * order) and a launch runs them one after the other.  Argument values are copied at node creation /
 * update; their sizes arrive through cudaKernelNodeParams::extra (see node_params in fdw_api.cu). */
struct cudaKernelNodeParams {
    void *func; dim3 gridDim, blockDim; unsigned sharedMemBytes; void **kernelParams; void **extra;
};
struct emu_graph;
struct emu_node;
typedef emu_graph *cudaGraph_t;
typedef emu_graph *cudaGraphExec_t;
typedef emu_node *cudaGraphNode_t;
cudaError_t cudaGraphCreate(cudaGraph_t *, unsigned);
cudaError_t cudaGraphDestroy(cudaGraph_t);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t);
cudaError_t cudaGraphAddEmptyNode(cudaGraphNode_t *, cudaGraph_t, const cudaGraphNode_t *, size_t);
cudaError_t cudaGraphAddKernelNode(cudaGraphNode_t *, cudaGraph_t, const cudaGraphNode_t *, size_t, const cudaKernelNodeParams *);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long);
cudaError_t cudaGraphExecKernelNodeSetParams(cudaGraphExec_t, cudaGraphNode_t, const cudaKernelNodeParams *);
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t);
extern long long emu_graph_launches; /* for the tests: how many graph launches have run */

/* "IPC" inside one process: the handle carries the pointer itself (two slab contexts of the
 * same test process can then push into each other's buffers) */
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *);
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned);
cudaError_t cudaIpcCloseMemHandle(void *);
#endif
