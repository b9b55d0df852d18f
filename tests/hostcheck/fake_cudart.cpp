// TEST INFRASTRUCTURE ONLY -- a host-memory stand-in for the few CUDA runtime entry points that
// dorylus_b200/csrc/engine.cu calls, so that the PRODUCT object file of the engine (the nvcc-compiled
// dorylus_b200/_obj/engine_cu.o, unchanged) can be linked and executed on a machine without a GPU.
// "Device" memory is malloc, copies are memcpy, streams and events are dummies, work is synchronous.
// Linked only into tests/hostcheck/_build/libdorylus_hostcheck.so (see build.py); nothing in the product
// path can load it: dorylus_b200/_lib.py opens dorylus_b200/libdorylus_b200.so by absolute path.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace {
thread_local dim3 g_grid, g_block;
thread_local size_t g_shmem = 0;
thread_local void *g_stream = nullptr;
int g_dummy_handles = 0;
}  // namespace

extern "C" {

// ---- what nvcc's host stubs need
void **__cudaRegisterFatBinary(void *) {
    static void *handle = nullptr;
    return &handle;
}
void __cudaRegisterFatBinaryEnd(void **) {}
void __cudaUnregisterFatBinary(void **) {}
void __cudaRegisterFunction(void **, const char *, char *, const char *, int, uint3 *, uint3 *, dim3 *, dim3 *, int *) {}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t shmem, void *stream) {
    g_grid = grid;
    g_block = block;
    g_shmem = shmem;
    g_stream = stream;
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3 *grid, dim3 *block, size_t *shmem, void *stream) {
    *grid = g_grid;
    *block = g_block;
    *shmem = g_shmem;
    *static_cast<void **>(stream) = g_stream;
    return cudaSuccess;
}
// engine.cu holds exactly one kernel of its own: publish_stats_kernel(const float *dev, volatile float *host),
// which copies two floats.
cudaError_t cudaLaunchKernel(const void *, dim3, dim3, void **args, size_t, cudaStream_t) {
    const float *dev = *static_cast<const float **>(args[0]);
    float *host = *static_cast<float **>(args[1]);
    host[0] = dev[0];
    host[1] = dev[1];
    return cudaSuccess;
}

// ---- device management
cudaError_t cudaGetDeviceCount(int *n) {
    *n = 8;  // one emulated 8-GPU box (host/run_onnode.sh hands rank i device i)
    return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    std::memset(p, 0, sizeof *p);
    std::strcpy(p->name, "hostcheck (no GPU)");
    p->major = 10;
    p->minor = 0;
    p->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) {
    *d = 0;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) {
    *v = 148;
    return cudaSuccess;
}
const char *cudaGetErrorString(cudaError_t) { return "hostcheck: emulated CUDA runtime"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }

// ---- memory
cudaError_t cudaMalloc(void **p, size_t n) {
    *p = std::malloc(n ? n : 1);
    if (*p) std::memset(*p, 0xA5, n);  // fresh device memory is NOT zero: poison it
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) {
    *p = std::calloc(n ? n : 1, 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p) {
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; ++r)
        std::memmove(static_cast<char *>(dst) + r * dpitch, static_cast<const char *>(src) + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) {
    std::memset(p, v, n);
    return cudaSuccess;
}

// ---- streams / events (everything above already completed when it returned)
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = reinterpret_cast<cudaStream_t>(&g_dummy_handles);
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = reinterpret_cast<cudaEvent_t>(&g_dummy_handles);
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) {
    *ms = 1.0f;
    return cudaSuccess;
}

// ---- peer memory: not emulated
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

}  // extern "C"
