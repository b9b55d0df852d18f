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

#include <cstdio>
#include <map>
#include <mutex>
#include <string>

namespace {
thread_local dim3 g_grid, g_block;
thread_local size_t g_shmem = 0;
thread_local void *g_stream = nullptr;
int g_dummy_handles = 0;

// ---- launch log: every kernel launch that reaches this runtime is checked against the limits the
// real one enforces at launch time (sm_100: 1024 threads per block, grid.x < 2^31, grid.y/z <= 65535,
// 227 KB of dynamic shared memory, portable clusters of <= 8 CTAs that divide the grid).  In the
// DORY_LAUNCHCHECK build the product's own launchers (spmm.cu, dense.cu, gat.cu) are linked in, so this
// is where their grid arithmetic and kernel selection get exercised without a GPU.
std::mutex g_log_mutex;
std::map<const void *, std::string> g_kernel_names;
unsigned long long g_launches = 0, g_violations = 0;
std::string g_first_violation;

void check_launch(const void *func, dim3 grid, dim3 block, size_t shmem, unsigned cx, unsigned cy, unsigned cz) {
    std::lock_guard<std::mutex> lk(g_log_mutex);
    ++g_launches;
    const char *why = nullptr;
    const unsigned long long threads = (unsigned long long)block.x * block.y * block.z;
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) why = "empty grid";
    else if (grid.x > 2147483647u || grid.y > 65535u || grid.z > 65535u) why = "grid dimension beyond the limit";
    else if (threads == 0 || threads > 1024 || block.x > 1024 || block.y > 1024 || block.z > 64) why = "block shape beyond the limit";
    else if (shmem > 232448) why = "dynamic shared memory beyond 227 KB";
    else if (cx * cy * cz == 0 || cx * cy * cz > 8) why = "cluster beyond the portable size";
    else if (grid.x % cx || grid.y % cy || grid.z % cz) why = "cluster shape does not divide the grid";
    if (why) {
        ++g_violations;
        if (g_first_violation.empty()) {
            auto it = g_kernel_names.find(func);
            char buf[512];
            std::snprintf(buf, sizeof buf, "%s: %s (grid %u x %u x %u, block %u x %u x %u, smem %zu, cluster %u x %u x %u)",
                          it == g_kernel_names.end() ? "?" : it->second.c_str(), why, grid.x, grid.y, grid.z, block.x,
                          block.y, block.z, shmem, cx, cy, cz);
            g_first_violation = buf;
        }
    }
}
}  // namespace

// Read (and reset) the launch log: number of launches seen, number that broke a limit, the first offender.
extern "C" void hostcheck_launch_log(unsigned long long *launches, unsigned long long *violations, char *first, size_t cap) {
    std::lock_guard<std::mutex> lk(g_log_mutex);
    *launches = g_launches;
    *violations = g_violations;
    if (first && cap) std::snprintf(first, cap, "%s", g_first_violation.c_str());
    g_launches = g_violations = 0;
    g_first_violation.clear();
}

#ifdef DORY_COMMCHECK
bool hostcheck_comm_kernel(const std::string &name, void **args);  // fake_nccl.cpp
#endif

extern "C" {

// ---- what nvcc's host stubs need
void **__cudaRegisterFatBinary(void *) {
    static void *handle = nullptr;
    return &handle;
}
void __cudaRegisterFatBinaryEnd(void **) {}
void __cudaUnregisterFatBinary(void **) {}
void __cudaRegisterFunction(void **, const char *hostFun, char *, const char *deviceName, int, uint3 *, uint3 *, dim3 *, dim3 *,
                            int *) {
    std::lock_guard<std::mutex> lk(g_log_mutex);
    g_kernel_names[hostFun] = deviceName ? deviceName : "?";
}
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
// The one kernel that is emulated here rather than in cpu_kernels.cpp is engine.cu's own
// publish_stats_kernel(const float *dev, volatile float *host), which copies two floats.  Every other
// kernel that reaches this point (DORY_LAUNCHCHECK build only) is checked and not executed.
cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t shmem, cudaStream_t) {
    check_launch(func, grid, block, shmem, 1, 1, 1);
    bool publish;
    std::string name;
    {
        std::lock_guard<std::mutex> lk(g_log_mutex);
        auto it = g_kernel_names.find(func);
        if (it != g_kernel_names.end()) name = it->second;
        publish = name.find("publish_stats") != std::string::npos;
    }
#ifdef DORY_COMMCHECK
    if (!publish && hostcheck_comm_kernel(name, args)) return cudaSuccess;
#endif
    if (publish) {
        const float *dev = *static_cast<const float **>(args[0]);
        float *host = *static_cast<float **>(args[1]);
        host[0] = dev[0];
        host[1] = dev[1];
    }
    return cudaSuccess;
}
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t *cfg, const void *func, void **) {
    unsigned cx = 1, cy = 1, cz = 1;
    for (unsigned i = 0; i < cfg->numAttrs; ++i)
        if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) {
            cx = cfg->attrs[i].val.clusterDim.x;
            cy = cfg->attrs[i].val.clusterDim.y;
            cz = cfg->attrs[i].val.clusterDim.z;
        }
    check_launch(func, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, cx, cy, cz);
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
#ifdef DORY_LAUNCHCHECK
    *p = std::calloc(n ? n : 1, 1);  // nothing computes on it: untouched pages cost nothing
#else
    *p = std::malloc(n ? n : 1);
    if (*p) std::memset(*p, 0xA5, n);  // fresh device memory is NOT zero: poison it
#endif
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
// (round 2: the exchange stream of the overlap path; never reached on the emulated runtime, present so that
// the engine object links)
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) {
    *s = reinterpret_cast<cudaStream_t>(&g_dummy_handles);
    return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) {
    *lo = 0;
    *hi = 0;
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

// ---- peer memory
#ifdef DORY_COMMCHECK
// ranks are threads of one process: the handle carries the pointer itself
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    std::memset(h, 0, sizeof *h);
    std::memcpy(h, &p, sizeof p);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    std::memcpy(p, &h, sizeof *p);
    return cudaSuccess;
}
#else
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
#endif
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

}  // extern "C"
