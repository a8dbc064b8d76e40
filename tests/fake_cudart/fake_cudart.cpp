// FAKE CUDA runtime -- test infrastructure, never shipped, never loaded by the product, the GPU tests or the bench.
//
// tools/sanitize_api_host.sh links the REAL host code of libgimic_b200.so (api.cu, host_basis.cpp and the nvcc-generated launch
// stubs of the kernels) against this stand-in for libcudart: "device" memory is zero-initialised host memory, copies are memcpy,
// every kernel launch is a no-op that reports success.  Nothing is computed (all results are zeros); what runs is the host-side
// orchestration of every C-ABI entry point -- context creation, staging buffers, the tile / batch / pool bookkeeping, the
// quadrature and property drivers -- under AddressSanitizer + UBSan in a container without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime_api.h>

namespace {
thread_local dim3 g_grid, g_block;
thread_local size_t g_shmem = 0;
thread_local void *g_stream = nullptr;
long g_launches = 0;
int g_dummy_stream, g_dummy_event;
}  // namespace

extern "C" {

void **__cudaRegisterFatBinary(void *) { static void *handle = nullptr; return &handle; }
void __cudaRegisterFatBinaryEnd(void **) {}
void __cudaUnregisterFatBinary(void **) {}
void __cudaRegisterFunction(void **, const char *, char *, const char *, int, uint3 *, uint3 *, dim3 *, dim3 *, int *) {}
void __cudaRegisterVar(void **, char *, char *, const char *, int, size_t, int, int) {}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t shmem, struct CUstream_st *stream) {
    g_grid = grid; g_block = block; g_shmem = shmem; g_stream = stream;
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3 *grid, dim3 *block, size_t *shmem, void *stream) {
    *grid = g_grid; *block = g_block; *shmem = g_shmem; *(void **)stream = g_stream;
    return cudaSuccess;
}

cudaError_t cudaLaunchKernel(const void *, dim3 grid, dim3 block, void **, size_t, cudaStream_t) {
    if (grid.x == 0 || block.x == 0) return cudaErrorInvalidConfiguration;
    ++g_launches;
    return cudaSuccess;
}
cudaError_t cudaGetDeviceCount(int *n) { *n = std::getenv("FAKE_CUDA_NO_DEVICE") ? 0 : 2; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return (d >= 0 && d < 2) ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *p, int) {
    std::memset(p, 0, sizeof *p);
    std::snprintf(p->name, sizeof p->name, "FAKE B200 (host memory, no-op kernels)");
    p->multiProcessorCount = 148; p->totalGlobalMem = (size_t)8 << 30; p->major = 10; p->minor = 0;
    p->sharedMemPerBlockOptin = 232448; p->sharedMemPerBlock = 49152; p->maxThreadsPerBlock = 1024; p->warpSize = 32;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr attr, int) {
    switch (attr) {
        case cudaDevAttrMultiProcessorCount: *v = 148; break;
        case cudaDevAttrMaxSharedMemoryPerBlock: *v = 49152; break;
        case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 232448; break;
        case cudaDevAttrMaxThreadsPerBlock: *v = 1024; break;
        case cudaDevAttrWarpSize: *v = 32; break;
        case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
        case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
        default: *v = 1024; break;
    }
    return cudaSuccess;
}
cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *a, const void *) {
    std::memset(a, 0, sizeof *a);
    a->ptxVersion = 100; a->binaryVersion = 100; a->maxThreadsPerBlock = 1024; a->numRegs = 32;
    return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *n, const void *, int, size_t, unsigned) { *n = 1; return cudaSuccess; }

cudaError_t cudaMalloc(void **p, size_t n) {
    if (n > ((size_t)6 << 30)) { *p = nullptr; return cudaErrorMemoryAllocation; }      // bounded: this is host memory
    *p = std::calloc(n ? n : 1, 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyToSymbol(const void *, const void *, size_t, size_t, cudaMemcpyKind) { return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) std::memset(d, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)&g_dummy_stream; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)&g_dummy_event; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake CUDA runtime error"; }

long fake_cuda_launch_count(void) { return g_launches; }

}  // extern "C"
