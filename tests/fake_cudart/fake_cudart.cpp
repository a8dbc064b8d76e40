// FAKE CUDA runtime -- test infrastructure, never shipped, never loaded by the product, the GPU tests or the bench.
//
// tools/sanitize_api_host.sh links the REAL host code of libgimic_b200.so (api.cu, host_basis.cpp and the nvcc-generated launch
// stubs of the kernels) against this stand-in for libcudart: "device" memory is zero-initialised host memory, copies are memcpy,
// every kernel launch is a no-op that reports success -- except two small preparation kernels (point gather, per-tile active-set
// count), which are emulated on the host so that api.cu's bookkeeping sees realistic tiles.  Nothing is computed (all results are zeros); what runs is the host-side
// orchestration of every C-ABI entry point -- context creation, staging buffers, the tile / batch / pool bookkeeping, the
// quadrature and property drivers -- under AddressSanitizer + UBSan in a container without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cmath>
#include <map>
#include <mutex>
#include <string>

#include <cuda_runtime_api.h>

#include "../../gimic_b200/csrc/kernels.cuh"      // argument structs of the two kernels that are emulated below

namespace {
thread_local dim3 g_grid, g_block;
thread_local size_t g_shmem = 0;
thread_local void *g_stream = nullptr;
long g_launches = 0;
int g_dummy_stream, g_dummy_event;
std::map<const void *, std::string> g_kernels;      // host stub -> mangled device name (__cudaRegisterFunction)
std::mutex g_mu;

// Two of the small preparation kernels are emulated on the host so that the tile / batch / pool bookkeeping of api.cu sees realistic tiles
// (everything else stays a no-op: in particular the sort, so tiles are runs of the caller's point order -- spatially incoherent, i.e.
// large active sets, which is what stresses the bookkeeping).
void emulate_gather_points(void **a) {              // k_gather_points(r, perm, n, rsx, rsy, rsz): here without the permutation
    const double *r = *(const double **)a[0];
    const long n = *(const long *)a[2];
    double *x = *(double **)a[3], *y = *(double **)a[4], *z = *(double **)a[5];
    for (long i = 0; i < n; ++i) { x[i] = r[3 * i]; y[i] = r[3 * i + 1]; z[i] = r[3 * i + 2]; }
}
void emulate_tile_count(void **a, unsigned ntiles) {   // k_tile_count(B, rsx, rsy, rsz, segs, geo, info), same counts as atom_active
    const gb::DevBasis &B = *(const gb::DevBasis *)a[0];
    const double *sx = *(const double **)a[1], *sy = *(const double **)a[2], *sz = *(const double **)a[3];
    const gb::TileSeg *segs = *(const gb::TileSeg **)a[4];
    gb::TileGeo *geo = *(gb::TileGeo **)a[5];
    gb::TileInfo *info = *(gb::TileInfo **)a[6];
    const int al = B.slot_align - 1;
    for (unsigned t = 0; t < ntiles; ++t) {
        const gb::TileSeg sg = segs[t];
        gb::TileGeo tg{1e300, 1e300, 1e300, -1e300, -1e300, -1e300, 0.0, 0.0};
        for (int p = 0; p < sg.npts; ++p) {
            const long q = sg.pt0 + p;
            tg.lox = std::fmin(tg.lox, sx[q]); tg.hix = std::fmax(tg.hix, sx[q]);
            tg.loy = std::fmin(tg.loy, sy[q]); tg.hiy = std::fmax(tg.hiy, sy[q]);
            tg.loz = std::fmin(tg.loz, sz[q]); tg.hiz = std::fmax(tg.hiz, sz[q]);
        }
        const double cx = 0.5 * (tg.lox + tg.hix), cy = 0.5 * (tg.loy + tg.hiy), cz = 0.5 * (tg.loz + tg.hiz);
        double gmax = 0.0; int imax = 0;
        for (int p = 0; p < sg.npts; ++p) {
            const long q = sg.pt0 + p;
            tg.rho = std::fmax(tg.rho, std::sqrt((sx[q] - cx) * (sx[q] - cx) + (sy[q] - cy) * (sy[q] - cy) + (sz[q] - cz) * (sz[q] - cz)));
            if (p + 1 < sg.npts) {
                const double g = std::sqrt((sx[q + 1] - sx[q]) * (sx[q + 1] - sx[q]) + (sy[q + 1] - sy[q]) * (sy[q + 1] - sy[q]) + (sz[q + 1] - sz[q]) * (sz[q + 1] - sz[q]));
                if (g > gmax) { gmax = g; imax = p; }
            }
        }
        int cnt = 0, nat = 0, nre = 0;
        for (int at = 0; at < B.natoms; ++at) {
            const double x = B.atom_xyz[3 * at], y = B.atom_xyz[3 * at + 1], z = B.atom_xyz[3 * at + 2];
            double d2 = 1e300;
            for (int p = 0; p < sg.npts; ++p) {
                const long q = sg.pt0 + p;
                d2 = std::fmin(d2, (sx[q] - x) * (sx[q] - x) + (sy[q] - y) * (sy[q] - y) + (sz[q] - z) * (sz[q] - z));
            }
            const double lim = std::sqrt(d2) - 1e-9;
            int nfun = 0;
            for (int s = B.atom_shell_off[at]; s < B.atom_shell_off[at + 1] && B.sh_thr[s] >= lim; ++s) nfun += (B.sh_l[s] + 1) * (B.sh_l[s] + 2) / 2;
            cnt += (nfun + al) & ~al; nat += nfun > 0; nre += nfun;
        }
        geo[t] = tg;
        info[t] = gb::TileInfo{(float)tg.rho, (float)gmax, imax, cnt, nat, nre};
    }
}
}  // namespace

extern "C" {

void **__cudaRegisterFatBinary(void *) { static void *handle = nullptr; return &handle; }
void __cudaRegisterFatBinaryEnd(void **) {}
void __cudaUnregisterFatBinary(void **) {}
void __cudaRegisterFunction(void **, const char *host_fun, char *, const char *device_name, int, uint3 *, uint3 *, dim3 *, dim3 *, int *) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_kernels[host_fun] = device_name ? device_name : "";
}
void __cudaRegisterVar(void **, char *, char *, const char *, int, size_t, int, int) {}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t shmem, struct CUstream_st *stream) {
    g_grid = grid; g_block = block; g_shmem = shmem; g_stream = stream;
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3 *grid, dim3 *block, size_t *shmem, void *stream) {
    *grid = g_grid; *block = g_block; *shmem = g_shmem; *(void **)stream = g_stream;
    return cudaSuccess;
}

cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t, cudaStream_t) {
    if (grid.x == 0 || block.x == 0) return cudaErrorInvalidConfiguration;
    std::string name;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        ++g_launches;
        auto it = g_kernels.find(func);
        if (it != g_kernels.end()) name = it->second;
    }
    if (!std::getenv("FAKE_CUDA_NO_EMULATION")) {
        if (name.find("k_gather_points") != std::string::npos) emulate_gather_points(args);
        else if (name.find("k_tile_count") != std::string::npos) emulate_tile_count(args, grid.x);
    }
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
