// What do the GIAO taps of k_jtensor cost beyond their own FP64 issue slots?  Isolates the consumer loop (8 warps per SM = 2 per
// scheduler, 16 rows x 16 columns x 4 planes per warp = 8 m16n8k4 accumulator tiles, fragments from shared memory every k4 step) and
// adds, every PERIOD k-steps, the 24 DFMAs a tap issues on the two D-plane accumulator tiles (Z_d += w_d * C).  DMMA and DFMA share the
// FP64 datapath; the question is whether alternating them costs more than the sum of their issue times.
//   MODE 0  no taps (the DMMA ceiling of this loop)
//   MODE 1  taps inline behind a data-dependent branch, like the kernel
//   MODE 2  like 1, but the two warps of a scheduler meet on a named barrier first, so that their DFMA bursts coincide
//   MODE 3  taps as straight-line predicated code (no branch)
//   MODE 4  like 1 with only 8 DFMAs per tap (the J path's tap)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_tap tools/dmma_tap.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int LDP = 132, LDB = 36, BK = 32;

__device__ __forceinline__ void mma4(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b0));
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_tap(double *out, int iters, unsigned mask_seed, const double *wtab) {
    extern __shared__ double sm[];
    double *sA = sm, *sB = sm + BK * LDP;       // A: [BK][LDP]; B: two pair-planes [BK][LDB]
    __shared__ double s_w[64 * 3];
    for (int i = threadIdx.x; i < BK * LDP + 2 * BK * LDB; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    if (threadIdx.x < 192) s_w[threadIdx.x] = wtab[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int row0 = warp * 16;
    double acc[4][2][4], zac[3][2][4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[q][h][i] = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) zac[d][h][i] = 0.0;
    double cx = s_w[0], cy = s_w[1], cz = s_w[2];
    int ia = 0;
    for (int it = 0; it < iters; ++it) {
        const unsigned m8 = MODE == 0 ? 0u : (mask_seed >> (it & 7)) | (mask_seed << (8 - (it & 7)));   // which of the 8 k-steps of this stage end an atom
#pragma unroll 8
        for (int ks = 0; ks < BK / 4; ++ks) {
            const double *pa = sA + (ks * 4 + t) * LDP + row0 + g;
            const double a0 = pa[0], a1 = pa[8];
            const double2 *pb = reinterpret_cast<const double2 *>(sB + (ks * 4 + t) * LDB) + g;
#pragma unroll
            for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double2 b = pb[pp * (BK * LDB / 2) + h * 8];
                    mma4(acc[2 * pp][h], a0, a1, b.x);
                    mma4(acc[2 * pp + 1][h], a0, a1, b.y);
                }
            const bool tap = (m8 >> ks) & 1u;
            if (MODE == 3) {
                const double wx = tap ? cx : 0.0, wy = tap ? cy : 0.0, wz = tap ? cz : 0.0;
                if (tap) {      // the compiler predicates this
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double cv = acc[0][h][i];
                            zac[0][h][i] = fma(wx, cv, zac[0][h][i]); zac[1][h][i] = fma(wy, cv, zac[1][h][i]); zac[2][h][i] = fma(wz, cv, zac[2][h][i]);
                        }
                    ia = (ia + 1) & 63; cx = s_w[3 * ia]; cy = s_w[3 * ia + 1]; cz = s_w[3 * ia + 2];
                }
            } else if (MODE != 0 && tap) {
                if (MODE == 2) asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double cv = acc[0][h][i];
                        zac[0][h][i] = fma(cx, cv, zac[0][h][i]);
                        if (MODE != 4) { zac[1][h][i] = fma(cy, cv, zac[1][h][i]); zac[2][h][i] = fma(cz, cv, zac[2][h][i]); }
                    }
                ia = (ia + 1) & 63; cx = s_w[3 * ia]; cy = s_w[3 * ia + 1]; cz = s_w[3 * ia + 2];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) s += acc[q][h][i];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) s += zac[d][h][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int run(const char *name, unsigned mask, double *out, const double *wtab, int nsm) {
    const size_t smem = (BK * LDP + 2 * BK * LDB) * sizeof(double);
    CK(cudaFuncSetAttribute(k_tap<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_tap<MODE><<<nsm, 256, smem>>>(out, iters, mask, wtab); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k_tap<MODE><<<nsm, 256, smem>>>(out, iters, mask, wtab); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    const int ntap = __builtin_popcount(mask & 0xff);
    const double mma = 2.0 * 16 * 8 * 4 * 8 * (BK / 4) * (double)iters * 8 * nsm;                       // 8 m16n8k4 per k-step per warp
    const double dfma = (MODE == 0 ? 0.0 : 2.0 * (MODE == 4 ? 8 : 24) * 32 * ntap) * (double)iters * 8 * nsm;
    printf("%-58s taps/stage %d : %8.3f ms  DMMA %6.2f TF  DMMA+DFMA %6.2f TF\n", name, MODE ? ntap : 0, best, mma / best * 1e-9, (mma + dfma) / best * 1e-9);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    double *out, *wtab; CK(cudaMalloc(&out, sizeof(double) * nsm * 256)); CK(cudaMalloc(&wtab, sizeof(double) * 192));
    double h[192]; for (int i = 0; i < 192; ++i) h[i] = 1e-3 * (i + 1);
    CK(cudaMemcpy(wtab, h, sizeof h, cudaMemcpyHostToDevice));
    printf("device %s, %d SMs; 8 warps/SM, 8 m16n8k4 per k-step per warp, 8 k-steps per stage\n", p.name, nsm);
    run<0>("no taps", 0, out, wtab, nsm);
    for (unsigned mask : {0x11u, 0x49u, 0x55u, 0xffu}) {     // 2, 3, 4, 8 taps per 8 k-steps (the flake: ~3.4)
        run<1>("taps behind a branch (kernel)", mask, out, wtab, nsm);
        run<2>("taps, the two warps of a scheduler synchronised", mask, out, wtab, nsm);
        run<3>("taps predicated", mask, out, wtab, nsm);
        run<4>("8-DFMA taps (J path)", mask, out, wtab, nsm);
    }
    return 0;
}
