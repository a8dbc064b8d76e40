// How fast can DMMA be fed?  Isolates the consumer loop of k_jtensor: 16 rows x (NT n8-tiles) per warp, k4 steps,
// fragments either re-loaded from shared memory every step (like the kernel) or kept in registers.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int LDP = 132, LDB = 20, BK = 32;

__device__ __forceinline__ void mma4(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(b0));
}
__device__ __forceinline__ void mma8(double (&c)[4], double a0, double a1, double a2, double a3, double b0, double b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

// MODE 0: LDS per step, k4;  1: LDS per step, k8;  2: registers only (k4);  3: LDS k4 with explicit double buffering of B frags
__device__ __forceinline__ void cp_async_8(unsigned smem, const void *g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem), "l"(g)); }
template <int NT, int MODE>
__global__ void __launch_bounds__(512, 1) k_feed(double *out, int iters, int nwarps_used, const double *gsrc = nullptr, int interfere = 0, long gstride = 0) {
    extern __shared__ double sm[];
    double *sA = sm, *sB = sm + BK * LDP;
    for (int i = threadIdx.x; i < BK * LDP + 7 * BK * LDB; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    if (warp >= nwarps_used) {
        if (!interfere || warp >= nwarps_used + 4) return;
        // producer-like traffic: per "stage" 28 x 8-byte gathers per lane (rows gstride apart) into a scratch tile, paced like the real kernel
        const int pw = warp - nwarps_used;
        unsigned dst0 = (unsigned)__cvta_generic_to_shared(sm + BK * LDP + 7 * BK * LDB) + (lane & 15) * 8;
        for (int it = 0; it < iters; ++it) {
            const double *src = gsrc + ((long)(it * 37 + pw * 8 + (lane >> 4)) % 4000) * gstride + (lane & 15) + (it % 64) * 16;
            for (int k = (lane >> 4) + 2 * pw; k < 32; k += 8)
                for (int q = 0; q < 7; ++q) cp_async_8(dst0 + (q * BK * LDB + k * LDB) * 8, src + (long)k * gstride + q * 4096 * gstride);
            asm volatile("cp.async.commit_group;"); 
            if (interfere == 1) asm volatile("cp.async.wait_group 2;");
            // pace: roughly one stage per 7168 cycles
            long long t0 = clock64(); while (clock64() - t0 < 6000) { }
        }
        asm volatile("cp.async.wait_group 0;");
        return;
    }
    const int row0 = (warp % 8) * 16;
    double acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll 4
            for (int ks = 0; ks < BK / 4; ++ks) {
                const double *pa = sA + (ks * 4 + t) * LDP + row0 + g;
                const double a0 = pa[0], a1 = pa[8];
                const double *pb = sB + (ks * 4 + t) * LDB + g;
#pragma unroll
                for (int i = 0; i < NT; ++i) mma4(acc[i], a0, a1, pb[(i >> 1) * BK * LDB + (i & 1) * 8]);
            }
        } else if (MODE == 1) {
#pragma unroll 2
            for (int ks = 0; ks < BK / 8; ++ks) {
                const double *pa = sA + (ks * 8 + t) * LDP + row0 + g;
                const double a0 = pa[0], a1 = pa[8], a2 = pa[4 * LDP], a3 = pa[4 * LDP + 8];
                const double *pb = sB + (ks * 8 + t) * LDB + g;
#pragma unroll
                for (int i = 0; i < NT; ++i) mma8(acc[i], a0, a1, a2, a3, pb[(i >> 1) * BK * LDB + (i & 1) * 8], pb[(i >> 1) * BK * LDB + 4 * LDB + (i & 1) * 8]);
            }
        } else if (MODE == 2) {
            double a0 = sA[lane], a1 = sA[lane + 32], b[NT];
#pragma unroll
            for (int i = 0; i < NT; ++i) b[i] = sB[lane + 32 * i];
#pragma unroll 4
            for (int ks = 0; ks < BK / 4; ++ks) {
#pragma unroll
                for (int i = 0; i < NT; ++i) mma4(acc[i], a0, a1, b[i]);
            }
        } else {
            // explicit software pipeline: fragments of step ks+1 are loaded while step ks computes
            double a0, a1, b[NT], na0, na1, nb[NT];
            { const double *pa = sA + t * LDP + row0 + g; a0 = pa[0]; a1 = pa[8]; const double *pb = sB + t * LDB + g;
#pragma unroll
              for (int i = 0; i < NT; ++i) b[i] = pb[(i >> 1) * BK * LDB + (i & 1) * 8]; }
#pragma unroll 2
            for (int ks = 0; ks < BK / 4; ++ks) {
                const int kn = (ks + 1) % (BK / 4);
                const double *pa = sA + (kn * 4 + t) * LDP + row0 + g; na0 = pa[0]; na1 = pa[8];
                const double *pb = sB + (kn * 4 + t) * LDB + g;
#pragma unroll
                for (int i = 0; i < NT; ++i) { nb[i] = pb[(i >> 1) * BK * LDB + (i & 1) * 8]; mma4(acc[i], a0, a1, b[i]); }
                a0 = na0; a1 = na1;
#pragma unroll
                for (int i = 0; i < NT; ++i) b[i] = nb[i];
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT, int MODE>
int run(const char *name, int nwarps, double *out, int nsm, const double *gsrc = nullptr, int interfere = 0) {
    size_t smem = (BK * LDP + 2 * 7 * BK * LDB) * sizeof(double);
    const int ITERS_PER_STAGE = 1; (void)ITERS_PER_STAGE;
    CK(cudaFuncSetAttribute(k_feed<NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_feed<NT, MODE><<<nsm, 512, smem>>>(out, iters, nwarps, gsrc, interfere, 10008); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k_feed<NT, MODE><<<nsm, 512, smem>>>(out, iters, nwarps, gsrc, interfere, 10008); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double fl = 2.0 * 16 * 8 * 4 * NT * (BK / 4) * (double)iters * nwarps * nsm;
    printf("%-34s NT=%2d warps/SM=%2d : %8.3f ms  %6.2f TFLOP/s\n", name, NT, nwarps, best, fl / best * 1e-9);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount; double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 512));
    double *gsrc; CK(cudaMalloc(&gsrc, sizeof(double) * 10008L * 4096 * 7 + 1024)); CK(cudaMemset(gsrc, 0, sizeof(double) * 10008L * 4096 * 7));
    printf("-- with 4 producer-like warps streaming 8-byte LDGSTS gathers (one 28.7 KB tile per ~6000 cycles per SM)\n");
    run<14, 0>("LDS k4 + LDGSTS interference", 8, out, nsm, gsrc, 1);
    run<14, 0>("LDS k4 + LDGSTS interference", 8, out, nsm, gsrc, 1);
    printf("-- alone\n");
    for (int w : {8}) {
        run<14, 0>("LDS per step, m16n8k4", w, out, nsm);
        run<14, 1>("LDS per step, m16n8k8", w, out, nsm);
        run<14, 2>("registers only, m16n8k4", w, out, nsm);
        run<14, 3>("LDS k4, explicit double buffer", w, out, nsm);
        run<7, 0>("LDS per step, m16n8k4", w, out, nsm);
        run<7, 3>("LDS k4, explicit double buffer", w, out, nsm);
    }
    return 0;
}
