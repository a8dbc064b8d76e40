// FP64 peak microbenchmarks for B200 (sm_100a): DFMA, DMMA shapes, and an
// m16n8k8 fragment-layout check. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
// The numbers feed the roofline denominators because
// MEASURED_PEAKS.json only carries bf16 and HBM-copy figures.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double *out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1684(double *out, int iters, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double *out, int iters, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double *out, int iters, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// layout check: one warp computes C(16x8) = A(16x8) * B(8x8) with the fragment map
// a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4); b0=(k=t,n=g) b1=(k=t+4,n=g);
// c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)
__global__ void k_layout1688(const double *A, const double *B, double *C) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a0 = A[g * 8 + t], a1 = A[(g + 8) * 8 + t], a2 = A[g * 8 + t + 4], a3 = A[(g + 8) * 8 + t + 4];
    double b0 = B[t * 8 + g], b1 = B[(t + 4) * 8 + g];
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1; C[(g + 8) * 8 + 2 * t] = c2; C[(g + 8) * 8 + 2 * t + 1] = c3;
}
// m8n8k4: a0=(g,t); b0=(k=t,n=g); c0=(g,2t), c1=(g,2t+1)
__global__ void k_layout884(const double *A, const double *B, double *C) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a0 = A[g * 4 + t], b0 = B[t * 8 + g], c0 = 0, c1 = 0;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a0), "d"(b0));
    C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int nsm = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 256));
    const int iters = 20000;
    for (int bps = 1; bps <= 8; bps *= 2) {
        int grid = nsm * bps;
        double ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 16 * iters * 256.0 * grid;
        printf("DFMA      blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    }
#define RUN_MMA(name, kern, flops_per_mma, nacc) \
    for (int bps = 1; bps <= 4; bps *= 2) { \
        int grid = nsm * bps; \
        double ms = time_ms([&] { kern<nacc><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); }, 5); \
        double fl = (double)(flops_per_mma) * nacc * (iters / 4) * 8.0 * grid; \
        printf("%-10s nacc %d blocks/SM %d: %.3f ms  %.2f TFLOP/s\n", name, nacc, bps, ms, fl / ms * 1e-9); \
    }
    RUN_MMA("m8n8k4", k_dmma884, 2 * 8 * 8 * 4, 4)
    RUN_MMA("m8n8k4", k_dmma884, 2 * 8 * 8 * 4, 8)
    RUN_MMA("m16n8k4", k_dmma1684, 2 * 16 * 8 * 4, 4)
    RUN_MMA("m16n8k8", k_dmma1688, 2 * 16 * 8 * 8, 2)
    RUN_MMA("m16n8k8", k_dmma1688, 2 * 16 * 8 * 8, 4)
    RUN_MMA("m16n8k8", k_dmma1688, 2 * 16 * 8 * 8, 8)
    RUN_MMA("m16n8k16", k_dmma16816, 2 * 16 * 8 * 16, 4)

    // layout checks
    double hA[128], hB[64], hC[128], *dA, *dB, *dC;
    for (int i = 0; i < 128; ++i) hA[i] = sin(0.37 * i) + 0.1 * i;
    for (int i = 0; i < 64; ++i) hB[i] = cos(0.91 * i) - 0.05 * i;
    CK(cudaMalloc(&dA, sizeof hA)); CK(cudaMalloc(&dB, sizeof hB)); CK(cudaMalloc(&dC, sizeof hC));
    CK(cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice));
    k_layout1688<<<1, 32>>>(dA, dB, dC); CK(cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 8; ++j) { double s = 0; for (int k = 0; k < 8; ++k) s += hA[i * 8 + k] * hB[k * 8 + j]; err = fmax(err, fabs(s - hC[i * 8 + j])); }
    printf("layout m16n8k8 max abs err %.3e (expect ~1e-15)\n", err);
    k_layout884<<<1, 32>>>(dA, dB, dC); CK(cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost));
    err = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += hA[i * 4 + k] * hB[k * 8 + j]; err = fmax(err, fabs(s - hC[i * 8 + j])); }
    printf("layout m8n8k4 max abs err %.3e (expect ~1e-15)\n", err);
    return 0;
}
