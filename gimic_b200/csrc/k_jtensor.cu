// k_jtensor: the contraction  T(r) = f(Phi(r), dPhi(r); D, P_x, P_y, P_z)  for tiles of 128 points.
//
// Replaces `contract` of the reference (src/libgimic/jtensor.F90:148-237: 7 nbf x nbf GEMVs + 28 dot
// products per point) by one FP64 tensor-core GEMM per tile with a fused epilogue:
//
//   X_q[p, nu] = sum_mu Phi[p, mu] * B_q[mu, nu],   q = 0..6,  mu/nu over the tile's ACTIVE functions
//     B_0 = D,  B_1..3 = P_x,P_y,P_z,  B_4..6[mu,nu] = D[mu,nu] (R_nu - R_mu)_d      (d = x,y,z)
//
// and, with e_0 = Phi, e_m = dPhi/dr_m, y = r x (X_4, X_5, X_6):
//
//   Tp(m,b) = sum_nu (X_{1+b} + y_b)[nu] e_m[nu]          (ppd + prsp1 + the d_m part of prsp2)
//   V_d     = sum_nu R_{nu,d} X_0[nu] e_0[nu]             (the (e_m x R) Phi part of prsp2, bfeval.f90:228-240)
//   rho     = sum_nu X_0[nu] e_0[nu]                      (diapam, jtensor.F90:168)
//   ct(m,b) = 1/2 [Tp(m,b) + sum_d eps(b,m,d) V_d]  +  eps(m,b,c) rho r_c / 2      (jtensor.F90:209-235)
//
// The identity behind B_4..6 (gauge-difference form of dendb/d2fvec, bfeval.f90:168-246):
//   -sum_nu dendb_b[nu] e_m[nu] + sum_nu denbf[nu] e_m[nu] g_{A(nu),b}
//        = sum_nu e_m[nu] sum_{c,d} eps(b,c,d) r_c sum_mu Phi_mu D[mu,nu] (R_{nu,d} - R_{mu,d}).
// Screened functions are exact zeros in the reference, so restricting mu, nu to the tile's active
// set changes nothing but the order of summation.
//
// Mapping: 256 threads = 8 warps; warp w owns rows 16w..16w+15 of the tile and all 7x2 n8 tiles of
// the current 16-wide nu chunk (56 fp64 accumulators / thread), so the per-point epilogue sums stay
// in registers for the whole tile and need only a 4-lane shuffle reduction at the end.
// mma.sync.m16n8k8.f64 lowers to 4 DMMA.8x8x4 on sm_100a (there is no FP64 tcgen05 kind).
#include "kernels.cuh"

namespace gb {

__device__ __forceinline__ void mma_16x8x8_f64(double (&c)[4], double a0, double a1, double a2, double a3, double b0, double b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_8(void *smem, const void *gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int NQ>
struct Smem {
    static constexpr int A_DOUBLES = BK * LDP;
    static constexpr int B_DOUBLES = NQ * BK * LDB;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
    static constexpr size_t BYTES = (size_t)STAGES * STAGE_DOUBLES * 8 + (size_t)FCAP * 4 + 16;
};

template <bool GIAO>
__global__ void __launch_bounds__(256, 1) k_jtensor(JtensorArgs a) {
    constexpr int NQ = GIAO ? NQ_GIAO : NQ_NOGIAO;
    using SM = Smem<NQ>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_stage = reinterpret_cast<double *>(smem_raw);
    int *s_fidx = reinterpret_cast<int *>(smem_raw + (size_t)STAGES * SM::STAGE_DOUBLES * 8);
    int *s_tile = s_fidx + FCAP;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = warp * 16;
    // loader roles
    const int ldn = tid & 15, ldk = tid >> 4;   // B gather: this thread fetches element (k = ldk, nu = ldn) of all NQ planes

    for (;;) {
        __syncthreads();
        if (tid == 0) *s_tile = atomicAdd(a.counter, 1);
        __syncthreads();
        const int tile = *s_tile;
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        const int rowA = row0 + g, rowB = row0 + g + 8;
        const bool vA = rowA < td.npts, vB = rowB < td.npts;

        if (td.nact == 0) {   // nothing within screening range: the reference returns exact zeros
            if (t == 0) {
                if (vA) { long o = a.perm[td.pt0 + rowA]; for (int i = 0; i < 9; ++i) a.tens[9 * o + i] = 0.0; if (a.edens) a.edens[o] = 0.0; }
                if (vB) { long o = a.perm[td.pt0 + rowB]; for (int i = 0; i < 9; ++i) a.tens[9 * o + i] = 0.0; if (a.edens) a.edens[o] = 0.0; }
            }
            continue;
        }
        const int nact = td.nact;
        const int nkc = nact / BK, nvc = nact / NV;
        const long NIT = (long)nkc * nvc;
        const double *panel = a.panel_pool + td.panel_off;
        const long plane = (long)nact * LDP;
        const int *gfidx = a.fidx_pool + td.fidx_off;
        const int *fidx = gfidx;
        if (nact <= FCAP) {
            for (int i = tid; i < nact; i += 256) s_fidx[i] = gfidx[i];
            fidx = s_fidx;
        }
        __syncthreads();

        // coordinates of this thread's two points (absolute, as r enters jtensor.F90:112 and bfeval.f90:168-189)
        const long pA = td.pt0 + (vA ? rowA : 0), pB = td.pt0 + (vB ? rowB : 0);
        const double rAx = a.rsx[pA], rAy = a.rsy[pA], rAz = a.rsz[pA];
        const double rBx = a.rsx[pB], rBy = a.rsy[pB], rBz = a.rsz[pB];

        double eA[13], eB[13];   // Tp(m,b) at [m + 3b], V_d at [9 + d], rho at [12]
#pragma unroll
        for (int i = 0; i < 13; ++i) { eA[i] = 0.0; eB[i] = 0.0; }

        auto issue = [&](long itl) {
            if (itl < NIT) {
                const int vc = (int)(itl / nkc), kc = (int)(itl - (long)vc * nkc);
                double *sA = s_stage + (size_t)(itl % STAGES) * SM::STAGE_DOUBLES;
                double *sB = sA + SM::A_DOUBLES;
                // A: Phi plane rows [kc*BK, kc*BK+BK) x 128 points, 16 B chunks
                const double *srcA = panel + (long)kc * BK * LDP;
#pragma unroll
                for (int i = 0; i < (BK * MT / 2) / 256; ++i) {
                    int c = tid + 256 * i, k = c >> 6, c2 = (c & 63) * 2;
                    cp_async_16(sA + k * LDP + c2, srcA + (long)k * LDP + c2);
                }
                // B: gathered element (mu, nu) of every operand plane
                const long mu = fidx[kc * BK + ldk], nu = fidx[vc * NV + ldn];
                const double *srcB = a.Bop + mu * a.ldb + nu;
                double *dstB = sB + ldk * LDB + ldn;
#pragma unroll
                for (int q = 0; q < NQ; ++q) cp_async_8(dstB + q * BK * LDB, srcB + q * a.plane_stride);
            }
            cp_async_commit();
        };

#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s) issue(s);

        double acc[NQ][2][4];
        int kc = 0, vc = 0;
        for (long it = 0; it < NIT; ++it) {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
            issue(it + STAGES - 1);
            if (kc == 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[q][h][i] = 0.0;
            }
            const double *sA = s_stage + (size_t)(it % STAGES) * SM::STAGE_DOUBLES;
            const double *sB = sA + SM::A_DOUBLES;
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                const double *pa = sA + (ks * 8 + t) * LDP + row0 + g;
                const double a0 = pa[0], a1 = pa[8], a2 = pa[4 * LDP], a3 = pa[4 * LDP + 8];
                const double *pb = sB + (ks * 8 + t) * LDB + g;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const double b0 = pb[q * BK * LDB + h * 8], b1 = pb[q * BK * LDB + 4 * LDB + h * 8];
                        mma_16x8x8_f64(acc[q][h], a0, a1, a2, a3, b0, b1);
                    }
            }
            if (++kc == nkc) {
                // ---- fused epilogue for nu chunk vc -------------------------------------------------
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int slot = vc * NV + h * 8 + 2 * t + j;
                        const double *pe = panel + (long)slot * LDP;
                        double Rx = 0, Ry = 0, Rz = 0;
                        if (GIAO) { const int f = fidx[slot]; Rx = a.fR[f]; Ry = a.fR[a.nbf + f]; Rz = a.fR[2 * a.nbf + f]; }
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int row = rr ? rowB : rowA;
                            double *e = rr ? eB : eA;
                            const double px = rr ? rBx : rAx, py = rr ? rBy : rAy, pz = rr ? rBz : rAz;
                            const int ci = 2 * rr + j;
                            const double e0 = pe[row], e1 = pe[plane + row], e2 = pe[2 * plane + row], e3 = pe[3 * plane + row];
                            const double t0 = acc[0][h][ci] * e0;
                            e[12] += t0;
                            double zx = acc[1][h][ci], zy = acc[2][h][ci], zz = acc[3][h][ci];
                            if (GIAO) {
                                e[9] += Rx * t0; e[10] += Ry * t0; e[11] += Rz * t0;
                                const double yx = acc[NQ - 3][h][ci], yy = acc[NQ - 2][h][ci], yz = acc[NQ - 1][h][ci];
                                zx += py * yz - pz * yy;   // (r x Y')_x
                                zy += pz * yx - px * yz;
                                zz += px * yy - py * yx;
                            }
                            e[0] += zx * e1; e[1] += zx * e2; e[2] += zx * e3;   // b = x: m = x,y,z
                            e[3] += zy * e1; e[4] += zy * e2; e[5] += zy * e3;
                            e[6] += zz * e1; e[7] += zz * e2; e[8] += zz * e3;
                        }
                    }
                kc = 0; ++vc;
            }
        }
        cp_async_wait<0>();

        // ---- reduce over the 4 lanes of a quad (they hold different nu), finalise, store ----------
#pragma unroll
        for (int i = 0; i < 13; ++i) {
            eA[i] += __shfl_xor_sync(0xffffffffu, eA[i], 1); eA[i] += __shfl_xor_sync(0xffffffffu, eA[i], 2);
            eB[i] += __shfl_xor_sync(0xffffffffu, eB[i], 1); eB[i] += __shfl_xor_sync(0xffffffffu, eB[i], 2);
        }
        if (t < 2) {
            const bool v = t ? vB : vA;
            if (v) {
                const double *e = t ? eB : eA;
                const double px = t ? rBx : rAx, py = t ? rBy : rAy, pz = t ? rBz : rAz;
                double ct[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) ct[i] = e[i];
                if (GIAO) {   // + sum_d eps(b,m,d) V_d  at ct[m + 3b]
                    ct[0 + 3 * 1] -= e[11]; ct[0 + 3 * 2] += e[10];
                    ct[1 + 3 * 0] += e[11]; ct[1 + 3 * 2] -= e[9];
                    ct[2 + 3 * 0] -= e[10]; ct[2 + 3 * 1] += e[9];
                }
#pragma unroll
                for (int i = 0; i < 9; ++i) ct[i] = a.paramag ? 0.5 * ct[i] : 0.0;      // ZETA, jtensor.F90:209-223
                const double rho = e[12];
                const double d1 = a.diamag ? rho * (0.5 * px) : 0.0, d2 = a.diamag ? rho * (0.5 * py) : 0.0,
                             d3 = a.diamag ? rho * (0.5 * pz) : 0.0;                     // dpd, jtensor.F90:187,225-228
                ct[0 + 3 * 1] += d3; ct[0 + 3 * 2] -= d2;                                // jtensor.F90:230-235
                ct[1 + 3 * 0] -= d3; ct[1 + 3 * 2] += d1;
                ct[2 + 3 * 0] += d2; ct[2 + 3 * 1] -= d1;
                const long o = a.perm[td.pt0 + (t ? rowB : rowA)];
#pragma unroll
                for (int i = 0; i < 9; ++i) a.tens[9 * o + i] = ct[i];
                if (a.edens) a.edens[o] = rho;
            }
        }
    }
}

size_t jtensor_smem_bytes(bool giao) { return giao ? Smem<NQ_GIAO>::BYTES : Smem<NQ_NOGIAO>::BYTES; }

void launch_jtensor(const JtensorArgs &a, bool giao, int nsm, cudaStream_t s) {
    if (a.ntiles <= 0) return;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_jtensor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<NQ_GIAO>::BYTES);
        cudaFuncSetAttribute(k_jtensor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<NQ_NOGIAO>::BYTES);
        configured = true;
    }
    int grid = a.ntiles < nsm ? a.ntiles : nsm;
    if (giao) k_jtensor<true><<<grid, 256, Smem<NQ_GIAO>::BYTES, s>>>(a);
    else k_jtensor<false><<<grid, 256, Smem<NQ_NOGIAO>::BYTES, s>>>(a);
}

// ---------------------------------------------------------------------------------------------
// Contraction operands in the internal (per-atom radius-sorted) function order, row-major [mu][nu]:
// planes 0..3 = D, Px, Py, Pz (optionally alpha +/- beta), planes 4..6 = D (R_nu - R_mu)_d.
// src is the dens.f90 layout: element (a,b) at a + nbf*b.
__global__ void k_build_operand(double *__restrict__ out, int nbf, int ldb, long long plane_stride, const double *__restrict__ srcA,
                                const double *__restrict__ srcB, double signB, const int *__restrict__ f2user,
                                const double *__restrict__ fR, int giao) {
    const int nu = blockIdx.x * blockDim.x + threadIdx.x, mu = blockIdx.y;
    if (nu >= nbf) return;
    const long un = f2user[nu], um = f2user[mu];
    const long src = um + (long)nbf * un, nn = (long)nbf * nbf;
    const long dst = (long)mu * ldb + nu;
    double d0 = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double v = srcA[q * nn + src];
        if (srcB) v += signB * srcB[q * nn + src];
        out[q * plane_stride + dst] = v;
        if (q == 0) d0 = v;
    }
    if (giao) {
#pragma unroll
        for (int d = 0; d < 3; ++d) out[(4 + d) * plane_stride + dst] = d0 * (fR[d * nbf + nu] - fR[d * nbf + mu]);
    }
}
void launch_build_operand(double *out, int nbf, int ldb, long long plane_stride, const double *srcA, const double *srcB, double signB,
                          const int *f2user, const double *fR, bool giao, cudaStream_t s) {
    dim3 grid((nbf + 127) / 128, nbf);
    k_build_operand<<<grid, 128, 0, s>>>(out, nbf, ldb, plane_stride, srcA, srcB, signB, f2user, fR, giao ? 1 : 0);
}

}  // namespace gb
