// k_jtensor: the contraction  T(r) = f(Phi(r), dPhi(r); D, P_x, P_y, P_z)  for tiles of 128 points.
//
// Replaces `contract` of the reference (src/libgimic/jtensor.F90:148-237: 7 nbf x nbf GEMVs + 28 dot
// products per point) by one FP64 tensor-core GEMM per tile with a fused epilogue:
//
//   X_q[p, nu] = sum_mu Phi[p, mu] * B_q[mu, nu],   q = 0..3,  mu/nu over the tile's ACTIVE functions
//     B_0 = D,  B_1..3 = P_x,P_y,P_z
//   Y_d[p, nu] = sum_mu Phi[p, mu] D[mu, nu] (R_nu - R_mu)_d                          (d = x,y,z; GIAO only)
//
// and, with e_0 = Phi, e_m = dPhi/dr_m, y = r x (Y_x, Y_y, Y_z):
//
//   Tp(m,b) = sum_nu (X_{1+b} + y_b)[nu] e_m[nu]          (ppd + prsp1 + the d_m part of prsp2)
//   V_d     = sum_nu R_{nu,d} X_0[nu] e_0[nu]             (the (e_m x R) Phi part of prsp2, bfeval.f90:228-240)
//   rho     = sum_nu X_0[nu] e_0[nu]                      (diapam, jtensor.F90:168)
//   ct(m,b) = 1/2 [Tp(m,b) + sum_d eps(b,m,d) V_d]  +  eps(m,b,c) rho r_c / 2      (jtensor.F90:209-235)
//
// The identity behind Y (gauge-difference form of dendb/d2fvec, bfeval.f90:168-246):
//   -sum_nu dendb_b[nu] e_m[nu] + sum_nu denbf[nu] e_m[nu] g_{A(nu),b}
//        = sum_nu e_m[nu] sum_{c,d} eps(b,c,d) r_c sum_mu Phi_mu D[mu,nu] (R_{nu,d} - R_{mu,d}).
// Y costs NO extra GEMM planes: R_mu is constant over the functions of one atom and the K slots are grouped by atom
// (each atom's run padded to whole k4 steps), so with C_A[nu] = the running D accumulator after atom A's last K step
//   Y_d = (R_nu - c)_d X_0 - Z_d,   Z_d = sum_A (R_A - c)_d (C_A - C_{A-1}) = sum_{A<n} C_A (R_A - R_{A+1})_d + C_n (R_n - c)_d
// (Abel summation; c = tile centre, all differences are O(screening radius) so nothing cancels for molecules far from the
// origin).  The consumer "taps" the D accumulator at every atom boundary: 3 DFMA per accumulator element instead of the
// 3 x K_A MMA columns the planes D(R_nu - R_mu)_d used to cost.  7 GEMM planes -> 4.
// Screened functions are exact zeros in the reference, so restricting mu, nu to the tile's active
// set changes nothing but the order of summation.
//
// Mapping: 8 consumer warps + 4 producer warps; consumer warp w owns rows 16w..16w+15 of the tile and all 4x2 n8 tiles of
// the current 16-wide nu chunk (32 + 24 fp64 accumulators / thread), so the per-point epilogue sums stay
// in registers for the whole tile and need only a 4-lane shuffle reduction at the end.
// mma.sync.m16n8k4.f64 lowers to 2 DMMA.8x8x4 on sm_100a (there is no FP64 tcgen05 kind).
#include "kernels.cuh"
#include <cstdlib>
#include <type_traits>
#ifndef KS_UNROLL
#define KS_UNROLL 8
#endif
constexpr int KSU = KS_UNROLL;   // unroll factor of the K-step loop (measured on the 1.2 M-point chunk: 2 -> 76.4 ms, 4 -> 74.4 ms, 8 -> 73.8 ms)

namespace gb {

// m16n8k4 = two independent DMMA.8x8x4 (rows 0-7 / 8-15).  The k8 form expands to four DMMAs of which the second pair
// depends on the first through the accumulator, back to back (measured: 37% fixed-latency "wait" stalls); with k4 the
// dependent pair is a separate instruction 14 accumulator tiles later.
__device__ __forceinline__ void mma_16x8x4_f64(double (&c)[4], double a0, double a1, double b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a0), "d"(a1), "d"(b0));
}
// ---- async-copy / mbarrier primitives (sm_90+ PTX; SASS: LDGSTS, UBLKCP, SYNCS) ---------------------
// L2_HINTS (A/B builds, tools/gpu_r02_z.sh): 1 = the Phi panels a CTA streams once per nu chunk are loaded with an evict_last L2 policy;
// 2 = additionally the density gathers with evict_first.  0 (the shipped build) = no hints.
#ifndef L2_HINTS
#define L2_HINTS 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async_16(uint32_t smem, const void *gmem) {
#if L2_HINTS >= 2
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "l"(l2_policy_evict_first()));
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem), "l"(gmem));
#endif
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t mbar) {   // arrive when this thread's prior cp.async land
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(mbar));
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(mbar), "r"(parity) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(uint32_t smem_dst, const void *gmem, uint32_t bytes, uint32_t mbar) {
#if L2_HINTS >= 1
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_dst), "l"(gmem), "r"(bytes), "r"(mbar), "l"(l2_policy_evict_last()) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gmem), "r"(bytes), "r"(mbar) : "memory");
#endif
}

constexpr int ROWLD = 17;   // doubles per row-table entry: 13 sums + 3 coordinates, padded to an odd stride (bank-conflict-free)
constexpr int NCONSUMER_WARPS = 8;
constexpr int NPRODUCER_WARPS = 4;   // one warpgroup, so that setmaxnreg can hand its registers to the consumers
constexpr int NTHREADS = (NCONSUMER_WARPS + NPRODUCER_WARPS) * 32;
// setmaxnreg only REDISTRIBUTES the registers the CTA was launched with: a request beyond that pool never completes and the kernel hangs
// (measured twice: 8 consumer warps at 240/32 in round 1, a 16-consumer-warp variant at 112/40 in round 2 -- that variant, 16 x 8 column
// tiles per warp at 112 registers, was measured 10 % slower, its epilogue spilling 444 B per thread, and removed:
// profiles/r02_ab_ncw_overlap.json).
#ifndef DEFAULT_EPI
#define DEFAULT_EPI 1
#endif
constexpr int CONSUMER_REGS = 232, PRODUCER_REGS = 40;   // (8*232 + 4*40) * 32 = 64512 < 65536.  Do NOT use the whole file: with
                                                          // 240/32 (= 65536) setmaxnreg.inc never succeeds and the kernel hangs (measured)
constexpr int ATAB_MAX = 1024;      // active atoms of a tile staged in shared memory (more: the taps read the table from global memory)
constexpr int KMASK_WORDS = 512;    // atom-end bits for up to 16384 K steps = 65536 slots

template <int NPP_, int NV_, int LDB_, int NTAB_ = 1>
struct SmemT {
    static constexpr int NTAB = NTAB_;                       // per-row epilogue tables (one per group of consumer warps that shares rows)
    static constexpr int A_DOUBLES = BK * LDP;
    static constexpr int NPP = NPP_;                         // pair-planes: tensor path (D,Px) (Py,Pz); J path (D, P.B)
    static constexpr int NVC = NV_;                          // nu columns per accumulator chunk
    static constexpr int LDB = LDB_;                         // smem row stride (doubles) of a pair-plane B tile: NVC x 2 + 4 pad
    static constexpr int PP_DOUBLES = BK * LDB;              // one pair-plane of a stage: [k][NVC nu x 2 + pad]
    static constexpr int B_DOUBLES = NPP * PP_DOUBLES;
    static constexpr int STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
    static constexpr size_t ROW_OFF = (size_t)STAGES * STAGE_DOUBLES * 8;        // per-row epilogue sums + point coordinates
    static constexpr size_t ATAB_OFF = ROW_OFF + (size_t)NTAB * (MT * ROWLD + 4) * 8;   // (+4: tile centre)          // GIAO tap weights of the tile's atoms [ATAB_MAX][3]
    static constexpr size_t KMASK_OFF = ATAB_OFF + (size_t)ATAB_MAX * 3 * 8;      // bit k4: K step k4 is the last of an atom
    static constexpr size_t BAR_OFF = KMASK_OFF + (size_t)KMASK_WORDS * 4;
    static constexpr size_t BYTES = BAR_OFF + 2 * STAGES * 8 + 16;
};
constexpr int NVJ = 32, LDB2J = 68;                // J path: 4 n8 tiles per chunk; 68 doubles = 34 16-byte units = 2 (mod 8) -> conflict-free LDS.128
using Smem = SmemT<(NQ + 1) / 2, NV, LDB2>;        // tensor path
using SmemJ = SmemT<1, NVJ, LDB2J>;                // J = T.B path: same 16 KB of B per stage, twice the columns per A fragment

// ---- output helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ long out_row(const JtensorArgs &a, long p) { return a.perm ? (long)a.perm[p] : p - a.out_base; }
__device__ __forceinline__ void store_zero(const JtensorArgs &a, long o) {
    if (a.tens) for (int i = 0; i < 9; ++i) a.tens[9 * o + i] = 0.0;
    if (a.jvec) for (int i = 0; i < 3; ++i) a.jvec[3 * o + i] = 0.0;
    if (a.jmod) a.jmod[o] = 0.0;
    if (a.acid) a.acid[o] = 0.0;
    if (a.edens) a.edens[o] = 0.0;
}
// signed |J| (jmod2_vtkplot, jfield.f90:446-489): the sign of (B x (r - (B.r) B)) . J; same statements as k_fields / k_jmod
__device__ __forceinline__ double signed_modulus(double vx, double vy, double vz, double cx, double cy, double cz, double bx, double by, double bz) {
    double jm = sqrt(vx * vx + vy * vy + vz * vz);
    const double d = bx * cx + by * cy + bz * cz;
    cx -= d * bx; cy -= d * by; cz -= d * bz;
    const double nx = by * cz - bz * cy, ny = bz * cx - bx * cz, nz = bx * cy - by * cx;   // cross_product(mag, coord)
    if (nx * vx + ny * vy + nz * vz < 0.0) jm = -1.0 * jm;
    return jm;
}
// get_acid (acid.f90:9-45) with the reference's DP33 = 0.3333333 (globals.f90:62)
__device__ __forceinline__ double acid_of(const double (&t)[9]) {
    const double xxmyy = (t[0] - t[4]) * (t[0] - t[4]), yymzz = (t[4] - t[8]) * (t[4] - t[8]), zzmxx = (t[8] - t[0]) * (t[8] - t[0]);
    const double xypyx = (t[3] + t[1]) * (t[3] + t[1]), xzpzx = (t[6] + t[2]) * (t[6] + t[2]), yzpzy = (t[7] + t[5]) * (t[7] + t[5]);
    return 0.3333333 * (xxmyy + yymzz + zzmxx) + 0.5 * (xypyx + xzpzx + yzpzy);
}

// From the 13 row sums of a point (Tp(m,b) at [m+3b], V_d at [9+d], rho at [12]) to the tensor and the fields derived from it
// (jtensor.F90:209-235; jfield.f90:167-184, 446-489; acid.f90).  Shared by the contraction's epilogue and k_slice_reduce.
template <bool GIAO>
__device__ __forceinline__ void finalise_store_tensor(const JtensorArgs &a, const double (&e)[13], double px, double py, double pz, long o) {
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
    if (a.tens) {
#pragma unroll
        for (int i = 0; i < 9; ++i) a.tens[9 * o + i] = ct[i];
    }
    if (a.edens) a.edens[o] = rho;
    // derived fields straight from the registers (what the separate k_fields pass computes from the stored tensor)
    if (a.jvec || a.jmod) {
        const double bx = a.B[0], by = a.B[1], bz = a.B[2];
        const double vx = ct[0] * bx + ct[3] * by + ct[6] * bz, vy = ct[1] * bx + ct[4] * by + ct[7] * bz,
                     vz = ct[2] * bx + ct[5] * by + ct[8] * bz;                      // matmul(reshape(tens,(3,3)), b), jfield.f90:167-184
        if (a.jvec) { a.jvec[3 * o] = vx; a.jvec[3 * o + 1] = vy; a.jvec[3 * o + 2] = vz; }
        if (a.jmod) a.jmod[o] = signed_modulus(vx, vy, vz, px, py, pz, bx, by, bz);
    }
    if (a.acid) a.acid[o] = acid_of(ct);
}
// J path: T_m = sum_b Tp(m,b) B_b at [m], V_d at [3+d], rho at [6]
__device__ __forceinline__ void finalise_store_j(const JtensorArgs &a, const double (&e)[13], double px, double py, double pz, long o) {
    const double bx = a.B[0], by = a.B[1], bz = a.B[2];
    // J_m = sum_b ct(m,b) B_b:  1/2 [T_m + (V x B)_m]  +  1/2 rho (B x r)_m   (jtensor.F90:209-235 contracted with B)
    double jx = e[0] + (e[4] * bz - e[5] * by), jy = e[1] + (e[5] * bx - e[3] * bz), jz = e[2] + (e[3] * by - e[4] * bx);
    jx = a.paramag ? 0.5 * jx : 0.0; jy = a.paramag ? 0.5 * jy : 0.0; jz = a.paramag ? 0.5 * jz : 0.0;
    const double rho = e[6];
    if (a.diamag) {
        jx += 0.5 * rho * (by * pz - bz * py); jy += 0.5 * rho * (bz * px - bx * pz); jz += 0.5 * rho * (bx * py - by * px);
    }
    if (a.jvec) { a.jvec[3 * o] = jx; a.jvec[3 * o + 1] = jy; a.jvec[3 * o + 2] = jz; }
    if (a.jmod) a.jmod[o] = signed_modulus(jx, jy, jz, px, py, pz, bx, by, bz);
    if (a.edens) a.edens[o] = rho;
}

// Keep a loop-invariant value in a register: without this ptxas re-derives the lane id and every shared-memory address from the special
// registers (S2R SR_TID.X, S2R/S2UR SR_CgaCtaId + LEA) at every pipeline stage -- three dependent ~40-cycle chains per stage that both
// warps of a scheduler run at the same time (2.4 % of the consumer's samples in profiles/r02_ncu_jtensor_e_source_regions.txt).
__device__ __forceinline__ void keep_in_register(uint32_t &v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void keep_in_register(int &v) { asm volatile("" : "+r"(v)); }

__device__ __forceinline__ void load_tap_weights(bool tab_sm, const double *s_atab, const double *gtab, int ia, double &x, double &y, double &z) {
    if (tab_sm) { const double *p = s_atab + 3 * ia; x = p[0]; y = p[1]; z = p[2]; }          // [ATAB_MAX][3] in shared memory
    else { const double *p = gtab + 4 * ia; x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2); }   // TileAtom = 4 doubles
}

// Tile bookkeeping shared by both roles: every thread of the CTA calls this once per tile (two CTA barriers).
__device__ __forceinline__ int next_tile(const JtensorArgs &a, int *s_tile) {
    __syncthreads();                                  // everybody is done with the previous tile (and with *s_tile)
    if (threadIdx.x == 0) *s_tile = atomicAdd(a.counter, 1);
    __syncthreads();
    return *s_tile;
}

template <class SM, int NCW>
__device__ __forceinline__ void producer_role(const JtensorArgs &a, uint32_t s_base, uint32_t bar_full, uint32_t bar_empty, int *s_tile) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t git = 0;
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        if (td.nact == 0) continue;
        const int nact = td.nact, nn = td.nn;
        const int nkc = (nact + BK - 1) / BK, nvc = (nn + SM::NVC - 1) / SM::NVC;   // nn is a multiple of 8: the last nu chunk may hold 8 columns
        const uint32_t NIT = (uint32_t)nkc * nvc;
        const double *panel = a.panel_pool + td.panel_off;
        const int *fidx = a.fidx_pool + td.fidx_off, *nlist = fidx + nact;
        {
        // ===================================== producer warps =====================================
        const int pw = warp - NCW;
        constexpr int LPW = 32 / SM::NVC > 0 ? 32 / SM::NVC : 1;    // k rows covered by one warp per pass (2 for 16 columns, 1 for 32)
        const int ldn = lane % SM::NVC, ldk0 = lane / SM::NVC + LPW * pw;    // this lane gathers nu slot ldn, k rows ldk0, ldk0 + LPW*4, ...
        int kc = 0, vc = 0;
        long nu = fidx[nlist[min(ldn, nn - 1)]];
        for (uint32_t itl = 0; itl < NIT; ++itl) {
            const uint32_t gi = git + itl, s = gi % STAGES, ph = (gi / STAGES) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const int kcnt = min(BK, nact - kc * BK);
            const uint32_t sA = s_base + (uint32_t)(s * SM::STAGE_DOUBLES * 8), sB = sA + SM::A_DOUBLES * 8;
            if (pw == 0 && lane == 0) {
                mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(kcnt * LDP * 8));
                tma_bulk_g2s(sA, panel + (long)kc * BK * LDP, (uint32_t)(kcnt * LDP * 8), bar_full + 8 * s);
            }
            const bool nu_ok = vc * SM::NVC + ldn < nn;
            const double *srcB = a.Bop + 2 * nu;
            const uint32_t dstB = sB + (uint32_t)(ldn * 16);
#pragma unroll 4
            for (int k = ldk0; k < (nu_ok ? kcnt : 0); k += LPW * NPRODUCER_WARPS) {
                const long mu = fidx[kc * BK + k];
                const double *src = srcB + 2 * mu * a.ldb;
                const uint32_t dst = dstB + (uint32_t)(k * SM::LDB * 8);
#pragma unroll
                for (int pp = 0; pp < SM::NPP; ++pp) cp_async_16(dst + (uint32_t)(pp * SM::PP_DOUBLES * 8), src + pp * a.plane_stride);
            }
            cp_async_arrive_noinc(bar_full + 8 * s);
            if (++kc == nkc) { kc = 0; ++vc; if (vc < nvc) nu = fidx[nlist[min(vc * SM::NVC + ldn, nn - 1)]]; }
        }
        }
        git += NIT;
    }
}

template <bool GIAO>
__device__ __forceinline__ void consumer_role(const JtensorArgs &a, const double *s_stage, double *s_rows, double *s_atab, uint32_t *s_kmask,
                                              uint32_t bar_full, uint32_t bar_empty, int *s_tile) {
    using SM = Smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = warp * 16;
    uint32_t git = 0;
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        const int rowA = row0 + g, rowB = row0 + g + 8;
        const bool vA = rowA < td.npts, vB = rowB < td.npts;
        if (td.nact == 0) {   // nothing within screening range: the reference returns exact zeros
            if (t < 2 && (t ? vB : vA)) store_zero(a, out_row(a, td.pt0 + (t ? rowB : rowA)));
            continue;
        }
        const int nact = td.nact, nn = td.nn;
        const int nkc = (nact + BK - 1) / BK, nvc = (nn + NV - 1) / NV;   // nn is a multiple of 8: the last nu chunk may hold 8 columns
        const uint32_t NIT = (uint32_t)nkc * nvc;
        const double *panel = a.panel_pool + td.panel_off;
        const long plane = (long)nact * LDP;
        const int *fidx = a.fidx_pool + td.fidx_off, *nlist = fidx + nact;
        {
        // ===================================== consumer warps =====================================
        // Row table: lane t=0 of a quad owns row A, lane t=1 row B.  [0..12] running sums Tp(m,b) at [m+3b], V_d at [9+d],
        // rho at [12]; [13..15] the point's absolute coordinates (as r enters jtensor.F90:112 and bfeval.f90:168-189).
        double *rowA_s = s_rows + rowA * ROWLD, *rowB_s = s_rows + rowB * ROWLD;
        if (t < 2) {
            double *rs = t ? rowB_s : rowA_s;
            const long p = td.pt0 + ((t ? vB : vA) ? (t ? rowB : rowA) : 0);
#pragma unroll
            for (int i = 0; i < 13; ++i) rs[i] = 0.0;
            rs[13] = a.rsx[p]; rs[14] = a.rsy[p]; rs[15] = a.rsz[p];
        }
        if (GIAO && threadIdx.x == 0) {   // tile centre, same expression as k_basis; read in the epilogue (after the bar.sync below)
            const TileGeo tg = a.geo[td.geo];
            s_rows[MT * ROWLD] = 0.5 * (tg.lox + tg.hix); s_rows[MT * ROWLD + 1] = 0.5 * (tg.loy + tg.hiy); s_rows[MT * ROWLD + 2] = 0.5 * (tg.loz + tg.hiz);
        }
        __syncwarp();
        double acc[NQ][2][4];
        double zac[3][2][4];                                        // Z_d (GIAO taps)
        const int nruns = td.nruns;
        // tap weights (dx,dy,dz) of run ia: shared memory (LDS; a generic pointer would cost a generic load per weight), or the
        // global table when the tile has more active atoms than the shared copy holds
        const bool tab_sm = nruns <= ATAB_MAX;
        const double *gtab = reinterpret_cast<const double *>(a.atab_pool + td.atab_off);
        double curx = 0, cury = 0, curz = 0;
        int ia = 0;
        int eslot[2][2];                                            // epilogue: K slot of this thread's 4 nu columns
        if (GIAO) {
            // stage the tile's atom table: weights to shared memory, atom ends as one bit per K step
            const double2 *atab = reinterpret_cast<const double2 *>(a.atab_pool + td.atab_off);   // TileAtom = 2 x double2
            const int ctid = threadIdx.x, nwords = (nact / 4 + 31) / 32;
            for (int w = ctid; w < nwords; w += NCONSUMER_WARPS * 32) s_kmask[w] = 0u;
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
            for (int r = ctid; r < nruns; r += NCONSUMER_WARPS * 32) {
                const double2 t0 = __ldg(atab + 2 * r), t1 = __ldg(atab + 2 * r + 1);
                if (r < ATAB_MAX) { s_atab[3 * r] = t0.x; s_atab[3 * r + 1] = t0.y; s_atab[3 * r + 2] = t1.x; }
                const int e = __double2loint(t1.y) - 1;
                atomicOr(&s_kmask[e >> 5], 1u << (e & 31));
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
        }
        int kc = 0, vc = 0;
        for (uint32_t it = 0; it < NIT; ++it) {
            const uint32_t gi = git + it, s = gi % STAGES, ph = (gi / STAGES) & 1;
            if (kc == 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[q][h][i] = 0.0;
                if (GIAO) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int i = 0; i < 4; ++i) zac[d][h][i] = 0.0;
                    ia = 0;
                    load_tap_weights(tab_sm, s_atab, gtab, 0, curx, cury, curz);
                }
                // K slots the epilogue of this chunk needs (loaded a whole K sweep early)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 2; ++j) eslot[h][j] = nlist[min(vc * NV + h * 8 + 2 * t + j, nn - 1)];
            }
            const int k4base = kc * (BK / 4);                                   // BK/4 = 8 K steps per stage: their bits share a word
            const uint32_t m8 = GIAO ? (s_kmask[k4base >> 5] >> (k4base & 31)) : 0u;
            const int nks = min(BK, nact - kc * BK) / 4;
            const bool h1 = vc * NV + 8 < nn;                       // second n8 tile of this chunk holds real columns
            const double *sA = s_stage + (size_t)s * SM::STAGE_DOUBLES;
            const double *sB = sA + SM::A_DOUBLES;
            mbar_wait(bar_full + 8 * s, ph);
#pragma unroll KSU
            for (int ks = 0; ks < nks; ++ks) {
                // fragments (m16n8k4.f64): a0 = A[row g][k t], a1 = A[row g+8][k t]; b0 = B[k t][n g]
                const double *pa = sA + (ks * 4 + t) * LDP + row0 + g;
                const double a0 = pa[0], a1 = pa[8];
                const double2 *pb = reinterpret_cast<const double2 *>(sB + (ks * 4 + t) * LDB2) + g;   // one LDS.128 = both planes of a pair
#pragma unroll
                for (int pp = 0; pp < SM::NPP; ++pp)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1 && !h1) continue;
                        const double2 b = pb[pp * (SM::PP_DOUBLES / 2) + h * 8];
                        mma_16x8x4_f64(acc[2 * pp][h], a0, a1, b.x);
                        mma_16x8x4_f64(acc[2 * pp + 1][h], a0, a1, b.y);
                    }
                if (GIAO && ((m8 >> ks) & 1u)) {
                    // last K step of an atom: Z_d += C_A * (R_A - R_next)_d  (see the header; C_A = acc[0] right now)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double cv = acc[0][h][i];
                            zac[0][h][i] = fma(curx, cv, zac[0][h][i]);
                            zac[1][h][i] = fma(cury, cv, zac[1][h][i]);
                            zac[2][h][i] = fma(curz, cv, zac[2][h][i]);
                        }
                    ia = min(ia + 1, nruns - 1);
                    load_tap_weights(tab_sm, s_atab, gtab, ia, curx, cury, curz);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);   // slot may be refilled
            if (++kc == nkc) {
                // ---- fused epilogue for nu chunk vc ---------------------------------------------
                double eA[13], eB[13];
#pragma unroll
                for (int i = 0; i < 13; ++i) { eA[i] = 0.0; eB[i] = 0.0; }
                const double rAx = rowA_s[13], rAy = rowA_s[14], rAz = rowA_s[15];
                const double rBx = rowB_s[13], rBy = rowB_s[14], rBz = rowB_s[15];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (h == 1 && !h1) continue;
                        const int slot = eslot[h][j];                         // K slot (= panel row) of this nu column
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
                                const double x0 = acc[0][h][ci];
                                const double cenx = s_rows[MT * ROWLD], ceny = s_rows[MT * ROWLD + 1], cenz = s_rows[MT * ROWLD + 2];
                                const double yx = (Rx - cenx) * x0 - zac[0][h][ci], yy = (Ry - ceny) * x0 - zac[1][h][ci],
                                             yz = (Rz - cenz) * x0 - zac[2][h][ci];
                                zx += py * yz - pz * yy;   // (r x Y')_x
                                zy += pz * yx - px * yz;
                                zz += px * yy - py * yx;
                            }
                            e[0] += zx * e1; e[1] += zx * e2; e[2] += zx * e3;   // b = x: m = x,y,z
                            e[3] += zy * e1; e[4] += zy * e2; e[5] += zy * e3;
                            e[6] += zz * e1; e[7] += zz * e2; e[8] += zz * e3;
                        }
                    }
                // reduce over the 4 lanes of the quad (they hold different nu) and add into the row table
#pragma unroll
                for (int i = 0; i < 13; ++i) {
                    eA[i] += __shfl_xor_sync(0xffffffffu, eA[i], 1); eA[i] += __shfl_xor_sync(0xffffffffu, eA[i], 2);
                    eB[i] += __shfl_xor_sync(0xffffffffu, eB[i], 1); eB[i] += __shfl_xor_sync(0xffffffffu, eB[i], 2);
                }
                if (t < 2) {
                    double *rs = t ? rowB_s : rowA_s;
#pragma unroll
                    for (int i = 0; i < 13; ++i) rs[i] += t ? eB[i] : eA[i];
                }
                kc = 0; ++vc;
            }
        }

        // ---- finalise and store ---------------------------------------------------------------------
        if (t < 2) {
            const bool v = t ? vB : vA;
            if (v) {
                const double *e = t ? rowB_s : rowA_s;
                const double px = e[13], py = e[14], pz = e[15];
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
                const long o = out_row(a, td.pt0 + (t ? rowB : rowA));
                if (a.tens) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) a.tens[9 * o + i] = ct[i];
                }
                if (a.edens) a.edens[o] = rho;
                // derived fields straight from the registers (what the separate k_fields pass computes from the stored tensor)
                if (a.jvec || a.jmod) {
                    const double bx = a.B[0], by = a.B[1], bz = a.B[2];
                    const double vx = ct[0] * bx + ct[3] * by + ct[6] * bz, vy = ct[1] * bx + ct[4] * by + ct[7] * bz,
                                 vz = ct[2] * bx + ct[5] * by + ct[8] * bz;                      // matmul(reshape(tens,(3,3)), b), jfield.f90:167-184
                    if (a.jvec) { a.jvec[3 * o] = vx; a.jvec[3 * o + 1] = vy; a.jvec[3 * o + 2] = vz; }
                    if (a.jmod) a.jmod[o] = signed_modulus(vx, vy, vz, px, py, pz, bx, by, bz);
                }
                if (a.acid) a.acid[o] = acid_of(ct);
            }
        }
        }
        git += NIT;
    }
}

template <bool GIAO>
__device__ __forceinline__ void consumer_role_j(const JtensorArgs &a, const double *s_stage, double *s_rows, double *s_atab, uint32_t *s_kmask,
                                              uint32_t bar_full, uint32_t bar_empty, int *s_tile) {
    using SM = SmemJ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = warp * 16;
    uint32_t git = 0;
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        const int rowA = row0 + g, rowB = row0 + g + 8;
        const bool vA = rowA < td.npts, vB = rowB < td.npts;
        if (td.nact == 0) {   // nothing within screening range: the reference returns exact zeros
            if (t < 2 && (t ? vB : vA)) store_zero(a, out_row(a, td.pt0 + (t ? rowB : rowA)));
            continue;
        }
        const int nact = td.nact, nn = td.nn;
        const int nkc = (nact + BK - 1) / BK, nvc = (nn + NVJ - 1) / NVJ;   // nn is a multiple of 8: the last nu chunk may hold 8 columns
        const uint32_t NIT = (uint32_t)nkc * nvc;
        const double *panel = a.panel_pool + td.panel_off;
        const long plane = (long)nact * LDP;
        const int *fidx = a.fidx_pool + td.fidx_off, *nlist = fidx + nact;
        {
        // ===================================== consumer warps =====================================
        // Row table: lane t=0 of a quad owns row A, lane t=1 row B.  [0..2] running sums T_m = sum_b Tp(m,b) B_b, [3..5] V_d,
        // [6] rho; [13..15] the point's absolute coordinates (as r enters jtensor.F90:112 and bfeval.f90:168-189).
        double *rowA_s = s_rows + rowA * ROWLD, *rowB_s = s_rows + rowB * ROWLD;
        if (t < 2) {
            double *rs = t ? rowB_s : rowA_s;
            const long p = td.pt0 + ((t ? vB : vA) ? (t ? rowB : rowA) : 0);
#pragma unroll
            for (int i = 0; i < 13; ++i) rs[i] = 0.0;
            rs[13] = a.rsx[p]; rs[14] = a.rsy[p]; rs[15] = a.rsz[p];
        }
        if (GIAO && threadIdx.x == 0) {   // tile centre, same expression as k_basis; read in the epilogue (after the bar.sync below)
            const TileGeo tg = a.geo[td.geo];
            s_rows[MT * ROWLD] = 0.5 * (tg.lox + tg.hix); s_rows[MT * ROWLD + 1] = 0.5 * (tg.loy + tg.hiy); s_rows[MT * ROWLD + 2] = 0.5 * (tg.loz + tg.hiz);
        }
        __syncwarp();
        constexpr int NH = NVJ / 8;                                 // n8 tiles per chunk
        double acc[2][NH][4];                                       // planes D and P.B
        double zac[NH][4];                                          // S = (B x r) . Z (GIAO taps with per-row weights)
        // w = B x r of this thread's two rows: B.(r x Y) = Y.(B x r)
        double wAx = 0, wAy = 0, wAz = 0, wBx = 0, wBy = 0, wBz = 0;
        if (GIAO) {
            const double bx = a.B[0], by = a.B[1], bz = a.B[2];
            const double ax = rowA_s[13], ay = rowA_s[14], az = rowA_s[15], cx = rowB_s[13], cy = rowB_s[14], cz = rowB_s[15];
            wAx = by * az - bz * ay; wAy = bz * ax - bx * az; wAz = bx * ay - by * ax;
            wBx = by * cz - bz * cy; wBy = bz * cx - bx * cz; wBz = bx * cy - by * cx;
        }
        const int nruns = td.nruns;
        // tap weights (dx,dy,dz) of run ia: shared memory (LDS; a generic pointer would cost a generic load per weight), or the
        // global table when the tile has more active atoms than the shared copy holds
        const bool tab_sm = nruns <= ATAB_MAX;
        const double *gtab = reinterpret_cast<const double *>(a.atab_pool + td.atab_off);
        double curx = 0, cury = 0, curz = 0;
        int ia = 0;
        int eslot[NVJ / 8][2];                                      // epilogue: K slot of this thread's nu columns
        if (GIAO) {
            // stage the tile's atom table: weights to shared memory, atom ends as one bit per K step
            const double2 *atab = reinterpret_cast<const double2 *>(a.atab_pool + td.atab_off);   // TileAtom = 2 x double2
            const int ctid = threadIdx.x, nwords = (nact / 4 + 31) / 32;
            for (int w = ctid; w < nwords; w += NCONSUMER_WARPS * 32) s_kmask[w] = 0u;
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
            for (int r = ctid; r < nruns; r += NCONSUMER_WARPS * 32) {
                const double2 t0 = __ldg(atab + 2 * r), t1 = __ldg(atab + 2 * r + 1);
                if (r < ATAB_MAX) { s_atab[3 * r] = t0.x; s_atab[3 * r + 1] = t0.y; s_atab[3 * r + 2] = t1.x; }
                const int e = __double2loint(t1.y) - 1;
                atomicOr(&s_kmask[e >> 5], 1u << (e & 31));
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
        }
        int kc = 0, vc = 0;
        for (uint32_t it = 0; it < NIT; ++it) {
            const uint32_t gi = git + it, s = gi % STAGES, ph = (gi / STAGES) & 1;
            if (kc == 0) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int h = 0; h < NH; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[q][h][i] = 0.0;
                if (GIAO) {
#pragma unroll
                    for (int h = 0; h < NH; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) zac[h][i] = 0.0;
                    ia = 0;
                    load_tap_weights(tab_sm, s_atab, gtab, 0, curx, cury, curz);
                }
                // K slots the epilogue of this chunk needs (loaded a whole K sweep early)
#pragma unroll
                for (int h = 0; h < NH; ++h)
#pragma unroll
                    for (int j = 0; j < 2; ++j) eslot[h][j] = nlist[min(vc * NVJ + h * 8 + 2 * t + j, nn - 1)];
            }
            const int k4base = kc * (BK / 4);                                   // BK/4 = 8 K steps per stage: their bits share a word
            const uint32_t m8 = GIAO ? (s_kmask[k4base >> 5] >> (k4base & 31)) : 0u;
            const int nks = min(BK, nact - kc * BK) / 4;
            const int nh = min(NH, (nn - vc * NVJ) / 8);            // n8 tiles of this chunk that hold real columns (nn is a multiple of 8)
            const double *sA = s_stage + (size_t)s * SM::STAGE_DOUBLES;
            const double *sB = sA + SM::A_DOUBLES;
            mbar_wait(bar_full + 8 * s, ph);
            // fragments (m16n8k4.f64): a0 = A[row g][k t], a1 = A[row g+8][k t]; b = B[k t][n g] (one LDS.128 = both planes).
            // Software-pipelined by hand: the fragments of step ks+1 are loaded before the MMAs of step ks (the tap branch between
            // the steps keeps the compiler from doing it; ncu showed every DMMA waiting on its own LDS, 29 % short-scoreboard stalls)
            const double *pa0 = sA + t * LDP + row0 + g;
            const double2 *pb0 = reinterpret_cast<const double2 *>(sB + t * LDB2J) + g;
            double a0 = pa0[0], a1 = pa0[8];
            double2 bf[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) bf[h] = pb0[h * 8];
#pragma unroll KSU
            for (int ks = 0; ks < nks; ++ks) {
                double na0 = 0, na1 = 0;
                double2 nb[NH];
                if (ks + 1 < nks) {
                    const double *pa = pa0 + (ks + 1) * 4 * LDP;
                    na0 = pa[0]; na1 = pa[8];
                    const double2 *pb = pb0 + (ks + 1) * 4 * (LDB2J / 2);
#pragma unroll
                    for (int h = 0; h < NH; ++h) nb[h] = pb[h * 8];
                }
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    if (h >= nh) continue;
                    mma_16x8x4_f64(acc[0][h], a0, a1, bf[h].x);
                    mma_16x8x4_f64(acc[1][h], a0, a1, bf[h].y);
                }
                a0 = na0; a1 = na1;
#pragma unroll
                for (int h = 0; h < NH; ++h) bf[h] = nb[h];
                if (GIAO && ((m8 >> ks) & 1u)) {
                    // last K step of an atom: S += C_A * ((B x r) . (R_A - R_next)), one weight per row (see the header)
                    const double oA = wAx * curx + wAy * cury + wAz * curz, oB = wBx * curx + wBy * cury + wBz * curz;
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        zac[h][0] = fma(oA, acc[0][h][0], zac[h][0]); zac[h][1] = fma(oA, acc[0][h][1], zac[h][1]);
                        zac[h][2] = fma(oB, acc[0][h][2], zac[h][2]); zac[h][3] = fma(oB, acc[0][h][3], zac[h][3]);
                    }
                    ia = min(ia + 1, nruns - 1);
                    load_tap_weights(tab_sm, s_atab, gtab, ia, curx, cury, curz);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);   // slot may be refilled
            if (++kc == nkc) {
                // ---- fused epilogue for nu chunk vc ---------------------------------------------
                double eA[7], eB[7];
#pragma unroll
                for (int i = 0; i < 7; ++i) { eA[i] = 0.0; eB[i] = 0.0; }
#pragma unroll
                for (int h = 0; h < NH; ++h)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (h >= nh) continue;
                        const int slot = eslot[h][j];                         // K slot (= panel row) of this nu column
                        const double *pe = panel + (long)slot * LDP;
                        double Rx = 0, Ry = 0, Rz = 0;
                        if (GIAO) { const int f = fidx[slot]; Rx = a.fR[f]; Ry = a.fR[a.nbf + f]; Rz = a.fR[2 * a.nbf + f]; }
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int row = rr ? rowB : rowA;
                            double *e = rr ? eB : eA;
                            const int ci = 2 * rr + j;
                            const double e0 = pe[row], e1 = pe[plane + row], e2 = pe[2 * plane + row], e3 = pe[3 * plane + row];
                            const double x0 = acc[0][h][ci];
                            const double t0 = x0 * e0;
                            e[6] += t0;
                            double z = acc[1][h][ci];
                            if (GIAO) {
                                e[3] += Rx * t0; e[4] += Ry * t0; e[5] += Rz * t0;
                                const double cenx = s_rows[MT * ROWLD], ceny = s_rows[MT * ROWLD + 1], cenz = s_rows[MT * ROWLD + 2];
                                const double wx = rr ? wBx : wAx, wy = rr ? wBy : wAy, wz = rr ? wBz : wAz;
                                z += (wx * (Rx - cenx) + wy * (Ry - ceny) + wz * (Rz - cenz)) * x0 - zac[h][ci];   // (B x r) . Y
                            }
                            e[0] += z * e1; e[1] += z * e2; e[2] += z * e3;
                        }
                    }
                // reduce over the 4 lanes of the quad (they hold different nu) and add into the row table
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    eA[i] += __shfl_xor_sync(0xffffffffu, eA[i], 1); eA[i] += __shfl_xor_sync(0xffffffffu, eA[i], 2);
                    eB[i] += __shfl_xor_sync(0xffffffffu, eB[i], 1); eB[i] += __shfl_xor_sync(0xffffffffu, eB[i], 2);
                }
                if (t < 2) {
                    double *rs = t ? rowB_s : rowA_s;
#pragma unroll
                    for (int i = 0; i < 7; ++i) rs[i] += t ? eB[i] : eA[i];
                }
                kc = 0; ++vc;
            }
        }

        // ---- finalise and store ---------------------------------------------------------------------
        if (t < 2) {
            const bool v = t ? vB : vA;
            if (v) {
                const double *e = t ? rowB_s : rowA_s;
                const double px = e[13], py = e[14], pz = e[15];
                const double bx = a.B[0], by = a.B[1], bz = a.B[2];
                // J_m = sum_b ct(m,b) B_b:  1/2 [T_m + (V x B)_m]  +  1/2 rho (B x r)_m   (jtensor.F90:209-235 contracted with B)
                double jx = e[0] + (e[4] * bz - e[5] * by), jy = e[1] + (e[5] * bx - e[3] * bz), jz = e[2] + (e[3] * by - e[4] * bx);
                jx = a.paramag ? 0.5 * jx : 0.0; jy = a.paramag ? 0.5 * jy : 0.0; jz = a.paramag ? 0.5 * jz : 0.0;
                const double rho = e[6];
                if (a.diamag) {
                    jx += 0.5 * rho * (by * pz - bz * py); jy += 0.5 * rho * (bz * px - bx * pz); jz += 0.5 * rho * (bx * py - by * px);
                }
                const long o = out_row(a, td.pt0 + (t ? rowB : rowA));
                if (a.jvec) { a.jvec[3 * o] = jx; a.jvec[3 * o + 1] = jy; a.jvec[3 * o + 2] = jz; }
                if (a.jmod) a.jmod[o] = signed_modulus(jx, jy, jz, px, py, pz, bx, by, bz);
                if (a.edens) a.edens[o] = rho;
            }
        }
        }
        git += NIT;
    }
}

// ---- tensor path with an EPILOGUE WARPGROUP (default) -------------------------------------------------------------------------------
// ncu source view of the kernel above (profiles/r02_ncu_jtensor_source_regions.txt): the consumer warps spend 85.6 % of their time in the
// K loop, where the DMMA pipe is the limit, 10.4 % in the per-chunk epilogue (half of it waiting for the Phi / dPhi rows it loads from
// global memory) and 4 % in the tile prologue -- and they do so all at the same time, so the DMMA pipe idles for those 14 %.
// Here the epilogue is a role of its own.  16 warps: 8 consumers (K loop only), 4 producers, 4 epilogue warps (one thread per point).
//   consumer, end of a nu chunk:  forms  x0 = X_0,  z_b = X_{1+b} + (r x Y)_b  for its 16 x 16 accumulator tile (the GIAO part needs the
//            accumulators of the taps, so it stays here: 18 DFMA per element), writes the 4 values per (point, column) to a 64.5 KB
//            exchange buffer in shared memory and goes on with the next chunk's K loop;
//   epilogue warps:  row r reads the chunk's (x0, z_x, z_y, z_z) from the buffer and Phi, dPhi of the chunk's 16 columns from the panel
//            (row-contiguous => one coalesced 256 B request per warp and plane, latency off the critical path), keeps the 13 sums of
//            jtensor.F90:160-235 in registers for the whole tile, finalises and stores.
// Two more mbarriers (buffer full / buffer free) couple the roles; the producers publish, per chunk, the panel row and the centre R of
// each column (ring of 6 chunks) so that neither the consumers nor the epilogue warps chase  nlist -> fidx -> fR  through global memory.
// Registers: (8 x 200 + 4 x 40 + 4 x 72) x 32 = 65536 = the 512 x 128 the CTA is launched with (setmaxnreg only redistributes that pool).
constexpr int NEPI_WARPS = 4;
constexpr int NTHREADS_E = (NCONSUMER_WARPS + NPRODUCER_WARPS + NEPI_WARPS) * 32;
constexpr int CONSUMER_REGS_E = 200, PRODUCER_REGS_E = 40, EPI_REGS_E = 72;
constexpr int CRING = 6;               // column tables in flight.  The table of chunk c is written with the chunk's first stage and read until the
                                       // epilogue warps finish chunk c, i.e. before the consumers start the K loop of chunk c+2; the producers run at most
                                       // STAGES stages ahead, so slot (c + CRING) % CRING is rewritten no earlier than stage (c+CRING)*nkc - STAGES >= (c+2)*nkc
constexpr int ATAB_MAX_E = 112;        // active atoms of a tile staged in shared memory (more: the taps read the table from global memory)

template <int NPP_, int NV_, int LDB_, int NXV_>
struct SmemET {
    static constexpr int NPP = NPP_, NVC = NV_, LDB = LDB_;
    static constexpr int NXV = NXV_;                         // values per (point, column) in the exchange buffer: 4 (x0, z_x, z_y, z_z) or 2 (x0, z)
    static constexpr int XLD = NXV * MT + 4;                 // doubles per column: [NXV][128 rows] + 4 pad (the 4 column pairs of a quad land 64 B apart)
    static constexpr int A_DOUBLES = BK * LDP, PP_DOUBLES = BK * LDB, B_DOUBLES = NPP * PP_DOUBLES, STAGE_DOUBLES = A_DOUBLES + B_DOUBLES;
    static constexpr size_t X_OFF = (size_t)STAGES * STAGE_DOUBLES * 8;                 // exchange buffer [NVC][XLD]
    static constexpr size_t COLR_OFF = X_OFF + (size_t)NVC * XLD * 8;                   // [CRING][NVC][4]: R_x, R_y, R_z of the column's centre
    static constexpr size_t COLSLOT_OFF = COLR_OFF + (size_t)CRING * NVC * 4 * 8;       // [CRING][NVC]: panel row (K slot) of the column
    static constexpr size_t ATAB_OFF = COLSLOT_OFF + (size_t)CRING * NVC * 4;
    static constexpr size_t KMASK_OFF = ATAB_OFF + (size_t)ATAB_MAX_E * 3 * 8;
    static constexpr size_t BAR_OFF = KMASK_OFF + (size_t)KMASK_WORDS * 4;              // full[STAGES], empty[STAGES], xfull, xempty
    static constexpr size_t BYTES = BAR_OFF + (2 * STAGES + 2) * 8 + 16;
};
using SmemE = SmemET<(NQ + 1) / 2, NV, LDB2, 4>;     // tensor path: 16 columns x 4 values
using SmemEJ = SmemET<1, NVJ, LDB2J, 2>;             // J path: 32 columns x 2 values (the same 64.5 KB)
static_assert(SmemE::BYTES <= 232448 && SmemEJ::BYTES <= 232448, "k_jtensor_e: shared memory over the 227 KB of a CTA");

template <class SM>
__device__ __forceinline__ void producer_role_e(const JtensorArgs &a, uint32_t s_base, uint32_t bar_full, uint32_t bar_empty, double *s_colR,
                                                int *s_colSlot, int *s_tile) {
    constexpr int NV = SM::NVC;                                    // columns per chunk (shadows the tensor path's constant)
    constexpr int LPW = 32 / NV > 0 ? 32 / NV : 1;                // k rows covered by one warp per pass (2 for 16 columns, 1 for 32)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t git = 0, gch = 0;           // stages / nu chunks issued so far (all tiles)
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        if (td.nact == 0 || td.col1 <= td.col0) continue;
        const int nact = td.nact, nn = td.col1 - td.col0;                 // the columns of this work item (a whole tile or a slice of it)
        const int nkc = (nact + BK - 1) / BK, nvc = (nn + NV - 1) / NV;   // nn is a multiple of 8: the last nu chunk may hold 8 columns
        const uint32_t NIT = (uint32_t)nkc * nvc;
        const double *panel = a.panel_pool + td.panel_off;
        const int *fidx = a.fidx_pool + td.fidx_off, *nlist = fidx + nact + td.col0;
        const int pw = warp - NCONSUMER_WARPS;
        const int ldn = lane % NV, ldk0 = lane / NV + LPW * pw;  // this lane gathers nu slot ldn, k rows ldk0, ldk0 + LPW*4, ...
        int kc = 0, vc = 0;
        int slot = nlist[min(ldn, nn - 1)];
        long nu = fidx[slot];
        for (uint32_t itl = 0; itl < NIT; ++itl) {
            const uint32_t gi = git + itl, s = gi % STAGES, ph = (gi / STAGES) & 1;
            double Rx = 0, Ry = 0, Rz = 0;
            const bool table = kc == 0 && pw == 0 && lane < NV;          // this lane publishes column ldn of chunk vc
            if (table) { Rx = a.fR[nu]; Ry = a.fR[a.nbf + nu]; Rz = a.fR[2 * a.nbf + nu]; }   // in flight while the gathers are issued
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const int kcnt = min(BK, nact - kc * BK);
            const uint32_t sA = s_base + (uint32_t)(s * SM::STAGE_DOUBLES * 8), sB = sA + SM::A_DOUBLES * 8;
            if (pw == 0 && lane == 0) {
                mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)(kcnt * LDP * 8));
                tma_bulk_g2s(sA, panel + (long)kc * BK * LDP, (uint32_t)(kcnt * LDP * 8), bar_full + 8 * s);
            }
            const bool nu_ok = vc * NV + ldn < nn;
            const double *srcB = a.Bop + 2 * nu;
            const uint32_t dstB = sB + (uint32_t)(ldn * 16);
#pragma unroll 4
            for (int k = ldk0; k < (nu_ok ? kcnt : 0); k += LPW * NPRODUCER_WARPS) {
                const long mu = fidx[kc * BK + k];
                const double *src = srcB + 2 * mu * a.ldb;
                const uint32_t dst = dstB + (uint32_t)(k * SM::LDB * 8);
#pragma unroll
                for (int pp = 0; pp < SM::NPP; ++pp) cp_async_16(dst + (uint32_t)(pp * SM::PP_DOUBLES * 8), src + pp * a.plane_stride);
            }
            if (table) {
                // visible to the consumers through this stage's full barrier (they read it at the END of the sweep) and, through the
                // consumers' arrival on the exchange barrier, to the epilogue warps.  Ring of CRING chunks: the slot written now was
                // last read two or more K sweeps ago.
                const int ring = (int)((gch + vc) % CRING) * NV + ldn;
                s_colSlot[ring] = slot;
                double *cr = s_colR + 4 * ring;
                cr[0] = Rx; cr[1] = Ry; cr[2] = Rz;
                __threadfence_block();       // ordinary stores before the (asynchronous) arrival below
            }
            cp_async_arrive_noinc(bar_full + 8 * s);
            if (++kc == nkc) { kc = 0; ++vc; if (vc < nvc) { slot = nlist[min(vc * NV + ldn, nn - 1)]; nu = fidx[slot]; } }
        }
        git += NIT; gch += nvc;
    }
}

template <bool GIAO>
__device__ __forceinline__ void consumer_role_e(const JtensorArgs &a, const double *s_stage, double *s_x, const double *s_colR, double *s_atab,
                                                uint32_t *s_kmask, uint32_t bar_full, uint32_t bar_empty, uint32_t bar_xfull, uint32_t bar_xempty,
                                                int *s_tile) {
    using SM = SmemE;
    int lane = threadIdx.x & 31, row0 = (threadIdx.x >> 5) * 16;
    keep_in_register(lane); keep_in_register(row0);
    keep_in_register(bar_full); keep_in_register(bar_empty); keep_in_register(bar_xfull); keep_in_register(bar_xempty);
    const int g = lane >> 2, t = lane & 3;
    uint32_t git = 0, gch = 0;
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        if (td.nact == 0 || td.col1 <= td.col0) continue;            // the epilogue warps write the zeros; empty slices do nothing
        const int rowA = row0 + g, rowB = row0 + g + 8;
        const int nact = td.nact, nn = td.col1 - td.col0;
        const int nkc = (nact + BK - 1) / BK, nvc = (nn + NV - 1) / NV;
        const uint32_t NIT = (uint32_t)nkc * nvc;
        // coordinates of this thread's two rows and the tile centre (GIAO: r x Y, and Y relative to the centre like the tap weights)
        double rAx = 0, rAy = 0, rAz = 0, rBx = 0, rBy = 0, rBz = 0, cenx = 0, ceny = 0, cenz = 0;
        if (GIAO) {
            const long pA = td.pt0 + (rowA < td.npts ? rowA : 0), pB = td.pt0 + (rowB < td.npts ? rowB : 0);
            rAx = a.rsx[pA]; rAy = a.rsy[pA]; rAz = a.rsz[pA]; rBx = a.rsx[pB]; rBy = a.rsy[pB]; rBz = a.rsz[pB];
            const TileGeo tg = a.geo[td.geo];                          // same expression as k_basis
            cenx = 0.5 * (tg.lox + tg.hix); ceny = 0.5 * (tg.loy + tg.hiy); cenz = 0.5 * (tg.loz + tg.hiz);
        }
        double acc[NQ][2][4];
        double zac[3][2][4];                                        // Z_d (GIAO taps)
        const int nruns = td.nruns;
        const bool tab_sm = nruns <= ATAB_MAX_E;
        const double *gtab = reinterpret_cast<const double *>(a.atab_pool + td.atab_off);
        double curx = 0, cury = 0, curz = 0;
        int ia = 0;
        if (GIAO) {
            // stage the tile's atom table: weights to shared memory, atom ends as one bit per K step
            const double2 *atab = reinterpret_cast<const double2 *>(a.atab_pool + td.atab_off);   // TileAtom = 2 x double2
            const int ctid = threadIdx.x, nwords = (nact / 4 + 31) / 32;
            for (int w = ctid; w < nwords; w += NCONSUMER_WARPS * 32) s_kmask[w] = 0u;
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
            for (int r = ctid; r < nruns; r += NCONSUMER_WARPS * 32) {
                const double2 t0 = __ldg(atab + 2 * r), t1 = __ldg(atab + 2 * r + 1);
                if (r < ATAB_MAX_E) { s_atab[3 * r] = t0.x; s_atab[3 * r + 1] = t0.y; s_atab[3 * r + 2] = t1.x; }
                const int e = __double2loint(t1.y) - 1;
                atomicOr(&s_kmask[e >> 5], 1u << (e & 31));
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
        }
        int kc = 0, vc = 0;
        for (uint32_t it = 0; it < NIT; ++it) {
            const uint32_t gi = git + it, s = gi % STAGES, ph = (gi / STAGES) & 1;
            if (kc == 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[q][h][i] = 0.0;
                if (GIAO) {
#pragma unroll
                    for (int d = 0; d < 3; ++d)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int i = 0; i < 4; ++i) zac[d][h][i] = 0.0;
                    ia = 0;
                    load_tap_weights(tab_sm, s_atab, gtab, 0, curx, cury, curz);
                }
            }
            const int k4base = kc * (BK / 4);                                   // BK/4 = 8 K steps per stage: their bits share a word
            const uint32_t m8 = GIAO ? (s_kmask[k4base >> 5] >> (k4base & 31)) : 0u;
            const int nks = min(BK, nact - kc * BK) / 4;
            const bool h1 = vc * NV + 8 < nn;                       // second n8 tile of this chunk holds real columns
            const double *sA = s_stage + (size_t)s * SM::STAGE_DOUBLES;
            const double *sB = sA + SM::A_DOUBLES;
            mbar_wait(bar_full + 8 * s, ph);
#pragma unroll KSU
            for (int ks = 0; ks < nks; ++ks) {
                // fragments (m16n8k4.f64): a0 = A[row g][k t], a1 = A[row g+8][k t]; b0 = B[k t][n g]
                const double *pa = sA + (ks * 4 + t) * LDP + row0 + g;
                const double a0 = pa[0], a1 = pa[8];
                const double2 *pb = reinterpret_cast<const double2 *>(sB + (ks * 4 + t) * LDB2) + g;   // one LDS.128 = both planes of a pair
#pragma unroll
                for (int pp = 0; pp < SM::NPP; ++pp)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1 && !h1) continue;
                        const double2 b = pb[pp * (SM::PP_DOUBLES / 2) + h * 8];
                        mma_16x8x4_f64(acc[2 * pp][h], a0, a1, b.x);
                        mma_16x8x4_f64(acc[2 * pp + 1][h], a0, a1, b.y);
                    }
                if (GIAO && ((m8 >> ks) & 1u)) {
                    // last K step of an atom: Z_d += C_A * (R_A - R_next)_d  (see the header; C_A = acc[0] right now)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double cv = acc[0][h][i];
                            zac[0][h][i] = fma(curx, cv, zac[0][h][i]);
                            zac[1][h][i] = fma(cury, cv, zac[1][h][i]);
                            zac[2][h][i] = fma(curz, cv, zac[2][h][i]);
                        }
                    ia = min(ia + 1, nruns - 1);
                    load_tap_weights(tab_sm, s_atab, gtab, ia, curx, cury, curz);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);   // slot may be refilled
            if (++kc == nkc) {
                // ---- hand nu chunk vc to the epilogue warps: (x0, z_x, z_y, z_z) per (row, column) ----------------------------
                const uint32_t c = gch + vc;
                mbar_wait(bar_xempty, (c & 1) ^ 1);                  // the epilogue warps are done with the previous chunk
                const double *colR = s_colR + (size_t)(c % CRING) * NV * 4;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 1 && !h1) continue;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = h * 8 + 2 * t + j;
                        double dRx = 0, dRy = 0, dRz = 0;
                        if (GIAO) { const double2 r01 = *reinterpret_cast<const double2 *>(colR + 4 * col); dRx = r01.x - cenx; dRy = r01.y - ceny; dRz = colR[4 * col + 2] - cenz; }
                        double *xc = s_x + (size_t)col * SM::XLD;
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int ci = 2 * rr + j, row = rr ? rowB : rowA;
                            const double x0 = acc[0][h][ci];
                            double zx = acc[1][h][ci], zy = acc[2][h][ci], zz = acc[3][h][ci];
                            if (GIAO) {
                                const double px = rr ? rBx : rAx, py = rr ? rBy : rAy, pz = rr ? rBz : rAz;
                                const double yx = dRx * x0 - zac[0][h][ci], yy = dRy * x0 - zac[1][h][ci], yz = dRz * x0 - zac[2][h][ci];
                                zx += py * yz - pz * yy;   // (r x Y')_x
                                zy += pz * yx - px * yz;
                                zz += px * yy - py * yx;
                            }
                            xc[row] = x0; xc[MT + row] = zx; xc[2 * MT + row] = zy; xc[3 * MT + row] = zz;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_xfull);
                kc = 0; ++vc;
            }
        }
        git += NIT; gch += nvc;
    }
}

// one thread per point of the tile
template <bool GIAO>
__device__ __forceinline__ void epilogue_role_e(const JtensorArgs &a, const double *s_x, const double *s_colR, const int *s_colSlot,
                                                uint32_t bar_xfull, uint32_t bar_xempty, int *s_tile) {
    const int lane = threadIdx.x & 31;
    const int row = threadIdx.x - (NCONSUMER_WARPS + NPRODUCER_WARPS) * 32;
    uint32_t gch = 0;
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        const bool valid = row < td.npts;
        if (td.col1 <= td.col0 && td.part >= 0) continue;            // empty slice
        if (td.nact == 0) {   // nothing within screening range: the reference returns exact zeros
            if (valid) store_zero(a, out_row(a, td.pt0 + row));
            continue;
        }
        const int nact = td.nact, nn = td.col1 - td.col0;
        const int nvc = (nn + NV - 1) / NV;
        const double *panel = a.panel_pool + td.panel_off + row;       // rows >= npts hold zeros (k_basis writes all MT rows)
        const long plane = (long)nact * LDP;
        const long p = td.pt0 + (valid ? row : 0);
        const double px = a.rsx[p], py = a.rsy[p], pz = a.rsz[p];     // absolute coordinates (as r enters jtensor.F90:112 and bfeval.f90:168-189)
        double e[13];                                                   // Tp(m,b) at [m+3b], V_d at [9+d], rho at [12]
#pragma unroll
        for (int i = 0; i < 13; ++i) e[i] = 0.0;
        for (int vc = 0; vc < nvc; ++vc) {
            const uint32_t c = gch + vc;
            const int ncol = min(NV, nn - vc * NV);
            const int *cslot = s_colSlot + (size_t)(c % CRING) * NV;
            const double *colR = s_colR + (size_t)(c % CRING) * NV * 4;
            mbar_wait(bar_xfull, c & 1);
            for (int c0 = 0; c0 < ncol; c0 += 4) {                     // ncol is 8 or 16: four columns' panel rows in flight at a time
                double ev[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double *pe = panel + (long)cslot[c0 + k] * LDP;
#pragma unroll
                    for (int q = 0; q < 4; ++q) ev[k][q] = pe[q * plane];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double *xc = s_x + (size_t)(c0 + k) * SmemE::XLD + row;
                    const double x0 = xc[0], zx = xc[MT], zy = xc[2 * MT], zz = xc[3 * MT];
                    const double t0 = x0 * ev[k][0];
                    e[12] += t0;
                    if (GIAO) { const double *cr = colR + 4 * (c0 + k); e[9] += cr[0] * t0; e[10] += cr[1] * t0; e[11] += cr[2] * t0; }
                    e[0] += zx * ev[k][1]; e[1] += zx * ev[k][2]; e[2] += zx * ev[k][3];   // b = x: m = x,y,z
                    e[3] += zy * ev[k][1]; e[4] += zy * ev[k][2]; e[5] += zy * ev[k][3];
                    e[6] += zz * ev[k][1]; e[7] += zz * ev[k][2]; e[8] += zz * ev[k][3];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xempty);
        }
        gch += nvc;
        // ---- a slice hands its row sums to k_slice_reduce; a whole tile is finalised and stored here -------------------------------
        if (td.part >= 0) {
            double *pp = a.part + ((size_t)td.part * MT + row) * PART_LD;
#pragma unroll
            for (int i = 0; i < 13; ++i) pp[i] = e[i];
        } else if (valid) {
            finalise_store_tensor<GIAO>(a, e, px, py, pz, out_row(a, td.pt0 + row));
        }
    }
}

// ---- J = T.B path with the epilogue warpgroup: operands (D, P.B), 32-column chunks, ONE tap weight per row, 2 values per (point, column) ----
template <bool GIAO>
__device__ __forceinline__ void consumer_role_ej(const JtensorArgs &a, const double *s_stage, double *s_x, const double *s_colR, double *s_atab,
                                                 uint32_t *s_kmask, uint32_t bar_full, uint32_t bar_empty, uint32_t bar_xfull, uint32_t bar_xempty,
                                                 int *s_tile) {
    using SM = SmemEJ;
    int lane = threadIdx.x & 31, row0 = (threadIdx.x >> 5) * 16;
    keep_in_register(lane); keep_in_register(row0);
    keep_in_register(bar_full); keep_in_register(bar_empty); keep_in_register(bar_xfull); keep_in_register(bar_xempty);
    const int g = lane >> 2, t = lane & 3;
    constexpr int NH = NVJ / 8;                                 // n8 tiles per chunk
    uint32_t git = 0, gch = 0;
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        if (td.nact == 0 || td.col1 <= td.col0) continue;            // the epilogue warps write the zeros; empty slices do nothing
        const int rowA = row0 + g, rowB = row0 + g + 8;
        const int nact = td.nact, nn = td.col1 - td.col0;
        const int nkc = (nact + BK - 1) / BK, nvc = (nn + NVJ - 1) / NVJ;
        const uint32_t NIT = (uint32_t)nkc * nvc;
        // w = B x r of this thread's two rows: B.(r x Y) = Y.(B x r); tile centre (Y relative to the centre like the tap weights)
        double wAx = 0, wAy = 0, wAz = 0, wBx = 0, wBy = 0, wBz = 0, cenx = 0, ceny = 0, cenz = 0;
        if (GIAO) {
            const long pA = td.pt0 + (rowA < td.npts ? rowA : 0), pB = td.pt0 + (rowB < td.npts ? rowB : 0);
            const double bx = a.B[0], by = a.B[1], bz = a.B[2];
            const double ax = a.rsx[pA], ay = a.rsy[pA], az = a.rsz[pA], cx = a.rsx[pB], cy = a.rsy[pB], cz = a.rsz[pB];
            wAx = by * az - bz * ay; wAy = bz * ax - bx * az; wAz = bx * ay - by * ax;
            wBx = by * cz - bz * cy; wBy = bz * cx - bx * cz; wBz = bx * cy - by * cx;
            const TileGeo tg = a.geo[td.geo];                          // same expression as k_basis
            cenx = 0.5 * (tg.lox + tg.hix); ceny = 0.5 * (tg.loy + tg.hiy); cenz = 0.5 * (tg.loz + tg.hiz);
        }
        double acc[2][NH][4];                                       // planes D and P.B
        double zac[NH][4];                                          // S = (B x r) . Z (GIAO taps with per-row weights)
        const int nruns = td.nruns;
        const bool tab_sm = nruns <= ATAB_MAX_E;
        const double *gtab = reinterpret_cast<const double *>(a.atab_pool + td.atab_off);
        double curx = 0, cury = 0, curz = 0;
        int ia = 0;
        if (GIAO) {
            // stage the tile's atom table: weights to shared memory, atom ends as one bit per K step
            const double2 *atab = reinterpret_cast<const double2 *>(a.atab_pool + td.atab_off);   // TileAtom = 2 x double2
            const int ctid = threadIdx.x, nwords = (nact / 4 + 31) / 32;
            for (int w = ctid; w < nwords; w += NCONSUMER_WARPS * 32) s_kmask[w] = 0u;
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
            for (int r = ctid; r < nruns; r += NCONSUMER_WARPS * 32) {
                const double2 t0 = __ldg(atab + 2 * r), t1 = __ldg(atab + 2 * r + 1);
                if (r < ATAB_MAX_E) { s_atab[3 * r] = t0.x; s_atab[3 * r + 1] = t0.y; s_atab[3 * r + 2] = t1.x; }
                const int e = __double2loint(t1.y) - 1;
                atomicOr(&s_kmask[e >> 5], 1u << (e & 31));
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NCONSUMER_WARPS * 32) : "memory");
        }
        int kc = 0, vc = 0;
        for (uint32_t it = 0; it < NIT; ++it) {
            const uint32_t gi = git + it, s = gi % STAGES, ph = (gi / STAGES) & 1;
            if (kc == 0) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int h = 0; h < NH; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) acc[q][h][i] = 0.0;
                if (GIAO) {
#pragma unroll
                    for (int h = 0; h < NH; ++h)
#pragma unroll
                        for (int i = 0; i < 4; ++i) zac[h][i] = 0.0;
                    ia = 0;
                    load_tap_weights(tab_sm, s_atab, gtab, 0, curx, cury, curz);
                }
            }
            const int k4base = kc * (BK / 4);                                   // BK/4 = 8 K steps per stage: their bits share a word
            const uint32_t m8 = GIAO ? (s_kmask[k4base >> 5] >> (k4base & 31)) : 0u;
            const int nks = min(BK, nact - kc * BK) / 4;
            const int nh = min(NH, (nn - vc * NVJ) / 8);            // n8 tiles of this chunk that hold real columns (nn is a multiple of 8)
            const double *sA = s_stage + (size_t)s * SM::STAGE_DOUBLES;
            const double *sB = sA + SM::A_DOUBLES;
            mbar_wait(bar_full + 8 * s, ph);
            // fragments (m16n8k4.f64), software-pipelined by hand: the fragments of step ks+1 are loaded before the MMAs of step ks
            const double *pa0 = sA + t * LDP + row0 + g;
            const double2 *pb0 = reinterpret_cast<const double2 *>(sB + t * LDB2J) + g;
            double a0 = pa0[0], a1 = pa0[8];
            double2 bf[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) bf[h] = pb0[h * 8];
#pragma unroll KSU
            for (int ks = 0; ks < nks; ++ks) {
                double na0 = 0, na1 = 0;
                double2 nb[NH];
                if (ks + 1 < nks) {
                    const double *pa = pa0 + (ks + 1) * 4 * LDP;
                    na0 = pa[0]; na1 = pa[8];
                    const double2 *pb = pb0 + (ks + 1) * 4 * (LDB2J / 2);
#pragma unroll
                    for (int h = 0; h < NH; ++h) nb[h] = pb[h * 8];
                }
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    if (h >= nh) continue;
                    mma_16x8x4_f64(acc[0][h], a0, a1, bf[h].x);
                    mma_16x8x4_f64(acc[1][h], a0, a1, bf[h].y);
                }
                a0 = na0; a1 = na1;
#pragma unroll
                for (int h = 0; h < NH; ++h) bf[h] = nb[h];
                if (GIAO && ((m8 >> ks) & 1u)) {
                    // last K step of an atom: S += C_A * ((B x r) . (R_A - R_next)), one weight per row (see the header)
                    const double oA = wAx * curx + wAy * cury + wAz * curz, oB = wBx * curx + wBy * cury + wBz * curz;
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        zac[h][0] = fma(oA, acc[0][h][0], zac[h][0]); zac[h][1] = fma(oA, acc[0][h][1], zac[h][1]);
                        zac[h][2] = fma(oB, acc[0][h][2], zac[h][2]); zac[h][3] = fma(oB, acc[0][h][3], zac[h][3]);
                    }
                    ia = min(ia + 1, nruns - 1);
                    load_tap_weights(tab_sm, s_atab, gtab, ia, curx, cury, curz);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * s);   // slot may be refilled
            if (++kc == nkc) {
                // ---- hand nu chunk vc to the epilogue warps: (x0, z) per (row, column), z = X_1 + (B x r) . Y ----------------------
                const uint32_t c = gch + vc;
                mbar_wait(bar_xempty, (c & 1) ^ 1);                  // the epilogue warps are done with the previous chunk
                const double *colR = s_colR + (size_t)(c % CRING) * NVJ * 4;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    if (h >= nh) continue;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = h * 8 + 2 * t + j;
                        double dRx = 0, dRy = 0, dRz = 0;
                        if (GIAO) { const double2 r01 = *reinterpret_cast<const double2 *>(colR + 4 * col); dRx = r01.x - cenx; dRy = r01.y - ceny; dRz = colR[4 * col + 2] - cenz; }
                        double *xc = s_x + (size_t)col * SM::XLD;
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int ci = 2 * rr + j, row = rr ? rowB : rowA;
                            const double x0 = acc[0][h][ci];
                            double z = acc[1][h][ci];
                            if (GIAO) {
                                const double wx = rr ? wBx : wAx, wy = rr ? wBy : wAy, wz = rr ? wBz : wAz;
                                z += (wx * dRx + wy * dRy + wz * dRz) * x0 - zac[h][ci];   // (B x r) . Y
                            }
                            xc[row] = x0; xc[MT + row] = z;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_xfull);
                kc = 0; ++vc;
            }
        }
        git += NIT; gch += nvc;
    }
}

template <bool GIAO>
__device__ __forceinline__ void epilogue_role_ej(const JtensorArgs &a, const double *s_x, const double *s_colR, const int *s_colSlot,
                                                 uint32_t bar_xfull, uint32_t bar_xempty, int *s_tile) {
    const int lane = threadIdx.x & 31;
    const int row = threadIdx.x - (NCONSUMER_WARPS + NPRODUCER_WARPS) * 32;
    uint32_t gch = 0;
    for (;;) {
        const int tile = next_tile(a, s_tile);
        if (tile >= a.ntiles) break;
        const TileDesc td = a.tiles[tile];
        const bool valid = row < td.npts;
        if (td.col1 <= td.col0 && td.part >= 0) continue;            // empty slice
        if (td.nact == 0) {   // nothing within screening range: the reference returns exact zeros
            if (valid) store_zero(a, out_row(a, td.pt0 + row));
            continue;
        }
        const int nact = td.nact, nn = td.col1 - td.col0;
        const int nvc = (nn + NVJ - 1) / NVJ;
        const double *panel = a.panel_pool + td.panel_off + row;       // rows >= npts hold zeros (k_basis writes all MT rows)
        const long plane = (long)nact * LDP;
        const long p = td.pt0 + (valid ? row : 0);
        const double px = a.rsx[p], py = a.rsy[p], pz = a.rsz[p];
        double e[13];                                                   // T_m = sum_b Tp(m,b) B_b at [m], V_d at [3+d], rho at [6] (7 used)
#pragma unroll
        for (int i = 0; i < 13; ++i) e[i] = 0.0;
        for (int vc = 0; vc < nvc; ++vc) {
            const uint32_t c = gch + vc;
            const int ncol = min(NVJ, nn - vc * NVJ);
            const int *cslot = s_colSlot + (size_t)(c % CRING) * NVJ;
            const double *colR = s_colR + (size_t)(c % CRING) * NVJ * 4;
            mbar_wait(bar_xfull, c & 1);
            for (int c0 = 0; c0 < ncol; c0 += 4) {                     // ncol is a multiple of 8: four columns' panel rows in flight at a time
                double ev[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double *pe = panel + (long)cslot[c0 + k] * LDP;
#pragma unroll
                    for (int q = 0; q < 4; ++q) ev[k][q] = pe[q * plane];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const double *xc = s_x + (size_t)(c0 + k) * SmemEJ::XLD + row;
                    const double x0 = xc[0], z = xc[MT];
                    const double t0 = x0 * ev[k][0];
                    e[6] += t0;
                    if (GIAO) { const double *cr = colR + 4 * (c0 + k); e[3] += cr[0] * t0; e[4] += cr[1] * t0; e[5] += cr[2] * t0; }
                    e[0] += z * ev[k][1]; e[1] += z * ev[k][2]; e[2] += z * ev[k][3];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xempty);
        }
        gch += nvc;
        if (td.part >= 0) {
            double *pp = a.part + ((size_t)td.part * MT + row) * PART_LD;
#pragma unroll
            for (int i = 0; i < 7; ++i) pp[i] = e[i];
        } else if (valid) {
            finalise_store_j(a, e, px, py, pz, out_row(a, td.pt0 + row));
        }
    }
}

template <bool GIAO, bool JVEC>
__global__ void __launch_bounds__(NTHREADS_E, 1) k_jtensor_e(JtensorArgs a) {
    using SM = typename std::conditional<JVEC, SmemEJ, SmemE>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_stage = reinterpret_cast<double *>(smem_raw);
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bar_full = s_base + (uint32_t)SM::BAR_OFF, bar_empty = bar_full + STAGES * 8;
    const uint32_t bar_xfull = bar_empty + STAGES * 8, bar_xempty = bar_xfull + 8;
    int *s_tile = reinterpret_cast<int *>(smem_raw + SM::BAR_OFF + (2 * STAGES + 2) * 8);
    double *s_x = reinterpret_cast<double *>(smem_raw + SM::X_OFF), *s_colR = reinterpret_cast<double *>(smem_raw + SM::COLR_OFF);
    int *s_colSlot = reinterpret_cast<int *>(smem_raw + SM::COLSLOT_OFF);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, NPRODUCER_WARPS * 32 + 1); mbar_init(bar_empty + 8 * s, NCONSUMER_WARPS); }
        mbar_init(bar_xfull, NCONSUMER_WARPS); mbar_init(bar_xempty, NEPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp >= NCONSUMER_WARPS + NPRODUCER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(EPI_REGS_E));
        if (JVEC) epilogue_role_ej<GIAO>(a, s_x, s_colR, s_colSlot, bar_xfull, bar_xempty, s_tile);
        else epilogue_role_e<GIAO>(a, s_x, s_colR, s_colSlot, bar_xfull, bar_xempty, s_tile);
    } else if (warp >= NCONSUMER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS_E));
        producer_role_e<SM>(a, s_base, bar_full, bar_empty, s_colR, s_colSlot, s_tile);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS_E));
        double *atab = reinterpret_cast<double *>(smem_raw + SM::ATAB_OFF);
        uint32_t *kmask = reinterpret_cast<uint32_t *>(smem_raw + SM::KMASK_OFF);
        if (JVEC) consumer_role_ej<GIAO>(a, s_stage, s_x, s_colR, atab, kmask, bar_full, bar_empty, bar_xfull, bar_xempty, s_tile);
        else consumer_role_e<GIAO>(a, s_stage, s_x, s_colR, atab, kmask, bar_full, bar_empty, bar_xfull, bar_xempty, s_tile);
    }
}

template <bool GIAO, bool JVEC>
static void launch_one_e(const JtensorArgs &a, int grid, cudaStream_t s) {
    constexpr size_t bytes = JVEC ? SmemEJ::BYTES : SmemE::BYTES;
    cudaFuncSetAttribute(k_jtensor_e<GIAO, JVEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    k_jtensor_e<GIAO, JVEC><<<grid, NTHREADS_E, bytes, s>>>(a);
}

// Pipeline: warps 8-11 are producers.  Per stage they (1) wait for the slot to be released by the 8 consumer warps
// (empty barrier), (2) issue ONE bulk-TMA copy of the contiguous Phi panel rows [kc*BK, +kcnt) x 132 doubles
// (expect_tx on the full barrier) and (3) gather the density elements B_q[fidx[k]][fidx[nu]] of all NQ planes with
// 8-byte cp.async, whose completion arrives on the same full barrier.  Consumer warps only wait(full) -> LDS + DMMA ->
// arrive(empty); nobody executes a CTA-wide barrier inside a tile.  The two roles are separate code paths so that
// setmaxnreg can give the consumers 232 registers (ptxas budgets each path by the setmaxnreg that dominates it).
template <bool GIAO, bool JVEC, int NCW>
__global__ void __launch_bounds__((NCW + NPRODUCER_WARPS) * 32, 1) k_jtensor(JtensorArgs a) {
    using SM = typename std::conditional<JVEC, SmemJ, Smem>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *s_stage = reinterpret_cast<double *>(smem_raw);
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bar_full = s_base + (uint32_t)SM::BAR_OFF, bar_empty = bar_full + STAGES * 8;
    int *s_tile = reinterpret_cast<int *>(smem_raw + SM::BAR_OFF + 2 * STAGES * 8);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, NPRODUCER_WARPS * 32 + 1); mbar_init(bar_empty + 8 * s, NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x >> 5) >= NCW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
        producer_role<SM, NCW>(a, s_base, bar_full, bar_empty, s_tile);
    } else {
        // (8*232 + 4*40) * 32 = 64512 = 384 x 168: exactly the launch allocation
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
        double *rows = reinterpret_cast<double *>(smem_raw + SM::ROW_OFF), *atab = reinterpret_cast<double *>(smem_raw + SM::ATAB_OFF);
        uint32_t *kmask = reinterpret_cast<uint32_t *>(smem_raw + SM::KMASK_OFF);
        if (JVEC) consumer_role_j<GIAO>(a, s_stage, rows, atab, kmask, bar_full, bar_empty, s_tile);
        else consumer_role<GIAO>(a, s_stage, rows, atab, kmask, bar_full, bar_empty, s_tile);
    }
}

size_t jtensor_smem_bytes() { return SmemEJ::BYTES > SmemE::BYTES ? SmemEJ::BYTES : SmemE::BYTES; }

template <bool GIAO, bool JVEC, int NCW>
static void launch_one(const JtensorArgs &a, int grid, cudaStream_t s) {
    constexpr size_t bytes = JVEC ? SmemJ::BYTES : Smem::BYTES;
    // per-device function attribute (a process may hold contexts on several GPUs): cheap enough to set on every launch
    cudaFuncSetAttribute(k_jtensor<GIAO, JVEC, NCW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    k_jtensor<GIAO, JVEC, NCW><<<grid, (NCW + NPRODUCER_WARPS) * 32, bytes, s>>>(a);
}

// adds the row sums of a tile's slices in slice order (fixed => reproducible), finalises and stores.  One CTA (MT threads) per tile.
template <bool GIAO>
__global__ void __launch_bounds__(MT) k_slice_reduce(JtensorArgs a, const TileDesc *__restrict__ tiles, int S, long long item_cost) {
    const TileDesc td = tiles[blockIdx.x];
    const int row = threadIdx.x;
    if (td.nact == 0 || row >= td.npts) return;                       // empty tiles were zero-filled by their first item
    const int w = slice_width(td, S, item_cost);
    const int nsl = (td.nn + w - 1) / w;
    const int ne = a.jpath ? 7 : 13;
    double e[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) e[i] = 0.0;
    for (int sl = 0; sl < nsl; ++sl) {
        const double *pp = a.part + ((size_t)(blockIdx.x * S + sl) * MT + row) * PART_LD;
        for (int i = 0; i < ne; ++i) e[i] += pp[i];
    }
    const long p = td.pt0 + row;
    const double px = a.rsx[p], py = a.rsy[p], pz = a.rsz[p];
    if (a.jpath) finalise_store_j(a, e, px, py, pz, out_row(a, p));
    else finalise_store_tensor<GIAO>(a, e, px, py, pz, out_row(a, p));
}
void launch_slice_reduce(const JtensorArgs &a, const TileDesc *tiles, int nt, int S, long long item_cost, bool giao, cudaStream_t s) {
    if (nt <= 0) return;
    if (giao) k_slice_reduce<true><<<nt, MT, 0, s>>>(a, tiles, S, item_cost); else k_slice_reduce<false><<<nt, MT, 0, s>>>(a, tiles, S, item_cost);
}

// GIMIC_B200_EPI=0 selects the round-1 mapping of the tensor path (epilogue inside the consumer warps) for A/B measurements
static bool tensor_path_epilogue_role() {
    static const bool on = [] { const char *e = std::getenv("GIMIC_B200_EPI"); return e ? std::atoi(e) != 0 : DEFAULT_EPI != 0; }();
    return on;
}

bool jtensor_supports_slices() { return tensor_path_epilogue_role(); }

void launch_jtensor(const JtensorArgs &a, bool giao, int nsm, cudaStream_t s) {
    if (a.ntiles <= 0) return;
    const int grid = a.ntiles < nsm ? a.ntiles : nsm;
    const bool jv = a.jpath != 0;        // J = T.B path: operands are ONE pair-plane (D, sum_b B_b P_b)
    if (tensor_path_epilogue_role()) {
        if (jv) { if (giao) launch_one_e<true, true>(a, grid, s); else launch_one_e<false, true>(a, grid, s); }
        else { if (giao) launch_one_e<true, false>(a, grid, s); else launch_one_e<false, false>(a, grid, s); }
        return;
    }
    if (jv) { if (giao) launch_one<true, true, 8>(a, grid, s); else launch_one<false, true, 8>(a, grid, s); return; }
    if (giao) launch_one<true, false, 8>(a, grid, s); else launch_one<false, false, 8>(a, grid, s);
}

// ---------------------------------------------------------------------------------------------
// Contraction operands in the internal (per-atom radius-sorted) function order as PAIR-PLANES, row-major [mu][nu][2]:
// pair 0 = (D, Px), pair 1 = (Py, Pz); optionally alpha +/- beta.
// One (mu,nu) element of a pair is a 16-byte unit: one cp.async.cg per gather, and runs of >=4 nu fill whole 64 B HBM atoms.
// src is the dens.f90 layout: element (a,b) at a + nbf*b.
__global__ void k_build_operand(double *__restrict__ out, int nbf, int ldb, long long plane_stride, const double *__restrict__ srcA,
                                const double *__restrict__ srcB, double signB, const int *__restrict__ f2user) {
    const int nu = blockIdx.x * blockDim.x + threadIdx.x;
    if (nu >= nbf) return;
    const long un = f2user[nu], nn = (long)nbf * nbf;
    for (int mu = blockIdx.y; mu < nbf; mu += gridDim.y) {   // gridDim.y is capped at 65535: rows beyond that by stride
        const long um = f2user[mu];
        const long src = um + (long)nbf * un;
        const long dst = 2 * ((long)mu * ldb + nu);
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = srcA[q * nn + src];
            if (srcB) v[q] += signB * srcB[q * nn + src];
        }
        for (int pp = 0; pp < 2; ++pp) *reinterpret_cast<double2 *>(out + pp * plane_stride + dst) = make_double2(v[2 * pp], v[2 * pp + 1]);
    }
}
// J = T.B path: ONE pair-plane (D, B_x P_x + B_y P_y + B_z P_z): the tensor is only ever contracted with this field direction.
// Built from the tensor-path operand planes (same internal order, elementwise), so the raw densities need not stay resident.
__global__ void k_operand_j(double2 *__restrict__ out, const double2 *__restrict__ p0, const double2 *__restrict__ p1, long count,
                            double bx, double by, double bz) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) {
        const double2 a = p0[i], b = p1[i];                     // (D, Px), (Py, Pz)
        out[i] = make_double2(a.x, bx * a.y + by * b.x + bz * b.y);
    }
}
void launch_operand_j(double *out, const double *op, long long plane_stride, const double *B3, cudaStream_t s) {
    const long count = plane_stride / 2;
    k_operand_j<<<148 * 8, 256, 0, s>>>(reinterpret_cast<double2 *>(out), reinterpret_cast<const double2 *>(op),
                                        reinterpret_cast<const double2 *>(op + plane_stride), count, B3[0], B3[1], B3[2]);
}
// total / spin-density operands of an open-shell context: alpha +- beta (jtensor.F90:86-88, 97-99; linear in D, P)
__global__ void k_operand_combine(double *__restrict__ out, const double *__restrict__ a, const double *__restrict__ b, double sg, long count) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < count; i += (long)gridDim.x * blockDim.x) out[i] = a[i] + sg * b[i];
}
void launch_operand_combine(double *out, const double *a, const double *b, double sg, long count, cudaStream_t s) {
    k_operand_combine<<<148 * 8, 256, 0, s>>>(out, a, b, sg, count);
}

void launch_build_operand(double *out, int nbf, int ldb, long long plane_stride, const double *srcA, const double *srcB, double signB,
                          const int *f2user, cudaStream_t s) {
    dim3 grid((nbf + 127) / 128, nbf < 65535 ? nbf : 65535);
    k_build_operand<<<grid, 128, 0, s>>>(out, nbf, ldb, plane_stride, srcA, srcB, signB, f2user);
}

}  // namespace gb
