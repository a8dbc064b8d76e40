// Field post-processing (HBM-bound) and plane quadrature.
//
//   k_fields        jvec = T.B (jfield.f90:167-184), signed |J| (jfield.f90:446-489), ACID (acid.f90:9-45)
//   k_quad_rows     inner (i) sums of integrate_current / integrate_modulus / integrate_acid
//                   (src/fgimic/integral.f90:123-148, 262-290, 473-487), one warp per (j,k) row
//   k_quad_final    outer (j,k) sums with the row weights, fixed-order tree => deterministic
#include "kernels.cuh"

namespace gb {

constexpr int FB = 256;   // points per block in k_fields

// Tensors are 9 doubles per point (AoS, the reference's tens(9,N)).  A block stages its 256x9 contiguous doubles through shared
// memory with 128-bit (double2) loads -- 72 B per point in, 24 + 8 + 8 B out, 24 B of coordinates for the signed modulus: the pass
// is HBM-bound, so every byte moves once, fully coalesced, in 16-byte units where alignment allows (a block's tensor slab starts at
// 256*9*8 B = a multiple of 16; the point count of the last block may be odd: scalar tail).  Each thread then works on one point
// (stride 9 is odd => conflict-free) and the vectors go back out through shared memory, again as double2.
// VEC = false: the same pass with 8-byte accesses, for caller pointers that are not 16-byte aligned (e.g. a view into a larger array).
template <bool VEC>
__global__ void __launch_bounds__(FB) k_fields(long n, const double *__restrict__ r, const double *__restrict__ tens, double bx,
                                               double by, double bz, double *__restrict__ jvec, double *__restrict__ jmod,
                                               double *__restrict__ acid) {
    __shared__ __align__(16) double s_t[FB * 9];
    __shared__ __align__(16) double s_r[FB * 3 + 1];
    const long base = (long)blockIdx.x * FB;
    const int cnt = (int)((n - base) < FB ? (n - base) : FB);
    if (!VEC) {
        for (int i = threadIdx.x; i < cnt * 9; i += FB) s_t[i] = tens[base * 9 + i];
        if (jmod) for (int i = threadIdx.x; i < cnt * 3; i += FB) s_r[i] = r[base * 3 + i];
    } else {
        const int nd = cnt * 9, n2 = nd >> 1;
        const double2 *src = reinterpret_cast<const double2 *>(tens + base * 9);     // base*9*8 B is 16-byte aligned (FB*9 even)
        double2 *dst = reinterpret_cast<double2 *>(s_t);
        for (int i = threadIdx.x; i < n2; i += FB) dst[i] = __ldcs(src + i);          // streamed: read once
        if ((nd & 1) && threadIdx.x == 0) s_t[nd - 1] = tens[base * 9 + nd - 1];
    }
    if (VEC && jmod) {
        const int nd = cnt * 3, n2 = nd >> 1;
        const double2 *src = reinterpret_cast<const double2 *>(r + base * 3);
        double2 *dst = reinterpret_cast<double2 *>(s_r);
        for (int i = threadIdx.x; i < n2; i += FB) dst[i] = __ldcs(src + i);
        if ((nd & 1) && threadIdx.x == 0) s_r[nd - 1] = r[base * 3 + nd - 1];
    }
    __syncthreads();
    double vx = 0, vy = 0, vz = 0, jm = 0, ac = 0;
    const int p = threadIdx.x;
    if (p < cnt) {
        const double *t = &s_t[9 * p];
        vx = t[0] * bx + t[3] * by + t[6] * bz;     // matmul(reshape(tens,(3,3)), b)
        vy = t[1] * bx + t[4] * by + t[7] * bz;
        vz = t[2] * bx + t[5] * by + t[8] * bz;
        if (acid) {
            const double xxmyy = (t[0] - t[4]) * (t[0] - t[4]), yymzz = (t[4] - t[8]) * (t[4] - t[8]), zzmxx = (t[8] - t[0]) * (t[8] - t[0]);
            const double xypyx = (t[3] + t[1]) * (t[3] + t[1]), xzpzx = (t[6] + t[2]) * (t[6] + t[2]), yzpzy = (t[7] + t[5]) * (t[7] + t[5]);
            ac = 0.3333333 * (xxmyy + yymzz + zzmxx) + 0.5 * (xypyx + xzpzx + yzpzy);   // DP33, globals.f90:62
        }
        if (jmod) {
            double cx = s_r[3 * p], cy = s_r[3 * p + 1], cz = s_r[3 * p + 2];
            jm = sqrt(vx * vx + vy * vy + vz * vz);
            const double d = bx * cx + by * cy + bz * cz;
            cx -= d * bx; cy -= d * by; cz -= d * bz;
            const double nx = by * cz - bz * cy, ny = bz * cx - bx * cz, nz = bx * cy - by * cx;   // cross_product(mag, coord)
            if (nx * vx + ny * vy + nz * vz < 0.0) jm = -1.0 * jm;
        }
    }
    if (jmod && p < cnt) __stcs(jmod + base + p, jm);
    if (acid && p < cnt) __stcs(acid + base + p, ac);
    if (jvec) {
        __syncthreads();
        if (p < cnt) { s_r[3 * p] = vx; s_r[3 * p + 1] = vy; s_r[3 * p + 2] = vz; }
        __syncthreads();
        const int nd = cnt * 3, n2 = nd >> 1;
        if (!VEC) { for (int i = threadIdx.x; i < nd; i += FB) jvec[base * 3 + i] = s_r[i]; return; }
        double2 *dst = reinterpret_cast<double2 *>(jvec + base * 3);
        const double2 *src = reinterpret_cast<const double2 *>(s_r);
        for (int i = threadIdx.x; i < n2; i += FB) __stcs(dst + i, src[i]);
        if ((nd & 1) && threadIdx.x == 0) jvec[base * 3 + nd - 1] = s_r[nd - 1];
    }
}
void launch_fields(long n, const double *r, const double *tens, const double *B3, double *jvec, double *jmod, double *acid, cudaStream_t s) {
    if (n <= 0) return;
    const bool vec = ((reinterpret_cast<uintptr_t>(tens) | reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(jvec)) & 15) == 0;
    if (vec) k_fields<true><<<(unsigned)((n + FB - 1) / FB), FB, 0, s>>>(n, r, tens, B3[0], B3[1], B3[2], jvec, jmod, acid);
    else k_fields<false><<<(unsigned)((n + FB - 1) / FB), FB, 0, s>>>(n, r, tens, B3[0], B3[1], B3[2], jvec, jmod, acid);
}

// signed |J| from J alone (jfield.f90:446-489), for the J = T.B path that never forms the tensor
__global__ void k_jmod(long n, const double *__restrict__ r, const double *__restrict__ jvec, double bx, double by, double bz,
                       double *__restrict__ jmod) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double vx = jvec[3 * i], vy = jvec[3 * i + 1], vz = jvec[3 * i + 2];
    double cx = r[3 * i], cy = r[3 * i + 1], cz = r[3 * i + 2];
    double jm = sqrt(vx * vx + vy * vy + vz * vz);
    const double d = bx * cx + by * cy + bz * cz;
    cx -= d * bx; cy -= d * by; cz -= d * bz;
    const double nx = by * cz - bz * cy, ny = bz * cx - bx * cz, nz = bx * cy - by * cx;   // cross_product(mag, coord)
    if (nx * vx + ny * vy + nz * vz < 0.0) jm = -1.0 * jm;
    jmod[i] = jm;
}
void launch_jmod(long n, const double *r, const double *jvec, const double *B3, double *jmod, cudaStream_t s) {
    if (n <= 0) return;
    k_jmod<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, r, jvec, B3[0], B3[1], B3[2], jmod);
}

// divj by central differences: r6 holds the 6 shifted copies (+x,-x,+y,-y,+z,-z) of every point
__global__ void k_shift_points(long n, const double *__restrict__ r, double h, double *__restrict__ r6) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        const double sg = (s & 1) ? -h : h;
        double *o = r6 + 3 * ((long)s * n + i);
        o[0] = x + (s / 2 == 0 ? sg : 0.0); o[1] = y + (s / 2 == 1 ? sg : 0.0); o[2] = z + (s / 2 == 2 ? sg : 0.0);
    }
}
void launch_shift_points(long n, const double *r, double h, double *r6, cudaStream_t s) {
    if (n <= 0) return;
    k_shift_points<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, r, h, r6);
}
__global__ void k_divj(long n, const double *__restrict__ jv6, double inv2h, double *__restrict__ divj) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = 0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) d += (jv6[3 * ((long)(2 * ax) * n + i) + ax] - jv6[3 * ((long)(2 * ax + 1) * n + i) + ax]) * inv2h;
    divj[i] = d;
}
void launch_divj(long n, const double *jv6, double h, double *divj, cudaStream_t s) {
    if (n <= 0) return;
    k_divj<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, jv6, 0.5 / h, divj);
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_quad_rows(QuadArgs q) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + warp;
    if (row >= q.nrows) return;
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = lane; i < q.p1; i += 32) {
        const long p = (long)row * q.p1 + i;
        const double *t = q.jvec ? nullptr : q.tens + 9 * p;
        const double dx = q.r[3 * p] - q.center[0], dy = q.r[3 * p + 1] - q.center[1], dz = q.r[3 * p + 2] - q.center[2];
        const bool inside = !(sqrt(dx * dx + dy * dy + dz * dz) > q.radius);      // integral.f90:125,128
        const double w = inside ? q.w1[i] : 0.0;
        double vx, vy, vz;
        if (q.jvec) { vx = q.jvec[3 * p]; vy = q.jvec[3 * p + 1]; vz = q.jvec[3 * p + 2]; }      // formed inside the contraction
        else {
            vx = t[0] * q.B[0] + t[3] * q.B[1] + t[6] * q.B[2];
            vy = t[1] * q.B[0] + t[4] * q.B[1] + t[7] * q.B[2];
            vz = t[2] * q.B[0] + t[5] * q.B[1] + t[8] * q.B[2];
        }
        const double nj = q.normal[0] * vx + q.normal[1] * vy + q.normal[2] * vz;
        if (q.what & 1) {                                                         // integrate_current
            const double jp = inside ? nj * w : 0.0;
            s[0] += jp;
            if (jp > 0.0) s[1] += jp; else s[2] += jp;
        }
        if (q.what & 2) {                                                         // integrate_modulus
            double sgn = 0.0;                                                     // outside the bound w = 0, so the stale sgn of the reference never matters
            if (inside) sgn = fabs(nj) < 1e-12 ? 0.0 : (nj > 0 ? 1.0 : -1.0);
            const double jp = sgn * sqrt(vx * vx + vy * vy + vz * vz);
            s[3] += jp * w;
            if (jp > 0.0) s[4] += jp * w; else s[5] += jp * w;
        }
        if ((q.what & 4) && t) {                                                  // integrate_acid (needs the tensor)
            const double xxmyy = (t[0] - t[4]) * (t[0] - t[4]), yymzz = (t[4] - t[8]) * (t[4] - t[8]), zzmxx = (t[8] - t[0]) * (t[8] - t[0]);
            const double xypyx = (t[3] + t[1]) * (t[3] + t[1]), xzpzx = (t[6] + t[2]) * (t[6] + t[2]), yzpzy = (t[7] + t[5]) * (t[7] + t[5]);
            s[6] += (0.3333333 * (xxmyy + yymzz + zzmxx) + 0.5 * (xypyx + xzpzx + yzpzy)) * w;
        }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) {
        const double wr = q.wrow[row];
#pragma unroll
        for (int k = 0; k < 7; ++k) q.row_partials[7 * (long)row + k] = s[k] * wr;
    }
}
__global__ void __launch_bounds__(256) k_quad_final(const double *__restrict__ part, int nrows, double *__restrict__ out7) {
    __shared__ double s[256];
    for (int k = 0; k < 7; ++k) {
        double v = 0;
        for (int r = threadIdx.x; r < nrows; r += 256) v += part[7 * (long)r + k];
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) out7[k] = s[0];
        __syncthreads();
    }
}
// ---------------------------------------------------------------------------------------------
// get_property (src/fgimic/jfield.f90:584-929): Biot-Savart shielding sigma_K at every nucleus and the magnetizability chi
// from the tensor field on a weighted point set (NumGrid).  Block (seg, k): partial sums over point segment `seg`
// (the per-atom grid blocks of nelpts.info) for nucleus k; k == natoms is the magnetizability.  Fixed-order tree => deterministic.
//   out[(k*nseg + seg)*5 + {0,1,2}] = sum w * integrand_{xx,yy,zz},  [3] = sum of positive w*(xx+yy+zz), [4] = negative
__global__ void __launch_bounds__(256) k_property(long n, const double *__restrict__ r, const double *__restrict__ w,
                                                  const double *__restrict__ tens, int natoms, const double *__restrict__ coords,
                                                  int nseg, const long *__restrict__ seg_end, double *__restrict__ out) {
    __shared__ double red[256];
    const int seg = blockIdx.x, k = blockIdx.y;
    const long lo = seg ? seg_end[seg - 1] : 0, hi = seg_end[seg];
    double acc[5] = {0, 0, 0, 0, 0};
    const bool chi = (k == natoms);
    const double cx = chi ? 0.0 : coords[3 * k], cy = chi ? 0.0 : coords[3 * k + 1], cz = chi ? 0.0 : coords[3 * k + 2];
    for (long i = lo + threadIdx.x; i < hi; i += 256) {
        const double dx = r[3 * i] - cx, dy = r[3 * i + 1] - cy, dz = r[3 * i + 2] - cz;
        const double *t = tens + 9 * i;
        double f;
        if (chi) f = 0.5;                                                                   // jfield.f90:834
        else f = 1.0e6 * (-1.0 / pow(dx * dx + dy * dy + dz * dz, 1.5) / (137.0359998 * 137.0359998));   // jfield.f90:702
        // jvec = T.(-e_b) = -T(:,b)   (jfield.f90:704-725)
        const double ixx = f * (dy * (-t[2]) - dz * (-t[1]));
        const double iyy = f * (dz * (-t[3]) - dx * (-t[5]));
        const double izz = f * (dx * (-t[7]) - dy * (-t[6]));
        const double wi = w[i], pd = ixx + iyy + izz;
        acc[0] += wi * ixx; acc[1] += wi * iyy; acc[2] += wi * izz;
        if (pd >= 0.0) acc[3] += pd * wi; else acc[4] += pd * wi;
    }
    for (int q = 0; q < 5; ++q) {
        red[threadIdx.x] = acc[q];
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) out[((long)k * nseg + seg) * 5 + q] = red[0];
        __syncthreads();
    }
}
void launch_property(long n, const double *r, const double *w, const double *tens, int natoms, const double *coords, int nseg,
                     const long *seg_end, double *out, cudaStream_t s) {
    if (nseg <= 0) return;
    k_property<<<dim3(nseg, natoms + 1), 256, 0, s>>>(n, r, w, tens, natoms, coords, nseg, seg_end, out);
}

// Pointwise integrands of get_property for ONE centre (the arrays the reference plots as sigma<k>.vtu, sigma_xx<k>.vtu, ...
// and intchi*.vtu, jfield.f90:786-808, 915-918): out[4*i + {0,1,2}] = integrand_{xx,yy,zz}, out[4*i + 3] = their sum (pdata).
// HBM-bound: 96 B in, 32 B out per point.
__global__ void __launch_bounds__(256) k_property_integrand(long n, const double *__restrict__ r, const double *__restrict__ tens,
                                                            int chi, double cx, double cy, double cz, double *__restrict__ out) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double dx = r[3 * i] - cx, dy = r[3 * i + 1] - cy, dz = r[3 * i + 2] - cz;
    const double *t = tens + 9 * i;
    double f;
    if (chi) f = 0.5;
    else f = 1.0e6 * (-1.0 / pow(dx * dx + dy * dy + dz * dz, 1.5) / (137.0359998 * 137.0359998));
    const double ixx = f * (dy * (-t[2]) - dz * (-t[1]));
    const double iyy = f * (dz * (-t[3]) - dx * (-t[5]));
    const double izz = f * (dx * (-t[7]) - dy * (-t[6]));
    double4 o; o.x = ixx; o.y = iyy; o.z = izz; o.w = ixx + iyy + izz;
    reinterpret_cast<double4 *>(out)[i] = o;
}
void launch_property_integrand(long n, const double *r, const double *tens, int chi, const double *c3, double *out, cudaStream_t s) {
    if (n <= 0) return;
    k_property_integrand<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, r, tens, chi, c3[0], c3[1], c3[2], out);
}

void launch_quadrature(const QuadArgs &q, cudaStream_t s) {
    if (q.nrows <= 0) { cudaMemsetAsync(q.out7, 0, 7 * sizeof(double), s); return; }
    k_quad_rows<<<(q.nrows + 3) / 4, 128, 0, s>>>(q);
    k_quad_final<<<1, 256, 0, s>>>(q.row_partials, q.nrows, q.out7);
}

}  // namespace gb
