// Preparation kernels: spatial sort keys, tile active sets (screening), basis-function panels.
//
// k_basis replaces calc_basis of the reference (src/libgimic/bfeval.f90:61-122, 295-338 with
// caos.f90:17-110 and basis.f90:118-136): one exp per primitive (the reference evaluates each four
// times), integer powers by repeated multiplication, and the *same* screening test
// sqrt(|r-R|^2) <= thr  so screened functions are exact zeros as in the reference.
#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

namespace gb {

__constant__ signed char c_lmn[2][6][21][3];

void upload_component_tables(const signed char *host_tab) { cudaMemcpyToSymbol(c_lmn, host_tab, 2 * 6 * 21 * 3); }

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread16(uint32_t v) {   // 16 bits -> every third bit of 48
    uint64_t x = v & 0xFFFFu;
    x = (x | (x << 16)) & 0x0000FF0000FFull;
    x = (x | (x << 8)) & 0x00F00F00F00Full;
    x = (x | (x << 4)) & 0x0C30C30C30C3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

// 48-bit 3-D Hilbert index (16 bits per axis, Skilling's axes-to-transpose).  A Hilbert curve is
// continuous, so any run of 128 consecutive sorted points is spatially compact; a Morton (Z) curve
// jumps, and the few tiles straddling a jump get a huge bounding sphere => nact ~ nbf => one CTA
// runs 100x longer than the rest (measured: SMs 5.7% active, profiles/r01_ncu_jtensor_a.txt).
__global__ void k_morton_keys(const double *__restrict__ r, long n, double lox, double loy, double loz, double inv_cell,
                              uint64_t *__restrict__ keys, int *__restrict__ vals) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t X[3];
    X[0] = (uint32_t)fmin(fmax((r[3 * i + 0] - lox) * inv_cell, 0.0), 65535.0);
    X[1] = (uint32_t)fmin(fmax((r[3 * i + 1] - loy) * inv_cell, 0.0), 65535.0);
    X[2] = (uint32_t)fmin(fmax((r[3 * i + 2] - loz) * inv_cell, 0.0), 65535.0);
    for (uint32_t Q = 1u << 15; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (X[d] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[d]) & P; X[0] ^= t; X[d] ^= t; }
        }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = 1u << 15; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    keys[i] = (spread16(X[0]) << 2) | (spread16(X[1]) << 1) | spread16(X[2]);
    vals[i] = (int)i;
}

void launch_morton_keys(const double *r, long n, const double *bbox_lo, double inv_cell, uint64_t *keys, int *vals, cudaStream_t s) {
    if (n <= 0) return;
    k_morton_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(r, n, bbox_lo[0], bbox_lo[1], bbox_lo[2], inv_cell, keys, vals);
}

size_t sort_temp_bytes(long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (const int *)nullptr, (int *)nullptr, (int)n, 0, 48);
    return bytes;
}
void launch_sort_pairs(void *temp, size_t temp_bytes, const uint64_t *kin, uint64_t *kout, const int *vin, int *vout, long n, cudaStream_t s) {
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout, (int)n, 0, 48, s);
}

__global__ void k_gather_points(const double *__restrict__ r, const int *__restrict__ perm, long n, double *__restrict__ rsx,
                                double *__restrict__ rsy, double *__restrict__ rsz) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    long p = perm[i];
    rsx[i] = r[3 * p + 0]; rsy[i] = r[3 * p + 1]; rsz[i] = r[3 * p + 2];
}
void launch_gather_points(const double *r, const int *perm, long n, double *rsx, double *rsy, double *rsz, cudaStream_t s) {
    if (n <= 0) return;
    k_gather_points<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(r, perm, n, rsx, rsy, rsz);
}

// gridpoint() of src/fgimic/grid.f90:498-511 for the flat index range [lo,hi), i fastest (grid.f90:478-495)
__global__ void k_grid_points(const double *__restrict__ ob, const double *__restrict__ p0, const double *__restrict__ p1,
                              const double *__restrict__ p2, int n0, int n1, int n2, long lo, long hi, double *__restrict__ r) {
    long idx = lo + blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= hi) return;
    long n01 = (long)n0 * n1;
    int k = (int)(idx / n01);
    int j = (int)((idx - k * n01) / n0);
    int i = (int)(idx - k * n01 - (long)j * n0);
    double a = p0[i], b = p1[j], c = p2[k];
    long o = idx - lo;
#pragma unroll
    for (int d = 0; d < 3; ++d) r[3 * o + d] = ob[d] + a * ob[3 + d] + b * ob[6 + d] + c * ob[9 + d];
}
void launch_grid_points(const double *ob, const double *p0, const double *p1, const double *p2, int n0, int n1, int n2, long lo, long hi,
                        double *r, cudaStream_t s) {
    long n = hi - lo;
    if (n <= 0) return;
    k_grid_points<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ob, p0, p1, p2, n0, n1, n2, lo, hi, r);
}

// ---------------------------------------------------------------------------------------------
// Screening of one atom against a tile: a shell can only be non-zero at some point p of the tile if |p - R| <= thr.
// Cheap reject by the distance to the tile's bounding box, then the exact minimum distance over the tile's points
// (staged in shared memory), so the active set is the exact union over the 128 points at shell granularity.  Shells are
// sorted by descending thr inside each atom, so the active set of an atom is a prefix.  k_tile_count and k_basis call this
// with identical inputs (same arithmetic => same counts); the per-point test sqrt(r2) <= thr is applied again in k_basis.
__device__ __forceinline__ void atom_active(const DevBasis &B, int a, const TileGeo &tg, const double *sx, const double *sy,
                                            const double *sz, int npts, int &nsh, int &nfun) {
    nsh = 0; nfun = 0;
    const double x = B.atom_xyz[3 * a], y = B.atom_xyz[3 * a + 1], z = B.atom_xyz[3 * a + 2];
    const double dx = fmax(fmax(tg.lox - x, x - tg.hix), 0.0), dy = fmax(fmax(tg.loy - y, y - tg.hiy), 0.0),
                 dz = fmax(fmax(tg.loz - z, z - tg.hiz), 0.0);
    const double mx = B.atom_maxthr[a];
    if (sqrt(dx * dx + dy * dy + dz * dz) - 1e-9 > mx) return;
    double d2 = 1e300;
    for (int p = 0; p < npts; ++p) {
        const double ex = sx[p] - x, ey = sy[p] - y, ez = sz[p] - z;
        d2 = fmin(d2, ex * ex + ey * ey + ez * ez);
    }
    const double lim = sqrt(d2) - 1e-9;
    if (lim > mx) return;
    int s1 = B.atom_shell_off[a + 1];
    for (int s = B.atom_shell_off[a]; s < s1; ++s) {
        if (B.sh_thr[s] >= lim) { int l = B.sh_l[s]; nfun += (l + 1) * (l + 2) / 2; ++nsh; }
        else break;
    }
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce_128(T v, Op op, T *s4) {   // 128 threads; result broadcast to all
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s4[threadIdx.x >> 5] = v;
    __syncthreads();
    return op(op(s4[0], s4[1]), op(s4[2], s4[3]));
}

__global__ void __launch_bounds__(128) k_tile_count(DevBasis B, const double *__restrict__ rsx, const double *__restrict__ rsy,
                                                    const double *__restrict__ rsz, const TileSeg *__restrict__ segs,
                                                    TileGeo *__restrict__ geo, TileInfo *__restrict__ info) {
    __shared__ double s4[4];
    __shared__ int i4[4];
    __shared__ unsigned long long u4[4];
    __shared__ double sx[MT], sy[MT], sz[MT];
    const TileSeg sg = segs[blockIdx.x];
    const bool valid = threadIdx.x < sg.npts;
    const long pt = sg.pt0 + (valid ? threadIdx.x : 0);
    const double x = rsx[pt], y = rsy[pt], z = rsz[pt];
    sx[threadIdx.x] = x; sy[threadIdx.x] = y; sz[threadIdx.x] = z;
    auto fmn = [](double a, double b) { return fmin(a, b); };
    auto fmx = [](double a, double b) { return fmax(a, b); };
    TileGeo tg;
    tg.lox = block_reduce_128(x, fmn, s4); tg.hix = block_reduce_128(x, fmx, s4);
    tg.loy = block_reduce_128(y, fmn, s4); tg.hiy = block_reduce_128(y, fmx, s4);
    tg.loz = block_reduce_128(z, fmn, s4); tg.hiz = block_reduce_128(z, fmx, s4);
    const double cx = 0.5 * (tg.lox + tg.hix), cy = 0.5 * (tg.loy + tg.hiy), cz = 0.5 * (tg.loz + tg.hiz);
    const double d = sqrt((x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz));
    const double rho = block_reduce_128(d, fmx, s4);
    tg.rho = rho; tg.pad_ = 0.0;
    // largest gap between consecutive points of the sorted run (where a curve jump would be cut)
    unsigned long long gi = 0;
    if ((int)threadIdx.x + 1 < sg.npts) {
        const double ex = rsx[pt + 1] - x, ey = rsy[pt + 1] - y, ez = rsz[pt + 1] - z;
        gi = ((unsigned long long)__float_as_uint((float)sqrt(ex * ex + ey * ey + ez * ez)) << 32) | threadIdx.x;
    }
    auto umx = [](unsigned long long a, unsigned long long b) { return a > b ? a : b; };
    gi = block_reduce_128(gi, umx, u4);
    int cnt = 0, nat = 0, nre = 0;
    __syncthreads();
    const int al = B.slot_align - 1;
    for (int a = threadIdx.x; a < B.natoms; a += 128) {
        int nsh, nfun; atom_active(B, a, tg, sx, sy, sz, sg.npts, nsh, nfun);
        cnt += (nfun + al) & ~al; nat += nfun > 0; nre += nfun;
    }
    auto iadd = [](int a, int b) { return a + b; };
    cnt = block_reduce_128(cnt, iadd, i4);
    nat = block_reduce_128(nat, iadd, i4);
    nre = block_reduce_128(nre, iadd, i4);
    if (threadIdx.x == 0) {
        geo[blockIdx.x] = tg;
        info[blockIdx.x] = TileInfo{(float)rho, __uint_as_float((unsigned)(gi >> 32)), (int)(gi & 0xffffffffu), cnt, nat, nre};
    }
}
void launch_tile_count(const DevBasis &B, const double *rsx, const double *rsy, const double *rsz, const TileSeg *segs, int ntiles,
                       TileGeo *geo, TileInfo *info, cudaStream_t s) {
    if (ntiles <= 0) return;
    k_tile_count<<<ntiles, 128, 0, s>>>(B, rsx, rsy, rsz, segs, geo, info);
}

// position in the Turbomole component order (gtodefs.f90:109-123) of the c-th component in the standard order (:86-106)
__device__ constexpr signed char TM_POS[6][21] = {
    {0},
    {0,1,2},
    {0,3,4,1,5,2},
    {0,3,4,5,9,7,1,6,8,2},
    {0,3,4,9,12,10,5,13,14,7,1,6,11,8,2},
    {0,3,4,9,15,10,11,18,19,13,5,16,20,17,7,1,6,12,14,8,2}};

// All components of one shell at one point, written to the 4 planes of the tile panel.  L is a template parameter so that
// after unrolling every power index is a compile-time constant (registers); the earlier version indexed px[lx] dynamically
// and generated 7 GB of local-memory traffic per launch (profiles/r01_ncu_final_summary.txt).
template <int L>
__device__ __forceinline__ void shell_to_panel(double rx, double ry, double rz, double q, double qp, bool on, bool tm,
                                               double *__restrict__ p0, long plane) {
    double px[L + 1], py[L + 1], pz[L + 1];
    px[0] = py[0] = pz[0] = 1.0;
#pragma unroll
    for (int k = 1; k <= L; ++k) { px[k] = px[k - 1] * rx; py[k] = py[k - 1] * ry; pz[k] = pz[k - 1] * rz; }
    int c = 0;
#pragma unroll
    for (int lx = L; lx >= 0; --lx)
#pragma unroll
        for (int ly = L - lx; ly >= 0; --ly, ++c) {
            const int lz = L - lx - ly;
            double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
            if (on) {
                const double ang = px[lx] * py[ly] * pz[lz];
                v0 = ang * q;                                      // cgto, caos.f90:31-36
                const double up = ang * qp;                        // dcgto, caos.f90:54-63: f_a x^(f-e_a) q - 2 x_a x^f q'
                v1 = (lx ? (double)lx * (px[lx > 0 ? lx - 1 : 0] * py[ly] * pz[lz]) * q : 0.0) - 2.0 * rx * up;
                v2 = (ly ? (double)ly * (px[lx] * py[ly > 0 ? ly - 1 : 0] * pz[lz]) * q : 0.0) - 2.0 * ry * up;
                v3 = (lz ? (double)lz * (px[lx] * py[ly] * pz[lz > 0 ? lz - 1 : 0]) * q : 0.0) - 2.0 * rz * up;
            }
            const long o = (long)(tm ? TM_POS[L][c] : c) * LDP;
            p0[o] = v0; p0[plane + o] = v1; p0[2 * plane + o] = v2; p0[3 * plane + o] = v3;
        }
}

// ---------------------------------------------------------------------------------------------
// k_basis: one CTA (128 threads = 128 points) per tile.
//  phase A: ordered compaction of the active atoms into (atom, nshell, slot0) runs + the slot->function index list
//  phase B: every thread evaluates its point for all active shells and writes the 4 planes
//           P0 = Phi, P1..3 = dPhi/dx,dy,dz at panel[(plane*nact + slot)*LDP + row]  (row-contiguous => coalesced)
__global__ void __launch_bounds__(128, 5) k_basis(DevBasis B, const TileDesc *__restrict__ tiles, const TileGeo *__restrict__ geo,
                                               const double *__restrict__ rsx, const double *__restrict__ rsy,
                                               const double *__restrict__ rsz, double *__restrict__ panel_pool,
                                               int *__restrict__ fidx_pool, TileAtom *__restrict__ atab_pool) {
    extern __shared__ int s_runs[];   // 4 ints per active atom: atom, shells, first K slot, first N column
    __shared__ int s_w[3][4];
    __shared__ int s_base[3];
    __shared__ int s_zero;            // a K slot whose panel rows are zero (padding), for the N-side padding columns
    const TileDesc td = tiles[blockIdx.x];
    if (td.nact == 0) return;
    const TileGeo tg = geo[td.geo];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *panel = panel_pool + td.panel_off;
    int *fidx = fidx_pool + td.fidx_off;
    const long plane = (long)td.nact * LDP;

    __shared__ double sx[MT], sy[MT], sz[MT];
    { const long p = td.pt0 + (tid < td.npts ? tid : 0); sx[tid] = rsx[p]; sy[tid] = rsy[p]; sz[tid] = rsz[p]; }
    if (tid == 0) { s_base[0] = 0; s_base[1] = 0; s_base[2] = 0; s_zero = td.nraw < td.nact ? td.nraw : 0x7fffffff; }
    __syncthreads();
    const int al = B.slot_align - 1;
    for (int a0 = 0; a0 < B.natoms; a0 += 128) {
        int a = a0 + tid, nsh = 0, nreal = 0;
        if (a < B.natoms) atom_active(B, a, tg, sx, sy, sz, td.npts, nsh, nreal);
        const int nfun = (nreal + al) & ~al;         // slots of the atom's run (functions + alignment padding)
        int flag = nfun > 0, sf = nfun, sr = flag, sn = nreal;   // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t1 = __shfl_up_sync(0xffffffffu, sf, o), t2 = __shfl_up_sync(0xffffffffu, sr, o), t3 = __shfl_up_sync(0xffffffffu, sn, o);
            if (lane >= o) { sf += t1; sr += t2; sn += t3; }
        }
        if (lane == 31) { s_w[0][wid] = sf; s_w[1][wid] = sr; s_w[2][wid] = sn; }
        __syncthreads();
        int offf = s_base[0], offr = s_base[1], offn = s_base[2];
        for (int w = 0; w < wid; ++w) { offf += s_w[0][w]; offr += s_w[1][w]; offn += s_w[2][w]; }
        if (flag) {
            int run = offr + sr - 1, slot0 = offf + sf - nfun;
            s_runs[4 * run] = a; s_runs[4 * run + 1] = nsh; s_runs[4 * run + 2] = slot0; s_runs[4 * run + 3] = offn + sn - nreal;
            if (nfun > nreal) atomicMin(&s_zero, slot0 + nreal);
        }
        __syncthreads();
        if (tid == 127) { s_base[0] = offf + sf; s_base[1] = offr + sr; s_base[2] = offn + sn; }
        __syncthreads();
    }
    const int nruns = s_base[1];
    // slot -> internal function index; pad slots point at a valid function (their Phi is zero)
    int *nlist = fidx + td.nact;   // N column -> K slot
    for (int rn = 0; rn < nruns; ++rn) {
        int a = s_runs[4 * rn], nsh = s_runs[4 * rn + 1], slot0 = s_runs[4 * rn + 2], col0 = s_runs[4 * rn + 3];
        int f0 = B.atom_func_off[a];
        int s_last = B.atom_shell_off[a] + nsh - 1, ll = B.sh_l[s_last];
        int nfun = B.sh_foff[s_last] - f0 + (ll + 1) * (ll + 2) / 2, nslot = (nfun + al) & ~al;
        for (int c = tid; c < nslot; c += 128) fidx[slot0 + c] = f0 + (c < nfun ? c : 0);
        for (int c = tid; c < nfun; c += 128) nlist[col0 + c] = slot0 + c;
    }
    for (int c = td.nraw + tid; c < td.nact; c += 128) fidx[c] = 0;
    for (int c = td.nreal + tid; c < td.nn; c += 128) nlist[c] = s_zero;   // nn > nreal implies that a padding slot exists
    // atom table for the GIAO taps of k_jtensor (see TileAtom)
    if (atab_pool) {
        TileAtom *atab = atab_pool + td.atab_off;
        const double cx = 0.5 * (tg.lox + tg.hix), cy = 0.5 * (tg.loy + tg.hiy), cz = 0.5 * (tg.loz + tg.hiz);
        for (int rn = tid; rn < nruns; rn += 128) {
            const int a = s_runs[4 * rn];
            const int slot_end = (rn + 1 < nruns) ? s_runs[4 * rn + 6] : td.nraw;
            double nx = cx, ny = cy, nz = cz;
            if (rn + 1 < nruns) { const int b = s_runs[4 * rn + 4]; nx = B.atom_xyz[3 * b]; ny = B.atom_xyz[3 * b + 1]; nz = B.atom_xyz[3 * b + 2]; }
            TileAtom ta;
            ta.dx = B.atom_xyz[3 * a] - nx; ta.dy = B.atom_xyz[3 * a + 1] - ny; ta.dz = B.atom_xyz[3 * a + 2] - nz;
            ta.kend4 = slot_end / 4; ta.atom = a;
            atab[rn] = ta;
        }
    }

    const int row = tid;
    const bool valid = row < td.npts;
    const long pt = td.pt0 + (valid ? row : 0);
    const double x = rsx[pt], y = rsy[pt], z = rsz[pt];
    const bool tm = B.turbomole != 0;

    for (int rn = 0; rn < nruns; ++rn) {
        const int a = s_runs[4 * rn], nsh = s_runs[4 * rn + 1], slot0 = s_runs[4 * rn + 2];
        const double rx = x - B.atom_xyz[3 * a], ry = y - B.atom_xyz[3 * a + 1], rz = z - B.atom_xyz[3 * a + 2];
        const double r2 = rx * rx + ry * ry + rz * rz;
        const double dist = sqrt(r2);                                  // filter_screened, basis.f90:127
        const int sA = B.atom_shell_off[a], f0 = B.atom_func_off[a];
        for (int s = sA; s < sA + nsh; ++s) {
            const int l = B.sh_l[s], np = B.sh_nprim[s], po = B.sh_prim_off[s];
            const int slot = slot0 + (B.sh_foff[s] - f0);
            double q = 0.0, qp = 0.0;
            const bool on = valid && (dist <= B.sh_thr[s]);            // basis.f90:130
            if (on) {
                for (int p = 0; p < np; ++p) {                         // cao2, caos.f90:94-110 (one exp, not four)
                    double al = B.alpha[po + p];
                    double e = B.ncc[po + p] * exp(-al * r2);
                    q += e; qp += al * e;
                }
            }
            double *p0 = panel + (long)slot * LDP + row;
            switch (l) {
                case 0: shell_to_panel<0>(rx, ry, rz, q, qp, on, tm, p0, plane); break;
                case 1: shell_to_panel<1>(rx, ry, rz, q, qp, on, tm, p0, plane); break;
                case 2: shell_to_panel<2>(rx, ry, rz, q, qp, on, tm, p0, plane); break;
                case 3: shell_to_panel<3>(rx, ry, rz, q, qp, on, tm, p0, plane); break;
                case 4: shell_to_panel<4>(rx, ry, rz, q, qp, on, tm, p0, plane); break;
                default: shell_to_panel<5>(rx, ry, rz, q, qp, on, tm, p0, plane); break;
            }
        }
        if (al) {   // zero rows of the run's alignment padding
            const int s_last = sA + nsh - 1, ll = B.sh_l[s_last];
            const int nfun = B.sh_foff[s_last] - f0 + (ll + 1) * (ll + 2) / 2, nslot = (nfun + al) & ~al;
            for (int c = nfun; c < nslot; ++c) {
                const long o = (long)(slot0 + c) * LDP + row;
                panel[o] = 0.0; panel[plane + o] = 0.0; panel[2 * plane + o] = 0.0; panel[3 * plane + o] = 0.0;
            }
        }
    }
    for (int c = td.nraw; c < td.nact; ++c) {
        const long o = (long)c * LDP + row;
        panel[o] = 0.0; panel[plane + o] = 0.0; panel[2 * plane + o] = 0.0; panel[3 * plane + o] = 0.0;
    }
}

// Dense evaluation for callers that want the basis vectors themselves (bfeval / dfdr of the reference,
// bfeval.f90:81-122,295-338): bf[i][f], dr[i][m][f] in the reference AO order, exact zeros where screened.
// One thread per (point, shell); not on the tensor hot path.
__global__ void k_basis_dense(DevBasis B, const int *__restrict__ f2user, long n, const double *__restrict__ r,
                              double *__restrict__ bf, double *__restrict__ dr) {
    const long i = blockIdx.x;
    const signed char(*lmn_tab)[21][3] = c_lmn[B.turbomole ? 1 : 0];
    for (int s = threadIdx.x; s < B.nshell; s += blockDim.x) {
        int a = 0;   // atom of shell s (binary search in atom_shell_off)
        { int lo = 0, hi = B.natoms; while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (B.atom_shell_off[mid] <= s) lo = mid; else hi = mid; } a = lo; }
        const double rx = r[3 * i] - B.atom_xyz[3 * a], ry = r[3 * i + 1] - B.atom_xyz[3 * a + 1], rz = r[3 * i + 2] - B.atom_xyz[3 * a + 2];
        const double r2 = rx * rx + ry * ry + rz * rz;
        const int l = B.sh_l[s], ncomp = (l + 1) * (l + 2) / 2;
        double q = 0.0, qp = 0.0;
        const bool on = sqrt(r2) <= B.sh_thr[s];
        if (on) for (int p = 0; p < B.sh_nprim[s]; ++p) { const double al = B.alpha[B.sh_prim_off[s] + p]; const double e = B.ncc[B.sh_prim_off[s] + p] * exp(-al * r2); q += e; qp += al * e; }
        double px[6], py[6], pz[6];
        px[0] = py[0] = pz[0] = 1.0;
        for (int k = 1; k < 6; ++k) { px[k] = px[k - 1] * rx; py[k] = py[k - 1] * ry; pz[k] = pz[k - 1] * rz; }
        for (int c = 0; c < ncomp; ++c) {
            const int lx = lmn_tab[l][c][0], ly = lmn_tab[l][c][1], lz = lmn_tab[l][c][2];
            double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (on) {
                const double ang = px[lx] * py[ly] * pz[lz], up = ang * qp;
                v0 = ang * q;
                v1 = (lx ? (double)lx * (px[lx - 1] * py[ly] * pz[lz]) * q : 0.0) - 2.0 * rx * up;
                v2 = (ly ? (double)ly * (px[lx] * py[ly - 1] * pz[lz]) * q : 0.0) - 2.0 * ry * up;
                v3 = (lz ? (double)lz * (px[lx] * py[ly] * pz[lz - 1]) * q : 0.0) - 2.0 * rz * up;
            }
            const long f = f2user[B.sh_foff[s] + c], nb = B.nbf;
            if (bf) bf[i * nb + f] = v0;
            if (dr) { dr[(i * 3 + 0) * nb + f] = v1; dr[(i * 3 + 1) * nb + f] = v2; dr[(i * 3 + 2) * nb + f] = v3; }
        }
    }
}
void launch_basis_dense(const DevBasis &B, const int *f2user, long n, const double *r, double *bf, double *dr, cudaStream_t s) {
    if (n <= 0) return;
    k_basis_dense<<<(unsigned)n, 128, 0, s>>>(B, f2user, n, r, bf, dr);
}

void launch_basis(const DevBasis &B, const TileDesc *tiles, int ntiles, const TileGeo *geo, const double *rsx, const double *rsy,
                  const double *rsz, double *panel_pool, int *fidx_pool, TileAtom *atab_pool, cudaStream_t s) {
    if (ntiles <= 0) return;
    size_t smem = (size_t)4 * B.natoms * sizeof(int);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_basis, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
    k_basis<<<ntiles, 128, smem, s>>>(B, tiles, geo, rsx, rsy, rsz, panel_pool, fidx_pool, atab_pool);
}

}  // namespace gb
