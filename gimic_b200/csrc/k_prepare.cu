// Preparation kernels: spatial sort keys, tile active sets (screening), basis-function panels.
//
// k_basis replaces calc_basis of the reference (src/libgimic/bfeval.f90:61-122, 295-338 with
// caos.f90:17-110 and basis.f90:118-136): one exp per primitive (the reference evaluates each four
// times), integer powers by repeated multiplication, and the *same* screening test
// sqrt(|r-R|^2) <= thr  so screened functions are exact zeros as in the reference.
#include <algorithm>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

namespace gb {

__constant__ signed char c_lmn[2][6][21][3];

void upload_component_tables(const signed char *host_tab) {
    cudaMemcpyToSymbol(c_lmn, host_tab, 2 * 6 * 21 * 3);
    cudaDeviceSynchronize();       // the copy runs on the legacy stream; the kernels that read the table run on non-blocking streams
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread16(uint32_t v) {   // 16 bits -> every third bit of 48
    uint64_t x = v & 0xFFFFu;
    x = (x | (x << 16)) & 0x0000FF0000FFull;
    x = (x | (x << 8)) & 0x00F00F00F00Full;
    x = (x | (x << 4)) & 0x0C30C30C30C3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

constexpr int KEY_DROP_BITS = 18;
// 48-bit 3-D Hilbert index (16 bits per axis, Skilling's axes-to-transpose), of which the key keeps the top 30 bits.  A Hilbert curve is
// continuous, so any run of 128 consecutive sorted points is spatially compact; a Morton (Z) curve
// jumps, and the few tiles straddling a jump get a huge bounding sphere => nact ~ nbf => one CTA
// runs 100x longer than the rest (measured: SMs 5.7% active, profiles/r01_ncu_jtensor_a.txt).
__global__ void k_morton_keys(const double *__restrict__ r, long n, double lox, double loy, double loz, double inv_cell,
                              uint32_t *__restrict__ keys, int *__restrict__ vals) {
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
    keys[i] = (uint32_t)(((spread16(X[0]) << 2) | (spread16(X[1]) << 1) | spread16(X[2])) >> KEY_DROP_BITS);
    vals[i] = (int)i;
}

void launch_morton_keys(const double *r, long n, const double *bbox_lo, double inv_cell, uint32_t *keys, int *vals, cudaStream_t s) {
    if (n <= 0) return;
    k_morton_keys<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(r, n, bbox_lo[0], bbox_lo[1], bbox_lo[2], inv_cell, keys, vals);
}

// Only the top 30 of the 48 index bits are kept (KEY_DROP_BITS = 18) and sorted on: 4 radix passes over 32-bit keys instead of 6 over
// 64-bit ones.  Cells of (box / 1024) ~ 0.1 bohr along the curve; the sort is stable, so points of one cell keep the caller's order --
// they are closer together than any tile is wide.
size_t sort_temp_bytes(long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const int *)nullptr, (int *)nullptr, (int)n, 0, 48 - KEY_DROP_BITS);
    return bytes;
}
void launch_sort_pairs(void *temp, size_t temp_bytes, const uint32_t *kin, uint32_t *kout, const int *vin, int *vout, long n, cudaStream_t s) {
    cub::DeviceRadixSort::SortPairs(temp, temp_bytes, kin, kout, vin, vout, (int)n, 0, 48 - KEY_DROP_BITS, s);
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
// Cheap reject by the distance to the tile's bounding box (one atom per lane), then, one atom per warp, a lower bound of the
// minimum distance over the tile's points (staged in shared memory, single precision about the tile centre, see below), so the
// active set is the union over the tile's points at shell granularity plus a ~1e-6 margin.  Shells are sorted by descending thr
// inside each atom, so the active set of an atom is a prefix.
// k_tile_split and k_basis call this with identical inputs; every operation below is written with explicit rounding-mode
// intrinsics so that both kernels compute bit-identical bounds (no context-dependent FMA contraction) => identical counts.  The
// exact per-point test sqrt(r2) <= thr is applied in k_basis (filter_screened, basis.f90:118-136).
// for_active_atoms: warp-cooperative sweep over the 32 atoms [base, base + 32).  Lane i loads atom base+i (position, largest radius,
// shell / function ranges) and tests its bounding box; the surviving atoms are then processed one at a time with their data broadcast
// by shuffles, so that the only dependent global-memory round per candidate is the one that fetches the shell radii.
// emit(atom, nsh, nfun) is called by all 32 lanes for every atom with nfun > 0.
// The minimum distance is searched in SINGLE precision on coordinates relative to the tile centre (FP64 vector instructions issue
// at a fraction of the FP32 rate and 64-bit shuffles are two instructions: the FP64 search was 45 % of k_tile_split's instructions,
// profiles/r02_ncu_tile_split.txt) and turned into a rigorous LOWER bound of the exact squared distance:
//   per component, fl32(x - c) - fl32(a - c) differs from x - a by at most 2^-23 (|x - c| + |a - c|), i.e. |delta| <= D :=
//   2.5e-7 (2 rho + d) with d the distance found, so d2 - (2 d D + D^2) - 3e-7 d2 (FMA chain) never exceeds the exact value.
// A lower bound only ever ADDS shells to the active set (by a relative margin of ~1e-6 in radius); the per-point test inside
// k_basis is the exact one.
struct TileFrame { double cx, cy, cz; float rho; };
__device__ __forceinline__ TileFrame tile_frame(const TileGeo &tg) {
    return TileFrame{0.5 * (tg.lox + tg.hix), 0.5 * (tg.loy + tg.hiy), 0.5 * (tg.loz + tg.hiz), __double2float_ru(tg.rho)};
}
template <class Emit>
__device__ __forceinline__ void for_active_atoms(const DevBasis &B, int base, const TileGeo &tg, const TileFrame &fr, const float *fx,
                                                 const float *fy, const float *fz, int np, Emit emit) {
    const int lane = threadIdx.x & 31;
    const int at = base + lane;
    // All tile-level tests compare SQUARED distances with (radius + 1e-9)^2 (precomputed): a shell is kept if (thr + 1e-9)^2 >= d2, a
    // superset of the per-point test sqrt(r2) <= thr of k_basis (the margin 2e-9 thr dwarfs the rounding of the squares).
    float ax_ = 0.f, ay_ = 0.f, az_ = 0.f;
    double mx = 0.0;
    int s0 = 0, s1 = 0, f0 = 0, f1 = 0;
    bool pass = false;
    if (at < B.natoms) {
        const double x = B.atom_xyz[3 * at], y = B.atom_xyz[3 * at + 1], z = B.atom_xyz[3 * at + 2];
        mx = B.atom_maxthr2e[at];
        s0 = B.atom_shell_off[at]; s1 = B.atom_shell_off[at + 1]; f0 = B.atom_func_off[at]; f1 = B.atom_func_off[at + 1];
        const double dx = fmax(fmax(tg.lox - x, x - tg.hix), 0.0), dy = fmax(fmax(tg.loy - y, y - tg.hiy), 0.0),
                     dz = fmax(fmax(tg.loz - z, z - tg.hiz), 0.0);
        pass = !(__fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx))) > mx);
        ax_ = __double2float_rn(x - fr.cx); ay_ = __double2float_rn(y - fr.cy); az_ = __double2float_rn(z - fr.cz);
    }
    unsigned bal = __ballot_sync(0xffffffffu, pass);
    while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        const float ax = __shfl_sync(0xffffffffu, ax_, src), ay = __shfl_sync(0xffffffffu, ay_, src), az = __shfl_sync(0xffffffffu, az_, src);
        const double amx = __shfl_sync(0xffffffffu, mx, src);
        const int as0 = __shfl_sync(0xffffffffu, s0, src), as1 = __shfl_sync(0xffffffffu, s1, src);
        const int af0 = __shfl_sync(0xffffffffu, f0, src), af1 = __shfl_sync(0xffffffffu, f1, src);
        float m = 3.0e38f;
        for (int p = lane; p < np; p += 32) {
            const float ex = __fsub_rn(fx[p], ax), ey = __fsub_rn(fy[p], ay), ez = __fsub_rn(fz[p], az);
            m = fminf(m, __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float dd = __fsqrt_ru(m), D = __fmul_ru(2.5e-7f, __fmaf_ru(2.0f, fr.rho, dd));
        const float slack = __fmaf_ru(3.0e-7f, m, __fmaf_ru(D, D, __fmul_ru(__fmul_ru(2.0f, dd), D)));
        const double d2 = (double)fmaxf(__fsub_rd(m, __fmul_ru(1.01f, slack)), 0.0f);      // <= the exact minimum squared distance
        if (d2 > amx) continue;
        // shells are sorted by descending radius inside an atom: the active ones are a prefix
        int cnt = 0, fend = af0;                  // active shells, function index after the last active shell
        for (int sb = as0; sb < as1; sb += 32) {
            const int s = sb + lane;
            const bool in = s < as1;
            const double thr = in ? B.sh_thr2e[s] : -1.0;
            const int fnext = in ? (s + 1 < as1 ? B.sh_foff[s + 1] : af1) : af1;
            const unsigned act = __ballot_sync(0xffffffffu, in && thr >= d2);
            const int k = act == 0xffffffffu ? 32 : __ffs(~act) - 1;
            if (k > 0) { cnt += k; fend = __shfl_sync(0xffffffffu, fnext, k - 1); }
            if (k < 32) break;
        }
        if (cnt > 0) emit(base + src, cnt, fend - af0);
    }
}

// ---- k_tile_split -----------------------------------------------------------------------------------------------------------
// One CTA (128 threads) per run of MT consecutive sorted points.  Depth-first, left piece first: evaluate a piece [a, b) of the
// run (bounding box, radius, largest consecutive gap, active slots / atoms / functions); a piece wider than split_radius whose
// largest gap exceeds half its radius is a candidate for a cut there (thin / planar point sets, cluster boundaries: a Hilbert run
// that leaves and re-enters the point cloud would otherwise drag in the active sets of both ends) and is cut if the two pieces
// cost the contraction less than the whole (piece_cost, kernels.cuh); everything else is emitted in order.
//
// minimum of 6 values over the 128 threads, in single precision: the callers pass values rounded DOWN, so the result bounds the exact
// minimum from below (the tile's box may grow by one float ulp per side, never shrink); result broadcast
__device__ __forceinline__ void block_min6(float (&v)[6], float (*s_red)[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] = fminf(v[i], __shfl_xor_sync(0xffffffffu, v[i], o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int i = 0; i < 6; ++i) s_red[threadIdx.x >> 5][i] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = fminf(fminf(s_red[0][i], s_red[1][i]), fminf(s_red[2][i], s_red[3][i]));
}

struct PieceEval { TileGeo tg; float rho, gmax; int imax, nraw, natom, nreal; };
__global__ void __launch_bounds__(128, 6) k_tile_split(DevBasis B, const double *__restrict__ rsx, const double *__restrict__ rsy,
                                                    const double *__restrict__ rsz, long n, double split_radius,
                                                    TileSeg *__restrict__ slot_seg, TileGeo *__restrict__ slot_geo,
                                                    TileInfo *__restrict__ slot_info, int *__restrict__ cnt_out) {
    constexpr int NSTACK = MAXSUB + SPLIT_DEPTH + 2;
    __shared__ double sx[MT], sy[MT], sz[MT];
    __shared__ float fx[MT], fy[MT], fz[MT];    // the piece in hand, relative to its centre
    __shared__ float s_red[4][6];
    __shared__ unsigned long long s_u[4];
    __shared__ int s_cnt[3];
    __shared__ int s_stack[NSTACK][4];          // first point, end, depth, "already evaluated"
    __shared__ PieceEval s_ev[NSTACK + 1];      // evaluated pieces on the stack; [NSTACK] = the piece in hand
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long run = blockIdx.x;
    const long pt0 = run * MT;
    const int np = (int)((n - pt0) < MT ? (n - pt0) : MT);
    { const long p = pt0 + (tid < np ? tid : 0); sx[tid] = rsx[p]; sy[tid] = rsy[p]; sz[tid] = rsz[p]; }
    int sp = 0, emitted = 0;               // identical in every thread (all decisions below are made on broadcast values)
    if (tid == 0) { s_stack[0][0] = 0; s_stack[0][1] = np; s_stack[0][2] = 0; s_stack[0][3] = 0; }
    sp = 1;
    __syncthreads();
    const int al = B.slot_align - 1;
    // box, radius, widest gap and active set of the points [a, b) of this run; every thread returns the same values
    auto evaluate = [&](int a, int b, PieceEval *dst) {
        PieceEval e;
        const bool in = tid >= a && tid < b;
        const double x = sx[tid], y = sy[tid], z = sz[tid];
        const float big = 3.0e38f;
        float v[6] = {in ? __double2float_rd(x) : big, in ? __double2float_rd(y) : big, in ? __double2float_rd(z) : big,
                      in ? __double2float_rd(-x) : big, in ? __double2float_rd(-y) : big, in ? __double2float_rd(-z) : big};
        block_min6(v, s_red);
        TileGeo &tg = e.tg;
        tg.lox = v[0]; tg.loy = v[1]; tg.loz = v[2]; tg.hix = -(double)v[3]; tg.hiy = -(double)v[4]; tg.hiz = -(double)v[5];
        const double cx = 0.5 * (tg.lox + tg.hix), cy = 0.5 * (tg.loy + tg.hiy), cz = 0.5 * (tg.loz + tg.hiz);
        fx[tid] = __double2float_rn(x - cx); fy[tid] = __double2float_rn(y - cy); fz[tid] = __double2float_rn(z - cz);   // visible after the barriers below
        // radius about the centre and the largest gap between consecutive points, packed as (float bits << 32 | index): one max-reduction
        // (squared distances as floats: monotone, so the maxima are those of the distances; the two square roots are taken once, below)
        unsigned long long key = 0;
        if (in) {
            const double d = (x - cx) * (x - cx) + (y - cy) * (y - cy) + (z - cz) * (z - cz);
            key = (unsigned long long)__float_as_uint((float)d) << 32;     // non-negative floats order like their bit patterns
        }
        unsigned long long gap = 0;
        if (tid >= a && tid + 1 < b) {
            const double ex = sx[tid + 1] - x, ey = sy[tid + 1] - y, ez = sz[tid + 1] - z;
            gap = ((unsigned long long)__float_as_uint((float)(ex * ex + ey * ey + ez * ez)) << 32) | (unsigned)(tid - a);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long k2 = __shfl_xor_sync(0xffffffffu, key, o), g2 = __shfl_xor_sync(0xffffffffu, gap, o);
            key = key > k2 ? key : k2; gap = gap > g2 ? gap : g2;
        }
        if (tid < 3) s_cnt[tid] = 0;
        __syncthreads();
        if (lane == 0) { s_u[wid] = key; }
        __syncthreads();
        key = s_u[0]; for (int w = 1; w < 4; ++w) key = key > s_u[w] ? key : s_u[w];
        __syncthreads();
        if (lane == 0) { s_u[wid] = gap; }
        __syncthreads();
        gap = s_u[0]; for (int w = 1; w < 4; ++w) gap = gap > s_u[w] ? gap : s_u[w];
        e.rho = sqrtf(__uint_as_float((unsigned)(key >> 32))); e.gmax = sqrtf(__uint_as_float((unsigned)(gap >> 32)));
        e.imax = (int)(gap & 0xffffffffu);
        tg.rho = (double)e.rho; tg.pad_ = 0.0;
        // active slots (atom runs aligned), atoms, functions
        const TileFrame fr = tile_frame(tg);
        int c0 = 0, c1 = 0, c2 = 0;
        for (int base = wid * 32; base < B.natoms; base += 128)
            for_active_atoms(B, base, tg, fr, fx + a, fy + a, fz + a, b - a, [&](int, int, int nfun) { c0 += (nfun + al) & ~al; c1 += 1; c2 += nfun; });
        if (lane == 0 && c1) { atomicAdd(&s_cnt[0], c0); atomicAdd(&s_cnt[1], c1); atomicAdd(&s_cnt[2], c2); }
        __syncthreads();
        e.nraw = s_cnt[0]; e.natom = s_cnt[1]; e.nreal = s_cnt[2];
        if (tid == 0) *dst = e;
        __syncthreads();                                   // s_cnt / s_u / s_red are reused by the next evaluation; *dst is visible
        return piece_cost(b - a, e.nraw, e.nreal, e.natom);
    };
    while (sp > 0) {
        --sp;
        const int a = s_stack[sp][0], b = s_stack[sp][1], depth = s_stack[sp][2];
        PieceEval *cur = &s_ev[NSTACK];
        if (s_stack[sp][3]) { if (tid == 0) *cur = s_ev[sp]; __syncthreads(); } else evaluate(a, b, cur);
        const float rho = cur->rho, gmax = cur->gmax;
        const int imax = cur->imax, nraw = cur->nraw;
        const long long cost = piece_cost(b - a, nraw, cur->nreal, cur->natom);
        const int npts = b - a;
        // A run of the sorted order that jumps (a curve piece that leaves a thin point set and comes back elsewhere) is cut at its
        // widest gap -- if the two pieces together cost the contraction less than the whole: their active sets must shrink by more
        // than the second pass over the (smaller) panels costs.  Round 2 first cut on the geometric test alone; plane grids then
        // fell into up to MAXSUB partially filled tiles whose active sets were barely smaller than the run's.
        bool split = depth < SPLIT_DEPTH && npts >= 16 && nraw > 0 && rho > (float)split_radius && gmax > 0.5f * rho &&
                     emitted + sp + 2 <= MAXSUB;
        const int m = a + imax + 1;
        if (split) {         // the pieces are evaluated straight into the stack slots they would take (slot sp has been read above)
            const long long cr = evaluate(m, b, &s_ev[sp]), cl = evaluate(a, m, &s_ev[sp + 1]);
            split = 20 * (cl + cr) < 17 * cost;
        }
        if (split) {
            if (tid == 0) {
                s_stack[sp][0] = m; s_stack[sp][1] = b; s_stack[sp][2] = depth + 1; s_stack[sp][3] = 1;                  // right piece: later
                s_stack[sp + 1][0] = a; s_stack[sp + 1][1] = m; s_stack[sp + 1][2] = depth + 1; s_stack[sp + 1][3] = 1;   // left piece: next
            }
            sp += 2;
        } else {
            if (tid == 0) {
                const long o = run * MAXSUB + emitted;
                slot_seg[o] = TileSeg{(int)(pt0 + a), npts};
                slot_geo[o] = cur->tg;
                slot_info[o] = TileInfo{rho, gmax, imax, nraw, cur->natom, cur->nreal};
            }
            ++emitted;
        }
        __syncthreads();
    }
    if (tid == 0) cnt_out[run] = emitted;
}

// Exclusive prefix sums (out may alias in; out[n] = total): tiles of SCAN_TILE elements per CTA (k_scan_partial_*), one CTA over the
// tile totals (k_scan_tops_*), offsets added back (k_scan_add_*).  The element count may live on the device (n_dev).  Round 2 first
// used ONE CTA for the whole array: 0.88 ms per plan of the 256^3 grid (131 072 runs), a quarter of the replicated per-rank plan.
constexpr int SCAN_TPB = 256, SCAN_IPT = 8, SCAN_TILE = SCAN_TPB * SCAN_IPT;
struct AddI { __device__ __forceinline__ int operator()(int a, int b) const { return a + b; } };
struct AddC { __device__ __forceinline__ TileCum operator()(TileCum a, TileCum b) const { return TileCum{a.cost + b.cost, a.panel + b.panel, a.fidx + b.fidx, a.atab + b.atab}; } };

// One CTA: exclusive scan of in[0, n) into out (may alias in); the total is left in s_part[blockDim.x - 1] and, if write_total, in out[n].
template <typename T, typename Add>
__device__ __forceinline__ void block_scan_exclusive(const T *in, T *out, int n, T zero, Add add, T *s_part /*[blockDim.x]*/, bool write_total) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int chunk = (n + nt - 1) / nt;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    T acc = zero;
    for (int i = lo; i < hi; ++i) acc = add(acc, in[i]);
    s_part[tid] = acc;
    __syncthreads();
    for (int o = 1; o < nt; o <<= 1) {       // Hillis-Steele over the per-thread sums
        T v = zero;
        if (tid >= o) v = s_part[tid - o];
        __syncthreads();
        if (tid >= o) s_part[tid] = add(v, s_part[tid]);
        __syncthreads();
    }
    T run = tid ? s_part[tid - 1] : zero;
    for (int i = lo; i < hi; ++i) { const T v = in[i]; out[i] = run; run = add(run, v); }
    if (write_total && tid == nt - 1) out[n] = s_part[nt - 1];
}
template <typename T, typename Add>
__device__ __forceinline__ void scan_partial_body(const T *in, T *out, int n, T zero, Add add, T *tops, T *s_part) {
    const int base = blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    const int m = min(SCAN_TILE, n - base);
    block_scan_exclusive<T>(in + base, out + base, m, zero, add, s_part, false);
    if (threadIdx.x == blockDim.x - 1) tops[blockIdx.x] = s_part[blockDim.x - 1];
}
template <typename T, typename Add>
__device__ __forceinline__ void scan_add_body(T *out, int n, Add add, const T *tops_excl) {
    const int base = blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    const int m = min(SCAN_TILE, n - base);
    const T off = tops_excl[blockIdx.x];
    for (int i = threadIdx.x; i < m; i += blockDim.x) out[base + i] = add(off, out[base + i]);
    if (base + m == n && threadIdx.x == 0) out[n] = tops_excl[(n + SCAN_TILE - 1) / SCAN_TILE];
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_partial_i(const int *in, int *out, int n, int *__restrict__ tops) {
    __shared__ int s_part[SCAN_TPB];
    scan_partial_body<int>(in, out, n, 0, AddI(), tops, s_part);
}
__global__ void __launch_bounds__(1024) k_scan_tops_i(int *tops, int n) {
    __shared__ int s_part[1024];
    const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    block_scan_exclusive<int>(tops, tops, nb, 0, AddI(), s_part, true);
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_add_i(int *out, int n, const int *__restrict__ tops, PlanSummary *sum, int cap) {
    scan_add_body<int>(out, n, AddI(), tops);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int tot = tops[(n + SCAN_TILE - 1) / SCAN_TILE]; sum->ntiles = tot < cap ? tot : cap; sum->overflow = tot > cap;
        // the drain-group table is filled by ALL blocks of k_plan_range: it is cleared here, by one thread of an earlier kernel
        for (int i = 0; i <= DRAIN_BATCHES * DRAIN_GROUPS; ++i) { sum->group_tile[i] = -1; sum->group_pt[i] = 0; }
    }
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_partial_c(const TileCum *in, TileCum *out, const PlanSummary *sum, TileCum *__restrict__ tops) {
    __shared__ TileCum s_part[SCAN_TPB];
    scan_partial_body<TileCum>(in, out, sum->ntiles, TileCum{0, 0, 0, 0}, AddC(), tops, s_part);
}
__global__ void __launch_bounds__(1024) k_scan_tops_c(TileCum *tops, const PlanSummary *sum) {
    __shared__ TileCum s_part[1024];
    const int nb = (sum->ntiles + SCAN_TILE - 1) / SCAN_TILE;
    block_scan_exclusive<TileCum>(tops, tops, nb, TileCum{0, 0, 0, 0}, AddC(), s_part, true);
}
__global__ void __launch_bounds__(SCAN_TPB) k_scan_add_c(TileCum *out, const PlanSummary *sum, const TileCum *__restrict__ tops) {
    scan_add_body<TileCum>(out, sum->ntiles, AddC(), tops);
}

// compact tiles in Hilbert order + their sizes; one thread per initial run
__global__ void k_tile_emit(const TileSeg *__restrict__ slot_seg, const TileGeo *__restrict__ slot_geo, const TileInfo *__restrict__ slot_info,
                            const int *__restrict__ cnt, const int *__restrict__ off, long nrun0, int cap, TileGeo *__restrict__ geo,
                            TileDesc *__restrict__ desc, TileCum *__restrict__ cum) {
    const long run = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (run >= nrun0) return;
    const int c = cnt[run], o = off[run];
    for (int j = 0; j < c && o + j < cap; ++j) {
        const int t = o + j;
        const TileSeg sg = slot_seg[run * MAXSUB + j];
        const TileInfo ti = slot_info[run * MAXSUB + j];
        geo[t] = slot_geo[run * MAXSUB + j];
        TileDesc td;
        td.pt0 = sg.pt0; td.npts = sg.npts; td.nraw = ti.nraw; td.nact = (ti.nraw + 7) / 8 * 8;
        td.nreal = ti.nreal; td.nn = (ti.nreal + 7) / 8 * 8; td.geo = t; td.nruns = ti.natom;
        td.panel_off = 0; td.fidx_off = 0; td.atab_off = 0;
        td.col0 = 0; td.col1 = td.nn; td.part = -1; td.pad_ = 0;
        desc[t] = td;
        TileCum tc;
        // scheduling cost in units of one DMMA column step: MMA k-steps x columns (4 planes) + GIAO taps + the tile's share of k_basis
        // (4 planes x 132 doubles written per K slot at ~4.8 TB/s against 32 TF for the contraction: ~110 units per slot) + a constant
        // per tile (descriptor, prologue, stores).  The last two were fitted on the 8-GPU run of the 256^3 grid, where ranks holding
        // many cheap tiles far from the molecule ran 4 % longer than ranks of equal flops (profiles/r02_bench_n8_grid_h.json).
        // Integers, so every rank computes the same prefix sums and hence the same partition.
        tc.cost = piece_cost(td.npts, td.nraw, td.nreal, td.nruns);
        tc.panel = 4LL * td.nact * LDP; tc.fidx = td.nact + td.nn; tc.atab = td.nruns;
        cum[t] = tc;
    }
}

// this rank's share: tiles whose exclusive cost prefix falls into [rank, rank + 1) x total / nranks; panel-pool batches of that range
__global__ void __launch_bounds__(256) k_plan_range(const TileDesc *__restrict__ desc, const TileCum *__restrict__ cum, int rank, int nranks,
                                                    long long pool_doubles, PlanSummary *sum, int drain_batches, int drain_min_tiles) {
    const int ntiles = sum->ntiles;
    __shared__ int s_lo, s_hi;
    __shared__ long long s_chunk;
    if (threadIdx.x == 0) {
        const long long total = cum[ntiles].cost;
        auto first_at_least = [&](long long target) {     // first t with cum[t].cost * nranks >= target (cum is non-decreasing)
            int lo = 0, hi = ntiles;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (cum[mid].cost * nranks >= target) hi = mid; else lo = mid + 1; }
            return lo;
        };
        const int tlo = rank == 0 ? 0 : first_at_least((long long)rank * total), thi = rank + 1 >= nranks ? ntiles : first_at_least((long long)(rank + 1) * total);
        s_lo = tlo; s_hi = thi;
        sum->tlo = tlo; sum->thi = thi;
        sum->cost_total = total; sum->cost_range = cum[thi].cost - cum[tlo].cost;
        sum->panel_range = cum[thi].panel - cum[tlo].panel;
        sum->pt_lo = tlo < thi ? desc[tlo].pt0 : 0;
        sum->pt_hi = tlo < thi ? (long long)desc[thi - 1].pt0 + desc[thi - 1].npts : 0;
        const long long nb = tlo < thi ? (cum[thi - 1].panel - cum[tlo].panel) / pool_doubles + 1 : 0;
        sum->nbatch = (int)(nb < MAX_BATCH ? nb : MAX_BATCH + 1);      // MAX_BATCH + 1: too many, the host refuses
        sum->max_nruns = 0; sum->max_tile_panel = 0;
        sum->sum_nact = sum->flops4 = sum->flops2 = sum->taps = sum->useful_mm = sum->useful_taps = 0.0;
        if (nb <= MAX_BATCH) { sum->batch_start[nb] = thi; sum->batch_pt[nb] = sum->pt_hi; }
        // drain groups: quarters of the (fullest) batch, never less than 128 MB of panels; only if the range has at most drain_batches
        // batches (with more, finished batches are drained while later ones run: measured at N = 2, 4 batches per rank, extra groups
        // only add launch tails) and at least drain_min_tiles tiles
        const long long full = sum->panel_range < pool_doubles ? sum->panel_range : pool_doubles;
        long long chunk = (full + DRAIN_GROUPS - 1) / DRAIN_GROUPS;
        if (chunk < (16LL << 20)) chunk = 16LL << 20;
        s_chunk = (nb <= drain_batches && thi - tlo >= drain_min_tiles) ? chunk : pool_doubles;
        sum->drain_chunk = s_chunk;          // every block writes the same values here; the group table below was cleared by k_scan_add_i
    }
    __syncthreads();
    const int tlo = s_lo, thi = s_hi;
    const long long base = cum[tlo].panel, chunk = s_chunk;
    for (int t = tlo + threadIdx.x + blockIdx.x * blockDim.x; t < thi; t += blockDim.x * gridDim.x) {
        const long long b = (cum[t].panel - base) / pool_doubles;
        if (b < MAX_BATCH && (t == tlo || (cum[t - 1].panel - base) / pool_doubles != b)) { sum->batch_start[b] = t; sum->batch_pt[b] = desc[t].pt0; }
        if (b < DRAIN_BATCHES) {
            const int g = drain_group(cum[t].panel - base, pool_doubles, chunk);
            const bool first = t == tlo || (cum[t - 1].panel - base) / pool_doubles != b || drain_group(cum[t - 1].panel - base, pool_doubles, chunk) != g;
            if (first) { sum->group_tile[b * DRAIN_GROUPS + g] = t; sum->group_pt[b * DRAIN_GROUPS + g] = desc[t].pt0; }
        }
    }
}

// offsets inside the batch, scheduling keys (batch, longest first) and the statistics of the range
__global__ void __launch_bounds__(256) k_plan_finalize(TileDesc *__restrict__ desc, const TileCum *__restrict__ cum, long long pool_doubles,
                                                       PlanSummary *sum, unsigned long long *__restrict__ keys, int *__restrict__ ord, int order_bits,
                                                       unsigned long long *__restrict__ gkeys) {
    const int tlo = sum->tlo, thi = sum->thi;
    const int t = tlo + blockIdx.x * blockDim.x + threadIdx.x;
    double st[6] = {0, 0, 0, 0, 0, 0};
    int mr = 0; long long mp = 0;
    if (t < thi && sum->nbatch <= MAX_BATCH) {
        const long long b = (cum[t].panel - cum[tlo].panel) / pool_doubles;
        const int t0 = sum->batch_start[b];
        TileDesc td = desc[t];
        td.panel_off = cum[t].panel - cum[t0].panel; td.fidx_off = cum[t].fidx - cum[t0].fidx; td.atab_off = cum[t].atab - cum[t0].atab;
        desc[t] = td;
        const long long c = cum[t + 1].cost - cum[t].cost;
        // Processing order inside a batch: costliest first (the contraction pulls tiles from an atomic counter: short tiles fill the
        // tail).  order_bits >= 0 (GIMIC_B200_ORDER_BITS, A/B runs): cost classes of 2^-order_bits relative width instead of exact costs;
        // the sort is stable, so tiles of one class keep their Hilbert order and spatial neighbours -- which gather nearly the same
        // density elements -- run at the same time.
        unsigned c32 = (unsigned)(c > 0xffffffffLL ? 0xffffffffLL : c);
        if (order_bits >= 0 && c > 0) {
            const int e = 63 - __clzll(c);
            const unsigned m = e > order_bits ? (unsigned)((c >> (e - order_bits)) & ((1LL << order_bits) - 1)) : (unsigned)(c & ((1LL << order_bits) - 1));
            c32 = ((unsigned)e << order_bits) | m;
        }
        keys[t - tlo] = ((unsigned long long)b << 32) | (0xffffffffu - c32);
        // second order, used only when host outputs are drained group by group: (batch, drain group, costliest first)
        const unsigned long long g = b < DRAIN_BATCHES ? (unsigned long long)drain_group(cum[t].panel - cum[tlo].panel, pool_doubles, sum->drain_chunk) : 0ull;
        gkeys[t - tlo] = ((unsigned long long)b << 36) | (g << 32) | (0xffffffffu - c32);
        ord[t - tlo] = t;
        st[0] = td.nact; st[1] = 2.0 * MT * 4.0 * td.nact * td.nn; st[2] = 2.0 * MT * 2.0 * td.nact * td.nn;
        st[3] = 2.0 * MT * (double)td.nn * td.nruns;                       // per tap weight (x3 tensor path, x1 J path)
        st[4] = 2.0 * td.npts * (double)td.nreal * td.nreal;               // per plane
        st[5] = 2.0 * td.npts * (double)td.nreal * td.nruns;
        mr = td.nruns; mp = cum[t + 1].panel - cum[t].panel;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 6; ++i) st[i] += __shfl_xor_sync(0xffffffffu, st[i], o);
        mr = max(mr, __shfl_xor_sync(0xffffffffu, mr, o));
        const long long m2 = __shfl_xor_sync(0xffffffffu, mp, o); mp = mp > m2 ? mp : m2;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&sum->sum_nact, st[0]); atomicAdd(&sum->flops4, st[1]); atomicAdd(&sum->flops2, st[2]); atomicAdd(&sum->taps, st[3]);
        atomicAdd(&sum->useful_mm, st[4]); atomicAdd(&sum->useful_taps, st[5]);
        atomicMax(&sum->max_nruns, mr); atomicMax(reinterpret_cast<unsigned long long *>(&sum->max_tile_panel), (unsigned long long)mp);
    }
}

__global__ void k_tile_gather(const TileDesc *__restrict__ desc, const int *__restrict__ ord, int nt, TileDesc *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) out[i] = desc[ord[i]];
}
// work items of sliced tiles: item i*S + s contracts columns [s*w, min(nn, (s+1)*w)) of tile i, w = slice_width (kernels.cuh); slices
// past the last column are empty (col1 <= col0: every role skips them)
__global__ void k_tile_slices(const TileDesc *__restrict__ tiles, int nt, int S, long long item_cost, TileDesc *__restrict__ items) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nt * S) return;
    const int i = j / S, s = j - i * S;
    TileDesc td = tiles[i];
    if (td.nact == 0) { if (s) { td.col0 = 0; td.col1 = 0; td.part = j; } items[j] = td; return; }     // slice 0 stores the zeros, the others do nothing
    const int w = slice_width(td, S, item_cost);
    td.col0 = min(s * w, td.nn); td.col1 = min(td.nn, td.col0 + w); td.part = j;
    items[j] = td;
}
void launch_tile_slices(const TileDesc *tiles, int nt, int S, long long item_cost, TileDesc *items, cudaStream_t s) {
    if (nt <= 0) return;
    k_tile_slices<<<(unsigned)((nt * S + 255) / 256), 256, 0, s>>>(tiles, nt, S, item_cost, items);
}

__global__ void k_perm_index(const int *__restrict__ perm, long n, long *__restrict__ index) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) index[i] = perm[i];
}

// Everything between the sorted points and the tile list, queued on one stream without a host round trip:
// split -> scan -> emit -> scan -> range/batches -> offsets/keys.  The host then reads PlanSummary once.
void launch_plan_tiles(const DevBasis &B, const double *rsx, const double *rsy, const double *rsz, long n, double split_radius,
                       int rank, int nranks, long long pool_doubles, const PlanBuffers &pb, cudaStream_t s) {
    const long nrun0 = (n + MT - 1) / MT;
    if (nrun0 <= 0) return;
    k_tile_split<<<(unsigned)nrun0, 128, 0, s>>>(B, rsx, rsy, rsz, n, split_radius, pb.slot_seg, pb.slot_geo, pb.slot_info, pb.cnt);
    const int nb0 = (int)((nrun0 + SCAN_TILE - 1) / SCAN_TILE), nbc = (pb.cap + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_partial_i<<<nb0, SCAN_TPB, 0, s>>>(pb.cnt, pb.off, (int)nrun0, pb.tops_i);
    k_scan_tops_i<<<1, 1024, 0, s>>>(pb.tops_i, (int)nrun0);
    k_scan_add_i<<<nb0, SCAN_TPB, 0, s>>>(pb.off, (int)nrun0, pb.tops_i, pb.summary, pb.cap);
    k_tile_emit<<<(unsigned)((nrun0 + 127) / 128), 128, 0, s>>>(pb.slot_seg, pb.slot_geo, pb.slot_info, pb.cnt, pb.off, nrun0, pb.cap, pb.geo, pb.desc, pb.cum);
    k_scan_partial_c<<<nbc, SCAN_TPB, 0, s>>>(pb.cum, pb.cum, pb.summary, pb.tops_c);
    k_scan_tops_c<<<1, 1024, 0, s>>>(pb.tops_c, pb.summary);
    k_scan_add_c<<<nbc, SCAN_TPB, 0, s>>>(pb.cum, pb.summary, pb.tops_c);
    // GIMIC_B200_DRAIN_BATCHES (default 1; 0 = never group, up to DRAIN_BATCHES), GIMIC_B200_DRAIN_MIN_TILES (default 2048): see PlanSummary
    const char *edb = std::getenv("GIMIC_B200_DRAIN_BATCHES"), *edt = std::getenv("GIMIC_B200_DRAIN_MIN_TILES");
    const int drain_batches = edb ? std::min(std::max(std::atoi(edb), 0), DRAIN_BATCHES) : 1, drain_min_tiles = edt ? std::max(std::atoi(edt), 1) : 2048;
    k_plan_range<<<64, 256, 0, s>>>(pb.desc, pb.cum, rank, nranks, pool_doubles, pb.summary, drain_batches, drain_min_tiles);
    static const int order_bits = [] { const char *e = std::getenv("GIMIC_B200_ORDER_BITS"); return e ? std::atoi(e) : -1; }();
    k_plan_finalize<<<(unsigned)((pb.cap + 255) / 256), 256, 0, s>>>(pb.desc, pb.cum, pool_doubles, pb.summary, pb.keys0, pb.ord0, order_bits, pb.gkeys0);
}
size_t plan_sort_temp_bytes(int nt) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (const int *)nullptr, (int *)nullptr, nt);
    return bytes;
}
// processing order of this rank's nt tiles: batch by batch, longest first inside a batch (the contraction kernel pulls tiles from an atomic counter)
void launch_plan_order(const PlanBuffers &pb, int tlo, int nt, bool grouped, void *sorttmp, size_t sorttmp_bytes, cudaStream_t s) {
    if (nt <= 0) return;
    (void)tlo;
    cub::DeviceRadixSort::SortPairs(sorttmp, sorttmp_bytes, pb.keys0, pb.keys1, pb.ord0, pb.ord1, nt, 0, 64, s);
    k_tile_gather<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(pb.desc, pb.ord1, nt, pb.tiles);
    if (grouped) {       // the drain-group order of the same tiles (keys1 / ord1 are free again: stream order)
        cub::DeviceRadixSort::SortPairs(sorttmp, sorttmp_bytes, pb.gkeys0, pb.keys1, pb.ord0, pb.ord1, nt, 0, 64, s);
        k_tile_gather<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(pb.desc, pb.ord1, nt, pb.gtiles);
    }
}
void launch_perm_index(const int *perm, long n, long *index, cudaStream_t s) {
    if (n <= 0) return;
    k_perm_index<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(perm, n, index);
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
// k_basis<NT>: NT threads = NT points, MT / NT CTAs per tile (NT = 128: one CTA per tile, the only instantiation.  A single-warp variant
// that ran beside the contraction kernel on a second stream was measured 5 % SLOWER end to end and removed: profiles/r02_ab_ncw_overlap.json).
//  phase A: ordered compaction of the active atoms into (atom, nshell, slot0) runs + the slot->function index list (written by part 0)
//  phase B: every thread evaluates its point for all active shells and writes the 4 planes
//           P0 = Phi, P1..3 = dPhi/dx,dy,dz at panel[(plane*nact + slot)*LDP + row]  (row-contiguous => coalesced)
template <int NT>
__global__ void __launch_bounds__(NT, 5) k_basis(DevBasis B, const TileDesc *__restrict__ tiles, const TileGeo *__restrict__ geo,
                                               const double *__restrict__ rsx, const double *__restrict__ rsy,
                                               const double *__restrict__ rsz, double *__restrict__ panel_pool,
                                               int *__restrict__ fidx_pool, TileAtom *__restrict__ atab_pool) {
    constexpr int Q = MT / NT;        // CTAs per tile
    extern __shared__ int s_dyn[];    // [natoms] packed (shells << 20 | functions) of every atom, then 4 ints per active atom (run)
    int *s_atom = s_dyn, *s_runs = s_dyn + B.natoms;   // run: atom, shells, first K slot, first N column
    __shared__ int s_w[3][4];
    __shared__ int s_base[3];
    __shared__ int s_zero;            // a K slot whose panel rows are zero (padding), for the N-side padding columns
    const TileDesc td = tiles[blockIdx.x / Q];
    const int part = blockIdx.x % Q;
    const int gy = blockIdx.y, G = gridDim.y;      // few tiles: the active atoms of a tile are dealt to G CTAs (launch_basis)
    if (td.nact == 0) return;
    const TileGeo tg = geo[td.geo];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double *panel = panel_pool + td.panel_off;
    int *fidx = fidx_pool + td.fidx_off;
    const long plane = (long)td.nact * LDP;

    __shared__ double sx[MT], sy[MT], sz[MT];
    __shared__ float fx[MT], fy[MT], fz[MT];          // the same points relative to the tile centre (for_active_atoms)
    const TileFrame fr = tile_frame(tg);
    for (int i = tid; i < MT; i += NT) {
        const long p = td.pt0 + (i < td.npts ? i : 0);
        const double x = rsx[p], y = rsy[p], z = rsz[p];
        sx[i] = x; sy[i] = y; sz[i] = z;
        fx[i] = __double2float_rn(x - fr.cx); fy[i] = __double2float_rn(y - fr.cy); fz[i] = __double2float_rn(z - fr.cz);
    }
    if (tid == 0) { s_base[0] = 0; s_base[1] = 0; s_base[2] = 0; s_zero = td.nraw < td.nact ? td.nraw : 0x7fffffff; }
    for (int a = tid; a < B.natoms; a += NT) s_atom[a] = 0;
    __syncthreads();
    // phase A1: active prefix of every atom (same arithmetic as k_tile_split => the counts the descriptors were sized with)
    for (int base = wid * 32; base < B.natoms; base += NT)
        for_active_atoms(B, base, tg, fr, fx, fy, fz, td.npts, [&](int aa, int nsh, int nfun) { if (lane == 0) s_atom[aa] = (nsh << 20) | nfun; });
    __syncthreads();
    const int al = B.slot_align - 1;
    for (int a0 = 0; a0 < B.natoms; a0 += NT) {
        int a = a0 + tid, nsh = 0, nreal = 0;
        if (a < B.natoms) { const int pk = s_atom[a]; nsh = pk >> 20; nreal = pk & 0xfffff; }
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
        if (tid == NT - 1) { s_base[0] = offf + sf; s_base[1] = offr + sr; s_base[2] = offn + sn; }
        __syncthreads();
    }
    const int nruns = s_base[1];
    // slot -> internal function index; pad slots point at a valid function (their Phi is zero)
    int *nlist = fidx + td.nact;   // N column -> K slot
    const bool lists = part == 0 && gy == 0;       // one CTA of the tile writes the index lists and the atom table
    for (int rn = 0; rn < nruns && lists; ++rn) {
        int a = s_runs[4 * rn], nsh = s_runs[4 * rn + 1], slot0 = s_runs[4 * rn + 2], col0 = s_runs[4 * rn + 3];
        int f0 = B.atom_func_off[a];
        int s_last = B.atom_shell_off[a] + nsh - 1, ll = B.sh_l[s_last];
        int nfun = B.sh_foff[s_last] - f0 + (ll + 1) * (ll + 2) / 2, nslot = (nfun + al) & ~al;
        for (int c = tid; c < nslot; c += NT) fidx[slot0 + c] = f0 + (c < nfun ? c : 0);
        for (int c = tid; c < nfun; c += NT) nlist[col0 + c] = slot0 + c;
    }
    for (int c = td.nraw + tid; c < td.nact && lists; c += NT) fidx[c] = 0;
    for (int c = td.nreal + tid; c < td.nn && lists; c += NT) nlist[c] = s_zero;   // nn > nreal implies that a padding slot exists
    // atom table for the GIAO taps of k_jtensor (see TileAtom)
    if (atab_pool && lists) {
        TileAtom *atab = atab_pool + td.atab_off;
        const double cx = 0.5 * (tg.lox + tg.hix), cy = 0.5 * (tg.loy + tg.hiy), cz = 0.5 * (tg.loz + tg.hiz);
        for (int rn = tid; rn < nruns; rn += NT) {
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

    const int row = part * NT + tid;
    const bool valid = row < td.npts;
    const long pt = td.pt0 + (valid ? row : 0);
    const double x = rsx[pt], y = rsy[pt], z = rsz[pt];
    const bool tm = B.turbomole != 0;

    for (int rn = gy; rn < nruns; rn += G) {
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
    for (int c = td.nraw; c < td.nact && gy == 0; ++c) {
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

// Diagnostic: the panels k_basis wrote for the contraction, scattered into dense reference-order arrays bf[i][f], dr[i][m][f]
// (zeros elsewhere) -- the very numbers k_jtensor's TMA copies and epilogue loads consume.  One CTA per tile.
__global__ void k_panel_scatter(const TileDesc *__restrict__ tiles, const double *__restrict__ panel_pool, const int *__restrict__ fidx_pool,
                                const int *__restrict__ perm, const int *__restrict__ f2user, int nbf, double *__restrict__ bf, double *__restrict__ dr) {
    const TileDesc td = tiles[blockIdx.x];
    if (td.nact == 0) return;
    const double *panel = panel_pool + td.panel_off;
    const int *fidx = fidx_pool + td.fidx_off, *nlist = fidx + td.nact;
    const long plane = (long)td.nact * LDP;
    for (int e = threadIdx.x; e < td.nreal * td.npts; e += blockDim.x) {
        const int col = e / td.npts, row = e - col * td.npts;
        const int slot = nlist[col];
        const long f = f2user[fidx[slot]], i = perm[td.pt0 + row];
        const double *p = panel + (long)slot * LDP + row;
        if (bf) bf[i * nbf + f] = p[0];
        if (dr) { dr[(i * 3 + 0) * nbf + f] = p[plane]; dr[(i * 3 + 1) * nbf + f] = p[2 * plane]; dr[(i * 3 + 2) * nbf + f] = p[3 * plane]; }
    }
}
void launch_panel_scatter(const TileDesc *tiles, int ntiles, const double *panel_pool, const int *fidx_pool, const int *perm, const int *f2user,
                          int nbf, double *bf, double *dr, cudaStream_t s) {
    if (ntiles <= 0) return;
    k_panel_scatter<<<ntiles, 256, 0, s>>>(tiles, panel_pool, fidx_pool, perm, f2user, nbf, bf, dr);
}

void launch_basis(const DevBasis &B, const TileDesc *tiles, int ntiles, int max_nruns, const TileGeo *geo, const double *rsx, const double *rsy,
                  const double *rsz, double *panel_pool, int *fidx_pool, TileAtom *atab_pool, cudaStream_t s) {
    if (ntiles <= 0) return;
    size_t smem = ((size_t)B.natoms + (size_t)4 * (max_nruns > 0 ? max_nruns : 1)) * sizeof(int);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_basis<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
    // one CTA per tile evaluates ~nact functions per thread one after the other: with fewer tiles than SMs (a plane of an integral)
    // the atoms of a tile are dealt to G CTAs instead, two CTAs' worth of work per SM (36x36 plane: 424 us with 15 CTAs)
    int nsm = 148, dev = 0;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int G = ntiles >= nsm ? 1 : std::min(16, (2 * nsm + ntiles - 1) / ntiles);
    k_basis<128><<<dim3((unsigned)ntiles, (unsigned)G), 128, smem, s>>>(B, tiles, geo, rsx, rsy, rsz, panel_pool, fidx_pool, atab_pool);
}

}  // namespace gb
