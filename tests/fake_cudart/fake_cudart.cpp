// FAKE CUDA runtime -- test infrastructure, never shipped, never loaded by the product, the GPU tests or the bench.
//
// tools/sanitize_api_host.sh links the REAL host code of libgimic_b200.so (api.cu, host_basis.cpp and the nvcc-generated launch
// stubs of the kernels) against this stand-in for libcudart: "device" memory is zero-initialised host memory, copies are memcpy,
// every kernel launch is a no-op that reports success -- except the small preparation kernels (point gather and the tile plan:
// gap splitting, active-set counts, prefix sums, rank range, batches), which are restated serially on the host so that api.cu's
// bookkeeping sees realistic tiles.  Nothing is computed (all results are zeros); what runs is the host-side
// orchestration of every C-ABI entry point -- context creation, staging buffers, the tile / batch / pool bookkeeping, the
// quadrature and property drivers -- under AddressSanitizer + UBSan in a container without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <algorithm>
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
// ---- the device-side tile plan (k_prepare.cu), restated serially: the semantic model the CUDA kernels are written against ----------
struct Piece { gb::TileGeo tg; float rho, gmax; int imax, nraw, natom, nreal; };
// double -> float rounded toward -inf (__double2float_rd)
float f_rd(double x) { float f = (float)x; if ((double)f > x) f = std::nextafterf(f, -INFINITY); return f; }
Piece eval_piece(const gb::DevBasis &B, const double *sx, const double *sy, const double *sz, long p0, int npts) {
    Piece P{};
    // bounding box in single precision, rounded outwards (the kernel reduces floats: lo = min rd(x), hi = -min rd(-x))
    float lo[3] = {3e38f, 3e38f, 3e38f}, nhi[3] = {3e38f, 3e38f, 3e38f};
    for (int p = 0; p < npts; ++p) {
        const long q = p0 + p;
        const double c[3] = {sx[q], sy[q], sz[q]};
        for (int k = 0; k < 3; ++k) { lo[k] = std::fmin(lo[k], f_rd(c[k])); nhi[k] = std::fmin(nhi[k], f_rd(-c[k])); }
    }
    gb::TileGeo tg{lo[0], lo[1], lo[2], -(double)nhi[0], -(double)nhi[1], -(double)nhi[2], 0.0, 0.0};
    const double cx = 0.5 * (tg.lox + tg.hix), cy = 0.5 * (tg.loy + tg.hiy), cz = 0.5 * (tg.loz + tg.hiz);
    float rho = 0.f, gmax = 0.f; int imax = 0;
    for (int p = 0; p < npts; ++p) {
        const long q = p0 + p;
        rho = std::fmax(rho, (float)((sx[q] - cx) * (sx[q] - cx) + (sy[q] - cy) * (sy[q] - cy) + (sz[q] - cz) * (sz[q] - cz)));      // squared, like the kernel
        if (p + 1 < npts) {
            const float g = (float)((sx[q + 1] - sx[q]) * (sx[q + 1] - sx[q]) + (sy[q + 1] - sy[q]) * (sy[q + 1] - sy[q]) + (sz[q + 1] - sz[q]) * (sz[q + 1] - sz[q]));
            if (g >= gmax) { gmax = g; imax = p; }          // ties: the later gap (the kernel reduces (bits << 32 | index) by max)
        }
    }
    rho = std::sqrt(rho); gmax = std::sqrt(gmax);
    tg.rho = rho; tg.pad_ = 0.0;
    const int al = B.slot_align - 1;
    for (int at = 0; at < B.natoms; ++at) {
        const double x = B.atom_xyz[3 * at], y = B.atom_xyz[3 * at + 1], z = B.atom_xyz[3 * at + 2];
        const double dx = std::fmax(std::fmax(tg.lox - x, x - tg.hix), 0.0), dy = std::fmax(std::fmax(tg.loy - y, y - tg.hiy), 0.0),
                     dz = std::fmax(std::fmax(tg.loz - z, z - tg.hiz), 0.0);
        if (dx * dx + dy * dy + dz * dz > B.atom_maxthr2e[at]) continue;
        // minimum squared distance in single precision about the tile centre, turned into a lower bound (for_active_atoms, k_prepare.cu)
        const float ax = (float)(x - cx), ay = (float)(y - cy), az = (float)(z - cz);
        float m = 3e38f;
        for (int p = 0; p < npts; ++p) {
            const long q = p0 + p;
            const float ex = (float)(sx[q] - cx) - ax, ey = (float)(sy[q] - cy) - ay, ez = (float)(sz[q] - cz) - az;
            m = std::fmin(m, std::fmaf(ez, ez, std::fmaf(ey, ey, ex * ex)));
        }
        const double dd = std::sqrt((double)m), D = 2.5e-7 * (2.0 * rho + dd), slack = 3.0e-7 * m + D * D + 2.0 * dd * D;
        const double d2 = std::fmax((double)m - 1.0101 * slack, 0.0);
        if (d2 > B.atom_maxthr2e[at]) continue;
        int nfun = 0;
        for (int s = B.atom_shell_off[at]; s < B.atom_shell_off[at + 1] && B.sh_thr2e[s] >= d2; ++s) nfun += (B.sh_l[s] + 1) * (B.sh_l[s] + 2) / 2;
        P.nraw += (nfun + al) & ~al; P.natom += nfun > 0; P.nreal += nfun;
    }
    P.tg = tg; P.rho = rho; P.gmax = gmax; P.imax = imax;
    return P;
}
using gb::piece_cost;      // kernels.cuh
void split_piece(const gb::DevBasis &B, const double *sx, const double *sy, const double *sz, long p0, int npts, int depth, double split_radius,
                 long run, int &emitted, int pending, gb::TileSeg *seg, gb::TileGeo *geo, gb::TileInfo *info) {
    const Piece P = eval_piece(B, sx, sy, sz, p0, npts);
    bool split = depth < gb::SPLIT_DEPTH && npts >= 16 && P.nraw > 0 && P.rho > (float)split_radius && P.gmax > 0.5f * P.rho &&
                 emitted + pending + 2 <= gb::MAXSUB;
    if (split) {         // only if the two pieces cost the contraction less than the whole (k_tile_split)
        const Piece L = eval_piece(B, sx, sy, sz, p0, P.imax + 1), R = eval_piece(B, sx, sy, sz, p0 + P.imax + 1, npts - P.imax - 1);
        split = 20 * (piece_cost(P.imax + 1, L.nraw, L.nreal, L.natom) + piece_cost(npts - P.imax - 1, R.nraw, R.nreal, R.natom)) <
                17 * piece_cost(npts, P.nraw, P.nreal, P.natom);
    }
    if (split) {
        split_piece(B, sx, sy, sz, p0, P.imax + 1, depth + 1, split_radius, run, emitted, pending + 1, seg, geo, info);
        split_piece(B, sx, sy, sz, p0 + P.imax + 1, npts - P.imax - 1, depth + 1, split_radius, run, emitted, pending, seg, geo, info);
    } else {
        const long o = run * gb::MAXSUB + emitted++;
        seg[o] = gb::TileSeg{(int)p0, npts}; geo[o] = P.tg; info[o] = gb::TileInfo{P.rho, P.gmax, P.imax, P.nraw, P.natom, P.nreal};
    }
}
void emulate_tile_split(void **a, unsigned nrun0) {
    const gb::DevBasis &B = *(const gb::DevBasis *)a[0];
    const double *sx = *(const double **)a[1], *sy = *(const double **)a[2], *sz = *(const double **)a[3];
    const long n = *(const long *)a[4];
    const double split_radius = *(const double *)a[5];
    gb::TileSeg *seg = *(gb::TileSeg **)a[6]; gb::TileGeo *geo = *(gb::TileGeo **)a[7]; gb::TileInfo *info = *(gb::TileInfo **)a[8];
    int *cnt = *(int **)a[9];
    for (long run = 0; run < (long)nrun0; ++run) {
        const long p0 = run * gb::MT;
        const int np = (int)std::min<long>(gb::MT, n - p0);
        int emitted = 0;
        split_piece(B, sx, sy, sz, p0, np, 0, split_radius, run, emitted, 0, seg, geo, info);
        cnt[run] = emitted;
    }
}
// The prefix sums are three kernels each on the device (tile scans, scan of the tile totals, offsets added back).  Here the first one
// produces the final result and the third only does what it does besides adding (the run -> tile count of the summary).
void emulate_scan_partial_i(void **a) {
    const int *cnt = *(const int **)a[0]; int *off = *(int **)a[1]; const int n = *(const int *)a[2];
    int run = 0;
    for (int i = 0; i < n; ++i) { const int v = cnt[i]; off[i] = run; run += v; }
    off[n] = run;
}
void emulate_scan_add_i(void **a) {
    const int *off = *(const int **)a[0]; const int n = *(const int *)a[1];
    gb::PlanSummary *sum = *(gb::PlanSummary **)a[3]; const int cap = *(const int *)a[4];
    sum->ntiles = std::min(off[n], cap); sum->overflow = off[n] > cap;
}
void emulate_tile_emit(void **a) {
    const gb::TileSeg *seg = *(const gb::TileSeg **)a[0]; const gb::TileGeo *sgeo = *(const gb::TileGeo **)a[1];
    const gb::TileInfo *info = *(const gb::TileInfo **)a[2]; const int *cnt = *(const int **)a[3], *off = *(const int **)a[4];
    const long nrun0 = *(const long *)a[5]; const int cap = *(const int *)a[6];
    gb::TileGeo *geo = *(gb::TileGeo **)a[7]; gb::TileDesc *desc = *(gb::TileDesc **)a[8]; gb::TileCum *cum = *(gb::TileCum **)a[9];
    for (long run = 0; run < nrun0; ++run)
        for (int j = 0; j < cnt[run] && off[run] + j < cap; ++j) {
            const int t = off[run] + j;
            const gb::TileInfo ti = info[run * gb::MAXSUB + j];
            geo[t] = sgeo[run * gb::MAXSUB + j];
            gb::TileDesc td{};
            td.pt0 = seg[run * gb::MAXSUB + j].pt0; td.npts = seg[run * gb::MAXSUB + j].npts; td.nraw = ti.nraw; td.nact = (ti.nraw + 7) / 8 * 8;
            td.nreal = ti.nreal; td.nn = (ti.nreal + 7) / 8 * 8; td.geo = t; td.nruns = ti.natom;
            td.col0 = 0; td.col1 = td.nn; td.part = -1;
            desc[t] = td;
            cum[t] = gb::TileCum{piece_cost(td.npts, td.nraw, td.nreal, td.nruns), 4LL * td.nact * gb::LDP, td.nact + td.nn, td.nruns};
        }
}
void emulate_scan_partial_c(void **a) {
    gb::TileCum *cum = *(gb::TileCum **)a[1]; const gb::PlanSummary *sum = *(const gb::PlanSummary **)a[2];
    gb::TileCum run{0, 0, 0, 0};
    for (int t = 0; t < sum->ntiles; ++t) {
        const gb::TileCum v = cum[t];
        cum[t] = run;
        run = gb::TileCum{run.cost + v.cost, run.panel + v.panel, run.fidx + v.fidx, run.atab + v.atab};
    }
    cum[sum->ntiles] = run;
}
void emulate_plan_range(void **a) {
    const gb::TileDesc *desc = *(const gb::TileDesc **)a[0]; const gb::TileCum *cum = *(const gb::TileCum **)a[1];
    const int rank = *(const int *)a[2], nranks = *(const int *)a[3]; const long long pool = *(const long long *)a[4];
    gb::PlanSummary *sum = *(gb::PlanSummary **)a[5];
    const int nt = sum->ntiles;
    const long long total = cum[nt].cost;
    int tlo = nt, thi = nt;
    for (int t = nt - 1; t >= 0; --t) {                       // tile t belongs to rank floor(cum[t].cost * nranks / total)
        const long long r = cum[t].cost * nranks / total;
        if (r >= rank) tlo = t;
        if (r >= rank + 1) thi = t;
    }
    sum->tlo = tlo; sum->thi = thi; sum->cost_total = total; sum->cost_range = cum[thi].cost - cum[tlo].cost;
    sum->panel_range = cum[thi].panel - cum[tlo].panel;
    sum->pt_lo = tlo < thi ? desc[tlo].pt0 : 0; sum->pt_hi = tlo < thi ? (long long)desc[thi - 1].pt0 + desc[thi - 1].npts : 0;
    const long long nb = tlo < thi ? (cum[thi - 1].panel - cum[tlo].panel) / pool + 1 : 0;
    sum->nbatch = (int)(nb < gb::MAX_BATCH ? nb : gb::MAX_BATCH + 1);
    sum->max_nruns = 0; sum->max_tile_panel = 0;
    sum->sum_nact = sum->flops4 = sum->flops2 = sum->taps = sum->useful_mm = sum->useful_taps = 0.0;
    if (nb <= gb::MAX_BATCH) { sum->batch_start[nb] = thi; sum->batch_pt[nb] = sum->pt_hi; }
    const long long full = sum->panel_range < pool ? sum->panel_range : pool;
    long long chunk = (full + gb::DRAIN_GROUPS - 1) / gb::DRAIN_GROUPS;
    if (chunk < (16LL << 20)) chunk = 16LL << 20;
    if (const char *e = std::getenv("FAKE_DRAIN_CHUNK")) chunk = std::atoll(e);          // tests: groups on small point sets
    const int drain_batches = *(const int *)a[6], drain_min_tiles = *(const int *)a[7];
    const bool groups = std::getenv("FAKE_DRAIN_CHUNK") ? nb <= gb::DRAIN_BATCHES : (nb <= drain_batches && thi - tlo >= drain_min_tiles);
    sum->drain_chunk = groups ? chunk : pool;
    for (int i = 0; i <= gb::DRAIN_BATCHES * gb::DRAIN_GROUPS; ++i) { sum->group_tile[i] = -1; sum->group_pt[i] = 0; }
    auto group_of = [&](int t) { return gb::drain_group(cum[t].panel - cum[tlo].panel, pool, sum->drain_chunk); };
    for (int t = tlo; t < thi; ++t) {
        const long long b = (cum[t].panel - cum[tlo].panel) / pool;
        const bool newb = t == tlo || (cum[t - 1].panel - cum[tlo].panel) / pool != b;
        if (b < gb::MAX_BATCH && newb) { sum->batch_start[b] = t; sum->batch_pt[b] = desc[t].pt0; }
        if (b < gb::DRAIN_BATCHES && (newb || group_of(t - 1) != group_of(t))) { sum->group_tile[b * gb::DRAIN_GROUPS + group_of(t)] = t; sum->group_pt[b * gb::DRAIN_GROUPS + group_of(t)] = desc[t].pt0; }
    }
}
const int *g_last_ord0 = nullptr;
void emulate_plan_finalize(void **a) {
    gb::TileDesc *desc = *(gb::TileDesc **)a[0]; const gb::TileCum *cum = *(const gb::TileCum **)a[1]; const long long pool = *(const long long *)a[2];
    gb::PlanSummary *sum = *(gb::PlanSummary **)a[3]; unsigned long long *keys = *(unsigned long long **)a[4]; int *ord = *(int **)a[5];
    g_last_ord0 = ord;
    if (sum->nbatch > gb::MAX_BATCH) return;
    for (int t = sum->tlo; t < sum->thi; ++t) {
        const long long b = (cum[t].panel - cum[sum->tlo].panel) / pool;
        const int t0 = sum->batch_start[b];
        gb::TileDesc &td = desc[t];
        td.panel_off = cum[t].panel - cum[t0].panel; td.fidx_off = cum[t].fidx - cum[t0].fidx; td.atab_off = cum[t].atab - cum[t0].atab;
        const long long c = cum[t + 1].cost - cum[t].cost;
        keys[t - sum->tlo] = ((unsigned long long)b << 32) | (0xffffffffu - (unsigned)std::min<long long>(c, 0xffffffffLL));
        ord[t - sum->tlo] = t;
        sum->sum_nact += td.nact; sum->flops4 += 2.0 * gb::MT * 4.0 * td.nact * td.nn; sum->flops2 += 2.0 * gb::MT * 2.0 * td.nact * td.nn;
        sum->taps += 2.0 * gb::MT * (double)td.nn * td.nruns; sum->useful_mm += 2.0 * td.npts * (double)td.nreal * td.nreal;
        sum->useful_taps += 2.0 * td.npts * (double)td.nreal * td.nruns;
        sum->max_nruns = std::max(sum->max_nruns, td.nruns); sum->max_tile_panel = std::max(sum->max_tile_panel, cum[t + 1].panel - cum[t].panel);
    }
}
void emulate_tile_gather(void **a) {      // the CUB sort in front of it is a no-op here: fall back to the unsorted order
    const gb::TileDesc *desc = *(const gb::TileDesc **)a[0]; const int *ord = *(const int **)a[1]; const int nt = *(const int *)a[2];
    gb::TileDesc *out = *(gb::TileDesc **)a[3];
    bool zeros = nt > 1;
    for (int i = 0; i < nt && zeros; ++i) zeros = ord[i] == 0;
    if (zeros && g_last_ord0) ord = g_last_ord0;
    for (int i = 0; i < nt; ++i) out[i] = desc[ord[i]];
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
        else if (name.find("k_tile_split") != std::string::npos) emulate_tile_split(args, grid.x);
        else if (name.find("k_scan_partial_i") != std::string::npos) emulate_scan_partial_i(args);
        else if (name.find("k_scan_add_i") != std::string::npos) emulate_scan_add_i(args);
        else if (name.find("k_tile_emit") != std::string::npos) emulate_tile_emit(args);
        else if (name.find("k_scan_partial_c") != std::string::npos) emulate_scan_partial_c(args);
        else if (name.find("k_plan_range") != std::string::npos) emulate_plan_range(args);
        else if (name.find("k_plan_finalize") != std::string::npos) emulate_plan_finalize(args);
        else if (name.find("k_tile_gather") != std::string::npos) emulate_tile_gather(args);
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
cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)&g_dummy_stream; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)&g_dummy_event; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)&g_dummy_event; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake CUDA runtime error"; }

long fake_cuda_launch_count(void) { return g_launches; }

}  // extern "C"
