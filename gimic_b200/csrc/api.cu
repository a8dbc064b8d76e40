// Context management, batching driver and the C ABI of libgimic_b200.so (see include/gimic_b200.h).
//
// There is deliberately no CPU fallback: every compute entry point fails with GIMIC_B200_ECUDA when
// no CUDA device / kernel image is available.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/gimic_b200.h"
#include "host_basis.hpp"
#include "kernels.cuh"

namespace gb { void upload_component_tables(const signed char *host_tab); }

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return fail(GIMIC_B200_ECUDA, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
    } while (0)

struct Buf {
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; return -1; } want = bytes; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct gimic_b200_ctx {
    int device = 0, nsm = 148;
    cudaStream_t stream = nullptr;
    gimic_b200_opts opts{};
    gb::HostBasis hb;
    gb::DevBasis db{};
    std::vector<void *> owned;
    int *d_f2user = nullptr;
    double *d_dens[2] = {nullptr, nullptr};   // dens_t%da / %db in the XDENS layout
    double *d_op[4] = {nullptr, nullptr, nullptr, nullptr};   // contraction operands per spin case
    double *d_opj[4] = {nullptr, nullptr, nullptr, nullptr};  // J = T.B path: one pair-plane (D, P.B) per spin case, for the field in opj_B
    double opj_B[4][3] = {};
    int nq = gb::NQ, ldb = 0; long long plane_stride = 0;
    double bbox_lo[3] = {0, 0, 0}; double inv_cell = 1.0;
    size_t pool_max_bytes = (size_t)8 << 30;
    // workspaces
    Buf keys0, keys1, vals0, vals1, sorttmp, rs, geo, nraw, segs, tiles, panel, fidx, atab, misc, r_in, tens_tmp, f_tmp, shift, jv6, gridbuf, quad;
    gb::TileInfo *h_info = nullptr; size_t h_info_cap = 0;
    gb::TileSeg *h_segs = nullptr;
    double split_radius = 2.5;   // bohr: tiles wider than this are cut at their largest consecutive gap if that shrinks them
    gb::TileDesc *h_tiles = nullptr;
    bool profiling = false;
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_call[2] = {};
    std::vector<cudaEvent_t> evpool;   // per-batch (basis, contract) stamps, resolved at the end of a call
    gimic_b200_stats stats{};
    std::string mol_path, xdens_path;   // for the legacy set_uhf-after-init path

    ~gimic_b200_ctx() {
        cudaSetDevice(device);
        for (void *p : owned) cudaFree(p);
        for (int i = 0; i < 2; ++i) if (d_dens[i]) cudaFree(d_dens[i]);
        for (int i = 0; i < 4; ++i) if (d_op[i]) cudaFree(d_op[i]);
        for (int i = 0; i < 4; ++i) if (d_opj[i]) cudaFree(d_opj[i]);
        for (Buf *b : {&keys0, &keys1, &vals0, &vals1, &sorttmp, &rs, &geo, &nraw, &segs, &tiles, &panel, &fidx, &atab, &misc, &r_in, &tens_tmp, &f_tmp, &shift, &jv6, &gridbuf, &quad}) b->release();
        if (h_info) cudaFreeHost(h_info);
        if (h_segs) cudaFreeHost(h_segs);
        if (h_tiles) cudaFreeHost(h_tiles);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        for (auto &e : evpool) cudaEventDestroy(e);
        for (auto &e : ev_call) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

template <typename T>
int upload(gimic_b200_ctx *c, const std::vector<T> &v, const T **out) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CUDA_TRY(cudaMalloc(&p, bytes));
    c->owned.push_back(p);
    if (!v.empty()) CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T *>(p);
    return 0;
}

// Builds the device tables: shells re-ordered inside each atom by descending screening radius so
// that a tile's active set is a per-atom prefix; functions keep their atom but follow the shells.
int build_device_basis(gimic_b200_ctx *c) {
    const gb::HostBasis &hb = c->hb;
    std::vector<int> order;   // internal shell order -> reference shell index
    std::vector<int> atom_shell_off(1, 0), atom_func_off(1, 0);
    for (int a = 0; a < hb.natoms; ++a) {
        std::vector<int> idx;
        for (int s = hb.atom_shell_off[a]; s < hb.atom_shell_off[a + 1]; ++s) idx.push_back(s);
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return hb.shells[x].thr > hb.shells[y].thr; });
        order.insert(order.end(), idx.begin(), idx.end());
        atom_shell_off.push_back((int)order.size());
        atom_func_off.push_back(hb.atom_func_off[a + 1]);
    }
    const int ns = (int)order.size();
    std::vector<int> sh_l(ns), sh_np(ns), sh_po(ns), sh_foff(ns), f2user(hb.nbf);
    std::vector<double> sh_thr(ns), maxthr(hb.natoms, 0.0), fR(3 * (size_t)hb.nbf);
    int f = 0;
    for (int i = 0; i < ns; ++i) {
        const gb::Shell &s = hb.shells[order[i]];
        sh_l[i] = s.l; sh_np[i] = s.nprim; sh_po[i] = s.prim_off; sh_thr[i] = s.thr; sh_foff[i] = f;
        maxthr[s.atom] = std::max(maxthr[s.atom], s.thr);
        for (int k = 0; k < s.ncomp; ++k, ++f) {
            f2user[f] = s.user_off + k;
            for (int d = 0; d < 3; ++d) fR[(size_t)d * hb.nbf + f] = hb.xyz[3 * s.atom + d];
        }
    }
    gb::DevBasis &d = c->db;
    d.natoms = hb.natoms; d.nbf = hb.nbf; d.nshell = ns; d.turbomole = hb.turbomole ? 1 : 0;
    d.slot_align = c->opts.giao ? gb::SLOT_ALIGN_GIAO : 1;
    if (int rc = upload(c, hb.xyz, &d.atom_xyz)) return rc;
    if (int rc = upload(c, maxthr, &d.atom_maxthr)) return rc;
    if (int rc = upload(c, atom_shell_off, &d.atom_shell_off)) return rc;
    if (int rc = upload(c, atom_func_off, &d.atom_func_off)) return rc;
    if (int rc = upload(c, sh_l, &d.sh_l)) return rc;
    if (int rc = upload(c, sh_np, &d.sh_nprim)) return rc;
    if (int rc = upload(c, sh_po, &d.sh_prim_off)) return rc;
    if (int rc = upload(c, sh_foff, &d.sh_foff)) return rc;
    if (int rc = upload(c, sh_thr, &d.sh_thr)) return rc;
    if (int rc = upload(c, hb.alpha, &d.alpha)) return rc;
    if (int rc = upload(c, hb.ncc, &d.ncc)) return rc;
    if (int rc = upload(c, fR, &d.fR)) return rc;
    const int *f2u = nullptr;
    if (int rc = upload(c, f2user, &f2u)) return rc;
    c->d_f2user = const_cast<int *>(f2u);

    signed char tab[2][6][21][3];
    std::memset(tab, 0, sizeof tab);
    for (int tm = 0; tm < 2; ++tm)
        for (int l = 0; l <= gb::MAX_L; ++l)
            for (int k = 0; k < (l + 1) * (l + 2) / 2; ++k) {
                int lmn[3]; gb::component_exponents(l, tm != 0, k, lmn);
                for (int x = 0; x < 3; ++x) tab[tm][l][k][x] = (signed char)lmn[x];
            }
    gb::upload_component_tables(&tab[0][0][0][0]);
    CUDA_TRY(cudaGetLastError());

    // Morton quantisation box: molecule +- 24 bohr (beyond every screening radius in practice; farther points clamp)
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int a = 0; a < hb.natoms; ++a)
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], hb.xyz[3 * a + k]); hi[k] = std::max(hi[k], hb.xyz[3 * a + k]); }
    double ext = 0;
    for (int k = 0; k < 3; ++k) { c->bbox_lo[k] = lo[k] - 24.0; ext = std::max(ext, hi[k] - lo[k] + 48.0); }
    c->inv_cell = 65536.0 / ext;
    return 0;
}

int init_device(gimic_b200_ctx *c) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(GIMIC_B200_ECUDA, "no CUDA device available: gimic-b200 has no CPU path");
    }
    if (c->opts.device >= 0) c->device = c->opts.device; else CUDA_TRY(cudaGetDevice(&c->device));
    CUDA_TRY(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
    c->nsm = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    for (auto &e : c->ev_call) CUDA_TRY(cudaEventCreate(&e));
    // panel pool: one batch (one k_basis + one k_jtensor launch) per ~pool of Phi/dPhi panels.  8 GB is the configuration of the
    // committed ncu captures and launch lists; GIMIC_B200_POOL_MB=24576 (one launch per 2M-point step) measured +1 %.
    c->pool_max_bytes = std::min<size_t>((size_t)8 << 30, std::max<size_t>((size_t)1 << 30, prop.totalGlobalMem / 8));
    if (const char *mb = std::getenv("GIMIC_B200_POOL_MB")) { long v = std::atol(mb); if (v > 0) c->pool_max_bytes = (size_t)v << 20; }
    return 0;
}

int get_operand(gimic_b200_ctx *c, int spincase, const double **op) {
    const bool uhf = c->opts.uhf != 0;
    if (spincase < 0 || spincase > 3) return fail(GIMIC_B200_EINVAL, "invalid spin case");
    if (!uhf) {
        if (spincase == GIMIC_B200_BETA) return fail(GIMIC_B200_ESPIN, "ctensor(): beta current requested, but not open-shell system!");
        if (spincase == GIMIC_B200_SPINDENS) return fail(GIMIC_B200_ESPIN, "ctensor(): spindens requested, but not open-shell system!");
        spincase = GIMIC_B200_ALPHA;
    }
    if (spincase < 0 || spincase > 3) return fail(GIMIC_B200_EINVAL, "invalid spin case");
    if (!c->d_op[spincase]) {
        const int nbf = c->hb.nbf;
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, (size_t)((c->nq + 1) / 2) * c->plane_stride * sizeof(double)));
        const double *A = c->d_dens[0], *Bm = nullptr; double sg = 0.0;
        if (spincase == GIMIC_B200_BETA) A = c->d_dens[1];
        if (spincase == GIMIC_B200_TOTAL) { Bm = c->d_dens[1]; sg = 1.0; }       // T_alpha + T_beta (jtensor.F90:86-88), by linearity in D, P
        if (spincase == GIMIC_B200_SPINDENS) { Bm = c->d_dens[1]; sg = -1.0; }   // T_alpha - T_beta (jtensor.F90:97-99)
        gb::launch_build_operand((double *)p, nbf, c->ldb, c->plane_stride, A, Bm, sg, c->d_f2user, c->stream);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->d_op[spincase] = (double *)p;
    }
    *op = c->d_op[spincase];
    return 0;
}

// Operand of the J = T.B path: the tensor is only ever contracted with B, so P_x, P_y, P_z enter as the single matrix
// sum_b B_b P_b (jfield.f90:167-184 applied before instead of after the contraction).  Rebuilt when B changes.
int get_operand_j(gimic_b200_ctx *c, int spincase, const double *B3, const double **op) {
    const bool uhf = c->opts.uhf != 0;
    if (spincase < 0 || spincase > 3) return fail(GIMIC_B200_EINVAL, "invalid spin case");
    if (!uhf) {
        if (spincase == GIMIC_B200_BETA) return fail(GIMIC_B200_ESPIN, "ctensor(): beta current requested, but not open-shell system!");
        if (spincase == GIMIC_B200_SPINDENS) return fail(GIMIC_B200_ESPIN, "ctensor(): spindens requested, but not open-shell system!");
        spincase = GIMIC_B200_ALPHA;
    }
    if (spincase < 0 || spincase > 3) return fail(GIMIC_B200_EINVAL, "invalid spin case");
    const bool same = c->d_opj[spincase] && c->opj_B[spincase][0] == B3[0] && c->opj_B[spincase][1] == B3[1] && c->opj_B[spincase][2] == B3[2];
    if (!same) {
        const int nbf = c->hb.nbf;
        if (!c->d_opj[spincase]) CUDA_TRY(cudaMalloc((void **)&c->d_opj[spincase], (size_t)c->plane_stride * sizeof(double)));
        const double *A = c->d_dens[0], *Bm = nullptr; double sg = 0.0;
        if (spincase == GIMIC_B200_BETA) A = c->d_dens[1];
        if (spincase == GIMIC_B200_TOTAL) { Bm = c->d_dens[1]; sg = 1.0; }
        if (spincase == GIMIC_B200_SPINDENS) { Bm = c->d_dens[1]; sg = -1.0; }
        gb::launch_build_operand_j(c->d_opj[spincase], nbf, c->ldb, A, Bm, sg, c->d_f2user, B3, c->stream);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 3; ++k) c->opj_B[spincase][k] = B3[k];
    }
    *op = c->d_opj[spincase];
    return 0;
}

int finish_create(gimic_b200_ctx *c, const double *dens_a, const double *dens_b, bool dens_on_device) {
    gb::finalize_basis(c->hb, c->opts.screening != 0, c->opts.screening_thrs);
    if (int rc = build_device_basis(c)) return rc;
    const size_t nn = (size_t)c->hb.nbf * c->hb.nbf;
    std::vector<double> cart[2];
    if (c->hb.spherical) {
        // spherical=on: the kernels stay cartesian; the SAO densities are folded with the projection once,
        // D_cart = po^T D_sao po (same bilinear form as projecting Phi, dPhi at every point, bfeval.f90:116-118,330-333)
        const size_t ss = (size_t)c->hb.nbf_sph * c->hb.nbf_sph;
        std::vector<double> stage;
        for (int sp = 0; sp < (c->opts.uhf ? 2 : 1); ++sp) {
            const double *src = sp ? dens_b : dens_a;
            if (!src) return fail(GIMIC_B200_EINVAL, sp ? "open-shell context needs beta densities" : "densities missing");
            if (dens_on_device) { stage.resize(4 * ss); CUDA_TRY(cudaMemcpy(stage.data(), src, 4 * ss * sizeof(double), cudaMemcpyDeviceToHost)); src = stage.data(); }
            cart[sp].resize(4 * nn);
            for (int m = 0; m < 4; ++m) gb::density_sph_to_cart(c->hb, src + m * ss, cart[sp].data() + m * nn);
        }
        dens_a = cart[0].data(); dens_b = c->opts.uhf ? cart[1].data() : nullptr; dens_on_device = false;
    }
    c->nq = gb::NQ;
    c->ldb = c->hb.nbf;
    c->plane_stride = 2LL * c->hb.nbf * c->ldb;          // doubles per pair-plane [nbf][ldb][2]
    for (int sp = 0; sp < (c->opts.uhf ? 2 : 1); ++sp) {
        const double *src = sp ? dens_b : dens_a;
        if (!src) return fail(GIMIC_B200_EINVAL, sp ? "open-shell context needs beta densities" : "densities missing");
        CUDA_TRY(cudaMalloc((void **)&c->d_dens[sp], 4 * nn * sizeof(double)));
        CUDA_TRY(cudaMemcpy(c->d_dens[sp], src, 4 * nn * sizeof(double), dens_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    }
    return 0;
}

// ---- the batched tensor pipeline ------------------------------------------------------------------
// d_r: device, 3 x n (AoS).  d_tens: device 9 x n.  d_edens: device n or null.
// jB3 / d_jvec non-null: the J = T.B path (3 x n output, operands (D, P.B)); d_tens is then unused.
int run_tensors(gimic_b200_ctx *c, long n, const double *d_r, int spincase, double *d_tens, double *d_edens,
                const double *jB3 = nullptr, double *d_jvec = nullptr) {
    using namespace gb;
    if (n <= 0) return 0;
    if (n > 2000000000L) return fail(GIMIC_B200_EINVAL, "more than 2e9 points in one call");
    const double *op = nullptr;
    const bool jpath = d_jvec != nullptr;
    if (int rc = jpath ? get_operand_j(c, spincase, jB3, &op) : get_operand(c, spincase, &op)) return rc;
    cudaStream_t st = c->stream;
    int ntiles = (int)((n + MT - 1) / MT);
    const bool prof = c->profiling;

    if (c->keys0.ensure(n * 8) || c->keys1.ensure(n * 8) || c->vals0.ensure(n * 4) || c->vals1.ensure(n * 4) ||
        c->rs.ensure((size_t)3 * n * 8) || c->misc.ensure(256))
        return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed");
    size_t tb = sort_temp_bytes(n);
    if (c->sorttmp.ensure(tb)) return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (sort)");
    double *rsx = c->rs.as<double>(), *rsy = rsx + n, *rsz = rsy + n;
    int *perm = c->vals1.as<int>();

    if (prof) cudaEventRecord(c->ev[0], st);
    launch_morton_keys(d_r, n, c->bbox_lo, c->inv_cell, c->keys0.as<uint64_t>(), c->vals0.as<int>(), st);
    launch_sort_pairs(c->sorttmp.p, tb, c->keys0.as<uint64_t>(), c->keys1.as<uint64_t>(), c->vals0.as<int>(), perm, n, st);
    launch_gather_points(d_r, perm, n, rsx, rsy, rsz, st);
    if (prof) cudaEventRecord(c->ev[1], st);
    c->stats.launches += 4;

    // Tiles: runs of MT consecutive points along the Hilbert curve.  A run that straddles a re-entry of the
    // curve (thin / planar point sets, cluster boundaries) is cut at its largest consecutive gap; a few rounds.
    std::vector<TileSeg> segs(ntiles);
    for (int t = 0; t < ntiles; ++t) segs[t] = TileSeg{t * MT, (int)std::min<long>(MT, n - (long)t * MT)};
    for (int round = 0; round < 8; ++round) {
        ntiles = (int)segs.size();
        if ((size_t)ntiles > c->h_info_cap) {
            if (c->h_info) cudaFreeHost(c->h_info);
            if (c->h_segs) cudaFreeHost(c->h_segs);
            if (c->h_tiles) cudaFreeHost(c->h_tiles);
            c->h_info = nullptr; c->h_segs = nullptr; c->h_tiles = nullptr; c->h_info_cap = 0;   // a failed allocation below must not leave stale pointers
            const size_t cap = (size_t)ntiles + ntiles / 4 + 64;
            CUDA_TRY(cudaMallocHost((void **)&c->h_info, cap * sizeof(TileInfo)));
            CUDA_TRY(cudaMallocHost((void **)&c->h_segs, cap * sizeof(TileSeg)));
            CUDA_TRY(cudaMallocHost((void **)&c->h_tiles, cap * sizeof(TileDesc)));
            c->h_info_cap = cap;
        }
        if (c->geo.ensure((size_t)ntiles * sizeof(TileGeo)) || c->nraw.ensure((size_t)ntiles * sizeof(TileInfo)) ||
            c->segs.ensure((size_t)ntiles * sizeof(TileSeg)) || c->tiles.ensure((size_t)ntiles * sizeof(TileDesc)))
            return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (tiles)");
        std::copy(segs.begin(), segs.end(), c->h_segs);
        CUDA_TRY(cudaMemcpyAsync(c->segs.p, c->h_segs, (size_t)ntiles * sizeof(TileSeg), cudaMemcpyHostToDevice, st));
        launch_tile_count(c->db, rsx, rsy, rsz, c->segs.as<TileSeg>(), ntiles, c->geo.as<TileGeo>(), c->nraw.as<TileInfo>(), st);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(c->h_info, c->nraw.p, (size_t)ntiles * sizeof(TileInfo), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        c->stats.launches += 1;
        std::vector<TileSeg> next;
        next.reserve(segs.size() + 16);
        bool any = false;
        for (int t = 0; t < ntiles; ++t) {
            const TileInfo &ti = c->h_info[t];
            const TileSeg &sg = segs[t];
            if (round < 7 && sg.npts >= 16 && ti.nraw > 0 && ti.rho > c->split_radius && ti.gmax > 0.5f * ti.rho) {
                next.push_back(TileSeg{sg.pt0, ti.imax + 1});
                next.push_back(TileSeg{sg.pt0 + ti.imax + 1, sg.npts - ti.imax - 1});
                any = true;
            } else next.push_back(sg);
        }
        if (!any) break;
        segs.swap(next);
    }
    if (prof) cudaEventRecord(c->ev[2], st);

    // host: tile descriptors, split into batches that fit the panel pool
    size_t max_tile = 0, total = 0;
    for (int t = 0; t < ntiles; ++t) {
        int nact = (c->h_info[t].nraw + 7) / 8 * 8;
        size_t d = (size_t)4 * nact * LDP;
        max_tile = std::max(max_tile, d); total += d;
    }
    size_t pool_doubles = std::max(max_tile, std::min(total, c->pool_max_bytes / 8));
    std::vector<int> batch_start(1, 0);
    size_t off = 0, foff = 0, fidx_max = 0, aoff = 0, atab_max = 0;
    double sum_nact = 0, flops = 0, useful = 0;
    const bool giao = c->opts.giao != 0;
    for (int t = 0; t < ntiles; ++t) {
        TileDesc &td = c->h_tiles[t];
        td.pt0 = segs[t].pt0; td.npts = segs[t].npts; td.geo = t; td.nruns = c->h_info[t].natom;
        td.nraw = c->h_info[t].nraw; td.nact = (td.nraw + 7) / 8 * 8;
        td.nreal = c->h_info[t].nreal; td.nn = (td.nreal + 7) / 8 * 8;
        if (td.nact > 65536) return fail(GIMIC_B200_EINVAL, "more than 65536 active basis-function slots in one tile of points (k_jtensor's K-step mask)");
        size_t d = (size_t)4 * td.nact * LDP;
        if (off + d > pool_doubles) { batch_start.push_back(t); fidx_max = std::max(fidx_max, foff); atab_max = std::max(atab_max, aoff); off = 0; foff = 0; aoff = 0; }
        td.panel_off = (long long)off; td.fidx_off = (long long)foff; td.atab_off = (long long)aoff;
        off += d; foff += td.nact + td.nn; aoff += td.nruns;
        sum_nact += td.nact;
        flops += 2.0 * MT * (jpath ? 2 : c->nq) * (double)td.nact * td.nn;   // DMMA: K runs over the nact slots, N over the nn columns (multiples of 8)
        if (giao) flops += 2.0 * MT * (jpath ? 1.0 : 3.0) * (double)td.nn * td.nruns;   // GIAO taps: 3 (J path: 1) DFMA per accumulator element per active atom
        useful += 2.0 * td.npts * (jpath ? 2 : c->nq) * (double)td.nreal * td.nreal;
        if (giao) useful += 2.0 * td.npts * (jpath ? 1.0 : 3.0) * (double)td.nreal * td.nruns;
    }
    fidx_max = std::max(fidx_max, foff); atab_max = std::max(atab_max, aoff);
    batch_start.push_back(ntiles);
    // inside a batch the contraction kernel pulls tiles from an atomic counter: longest first (cost ~ nact^2)
    static const int sched = [] { const char *e = std::getenv("GIMIC_B200_SCHED"); return e ? std::atoi(e) : 0; }();
    for (size_t b = 0; b + 1 < batch_start.size(); ++b) {
        TileDesc *t0 = c->h_tiles + batch_start[b], *t1 = c->h_tiles + batch_start[b + 1];
        if (sched == 0) {          // longest first
            std::stable_sort(t0, t1, [](const TileDesc &x, const TileDesc &y) { return x.nact > y.nact; });
        } else if (sched == 2) {   // Hilbert order, except that the heaviest 1/8 of the tiles go first (tail protection)
            std::vector<int> v; for (TileDesc *t = t0; t < t1; ++t) v.push_back(t->nact);
            if (!v.empty()) {
                std::nth_element(v.begin(), v.begin() + v.size() / 8, v.end(), std::greater<int>());
                const int cut = v[v.size() / 8];
                std::stable_partition(t0, t1, [cut](const TileDesc &x) { return x.nact > cut; });
            }
        }                          // sched == 1: plain Hilbert order
    }
    if (c->panel.ensure(std::max<size_t>(pool_doubles, 2) * 8) || c->fidx.ensure(std::max<size_t>(fidx_max, 1) * 4) ||
        c->atab.ensure(std::max<size_t>(atab_max, 1) * sizeof(TileAtom)))
        return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (panel pool)");
    CUDA_TRY(cudaMemcpyAsync(c->tiles.p, c->h_tiles, (size_t)ntiles * sizeof(TileDesc), cudaMemcpyHostToDevice, st));

    const size_t nbatch = batch_start.size() - 1;
    if (prof) while (c->evpool.size() < 3 * nbatch) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); c->evpool.push_back(e); }
    for (size_t b = 0; b < nbatch; ++b) {
        const int t0 = batch_start[b], nb = batch_start[b + 1] - t0;
        if (nb <= 0) continue;
        if (prof) cudaEventRecord(c->evpool[3 * b], st);
        launch_basis(c->db, c->tiles.as<TileDesc>() + t0, nb, c->geo.as<TileGeo>(), rsx, rsy, rsz, c->panel.as<double>(), c->fidx.as<int>(),
                     giao ? c->atab.as<TileAtom>() : nullptr, st);
        if (prof) cudaEventRecord(c->evpool[3 * b + 1], st);
        CUDA_TRY(cudaMemsetAsync(c->misc.p, 0, 4, st));
        JtensorArgs a;
        a.tiles = c->tiles.as<TileDesc>() + t0; a.ntiles = nb; a.counter = c->misc.as<int>();
        a.panel_pool = c->panel.as<double>(); a.fidx_pool = c->fidx.as<int>(); a.atab_pool = c->atab.as<TileAtom>(); a.geo = c->geo.as<TileGeo>();
        a.Bop = op; a.plane_stride = c->plane_stride; a.ldb = c->ldb; a.fR = c->db.fR; a.nbf = c->hb.nbf;
        a.rsx = rsx; a.rsy = rsy; a.rsz = rsz; a.perm = perm; a.tens = d_tens; a.edens = d_edens;
        a.jvec = d_jvec; for (int k = 0; k < 3; ++k) a.B[k] = jpath ? jB3[k] : 0.0;
        a.paramag = c->opts.paramag; a.diamag = c->opts.diamag;
        launch_jtensor(a, c->opts.giao != 0, c->nsm, st);
        CUDA_TRY(cudaGetLastError());
        c->stats.launches += 2;
        c->stats.contract_launches += 1;
        if (prof) cudaEventRecord(c->evpool[3 * b + 2], st);
    }
    if (prof) {
        CUDA_TRY(cudaStreamSynchronize(st));
        float m = 0;
        cudaEventElapsedTime(&m, c->ev[0], c->ev[1]); c->stats.ms_sort += m;
        cudaEventElapsedTime(&m, c->ev[1], c->ev[2]); c->stats.ms_tiles += m;
        for (size_t b = 0; b < nbatch; ++b) {
            if (batch_start[b + 1] - batch_start[b] <= 0) continue;
            cudaEventElapsedTime(&m, c->evpool[3 * b], c->evpool[3 * b + 1]); c->stats.ms_basis += m;
            cudaEventElapsedTime(&m, c->evpool[3 * b + 1], c->evpool[3 * b + 2]); c->stats.ms_contract += m;
        }
    }
    c->stats.n_points += n; c->stats.n_tiles += ntiles; c->stats.sum_nact += sum_nact; c->stats.executed_flops += flops; c->stats.useful_flops += useful;
    const double nbf = c->hb.nbf;
    c->stats.dense_flops += (double)n * (c->opts.giao ? 14.0 * nbf * nbf + 56.0 * nbf : 8.0 * nbf * nbf + 20.0 * nbf);
    return 0;
}

void reset_stats(gimic_b200_ctx *c) { c->stats = gimic_b200_stats{}; }

// host<->device staging helpers --------------------------------------------------------------------
int stage_in(gimic_b200_ctx *c, Buf &b, const double *src, size_t count, int flags, const double **dev) {
    if (flags & GIMIC_B200_DEVICE_PTR) { *dev = src; return 0; }
    if (b.ensure(std::max<size_t>(count, 1) * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (input staging)");
    CUDA_TRY(cudaMemcpyAsync(b.p, src, count * 8, cudaMemcpyHostToDevice, c->stream));
    *dev = b.as<double>();
    return 0;
}

int grid_upload(gimic_b200_ctx *c, const gimic_b200_grid *g, const double **ob, const double **p0, const double **p1, const double **p2,
                const double **w0) {
    const long n0 = g->npts[0], n1 = g->npts[1], n2 = g->npts[2];
    if (n0 <= 0 || n1 <= 0 || n2 <= 0) return fail(GIMIC_B200_EINVAL, "grid with no points");
    std::vector<double> h(12 + 2 * n0 + n1 + n2);
    for (int i = 0; i < 3; ++i) h[i] = g->origin[i];
    for (int i = 0; i < 9; ++i) h[3 + i] = g->basv[i];
    for (long i = 0; i < n0; ++i) { h[12 + i] = g->pts[0][i]; h[12 + n0 + n1 + n2 + i] = g->wgt[0] ? g->wgt[0][i] : 1.0; }
    for (long i = 0; i < n1; ++i) h[12 + n0 + i] = g->pts[1][i];
    for (long i = 0; i < n2; ++i) h[12 + n0 + n1 + i] = g->pts[2][i];
    if (c->gridbuf.ensure(h.size() * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (grid)");
    CUDA_TRY(cudaMemcpyAsync(c->gridbuf.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));   // h goes out of scope
    const double *d = c->gridbuf.as<double>();
    *ob = d; *p0 = d + 12; *p1 = d + 12 + n0; *p2 = d + 12 + n0 + n1; *w0 = d + 12 + n0 + n1 + n2;
    return 0;
}

}  // namespace

// ===================================================================================================
extern "C" {

const char *gimic_b200_last_error(void) { return g_err.c_str(); }
const char *gimic_b200_version(void) { return "gimic-b200 0.1 (sm_100a)"; }

void gimic_b200_default_opts(gimic_b200_opts *o) {
    if (!o) return;
    o->uhf = 0; o->giao = 1; o->diamag = 1; o->paramag = 1; o->screening = 1;
    o->screening_thrs = 1e-6;   // SCREEN_THRS, globals.f90:56 (gimic_interface.f90:41)
    o->device = -1; o->spherical = 0;
}

int gimic_b200_create(gimic_b200_handle *h, const char *mol, const char *xdens, const gimic_b200_opts *opts) {
    if (!h || !mol || !xdens) return fail(GIMIC_B200_EINVAL, "null argument");
    *h = nullptr;
    try {
    std::unique_ptr<gimic_b200_ctx> holder(new gimic_b200_ctx());
    gimic_b200_ctx *c = holder.get();
    if (opts) c->opts = *opts; else gimic_b200_default_opts(&c->opts);
    c->mol_path = mol; c->xdens_path = xdens;
    std::string err;
    if (!gb::parse_mol(mol, c->hb, err)) return fail(GIMIC_B200_EIO, err);
    c->hb.spherical = c->opts.spherical != 0;
    // spherical=on: XDENS is over the 2l+1 components per shell (get_ncgto, intgrl.f90:134-138)
    const int nbf = c->hb.spherical ? c->hb.nbf_sph : c->hb.nbf, nmat = c->opts.uhf ? 8 : 4;
    std::vector<double> dens;
    if (!gb::read_xdens(xdens, nbf, nmat, dens, err)) return fail(GIMIC_B200_EIO, err);
    const size_t nn = (size_t)nbf * nbf;
    if (c->opts.uhf)   // "scaling perturbed densities by 0.5", dens.f90:94-98
        for (int sp = 0; sp < 2; ++sp) for (int b = 1; b < 4; ++b) { double *m = &dens[(sp * 4 + b) * nn]; for (size_t i = 0; i < nn; ++i) m[i] /= 2.0; }
    if (c->hb.turbomole) {   // reorder_dens, dens.f90:100-106,210-234: new(sv(i),sv(j)) = old(i,j)
        std::vector<int> sv;
        if (c->hb.spherical) gb::turbomole_permutation_sph(c->hb, sv); else gb::turbomole_permutation(c->hb, sv);
        std::vector<double> tmp(nn);
        for (int m = 0; m < nmat; ++m) {
            double *src = &dens[m * nn];
            for (int j = 0; j < nbf; ++j) for (int i = 0; i < nbf; ++i) tmp[(size_t)sv[i] + (size_t)nbf * sv[j]] = src[(size_t)i + (size_t)nbf * j];
            std::copy(tmp.begin(), tmp.end(), src);
        }
    }
    int rc = init_device(c);
    if (!rc) rc = finish_create(c, dens.data(), c->opts.uhf ? dens.data() + 4 * nn : nullptr, false);
    if (rc) return rc;
    *h = holder.release();
    return 0;
    } catch (const std::bad_alloc &) { return fail(GIMIC_B200_ENOMEM, "out of host memory"); }
      catch (const std::exception &e) { return fail(GIMIC_B200_EINVAL, e.what()); }
}

int gimic_b200_create_from_arrays(gimic_b200_handle *h, int natoms, const double *coords, const int *nctr_per_atom, const int *ctr_l,
                                  const int *ctr_npf, const double *xp, const double *cc, int turbomole_order, const double *dens_alpha,
                                  const double *dens_beta, int dens_flags, const gimic_b200_opts *opts) {
    if (!h || !coords || !nctr_per_atom || !ctr_l || !ctr_npf || !xp || !cc || !dens_alpha) return fail(GIMIC_B200_EINVAL, "null argument");
    *h = nullptr;
    try {
    std::unique_ptr<gimic_b200_ctx> holder(new gimic_b200_ctx());
    gimic_b200_ctx *c = holder.get();
    if (opts) c->opts = *opts; else gimic_b200_default_opts(&c->opts);
    std::string err;
    if (!gb::basis_from_arrays(natoms, coords, nctr_per_atom, ctr_l, ctr_npf, xp, cc, turbomole_order, c->hb, err)) return fail(GIMIC_B200_EINVAL, err);
    c->hb.spherical = c->opts.spherical != 0;
    int rc = init_device(c);
    if (!rc) rc = finish_create(c, dens_alpha, dens_beta, (dens_flags & GIMIC_B200_DEVICE_PTR) != 0);
    if (rc) return rc;
    *h = holder.release();
    return 0;
    } catch (const std::bad_alloc &) { return fail(GIMIC_B200_ENOMEM, "out of host memory"); }
      catch (const std::exception &e) { return fail(GIMIC_B200_EINVAL, e.what()); }
}

int gimic_b200_destroy(gimic_b200_handle h) { delete h; return 0; }
int gimic_b200_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(GIMIC_B200_ECUDA, "no CUDA device available: gimic-b200 has no CPU path"); }
    return ndev;
}
int gimic_b200_nbf(gimic_b200_handle h) { return h ? (h->hb.spherical ? h->hb.nbf_sph : h->hb.nbf) : fail(GIMIC_B200_EINVAL, "null handle"); }
int gimic_b200_natoms(gimic_b200_handle h) { return h ? h->hb.natoms : fail(GIMIC_B200_EINVAL, "null handle"); }
int gimic_b200_is_uhf(gimic_b200_handle h) { return h ? h->opts.uhf : fail(GIMIC_B200_EINVAL, "null handle"); }
int gimic_b200_atom_coords(gimic_b200_handle h, double *xyz) {
    if (!h || !xyz) return fail(GIMIC_B200_EINVAL, "null argument");
    std::copy(h->hb.xyz.begin(), h->hb.xyz.end(), xyz);
    return 0;
}
int gimic_b200_set_profiling(gimic_b200_handle h, int enable) { if (!h) return fail(GIMIC_B200_EINVAL, "null handle"); h->profiling = enable != 0; return 0; }
int gimic_b200_get_stats(gimic_b200_handle h, gimic_b200_stats *out) { if (!h || !out) return fail(GIMIC_B200_EINVAL, "null argument"); *out = h->stats; return 0; }

int gimic_b200_calc_fields(gimic_b200_handle c, long n, const double *r, const double *B3, int spincase, double *tens, double *jvec,
                           double *jmod, double *acid, double *edens, double *divj, double divj_h, int flags) {
    if (!c || (n > 0 && !r)) return fail(GIMIC_B200_EINVAL, "null argument");
    if ((jvec || jmod || divj) && !B3) return fail(GIMIC_B200_EINVAL, "jvec/jmod/divj need the magnetic field direction");
    if (n < 0) return fail(GIMIC_B200_EINVAL, "negative point count");
    CUDA_TRY(cudaSetDevice(c->device));
    reset_stats(c);
    if (n == 0) return 0;
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    cudaStream_t st = c->stream;
    cudaEventRecord(c->ev_call[0], st);
    const double *d_r = nullptr;
    if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, flags, &d_r)) return rc;
    // Only J (and |J|, rho) wanted: contract with B before instead of after the GEMM -- 2 operand planes instead of 4 and one
    // tap weight per row instead of three (compute_jvectors, jfield.f90:167-184, folded into the contraction).
    const bool jpath = !tens && !acid && !divj && (jvec || jmod);
    double *d_tens = tens;
    if (!jpath && (!dev || !tens)) { if (c->tens_tmp.ensure((size_t)9 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (tensors)"); d_tens = c->tens_tmp.as<double>(); }
    // scalar/vector field outputs on the device
    const size_t nf = (size_t)n;
    double *d_jvec = jvec, *d_jmod = jmod, *d_acid = acid, *d_edens = edens, *d_divj = divj;
    if (!dev) {
        if (c->f_tmp.ensure(nf * 8 * 7)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (fields)");
        double *f = c->f_tmp.as<double>();
        d_jvec = jvec ? f : nullptr; d_jmod = jmod ? f + 3 * nf : nullptr; d_acid = acid ? f + 4 * nf : nullptr;
        d_edens = edens ? f + 5 * nf : nullptr; d_divj = divj ? f + 6 * nf : nullptr;
    }
    if (jpath) {
        if (!d_jvec) {   // |J| alone: J goes to scratch
            if (c->jv6.ensure((size_t)3 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (jvec)");
            d_jvec = c->jv6.as<double>();
        }
        if (int rc = run_tensors(c, n, d_r, spincase, nullptr, d_edens, B3, d_jvec)) return rc;
        if (d_jmod) { gb::launch_jmod(n, d_r, d_jvec, B3, d_jmod, st); c->stats.launches += 1; }
        if (!jvec) d_jvec = nullptr;
    } else if (int rc = run_tensors(c, n, d_r, spincase, d_tens, d_edens)) return rc;
    if (c->profiling) cudaEventRecord(c->ev[0], st);
    if (!jpath && (d_jvec || d_jmod || d_acid)) {
        const double zero3[3] = {0, 0, 0};
        gb::launch_fields(n, d_r, d_tens, B3 ? B3 : zero3, d_jvec, d_jmod, d_acid, st);
        c->stats.launches += 1;
    }
    if (c->profiling) { cudaEventRecord(c->ev[1], st); cudaEventSynchronize(c->ev[1]); float m = 0; cudaEventElapsedTime(&m, c->ev[0], c->ev[1]); c->stats.ms_fields += m; }
    if (d_divj) {
        // div J by central differences of J = T.B at r +- h e_a: 6 more tensor passes (no reference semantics at this commit)
        const double hstep = divj_h > 0 ? divj_h : 1e-3;
        if (c->shift.ensure((size_t)18 * n * 8) || c->jv6.ensure((size_t)(54 + 18) * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (divj)");
        double *r6 = c->shift.as<double>(), *t6 = c->jv6.as<double>(), *v6 = t6 + (size_t)54 * n;
        gb::launch_shift_points(n, d_r, hstep, r6, st);
        gimic_b200_stats keep = c->stats;
        if (int rc = run_tensors(c, 6 * n, r6, spincase, t6, nullptr)) return rc;
        keep.launches = c->stats.launches; c->stats = keep;   // statistics describe the primary pass only
        gb::launch_fields(6 * n, r6, t6, B3, v6, nullptr, nullptr, st);
        gb::launch_divj(n, v6, hstep, d_divj, st);
        c->stats.launches += 3;
    }
    CUDA_TRY(cudaGetLastError());
    if (!dev) {
        if (tens) CUDA_TRY(cudaMemcpyAsync(tens, d_tens, (size_t)9 * n * 8, cudaMemcpyDeviceToHost, st));
        if (jvec) CUDA_TRY(cudaMemcpyAsync(jvec, d_jvec, nf * 24, cudaMemcpyDeviceToHost, st));
        if (jmod) CUDA_TRY(cudaMemcpyAsync(jmod, d_jmod, nf * 8, cudaMemcpyDeviceToHost, st));
        if (acid) CUDA_TRY(cudaMemcpyAsync(acid, d_acid, nf * 8, cudaMemcpyDeviceToHost, st));
        if (edens) CUDA_TRY(cudaMemcpyAsync(edens, d_edens, nf * 8, cudaMemcpyDeviceToHost, st));
        if (divj) CUDA_TRY(cudaMemcpyAsync(divj, d_divj, nf * 8, cudaMemcpyDeviceToHost, st));
    }
    cudaEventRecord(c->ev_call[1], st);
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->stats.ms_total, c->ev_call[0], c->ev_call[1]);
    return 0;
}

int gimic_b200_calc_jtensors(gimic_b200_handle h, long n, const double *r, int spincase, double *tens, int flags) {
    if (!tens && n > 0) return fail(GIMIC_B200_EINVAL, "null argument");
    return gimic_b200_calc_fields(h, n, r, nullptr, spincase, tens, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, flags);
}

int gimic_b200_fields_from_tensors(gimic_b200_handle c, long n, const double *r, const double *tens, const double *B3, double *jvec,
                                   double *jmod, double *acid, int flags) {
    if (!c || !tens || !B3 || (jmod && !r)) return fail(GIMIC_B200_EINVAL, "null argument");
    if (n <= 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    cudaStream_t st = c->stream;
    const double *d_r = r, *d_t = nullptr;
    if (r) { if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, flags, &d_r)) return rc; }
    if (int rc = stage_in(c, c->tens_tmp, tens, (size_t)9 * n, flags, &d_t)) return rc;
    const size_t nf = (size_t)n;
    double *d_jvec = jvec, *d_jmod = jmod, *d_acid = acid;
    if (!dev) {
        if (c->f_tmp.ensure(nf * 8 * 5)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (fields)");
        double *f = c->f_tmp.as<double>();
        d_jvec = jvec ? f : nullptr; d_jmod = jmod ? f + 3 * nf : nullptr; d_acid = acid ? f + 4 * nf : nullptr;
    }
    gb::launch_fields(n, d_r, d_t, B3, d_jvec, d_jmod, d_acid, st);
    CUDA_TRY(cudaGetLastError());
    if (!dev) {
        if (jvec) CUDA_TRY(cudaMemcpyAsync(jvec, d_jvec, nf * 24, cudaMemcpyDeviceToHost, st));
        if (jmod) CUDA_TRY(cudaMemcpyAsync(jmod, d_jmod, nf * 8, cudaMemcpyDeviceToHost, st));
        if (acid) CUDA_TRY(cudaMemcpyAsync(acid, d_acid, nf * 8, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int gimic_b200_jmod_from_jvec(gimic_b200_handle c, long n, const double *r, const double *jvec, const double *B3, double *jmod, int flags) {
    if (!c || !r || !jvec || !B3 || !jmod) return fail(GIMIC_B200_EINVAL, "null argument");
    if (n <= 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    const double *d_r, *d_j;
    if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, flags, &d_r)) return rc;
    if (int rc = stage_in(c, c->tens_tmp, jvec, (size_t)3 * n, flags, &d_j)) return rc;
    double *d_m = jmod;
    if (!dev) { if (c->f_tmp.ensure((size_t)n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (jmod)"); d_m = c->f_tmp.as<double>(); }
    gb::launch_jmod(n, d_r, d_j, B3, d_m, c->stream);
    CUDA_TRY(cudaGetLastError());
    if (!dev) CUDA_TRY(cudaMemcpyAsync(jmod, d_m, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int gimic_b200_calc_jtensors_grid(gimic_b200_handle c, const gimic_b200_grid *g, long lo, long hi, int spincase, double *tens, int flags) {
    if (!c || !g || !tens) return fail(GIMIC_B200_EINVAL, "null argument");
    const long ntot = (long)g->npts[0] * g->npts[1] * g->npts[2];
    if (lo < 0 || hi > ntot || lo > hi) return fail(GIMIC_B200_EINVAL, "grid index range out of bounds");
    const long n = hi - lo;
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    reset_stats(c);
    const double *ob, *p0, *p1, *p2, *w0;
    if (int rc = grid_upload(c, g, &ob, &p0, &p1, &p2, &w0)) return rc;
    cudaEventRecord(c->ev_call[0], c->stream);
    if (c->r_in.ensure((size_t)3 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (grid points)");
    gb::launch_grid_points(ob, p0, p1, p2, g->npts[0], g->npts[1], g->npts[2], lo, hi, c->r_in.as<double>(), c->stream);
    c->stats.launches += 1;
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    double *d_tens = tens;
    if (!dev) { if (c->tens_tmp.ensure((size_t)9 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (tensors)"); d_tens = c->tens_tmp.as<double>(); }
    if (int rc = run_tensors(c, n, c->r_in.as<double>(), spincase, d_tens, nullptr)) return rc;
    if (!dev) CUDA_TRY(cudaMemcpyAsync(tens, d_tens, (size_t)9 * n * 8, cudaMemcpyDeviceToHost, c->stream));
    cudaEventRecord(c->ev_call[1], c->stream);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->stats.ms_total, c->ev_call[0], c->ev_call[1]);
    return 0;
}

}  // extern "C" (reopened below)

// Plane / volume quadrature for `ng` grids in ONE tensor pass: the points of all grids (rows j in [jlo_g, jhi_g) of each) are
// concatenated, run through sort -> tiles -> basis -> contraction together, and reduced per grid.  A 36x36 Gauss plane is only
// 10 tiles -- a current-profile scan of hundreds of such planes fills the 148 SMs only when batched.
namespace {
int integrate_many(gimic_b200_ctx *c, int ng, const gimic_b200_grid *grids, const double *B3s, int spincase, int what,
                   const int *jlos, const int *jhis, double *out7s) {
    for (int k = 0; k < 7 * ng; ++k) out7s[k] = 0.0;
    CUDA_TRY(cudaSetDevice(c->device));
    reset_stats(c);
    // layout of the grid tables: per grid [origin 3 | basv 9 | pts0 | pts1 | pts2 | wgt0]
    std::vector<size_t> goff(ng + 1, 0), roff(ng + 1, 0), rowoff(ng + 1, 0);
    for (int g = 0; g < ng; ++g) {
        const gimic_b200_grid &G = grids[g];
        const long n0 = G.npts[0], n1 = G.npts[1], n2 = G.npts[2];
        if (n0 <= 0 || n1 <= 0 || n2 <= 0) return fail(GIMIC_B200_EINVAL, "grid with no points");
        if (jlos[g] < 0 || jhis[g] > n1 || jlos[g] > jhis[g]) return fail(GIMIC_B200_EINVAL, "row range out of bounds");
        const size_t nrows = (size_t)(jhis[g] - jlos[g]) * n2;
        goff[g + 1] = goff[g] + 12 + 2 * n0 + n1 + n2;
        rowoff[g + 1] = rowoff[g] + nrows;
        roff[g + 1] = roff[g] + nrows * n0;
    }
    const size_t n = roff[ng], nrows_tot = rowoff[ng];
    if (n == 0) return 0;
    std::vector<double> h(goff[ng]), wrow(nrows_tot);
    for (int g = 0; g < ng; ++g) {
        const gimic_b200_grid &G = grids[g];
        const long n0 = G.npts[0], n1 = G.npts[1], n2 = G.npts[2];
        double *q = h.data() + goff[g];
        for (int i = 0; i < 3; ++i) q[i] = G.origin[i];
        for (int i = 0; i < 9; ++i) q[3 + i] = G.basv[i];
        for (long i = 0; i < n0; ++i) { q[12 + i] = G.pts[0][i]; q[12 + n0 + n1 + n2 + i] = G.wgt[0] ? G.wgt[0][i] : 1.0; }
        for (long i = 0; i < n1; ++i) q[12 + n0 + i] = G.pts[1][i];
        for (long i = 0; i < n2; ++i) q[12 + n0 + n1 + i] = G.pts[2][i];
        const int nj = jhis[g] - jlos[g];
        for (int k = 0; k < n2; ++k) for (int j = 0; j < nj; ++j)
            wrow[rowoff[g] + (size_t)k * nj + j] = (G.wgt[1] ? G.wgt[1][jlos[g] + j] : 1.0) * (G.wgt[2] ? G.wgt[2][k] : 1.0);
    }
    if (c->gridbuf.ensure(h.size() * 8) || c->r_in.ensure(3 * n * 8) || c->tens_tmp.ensure(9 * n * 8) ||
        c->quad.ensure((8 * nrows_tot + 7 * (size_t)ng + 8) * 8))
        return fail(GIMIC_B200_ENOMEM, "device allocation failed (integration)");
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(c->gridbuf.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice, st));
    double *d_r = c->r_in.as<double>();
    double *d_part = c->quad.as<double>(), *d_wrow = d_part + 7 * nrows_tot, *d_out = d_wrow + nrows_tot;
    CUDA_TRY(cudaMemcpyAsync(d_wrow, wrow.data(), nrows_tot * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_out, 0, 7 * (size_t)ng * 8, st));
    for (int g = 0; g < ng; ++g) {   // rows (k, j in [jlo,jhi)), i fastest: the loop nest of integral.f90:113-123
        const gimic_b200_grid &G = grids[g];
        const int p1 = G.npts[0], p2 = G.npts[1], p3 = G.npts[2], nj = jhis[g] - jlos[g];
        const double *d = c->gridbuf.as<double>() + goff[g];
        for (int k = 0; k < p3 && nj > 0; ++k) {
            const long lo = ((long)k * p2 + jlos[g]) * p1, hi = ((long)k * p2 + jhis[g]) * p1;
            gb::launch_grid_points(d, d + 12, d + 12 + p1, d + 12 + p1 + p2, p1, p2, p3, lo, hi, d_r + 3 * (roff[g] + (size_t)k * nj * p1), st);
            c->stats.launches += 1;
        }
    }
    if (int rc = run_tensors(c, (long)n, d_r, spincase, c->tens_tmp.as<double>(), nullptr)) return rc;
    for (int g = 0; g < ng; ++g) {
        const gimic_b200_grid &G = grids[g];
        const int p1 = G.npts[0], p2 = G.npts[1], nrows = (int)(rowoff[g + 1] - rowoff[g]);
        if (nrows == 0) continue;
        gb::QuadArgs q;
        q.tens = c->tens_tmp.as<double>() + 9 * roff[g]; q.p1 = p1; q.nrows = nrows; q.r = d_r + 3 * roff[g];
        q.w1 = c->gridbuf.as<double>() + goff[g] + 12 + p1 + p2 + G.npts[2]; q.wrow = d_wrow + rowoff[g];
        auto gp = [&](int i, int j, int k, double *r) {   // gridpoint, grid.f90:498-511 (0-based here)
            for (int d = 0; d < 3; ++d) r[d] = G.origin[d] + G.pts[0][i] * G.basv[d] + G.pts[1][j] * G.basv[3 + d] + G.pts[2][k] * G.basv[6 + d];
        };
        double v1[3], v2[3];
        gp(p1 - 1, 0, 0, v1); gp(0, p2 - 1, 0, v2);   // grid_center, grid.f90:529-541
        for (int d = 0; d < 3; ++d) { q.center[d] = (v1[d] + v2[d]) * 0.5; q.B[d] = B3s[3 * g + d]; q.normal[d] = G.basv[6 + d]; }
        q.radius = (G.radius > 0.0) ? G.radius : 1e300;
        q.what = what; q.row_partials = d_part + 7 * rowoff[g]; q.out7 = d_out + 7 * g;
        gb::launch_quadrature(q, st);
        c->stats.launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out7s, d_out, 7 * (size_t)ng * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}
}  // namespace

extern "C" {

int gimic_b200_integrate(gimic_b200_handle c, const gimic_b200_grid *g, const double *B3, int spincase, int what, int jlo, int jhi,
                         double *out7) {
    if (!c || !g || !B3 || !out7) return fail(GIMIC_B200_EINVAL, "null argument");
    return integrate_many(c, 1, g, B3, spincase, what, &jlo, &jhi, out7);
}

int gimic_b200_integrate_batch(gimic_b200_handle c, int ngrids, const gimic_b200_grid *grids, const double *B3s, int spincase,
                               int what, double *out7s) {
    if (!c || !grids || !B3s || !out7s || ngrids < 0) return fail(GIMIC_B200_EINVAL, "null argument");
    if (ngrids == 0) return 0;
    std::vector<int> lo(ngrids, 0), hi(ngrids);
    for (int g = 0; g < ngrids; ++g) hi[g] = grids[g].npts[1];
    return integrate_many(c, ngrids, grids, B3s, spincase, what, lo.data(), hi.data(), out7s);
}

int gimic_b200_calc_basis(gimic_b200_handle c, long n, const double *r, double *bf, double *dr, int flags) {
    if (!c || (n > 0 && !r) || (!bf && !dr)) return fail(GIMIC_B200_EINVAL, "null argument");
    if (n <= 0) return 0;
    if (n > 2147483647L) return fail(GIMIC_B200_EINVAL, "too many points");
    CUDA_TRY(cudaSetDevice(c->device));
    bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    const size_t nb = (size_t)c->hb.nbf;
    const double *d_r = nullptr;
    if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, flags, &d_r)) return rc;
    if (c->hb.spherical) {
        // sbf = po . bf, sdr = po . dr (bfeval.f90:116-118, 330-333): cartesian vectors from the kernel, the block-diagonal
        // projection on the host (this entry point is a diagnostic; the tensor path folds po into the densities instead)
        if (dev) return fail(GIMIC_B200_EINVAL, "calc_basis with spherical=on needs host output buffers");
        const size_t ns = (size_t)c->hb.nbf_sph;
        if (c->f_tmp.ensure((size_t)n * nb * 4 * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (basis vectors)");
        double *d_b = c->f_tmp.as<double>(), *d_d = d_b + (size_t)n * nb;
        gb::launch_basis_dense(c->db, c->d_f2user, n, d_r, d_b, d_d, c->stream);
        CUDA_TRY(cudaGetLastError());
        std::vector<double> hc((size_t)n * nb * 4);
        CUDA_TRY(cudaMemcpyAsync(hc.data(), d_b, hc.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        std::vector<double> po[gb::MAX_L + 1];
        for (int l = 0; l <= gb::MAX_L; ++l) gb::c2s_rows(l, c->hb.turbomole, po[l]);
        for (long v = 0; v < 4 * n; ++v) {   // v < n: bf rows; then dr rows (3 per point), same order as the cartesian output
            double *out = v < n ? (bf ? bf + (size_t)v * ns : nullptr) : (dr ? dr + (size_t)(v - n) * ns : nullptr);
            if (!out) continue;
            const double *in = hc.data() + (size_t)v * nb;
            for (const gb::Shell &s : c->hb.shells)
                for (int q = 0; q < s.nsph; ++q) {
                    double acc = 0.0;
                    for (int k = 0; k < s.ncomp; ++k) acc += po[s.l][(size_t)q * s.ncomp + k] * in[s.user_off + k];
                    out[s.sph_off + q] = acc;
                }
        }
        return 0;
    }
    double *d_bf = bf, *d_dr = dr;
    if (!dev) {
        if (c->f_tmp.ensure((size_t)n * nb * 4 * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (basis vectors)");
        d_bf = bf ? c->f_tmp.as<double>() : nullptr;
        d_dr = dr ? c->f_tmp.as<double>() + (size_t)n * nb : nullptr;
    }
    gb::launch_basis_dense(c->db, c->d_f2user, n, d_r, d_bf, d_dr, c->stream);
    CUDA_TRY(cudaGetLastError());
    if (!dev) {
        if (bf) CUDA_TRY(cudaMemcpyAsync(bf, d_bf, (size_t)n * nb * 8, cudaMemcpyDeviceToHost, c->stream));
        if (dr) CUDA_TRY(cudaMemcpyAsync(dr, d_dr, (size_t)n * nb * 24, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int gimic_b200_property(gimic_b200_handle c, long n, const double *r, const double *w, const double *tens, int natoms,
                        const double *coords, int nseg, const long *seg_end, double *part, int flags) {
    if (!c || !r || !w || !tens || !coords || !seg_end || !part || natoms < 0 || nseg <= 0 || n < 0) return fail(GIMIC_B200_EINVAL, "bad argument");
    if (seg_end[nseg - 1] != n) return fail(GIMIC_B200_EINVAL, "segment ends must be cumulative and finish at n");
    for (int i = 1; i < nseg; ++i) if (seg_end[i] < seg_end[i - 1]) return fail(GIMIC_B200_EINVAL, "segment ends must be non-decreasing");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const double *d_r, *d_t, *d_w;
    if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, flags, &d_r)) return rc;
    if (int rc = stage_in(c, c->tens_tmp, tens, (size_t)9 * n, flags, &d_t)) return rc;
    if (int rc = stage_in(c, c->f_tmp, w, (size_t)n, flags, &d_w)) return rc;
    const size_t nout = (size_t)(natoms + 1) * nseg * 5;
    if (c->quad.ensure((nout + 3 * (size_t)natoms + 1) * 8 + (size_t)nseg * 8 + 64)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (property)");
    double *d_out = c->quad.as<double>(), *d_xyz = d_out + nout;
    long *d_seg = reinterpret_cast<long *>(d_xyz + 3 * (size_t)natoms + 1);
    if (natoms) CUDA_TRY(cudaMemcpyAsync(d_xyz, coords, (size_t)3 * natoms * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_seg, seg_end, (size_t)nseg * 8, cudaMemcpyHostToDevice, st));
    gb::launch_property(n, d_r, d_w, d_t, natoms, d_xyz, nseg, d_seg, d_out, st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(part, d_out, nout * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int gimic_b200_property_integrand(gimic_b200_handle c, long n, const double *r, const double *tens, const double *centre3,
                                  double *out4, int flags) {
    if (!c || !r || !tens || !out4) return fail(GIMIC_B200_EINVAL, "null argument");
    if (n <= 0) return 0;
    CUDA_TRY(cudaSetDevice(c->device));
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    const double *d_r, *d_t;
    if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, flags, &d_r)) return rc;
    if (int rc = stage_in(c, c->tens_tmp, tens, (size_t)9 * n, flags, &d_t)) return rc;
    double *d_o = out4;
    if (!dev) { if (c->f_tmp.ensure((size_t)4 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (integrand)"); d_o = c->f_tmp.as<double>(); }
    const double zero[3] = {0, 0, 0};
    gb::launch_property_integrand(n, d_r, d_t, centre3 ? 0 : 1, centre3 ? centre3 : zero, d_o, c->stream);
    CUDA_TRY(cudaGetLastError());
    if (!dev) CUDA_TRY(cudaMemcpyAsync(out4, d_o, (size_t)4 * n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int gimic_b200_gauss_points(double a, double b, int npts, int order, int quadrature, double *pts, double *wgts) {
    if (!pts || !wgts || npts <= 0) return fail(GIMIC_B200_EINVAL, "bad argument");
    int rc = gb::gauss_blocks(a, b, npts, order, quadrature, pts, wgts);
    if (rc == -1) return fail(GIMIC_B200_EINVAL, "*** integration did not converge!");
    if (rc) return fail(GIMIC_B200_EINVAL, "gaussgrid(): npts is not dividable by ngp!");
    return 0;
}

int gimic_b200_convert_xdens(const char *xdens_text, int nbf, int nmat, const char *xdens_binary) {
    if (!xdens_text || !xdens_binary || nbf <= 0 || (nmat != 4 && nmat != 8)) return fail(GIMIC_B200_EINVAL, "bad argument");
    try {
    std::vector<double> v; std::string err;
    if (!gb::read_xdens(xdens_text, nbf, nmat, v, err)) return fail(GIMIC_B200_EIO, err);
    if (!gb::write_xdens_binary(xdens_binary, nbf, nmat, v.data(), err)) return fail(GIMIC_B200_EIO, err);
    return 0;
    } catch (const std::bad_alloc &) { return fail(GIMIC_B200_ENOMEM, "out of host memory"); }
      catch (const std::exception &e) { return fail(GIMIC_B200_EINVAL, e.what()); }
}

long gimic_b200_format_e(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap) {
    if (n < 0 || (n > 0 && (!v || !out)) || w <= 0 || w > 40 || d <= 0 || d > 30) { fail(GIMIC_B200_EINVAL, "bad argument"); return GIMIC_B200_EINVAL; }
    long rc = gb::format_fortran_e(n, v, w, d, per_line, first_count, prefix, out, cap);
    if (rc < 0) { fail(GIMIC_B200_EINVAL, "output buffer too small"); return GIMIC_B200_EINVAL; }
    return rc;
}

long gimic_b200_format_f(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap) {
    if (n < 0 || (n > 0 && (!v || !out)) || w <= 0 || w > 40 || d < 0 || d > 30) { fail(GIMIC_B200_EINVAL, "bad argument"); return GIMIC_B200_EINVAL; }
    long rc = gb::format_fortran(n, v, 'F', w, d, per_line, first_count, prefix, out, cap);
    if (rc < 0) { fail(GIMIC_B200_EINVAL, "output buffer too small"); return GIMIC_B200_EINVAL; }
    return rc;
}

int gimic_b200_mol_geometry(const char *mol, int max_atoms, double *xyz, char *symbols2) {
    if (!mol) return fail(GIMIC_B200_EINVAL, "null argument");
    try {
    gb::HostBasis hb; std::string err;
    if (!gb::parse_mol(mol, hb, err)) return fail(GIMIC_B200_EIO, err);
    for (int a = 0; a < hb.natoms && a < max_atoms; ++a) {
        if (xyz) for (int k = 0; k < 3; ++k) xyz[3 * a + k] = hb.xyz[3 * a + k];
        if (symbols2) { symbols2[2 * a] = hb.symbol[a].size() > 0 ? hb.symbol[a][0] : ' '; symbols2[2 * a + 1] = hb.symbol[a].size() > 1 ? hb.symbol[a][1] : ' '; }
    }
    return hb.natoms;
    } catch (const std::bad_alloc &) { return fail(GIMIC_B200_ENOMEM, "out of host memory"); }
      catch (const std::exception &e) { return fail(GIMIC_B200_EINVAL, e.what()); }
}

int gimic_b200_mol_summary(const char *mol, int *info5) {
    if (!mol || !info5) return fail(GIMIC_B200_EINVAL, "null argument");
    try {
    gb::HostBasis hb; std::string err;
    if (!gb::parse_mol(mol, hb, err)) return fail(GIMIC_B200_EIO, err);
    info5[0] = hb.natoms; info5[1] = hb.ngto; info5[2] = hb.nbf; info5[3] = hb.turbomole ? 1 : 0; info5[4] = hb.nbf_sph;
    return 0;
    } catch (const std::bad_alloc &) { return fail(GIMIC_B200_ENOMEM, "out of host memory"); }
      catch (const std::exception &e) { return fail(GIMIC_B200_EINVAL, e.what()); }
}

int gimic_b200_c2s_rows(int l, int turbomole_order, double *po) {
    if (!po || l < 0 || l > gb::MAX_L) return fail(GIMIC_B200_EINVAL, "bad argument");
    std::vector<double> rows;
    gb::c2s_rows(l, turbomole_order != 0, rows);
    std::copy(rows.begin(), rows.end(), po);
    return 0;
}

// ------------------------------------------------------------------------ legacy single-point boundary
namespace {
gimic_b200_ctx *g_default = nullptr;
double g_magnet[3] = {0, 0, 0};
int g_spin = GIMIC_B200_TOTAL;
double g_screen = 1e-6;
[[noreturn]] void legacy_stop(const char *what) {   // the Fortran side would `stop`
    std::fprintf(stderr, " *** gimic-b200: %s: %s\n", what, g_err.c_str());
    std::exit(1);
}
}  // namespace

void gimic_init(const char *mol, const char *xdens) {
    if (g_default) { delete g_default; g_default = nullptr; }
    gimic_b200_opts o; gimic_b200_default_opts(&o);          // gimic_interface.f90:37-51
    g_magnet[0] = g_magnet[1] = g_magnet[2] = 0.0; g_spin = GIMIC_B200_TOTAL;
    if (gimic_b200_create(&g_default, mol, xdens, &o)) legacy_stop("gimic_init");
}
void gimic_finalize(void) { delete g_default; g_default = nullptr; }
void gimic_set_uhf(int *uhf) {
    // The reference only flips settings%is_uhf (gimic_interface.f90:80-86) and then reads unallocated beta
    // densities; here the context is rebuilt from the same files as an open-shell one.
    if (!g_default || !uhf) return;
    const int want = *uhf != 0;
    if (want == (g_default->opts.uhf != 0)) return;
    gimic_b200_opts o = g_default->opts; o.uhf = want;
    std::string m = g_default->mol_path, x = g_default->xdens_path;
    delete g_default; g_default = nullptr;
    if (gimic_b200_create(&g_default, m.c_str(), x.c_str(), &o)) legacy_stop("gimic_set_uhf");
}
void gimic_set_magnet(const double *b) { if (b) for (int i = 0; i < 3; ++i) g_magnet[i] = b[i]; }
void gimic_set_spin(const char *s) {
    if (!s) return;
    if (!std::strcmp(s, "alpha")) g_spin = GIMIC_B200_ALPHA;
    else if (!std::strcmp(s, "beta")) g_spin = GIMIC_B200_BETA;
    else if (!std::strcmp(s, "total")) g_spin = GIMIC_B200_TOTAL;
    else if (!std::strcmp(s, "spindens")) g_spin = GIMIC_B200_SPINDENS;
    else { g_err = s; legacy_stop("Invalid spin case."); }   // gimic_interface.f90:111-112
}
void gimic_set_screening(const double *thrs) { if (thrs) g_screen = *thrs; }   // like the reference: stored, radii are not rebuilt (gimic_interface.f90:116-119)
void gimic_calc_jtensor(const double *r, double *jt) {
    if (!g_default) { g_err = "gimic_init() has not been called"; legacy_stop("gimic_calc_jtensor"); }
    if (gimic_b200_calc_jtensors(g_default, 1, r, g_spin, jt, 0)) legacy_stop("gimic_calc_jtensor");
}
void gimic_calc_jvector(const double *r, double *jv) {
    if (!g_default) { g_err = "gimic_init() has not been called"; legacy_stop("gimic_calc_jvector"); }
    if (gimic_b200_calc_fields(g_default, 1, r, g_magnet, g_spin, nullptr, jv, nullptr, nullptr, nullptr, nullptr, 0.0, 0)) legacy_stop("gimic_calc_jvector");
}
void gimic_calc_modj(const double *r, double *d) {
    // reference: stop 'gimic_calc_modj(): NOT IMPLEMENTED YET!' (gimic_interface.f90:153-163); here |J| for the stored magnet
    if (!g_default) { g_err = "gimic_init() has not been called"; legacy_stop("gimic_calc_modj"); }
    double jv[3];
    if (gimic_b200_calc_fields(g_default, 1, r, g_magnet, g_spin, nullptr, jv, nullptr, nullptr, nullptr, nullptr, 0.0, 0)) legacy_stop("gimic_calc_modj");
    *d = std::sqrt(jv[0] * jv[0] + jv[1] * jv[1] + jv[2] * jv[2]);
}
void gimic_get_gauss_points(double *a, double *b, int *npts, int *order, double *pts, double *wgts) {
    if (gimic_b200_gauss_points(*a, *b, *npts, *order, 0, pts, wgts)) legacy_stop("gimic_get_gauss_points");
}
void mkgausspoints(double *a, double *b, int *npts, int *order, double *pts, double *wgts) { gimic_get_gauss_points(a, b, npts, order, pts, wgts); }

}  // extern "C"
