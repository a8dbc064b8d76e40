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

// GIMIC_B200_POISON=1 (debugging): every fresh workspace allocation is filled with 0xFF bytes (NaN as doubles, -1 as ints), so that a
// read of memory nobody wrote shows up as NaN in the results instead of depending on what the allocator handed out.
bool poison_workspaces() {
    static const bool on = [] { const char *e = std::getenv("GIMIC_B200_POISON"); return e && e[0] != '0'; }();
    return on;
}
struct Buf {
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; return -1; } want = bytes; }
        cap = want;
        if (poison_workspaces()) cudaMemset(p, 0xFF, want);
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct Plan {                        // one sorted, tiled point set (and this rank's share of it) -- see build_plan()
    bool valid = false;
    long n = 0;                      // points of the whole set
    int rank = 0, nranks = 1;
    gb::PlanSummary sum{};           // host copy
    long count() const { return valid ? (long)(sum.pt_hi - sum.pt_lo) : 0; }
    int ntiles() const { return valid ? sum.thi - sum.tlo : 0; }
};

struct gimic_b200_ctx {
    int device = 0, nsm = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    gimic_b200_opts opts{};
    gb::HostBasis hb;
    gb::DevBasis db{};
    std::vector<void *> owned;
    int *d_f2user = nullptr;
    double *d_dens[2] = {nullptr, nullptr};   // dens_t%da / %db in the XDENS layout (released once every operand set that needs them is built)
    double *d_op[4] = {nullptr, nullptr, nullptr, nullptr};   // contraction operands per spin case
    double *d_opj[4] = {nullptr, nullptr, nullptr, nullptr};  // J = T.B path: one pair-plane (D, P.B) per spin case, for the field in opj_B
    double opj_B[4][3] = {};
    int nq = gb::NQ, ldb = 0; long long plane_stride = 0;
    double bbox_lo[3] = {0, 0, 0}; double inv_cell = 1.0;
    size_t pool_max_bytes = (size_t)8 << 30;      // set per device in create_common
    // workspaces
    Buf keys0, keys1, vals0, vals1, sorttmp, rs, panel, fidx, atab, misc, r_in, r_in2, tens_tmp, tens_tmp2, f_tmp, f_tmp2, shift, jv6, gridbuf, quad;
    Buf p_seg, p_geo, p_info, p_cnt, p_off, geo, desc, cum, pkeys0, pkeys1, pord0, pord1, tiles, d_summary, p_tops;   // tile plan (k_prepare.cu)
    Buf items, part;                   // work items / partial row sums of sliced tiles (few tiles: see launch_tile_slices)
    Buf gtiles, pgkeys0;               // the plan's tiles in (batch, drain group, cost) order and their sort keys
    gb::PlanSummary *h_summary = nullptr;   // pinned
    Plan plan;
    double split_radius = 2.5;   // bohr: tiles wider than this are cut at their largest consecutive gap if that shrinks them
    bool profiling = false;
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_call[2] = {};
    cudaEvent_t ev_plan[2] = {};       // span of the last gimic_b200_partition_* call
    cudaEvent_t ev_chunk[4] = {};      // [0,1] results of buffer 0/1 ready (compute stream), [2,3] buffer 0/1 drained (copy stream)
    std::vector<cudaEvent_t> evpool;   // per-batch (basis, contract) stamps, resolved at the end of a call
    std::vector<cudaEvent_t> evbatch;  // per-batch "rows ready" events of the overlapped device->host drain
    gimic_b200_stats stats{};
    std::string mol_path, xdens_path;   // for the legacy set_uhf-after-init path

    ~gimic_b200_ctx() {
        cudaSetDevice(device);
        for (void *p : owned) cudaFree(p);
        for (int i = 0; i < 2; ++i) if (d_dens[i]) cudaFree(d_dens[i]);
        for (int i = 0; i < 4; ++i) if (d_op[i]) cudaFree(d_op[i]);
        for (int i = 0; i < 4; ++i) if (d_opj[i]) cudaFree(d_opj[i]);
        for (Buf *b : {&keys0, &keys1, &vals0, &vals1, &sorttmp, &rs, &panel, &fidx, &atab, &misc, &r_in, &r_in2, &tens_tmp, &tens_tmp2, &f_tmp, &f_tmp2,
                       &shift, &jv6, &gridbuf, &quad, &p_seg, &p_geo, &p_info, &p_cnt, &p_off, &geo, &desc, &cum, &pkeys0, &pkeys1, &pord0, &pord1,
                       &tiles, &d_summary, &p_tops, &items, &part, &gtiles, &pgkeys0}) b->release();
        if (h_summary) cudaFreeHost(h_summary);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        for (auto &e : evpool) cudaEventDestroy(e);
        for (auto &e : evbatch) cudaEventDestroy(e);
        for (auto &e : ev_call) if (e) cudaEventDestroy(e);
        for (auto &e : ev_plan) if (e) cudaEventDestroy(e);
        for (auto &e : ev_chunk) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
    }
};

namespace {

template <typename T>
int upload(gimic_b200_ctx *c, const std::vector<T> &v, const T **out) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CUDA_TRY(cudaMalloc(&p, bytes));
    c->owned.push_back(p);
    // On the context's stream: a plain cudaMemcpy from pageable memory may return while the DMA of its last staging chunk is still in
    // flight, and the context's streams are non-blocking (not ordered behind the legacy stream that copy runs on).
    if (!v.empty()) {
        CUDA_TRY(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
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
    {
        std::vector<double> t2(ns), m2(hb.natoms);
        for (int i = 0; i < ns; ++i) t2[i] = (sh_thr[i] + 1e-9) * (sh_thr[i] + 1e-9);
        for (int a = 0; a < hb.natoms; ++a) m2[a] = (maxthr[a] + 1e-9) * (maxthr[a] + 1e-9);
        if (int rc = upload(c, t2, &d.sh_thr2e)) return rc;
        if (int rc = upload(c, m2, &d.atom_maxthr2e)) return rc;
    }
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
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    for (auto &e : c->ev_call) CUDA_TRY(cudaEventCreate(&e));
    for (auto &e : c->ev_plan) CUDA_TRY(cudaEventCreate(&e));
    for (auto &e : c->ev_chunk) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaMallocHost((void **)&c->h_summary, sizeof(gb::PlanSummary)));
    // panel pool: one batch (one k_basis + one k_jtensor launch) per ~pool of Phi/dPhi panels.  8 GB is the configuration of the
    // committed ncu captures and launch lists; GIMIC_B200_POOL_MB=24576 (one launch per 2M-point step) measured +1 %.
    // panel pool of one batch: an eighth of the device memory, at most 24 GB (B200: 22 GB).  Fewer, larger batches mean fewer launch tails:
    // whole 256^3 grid at nbf = 10^4: 4 GB 758.7 ms, 8 GB 747.3 ms, 24 GB 743.8 ms, 48 GB 742.6 ms (profiles/r02_pool_sweep.txt).  The pool is
    // allocated at the size the point set needs, never more.
    c->pool_max_bytes = std::min<size_t>((size_t)24 << 30, std::max<size_t>((size_t)1 << 30, prop.totalGlobalMem / 8));
    if (const char *mb = std::getenv("GIMIC_B200_POOL_MB")) { long v = std::atol(mb); if (v > 0) c->pool_max_bytes = (size_t)v << 20; }
    return 0;
}

int check_spincase(gimic_b200_ctx *c, int &spincase) {
    if (spincase < 0 || spincase > 3) return fail(GIMIC_B200_EINVAL, "invalid spin case");
    if (!c->opts.uhf) {
        if (spincase == GIMIC_B200_BETA) return fail(GIMIC_B200_ESPIN, "ctensor(): beta current requested, but not open-shell system!");
        if (spincase == GIMIC_B200_SPINDENS) return fail(GIMIC_B200_ESPIN, "ctensor(): spindens requested, but not open-shell system!");
        spincase = GIMIC_B200_ALPHA;
    }
    return 0;
}

// Contraction operands of a spin case.  alpha (and beta) are built from the densities at creation; total / spin density of an
// open-shell context are alpha +- beta (jtensor.F90:86-88, 97-99: linear in D and P), built on first use.
int get_operand(gimic_b200_ctx *c, int spincase, const double **op) {
    if (int rc = check_spincase(c, spincase)) return rc;
    if (!c->d_op[spincase]) {
        if (spincase != GIMIC_B200_TOTAL && spincase != GIMIC_B200_SPINDENS) return fail(GIMIC_B200_EINVAL, "operands of this spin case were not built");
        const size_t count = (size_t)((c->nq + 1) / 2) * c->plane_stride;
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, count * sizeof(double)));
        gb::launch_operand_combine((double *)p, c->d_op[GIMIC_B200_ALPHA], c->d_op[GIMIC_B200_BETA], spincase == GIMIC_B200_TOTAL ? 1.0 : -1.0,
                                   (long)count, c->stream);
        CUDA_TRY(cudaGetLastError());
        c->d_op[spincase] = (double *)p;
    }
    *op = c->d_op[spincase];
    return 0;
}

// Operand of the J = T.B path: the tensor is only ever contracted with B, so P_x, P_y, P_z enter as the single matrix
// sum_b B_b P_b (jfield.f90:167-184 applied before instead of after the contraction).  Rebuilt when B changes.
int get_operand_j(gimic_b200_ctx *c, int spincase, const double *B3, const double **op) {
    if (int rc = check_spincase(c, spincase)) return rc;
    const bool same = c->d_opj[spincase] && c->opj_B[spincase][0] == B3[0] && c->opj_B[spincase][1] == B3[1] && c->opj_B[spincase][2] == B3[2];
    if (!same) {
        const double *full = nullptr;
        if (int rc = get_operand(c, spincase, &full)) return rc;
        if (!c->d_opj[spincase]) CUDA_TRY(cudaMalloc((void **)&c->d_opj[spincase], (size_t)c->plane_stride * sizeof(double)));
        gb::launch_operand_j(c->d_opj[spincase], full, c->plane_stride, B3, c->stream);
        CUDA_TRY(cudaGetLastError());
        for (int k = 0; k < 3; ++k) c->opj_B[spincase][k] = B3[k];
    }
    *op = c->d_opj[spincase];
    return 0;
}

int finish_create(gimic_b200_ctx *c, const double *dens_a, const double *dens_b, bool dens_on_device) {
    gb::finalize_basis(c->hb, c->opts.screening != 0, c->opts.screening_thrs);
    // With screening off every tile has all nbf functions active: k_jtensor's K-step mask covers 65536 slots per tile.
    if (!c->opts.screening && c->hb.nbf + 3 * c->hb.natoms > 65536)
        return fail(GIMIC_B200_EINVAL, "more than 65536 basis-function slots with Advanced.screening off: a tile of points cannot hold all functions; "
                                       "switch screening on (exact: screened functions are zeros in the reference too)");
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
    // The densities are only needed to build the alpha / beta operand planes (internal function order, pair-planes); the raw
    // XDENS-layout copy is released right away (3.2 GB per spin at nbf = 10^4).
    for (int sp = 0; sp < (c->opts.uhf ? 2 : 1); ++sp) {
        const double *src = sp ? dens_b : dens_a;
        if (!src) return fail(GIMIC_B200_EINVAL, sp ? "open-shell context needs beta densities" : "densities missing");
        const double *d_src = src;
        if (!dens_on_device) {
            CUDA_TRY(cudaMalloc((void **)&c->d_dens[sp], 4 * nn * sizeof(double)));
            // stream-ordered before k_build_operand (a plain cudaMemcpy from pageable memory returns before its DMA has landed, and
            // c->stream is non-blocking: round 2 found the beta operand of a small open-shell case built from a partly stale buffer
            // once in ~4 runs of the whole parity file)
            CUDA_TRY(cudaMemcpyAsync(c->d_dens[sp], src, 4 * nn * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            d_src = c->d_dens[sp];
        }
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, (size_t)((c->nq + 1) / 2) * c->plane_stride * sizeof(double)));
        c->d_op[sp ? GIMIC_B200_BETA : GIMIC_B200_ALPHA] = (double *)p;
        gb::launch_build_operand((double *)p, c->hb.nbf, c->ldb, c->plane_stride, d_src, nullptr, 0.0, c->d_f2user, c->stream);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (c->d_dens[sp]) { cudaFree(c->d_dens[sp]); c->d_dens[sp] = nullptr; }
    }
    return 0;
}

// ---- the batched tensor pipeline ------------------------------------------------------------------
// Panel doubles of one batch: never smaller than the largest possible tile (every function active, every atom's run padded) -- consecutive
// tiles then never skip a batch number, which the batch table of the plan relies on.
long long plan_pool_doubles(const gimic_b200_ctx *c) {
    const long long max_tile = 4LL * (((long long)c->hb.nbf + 3LL * c->hb.natoms + 7) / 8 * 8) * gb::LDP;
    return std::max((long long)(c->pool_max_bytes / 8), max_tile);
}

// build_plan: Hilbert sort of the points, tiles (with gap splitting), per-tile active-set sizes, this rank's equal-cost share
// of the tile list, panel-pool batches and the processing order -- all on the device; the host reads one PlanSummary.
// d_r: device, 3 x n (AoS).
int build_plan(gimic_b200_ctx *c, long n, const double *d_r, int rank, int nranks) {
    using namespace gb;
    c->plan.valid = false;
    if (n <= 0) return fail(GIMIC_B200_EINVAL, "no points");
    if (n > 2000000000L) return fail(GIMIC_B200_EINVAL, "more than 2e9 points in one call");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(GIMIC_B200_EINVAL, "rank / nranks out of range");
    cudaStream_t st = c->stream;
    const bool prof = c->profiling;
    const long nrun0 = (n + MT - 1) / MT;

    if (c->keys0.ensure(n * 4) || c->keys1.ensure(n * 4) || c->vals0.ensure(n * 4) || c->vals1.ensure(n * 4) ||
        c->rs.ensure((size_t)3 * n * 8) || c->misc.ensure(256) || c->d_summary.ensure(sizeof(PlanSummary)))
        return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed");
    size_t tb = sort_temp_bytes(n);
    if (c->sorttmp.ensure(tb)) return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (sort)");
    double *rsx = c->rs.as<double>(), *rsy = rsx + n, *rsz = rsy + n;
    int *perm = c->vals1.as<int>();

    if (prof) cudaEventRecord(c->ev[0], st);
    launch_morton_keys(d_r, n, c->bbox_lo, c->inv_cell, c->keys0.as<uint32_t>(), c->vals0.as<int>(), st);
    launch_sort_pairs(c->sorttmp.p, tb, c->keys0.as<uint32_t>(), c->keys1.as<uint32_t>(), c->vals0.as<int>(), perm, n, st);
    launch_gather_points(d_r, perm, n, rsx, rsy, rsz, st);
    if (prof) cudaEventRecord(c->ev[1], st);
    c->stats.launches += 4;

    // Gap splitting rarely adds more than a few per cent of tiles; if a thin point set needs more, the summary says so and
    // the plan is rebuilt with room for every possible piece.
    const long long pool_doubles = plan_pool_doubles(c);
    long cap = nrun0 + nrun0 / 2 + 1024;
    PlanSummary &S = c->plan.sum;
    for (int attempt = 0; attempt < 2; ++attempt) {
        cap = std::min<long>(cap, nrun0 * MAXSUB);
        if (c->p_seg.ensure((size_t)nrun0 * MAXSUB * sizeof(TileSeg)) || c->p_geo.ensure((size_t)nrun0 * MAXSUB * sizeof(TileGeo)) ||
            c->p_info.ensure((size_t)nrun0 * MAXSUB * sizeof(TileInfo)) || c->p_cnt.ensure((size_t)nrun0 * 4) || c->p_off.ensure((size_t)(nrun0 + 1) * 4) ||
            c->geo.ensure((size_t)cap * sizeof(TileGeo)) || c->desc.ensure((size_t)cap * sizeof(TileDesc)) || c->cum.ensure((size_t)(cap + 1) * sizeof(TileCum)) ||
            c->pkeys0.ensure((size_t)cap * 8) || c->pkeys1.ensure((size_t)cap * 8) || c->pord0.ensure((size_t)cap * 4) || c->pord1.ensure((size_t)cap * 4) ||
            c->tiles.ensure((size_t)cap * sizeof(TileDesc)) || c->gtiles.ensure((size_t)cap * sizeof(TileDesc)) || c->pgkeys0.ensure((size_t)cap * 8) || c->p_tops.ensure((size_t)(nrun0 * MAXSUB / 2048 + 4) * (sizeof(TileCum) + sizeof(int))))
            return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (tiles)");
        PlanBuffers pb;
        pb.slot_seg = c->p_seg.as<TileSeg>(); pb.slot_geo = c->p_geo.as<TileGeo>(); pb.slot_info = c->p_info.as<TileInfo>();
        pb.cnt = c->p_cnt.as<int>(); pb.off = c->p_off.as<int>(); pb.geo = c->geo.as<TileGeo>(); pb.desc = c->desc.as<TileDesc>();
        pb.cum = c->cum.as<TileCum>(); pb.keys0 = c->pkeys0.as<unsigned long long>(); pb.keys1 = c->pkeys1.as<unsigned long long>();
        pb.ord0 = c->pord0.as<int>(); pb.ord1 = c->pord1.as<int>(); pb.tiles = c->tiles.as<TileDesc>();
        pb.gkeys0 = c->pgkeys0.as<unsigned long long>(); pb.gtiles = c->gtiles.as<TileDesc>();
        pb.summary = c->d_summary.as<PlanSummary>(); pb.cap = (int)cap;
        pb.tops_c = c->p_tops.as<TileCum>(); pb.tops_i = reinterpret_cast<int *>(pb.tops_c + (nrun0 * MAXSUB / 2048 + 4));
        launch_plan_tiles(c->db, rsx, rsy, rsz, n, c->split_radius, rank, nranks, pool_doubles, pb, st);
        CUDA_TRY(cudaGetLastError());
        c->stats.launches += 10;
        CUDA_TRY(cudaMemcpyAsync(c->h_summary, pb.summary, sizeof(PlanSummary), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));          // the ONE host round trip of a plan
        S = *c->h_summary;
        if (!S.overflow) {
            if (S.nbatch > MAX_BATCH) return fail(GIMIC_B200_EINVAL, "point set needs more than 4096 panel-pool batches: pass fewer points per call or raise GIMIC_B200_POOL_MB");
            const int nt = S.thi - S.tlo;
            if (nt > 0) {
                const size_t sb = plan_sort_temp_bytes(nt);
                if (c->sorttmp.ensure(sb)) return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (tile sort)");
                launch_plan_order(pb, S.tlo, nt, S.drain_chunk < pool_doubles, c->sorttmp.p, sb, st);
                CUDA_TRY(cudaGetLastError());
                c->stats.launches += S.drain_chunk < pool_doubles ? 4 : 2;
            }
            break;
        }
        if (attempt == 1) return fail(GIMIC_B200_EINVAL, "tile plan overflow");
        cap = nrun0 * MAXSUB;
    }
    if (prof) cudaEventRecord(c->ev[2], st);
    c->plan.valid = true; c->plan.n = n; c->plan.rank = rank; c->plan.nranks = nranks;
    if (prof) {
        CUDA_TRY(cudaEventSynchronize(c->ev[2]));
        float m = 0;
        cudaEventElapsedTime(&m, c->ev[0], c->ev[1]); c->stats.ms_sort += m;
        cudaEventElapsedTime(&m, c->ev[1], c->ev[2]); c->stats.ms_tiles += m;
    }
    return 0;
}

// What one pass over the planned tiles writes (device pointers; any may be null).  jpath: contract with B before the GEMM
// (operands (D, P.B), jvec / jmod / edens only).
struct Outputs {
    double *tens = nullptr, *jvec = nullptr, *jmod = nullptr, *acid = nullptr, *edens = nullptr;
    const double *B3 = nullptr;       // host, needed for jvec / jmod
    bool jpath = false;
};

// exec_plan: basis panels + contraction for the planned tiles, batch by batch.  compact == false: results go to the caller's
// point order (row perm[p] of the outputs, which hold plan.n rows); compact == true: row p - pt_lo (plan.count() rows).
int exec_plan(gimic_b200_ctx *c, int spincase, const Outputs &o, bool compact, const std::function<int(int, long, long)> *after_batch = nullptr) {
    using namespace gb;
    const Plan &P = c->plan;
    if (!P.valid) return fail(GIMIC_B200_EINVAL, "no tile plan: call gimic_b200_partition_points / _grid first");
    const double *op = nullptr;
    if (int rc = o.jpath ? get_operand_j(c, spincase, o.B3, &op) : get_operand(c, spincase, &op)) return rc;
    if ((o.jvec || o.jmod) && !o.B3) return fail(GIMIC_B200_EINVAL, "jvec / jmod need the magnetic field direction");
    const PlanSummary &S = P.sum;
    const int nt = S.thi - S.tlo;
    if (nt <= 0) return 0;
    cudaStream_t st = c->stream;
    const bool prof = c->profiling, giao = c->opts.giao != 0;
    const long n = P.n;
    const double *rsx = c->rs.as<double>(), *rsy = rsx + n, *rsz = rsy + n;
    const long long pool_cap = plan_pool_doubles(c);
    const size_t pool_doubles = (size_t)std::max<long long>(2, std::min<long long>(S.panel_range, pool_cap + S.max_tile_panel));
    // index / atom-table pools by their bound relative to the panel pool: a tile holds 4*LDP panel doubles per K slot, at most
    // 2 index ints per slot (slot -> function, column -> slot) and at most one atom run per slot
    const size_t slots = pool_doubles / (4 * LDP) + 16;
    const int nbatch = S.nbatch;
    if (c->panel.ensure(pool_doubles * 8) || c->fidx.ensure(2 * slots * 4) || c->atab.ensure(slots * sizeof(TileAtom)))
        return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (panel pool)");
    if (prof) while (c->evpool.size() < 3 * (size_t)nbatch) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); c->evpool.push_back(e); }
    for (int b = 0; b < nbatch; ++b) {
        const int t0 = S.batch_start[b] - S.tlo, nb = S.batch_start[b + 1] - S.batch_start[b];
        if (nb <= 0) continue;
        if (prof) cudaEventRecord(c->evpool[3 * b], st);
        launch_basis(c->db, c->tiles.as<TileDesc>() + t0, nb, S.max_nruns, c->geo.as<TileGeo>(), rsx, rsy, rsz, c->panel.as<double>(), c->fidx.as<int>(),
                     giao ? c->atab.as<TileAtom>() : nullptr, st);
        if (prof) cudaEventRecord(c->evpool[3 * b + 1], st);
        CUDA_TRY(cudaMemsetAsync(c->misc.p, 0, 4, st));
        // Less than one tile per SM in the whole point set (a plane of an integral, a handful of points): cut the tiles into column
        // slices so that the work items fill the SMs twice over, each of about the same cost (slice_width, kernels.cuh); the slices'
        // row sums are added by k_slice_reduce.  Slot count and item cost depend only on the WHOLE set's tiles (the same on every
        // rank of a partition, so the ranks' results stay bit-identical to a single-rank run) and on the device.
        int nsl = 1;
        long long item_cost = 1;
        const char *slices_env = std::getenv("GIMIC_B200_SLICES");      // "0": never slice (tests compare the two paths)
        if (S.ntiles <= c->nsm && jtensor_supports_slices() && !(slices_env && slices_env[0] == '0')) {
            nsl = 16;
            item_cost = std::max<long long>(1, S.cost_total / (2LL * c->nsm));
        }
        if (nsl > 1 && (c->items.ensure((size_t)nb * nsl * sizeof(TileDesc)) || c->part.ensure((size_t)nb * nsl * MT * PART_LD * 8)))
            return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (tile slices)");
        if (nsl > 1) { launch_tile_slices(c->tiles.as<TileDesc>() + t0, nb, nsl, item_cost, c->items.as<TileDesc>(), st); c->stats.launches += 2; }
        JtensorArgs a;
        a.tiles = nsl > 1 ? c->items.as<TileDesc>() : c->tiles.as<TileDesc>() + t0; a.ntiles = nb * nsl; a.counter = c->misc.as<int>();
        a.part = nsl > 1 ? c->part.as<double>() : nullptr;
        a.panel_pool = c->panel.as<double>(); a.fidx_pool = c->fidx.as<int>(); a.atab_pool = c->atab.as<TileAtom>(); a.geo = c->geo.as<TileGeo>();
        a.Bop = op; a.plane_stride = c->plane_stride; a.ldb = c->ldb; a.fR = c->db.fR; a.nbf = c->hb.nbf;
        a.rsx = rsx; a.rsy = rsy; a.rsz = rsz; a.perm = compact ? nullptr : c->vals1.as<int>(); a.out_base = compact ? S.pt_lo : 0;
        a.tens = o.tens; a.edens = o.edens; a.jvec = o.jvec; a.jmod = o.jmod; a.acid = o.acid; a.jpath = o.jpath ? 1 : 0;
        for (int k = 0; k < 3; ++k) a.B[k] = o.B3 ? o.B3[k] : 0.0;
        a.paramag = c->opts.paramag; a.diamag = c->opts.diamag;
        // A caller that drains the rows to the host while the GPU works (after_batch) gets them group by group when the range has few
        // batches: a second tile list is ordered (batch, drain group, costliest first), a group is a contiguous run of Hilbert-ordered
        // tiles = of compact output rows, and each group is its own launch.  Otherwise one launch per batch over the (batch, cost) list.
        const bool grouped = after_batch && nsl == 1 && b < DRAIN_BATCHES && S.drain_chunk < pool_cap;
        if (!grouped) {
            launch_jtensor(a, giao, c->nsm, st);
            if (nsl > 1) launch_slice_reduce(a, c->tiles.as<TileDesc>() + t0, nb, nsl, item_cost, giao, st);
            CUDA_TRY(cudaGetLastError());
            c->stats.launches += 2;
            c->stats.contract_launches += 1;
            if (prof) cudaEventRecord(c->evpool[3 * b + 2], st);
            // a batch is a contiguous run of Hilbert-ordered tiles = of compact output rows: the caller may start draining them
            if (after_batch) if (int rc = (*after_batch)(b, (long)(S.batch_pt[b] - S.pt_lo), (long)(S.batch_pt[b + 1] - S.pt_lo))) return rc;
        } else {
            int done = 0;
            for (int g = 0; g < DRAIN_GROUPS; ++g) {
                const int gt = S.group_tile[b * DRAIN_GROUPS + g];
                if (gt < 0) continue;
                int next_tile = S.batch_start[b + 1]; long long next_pt = S.batch_pt[b + 1];
                for (int g2 = g + 1; g2 < DRAIN_GROUPS; ++g2)
                    if (S.group_tile[b * DRAIN_GROUPS + g2] >= 0) { next_tile = S.group_tile[b * DRAIN_GROUPS + g2]; next_pt = S.group_pt[b * DRAIN_GROUPS + g2]; break; }
                const int ng = next_tile - gt;
                if (ng <= 0) continue;
                if (done) CUDA_TRY(cudaMemsetAsync(c->misc.p, 0, 4, st));
                a.tiles = c->gtiles.as<TileDesc>() + t0 + done; a.ntiles = ng;
                launch_jtensor(a, giao, c->nsm, st);
                CUDA_TRY(cudaGetLastError());
                c->stats.launches += 1;
                c->stats.contract_launches += 1;
                done += ng;
                if (int rc = (*after_batch)(b * DRAIN_GROUPS + g, (long)(S.group_pt[b * DRAIN_GROUPS + g] - S.pt_lo), (long)(next_pt - S.pt_lo))) return rc;
            }
            c->stats.launches += 1;
            if (done != nb) return fail(GIMIC_B200_EINVAL, "internal error: drain groups do not cover the batch");
            if (prof) cudaEventRecord(c->evpool[3 * b + 2], st);
        }
    }
    if (prof) {
        CUDA_TRY(cudaStreamSynchronize(st));
        float m = 0;
        for (int b = 0; b < nbatch; ++b) {
            if (S.batch_start[b + 1] - S.batch_start[b] <= 0) continue;
            cudaEventElapsedTime(&m, c->evpool[3 * b], c->evpool[3 * b + 1]); c->stats.ms_basis += m;
            cudaEventElapsedTime(&m, c->evpool[3 * b + 1], c->evpool[3 * b + 2]); c->stats.ms_contract += m;
        }
    }
    const double tapw = giao ? (o.jpath ? 1.0 : 3.0) : 0.0, planes = o.jpath ? 2.0 : 4.0;
    c->stats.n_points += P.count(); c->stats.n_tiles += nt; c->stats.sum_nact += S.sum_nact;
    c->stats.panel_bytes += 8.0 * (double)S.panel_range;
    c->stats.executed_flops += (o.jpath ? S.flops2 : S.flops4) + tapw * S.taps;
    c->stats.useful_flops += planes * S.useful_mm + tapw * S.useful_taps;
    const double nbf = c->hb.nbf;
    c->stats.dense_flops += (double)P.count() * (c->opts.giao ? 14.0 * nbf * nbf + 56.0 * nbf : 8.0 * nbf * nbf + 20.0 * nbf);
    return 0;
}

// one plan over all n points on this device, results in the caller's point order
int run_tensors(gimic_b200_ctx *c, long n, const double *d_r, int spincase, const Outputs &o) {
    if (n <= 0) return 0;
    int sc = spincase;
    if (int rc = check_spincase(c, sc)) return rc;       // before any work is queued
    if (int rc = build_plan(c, n, d_r, 0, 1)) return rc;
    const int rc = exec_plan(c, spincase, o, false);
    c->plan.valid = false;                               // the workspaces are reused by the next call
    return rc;
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

}  // extern "C" (reopened below)

namespace {
// One pass over `n` device-resident points: tensors and/or derived fields, all outputs device pointers in the caller's order.
// divj (no reference semantics at this commit): central differences of J = T.B at r +- h e_a, 6n more points on the J path.
int fields_on_device(gimic_b200_ctx *c, long n, const double *d_r, const double *B3, int spincase, double *d_tens, double *d_jvec,
                     double *d_jmod, double *d_acid, double *d_edens, double *d_divj, double divj_h) {
    cudaStream_t st = c->stream;
    // Only J (and |J|, rho) wanted: contract with B before instead of after the GEMM -- 2 operand planes instead of 4 and one
    // tap weight per row instead of three (compute_jvectors, jfield.f90:167-184, folded into the contraction).
    Outputs o;
    o.jpath = !d_tens && !d_acid && (d_jvec || d_jmod);
    o.tens = d_tens; o.jvec = d_jvec; o.jmod = d_jmod; o.acid = d_acid; o.edens = d_edens; o.B3 = B3;
    if (d_tens || d_jvec || d_jmod || d_acid || d_edens) {
        if (!o.jpath && !d_tens && !d_acid && !d_jvec && !d_jmod) {       // edens alone: the cheapest pass that forms rho
            static const double z3[3] = {0, 0, 0};
            o.jpath = true; if (!o.B3) o.B3 = z3;
        }
        if (int rc = run_tensors(c, n, d_r, spincase, o)) return rc;
    }
    if (d_divj) {
        const double hstep = divj_h > 0 ? divj_h : 1e-3;
        if (c->shift.ensure((size_t)18 * n * 8) || c->jv6.ensure((size_t)18 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (divj)");
        double *r6 = c->shift.as<double>(), *v6 = c->jv6.as<double>();
        gb::launch_shift_points(n, d_r, hstep, r6, st);
        gimic_b200_stats keep = c->stats;
        Outputs o6; o6.jpath = true; o6.jvec = v6; o6.B3 = B3;
        if (int rc = run_tensors(c, 6 * n, r6, spincase, o6)) return rc;
        keep.launches = c->stats.launches; c->stats = keep;   // statistics describe the primary pass only
        gb::launch_divj(n, v6, hstep, d_divj, st);
        c->stats.launches += 2;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Host buffers: the points are processed in chunks of the caller's order so that the device->host copy of chunk k (copy
// stream, pinned or pageable destination) overlaps the contraction of chunk k+1; two sets of device staging buffers.
constexpr long CHUNK_POINTS = 1L << 21;
}  // namespace

extern "C" {

int gimic_b200_calc_fields(gimic_b200_handle c, long n, const double *r, const double *B3, int spincase, double *tens, double *jvec,
                           double *jmod, double *acid, double *edens, double *divj, double divj_h, int flags) {
    if (!c || (n > 0 && !r)) return fail(GIMIC_B200_EINVAL, "null argument");
    if ((jvec || jmod || divj) && !B3) return fail(GIMIC_B200_EINVAL, "jvec/jmod/divj need the magnetic field direction");
    if (n < 0) return fail(GIMIC_B200_EINVAL, "negative point count");
    CUDA_TRY(cudaSetDevice(c->device));
    reset_stats(c);
    if (n == 0) return 0;
    { int sc = spincase; if (int rc = check_spincase(c, sc)) return rc; }
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    cudaStream_t st = c->stream, cs = c->copy_stream;
    cudaEventRecord(c->ev_call[0], st);
    if (dev) {
        if (int rc = fields_on_device(c, n, r, B3, spincase, tens, jvec, jmod, acid, edens, divj, divj_h)) return rc;
    } else {
        const long chunk = std::min<long>(n, CHUNK_POINTS);
        const size_t nf = (size_t)chunk;
        Buf *rin[2] = {&c->r_in, &c->r_in2}, *tt[2] = {&c->tens_tmp, &c->tens_tmp2}, *ft[2] = {&c->f_tmp, &c->f_tmp2};
        const int nbuf = n > chunk ? 2 : 1;
        for (int k = 0; k < nbuf; ++k)
            if (rin[k]->ensure(nf * 24) || (tens && tt[k]->ensure(nf * 72)) || ft[k]->ensure(nf * 8 * 7)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (staging)");
        int it = 0;
        for (long lo = 0; lo < n; lo += chunk, ++it) {
            const long m = std::min<long>(chunk, n - lo);
            const int k = it & 1;
            if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[2 + k], 0));     // buffer k drained by the copy stream
            double *d_r = rin[k]->as<double>();
            CUDA_TRY(cudaMemcpyAsync(d_r, r + 3 * lo, (size_t)m * 24, cudaMemcpyHostToDevice, st));
            double *f = ft[k]->as<double>();
            double *d_tens = tens ? tt[k]->as<double>() : nullptr;
            double *d_jvec = jvec ? f : nullptr, *d_jmod = jmod ? f + 3 * nf : nullptr, *d_acid = acid ? f + 4 * nf : nullptr,
                   *d_edens = edens ? f + 5 * nf : nullptr, *d_divj = divj ? f + 6 * nf : nullptr;
            if (int rc = fields_on_device(c, m, d_r, B3, spincase, d_tens, d_jvec, d_jmod, d_acid, d_edens, d_divj, divj_h)) return rc;
            CUDA_TRY(cudaEventRecord(c->ev_chunk[k], st));
            CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_chunk[k], 0));
            if (tens) CUDA_TRY(cudaMemcpyAsync(tens + 9 * lo, d_tens, (size_t)m * 72, cudaMemcpyDeviceToHost, cs));
            if (jvec) CUDA_TRY(cudaMemcpyAsync(jvec + 3 * lo, d_jvec, (size_t)m * 24, cudaMemcpyDeviceToHost, cs));
            if (jmod) CUDA_TRY(cudaMemcpyAsync(jmod + lo, d_jmod, (size_t)m * 8, cudaMemcpyDeviceToHost, cs));
            if (acid) CUDA_TRY(cudaMemcpyAsync(acid + lo, d_acid, (size_t)m * 8, cudaMemcpyDeviceToHost, cs));
            if (edens) CUDA_TRY(cudaMemcpyAsync(edens + lo, d_edens, (size_t)m * 8, cudaMemcpyDeviceToHost, cs));
            if (divj) CUDA_TRY(cudaMemcpyAsync(divj + lo, d_divj, (size_t)m * 8, cudaMemcpyDeviceToHost, cs));
            CUDA_TRY(cudaEventRecord(c->ev_chunk[2 + k], cs));
        }
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[2], 0));
        if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[3], 0));
    }
    cudaEventRecord(c->ev_call[1], st);
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaStreamSynchronize(cs));
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
    reset_stats(c);
    cudaEventRecord(c->ev[0], st);
    gb::launch_fields(n, d_r, d_t, B3, d_jvec, d_jmod, d_acid, st);
    cudaEventRecord(c->ev[1], st);
    CUDA_TRY(cudaGetLastError());
    c->stats.launches = 1; c->stats.n_points = n;
    if (!dev) {
        if (jvec) CUDA_TRY(cudaMemcpyAsync(jvec, d_jvec, nf * 24, cudaMemcpyDeviceToHost, st));
        if (jmod) CUDA_TRY(cudaMemcpyAsync(jmod, d_jmod, nf * 8, cudaMemcpyDeviceToHost, st));
        if (acid) CUDA_TRY(cudaMemcpyAsync(acid, d_acid, nf * 8, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c->stats.ms_fields, c->ev[0], c->ev[1]);      // the kernel alone (HBM roofline of the field pass)
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
    { int sc = spincase; if (int rc = check_spincase(c, sc)) return rc; }
    const double *ob, *p0, *p1, *p2, *w0;
    if (int rc = grid_upload(c, g, &ob, &p0, &p1, &p2, &w0)) return rc;
    cudaStream_t st = c->stream, cs = c->copy_stream;
    cudaEventRecord(c->ev_call[0], st);
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    if (dev) {
        if (c->r_in.ensure((size_t)3 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (grid points)");
        gb::launch_grid_points(ob, p0, p1, p2, g->npts[0], g->npts[1], g->npts[2], lo, hi, c->r_in.as<double>(), st);
        c->stats.launches += 1;
        Outputs o; o.tens = tens;
        if (int rc = run_tensors(c, n, c->r_in.as<double>(), spincase, o)) return rc;
    } else {
        // host output: chunks of the flat index range, the copy of chunk k overlaps the contraction of chunk k+1 (see calc_fields)
        const long chunk = std::min<long>(n, CHUNK_POINTS);
        Buf *rin[2] = {&c->r_in, &c->r_in2}, *tt[2] = {&c->tens_tmp, &c->tens_tmp2};
        const int nbuf = n > chunk ? 2 : 1;
        for (int k = 0; k < nbuf; ++k)
            if (rin[k]->ensure((size_t)chunk * 24) || tt[k]->ensure((size_t)chunk * 72)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (staging)");
        int it = 0;
        for (long a = lo; a < hi; a += chunk, ++it) {
            const long m = std::min<long>(chunk, hi - a);
            const int k = it & 1;
            if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[2 + k], 0));
            gb::launch_grid_points(ob, p0, p1, p2, g->npts[0], g->npts[1], g->npts[2], a, a + m, rin[k]->as<double>(), st);
            c->stats.launches += 1;
            Outputs o; o.tens = tt[k]->as<double>();
            if (int rc = run_tensors(c, m, rin[k]->as<double>(), spincase, o)) return rc;
            CUDA_TRY(cudaEventRecord(c->ev_chunk[k], st));
            CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_chunk[k], 0));
            CUDA_TRY(cudaMemcpyAsync(tens + 9 * (a - lo), o.tens, (size_t)m * 72, cudaMemcpyDeviceToHost, cs));
            CUDA_TRY(cudaEventRecord(c->ev_chunk[2 + k], cs));
        }
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[2], 0));
        if (it >= 2) CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[3], 0));
    }
    cudaEventRecord(c->ev_call[1], st);
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaStreamSynchronize(cs));
    cudaEventElapsedTime(&c->stats.ms_total, c->ev_call[0], c->ev_call[1]);
    return 0;
}

// ---- cost-balanced partition (multi-GPU) -----------------------------------------------------------------------------------
// The reference splits the flat point index into equal-count contiguous slabs (schedule(), src/fgimic/parallel.F90:66-84, used
// by jfield.f90:90-104).  The cost of a point is ~ (active functions)^2, which on a planar molecule differs 5x between slabs
// near and far from the plane.  Here every rank sorts and tiles the WHOLE point set (identical on all ranks: same input, same
// deterministic kernels, integer costs), and rank r takes the run of Hilbert-ordered tiles whose cumulative cost lies in
// [r, r+1) x total / nranks.  The tiles -- hence every result bit -- are the same as in a single-rank run.
int gimic_b200_partition_points(gimic_b200_handle c, long n, const double *r, int flags, int rank, int nranks, long *count) {
    if (!c || (n > 0 && !r) || !count) return fail(GIMIC_B200_EINVAL, "null argument");
    if (n <= 0) return fail(GIMIC_B200_EINVAL, "no points");
    CUDA_TRY(cudaSetDevice(c->device));
    reset_stats(c);
    cudaEventRecord(c->ev_plan[0], c->stream);
    const double *d_r = nullptr;
    if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, flags, &d_r)) return rc;
    if (int rc = build_plan(c, n, d_r, rank, nranks)) return rc;
    cudaEventRecord(c->ev_plan[1], c->stream);
    *count = c->plan.count();
    return 0;
}

int gimic_b200_partition_grid(gimic_b200_handle c, const gimic_b200_grid *g, int rank, int nranks, long *count) {
    if (!c || !g || !count) return fail(GIMIC_B200_EINVAL, "null argument");
    const long n = (long)g->npts[0] * g->npts[1] * g->npts[2];
    CUDA_TRY(cudaSetDevice(c->device));
    reset_stats(c);
    const double *ob, *p0, *p1, *p2, *w0;
    cudaEventRecord(c->ev_plan[0], c->stream);
    if (int rc = grid_upload(c, g, &ob, &p0, &p1, &p2, &w0)) return rc;
    if (c->r_in.ensure((size_t)3 * n * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (grid points)");
    gb::launch_grid_points(ob, p0, p1, p2, g->npts[0], g->npts[1], g->npts[2], 0, n, c->r_in.as<double>(), c->stream);
    c->stats.launches += 1;
    if (int rc = build_plan(c, n, c->r_in.as<double>(), rank, nranks)) return rc;
    cudaEventRecord(c->ev_plan[1], c->stream);
    *count = c->plan.count();
    return 0;
}

int gimic_b200_partition_info(gimic_b200_handle c, long *info8) {
    if (!c || !info8) return fail(GIMIC_B200_EINVAL, "null argument");
    if (!c->plan.valid) return fail(GIMIC_B200_EINVAL, "no partition: call gimic_b200_partition_points / _grid first");
    const gb::PlanSummary &S = c->plan.sum;
    info8[0] = c->plan.n; info8[1] = c->plan.count(); info8[2] = S.ntiles; info8[3] = S.thi - S.tlo;
    info8[4] = S.cost_total; info8[5] = S.cost_range; info8[6] = S.nbatch; info8[7] = S.tlo;
    return 0;
}

int gimic_b200_partition_calc(gimic_b200_handle c, const double *B3, int spincase, long *index, double *tens, double *jvec, double *jmod,
                              double *acid, double *edens, int flags) {
    if (!c) return fail(GIMIC_B200_EINVAL, "null handle");
    if (!c->plan.valid) return fail(GIMIC_B200_EINVAL, "no partition: call gimic_b200_partition_points / _grid first");
    if ((jvec || jmod) && !B3) return fail(GIMIC_B200_EINVAL, "jvec/jmod need the magnetic field direction");
    CUDA_TRY(cudaSetDevice(c->device));
    { int sc = spincase; if (int rc = check_spincase(c, sc)) return rc; }
    const long m = c->plan.count();
    { const gimic_b200_stats keep = c->stats; reset_stats(c);      // sort / tile times of the partition call stay part of the picture
      c->stats.ms_sort = keep.ms_sort; c->stats.ms_tiles = keep.ms_tiles; c->stats.launches = keep.launches; }
    if (m == 0) return 0;
    const bool dev = (flags & GIMIC_B200_DEVICE_PTR) != 0;
    cudaStream_t st = c->stream;
    cudaEventRecord(c->ev_call[0], st);
    const size_t nf = (size_t)m;
    Outputs o;
    o.jpath = !tens && !acid && (jvec || jmod || edens);
    static const double z3[3] = {0, 0, 0};
    o.B3 = B3 ? B3 : z3;
    long *d_index = index;
    if (dev) { o.tens = tens; o.jvec = jvec; o.jmod = jmod; o.acid = acid; o.edens = edens; }
    else {
        if ((tens && c->tens_tmp.ensure(nf * 72)) || c->f_tmp.ensure(nf * 8 * 7)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (partition outputs)");
        double *f = c->f_tmp.as<double>();
        o.tens = tens ? c->tens_tmp.as<double>() : nullptr;
        o.jvec = jvec ? f : nullptr; o.jmod = jmod ? f + 3 * nf : nullptr; o.acid = acid ? f + 4 * nf : nullptr; o.edens = edens ? f + 5 * nf : nullptr;
        d_index = index ? reinterpret_cast<long *>(f + 6 * nf) : nullptr;
    }
    cudaStream_t cs = c->copy_stream;
    if (index) { gb::launch_perm_index(c->vals1.as<int>() + c->plan.sum.pt_lo, m, d_index, st); c->stats.launches += 1; }
    // host outputs: the rows of a finished panel batch are copied out on the copy stream while the next batch is contracted
    std::vector<cudaEvent_t> &evb = c->evbatch;
    const std::function<int(int, long, long)> drain = [&](int b, long lo, long hi) -> int {
        if (hi <= lo) return 0;
        while (evb.size() <= (size_t)b) { cudaEvent_t e; CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); evb.push_back(e); }
        CUDA_TRY(cudaEventRecord(evb[b], st));
        CUDA_TRY(cudaStreamWaitEvent(cs, evb[b], 0));
        const size_t k = (size_t)(hi - lo);
        if (tens) CUDA_TRY(cudaMemcpyAsync(tens + 9 * lo, o.tens + 9 * lo, k * 72, cudaMemcpyDeviceToHost, cs));
        if (jvec) CUDA_TRY(cudaMemcpyAsync(jvec + 3 * lo, o.jvec + 3 * lo, k * 24, cudaMemcpyDeviceToHost, cs));
        if (jmod) CUDA_TRY(cudaMemcpyAsync(jmod + lo, o.jmod + lo, k * 8, cudaMemcpyDeviceToHost, cs));
        if (acid) CUDA_TRY(cudaMemcpyAsync(acid + lo, o.acid + lo, k * 8, cudaMemcpyDeviceToHost, cs));
        if (edens) CUDA_TRY(cudaMemcpyAsync(edens + lo, o.edens + lo, k * 8, cudaMemcpyDeviceToHost, cs));
        return 0;
    };
    if (o.tens || o.jvec || o.jmod || o.acid || o.edens)
        if (int rc = exec_plan(c, spincase, o, true, dev ? nullptr : &drain)) return rc;
    CUDA_TRY(cudaGetLastError());
    if (!dev) {
        if (index) CUDA_TRY(cudaMemcpyAsync(index, d_index, nf * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(c->ev_chunk[2], cs));
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_chunk[2], 0));      // the call's end stamp covers the last copy
    }
    cudaEventRecord(c->ev_call[1], st);
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaStreamSynchronize(cs));
    float ms = 0; cudaEventElapsedTime(&ms, c->ev_call[0], c->ev_call[1]);
    c->stats.ms_total = ms;
    cudaEventElapsedTime(&c->stats.ms_plan, c->ev_plan[0], c->ev_plan[1]);
    cudaEventElapsedTime(&c->stats.ms_span, c->ev_plan[0], c->ev_call[1]);
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
    // current and |J| only need J = T.B: when every plane has the same field direction the contraction forms it itself (operands
    // (D, sum_b B_b P_b): half the planes, half the flops); the ACID integral needs the tensor
    bool jpath = !(what & 4);
    for (int g = 1; g < ng && jpath; ++g) for (int d = 0; d < 3; ++d) jpath = jpath && B3s[3 * g + d] == B3s[d];
    {
        Outputs o;
        if (jpath) { o.jvec = c->tens_tmp.as<double>(); o.B3 = B3s; o.jpath = true; } else o.tens = c->tens_tmp.as<double>();
        if (int rc = run_tensors(c, (long)n, d_r, spincase, o)) return rc;
    }
    for (int g = 0; g < ng; ++g) {
        const gimic_b200_grid &G = grids[g];
        const int p1 = G.npts[0], p2 = G.npts[1], nrows = (int)(rowoff[g + 1] - rowoff[g]);
        if (nrows == 0) continue;
        gb::QuadArgs q;
        q.tens = jpath ? nullptr : c->tens_tmp.as<double>() + 9 * roff[g]; q.jvec = jpath ? c->tens_tmp.as<double>() + 3 * roff[g] : nullptr;
        q.p1 = p1; q.nrows = nrows; q.r = d_r + 3 * roff[g];
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

// The same vectors through the HOT-PATH kernels: Hilbert sort, tiles, k_basis panels (what the contraction consumes), scattered
// back into the dense layout.  Exists so that tests can hold k_basis itself -- not only k_basis_dense -- to the oracle.
int gimic_b200_calc_basis_tiles(gimic_b200_handle c, long n, const double *r, double *bf, double *dr, int *tile_info3) {
    if (!c || (n > 0 && !r) || (!bf && !dr)) return fail(GIMIC_B200_EINVAL, "null argument");
    if (n <= 0) return 0;
    if (c->hb.spherical) return fail(GIMIC_B200_EINVAL, "calc_basis_tiles: cartesian contexts only (spherical=on folds the projection into the densities)");
    CUDA_TRY(cudaSetDevice(c->device));
    reset_stats(c);
    const size_t nb = (size_t)c->hb.nbf;
    const double *d_r = nullptr;
    if (int rc = stage_in(c, c->r_in, r, (size_t)3 * n, 0, &d_r)) return rc;
    if (int rc = build_plan(c, n, d_r, 0, 1)) return rc;
    c->plan.valid = false;
    const gb::PlanSummary &S = c->plan.sum;
    if (S.nbatch != 1) return fail(GIMIC_B200_EINVAL, "calc_basis_tiles: the panels of all points must fit one pool batch (pass fewer points)");
    if (c->f_tmp.ensure((size_t)n * nb * 4 * 8)) return fail(GIMIC_B200_ENOMEM, "device allocation failed (basis vectors)");
    double *d_bf = c->f_tmp.as<double>(), *d_dr = d_bf + (size_t)n * nb;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemsetAsync(d_bf, 0, (size_t)n * nb * 4 * 8, st));
    const size_t pool_doubles = (size_t)std::max<long long>(2, S.panel_range), slots = pool_doubles / (4 * gb::LDP) + 16;
    if (c->panel.ensure(pool_doubles * 8) || c->fidx.ensure(2 * slots * 4) || c->atab.ensure(slots * sizeof(gb::TileAtom)))
        return fail(GIMIC_B200_ENOMEM, "device workspace allocation failed (panel pool)");
    const int nt = S.thi - S.tlo;
    const double *rsx = c->rs.as<double>(), *rsy = rsx + n, *rsz = rsy + n;
    gb::launch_basis(c->db, c->tiles.as<gb::TileDesc>(), nt, S.max_nruns, c->geo.as<gb::TileGeo>(), rsx, rsy, rsz, c->panel.as<double>(), c->fidx.as<int>(),
                     c->opts.giao ? c->atab.as<gb::TileAtom>() : nullptr, st);
    gb::launch_panel_scatter(c->tiles.as<gb::TileDesc>(), nt, c->panel.as<double>(), c->fidx.as<int>(), c->vals1.as<int>(), c->d_f2user, (int)nb,
                             bf ? d_bf : nullptr, dr ? d_dr : nullptr, st);
    CUDA_TRY(cudaGetLastError());
    if (bf) CUDA_TRY(cudaMemcpyAsync(bf, d_bf, (size_t)n * nb * 8, cudaMemcpyDeviceToHost, st));
    if (dr) CUDA_TRY(cudaMemcpyAsync(dr, d_dr, (size_t)n * nb * 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (tile_info3) { tile_info3[0] = nt; tile_info3[1] = (int)(S.sum_nact / std::max(nt, 1)); tile_info3[2] = S.max_nruns; }
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
