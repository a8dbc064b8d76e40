// TEST DOUBLE of libgimic_b200.so -- test infrastructure, never shipped, never built in-tree.
//
// tests/test_native_driver_mock.py compiles this file into a TEMPORARY directory as "libgimic_b200.so" and puts that directory in
// front of the loader path, so that the host-side layers ABOVE the C ABI -- the native C++ driver (libgimic_b200_driver.so,
// gimic-b200) and the Python driver -- can be run end to end in a container without a GPU.  The compute entry points of
// include/gimic_b200.h are answered by the CPU oracle (oracle/libgimic_oracle.so, itself test infrastructure); the host-only entry
// points (MOL geometry, Gauss nodes, Fortran number formatting) are the product's own host_basis.cpp.
//
// What this buys: (1) the orchestration of driver.cpp (run modes, spin-case combination, file formats, report text, scan mode,
// the multi-device thread partition) is checked byte for byte against the Python driver on CPU; (2) the native driver + oracle
// reproduces the reference's golden files, i.e. the writers are pinned to the reference without a GPU in the loop.
// What it does NOT do: say anything about the CUDA kernels -- those are held to the oracle by tests/test_gpu_*.py on a B200.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../gimic_b200/csrc/host_basis.hpp"
#include "../../include/gimic_b200.h"

extern "C" {   // oracle/gimic_oracle.cpp
void go_next_spherical(int on);
void *go_create_from_files(const char *mol, const char *xdens, int uhf, int use_screening, double screening_thrs, int giao, int diamag, int paramag,
                           char *errbuf, int errlen);
void go_destroy(void *h);
int go_nbf(void *h);
int go_natoms(void *h);
void go_atom_coords(void *h, double *out);
int go_ctensor(void *h, long n, const double *r, const char *spincase, double *tens, double *edens, int nthreads);
void go_jvectors(long n, const double *tens, const double *b, double *vec);
void go_jmod_signed(long n, const double *r, const double *vec, const double *mag, double *out);
void go_acid_field(long n, const double *tens, double *out);
}

struct gimic_b200_ctx {
    void *o = nullptr;
    gimic_b200_opts opts{};
    gimic_b200_stats stats{};
    std::vector<double> part_r;      // the owned points of the last gimic_b200_partition_* call
    std::vector<long> part_index;    // and their caller indices
    bool part_valid = false;
    ~gimic_b200_ctx() { if (o) go_destroy(o); }
};

namespace {
thread_local std::string g_err;
std::mutex g_create;
int fail(int code, const std::string &m) { g_err = m; return code; }
const char *SPIN[4] = {"alpha", "beta", "total", "spindens"};

int tensors(gimic_b200_ctx *c, long n, const double *r, int spincase, double *tens, double *edens) {
    if (spincase < 0 || spincase > 3) return fail(GIMIC_B200_EINVAL, "invalid spin case");
    const int rc = go_ctensor(c->o, n, r, SPIN[spincase], tens, edens, 2);
    if (rc == -2) return fail(GIMIC_B200_ESPIN, "ctensor(): beta/spindens requested, but not open-shell system!");
    if (rc) return fail(GIMIC_B200_EINVAL, "oracle ctensor failed");
    c->stats.n_points += n; c->stats.launches += 1;
    return 0;
}

void grid_point(const gimic_b200_grid *g, long i, long j, long k, double *r) {   // gridpoint, grid.f90:498-511
    for (int d = 0; d < 3; ++d) r[d] = g->origin[d] + g->pts[0][i] * g->basv[d] + g->pts[1][j] * g->basv[3 + d] + g->pts[2][k] * g->basv[6 + d];
}
}  // namespace

extern "C" {

const char *gimic_b200_last_error(void) { return g_err.c_str(); }
const char *gimic_b200_version(void) { return "gimic-b200 TEST DOUBLE (CPU oracle backend, tests only)"; }
void gimic_b200_default_opts(gimic_b200_opts *o) {
    if (!o) return;
    o->uhf = 0; o->giao = 1; o->diamag = 1; o->paramag = 1; o->screening = 1; o->screening_thrs = 1e-6; o->device = -1; o->spherical = 0;
}
int gimic_b200_device_count(void) { return 2; }

int gimic_b200_create(gimic_b200_handle *h, const char *mol, const char *xdens, const gimic_b200_opts *opts) {
    if (!h || !mol || !xdens) return fail(GIMIC_B200_EINVAL, "null argument");
    *h = nullptr;
    gimic_b200_ctx *c = new gimic_b200_ctx();
    if (opts) c->opts = *opts; else gimic_b200_default_opts(&c->opts);
    char err[512] = "";
    {
        std::lock_guard<std::mutex> lk(g_create);
        go_next_spherical(c->opts.spherical);
        c->o = go_create_from_files(mol, xdens, c->opts.uhf, c->opts.screening, c->opts.screening_thrs, c->opts.giao, c->opts.diamag, c->opts.paramag, err, 512);
    }
    if (!c->o) { delete c; return fail(GIMIC_B200_EIO, err); }
    *h = c;
    return 0;
}
int gimic_b200_create_from_arrays(gimic_b200_handle *, int, const double *, const int *, const int *, const int *, const double *, const double *, int,
                                  const double *, const double *, int, const gimic_b200_opts *) {
    return fail(GIMIC_B200_EINVAL, "test double: create_from_arrays is not provided");
}
int gimic_b200_destroy(gimic_b200_handle h) { delete h; return 0; }
int gimic_b200_nbf(gimic_b200_handle h) { return h ? go_nbf(h->o) : fail(GIMIC_B200_EINVAL, "null handle"); }
int gimic_b200_natoms(gimic_b200_handle h) { return h ? go_natoms(h->o) : fail(GIMIC_B200_EINVAL, "null handle"); }
int gimic_b200_is_uhf(gimic_b200_handle h) { return h ? h->opts.uhf : fail(GIMIC_B200_EINVAL, "null handle"); }
int gimic_b200_atom_coords(gimic_b200_handle h, double *xyz) {
    if (!h || !xyz) return fail(GIMIC_B200_EINVAL, "null argument");
    go_atom_coords(h->o, xyz);
    return 0;
}
int gimic_b200_set_profiling(gimic_b200_handle h, int) { return h ? 0 : fail(GIMIC_B200_EINVAL, "null handle"); }
int gimic_b200_get_stats(gimic_b200_handle h, gimic_b200_stats *out) { if (!h || !out) return fail(GIMIC_B200_EINVAL, "null argument"); *out = h->stats; return 0; }

int gimic_b200_fields_from_tensors(gimic_b200_handle c, long n, const double *r, const double *tens, const double *B3, double *jvec, double *jmod,
                                   double *acid, int flags) {
    if (!c || !tens || !B3 || (jmod && !r) || flags) return fail(GIMIC_B200_EINVAL, "bad argument");
    if (n <= 0) return 0;
    std::vector<double> jv;
    double *pj = jvec;
    if (!pj && jmod) { jv.resize((size_t)3 * n); pj = jv.data(); }
    if (pj) go_jvectors(n, tens, B3, pj);
    if (jmod) go_jmod_signed(n, r, pj, B3, jmod);
    if (acid) go_acid_field(n, tens, acid);
    return 0;
}

int gimic_b200_jmod_from_jvec(gimic_b200_handle c, long n, const double *r, const double *jvec, const double *B3, double *jmod, int flags) {
    if (!c || !r || !jvec || !B3 || !jmod || flags) return fail(GIMIC_B200_EINVAL, "bad argument");
    if (n > 0) go_jmod_signed(n, r, jvec, B3, jmod);
    return 0;
}

int gimic_b200_calc_fields(gimic_b200_handle c, long n, const double *r, const double *B3, int spincase, double *tens, double *jvec, double *jmod,
                           double *acid, double *edens, double *divj, double divj_h, int flags) {
    if (!c || (n > 0 && !r) || flags) return fail(GIMIC_B200_EINVAL, "bad argument");
    if ((jvec || jmod || divj) && !B3) return fail(GIMIC_B200_EINVAL, "jvec/jmod/divj need the magnetic field direction");
    if (n < 0) return fail(GIMIC_B200_EINVAL, "negative point count");
    c->stats = gimic_b200_stats{};
    if (n == 0) return 0;
    std::vector<double> t;
    double *pt = tens;
    if (!pt) { t.resize((size_t)9 * n); pt = t.data(); }
    if (int rc = tensors(c, n, r, spincase, pt, edens)) return rc;
    if (jvec || jmod || acid)
        if (int rc = gimic_b200_fields_from_tensors(c, n, r, pt, B3 ? B3 : pt, jvec, jmod, acid, 0)) return rc;
    if (divj) {   // central differences of J = T.B, like the product
        const double hs = divj_h > 0 ? divj_h : 1e-3;
        std::vector<double> rs((size_t)3 * n), ts((size_t)9 * n), jp((size_t)3 * n), jm((size_t)3 * n);
        for (long i = 0; i < n; ++i) divj[i] = 0.0;
        for (int a = 0; a < 3; ++a) {
            for (int sgn = 0; sgn < 2; ++sgn) {
                for (long i = 0; i < 3 * n; ++i) rs[(size_t)i] = r[i];
                for (long i = 0; i < n; ++i) rs[(size_t)(3 * i + a)] += sgn ? -hs : hs;
                if (int rc = tensors(c, n, rs.data(), spincase, ts.data(), nullptr)) return rc;
                go_jvectors(n, ts.data(), B3, sgn ? jm.data() : jp.data());
            }
            for (long i = 0; i < n; ++i) divj[i] += (jp[(size_t)(3 * i + a)] - jm[(size_t)(3 * i + a)]) / (2.0 * hs);
        }
    }
    return 0;
}

int gimic_b200_calc_jtensors(gimic_b200_handle h, long n, const double *r, int spincase, double *tens, int flags) {
    if (!tens && n > 0) return fail(GIMIC_B200_EINVAL, "null argument");
    return gimic_b200_calc_fields(h, n, r, nullptr, spincase, tens, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, flags);
}

int gimic_b200_calc_jtensors_grid(gimic_b200_handle c, const gimic_b200_grid *g, long lo, long hi, int spincase, double *tens, int flags) {
    if (!c || !g || !tens || flags) return fail(GIMIC_B200_EINVAL, "bad argument");
    const long n0 = g->npts[0], n1 = g->npts[1], ntot = n0 * n1 * g->npts[2];
    if (lo < 0 || hi > ntot || lo > hi) return fail(GIMIC_B200_EINVAL, "grid index range out of bounds");
    std::vector<double> r((size_t)3 * (hi - lo));
    for (long q = lo; q < hi; ++q) grid_point(g, q % n0, (q / n0) % n1, q / (n0 * n1), &r[(size_t)3 * (q - lo)]);
    c->stats = gimic_b200_stats{};
    return hi > lo ? tensors(c, hi - lo, r.data(), spincase, tens, nullptr) : 0;
}

// The product gives a rank a run of Hilbert-ordered tiles (an arbitrary, non-contiguous set of caller indices); the test double
// hands out blocks of 7 consecutive points round-robin, so that callers which assumed contiguous slabs would fail.
static int mock_partition(gimic_b200_ctx *c, long n, const double *r, int rank, int nranks, long *count) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(GIMIC_B200_EINVAL, "rank / nranks out of range");
    c->part_r.clear(); c->part_index.clear();
    for (long i = 0; i < n; ++i)
        if ((i / 7) % nranks == rank) { c->part_index.push_back(i); for (int d = 0; d < 3; ++d) c->part_r.push_back(r[3 * i + d]); }
    c->part_valid = true;
    *count = (long)c->part_index.size();
    return 0;
}
int gimic_b200_partition_points(gimic_b200_handle c, long n, const double *r, int flags, int rank, int nranks, long *count) {
    if (!c || !r || !count || flags || n <= 0) return fail(GIMIC_B200_EINVAL, "bad argument");
    return mock_partition(c, n, r, rank, nranks, count);
}
int gimic_b200_partition_grid(gimic_b200_handle c, const gimic_b200_grid *g, int rank, int nranks, long *count) {
    if (!c || !g || !count) return fail(GIMIC_B200_EINVAL, "bad argument");
    const long n0 = g->npts[0], n1 = g->npts[1], ntot = n0 * n1 * g->npts[2];
    std::vector<double> r((size_t)3 * ntot);
    for (long q = 0; q < ntot; ++q) grid_point(g, q % n0, (q / n0) % n1, q / (n0 * n1), &r[(size_t)3 * q]);
    return mock_partition(c, ntot, r.data(), rank, nranks, count);
}
int gimic_b200_partition_info(gimic_b200_handle c, long *info8) {
    if (!c || !info8 || !c->part_valid) return fail(GIMIC_B200_EINVAL, "no partition");
    for (int i = 0; i < 8; ++i) info8[i] = 0;
    info8[1] = (long)c->part_index.size();
    return 0;
}
int gimic_b200_partition_calc(gimic_b200_handle c, const double *B3, int spincase, long *index, double *tens, double *jvec, double *jmod, double *acid,
                              double *edens, int flags) {
    if (!c || flags) return fail(GIMIC_B200_EINVAL, "bad argument");
    if (!c->part_valid) return fail(GIMIC_B200_EINVAL, "no partition: call gimic_b200_partition_points / _grid first");
    const long m = (long)c->part_index.size();
    if (index) for (long i = 0; i < m; ++i) index[i] = c->part_index[(size_t)i];
    if (m == 0) return 0;
    const std::vector<double> r = c->part_r;      // calc_fields resets nothing of the plan, but keep the inputs stable
    const int rc = gimic_b200_calc_fields(c, m, r.data(), B3, spincase, tens, jvec, jmod, acid, edens, nullptr, 0.0, 0);
    return rc;
}

int gimic_b200_calc_basis_tiles(gimic_b200_handle, long, const double *, double *, double *, int *) {
    return fail(GIMIC_B200_EINVAL, "test double: calc_basis_tiles is not provided");
}
int gimic_b200_calc_basis(gimic_b200_handle, long, const double *, double *, double *, int) {
    return fail(GIMIC_B200_EINVAL, "test double: calc_basis is not provided");
}

// integrate_current / integrate_modulus / integrate_acid (integral.f90:50-511) on rows [jlo, jhi)
int gimic_b200_integrate(gimic_b200_handle c, const gimic_b200_grid *g, const double *B3, int spincase, int what, int jlo, int jhi, double *out7) {
    if (!c || !g || !B3 || !out7) return fail(GIMIC_B200_EINVAL, "null argument");
    const long p1 = g->npts[0], p2 = g->npts[1], p3 = g->npts[2];
    if (p1 <= 0 || p2 <= 0 || p3 <= 0) return fail(GIMIC_B200_EINVAL, "grid with no points");
    if (jlo < 0 || jhi > p2 || jlo > jhi) return fail(GIMIC_B200_EINVAL, "row range out of bounds");
    for (int k = 0; k < 7; ++k) out7[k] = 0.0;
    const long nrow = jhi - jlo, n = p1 * nrow * p3;
    if (n == 0) return 0;
    std::vector<double> r((size_t)3 * n), w((size_t)n), t((size_t)9 * n);
    size_t q = 0;
    for (long k = 0; k < p3; ++k) for (long j = jlo; j < jhi; ++j) for (long i = 0; i < p1; ++i, ++q) {
        grid_point(g, i, j, k, &r[3 * q]);
        w[q] = (g->wgt[0] ? g->wgt[0][i] : 1.0) * (g->wgt[1] ? g->wgt[1][j] : 1.0) * (g->wgt[2] ? g->wgt[2][k] : 1.0);
    }
    if (int rc = tensors(c, n, r.data(), spincase, t.data(), nullptr)) return rc;
    double a[3], b[3], center[3];
    grid_point(g, p1 - 1, 0, 0, a); grid_point(g, 0, p2 - 1, 0, b);                  // grid_center, grid.f90:529-541
    for (int d = 0; d < 3; ++d) center[d] = (a[d] + b[d]) * 0.5;
    const double bound = g->radius > 0.0 ? g->radius : 1e300;
    const double *nrm = &g->basv[6];
    for (long p = 0; p < n; ++p) {
        const double *rr = &r[(size_t)3 * p], *T = &t[(size_t)9 * p];
        const double dist = std::sqrt((rr[0] - center[0]) * (rr[0] - center[0]) + (rr[1] - center[1]) * (rr[1] - center[1]) + (rr[2] - center[2]) * (rr[2] - center[2]));
        double J[3];
        go_jvectors(1, T, B3, J);
        double jp = nrm[0] * J[0] + nrm[1] * J[1] + nrm[2] * J[2];
        double jm = std::sqrt(J[0] * J[0] + J[1] * J[1] + J[2] * J[2]);
        double ac = 0.0;
        if (what & 4) go_acid_field(1, T, &ac);
        if (dist > bound) { jp = 0.0; jm = 0.0; ac = 0.0; }
        if (what & 1) { out7[0] += w[(size_t)p] * jp; if (jp > 0.0) out7[1] += w[(size_t)p] * jp; else out7[2] += w[(size_t)p] * jp; }
        if (what & 2) {
            const double s = std::fabs(jp) < 1e-12 ? 0.0 : (jp > 0.0 ? jm : -jm);
            out7[3] += w[(size_t)p] * s; if (s > 0.0) out7[4] += w[(size_t)p] * s; else out7[5] += w[(size_t)p] * s;
        }
        if (what & 4) out7[6] += w[(size_t)p] * ac;
    }
    return 0;
}

int gimic_b200_integrate_batch(gimic_b200_handle c, int ngrids, const gimic_b200_grid *grids, const double *B3s, int spincase, int what, double *out7s) {
    if (!c || !grids || !B3s || !out7s || ngrids < 0) return fail(GIMIC_B200_EINVAL, "null argument");
    for (int g = 0; g < ngrids; ++g)
        if (int rc = gimic_b200_integrate(c, &grids[g], &B3s[3 * g], spincase, what, 0, grids[g].npts[1], &out7s[7 * g])) return rc;
    return 0;
}

// get_property integrands (jfield.f90:689-760, 820-905), as documented in include/gimic_b200.h
static void integrands(const double *r, const double *T, const double *centre, double *f3) {
    const double Jb[3][3] = {{-T[0], -T[1], -T[2]}, {-T[3], -T[4], -T[5]}, {-T[6], -T[7], -T[8]}};   // J_b = T.(-e_b): column b of T, negated
    double d[3] = {r[0], r[1], r[2]}, f = 0.5;
    if (centre) {
        for (int k = 0; k < 3; ++k) d[k] = r[k] - centre[k];
        const double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        f = 1.0e6 * (-1.0 / (d2 * std::sqrt(d2)) / (137.0359998 * 137.0359998));
    }
    f3[0] = f * (d[1] * Jb[0][2] - d[2] * Jb[0][1]);
    f3[1] = f * (d[2] * Jb[1][0] - d[0] * Jb[1][2]);
    f3[2] = f * (d[0] * Jb[2][1] - d[1] * Jb[2][0]);
}

int gimic_b200_property(gimic_b200_handle c, long n, const double *r, const double *w, const double *tens, int natoms, const double *coords, int nseg,
                        const long *seg_end, double *part, int flags) {
    if (!c || !r || !w || !tens || !coords || !seg_end || !part || natoms < 0 || nseg <= 0 || n < 0 || flags) return fail(GIMIC_B200_EINVAL, "bad argument");
    if (seg_end[nseg - 1] != n) return fail(GIMIC_B200_EINVAL, "segment ends must be cumulative and finish at n");
    for (long i = 0; i < (long)(natoms + 1) * nseg * 5; ++i) part[i] = 0.0;
    for (int k = 0; k <= natoms; ++k) {
        long p = 0;
        for (int s = 0; s < nseg; ++s)
            for (; p < seg_end[s]; ++p) {
                double f3[3];
                integrands(&r[3 * p], &tens[9 * p], k < natoms ? &coords[3 * k] : nullptr, f3);
                double *o = &part[((size_t)k * nseg + s) * 5];
                const double tot = w[p] * (f3[0] + f3[1] + f3[2]);
                for (int q = 0; q < 3; ++q) o[q] += w[p] * f3[q];
                if (tot > 0.0) o[3] += tot; else o[4] += tot;
            }
    }
    return 0;
}

int gimic_b200_property_integrand(gimic_b200_handle c, long n, const double *r, const double *tens, const double *centre3, double *out4, int flags) {
    if (!c || !r || !tens || !out4 || flags) return fail(GIMIC_B200_EINVAL, "bad argument");
    for (long p = 0; p < n; ++p) {
        integrands(&r[3 * p], &tens[9 * p], centre3, &out4[4 * p]);
        out4[4 * p + 3] = out4[4 * p] + out4[4 * p + 1] + out4[4 * p + 2];
    }
    return 0;
}

// ---- host-only entry points: the product's own code (gimic_b200/csrc/host_basis.cpp) ----------------------------------------
int gimic_b200_gauss_points(double a, double b, int npts, int order, int quadrature, double *pts, double *wgts) {
    if (!pts || !wgts || npts <= 0) return fail(GIMIC_B200_EINVAL, "bad argument");
    const int rc = gb::gauss_blocks(a, b, npts, order, quadrature, pts, wgts);
    if (rc == -1) return fail(GIMIC_B200_EINVAL, "*** integration did not converge!");
    if (rc) return fail(GIMIC_B200_EINVAL, "gaussgrid(): npts is not dividable by ngp!");
    return 0;
}
int gimic_b200_convert_xdens(const char *, int, int, const char *) { return fail(GIMIC_B200_EINVAL, "test double: convert_xdens is not provided"); }
long gimic_b200_format_e(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap) {
    if (n < 0 || (n > 0 && (!v || !out)) || w <= 0 || w > 40 || d <= 0 || d > 30) { fail(GIMIC_B200_EINVAL, "bad argument"); return GIMIC_B200_EINVAL; }
    const long rc = gb::format_fortran_e(n, v, w, d, per_line, first_count, prefix, out, cap);
    return rc < 0 ? (long)fail(GIMIC_B200_EINVAL, "output buffer too small") : rc;
}
long gimic_b200_format_f(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap) {
    if (n < 0 || (n > 0 && (!v || !out)) || w <= 0 || w > 40 || d < 0 || d > 30) { fail(GIMIC_B200_EINVAL, "bad argument"); return GIMIC_B200_EINVAL; }
    const long rc = gb::format_fortran(n, v, 'F', w, d, per_line, first_count, prefix, out, cap);
    return rc < 0 ? (long)fail(GIMIC_B200_EINVAL, "output buffer too small") : rc;
}
int gimic_b200_mol_geometry(const char *mol, int max_atoms, double *xyz, char *symbols2) {
    if (!mol) return fail(GIMIC_B200_EINVAL, "null argument");
    gb::HostBasis hb; std::string err;
    if (!gb::parse_mol(mol, hb, err)) return fail(GIMIC_B200_EIO, err);
    for (int a = 0; a < hb.natoms && a < max_atoms; ++a) {
        if (xyz) for (int k = 0; k < 3; ++k) xyz[3 * a + k] = hb.xyz[3 * a + k];
        if (symbols2) { symbols2[2 * a] = hb.symbol[a].size() > 0 ? hb.symbol[a][0] : ' '; symbols2[2 * a + 1] = hb.symbol[a].size() > 1 ? hb.symbol[a][1] : ' '; }
    }
    return hb.natoms;
}
int gimic_b200_mol_summary(const char *mol, int *info5) {
    if (!mol || !info5) return fail(GIMIC_B200_EINVAL, "null argument");
    gb::HostBasis hb; std::string err;
    if (!gb::parse_mol(mol, hb, err)) return fail(GIMIC_B200_EIO, err);
    info5[0] = hb.natoms; info5[1] = hb.ngto; info5[2] = hb.nbf; info5[3] = hb.turbomole ? 1 : 0; info5[4] = hb.nbf_sph;
    return 0;
}
int gimic_b200_c2s_rows(int l, int turbomole_order, double *po) {
    if (!po || l < 0 || l > gb::MAX_L) return fail(GIMIC_B200_EINVAL, "bad argument");
    std::vector<double> rows;
    gb::c2s_rows(l, turbomole_order != 0, rows);
    std::copy(rows.begin(), rows.end(), po);
    return 0;
}

// legacy symbols: present so that the loader of gimic_b200/_lib.py finds them; not exercised through the test double
void gimic_init(const char *, const char *) {}
void gimic_finalize(void) {}
void gimic_set_uhf(int *) {}
void gimic_set_magnet(const double *) {}
void gimic_set_spin(const char *) {}
void gimic_set_screening(const double *) {}
void gimic_calc_jtensor(const double *, double *) {}
void gimic_calc_jvector(const double *, double *) {}
void gimic_calc_modj(const double *, double *) {}
void gimic_get_gauss_points(double *a, double *b, int *npts, int *order, double *pts, double *wgts) { gimic_b200_gauss_points(*a, *b, *npts, *order, 0, pts, wgts); }
void mkgausspoints(double *a, double *b, int *npts, int *order, double *pts, double *wgts) { gimic_get_gauss_points(a, b, npts, order, pts, wgts); }

}  // extern "C"
