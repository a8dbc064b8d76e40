// Run modes: the compiled counterpart of `program gimic` (src/fgimic/gimic.F90:60-261: initialize, driver, run_cdens,
// run_integral), jvector_plots (src/fgimic/jfield.f90:250-443), the report printing of src/fgimic/integral.f90:167-183,
// 306-322,502-510 and of get_property (jfield.f90:584-929).  All arithmetic on the path happens behind the C ABI
// (include/gimic_b200.h); this file only orchestrates calls and lays out text.
#include <cctype>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <ctime>
#include <fstream>
#include <memory>
#include <sstream>
#include <thread>
#include <sys/resource.h>
#include <sys/stat.h>

#include "../../../include/gimic_b200_driver.h"
#include "native_driver.hpp"

namespace gbd {

namespace {

thread_local std::string g_error;

void check(int rc) { if (rc < 0) throw DriverError(gimic_b200_last_error()); }

struct Context {
    gimic_b200_handle h = nullptr;
    std::string key;
    ~Context() { if (h) gimic_b200_destroy(h); }
};

const char *SPIN_LABEL[4] = {"alpha", "beta", "total", "spin"};

// printf writes a NaN as "nan" or "-nan"; gfortran's F and E edits write "NaN" (and never a sign).  Same width, so columns stay put.
void gfortran_nan(char *s) {
    for (char *p = s; (p = std::strstr(p, "nan")) != nullptr; p += 3) {
        const bool word = (p > s && std::isalpha((unsigned char)p[-1])) || std::isalpha((unsigned char)p[3]);
        if (word) continue;
        p[0] = 'N'; p[2] = 'N';
        if (p > s && p[-1] == '-') p[-1] = ' ';
    }
}

struct Printer {
    FILE *f;
    void raw(const std::string &s) const { std::fwrite(s.data(), 1, s.size(), f); }
    void say(const std::string &s = std::string()) const { raw(s.empty() ? std::string("\n") : " " + s + "\n"); }   // msg_out lines
    void pf(const char *fmtstr, ...) const __attribute__((format(printf, 2, 3))) {
        char buf[512];
        va_list ap;
        va_start(ap, fmtstr);
        vsnprintf(buf, sizeof buf, fmtstr, ap);
        va_end(ap);
        gfortran_nan(buf);
        raw(buf);
    }
};

std::string sfmt(const char *fmtstr, ...) __attribute__((format(printf, 1, 2)));
std::string sfmt(const char *fmtstr, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmtstr);
    vsnprintf(buf, sizeof buf, fmtstr, ap);
    va_end(ap);
    gfortran_nan(buf);
    return buf;
}

std::string real_path(const std::string &p) {
    char buf[PATH_MAX];
    return ::realpath(p.c_str(), buf) ? std::string(buf) : p;
}

using Sums = std::array<double, 7>;

class Run {
  public:
    Input inp;
    std::string workdir;
    RunOptions opt;
    Printer out;
    bool uhf = false;
    std::shared_ptr<Context> ctx;                   // primary context (devices[0])
    std::vector<std::shared_ptr<Context>> peers;    // contexts on the other listed devices (multi-device run)
    std::string context_key;
    std::vector<std::string> symbols;
    std::vector<double> xyz;
    GridSpec grid;
    Vec3 magnet{{0, 0, 0}};
    std::vector<std::string> magnet_log;            // what get_magnet prints every time it is called (magnet.f90:66-86)
    int summary[5] = {0, 0, 0, 0, 0};               // natoms, primitive GTOs, contracted GTOs, TURBOMOLE flag, spherical count (gimic_b200_mol_summary)
    std::map<int, Sums> results;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();   // stockas_klocka reports the times of the whole run
    double cpu0[2] = {0, 0};

    static void cpu_times(double *us) {
        struct rusage ru;
        getrusage(RUSAGE_SELF, &ru);
        us[0] = (double)ru.ru_utime.tv_sec + 1e-6 * (double)ru.ru_utime.tv_usec;
        us[1] = (double)ru.ru_stime.tv_sec + 1e-6 * (double)ru.ru_stime.tv_usec;
    }

    // `find_shared`: returns an existing context for a key (scan mode) or nullptr
    template <class Finder>
    Run(const std::string &inpfile, const RunOptions &o, FILE *report, Finder find_shared) : opt(o), out{report} {
        cpu_times(cpu0);
        workdir = o.workdir.empty() ? dirname_of(inpfile) : o.workdir;
        inp = parse_file(inpfile);
        if (o.dryrun) inp.force_flag("dryrun", true);                 // the -y switch overrides the keyword (src/gimic.in:139-140)
        if (!o.title.empty()) inp.force_str("title", o.title);
        uhf = inp.flag("openshell");
        const std::string mol = path(inp.str("basis")), xdens = path(inp.str("xdens"));
        context_key = real_path(mol) + "\n" + real_path(xdens) +
                      sfmt("\n%d%d%d%d%d%d %.17g", (int)uhf, (int)inp.flag("Advanced.GIAO"), (int)inp.flag("Advanced.diamag"),
                           (int)inp.flag("Advanced.paramag"), (int)inp.flag("Advanced.screening"), (int)inp.flag("Advanced.spherical"),
                           inp.real("Advanced.screening_thrs"));
        // geometry from the library's own MOL reader (host only); a dry run needs nothing else (gimic.F90:142-159)
        const int natoms = gimic_b200_mol_geometry(mol.c_str(), 0, nullptr, nullptr);
        check(natoms);
        xyz.assign((size_t)natoms * 3, 0.0);
        std::string sym((size_t)natoms * 2, ' ');
        check(gimic_b200_mol_geometry(mol.c_str(), natoms, xyz.data(), &sym[0]));
        for (int a = 0; a < natoms; ++a) symbols.push_back(sym.substr(2 * (size_t)a, 2));
        check(gimic_b200_mol_summary(mol.c_str(), summary));
        if (!inp.flag("dryrun")) {
            ctx = find_shared(context_key);
            gimic_b200_opts go;
            gimic_b200_default_opts(&go);
            go.uhf = uhf; go.giao = inp.flag("Advanced.GIAO"); go.diamag = inp.flag("Advanced.diamag");
            go.paramag = inp.flag("Advanced.paramag"); go.screening = inp.flag("Advanced.screening");
            go.screening_thrs = inp.real("Advanced.screening_thrs"); go.device = o.devices.empty() ? o.device : o.devices[0];
            go.spherical = inp.flag("Advanced.spherical");
            if (!ctx) {
                ctx = std::make_shared<Context>();
                ctx->key = context_key;
                if (o.devices.size() > 1) {
                    // one context per listed GPU, created concurrently (each reads MOL / XDENS itself: use the binary XDENS cache
                    // for large cases); densities are replicated, nothing is exchanged between devices afterwards
                    peers.resize(o.devices.size() - 1);
                    std::vector<std::string> errs(o.devices.size());
                    std::vector<std::thread> th;
                    for (size_t d = 0; d < o.devices.size(); ++d)
                        th.emplace_back([&, d] {
                            gimic_b200_opts gd = go;
                            gd.device = o.devices[d];
                            auto c = d == 0 ? ctx : (peers[d - 1] = std::make_shared<Context>());
                            c->key = context_key;
                            if (gimic_b200_create(&c->h, mol.c_str(), xdens.c_str(), &gd) < 0) errs[d] = gimic_b200_last_error();
                        });
                    for (auto &t : th) t.join();
                    for (const auto &e : errs) if (!e.empty()) throw DriverError(e);
                } else {
                    check(gimic_b200_create(&ctx->h, mol.c_str(), xdens.c_str(), &go));
                }
            }
            check(gimic_b200_atom_coords(ctx->h, xyz.data()));
        }
        grid = grid_from_input(inp, xyz, workdir);
        magnet = get_magnet(grid, inp.str("magnet_axis"), inp.vec3("magnet"), &magnet_log);
    }

    std::string path(const std::string &name) const { return (!name.empty() && name[0] == '/') ? name : join_path(workdir, name); }

    void integral_cases(std::vector<int> &cases, int &what) const {
        cases = {GIMIC_B200_TOTAL};
        if (uhf) { cases.push_back(GIMIC_B200_ALPHA); cases.push_back(GIMIC_B200_BETA); cases.push_back(GIMIC_B200_SPINDENS); }
        what = 1 | (inp.flag("Essential.jmod") ? 2 : 0) | (inp.flag("Essential.acid") ? 4 : 0);
    }

    void field_line() const {                       // integrate_* call get_magnet again (integral.f90:85,225)
        for (const std::string &line : magnet_log) out.raw(line + "\n");
    }

    void run(const std::map<int, Sums> *pre = nullptr) {
        run_body(pre);
        // finalize() + stockas_klocka (gimic.F90:43-52,134-138; grid.f90:453; basis.f90:348-349; timer.f90:13-43).  The jobscripts
        // recognise a finished slice by the word "wall" in gimic.N.out (jobscripts/src/current-profile-local-submit:52).
        double cpu[2];
        cpu_times(cpu);
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        out.say("*** Deallocated grid data");
        out.say("INFO: Deallocated basis set and atom data");
        out.say();
        out.say(std::string(70, '-'));
        out.say(sfmt("   wall time:%9.2fsec (%6.1f h )", wall, wall / 3600.0));
        out.say(sfmt("        user:%9.2fsec (%6.1f h )", cpu[0] - cpu0[0], (cpu[0] - cpu0[0]) / 3600.0));
        out.say(sfmt("         sys:%9.2fsec (%6.1f h )", cpu[1] - cpu0[1], (cpu[1] - cpu0[1]) / 3600.0));
        out.say(std::string(70, '-'));
        out.say(fdate());
        out.say("Hello World! (tm)");
        out.say();
        out.say("done.");
        out.say();
    }

    static std::string fdate() {
        char date[64];
        const std::time_t now = std::time(nullptr);
        std::tm tmv;
        localtime_r(&now, &tmv);
        std::strftime(date, sizeof date, "%a %b %e %H:%M:%S %Y", &tmv);
        return date;
    }

    void run_body(const std::map<int, Sums> *pre) {
        // initialize(), gimic.F90:107-131
        out.say();
        out.say(fdate());
        std::string title = inp.str("title");
        while (!title.empty() && std::isspace((unsigned char)title.front())) title.erase(title.begin());
        while (!title.empty() && std::isspace((unsigned char)title.back())) title.pop_back();
        out.say(title.empty() ? std::string(" TITLE:") : " TITLE: " + title);      // msg_out trims trailing blanks
        out.say();
        if (!inp.flag("Advanced.GIAO")) { out.say("INFO: GIAOs not used!"); out.say(); }
        if (!inp.flag("Advanced.diamag")) { out.say("INFO: Diamagnetic contributions not calculated!"); out.say(); }
        if (!inp.flag("Advanced.paramag")) { out.say("INFO: Paramagnetic contributions not calculated!"); out.say(); }
        if (!inp.flag("Advanced.diamag") && !inp.flag("Advanced.paramag")) {
            out.say("    ...this does not make sense..."); out.say();
            throw DriverError("neither diamagnetic nor paramagnetic contributions requested: nothing to calculate (gimic.F90:124-130)");
        }
        // driver(), gimic.F90:141-165: what new_basis (intgrl.f90:48-60, basis.f90:44-80), read_dens (dens.f90:94-103), new_grid and
        // plot_grid_xyz (grid.f90:608-609) print on the way
        if (summary[3]) { out.say("INFO: Detected TURBOMOLE input"); out.say(); }
        out.say(sfmt("Number of atoms =%4d", summary[0])); out.say();
        out.say("Normalizing basis"); out.say();
        out.say(sfmt("  Total number of primitive  GTO's %6d", summary[1]));
        out.say(sfmt("  Total number of contracted GTO's %6d", summary[2])); out.say();
        if (inp.flag("Advanced.screening") && inp.real("Advanced.screening_thrs") > 0.0) {
            out.say("*** Calculating screening coefficients");
            out.say("INFO: Screening threshold: " + fortran_e(inp.real("Advanced.screening_thrs"), 12, 4)); out.say();
        } else {
            out.say("INFO: Screening is not used");
        }
        if (!inp.flag("dryrun")) {
            if (uhf) out.say("INFO: scaling perturbed densities by 0.d5");
            if (summary[3]) out.say("INFO: Reordering densities [TURBOMOLE]");
        }
        for (const std::string &line : grid.log) out.raw(line + "\n");
        if (root()) write_mol_xyz(join_path(workdir, "mol.xyz"), symbols, xyz);
        if (root()) write_grid_xyz(join_path(workdir, "grid.xyz"), grid, symbols, xyz);
        out.say("*** Grid plot in grid.xyz");
        field_line();
        out.say(std::string("INFO: ") + (uhf ? "Open-shell calculation" : "Closed-shell calculation"));
        out.say();
        const std::string calc = inp.str("calc");
        if (inp.flag("dryrun")) {
            // gimic.F90:174-185,196-204,222-230: the note, then the run mode's banner, then return before any arithmetic
            out.say("*** Dry run, not calculating ...");
            out.say();
            if (calc == "cdens") { out.say("Calculating current density"); out.say("*****************************************"); }
            else if (calc == "integral") { out.say("Integrating current density"); out.say("*****************************************"); }
            return;
        }
        if (calc == "cdens") run_cdens();
        else if (calc == "integral") run_integral(pre);
        else if (calc == "edens" || calc == "divj") run_scalar(calc);
    }

  private:
    // Runs fn(handle, device index, lo, hi) for one contiguous slab of range(n) per device, each on its own host thread (the C ABI is thread-safe
    // per handle, errors are thread-local).  Slabs are the block partition of schedule() (parallel.F90:66-84); outputs go to
    // disjoint parts of caller-owned host arrays, so the devices exchange nothing.
    // One process per GPU (opt.nranks > 1): this rank's slab only, as device 0; the caller completes the result with a collective.
    template <class Fn>
    void over_devices(long n, Fn fn, long *my_lo = nullptr, long *my_hi = nullptr) const {
        const size_t nd = 1 + peers.size();
        if (my_lo) { *my_lo = 0; *my_hi = n; }
        if (ranked()) {
            const long base = n / opt.nranks, rem = n % opt.nranks;
            const long lo = opt.rank * base + std::min<long>(opt.rank, rem), hi = lo + base + (opt.rank < rem ? 1 : 0);
            if (my_lo) { *my_lo = lo; *my_hi = hi; }
            if (hi > lo) check(fn(ctx->h, (size_t)0, lo, hi));
            return;
        }
        if (nd == 1) { check(fn(ctx->h, (size_t)0, 0L, n)); return; }
        std::vector<std::string> errs(nd);
        std::vector<std::thread> th;
        for (size_t d = 0; d < nd; ++d) {
            const long base = n / (long)nd, rem = n % (long)nd;
            const long lo = (long)d * base + std::min<long>((long)d, rem), hi = lo + base + ((long)d < rem ? 1 : 0);
            if (hi <= lo) continue;
            gimic_b200_handle h = d == 0 ? ctx->h : peers[d - 1]->h;
            th.emplace_back([&errs, fn, h, lo, hi, d] { if (fn(h, d, lo, hi) < 0) errs[d] = gimic_b200_last_error(); });
        }
        for (auto &t : th) t.join();
        for (const auto &e : errs) if (!e.empty()) throw DriverError(e);
    }

    // Several devices: every device tiles the whole point set and evaluates its equal-COST share of the tiles (gimic_b200_partition_*;
    // the equal-count slabs of schedule() leave the devices far from the molecule idle on a planar system); the rows come back with
    // their point indices and are scattered into the caller's arrays -- disjoint index sets, so the host threads exchange nothing.
    bool ranked() const { return opt.nranks > 1; }
    bool root() const { return opt.rank == 0; }
    // rows [index[i]] of a [n][ncols] array are this rank's: complete `full` on every rank
    void gather_rows(long n, int ncols, const std::vector<long> &index, const std::vector<double> &rows, double *full) const {
        if (!opt.allgather_rows) throw DriverError("a run over several processes needs the allgather_rows callback (gimic_b200_run_opts)");
        if (opt.allgather_rows(opt.user, n, ncols, (long)index.size(), index.data(), rows.data(), full) < 0)
            throw DriverError("the launcher's allgather_rows callback failed");
    }
    void partitioned(const double *rp, long n, const gimic_b200_grid *g, int spincase, double *tens, double *jvec, double *jmod, double *edens) const {
        const size_t nd = 1 + peers.size();
        const double *B = magnet.data();
        if (ranked()) {
            long cnt = 0;
            check(g ? gimic_b200_partition_grid(ctx->h, g, opt.rank, opt.nranks, &cnt) : gimic_b200_partition_points(ctx->h, n, rp, 0, opt.rank, opt.nranks, &cnt));
            std::vector<long> idx((size_t)cnt);
            std::vector<double> t(tens ? (size_t)cnt * 9 : 0), v(jvec ? (size_t)cnt * 3 : 0), m(jmod ? (size_t)cnt : 0), e(edens ? (size_t)cnt : 0);
            if (cnt > 0)
                check(gimic_b200_partition_calc(ctx->h, B, spincase, idx.data(), tens ? t.data() : nullptr, jvec ? v.data() : nullptr, jmod ? m.data() : nullptr,
                                                nullptr, edens ? e.data() : nullptr, 0));
            if (tens) gather_rows(n, 9, idx, t, tens);
            if (jvec) gather_rows(n, 3, idx, v, jvec);
            if (jmod) gather_rows(n, 1, idx, m, jmod);
            if (edens) gather_rows(n, 1, idx, e, edens);
            return;
        }
        std::vector<std::string> errs(nd);
        std::vector<std::thread> th;
        for (size_t d = 0; d < nd; ++d) {
            gimic_b200_handle h = d == 0 ? ctx->h : peers[d - 1]->h;
            th.emplace_back([=, &errs] {
                long cnt = 0;
                const int rc = g ? gimic_b200_partition_grid(h, g, (int)d, (int)nd, &cnt) : gimic_b200_partition_points(h, n, rp, 0, (int)d, (int)nd, &cnt);
                if (rc < 0) { errs[d] = gimic_b200_last_error(); return; }
                if (cnt == 0) return;
                std::vector<long> idx((size_t)cnt);
                std::vector<double> t(tens ? (size_t)cnt * 9 : 0), v(jvec ? (size_t)cnt * 3 : 0), m(jmod ? (size_t)cnt : 0), e(edens ? (size_t)cnt : 0);
                if (gimic_b200_partition_calc(h, B, spincase, idx.data(), tens ? t.data() : nullptr, jvec ? v.data() : nullptr, jmod ? m.data() : nullptr,
                                              nullptr, edens ? e.data() : nullptr, 0) < 0) { errs[d] = gimic_b200_last_error(); return; }
                for (long i = 0; i < cnt; ++i) {
                    const size_t o = (size_t)idx[(size_t)i];
                    if (tens) for (int k = 0; k < 9; ++k) tens[9 * o + k] = t[(size_t)i * 9 + k];
                    if (jvec) for (int k = 0; k < 3; ++k) jvec[3 * o + k] = v[(size_t)i * 3 + k];
                    if (jmod) jmod[o] = m[(size_t)i];
                    if (edens) edens[o] = e[(size_t)i];
                }
            });
        }
        for (auto &t : th) t.join();
        for (const auto &e : errs) if (!e.empty()) throw DriverError(e);
    }

    // calc_jtensors (jfield.f90:62-138) on the whole grid
    std::vector<double> tensors(int spincase) const {
        const long n = grid.n();
        std::vector<double> t((size_t)n * 9);
        double *tp = t.data();
        if (grid.is_file()) {
            const double *xp = grid.xdata.data();
            if (peers.empty() && !ranked()) check(gimic_b200_calc_jtensors(ctx->h, n, xp, spincase, tp, 0));
            else partitioned(xp, n, nullptr, spincase, tp, nullptr, nullptr, nullptr);
        } else {
            const gimic_b200_grid g = grid.cstruct();
            if (peers.empty() && !ranked()) check(gimic_b200_calc_jtensors_grid(ctx->h, &g, 0, n, spincase, tp, 0));
            else partitioned(nullptr, n, &g, spincase, tp, nullptr, nullptr, nullptr);
        }
        return t;
    }

    // J (and signed |J|, rho, div J) straight from the contraction; div J (central differences, 6 more passes) keeps the slab split
    void point_fields(const std::vector<double> &r, int spincase, double *jvec, double *jmod, double *edens, double *divj) const {
        const double *rp = r.data(), *B = magnet.data();
        if ((!peers.empty() || ranked()) && !divj) { partitioned(rp, (long)r.size() / 3, nullptr, spincase, nullptr, jvec, jmod, edens); return; }
        const long n = (long)r.size() / 3;
        long lo = 0, hi = n;
        over_devices(n, [=](gimic_b200_handle h, size_t, long lo, long hi) {
            return gimic_b200_calc_fields(h, hi - lo, rp + 3 * lo, B, spincase, nullptr, jvec ? jvec + 3 * lo : nullptr, jmod ? jmod + lo : nullptr, nullptr,
                                          edens ? edens + lo : nullptr, divj ? divj + lo : nullptr, 1e-3, 0);
        }, &lo, &hi);
        if (ranked()) {            // every rank filled its slab of the arrays in place: exchange the slabs
            std::vector<long> idx((size_t)(hi - lo));
            for (long i = lo; i < hi; ++i) idx[(size_t)(i - lo)] = i;
            auto slab = [&](double *a, int nc) { if (a) { const std::vector<double> mine(a + nc * lo, a + nc * hi); gather_rows(n, nc, idx, mine, a); } };
            slab(jvec, 3); slab(jmod, 1); slab(edens, 1); slab(divj, 1);
        }
    }

    static std::vector<double> combine(const std::vector<double> &a, const std::vector<double> &b, double sign) {
        std::vector<double> o(a.size());
        for (size_t i = 0; i < a.size(); ++i) o[i] = a[i] + sign * b[i];
        return o;
    }

    // run_cdens (gimic.F90:196-220) + jvector_plots (jfield.f90:250-443)
    void run_cdens() {
        out.say("Calculating current density");
        out.say("*****************************************");
        std::vector<std::pair<int, std::string>> cases = {{GIMIC_B200_TOTAL, ""}};
        if (uhf) { cases.push_back({GIMIC_B200_ALPHA, "alpha"}); cases.push_back({GIMIC_B200_BETA, "beta"}); cases.push_back({GIMIC_B200_SPINDENS, "spindens"}); }
        const bool want_jmod = inp.flag("Essential.jmod") && grid.is_3d();
        const bool want_acid = inp.flag("Essential.acid") && grid.is_3d();
        const bool prop = inp.flag("Essential.prop");
        // Only J (and |J|) is written when neither ACID nor the property quadrature is asked for: the library then contracts
        // with B inside the GEMM (2 operand planes instead of 4) and never forms the tensors.
        const bool j_only = !want_acid && !prop;
        const std::vector<double> r = grid.points();
        const long n = grid.n();
        std::map<int, std::vector<double>> cache;          // tensors (or J vectors on the J path) per spin case
        if (uhf) {
            // linear in the densities: alpha and beta are evaluated once, total = alpha + beta and spindens = alpha - beta
            // exactly as ctensor combines them (jtensor.F90:86-99); the reference re-evaluates everything per spin case
            for (int sc : {GIMIC_B200_ALPHA, GIMIC_B200_BETA}) {
                if (j_only) {
                    cache[sc].assign((size_t)n * 3, 0.0);
                    point_fields(r, sc, cache[sc].data(), nullptr, nullptr, nullptr);
                } else {
                    cache[sc] = tensors(sc);
                }
            }
            cache[GIMIC_B200_TOTAL] = combine(cache[GIMIC_B200_ALPHA], cache[GIMIC_B200_BETA], 1.0);
            cache[GIMIC_B200_SPINDENS] = combine(cache[GIMIC_B200_ALPHA], cache[GIMIC_B200_BETA], -1.0);
        }
        for (const auto &cs : cases) {
            const int sc = cs.first;
            const std::string &tag = cs.second;
            std::vector<double> tens, jv, jmod, acid;
            if (j_only) {
                if (want_jmod) jmod.assign((size_t)n, 0.0);
                if (uhf) {
                    jv = cache[sc];
                    if (want_jmod) check(gimic_b200_jmod_from_jvec(ctx->h, n, r.data(), jv.data(), magnet.data(), jmod.data(), 0));
                } else {
                    jv.assign((size_t)n * 3, 0.0);
                    point_fields(r, sc, jv.data(), want_jmod ? jmod.data() : nullptr, nullptr, nullptr);
                }
            } else {
                tens = uhf ? cache[sc] : tensors(sc);
                jv.assign((size_t)n * 3, 0.0);
                if (want_jmod) jmod.assign((size_t)n, 0.0);
                if (want_acid) acid.assign((size_t)n, 0.0);
                check(gimic_b200_fields_from_tensors(ctx->h, n, r.data(), tens.data(), magnet.data(), jv.data(), want_jmod ? jmod.data() : nullptr,
                                                     want_acid ? acid.data() : nullptr, 0));
            }
            out.raw(" magnetic field\n" + ld_real(magnet[0]) + ld_real(magnet[1]) + ld_real(magnet[2]) + "\n \n");
            const bool regular = grid.mode == "std" || grid.mode == "base" || grid.mode == "bond";
            if (grid.gauss && !grid.is_file())
                if (root()) write_jmod_txt(join_path(workdir, "jmod" + tag + ".txt"), grid, jv, regular && (grid.mode == "bond" || grid.gtype == "even"));
            if (grid.is_3d()) {
                if (root() && inp.flag("Essential.acid")) write_vti_scalar(join_path(workdir, "acid.vti"), grid, acid, opt.vtk_appended);
                if (root() && inp.flag("Essential.jmod")) write_vti_scalar(join_path(workdir, "jmod" + tag + ".vti"), grid, jmod, opt.vtk_appended);
            }
            if (prop) run_property(tens);
            if (regular && grid.gtype == "even") {
                if (root()) write_vti_vector(join_path(workdir, "jvec" + tag + ".vti"), grid, radius_masked_vectors(grid, jv), opt.vtk_appended);
            } else if (((grid.mode == "std" || grid.mode == "base") && grid.gauss) || grid.is_file()) {
                const std::string ele = join_path(workdir, "grid.1.ele");
                if (root() && file_exists(ele)) write_vtu(join_path(workdir, "jvec.vtu"), r, "vectors", 3, jv, read_ele(ele));
                else out.raw(" not writing a vtu file, because the file grid.1.ele was not found.\n");
            }
        }
    }

    // get_property (jfield.f90:584-929): needs coord.au, gridfile.grd, grid_w.grd (and nelpts.info) in the work dir; the tensor
    // field must have been computed on the points of gridfile.grd (Grid(file)).
    void run_property(const std::vector<double> &tens) {
        const std::string need[3] = {join_path(workdir, "coord.au"), join_path(workdir, "gridfile.grd"), join_path(workdir, "grid_w.grd")};
        if (!(file_exists(need[0]) && file_exists(need[1]) && file_exists(need[2]))) {
            out.raw(" at least one of the files coord.au, gridfile.grd, and grid_w.grd is missing.Therefore any property calculation is skipped.\n");
            return;
        }
        std::vector<double> coord = read_numbers(need[0]), grd = read_numbers(need[1]);
        const std::vector<double> wg = read_numbers(need[2]);
        coord.resize(coord.size() / 3 * 3); grd.resize(grd.size() / 3 * 3);
        const long n = (long)grd.size() / 3;
        const int nat = (int)(coord.size() / 3);
        if ((long)wg.size() < n || (long)tens.size() < 9 * n) throw DriverError("get_property: gridfile.grd, grid_w.grd and the tensor field differ in length");
        std::vector<long> seg_end;
        const std::string nel = join_path(workdir, "nelpts.info");
        if (file_exists(nel)) {
            const std::vector<double> v = read_numbers(nel);
            long sum = 0;
            for (size_t i = 1; i < v.size(); i += 2) { sum += (long)v[i]; seg_end.push_back(sum); }
            if (sum != n) seg_end.clear();
        }
        if (seg_end.empty()) seg_end.push_back(n);
        const int nseg = (int)seg_end.size();
        std::vector<double> part((size_t)(nat + 1) * nseg * 5, 0.0);
        check(gimic_b200_property(ctx->h, n, grd.data(), wg.data(), tens.data(), nat, coord.data(), nseg, seg_end.data(), part.data(), 0));
        // running sums at the segment ends (the reference's scont), per-block contributions = their differences
        struct Res { double xyz[3], iso, pos, neg; std::vector<std::array<double, 3>> atoms; };
        std::vector<Res> res((size_t)nat + 1);
        for (int k = 0; k <= nat; ++k) {
            double cum[5] = {0, 0, 0, 0, 0};
            std::array<double, 3> prev{{0, 0, 0}};
            Res &R = res[(size_t)k];
            for (int s = 0; s < nseg; ++s) {
                for (int q = 0; q < 5; ++q) cum[q] += part[((size_t)k * nseg + s) * 5 + q];
                const std::array<double, 3> sc{{(cum[0] + cum[1] + cum[2]) / 3.0, cum[3] / 3.0, cum[4] / 3.0}};
                R.atoms.push_back({{sc[0] - prev[0], sc[1] - prev[1], sc[2] - prev[2]}});
                prev = sc;
            }
            for (int q = 0; q < 3; ++q) R.xyz[q] = cum[q];
            R.iso = (cum[0] + cum[1] + cum[2]) / 3.0; R.pos = cum[3] / 3.0; R.neg = cum[4] / 3.0;
        }
        // integrand plots (only when the TetGen cell file is there, jfield.f90:677-686, 786-808, 911-919)
        const std::string ele = join_path(workdir, "grid.1.ele");
        const bool have_cells = file_exists(ele);
        std::vector<long> cells;
        if (have_cells) cells = read_ele(ele);
        auto plot_integrands = [&](const double *centre, const std::vector<std::string> &names) {
            std::vector<double> f4((size_t)n * 4, 0.0), col((size_t)n);
            check(gimic_b200_property_integrand(ctx->h, n, grd.data(), tens.data(), centre, f4.data(), 0));
            const int cols[4] = {3, 0, 1, 2};
            for (int q = 0; q < 4; ++q) {
                for (long i = 0; i < n; ++i) col[(size_t)i] = f4[4 * (size_t)i + cols[q]];
                if (root()) write_vtu(join_path(workdir, names[(size_t)q]), grd, "scalars", 1, col, cells);
            }
        };
        // write(*,*) " " before the shielding tables, write(*,*) "" before the chi table (jfield.f90:762,890)
        auto table = [&](const std::vector<std::array<double, 3>> &contrib, const char *lead) {
            out.raw(std::string(lead) + "\n atom contributions, total, positive, negative\n");
            double cs[3] = {0, 0, 0};
            for (size_t l = 0; l < contrib.size(); ++l) {
                out.pf("atom %5zu%14.6f%14.6f%14.6f\n", l + 1, contrib[l][0], contrib[l][1], contrib[l][2]);
                for (int q = 0; q < 3; ++q) cs[q] += contrib[l][q];
            }
            out.pf("%10s%14.6f%14.6f%14.6f\n", "sum ", cs[0], cs[1], cs[2]);
            out.raw(" ****************************************************\n");
        };
        out.pf(" npts%12ld\n", n);
        for (int k = 0; k < nat; ++k) {
            const Res &R = res[(size_t)k];
            out.pf(" atom %12d\n in ppm\n", k + 1);
            const char *lbl[3] = {"sigma_xx ", "sigma_yy ", "sigma_zz "};
            for (int q = 0; q < 3; ++q) out.pf(" %10s  %14.6f\n", lbl[q], R.xyz[q]);
            out.pf("%30s  %14.6f\n", "shielding constant    = ", R.iso);
            out.pf("%30s  %14.6f\n", "positive contribution = ", R.pos);
            out.pf("%30s  %14.6f\n", "negative contribution = ", R.neg);
            out.pf("%30s  %14.6f\n", "sum = ", R.pos + R.neg);
            table(R.atoms, "  ");
            if (have_cells) {
                const std::string id = std::to_string(k + 1);
                const std::vector<std::string> names = {"sigma" + id + ".vtu", "sigma_xx" + id + ".vtu", "sigma_yy" + id + ".vtu", "sigma_zz" + id + ".vtu"};
                for (size_t q = 1; q < 4; ++q) out.pf(" %-70s\n", names[q].c_str());    // print *, filename (character(len=70), jfield.f90:606,801)
                plot_integrands(&coord[3 * (size_t)k], names);
            }
        }
        const Res &X = res[(size_t)nat];
        out.raw(" \n \n");
        const char *clbl[3] = {"chi_xx ", "chi_yy ", "chi_zz "};
        for (int q = 0; q < 3; ++q) out.pf(" %7s  %14.8f\n", clbl[q], X.xyz[q]);
        out.raw(" in au\n");
        // (X,A30,2X,F14.6) with 32-character labels, jfield.f90:876-878: the A30 edit descriptor keeps the leftmost 30 characters, so the
        // reference prints these three lines without their '= ' (test/benzene/magnetizability/reference/stdout)
        out.pf(" %30.30s  %14.6f\n", "isotropic magnetizability chi = ", X.iso);
        out.pf(" %30.30s  %14.6f\n", "positive contribution         = ", X.pos);
        out.pf(" %30.30s  %14.6f\n", "negative contribution         = ", X.neg);
        out.pf(" %30s  %14.6f\n \n", "sum ", X.pos + X.neg);
        const double fac = 7.89104e-29;                               // fac_au2simag, jfield.f90:606
        out.raw(" in SI units J/T^2 \n conversion factor: 7.89104*10^-29 J/T^2 \n \n");
        const std::pair<const char *, double> si[4] = {{"isotropic magnetizability = ", X.iso}, {"positive contribution     = ", X.pos},
                                                       {"negative contribution     = ", X.neg}, {"sum ", X.pos + X.neg}};
        for (const auto &p : si) out.pf("%30s  %s\n", p.first, fortran_e(p.second * fac, 14, 6).c_str());
        out.raw(" ****************************************************\n");
        table(X.atoms, " ");
        if (have_cells) plot_integrands(nullptr, {"intchi.vtu", "intchi_xx.vtu", "intchi_yy.vtu", "intchi_zz.vtu"});
    }

    // run_integral (gimic.F90:222-261) with the report formats of integral.f90:167-183,306-322,502-510
    void run_integral(const std::map<int, Sums> *pre) {
        out.say("Integrating current density");
        out.say("*****************************************");
        std::vector<int> cases;
        int what;
        integral_cases(cases, what);
        if (pre && !pre->empty()) {
            results = *pre;
        } else {
            const gimic_b200_grid g = grid.cstruct();
            for (int sc : cases) {
                // rows j split over the devices like schedule() (parallel.F90:66-84); the <= 7 partial sums are added on the host
                // in device order (the collect_sum of integral.f90:157-161), so the result does not depend on thread timing
                const size_t nd = 1 + peers.size();
                std::vector<Sums> part(nd, Sums{});
                const int w = sc == GIMIC_B200_TOTAL ? what : (what & 3);
                const double *B = magnet.data();
                Sums *pp = part.data();
                over_devices((long)grid.npts[1], [=](gimic_b200_handle h, size_t d, long lo, long hi) {
                    return gimic_b200_integrate(h, &g, B, sc, w, (int)lo, (int)hi, pp[d].data());
                });
                Sums s = part[0];
                for (size_t d = 1; d < nd; ++d) for (int k = 0; k < 7; ++k) s[(size_t)k] += part[d][(size_t)k];
                if (ranked()) {
                    if (!opt.allreduce_sum) throw DriverError("a run over several processes needs the allreduce_sum callback (gimic_b200_run_opts)");
                    if (opt.allreduce_sum(opt.user, s.data(), 7) < 0) throw DriverError("the launcher's allreduce_sum callback failed");
                }
                results[sc] = s;
            }
        }
        const std::string bar(60, '*');
        const double bound = grid.radius;
        auto note_spin = [&](int sc) { if (uhf) out.say(std::string("*** Integrating ") + SPIN_LABEL[sc] + " density"); };
        auto block = [&](const char *lbl_au, const char *lbl_si, double x, double p, double m) {
            out.say();
            out.say(bar);
            out.say(sfmt("%s%13.6f", lbl_au, x));
            out.say(sfmt("      Positive contribution:%13.6f  (%11.6f )", p, au2si(p)));
            out.say(sfmt("      Negative contribution:%13.6f  (%11.6f )", m, au2si(m)));
            out.say();
            out.say(sfmt("%s%13.6f", lbl_si, au2si(x)));
            out.say(sfmt("      (conversion factor)  :%13.6f", au2si(1.0)));
            out.say(bar);
            out.say();
        };
        auto one = [&](int sc, int off, const char *lbl_au, const char *lbl_si) {
            note_spin(sc);
            field_line();
            if (bound < 1.0e10) {                   // write(str_g, *) 'Integration bound set to radius ', bound (integral.f90:93,231); msg_out trims
                std::string v = ld_real(bound);
                while (!v.empty() && v.back() == ' ') v.pop_back();
                out.say(" Integration bound set to radius " + v);
            }
            const Sums &s = results.at(sc);
            block(lbl_au, lbl_si, s[off], s[off + 1], s[off + 2]);
        };
        if (inp.flag("Essential.jmod")) {
            out.say("*** Integrating |J|");
            for (int sc : cases) one(sc, 3, "Induced mod current (au)   :", "Induced mod current (nA/T) :");
            out.say();
        } else {
            out.raw(" Jmod integration skipped.\n");              // write(*,*) "...", gimic.F90:242
        }
        out.say("*** Integrating current");
        for (int sc : cases) one(sc, 0, "   Induced current (au)    :", "   Induced current (nA/T)  :");
        out.say();
        if (inp.flag("Essential.acid")) {
            out.say("*** Integrating ACID density");
            const double acid = std::sqrt(results.at(GIMIC_B200_TOTAL)[6]);
            out.say();
            out.say(bar);
            out.say(sfmt("   ACID (au) sqrt(delta J^2):%13.6f", acid));
            out.say(sfmt("   ACID (nA/T)              :%13.6f", au2si(acid)));
            out.say();
            out.say(bar);
            out.say();
            out.say();                              // call nl after integrate_acid, gimic.F90:257
        }
    }

    // edens / divj: whitelisted by the reference front end (src/gimic.in:267) but not implemented at this commit.  Defined in
    // the library as rho = Phi^T D Phi and div(T.B) (central differences); <calc>.vti on even image grids, '<x y z value>'
    // rows otherwise.  No reference output exists: parity unpinned.
    void run_scalar(const std::string &calc) {
        const std::vector<double> r = grid.points();
        const long n = grid.n();
        std::vector<double> v((size_t)n, 0.0);
        const bool ed = calc == "edens";
        point_fields(r, GIMIC_B200_TOTAL, nullptr, nullptr, ed ? v.data() : nullptr, ed ? nullptr : v.data());
        if (!root()) return;
        if (!grid.is_file() && grid.gtype == "even" && grid.npts[0] > 1 && grid.npts[1] > 1)
            write_vti_scalar(join_path(workdir, calc + ".vti"), grid, v, opt.vtk_appended);
        else
            write_points_txt(join_path(workdir, calc + ".txt"), r, v);
    }
};

int fail(int code, const std::string &msg) { g_error = msg; return code; }

template <class F>
int guarded(F &&body) {
    try {
        g_error.clear();
        body();
        return 0;
    } catch (const InputError &e) {
        return fail(std::string(e.what()).find("cannot open") == 0 ? GIMIC_B200_EIO : GIMIC_B200_EINVAL, e.what());
    } catch (const DriverError &e) {
        // errors that come out of the library keep its message; classify by what the library reported last
        const std::string lib = gimic_b200_last_error();
        int code = GIMIC_B200_EINVAL;
        if (!lib.empty() && lib == e.what()) {
            if (lib.find("CUDA") != std::string::npos || lib.find("cuda") != std::string::npos) code = GIMIC_B200_ECUDA;
            else if (lib.find("open") != std::string::npos || lib.find("read") != std::string::npos) code = GIMIC_B200_EIO;
        } else if (std::string(e.what()).find("cannot open") == 0 || std::string(e.what()).find("cannot write") == 0) {
            code = GIMIC_B200_EIO;
        }
        return fail(code, e.what());
    } catch (const std::bad_alloc &) {
        return fail(GIMIC_B200_ENOMEM, "out of host memory");
    } catch (const std::exception &e) {
        return fail(GIMIC_B200_EINVAL, e.what());
    }
}

std::string stem_of(const std::string &path) {
    const size_t slash = path.find_last_of('/');
    const size_t dot = path.find_last_of('.');
    if (dot == std::string::npos || (slash != std::string::npos && dot < slash) || dot == (slash == std::string::npos ? 0 : slash + 1)) return path;
    return path.substr(0, dot);
}

}  // namespace

// au2si, globals.f90:309-332 (nA/T per atomic unit of dJ/dB)
double au2si(double au) {
    const double aulength = 0.52917726e-10, auspeedoflight = 137.03599e0, speedoflight = 299792458.0;
    const double aucharge = 1.60217733e-19, hbar = 1.05457267e-34;
    const double autime = aulength * auspeedoflight / speedoflight;
    const double autesla = hbar / aucharge / aulength / aulength;
    return au * (aucharge / autime / autesla) * 1.0e9;
}

bool file_exists(const std::string &path) { struct stat st; return ::stat(path.c_str(), &st) == 0; }

std::string join_path(const std::string &dir, const std::string &name) {
    if (!name.empty() && name[0] == '/') return name;
    if (dir.empty()) return name;
    return dir.back() == '/' ? dir + name : dir + "/" + name;
}

std::string dirname_of(const std::string &path) {
    const std::string abs = real_path(path);
    const size_t slash = abs.find_last_of('/');
    if (slash == std::string::npos) return ".";
    return slash == 0 ? "/" : abs.substr(0, slash);
}

const std::string &last_error_message() { return g_error; }

int run_input(const std::string &inpfile, const RunOptions &opt, FILE *out) {
    return guarded([&] {
        Run run(inpfile, opt, out, [](const std::string &) { return std::shared_ptr<Context>(); });
        run.run();
        std::fflush(out);
    });
}

// A current-profile scan (jobscripts/src/current-profile-local-submit: `gimic gimic.N.inp > gimic.N.out` for hundreds of thin
// slices, one process and one MOL/XDENS read each) as ONE context and ONE tensor pass per spin case: all inputs that share
// basis, densities and Advanced settings are integrated by gimic_b200_integrate_batch.
int run_scan(const std::vector<std::string> &inpfiles, const RunOptions &opt) {
    return guarded([&] {
        std::vector<std::unique_ptr<Run>> runs;
        for (const std::string &f : inpfiles) {
            RunOptions ro = opt;
            ro.workdir.clear();                                   // every input runs in its own directory
            if (!ro.devices.empty()) { ro.device = ro.devices[0]; ro.devices.clear(); }   // a scan batches all planes on one device
            auto finder = [&](const std::string &key) {
                for (const auto &r : runs) if (r->ctx && r->context_key == key) return r->ctx;
                return std::shared_ptr<Context>();
            };
            // the report file is only open while it is written (a scan can have more slices than the process may hold open files)
            runs.emplace_back(new Run(f, ro, nullptr, finder));
        }
        std::vector<std::map<int, Sums>> pre(runs.size());
        std::vector<Context *> order;                             // contexts in order of first appearance
        std::map<Context *, std::vector<size_t>> by_ctx;
        for (size_t i = 0; i < runs.size(); ++i) {
            const Run &r = *runs[i];
            if (r.inp.str("calc") != "integral" || r.inp.flag("dryrun")) continue;
            if (!by_ctx.count(r.ctx.get())) order.push_back(r.ctx.get());
            by_ctx[r.ctx.get()].push_back(i);
        }
        for (Context *c : order) {
            const std::vector<size_t> &ids = by_ctx[c];
            std::vector<int> cases;
            int what0;
            runs[ids[0]]->integral_cases(cases, what0);
            std::vector<gimic_b200_grid> grids;
            std::vector<double> Bs;
            for (size_t i : ids) {
                grids.push_back(runs[i]->grid.cstruct());
                for (int k = 0; k < 3; ++k) Bs.push_back(runs[i]->magnet[(size_t)k]);
            }
            for (int sc : cases) {
                int what = 0;
                for (size_t i : ids) {
                    std::vector<int> tmp;
                    int w;
                    runs[i]->integral_cases(tmp, w);
                    what |= sc == GIMIC_B200_TOTAL ? w : (w & 3);
                }
                std::vector<double> sums(ids.size() * 7, 0.0);
                check(gimic_b200_integrate_batch(c->h, (int)ids.size(), grids.data(), Bs.data(), sc, what, sums.data()));
                for (size_t q = 0; q < ids.size(); ++q) {
                    Sums s;
                    for (int k = 0; k < 7; ++k) s[(size_t)k] = sums[q * 7 + (size_t)k];
                    pre[ids[q]][sc] = s;
                }
            }
        }
        for (size_t i = 0; i < runs.size(); ++i) {
            const std::string rep = stem_of(inpfiles[i]) + ".out";
            struct Report { FILE *f; ~Report() { if (f) std::fclose(f); } } report{std::fopen(rep.c_str(), "w")};
            if (!report.f) throw DriverError("cannot write " + rep);
            runs[i]->out.f = report.f;
            runs[i]->run(pre[i].empty() ? nullptr : &pre[i]);
            runs[i]->out.f = nullptr;
        }
        // current_profile.dat next to the first input: slice position (index x delta when the jobscripts' calculation.dat is there,
        // jobscripts/src/current-profile-header:38, else the index), net / diatropic / paratropic current in nA/T -- what
        // jobscripts/src/gradient.sh.in:38-47 assembles by grepping every gimic.N.out, here from the unrounded sums (8 decimals)
        std::vector<const Run *> rows;
        for (const auto &r : runs)
            if (r->inp.str("calc") == "integral" && !r->inp.flag("dryrun") && r->results.count(GIMIC_B200_TOTAL)) rows.push_back(r.get());
        if (rows.size() >= 2) {
            bool has_delta = false;
            double delta = 0.0;
            {
                std::ifstream cf(join_path(rows[0]->workdir, "calculation.dat"));
                std::stringstream ss;
                if (cf) ss << cf.rdbuf();
                const std::string txt = ss.str();
                const size_t p = txt.find("delta=");
                if (p != std::string::npos) {
                    std::string tok = txt.substr(p + 6, 40);
                    for (char &ch : tok) if (ch == 'd' || ch == 'D') ch = 'e';
                    char *end = nullptr;
                    delta = std::strtod(tok.c_str(), &end);
                    has_delta = end != tok.c_str();
                }
            }
            const std::string ppath = join_path(rows[0]->workdir, "current_profile.dat");
            FILE *pf = std::fopen(ppath.c_str(), "w");
            if (!pf) throw DriverError("cannot write " + ppath);
            for (size_t k = 0; k < rows.size(); ++k) {
                const Sums &sm = rows[k]->results.at(GIMIC_B200_TOTAL);
                if (has_delta) std::fprintf(pf, "%5.2f", (double)k * delta); else std::fprintf(pf, "%5zu", k);
                std::fprintf(pf, "\t% .8f\t% .8f\t% .8f\n", au2si(sm[0]), au2si(sm[1]), au2si(sm[2]));
            }
            std::fclose(pf);
        }
    });
}

// XDENS text -> binary cache next to it (<xdens>.bin) for the basis / open-shell / spherical settings of a gimic.inp.  A 10^4-function
// XDENS is 3.2 GB of text (4e8 list-directed reads in read_dens, dens.f90:129-135); the cache is read at disk speed.
int cache_xdens(const std::string &inpfile, const std::string &workdir_in, std::string &written) {
    return guarded([&] {
        const std::string workdir = workdir_in.empty() ? dirname_of(inpfile) : workdir_in;
        const Input inp = parse_file(inpfile);
        auto resolve = [&](const std::string &n) { return (!n.empty() && n[0] == '/') ? n : join_path(workdir, n); };
        const std::string mol = resolve(inp.str("basis")), xdens = resolve(inp.str("xdens"));
        int info[5];
        check(gimic_b200_mol_summary(mol.c_str(), info));
        const int nbf = inp.flag("Advanced.spherical") ? info[4] : info[2];
        written = xdens + ".bin";
        check(gimic_b200_convert_xdens(xdens.c_str(), nbf, inp.flag("openshell") ? 8 : 4, written.c_str()));
    });
}

// Write an array computed elsewhere (e.g. by the reference's Fortran loops calling the batched C ABI) on the grid of a gimic.inp
int write_field(const std::string &inpfile, const std::string &workdir_in, const std::string &kind, const double *data, long n, const std::string &name,
                bool appended) {
    return guarded([&] {
        const std::string workdir = workdir_in.empty() ? dirname_of(inpfile) : workdir_in;
        const Input inp = parse_file(inpfile);
        const std::string basis = inp.str("basis");
        const std::string mol = (!basis.empty() && basis[0] == '/') ? basis : join_path(workdir, basis);
        const int natoms = gimic_b200_mol_geometry(mol.c_str(), 0, nullptr, nullptr);
        check(natoms);
        std::vector<double> xyz((size_t)natoms * 3, 0.0);
        check(gimic_b200_mol_geometry(mol.c_str(), natoms, xyz.data(), nullptr));
        const GridSpec grid = grid_from_input(inp, xyz, workdir);
        const long np = grid.n();
        const int ncomp = (kind == "vti_vector" || kind == "vtu_vector" || kind == "jmod_txt") ? 3 : 1;
        if (!data || n != np * ncomp) throw DriverError("write_field: expected " + std::to_string(np * ncomp) + " values for " + kind + ", got " + std::to_string(n));
        const std::vector<double> v(data, data + n);
        const std::string path = join_path(workdir, name);
        if (kind == "vti_scalar") write_vti_scalar(path, grid, v, appended);
        else if (kind == "vti_vector") write_vti_vector(path, grid, radius_masked_vectors(grid, v), appended);
        else if (kind == "jmod_txt") write_jmod_txt(path, grid, v, grid.mode == "bond" || (grid.mode != "file" && grid.gtype == "even"));
        else if (kind == "vtu_vector" || kind == "vtu_scalar") {
            const std::string ele = join_path(workdir, "grid.1.ele");
            if (!file_exists(ele)) throw DriverError("cannot open " + ele);
            write_vtu(path, grid.points(), ncomp == 3 ? "vectors" : "scalars", ncomp, v, read_ele(ele));
        } else throw DriverError("write_field: unknown kind '" + kind + "'");
    });
}

// The grid and the field direction a gimic.inp describes (grid.f90 new_grid + magnet.f90 get_magnet), for a caller that feeds the
// batched C ABI itself.  Host only.
int input_grid(const std::string &inpfile, const std::string &workdir_in, gimic_b200_grid_info *info, double *pts, double *wgt, long cap) {
    return guarded([&] {
        const std::string workdir = workdir_in.empty() ? dirname_of(inpfile) : workdir_in;
        const Input inp = parse_file(inpfile);
        const std::string basis = inp.str("basis");
        const std::string mol = (!basis.empty() && basis[0] == '/') ? basis : join_path(workdir, basis);
        const int natoms = gimic_b200_mol_geometry(mol.c_str(), 0, nullptr, nullptr);
        check(natoms);
        std::vector<double> xyz((size_t)natoms * 3, 0.0);
        check(gimic_b200_mol_geometry(mol.c_str(), natoms, xyz.data(), nullptr));
        const GridSpec grid = grid_from_input(inp, xyz, workdir);
        const Vec3 b = get_magnet(grid, inp.str("magnet_axis"), inp.vec3("magnet"));
        std::memset(info, 0, sizeof *info);
        info->is_file = grid.is_file() ? 1 : 0;
        info->npoints = grid.n();
        for (int k = 0; k < 3; ++k) {
            info->npts[k] = grid.npts[k]; info->origin[k] = grid.origin[(size_t)k]; info->magnet[k] = b[(size_t)k];
            info->lengths[k] = grid.lengths[(size_t)k]; info->center_bond[k] = grid.center_bond[(size_t)k];
            for (int c = 0; c < 3; ++c) info->basv[3 * k + c] = grid.basv[k][c];
        }
        info->radius = grid.radius; info->has_center_bond = grid.has_center_bond ? 1 : 0;
        // axis coordinates and weights one axis after the other (npts[0] + npts[1] + npts[2] values each); file grids: the 3 x n points in pts
        const long need = grid.is_file() ? 3 * grid.n() : (long)grid.npts[0] + grid.npts[1] + grid.npts[2];
        if (pts || wgt) {
            if (cap < need) throw DriverError("input_grid: room for " + std::to_string(cap) + " values, the grid needs " + std::to_string(need));
            if (grid.is_file()) { if (pts) std::copy(grid.xdata.begin(), grid.xdata.end(), pts); }
            else {
                long o = 0;
                for (int k = 0; k < 3; ++k)
                    for (int i = 0; i < grid.npts[k]; ++i, ++o) { if (pts) pts[o] = grid.pts[k][(size_t)i]; if (wgt) wgt[o] = grid.wgt[k][(size_t)i]; }
            }
        }
    });
}

}  // namespace gbd

// ------------------------------------------------------------------------------------------------------- C ABI
extern "C" {

int gimic_b200_input_grid(const char *inpfile, const char *workdir, gimic_b200_grid_info *info, double *pts, double *wgt, long cap) {
    if (!inpfile || !info) { gbd::g_error = "null argument"; return GIMIC_B200_EINVAL; }
    return gbd::input_grid(inpfile, workdir ? workdir : "", info, pts, wgt, cap);
}

int gimic_b200_run(const char *inpfile, const gimic_b200_run_opts *opts);

int gimic_b200_run_input(const char *inpfile, const char *workdir, int device, int flags, const char *report_path) {
    gimic_b200_run_opts o;
    std::memset(&o, 0, sizeof o);
    o.flags = flags; o.device = device; o.workdir = workdir; o.report_path = report_path;
    return gimic_b200_run(inpfile, &o);
}

int gimic_b200_run(const char *inpfile, const gimic_b200_run_opts *opts) {
    if (!inpfile) { gbd::g_error = "null argument"; return GIMIC_B200_EINVAL; }
    gimic_b200_run_opts d;
    std::memset(&d, 0, sizeof d);
    d.device = -1;
    if (opts) d = *opts;
    if (d.ndevices > 0 && !d.devices) { gbd::g_error = "bad argument: ndevices > 0 without a device list"; return GIMIC_B200_EINVAL; }
    gbd::RunOptions o;
    o.dryrun = (d.flags & GIMIC_B200_RUN_DRYRUN) != 0;
    o.vtk_appended = (d.flags & GIMIC_B200_RUN_VTK_APPENDED) != 0;
    o.device = d.device;
    if (d.ndevices < 0 && !o.dryrun) {                            // every GPU of the node
        const int nd = gimic_b200_device_count();
        if (nd < 0) { gbd::g_error = gimic_b200_last_error(); return nd; }
        for (int k = 0; k < nd; ++k) o.devices.push_back(k);
    } else if (d.ndevices > 0) {
        o.devices.assign(d.devices, d.devices + d.ndevices);
    }
    if (!o.devices.empty()) o.device = o.devices[0];
    if (d.title) o.title = d.title;
    if (d.workdir) o.workdir = d.workdir;
    if (d.nranks > 1) {
        if (d.rank < 0 || d.rank >= d.nranks) { gbd::g_error = "bad argument: rank outside [0, nranks)"; return GIMIC_B200_EINVAL; }
        if (!o.devices.empty() && o.devices.size() > 1) { gbd::g_error = "bad argument: a rank of a multi-process run drives one device"; return GIMIC_B200_EINVAL; }
        if (!o.dryrun && (!d.allgather_rows || !d.allreduce_sum)) { gbd::g_error = "bad argument: nranks > 1 needs both collective callbacks"; return GIMIC_B200_EINVAL; }
        o.devices.clear();
        o.rank = d.rank; o.nranks = d.nranks; o.allgather_rows = d.allgather_rows; o.allreduce_sum = d.allreduce_sum; o.user = d.user;
    }
    FILE *out = stdout;
    const char *report = o.rank > 0 ? "/dev/null" : d.report_path;       // rank 0 alone reports
    if (report) {
        out = std::fopen(report, "w");
        if (!out) { gbd::g_error = std::string("cannot write ") + report; return GIMIC_B200_EIO; }
    }
    const int rc = gbd::run_input(inpfile, o, out);
    if (report) std::fclose(out); else std::fflush(out);
    return rc;
}

int gimic_b200_run_scan(int n, const char *const *inpfiles, int device, int flags) {
    if (n < 0 || (n > 0 && !inpfiles)) { gbd::g_error = "bad argument"; return GIMIC_B200_EINVAL; }
    std::vector<std::string> files;
    for (int i = 0; i < n; ++i) {
        if (!inpfiles[i]) { gbd::g_error = "null input file name"; return GIMIC_B200_EINVAL; }
        files.emplace_back(inpfiles[i]);
    }
    gbd::RunOptions o;
    o.dryrun = (flags & GIMIC_B200_RUN_DRYRUN) != 0;
    o.vtk_appended = (flags & GIMIC_B200_RUN_VTK_APPENDED) != 0;
    o.device = device;
    return gbd::run_scan(files, o);
}

int gimic_b200_write_field(const char *inpfile, const char *workdir, const char *kind, const double *data, long n, const char *filename, int flags) {
    if (!inpfile || !kind || !filename) { gbd::g_error = "null argument"; return GIMIC_B200_EINVAL; }
    return gbd::write_field(inpfile, workdir ? workdir : "", kind, data, n, filename, (flags & GIMIC_B200_RUN_VTK_APPENDED) != 0);
}

int gimic_b200_cache_xdens(const char *inpfile, const char *workdir, char *written, int cap) {
    if (!inpfile) { gbd::g_error = "null argument"; return GIMIC_B200_EINVAL; }
    std::string out;
    const int rc = gbd::cache_xdens(inpfile, workdir ? workdir : "", out);
    if (rc == 0 && written && cap > 0) std::snprintf(written, (size_t)cap, "%s", out.c_str());
    return rc;
}

const char *gimic_b200_driver_last_error(void) { return gbd::g_error.c_str(); }

}  // extern "C"
