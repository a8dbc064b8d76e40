// Calls every entry point of include/gimic_b200.h (incl. error paths and the legacy symbols) against the FAKE CUDA runtime of this
// directory; built with AddressSanitizer + UBSan by tools/sanitize_api_host.sh.  Results are zeros (kernels are no-ops): the point is
// the host-side orchestration of api.cu / host_basis.cpp.  Test infrastructure only.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "gimic_b200.h"

static int failures = 0;
#define EXPECT(cond) do { if (!(cond)) { std::printf("FAILED %s:%d  %s   [%s]\n", __FILE__, __LINE__, #cond, gimic_b200_last_error()); ++failures; } } while (0)

int main(int argc, char **argv) {
    if (argc < 5) { std::printf("usage: api_harness MOL XDENS MOL_UHF XDENS_UHF\n"); return 2; }
    gimic_b200_opts o;
    gimic_b200_default_opts(&o);
    o.screening_thrs = 1e-8;
    EXPECT(gimic_b200_device_count() == 2 || std::getenv("FAKE_CUDA_NO_DEVICE"));
    // ---- closed shell from files
    gimic_b200_handle h = nullptr;
    if (gimic_b200_create(&h, argv[1], argv[2], &o) != 0) {          // e.g. FAKE_CUDA_NO_DEVICE=1: the library must refuse, not compute
        std::printf("create refused: %s\n", gimic_b200_last_error());
        return 3;
    }
    EXPECT(gimic_b200_destroy(h) == 0);
    EXPECT(gimic_b200_create(&h, argv[1], "/nonexistent", &o) == GIMIC_B200_EIO && h == nullptr);
    EXPECT(gimic_b200_create(&h, argv[1], argv[2], &o) == 0);
    const int nbf = gimic_b200_nbf(h), nat = gimic_b200_natoms(h);
    EXPECT(nbf == 168 && nat == 8 && gimic_b200_is_uhf(h) == 0);
    std::vector<double> xyz(3 * nat);
    EXPECT(gimic_b200_atom_coords(h, xyz.data()) == 0);
    const long n = 5000;                                    // several tiles
    std::vector<double> r(3 * n), tens(9 * n, 1.0), jv(3 * n, 1.0), jm(n, 1.0), ac(n, 1.0), ed(n, 1.0), dj(n, 1.0);
    for (long i = 0; i < 3 * n; ++i) r[i] = 8.0 * std::sin(0.37 * (double)i);
    const double B[3] = {0.0, 0.0, 1.0};
    EXPECT(gimic_b200_set_profiling(h, 1) == 0);
    EXPECT(gimic_b200_calc_jtensors(h, n, r.data(), GIMIC_B200_TOTAL, tens.data(), 0) == 0);
    EXPECT(gimic_b200_calc_jtensors(h, 0, r.data(), GIMIC_B200_TOTAL, tens.data(), 0) == 0);
    EXPECT(gimic_b200_calc_jtensors(h, -1, r.data(), GIMIC_B200_TOTAL, tens.data(), 0) == GIMIC_B200_EINVAL);
    EXPECT(gimic_b200_calc_jtensors(h, n, r.data(), GIMIC_B200_BETA, tens.data(), 0) == GIMIC_B200_ESPIN);
    EXPECT(gimic_b200_calc_jtensors(h, n, r.data(), 7, tens.data(), 0) == GIMIC_B200_EINVAL);
    EXPECT(gimic_b200_calc_fields(h, n, r.data(), B, GIMIC_B200_TOTAL, tens.data(), jv.data(), jm.data(), ac.data(), ed.data(), dj.data(), 1e-3, 0) == 0);
    EXPECT(gimic_b200_calc_fields(h, n, r.data(), B, GIMIC_B200_TOTAL, nullptr, jv.data(), jm.data(), nullptr, ed.data(), nullptr, 0.0, 0) == 0);   // J path
    EXPECT(gimic_b200_calc_fields(h, n, r.data(), B, GIMIC_B200_TOTAL, nullptr, nullptr, jm.data(), nullptr, nullptr, nullptr, 0.0, 0) == 0);        // |J| alone
    EXPECT(gimic_b200_calc_fields(h, n, r.data(), nullptr, GIMIC_B200_TOTAL, nullptr, jv.data(), nullptr, nullptr, nullptr, nullptr, 0.0, 0) == GIMIC_B200_EINVAL);
    EXPECT(gimic_b200_calc_fields(h, n, r.data(), B, GIMIC_B200_TOTAL, tens.data(), jv.data(), jm.data(), ac.data(), ed.data(), dj.data(), 1e-3,
                                  GIMIC_B200_DEVICE_PTR) == 0);                                     // "device" pointers (host memory under the fake runtime)
    EXPECT(gimic_b200_fields_from_tensors(h, n, r.data(), tens.data(), B, jv.data(), jm.data(), ac.data(), 0) == 0);
    EXPECT(gimic_b200_fields_from_tensors(h, n, nullptr, tens.data(), B, jv.data(), nullptr, ac.data(), 0) == 0);
    EXPECT(gimic_b200_jmod_from_jvec(h, n, r.data(), jv.data(), B, jm.data(), 0) == 0);
    std::vector<double> bf((size_t)64 * nbf), dr((size_t)64 * 3 * nbf);
    EXPECT(gimic_b200_calc_basis(h, 64, r.data(), bf.data(), dr.data(), 0) == 0);
    EXPECT(gimic_b200_calc_basis(h, 64, r.data(), nullptr, nullptr, 0) == GIMIC_B200_EINVAL);
    gimic_b200_stats st;
    EXPECT(gimic_b200_get_stats(h, &st) == 0);
    // ---- grids, quadrature, batch
    std::vector<double> p0(36), p1(36), p2(1, 0.0), w0(36), w1(36), w2(1, 1.0);
    EXPECT(gimic_b200_gauss_points(0.0, 10.0, 36, 9, 0, p0.data(), w0.data()) == 0);
    EXPECT(gimic_b200_gauss_points(0.0, 7.25, 36, 9, 1, p1.data(), w1.data()) == 0);
    EXPECT(gimic_b200_gauss_points(0.0, 7.25, 35, 9, 0, p1.data(), w1.data()) == GIMIC_B200_EINVAL);
    EXPECT(gimic_b200_gauss_points(0.0, 7.25, 36, 9, 0, p1.data(), w1.data()) == 0);
    gimic_b200_grid g;
    const double basv[9] = {0, 0, -1, 0, 1, 0, 1, 0, 0};
    for (int i = 0; i < 3; ++i) g.origin[i] = -2.0;
    for (int i = 0; i < 9; ++i) g.basv[i] = basv[i];
    g.npts[0] = 36; g.npts[1] = 36; g.npts[2] = 1;
    g.pts[0] = p0.data(); g.pts[1] = p1.data(); g.pts[2] = p2.data(); g.wgt[0] = w0.data(); g.wgt[1] = w1.data(); g.wgt[2] = w2.data();
    g.radius = 3.0;
    std::vector<double> tg((size_t)9 * 36 * 36);
    EXPECT(gimic_b200_calc_jtensors_grid(h, &g, 0, 36 * 36, GIMIC_B200_TOTAL, tg.data(), 0) == 0);
    EXPECT(gimic_b200_calc_jtensors_grid(h, &g, 100, 700, GIMIC_B200_TOTAL, tg.data(), 0) == 0);
    EXPECT(gimic_b200_calc_jtensors_grid(h, &g, 100, 36 * 36 + 1, GIMIC_B200_TOTAL, tg.data(), 0) == GIMIC_B200_EINVAL);
    double out7[7];
    EXPECT(gimic_b200_integrate(h, &g, B, GIMIC_B200_TOTAL, 7, 0, 36, out7) == 0);
    EXPECT(gimic_b200_integrate(h, &g, B, GIMIC_B200_TOTAL, 1, 10, 20, out7) == 0);
    EXPECT(gimic_b200_integrate(h, &g, B, GIMIC_B200_TOTAL, 1, 20, 10, out7) == GIMIC_B200_EINVAL);
    std::vector<gimic_b200_grid> gs(5, g);
    std::vector<double> Bs(15, 0.0), outs(35);
    for (int k = 0; k < 5; ++k) { Bs[3 * k + 2] = 1.0; gs[k].origin[0] += 0.3 * k; }
    EXPECT(gimic_b200_integrate_batch(h, 5, gs.data(), Bs.data(), GIMIC_B200_TOTAL, 3, outs.data()) == 0);
    EXPECT(gimic_b200_integrate_batch(h, 0, gs.data(), Bs.data(), GIMIC_B200_TOTAL, 3, outs.data()) == 0);
    // ---- cost-balanced partition: every rank tiles the whole set; the shares are disjoint and cover it
    {
        long total = 0, info[8];
        std::vector<long> idx(n);
        EXPECT(gimic_b200_partition_calc(h, B, GIMIC_B200_TOTAL, idx.data(), tens.data(), nullptr, nullptr, nullptr, nullptr, 0) == GIMIC_B200_EINVAL);   // no plan yet
        long cost_sum = 0, tiles_sum = 0;
        for (int rk = 0; rk < 3; ++rk) {
            long cnt = -1;
            EXPECT(gimic_b200_partition_points(h, n, r.data(), 0, rk, 3, &cnt) == 0 && cnt >= 0);
            EXPECT(gimic_b200_partition_info(h, info) == 0 && info[0] == n && info[1] == cnt && info[3] <= info[2] && info[5] <= info[4]);
            EXPECT(gimic_b200_partition_calc(h, B, GIMIC_B200_TOTAL, idx.data(), tens.data(), jv.data(), jm.data(), ac.data(), ed.data(), 0) == 0);
            EXPECT(gimic_b200_partition_calc(h, B, GIMIC_B200_TOTAL, idx.data(), nullptr, jv.data(), jm.data(), nullptr, nullptr, 0) == 0);   // plan reused, J path
            if (std::getenv("FAKE_DRAIN_CHUNK") && info[3] > 1) {      // host outputs + drain groups: one contraction launch per group of tiles
                gimic_b200_stats st;
                const bool every_tile_its_own = std::atol(std::getenv("FAKE_DRAIN_CHUNK")) < 50000;     // a c4h4 tile holds ~88 000 panel doubles
                EXPECT(gimic_b200_get_stats(h, &st) == 0 && st.contract_launches >= (every_tile_its_own ? 2 : 1) && st.contract_launches <= 4);
            }
            total += cnt; cost_sum += info[5]; tiles_sum += info[3];
            // balanced to within one tile: no share is more than the mean + the largest tile cost (bounded here by the total / 2)
            EXPECT(3 * info[5] <= info[4] + 3 * (info[4] / 2));
        }
        EXPECT(total == n && cost_sum == info[4] && tiles_sum == info[2]);
        long cnt = 0;
        EXPECT(gimic_b200_partition_points(h, n, r.data(), 0, 3, 3, &cnt) == GIMIC_B200_EINVAL);
        EXPECT(gimic_b200_partition_grid(h, &g, 1, 2, &cnt) == 0 && cnt > 0 && cnt < 36 * 36);
        EXPECT(gimic_b200_partition_calc(h, B, GIMIC_B200_TOTAL, idx.data(), tens.data(), nullptr, nullptr, nullptr, nullptr, 0) == 0);
        EXPECT(gimic_b200_calc_jtensors(h, 100, r.data(), GIMIC_B200_TOTAL, tens.data(), 0) == 0);      // any compute call drops the plan
        EXPECT(gimic_b200_partition_calc(h, B, GIMIC_B200_TOTAL, idx.data(), tens.data(), nullptr, nullptr, nullptr, nullptr, 0) == GIMIC_B200_EINVAL);
    }
    // ---- property
    std::vector<double> w(n, 0.01);
    long seg[3] = {1000, 1000, n};
    std::vector<double> part((size_t)(nat + 1) * 3 * 5), f4((size_t)4 * n);
    EXPECT(gimic_b200_property(h, n, r.data(), w.data(), tens.data(), nat, xyz.data(), 3, seg, part.data(), 0) == 0);
    long bad_seg[2] = {10, n - 1};
    EXPECT(gimic_b200_property(h, n, r.data(), w.data(), tens.data(), nat, xyz.data(), 2, bad_seg, part.data(), 0) == GIMIC_B200_EINVAL);
    EXPECT(gimic_b200_property_integrand(h, n, r.data(), tens.data(), xyz.data(), f4.data(), 0) == 0);
    EXPECT(gimic_b200_property_integrand(h, n, r.data(), tens.data(), nullptr, f4.data(), 0) == 0);
    EXPECT(gimic_b200_destroy(h) == 0);
    // ---- open shell from files: all four spin cases, both paths
    o.uhf = 1;
    EXPECT(gimic_b200_create(&h, argv[3], argv[4], &o) == 0);
    for (int sc = 0; sc < 4; ++sc) {
        EXPECT(gimic_b200_calc_jtensors(h, 700, r.data(), sc, tens.data(), 0) == 0);
        EXPECT(gimic_b200_calc_fields(h, 700, r.data(), B, sc, nullptr, jv.data(), nullptr, nullptr, nullptr, nullptr, 0.0, 0) == 0);
    }
    const double B2[3] = {0.0, 1.0, 0.0};
    EXPECT(gimic_b200_calc_fields(h, 700, r.data(), B2, GIMIC_B200_TOTAL, nullptr, jv.data(), nullptr, nullptr, nullptr, nullptr, 0.0, 0) == 0);   // operand rebuilt for a new B
    // ---- in-memory construction from what the file context holds: a synthetic two-atom molecule with s..f shells, UHF, spherical on/off
    const double coords[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 2.5};
    const int nctr[2] = {4, 2}, cl[6] = {0, 1, 2, 3, 0, 1}, cn[6] = {2, 1, 1, 1, 1, 1};
    const double xp[7] = {3.0, 0.5, 0.8, 0.6, 0.9, 0.4, 0.7}, cc[7] = {0.4, 0.7, 1.0, 1.0, 1.0, 1.0, 1.0};
    for (int sph = 0; sph < 2; ++sph) {
        const int nb = sph ? (1 + 3 + 5 + 7 + 1 + 3) : (1 + 3 + 6 + 10 + 1 + 3);
        std::vector<double> da((size_t)4 * nb * nb), db((size_t)4 * nb * nb);
        for (size_t i = 0; i < da.size(); ++i) { da[i] = std::cos(0.1 * (double)i); db[i] = std::sin(0.2 * (double)i); }
        gimic_b200_handle a = nullptr;
        gimic_b200_opts oa = o;
        oa.spherical = sph;
        EXPECT(gimic_b200_create_from_arrays(&a, 2, coords, nctr, cl, cn, xp, cc, sph, da.data(), db.data(), 0, &oa) == 0);
        EXPECT(gimic_b200_nbf(a) == nb);
        EXPECT(gimic_b200_calc_jtensors(a, 300, r.data(), GIMIC_B200_SPINDENS, tens.data(), 0) == 0);
        if (!sph) EXPECT(gimic_b200_calc_basis(a, 16, r.data(), bf.data(), dr.data(), 0) == 0);
        else { std::vector<double> sb((size_t)16 * nb), sd((size_t)48 * nb); EXPECT(gimic_b200_calc_basis(a, 16, r.data(), sb.data(), sd.data(), 0) == 0); }
        EXPECT(gimic_b200_destroy(a) == 0);
        oa.uhf = 1;
        EXPECT(gimic_b200_create_from_arrays(&a, 2, coords, nctr, cl, cn, xp, cc, 0, da.data(), nullptr, 0, &oa) == GIMIC_B200_EINVAL && a == nullptr);   // beta missing
    }
    const int bad_l[6] = {0, 1, 2, 9, 0, 1};
    gimic_b200_handle a = nullptr;
    EXPECT(gimic_b200_create_from_arrays(&a, 2, coords, nctr, bad_l, cn, xp, cc, 0, tens.data(), nullptr, 0, &o) == GIMIC_B200_EINVAL);
    EXPECT(gimic_b200_destroy(h) == 0);
    // ---- host-only entry points
    int info[5];
    EXPECT(gimic_b200_mol_summary(argv[1], info) == 0 && info[0] == 8 && info[2] == 168 && info[3] == 1);
    char sym[16]; double gxyz[24];
    EXPECT(gimic_b200_mol_geometry(argv[1], 8, gxyz, sym) == 8);
    std::vector<double> po(21 * 11);
    for (int l = 0; l <= 5; ++l) EXPECT(gimic_b200_c2s_rows(l, l & 1, po.data()) == 0);
    EXPECT(gimic_b200_c2s_rows(6, 0, po.data()) == GIMIC_B200_EINVAL);
    std::vector<char> txt(4096);
    EXPECT(gimic_b200_format_e(9, r.data(), 14, 6, 4, 1, nullptr, txt.data(), 4096) > 0);
    EXPECT(gimic_b200_format_f(8, r.data(), 11, 7, 4, 0, "  ", txt.data(), 4096) > 0);
    EXPECT(gimic_b200_format_e(9, r.data(), 14, 6, 4, 1, nullptr, txt.data(), 10) == GIMIC_B200_EINVAL);
    const std::string bin = std::string(argv[2]) + ".harness.bin";
    EXPECT(gimic_b200_convert_xdens(argv[2], 168, 4, bin.c_str()) == 0);
    o.uhf = 0;
    EXPECT(gimic_b200_create(&h, argv[1], bin.c_str(), &o) == 0);          // the binary cache is recognised
    EXPECT(gimic_b200_destroy(h) == 0);
    std::remove(bin.c_str());
    // ---- legacy boundary
    gimic_init(argv[1], argv[2]);
    double jt9[9], jv3[3], mj, thr = 1e-7;
    int uhf0 = 0;
    gimic_set_magnet(B); gimic_set_spin("total"); gimic_set_screening(&thr); gimic_set_uhf(&uhf0);
    gimic_calc_jtensor(r.data(), jt9); gimic_calc_jvector(r.data(), jv3); gimic_calc_modj(r.data(), &mj);
    double a0 = 0.0, b0 = 2.0; int np = 18, ord = 9; double gp[18], gw[18];
    gimic_get_gauss_points(&a0, &b0, &np, &ord, gp, gw); mkgausspoints(&a0, &b0, &np, &ord, gp, gw);
    gimic_finalize();
    std::printf("api harness: %d failure(s)\n", failures);
    return failures ? 1 : 0;
}
