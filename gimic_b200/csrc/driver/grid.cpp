// Grid geometry and magnetic-field direction: compiled counterpart of src/fgimic/grid.f90 and magnet.f90.
// std/base grids (grid.f90:140-163), bond grids (:165-276), file grids (:543-576), even / gauss / lobatto point distributions
// (:291-373), rotation (:697-769), grid_center (:529-541), get_magnet / check_field (magnet.f90:11-86).
// Quadrature nodes come from the library (gimic_b200_gauss_points = setup_gauss_data, gaussint.f90:267-319).
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "native_driver.hpp"

namespace gbd {

namespace {

const double PII = 3.141592653589793;   // globals.f90:41
const double TINY = 2.2250738585072014e-308;

double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
Vec3 cross(const Vec3 &a, const Vec3 &b) { return Vec3{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}}; }
Vec3 unit(const Vec3 &v) { double n = std::sqrt(dot(v.data(), v.data())); return Vec3{{v[0] / n, v[1] / n, v[2] / n}}; }
Vec3 sub(const Vec3 &a, const Vec3 &b) { return Vec3{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
bool any(const Vec3 &v) { return v[0] != 0.0 || v[1] != 0.0 || v[2] != 0.0; }
double nint(double x) { return std::round(x); }   // Fortran NINT: halves away from zero (also for negative arguments)

// R = R_x . R_y . R_z with the sign conventions of grid.f90:716-757
void rotation_matrix(const Vec3 &deg, double R[3][3]) {
    const double rx = deg[0] / 180.0 * PII, ry = deg[1] / 180.0 * PII, rz = deg[2] / 180.0 * PII;
    const double Rz[3][3] = {{std::cos(rz), std::sin(rz), 0.0}, {-std::sin(rz), std::cos(rz), 0.0}, {0.0, 0.0, 1.0}};
    const double Ry[3][3] = {{std::cos(ry), 0.0, -std::sin(ry)}, {0.0, 1.0, 0.0}, {std::sin(ry), 0.0, std::cos(ry)}};
    const double Rx[3][3] = {{1.0, 0.0, 0.0}, {0.0, std::cos(rx), std::sin(rx)}, {0.0, -std::sin(rx), std::cos(rx)}};
    double T[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0.0; for (int k = 0; k < 3; ++k) s += Ry[i][k] * Rz[k][j]; T[i][j] = s; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0.0; for (int k = 0; k < 3; ++k) s += Rx[i][k] * T[k][j]; R[i][j] = s; }
}
Vec3 matvec(const double R[3][3], const Vec3 &v) {
    Vec3 o;
    for (int i = 0; i < 3; ++i) { double s = 0.0; for (int k = 0; k < 3; ++k) s += R[i][k] * v[k]; o[i] = s; }
    return o;
}

std::string f3(const double *v, int w = 12, int d = 6) {
    char buf[128];
    snprintf(buf, sizeof buf, "%*.*f%*.*f%*.*f", w, d, v[0], w, d, v[1], w, d, v[2]);
    return buf;
}
std::string i3(long a, long b, long c) {
    char buf[64];
    snprintf(buf, sizeof buf, "%5ld%5ld%5ld", a, b, c);
    return buf;
}

void make_axes(GridSpec &g, const Vec3 &step, const AxisOpts &o) {
    const Vec3 &l = g.lengths;
    if (o.gtype == "even") {                                  // setup_even_grid, grid.f90:351-373
        for (int d = 0; d < 3; ++d) {
            long n;
            if (d == 2 && (std::fabs(l[2]) < TINY || std::fabs(step[2]) < TINY)) n = 1;
            else n = (long)nint(l[d] / step[d]) + 1;
            if (n < 1) throw DriverError("Grid has no points along axis " + std::to_string(d + 1));
            g.pts[d].resize(n); g.wgt[d].assign(n, 1.0);
            for (long i = 0; i < n; ++i) g.pts[d][i] = (double)i * step[d];
        }
        return;
    }
    if (o.gtype != "gauss" && o.gtype != "lobatto") throw DriverError("Unknown grid type: " + o.gtype);
    if (o.gauss_order < 1) throw DriverError("gauss_order must be positive");
    long npts[3];                                             // setup_gauss_grid, grid.f90:291-349
    bool fixed = false;
    g.log.push_back(" INFO: Integration grid selected.");
    for (int d = 0; d < 3; ++d) {
        if (o.has_grid_points) npts[d] = o.grid_points[d];
        else if (std::fabs(o.spacing[d]) < 1e-10 || o.spacing[d] < 0.0) npts[d] = 0;
        else npts[d] = (long)nint(l[d] / o.spacing[d]);
        if (!(npts[d] > 1)) npts[d] = 0;
        long rem = npts[d] % o.gauss_order;
        if (rem != 0) { npts[d] = npts[d] - rem + o.gauss_order; fixed = true; }
    }
    if (fixed) g.log.push_back(" INFO: Adjusted number of grid points for quadrature: " + i3(npts[0], npts[1], npts[2]));
    for (int d = 0; d < 3; ++d) {
        const long n = npts[d] > 0 ? npts[d] : 1;
        g.pts[d].assign(n, 0.0); g.wgt[d].assign(n, 0.0);
        int rc = gimic_b200_gauss_points(0.0, l[d], (int)n, npts[d] > 0 ? o.gauss_order : 1, o.gtype == "lobatto" ? 1 : 0, g.pts[d].data(),
                                         g.wgt[d].data());
        if (rc) throw DriverError(gimic_b200_last_error());
    }
}

GridSpec finish(const Vec3 &origin_in, const Vec3 b_in[3], const Vec3 &lengths, const std::string &mode, const Vec3 &ortho, double radius,
                const Vec3 &step, const AxisOpts &o, double out_len, double down_len, const Vec3 *center_bond) {
    GridSpec g;
    g.mode = mode; g.gtype = o.gtype; g.gauss = (o.gtype == "gauss" || o.gtype == "lobatto");
    g.ortho = ortho; g.lengths = lengths; g.radius = radius;
    if (center_bond) { g.has_center_bond = true; g.center_bond = *center_bond; }
    for (int v = 0; v < 3; ++v) for (int c = 0; c < 3; ++c) g.basv[v][c] = b_in[v][c];
    auto normalise = [&]() {                                  // grid.f90:278-288
        for (int v = 0; v < 3; ++v) {
            double n = std::sqrt(dot(g.basv[v], g.basv[v]));
            if (n > 0.0) for (int c = 0; c < 3; ++c) g.basv[v][c] = g.basv[v][c] / n;
        }
    };
    if (mode == "bond") {                                     // the block setup_bond_grid prints, grid.f90:258-275 (before any rotation)
        g.log.push_back("");
        g.log.push_back(" Integration grid data");
        g.log.push_back(" " + std::string(48, '-'));
        g.log.push_back(" center " + f3(center_bond ? center_bond->data() : origin_in.data()));
        g.log.push_back(" origin " + f3(origin_in.data()));
        g.log.push_back(" basv1  " + f3(g.basv[0]));
        g.log.push_back(" basv2  " + f3(g.basv[1]));
        g.log.push_back(" basv3  " + f3(g.basv[2]));
        g.log.push_back(" lenghts" + f3(lengths.data()));
        g.log.push_back(" magnet " + f3(ortho.data()));
        g.log.push_back("");
    }
    g.log.push_back(" Grid mode = " + mode);
    normalise();
    if (std::fabs(dot(g.basv[0], g.basv[1])) > 1e-10) {      // ortho_coordsys, grid.f90:400-428
        Vec3 t = cross(Vec3{{g.basv[0][0], g.basv[0][1], g.basv[0][2]}}, Vec3{{g.basv[2][0], g.basv[2][1], g.basv[2][2]}});
        double n = std::sqrt(dot(t.data(), t.data()));
        for (int c = 0; c < 3; ++c) g.basv[1][c] = t[c] / n;
        normalise();
    }
    g.origin = origin_in;
    if (o.has_rotation) {                                     // grid.f90:92-114
        Vec3 ref;
        if (o.has_rotation_origin) ref = o.rotation_origin;
        else for (int c = 0; c < 3; ++c) ref[c] = g.origin[c] + out_len * g.basv[1][c] + down_len * g.basv[0][c];
        double R[3][3];
        rotation_matrix(o.rotation, R);
        for (int v = 0; v < 3; ++v) {
            Vec3 r = matvec(R, Vec3{{g.basv[v][0], g.basv[v][1], g.basv[v][2]}});
            for (int c = 0; c < 3; ++c) g.basv[v][c] = r[c];
        }
        Vec3 r = matvec(R, sub(g.origin, ref));
        for (int c = 0; c < 3; ++c) g.origin[c] = r[c] + ref[c];
        const double rad[3] = {o.rotation[0] / 180.0 * PII, o.rotation[1] / 180.0 * PII, o.rotation[2] / 180.0 * PII};
        g.log.push_back(" INFO: Rotation is: " + f3(rad, 9, 5));
    }
    make_axes(g, step, o);
    for (int d = 0; d < 3; ++d) g.npts[d] = (int)g.pts[d].size();
    char buf[96];
    g.log.push_back("   Number of grid points <v1,v2>:" + i3(g.npts[0], g.npts[1], g.npts[2]));
    snprintf(buf, sizeof buf, "   Total number of grid points  :%10ld", g.n());
    g.log.push_back(buf);
    g.log.push_back("");
    return g;
}

}  // namespace

Vec3 GridSpec::gridpoint(int i, int j, int k) const {
    if (is_file()) return Vec3{{xdata[3 * (size_t)i], xdata[3 * (size_t)i + 1], xdata[3 * (size_t)i + 2]}};
    Vec3 r;
    for (int c = 0; c < 3; ++c) r[c] = origin[c] + pts[0][i] * basv[0][c] + pts[1][j] * basv[1][c] + pts[2][k] * basv[2][c];
    return r;
}

Vec3 GridSpec::center() const {
    Vec3 a = gridpoint(npts[0] - 1, 0, 0), b = gridpoint(0, npts[1] - 1, 0);
    return Vec3{{(a[0] + b[0]) * 0.5, (a[1] + b[1]) * 0.5, (a[2] + b[2]) * 0.5}};
}

std::vector<double> GridSpec::points() const {
    if (is_file()) return xdata;
    std::vector<double> r((size_t)n() * 3);
    size_t q = 0;
    for (int k = 0; k < npts[2]; ++k) for (int j = 0; j < npts[1]; ++j) for (int i = 0; i < npts[0]; ++i) {
        Vec3 p = gridpoint(i, j, k);
        r[q++] = p[0]; r[q++] = p[1]; r[q++] = p[2];
    }
    return r;
}

gimic_b200_grid GridSpec::cstruct() const {
    gimic_b200_grid g;
    for (int i = 0; i < 3; ++i) {
        g.origin[i] = origin[i]; g.npts[i] = npts[i]; g.pts[i] = pts[i].data(); g.wgt[i] = wgt[i].data();
        for (int c = 0; c < 3; ++c) g.basv[c + 3 * i] = basv[i][c];
    }
    g.radius = radius;
    return g;
}

// setup_std_grid, grid.f90:140-163
GridSpec std_grid(const Vec3 &origin, const Vec3 &ivec, const Vec3 &jvec, const Vec3 &lengths, const AxisOpts &o, const std::string &mode) {
    Vec3 step;
    for (int d = 0; d < 3; ++d) step[d] = o.has_spacing ? o.spacing[d] : lengths[d] / (double)(o.grid_points[d] - 1);
    const Vec3 b3 = cross(ivec, jvec);
    const Vec3 b[3] = {ivec, jvec, b3};
    return finish(origin, b, lengths, mode, unit(b3), -1.0, step, o, 0.0, 0.0, nullptr);
}

// setup_bond_grid, grid.f90:165-276 (height/width as in the input file; the first entries are negated, :212-213)
GridSpec bond_grid(const Vec3 &c1, const Vec3 &c2, const Vec3 &fix, double distance, const std::array<double, 2> &height,
                   const std::array<double, 2> &width, const AxisOpts &o, bool has_radius, double radius, bool has_magnet, const Vec3 &magnet) {
    const double hgt[2] = {-height[0], height[1]}, wdt[2] = {-width[0], width[1]};
    const Vec3 lengths{{hgt[0] + hgt[1], wdt[0] + wdt[1], 0.0}};
    if (wdt[0] + wdt[1] < 0.0 || hgt[0] + hgt[1] < 0.0) throw DriverError("Grid width/height < 0!");
    Vec3 v1 = sub(c1, fix), v2 = sub(c2, fix);
    Vec3 ortho = cross(v1, v2);
    if (!any(ortho)) throw DriverError("Basis vectors are linearly dependent, field direction undefined!");
    ortho = unit(ortho);
    const Vec3 v3 = unit(sub(v2, v1));
    v1 = Vec3{{-ortho[0], -ortho[1], -ortho[2]}};
    v2 = unit(cross(v3, v1));
    Vec3 oo, origin;
    for (int c = 0; c < 3; ++c) { oo[c] = c1[c] + distance * v3[c]; origin[c] = oo[c] - wdt[1] * v2[c] - hgt[1] * v1[c]; }
    if (has_magnet) ortho = unit(magnet);                     // top-level `magnet` overrides the plane normal, :253-256
    const double rad = has_radius ? radius : (double)1.0e10f; // 1.e10 is a real(4) literal, grid.f90:199
    // Reference quirk, replicated: setup_bond_grid never reads Grid.spacing / Grid.grid_points into grid%step, so an *even* bond
    // grid always has the default step of 1 bohr (new_grid sets step=1.d0, grid.f90:66; test/benzene/int-grid-bond-even has
    // 11 x 8 points for grid_points=[40,40,0]).  Gauss grids do use grid_points.
    const Vec3 step{{1.0, 1.0, 1.0}};
    const Vec3 b[3] = {v1, v2, v3};
    return finish(origin, b, lengths, "bond", ortho, rad, step, o, width[1], height[1], &oo);
}

// extgrid, grid.f90:543-576: an explicit point list; basis vectors are zero (so get_magnet never flips B)
GridSpec file_grid(std::vector<double> xyz) {
    GridSpec g;
    g.mode = "file"; g.gtype = "file"; g.gauss = false;
    xyz.resize(xyz.size() / 3 * 3);
    g.xdata = std::move(xyz);
    for (int d = 0; d < 3; ++d) { g.pts[d].assign(1, 0.0); g.wgt[d].assign(1, 1.0); }
    g.npts[0] = (int)(g.xdata.size() / 3); g.npts[1] = g.npts[2] = 1;
    char buf[96];
    snprintf(buf, sizeof buf, "   Total number of grid points  :%10ld", g.n());   // extgrid, grid.f90:572-575
    g.log.push_back(buf);
    g.log.push_back("");
    return g;
}

// get_magnet + check_field, magnet.f90:11-86
Vec3 get_magnet(const GridSpec &g, const std::string &magnet_axis, const Vec3 &magnet, std::vector<std::string> *log) {
    std::string axis = magnet_axis;
    while (!axis.empty() && std::isspace((unsigned char)axis.front())) axis.erase(axis.begin());
    while (!axis.empty() && std::isspace((unsigned char)axis.back())) axis.pop_back();
    bool ortho = false;
    double d = 1.0;
    Vec3 mag{{0, 0, 0}};
    if (!axis.empty()) {
        if (axis[0] == '-') { d = -1.0; axis = axis.substr(1); }
        const char a = axis.empty() ? '\0' : axis[0];
        if (a == 'i' || a == 'j' || a == 'k') { const int v = a - 'i'; for (int c = 0; c < 3; ++c) mag[c] = g.basv[v][c] * d; }
        else if (a == 'x' || a == 'y' || a == 'z') { for (int c = 0; c < 3; ++c) mag[c] = (c == a - 'x' ? 1.0 : 0.0) * d; }   // (/D0,D0,D1/)*dir, magnet.f90:38-43
        else if (a == 'X') { ortho = true; for (int c = 0; c < 3; ++c) mag[c] = g.ortho[c] * d; }
        else throw DriverError("Invalid axis specifier: " + axis);
    } else {
        mag = magnet;
    }
    if (!any(mag)) throw DriverError("Magnetic field is zero, not wasting more CPU.");
    if (!ortho) {                                             // check_field, magnet.f90:66-86
        const double x = dot(g.basv[2], mag.data());
        if (x > 0.0) {
            for (int c = 0; c < 3; ++c) mag[c] = -mag[c];
            if (log) log->push_back(" INFO: Left handed coordinate system, reversing magnetic field");
        }
        if (log) {
            if (std::fabs(x) - 1.0 > 1e-12 && std::fabs(x) > 1e-12) log->push_back(" WARNING: Magnetic field not orthogonal to grid");
            char buf[96];
            snprintf(buf, sizeof buf, "    Magnetic field <x,y,z> =%10.5f%10.5f%10.5f", mag[0], mag[1], mag[2]);
            log->push_back(buf);
            log->push_back("");
        }
    }
    return mag;
}

std::vector<double> read_numbers(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw DriverError("cannot open " + path);
    std::vector<double> v;
    std::string line;
    while (std::getline(f, line)) {
        size_t h = line.find('#');
        if (h != std::string::npos) line.resize(h);
        const char *p = line.c_str();
        char *end = nullptr;
        while (true) {
            double x = std::strtod(p, &end);
            if (end == p) break;
            v.push_back(x);
            p = end;
        }
    }
    return v;
}

// new_grid, grid.f90:50-138
GridSpec grid_from_input(const Input &inp, const std::vector<double> &atom_coords, const std::string &workdir) {
    auto S = [&](const char *k) { return inp.is_set(std::string("Grid.") + k); };
    const std::string mode = inp.grid_arg;
    if (mode == "file") {
        const std::string path = join_path(workdir, S("file") ? inp.str("Grid.file") : std::string("GRIDDATA"));
        return file_grid(read_numbers(path));
    }
    AxisOpts o;
    o.gtype = inp.str("Grid.type");
    o.gauss_order = (int)inp.integer("Grid.gauss_order");
    if (S("rotation")) { o.has_rotation = true; o.rotation = inp.vec3("Grid.rotation"); }
    if (S("rotation_origin")) { o.has_rotation_origin = true; o.rotation_origin = inp.vec3("Grid.rotation_origin"); }
    if (S("grid_points")) {
        const auto &gp = inp.integers("Grid.grid_points");
        if (gp.size() < 3) throw InputError("'Grid.grid_points' needs three values");
        o.has_grid_points = true; o.grid_points = {{gp[0], gp[1], gp[2]}};
    }
    if (S("spacing")) { o.has_spacing = true; o.spacing = inp.vec3("Grid.spacing"); }
    const long natoms = (long)atom_coords.size() / 3;
    auto atom = [&](long idx1) {
        if (idx1 < 1 || idx1 > natoms) throw DriverError("atom index " + std::to_string(idx1) + " out of range in the Grid section");
        return Vec3{{atom_coords[3 * (idx1 - 1)], atom_coords[3 * (idx1 - 1) + 1], atom_coords[3 * (idx1 - 1) + 2]}};
    };
    if (mode == "std" || mode == "base")
        return std_grid(inp.vec3("Grid.origin"), inp.vec3("Grid.ivec"), inp.vec3("Grid.jvec"), inp.vec3("Grid.lengths"), o, mode);
    if (mode == "bond") {
        Vec3 c1, c2;
        if (S("bond")) {
            const auto &b = inp.integers("Grid.bond");
            if (b.size() < 2) throw InputError("'Grid.bond' needs two atom indices");
            c1 = atom(b[0]); c2 = atom(b[1]);
        } else { c1 = inp.vec3("Grid.coord1"); c2 = inp.vec3("Grid.coord2"); }
        const Vec3 fix = S("fixpoint") ? atom(inp.integer("Grid.fixpoint")) : inp.vec3("Grid.fixcoord");
        const auto &h = inp.reals("Grid.height"); const auto &w = inp.reals("Grid.width");
        if (h.size() < 2 || w.size() < 2) throw InputError("'Grid.height' and 'Grid.width' need two values");
        return bond_grid(c1, c2, fix, inp.real("Grid.distance"), {{h[0], h[1]}}, {{w[0], w[1]}}, o, S("radius"), inp.real("Grid.radius"),
                         inp.is_set("magnet"), inp.is_set("magnet") ? inp.vec3("magnet") : Vec3{{0, 0, 0}});
    }
    throw DriverError("Unknown grid type: " + mode);
}

}  // namespace gbd
