// Output files in the reference's formats: src/fgimic/vtkplot.f90 (ImageData / UnstructuredGrid, ASCII e14.6 / e20.10),
// jmod.txt (jfield.f90:356-364,540), mol.xyz (basis.f90 write_xyz) and grid.xyz (grid.f90:586-672).  Headers that the
// reference writes list-directed are emitted with gfortran's spacing; the bulk number blocks go through the library's
// threaded formatter (gimic_b200_format_e / _format_f).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "native_driver.hpp"

namespace gbd {

const double AU2A = (double)0.52917726f;     // globals.f90:51 is a single-precision literal

namespace {

std::string fmt(const char *f, ...) __attribute__((format(printf, 1, 2)));
std::string fmt(const char *f, ...) {
    char buf[256];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

struct OutFile {
    FILE *f;
    explicit OutFile(const std::string &path) : f(std::fopen(path.c_str(), "wb")) { if (!f) throw DriverError("cannot write " + path); }
    ~OutFile() { if (f) std::fclose(f); }
    void w(const std::string &s) { if (!s.empty() && std::fwrite(s.data(), 1, s.size(), f) != s.size()) throw DriverError("write failed"); }
    void raw(const void *p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) throw DriverError("write failed"); }
};

std::string rjust(const std::string &s, size_t w) { return s.size() >= w ? s : std::string(w - s.size(), ' ') + s; }

void vti_geometry(const GridSpec &g, Vec3 &qmin, Vec3 &step) {                 // vtkplot.f90:33-38
    qmin = g.gridpoint(0, 0, 0);
    const Vec3 qmax = g.gridpoint(g.npts[0] - 1, g.npts[1] - 1, g.npts[2] - 1);
    for (int i = 0; i < 3; ++i) {
        step[i] = qmax[i] - qmin[i];
        if (step[i] > 1e-8) step[i] = step[i] / (double)(g.npts[i] - 1);
    }
}

void vti_header(OutFile &f, const GridSpec &g, const Vec3 &qmin, const Vec3 &step, const char *name, int ncomp) {
    std::string ext;
    for (int d = 0; d < 3; ++d) ext += ld_int(0) + ld_int(g.npts[d] - 1);
    f.w("<?xml version=\"1.0\"?>\n");
    f.w(" <VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\">\n");
    // gfortran separates a numeric list item from a following character item by one blank (test/*/reference/*.vti)
    f.w("   <ImageData WholeExtent=\"" + ext + " \" Origin=\"" + ld_real(qmin[0]) + ld_real(qmin[1]) + ld_real(qmin[2]) + " \" Spacing=\"" +
        ld_real(step[0]) + ld_real(step[1]) + ld_real(step[2]) + " \">\n");
    f.w("   <Piece Extent=\"" + ext + " \">\n");
    f.w("   <PointData Scalars=\"scalars\">\n");
    f.w(fmt("   <DataArray Name=\"%s\" type=\"Float64\" NumberOfComponents=\"%d\" Format=\"ascii\">\n", name, ncomp));
}

// |J| averaged over the corners of each cell (vtkplot.f90:132-226); empty for grids without cells.  vec: [k][j][i][3]
std::vector<double> cell_average_norm(const std::vector<double> &vec, int p1, int p2, int p3) {
    auto at = [&](int k, int j, int i, int c) { return vec[(((size_t)k * p2 + j) * p1 + i) * 3 + c]; };
    const bool t1 = p1 > 1, t2 = p2 > 1, t3 = p3 > 1;
    if ((int)t1 + (int)t2 + (int)t3 < 2) return {};
    const int n1 = t1 ? p1 - 1 : 1, n2 = t2 ? p2 - 1 : 1, n3 = t3 ? p3 - 1 : 1;
    const double div = (t1 && t2 && t3) ? 8.0 : 4.0;
    std::vector<double> out((size_t)n1 * n2 * n3);
    size_t q = 0;
    for (int k = 0; k < n3; ++k) for (int j = 0; j < n2; ++j) for (int i = 0; i < n1; ++i) {
        double a[3];
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;                                    // corner order: k offset outermost, i offset innermost
            for (int dk = 0; dk <= (t3 ? 1 : 0); ++dk) for (int dj = 0; dj <= (t2 ? 1 : 0); ++dj) for (int di = 0; di <= (t1 ? 1 : 0); ++di)
                s = s + at(k + dk, j + dj, i + di, c);
            a[c] = s / div;
        }
        out[q++] = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    }
    return out;
}

// EXTRA (not a reference format): the same ImageData file with raw appended Float64 blocks instead of ASCII numbers
void vti_appended(const std::string &path, const GridSpec &g, const char *name, int ncomp, const std::vector<double> &data,
                  const std::vector<double> *cell) {
    Vec3 qmin, step;
    vti_geometry(g, qmin, step);
    std::string ext = fmt("0 %d 0 %d 0 %d", g.npts[0] - 1, g.npts[1] - 1, g.npts[2] - 1);
    std::string head = "<?xml version=\"1.0\"?>\n<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
    head += "  <ImageData WholeExtent=\"" + ext + "\" Origin=\"" + py_repr(qmin[0]) + " " + py_repr(qmin[1]) + " " + py_repr(qmin[2]) + "\" Spacing=\"" +
            py_repr(step[0]) + " " + py_repr(step[1]) + " " + py_repr(step[2]) + "\">\n";
    head += "    <Piece Extent=\"" + ext + "\">\n";
    head += std::string("      <PointData ") + (ncomp == 3 ? "Vectors" : "Scalars") + "=\"" + name + "\">\n";
    head += fmt("        <DataArray Name=\"%s\" type=\"Float64\" NumberOfComponents=\"%d\" format=\"appended\" offset=\"0\"/>\n", name, ncomp);
    head += "      </PointData>\n";
    if (cell) {
        head += "      <CellData Scalars=\"cell_norm\">\n";
        head += fmt("        <DataArray Name=\"cell_norm\" type=\"Float64\" NumberOfComponents=\"1\" format=\"appended\" offset=\"%zu\"/>\n",
                    (size_t)8 + data.size() * 8);
        head += "      </CellData>\n";
    }
    head += "    </Piece>\n  </ImageData>\n  <AppendedData encoding=\"raw\">\n_";
    OutFile f(path);
    f.w(head);
    uint64_t nb = data.size() * 8;
    f.raw(&nb, 8); f.raw(data.data(), nb);
    if (cell) { nb = cell->size() * 8; f.raw(&nb, 8); f.raw(cell->data(), nb); }
    f.w("\n  </AppendedData>\n</VTKFile>\n");
}

}  // namespace

// Shortest round-trip decimal with Python's repr() layout: fixed notation for 1e-4 <= |x| < 1e16, else d.ddde+XX
std::string py_repr(double x) {
    if (x != x) return "nan";
    if (std::isinf(x)) return x < 0 ? "-inf" : "inf";
    if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[48];
    int prec = 1;
    for (; prec <= 17; ++prec) {
        snprintf(buf, sizeof buf, "%.*e", prec - 1, x);
        if (std::strtod(buf, nullptr) == x) break;
    }
    std::string m(buf);
    const size_t epos = m.find('e');
    const int e = std::atoi(m.c_str() + epos + 1);
    if (e >= -4 && e < 16) {
        const int dec = std::max(prec - 1 - e, 1);           // digits after the point; at least ".0"
        snprintf(buf, sizeof buf, "%.*f", dec, x);
        return buf;
    }
    return m.substr(0, epos) + fmt("e%+03d", e);
}

// Fortran Ew.d: 0.dddddE+ee (gfortran drops the 'E' when the exponent needs three digits; asterisks on overflow)
std::string fortran_e(double x, int w, int d) {
    std::string s;
    if (!std::isfinite(x)) {
        s = std::isnan(x) ? "NaN" : (x < 0 ? "-Infinity" : "Infinity");
    } else if (x == 0.0) {
        s = std::string(std::signbit(x) ? "-" : "") + "0." + std::string(d, '0') + "E+00";      // gfortran keeps the sign of a negative zero
    } else {
        char buf[64];
        snprintf(buf, sizeof buf, "%.*E", d - 1, std::fabs(x));
        std::string m(buf);
        const size_t epos = m.find('E');
        int e = std::atoi(m.c_str() + epos + 1) + 1;
        std::string digits;
        for (size_t i = 0; i < epos; ++i) if (m[i] != '.') digits.push_back(m[i]);
        s = std::string(x < 0 ? "-" : "") + "0." + digits + (std::abs(e) < 100 ? fmt("E%+03d", e) : fmt("%+04d", e));
    }
    return (int)s.size() > w ? std::string(w, '*') : rjust(s, w);
}

// gfortran list-directed real(8): 17 significant digits, F form for 1e-1 <= |x| < 1e16
std::string ld_real(double x) {
    const double ax = std::fabs(x);
    char buf[64];
    // e.g. the z spacing of a tilted 2-D grid: vtkplot.f90:33-38 divides a non-zero extent by npts-1 = 0; gfortran prints
    // non-finite values right-justified in the same 25-column field
    if (x != x) return rjust("NaN", 26);
    if (std::isinf(x)) return rjust(x < 0 ? "-Infinity" : "Infinity", 26);
    if (ax != 0.0 && !(0.1 <= ax && ax < 1e16)) {
        snprintf(buf, sizeof buf, "%.16E", ax);
        std::string m(buf);
        const size_t epos = m.find('E');
        const int e = std::atoi(m.c_str() + epos + 1);
        return "  " + std::string(x < 0 ? "-" : "") + m.substr(0, epos) + fmt("E%+04d", e) + " ";
    }
    int dec = ax == 0.0 ? 16 : 17;                 // gfortran prints zero as 0.0000000000000000 (test/benzene/2d/reference/jvec.vti)
    if (ax >= 1.0) {
        snprintf(buf, sizeof buf, "%.0f", std::floor(ax));
        dec = 17 - (int)std::strlen(buf);
    }
    snprintf(buf, sizeof buf, "%.*f", dec, ax);
    return rjust(std::string(x < 0 ? "-" : "") + buf, 21) + "     ";
}

std::string ld_int(long i) { return fmt("%12ld", i); }

std::string format_block(char kind, const double *v, long n, int w, int d, int per_line, int first, const std::string &prefix) {
    if (n <= 0) return std::string();
    const long nlines = n / (per_line > 0 ? per_line : 1) + 2;
    const long cap = n * w + nlines * ((long)prefix.size() + 1) + 16;
    std::string out((size_t)cap, '\0');
    const long got = (kind == 'E' ? gimic_b200_format_e : gimic_b200_format_f)(n, v, w, d, per_line, first, prefix.c_str(), &out[0], cap);
    if (got < 0) throw DriverError("number formatting failed");
    out.resize((size_t)got);
    return out;
}

// write_vtk_imagedata, vtkplot.f90:14-86: values[i + p1*(j + p2*k)], e14.6, line break when mod(l,4)==0
void write_vti_scalar(const std::string &path, const GridSpec &g, const std::vector<double> &values, bool appended) {
    if (appended) return vti_appended(path, g, "scalars", 1, values, nullptr);
    Vec3 qmin, step;
    vti_geometry(g, qmin, step);
    OutFile f(path);
    vti_header(f, g, qmin, step, "scalars", 1);
    f.w(format_block('E', values.data(), (long)values.size(), 14, 6, 4, 1));
    f.w("\n    </DataArray>\n    </PointData>\n    </Piece>\n    </ImageData>\n </VTKFile>\n");
}

// write_vtk_vector_imagedata, vtkplot.f90:88-234: 3e14.6 per point + CellData of cell-averaged |J|
void write_vti_vector(const std::string &path, const GridSpec &g, const std::vector<double> &vec, bool appended) {
    const std::vector<double> nrm = cell_average_norm(vec, g.npts[0], g.npts[1], g.npts[2]);
    if (appended) return vti_appended(path, g, "vectors", 3, vec, nrm.empty() ? nullptr : &nrm);
    Vec3 qmin, step;
    vti_geometry(g, qmin, step);
    OutFile f(path);
    vti_header(f, g, qmin, step, "vectors", 3);
    f.w(format_block('E', vec.data(), (long)vec.size(), 14, 6, 3));
    f.w("\n    </DataArray>\n    </PointData>\n    <CellData Scalars=\"foo\">\n");
    if (!nrm.empty()) f.w(format_block('E', nrm.data(), (long)nrm.size(), 14, 6, 1));
    f.w("    </CellData>\n    </Piece>\n    </ImageData>\n </VTKFile>\n");
}

// cdens with the `radius` keyword (jfield.f90:310-346): on 2-D bond grids the vectors written to jvec.vti are zeroed where
// |coord*AU2A - center| > radius.  The reference compares Angstrom coordinates with the bohr centre and radius (unit slip,
// SURVEY A.10); replicated as is.
std::vector<double> radius_masked_vectors(const GridSpec &g, const std::vector<double> &vec) {
    if (g.mode != "bond" || g.is_3d() || !(g.radius > 0.1) || g.radius >= 1.0e10) return vec;
    const std::vector<double> r = g.points();
    const Vec3 c = g.center();
    std::vector<double> out = vec;
    for (size_t p = 0; p < out.size() / 3; ++p) {
        const double dx = r[3 * p] * AU2A - c[0], dy = r[3 * p + 1] * AU2A - c[1], dz = r[3 * p + 2] * AU2A - c[2];
        if (std::sqrt(dx * dx + dy * dy + dz * dz) > g.radius) out[3 * p] = out[3 * p + 1] = out[3 * p + 2] = 0.0;
    }
    return out;
}

// TetGen .ele: first line 'ncells 4 0', then 'idx n1 n2 n3 n4' (jfield.f90:421-431)
std::vector<long> read_ele(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw DriverError("cannot open " + path);
    std::string line;
    std::getline(f, line);
    long ncells = std::atol(line.c_str());
    std::vector<long> cells;
    cells.reserve((size_t)ncells * 4);
    for (long c = 0; c < ncells; ++c) {
        long idx, n[4];
        if (!(f >> idx >> n[0] >> n[1] >> n[2] >> n[3])) throw DriverError("short read in " + path);
        for (int k = 0; k < 4; ++k) cells.push_back(n[k]);
    }
    return cells;
}

// write_vtk_vector_unstructuredgrid / write_vtk_scalar_unstructuredgrid, vtkplot.f90:241-391
void write_vtu(const std::string &path, const std::vector<double> &points, const std::string &name, int ncomp, const std::vector<double> &data,
               const std::vector<long> &cells) {
    const long np = (long)points.size() / 3, nc = (long)cells.size() / 4;
    OutFile f(path);
    f.w("<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <UnstructuredGrid>\n");
    f.w(fmt("    <Piece NumberOfPoints=\"%10ld\" NumberOfCells=\"%10ld\">\n      <Points>\n", np, nc));
    f.w("        <DataArray type=\"Float32\" NumberOfComponents=\"3\" Format=\"ascii\">\n");
    f.w(format_block('E', points.data(), np * 3, 20, 10, 3, 0, "        "));
    f.w("        </DataArray>\n      </Points>\n      <PointData Scalars=\"scalars\">\n");
    f.w(fmt("        <DataArray Name=\"%s\" type=\"Float64\" NumberOfComponents=\"%d\" Format=\"ascii\">\n", name.c_str(), ncomp));
    f.w(format_block('E', data.data(), np * ncomp, 20, 10, ncomp, 0, "        "));
    f.w("        </DataArray>\n      </PointData>\n      <Cells>\n        <DataArray type=\"Int32\" Name=\"connectivity\" Format=\"ascii\">\n");
    std::string s;
    for (long c = 0; c < nc; ++c) {
        s += "        ";
        for (int k = 0; k < 4; ++k) s += fmt("%10ld", cells[4 * c + k] - 1);
        s += "\n";
    }
    f.w(s);
    f.w("        </DataArray>\n        <DataArray type=\"Int32\" Name=\"offsets\" Format=\"ascii\">\n        ");
    s.clear();
    for (long c = 0; c < nc; ++c) s += fmt("%10ld", 4 * (c + 1));
    f.w(s + "\n");
    f.w("        </DataArray>\n        <DataArray type=\"Int32\" Name=\"types\" Format=\"ascii\">\n        ");
    s.clear();
    for (long c = 0; c < nc; ++c) s += fmt("%5d", 10);
    f.w(s + "\n");
    f.w("        </DataArray>\n      </Cells>\n      <CellData Scalars=\"foo\">\n        ");
    s.clear();
    for (long c = 0; c < nc; ++c) s += " 0.0";
    f.w(s + "\n");
    f.w("      </CellData>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n");
}

// jmod<tag>.txt on Gauss grids (jfield.f90:294-301,356-376,531-541): '(6f11.7)' of coord*AU2A and |J|; a blank line after
// each i-row on regular grids
void write_jmod_txt(const std::string &path, const GridSpec &g, const std::vector<double> &vec, bool regular) {
    const std::vector<double> r = g.points();
    const long n = (long)vec.size() / 3;
    std::vector<double> rows((size_t)n * 4);
    for (long p = 0; p < n; ++p) {
        for (int c = 0; c < 3; ++c) rows[4 * p + c] = r[3 * p + c] * AU2A;
        const double x = vec[3 * p], y = vec[3 * p + 1], z = vec[3 * p + 2];
        rows[4 * p + 3] = std::sqrt(x * x + y * y + z * z);
    }
    const std::string txt = format_block('F', rows.data(), n * 4, 11, 7, 4);          // n lines of 44 characters + newline
    OutFile f(path);
    if (!regular) { f.w(txt); return; }
    const long p1 = g.npts[0];
    const size_t row = 45;
    for (long k = 0; k < n; k += p1) {
        const long m = std::min(p1, n - k);
        f.raw(txt.data() + (size_t)k * row, (size_t)m * row);
        if (k + p1 <= n) f.w("\n");
    }
}

namespace {
// F16.10; a coordinate that is not finite (grid_points = 1 along a non-zero length: step = l/0, grid.f90:157) prints like gfortran
std::string f16(double x) {
    if (std::isfinite(x)) return fmt("%16.10f", x);
    return rjust(std::isnan(x) ? "NaN" : (x < 0 ? "-Infinity" : "Infinity"), 16);
}
std::string xyz_line(const std::string &sym, const double *c) {
    return sym + f16(c[0] * AU2A) + f16(c[1] * AU2A) + f16(c[2] * AU2A) + "\n";
}
}  // namespace

// write_xyz (basis.f90): natoms, blank, 'sym x y z' in Angstrom
void write_mol_xyz(const std::string &path, const std::vector<std::string> &symbols, const std::vector<double> &coords) {
    OutFile f(path);
    f.w(fmt("%12zu\n\n", symbols.size()));
    for (size_t a = 0; a < symbols.size(); ++a) f.w(xyz_line(symbols[a], &coords[3 * a]));
}

// plot_grid_xyz, grid.f90:586-672: atoms + grid corners ('X') + field-direction marker ('Be')
void write_grid_xyz(const std::string &path, const GridSpec &g, const std::vector<std::string> &symbols, const std::vector<double> &coords) {
    const int p1 = g.npts[0], p2 = g.npts[1], p3 = g.npts[2];
    std::vector<Vec3> corners;
    // file grids take the p3 == 1 branch like every grid with npts = (n, 1, 1): gridpoint(i,j,k) = xdata(:,i), so the four 'X' lines are the
    // first and the last point twice, and the count line says natoms+5 although no 'Be' line follows (no case for 'file', grid.f90:660-667)
    if (p3 > 1) {
        const int idx[8][3] = {{0, 0, 0}, {p1 - 1, 0, 0}, {0, p2 - 1, 0}, {0, 0, p3 - 1}, {p1 - 1, p2 - 1, 0}, {p1 - 1, 0, p3 - 1},
                               {0, p2 - 1, p3 - 1}, {p1 - 1, p2 - 1, p3 - 1}};
        for (auto &i : idx) corners.push_back(g.gridpoint(i[0], i[1], i[2]));
    } else {
        const int idx[4][3] = {{0, 0, 0}, {p1 - 1, 0, 0}, {0, p2 - 1, 0}, {p1 - 1, p2 - 1, 0}};
        for (auto &i : idx) corners.push_back(g.gridpoint(i[0], i[1], i[2]));
    }
    bool has_marker = false;
    Vec3 marker{{0, 0, 0}};
    if (g.mode == "std" || g.mode == "base") { has_marker = true; for (int c = 0; c < 3; ++c) marker[c] = g.origin[c] + g.basv[2][c] * 2.0; }
    else if (g.mode == "bond") { has_marker = true; for (int c = 0; c < 3; ++c) marker[c] = g.origin[c] + g.ortho[c] * 2.0; }
    OutFile f(path);
    f.w(fmt("%12zu\n\n", symbols.size() + corners.size() + 1));
    for (size_t a = 0; a < symbols.size(); ++a) f.w(xyz_line(symbols[a], &coords[3 * a]));
    for (auto &c : corners) f.w(xyz_line("X ", c.data()));
    if (has_marker) f.w(xyz_line("Be ", marker.data()));
}

// edens / divj on grids that are not images: 'x y z value' rows (bohr), %20.12e
void write_points_txt(const std::string &path, const std::vector<double> &r, const std::vector<double> &v) {
    OutFile f(path);
    std::string s;
    for (size_t p = 0; p < v.size(); ++p) {
        s += fmt("%20.12e %20.12e %20.12e %20.12e\n", r[3 * p], r[3 * p + 1], r[3 * p + 2], v[p]);
        if (s.size() > (1u << 20)) { f.w(s); s.clear(); }
    }
    f.w(s);
}

}  // namespace gbd
