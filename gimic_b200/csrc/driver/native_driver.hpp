// Native (C++17, host-only) driver above the C ABI: the compiled counterpart of the reference's `program gimic`
// (src/fgimic/gimic.F90), grid.f90, magnet.f90, vtkplot.f90 and the report formats of integral.f90 / jfield.f90,
// plus the keyword schema and cross-checks of the front end (src/gimic.in:59-283).
//
// Everything in this directory calls ONLY the public entry points of include/gimic_b200.h: the arithmetic of the hot path
// stays behind the C ABI (CUDA, no CPU fallback); what lives here is input, geometry, orchestration and file formats.
#pragma once
#include <array>
#include <cstdio>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/gimic_b200.h"

namespace gbd {

using Vec3 = std::array<double, 3>;

struct InputError : std::runtime_error { using std::runtime_error::runtime_error; };
struct DriverError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---------------------------------------------------------------------------------------------- input (src/gimic.in)
enum class Type { STR, INT, DBL, BOOL, INT_ARRAY, DBL_ARRAY };

struct Value {
    Type type = Type::STR;
    bool none = true;                 // no default and not set
    std::string s;
    bool b = false;
    long i = 0;
    double x = 0.0;
    std::vector<long> iv;
    std::vector<double> dv;
};

class Input {
  public:
    Input();
    std::string grid_arg = "std";     // grid.set_arg('STR', ('std',)), src/gimic.in:92
    bool grid_present = false;

    bool is_set(const std::string &path) const { return set_.count(path) != 0; }
    const Value &get(const std::string &path) const;
    std::string str(const std::string &path) const { return get(path).s; }
    bool flag(const std::string &path) const { return get(path).b; }
    long integer(const std::string &path) const { return get(path).i; }
    double real(const std::string &path) const { return get(path).x; }
    const std::vector<double> &reals(const std::string &path) const { return get(path).dv; }
    const std::vector<long> &integers(const std::string &path) const { return get(path).iv; }
    Vec3 vec3(const std::string &path) const;

    void assign(const std::string &sect, const std::string &key, const std::vector<std::string> &raw, bool is_list);
    void force_flag(const std::string &path, bool v);       // command-line overrides (src/gimic.in:139-140)
    void force_str(const std::string &path, const std::string &v);

  private:
    std::map<std::string, Value> values_;
    std::set<std::string> set_;
};

Input parse_text(const std::string &text);
Input parse_file(const std::string &path);

// ---------------------------------------------------------------------------------------------- grids (grid.f90, magnet.f90)
struct GridSpec {
    Vec3 origin{{0, 0, 0}};
    double basv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // basv[v][c]: component c of basis vector v
    std::vector<double> pts[3], wgt[3];
    int npts[3] = {1, 1, 1};
    double radius = -1.0;
    std::string mode = "std", gtype = "even";
    bool gauss = false;
    Vec3 ortho{{0, 0, 0}}, lengths{{0, 0, 0}};
    bool has_center_bond = false;
    Vec3 center_bond{{0, 0, 0}};
    std::vector<double> xdata;        // file grids: explicit points, 3 per point
    std::vector<std::string> log;     // what new_grid prints while it sets the grid up (grid.f90:87,131-137,259-275,301,328-332,710-711)

    long n() const { return (long)npts[0] * npts[1] * npts[2]; }
    bool is_file() const { return mode == "file"; }
    bool is_3d() const { return npts[0] > 1 && npts[1] > 1 && npts[2] > 1; }
    Vec3 gridpoint(int i, int j, int k) const;           // 0-based gridpoint(), grid.f90:498-511
    Vec3 center() const;                                 // grid_center, grid.f90:529-541
    std::vector<double> points() const;                  // 3 x n, i fastest (get_grid_index, grid.f90:478-495)
    gimic_b200_grid cstruct() const;                     // pointers into this object: keep it alive
};

struct AxisOpts {
    std::string gtype = "even";
    bool has_spacing = false, has_grid_points = false, has_rotation = false, has_rotation_origin = false;
    Vec3 spacing{{0, 0, 0}};
    std::array<long, 3> grid_points{{0, 0, 0}};
    int gauss_order = 7;
    Vec3 rotation{{0, 0, 0}}, rotation_origin{{0, 0, 0}};
};

GridSpec std_grid(const Vec3 &origin, const Vec3 &ivec, const Vec3 &jvec, const Vec3 &lengths, const AxisOpts &o, const std::string &mode);
GridSpec bond_grid(const Vec3 &c1, const Vec3 &c2, const Vec3 &fix, double distance, const std::array<double, 2> &height,
                   const std::array<double, 2> &width, const AxisOpts &o, bool has_radius, double radius, bool has_magnet, const Vec3 &magnet);
GridSpec file_grid(std::vector<double> xyz);
// log (optional) receives what check_field prints (magnet.f90:66-86): nothing for the 'X' specifier, else the notes about a reversed /
// non-orthogonal field and the 'Magnetic field <x,y,z>' line
Vec3 get_magnet(const GridSpec &g, const std::string &magnet_axis, const Vec3 &magnet, std::vector<std::string> *log = nullptr);
GridSpec grid_from_input(const Input &inp, const std::vector<double> &atom_coords, const std::string &workdir);
std::vector<double> read_numbers(const std::string &path);   // whitespace-separated reals, '#' comments (np.loadtxt-like)

// ---------------------------------------------------------------------------------------------- writers (vtkplot.f90 & co)
extern const double AU2A;                                     // real(4) literal 0.52917726, globals.f90:51
std::string fortran_e(double x, int w, int d);
std::string ld_real(double x);                                // gfortran list-directed real(8)
std::string ld_int(long i);
std::string py_repr(double x);                             // shortest round-trip decimal, laid out like Python's repr(float)
std::string format_block(char kind, const double *v, long n, int w, int d, int per_line, int first = 0, const std::string &prefix = "");
void write_vti_scalar(const std::string &path, const GridSpec &g, const std::vector<double> &values, bool appended);
void write_vti_vector(const std::string &path, const GridSpec &g, const std::vector<double> &vec, bool appended);
std::vector<double> radius_masked_vectors(const GridSpec &g, const std::vector<double> &vec);
std::vector<long> read_ele(const std::string &path);          // 4 node indices (1-based) per cell
void write_vtu(const std::string &path, const std::vector<double> &points, const std::string &name, int ncomp,
               const std::vector<double> &data, const std::vector<long> &cells);
void write_jmod_txt(const std::string &path, const GridSpec &g, const std::vector<double> &vec, bool regular);
void write_mol_xyz(const std::string &path, const std::vector<std::string> &symbols, const std::vector<double> &coords);
void write_grid_xyz(const std::string &path, const GridSpec &g, const std::vector<std::string> &symbols, const std::vector<double> &coords);
void write_points_txt(const std::string &path, const std::vector<double> &r, const std::vector<double> &v);   // edens/divj on non-image grids

// ---------------------------------------------------------------------------------------------- driver (gimic.F90)
double au2si(double au);                                      // globals.f90:309-332
bool file_exists(const std::string &path);
std::string join_path(const std::string &dir, const std::string &name);
std::string dirname_of(const std::string &path);

struct RunOptions {
    bool dryrun = false;         // -y
    std::string title;           // -t: overrides the `title` keyword when non-empty (src/gimic.in:135-136)
    bool vtk_appended = false;   // --vtk appended (extra)
    int device = -1;
    std::vector<int> devices;    // more than one entry: single-process multi-device run (one context per listed GPU, one host
                                 // thread each); point slabs / plane rows are split like schedule() of parallel.F90:66-84
    std::string workdir;         // default: the directory of the input file
    // one process per GPU (gimic_b200_run_opts::rank / nranks): the two collectives of the run, supplied by the launcher
    int rank = 0, nranks = 1;
    int (*allgather_rows)(void *user, long n_total, int ncols, long count, const long *index, const double *rows, double *full) = nullptr;
    int (*allreduce_sum)(void *user, double *v, int n) = nullptr;
    void *user = nullptr;
};

// one gimic.inp: returns 0 or a negative GIMIC_B200_E* code (message from last_error_message())
int run_input(const std::string &inpfile, const RunOptions &opt, FILE *out);
// many inputs sharing contexts, integrals batched per context (jobscripts/src/current-profile-*): reports go to <stem>.out
int run_scan(const std::vector<std::string> &inpfiles, const RunOptions &opt);
// writers alone: lay `data` out on the grid of a gimic.inp (kind: vti_scalar | vti_vector | jmod_txt | vtu_vector | vtu_scalar)
int write_field(const std::string &inpfile, const std::string &workdir, const std::string &kind, const double *data, long n, const std::string &name,
                bool appended);
int cache_xdens(const std::string &inpfile, const std::string &workdir, std::string &written);   // XDENS text -> <xdens>.bin
const std::string &last_error_message();

}  // namespace gbd
