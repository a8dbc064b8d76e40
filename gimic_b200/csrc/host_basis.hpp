// Host-side molecule / basis / density ingestion for gimic-b200.
//
// Produces the flat structure-of-arrays shell tables the CUDA kernels consume.  Replaces, for the
// grid hot path, the reference's  src/libgimic/intgrl.f90 (MOL parser), basis.f90 (normalisation,
// screening radii), gtodefs.f90 (cartesian component tables), reorder.f90 + dens.f90 (XDENS reader,
// Turbomole permutation, UHF halving).
#pragma once
#include <string>
#include <vector>

namespace gb {

constexpr int MAX_L = 5;                 // src/libgimic/globals.f90:30
constexpr int MAX_SHELLS_PER_ATOM = 99;  // posvec(99), src/libgimic/bfeval.f90:86

struct Shell {            // one segmented contraction, reference order (atom -> file order)
    int atom, l, nprim, prim_off, ncomp;
    int user_off;         // first function index in the reference's AO order
    int nsph, sph_off;    // spherical=on: 2l+1 components and first index in the reference's SAO order
    double thr;           // screening radius (basis.f90:90-112); 1e10 when screening is off
};

struct HostBasis {
    bool turbomole = false;              // line 2 of MOL == "TURBOMOLE" (intgrl.f90:47-53)
    bool spherical = false;              // Advanced.spherical: densities are given over 2l+1 components per shell
    int natoms = 0, nbf = 0, nprim_total = 0, ngto = 0;
    int nbf_sph = 0;                     // sum of 2l+1 (get_ncgto of the reference when spherical=on)
    std::vector<double> xyz;             // 3*natoms
    std::vector<double> charge;
    std::vector<std::string> symbol;
    std::vector<Shell> shells;           // reference order
    std::vector<int> atom_shell_off;     // natoms+1
    std::vector<int> atom_func_off;      // natoms+1, reference AO order
    std::vector<double> alpha, cc, ncc;  // primitives, shell-major
};

// Parses an INTGRL/MOL file.  Returns false and sets err on failure.
bool parse_mol(const std::string &path, HostBasis &b, std::string &err);
// From flat arrays (synthetic benchmarks / bindings that already hold the basis).
bool basis_from_arrays(int natoms, const double *coords, const int *nctr_per_atom, const int *ctr_l,
                       const int *ctr_npf, const double *xp, const double *cc, int turbomole_order,
                       HostBasis &b, std::string &err);
// Normalised contraction coefficients (basis.f90:164-191) and screening radii (basis.f90:90-112).
void finalize_basis(HostBasis &b, bool use_screening, double screening_thrs);

// cartesian exponents of component c of a shell with angular momentum l, in the reference's
// standard (gtodefs.f90:86-106) or Turbomole (gtodefs.f90:109-123) component order
void component_exponents(int l, bool turbomole, int c, int lmn[3]);

// XDENS: one number per line; nmat = 4 (closed shell) or 8 (UHF) matrices of nbf*nbf values,
// element (a,b) at a + nbf*b.  Returns them in file order.
// The text is parsed by all host threads; a file starting with "GB2XDENS" is the binary cache written by
// write_xdens_binary (header: magic, int64 nbf, int64 nmat; then the same values as raw doubles).
bool read_xdens(const std::string &path, int nbf, int nmat, std::vector<double> &out, std::string &err);
bool write_xdens_binary(const std::string &path, int nbf, int nmat, const double *vals, std::string &err);
// Permutation of reorder.f90:54-96: sv[i] = atom-major index of the i-th function in Turbomole's
// "all s, all p, ..." AO order, so that new(sv[i], sv[j]) = old(i, j).
void turbomole_permutation(const HostBasis &b, std::vector<int> &sv);

// Same permutation over the spherical components (spherical=on: reorder.f90 runs on the 2l+1 counts).
void turbomole_permutation_sph(const HostBasis &b, std::vector<int> &sv);

// Cartesian -> spherical projection rows of cao2sao.f90:163-231 for one angular momentum: po[(m+l)*ncart + c],
// m = -l..l, c in this molecule's cartesian component order.  Integer-valued, not normalised (like the reference).
void c2s_rows(int l, bool turbomole, std::vector<double> &po);
// spherical=on: dcart = po^T dsph po, with po = blockdiag(c2s_rows(l_shell)) (bfeval.f90:116-118 applies po to the basis
// vectors at every point; applying it to the densities once is the same bilinear form).  Both column-major.
void density_sph_to_cart(const HostBasis &b, const double *dsph, double *dcart);

// Bulk number formatting for the output files (vtkplot.f90 writes every value with a Fortran Ew.d edit descriptor; at 256^3
// points that is 5e7 values).  Lines hold `per_line` values (the first line `first_count` if > 0), each line starts with
// `prefix` and complete lines end with '\n'.  Threaded; returns the number of bytes written, or -1 if `cap` is too small.
long format_fortran_e(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap);
// same with kind = 'E' (Ew.d) or 'F' (Fw.d)
long format_fortran(long n, const double *v, char kind, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap);

// Gauss-Legendre / Lobatto nodes in the piecewise-block layout of setup_gauss_data
// (src/libgimic/gaussint.f90:267-319).  quadrature: 0 = gauss, 1 = lobatto.
int gauss_blocks(double a, double b, int npts, int order, int quadrature, double *pts, double *wgts);

}  // namespace gb
