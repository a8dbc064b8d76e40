/* gimic_b200.h -- C ABI of the B200-native GIMIC grid hot path (libgimic_b200.so).
 *
 * Two groups of entry points:
 *
 *  (1) LEGACY symbols: signature-identical to the reference's C boundary, so existing callers
 *      (GimicInterface.cpp, gimic.pyx, gengauss.pyx, pygimic) link unchanged:
 *        src/libgimic/gimic_interface.h:9-18   (Fortran side: gimic_interface.f90:26-163)
 *        src/libgimic/gausspoints.h:10-13      (Fortran side: gausspoints.f90:13-29, gausspoints.c:4-7)
 *      They operate on one hidden default context, one point per call, and -- like the Fortran
 *      `stop` they replace -- terminate the process on error after printing the message.
 *
 *  (2) BATCHED, handle-based API (gimic_b200_*): what a driver should call.  It replaces the
 *      Fortran-internal boundary the reference's grid loops use:
 *        new_jtensor/ctensor/del_jtensor   src/libgimic/jtensor.F90:39-103
 *        calc_jtensors / compute_jvectors  src/fgimic/jfield.f90:62-184
 *        jmod2_vtkplot / acid_vtkplot      src/fgimic/jfield.f90:446-489, 556-582 (field arithmetic only)
 *        integrate_current/_modulus/_acid  src/fgimic/integral.f90:50-511
 *      All functions return 0 on success or a negative GIMIC_B200_E* code; the message is
 *      available from gimic_b200_last_error() (thread-local).  Buffers are caller-owned, plain
 *      pointers; host memory unless GIMIC_B200_DEVICE_PTR is set in `flags`, in which case they are
 *      device pointers on the handle's GPU and no host<->device copy is made.
 *
 * Conventions (identical to the reference):
 *   r      3 x n, point-major: r[3*i + c]                      (bohr)
 *   tens   9 x n, point-major: tens[9*i + m + 3*b] = dJ_m/dB_b  (jtensor.F90:66-103, column-major 3x3)
 *   jvec   3 x n: jvec[3*i + m] = sum_b tens(m,b) B_b           (jfield.f90:167-184)
 *   densities: XDENS layout, column-major nbf x nbf, element (a,b) at a + nbf*b, matrices in the
 *              order D, P_x, P_y, P_z (alpha) then the same four for beta (dens.f90:79-90)
 *   AO order: atom -> contraction (file order) -> cartesian component (gtodefs.f90:86-123)
 */
#ifndef GIMIC_B200_H
#define GIMIC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ (1) legacy boundary -------- */
void gimic_init(const char *mol, const char *xdens);      /* gimic_interface.h:9  */
void gimic_finalize(void);                                /* gimic_interface.h:10 */
void gimic_set_uhf(int *uhf);                             /* gimic_interface.h:11 */
void gimic_set_magnet(const double *b3);                  /* gimic_interface.h:12 */
void gimic_set_spin(const char *spincase);                /* gimic_interface.h:13 */
void gimic_set_screening(const double *thrs);             /* gimic_interface.h:14 */
void gimic_calc_jtensor(const double *r3, double *jt9);   /* gimic_interface.h:15 */
void gimic_calc_jvector(const double *r3, double *jv3);   /* gimic_interface.h:16 */
void gimic_calc_modj(const double *r3, double *modj);     /* gimic_interface.h:17 (reference: STOP 'NOT IMPLEMENTED') */
void gimic_get_gauss_points(double *a, double *b, int *npts, int *order, double *pts, double *wgts); /* gausspoints.h:10 */
void mkgausspoints(double *a, double *b, int *npts, int *order, double *pts, double *wgts);          /* gausspoints.h:12 */

/* ------------------------------------------------------------------ (2) batched API ------------ */
typedef struct gimic_b200_ctx *gimic_b200_handle;

enum {
    GIMIC_B200_OK = 0,
    GIMIC_B200_EINVAL = -1,   /* bad argument */
    GIMIC_B200_EIO = -2,      /* MOL / XDENS unreadable or malformed */
    GIMIC_B200_ECUDA = -3,    /* CUDA runtime error (no CPU fallback exists) */
    GIMIC_B200_ESPIN = -4,    /* beta/spindens requested on a closed-shell context (jtensor.F90:74-96) */
    GIMIC_B200_ENOMEM = -5
};

enum { GIMIC_B200_ALPHA = 0, GIMIC_B200_BETA = 1, GIMIC_B200_TOTAL = 2, GIMIC_B200_SPINDENS = 3 };

enum { GIMIC_B200_DEVICE_PTR = 1 };   /* flags: r / output buffers are device pointers */

typedef struct {
    int uhf;                 /* open shell: XDENS holds 8 matrices, P_b halved on read (dens.f90:94-98) */
    int giao;                /* Advanced.GIAO     (jtensor.F90:115,180,194,210) */
    int diamag;              /* Advanced.diamag   (jtensor.F90:225) */
    int paramag;             /* Advanced.paramag  (jtensor.F90:220) */
    int screening;           /* Advanced.screening */
    double screening_thrs;   /* Advanced.screening_thrs; gimic_init uses 1e-6 (globals.f90:56) */
    int device;              /* CUDA device ordinal; -1 = current device */
    int spherical;           /* Advanced.spherical: XDENS / density arrays are over 2l+1 components per shell in the
                                reference's (non-normalised, m = -l..l) convention of cao2sao.f90; cartesian when 0 */
} gimic_b200_opts;

/* defaults of gimic_init (gimic_interface.f90:39-51): closed shell, GIAO/diamag/paramag on,
 * screening on with 1e-6, current device */
void gimic_b200_default_opts(gimic_b200_opts *opts);

/* Reads MOL (INTGRL format, intgrl.f90) and XDENS (dens.f90), applies the Turbomole reorder and the
 * UHF halving exactly as read_dens does, and uploads basis tables + contraction operands. */
int gimic_b200_create(gimic_b200_handle *h, const char *mol, const char *xdens, const gimic_b200_opts *opts);

/* Same, from memory.  Shell arrays are flat in AO order; dens_alpha/dens_beta point to 4 matrices
 * each in the XDENS layout, already in atom-major AO order and (for UHF) already halved -- i.e. the
 * contents of dens_t%da / dens_t%db (dens.f90:13-18).  dens_beta may be NULL for closed shell.
 * dens_flags: GIMIC_B200_DEVICE_PTR if the density pointers are device memory. */
int gimic_b200_create_from_arrays(gimic_b200_handle *h, int natoms, const double *coords, const int *nctr_per_atom,
                                  const int *ctr_l, const int *ctr_npf, const double *xp, const double *cc,
                                  int turbomole_order, const double *dens_alpha, const double *dens_beta,
                                  int dens_flags, const gimic_b200_opts *opts);
int gimic_b200_destroy(gimic_b200_handle h);

/* Number of CUDA devices visible to the process (for callers that place one context per GPU), or GIMIC_B200_ECUDA. */
int gimic_b200_device_count(void);

int gimic_b200_nbf(gimic_b200_handle h);
int gimic_b200_natoms(gimic_b200_handle h);
int gimic_b200_atom_coords(gimic_b200_handle h, double *xyz /* 3 x natoms */);
int gimic_b200_is_uhf(gimic_b200_handle h);

/* calc_jtensors (jfield.f90:62-138): tens(:,i) = ctensor(r_i, spincase) for n points. */
int gimic_b200_calc_jtensors(gimic_b200_handle h, long n, const double *r, int spincase, double *tens, int flags);

/* Tensors + derived fields in one pass.  Any output may be NULL (skipped).
 *   jvec  = T.B                       (jfield.f90:167-184)
 *   jmod  = signed |J|                (jfield.f90:446-489: sign of (B x (r - (B.r)B)) . J)
 *   acid  = get_acid(T)               (acid.f90:9-45, with the reference's 0.3333333)
 *   edens = Phi^T D Phi ("diapam", jtensor.F90:168)        -- no reference run mode at this commit
 *   divj  = div(T.B) by central differences of step divj_h  -- no reference run mode at this commit */
/* When only jvec and/or jmod (and edens) are requested (tens, acid, divj NULL) the tensor is never formed: the contraction
 * runs with the operand pair (D, sum_b B_b P_b) -- compute_jvectors (jfield.f90:167-184) applied before instead of after
 * the GEMM -- which halves the tensor-core work.  Same J within the 1e-10 / 1e-12 tolerance. */
int gimic_b200_calc_fields(gimic_b200_handle h, long n, const double *r, const double *B3, int spincase,
                           double *tens, double *jvec, double *jmod, double *acid, double *edens, double *divj,
                           double divj_h, int flags);

/* Basis vectors themselves (calc_basis / bfeval / dfdr, src/libgimic/bfeval.f90:61-122,295-338), reference AO order,
 * exact zeros where a contraction is screened:  bf[i*nbf + f] = Phi_f(r_i),  dr[(3*i + m)*nbf + f] = dPhi_f/dr_m.
 * Either output may be NULL.  (The London/GIAO vectors db, d2 are products of these with r x R_A, bfeval.f90:168-293.) */
int gimic_b200_calc_basis(gimic_b200_handle h, long n, const double *r, double *bf, double *dr, int flags);

/* The same vectors evaluated by the HOT-PATH kernels (spatial sort, tiles, the k_basis panels the contraction consumes) and scattered
 * back into the dense layout of gimic_b200_calc_basis: a diagnostic that lets tests hold the panel kernel itself to bfeval.f90.
 * Host buffers, cartesian contexts; tile_info3 (may be NULL) = { tiles, mean padded active slots per tile, max active atoms of a tile }. */
int gimic_b200_calc_basis_tiles(gimic_b200_handle h, long n, const double *r, double *bf, double *dr, int *tile_info3);

/* Field arithmetic alone on existing tensors (the HBM-bound pass). */
int gimic_b200_fields_from_tensors(gimic_b200_handle h, long n, const double *r, const double *tens,
                                   const double *B3, double *jvec, double *jmod, double *acid, int flags);

/* Signed modulus from J alone (jfield.f90:446-489), e.g. for J vectors combined by linearity (total = alpha + beta). */
int gimic_b200_jmod_from_jvec(gimic_b200_handle h, long n, const double *r, const double *jvec, const double *B3, double *jmod, int flags);

/* Regular grid (grid_t of src/fgimic/grid.f90:19-32 reduced to what gridpoint/get_weight need):
 * r(i,j,k) = origin + pts[0][i] basv(:,1) + pts[1][j] basv(:,2) + pts[2][k] basv(:,3)  (grid.f90:498-511) */
typedef struct {
    double origin[3];
    double basv[9];          /* basv[c + 3*v] = component c of basis vector v */
    int npts[3];
    const double *pts[3];    /* host arrays, npts[d] each */
    const double *wgt[3];    /* quadrature weights (1.0 on even grids), host arrays */
    double radius;           /* integration bound around grid_center (integral.f90:91,128); <=0 or >=1e10: none */
} gimic_b200_grid;

/* Tensors on the flat index range [lo, hi) of a regular grid (i fastest, grid.f90:478-495);
 * points are generated on the device.  tens holds (hi-lo) tensors. */
int gimic_b200_calc_jtensors_grid(gimic_b200_handle h, const gimic_b200_grid *g, long lo, long hi, int spincase,
                                  double *tens, int flags);

/* Cost-balanced multi-GPU partition (replaces schedule(), src/fgimic/parallel.F90:66-84, which gives every rank an equal COUNT of
 * consecutive flat indices, jfield.f90:90-104).  Every rank calls _partition_points / _partition_grid with the SAME complete point
 * set and its own rank: the points are sorted along a Hilbert curve and tiled identically on every rank, and rank r is given the run
 * of tiles whose cumulative cost (~ active functions^2, integers) lies in [r, r+1) x total / nranks.  *count = points this rank owns.
 * _partition_calc then evaluates them: output row i belongs to caller point index[i] (flat grid index for _partition_grid); any output
 * may be NULL; only jvec / jmod / edens requested => the tensor is never formed (see gimic_b200_calc_fields).  The plan stays valid for
 * further _partition_calc calls (other spin cases, other fields) until the next compute call on the handle.  Tiles, and therefore every
 * result bit, are those of a single-rank run.  index / outputs: host, or device with GIMIC_B200_DEVICE_PTR.
 * _partition_info: { points, owned points, tiles, owned tiles, total cost, owned cost, panel batches, first owned tile }. */
int gimic_b200_partition_points(gimic_b200_handle h, long n, const double *r, int flags, int rank, int nranks, long *count);
int gimic_b200_partition_grid(gimic_b200_handle h, const gimic_b200_grid *g, int rank, int nranks, long *count);
int gimic_b200_partition_calc(gimic_b200_handle h, const double *B3, int spincase, long *index, double *tens, double *jvec,
                              double *jmod, double *acid, double *edens, int flags);
int gimic_b200_partition_info(gimic_b200_handle h, long *info8);

/* Plane/volume quadrature of integral.f90 over rows j in [jlo, jhi) of the grid (all i, all k):
 *   out[0..2] = sum, positive part, negative part of  w (n.T.B)        (integrate_current)
 *   out[3..5] = same for the signed modulus sgn(n.J)|J|               (integrate_modulus)
 *   out[6]    = sum of w * get_acid(T)  (caller takes sqrt after reducing; integrate_acid)
 * `what` is a bit mask: 1 current, 2 modulus, 4 acid.  With jlo=0, jhi=npts[1] this is the whole
 * integral; a multi-GPU driver gives each rank a slab of rows and sums out[] (one all-reduce). */
int gimic_b200_integrate(gimic_b200_handle h, const gimic_b200_grid *g, const double *B3, int spincase, int what,
                         int jlo, int jhi, double *out7);

/* The same quadrature for ngrids grids in one tensor pass (all rows of each grid): the points of all planes go through the
 * sort/tile/basis/contraction pipeline together.  This is what a current-profile scan (jobscripts/src/current-profile-*:
 * hundreds of gimic.N.inp integrals over thin slices of one plane) should call instead of ngrids separate runs.
 * B3s: 3 doubles per grid; out7s: 7 doubles per grid, laid out as in gimic_b200_integrate. */
int gimic_b200_integrate_batch(gimic_b200_handle h, int ngrids, const gimic_b200_grid *grids, const double *B3s, int spincase,
                               int what, double *out7s);

/* get_property (src/fgimic/jfield.f90:584-929): shielding and magnetizability quadrature of an existing tensor field on a
 * weighted point set (NumGrid: r = gridfile.grd, w = grid_w.grd, coords = coord.au, segments = the per-atom grid blocks of
 * nelpts.info as cumulative end indices).  part[(k*nseg + s)*5 + q]: for nucleus k (k == natoms: magnetizability) and point
 * segment s the sums  q=0,1,2: w * integrand_xx,yy,zz  (sigma in ppm: 1e6 * (-1/|d|^3/c^2) (d x J_b)_b, chi: 1/2 (r x J_b)_b,
 * J_b = T.(-e_b)),  q=3 / q=4: positive / negative part of w*(xx+yy+zz).  The reference's running totals and per-atom
 * contributions are prefix sums of these over s.  r, w, tens: host or device (flags); coords, seg_end, part: host. */
int gimic_b200_property(gimic_b200_handle h, long n, const double *r, const double *w, const double *tens, int natoms,
                        const double *coords, int nseg, const long *seg_end, double *part, int flags);

/* The per-point integrands of get_property for one centre -- what the reference writes as sigma<k>.vtu, sigma_xx<k>.vtu, ... and
 * intchi*.vtu (jfield.f90:786-808, 915-918): out4[4*i + 0..2] = integrand_xx,yy,zz at point i, out4[4*i + 3] = their sum.
 * centre3 = nucleus position (shielding, ppm) or NULL (magnetizability, au). */
int gimic_b200_property_integrand(gimic_b200_handle h, long n, const double *r, const double *tens, const double *centre3,
                                  double *out4, int flags);

/* Gauss-Legendre (quadrature=0) / Lobatto (1) nodes in the block layout of setup_gauss_data
 * (gaussint.f90:267-319); host only. */
int gimic_b200_gauss_points(double a, double b, int npts, int order, int quadrature, double *pts, double *wgts);

/* Geometry of an INTGRL/MOL file without creating a context (host only; e.g. for a dry run that only lays out the grid):
 * returns the number of atoms (or a negative GIMIC_B200_E* code) and fills up to max_atoms entries of xyz (3 per atom, bohr)
 * and symbols2 (the two-character element field of intgrl.f90:111, not NUL-terminated); either may be NULL. */
int gimic_b200_mol_geometry(const char *mol, int max_atoms, double *xyz, char *symbols2);

/* What new_basis prints about a MOL file (basis.f90:44-62, intgrl.f90:48-60), host only: info5 = { number of atoms, total number of
 * primitive GTOs (sum of npf x components), number of contracted cartesian GTOs, 1 if the file says TURBOMOLE, number of
 * contracted spherical GTOs (the dimension of the XDENS matrices when Advanced.spherical is on) }. */
int gimic_b200_mol_summary(const char *mol, int *info5);

/* XDENS text (dens.f90:129-135: one real per line, 4 or 8 matrices of nbf x nbf) -> binary cache that gimic_b200_create
 * recognises by its "GB2XDENS" magic (int64 nbf, int64 nmat, raw doubles in file order).  Values are stored exactly as the
 * text reader parses them, before UHF halving / Turbomole reordering.  nbf = the dimension of the matrices in the file
 * (spherical count when the file is over spherical components).  Host only. */
int gimic_b200_convert_xdens(const char *xdens_text, int nbf, int nmat, const char *xdens_binary);

/* Bulk formatting of n doubles with the Fortran edit descriptor Ew.d (vtkplot.f90 writes e14.6 / e20.10 for every value; at
 * 256^3 points text formatting, not the GPU, is the wall-clock bottleneck).  Lines hold per_line values (the first line
 * first_count values if first_count > 0), start with `prefix` (may be NULL) and end with a newline when complete.  Threaded
 * over the host cores.  Returns the number of bytes written to out (capacity cap), or a negative GIMIC_B200_E* code. */
long gimic_b200_format_e(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap);
/* Same for the fixed-point descriptor Fw.d (jmod.txt: '(6f11.7)', jfield.f90:540). */
long gimic_b200_format_f(long n, const double *v, int w, int d, int per_line, int first_count, const char *prefix, char *out, long cap);

/* Cartesian -> spherical projection of cao2sao.f90:163-231 for angular momentum l (0..5), as used when opts.spherical is
 * set: po[(m + l) * ncart + c], m = -l..l, c in the standard (turbomole_order = 0) or Turbomole cartesian component order;
 * integer-valued rows, bug-compatible with the reference (see host_basis.cpp); host only. */
int gimic_b200_c2s_rows(int l, int turbomole_order, double *po);

/* Last-call statistics for benches / roofline accounting. */
typedef struct {
    long n_points;            /* points processed */
    long n_tiles;             /* point tiles */
    double sum_nact;          /* sum over tiles of the padded active-function count */
    double executed_flops;    /* FP64 flops actually issued, summed over tiles: DMMA 2*128*planes*nact*nn (padded K slots x padded N columns,
                                 all 128 rows of a tile) + GIAO-tap DFMA 2*128*(3|1)*nn*active_atoms */
    double dense_flops;       /* n_points * (14 nbf^2 + 56 nbf), the reference's algorithmic work */
    float ms_sort, ms_tiles, ms_basis, ms_contract, ms_fields;   /* CUDA-event times per stage (profiling on) */
    float ms_total;           /* CUDA-event time of the whole call on the library stream, copies included (calc_fields/calc_jtensors) */
    long launches;            /* kernel launches issued */
    long contract_launches;   /* launches of the contraction kernel (k_jtensor) among them */
    double useful_flops;      /* the same count without padding: real points of each tile x its active functions on both sides
                                 (2*npts*planes*nreal^2 + taps); executed - useful = work spent on K/N padding and partial tiles */
    float ms_plan;            /* gimic_b200_partition_*: CUDA-event time of the plan (point generation / upload, sort, tiles, partition) */
    float ms_span;            /* from the start of the last gimic_b200_partition_* call to the end of gimic_b200_partition_calc */
    double panel_bytes;       /* bytes of basis-function panels k_basis wrote (4 planes x padded K slots x 132 doubles per tile) */
} gimic_b200_stats;
int gimic_b200_get_stats(gimic_b200_handle h, gimic_b200_stats *out);
int gimic_b200_set_profiling(gimic_b200_handle h, int enable);  /* per-stage CUDA-event timing (adds syncs) */

const char *gimic_b200_last_error(void);
const char *gimic_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GIMIC_B200_H */
