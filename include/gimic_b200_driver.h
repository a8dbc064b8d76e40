/* gimic_b200_driver.h -- C ABI of the native run-mode driver (libgimic_b200_driver.so, host-only C++17).
 *
 * What it replaces: the reference's executable pair -- the `gimic` front end (src/gimic.in:25-159: parse gimic.inp with
 * getkw, validate with check_top/check_grid :161-283, hand the keywords to gimic.bin) and `program gimic`
 * (src/fgimic/gimic.F90:7-58 main, :60-261 initialize / driver / run_cdens / run_integral), together with the grid set-up
 * (grid.f90, magnet.f90), the writers (vtkplot.f90, jfield.f90:250-443) and the report formats (integral.f90:167-183,
 * 306-322,502-510; jfield.f90:584-929).  Same gimic.inp, same files in the work directory (mol.xyz, grid.xyz, jvec*.vti,
 * jmod*.vti, jmod*.txt, acid.vti, jvec.vtu, sigma*.vtu, intchi*.vtu), same stdout report.
 *
 * The driver is a CLIENT of include/gimic_b200.h: every number on the hot path comes out of libgimic_b200.so (CUDA, no CPU
 * fallback); this layer does input, geometry, orchestration and text.  `gimic-b200` (gimic_b200/gimic-b200) is the
 * command-line program over these entry points.
 *
 * All functions return 0 or a negative GIMIC_B200_E* code (include/gimic_b200.h); the message is available from
 * gimic_b200_driver_last_error() (thread-local).
 */
#ifndef GIMIC_B200_DRIVER_H
#define GIMIC_B200_DRIVER_H

#include "gimic_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    GIMIC_B200_RUN_DRYRUN = 1,        /* -y / dryrun=on (src/gimic.in:53-54,139-140; gimic.F90:174-185): lay out the grid, write
                                         mol.xyz / grid.xyz, print the banners, calculate nothing; needs no GPU */
    GIMIC_B200_RUN_VTK_APPENDED = 2   /* extra: .vti files with raw appended Float64 blocks instead of ASCII e14.6 */
};

/* One gimic.inp: `gimic gimic.inp > report` (src/gimic.in:116-159 + program gimic).  workdir NULL: the directory of the
 * input file (basis / xdens / grid files are resolved against it, outputs are written into it).  device: CUDA ordinal or -1.
 * report_path NULL: the report goes to stdout. */
int gimic_b200_run_input(const char *inpfile, const char *workdir, int device, int flags, const char *report_path);

/* The general form.  Zero-initialise the struct, then set what is needed (device = -1 selects the current device).
 *   ndevices > 0 (list in `devices`) or ndevices < 0 (every GPU of the node): the run uses several GPUs from ONE process -- a
 *   context per device (densities replicated), one host thread each; cdens splits the flat point index into contiguous slabs,
 *   integral mode splits the plane rows j -- the block partition of schedule() (src/fgimic/parallel.F90:66-84) -- and the <= 7
 *   partial sums are added on the host in device order.  Nothing is exchanged between the devices.  (rank / nranks below are
 *   the one-process-per-GPU form of the same run: `python -m gimic_b200` under torchrun, collectives over NCCL.)
 *   title: the -t switch of the front end (src/gimic.in:135-136), overrides the `title` keyword. */
typedef struct {
    int flags;                /* GIMIC_B200_RUN_* */
    int device;               /* CUDA ordinal for a single-device run, -1 = current device */
    int ndevices;             /* 0: single device; > 0: `devices` lists the GPUs; < 0: all GPUs of the node */
    const int *devices;
    const char *workdir;      /* NULL: the directory of the input file */
    const char *title;        /* NULL: the `title` keyword */
    const char *report_path;  /* NULL: the report goes to stdout */
    /* nranks > 1: this process is rank `rank` of `nranks` cooperating processes, one per GPU (torchrun; `python -m gimic_b200` fills
     * these in).  Every rank parses the input and builds its own context; cdens / edens evaluate the rank's equal-COST share of
     * the tiles (gimic_b200_partition_*), div J its slab of points, integral mode its slab of plane rows (schedule(),
     * parallel.F90:66-84).  The two callbacks are the only communication: allgather_rows completes a [n_total][ncols] array of
     * which this rank holds `count` rows (row numbers in `index`, values in `rows`) on every rank; allreduce_sum adds v[0..n) over
     * the ranks (the collect_sum of integral.f90:157-161).  Both return 0 or a negative code.  Rank 0 alone writes the report and
     * the files (like the reference's MPI path, jfield.f90:90-137). */
    int rank, nranks;
    int (*allgather_rows)(void *user, long n_total, int ncols, long count, const long *index, const double *rows, double *full);
    int (*allreduce_sum)(void *user, double *v, int n);
    void *user;
} gimic_b200_run_opts;
int gimic_b200_run(const char *inpfile, const gimic_b200_run_opts *opts);

/* A current-profile scan (jobscripts/src/current-profile-local-submit: `gimic gimic.N.inp > gimic.N.out` for every slice):
 * inputs that agree on basis, densities and Advanced settings share ONE device context, and all their plane integrals go
 * through ONE tensor pass per spin case (gimic_b200_integrate_batch).  Each report is written to <input stem>.out, and
 * current_profile.dat (slice position, net / diatropic / paratropic current in nA/T; the table jobscripts/src/gradient.sh.in:38-47
 * pastes together from the reports) next to the first input. */
int gimic_b200_run_scan(int n, const char *const *inpfiles, int device, int flags);

/* Writers alone (vtkplot.f90:14-391, jfield.f90:356-376,531-541): lay `data` out on the grid that gimic.inp describes and
 * write it as `filename` in the work directory.  kind: "vti_scalar" (n = npoints), "vti_vector" (3 x npoints, with the
 * CellData block and the radius mask of 2-D bond grids), "jmod_txt" (3 x npoints), "vtu_vector" / "vtu_scalar" (needs
 * grid.1.ele).  For a caller that computed a field through the batched C ABI itself (e.g. the reference's Fortran loops). */
int gimic_b200_write_field(const char *inpfile, const char *workdir, const char *kind, const double *data, long n,
                           const char *filename, int flags);

/* The grid and the field direction that gimic.inp describes (grid.f90 new_grid: std / base / bond / file grids, even / gauss /
 * lobatto axes, rotation; magnet.f90 get_magnet incl. the bond-grid default and the check_field reversal), for a caller that feeds
 * the batched C ABI itself (gimic_b200_grid takes exactly these arrays).  Host only.  pts / wgt (each may be NULL) receive the axis
 * coordinates and weights, axis 0 first (npts[0] + npts[1] + npts[2] values); for a file grid pts receives the 3 x npoints
 * coordinates instead.  Call with pts = wgt = NULL first to size them. */
typedef struct {
    int is_file;              /* 1: Grid(file), explicit points */
    int npts[3];
    long npoints;
    double origin[3];
    double basv[9];           /* basv[3*v + c]: component c of basis vector v (orthonormal) */
    double lengths[3];
    double magnet[3];         /* unit vector of the external field */
    double radius;            /* integration cut-off of bond grids (1e10: none), integral.f90:93 */
    int has_center_bond;
    double center_bond[3];
} gimic_b200_grid_info;
int gimic_b200_input_grid(const char *inpfile, const char *workdir, gimic_b200_grid_info *info, double *pts, double *wgt, long cap);

/* Parse the text XDENS named in gimic.inp once (threaded) and write the binary cache <xdens>.bin beside it, sized from the input's
 * basis / openshell / Advanced.spherical settings; point `xdens=` at the .bin afterwards (gimic_b200_create recognises it).  Replaces
 * the 4 x nbf^2 (8 x for open shell) list-directed reads of read_dens (src/libgimic/dens.f90:56-135) on every later run.  Host only.
 * `written` (may be NULL) receives the path of the cache file. */
int gimic_b200_cache_xdens(const char *inpfile, const char *workdir, char *written, int cap);

const char *gimic_b200_driver_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GIMIC_B200_DRIVER_H */
