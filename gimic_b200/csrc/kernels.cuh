// Shared declarations for the gimic-b200 CUDA kernels (sm_100a).
//
// Data flow of one batched tensor evaluation (replaces the per-point loop of
// src/fgimic/jfield.f90:114-129 + src/libgimic/jtensor.F90:66-237 + bfeval.f90:61-338):
//
//   points --k_morton_keys/sort/k_gather--> spatially sorted points, tiles of MT=128 points
//   k_tile_count      per tile: bounding sphere + conservative active-function count (screening)
//   k_basis           per tile: Phi, dPhi/dx,dy,dz of the active functions -> K-major panels in HBM/L2
//   k_jtensor         per tile: DMMA contraction of the Phi panel with the gathered density
//                     operands [D | Px | Py | Pz]; the three London/GIAO products sum_mu Phi_mu D[mu,nu] (R_nu - R_mu)_d are
//                     taken from the running D accumulator at atom boundaries of the K loop (no extra GEMM planes); fused
//                     epilogue that reduces against Phi / dPhi to the 3x3 tensor (never stores X)
//   k_fields          T -> jvec, signed |J|, ACID (HBM-bound pass)
//   k_quad_rows/final Gauss-Legendre plane quadrature
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace gb {

constexpr int MT = 128;    // points per tile (8 m16 row blocks, one per warp of k_jtensor)
constexpr int LDP = 132;   // panel row stride in doubles: 132 = 4 (mod 16) -> conflict-free A-fragment LDS.64
constexpr int BK = 32;     // K slots per pipeline stage (the last stage of a K sweep may hold 16)
constexpr int NV = 16;     // nu slots per accumulator chunk (2 n8 tiles per operand matrix)
constexpr int LDB2 = 36;   // smem row stride (doubles) of a pair-plane B tile: 16 nu x 2 + 4 pad = 18 16-byte units = 2 (mod 8) -> conflict-free LDS.128
constexpr int STAGES = 3;  // mbarrier pipeline depth (3 x 51.2 KB)
constexpr int NQ = 4;      // GEMM operand planes: D, Px, Py, Pz (two 16-byte pair-planes)
constexpr int SLOT_ALIGN_GIAO = 4;   // with GIAO every atom's slot run is padded to whole k4 MMA steps (atom-boundary taps)

struct DevBasis {
    int natoms, nbf, nshell;
    const double *atom_xyz;       // [natoms][3]
    const double *atom_maxthr;    // [natoms] largest screening radius on the atom
    const int *atom_shell_off;    // [natoms+1] into the internal (radius-sorted) shell order
    const int *atom_func_off;     // [natoms+1] first internal function of the atom
    const int *sh_l, *sh_nprim, *sh_prim_off, *sh_foff;   // [nshell] internal order; sh_foff = first internal function
    const double *sh_thr;         // [nshell] screening radius (descending within an atom)
    const double *sh_thr2e, *atom_maxthr2e;   // (radius + 1e-9)^2 per shell / of the atom's widest shell: the tile-level tests compare squared distances (no sqrt)
    const double *alpha, *ncc;    // primitives
    const double *fR;             // [3][nbf] centre coordinates of each internal function (SoA)
    int turbomole;                // component order (gtodefs.f90:109-123) instead of the standard one
    int slot_align;               // 1, or SLOT_ALIGN_GIAO: each active atom occupies a multiple of this many tile slots
};

struct TileDesc {
    int pt0, npts;      // range in the sorted point list
    int nact;           // K slots: atom runs (each aligned to slot_align) padded to a multiple of 8 (0: nothing within screening range)
    int nraw;           // K slots before the final padding
    int nn, nreal;      // N columns: the nreal active functions padded to a multiple of 8 (no per-atom padding on this side)
    int geo, nruns;     // index into the TileGeo array; number of active atoms (slot runs)
    long long panel_off;  // doubles, into the panel pool: 4 planes x nact x LDP
    long long fidx_off;   // ints, into the index pool: nact slot -> function indices, then nn column -> K-slot indices
    long long atab_off;   // TileAtom entries, into the atom-table pool (nruns entries, atom order = slot order)
    int col0, col1;       // N columns this work item contracts: [0, nn) for a whole tile; a slice of it when few tiles share the GPU
    int part, pad_;       // >= 0: the item is a slice -- its 13 row sums go to partial block `part` (k_slice_reduce adds the slices); -1: store results
};

// One active atom of a tile, in slot order.  kend4 = end of its slot run in units of 4 slots (one m16n8k4 K step).
// (dx,dy,dz) = R_A - R_next for all but the last run, R_A - tile centre for the last: the Abel-summed weights of the
// GIAO taps in k_jtensor (sum_A (R_A - c) P_A = sum_A C_A (R_A - R_{A+1}) + C_n (R_n - c), C_A = running K sum after atom A).
struct __align__(32) TileAtom { double dx, dy, dz; int kend4, atom; };

struct TileGeo { double lox, loy, loz, hix, hiy, hiz, rho, pad_; };   // axis-aligned bounding box of the tile's points (+ radius about its centre)
struct TileSeg { int pt0, npts; };                         // a tile = npts <= MT consecutive points of the sorted list
struct TileInfo { float rho, gmax; int imax, nraw, natom, nreal; };   // radius, largest consecutive gap (and where), active slots (atom runs aligned), active atoms, active functions

// ---- device-side tile plan (k_prepare.cu): no host round trip between the sort and the first k_basis launch except ONE small
// summary copy.  A run of MT consecutive Hilbert-sorted points that straddles a re-entry of the curve is cut at its largest
// consecutive gap, recursively (depth <= SPLIT_DEPTH, at most MAXSUB pieces), inside one CTA (k_tile_split).
constexpr int MAXSUB = 16;
constexpr int SPLIT_DEPTH = 7;
constexpr int MAX_BATCH = 4096;     // panel-pool batches one call may need
struct TileCum { long long cost, panel, fidx, atab; };   // per tile: scheduling cost, panel doubles, index ints, atom-table entries (and their exclusive prefix sums)
constexpr int DRAIN_BATCHES = 4, DRAIN_GROUPS = 4;
struct PlanSummary {                // device -> host, once per plan
    int ntiles;                     // tiles of the whole point set (Hilbert order)
    int tlo, thi;                   // the tile range this rank owns: equal shares of the cumulative cost
    int nbatch, max_nruns, overflow;
    long long pt_lo, pt_hi;         // sorted-point range of [tlo, thi)
    long long panel_range;          // panel doubles of the range
    long long max_tile_panel;
    long long cost_total, cost_range;
    double sum_nact, flops4, flops2, taps, useful_mm, useful_taps;   // statistics of the range (see gimic_b200_stats)
    // Drain groups (only when the range has at most DRAIN_BATCHES batches): a batch is cut into up to DRAIN_GROUPS runs of Hilbert-ordered
    // tiles of `drain_chunk` panel doubles each; a second tile list is ordered (batch, group, costliest first), so a caller that wants
    // its rows on the host can launch the contraction group by group and copy a finished group's rows out while the next one runs.
    long long drain_chunk;
    int group_tile[DRAIN_BATCHES * DRAIN_GROUPS + 1];        // first tile (Hilbert order, absolute) of group g of batch b at [b * DRAIN_GROUPS + g]; -1 = empty
    long long group_pt[DRAIN_BATCHES * DRAIN_GROUPS + 1];    // its first sorted point
    int batch_start[MAX_BATCH + 1]; // tile index (Hilbert order, absolute) where each batch begins; [nbatch] = thi
    long long batch_pt[MAX_BATCH + 1];   // first sorted point of each batch; [nbatch] = pt_hi (a batch is a contiguous run of compact output rows)
};

// drain group of a tile inside its batch (PlanSummary::drain_chunk): its panel offset in the batch over the chunk size
__host__ __device__ inline int drain_group(long long panel_prefix_in_range, long long pool_doubles, long long chunk) {
    const long long off = panel_prefix_in_range % pool_doubles;
    const long long g = off / chunk;
    return (int)(g < DRAIN_GROUPS - 1 ? g : DRAIN_GROUPS - 1);
}

// ---- launch wrappers (defined in k_prepare.cu / k_jtensor.cu / k_fields.cu) ----------------------
void launch_morton_keys(const double *r, long n, const double *bbox_lo, double inv_cell, uint32_t *keys, int *vals, cudaStream_t s);
void launch_gather_points(const double *r, const int *perm, long n, double *rsx, double *rsy, double *rsz, cudaStream_t s);
void launch_grid_points(const double *origin_basv /*12 doubles, device*/, const double *p0, const double *p1, const double *p2,
                        int n0, int n1, int n2, long lo, long hi, double *r, cudaStream_t s);
struct PlanBuffers {                // device workspaces of one plan (all sized by the number of initial runs nrun0 = ceil(n / MT))
    TileSeg *slot_seg; TileGeo *slot_geo; TileInfo *slot_info;   // [nrun0 * MAXSUB] pieces of every run, in order
    int *cnt; int *off;                                          // [nrun0] pieces per run, [nrun0 + 1] their exclusive prefix sum
    TileGeo *geo; TileDesc *desc;                                // [cap] compact tiles in Hilbert order
    TileCum *cum;                                                // [cap + 1] sizes -> exclusive prefix sums (in place), [ntiles] = totals
    unsigned long long *keys0, *keys1; int *ord0, *ord1;         // [cap] scheduling keys / tile order (CUB sort)
    TileDesc *tiles;                                             // [cap] this rank's tiles in processing order (batch by batch, longest first)
    unsigned long long *gkeys0; TileDesc *gtiles;                // [cap] the same tiles ordered (batch, drain group, longest first): only used when host outputs are drained group by group
    PlanSummary *summary;                                        // device copy
    int *tops_i; TileCum *tops_c;                                // [cap / 2048 + 2] tile totals of the prefix sums
    int cap;
};
void launch_plan_tiles(const DevBasis &B, const double *rsx, const double *rsy, const double *rsz, long n, double split_radius,
                       int rank, int nranks, long long pool_doubles, const PlanBuffers &pb, cudaStream_t s);
void launch_plan_order(const PlanBuffers &pb, int tlo, int nt, bool grouped, void *sorttmp, size_t sorttmp_bytes, cudaStream_t s);
size_t plan_sort_temp_bytes(int nt);
void launch_perm_index(const int *perm, long n, long *index, cudaStream_t s);
void launch_basis(const DevBasis &B, const TileDesc *tiles, int ntiles, int max_nruns, const TileGeo *geo, const double *rsx, const double *rsy,
                  const double *rsz, double *panel_pool, int *fidx_pool, TileAtom *atab_pool, cudaStream_t s);
void launch_panel_scatter(const TileDesc *tiles, int ntiles, const double *panel_pool, const int *fidx_pool, const int *perm, const int *f2user,
                          int nbf, double *bf, double *dr, cudaStream_t s);
void launch_basis_dense(const DevBasis &B, const int *f2user, long n, const double *r, double *bf, double *dr, cudaStream_t s);
size_t sort_temp_bytes(long n);
void launch_sort_pairs(void *temp, size_t temp_bytes, const uint32_t *kin, uint32_t *kout, const int *vin, int *vout, long n, cudaStream_t s);

struct JtensorArgs {
    const TileDesc *tiles; int ntiles; int *counter;
    const double *panel_pool; const int *fidx_pool; const TileAtom *atab_pool; const TileGeo *geo;
    const double *Bop; long long plane_stride; int ldb;     // pair-plane operands [2][nbf][ldb][2] = (D,Px), (Py,Pz); plane_stride in doubles
    const double *fR; int nbf;
    const double *rsx, *rsy, *rsz; const int *perm;         // perm == null: compact output, sorted point p goes to row p - out_base
    long long out_base;
    double *tens; double *edens;                            // outputs (any may be null): 9 x n tensors, n densities,
    double *jvec, *jmod, *acid; double B[3];                // 3 x n J = T.B, n signed |J| (jfield.f90:446-489), n ACID (acid.f90:9-45); field direction
    int jpath;                                              // J = T.B path: operands (D, P.B), the tensor is never formed (tens, acid unavailable)
    double *part;                                           // [items][MT][PART_LD] partial row sums of sliced tiles (see TileDesc::part)
    int paramag, diamag;
};
void launch_jtensor(const JtensorArgs &a, bool giao, int nsm, cudaStream_t s);
// Few tiles (a single plane of an integral, a handful of points): one tile per SM leaves most of the GPU idle and the call waits for one
// tile's contraction (12.8 ms for a 36 x 36 plane at nbf = 10^4).  The tiles are then cut into S column slices, each a work item of
// its own (same panels, same K sweep, a subset of the nu chunks); k_slice_reduce adds the slices' row sums in slice order and stores.
constexpr int PART_LD = 13;          // doubles per row of a partial block
constexpr int SLICE_COLS = 32;       // slice boundaries are multiples of this many columns (the widest nu chunk of the kernels)
bool jtensor_supports_slices();      // the epilogue-warpgroup kernels do; the round-1 mapping (GIMIC_B200_EPI=0) does not
// what a tile costs the contraction (one DMMA column step = 1): MMA k-steps x columns (4 planes) + GIAO taps + the tile's share of
// k_basis + a fixed part.  A tile costs the same whether its 128 rows are all points or not: letting the consumer warps without
// valid rows skip their MMAs was measured (round 2, calls L/M) -- the extra branch cost the full-tile loop 4.5 % and a 36x36 plane
// gained nothing.
__host__ __device__ inline long long piece_cost(int npts, int nraw, int nreal, int natom) {
    const long long nact = (nraw + 7) & ~7, nn = (nreal + 7) & ~7;
    (void)npts;
    return 4LL * nact * nn + 3LL * nn * natom + 110LL * nact + (nact ? 8192 : 256);
}
// Column slices of a tile (few tiles in the whole point set): a tile gets as many slices as its cost holds `item_cost` units, at most
// S (the item slots per tile) and at most one per SLICE_COLS columns -- every work item then costs about the same, whatever the
// spread of active-set sizes between the tiles.  Returns the slice width in columns; slice s covers [s*w, min(nn, (s+1)*w)).
__host__ __device__ inline int slice_width(const TileDesc &td, int S, long long item_cost) {
    const long long c = piece_cost(td.npts, td.nraw, td.nreal, td.nruns);
    long long k = (c + item_cost - 1) / item_cost;
    const int kmax = td.nn / SLICE_COLS > 1 ? td.nn / SLICE_COLS : 1;
    if (k > S) k = S;
    if (k > kmax) k = kmax;
    if (k < 1) k = 1;
    return ((td.nn + (int)k - 1) / (int)k + SLICE_COLS - 1) / SLICE_COLS * SLICE_COLS;
}
void launch_tile_slices(const TileDesc *tiles, int nt, int S, long long item_cost, TileDesc *items, cudaStream_t s);
void launch_slice_reduce(const JtensorArgs &a, const TileDesc *tiles, int nt, int S, long long item_cost, bool giao, cudaStream_t s);
size_t jtensor_smem_bytes();

void launch_build_operand(double *out, int nbf, int ldb, long long plane_stride, const double *srcA, const double *srcB, double signB,
                          const int *f2user, cudaStream_t s);

void launch_operand_j(double *out, const double *op, long long plane_stride, const double *B3 /*host*/, cudaStream_t s);
void launch_operand_combine(double *out, const double *a, const double *b, double sg, long count, cudaStream_t s);
void launch_jmod(long n, const double *r, const double *jvec, const double *B3 /*host*/, double *jmod, cudaStream_t s);
void launch_fields(long n, const double *r, const double *tens, const double *B3 /*host values*/, double *jvec, double *jmod, double *acid, cudaStream_t s);
void launch_divj(long n, const double *jv6 /* [6][n][3] shifted jvecs */, double h, double *divj, cudaStream_t s);
void launch_shift_points(long n, const double *r, double h, double *r6, cudaStream_t s);
struct QuadArgs {
    const double *tens; int p1, nrows;            // rows = (j,k) pairs handled by this call, i fastest
    const double *jvec = nullptr;                 // if set: J = T.B as delivered by the contraction's J path (3 per point), tens unused
    const double *r;                              // points (3 x p1*nrows)
    const double *w1, *wrow;                      // w_i [p1]; per-row weight w_j*w_k [nrows]
    double B[3], normal[3], center[3], radius; int what;
    double *row_partials;                         // [nrows][7]
    double *out7;                                 // device, 7 doubles
};
void launch_quadrature(const QuadArgs &q, cudaStream_t s);
void launch_property(long n, const double *r, const double *w, const double *tens, int natoms, const double *coords, int nseg,
                     const long *seg_end, double *out, cudaStream_t s);

void launch_property_integrand(long n, const double *r, const double *tens, int chi, const double *c3, double *out, cudaStream_t s);

}  // namespace gb
