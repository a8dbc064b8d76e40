#!/usr/bin/env python3
"""bench.py -- J^B tensor grid points per second at nbf ~ 10^4 (BASELINE.json metric).

Workload (configs[4], "nanoring-size synthetic"): 278 carbon-like centres x 36 cartesian functions
(def2-TZVP carbon shells) = 10 008 basis functions, compact hexagonal flake, seeded random symmetric D
and antisymmetric P_x,P_y,P_z; cdens tensors on the COMPLETE 256^3 even grid over the bounding box + 8 bohr.
One STEP = one pass of the whole hot path over the whole grid through the product's multi-GPU calls:
every rank runs gimic_b200_partition_grid (points generated on the device, Hilbert sort, tiles with gap
splitting, active-set sizes, this rank's equal-COST share of the tile list) and gimic_b200_partition_calc
(basis panels + FP64 DMMA contraction + fused tensor epilogue on the owned tiles).  Total work is fixed
as N grows (strong scaling); there is no data-path collective in cdens mode.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU (oracle port)
  python bench.py --mode octant ...                        # round-1 extra: one 128^3 octant per rank (weak scaling)

Timing: CUDA events on the library's stream from the start of the partition call to the end of the calc call
(stats.ms_span), max over ranks.  `value` has the outputs resident in HBM; `e2e` goes through the same C-ABI calls
with pinned HOST output buffers (the rows of a finished panel batch are copied out while the next batch is
contracted; the copy is inside the timed span).  After the headline, rank 0 at N=1 also times the other stages
of the path (field pass, basis kernel, one plane integral, config 4) and, at N>1, the NCCL-reduced plane integral.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSLAB = 8


def build_workload(natoms, grid_n, geometry="flake", general_p=False):
    from gimic_b200 import synthetic
    sh, dens, nbf = synthetic.synthetic_case(natoms, geometry, seed=1234, general_p=general_p)
    origin, basv, pts = synthetic.box_grid(sh["coords"], (grid_n, grid_n, grid_n))
    return sh, dens, nbf, origin, basv, pts


def slab_points(origin, basv, pts, slab, nslab=NSLAB):
    """points of octant `slab` of the grid (bit 0: upper x half, bit 1: upper y half, bit 2: upper z half),
    i fastest inside the brick like the reference's flat index (grid.f90:478-511)"""
    assert nslab == 8
    h = [len(p) // 2 for p in pts]
    sel = [np.arange(h[d]) + (h[d] if (slab >> d) & 1 else 0) for d in range(3)]
    x = origin[0] + pts[0][sel[0]]; y = origin[1] + pts[1][sel[1]]; z = origin[2] + pts[2][sel[2]]
    r = np.empty((len(z), len(y), len(x), 3))
    r[..., 0] = x[None, None, :]; r[..., 1] = y[None, :, None]; r[..., 2] = z[:, None, None]
    return r.reshape(-1, 3)


def flat_points(origin, pts, idx):
    """coordinates of flat grid indices (i fastest, grid.f90:478-511) of an axis-aligned grid"""
    n0, n1 = len(pts[0]), len(pts[1])
    i = idx % n0; j = (idx // n0) % n1; k = idx // (n0 * n1)
    return np.ascontiguousarray(np.stack([origin[0] + pts[0][i], origin[1] + pts[1][j], origin[2] + pts[2][k]], 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def fp64_peak():
    """FP64 roofline denominator.  MEASURED_PEAKS.json (driver-written) only carries bf16 and HBM-copy figures, so the
    FP64 peak is this repo's own measurement on the pool's B200s: cuBLAS DGEMM 8192^3 (profiles/r01_dgemm_peak.json);
    the DMMA issue-rate microbenchmark (profiles/r01_fp64_peaks.txt) gives 37.19 TF."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_dgemm_peak.json")))
        return float(d["dgemm_8192_tflops_sustained"]), "measured cuBLAS DGEMM 8192^3 sustained (profiles/r01_dgemm_peak.json); MEASURED_PEAKS.json has no FP64 entry"
    except Exception:
        return 37.2, "nominal 148 SM x 64 FMA/clk x 1.965 GHz (no measured file)"


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6540.8, "fallback: the pool's measured copy bandwidth of round 1 (MEASURED_PEAKS.json absent)"


def ncu_traffic(nbf):
    """dram bytes per k_jtensor launch from the committed `ncu --set full` capture (NOT measured in this run)"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = d.get(str(nbf), {})
        return e.get("k_jtensor_dram_bytes_per_launch"), e.get("source", "profiles/ncu_traffic.json")
    except Exception:
        return None, None


def gauss_axis(a, b, npts, order=9):
    from gimic_b200.gengauss import gausspoints
    p = np.zeros(npts); w = np.zeros(npts)
    gausspoints(a, b, order, p, w)
    return p, w


def flake_plane(sh, npts):
    """a Gauss-Legendre plane (y = 0, spanned by x and z) through the flake, npts x npts points (integral mode, integral.f90)"""
    import gimic_b200
    c = sh["coords"]
    lx = float(c[:, 0].max() - c[:, 0].min()) + 10.0
    p0, w0 = gauss_axis(0.0, lx, npts); p1, w1 = gauss_axis(0.0, 16.0, npts)
    return gimic_b200.Grid([c[:, 0].min() - 5.0, 0.0, -8.0], [[1, 0, 0], [0, 0, 1], [0, -1, 0]], [p0, p1, np.zeros(1)], [w0, w1, np.ones(1)])


def run_cpu(sh, dens, r, nthreads, repeats=1):
    """the reference algorithm on the host CPU: oracle port (C++/OpenMP restatement, thread-private scratch,
    static schedule over points like jfield.f90:114-129).  Returns (points/s, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from gimic_b200 import synthetic
    o = run_cpu.cache.get("o")
    if o is None:
        o = O.Oracle.from_arrays(dens_a=synthetic.dens_to_colmajor(dens), **sh)
        run_cpu.cache["o"] = o
    # torchrun exports OMP_NUM_THREADS=1; the CPU baseline is meant to use every host core it can
    cores = nthreads or max(O.max_threads(), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    best = 1e300
    for _ in range(repeats):
        t0 = time.perf_counter()
        o.ctensor(r, "total", nthreads=cores)
        best = min(best, time.perf_counter() - t0)
    return r.shape[0] / best, cores, best


run_cpu.cache = {}


def stage_records(g, grid, dev, t_dev_full, index_full, sh, B):
    """N=1 extras, same process: the other stages of the path, each against the roofline that bounds it."""
    import ctypes as C
    import torch
    import gimic_b200
    from gimic_b200 import _lib
    L = _lib.lib()
    out = {}
    hbm, hbm_src = hbm_peak()
    # (1) field pass (k_fields: jvec + signed |J| + ACID from stored tensors) on the headline grid's tensors, HBM roofline
    n = t_dev_full.shape[0]
    pts_all = flat_points(grid.origin, grid.pts, index_full.cpu().numpy())
    r_dev = torch.from_numpy(pts_all).to(dev)
    jvec = torch.empty((n, 3), dtype=torch.float64, device=dev); jmod = torch.empty(n, dtype=torch.float64, device=dev)
    acid = torch.empty(n, dtype=torch.float64, device=dev)
    ms = []
    for _ in range(6):
        _lib.check(L.gimic_b200_fields_from_tensors(g._h, n, C.c_void_p(r_dev.data_ptr()), C.c_void_p(t_dev_full.data_ptr()), B.ctypes.data_as(_lib.dp),
                                                    C.c_void_p(jvec.data_ptr()), C.c_void_p(jmod.data_ptr()), C.c_void_p(acid.data_ptr()), _lib.DEVICE_PTR))
        ms.append(g.stats()["ms_fields"])
    bpp = 72 + 24 + 24 + 8 + 8
    best = min(ms[2:])
    out["fields"] = {"kernel": "k_fields (jvec + signed |J| + ACID from stored tensors)", "points": n, "bytes_per_point": bpp, "ms": best,
                     "roofline": {"bound": "hbm", "achieved": n * bpp / (best * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                  "frac": n * bpp / (best * 1e-3) / 1e9 / hbm, "peak_source": hbm_src},
                     "note": "inputs 1.6 GB > L2; in the cdens driver these fields are written by the contraction's epilogue and this pass is not run"}
    del jvec, jmod, acid, r_dev
    # (2) one 36 x 36 Gauss plane integral (integral mode, latency-bound): wall-clock through the C ABI incl. grid upload and result copy
    plane = flake_plane(sh, 36)
    g.integrate(plane, B, "total", 3)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); s7 = g.integrate(plane, B, "total", 3); ts.append(time.perf_counter() - t0)
    st = g.stats()
    g.set_profiling(True)
    g.integrate(plane, B, "total", 3)
    sp = g.stats()
    g.set_profiling(False)
    out["integral_36x36"] = {"what": "gimic_b200_integrate: one 36x36 Gauss-Legendre plane through the flake (current + modulus), nbf=10008",
                             "wall_us": min(ts) * 1e6, "points": 1296, "tiles": int(st["n_tiles"]), "mean_active_slots": st["sum_nact"] / max(st["n_tiles"], 1),
                             "executed_over_useful_flops": st["executed_flops"] / st["useful_flops"] if st["useful_flops"] else None,
                             "stage_ms_profiled_call": {k: sp[k] for k in ("ms_total", "ms_sort", "ms_tiles", "ms_basis", "ms_contract")},
                             "launches": int(sp["launches"]), "sums": [float(v) for v in s7[:6]]}
    return out


def config4_record(dev):
    """configs[3]: coronene-size synthetic (42 centres, nbf = 1512), ACID + jmod (+ jvec) on a 128^3 grid, 1 GPU, with a parity sample"""
    import torch
    import gimic_b200
    from gimic_b200 import synthetic
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    sh, dens, nbf = synthetic.synthetic_case(42, "flake", seed=1234)
    flat = synthetic.dens_to_colmajor(dens)
    g4 = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, device=dev.index, **sh)
    origin, basv, pts = synthetic.box_grid(sh["coords"], (128, 128, 128))
    grid = gimic_b200.Grid(origin, basv, pts)
    B = np.array([0.0, 0.0, 1.0])
    ms = []
    for it in range(5):
        g4.partition(grid, 0, 1)
        res = g4.partition_calc(B, "total", jvec=True, jmod=True, acid=True, device=dev)
        if it >= 2:
            ms.append(g4.stats()["ms_span"])
    st = g4.stats()
    n = grid.n
    # parity of the timed result on a CPU sample
    rng = np.random.default_rng(5)
    pick = rng.choice(n, size=64, replace=False)
    inv = torch.empty(n, dtype=torch.int64, device=dev); inv[res["index"]] = torch.arange(n, device=dev)
    rows = inv[torch.from_numpy(pick).to(dev)]
    o = O.Oracle.from_arrays(dens_a=flat, **sh)
    rp = flat_points(origin, pts, pick)
    tref = o.ctensor(rp, "total")
    jref = O.jvectors(tref, B)
    jv = res["jvec"][rows].cpu().numpy()
    scale = np.abs(tref).max(axis=1, keepdims=True)
    perr = float((np.abs(jv - jref) / (1e-10 * np.maximum(np.abs(jref), 1e-3 * scale) + 1e-12)).max())
    aref = O.acid_field(tref)
    ac = res["acid"][rows].cpu().numpy()
    aerr = float((np.abs(ac - aref) / (1e-10 * np.maximum(np.abs(aref), 1e-3 * scale[:, 0] ** 2) + 1e-12)).max())
    rec = {"workload": f"synthetic flake nbf={nbf}, 128^3 even grid, jvec + signed |J| + ACID written by the contraction's epilogue (no separate field pass)",
           "points": n, "ms_per_pass": min(ms), "points_per_s": n / (min(ms) * 1e-3), "executed_tflops": st["executed_flops"] / (min(ms) * 1e-3) / 1e12,
           "parity_jvec_max_scaled_err_64pts": perr, "parity_acid_max_scaled_err_64pts": aerr,
           "parity_scale": "1e-10 relative at the scale of the point's tensor (random unphysical densities: J components cancel), 1e-12 absolute"}
    g4.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="grid", choices=["grid", "octant"],
                    help="grid: the whole 256^3 grid per step, cost-balanced over the ranks (strong scaling, headline); "
                         "octant: round-1 extra, one 128^3 octant per rank (weak scaling)")
    ap.add_argument("--natoms", type=int, default=278)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--cpu-points", type=int, default=192, help="bounded CPU sample per step / for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the stage records after the headline")
    ap.add_argument("--geometry", default="flake", choices=["flake", "ring"],
                    help="flake: compact hexagonal flake (headline); ring: radius-120-bohr ring, mostly empty box (SURVEY 8d ii)")
    ap.add_argument("--general-p", action="store_true", help="general (not antisymmetric) perturbed densities P_b")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = max(args.steps, 1)
    octant = args.mode == "octant"
    ntot = args.grid ** 3

    geo = "hex flake" if args.geometry == "flake" else "ring (radius 120 bohr)"
    step_desc = (f"step = octant (rank mod 8) = {ntot // NSLAB} points/GPU (weak scaling extra)" if octant else
                 f"step = the whole grid = {ntot} points, split over the ranks by cumulative tile cost (gimic_b200_partition_grid)")
    cfg = {"workload": f"synthetic {geo} {args.natoms} C-like centres x 36 fn (nbf={args.natoms * 36}), cdens J^B tensors, "
                       f"{args.grid}^3 even grid over bbox+8 bohr, {step_desc}",
           "nbf": args.natoms * 36, "grid": [args.grid] * 3, "points_per_step": ntot // NSLAB * world if octant else ntot,
           "spincase": "total (closed shell)", "giao": True, "screening_thrs": 1e-8,
           "densities": "seeded random symmetric D, " + ("general" if args.general_p else "antisymmetric") + " P_x,P_y,P_z",
           "cache": "inputs larger than L2 (contraction operand 4*nbf^2*8 B = %.1f GB, panels streamed)" % (4 * (args.natoms * 36) ** 2 * 8 / 1e9),
           "parallelism": f"{world} rank(s), one per GPU; equal-cost runs of Hilbert-ordered tiles per rank, no data-path collective"}
    scaling = "weak" if octant else "strong"

    # ------------------------------------------------------------------ reference arm (CPU) -------
    if args.impl == "reference":
        if rank != 0:
            return 0
        sh, dens, nbf, origin, basv, pts = build_workload(args.natoms, args.grid, args.geometry, args.general_p)
        rng = np.random.default_rng(77)
        if octant:
            r = slab_points(origin, basv, pts, 0)
            sample = np.ascontiguousarray(r[rng.choice(r.shape[0], size=args.cpu_points, replace=False)])
        else:
            sample = flat_points(origin, pts, rng.choice(ntot, size=args.cpu_points, replace=False))
        for _ in range(args.warmup):
            run_cpu(sh, dens, sample[: max(16, args.cpu_points // 8)], 0)
        times = []
        for _ in range(K):
            pps, cores, dt = run_cpu(sh, dens, sample, 0)
            times.append(dt)
        dt = float(np.mean(times))
        val = args.cpu_points / dt
        line = {"impl": "reference", "metric": "J^B tensor grid points/sec", "value": val, "unit": "points/s", "n_gpus": args.gpus,
                "steps": K, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": val, "unit": "points/s", "cores": cores, "kind": "port",
                                 "sample": f"{args.cpu_points} random points of the step's grid per step; C++/OpenMP restatement of the "
                                           "reference algorithm (dense 7 GEMV + 28 DOT per point), all host threads"},
                "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ CUDA arm --------------------
    import ctypes as C
    import torch
    import gimic_b200
    from gimic_b200 import synthetic, _lib
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: gimic-b200 has no CPU path"}))
        return 1
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    sh, dens, nbf, origin, basv, pts = build_workload(args.natoms, args.grid, args.geometry, args.general_p)
    flat = synthetic.dens_to_colmajor(dens)
    g = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, device=local_rank, **sh)
    del flat
    grid = gimic_b200.Grid(origin, basv, pts)
    B = np.array([0.0, 0.0, 1.0])
    KEYS = ("ms_total", "ms_plan", "ms_span", "ms_contract", "ms_basis", "ms_sort", "ms_tiles", "launches", "contract_launches", "executed_flops",
            "useful_flops", "dense_flops", "sum_nact", "n_tiles", "n_points", "panel_bytes")
    L = _lib.lib()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if octant:
        r_np = slab_points(origin, basv, pts, rank % NSLAB)
        n_local = r_np.shape[0]
        r_host = torch.empty((n_local, 3), dtype=torch.float64, pin_memory=True); r_host.numpy()[:] = r_np
        t_host = torch.empty((n_local, 9), dtype=torch.float64, pin_memory=True)
        r_dev = r_host.to(dev); t_dev = torch.empty((n_local, 9), dtype=torch.float64, device=dev)
        idx_dev = None

        def one_step(host):
            if host:
                g.jtensors(r_host.numpy(), "total", out=t_host.numpy())
            else:
                g.jtensors(r_dev, "total", out=t_dev)
            s = g.stats(); s["ms_span"] = s["ms_total"]
            return s
        h2d, d2h = n_local * 24, n_local * 72
    else:
        n_local = g.partition(grid, rank, world)
        idx_dev = torch.empty(n_local, dtype=torch.int64, device=dev)
        t_dev = torch.empty((n_local, 9), dtype=torch.float64, device=dev)
        t_host = torch.empty((n_local, 9), dtype=torch.float64, pin_memory=True)
        idx_host = torch.empty(n_local, dtype=torch.int64, pin_memory=True)

        def one_step(host):
            cnt = g.partition(grid, rank, world)
            assert cnt == n_local
            if host:
                _lib.check(L.gimic_b200_partition_calc(g._h, None, _lib.TOTAL, C.c_void_p(idx_host.data_ptr()), C.c_void_p(t_host.data_ptr()),
                                                       None, None, None, None, 0))
            else:
                _lib.check(L.gimic_b200_partition_calc(g._h, None, _lib.TOTAL, C.c_void_p(idx_dev.data_ptr()), C.c_void_p(t_dev.data_ptr()),
                                                       None, None, None, None, _lib.DEVICE_PTR))
            return g.stats()
        h2d = int(8 * (12 + 2 * len(pts[0]) + len(pts[1]) + len(pts[2])))     # the grid description (points are generated on the device)
        d2h = n_local * 80                                                      # tensors + point indices

    def run_steps(k, host):
        tot = {key: 0.0 for key in KEYS}
        for _ in range(k):
            s = one_step(host)
            for key in KEYS:
                tot[key] += s[key]
        return tot

    # profiling (per-stage events, extra syncs) only in a separate pass: the timed steps run without it
    g.set_profiling(False)
    run_steps(W, host=False)
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    t0 = time.perf_counter()
    tot = run_steps(K, host=False)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    run_steps(1, host=True)
    barrier()
    tot_h = run_steps(K, host=True)
    barrier()
    g.set_profiling(True)
    prof = run_steps(2, host=False)       # stage times and the contraction's launch duration (same work, CUDA events per stage)
    g.set_profiling(False)
    barrier()

    # points with at least one unscreened basis function (all others are exact zeros and cost nothing; SURVEY 8d caveat)
    n_active = torch.tensor([float((t_dev.abs().amax(dim=1) > 0).sum())], dtype=torch.float64, device=dev)
    ms = torch.tensor([tot["ms_span"] / K, tot_h["ms_span"] / K], dtype=torch.float64, device=dev)
    per_rank = torch.tensor([tot["ms_span"] / K, prof["executed_flops"] / 2, float(n_local), prof["ms_contract"] / 2], dtype=torch.float64, device=dev)
    gathered = [per_rank.clone() for _ in range(world)]
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_active, op=dist.ReduceOp.SUM)
        dist.all_gather(gathered, per_rank)
    ms_step, ms_step_h = float(ms[0]), float(ms[1])
    npts_step = (ntot // NSLAB * world) if octant else ntot
    value = npts_step / (ms_step * 1e-3)
    e2e = npts_step / (ms_step_h * 1e-3)

    extras = {}
    if dist is not None and not octant and not args.no_extras:
        # integral mode at N > 1: plane rows split over the ranks, ONE NCCL all-reduce of the 7 partial sums
        # (the reference's collect_sum calls are commented out, integral.f90:157-161)
        from gimic_b200.gimic import integrate_distributed
        plane = flake_plane(sh, 360)
        integrate_distributed(g, plane, B, "total", 3)
        barrier(); t1 = time.perf_counter()
        part = integrate_distributed(g, plane, B, "total", 3)
        barrier(); dt_int = time.perf_counter() - t1
        if rank == 0:
            t2 = time.perf_counter(); ref7 = g.integrate(plane, B, "total", 3); dt_one = time.perf_counter() - t2      # the same plane on one GPU
            extras["integral_nccl"] = {"what": f"integral mode: 360x360 Gauss plane through the flake (129600 points), rows split over {world} ranks, "
                                               "one NCCL all-reduce of 7 doubles",
                                       "wall_ms": dt_int * 1e3, "wall_ms_one_gpu": dt_one * 1e3, "current": float(part[0]),
                                       "max_rel_dev_vs_one_gpu": float(np.max(np.abs(part[:6] - ref7[:6]) / np.maximum(np.abs(ref7[:6]), 1e-300)))}

    if rank == 0:
        peak, peak_src = fp64_peak()
        t_contract = prof["ms_contract"] * 1e-3
        achieved = prof["executed_flops"] / t_contract / 1e12 if t_contract > 0 else None
        traffic, traffic_src = ncu_traffic(nbf)
        roof = {"bound": "tensor", "kernel": "k_jtensor<GIAO> (FP64 DMMA contraction of Phi with [D|Px|Py|Pz], GIAO terms by atom-boundary taps of the D accumulator, fused tensor epilogue)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "flops": "EXECUTED FP64 flops per tile: DMMA 2*128*4*nact*nn + GIAO-tap DFMA 2*128*3*nn*natoms_active (screened-function skipping on; exact zeros in the reference)",
                # the same flops without K/N padding and partial tiles (real points x active functions): what a perfect tiling would issue
                "useful_tflops": (prof["useful_flops"] / t_contract / 1e12) if t_contract > 0 else None,
                "useful_frac": (prof["useful_flops"] / t_contract / 1e12 / peak) if t_contract > 0 else None,
                "executed_over_useful": (prof["executed_flops"] / prof["useful_flops"]) if prof["useful_flops"] else None,
                "avg_launch_ms": prof["ms_contract"] / max(prof["contract_launches"], 1), "launches_timed": prof["contract_launches"],
                "timed_in": "a separate 2-step pass of the same work with per-stage CUDA events on the library stream (rank 0)",
                "share_of_step": prof["ms_contract"] / prof["ms_span"] if prof["ms_span"] else None,
                "dense_equivalent_tflops": (14.0 * nbf * nbf + 56.0 * nbf) * npts_step / (ms_step * 1e-3) / 1e12,
                "skip_ratio_dense_over_executed": tot["dense_flops"] / tot["executed_flops"] if tot["executed_flops"] else None,
                "mean_active_functions_per_tile": tot["sum_nact"] / max(tot["n_tiles"], 1),
                "frac_points_with_active_functions": float(n_active) / npts_step,
                "active_points_per_s": float(n_active) / (ms_step * 1e-3)}
        stage = {k2: prof[k2] / 2 for k2 in ("ms_plan", "ms_sort", "ms_tiles", "ms_basis", "ms_contract", "ms_span")}
        stage["basis_panel_gbs"] = prof["panel_bytes"] / (prof["ms_basis"] * 1e-3) / 1e9 if prof["ms_basis"] else None
        stage["basis_frac_of_hbm"] = (stage["basis_panel_gbs"] / hbm_peak()[0]) if stage["basis_panel_gbs"] else None
        stage["non_contraction_share"] = 1.0 - prof["ms_contract"] / prof["ms_span"] if prof["ms_span"] else None
        balance = {"per_rank_ms": [float(x[0]) for x in gathered], "per_rank_executed_flops": [float(x[1]) for x in gathered],
                   "per_rank_points": [int(x[2]) for x in gathered], "per_rank_contract_ms": [float(x[3]) for x in gathered]}
        fl = balance["per_rank_executed_flops"]
        balance["flops_max_over_mean"] = max(fl) / (sum(fl) / len(fl)) if sum(fl) else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rng = np.random.default_rng(77)
            if octant:
                pick = rng.choice(n_local, size=args.cpu_points, replace=False)
                sample = np.ascontiguousarray(r_np[pick])
                got = t_dev[torch.from_numpy(pick).to(dev)].cpu().numpy()
            else:
                pick = rng.choice(ntot, size=args.cpu_points, replace=False)
                sample = flat_points(origin, pts, pick)
                inv = torch.empty(ntot, dtype=torch.int64, device=dev); inv[idx_dev] = torch.arange(ntot, device=dev)
                got = t_dev[inv[torch.from_numpy(pick).to(dev)]].cpu().numpy()      # rows of the TIMED result
                del inv
            run_cpu(sh, dens, sample[:16], 0)
            pps, cores, dt = run_cpu(sh, dens, sample, 0)
            # parity of the timed GPU result on the CPU-evaluated sample (same tolerance as the tests)
            ref = run_cpu.cache["o"].ctensor(sample, "total")
            perr = float((np.abs(got - ref) / (1e-10 * np.abs(ref) + 1e-12)).max())
            cpu = {"value": pps, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_points} random points of the step's grid, {dt:.1f} s; C++/OpenMP restatement of the reference "
                             "algorithm (dense 7 GEMV + 28 DOT per point), all host threads",
                   "parity_max_scaled_err": perr}
        if world == 1 and not octant and not args.no_extras:
            try:
                extras.update(stage_records(g, grid, dev, t_dev, idx_dev, sh, B))
            except Exception as e:      # extras must never cost the headline line
                extras["stage_records_error"] = repr(e)
            try:
                del t_dev
                torch.cuda.empty_cache()
                extras["config4"] = config4_record(dev)
            except Exception as e:
                extras["config4_error"] = repr(e)
        line = {"metric": "J^B tensor grid points/sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e, "unit": "points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": ms_step_h,
                        "note": "same C-ABI calls with pinned host outputs; rows of a finished panel batch are copied out while the next is contracted; "
                                "h2d = the grid description (the points are generated on the device from it, grid.f90:498-511)"},
                "gpu_launches": int(tot["launches"]), "roofline": roof, "stage_ms_per_step": stage, "balance": balance, "cpu_baseline": cpu,
                "stages": extras, "wall_s_timed_region": wall}
        print(json.dumps(line))
    g.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
