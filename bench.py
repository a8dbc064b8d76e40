#!/usr/bin/env python3
"""bench.py -- J^B tensor grid points per second at nbf ~ 10^4 (BASELINE.json metric).

Workload (configs[4], "nanoring-size synthetic"): 278 carbon-like centres x 36 cartesian functions
(def2-TZVP carbon shells) = 10 008 basis functions, compact hexagonal flake, seeded random symmetric D
and antisymmetric P_x,P_y,P_z; cdens tensors on a 256^3 even grid over the bounding box + 8 bohr.
One STEP = one pass of the whole hot path (spatial sort -> tile screening -> basis panels -> DMMA
contraction + fused tensor epilogue) over one brick of that grid: octant (rank mod 8) = a contiguous
128^3 sub-grid at full resolution (2 097 152 points), so that 8 ranks cover the full grid once per step.
The flake is (approximately) mirror-symmetric in x, y and z, so the octants carry near-equal work; the
reported time is the max over ranks (weak scaling; there is no data-path collective in cdens mode).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU (oracle port)

Timing: CUDA events on the library's stream around every call (stats.ms_total), max over ranks.
`value` has inputs/outputs resident in HBM; `e2e` goes through the same C-ABI call with pinned HOST buffers
(H2D of the points and D2H of the tensors inside the timed region).
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


def ncu_traffic(nbf):
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return d.get(str(nbf), {}).get("k_jtensor_dram_bytes_per_launch")
    except Exception:
        return None


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--natoms", type=int, default=278)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--cpu-points", type=int, default=192, help="bounded CPU sample per step / for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--geometry", default="flake", choices=["flake", "ring"],
                    help="flake: compact hexagonal flake (headline); ring: radius-120-bohr ring, mostly empty box (SURVEY 8d ii)")
    ap.add_argument("--general-p", action="store_true", help="general (not antisymmetric) perturbed densities P_b")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = max(args.steps, 1)

    geo = "hex flake" if args.geometry == "flake" else "ring (radius 120 bohr)"
    cfg = {"workload": f"synthetic {geo} {args.natoms} C-like centres x 36 fn (nbf={args.natoms * 36}), cdens J^B tensors, "
                       f"{args.grid}^3 even grid over bbox+8 bohr, step = octant (rank mod 8) = {args.grid ** 3 // NSLAB} points/GPU",
           "nbf": args.natoms * 36, "grid": [args.grid] * 3, "points_per_step_per_gpu": args.grid ** 3 // NSLAB,
           "spincase": "total (closed shell)", "giao": True, "screening_thrs": 1e-8,
           "densities": "seeded random symmetric D, " + ("general" if args.general_p else "antisymmetric") + " P_x,P_y,P_z",
           "cache": "inputs larger than L2 (contraction operand 4*nbf^2*8 B = %.1f GB, panels streamed)" % (4 * (args.natoms * 36) ** 2 * 8 / 1e9),
           "parallelism": f"grid slabs over {world} GPU(s), no data-path collective"}

    # ------------------------------------------------------------------ reference arm (CPU) -------
    if args.impl == "reference":
        if rank != 0:
            return 0
        sh, dens, nbf, origin, basv, pts = build_workload(args.natoms, args.grid, args.geometry, args.general_p)
        r = slab_points(origin, basv, pts, 0)
        rng = np.random.default_rng(77)
        sample = np.ascontiguousarray(r[rng.choice(r.shape[0], size=args.cpu_points, replace=False)])
        for _ in range(args.warmup):
            run_cpu(sh, dens, sample[: max(16, args.cpu_points // 8)], 0)
        times = []
        for _ in range(K):
            pps, cores, dt = run_cpu(sh, dens, sample, 0)
            times.append(dt)
        dt = float(np.mean(times))
        val = args.cpu_points / dt
        line = {"impl": "reference", "metric": "J^B tensor grid points/sec", "value": val, "unit": "points/s", "n_gpus": args.gpus,
                "steps": K, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": val, "unit": "points/s", "cores": cores, "kind": "port",
                                 "sample": f"{args.cpu_points} random points of the step's slab per step; C++/OpenMP restatement of the "
                                           "reference algorithm (dense 7 GEMV + 28 DOT per point), all host threads"},
                "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ CUDA arm --------------------
    import torch
    import gimic_b200
    from gimic_b200 import synthetic
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
    r_np = slab_points(origin, basv, pts, rank % NSLAB)
    n = r_np.shape[0]
    r_host = torch.empty((n, 3), dtype=torch.float64, pin_memory=True); r_host.numpy()[:] = r_np
    t_host = torch.empty((n, 9), dtype=torch.float64, pin_memory=True)
    r_dev = r_host.to(dev); t_dev = torch.empty((n, 9), dtype=torch.float64, device=dev)
    g.set_profiling(True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(k, host):
        tot = {"ms_total": 0.0, "ms_contract": 0.0, "ms_basis": 0.0, "ms_sort": 0.0, "ms_tiles": 0.0, "launches": 0,
               "contract_launches": 0, "executed_flops": 0.0, "useful_flops": 0.0, "dense_flops": 0.0, "sum_nact": 0.0, "n_tiles": 0}
        for _ in range(k):
            if host:
                g.jtensors(r_host.numpy(), "total", out=t_host.numpy())
            else:
                g.jtensors(r_dev, "total", out=t_dev)
            s = g.stats()
            for key in tot:
                tot[key] += s[key]
        return tot

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

    # points with at least one unscreened basis function (all others are exact zeros and cost nothing; SURVEY 8d caveat)
    n_active = torch.tensor([float((t_dev.abs().amax(dim=1) > 0).sum())], dtype=torch.float64, device=dev)
    ms = torch.tensor([tot["ms_total"] / K, tot_h["ms_total"] / K], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_active, op=dist.ReduceOp.SUM)
    ms_step, ms_step_h = float(ms[0]), float(ms[1])
    value = world * n / (ms_step * 1e-3)
    e2e = world * n / (ms_step_h * 1e-3)

    if rank == 0:
        peak, peak_src = fp64_peak()
        t_contract = tot["ms_contract"] * 1e-3
        achieved = tot["executed_flops"] / t_contract / 1e12 if t_contract > 0 else None
        roof = {"bound": "tensor", "kernel": "k_jtensor<GIAO> (FP64 DMMA contraction of Phi with [D|Px|Py|Pz], GIAO terms by atom-boundary taps of the D accumulator, fused tensor epilogue)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "traffic": ncu_traffic(nbf), "peak_source": peak_src,
                "flops": "EXECUTED FP64 flops per tile: DMMA 2*128*4*nact*nn + GIAO-tap DFMA 2*128*3*nn*natoms_active (screened-function skipping on; exact zeros in the reference)",
                # the same flops without K/N padding and partial tiles (real points x active functions): what a perfect tiling would issue
                "useful_tflops": (tot["useful_flops"] / t_contract / 1e12) if t_contract > 0 else None,
                "useful_frac": (tot["useful_flops"] / t_contract / 1e12 / peak) if t_contract > 0 else None,
                "executed_over_useful": (tot["executed_flops"] / tot["useful_flops"]) if tot["useful_flops"] else None,
                "avg_launch_ms": tot["ms_contract"] / max(tot["contract_launches"], 1), "launches_timed": tot["contract_launches"],
                "share_of_step": tot["ms_contract"] / tot["ms_total"] if tot["ms_total"] else None,
                "dense_equivalent_tflops": tot["dense_flops"] / (tot["ms_total"] * 1e-3) / 1e12,
                "skip_ratio_dense_over_executed": tot["dense_flops"] / tot["executed_flops"] if tot["executed_flops"] else None,
                "mean_active_functions_per_tile": tot["sum_nact"] / max(tot["n_tiles"], 1),
                "frac_points_with_active_functions": float(n_active) / (world * n),
                "active_points_per_s": float(n_active) / (ms_step * 1e-3),
                "stage_ms_per_step": {k2: tot[k2] / K for k2 in ("ms_sort", "ms_tiles", "ms_basis", "ms_contract", "ms_total")}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rng = np.random.default_rng(77)
            sample = np.ascontiguousarray(r_np[rng.choice(n, size=args.cpu_points, replace=False)])
            run_cpu(sh, dens, sample[:16], 0)
            pps, cores, dt = run_cpu(sh, dens, sample, 0)
            # parity of the timed GPU result on the CPU-evaluated sample (same tolerance as the tests)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            ref = run_cpu.cache["o"].ctensor(sample, "total")
            got = g.jtensors(sample, "total")
            perr = float((np.abs(got - ref) / (1e-10 * np.abs(ref) + 1e-12)).max())
            cpu = {"value": pps, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_points} random points of the step's slab, {dt:.1f} s; C++/OpenMP restatement of the reference "
                             "algorithm (dense 7 GEMV + 28 DOT per point), all host threads",
                   "parity_max_scaled_err": perr}
        line = {"metric": "J^B tensor grid points/sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e, "unit": "points/s", "h2d_bytes_per_step": int(n * 24), "d2h_bytes_per_step": int(n * 72),
                        "ms_per_step": ms_step_h},
                "gpu_launches": int(tot["launches"]), "roofline": roof, "cpu_baseline": cpu,
                "wall_s_timed_region": wall}
        print(json.dumps(line))
    g.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
