"""One call over the complete 256^3 grid at nbf=10008 on one GPU (config 5 at N=1): timing + consistency with slab calls."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, gimic_b200
from gimic_b200 import synthetic
natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 278
n1 = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sh, dens, nbf, origin, basv, pts = bench.build_workload(natoms, n1)
g = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=synthetic.dens_to_colmajor(dens), **sh)
grid = gimic_b200.Grid(origin, basv, pts)
dev = torch.device("cuda", 0)
out = torch.empty((grid.n, 9), dtype=torch.float64, device=dev)
g.set_profiling(True)
res = {}
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g.jtensors_grid(grid, 0, grid.n, "total", out=out)
    dt = time.perf_counter() - t0
    st = g.stats()
    res = dict(points=grid.n, nbf=nbf, wall_s=dt, points_per_s=grid.n / dt, ms_total=st["ms_total"], ms_contract=st["ms_contract"],
               ms_basis=st["ms_basis"], ms_sort=st["ms_sort"], ms_tiles=st["ms_tiles"], n_tiles=st["n_tiles"],
               contract_tflops=st["executed_flops"] / (st["ms_contract"] * 1e-3) / 1e12, mean_nact=st["sum_nact"] / st["n_tiles"],
               contract_launches=st["contract_launches"])
# consistency: three random slabs recomputed on their own
rng = np.random.default_rng(0)
worst = 0.0
for _ in range(3):
    lo = int(rng.integers(0, grid.n - 50000)); hi = lo + 50000
    part = g.jtensors_grid(grid, lo, hi, "total")
    ref = out[lo:hi].cpu().numpy()
    worst = max(worst, float((np.abs(part - ref) / (1e-10 * np.abs(ref) + 1e-12)).max()))
res["slab_vs_full_max_scaled_err"] = worst
res["frac_points_all_zero"] = float((out.abs().sum(1) == 0).double().mean())
print(json.dumps(res))
